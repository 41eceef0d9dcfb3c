#!/bin/bash
# round 2, GPU session 17: vector operand staging (LDG.128) in the wide pair kernels and bgx_linear; padded last layer
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pair.py tests/test_gpu_autograd.py -m gpu -q > $O/s17_tests.log 2>&1
echo "rc=$?" >> $O/s17_tests.log
BGX_BACKWARD_GEMM=tcgen05 timeout 300 python tools/profile_train.py > $O/s17_train_profile_tcgen05.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-train --steps 10 > $O/s17_bench.json 2> $O/s17_bench.err
echo done
