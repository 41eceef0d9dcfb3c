#!/bin/bash
# round 2, GPU session 23: gemm_tn fast-path addressing, wave-aligned e2e chunks, host pipeline test
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_autograd.py tests/test_gpu_pair.py tests/test_gpu_pipeline.py -m gpu -q > $O/s23_tests.log 2>&1
echo "rc=$?" >> $O/s23_tests.log
timeout 300 python tools/bench_train_kernels.py > $O/s23_train_kernels.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-sweep --steps 10 > $O/s23_bench.json 2> $O/s23_bench.err
echo done
