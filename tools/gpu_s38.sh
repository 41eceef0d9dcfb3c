#!/bin/bash
# round 2, GPU session 38: cp.async prefetch of the next operand block in the K-split linear kernel
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pair.py tests/test_gpu_autograd.py -m gpu -q > $O/s38_tests.log 2>&1
echo "rc=$?" >> $O/s38_tests.log
timeout 300 python tools/bench_train_kernels.py > $O/s38_train_kernels.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-sweep --steps 10 > $O/s38_bench.json 2> $O/s38_bench.err
echo done
