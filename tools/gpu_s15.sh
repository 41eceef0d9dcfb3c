#!/bin/bash
# round 2, GPU session 15: bgx_linear (tcgen05 linear layer) + tcgen05 backward mode
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pair.py -x -q -k "linear" > $O/s15_linear.log 2>&1
echo "rc=$?" >> $O/s15_linear.log
timeout 900 python -m pytest tests/test_gpu_autograd.py tests/test_gpu_fullsize.py tests/test_gpu_pair.py -q > $O/s15_tests.log 2>&1
echo "rc=$?" >> $O/s15_tests.log
timeout 600 python bench.py --steps 10 --no-sweep --no-cpu-baseline --no-e2e > $O/s15_bench.json 2> $O/s15_bench.err
echo done
