#!/bin/bash
# round 2, GPU session 35: bgx_spline_coupling_backward (one host call per block backward)
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_autograd.py -m gpu -q > $O/s35_tests.log 2>&1
echo "rc=$?" >> $O/s35_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-sweep --steps 10 > $O/s35_bench.json 2> $O/s35_bench.err
timeout 300 python tools/profile_train.py > $O/s35_train_profile.txt 2>&1
echo done
