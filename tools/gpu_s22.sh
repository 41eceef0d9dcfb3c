#!/bin/bash
# round 2, GPU session 22: native recompute / backward drivers (bgx_mlp_forward_train, bgx_mlp_backward)
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_autograd.py tests/test_gpu_pair.py -m gpu -q -x > $O/s22_tests.log 2>&1
echo "rc=$?" >> $O/s22_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-sweep --steps 10 > $O/s22_bench.json 2> $O/s22_bench.err
timeout 300 python tools/profile_train.py > $O/s22_train_profile.txt 2>&1
echo done
