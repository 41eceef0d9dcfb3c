#!/bin/bash
# round 2, GPU session 14: tensor-core backward GEMMs — gradient parity + KL step timing; full suite
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_autograd.py -x -q > $O/s14_autograd.log 2>&1
echo "rc=$?" >> $O/s14_autograd.log
timeout 300 python tools/profile_train.py > $O/s14_train_profile.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > $O/s14_all_tests.log 2>&1
echo "rc=$?" >> $O/s14_all_tests.log
timeout 900 python bench.py --steps 10 > $O/s14_bench.json 2> $O/s14_bench.err
echo done
