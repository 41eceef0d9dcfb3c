#!/bin/bash
# round 2, GPU session 20: line-coalesced gemm_tn converters, vector spline backward I/O
set +e
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pair.py -m gpu -q  > $O/s20_tn_tests.log 2>&1
echo "rc=$?" >> $O/s20_tn_tests.log
timeout 600 python -m pytest tests/test_gpu_autograd.py -m gpu -q > $O/s20_autograd_tests.log 2>&1
echo "rc=$?" >> $O/s20_autograd_tests.log
timeout 300 python tools/bench_train_kernels.py > $O/s20_train_kernels.txt 2>&1
BGX_BACKWARD_GEMM=tcgen05 timeout 300 python tools/profile_train.py > $O/s20_train_profile_tcgen05.txt 2>&1
echo done
