#!/bin/bash
# round 2, GPU session 18: bgx_gemm_tn (weight gradients on tcgen05)
set +e
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pair.py -m gpu -q -k "weight_gradient" > $O/s18_tn_tests.log 2>&1
echo "rc=$?" >> $O/s18_tn_tests.log
timeout 600 python -m pytest tests/test_gpu_autograd.py -m gpu -q > $O/s18_autograd_tests.log 2>&1
echo "rc=$?" >> $O/s18_autograd_tests.log
timeout 300 python tools/bench_train_kernels.py > $O/s18_train_kernels.txt 2>&1
BGX_BACKWARD_GEMM=tcgen05 timeout 300 python tools/profile_train.py > $O/s18_train_profile_tcgen05.txt 2>&1
echo done
