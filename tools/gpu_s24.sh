#!/bin/bash
# round 2, GPU session 24: e2e chunk sweep; autograd tests
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_autograd.py -m gpu -q > $O/s24_tests.log 2>&1
echo "rc=$?" >> $O/s24_tests.log
timeout 600 python tools/e2e_chunks.py > $O/s24_e2e_chunks.txt 2>&1
echo done
