#!/bin/bash
# round 2, GPU session 26: CUDA-graph replay of the host pipeline
set +e
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q > $O/s26_tests.log 2>&1
echo "rc=$?" >> $O/s26_tests.log
timeout 600 python tools/e2e_chunks.py > $O/s26_e2e_chunks.txt 2>&1
echo done
