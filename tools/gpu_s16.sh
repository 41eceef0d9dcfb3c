#!/bin/bash
# round 2, GPU session 16: canonical log-det cells (slot-rotating roles), bgx_linear tests, tcgen05 backward profile
set +e
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/s16_all_tests.log 2>&1
echo "rc=$?" >> $O/s16_all_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 10 > $O/s16_bench.json 2> $O/s16_bench.err
BGX_BACKWARD_GEMM=tcgen05 timeout 300 python tools/profile_train.py > $O/s16_train_profile_tcgen05.txt 2>&1
echo done
