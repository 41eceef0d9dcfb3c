#!/bin/bash
# round 2, GPU session 34: role offset between the slots, A/B in one session
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pair.py tests/test_gpu_fullsize.py -m gpu -q > $O/s34_tests.log 2>&1
echo "rc=$?" >> $O/s34_tests.log
for i in 1 2 3; do
for rs in 1 2 3; do
BGX_PAIR_ROLE_STRIDE=$rs timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 10 > $O/s34_bench_rs${rs}_$i.json 2> $O/s34_bench.err
done
done
echo done
