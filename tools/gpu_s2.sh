#!/bin/bash
# round 2, GPU session 2: pair kernel v2 (compact code, sleeping waits, reducer warp)
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pair.py -x -q > $O/s2_pair_tests.log 2>&1
echo "pair tests rc=$?" >> $O/s2_pair_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 10 > $O/s2_bench_pair.json 2> $O/s2_bench_pair.err
BGX_SPLINE_KERNEL=pair_wide timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 10 > $O/s2_bench_pairwide.json 2> $O/s2_bench_pairwide.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spline_coupling_pair -s 8 -c 1 -o $O/s2_pair python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/s2_ncu.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/s2_all_tests.log 2>&1
echo "all tests rc=$?" >> $O/s2_all_tests.log
echo done
