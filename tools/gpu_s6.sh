#!/bin/bash
# round 2, GPU session 6: pair kernel with 24 epilogue warps (EPW=6) vs 16 (EPW=4)
set +e
O=gpurun_out
mkdir -p $O
BGX_PAIR_EPW=6 timeout 900 python -m pytest tests/test_gpu_pair.py -x -q > $O/s6_tests_epw6.log 2>&1
echo "rc=$?" >> $O/s6_tests_epw6.log
for epw in 4 6; do
  BGX_PAIR_EPW=$epw BGX_SPLINE_KERNEL=pair timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 10 > $O/s6_bench_pair_epw$epw.json 2> $O/s6_bench_pair_epw$epw.err
  BGX_PAIR_EPW=$epw BGX_SPLINE_KERNEL=pair_wide timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 10 > $O/s6_bench_pairwide_epw$epw.json 2> $O/s6_bench_pairwide_epw$epw.err
done
BGX_PAIR_EPW=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:spline_coupling_pair -s 8 -c 1 -o $O/s6_pair_epw6 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-sweep --no-train > $O/s6_ncu.log 2>&1
echo done
