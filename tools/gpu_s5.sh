#!/bin/bash
# round 2, GPU session 5: packed-fp32 (f32x2) epilogues — correctness, A/B, sweep, ncu
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/s5_all_tests.log 2>&1
echo "all tests rc=$?" >> $O/s5_all_tests.log
BGX_T2_PACKED=0 timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 10 > $O/s5_bench_scalar.json 2> $O/s5_bench_scalar.err
timeout 900 python bench.py --steps 10 > $O/s5_bench.json 2> $O/s5_bench.err
BGX_SPLINE_KERNEL=pair timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 10 > $O/s5_bench_pair.json 2> $O/s5_bench_pair.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spline_coupling_tc2 -s 8 -c 1 -o $O/s5_tc2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-sweep --no-train > $O/s5_ncu.log 2>&1
echo done
