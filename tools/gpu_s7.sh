#!/bin/bash
# round 2, GPU session 7: IC kernels with bulk-TMA tile I/O (parity + A/B), pair-kernel profiles EPW 4 / 6
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ic.py tests/test_gpu_relic.py tests/test_gpu_cdf.py tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -x -q > $O/s7_ic_tests.log 2>&1
echo "rc=$?" >> $O/s7_ic_tests.log
BGX_IC_BULK=0 timeout 300 python tools/bench_ic.py > $O/s7_ic_staged.json 2> $O/s7_ic_staged.err
BGX_IC_BULK=1 timeout 300 python tools/bench_ic.py > $O/s7_ic_bulk.json 2> $O/s7_ic_bulk.err
for epw in 6 4; do
BGX_PAIR_EPW=$epw BGX_SPLINE_KERNEL=pair timeout 600 ncu --set full --clock-control none --import-source on -k regex:spline_coupling_pair -s 8 -c 1 -o $O/s7_pair_epw$epw python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-sweep --no-train > $O/s7_ncu_epw$epw.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ic_ -c 4 -o $O/s7_ic python tools/run_ic.py > $O/s7_ncu_ic.log 2>&1
echo done
