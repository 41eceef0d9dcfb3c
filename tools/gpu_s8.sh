#!/bin/bash
# round 2, GPU session 8: lean barrier waits everywhere — full test-suite, A/B of the spline kernels, e2e chunking
set +e
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/s8_all_tests.log 2>&1
echo "all tests rc=$?" >> $O/s8_all_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 10 > $O/s8_bench_tc2.json 2> $O/s8_bench_tc2.err
for epw in 4 6; do
  BGX_PAIR_EPW=$epw BGX_SPLINE_KERNEL=pair timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 10 > $O/s8_bench_pair_epw$epw.json 2> $O/s8_bench_pair_epw$epw.err
  BGX_PAIR_EPW=$epw BGX_SPLINE_KERNEL=pair_wide timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 10 > $O/s8_bench_pairwide_epw$epw.json 2> $O/s8_bench_pairwide_epw$epw.err
done
BGX_E2E_CHUNK=65536 BGX_E2E_STREAMS=4 timeout 300 python bench.py --no-cpu-baseline --no-sweep --no-train --steps 10 > $O/s8_bench_e2e_c64k.json 2> $O/s8_bench_e2e_c64k.err
BGX_E2E_CHUNK=32768 BGX_E2E_STREAMS=4 timeout 300 python bench.py --no-cpu-baseline --no-sweep --no-train --steps 10 > $O/s8_bench_e2e_c32k.json 2> $O/s8_bench_e2e_c32k.err
echo done
