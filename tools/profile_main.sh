#!/bin/bash
# Trimmed evidence capture (run under gpurun, 1 GPU): the dominant kernels + the bench lines.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 300 $NCU --set full --import-source on -k regex:spline_coupling_tc2 -s 10 -c 1 -o gpurun_out/r1_spline_tc2 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1_spline_tc2.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:affine_coupling_tc -s 10 -c 1 -o gpurun_out/r1_affine_tc \
  python bench.py --workload ala2_affine_d66_8blk --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1_affine_tc.log 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum -s 60 -c 60 --csv --log-file gpurun_out/r1_launches_spline.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum -s 100 -c 60 --csv --log-file gpurun_out/r1_launches_affine.csv \
  python bench.py --workload ala2_affine_d66_8blk --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --extras > gpurun_out/r1_bench_spline.json 2> gpurun_out/r1_bench_spline.err
timeout 300 python bench.py --workload ala2_affine_d66_8blk --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_affine.json 2> gpurun_out/r1_bench_affine.err
BGX_PRECISION=bf16x6 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1_bench_spline_bf16x6.json 2>/dev/null
BGX_TC_SINGLE_CTA=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1_bench_spline_single_cta.json 2>/dev/null
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference.json 2>/dev/null
timeout 200 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r1_smoke.log 2>&1; tail -2 gpurun_out/r1_smoke.log
ls gpurun_out | grep r1_ | wc -l
