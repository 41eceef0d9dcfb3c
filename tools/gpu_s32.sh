#!/bin/bash
# round 2, GPU session 32: source-level instruction counts of the pair kernel (one launch), checks of the phase-bit edits
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pair.py tests/test_gpu_autograd.py -m gpu -q > $O/s32_tests.log 2>&1
echo "rc=$?" >> $O/s32_tests.log
timeout 400 ncu --clock-control none --set full --import-source on -k regex:spline_coupling_pair -s 10 -c 1 -o $O/s32_spline_pair \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --no-train > $O/s32_ncu.log 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-train --steps 5 > $O/s32_bench.json 2> $O/s32_bench.err
echo done
