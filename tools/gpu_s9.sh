#!/bin/bash
# round 2, GPU session 9: what bounds the pair kernel? MMAs only / epilogue only timing experiments (results are garbage)
set +e
O=gpurun_out
mkdir -p $O
for epw in 4 6; do for dbg in 0 1 2 4 6 7; do
  BGX_PAIR_DEBUG=$dbg BGX_PAIR_EPW=$epw BGX_SPLINE_KERNEL=pair timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 10 > $O/s9_epw${epw}_dbg$dbg.json 2> $O/s9_epw${epw}_dbg$dbg.err
done; done
timeout 1200 python -m pytest tests -m gpu -q > $O/s9_all_tests.log 2>&1
echo "all tests rc=$?" >> $O/s9_all_tests.log
echo done
