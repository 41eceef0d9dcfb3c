#!/bin/bash
# round 2, GPU session 10: per-role wait accounting of the pair kernel; training-step profile
set +e
O=gpurun_out
mkdir -p $O
for epw in 4 6; do for dbg in 0 7 6 1; do
  echo "== EPW $epw DEBUG $dbg" >> $O/s10_trace.txt
  BGX_PAIR_DEBUG=$dbg BGX_PAIR_EPW=$epw timeout 120 python tools/trace_pair.py >> $O/s10_trace.txt 2>&1
done; done
timeout 300 python tools/profile_train.py > $O/s10_train_profile.txt 2>&1
echo done
