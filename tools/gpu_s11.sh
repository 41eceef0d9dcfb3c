#!/bin/bash
# round 2, GPU session 11: per-role wait accounting; weight stages shared by TMA multicast across a cluster (1 / 2 / 4 CTAs)
set +e
O=gpurun_out
mkdir -p $O
rm -f $O/s11_trace.txt
for cl in 1 2 4; do for dbg in 0 7 6 1; do
  echo "== CLUSTER $cl DEBUG $dbg" >> $O/s11_trace.txt
  BGX_PAIR_CLUSTER=$cl BGX_PAIR_DEBUG=$dbg timeout 120 python tools/trace_pair.py >> $O/s11_trace.txt 2>&1
done; done
for cl in 2 4; do
  BGX_PAIR_CLUSTER=$cl BGX_SPLINE_KERNEL=pair timeout 600 python -m pytest tests/test_gpu_pair.py -x -q > $O/s11_tests_cl$cl.log 2>&1
  echo "rc=$?" >> $O/s11_tests_cl$cl.log
done
for cl in 1 2 4; do
  BGX_PAIR_CLUSTER=$cl BGX_SPLINE_KERNEL=pair timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 10 > $O/s11_bench_pair_cl$cl.json 2> $O/s11_bench_pair_cl$cl.err
  BGX_PAIR_CLUSTER=$cl BGX_SPLINE_KERNEL=pair_wide timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 10 > $O/s11_bench_pairwide_cl$cl.json 2> $O/s11_bench_pairwide_cl$cl.err
  BGX_PAIR_CLUSTER=$cl timeout 300 python bench.py --workload spline_d384_8blk --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 5 > $O/s11_bench_d384_cl$cl.json 2> $O/s11_bench_d384_cl$cl.err
done
timeout 300 python tools/profile_train.py > $O/s11_train_profile.txt 2>&1
echo done
