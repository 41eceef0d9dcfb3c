#!/bin/bash
# round 2, GPU session 13: un-instrumented build — pair kernel with weight multicast (cluster 1 / 2) vs the two-CTAs-per-SM kernel; sweep with cluster 2
set +e
O=gpurun_out
mkdir -p $O
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 10 > $O/s13_bench_tc2.json 2> $O/s13_bench_tc2.err
for cl in 1 2; do
  BGX_PAIR_CLUSTER=$cl BGX_SPLINE_KERNEL=pair timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 10 > $O/s13_bench_pair_cl$cl.json 2> $O/s13_bench_pair_cl$cl.err
  BGX_PAIR_CLUSTER=$cl timeout 300 python bench.py --workload spline_d384_8blk --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 5 > $O/s13_bench_d384_cl$cl.json 2> $O/s13_bench_d384_cl$cl.err
  BGX_PAIR_CLUSTER=$cl timeout 300 python bench.py --workload spline_d3072_8blk --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 3 > $O/s13_bench_d3072_cl$cl.json 2> $O/s13_bench_d3072_cl$cl.err
done
BGX_PAIR_CLUSTER=2 timeout 900 python -m pytest tests -m gpu -q > $O/s13_all_tests_cl2.log 2>&1
echo "rc=$?" >> $O/s13_all_tests_cl2.log
echo done
