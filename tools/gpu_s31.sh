#!/bin/bash
# round 2, GPU session 31: log-det shares accumulated in shared-memory cells; phase bits in registers
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pair.py tests/test_gpu_fullsize.py tests/test_gpu_coupling.py -m gpu -q > $O/s31_tests.log 2>&1
echo "rc=$?" >> $O/s31_tests.log
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 10 > $O/s31_bench_$i.json 2> $O/s31_bench.err
done
echo done
