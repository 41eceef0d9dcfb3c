#!/bin/bash
# round 2, GPU session 33: four try_wait attempts per counter check
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pair.py tests/test_gpu_fullsize.py tests/test_gpu_coupling.py -m gpu -q > $O/s33_tests.log 2>&1
echo "rc=$?" >> $O/s33_tests.log
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-sweep --no-train --steps 10 > $O/s33_bench_$i.json 2> $O/s33_bench.err
done
echo done
