#!/bin/bash
# round 2: final numbers of the final build (un-profiled), full GPU suite, smoke
set +e
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/final_gpu_tests.log 2>&1
echo "rc=$?" >> $O/final_gpu_tests.log
timeout 1200 python bench.py --steps 10 --warmup 3 > $O/final_bench.json 2> $O/final_bench.err
timeout 300 python tools/bench_train_kernels.py > $O/final_train_kernels.txt 2>&1
timeout 300 python tools/profile_train.py > $O/final_train_profile.txt 2>&1
timeout 200 python -c 'import __graft_entry__ as g; g.smoke()' > $O/final_smoke.log 2>&1
tail -2 $O/final_gpu_tests.log; tail -1 $O/final_smoke.log
