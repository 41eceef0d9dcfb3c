#!/bin/bash
# round 2, GPU session 1: pair kernel correctness, A/B bench, MMA probe, ncu capture, sanitizer
set +e
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/s1_smi.txt
timeout 900 python -m pytest tests/test_gpu_pair.py -x -q > $O/s1_pair_tests.log 2>&1
echo "pair tests rc=$?" >> $O/s1_pair_tests.log
timeout 120 python tools/mma_rate.py > $O/s1_mma_rate.txt 2>&1
BGX_SPLINE_KERNEL=tc2 timeout 300 python bench.py --no-cpu-baseline --steps 10 > $O/s1_bench_tc2.json 2> $O/s1_bench_tc2.err
timeout 300 python bench.py --no-cpu-baseline --steps 10 > $O/s1_bench_pair.json 2> $O/s1_bench_pair.err
BGX_SPLINE_KERNEL=pair_wide timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 10 > $O/s1_bench_pairwide.json 2> $O/s1_bench_pairwide.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spline_coupling_pair -s 8 -c 1 -o $O/s1_pair python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/s1_ncu.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/s1_all_tests.log 2>&1
echo "all tests rc=$?" >> $O/s1_all_tests.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_pair.py -x -q -k "narrow or wide" > $O/s1_memcheck.log 2>&1
echo done
