#!/bin/bash
# round 2, GPU session 3: affine pair kernel, config-5 sweep, full default bench line, f32x2 probe, racecheck
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pair.py tests/test_gpu_autograd.py -x -q > $O/s3_tests.log 2>&1
echo "tests rc=$?" >> $O/s3_tests.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/f32x2_probe tools/f32x2_probe.cu && /tmp/f32x2_probe > $O/s3_f32x2_probe.txt 2>&1
timeout 900 python bench.py --steps 10 > $O/s3_bench.json 2> $O/s3_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/s3_bench_ref.json 2> $O/s3_bench_ref.err
timeout 900 python -m pytest tests -m gpu -x -q > $O/s3_all_tests.log 2>&1
echo "all tests rc=$?" >> $O/s3_all_tests.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_pair.py -x -q -k "narrow or golden" > $O/s3_racecheck.log 2>&1
echo done
