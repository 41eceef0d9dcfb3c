#!/bin/bash
# Round-2 evidence capture (run under gpurun, 1 GPU).  Writes everything to gpurun_out/r2_*.
set -u
O=gpurun_out
mkdir -p $O
B=${1:-1048576}
NCU="ncu --clock-control none"
# the numbers themselves first (never taken under the profiler)
timeout 1200 python bench.py --steps 10 --warmup 3 > $O/r2_bench.json 2> $O/r2_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err
timeout 300 python bench.py --workload ala2_affine_d66_8blk --steps 10 --warmup 3 --no-cpu-baseline --no-sweep --no-train > $O/r2_bench_affine.json 2> $O/r2_bench_affine.err
BGX_SPLINE_KERNEL=tc2 timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e --no-sweep --no-train > $O/r2_bench_spline_tc2.json 2>/dev/null
timeout 300 python tools/bench_ic.py > $O/r2_bench_ic.json 2> $O/r2_bench_ic.err
BGX_IC_BULK=0 timeout 300 python tools/bench_ic.py > $O/r2_bench_ic_staged.json 2>/dev/null
timeout 120 python tools/mma_rate.py > $O/r2_mma_rate_probe.txt 2>&1
# training path: kernel times, step profiles per backward mode, end-to-end chunking
timeout 300 python tools/bench_train_kernels.py > $O/r2_train_kernels.txt 2>&1
timeout 300 python tools/profile_train.py > $O/r2_train_profile.txt 2>&1
BGX_BACKWARD_GEMM=fp32 timeout 300 python tools/profile_train.py > $O/r2_train_profile_fp32.txt 2>&1
BGX_BACKWARD_GEMM=tf32 timeout 300 python tools/profile_train.py > $O/r2_train_profile_tf32.txt 2>&1
timeout 600 python tools/e2e_chunks.py > $O/r2_e2e_chunks.txt 2>&1
# full captures of the dominant kernels at the bench batch size (one launch each)
timeout 400 $NCU --set full --import-source on -k regex:spline_coupling_pair -s 10 -c 1 -o $O/r2_spline_pair \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --no-train --batch-per-gpu $B > $O/r2_spline_pair.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:spline_coupling_pair -s 10 -c 1 -o $O/r2_spline_pair_d384 \
  python bench.py --workload spline_d384_8blk --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --no-train --batch-per-gpu 262144 > $O/r2_spline_pair_d384.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:affine_coupling_pair -s 10 -c 1 -o $O/r2_affine_pair_d384 \
  python bench.py --workload affine_d384_8blk --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --no-train --batch-per-gpu 262144 > $O/r2_affine_pair_d384.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:ic_ -s 4 -c 2 -o $O/r2_ic python tools/run_ic.py > $O/r2_ic.log 2>&1
# the training kernels at the KL step's shapes (first launches of the bench tool: K=128 -> N=828, then K=828 -> N=128; N=825 x K=128)
timeout 300 $NCU --set full --import-source on -k regex:linear_tc_kernel -s 3 -c 1 -o $O/r2_linear_tc_n828 env BGX_NO_TORCH_PROFILER=1 python tools/bench_train_kernels.py > $O/r2_linear_tc.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:linear_tc_kernel -s 22 -c 1 -o $O/r2_linear_tc_k828 env BGX_NO_TORCH_PROFILER=1 python tools/bench_train_kernels.py >> $O/r2_linear_tc.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:gemm_tn_kernel -s 3 -c 1 -o $O/r2_gemm_tn env BGX_NO_TORCH_PROFILER=1 python tools/bench_train_kernels.py > $O/r2_gemm_tn.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:spline_backward_kernel -s 3 -c 1 -o $O/r2_spline_bwd env BGX_NO_TORCH_PROFILER=1 python tools/bench_train_kernels.py > $O/r2_spline_bwd.log 2>&1
# launch list of the bench command (shares of the step)
timeout 300 $NCU --metrics gpu__time_duration.sum -s 60 -c 60 --csv --log-file $O/r2_launches_spline.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --no-train > /dev/null 2>&1
# sanitizers on the new kernels (small batches)
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_pair.py tests/test_gpu_ic.py tests/test_gpu_relic.py -x -q > $O/r2_memcheck.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_pair.py -x -q -k "narrow or affine_wide" > $O/r2_racecheck.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_autograd.py -x -q -k "drivers or gemm_modes" > $O/r2_memcheck_train.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_pair.py -x -q -k "weight_gradient or linear_layer" > $O/r2_racecheck_train.log 2>&1
timeout 200 python -c 'import __graft_entry__ as g; g.smoke()' > $O/r2_smoke.log 2>&1; tail -2 $O/r2_smoke.log
timeout 1200 python -m pytest tests -m gpu -q > $O/r2_gpu_tests.log 2>&1; tail -2 $O/r2_gpu_tests.log
ls -la $O | grep r2_ | tail -30
