#!/bin/bash
# round 2, GPU session 4 (2 GPUs): KL training step with overlapped per-block all-reduce, e2e copy ceiling at N=2
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-sweep > $O/s4_bench_n2.json 2> $O/s4_bench_n2.err
timeout 600 python bench.py --steps 10 --no-sweep --no-cpu-baseline > $O/s4_bench_n1.json 2> $O/s4_bench_n1.err
echo done
