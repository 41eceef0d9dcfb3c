#!/bin/bash
# round 2, GPU session 52: default bench line under torchrun at N GPUs (no sweep / cpu baseline)
set +e
O=gpurun_out
mkdir -p $O
N=${1:-2}
python -c "import torch; print('devices', torch.cuda.device_count())"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 10 --warmup 3 --no-sweep --no-cpu-baseline > $O/s52_bench_${N}gpu.json 2> $O/s52_bench_${N}gpu.err
echo "rc=$?"
nvidia-smi -L | head -8
tail -3 $O/s52_bench_${N}gpu.err
echo done
