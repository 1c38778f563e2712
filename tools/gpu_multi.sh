#!/bin/bash
# Multi-GPU pass (run through `gpurun --gpus N`): bash tools/gpu_multi.sh N  -> config 1 (weak), 3 and 4 (strong) under torchrun,
# plus the 2-rank NCCL test of distributed.caption_sharded with the real engine.
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_distributed_gpu.py -m gpu -q --timeout 900 -s > gpurun_out/t_dist.log 2>&1
tail -3 gpurun_out/t_dist.log
P=29510
for C in 1 4 3; do
  P=$((P + 1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --config $C \
    --steps 3 --warmup 3 > gpurun_out/bench_c${C}_${N}gpu.json 2> gpurun_out/bench_c${C}_${N}gpu.err
  tail -c 200 gpurun_out/bench_c${C}_${N}gpu.err
done
