mkdir -p gpurun_out
N=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_c1_${N}gpu.json 2> gpurun_out/bench_c1_${N}gpu.err
tail -c 200 gpurun_out/bench_c1_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config 4 --steps 3 --warmup 3 > gpurun_out/bench_c4_${N}gpu.json 2> gpurun_out/bench_c4_${N}gpu.err
tail -c 200 gpurun_out/bench_c4_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --config 3 --steps 3 --warmup 3 > gpurun_out/bench_c3_${N}gpu.json 2> gpurun_out/bench_c3_${N}gpu.err
tail -c 200 gpurun_out/bench_c3_${N}gpu.err
N=4
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --config 3 --steps 3 --warmup 3 > gpurun_out/bench_c3_${N}gpu.json 2> gpurun_out/bench_c3_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_c1_${N}gpu.json 2> gpurun_out/bench_c1_${N}gpu.err
