#!/bin/bash
# One-GPU pass of a round (run through gpurun): GPU tests, the bench lines of BASELINE configs 1-4 and the decoder phase trace.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [quick]'      quick = decoder tests + config-1 bench + trace only
mkdir -p gpurun_out
if [ "$1" = "quick" ]; then
  python -m pytest tests/test_gpu_parity.py tests/test_bench_parity.py -m gpu -q --timeout 600 -x -k "cluster or bench_config" > gpurun_out/t_dec.log 2>&1
  tail -3 gpurun_out/t_dec.log
else
  python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/t_all.log 2>&1
  tail -4 gpurun_out/t_all.log
fi
CNB_DEC_TRACE=1 python tools/decode_trace.py --batch 64 > gpurun_out/trace16.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
tail -c 300 gpurun_out/bench_c1.err
[ "$1" = "quick" ] && exit 0
python bench.py --config 2 --steps 3 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python bench.py --config 4 --steps 2 --warmup 3 --cpu-sample 2 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python bench.py --config 3 --steps 2 --warmup 3 --cpu-sample 2 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --vocab-words 8174 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c1_v8192.json 2> gpurun_out/bench_c1_v8192.err
tail -c 300 gpurun_out/bench_c3.err
CNB_PARITY_CLIPS=64 python -m pytest tests/test_bench_parity.py -m gpu -q --timeout 900 -s -k "bench_config" > gpurun_out/t_parity64.log 2>&1
grep "parity\]" gpurun_out/t_parity64.log | cut -c1-400
