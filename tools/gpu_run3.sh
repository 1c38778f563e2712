mkdir -p gpurun_out
CNB_DEC_TRACE=1 python tools/decode_trace.py --batch 64 > gpurun_out/trace16.log 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_bench_parity.py -m gpu -q --timeout 600 -x -k "cluster or bench_config" > gpurun_out/t_dec.log 2>&1
tail -3 gpurun_out/t_dec.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
tail -c 300 gpurun_out/bench_b.err
