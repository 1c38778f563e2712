mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
tail -c 600 gpurun_out/bench_a.err
CNB_DEC_TRACE=1 python tools/decode_trace.py --batch 64 > gpurun_out/trace16.log 2>&1
CNB_DEC_TRACE=1 CNB_DEC_NR=32 python tools/decode_trace.py --batch 64 > gpurun_out/trace32.log 2>&1
python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/t_all.log 2>&1
tail -5 gpurun_out/t_all.log
