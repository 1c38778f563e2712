mkdir -p gpurun_out
CNB_DEC_TRACE=1 CNB_DEC_NR=32 CNB_DEC_FILL=1 python tools/decode_trace.py --batch 64 > gpurun_out/trace32full.log 2>&1
CNB_DEC_TRACE=1 CNB_DEC_NR=16 CNB_DEC_FILL=1 python tools/decode_trace.py --batch 64 > gpurun_out/trace16full.log 2>&1
CNB_DEC_NR=32 CNB_DEC_FILL=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nr32.json 2> gpurun_out/bench_nr32.err
tail -c 300 gpurun_out/bench_nr32.err
