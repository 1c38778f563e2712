#!/bin/bash
# One-GPU profiling pass whose outputs are summarised under profiles/ (run through gpurun):
#   1. ncu launch list of a short bench run   2. ncu --set full captures of the kernels that changed this round
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dwconv_ln_w7 -c 1 -o gpurun_out/full_dw_s4 -f \
  python tools/run_once.py caption --batch 64 --reps 1 > gpurun_out/ncu_full_dw4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decoder_cluster -c 1 -o gpurun_out/full_dec -f \
  python tools/run_once.py caption --batch 64 --reps 1 > gpurun_out/ncu_full_dec.log 2>&1
ls -la gpurun_out | tail -8
