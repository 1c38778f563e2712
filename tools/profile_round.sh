#!/bin/bash
# One-GPU measurement pass whose outputs are summarised under profiles/ (run through gpurun):
#   1. the bench line (with the CPU-reference leg)   2. ncu launch list of a short bench run
#   3. ncu --set full captures of the top kernels (one launch each): dwconv s1, fused stage-1 MLP, first tcgen05 GEMMs, cluster decoder
set -x
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dwconv_ln_tma -c 1 -o gpurun_out/full_dw_s1 \
  python tools/run_once.py caption --batch 64 --reps 1 > gpurun_out/ncu_full1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 2 -o gpurun_out/full_gemm_s1 \
  python tools/run_once.py caption --batch 64 --reps 1 > gpurun_out/ncu_full2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_fused -c 1 -o gpurun_out/full_mlp_fused \
  python tools/run_once.py caption --batch 64 --reps 1 > gpurun_out/ncu_full4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decoder_cluster -c 1 -o gpurun_out/full_dec \
  python tools/run_once.py caption --batch 64 --reps 1 > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out | tail -20
tail -c 300 gpurun_out/bench_full.err
