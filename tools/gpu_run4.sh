mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:decoder_cluster -s 1 -c 1 -o gpurun_out/dec_prof -f python tools/decode_trace.py --batch 64 > gpurun_out/ncu_dec.log 2>&1
tail -3 gpurun_out/ncu_dec.log
ls -la gpurun_out/*.ncu-rep
