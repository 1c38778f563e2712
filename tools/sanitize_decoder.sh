mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 800 -x -k "decoder_cluster_vs_oracle and (5-1-12 or 7-5-16 or 6-3-20-94)" > gpurun_out/san_dec.log 2>&1
tail -15 gpurun_out/san_dec.log
