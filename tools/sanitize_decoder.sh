# compute-sanitizer memcheck over the cluster decoder (three shapes) and one small end-to-end caption (every encoder kernel incl. the
# stage-4 row-ring depthwise kernel, the split-precision projection, the tag head):  gpurun --timeout 1500 -- 'bash tools/sanitize_decoder.sh'
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 800 -x -k "decoder_cluster_vs_oracle and (5-1-12 or 7-5-16 or 6-3-20-94)" > gpurun_out/san_dec.log 2>&1
tail -6 gpurun_out/san_dec.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/run_once.py caption --batch 5 --seconds 3.1 --reps 1 > gpurun_out/san_caption.log 2>&1
tail -6 gpurun_out/san_caption.log
