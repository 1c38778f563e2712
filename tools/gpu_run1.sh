mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -s -k "decoder_cluster_vs_oracle" > gpurun_out/t_cluster.log 2>&1
tail -5 gpurun_out/t_cluster.log
python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -s -k "tcgen05 or fused_mlp or encoder_outputs or graph_and_eager or beam_search_vs_oracle" > gpurun_out/t_enc.log 2>&1
tail -5 gpurun_out/t_enc.log
python -m pytest tests/test_bench_parity.py -m gpu -q --timeout 600 -s > gpurun_out/t_bench.log 2>&1
tail -5 gpurun_out/t_bench.log
