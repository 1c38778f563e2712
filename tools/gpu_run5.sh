mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/t_all.log 2>&1
tail -4 gpurun_out/t_all.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
tail -c 300 gpurun_out/bench_c1.err
python bench.py --config 2 --steps 3 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
tail -c 300 gpurun_out/bench_c2.err
python bench.py --config 4 --steps 2 --warmup 3 --cpu-sample 2 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
tail -c 300 gpurun_out/bench_c4.err
python bench.py --config 3 --steps 2 --warmup 3 --cpu-sample 2 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
tail -c 300 gpurun_out/bench_c3.err
