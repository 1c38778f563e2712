#!/bin/bash
# usage: tools/sweep_env.sh VAR v1 v2 ...   -> one bench line summary per value of the environment variable
var=$1; shift
for v in "$@"; do
  env $var=$v timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_${var}_$v.json 2> gpurun_out/bench_${var}_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${var}_$v.json").read().strip().splitlines()[-1])
k=d["kernels"]
print("$var=$v:", round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), " ".join(f"{n}={k[n]['ms_per_step']:.3f}" for n in k if not n.startswith("gemm_pw") and not n.startswith("dec")))
PY
done
