#!/bin/bash
# usage: tools/quick_bench.sh TAG   -> fused-MLP numerics check + one bench line summary (gpurun_out/bench_TAG.json)
tag=${1:-quick}
timeout 120 python tools/fused_check.py 2>&1 | tail -4
timeout 250 python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/b.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
k=d["kernels"]
print(round(d["value"]), round(d["ms_per_step"],3), "seq", round(d["value_sequential"]["ms_per_step"],3), "e2e", round(d["e2e"]["value"]))
print(" ".join(n+"=%.3f"%k[n]["ms_per_step"] for n in k))
PY
tail -3 gpurun_out/b.err
