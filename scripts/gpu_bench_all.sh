#!/bin/bash
# bench line of every BASELINE.json config on one GPU -> gpurun_out/<tag>_cfg<k>.json
tag=${1:-bench}; shift
mkdir -p gpurun_out
for c in 3 0 1 2 4; do
  python bench.py --config $c --steps 20 --warmup 5 "$@" > gpurun_out/${tag}_cfg$c.json 2> gpurun_out/${tag}_cfg$c.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_cfg$c.json").read().strip().splitlines()[-1])
    k={a: round(b["ms"],4) for a,b in d["roofline"]["kernels"].items()}
    print($c, round(d["value"],1), "GTexel/s  ms", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), k, "e2e", d["e2e"] and round(d["e2e"]["value"],1), d["e2e"] and d["e2e"].get("frames_per_s"), "launches", d["gpu_launches"])
except Exception as e:
    print($c, "failed", e); print(open("gpurun_out/${tag}_cfg$c.err").read()[-1500:])
PY
done
