#!/bin/bash
# round-2 call 1: pipe-rate microbench (DPX ops), baseline bench of every config
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2_smi.txt 2>&1
nproc >> gpurun_out/r2_smi.txt; free -g >> gpurun_out/r2_smi.txt
./profiles/microbench/pipe_rates > gpurun_out/r2_pipe_rates.txt 2>&1
for c in 3 0 1 2 4; do
  python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2_base_cfg$c.json 2> gpurun_out/r2_base_cfg$c.err
done
tail -n 25 gpurun_out/r2_pipe_rates.txt
for c in 3 0 1 2 4; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_base_cfg$c.json").read().strip().splitlines()[-1])
    print($c, round(d["value"],1), "ms", round(d["ms_per_step"],4), d["roofline"]["kernel_ms_all"], "e2e", d["e2e"] and round(d["e2e"]["value"],1))
except Exception as e: print($c, "failed", e)
PY
done
