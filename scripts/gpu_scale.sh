#!/bin/bash
# bench.py at N GPUs (strong scaling of the configured batch), configs 3 and 4 -> gpurun_out/<tag>_n<N>_cfg<k>.json
N=${1:-8}; tag=${2:-scale}
for c in 3 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$c bench.py --gpus $N --config $c --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/${tag}_n${N}_cfg$c.err | tail -1 > gpurun_out/${tag}_n${N}_cfg$c.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_n${N}_cfg$c.json"))
    print("cfg$c N=$N", round(d["value"],1), "ms", round(d["ms_per_step"],4), d["scaling"], "img/gpu", d["config"]["images_per_gpu"], "other", d["other_scaling"] and (round(d["other_scaling"]["value"],1), round(d["other_scaling"]["ms_per_step"],4)), "e2e", round(d["e2e"]["value"],1), d["e2e"].get("frames_per_s"), "res", round(d["e2e_resident"]["value"],1))
except Exception as e:
    print("failed", e); print(open("gpurun_out/${tag}_n${N}_cfg$c.err").read()[-1500:])
PY
done
