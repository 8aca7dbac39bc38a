#!/bin/bash
# A/B timing of prebuilt library variants (gst_b200/lib/var/*.so) on the GPU box.  bench.py checks 32 distinct
# images bit for bit before it times anything, so a wrong variant fails instead of printing a number.
# usage: bash scripts/ab.sh [reps] [extra bench args]
reps=${1:-2}; shift
for v in gst_b200/lib/var/*.so; do
  for rep in $(seq $reps); do
    GST_LIB=$PWD/$v python bench.py --no-e2e --no-cpu-baseline --steps 20 "$@" 2>gpurun_out/ab_$(basename $v .so).err | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$(basename $v .so)', round(d['value'],1), 'ms', round(d['ms_per_step'],4), {k: round(v['ms'],4) for k,v in d['roofline']['kernels'].items()})
except Exception as e: print('$(basename $v .so)', 'FAILED', e)"
  done
done
