#!/bin/bash
# One GPU iteration: parity tests, short bench (kernel times), ncu capture of the two heavy kernels.
# usage (on the GPU box): bash scripts/gpu_iter.sh <tag> [notest]
tag=${1:-iter}
if [ "$2" != "notest" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
python bench.py --no-e2e --no-cpu-baseline --steps 20 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('GTexel/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), {k: round(v['ms'],4) for k,v in d['roofline']['kernels'].items()})"
ncu --set full --import-source on --clock-control none -k regex:"rans_streams|wavelet_assemble" -c 2 -o gpurun_out/prof_$tag python bench.py --no-e2e --no-cpu-baseline --steps 1 --warmup 0 --images 256 > /dev/null 2>&1
ls gpurun_out/
