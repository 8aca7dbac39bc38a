#!/bin/bash
# A/B timing of prebuilt library variants (gst_b200/lib/var/*.so) on the GPU box: swaps each in as
# libgst_cuda.so and prints the kernel times of a short bench run.
cp gst_b200/lib/libgst_cuda.so /tmp/lib_orig.so
for v in gst_b200/lib/var/*.so; do
  cp "$v" gst_b200/lib/libgst_cuda.so
  for rep in 1 2; do
    python bench.py --no-e2e --no-cpu-baseline --steps 20 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$(basename $v)', round(d['value'],1), {k: round(v,4) for k,v in d['roofline']['kernel_ms_all'].items()})"
  done
done
cp /tmp/lib_orig.so gst_b200/lib/libgst_cuda.so
