"""Experiment: one LoadCompressedDXTs-style step issued as P pages on P streams (the reference's photos_sf loop
uses its 4 out-of-order queues the same way).  Prints ms per step (wall clock over K steps, device synchronised)."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import gst_b200, gst_fixtures as fx
from gst_b200.capi import check, lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
K = 50
dec = gst_b200.Decoder(0)
streams = fx.encode_images(2048, 2048, [30000 + i for i in range(32)])
files = [g for g, _ in streams]
N = 2048 * 2048 // 16
for P in (1, 2, 4, 8):
    per = B // P
    pages = []
    for pg in range(P):
        blobs = [files[(pg * per + i) % 32] for i in range(per)]
        packed, hdrs = gst_b200.pack_batch(blobs)
        d_cmp, d_out = dec.malloc(packed.size), dec.malloc(8 * N * per)
        dec.upload(d_cmp, packed)
        harr = (gst_b200.capi.gst_header * per)(*[h.to_c() for h in hdrs])
        pages.append((harr, per, d_cmp, d_out))
    strs = [lib().gst_stream_next(dec.ctx) for _ in range(min(P, 4))] if P > 1 else [dec.GetDefaultCommandQueue()]
    def step():
        for i, (harr, per, d_cmp, d_out) in enumerate(pages):
            check(lib().gst_load_dxt_batch(dec.ctx, harr, per, strs[i % len(strs)], d_cmp.ptr, d_cmp.nbytes, d_out.ptr, None, 0, None))
    for _ in range(5):
        step()
    dec.sync()
    t0 = time.perf_counter()
    for _ in range(K):
        step()
    dec.sync()
    dt = (time.perf_counter() - t0) / K * 1e3
    print(f"B={B} pages={P}: {dt:.4f} ms/step  {B * 2048 * 2048 / dt / 1e6:.1f} GTexel/s", flush=True)
    for _, _, a, b in pages:
        a.free(); b.free()
dec.close()
