"""Where does a streamed frame's time go?  gst_streamer_play over 600 frames of 1920x1024 with and without the
read-back, staged and direct uploads, several group sizes and depths.  Prints frames per second."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import gst_b200, gst_fixtures as fx

dec = gst_b200.Decoder(0)
w, h, n = 1920, 1024, 600
frames = fx.encode_motion(w, h, 40000, 32)
per = w * h // 2
pins = []
for g, _ in frames:
    pb = dec.pinned(g.size); pb.array[:] = g; pins.append(pb)
ptrs = (C.c_void_p * n)(*[pins[f % 32].ptr for f in range(n)])
lens = (C.c_size_t * n)(*[pins[f % 32].nbytes for f in range(n)])
host = dec.pinned(per * n)
d_out = dec.malloc(per * n)
for depth in (4, 8):
    st = gst_b200.FrameStreamer(dec, w, h, depth=depth)
    for group in (1, 4, 16):
        for name, kw in (("decode only (frames stay in a device buffer)", dict(dev_out=d_out.ptr)),
                         ("decode + read-back", dict(host_out=host.ptr)),
                         ("decode + read-back, staged upload", dict(host_out=host.ptr, direct=False))):
            kw.setdefault("direct", True)
            st.play(ptrs, lens, n, group=group, **kw)
            t0 = time.perf_counter()
            for _ in range(3):
                st.play(ptrs, lens, n, group=group, **kw)
            dt = (time.perf_counter() - t0) / 3
            print(f"depth {depth} group {group:2d} {name:48s} {n / dt:9.0f} frames/s  {dt / n * 1e6:6.1f} us/frame", flush=True)
    st.close()
dec.close()
