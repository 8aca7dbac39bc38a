"""Debug helper: decode one stream with the stage taps and report where each stage first differs
from the CPU oracle."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import gst_b200, gst_fixtures as fx

dec = gst_b200.Decoder(0)
gst = fx.golden_test1()[0] if len(sys.argv) < 2 else np.fromfile(sys.argv[1], dtype=np.uint8)
res = dec.decode_tapped([gst])
o = fx.oracle_decode(gst)
N = res["hdrs"][0].num_blocks
bx = res["hdrs"][0].width // 4
for name, got, want in (("symbols", res["symbols"], o["symbols"]), ("indices", res["indices"], o["indices"]),
                        ("planes", res["planes"], o["planes"]), ("dxt", res["dxt"], o["out"])):
    bad = np.flatnonzero(got != want)
    print(f"{name}: {bad.size} of {want.size} differ", "first:", bad[:16])
    if bad.size and name == "planes":
        pl = bad // N
        print("  per plane:", np.bincount(pl, minlength=6))
        i = bad[0] % N
        print("  first bad plane", bad[0] // N, "row", i // bx, "col", i % bx)
        p0 = got[:N].reshape(-1, bx); w0 = want[:N].reshape(-1, bx)
        print("  got rows 0..3 cols 0..15\n", p0[:4, :16], "\n  want\n", w0[:4, :16])
        diff = (p0 != w0)
        print("  bad rows (tile 0):", np.flatnonzero(diff[:32, :32].any(axis=1)), "bad cols:", np.flatnonzero(diff[:32, :32].any(axis=0)))
    if bad.size and name == "indices":
        print("  got", got[:8], "want", want[:8], " ... got", got[bad[0]:bad[0]+4], "want", want[bad[0]:bad[0]+4])
