"""Regenerates the committed golden fixtures.  Needs /root/reference (this container only).

  test1.gst / test1.dxt : codec/test/test1.png through the UNMODIFIED reference encoder
      (GenTC::DXTImage + GenTC::CompressDXT) and the encoder's PhysicalBlocks() -- the
      identity codec/test/codec_test.cpp:36-48 checks the GPU decoder against.
  synth512_s7.gst / .dxt : one seeded synthetic 512x512 image (tests/gst_fixtures.synth_image)
      through the same encoder, so GPU-box tests have a second real stream even if the
      reference-linked library did not travel.
  golden.json : sha256 of every intermediate of the reference-linked stitched decoder
      (oracle/ref_glue.cpp) on those two streams.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import gst_fixtures as fx  # noqa: E402

REF = os.environ.get("GST_REFERENCE", "/root/reference")


def main():
    L = fx.ref()
    assert L is not None, "build oracle/_ref first (oracle/build_ref.sh)"
    png = os.path.join(REF, "codec", "test", "test1.png").encode()
    cap = 1 << 22
    gst = np.empty(cap, np.uint8)
    dxt = np.empty(cap, np.uint8)
    w, h, n = C.c_int(), C.c_int(), C.c_size_t()
    rc = L.gstref_encode_file(png, C.byref(w), C.byref(h), gst.ctypes.data, cap, C.byref(n), dxt.ctypes.data, cap)
    assert rc == 0, rc
    gst, dxt = gst[: n.value], dxt[: w.value * h.value // 2]
    gst.tofile(os.path.join(HERE, "test1.gst"))
    dxt.tofile(os.path.join(HERE, "test1.dxt"))
    g2, d2 = fx.encode_rgb(fx.synth_image(512, 512, 7))
    g2.tofile(os.path.join(HERE, "synth512_s7.gst"))
    d2.tofile(os.path.join(HERE, "synth512_s7.dxt"))
    meta = {}
    for name, (g, d) in {"test1": (gst, dxt), "synth512_s7": (g2, d2)}.items():
        r = fx.ref_decode(g)
        meta[name] = dict(header=r["header"], gst_bytes=int(g.size), gst=fx.sha(g), physical_blocks=fx.sha(d),
                          ref_dxt=fx.sha(r["out"]), ref_symbols=fx.sha(r["symbols"]), ref_planes=fx.sha(r["planes"]),
                          ref_indices=fx.sha(r["indices"]),
                          ref_dxt_equals_physical_blocks=bool(np.array_equal(r["out"], d)))
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
