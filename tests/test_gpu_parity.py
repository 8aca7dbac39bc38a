"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same
.gst inputs.  Everything on this path is integer, so the bar is bit-exact."""
import numpy as np
import pytest

import gst_fixtures as fx

pytestmark = pytest.mark.gpu


def _check_stages(res, i, gst):
    """Compare every tapped stage of image i of a decode_tapped() result with the oracle."""
    o = fx.oracle_decode(gst)
    hdrs = res["hdrs"]
    N = hdrs[0].num_blocks
    sym_off = sum(7 * N + h.palette_bytes for h in hdrs[:i])
    S = 7 * N + hdrs[i].palette_bytes
    assert np.array_equal(res["symbols"][sym_off:sym_off + S], o["symbols"]), "stage 2 (rANS symbols) differs"
    assert np.array_equal(res["indices"][i * N:(i + 1) * N], o["indices"]), "stage 3 (palette indices) differs"
    assert np.array_equal(res["planes"][i * 6 * N:(i + 1) * 6 * N], o["planes"]), "stage 4 (wavelet planes) differs"
    assert np.array_equal(res["dxt"][i * 8 * N:(i + 1) * 8 * N], o["out"]), "stage 5 (DXT1 blocks) differs"


def test_codec_test_identity(decoder):
    """codec/test/codec_test.cpp:36-48: decoder output == encoder's PhysicalBlocks() on test1.png."""
    gst, golden = fx.golden_test1()
    out = decoder.DecompressDXT(gst)
    assert out.size == golden.size
    bad = np.flatnonzero(out.view(np.uint64) != golden.view(np.uint64))
    assert bad.size == 0, f"{bad.size} of {golden.size // 8} blocks differ, first at {bad[:8]}"


def test_stages_golden_streams(decoder):
    for name in ("test1", "synth512_s7"):
        gst = np.fromfile(f"{fx.GOLDEN_DIR}/{name}.gst", dtype=np.uint8)
        res = decoder.decode_tapped([gst])
        _check_stages(res, 0, gst)
