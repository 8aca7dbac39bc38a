"""CPU tests: the oracle (oracle/gst_oracle.c) pinned against the reference's own golden
vectors and, where oracle/_ref/libgst_ref.so is available, against the reference's own code."""
import json
import os

import numpy as np
import pytest

import gst_fixtures as fx


def _golden(name):
    gst = np.fromfile(os.path.join(fx.GOLDEN_DIR, name + ".gst"), dtype=np.uint8)
    dxt = np.fromfile(os.path.join(fx.GOLDEN_DIR, name + ".dxt"), dtype=np.uint8)
    meta = json.load(open(os.path.join(fx.GOLDEN_DIR, "golden.json")))[name]
    return gst, dxt, meta


@pytest.mark.parametrize("name", ["test1", "synth512_s7"])
def test_oracle_reproduces_encoder_blocks(name):
    """codec/test/codec_test.cpp:36-48: decoded blocks == the encoder's PhysicalBlocks()."""
    gst, dxt, meta = _golden(name)
    assert fx.sha(gst) == meta["gst"]
    o = fx.oracle_decode(gst)
    assert np.array_equal(o["out"], dxt)
    # every intermediate equals what the reference's own CPU code produced (oracle/ref_glue.cpp)
    assert fx.sha(o["symbols"]) == meta["ref_symbols"]
    assert fx.sha(o["planes"]) == meta["ref_planes"]
    assert fx.sha(o["indices"]) == meta["ref_indices"]
    assert fx.sha(o["out"]) == meta["ref_dxt"]


def test_oracle_vs_reference_library(ref_lib):
    """Same stream through oracle/gst_oracle.c and through ans::DecodeInterleaved +
    GenTC::InverseWavelet2D linked unmodified."""
    for gst in (fx.golden_test1()[0], fx.encode_image(512, 512, 10000)[0]):
        o, r = fx.oracle_decode(gst), fx.ref_decode(gst)
        for k in ("symbols", "planes", "indices", "out"):
            assert np.array_equal(o[k], r[k]), k


def test_oracle_inverse_wavelet_kat():
    """codec/test/wavelet_test.cpp:129-158, applied through the oracle's lifting on a tile whose
    top-left 4x4 corner holds the vector (only the level-4 pass is exercised here, via the
    reference library when present; the oracle's tile routine runs all five levels, so compare
    the single level through the exposed 1-D building block instead)."""
    xs = np.array([63, 64, 0, -1, 66, 60, 6, 9, 0, 2, -2, -2, 7, -18, 16, 36], dtype=np.int32).reshape(4, 4)
    expected = np.array([63, 63, 63, 63, 63, 63, 64, 63, 63, 65, 62, 64, 62, 65, 31, 69], dtype=np.int32).reshape(4, 4)

    def tdiv(a, b):
        return int(a / b) if a * b >= 0 else -int((-a) / b)  # C truncating division

    def lift(v):
        n, mid = len(v), len(v) // 2
        out = [0] * n

        def mirror(i):
            if i >= n:
                i = i - (i - n + 2)
            return abs(i)
        for i in range(0, n, 2):
            out[i] = v[i // 2] - tdiv(v[mid + mirror(i - 1) // 2] + v[mid + mirror(i + 1) // 2] + 2, 4)
        for i in range(1, n, 2):
            out[i] = v[mid + i // 2] + tdiv(out[mirror(i - 1)] + out[mirror(i + 1)], 2)
        return out
    rows = np.array([lift(list(r)) for r in xs])
    both = np.array([lift(list(c)) for c in rows.T]).T
    assert np.array_equal(both, expected)


def test_oracle_tile_matches_reference_wavelet(ref_lib):
    """Random tiles: oracle 5-level tile transform == InverseWavelet2D applied for dim 2..32
    (the stitching of oracle/ref_glue.cpp), including the (char) truncation."""
    rng = np.random.default_rng(3)
    O = fx.oracle()
    for trial in range(8):
        tile = rng.integers(0, 256, size=1024, dtype=np.uint8) if trial % 2 else \
            np.clip(np.rint(rng.laplace(0, 5, size=1024)) + 128, 0, 255).astype(np.uint8)
        got = np.empty(1024, dtype=np.int8)
        O.gsto_inverse_wavelet_tile(tile.ctypes.data, got.ctypes.data)
        blk = tile.astype(np.int16) - 128
        for d in (2, 4, 8, 16, 32):
            dst = blk.copy()
            ref_lib.gstref_inverse_wavelet2d(blk.ctypes.data, dst.ctypes.data, d, 64)
            blk = dst
        assert np.array_equal(got, blk.astype(np.int8))


def test_oracle_table_matches_reference_expansion():
    """ans/ans_ocl_test.cpp:64-155: the table is the plain expansion of the normalised
    frequencies (symbol, freq, cumulative freq per slot)."""
    O = fx.oracle()
    for F in ([614, 410, 205, 614, 205], [80, 300, 2, 14, 1, 1, 1, 1649], [2048], [1] * 255 + [1793]):
        assert sum(F) == 2048
        freqs = np.zeros(256, dtype=np.uint16)
        freqs[:len(F)] = F
        tf, tc, ts = np.empty(2048, np.uint16), np.empty(2048, np.uint16), np.empty(2048, np.uint8)
        O.gsto_build_table(freqs.ctypes.data, tf.ctypes.data, tc.ctypes.data, ts.ctypes.data)
        es = np.repeat(np.arange(len(F)), F)
        cum = np.concatenate([[0], np.cumsum(F)[:-1]])
        assert np.array_equal(ts, es.astype(np.uint8))
        assert np.array_equal(tf, np.asarray(F, dtype=np.uint16)[es])
        assert np.array_equal(tc, cum.astype(np.uint16)[es])


def test_oracle_rans_group_vs_reference(ref_lib):
    """ans/ans_test.cpp:200-246 style round trip: reference EncodeInterleaved -> oracle group
    decode, for 32, 24 and 1 interleaved lanes (ans/ans_ocl_test.cpp:157-275)."""
    rng = np.random.default_rng(0)
    O = fx.oracle()
    F = fx.ref_histogram(np.array([12, 14, 17, 1, 1, 2, 372], dtype=np.uint32))
    p = F / F.sum()
    freqs = np.zeros(256, dtype=np.uint16)
    freqs[:F.size] = F
    tf, tc, ts = np.empty(2048, np.uint16), np.empty(2048, np.uint16), np.empty(2048, np.uint8)
    O.gsto_build_table(freqs.ctypes.data, tf.ctypes.data, tc.ctypes.data, ts.ctypes.data)
    for lanes in (32, 24, 1):
        syms = rng.choice(F.size, size=lanes * 256, p=p).astype(np.uint8)
        enc = fx.ref_encode_interleaved(syms, F, lanes)
        if enc.size & 3:
            enc = np.concatenate([np.zeros(2, np.uint8), enc])
        data = np.concatenate([np.array([4 + enc.size], dtype="<u4").view(np.uint8), enc])
        out = np.empty(lanes * 256, dtype=np.uint8)
        O.gsto_ans_decode_group(tf.ctypes.data, tc.ctypes.data, ts.ctypes.data, data.ctypes.data, 0, lanes,
                                out.ctypes.data)
        assert np.array_equal(out, syms)


def test_make_gst_round_trip(ref_lib):
    """tests/gst_fixtures.make_gst builds containers the oracle and the reference-linked decoder
    both accept, with arbitrary symbol planes."""
    gst = fx.random_gst(256, 512, seed=5, palette_entries=3000, plane_mode="uniform")
    o, r = fx.oracle_decode(gst), fx.ref_decode(gst)
    for k in ("symbols", "planes", "indices", "out"):
        assert np.array_equal(o[k], r[k]), k
    assert o["indices"].min() >= 0 and o["indices"].max() < 3000


def test_oracle_rgb_mode_consistent_with_dxt():
    """assemble_rgb (codec/assemble.cl:83-129) decodes the same endpoints / indices as
    assemble_dxt: texel colour k of a block is palette entry (word >> 2k) & 3 of the 565
    endpoints expanded to 888."""
    gst, _ = fx.golden_test1()
    d = fx.oracle_decode(gst)["out"].view("<u2").reshape(-1, 4)
    rgb = fx.oracle_decode(gst, mode=1, taps=False)["out"].reshape(512, 512, 3)
    for blk in (0, 77, 5000, 16383):
        ep = [int(d[blk, 0]), int(d[blk, 1])]
        word = int(d[blk, 2]) | (int(d[blk, 3]) << 16)
        pal = []
        for e in ep:
            r, g, b = e >> 11, (e >> 5) & 63, e & 31
            pal.append(((r << 3) | (r >> 2), (g << 2) | (g >> 4), (b << 3) | (b >> 2)))
        pal.append(tuple((2 * a + b) // 3 for a, b in zip(pal[0], pal[1])))
        pal.append(tuple((a + 2 * b) // 3 for a, b in zip(pal[0], pal[1])))
        by, bx = divmod(blk, 128)
        for k in range(16):
            want = pal[(word >> (2 * k)) & 3]
            assert tuple(int(x) for x in rgb[4 * by + k // 4, 4 * bx + k % 4]) == want
