"""The five ANS_OpenCL tests of ans/ans_ocl_test.cpp:64-345 against the CUDA rANS decoder
(AnsDecoder = ans::ocl::OpenCLDecoder), plus lane counts in between."""
import numpy as np
import pytest

import gst_b200
import gst_fixtures as fx

pytestmark = pytest.mark.gpu


def _expected_table(F):
    nf = fx.ref_histogram(np.asarray(F, dtype=np.uint32)) if fx.ref() is not None else gst_b200.normalize_frequencies(F)
    assert nf.sum() == gst_b200.kANSTableSize
    sym = np.repeat(np.arange(nf.size), nf)
    cum = np.concatenate([[0], np.cumsum(nf)[:-1]])
    return sym.astype(np.uint8), nf[sym].astype(np.uint16), cum[sym].astype(np.uint16)


def _gen_symbols(rng, F, n):
    F = np.asarray(F, dtype=np.float64)
    return rng.choice(F.size, size=n, p=F / F.sum()).astype(np.uint8)


def _encode(symbols, F, lanes):
    nf = fx.ref_histogram(np.asarray(F, dtype=np.uint32))
    enc = fx.ref_encode_interleaved(symbols, nf, lanes)
    return enc[: enc.size - 4 * lanes], enc[enc.size - 4 * lanes:].view("<u4").copy()


def test_initialization(decoder):
    """ans/ans_ocl_test.cpp:64-108."""
    F = [3, 2, 1, 4, 3]
    d = gst_b200.AnsDecoder(decoder, F, 1)
    sym, fr, cum = _expected_table(F)
    assert np.array_equal(d.GetSymbols(), sym)
    assert np.array_equal(d.GetFrequencies(), fr)
    assert np.array_equal(d.GetCumulativeFrequencies(), cum)


def test_table_rebuilding(decoder):
    """ans/ans_ocl_test.cpp:110-155."""
    d = gst_b200.AnsDecoder(decoder, [3, 2, 1, 4, 3, 406], 1)
    new_F = [80, 300, 2, 14, 1, 1, 1, 20]
    d.RebuildTable(new_F)
    sym, fr, cum = _expected_table(new_F)
    assert np.array_equal(d.GetSymbols(), sym)
    assert np.array_equal(d.GetFrequencies(), fr)
    assert np.array_equal(d.GetCumulativeFrequencies(), cum)


def test_build_tables_from_stream_frequencies(decoder):
    """Stage 1 on the four frequency blocks of a real .gst, against the oracle's restatement of
    ans/build_table.cl."""
    gst = fx.golden_test1()[0]
    freqs = gst[28:28 + 2048].view("<u2").reshape(4, 256)
    sym, fr, cum = decoder.build_tables(freqs)
    O = fx.oracle()
    for t in range(4):
        tf, tc, ts = np.empty(2048, np.uint16), np.empty(2048, np.uint16), np.empty(2048, np.uint8)
        f = np.ascontiguousarray(freqs[t])
        O.gsto_build_table(f.ctypes.data, tf.ctypes.data, tc.ctypes.data, ts.ctypes.data)
        assert np.array_equal(sym[t], ts) and np.array_equal(fr[t], tf) and np.array_equal(cum[t], tc)


def test_decode_single_stream(decoder, ref_lib):
    """ans/ans_ocl_test.cpp:157-218."""
    rng = np.random.default_rng(0)
    F = [12, 14, 17, 1, 1, 2, 372]
    symbols = _gen_symbols(rng, F, 256)
    data, states = _encode(symbols, F, 1)
    d = gst_b200.AnsDecoder(decoder, F, 1)
    out = d.Decode(int(states[0]), data)
    assert np.array_equal(out, symbols)


@pytest.mark.parametrize("lanes", [24, 32, 2, 7, 31])
def test_decode_interleaved_streams(decoder, ref_lib, lanes):
    """ans/ans_ocl_test.cpp:220-275 (24 interleaved streams), and other lane counts."""
    rng = np.random.default_rng(lanes)
    F = [32, 186, 54, 8, 1, 1, 1, 12, 500]
    symbols = _gen_symbols(rng, F, 256 * lanes)
    data, states = _encode(symbols, F, lanes)
    d = gst_b200.AnsDecoder(decoder, F, lanes)
    out = d.Decode(states, data)
    assert out.shape == (lanes, 256)
    assert np.array_equal(out.reshape(-1), symbols)


def test_decode_multiple_groups(decoder, ref_lib):
    """ans/ans_ocl_test.cpp:277-345: independent groups of interleaved streams in one launch."""
    rng = np.random.default_rng(2)
    F = [65, 4, 6, 132, 135, 64, 879, 87, 456, 13, 2, 12, 33, 16, 546, 987, 98, 74, 65, 43, 21, 32, 1]
    lanes, groups = 16, 5
    all_syms, datas, states = [], [], []
    for g in range(groups):
        s = _gen_symbols(rng, F, 256 * lanes)
        dt, st = _encode(s, F, lanes)
        all_syms.append(s)
        datas.append(dt)
        states.append(st)
    d = gst_b200.AnsDecoder(decoder, F, lanes)
    out = d.Decode(np.concatenate(states), datas)
    assert np.array_equal(out.reshape(-1), np.concatenate(all_syms))


def test_incompressible_group(decoder, ref_lib):
    """Near-uniform symbols: almost every lane renormalises on every step, the worst case for
    the staging ring (64 B consumed per symbol)."""
    rng = np.random.default_rng(5)
    F = [8] * 256
    symbols = rng.integers(0, 256, size=8192).astype(np.uint8)
    data, states = _encode(symbols, F, 32)
    d = gst_b200.AnsDecoder(decoder, F, 32)
    assert np.array_equal(d.Decode(states, data).reshape(-1), symbols)
