"""GPU rANS stream encoder against the reference's own CPU encoder:
gst_ans_encode_stream must emit, byte for byte, what ByteEncoder::EncodeBytes
(codec/entropy.cpp:174-265 = ans::ocl::NormalizeFrequencies + ans::EncodeInterleaved per group) emits for
the same symbols, and the CUDA decoder must read it back."""
import numpy as np
import pytest

import gst_b200
import gst_fixtures as fx

pytestmark = pytest.mark.gpu


def _symbol_sets():
    rng = np.random.default_rng(11)
    yield "laplace", np.clip(np.rint(rng.laplace(0.0, 5.0, size=4 * 8192)) + 128, 0, 255).astype(np.uint8)
    yield "uniform", rng.integers(0, 256, size=2 * 8192, dtype=np.uint8)          # ~2 bytes per symbol: most steps emit
    yield "constant", np.full(8192, 77, dtype=np.uint8)                           # F = 2048: never emits
    yield "two_symbols", rng.choice(np.array([3, 250], np.uint8), p=[0.999, 0.001], size=3 * 8192)
    yield "sparse_high", np.where(rng.random(8192) < 0.01, 255, 0).astype(np.uint8)
    yield "one_group_odd_words", np.clip(np.rint(rng.laplace(0.0, 1.2, size=8192)) + 128, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("name,symbols", list(_symbol_sets()), ids=[n for n, _ in _symbol_sets()])
def test_stream_matches_reference_encoder(decoder, ref_lib, name, symbols):
    want_freqs, want_stream = fx.encode_stream(symbols)
    freqs, stream = gst_b200.encode_stream(decoder, symbols)
    assert np.array_equal(freqs, want_freqs), "normalised frequencies differ from ans::ocl::NormalizeFrequencies"
    assert stream.size == want_stream.size, f"stream is {stream.size} bytes, the reference encoder emits {want_stream.size}"
    bad = np.flatnonzero(stream != want_stream)
    assert bad.size == 0, f"{bad.size} bytes differ from ByteEncoder::EncodeBytes, first at {bad[:8]}"


def test_encode_decode_round_trip_through_the_gst_container(decoder):
    """Random symbol planes -> GPU-encoded .gst -> CUDA decode == CPU oracle decode of the same container,
    and the decoded symbols are the ones that went in."""
    rng = np.random.default_rng(5)
    w, h = 512, 256
    n = (w // 4) * (h // 4)
    planes = np.clip(np.rint(rng.laplace(0.0, 6.0, size=6 * n)) + 128, 0, 255).astype(np.uint8)
    entries = 3000
    palette = np.zeros(16384, dtype=np.uint8)
    palette[: 4 * entries] = rng.integers(0, 256, size=4 * entries, dtype=np.uint8)
    # a bounded random walk over the palette: every delta fits a byte and every index stays in range
    idx = np.empty(n, dtype=np.int64)
    cur = 0
    steps = rng.integers(-128, 128, size=n)
    for i in range(n):
        nxt = cur + int(steps[i])
        if nxt < 0 or nxt >= entries:
            nxt = cur - int(steps[i])
        if nxt < 0 or nxt >= entries:
            nxt = cur
        idx[i] = cur = nxt
    deltas = np.diff(np.concatenate([[0], idx]))
    index_syms = (deltas + 128).astype(np.uint8)
    gst = gst_b200.build_gst(decoder, w, h, planes[: 2 * n], planes[2 * n:], palette, index_syms)
    res = decoder.decode_tapped([gst])
    o = fx.oracle_decode(gst)
    assert np.array_equal(res["symbols"][: 6 * n], planes), "decoded plane symbols are not the encoded ones"
    assert np.array_equal(res["symbols"], o["symbols"])
    assert np.array_equal(res["dxt"], o["out"])


def test_gpu_encoded_container_equals_cpu_encoded_container(decoder, ref_lib):
    rng = np.random.default_rng(9)
    w, h = 512, 256  # N = 8192 blocks: the smallest size whose index stream fills a group
    n = (w // 4) * (h // 4)
    y = rng.integers(100, 160, size=2 * n, dtype=np.uint8)
    c = rng.integers(120, 136, size=4 * n, dtype=np.uint8)
    pal = rng.integers(0, 256, size=8192, dtype=np.uint8)
    isym = rng.integers(126, 131, size=n, dtype=np.uint8)
    a = gst_b200.build_gst(decoder, w, h, y, c, pal, isym)
    b = fx.make_gst(w, h, y, c, pal, isym)
    assert np.array_equal(a, b)


def test_encoder_argument_checks(decoder):
    with pytest.raises(gst_b200.GstError):
        gst_b200.encode_stream(decoder, np.zeros(8191, np.uint8))
    with pytest.raises(gst_b200.GstError):
        gst_b200.encode_stream(decoder, np.zeros(0, np.uint8))


def test_many_distinct_gpu_encoded_images_in_one_batch(decoder):
    """48 DISTINCT 512x512 containers (random symbol planes, palettes and index walks, entropy-coded by the GPU
    encoder in a fraction of the time the CPU encoder needs) decoded in one LoadCompressedDXTs call: every
    image against the CPU oracle."""
    w = h = 512
    n = (w // 4) * (h // 4)
    files = []
    for seed in range(48):
        rng = np.random.default_rng(1000 + seed)
        scale = [1.5, 4.0, 9.0, 20.0][seed % 4]
        planes = np.clip(np.rint(rng.laplace(0.0, scale, size=6 * n)) + 128, 0, 255).astype(np.uint8)
        entries = int(rng.integers(300, 6000))
        pal_bytes = -(-4 * entries // 8192) * 8192
        palette = np.zeros(pal_bytes, dtype=np.uint8)
        palette[: 4 * entries] = rng.integers(0, 256, size=4 * entries, dtype=np.uint8)
        # a triangle-wave fold of a random walk: every index in range, every delta within +-100
        walk = np.cumsum(rng.integers(-100, 101, size=n))
        period = 2 * (entries - 1)
        m = np.mod(walk, period)
        idx = np.where(m < entries, m, period - m)
        deltas = np.diff(np.concatenate([[0], idx]))
        assert np.abs(deltas).max() <= 127 + 1 and idx.min() >= 0 and idx.max() < entries
        deltas = np.clip(deltas, -128, 127)
        index_syms = (deltas + 128).astype(np.uint8)
        files.append(gst_b200.build_gst(decoder, w, h, planes[: 2 * n], planes[2 * n:], palette, index_syms))
    assert len({f.tobytes() for f in files}) == 48
    out = decoder.DecompressDXTs(files, page=48).reshape(48, 8 * n)   # one page = one LoadCompressedDXTs call
    for i, f in enumerate(files):
        want = fx.oracle_decode(f, taps=False)["out"]
        assert np.array_equal(out[i], want), f"image {i} differs from the CPU oracle"
