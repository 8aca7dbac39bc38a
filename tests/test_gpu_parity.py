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


def test_batch_mixed_streams_all_stages(decoder):
    """LoadCompressedDXTs with several different 512x512 streams in one call (photos_sf
    layout): every stage of every image against the oracle."""
    files = [fx.golden_test1()[0], np.fromfile(f"{fx.GOLDEN_DIR}/synth512_s7.gst", dtype=np.uint8)]
    files += [fx.encode_image(512, 512, 10000 + i)[0] for i in range(3)]
    files += [fx.random_gst(512, 512, seed=21, palette_entries=5000)]  # palette_bytes = 24576
    res = decoder.decode_tapped(files)
    assert len({h.palette_bytes for h in res["hdrs"]}) > 1
    for i, f in enumerate(files):
        _check_stages(res, i, f)


@pytest.mark.parametrize("mode", ["uniform", "extreme", "laplace"])
def test_adversarial_symbol_planes(decoder, mode):
    """Symbol planes no encoder would emit (uniform / extreme bytes): exercises the int16
    work-tile bound of the wavelet, the (char) truncation and the unmasked 565 pack."""
    gst = fx.random_gst(512, 256, seed=3 + len(mode), palette_entries=2048, plane_mode=mode)
    res = decoder.decode_tapped([gst])
    _check_stages(res, 0, gst)


@pytest.mark.parametrize("wh", [(256, 512), (1024, 128), (128, 1024), (1920, 1024), (640, 1024)])
def test_geometries(decoder, wh, ref_lib):
    """Tile-major planes vs raster indices on non-square / non-power-of-two block grids
    (tiles_x = 2, 8, 1, 15, 5)."""
    w, h = wh
    if (w, h) == (1920, 1024):
        gst, golden = fx.encode_image(w, h, 40000)
    else:
        gst, golden = fx.random_gst(w, h, seed=w + h, palette_entries=1500), None
    res = decoder.decode_tapped([gst, gst])
    _check_stages(res, 0, gst)
    _check_stages(res, 1, gst)
    if golden is not None:
        assert fx.matches_golden(res["dxt"][: 8 * res["hdrs"][0].num_blocks], golden)


def test_rgb_output_mode(decoder):
    """LoadRGB / assemble_rgb (codec/assemble.cl:83-129) against the oracle, single and batched."""
    g1 = fx.golden_test1()[0]
    g2 = fx.random_gst(512, 512, seed=9, palette_entries=700, plane_mode="uniform")
    for g in (g1, g2):
        want = fx.oracle_decode(g, mode=1, taps=False)["out"]
        assert np.array_equal(decoder.DecompressDXT(g, mode=1), want)
    out = decoder.DecompressDXTs([g1, g2, g1], page=2, mode=1)
    per = 512 * 512 * 3
    assert np.array_equal(out[per:2 * per], fx.oracle_decode(g2, mode=1, taps=False)["out"])
    assert np.array_equal(out[2 * per:], out[:per])


def test_large_single_texture(decoder):
    """configs[2]: one 4096x4096 texture (128 rANS groups per plane)."""
    gst, golden = fx.encode_image(4096, 4096, 20000)
    out = decoder.DecompressDXT(gst)
    assert fx.matches_golden(out, golden)
    res = decoder.decode_tapped([gst])
    _check_stages(res, 0, gst)


def test_big_batch_tiled(decoder):
    """configs[3] shape at a size that runs in seconds: 96 x 2048x2048 tiled from 4 distinct
    streams; every output image must equal the encoder's PhysicalBlocks() of its source, and
    equal inputs must give equal outputs wherever they sit in the batch."""
    srcs = [fx.encode_image(2048, 2048, 30000 + i) for i in range(4)]
    order = [(7 * i) % 4 for i in range(96)]
    out = decoder.DecompressDXTs([srcs[j][0] for j in order], page=96)
    per = 2048 * 2048 // 2
    for pos, j in enumerate(order):
        assert fx.matches_golden(out[pos * per:(pos + 1) * per], srcs[j][1]), f"image {pos}"
    want = fx.oracle_decode(srcs[0][0], taps=False)["out"]
    assert np.array_equal(out[:per], want)
    # the same batch in pages of 16 over the four work streams
    out2 = decoder.DecompressDXTs([srcs[j][0] for j in order], page=16)
    assert np.array_equal(out, out2)


def test_async_api_events_and_scratch_arena(decoder):
    """LoadCompressedDXT(s) on caller-owned device buffers and a work queue, ordered by events,
    with PreallocateDecompressor (bump arena, never reset: codec/decoder.cpp:74-82)."""
    import gst_b200
    files = [fx.encode_image(512, 512, 10000 + i)[0] for i in range(4)]
    goldens = [fx.encode_image(512, 512, 10000 + i)[1] for i in range(4)]
    packed, hdrs = gst_b200.pack_batch(files)
    need = sum(gst_b200.required_scratch_mem(h) for h in hdrs)
    decoder.PreallocateDecompressor(2 * need)
    try:
        q_up, q = decoder.GetNextQueue(), decoder.GetNextQueue()
        assert q_up != q
        d_cmp, d_out = decoder.malloc(packed.size), decoder.malloc(4 * 131072)
        pin = decoder.pinned(packed.size)
        pin.array[:] = packed
        decoder.upload(d_cmp, pin, stream=q_up)
        copied = decoder.record(q_up)
        for _ in range(2):  # two calls fit the arena
            done = decoder.LoadCompressedDXTs(hdrs, q, d_cmp, d_out, init=[copied])
            done.wait()
            out = decoder.download(d_out)
            assert np.array_equal(out, np.concatenate(goldens))
            done.destroy()
        with pytest.raises(gst_b200.GstError) as e:  # the arena is never reset
            for _ in range(64):
                decoder.LoadCompressedDXTs(hdrs, q, d_cmp, d_out, want_event=False)
        assert e.value.code == -4
        decoder.sync()
    finally:
        decoder.FreeDecompressor()
    # without an arena: single-image entry point
    one, h1 = gst_b200.pack_batch(files[:1])
    d1 = decoder.malloc(one.size)
    decoder.upload(d1, one)
    ev = decoder.LoadCompressedDXT(h1[0], decoder.GetDefaultCommandQueue(), d1, d_out)
    ev.wait()
    assert np.array_equal(decoder.download(d_out, 131072), goldens[0])


def test_errors_are_reported_not_swallowed(decoder):
    import gst_b200
    files = [fx.golden_test1()[0]]
    packed, hdrs = gst_b200.pack_batch(files)
    d_cmp, d_out = decoder.malloc(packed.size), decoder.malloc(131072)
    decoder.upload(d_cmp, packed)
    with pytest.raises(gst_b200.GstError):  # buffer shorter than the headers say
        decoder.LoadCompressedDXTs(hdrs, decoder.GetDefaultCommandQueue(), d_cmp, d_out, cmp_bytes=packed.size - 8)
    with pytest.raises(gst_b200.GstError):
        decoder.DecompressDXT(files[0][:5000])


def test_deterministic(decoder):
    gst = fx.encode_image(2048, 2048, 30001)[0]
    a = decoder.DecompressDXTs([gst] * 8, page=8)
    b = decoder.DecompressDXTs([gst] * 8, page=3)
    assert np.array_equal(a, b)


def test_config1_batch_128_small_textures(decoder):
    """BASELINE.json configs[1]: a batch of 128 512x512 textures in ONE call (the photos_sf layout,
    demo/photos_sf.cpp:753-795, with all 128 in a single page): 8 distinct reference-encoded images
    tiled, every output equal to the encoder's PhysicalBlocks() of its source."""
    srcs = [fx.encode_image(512, 512, 10000 + i) for i in range(8)]
    order = [(5 * i + 3) % 8 for i in range(128)]
    out = decoder.DecompressDXTs([srcs[j][0] for j in order], page=128)
    per = 512 * 512 // 2
    for pos, j in enumerate(order):
        assert np.array_equal(out[pos * per:(pos + 1) * per], srcs[j][1]), f"image {pos}"
    assert np.array_equal(out[:per], fx.oracle_decode(srcs[order[0]][0], taps=False)["out"])


def test_config4_frame_sequence_streamed(decoder):
    """BASELINE.json configs[4]: a 1920x1024 frame sequence streamed one frame per page through
    the work streams (demo/demo.cpp:145-243 decodes frameNNNN.gtc one at a time); frames cycle
    over 3 distinct reference-encoded images, outputs stay in device memory and are read back
    once at the end."""
    srcs = [fx.encode_image(1920, 1024, 40000 + i) for i in range(3)]
    n = 24
    order = [i % 3 for i in range(n)]
    per = 1920 * 1024 // 2
    d_out = decoder.malloc(per * n)
    decoder.memset(d_out, 0xEE)
    pins = []
    for g, _ in srcs:
        pb = decoder.pinned(g.size)
        pb.array[:] = g
        pins.append(pb)
    decoder.LoadHostBatch([pins[j] for j in order], d_out, page=1)
    got = decoder.download(d_out)
    for pos, j in enumerate(order):
        assert fx.matches_golden(got[pos * per:(pos + 1) * per], srcs[j][1]), f"frame {pos}"
    d_out.free()


@pytest.mark.parametrize("pinned", [False, True])
def test_load_host_batch_resident(decoder, pinned):
    """gst_load_host_batch with pageable and with page-locked sources: identical textures on the
    device, ragged last page included."""
    srcs = [fx.encode_image(512, 512, 10000 + i) for i in range(4)] + [(fx.golden_test1())]
    order = [(3 * i) % 5 for i in range(23)]  # ragged last page
    files = []
    for j in order:
        g = srcs[j][0]
        if pinned:
            pb = decoder.pinned(g.size)
            pb.array[:] = g
            files.append(pb)
        else:
            files.append(g)
    per = 512 * 512 // 2
    d_out = decoder.malloc(per * len(order))
    decoder.memset(d_out, 0x55)
    decoder.LoadHostBatch(files, d_out, page=4)
    got = decoder.download(d_out)
    for pos, j in enumerate(order):
        assert np.array_equal(got[pos * per:(pos + 1) * per], srcs[j][1]), f"image {pos}"
    d_out.free()


@pytest.mark.parametrize("mode", [0, 1])
def test_frame_streamer(decoder, mode):
    """gst_streamer_*: the demo player loop (demo/demo.cpp:145-243) with 3 frames in flight, DXT1
    and RGB8 (LoadRGB) outputs; every frame checked, including after the slots wrapped around."""
    import gst_b200
    srcs = [fx.encode_image(1920, 1024, 40000 + i) for i in range(3)]
    want = [fx.oracle_decode(g, mode=mode, taps=False)["out"] for g, _ in srcs]
    st = gst_b200.FrameStreamer(decoder, 1920, 1024, depth=3, mode=mode)
    try:
        tickets = []
        for f in range(10):
            tickets.append(st.submit(srcs[f % 3][0]))
            if f >= 2:  # consume two frames behind the producer
                k = f - 2
                got = st.read(tickets[k])
                assert np.array_equal(got, want[k % 3]), f"frame {k}"
        for k in (8, 9):
            assert np.array_equal(st.read(tickets[k]), want[k % 3]), f"frame {k}"
        with pytest.raises(gst_b200.GstError):
            st.wait(tickets[0])  # long gone: its slot holds a later frame
        with pytest.raises(gst_b200.GstError):
            st.submit(srcs[0][0][:100])  # truncated frame
        with pytest.raises(gst_b200.GstError):
            st.submit(fx.golden_test1()[0])  # 512x512 into a 1920x1024 streamer
    finally:
        st.close()


def _sweeping_gst(width, height, palette_entries, seed):
    """Like fx.random_gst, but the palette index sweeps the whole palette (a triangle wave with steps
    of up to 127), so indices above 2^16 really occur."""
    rng = np.random.default_rng(seed)
    n = (width // 4) * (height // 4)
    planes = np.clip(np.rint(rng.laplace(0.0, 6.0, size=6 * n)) + 128, 0, 255).astype(np.uint8)
    pal_bytes = -(-palette_entries * 4 // fx.GROUP) * fx.GROUP
    palette = np.zeros(pal_bytes, dtype=np.uint8)
    palette[: palette_entries * 4] = rng.integers(0, 256, size=palette_entries * 4, dtype=np.uint8)
    idx = np.empty(n, dtype=np.int64)
    cur, direction = 0, 1
    steps = rng.integers(90, 128, size=n)
    for i in range(n):
        nxt = cur + direction * int(steps[i])
        if nxt < 0 or nxt >= palette_entries:
            direction = -direction
            nxt = cur + direction * int(steps[i])
        idx[i] = cur = nxt
    deltas = np.diff(np.concatenate([[0], idx])) + 128
    assert deltas.min() >= 0 and deltas.max() <= 255 and idx.max() > 66000
    return fx.make_gst(width, height, planes[: 2 * n], planes[2 * n:], palette, deltas.astype(np.uint8))


def test_wide_palette_uses_32_bit_index_sums(decoder):
    """More than 65536 palette entries: the per-block index suffix sums no longer fit 16 bits and the
    u32 form of S is used (the reference's decoded_indices are int32, codec/decoder.cpp:302).  Two
    images in one call, one of them with a small palette, so the batch-wide choice is exercised."""
    big = _sweeping_gst(1024, 1024, 70000, seed=77)   # N = 65536 blocks, indices up to ~70000
    small = fx.random_gst(1024, 1024, seed=78, palette_entries=900)
    res = decoder.decode_tapped([small, big])
    _check_stages(res, 0, small)
    _check_stages(res, 1, big)
    assert int(res["indices"][65536:].max()) > 66000


def test_malformed_streams_do_not_fault(decoder):
    """The reference has no bounds checks in its kernels (a malformed stream is
    undefined behaviour).  Here garbage payloads, garbage group offsets and garbage frequency
    tables must come back as ordinary (meaningless) output without a CUDA fault, and the context
    must still decode a good stream afterwards."""
    good, golden = fx.golden_test1()
    rng = np.random.default_rng(5)
    bad = []
    a = good.copy(); a[28 + 2048:] = rng.integers(0, 256, size=a.size - 28 - 2048, dtype=np.uint8); bad.append(a)   # payload + offsets
    b = good.copy(); b[28:28 + 2048] = rng.integers(0, 256, size=2048, dtype=np.uint8); bad.append(b)              # frequency tables
    c = good.copy(); c[28 + 2048:28 + 2048 + 64] = 0xFF; bad.append(c)                                            # huge group offsets
    d = good.copy(); d[28 + 2048:28 + 2048 + 64] = 0x00; bad.append(d)                                            # zero group offsets
    for k, f in enumerate(bad):
        out = decoder.DecompressDXT(f)
        assert out.size == golden.size, f"case {k}"
    out = decoder.DecompressDXTs(bad + [good], page=5)
    assert np.array_equal(out[4 * golden.size:], golden), "a good image next to malformed ones must still decode"
    assert np.array_equal(decoder.DecompressDXT(good), golden)
