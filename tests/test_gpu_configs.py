"""BASELINE.json configs[3] and configs[4] at their STATED sizes, through the C ABI, on the GPU.

configs[3]: one LoadCompressedDXTs call over 1024 x 2048x2048 textures built from 64 DISTINCT containers
(random symbol planes / palettes / index walks, entropy-coded by the GPU rANS encoder -- SURVEY.md 8f row 4 -- so the
set takes seconds, not the 3 s per image of the CPU encoder); every one of the 1024 outputs against the CPU oracle.

configs[4]: 600 frames of the 2 px/frame translated motion sequence through gst_streamer_*; the
sequence is 64 consecutive reference-encoded frames played forwards and backwards; every distinct frame against the
encoder's PhysicalBlocks(), 32 of them against the CPU oracle as well, and all 600 against a second decode path
(the batched host-to-host decode)."""
import ctypes as C

import numpy as np
import pytest

import gst_b200
import gst_fixtures as fx
from gst_b200.capi import check, lib

pytestmark = pytest.mark.gpu


def _random_container(dec, w, h, seed):
    n = (w // 4) * (h // 4)
    rng = np.random.default_rng(77000 + seed)
    scale = [1.5, 3.0, 6.0, 12.0][seed % 4]
    planes = np.clip(np.rint(rng.laplace(0.0, scale, size=6 * n)) + 128, 0, 255).astype(np.uint8)
    entries = int(rng.integers(1500, 9000))
    pal_bytes = -(-4 * entries // 8192) * 8192
    palette = np.zeros(pal_bytes, dtype=np.uint8)
    palette[: 4 * entries] = rng.integers(0, 256, size=4 * entries, dtype=np.uint8)
    walk = np.cumsum(rng.integers(-100, 101, size=n))
    period = 2 * (entries - 1)
    m = np.mod(walk, period)
    idx = np.where(m < entries, m, period - m)
    deltas = np.clip(np.diff(np.concatenate([[0], idx])), -128, 127)
    return gst_b200.build_gst(dec, w, h, planes[: 2 * n], planes[2 * n:], palette, (deltas + 128).astype(np.uint8))


def test_config3_1024_textures_64_distinct_every_image_against_the_oracle(decoder):
    w = h = 2048
    per = w * h // 2
    distinct = [_random_container(decoder, w, h, s) for s in range(64)]
    assert len({f.tobytes() for f in distinct}) == 64
    want = [fx.oracle_decode(f, taps=False)["out"] for f in distinct]
    order = [(37 * i + i // 64) % 64 for i in range(1024)]
    assert set(order) == set(range(64))
    packed, hdrs = gst_b200.pack_batch([distinct[j] for j in order])
    d_cmp, d_out = decoder.malloc(packed.size), decoder.malloc(per * 1024)
    decoder.upload(d_cmp, packed)
    decoder.memset(d_out, 0xEE)
    q = decoder.GetDefaultCommandQueue()
    decoder.LoadCompressedDXTs(hdrs, q, d_cmp, d_out, want_event=False)     # ONE call, 1024 images
    decoder.sync(q)
    got = decoder.download(d_out).reshape(1024, per)
    for pos, j in enumerate(order):
        assert np.array_equal(got[pos], want[j]), f"image {pos} (container {j}) differs from the CPU oracle"
    d_cmp.free()
    d_out.free()


def test_config4_600_frame_motion_sequence_streamed(decoder):
    w, h, n_frames, n_distinct, depth = 1920, 1024, 600, 64, 4
    per = w * h // 2
    frames = fx.encode_motion(w, h, 40000, n_distinct)       # base image translated by 2 px per frame
    assert len({g.tobytes() for g, _ in frames}) == n_distinct
    period = 2 * n_distinct - 2
    seq = [(f % period) if (f % period) < n_distinct else period - (f % period) for f in range(n_frames)]
    pins = []
    for g, _ in frames:
        pb = decoder.pinned(g.size)
        pb.array[:] = g
        pins.append(pb)
    out = decoder.pinned(per * n_frames)
    out.array[:] = 0xEE
    ptrs = (C.c_void_p * n_frames)(*[pins[j].ptr for j in seq])
    lens = (C.c_size_t * n_frames)(*[pins[j].nbytes for j in seq])
    st = gst_b200.FrameStreamer(decoder, w, h, depth=depth)
    try:
        st.play(ptrs, lens, n_frames, host_out=out.ptr, direct=True)     # gst_streamer_play: upload, decode, read back
    finally:
        st.close()
    got = out.array.reshape(n_frames, per)
    # every distinct frame against the encoder's PhysicalBlocks(); 32 of them against the CPU oracle as well
    first = {j: seq.index(j) for j in range(n_distinct)}
    for j, (g, golden) in enumerate(frames):
        assert fx.matches_golden(got[first[j]], golden), f"frame {first[j]} (source {j}) differs from PhysicalBlocks()"
        if j % 2 == 0:
            assert np.array_equal(got[first[j]], fx.oracle_decode(g, taps=False)["out"]), f"source {j} differs from the oracle"
    # all 600 against a second decode path: the batched host-to-host decode of the 64 sources
    second = decoder.DecompressDXTs([g for g, _ in frames], page=16).reshape(n_distinct, per)
    for f, j in enumerate(seq):
        assert np.array_equal(got[f], second[j]), f"frame {f} differs from the batched decode of its source"


def test_streamer_submit_ex_direct_and_host_output(decoder):
    """gst_streamer_submit_ex: upload straight from the caller's pinned buffer, read-back on the slot's stream; the
    slot is reused (depth 2, 7 frames) while earlier read-backs may still be in flight."""
    srcs = [fx.encode_image(512, 512, 10000 + i) for i in range(3)]
    per = 512 * 512 // 2
    pins = []
    for g, _ in srcs:
        pb = decoder.pinned(g.size)
        pb.array[:] = g
        pins.append(pb)
    out = decoder.pinned(per * 7)
    out.array[:] = 0
    st = gst_b200.FrameStreamer(decoder, 512, 512, depth=2)
    try:
        tickets = [st.submit(pins[f % 3], host_out=out.ptr + f * per, direct=(f % 2 == 0)) for f in range(7)]
        for t in tickets[-2:]:
            st.wait(t)
    finally:
        st.close()
    for f in range(7):
        assert np.array_equal(out.array[f * per:(f + 1) * per], srcs[f % 3][1]), f"frame {f}"


def test_host_batch_rejects_pages_of_other_dimensions(decoder):
    """the output stride comes from the first file; a later page of larger images must be refused, not
    decoded past the end of the staging buffers."""
    small = fx.encode_image(512, 512, 10000)[0]
    big = fx.encode_image(2048, 2048, 30000)[0]
    with pytest.raises(gst_b200.GstError) as e:
        decoder.DecompressDXTs([small, small, big, big], page=2)
    assert e.value.code == -1 and "dimensions" in str(e.value)
    d_out = decoder.malloc(4 * 512 * 512 // 2)
    with pytest.raises(gst_b200.GstError):
        decoder.LoadHostBatch([small, small, big, big], d_out, page=2)
    d_out.free()
    # the context is still usable
    assert np.array_equal(decoder.DecompressDXT(small), fx.encode_image(512, 512, 10000)[1])


def test_concurrent_host_batches_share_the_slot_pool(decoder):
    """Two threads call the host batch loader at once (the reference's pool threads do, demo/photos_sf.cpp:747-830):
    both results are right, and together they take less than twice one call."""
    import threading
    srcs = [fx.encode_image(512, 512, 10000 + i) for i in range(8)]
    files = [srcs[i % 8][0] for i in range(64)]
    want = np.concatenate([srcs[i % 8][1] for i in range(64)])
    outs = [None, None]

    def run(k):
        outs[k] = decoder.DecompressDXTs(files, page=8)

    ts = [threading.Thread(target=run, args=(k,)) for k in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert np.array_equal(outs[0], want) and np.array_equal(outs[1], want)


def test_launch_count_small_and_large_calls(decoder):
    small = gst_b200.parse_header(fx.golden_test1()[0])
    big = gst_b200.parse_header(fx.encode_image(2048, 2048, 30000)[0])
    one = (gst_b200.capi.gst_header * 1)(small.to_c())
    many = (gst_b200.capi.gst_header * 128)(*[big.to_c()] * 128)
    assert lib().gst_launches_for_batch(one, 1) == 2      # tables built by the consuming CTAs
    assert lib().gst_launches_for_batch(many, 32) == 2    # 32 x 2048^2: still under 16 384 groups
    assert lib().gst_launches_for_batch(many, 128) == 3
    assert lib().gst_launches_per_batch() == 3


def test_status_flag_reports_clamped_palette_indices(decoder):
    """A stream whose index deltas walk past the end of its palette (the reference would read out of bounds): the
    decode is memory-safe, and gst_status_flags says that an index was clamped; a well-formed stream sets nothing."""
    w = h = 512
    n = (w // 4) * (h // 4)
    rng = np.random.default_rng(5)
    planes = np.clip(np.rint(rng.laplace(0.0, 4.0, size=6 * n)) + 128, 0, 255).astype(np.uint8)
    palette = rng.integers(0, 256, size=8192, dtype=np.uint8)          # 2048 entries
    good = np.full(n, 128, dtype=np.uint8)                              # every delta 0: index 0 everywhere
    bad = np.full(n, 128 + 100, dtype=np.uint8)                         # +100 per block: far beyond 2048 entries
    decoder.status_flags(clear=True)
    out = decoder.DecompressDXT(gst_b200.build_gst(decoder, w, h, planes[: 2 * n], planes[2 * n:], palette, good))
    assert decoder.status_flags() == 0
    assert np.array_equal(out.view("<u4")[1::2], np.full(n, palette.view("<u4")[0]))
    out = decoder.DecompressDXT(gst_b200.build_gst(decoder, w, h, planes[: 2 * n], planes[2 * n:], palette, bad))
    assert decoder.status_flags(clear=True) & 1
    assert decoder.status_flags() == 0
    # block 30 has index 31 * 100 = 3100 > 2047: it carries the last palette entry
    assert out.view("<u4")[1::2][30] == palette.view("<u4")[-1]
    assert out.view("<u4")[1::2][10] == palette.view("<u4")[1100]


def test_host_batch_direct_upload_of_pinned_files(decoder):
    """gst_ctx_set_direct_upload: pinned files are DMA-ed from where they lie, pageable ones staged, in one page."""
    srcs = [fx.encode_image(512, 512, 10000 + i) for i in range(6)]
    per = 512 * 512 // 2
    files = []
    for i, (g, _) in enumerate(srcs):
        if i % 2 == 0:
            pb = decoder.pinned(g.size)
            pb.array[:] = g
            files.append(pb)
        else:
            files.append(g)
    d_out = decoder.malloc(per * 6)
    check(lib().gst_ctx_set_direct_upload(decoder.ctx, 1))
    try:
        for page in (6, 4, 1):
            decoder.memset(d_out, 0xEE)
            decoder.LoadHostBatch(files, d_out, page=page)
            got = decoder.download(d_out).reshape(6, per)
            for i, (_, golden) in enumerate(srcs):
                assert np.array_equal(got[i], golden), f"page {page}, image {i}"
    finally:
        check(lib().gst_ctx_set_direct_upload(decoder.ctx, 0))
        d_out.free()


@pytest.mark.parametrize("group,direct", [(1, True), (3, False), (4, True)])
def test_streamer_play_groups(decoder, group, direct):
    """gst_streamer_play with `group` frames per decode call: a ragged last group, staged and direct uploads, the
    frames left in a caller's device buffer and in host memory."""
    srcs = [fx.encode_image(512, 512, 10000 + i) for i in range(5)]
    per = 512 * 512 // 2
    n = 11
    pins = []
    for g, _ in srcs:
        pb = decoder.pinned(g.size)
        pb.array[:] = g
        pins.append(pb)
    ptrs = (C.c_void_p * n)(*[pins[f % 5].ptr for f in range(n)])
    lens = (C.c_size_t * n)(*[pins[f % 5].nbytes for f in range(n)])
    host = decoder.pinned(per * n)
    host.array[:] = 0
    d_out = decoder.malloc(per * n)
    decoder.memset(d_out, 0xEE)
    st = gst_b200.FrameStreamer(decoder, 512, 512, depth=2)
    try:
        st.play(ptrs, lens, n, host_out=host.ptr, dev_out=d_out.ptr, direct=direct, group=group)
        # single submissions still work on the same streamer afterwards
        t = st.submit(srcs[0][0])
        assert np.array_equal(st.read(t), srcs[0][1])
    finally:
        st.close()
    dev = decoder.download(d_out)
    for f in range(n):
        assert np.array_equal(host.array[f * per:(f + 1) * per], srcs[f % 5][1]), f"host frame {f}"
        assert np.array_equal(dev[f * per:(f + 1) * per], srcs[f % 5][1]), f"device frame {f}"
    d_out.free()


def test_back_to_back_calls_reuse_the_hand_over_counters_and_scratch(monkeypatch):
    """The image-granular hand-over between the entropy-decode kernel and the tile kernel takes its counters from a
    zeroed per-stream pool that is zeroed again, in stream order, when it is used up, and the tile warps read scratch
    the previous call of the stream used for OTHER data.  Two different batches of small images (sixteen images' index
    totals share a cache line) alternate on one queue without any host synchronisation in between, on a context whose
    pool lasts four calls; every output is checked and no flag may be raised."""
    monkeypatch.setenv("GST_SYNC_POOL_WORDS", "256")
    decoder = gst_b200.Decoder(0)
    w = h = 512
    per = w * h // 2
    sets = []
    for s in range(2):
        files = [_random_container(decoder, w, h, 100 * s + i) for i in range(40)]
        want = np.stack([fx.oracle_decode(f, taps=False)["out"] for f in files])
        packed, hdrs = gst_b200.pack_batch(files)
        d_cmp = decoder.malloc(packed.size)
        decoder.upload(d_cmp, packed)
        sets.append((hdrs, d_cmp, want))
    rounds = 12
    outs = [decoder.malloc(per * 40) for _ in range(2 * rounds)]
    for o in outs:
        decoder.memset(o, 0xEE)
    decoder.status_flags(clear=True)
    q = decoder.GetDefaultCommandQueue()
    for r in range(2 * rounds):
        hdrs, d_cmp, _ = sets[r % 2]
        decoder.LoadCompressedDXTs(hdrs, q, d_cmp, outs[r], want_event=False)
    decoder.sync(q)
    assert decoder.status_flags() == 0
    for r in range(2 * rounds):
        got = decoder.download(outs[r]).reshape(40, per)
        assert np.array_equal(got, sets[r % 2][2]), f"call {r} differs from the CPU oracle"
    for o in outs:
        o.free()
    for _, d_cmp, _ in sets:
        d_cmp.free()
    decoder.close()


def test_back_to_back_large_calls_three_kernel_path(decoder):
    """The same for calls large enough to take the three-kernel path (separate table build, both later kernels launched
    as programmatic dependents): 300 x 1024x1024 per call, two different sets alternating on one queue."""
    w = h = 1024
    per = w * h // 2
    n = 300
    sets = []
    for s in range(2):
        distinct = [_random_container(decoder, w, h, 500 + 10 * s + i) for i in range(8)]
        want = [fx.oracle_decode(f, taps=False)["out"] for f in distinct]
        order = [(3 * i + s) % 8 for i in range(n)]
        packed, hdrs = gst_b200.pack_batch([distinct[j] for j in order])
        d_cmp = decoder.malloc(packed.size)
        decoder.upload(d_cmp, packed)
        sets.append((hdrs, d_cmp, want, order))
    outs = [decoder.malloc(per * n) for _ in range(4)]
    for o in outs:
        decoder.memset(o, 0xEE)
    decoder.status_flags(clear=True)
    q = decoder.GetDefaultCommandQueue()
    for r in range(4):
        hdrs, d_cmp, _, _ = sets[r % 2]
        decoder.LoadCompressedDXTs(hdrs, q, d_cmp, outs[r], want_event=False)
    decoder.sync(q)
    assert decoder.status_flags() == 0
    for r in range(4):
        got = decoder.download(outs[r]).reshape(n, per)
        _, _, want, order = sets[r % 2]
        for pos, j in enumerate(order):
            assert np.array_equal(got[pos], want[j]), f"call {r}, image {pos} differs from the CPU oracle"
    for o in outs:
        o.free()
    for _, d_cmp, _, _ in sets:
        d_cmp.free()
