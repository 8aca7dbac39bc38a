"""CPU tests of the C-ABI library: it loads, exports every declared symbol, refuses to run
without a GPU, and its host-only logic (container parsing, batch packing, scratch sizing,
frequency normalisation) matches the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gst_fixtures as fx

ROOT = fx.ROOT


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    import gst_b200
    from gst_b200 import capi
    handle = gst_b200.load_library()
    text = open(os.path.join(ROOT, "include", "gst_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(gst_[a-z0-9_]+)\s*\(", text))
    assert len(declared) >= 35
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in gst_cuda.h but not exported"
        assert name in capi.PROTOTYPES, f"{name} has no ctypes prototype in gst_b200/capi.py"
    assert set(capi.PROTOTYPES) <= declared


def test_no_cpu_fallback_without_device():
    import gst_b200
    if _have_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(gst_b200.GstError) as e:
        gst_b200.Decoder(0)
    assert e.value.code == -2  # GST_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under gst_b200/ or include/ may mention it."""
    for base in ("gst_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                    src = open(os.path.join(dirpath, f), errors="replace").read()
                    assert "oracle" not in src.lower() and "gst_fixtures" not in src, os.path.join(dirpath, f)


def test_parse_header_and_container_checks():
    import gst_b200
    gst, _ = fx.golden_test1()
    h = gst_b200.parse_header(gst)
    assert (h.width, h.height, h.palette_bytes) == (512, 512, 8192)
    assert 28 + 2048 + h.payload_bytes == gst.size  # codec/encoder.cpp:122-144
    assert gst_b200.GenTCHeader.LoadFrom(gst).y_cmp_sz == h.y_cmp_sz
    with pytest.raises(gst_b200.GstError):
        gst_b200.parse_header(gst[:1000])          # truncated
    bad = gst.copy()
    bad[:4] = np.array([500], dtype="<u4").view(np.uint8)  # width not a multiple of 128
    with pytest.raises(gst_b200.GstError):
        gst_b200.parse_header(bad)
    bad = gst.copy()
    bad[8:12] = np.array([8000], dtype="<u4").view(np.uint8)  # palette_bytes % 8192
    with pytest.raises(gst_b200.GstError):
        gst_b200.parse_header(bad)


def test_required_scratch_matches_reference_formula():
    """codec/decoder.cpp:41-47: 4*2048*6 + 17*W*H/16 + palette_bytes."""
    import gst_b200
    h = gst_b200.parse_header(fx.golden_test1()[0])
    assert gst_b200.required_scratch_mem(h) == 4 * 2048 * 6 + 17 * 512 * 512 // 16 + 8192


def test_pack_batch_layout_matches_photos_sf():
    """demo/photos_sf.cpp:753-795 / codec/decoder.cpp:430-476: [out_off 4n][in_off 4n] padded
    to 512 | n x 2048 freqs | payloads, offsets as running sums."""
    import gst_b200
    files = [fx.golden_test1()[0], np.fromfile(os.path.join(fx.GOLDEN_DIR, "synth512_s7.gst"), dtype=np.uint8)]
    files = files + [files[0]]
    packed, hdrs = gst_b200.pack_batch(files)
    n, N = 3, 16384
    off_region = 512
    assert packed.size == off_region + n * 2048 + sum(h.payload_bytes for h in hdrs)
    offs = packed[:8 * n * 4].view("<u4")
    out_off, in_off = offs[:4 * n], offs[4 * n:8 * n]
    io = oo = 0
    for i, h in enumerate(hdrs):
        for s, (isz, osz) in enumerate(zip((h.y_cmp_sz, h.chroma_cmp_sz, h.palette_sz, h.indices_sz),
                                           (2 * N, 4 * N, h.palette_bytes, N))):
            assert in_off[4 * i + s] == io and out_off[4 * i + s] == oo
            io += isz
            oo += osz
        assert np.array_equal(packed[off_region + 2048 * i: off_region + 2048 * (i + 1)], files[i][28:28 + 2048])
    payload = packed[off_region + 2048 * n:]
    pos = 0
    for f, h in zip(files, hdrs):
        assert np.array_equal(payload[pos:pos + h.payload_bytes], f[28 + 2048:28 + 2048 + h.payload_bytes])
        pos += h.payload_bytes
    # a single image reproduces UploadData: 8 offsets at byte 0, file minus header at byte 512
    one, _ = gst_b200.pack_batch(files[:1])
    assert np.array_equal(one[512:], files[0][28:])


def test_pack_batch_rejects_mixed_dimensions():
    import gst_b200
    a = fx.golden_test1()[0]
    b = a.copy()
    b[4:8] = np.array([1024], dtype="<u4").view(np.uint8)
    with pytest.raises(gst_b200.GstError):
        gst_b200.pack_batch([a, b])


def test_normalize_frequencies_golden_vectors():
    """ans/histogram_test.cpp:64-103."""
    import gst_b200
    nf = gst_b200.normalize_frequencies
    assert list(nf([1, 0, 2, 1], 256)) == [64, 0, 128, 64]
    assert list(nf([1, 1, 2], 256)) == [64, 64, 128]
    assert list(nf([1, 2, 3, 4, 5, 6, 7, 8, 9, 10], 256)) == [5, 9, 14, 19, 23, 28, 33, 37, 42, 46]
    assert list(nf([1, 2, 3, 4, 5, 6, 7, 8, 9, 10], 11)) == [1, 1, 1, 1, 1, 1, 1, 1, 1, 2]
    with pytest.raises(gst_b200.GstError):
        nf([0, 0, 0], 256)


def test_normalize_frequencies_matches_reference(ref_lib):
    import gst_b200
    rng = np.random.default_rng(11)
    for trial in range(50):
        n = int(rng.integers(1, 257))
        counts = rng.integers(0, 5000, size=n).astype(np.uint32)
        if trial % 3 == 0:
            counts[rng.integers(0, n, size=n // 2)] = 0
        if counts.sum() == 0:
            counts[0] = 1
        got = gst_b200.normalize_frequencies(counts)
        assert got.sum() == 2048
        assert np.array_equal(got, fx.ref_histogram(counts))
    # the byte histograms of real streams
    o = fx.oracle_decode(fx.golden_test1()[0])
    counts = np.bincount(o["symbols"][:32768], minlength=256).astype(np.uint32)
    assert np.array_equal(gst_b200.normalize_frequencies(counts), fx.ref_histogram(counts))


def test_launch_count_and_stream_size_checks_are_host_logic():
    """gst_launches_for_batch (2 launches for calls of at most 16384 rANS groups, else 3) and the per-stream minimum
    size every entry point now checks need no device."""
    import gst_b200
    from gst_b200.capi import gst_header, lib
    h = gst_b200.parse_header(fx.golden_test1()[0]).to_c()
    assert lib().gst_launches_for_batch((gst_header * 1)(h), 1) == 2
    assert lib().gst_launches_for_batch((gst_header * 128)(*[h] * 128), 128) == 2       # 128 x 15 groups
    assert lib().gst_launches_for_batch((gst_header * 512)(*[h] * 512), 512) == 2       # 7 680 groups
    assert lib().gst_launches_for_batch((gst_header * 2048)(*[h] * 2048), 2048) == 3
    bad = gst_header(width=512, height=512, palette_bytes=8192, y_cmp_sz=64, chroma_cmp_sz=64, palette_sz=64, indices_sz=64)
    assert lib().gst_launches_for_batch((gst_header * 1)(bad), 1) == 0                  # 4 groups cannot fit in 64 bytes
    assert lib().gst_packed_size((gst_header * 1)(bad), 1) == 0
    assert b"cannot hold" in lib().gst_last_error()
