"""Test / benchmark infrastructure: fixture generation and the CPU checkers.

* `oracle()`  -> ctypes handle of oracle/libgst_oracle.so (the plain-C restatement)
* `ref()`     -> ctypes handle of oracle/_ref/libgst_ref.so (the reference's own CPU sources,
                 compiled in place by oracle/build_ref.sh) or None when it was never built
* `encode_image(w, h, seed)` -> (.gst bytes, golden DXT1 bytes) through the UNMODIFIED
  reference encoder (GenTC::CompressDXT), cached under tests/_cache/
* `make_gst(...)` -> a .gst container built from arbitrary symbol planes (reference rANS
  encoder per group), for edge cases a real image never produces

Nothing here is imported by the product package (gst_b200/).
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
CACHE_DIR = os.path.join(ROOT, "tests", "_cache")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

GROUP = 8192
_oracle = None
_ref = None
_ref_tried = False

_u8p = C.POINTER(C.c_uint8)


def _ptr(a, t=C.c_void_p):
    return a.ctypes.data_as(t) if a is not None else None


def oracle():
    """Build (gcc, a second) and load the plain-C oracle."""
    global _oracle
    if _oracle is not None:
        return _oracle
    so = os.path.join(ORACLE_DIR, "libgst_oracle.so")
    src = os.path.join(ORACLE_DIR, "gst_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c99", "-shared", "-o", so, src])
    L = C.CDLL(so)
    L.gsto_decode.restype = C.c_int
    L.gsto_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gsto_build_table.restype = None
    L.gsto_build_table.argtypes = [C.c_void_p] * 4
    L.gsto_ans_decode_group.restype = None
    L.gsto_ans_decode_group.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    L.gsto_ans_decode_stream.restype = None
    L.gsto_ans_decode_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.gsto_decode_indices.restype = None
    L.gsto_decode_indices.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    L.gsto_inverse_wavelet_tile.restype = None
    L.gsto_inverse_wavelet_tile.argtypes = [C.c_void_p, C.c_void_p]
    L.gsto_inverse_wavelet_plane.restype = None
    L.gsto_inverse_wavelet_plane.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    L.gsto_assemble_dxt.restype = C.c_int
    L.gsto_assemble_dxt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    L.gsto_assemble_rgb.restype = C.c_int
    L.gsto_assemble_rgb.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p]
    _oracle = L
    return L


def ref():
    """The reference-linked checker, or None if oracle/_ref/libgst_ref.so does not exist and
    cannot be built here (no /root/reference on the GPU box: the prebuilt file travels)."""
    global _ref, _ref_tried
    if _ref_tried:
        return _ref
    _ref_tried = True
    so = os.path.join(ORACLE_DIR, "_ref", "libgst_ref.so")
    if not os.path.exists(so) and os.path.isdir(os.environ.get("GST_REFERENCE", "/root/reference")):
        subprocess.call(["bash", os.path.join(ORACLE_DIR, "build_ref.sh")])
    if not os.path.exists(so):
        return None
    L = C.CDLL(so)
    L.gstref_encode_rgb.restype = C.c_int
    L.gstref_encode_rgb.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p]
    L.gstref_encode_file.restype = C.c_int
    L.gstref_encode_file.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_size_t,
                                     C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t]
    L.gstref_decode.restype = C.c_int
    L.gstref_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gstref_decode_batch.restype = C.c_int
    L.gstref_decode_batch.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_int, C.POINTER(C.c_void_p), C.c_int]
    L.gstref_generate_histogram.restype = C.c_int
    L.gstref_generate_histogram.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.gstref_encode_interleaved.restype = C.c_int
    L.gstref_encode_interleaved.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                            C.POINTER(C.c_size_t)]
    L.gstref_decode_interleaved.restype = C.c_int
    L.gstref_decode_interleaved.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    for name in ("gstref_inverse_wavelet2d", "gstref_forward_wavelet2d"):
        fn = getattr(L, name)
        fn.restype = None
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    for name in ("gstref_inverse_wavelet1d", "gstref_forward_wavelet1d"):
        fn = getattr(L, name)
        fn.restype = None
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    _ref = L
    return L


# ------------------------------------------------------------------------------------------
def synth_image(width, height, seed):
    """per channel a sum of 6 sinusoids (0.5-12 cycles per image,
    random phase / amplitude) normalised to 0..255, plus Gaussian noise sigma = 3."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:height, 0:width].astype(np.float32)
    img = np.empty((height, width, 3), dtype=np.uint8)
    for c in range(3):
        acc = np.zeros((height, width), dtype=np.float32)
        for _ in range(6):
            fx, fy = rng.uniform(0.5, 12.0, size=2)
            ph = rng.uniform(0, 2 * np.pi)
            amp = rng.uniform(0.3, 1.0)
            acc += amp * np.sin(2 * np.pi * (fx * x / width + fy * y / height) + ph).astype(np.float32)
        acc = (acc - acc.min()) / max(float(acc.max() - acc.min()), 1e-6) * 255.0
        acc += rng.normal(0.0, 3.0, size=acc.shape).astype(np.float32)
        img[:, :, c] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    return img


def encode_rgb(img):
    """RGB8 array -> (.gst bytes, encoder's PhysicalBlocks() bytes) via the reference encoder."""
    L = ref()
    if L is None:
        raise RuntimeError("oracle/_ref/libgst_ref.so is not available")
    h, w, _ = img.shape
    img = np.ascontiguousarray(img, dtype=np.uint8)
    cap = w * h * 2 + (1 << 20)
    gst = np.empty(cap, dtype=np.uint8)
    dxt = np.empty(w * h // 2, dtype=np.uint8)
    n = C.c_size_t()
    rc = L.gstref_encode_rgb(w, h, _ptr(img), _ptr(gst), cap, C.byref(n), _ptr(dxt))
    if rc != 0:
        raise RuntimeError(f"gstref_encode_rgb failed: {rc}")
    return gst[: n.value].copy(), dxt


class GoldenSha(str):
    """sha256 of a golden DXT1 image too large to keep on disk (see encode_image)."""


BIG_PIXELS = 1024 * 1024  # goldens above this are cached as a sha256, not as 0.5 B/texel of blocks


def matches_golden(out, golden):
    """out == the encoder's PhysicalBlocks(), given either the blocks or their GoldenSha."""
    if isinstance(golden, GoldenSha):
        return sha(out) == golden
    return bool(np.array_equal(out, golden))


def _tag(width, height, seed, noise_only=False, shift=None):
    if shift is not None:
        return f"motion_{width}x{height}_s{seed}_x{shift}"
    return f"{'noise' if noise_only else 'synth'}_{width}x{height}_s{seed}"


def encode_image(width, height, seed, noise_only=False, shift=None):
    """Seeded synthetic image through the reference encoder, cached on disk.  Returns
    (.gst bytes, golden) where golden is the encoder's PhysicalBlocks() as a uint8 array, or,
    for images above 1 Mpixel, its GoldenSha (keeps the cache that travels to the GPU box small).
    shift: the image translated horizontally by `shift` pixels (wrapping) -- a frame of the motion
    sequence of SURVEY.md section 8(d)."""
    os.makedirs(CACHE_DIR, exist_ok=True)
    tag = _tag(width, height, seed, noise_only, shift)
    pg, pd, ps = (os.path.join(CACHE_DIR, tag + ext) for ext in (".gst", ".dxt", ".sha"))
    big = width * height > BIG_PIXELS
    if os.path.exists(pg) and os.path.exists(ps if big else pd):
        gst = np.fromfile(pg, dtype=np.uint8)
        return gst, (GoldenSha(open(ps).read().strip()) if big else np.fromfile(pd, dtype=np.uint8))
    if noise_only:
        img = np.random.default_rng(seed).integers(0, 256, size=(height, width, 3), dtype=np.uint8)
    else:
        img = synth_image(width, height, seed)
        if shift:
            img = np.ascontiguousarray(np.roll(img, shift, axis=1))
    gst, dxt = encode_rgb(img)
    gst.tofile(pg + ".tmp")
    os.replace(pg + ".tmp", pg)
    if big:
        with open(ps + ".tmp", "w") as f:
            f.write(sha(dxt))
        os.replace(ps + ".tmp", ps)
        return gst, GoldenSha(sha(dxt))
    dxt.tofile(pd + ".tmp")
    os.replace(pd + ".tmp", pd)
    return gst, dxt


def encode_images(width, height, seeds, workers=None):
    """encode_image for many seeds, missing ones on parallel worker processes (the reference
    encoder is ~3 s per 2048x2048 image per core and narrates through a global std::cout, so
    the workers are separate processes, not threads)."""
    ext = ".sha" if width * height > BIG_PIXELS else ".dxt"
    missing = [s for s in seeds if not (os.path.exists(os.path.join(CACHE_DIR, f"synth_{width}x{height}_s{s}.gst")) and
                                        os.path.exists(os.path.join(CACHE_DIR, f"synth_{width}x{height}_s{s}{ext}")))]
    if len(missing) > 1:
        import sys
        workers = workers or min(len(missing), os.cpu_count() or 1)
        procs = []
        for k in range(workers):
            part = missing[k::workers]
            if part:
                procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--encode", str(width),
                                               str(height)] + [str(s) for s in part]))
        for pr in procs:
            if pr.wait() != 0:
                raise RuntimeError("fixture encoder worker failed")
    return [encode_image(width, height, s) for s in seeds]


def encode_motion(width, height, seed, n_frames, step=2, workers=None):
    """Frames 0 .. n_frames-1 of the motion sequence: the seeded base image translated by `step` pixels per frame
    (SURVEY.md section 8d), each through the reference encoder; missing ones on parallel worker processes."""
    ext = ".sha" if width * height > BIG_PIXELS else ".dxt"
    shifts = [k * step for k in range(n_frames)]
    missing = [x for x in shifts if not (os.path.exists(os.path.join(CACHE_DIR, _tag(width, height, seed, shift=x) + ".gst")) and
                                         os.path.exists(os.path.join(CACHE_DIR, _tag(width, height, seed, shift=x) + ext)))]
    if len(missing) > 1:
        import sys
        workers = workers or min(len(missing), os.cpu_count() or 1)
        procs = []
        for k in range(workers):
            part = missing[k::workers]
            if part:
                procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--encode-motion", str(width),
                                               str(height), str(seed)] + [str(x) for x in part]))
        for pr in procs:
            if pr.wait() != 0:
                raise RuntimeError("fixture encoder worker failed")
    return [encode_image(width, height, seed, shift=x) for x in shifts]


def golden_test1():
    """codec/test/test1.png encoded by the reference encoder: (.gst, PhysicalBlocks) from the
    committed fixture (tests/golden/make_golden.py wrote it)."""
    gst = np.fromfile(os.path.join(GOLDEN_DIR, "test1.gst"), dtype=np.uint8)
    dxt = np.fromfile(os.path.join(GOLDEN_DIR, "test1.dxt"), dtype=np.uint8)
    return gst, dxt


# ------------------------------------------------------------------------------------------
def header_of(gst):
    v = np.frombuffer(bytes(gst[:28]), dtype="<u4")
    return dict(zip(("width", "height", "palette_bytes", "y_cmp_sz", "chroma_cmp_sz", "palette_sz", "indices_sz"),
                    (int(x) for x in v)))


def oracle_decode(gst, mode=0, taps=True):
    """Plain-C oracle: dict(out, symbols, planes, indices)."""
    L = oracle()
    gst = np.ascontiguousarray(gst, dtype=np.uint8)
    h = header_of(gst)
    n = (h["width"] // 4) * (h["height"] // 4)
    out = np.empty(h["width"] * h["height"] * 3 if mode else 8 * n, dtype=np.uint8)
    sym = np.empty(7 * n + h["palette_bytes"], dtype=np.uint8) if taps else None
    planes = np.empty(6 * n, dtype=np.int8) if taps else None
    idx = np.empty(n, dtype=np.int32) if taps else None
    rc = L.gsto_decode(_ptr(gst), gst.size, mode, _ptr(out), _ptr(sym), _ptr(planes), _ptr(idx))
    if rc != 0:
        raise RuntimeError(f"gsto_decode failed: {rc}")
    return dict(out=out, symbols=sym, planes=planes, indices=idx, header=h)


def ref_decode(gst, taps=True):
    """Reference-linked stitched decoder (oracle/ref_glue.cpp): dict(out, symbols, planes, indices)."""
    L = ref()
    gst = np.ascontiguousarray(gst, dtype=np.uint8)
    h = header_of(gst)
    n = (h["width"] // 4) * (h["height"] // 4)
    out = np.empty(8 * n, dtype=np.uint8)
    sym = np.empty(7 * n + h["palette_bytes"], dtype=np.uint8) if taps else None
    planes = np.empty(6 * n, dtype=np.int8) if taps else None
    idx = np.empty(n, dtype=np.int32) if taps else None
    rc = L.gstref_decode(_ptr(gst), gst.size, _ptr(out), _ptr(sym), _ptr(planes), _ptr(idx))
    if rc != 0:
        raise RuntimeError(f"gstref_decode failed: {rc}")
    return dict(out=out, symbols=sym, planes=planes, indices=idx, header=h)


# ------------------------------------------------------------------------------------------
def ref_histogram(counts, M=2048):
    L = ref()
    c = np.ascontiguousarray(counts, dtype=np.uint32)
    out = np.zeros_like(c)
    rc = L.gstref_generate_histogram(_ptr(c), c.size, M, _ptr(out))
    if rc != 0:
        raise RuntimeError("gstref_generate_histogram failed")
    return out


def ref_encode_interleaved(symbols, F, num_streams):
    """ans::EncodeInterleaved with the OpenCL options; F must sum to 2048.
    Returns the encoded bytes: [renorm words][num_streams u32 states]."""
    L = ref()
    s = np.ascontiguousarray(symbols, dtype=np.uint8)
    f = np.ascontiguousarray(F, dtype=np.uint32)
    cap = 4 * s.size + 4 * num_streams + 64
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t()
    rc = L.gstref_encode_interleaved(_ptr(s), s.size, _ptr(f), f.size, num_streams, _ptr(out), cap, C.byref(n))
    if rc != 0:
        raise RuntimeError(f"gstref_encode_interleaved failed: {rc}")
    return out[: n.value].copy()


def encode_stream(symbols):
    """ByteEncoder::EncodeBytes (codec/entropy.cpp:174-265) rebuilt around the reference rANS
    encoder: returns (512-byte freq block, stream bytes = [u32 end offsets][groups])."""
    s = np.ascontiguousarray(symbols, dtype=np.uint8)
    assert s.size % GROUP == 0
    counts = np.bincount(s, minlength=256).astype(np.uint32)
    nz = int(np.max(np.nonzero(counts)[0])) + 1
    F = ref_histogram(counts[:nz])
    freqs = np.zeros(256, dtype="<u2")
    freqs[:nz] = F
    n_groups = s.size // GROUP
    chunks, offsets, cum = [], [], 4 * n_groups
    for g in range(n_groups):
        enc = ref_encode_interleaved(s[g * GROUP:(g + 1) * GROUP], F, 32)
        if enc.size & 3:
            enc = np.concatenate([np.zeros(2, np.uint8), enc])  # entropy.cpp:213-228
        cum += enc.size
        offsets.append(cum)
        chunks.append(enc)
    stream = np.concatenate([np.array(offsets, dtype="<u4").view(np.uint8)] + chunks)
    pad = (-stream.size) % 4
    if pad:
        stream = np.concatenate([stream, np.zeros(pad, np.uint8)])
    return freqs.view(np.uint8), stream


def make_gst(width, height, y_syms, chroma_syms, palette, index_syms):
    """Assemble a .gst container (codec/encoder.cpp:122-144) from raw symbol arrays:
    y_syms 2N bytes, chroma_syms 4N, palette P bytes (multiple of 8192), index_syms N."""
    n = (width // 4) * (height // 4)
    assert y_syms.size == 2 * n and chroma_syms.size == 4 * n and index_syms.size == n and palette.size % GROUP == 0
    parts = [encode_stream(a) for a in (y_syms, chroma_syms, palette, index_syms)]
    hdr = np.array([width, height, palette.size] + [p[1].size for p in parts], dtype="<u4").view(np.uint8)
    return np.concatenate([hdr] + [p[0] for p in parts] + [p[1] for p in parts])


def random_gst(width, height, seed, palette_entries=2048, plane_mode="laplace"):
    """A synthetic .gst whose symbol planes are random (not an image): exercises value ranges
    the reference encoder never emits.  Returns the container bytes."""
    rng = np.random.default_rng(seed)
    n = (width // 4) * (height // 4)
    if plane_mode == "uniform":
        planes = rng.integers(0, 256, size=6 * n, dtype=np.uint8)
    elif plane_mode == "extreme":
        planes = rng.choice(np.array([0, 255, 128, 1, 254], dtype=np.uint8), size=6 * n)
    else:
        planes = np.clip(np.rint(rng.laplace(0.0, 6.0, size=6 * n)) + 128, 0, 255).astype(np.uint8)
    pal_bytes = -(-palette_entries * 4 // GROUP) * GROUP
    palette = np.zeros(pal_bytes, dtype=np.uint8)
    palette[: palette_entries * 4] = rng.integers(0, 256, size=palette_entries * 4, dtype=np.uint8)
    # index deltas: a random walk kept inside [0, palette_entries)
    idx = np.empty(n, dtype=np.int64)
    cur = 0
    steps = rng.integers(-127, 128, size=n)
    for i in range(n):
        nxt = cur + int(steps[i])
        if nxt < 0 or nxt >= palette_entries:
            nxt = cur - int(steps[i])
            if nxt < 0 or nxt >= palette_entries:
                nxt = cur
        idx[i] = nxt
        cur = nxt
    deltas = np.diff(np.concatenate([[0], idx])) + 128
    assert deltas.min() >= 0 and deltas.max() <= 255
    return make_gst(width, height, planes[: 2 * n], planes[2 * n:], palette, deltas.astype(np.uint8))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 4 and sys.argv[1] == "--encode":
        for seed in sys.argv[4:]:
            encode_image(int(sys.argv[2]), int(sys.argv[3]), int(seed))
    if len(sys.argv) > 5 and sys.argv[1] == "--encode-motion":
        for x in sys.argv[5:]:
            encode_image(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), shift=int(x))
