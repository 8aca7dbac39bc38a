"""Host-side mirror of the reference decoder interface, on top of the C ABI.

Names follow the reference (codec/decoder.h:15-36, codec/codec_base.h:9-22, ans/ans.h:72-79,
ans/ans_ocl.h:26-72) so the parity tests read like the reference's own tests:

    reference (C++ / OpenCL)                         here
    ---------------------------------------------    ---------------------------------------
    gpu::GPUContext::InitializeOpenCL + InitializeDecoder   Decoder(device)
    GenTC::DecompressDXT(ctx, bytes)                 Decoder.DecompressDXT(bytes)
    GenTC::LoadCompressedDXT(s)(ctx, hdrs, q, cmp, out, n, ev)  Decoder.LoadCompressedDXTs(...)
    GenTC::LoadRGB(s)                                Decoder.LoadRGBs(...)
    GenTC::RequiredScratchMem(hdr)                   required_scratch_mem(hdr)
    GenTC::PreallocateDecompressor / FreeDecompressor  Decoder.PreallocateDecompressor / FreeDecompressor
    ans::ocl::OpenCLDecoder                          AnsDecoder
    ans::ocl::NormalizeFrequencies                   normalize_frequencies

Everything here is glue: the work is done by libgst_cuda.so.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import check, gst_header, lib

kANSTableSize = 1 << 11          # ans/ans.h:73
kNumEncodedSymbols = 256         # ans/ans.h:74
kThreadsPerEncodingGroup = 32    # ans/ans.h:75
kWaveletBlockDim = 32            # codec/codec_base.h:22
kHeaderBytes = 28


class GenTCHeader:
    """codec/codec_base.h:9-20."""
    FIELDS = ("width", "height", "palette_bytes", "y_cmp_sz", "chroma_cmp_sz", "palette_sz", "indices_sz")

    def __init__(self, **kw):
        for f in self.FIELDS:
            setattr(self, f, int(kw.get(f, 0)))

    @classmethod
    def LoadFrom(cls, buf):
        """codec/codec_base.cpp:18-24: raw little-endian copy of 7 u32."""
        v = np.frombuffer(bytes(buf[:kHeaderBytes]), dtype="<u4")
        if v.size != 7:
            raise capi.GstError(-1, "buffer too short for a GenTCHeader")
        return cls(**dict(zip(cls.FIELDS, (int(x) for x in v))))

    def to_c(self):
        return gst_header(*(getattr(self, f) for f in self.FIELDS))

    @classmethod
    def from_c(cls, h):
        return cls(**{f: getattr(h, f) for f in cls.FIELDS})

    @property
    def num_blocks(self):
        return (self.width // 4) * (self.height // 4)

    @property
    def dxt_bytes(self):
        return self.width * self.height // 2

    @property
    def rgb_bytes(self):
        return self.width * self.height * 3

    @property
    def payload_bytes(self):
        return self.y_cmp_sz + self.chroma_cmp_sz + self.palette_sz + self.indices_sz

    def __repr__(self):
        return "GenTCHeader(" + ", ".join(f"{f}={getattr(self, f)}" for f in self.FIELDS) + ")"


def _as_u8(buf):
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    return np.ascontiguousarray(a, dtype=np.uint8)


def _hdr_array(hdrs):
    arr = (gst_header * len(hdrs))()
    for i, h in enumerate(hdrs):
        arr[i] = h.to_c() if isinstance(h, GenTCHeader) else h
    return arr


def parse_header(gst_bytes):
    """GenTCHeader::LoadFrom + the container checks (gst_parse_header)."""
    a = _as_u8(gst_bytes)
    h = gst_header()
    check(lib().gst_parse_header(a.ctypes.data, a.size, C.byref(h)))
    return GenTCHeader.from_c(h)


def required_scratch_mem(hdr):
    """GenTC::RequiredScratchMem (codec/decoder.cpp:41-47)."""
    h = hdr.to_c()
    return int(lib().gst_required_scratch(C.byref(h)))


def pack_batch(files, out=None):
    """Build the device input buffer of LoadCompressedDXTs (demo/photos_sf.cpp:753-795).
    Returns (packed uint8 array, [GenTCHeader])."""
    arrs = [_as_u8(f) for f in files]
    n = len(arrs)
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    lens = (C.c_size_t * n)(*[a.size for a in arrs])
    hdrs = (gst_header * n)()
    for i, a in enumerate(arrs):
        check(lib().gst_parse_header(a.ctypes.data, a.size, C.byref(hdrs[i])))
    size = int(lib().gst_packed_size(hdrs, n))
    if size == 0:
        raise capi.GstError(-1, lib().gst_last_error().decode())
    if out is None:
        out = np.empty(size, dtype=np.uint8)
    check(lib().gst_pack_batch(ptrs, lens, n, out.ctypes.data, out.size, hdrs))
    return out[:size], [GenTCHeader.from_c(h) for h in hdrs]


def normalize_frequencies(counts, target_sum=kANSTableSize):
    """ans::ocl::NormalizeFrequencies (ans/ans_ocl_encode.cpp:6-8) =
    ans::GenerateHistogram(counts, 2048) (ans/histogram.cpp:41-123)."""
    c = np.ascontiguousarray(counts, dtype=np.uint32)
    out = np.zeros_like(c)
    check(lib().gst_normalize_frequencies(c.ctypes.data_as(C.POINTER(C.c_uint32)), c.size, int(target_sum),
                                          out.ctypes.data_as(C.POINTER(C.c_uint32))))
    return out


class DeviceBuffer:
    def __init__(self, dec, nbytes):
        self.dec, self.nbytes = dec, int(nbytes)
        p = C.c_void_p()
        check(lib().gst_malloc(dec.ctx, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def free(self):
        if self.ptr:
            lib().gst_free(self.dec.ctx, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy uint8 array (.array)."""

    def __init__(self, dec, nbytes):
        self.dec, self.nbytes = dec, int(nbytes)
        p = C.c_void_p()
        check(lib().gst_host_alloc(dec.ctx, self.nbytes, C.byref(p)))
        self.ptr = p.value
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr))[: self.nbytes]

    def free(self):
        if self.ptr:
            self.array = None
            lib().gst_host_free(self.dec.ctx, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Event:
    def __init__(self, handle):
        self.handle = handle

    def wait(self):
        check(lib().gst_event_wait(self.handle))

    def elapsed_ms(self, later):
        ms = C.c_float()
        check(lib().gst_event_elapsed_ms(self.handle, later.handle, C.byref(ms)))
        return ms.value

    def destroy(self):
        if self.handle:
            lib().gst_event_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Decoder:
    """One decoder context per GPU (replaces gpu::GPUContext + GenTC::InitializeDecoder)."""

    def __init__(self, device=0):
        self.ctx = None
        p = C.c_void_p()
        check(lib().gst_ctx_create(int(device), C.byref(p)))
        self.ctx = p.value
        self.device = int(device)

    def close(self):
        if self.ctx:
            lib().gst_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- streams / events / memory -------------------------------------------------------
    def GetDefaultCommandQueue(self):
        return lib().gst_stream_default(self.ctx)

    def GetNextQueue(self):
        return lib().gst_stream_next(self.ctx)

    def FlushAllQueues(self):
        check(lib().gst_ctx_sync(self.ctx))

    def sync(self, stream=None):
        if stream is None:
            check(lib().gst_ctx_sync(self.ctx))
        else:
            check(lib().gst_stream_sync(self.ctx, stream))

    def record(self, stream):
        ev = C.c_void_p()
        check(lib().gst_event_record(self.ctx, stream, C.byref(ev)))
        return Event(ev.value)

    def malloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    def pinned(self, nbytes):
        return PinnedBuffer(self, nbytes)

    def upload(self, dbuf, host, stream=None, offset=0):
        a = _as_u8(host) if not isinstance(host, PinnedBuffer) else host.array
        s = stream if stream is not None else self.GetDefaultCommandQueue()
        check(lib().gst_upload_async(self.ctx, s, dbuf.ptr + offset, a.ctypes.data, a.size))
        if stream is None:
            self.sync(s)

    def download(self, dbuf, nbytes=None, stream=None, out=None, offset=0):
        n = dbuf.nbytes - offset if nbytes is None else int(nbytes)
        if out is None:
            out = np.empty(n, dtype=np.uint8)
        s = stream if stream is not None else self.GetDefaultCommandQueue()
        check(lib().gst_download_async(self.ctx, s, out.ctypes.data, dbuf.ptr + offset, n))
        if stream is None:
            self.sync(s)
        return out

    def download_2d(self, pinned, dbuf, width, rows, src_pitch, dst_pitch, src_offset=0, dst_offset=0, stream=None):
        """rows x width bytes, strided on both sides, into a PinnedBuffer (blocking unless a stream is given)."""
        s = stream if stream is not None else self.GetDefaultCommandQueue()
        check(lib().gst_download_2d_async(self.ctx, s, pinned.ptr + dst_offset, dst_pitch, dbuf.ptr + src_offset,
                                          src_pitch, width, rows))
        if stream is None:
            self.sync(s)

    def memset(self, dbuf, value, stream=None):
        s = stream if stream is not None else self.GetDefaultCommandQueue()
        check(lib().gst_memset_async(self.ctx, s, dbuf.ptr, int(value), dbuf.nbytes))
        if stream is None:
            self.sync(s)

    # -- profiling -----------------------------------------------------------------------
    KERNEL_NAMES = ("build_tables", "rans_streams", "wavelet_assemble")

    def status_flags(self, clear=True):
        """gst_status_flags: bit 0 = a palette index was clamped, bit 1 = a palette region was out of range,
        bit 2 = the kernels' internal hand-over timed out (never expected)."""
        f = C.c_uint32()
        check(lib().gst_status_flags(self.ctx, C.byref(f), 1 if clear else 0))
        return f.value

    def profile(self, on=True):
        check(lib().gst_profile_enable(self.ctx, 1 if on else 0))

    def profile_read(self):
        """-> ({kernel name: total ms}, number of decode calls covered)."""
        n = lib().gst_launches_per_batch()
        ms = (C.c_double * n)()
        calls = C.c_uint64()
        check(lib().gst_profile_read(self.ctx, ms, n, C.byref(calls)))
        return dict(zip(self.KERNEL_NAMES, list(ms))), int(calls.value)

    # -- scratch -------------------------------------------------------------------------
    def PreallocateDecompressor(self, req_sz):
        check(lib().gst_preallocate(self.ctx, int(req_sz)))

    def FreeDecompressor(self):
        check(lib().gst_free_scratch(self.ctx))

    # -- decode --------------------------------------------------------------------------
    def _load(self, fn, hdrs, queue, cmp_data, output, init, want_event, cmp_bytes):
        harr = _hdr_array(hdrs)
        n_wait = len(init) if init else 0
        wait = (C.c_void_p * n_wait)(*[e.handle for e in init]) if n_wait else None
        done = C.c_void_p()
        cmp_ptr = cmp_data.ptr if isinstance(cmp_data, DeviceBuffer) else int(cmp_data)
        out_ptr = output.ptr if isinstance(output, DeviceBuffer) else int(output)
        nbytes = cmp_bytes if cmp_bytes is not None else cmp_data.nbytes
        check(fn(self.ctx, harr, len(hdrs), queue, cmp_ptr, nbytes, out_ptr, wait, n_wait,
                 C.byref(done) if want_event else None))
        return Event(done.value) if want_event else None

    def LoadCompressedDXTs(self, hdrs, queue, cmp_data, output, init=(), want_event=True, cmp_bytes=None):
        """codec/decoder.h:23-25.  Returns the completion Event (caller owns it)."""
        return self._load(lib().gst_load_dxt_batch, hdrs, queue, cmp_data, output, init, want_event, cmp_bytes)

    def LoadCompressedDXT(self, hdr, queue, cmp_data, output, init=(), want_event=True, cmp_bytes=None):
        """codec/decoder.h:19-21."""
        return self.LoadCompressedDXTs([hdr], queue, cmp_data, output, init, want_event, cmp_bytes)

    def LoadRGBs(self, hdrs, queue, cmp_data, output, init=(), want_event=True, cmp_bytes=None):
        """codec/decoder.h:31-33."""
        return self._load(lib().gst_load_rgb_batch, hdrs, queue, cmp_data, output, init, want_event, cmp_bytes)

    def LoadRGB(self, hdr, queue, cmp_data, output, init=(), want_event=True, cmp_bytes=None):
        """codec/decoder.h:27-29."""
        return self.LoadRGBs([hdr], queue, cmp_data, output, init, want_event, cmp_bytes)

    def DecompressDXT(self, cmp_data, mode=0):
        """codec/decoder.h:16-17: host .gst bytes -> host DXT1 blocks (uint8 array of W*H/2)."""
        a = _as_u8(cmp_data)
        h = parse_header(a)
        out = np.empty(h.rgb_bytes if mode else h.dxt_bytes, dtype=np.uint8)
        check(lib().gst_decompress_host(self.ctx, a.ctypes.data, a.size, int(mode), out.ctypes.data, out.size))
        return out

    def DecompressDXTs(self, files, page=16, mode=0, out=None):
        """Batched host-to-host decode (pages pipelined over the work streams)."""
        arrs = [_as_u8(f) for f in files]
        n = len(arrs)
        h = parse_header(arrs[0])
        per = h.rgb_bytes if mode else h.dxt_bytes
        if out is None:
            out = np.empty(per * n, dtype=np.uint8)
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        lens = (C.c_size_t * n)(*[a.size for a in arrs])
        check(lib().gst_decompress_host_batch(self.ctx, ptrs, lens, n, int(page), int(mode), out.ctypes.data, out.size))
        return out

    def LoadHostBatch(self, files, output, page=16, mode=0):
        """The headless photos_sf loader (demo/photos_sf.cpp:688-885): host .gst buffers (numpy
        arrays or PinnedBuffers) -> textures in the caller's DeviceBuffer `output`."""
        n = len(files)
        ptrs = (C.c_void_p * n)(*[f.ptr if isinstance(f, PinnedBuffer) else _as_u8(f).ctypes.data for f in files])
        lens = (C.c_size_t * n)(*[f.nbytes if isinstance(f, PinnedBuffer) else _as_u8(f).size for f in files])
        check(lib().gst_load_host_batch(self.ctx, ptrs, lens, n, int(page), int(mode), output.ptr, output.nbytes))

    def decode_tapped(self, files):
        """Decode a batch and also return the stage intermediates (parity tests):
        dict(dxt, symbols, planes, indices, hdrs)."""
        packed, hdrs = pack_batch(files)
        n, N = len(hdrs), hdrs[0].num_blocks
        sym_bytes = sum(7 * N + h.palette_bytes for h in hdrs)
        d_cmp, d_out = self.malloc(packed.size), self.malloc(8 * N * n)
        d_sym, d_pl, d_idx = self.malloc(sym_bytes), self.malloc(6 * N * n), self.malloc(4 * N * n)
        s = self.GetDefaultCommandQueue()
        self.upload(d_cmp, packed)
        for b in (d_out, d_sym, d_pl, d_idx):
            self.memset(b, 0xCD)
        check(lib().gst_load_dxt_batch_tapped(self.ctx, _hdr_array(hdrs), n, s, d_cmp.ptr, d_cmp.nbytes, d_out.ptr,
                                              d_sym.ptr, d_pl.ptr, d_idx.ptr))
        self.sync(s)
        res = dict(
            hdrs=hdrs,
            dxt=self.download(d_out, 8 * N * n),
            symbols=self.download(d_sym, sym_bytes),
            planes=self.download(d_pl, 6 * N * n).view(np.int8),
            indices=self.download(d_idx, 4 * N * n).view(np.int32),
        )
        for b in (d_cmp, d_out, d_sym, d_pl, d_idx):
            b.free()
        return res

    def build_tables(self, freqs_u16):
        """Stage 1 on n x 256 u16 frequencies -> (symbols, freqs, cum_freqs) arrays [n, 2048]."""
        f = np.ascontiguousarray(freqs_u16, dtype=np.uint16).reshape(-1, 256)
        n = f.shape[0]
        d_f, d_t = self.malloc(f.nbytes), self.malloc(n * kANSTableSize * 4)
        self.upload(d_f, f.view(np.uint8).reshape(-1))
        s = self.GetDefaultCommandQueue()
        check(lib().gst_build_tables(self.ctx, s, d_f.ptr, n, d_t.ptr))
        t = self.download(d_t, n * kANSTableSize * 4).view(np.uint32).reshape(n, kANSTableSize)
        d_f.free()
        d_t.free()
        # packed entry (csrc/gst_kernels.cuh): freq | sym << 12 | (slot - cum) << 20
        slot = np.arange(kANSTableSize, dtype=np.int64)[None, :]
        freq = (t & 0xFFF).astype(np.uint16)
        cum = slot - (t >> 20).astype(np.int64)
        return ((t >> 12) & 0xFF).astype(np.uint8), freq, cum.astype(np.uint16)


class FrameStreamer:
    """gst_streamer_*: the headless demo player (demo/demo.cpp:145-243) with `depth` frames in flight."""

    def __init__(self, dec, width, height, depth=4, mode=0):
        self.dec, self.frame_bytes = dec, (width * height * 3 if mode else width * height // 2)
        h = C.c_void_p()
        check(lib().gst_streamer_create(dec.ctx, width, height, depth, int(mode), C.byref(h)))
        self.handle = h

    def submit(self, frame, out=None, host_out=None, direct=False):
        """frame: .gst bytes (numpy array or PinnedBuffer); out: optional DeviceBuffer; host_out: optional host
        address the decoded frame is also copied to (on the slot's stream, covered by wait()); direct: upload
        straight from `frame`, which must then stay valid until wait().  Returns the ticket."""
        ptr, n = (frame.ptr, frame.nbytes) if isinstance(frame, PinnedBuffer) else (_as_u8(frame).ctypes.data, _as_u8(frame).size)
        t = C.c_uint64()
        if host_out is None and not direct:
            check(lib().gst_streamer_submit(self.handle, ptr, n, out.ptr if out is not None else None, C.byref(t)))
        else:
            check(lib().gst_streamer_submit_ex(self.handle, ptr, n, out.ptr if out is not None else None, host_out,
                                               1 if direct else 0, C.byref(t)))
        return t.value

    def play(self, frame_ptrs, frame_lens, n, host_out=None, dev_out=None, direct=False, group=0):
        """gst_streamer_play: the whole frame loop in one call (ctypes arrays of addresses and sizes); `group`
        frames per decode call (0 = the library's default, 8)."""
        flags = (1 if direct else 0) | ((int(group) & 0xFF) << 8)
        check(lib().gst_streamer_play(self.handle, frame_ptrs, frame_lens, n, dev_out, host_out, flags))

    def wait(self, ticket):
        """Blocks until the frame is decoded; returns its device address."""
        p = C.c_void_p()
        check(lib().gst_streamer_wait(self.handle, ticket, C.byref(p)))
        return p.value

    def read(self, ticket):
        """wait() + download of the frame (test helper)."""
        ptr = self.wait(ticket)
        out = np.empty(self.frame_bytes, dtype=np.uint8)
        s = self.dec.GetDefaultCommandQueue()
        check(lib().gst_download_async(self.dec.ctx, s, out.ctypes.data, ptr, out.size))
        self.dec.sync(s)
        return out

    def close(self):
        if self.handle:
            lib().gst_streamer_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def encode_stream(dec, symbols):
    """ByteEncoder::EncodeBytes (codec/entropy.cpp:174-265) on the GPU: byte symbols (a multiple of
    8192 of them) -> (512-byte frequency block, stream bytes = [u32 end offsets][rANS groups]).
    Fixture tooling: it entropy-codes test inputs without the CPU encoder."""
    s = _as_u8(symbols)
    freqs = np.empty(256, dtype="<u2")
    cap = lib().gst_ans_encode_bound(s.size)
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t()
    check(lib().gst_ans_encode_stream(dec.ctx, s.ctypes.data, s.size, freqs.ctypes.data, out.ctypes.data, cap, C.byref(n)))
    return freqs.view(np.uint8).copy(), out[: n.value].copy()


def build_gst(dec, width, height, y_syms, chroma_syms, palette, index_syms):
    """A .gst container (codec/encoder.cpp:122-144) from raw symbol arrays, entropy-coded on the GPU:
    y_syms 2N bytes (Y1 || Y2), chroma_syms 4N (Co1 || Cg1 || Co2 || Cg2), palette P bytes (a multiple of
    8192), index_syms N, with N = (width / 4) * (height / 4)."""
    n = (width // 4) * (height // 4)
    arrs = [_as_u8(a) for a in (y_syms, chroma_syms, palette, index_syms)]
    if [a.size for a in arrs[:2]] + [arrs[3].size] != [2 * n, 4 * n, n] or arrs[2].size % 8192:
        raise ValueError("symbol arrays do not match the image size")
    parts = [encode_stream(dec, a) for a in arrs]
    hdr = np.array([width, height, arrs[2].size] + [p[1].size for p in parts], dtype="<u4").view(np.uint8)
    return np.concatenate([hdr] + [p[0] for p in parts] + [p[1] for p in parts])


class AnsDecoder:
    """ans::ocl::OpenCLDecoder (ans/ans_ocl.h:26-72)."""

    def __init__(self, dec, F, num_interleaved):
        self.dec = dec
        self.handle = None
        f = np.ascontiguousarray(F, dtype=np.uint32)
        p = C.c_void_p()
        check(lib().gst_ans_create(dec.ctx, f.ctypes.data_as(C.POINTER(C.c_uint32)), f.size, int(num_interleaved),
                                   C.byref(p)))
        self.handle = p.value
        self.num_interleaved = int(num_interleaved)

    def RebuildTable(self, F):
        f = np.ascontiguousarray(F, dtype=np.uint32)
        check(lib().gst_ans_rebuild(self.handle, f.ctypes.data_as(C.POINTER(C.c_uint32)), f.size))

    def _table(self):
        sym = np.empty(kANSTableSize, np.uint8)
        fr = np.empty(kANSTableSize, np.uint16)
        cum = np.empty(kANSTableSize, np.uint16)
        check(lib().gst_ans_table(self.handle, sym.ctypes.data, fr.ctypes.data, cum.ctypes.data))
        return sym, fr, cum

    def GetSymbols(self):
        return self._table()[0]

    def GetFrequencies(self):
        return self._table()[1]

    def GetCumulativeFrequencies(self):
        return self._table()[2]

    def Decode(self, states, data):
        """The three overloads of ans/ans_ocl.cpp:155-345.
        Decode(state:int, bytes)                 -> 256 symbols
        Decode([states], bytes)                  -> [len(states)][256] (one interleaved group)
        Decode([states], [bytes per group])      -> [len(states)][256] (groups of num_interleaved)"""
        single = np.isscalar(states)
        st = np.ascontiguousarray([states] if single else states, dtype=np.uint32)
        if isinstance(data, (bytes, bytearray, np.ndarray)):
            groups, lanes = [_as_u8(data)], st.size
        else:
            groups, lanes = [_as_u8(d) for d in data], self.num_interleaved
            assert st.size == lanes * len(groups)
        g = len(groups)
        ptrs = (C.c_void_p * g)(*[a.ctypes.data if a.size else None for a in groups])
        lens = (C.c_size_t * g)(*[a.size for a in groups])
        out = np.empty(st.size * kNumEncodedSymbols, dtype=np.uint8)
        check(lib().gst_ans_decode(self.handle, lanes, st.ctypes.data_as(C.POINTER(C.c_uint32)), ptrs, lens, g,
                                   out.ctypes.data))
        return out if single else out.reshape(st.size, kNumEncodedSymbols)

    def close(self):
        if self.handle:
            lib().gst_ans_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
