"""ctypes binding of include/gst_cuda.h (one prototype per exported symbol)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# GST_LIB: another build of the same library (A/B timing of kernel variants, scripts/ab.sh)
LIB_PATH = os.environ.get("GST_LIB") or os.path.join(HERE, "lib", "libgst_cuda.so")


class GstError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"gst error {code}: {message}")
        self.code = code


class gst_header(C.Structure):
    """codec/codec_base.h:9-20 GenTCHeader."""
    _fields_ = [(n, C.c_uint32) for n in (
        "width", "height", "palette_bytes", "y_cmp_sz", "chroma_cmp_sz", "palette_sz", "indices_sz")]


_vp, _u32, _sz, _int = C.c_void_p, C.c_uint32, C.c_size_t, C.c_int
_pp = C.POINTER(C.c_void_p)
_hdr_p = C.POINTER(gst_header)

# name -> (restype, argtypes); must list every symbol include/gst_cuda.h declares
PROTOTYPES = {
    "gst_last_error": (C.c_char_p, []),
    "gst_ctx_create": (_int, [_int, _pp]),
    "gst_ctx_destroy": (None, [_vp]),
    "gst_ctx_device": (_int, [_vp]),
    "gst_stream_default": (_vp, [_vp]),
    "gst_stream_next": (_vp, [_vp]),
    "gst_ctx_sync": (_int, [_vp]),
    "gst_stream_sync": (_int, [_vp, _vp]),
    "gst_malloc": (_int, [_vp, _sz, _pp]),
    "gst_free": (_int, [_vp, _vp]),
    "gst_host_alloc": (_int, [_vp, _sz, _pp]),
    "gst_host_free": (_int, [_vp, _vp]),
    "gst_upload_async": (_int, [_vp, _vp, _vp, _vp, _sz]),
    "gst_download_async": (_int, [_vp, _vp, _vp, _vp, _sz]),
    "gst_download_2d_async": (_int, [_vp, _vp, _vp, _sz, _vp, _sz, _sz, _sz]),
    "gst_memset_async": (_int, [_vp, _vp, _vp, _int, _sz]),
    "gst_event_record": (_int, [_vp, _vp, _pp]),
    "gst_event_wait": (_int, [_vp]),
    "gst_event_elapsed_ms": (_int, [_vp, _vp, C.POINTER(C.c_float)]),
    "gst_event_destroy": (None, [_vp]),
    "gst_parse_header": (_int, [_vp, _sz, _hdr_p]),
    "gst_packed_size": (_sz, [_hdr_p, _u32]),
    "gst_pack_batch": (_int, [_pp, C.POINTER(_sz), _u32, _vp, _sz, _hdr_p]),
    "gst_required_scratch": (_sz, [_hdr_p]),
    "gst_preallocate": (_int, [_vp, _sz]),
    "gst_free_scratch": (_int, [_vp]),
    "gst_load_dxt_batch": (_int, [_vp, _hdr_p, _u32, _vp, _vp, _sz, _vp, _pp, _u32, _pp]),
    "gst_load_rgb_batch": (_int, [_vp, _hdr_p, _u32, _vp, _vp, _sz, _vp, _pp, _u32, _pp]),
    "gst_decompress_host": (_int, [_vp, _vp, _sz, _int, _vp, _sz]),
    "gst_decompress_host_batch": (_int, [_vp, _pp, C.POINTER(_sz), _u32, _u32, _int, _vp, _sz]),
    "gst_load_host_batch": (_int, [_vp, _pp, C.POINTER(_sz), _u32, _u32, _int, _vp, _sz]),
    "gst_streamer_create": (_int, [_vp, _u32, _u32, _u32, _int, C.POINTER(_vp)]),
    "gst_streamer_submit": (_int, [_vp, _vp, _sz, _vp, C.POINTER(C.c_uint64)]),
    "gst_streamer_submit_ex": (_int, [_vp, _vp, _sz, _vp, _vp, _u32, C.POINTER(C.c_uint64)]),
    "gst_streamer_play": (_int, [_vp, _pp, C.POINTER(_sz), _u32, _vp, _vp, _u32]),
    "gst_streamer_wait": (_int, [_vp, C.c_uint64, C.POINTER(_vp)]),
    "gst_streamer_destroy": (None, [_vp]),
    "gst_load_dxt_batch_tapped": (_int, [_vp, _hdr_p, _u32, _vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "gst_normalize_frequencies": (_int, [C.POINTER(_u32), _u32, _u32, C.POINTER(_u32)]),
    "gst_ans_create": (_int, [_vp, C.POINTER(_u32), _u32, _u32, _pp]),
    "gst_ans_rebuild": (_int, [_vp, C.POINTER(_u32), _u32]),
    "gst_ans_table": (_int, [_vp, _vp, _vp, _vp]),
    "gst_ans_decode": (_int, [_vp, _u32, C.POINTER(_u32), _pp, C.POINTER(_sz), _u32, _vp]),
    "gst_ans_destroy": (None, [_vp]),
    "gst_ans_encode_bound": (_sz, [_sz]),
    "gst_ans_encode_stream": (_int, [_vp, _vp, _sz, _vp, _vp, _sz, C.POINTER(_sz)]),
    "gst_build_tables": (_int, [_vp, _vp, _vp, _u32, _vp]),
    "gst_ctx_set_direct_upload": (_int, [_vp, _int]),
    "gst_status_flags": (_int, [_vp, C.POINTER(_u32), _int]),
    "gst_launches_per_batch": (_int, []),
    "gst_launches_for_batch": (_int, [_hdr_p, _u32]),
    "gst_profile_enable": (_int, [_vp, _int]),
    "gst_profile_read": (_int, [_vp, C.POINTER(C.c_double), _u32, C.POINTER(C.c_uint64)]),
}

_lib = None


def load_library(path=None):
    """dlopen libgst_cuda.so and attach the prototypes.  Raises if the library is missing:
    there is no fallback implementation."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise GstError(-2, f"{p} not found: build it with `python -m gst_b200.build` "
                           "(the decode path has no CPU fallback)")
    handle = C.CDLL(p)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(handle, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = handle
    return handle


def lib():
    return load_library()


def check(rc):
    if rc != 0:
        raise GstError(rc, lib().gst_last_error().decode("utf-8", "replace"))
