"""gst_b200 -- a B200-native (sm_100a) decoder for GammaUNC/GST `.gst` streams.

The product is the C-ABI shared library `gst_b200/lib/libgst_cuda.so` (declared in
`include/gst_cuda.h`, C++ facade in `include/gst_decoder.hpp`).  This package is the thin
Python binding used by the tests and the benchmark; it never falls back to a CPU path: if
the library is missing or no sm_100 device is present, calls raise.
"""
from .capi import GstError, lib, load_library  # noqa: F401
from .decoder import (  # noqa: F401
    AnsDecoder,
    Decoder,
    FrameStreamer,
    GenTCHeader,
    build_gst,
    encode_stream,
    kANSTableSize,
    kNumEncodedSymbols,
    kThreadsPerEncodingGroup,
    kWaveletBlockDim,
    normalize_frequencies,
    pack_batch,
    parse_header,
    required_scratch_mem,
)

__all__ = [
    "AnsDecoder", "Decoder", "FrameStreamer", "build_gst", "encode_stream", "GenTCHeader", "GstError", "lib", "load_library",
    "normalize_frequencies", "pack_batch", "parse_header", "required_scratch_mem",
    "kANSTableSize", "kNumEncodedSymbols", "kThreadsPerEncodingGroup", "kWaveletBlockDim",
]
