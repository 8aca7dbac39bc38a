"""The C++ facade (include/gst_decoder.hpp) compiles against the C ABI, and -- on the GPU box --
its port of codec/test/codec_test.cpp passes."""
import os
import subprocess

import pytest

import gst_fixtures as fx

SRC = os.path.join(fx.ROOT, "tests", "cpp", "codec_test.cpp")
EXE = os.path.join(fx.ROOT, "tests", "cpp", "codec_test")
LIB_DIR = os.path.join(fx.ROOT, "gst_b200", "lib")


def _build():
    import gst_b200
    gst_b200.load_library()
    newest = max(os.path.getmtime(p) for p in (SRC, os.path.join(fx.ROOT, "include", "gst_decoder.hpp"),
                                              os.path.join(fx.ROOT, "include", "gst_cuda.h")))
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < newest:
        subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(fx.ROOT, "include"), SRC, "-o", EXE,
                               "-L", LIB_DIR, "-lgst_cuda", "-Wl,-rpath," + LIB_DIR])
    return EXE


def test_cpp_facade_builds_and_fails_loudly_without_gpu():
    exe = _build()
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    r = subprocess.run([exe, os.path.join(fx.GOLDEN_DIR, "test1.gst"), os.path.join(fx.GOLDEN_DIR, "test1.dxt")],
                       capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


def test_facade_keeps_the_reference_signatures():
    """tests/cpp/signature_check.cpp: every codec/decoder.h declaration, verbatim, binds to the facade."""
    src = os.path.join(fx.ROOT, "tests", "cpp", "signature_check.cpp")
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-fsyntax-only", "-I", os.path.join(fx.ROOT, "include"), src])


@pytest.mark.gpu
def test_cpp_codec_test():
    exe = _build()
    r = subprocess.run([exe, os.path.join(fx.GOLDEN_DIR, "test1.gst"), os.path.join(fx.GOLDEN_DIR, "test1.dxt")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout
