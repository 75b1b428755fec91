"""The translation unit NVRTC builds when a bin-local frequency kernel is fused into the stretch kernel
(rocoder_b200/csrc/rcg_jit_entry.h) compiles for sm_100a from the embedded header text alone -- no CUDA toolkit
headers, no source tree -- for every window size the engine dispatches.  NVRTC needs no GPU."""
import ctypes as C
import os
import re

import pytest

import oracle_lib as ol

CSRC = os.path.join(ol.ROOT, "rocoder_b200", "csrc")
USER = ("__device__ float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins, float2 x) {\n"
        "    return make_float2(x.x * 2.0f, x.y * (bin & 1 ? 1.0f : 0.5f));\n}\n")


def _nvrtc():
    for name in ("libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    pytest.skip("libnvrtc not available")


def _embedded_headers():
    """parse rcg_jit_sources.inc back into {name: text}: what the library hands to nvrtcCreateProgram"""
    from rocoder_b200 import build
    inc = open(build.generate_jit_sources()).read()
    names = re.search(r"kJitHeaderNames\[\] = \{(.*?)\};", inc).group(1)
    names = [n.strip().strip('"') for n in names.split(",")]
    body = inc[inc.index("kJitHeaderSources[] = {"):]
    texts, cur = [], []
    for line in body.split("\n    ,\n")[:-1]:
        texts.append("".join(re.findall(r'R"RCGJIT\((.*?)\)RCGJIT"', line, flags=re.S)))
    assert len(texts) == len(names)
    return dict(zip(names, texts))


def test_embedded_headers_equal_the_sources():
    for name, text in _embedded_headers().items():
        assert text == open(os.path.join(CSRC, name)).read(), name


@pytest.mark.parametrize("log2h,mode", [(5, 0), (8, 2), (11, 5), (13, 0), (13, 2), (13, 4), (14, 3), (15, 0), (15, 2)])
def test_fused_translation_unit_compiles_with_nvrtc(log2h, mode):
    """every window size, generic and with the launch configuration folded in (RCG_JIT_MODE: pitch 1 / other /
    -2 / -3, what rcg_engine.cu's fused_kernel_mode passes for odd and negative pitch multiples)"""
    nv = _nvrtc()
    hdr = _embedded_headers()
    names = (C.c_char_p * len(hdr))(*[n.encode() for n in hdr])
    srcs = (C.c_char_p * len(hdr))(*[t.encode() for t in hdr.values()])
    src = (USER + f"\n#define RCG_JIT_LOG2H {log2h}\n#define RCG_JIT_MODE {mode}\n#include \"rcg_jit_entry.h\"\n").encode()
    prog = C.c_void_p()
    assert nv.nvrtcCreateProgram(C.byref(prog), src, b"fused_frequency_kernel.cu", len(hdr), srcs, names) == 0
    opts = (C.c_char_p * 3)(b"--gpu-architecture=sm_100a", b"--std=c++17", b"-default-device")
    rc = nv.nvrtcCompileProgram(prog, 3, opts)
    n = C.c_size_t()
    nv.nvrtcGetProgramLogSize(prog, C.byref(n))
    log = C.create_string_buffer(max(n.value, 1))
    nv.nvrtcGetProgramLog(prog, log)
    assert rc == 0, log.value.decode(errors="replace")[:2000]
    nv.nvrtcGetCUBINSize(prog, C.byref(n))
    assert n.value > 10000
    cubin = C.create_string_buffer(n.value)
    nv.nvrtcGetCUBIN(prog, cubin)
    assert b"rcg_jit_stretch" in cubin.raw
    nv.nvrtcDestroyProgram(C.byref(prog))


def test_user_compile_error_is_reported():
    nv = _nvrtc()
    hdr = _embedded_headers()
    names = (C.c_char_p * len(hdr))(*[n.encode() for n in hdr])
    srcs = (C.c_char_p * len(hdr))(*[t.encode() for t in hdr.values()])
    src = b"float2 apply(int oops) { return nothing; }\n#define RCG_JIT_LOG2H 9\n#include \"rcg_jit_entry.h\"\n"
    prog = C.c_void_p()
    assert nv.nvrtcCreateProgram(C.byref(prog), src, b"fused_frequency_kernel.cu", len(hdr), srcs, names) == 0
    opts = (C.c_char_p * 3)(b"--gpu-architecture=sm_100a", b"--std=c++17", b"-default-device")
    assert nv.nvrtcCompileProgram(prog, 3, opts) != 0
    nv.nvrtcDestroyProgram(C.byref(prog))
