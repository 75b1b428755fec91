"""The C-ABI library loads without a GPU, exports every symbol include/rocoder_b200.h declares,
its host arithmetic is bit-exact with the oracle, and GPU entry points fail loudly (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as ol
import rocoder_b200 as rb
from rocoder_b200 import _abi


def _declared_symbols():
    text = open(os.path.join(ol.ROOT, "include", "rocoder_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rcg_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    names = _declared_symbols()
    assert len(names) >= 25
    L = C.CDLL(rb.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.rcg_abi_version() == 2


def test_binding_covers_the_header():
    L = _abi.lib()
    for n in _declared_symbols():
        assert getattr(L, n).argtypes is not None, n


@pytest.mark.parametrize("f", [0.5, 0.75, 1.0, 1.3, 1.7, 2.0, 3.0, 4.0, 8.0, 13.37, 20.0, 100.0])
@pytest.mark.parametrize("p", [1, 2, 3, 5, 7, -2, -3, -4, -7])
@pytest.mark.parametrize("W", [64, 1024, 16384, 65536])
def test_plan_bit_exact_with_oracle(f, p, W):
    want = ol.constants(f, p, W, amplitude=0.7)
    ok = 1 <= want["sample_step_len"] <= W
    if not ok:
        with pytest.raises(rb.RcgError) as ei:
            rb.plan(f, p, W, amplitude=0.7)
        assert ei.value.status == _abi.RCG_ERR_STEP
        return
    got = rb.plan(f, p, W, amplitude=0.7)
    for key in want:
        assert got[key] == want[key], key
    for n in (0, 1, W - 1, W, W + 1, 10 * W + 17, 1234567):
        assert rb.expected_counts(n, f, p, W) == ol.expected_counts(n, f, p, W)


def test_invalid_arguments_return_error_codes_not_panics():
    for p, status in ((0, _abi.RCG_ERR_PITCH), (-1, _abi.RCG_ERR_PITCH), (128, _abi.RCG_ERR_PITCH), (-128, _abi.RCG_ERR_PITCH)):
        with pytest.raises(rb.RcgError) as ei:
            rb.plan(1.0, p, 1024)
        assert ei.value.status == status
    for W in (1001, 10001, 63, 32, 3, 0, 2 ** 20 + 2, 2 ** 21):   # odd, too short, too long
        with pytest.raises(rb.RcgError) as ei:
            rb.plan(1.0, 1, W)
        assert ei.value.status == _abi.RCG_ERR_WINDOW
    with pytest.raises(rb.RcgError) as ei:  # step > W: `-f 1 -p -4` underflows usize in the reference
        rb.plan(1.0, -4, 1024)
    assert ei.value.status == _abi.RCG_ERR_STEP
    assert b"step" in _abi.lib().rcg_status_string(_abi.RCG_ERR_STEP)


def test_tables_bit_exact_with_oracle():
    for n in (2, 32, 1000, 16384):
        assert np.array_equal(rb.hanning(n), ol.hanning(n))
    for n in (16, 8192):
        assert np.array_equal(rb.hanning_crossfade_compensation(n), ol.crossfade(n))
    env = rb.hanning_crossfade_compensation(8192)
    assert 0.42 < env.min() < 0.4205 and 0.5795 < env.max() < 0.58  # 0.4204..0.5796


def test_channel_bound():
    """stretcher.rs:82-85"""
    assert rb.channel_bound(16384, 44100, 1.0) == 1
    assert rb.channel_bound(16384, 44100, 0.1) == 4
    assert rb.channel_bound(65536, 48000, 0.5) == 3
    s = ol.OracleStretcher(1.0, 1, 16384, sample_rate=44100, buffer_dur=0.1)
    assert s.channel_bound() == 4


def test_engine_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rb.RcgError) as ei:
        rb.Engine(2, 48000)
    assert ei.value.status in (_abi.RCG_ERR_NO_DEVICE, _abi.RCG_ERR_CUDA)
    with pytest.raises(rb.RcgError):
        rb.stretch(np.zeros((1, 100), np.float32), 48000)


def test_product_does_not_reference_the_oracle():
    """the product path may not import, link or execute anything under oracle/ or tests/"""
    pkg = os.path.join(ol.ROOT, "rocoder_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".hpp")) and f != "build.py":
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in text and "librocoder_oracle" not in text and "np_oracle" not in text, f
                assert "rcg_emul" not in text or f == "rcg_frame.h", f
    import subprocess
    out = subprocess.run(["ldd", rb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "emul" not in out


def test_header_is_plain_c(tmp_path):
    """the boundary is a C ABI: the header must compile as C99 and as C++"""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "rocoder_b200.h"\nint main(void) { rcg_config c; c.struct_size = sizeof c; return rcg_abi_version() == RCG_ABI_VERSION ? 0 : (int)c.struct_size; }\n')
    inc = os.path.join(ol.ROOT, "include")
    libdir = os.path.dirname(rb.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", inc, str(src), "-o", str(tmp_path / "t"),
                           "-L", libdir, "-lrocoder_b200", f"-Wl,-rpath,{libdir}"])
    assert subprocess.run([str(tmp_path / "t")]).returncode == 0
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-I", inc, "-x", "c++", os.path.join(inc, "rocoder_b200.hpp")])
