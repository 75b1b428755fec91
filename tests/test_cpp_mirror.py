"""include/rocoder_b200.hpp: the C++ mirror of Stretcher / StretcherProcessor / AudioBus compiles
against the C ABI (CPU) and, on a GPU, reproduces the reference's own Stretcher tests and the
batched engine's output."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
import rocoder_b200 as rb

SRC = os.path.join(ol.ROOT, "tests", "cpp", "stretcher_mirror.cpp")


def _build(tmp_path):
    exe = str(tmp_path / "stretcher_mirror")
    libdir = os.path.dirname(rb.LIB_PATH)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ol.ROOT, "include"), SRC, "-o", exe,
                           "-L", libdir, "-lrocoder_b200", f"-Wl,-rpath,{libdir}", "-lpthread"])
    return exe


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu(tmp_path):
    exe = _build(tmp_path)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([exe, "ensure"], capture_output=True, text=True)
    assert r.returncode != 0  # rcg_create throws: no device, no fallback


@pytest.mark.gpu
def test_cpp_reference_stretcher_tests(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "ensure"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("f,p,W,n", [(2.0, 2, 1024, 30000), (3.0, -2, 512, 9000), (1.0, 1, 16384, 70000)])
def test_cpp_processor_equals_batched_engine(tmp_path, f, p, W, n):
    exe = _build(tmp_path)
    rng = np.random.default_rng(3)
    x = ((rng.random((2, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)
    x.tofile(tmp_path / "in.f32")
    r = subprocess.run([exe, "run", str(tmp_path / "in.f32"), "2", str(n), str(f), str(p), str(W), "21",
                        str(tmp_path / "out.f32")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=21) as e:
        want = e.process_all(x)
    got = np.fromfile(tmp_path / "out.f32", dtype=np.float32).reshape(2, -1)
    assert np.array_equal(got, want)
