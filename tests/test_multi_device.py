"""Several GPUs behind ONE engine (rcg_config.n_devices; reference shape: one call per file for all channels,
main.rs:133-156).  The C++ test uses the plain C ABI (no torch); the Python tests cover what rides on top:
streaming routed by channel, frequency kernels on every shard, the host-`apply` rule."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
import plugins
import rocoder_b200 as rb
from rocoder_b200 import _abi

SRC = os.path.join(ol.ROOT, "tests", "cpp", "multi_device_test.cpp")


def _build(tmp_path):
    exe = str(tmp_path / "multi_device_test")
    libdir = os.path.dirname(rb.LIB_PATH)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ol.ROOT, "include"), SRC, "-o", exe,
                           "-L", libdir, "-lrocoder_b200", f"-Wl,-rpath,{libdir}"])
    return exe


def test_multi_device_test_compiles_and_config_is_validated(tmp_path):
    """no GPU here: the C++ test builds against the header, reports 'no device' (exit 77), and rcg_create checks the
    device list before it touches CUDA"""
    exe = _build(tmp_path)
    import torch
    if not torch.cuda.is_available():
        assert subprocess.run([exe], capture_output=True).returncode == 77
    cfg = _abi.make_config(2, 48000, 1024, devices=[0, 0])
    cfg.n_devices = _abi.RCG_MAX_DEVICES + 1
    h = C.c_void_p()
    assert _abi.lib().rcg_create(C.byref(cfg), C.byref(h)) == _abi.RCG_ERR_INVALID_ARG
    # an ABI-1 caller (struct ends before n_devices) is still served: it gets past the size check
    cfg = _abi.make_config(2, 48000, 1024)
    cfg.struct_size = _abi.rcg_config.n_devices.offset
    rc = _abi.lib().rcg_create(C.byref(cfg), C.byref(h))
    assert rc in (_abi.RCG_OK, _abi.RCG_ERR_NO_DEVICE, _abi.RCG_ERR_CUDA)
    if rc == _abi.RCG_OK:
        _abi.lib().rcg_destroy(h)
    cfg.struct_size = 12
    assert _abi.lib().rcg_create(C.byref(cfg), C.byref(h)) == _abi.RCG_ERR_INVALID_ARG


def _input(n, channels=2, seed=1):
    rng = np.random.default_rng(seed)
    return ((rng.random((channels, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_cpp_one_vs_many_devices_bit_identical(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    print(r.stdout[-4000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 failure(s)" in r.stdout


@pytest.mark.gpu
def test_group_streaming_plugins_and_host_rule():
    n, f, p, W = 200000, 2.0, 1, 2048
    x = _input(n, channels=4, seed=5)
    devs = [0, 0, 0] if rb.device_count() < 3 else [0, 1, 2]
    flags = rb.RCG_FLAG_SHARD_ALWAYS
    with rb.Engine(4, 48000, W, f, 1.0, p, seed=9) as e:
        whole = e.process_all(x)
    # streaming: channel c is served by device c mod 3
    with rb.Engine(4, 48000, W, f, 1.0, p, seed=9, devices=devs, flags=flags, max_lookahead_calls=4) as g:
        for c in range(4):
            g.send(c, x[c])
            g.close_input(c)
        for c in range(4):
            wins = []
            while not g.is_done(c):
                wins.append(g.next_window(c))
            assert np.array_equal(np.concatenate(wins), whole[c])
    # device frequency kernels (gather and fused forms) are installed on every shard
    for push in ("push_device_frequency_kernel", "push_local_device_frequency_kernel"):
        src = plugins.DEVICE_TILT if push == "push_device_frequency_kernel" else plugins.LOCAL_TILT
        with rb.Engine(4, 48000, W, f, 1.0, p, seed=9) as e:
            getattr(e, push)(src)
            one = e.process_all(x)
        with rb.Engine(4, 48000, W, f, 1.0, p, seed=9, devices=devs, flags=flags) as g:
            getattr(g, push)(src)
            many = g.process_all(x)
            part = g.last_partition()
        assert part["devices_used"] == 3 and part["split"] == "time" and part["exchanges"] == 2
        assert np.array_equal(one, many) and not np.array_equal(one, whole)
    # a host `apply` on top of the stack is called from one thread in the reference's order: one device
    calls = []

    def gain(ms, spec):
        calls.append(1)
        return spec * np.complex64(0.5)
    with rb.Engine(4, 48000, W, f, 1.0, p, seed=9, devices=devs, flags=flags) as g:
        g.push_frequency_kernel(gain)
        got = g.process_all(x[:, :40000])
        assert g.last_partition()["devices_used"] == 1
    ref = ol.process_all(x[:, :40000], f, p, W, seed=9, precision=64, plugin=plugins.half_gain)
    assert len(calls) == ol.expected_counts(40000, f, p, W)[1] * 4
    assert np.abs(got - ref).max() <= 1e-4


@pytest.mark.gpu
def test_group_kernel_stacks_stay_equal_after_a_failing_host_kernel():
    """a host kernel that panics is dropped by the one shard that ran it (host kernels keep a job on the first
    device); the other shards must drop it too, or the next multi-device job would mix two different stacks"""
    n, f, p, W = 200000, 2.0, 1, 2048
    x = _input(n, channels=2, seed=8)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=9) as e:
        e.push_device_frequency_kernel(plugins.DEVICE_TILT)
        want = e.process_all(x)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=9, devices=[0, 0, 0], flags=rb.RCG_FLAG_SHARD_ALWAYS) as g:
        g.push_device_frequency_kernel(plugins.DEVICE_TILT)
        calls = []

        def bad(ms, spec):
            calls.append(1)
            raise RuntimeError("panic")
        g.push_frequency_kernel(bad)
        first = g.process_all(x)                       # host kernel on top: one device; it fails on its first frame
        assert g.last_partition()["devices_used"] == 1 and len(calls) == 1
        assert g.frequency_kernel_count() == 1
        second = g.process_all(x)                      # now the device kernel is on top everywhere: all three devices
        assert g.last_partition()["devices_used"] == 3 and len(calls) == 1
    assert np.array_equal(first, want) and np.array_equal(second, want)
