"""hot swap of frequency kernels from a watched source file (reference
hotswapper.rs:14-86, fft.rs:76-108).  The watcher itself is host C++ and is tested on CPU; the swap
into a running engine is tested on GPU against the oracle's plugin call site."""
import os
import subprocess
import time

import numpy as np
import pytest

import oracle_lib as ol
import plugins


def test_watcher_build_swap_fallback_on_cpu(tmp_path):
    exe = str(tmp_path / "hotswap_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", os.path.join(ol.ROOT, "tests", "cpp", "hotswap_test.cpp"),
                           "-o", exe, "-ldl"])
    env = dict(os.environ, TMPDIR=str(tmp_path))
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert not [f for f in os.listdir(tmp_path) if f.startswith("rcg_kernel_")]   # no build products left behind


def _input(n, seed=1):
    rng = np.random.default_rng(seed)
    return ((rng.random((2, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)


def _host_src(gain):
    return ("#include <stddef.h>\nint apply(size_t ms, const float* in, size_t n, float* out) {\n"
            f"  for (size_t i = 0; i < 2 * n; i++) out[i] = in[i] * {gain:.1f}f;\n  return 0;\n}}\n")


def _dev_src(gain):
    return ("__device__ float2 apply(size_t ms, unsigned bin, unsigned n, const float2* s) {\n"
            f"  return make_float2(s[bin].x * {gain:.1f}f, s[bin].y * {gain:.1f}f);\n}}\n")


def _rewrite(path, text):
    tmp = str(path) + ".tmp"
    open(tmp, "w").write(text)
    os.replace(tmp, path)


def _wait(pred, timeout=60.0):
    t0 = time.time()
    while time.time() - t0 < timeout:
        if pred():
            return True
        time.sleep(0.02)
    return pred()


def _gain_plugin(g):
    def plug(ms, pin, n, pout):
        a, b = plugins._views(pin, n, pout)
        b[:] = a * np.float32(g)
        return 0
    return plug


def _local_src(gain):
    return ("__device__ float2 apply(size_t ms, unsigned bin, unsigned n, float2 x) {\n"
            f"  return make_float2(x.x * {gain:.1f}f, x.y * {gain:.1f}f);\n}}\n")


@pytest.mark.gpu
@pytest.mark.parametrize("suffix,src", [(".c", _host_src), (".cu", _dev_src), (".local.cu", _local_src)],
                         ids=["host", "device", "device-local"])
@pytest.mark.parametrize("p", [1, 2])
def test_gpu_hot_swap_from_watched_file(tmp_path, suffix, src, p):
    import rocoder_b200 as rb
    x = _input(9000)
    f, W = 2.0, 1024
    path = tmp_path / ("kernel" + suffix)
    path.write_text(src(2.0))
    refs = {g: ol.process_all(x, f, p, W, seed=4, precision=64, plugin=_gain_plugin(g)) for g in (2.0, 3.0)}

    def close(a, b):
        return np.abs(a - b).max() <= 1e-4 and ol.snr_db(b, a) > 100.0

    with rb.Engine(2, 48000, W, f, 1.0, p, seed=4) as e:
        e.watch_frequency_kernel(str(path), poll_ms=20)
        assert e.frequency_kernel_count() == 1                    # first build is synchronous (hotswapper.rs:17)
        assert e.watch_stats() == {"builds_ok": 1, "builds_failed": 0, "swaps": 1}
        assert close(e.process_all(x), refs[2.0])
        # modified -> rebuilt in the background -> newest kernel from the next batch on (fft.rs:78-81)
        _rewrite(path, src(3.0))
        assert _wait(lambda: e.watch_stats()["builds_ok"] == 2), e.watch_log()
        assert e.frequency_kernel_count() == 1                    # not yet taken
        assert close(e.process_all(x), refs[3.0])
        assert e.frequency_kernel_count() == 2                    # libraries accumulate (fft.rs:22)
        # a build that fails changes nothing (hotswapper.rs:36-42)
        _rewrite(path, "this does not compile\n")
        assert _wait(lambda: e.watch_stats()["builds_failed"] == 1)
        assert e.watch_log() != ""
        assert e.poll_frequency_kernels() == 0
        assert close(e.process_all(x), refs[3.0])
        assert e.watch_stats()["swaps"] == 2
        with pytest.raises(rb.RcgError):
            e.watch_frequency_kernel(str(path))                   # one watched file per engine


@pytest.mark.gpu
def test_gpu_watch_starts_without_a_working_kernel(tmp_path):
    """Before the first successful compile frames pass through unmodified (fft.rs:82-85)."""
    import rocoder_b200 as rb
    x = _input(9000)
    path = tmp_path / "late.cpp"
    path.write_text("int apply(;\n")
    with rb.Engine(2, 48000, 1024, 2.0, 1.0, 2, seed=4) as plain:
        base = plain.process_all(x)
    with rb.Engine(2, 48000, 1024, 2.0, 1.0, 2, seed=4) as e:
        e.watch_frequency_kernel(str(path), poll_ms=20)
        assert e.frequency_kernel_count() == 0 and e.watch_stats()["builds_failed"] == 1
        got = e.process_all(x)
        assert np.abs(got - base).max() < 2e-6                    # full-rate path vs half-rate path rounding
        _rewrite(path, '#include <stddef.h>\nextern "C" ' + _host_src(2.0).split("\n", 1)[1])
        assert _wait(lambda: e.watch_stats()["builds_ok"] == 1), e.watch_log()
        ref = ol.process_all(x, 2.0, 2, 1024, seed=4, precision=64, plugin=_gain_plugin(2.0))
        got = e.process_all(x)
        assert np.abs(got - ref).max() <= 1e-4 and ol.snr_db(ref, got) > 100.0


@pytest.mark.gpu
def test_gpu_watch_rejects_unknown_file_type_and_late_start(tmp_path):
    import rocoder_b200 as rb
    x = _input(9000)
    (tmp_path / "k.txt").write_text("fn apply() {}\n")
    (tmp_path / "k.c").write_text(_host_src(2.0))
    with rb.Engine(2, 48000, 1024, 2.0, 1.0, 2, seed=4) as e:
        with pytest.raises(rb.RcgError):
            e.watch_frequency_kernel(str(tmp_path / "k.txt"))
        e.process_all(x)   # the half-rate path has now run: the carried half window is even samples only
        with pytest.raises(rb.RcgError):
            e.watch_frequency_kernel(str(tmp_path / "k.c"))


@pytest.mark.gpu
def test_gpu_rust_kernel_without_a_rust_compiler(tmp_path, monkeypatch):
    """a reference-style .rs kernel is accepted; with no rustc the build fails like any other failed build
    (hotswapper.rs:36-42) and frames pass through (fft.rs:82-85).  The shim itself is checked on CPU
    (tests/cpp/hotswap_test.cpp)."""
    import rocoder_b200 as rb
    monkeypatch.setenv("RCG_RUSTC", "/nonexistent/rustc")
    x = _input(9000)
    path = tmp_path / "kernel.rs"
    path.write_text("#[no_mangle]\npub fn apply(_t: usize, input: Vec<(f32, f32)>) -> Vec<(f32, f32)> { input }\n")
    with rb.Engine(2, 48000, 1024, 2.0, 1.0, 1, seed=4) as plain:
        base = plain.process_all(x)
    with rb.Engine(2, 48000, 1024, 2.0, 1.0, 1, seed=4) as e:
        e.watch_frequency_kernel(str(path), poll_ms=20)
        assert e.frequency_kernel_count() == 0 and e.watch_stats()["builds_failed"] == 1 and e.watch_log() != ""
        assert np.abs(e.process_all(x) - base).max() < 2e-6


@pytest.mark.gpu
def test_gpu_hot_swap_while_streaming(tmp_path):
    """next_window picks the new kernel up between windows, like the per-frame poll at fft.rs:78-81."""
    import rocoder_b200 as rb
    f, p, W = 2.0, 1, 512
    x = _input(20000)
    path = tmp_path / "k.c"
    path.write_text(_host_src(2.0))
    ref2 = ol.process_all(x, f, p, W, seed=4, precision=64, plugin=_gain_plugin(2.0))
    ref3 = ol.process_all(x, f, p, W, seed=4, precision=64, plugin=_gain_plugin(3.0))
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=4, max_lookahead_calls=1) as e:   # one window per launch
        e.watch_frequency_kernel(str(path), poll_ms=20)
        for c in range(2):
            e.send(c, x[c]); e.close_input(c)
        olen = e.plan["out_len_per_call"]
        out = [[], []]
        for i in range(10):
            for c in range(2):
                out[c].append(e.next_window(c))
        _rewrite(path, _host_src(3.0))
        assert _wait(lambda: e.watch_stats()["builds_ok"] == 2), e.watch_log()
        while not e.is_done(0):
            for c in range(2):
                out[c].append(e.next_window(c))
        got = np.stack([np.concatenate(o) for o in out])
    assert got.shape == ref2.shape
    n0 = 10 * olen
    # windows emitted before the swap carry gain 2; after the swap (plus the one overlap block that mixes
    # a gain-2 tail into a gain-3 frame) gain 3
    assert np.abs(got[:, : n0 - W] - ref2[:, : n0 - W]).max() <= 1e-4
    assert np.abs(got[:, n0 + W:] - ref3[:, n0 + W:]).max() <= 1e-4
