"""Frequency-kernel path (reference fft.rs:76-108): the kernel's analysis / synthesis phases on the
host (CPU) and through the C ABI (GPU) against the oracle's literal plugin call site."""
import numpy as np
import pytest

import oracle_lib as ol
import plugins


def _input(n, seed=1):
    rng = np.random.default_rng(seed)
    return ((rng.random((2, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)


CASES = [(6000, 2.0, 1, 256), (9000, 3.0, 2, 1024), (7000, 4.0, -2, 512), (40000, 2.0, 1, 16384)]
PAIR_CASES = [(150000, 2.0, 1, 65536), (150000, 3.0, 2, 65536), (150000, 2.0, -2, 65536)]   # frame across a CTA pair
PLUGS = [plugins.amplitude_x2, plugins.lowpass_asymmetric, plugins.tilt]


@pytest.mark.parametrize("n,f,p,W", CASES[:3])
@pytest.mark.parametrize("plug", PLUGS, ids=lambda f: f.__name__)
def test_emulated_plugin_path_matches_oracle(n, f, p, W, plug):
    x = _input(n)
    ref = ol.process_all(x, f, p, W, seed=4, precision=64, plugin=plug)
    got = ol.emul_process_all_plugin(x, f, p, W, plug, seed=4)
    assert np.abs(got - ref).max() <= 1e-4 and ol.snr_db(ref, got) > 100.0


@pytest.mark.parametrize("n,f,p,W", PAIR_CASES)
@pytest.mark.parametrize("plug", [plugins.lowpass_asymmetric, plugins.tilt], ids=lambda f: f.__name__)
def test_emulated_plugin_path_matches_oracle_w65536(n, f, p, W, plug):
    x = _input(n)
    ref = ol.process_all(x, f, p, W, seed=4, precision=64, plugin=plug)
    got = ol.emul_process_all_plugin(x, f, p, W, plug, seed=4)
    assert np.abs(got - ref).max() <= 1e-4 and ol.snr_db(ref, got) > 100.0


@pytest.mark.parametrize("n,f,p,W", [(6000, 2.0, 1, 256), (9000, 3.0, 2, 1024), (7000, 4.0, -2, 512), (5000, 1.0, 12, 128),
                                     (40000, 8.0, 2, 16384), (100, 1.0, 1, 64), (150000, 2.0, 1, 65536)])
def test_emulated_fused_local_plugin_matches_oracle(n, f, p, W):
    """the hook that evaluates a bin-local frequency kernel inside the spectral step (what NVRTC compiles into the
    stretch kernel, rcg_jit_entry.h) executed on the host -- full-rate, half-rate and CTA-pair paths, special bins
    included -- against the oracle calling the same kernel (`tilt`) on the whole spectrum"""
    x = _input(n)
    ref = ol.process_all(x, f, p, W, seed=4, precision=64, plugin=plugins.tilt)
    got = ol.emul_process_all(x, f, p, W, seed=4, lib=ol.emulator_local_plugin())
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-4 and (not ref.any() or ol.snr_db(ref, got) > 100.0)
    plain = ol.emul_process_all(x, f, p, W, seed=4)
    assert not np.array_equal(plain, got)
    if W <= 16384:   # phases stay order-independent with the hook compiled in
        E = ol.emulator_local_plugin()
        E.rcg_emul_reverse_threads(1)
        try:
            again = ol.emul_process_all(x, f, p, W, seed=4, lib=E)
        finally:
            E.rcg_emul_reverse_threads(0)
        assert np.array_equal(got, again)


def test_identity_plugin_equals_fused_path():
    """pass-through spectra (no kernel loaded yet, fft.rs:82-85) reproduce the fused kernel up to
    the rounding of the two different magnitude formulas"""
    x = _input(8000)
    def ident(ms, pin, n, pout):
        a, b = plugins._views(pin, n, pout)
        b[:] = a
        return 0
    a = ol.emul_process_all(x, 2.0, 2, 512, seed=3)
    b = ol.emul_process_all_plugin(x, 2.0, 2, 512, ident, seed=3)
    assert np.abs(a - b).max() < 1e-6


def test_amplitude_plugin_doubles_output():
    """only magnitudes survive (fft.rs:67) and everything downstream is linear in them"""
    x = _input(8000)
    a = ol.process_all(x, 2.0, 1, 512, seed=3)
    b = ol.process_all(x, 2.0, 1, 512, seed=3, plugin=plugins.amplitude_x2)
    assert np.array_equal(b, a * np.float32(2.0))


def test_oracle_plugin_panic_falls_back():
    """fft.rs:101-106: a panicking kernel is popped and the frame passes through"""
    x = _input(5000)
    a = ol.process_all(x, 2.0, 1, 256, seed=3)
    b = ol.process_all(x, 2.0, 1, 256, seed=3, plugin=plugins.panics)
    assert np.array_equal(a, b)


# ---------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("n,f,p,W", CASES + PAIR_CASES)
@pytest.mark.parametrize("plug", PLUGS, ids=lambda f: f.__name__)
def test_gpu_plugin_path_matches_oracle(n, f, p, W, plug):
    import ctypes as C
    import rocoder_b200 as rb
    from rocoder_b200 import _abi
    x = _input(n)
    ref = ol.process_all(x, f, p, W, seed=4, precision=64, plugin=plug)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=4) as e:
        cb = _abi.APPLY_FN(plug)
        _abi.check(_abi.lib().rcg_push_frequency_kernel(e._h, cb))
        assert e.frequency_kernel_count() == 1
        got = e.process_all(x)
    assert np.abs(got - ref).max() <= 1e-4 and ol.snr_db(ref, got) > 100.0


@pytest.mark.gpu
def test_gpu_plugin_python_callable_and_fallback():
    import rocoder_b200 as rb
    x = _input(9000)
    with rb.Engine(2, 48000, 1024, 2.0, 1.0, 1, seed=4) as e:
        plain = e.process_all(x)
    with rb.Engine(2, 48000, 1024, 2.0, 1.0, 1, seed=4) as e:
        e.push_frequency_kernel(lambda ms, spec: spec * np.float32(2.0))
        doubled = e.process_all(x)
    assert np.abs(doubled - plain * np.float32(2.0)).max() < 2e-6   # fused vs plugin path: different rounding
    with rb.Engine(2, 48000, 1024, 2.0, 1.0, 1, seed=4) as e:
        e.push_frequency_kernel(lambda ms, spec: spec * np.float32(2.0))   # older kernel
        def bad(ms, spec):
            raise RuntimeError("panic")
        e.push_frequency_kernel(bad)                                       # newest kernel panics
        out = e.process_all(x)
        assert e.frequency_kernel_count() == 1                             # popped (fft.rs:104)
    assert np.array_equal(out, doubled)


@pytest.mark.gpu
def test_gpu_plugin_streaming_and_shared_object(tmp_path):
    """hot-swap role: compile a C kernel to a shared object, load it (rcg_load_frequency_kernel)"""
    import subprocess
    import rocoder_b200 as rb
    src = tmp_path / "kernel.c"
    src.write_text("#include <stddef.h>\n"
                   "int apply(size_t ms, const float* in, size_t n, float* out) {\n"
                   "  for (size_t i = 0; i < 2 * n; i++) out[i] = in[i] * 2.0f; return 0; }\n")
    so = tmp_path / "kernel.so"
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", str(so), str(src)])
    x = _input(9000)[:1]
    ref = ol.process_all(x, 2.0, 1, 1024, seed=4, precision=64, plugin=plugins.amplitude_x2)
    with rb.Engine(1, 48000, 1024, 2.0, 1.0, 1, seed=4, max_lookahead_calls=2) as e:
        e.load_frequency_kernel(str(so))
        e.send(0, x[0])
        e.close_input(0)
        wins = []
        while not e.is_done(0):
            wins.append(e.next_window(0))
    got = np.concatenate(wins)[None, :]
    assert np.abs(got - ref).max() <= 1e-4 and ol.snr_db(ref, got) > 100.0


DEVICE_X2, DEVICE_TILT, DEVICE_LOWPASS = plugins.DEVICE_X2, plugins.DEVICE_TILT, plugins.DEVICE_LOWPASS


@pytest.mark.gpu
@pytest.mark.parametrize("src,host", [(DEVICE_X2, plugins.amplitude_x2), (DEVICE_TILT, plugins.tilt),
                                      (DEVICE_LOWPASS, plugins.lowpass_asymmetric)], ids=["x2", "tilt", "lowpass"])
@pytest.mark.parametrize("n,f,p,W", [(9000, 3.0, 2, 1024), (60000, 2.0, 1, 16384), (150000, 3.0, 2, 65536)])
def test_gpu_device_frequency_kernel_matches_oracle(src, host, n, f, p, W):
    """NVRTC-compiled device `apply` (spectra stay in HBM) == the oracle with the same host kernel"""
    import rocoder_b200 as rb
    x = _input(n)
    ref = ol.process_all(x, f, p, W, seed=4, precision=64, plugin=host)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=4) as e:
        e.push_device_frequency_kernel(src)
        got = e.process_all(x)
        assert e.launch_count() >= 3   # analysis + apply + synthesis
    assert np.abs(got - ref).max() <= 1e-4 and ol.snr_db(ref, got) > 100.0


@pytest.mark.gpu
def test_gpu_device_kernel_compile_error_keeps_previous():
    """hotswapper.rs:36-42: a kernel that fails to compile is dropped, the old one stays"""
    import rocoder_b200 as rb
    x = _input(9000)
    with rb.Engine(2, 48000, 1024, 2.0, 1.0, 1, seed=4) as e:
        e.push_device_frequency_kernel(DEVICE_X2)
        with pytest.raises(rb.RcgError) as ei:
            e.push_device_frequency_kernel("this is not CUDA")
        assert ei.value.status == 9 and e.frequency_kernel_count() == 1
        got = e.process_all(x)
    ref = ol.process_all(x, 2.0, 1, 1024, seed=4, precision=64, plugin=plugins.amplitude_x2)
    assert np.abs(got - ref).max() <= 1e-4


LOCAL_X2, LOCAL_TILT, LOCAL_LOWPASS = plugins.LOCAL_X2, plugins.LOCAL_TILT, plugins.LOCAL_LOWPASS


@pytest.mark.gpu
@pytest.mark.parametrize("src,host", [(LOCAL_X2, plugins.amplitude_x2), (LOCAL_TILT, plugins.tilt),
                                      (LOCAL_LOWPASS, plugins.lowpass_asymmetric)], ids=["x2", "tilt", "lowpass"])
@pytest.mark.parametrize("n,f,p,W", [(3000, 2.0, 1, 64), (9000, 3.0, 2, 1024), (20000, 4.0, -2, 4096), (60000, 2.0, 1, 16384),
                                     (150000, 2.0, 3, 32768), (200000, 2.0, 1, 65536), (200000, 3.0, -2, 65536)])
def test_gpu_fused_local_frequency_kernel_matches_oracle(src, host, n, f, p, W):
    """bin-local device `apply` compiled INTO the stretch kernel by NVRTC: one launch, no spectra in memory, same
    result as the oracle calling the equivalent host kernel on the full spectrum (non-Hermitian outputs included)"""
    import rocoder_b200 as rb
    x = _input(n)
    ref = ol.process_all(x, f, p, W, seed=4, precision=64, plugin=host)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=4) as e:
        e.push_local_device_frequency_kernel(src)
        assert e.frequency_kernel_count() == 1
        l0 = e.launch_count()
        got = e.process_all(x)
        assert e.launch_count() - l0 <= 2          # the fused kernel (per chunk), not analysis + apply + synthesis
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-4 and ol.snr_db(ref, got) > 100.0


@pytest.mark.gpu
def test_gpu_fused_local_kernel_errors_and_stack():
    import rocoder_b200 as rb
    x = _input(9000)
    ref2 = ol.process_all(x, 2.0, 1, 1024, seed=4, precision=64, plugin=plugins.amplitude_x2)
    with rb.Engine(2, 48000, 1024, 2.0, 1.0, 1, seed=4) as e:
        e.push_device_frequency_kernel(DEVICE_TILT)                   # older kernel, gather form
        e.push_local_device_frequency_kernel(LOCAL_X2)                # newest kernel wins (fft.rs:82 `kernels.last()`)
        with pytest.raises(rb.RcgError) as ei:
            e.push_local_device_frequency_kernel("float2 apply(int oops) { return nothing; }")
        assert ei.value.status == 9 and e.frequency_kernel_count() == 2
        got = e.process_all(x)
    assert np.abs(got - ref2).max() <= 1e-4


@pytest.mark.gpu
def test_gpu_fused_local_kernel_keeps_half_rate_and_can_follow_processing():
    """a bin-local kernel runs inside the stretch kernel, so the even-pitch half-rate path stays on and the kernel
    may be pushed after audio was processed; a gather-form kernel may not (the carried half window differs)"""
    import rocoder_b200 as rb
    x = _input(40000)
    f, p, W = 8.0, 2, 4096
    ref = ol.process_all(x, f, p, W, seed=4, precision=64, plugin=plugins.tilt)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=4) as e:
        plain = e.process_all(x)                                      # half-rate fused path
        e.push_local_device_frequency_kernel(LOCAL_TILT)              # allowed after processing
        got = e.process_all(x)
        with pytest.raises(rb.RcgError) as ei:
            e.push_device_frequency_kernel(DEVICE_TILT)
        assert ei.value.status == 10
    assert np.abs(got - ref).max() <= 1e-4 and ol.snr_db(ref, got) > 100.0
    assert not np.array_equal(got, plain)


@pytest.mark.gpu
def test_gpu_fused_local_kernel_streaming_equals_batch():
    import rocoder_b200 as rb
    x = _input(30000)
    with rb.Engine(2, 48000, 2048, 3.0, 1.0, 1, seed=4) as e:
        e.push_local_device_frequency_kernel(LOCAL_TILT)
        batch = e.process_all(x)
    with rb.Engine(2, 48000, 2048, 3.0, 1.0, 1, seed=4, max_lookahead_calls=3) as e:
        e.push_local_device_frequency_kernel(LOCAL_TILT)
        for c in range(2):
            e.send(c, x[c]); e.close_input(c)
        outs = [[], []]
        while not e.is_done(0):
            for c in range(2):
                outs[c].append(e.next_window(c))
    assert np.array_equal(np.stack([np.concatenate(o) for o in outs]), batch)


@pytest.mark.gpu
@pytest.mark.parametrize("beneath", ["gather", "fused"])
@pytest.mark.parametrize("good", [0, 7, 23])
def test_gpu_mixed_stack_retries_with_the_kernel_beneath_mid_batch(beneath, good):
    """fft.rs:101-106 on a mixed stack: a host kernel that panics in the middle of a batch is dropped and the SAME frame
    is retried with the kernel beneath it -- here a device kernel (gather form, or a fused bin-local one through its
    per-bin twin); frames already processed keep the host kernel's result.  The oracle runs the equivalent single
    kernel 'half_gain for the first `good` frames in the reference's call order, tilt afterwards'."""
    import rocoder_b200 as rb
    n, f, p, W = 30000, 2.0, 1, 1024
    x = _input(n)
    ref = ol.process_all(x, f, p, W, seed=4, precision=64, plugin=plugins.GainThenTilt(good))
    host = plugins.FailsAfter(good)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=4) as e:
        if beneath == "gather":
            e.push_device_frequency_kernel(plugins.DEVICE_TILT)
        else:
            e.push_local_device_frequency_kernel(plugins.LOCAL_TILT)
        from rocoder_b200 import _abi
        cb = _abi.APPLY_FN(host)
        e._keep.append(cb)     # the callback object must outlive the engine
        _abi.check(_abi.lib().rcg_push_frequency_kernel(e._h, cb))
        assert e.frequency_kernel_count() == 2
        got = e.process_all(x)
        assert e.frequency_kernel_count() == 1      # the host kernel is gone for good
        again = e.process_all(x)                    # and the next job runs on the device kernel alone
    assert host.calls == good + 1
    assert np.abs(got - ref).max() <= 1e-4 and ol.snr_db(ref, got) > 100.0
    tilt_only = ol.process_all(x, f, p, W, seed=4, precision=64, plugin=plugins.tilt)
    assert np.abs(again - tilt_only).max() <= 1e-4 and ol.snr_db(tilt_only, again) > 100.0


@pytest.mark.gpu
def test_gpu_bad_segment_argument_does_not_drop_the_kernel():
    """a caller error in a segment description is not a failing frequency kernel: the stack stays as it was"""
    import torch
    import rocoder_b200 as rb
    x = torch.zeros((2, 20000), device="cuda")
    out = torch.zeros((2, 50000), device="cuda")
    with rb.Engine(2, 48000, 1024, 2.0, 1.0, 1, seed=4) as e:
        e.push_local_device_frequency_kernel(plugins.LOCAL_X2)
        with pytest.raises(rb.RcgError):   # output of a direct-store mode at an odd (4-byte aligned) offset
            e.process_segment(x, x.stride(0), 0, 20000, out[:, 1:], out.stride(0), 0, 2,
                              stream=torch.cuda.current_stream().cuda_stream)
        with pytest.raises(rb.RcgError):   # deferred first block without a place for its head
            e.process_segment(x, x.stride(0), 0, 20000, out, out.stride(0), 2, 2, defer_first=True,
                              stream=torch.cuda.current_stream().cuda_stream)
        assert e.frequency_kernel_count() == 1


GATE_LOCAL = """
__device__ float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins, float2 x) {
    return elapsed_ms >= GATE_MS ? x : make_float2(0.0f, 0.0f);     // silence until the clock reaches GATE_MS
}
"""
GATE_GATHER = """
__device__ float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins, const float2* spectrum) {
    return elapsed_ms >= GATE_MS ? spectrum[bin] : make_float2(0.0f, 0.0f);
}
"""


@pytest.mark.gpu
@pytest.mark.parametrize("form", ["fused", "gather"])
@pytest.mark.parametrize("f,p,W", [(2.0, 1, 1024), (4.0, 2, 4096)])
def test_gpu_frame_clock_for_device_kernels(form, f, p, W):
    """RCG_FLAG_FRAME_CLOCK: a device `apply` sees elapsed_ms advance frame by frame (the reference reads the clock
    before every frame, fft.rs:89-92) -- clock base + the frame's position in output time.  A kernel that gates on the
    clock silences exactly the frames before the gate and leaves the others as the identity kernel leaves them, in
    every batch / launch split (one shot, streaming, three devices)."""
    import rocoder_b200 as rb
    n, rate, base = 60000, 48000, 5_000_000
    x = _input(n)
    with rb.Engine(2, rate, W, f, 1.0, p, seed=4) as e:
        plan = e.plan
        calls, frames, out_len = e.expected_counts(n)
    k, olen, H = plan["frames_per_call"], plan["out_len_per_call"], W // 2
    gate_frame = frames // 2 + 1
    num, den = 1000 * olen, k * rate
    gate_ms = base + gate_frame * num // den
    while (gate_frame - 1) * num // den == gate_frame * num // den:     # the gate must fall between two distinct clock values
        gate_frame += 1
        gate_ms = base + gate_frame * num // den
    src = ("#define GATE_MS %dull\n" % gate_ms) + (GATE_LOCAL if form == "fused" else GATE_GATHER)
    ident = "#define GATE_MS 0ull\n" + (GATE_LOCAL if form == "fused" else GATE_GATHER)
    push = "push_local_device_frequency_kernel" if form == "fused" else "push_device_frequency_kernel"

    def run(source, **kw):
        with rb.Engine(2, rate, W, f, 1.0, p, seed=4, flags=rb.RCG_FLAG_FRAME_CLOCK | kw.pop("flags", 0), **kw) as e:
            e.set_clock_base(base)
            getattr(e, push)(source)
            return e.process_all(x)
    open_ = run(ident)
    gated = run(src)
    # output block m mixes frames m and m-1: blocks before the gate frame are silent, blocks after it untouched
    per_block = olen / k                     # output samples per frame block (exact for the pitches used here)
    lo, hi = int((gate_frame - 1) * per_block), int((gate_frame + 1) * per_block)
    assert not gated[:, :lo].any() and open_[:, :lo].any()
    assert np.array_equal(gated[:, hi:], open_[:, hi:]) and gated[:, hi:].any()
    # the clock belongs to the frame, not to the launch: other splits of the same job give the same samples
    assert np.array_equal(run(src, devices=[0, 0, 0], flags=rb.RCG_FLAG_SHARD_ALWAYS), gated)
    with rb.Engine(1, rate, W, f, 1.0, p, seed=4, flags=rb.RCG_FLAG_FRAME_CLOCK, max_lookahead_calls=3) as e:
        e.set_clock_base(base)
        getattr(e, push)(src)
        e.send(0, x[0]); e.close_input(0)
        wins = []
        while not e.is_done(0):
            wins.append(e.next_window(0))
    assert np.array_equal(np.concatenate(wins), gated[0])
