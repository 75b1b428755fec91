"""WAV reader/writer, -s/-d clipping, duration parser and the CLI with the
reference's flags.  Host helpers are checked on CPU against the reference's own unit tests; the
CLI end to end on GPU against the oracle."""
import os
import struct
import subprocess
import wave

import numpy as np
import pytest

import oracle_lib as ol


def _write_wav(path, x, rate, bits):
    """x: [C, N] float32 in [-1, 1]; integer formats quantise like a normal encoder"""
    ch, n = x.shape
    inter = x.T.reshape(-1)
    if bits == 32.5:   # 32-bit float
        hdr = b"RIFF" + struct.pack("<I", 36 + 4 * inter.size) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 3, ch, rate, rate * ch * 4, ch * 4, 32)
        open(path, "wb").write(hdr + b"data" + struct.pack("<I", 4 * inter.size) + inter.astype("<f4").tobytes())
        return x.copy()
    if bits == 16:
        q = np.clip(np.round(inter * 32767), -32768, 32767).astype("<i2")
        with wave.open(str(path), "wb") as w:
            w.setnchannels(ch); w.setsampwidth(2); w.setframerate(rate); w.writeframes(q.tobytes())
        return (q.astype(np.float32) / np.float32(32767)).reshape(n, ch).T.copy()       # audio.rs:20-22
    if bits == 24:
        q = np.clip(np.round(inter * 8388607), -8388608, 8388607).astype(np.int32)
        raw = b"".join(struct.pack("<i", int(v))[:3] for v in q)
        with wave.open(str(path), "wb") as w:
            w.setnchannels(ch); w.setsampwidth(3); w.setframerate(rate); w.writeframes(raw)
        return (q.astype(np.float32) / np.float32(8388608.0)).reshape(n, ch).T.copy()   # audio.rs:23-25
    if bits == 8:
        q = np.clip(np.round(inter * 127), -128, 127).astype(np.int16)
        with wave.open(str(path), "wb") as w:
            w.setnchannels(ch); w.setsampwidth(1); w.setframerate(rate); w.writeframes((q + 128).astype(np.uint8).tobytes())
        return (q.astype(np.float32) / np.float32(127)).reshape(n, ch).T.copy()         # audio.rs:17-19
    raise ValueError(bits)


def _read_wav_f32(path):
    b = open(path, "rb").read()
    assert b[:4] == b"RIFF" and b[8:12] == b"WAVE"
    fmt, ch, rate, _, _, bits = struct.unpack("<HHIIHH", b[20:36])
    assert (fmt, bits) == (3, 32)   # audio_files.rs:203-212: 32-bit float output
    n = struct.unpack("<I", b[40:44])[0]
    data = np.frombuffer(b[44:44 + n], dtype="<f4")
    return data.reshape(-1, ch).T.copy(), rate


def test_audio_host_helpers_against_reference_unit_tests(tmp_path):
    exe = str(tmp_path / "audio_host_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", os.path.join(ol.ROOT, "tests", "cpp", "audio_host_test.cpp"), "-o", exe])
    r = subprocess.run([exe, str(tmp_path / "rt.wav")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_cli_builds_and_rejects_device_modes():
    from rocoder_b200 import build
    exe = build.build_cli()
    r = subprocess.run([exe, "-o", "x.wav"], capture_output=True, text=True)
    assert r.returncode == 1 and "recording" in r.stderr          # main.rs:172-175 records from a device
    r = subprocess.run([exe, "-i", "x.wav"], capture_output=True, text=True)
    assert r.returncode == 1 and "playback" in r.stderr           # main.rs:205-207 plays to a device


@pytest.mark.gpu
@pytest.mark.parametrize("bits", [8, 16, 24, 32.5])
def test_cli_file_to_file_matches_oracle(tmp_path, bits):
    """BASELINE config 1's shape: `rocoder -i in.wav -f .. -p .. -w .. -o out.wav` incl. -s/-d/-a/--rotate-channels"""
    from rocoder_b200 import build
    exe = build.build_cli()
    rate, n = 44100, 44100 * 2
    rng = np.random.default_rng(int(bits))
    x = ((rng.random((2, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)
    decoded = _write_wav(tmp_path / "in.wav", x, rate, bits)
    args = ["-f", "2", "-p", "2", "-w", "4096", "-a", "0.5", "-s", "0.25", "-d", "1.5", "--rotate-channels", "--seed", "77"]
    r = subprocess.run([exe, "-i", str(tmp_path / "in.wav"), "-o", str(tmp_path / "out.wav"), *args], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got, out_rate = _read_wav_f32(tmp_path / "out.wav")
    assert out_rate == rate
    s0, s1 = int(0.25 * rate), int(0.25 * rate) + int(1.5 * rate)   # audio.rs:115-130
    clipped = decoded[:, s0:s1][::-1].copy()                          # rotate_right(1) on 2 channels = swap
    ref = ol.process_all(clipped, 2.0, 2, 4096, seed=77, amplitude=0.5, precision=64, sample_rate=rate)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-4 and ol.snr_db(ref, got) > 100.0


@pytest.mark.gpu
def test_cli_frequency_kernel_source(tmp_path):
    from rocoder_b200 import build
    import plugins
    exe = build.build_cli()
    x = ((np.random.default_rng(1).random((1, 30000), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)
    _write_wav(tmp_path / "in.wav", x, 48000, 32.5)
    src = tmp_path / "kernel.cpp"
    src.write_text('#include <stddef.h>\nextern "C" int apply(size_t ms, const float* in, size_t n, float* out) {\n'
                   "  for (size_t i = 0; i < 2 * n; i++) out[i] = in[i] * 2.0f; return 0; }\n")
    r = subprocess.run([exe, "-i", str(tmp_path / "in.wav"), "-o", str(tmp_path / "out.wav"), "-f", "2", "-w", "1024",
                        "--freq-kernel", str(src), "--seed", "5"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got, _ = _read_wav_f32(tmp_path / "out.wav")
    ref = ol.process_all(x, 2.0, 1, 1024, seed=5, precision=64, plugin=plugins.amplitude_x2)
    assert np.abs(got - ref).max() <= 1e-4 and ol.snr_db(ref, got) > 100.0


@pytest.mark.gpu
def test_cli_device_list_partitions_the_job(tmp_path):
    """`--device 0,0,...`: one process, every listed device takes its share (rcg_config.n_devices); the file written
    is bit-identical to the single-device run.  (The same ordinal twice shares that GPU: enough to run on any box.)"""
    from rocoder_b200 import build
    import json
    exe = build.build_cli()
    rate, n = 48000, 48000 * 30                     # 30 s stereo: big enough for the engine to use both devices
    rng = np.random.default_rng(3)
    x = ((rng.random((2, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)
    _write_wav(tmp_path / "in.wav", x, rate, 32.5)
    outs = []
    for devs, want in (("0", (1, "none", 0)), ("0,0", (2, "channel", 0)), ("0,0,0", (3, "time", 2))):
        r = subprocess.run([exe, "-i", str(tmp_path / "in.wav"), "-o", str(tmp_path / f"out_{len(outs)}.wav"), "-f", "4", "-p", "-2",
                            "-w", "4096", "--seed", "9", "--device", devs, "--bench"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        info = json.loads(r.stdout.strip().splitlines()[-1])
        assert (info["devices_used"], info["split"], info["exchanges"]) == want
        outs.append(_read_wav_f32(tmp_path / f"out_{len(outs)}.wav")[0])
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    r = subprocess.run([exe, "-i", str(tmp_path / "in.wav"), "-o", str(tmp_path / "o.wav"), "--device", "0,99"], capture_output=True, text=True)
    assert r.returncode == 1 and "device" in r.stderr
