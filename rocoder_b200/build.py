"""Build the native pieces in-tree.

  product : rocoder_b200/librocoder_b200.so   (nvcc, sm_100a only)
  test infrastructure (never imported by the product):
            oracle/build/librocoder_oracle.so (gcc; the parity oracle)
            tests/emul/librcg_emul.so         (g++; kernel phases executed on the host)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "rocoder_b200", "csrc")
LIB = os.path.join(ROOT, "rocoder_b200", "librocoder_b200.so")
ORACLE_LIB = os.path.join(ROOT, "oracle", "build", "librocoder_oracle.so")
ORACLE_FAST_LIB = os.path.join(ROOT, "oracle", "build", "librocoder_oracle_native.so")
EMUL_LIB = os.path.join(ROOT, "tests", "emul", "librcg_emul.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared",
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources(d: str, exts: tuple[str, ...]) -> list[str]:
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def _run(cmd: list[str], cwd: str | None = None) -> None:
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def build_product(force: bool = False, verbose_ptxas: bool = False) -> str:
    deps = _sources(CSRC, (".cu", ".h", ".cuh")) + [os.path.join(ROOT, "include", "rocoder_b200.h")]
    if not force and _newer(LIB, deps):
        return LIB
    cus = _sources(CSRC, (".cu",))
    flags = list(NVCC_FLAGS) + (["-Xptxas", "-v"] if verbose_ptxas else [])
    _run([nvcc_path(), *flags, "-o", LIB, *cus, "-ldl"])
    return LIB


def build_oracle(force: bool = False) -> str:
    d = os.path.join(ROOT, "oracle")
    deps = [os.path.join(d, f) for f in ("rocoder_oracle.c", "oracle_core.inc", "rocoder_oracle.h", "Makefile")]
    if force or not _newer(ORACLE_LIB, deps):
        _run(["make", "-B", "-C", d])
    return ORACLE_LIB


def build_oracle_native(force: bool = False) -> str:
    """-O3 -march=native build for the timed CPU baseline; built on the box it runs on."""
    d = os.path.join(ROOT, "oracle")
    deps = [os.path.join(d, f) for f in ("rocoder_oracle.c", "oracle_core.inc", "rocoder_oracle.h")]
    if force or not _newer(ORACLE_FAST_LIB, deps):
        os.makedirs(os.path.dirname(ORACLE_FAST_LIB), exist_ok=True)
        _run(["gcc", "-O3", "-march=native", "-fPIC", "-std=gnu99", "-ffp-contract=off", "-fno-math-errno",
              "-shared", "-o", ORACLE_FAST_LIB, os.path.join(d, "rocoder_oracle.c"), "-lm"])
    return ORACLE_FAST_LIB


def build_emulator(force: bool = False) -> str:
    src = os.path.join(ROOT, "tests", "emul", "rcg_emul.cpp")
    deps = [src] + _sources(CSRC, (".h",))
    if force or not _newer(EMUL_LIB, deps):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", EMUL_LIB, src])
    return EMUL_LIB


CLI = os.path.join(ROOT, "rocoder_b200", "rocoder_gpu")


def build_cli(force: bool = False) -> str:
    """tools/rocoder_gpu.cpp: the reference's CLI flags for the file-to-file path"""
    src = os.path.join(ROOT, "tools", "rocoder_gpu.cpp")
    deps = [src, os.path.join(ROOT, "rocoder_b200", "host", "rocoder_audio.hpp"), os.path.join(ROOT, "include", "rocoder_b200.h")]
    build_product()
    if force or not _newer(CLI, deps + [LIB]):
        libdir = os.path.dirname(LIB)
        _run(["g++", "-O2", "-std=c++17", src, "-o", CLI, "-L", libdir, "-lrocoder_b200", f"-Wl,-rpath,{libdir}", "-lpthread"])
    return CLI


def build_all(force: bool = False) -> None:
    build_product(force)
    build_cli(force)
    build_oracle(force)
    build_emulator(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", LIB, ORACLE_LIB, EMUL_LIB)
