"""In-tree build of the sm_100a library (and the device self-test binary).

    python -m video_fingerprint_b200.build            # libvfp_b200.so
    python -m video_fingerprint_b200.build selftest   # build/selftest_gemm

nvcc cross-compiles for sm_100a without a GPU; the resulting .so lives next to this file so it travels
with the source tree (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libvfp_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
]


def _sources_mtime() -> float:
    newest = 0.0
    for d in (CSRC, os.path.join(ROOT, "include")):
        for name in os.listdir(d):
            newest = max(newest, os.path.getmtime(os.path.join(d, name)))
    return newest


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _sources_mtime():
        return LIB_PATH
    tmp = LIB_PATH + ".tmp%d" % os.getpid()   # written aside and renamed: a snapshot of the tree never sees a half-written library
    cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-o", tmp, os.path.join(CSRC, "vfp_b200.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(res.stderr)
    return LIB_PATH


DEVICE_TOOLS = ("selftest_gemm", "selftest_ts_mma", "microbench_tensor", "microbench_handoff", "microbench_tmem")


def build_selftest(name: str = "selftest_gemm") -> str:
    """Device self-tests and micro-benchmarks under tests/cuda/ (stand-alone binaries, run on a B200)."""
    out_dir = os.path.join(ROOT, "build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, name)
    src = os.path.join(ROOT, "tests", "cuda", name + ".cu")
    if os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(src), _sources_mtime()):
        return out
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", out, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out


def build_tools() -> list:
    return [build_selftest(n) for n in DEVICE_TOOLS]


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "selftest":
        print("\n".join(build_tools()))
    else:
        print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
