"""In-tree build of libcogsworth_b200.so with nvcc for sm_100a (no torch, no JIT cache)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["cw_api.cu", "cw_leapfrog.cu", "cw_dop853.cu", "cw_diag.cu"]
HEADERS = ["cw_device.cuh", "cw_kernels.cuh", "../../include/cogsworth_b200.h", "../../include/cw_dop853_coeffs.h"]
OUT = os.environ.get("CW_BUILD_OUT") or os.path.join(HERE, "libcogsworth_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return OUT
    extra = []
    for knob in ("CW_TRAJ_NOSTORE", "CW_TRAJ_DEBUG", "CW_LF_REGS", "CW_LF_GEN_REGS", "CW_DP_MINB", "CW_STORE_UNROLL"):      # experiment knobs (see cw_traj.cuh / cw_leapfrog.cu)
        if os.environ.get(knob):
            extra.append(f"-D{knob}=" + os.environ[knob])
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    env = dict(os.environ)
    env.pop("CC", None)      # the image exports a CC without libgomp specs; nvcc wants the system g++
    env.pop("CXX", None)
    subprocess.check_call(cmd, cwd=CSRC, env=env)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
