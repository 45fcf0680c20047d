"""In-tree build of libcogsworth_b200.so with nvcc for sm_100a (no torch, no JIT cache).

Every translation unit is compiled by its own nvcc process (in parallel) into ``csrc/_obj/`` and re-compiled only
when it, or a header, is newer than its object file; the objects are then linked into the shared library.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["cw_api.cu", "cw_host.cu", "cw_leapfrog.cu", "cw_dop853.cu", "cw_diag.cu"]
HEADERS = ["cw_device.cuh", "cw_kernels.cuh", "cw_traj.cuh", "cw_host.h", "../../include/cogsworth_b200.h",
           "../../include/cw_dop853_coeffs.h"]
OUT = os.environ.get("CW_BUILD_OUT") or os.path.join(HERE, "libcogsworth_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread",
              "-lcuda"]
# experiment knobs (see cw_traj.cuh / cw_leapfrog.cu / cw_dop853.cu): -DNAME=VALUE when set in the environment
KNOBS = ("CW_TRAJ_NOSTORE", "CW_TRAJ_DEBUG", "CW_LF_REGS", "CW_LF_GEN_REGS", "CW_DP_MINB", "CW_STORE_UNROLL",
         "CW_LF_VARIANT", "CW_DP_VARIANT", "CW_TRAJ_VARIANT", "CW_CHECKED", "CW_LF_BLOCK", "CW_TRAJ_S", "CW_LF_STORE_REGS", "CW_TRAJ_TMEM", "CW_TRAJ_WIDE", "CW_TRAJ_TSTAGE", "CW_DP_SPEC", "CW_DP_BLOCK")


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _knob_flags():
    return [f"-D{k}=" + os.environ[k] for k in KNOBS if os.environ.get(k)]


def _obj_dir():
    # objects of a variant build (CW_BUILD_OUT / knobs) must not be mixed with the default build's
    tag = os.path.splitext(os.path.basename(OUT))[0]
    d = os.path.join(CSRC, "_obj", tag)
    os.makedirs(d, exist_ok=True)
    return d


def _newest_header() -> float:
    return max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)


def _stale_sources(force: bool):
    d, th = _obj_dir(), _newest_header()
    flags_file = os.path.join(d, "flags.txt")
    flags = " ".join(NVCC_FLAGS + _knob_flags())
    if not os.path.exists(flags_file) or open(flags_file).read() != flags:
        force = True
    out = []
    for s in SOURCES:
        o = os.path.join(d, s[:-3] + ".o")
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(os.path.join(CSRC, s)), th):
            out.append(s)
    return out, flags_file, flags


def is_stale() -> bool:
    if not os.path.exists(OUT):
        return True
    stale, _, _ = _stale_sources(False)
    if stale:
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(_obj_dir(), s[:-3] + ".o")) > t for s in SOURCES)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return OUT
    env = dict(os.environ)
    env.pop("CC", None)      # the image exports a CC without libgomp specs; nvcc wants the system g++
    env.pop("CXX", None)
    nvcc, d = _nvcc(), _obj_dir()
    stale, flags_file, flags = _stale_sources(force)
    extra = _knob_flags() + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", "-o", os.path.join(d, src[:-3] + ".o"), src]
        r = subprocess.run(cmd, cwd=CSRC, env=env, capture_output=True, text=True)
        return src, r

    with ThreadPoolExecutor(max_workers=max(1, min(len(stale), os.cpu_count() or 1))) as ex:
        results = list(ex.map(compile_one, stale))
    for src, r in results:
        if verbose or r.returncode != 0:
            print(f"---- {src}\n{r.stdout}{r.stderr}")
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    open(flags_file, "w").write(flags)
    objs = [os.path.join(d, s[:-3] + ".o") for s in SOURCES]
    subprocess.check_call([nvcc] + LINK_FLAGS + ["-o", OUT] + objs, cwd=CSRC, env=env)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
