"""ctypes binding of libcogsworth_b200.so (C ABI in include/cogsworth_b200.h).

This is the only place the product touches native code.  The library is built in-tree by
``__graft_entry__.build()`` / ``cogsworth_b200.build.build()``; if it is missing, or there is no
CUDA device, everything here raises -- there is deliberately NO CPU fallback on this path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# COGSWORTH_B200_LIB selects another build of the same ABI (A/B experiments); default = the in-tree build
LIB_PATH = os.environ.get("COGSWORTH_B200_LIB") or os.path.join(_HERE, "libcogsworth_b200.so")

LEAPFROG, DOP853, DOP853_DENSE = 0, 1, 2
GRID_APPEND_T2 = 1
MEM_HOST, MEM_DEVICE = 0, 1

ORBIT_STATUS = {0: "ok", -1: "inconsistent input", -2: "larger nmax is needed",
                -3: "step size becomes too small", -4: "problem is probably stiff",
                -5: "non-finite error norm", -6: "trajectory slot too small"}

EXPORTS = ["cw_create", "cw_destroy", "cw_device_count", "cw_set_potential", "cw_set_potential_ex", "cw_set_potential_general",
           "cw_set_potential_time", "cw_count_nodes",
           "cw_integrate", "cw_gradient", "cw_potential_value", "cw_circular_velocity", "cw_pointwise_t", "cw_strerror",
           "cw_last_error", "cw_abi_version", "cw_last_launch_count", "cw_last_kernel_ms",
           "cw_last_total_kernel_ms", "cw_measure_fp64_peak", "cw_measure_scatter_write", "cw_selftest_math"]


class CwOpts(C.Structure):
    _fields_ = [("integrator", C.c_int32), ("mem_space", C.c_int32), ("atol", C.c_double),
                ("rtol", C.c_double), ("nmax", C.c_int64), ("stream", C.c_void_p),
                ("no_sync", C.c_int32), ("reserved", C.c_int32)]


class CwBatch(C.Structure):
    _fields_ = [("n_orbits", C.c_int64), ("w0", C.c_void_p), ("t1", C.c_void_p), ("t2", C.c_void_p),
                ("dt", C.c_void_p), ("to_myr", C.c_void_p), ("grid_flags", C.c_void_p),
                ("ev_off", C.c_void_p), ("ev_time", C.c_void_p), ("ev_dv", C.c_void_p)]


class CwResult(C.Structure):
    _fields_ = [("w_final", C.c_void_p), ("status", C.c_void_p), ("n_nodes", C.c_void_p),
                ("ev_step", C.c_void_p), ("traj_off", C.c_void_p), ("traj_pos", C.c_void_p),
                ("traj_vel", C.c_void_p), ("traj_t", C.c_void_p), ("stats", C.c_void_p)]


class NativeLibraryError(RuntimeError):
    pass


_LIB = None


def load():
    """Load the shared library; raises NativeLibraryError if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  cogsworth_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.cw_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(C.c_int)]
    L.cw_create.restype = C.c_int
    L.cw_destroy.argtypes = [vp]
    L.cw_destroy.restype = None
    L.cw_device_count.argtypes = [vp]
    L.cw_device_count.restype = C.c_int
    L.cw_set_potential.argtypes = [vp, C.c_double, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    L.cw_set_potential.restype = C.c_int
    L.cw_set_potential_ex.argtypes = [vp, C.c_double, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_int]
    L.cw_set_potential_ex.restype = C.c_int
    L.cw_set_potential_general.argtypes = [vp, C.c_double, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_int,
                                           C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                           C.POINTER(C.c_double), C.c_int]
    L.cw_set_potential_general.restype = C.c_int
    L.cw_set_potential_time.argtypes = [vp, C.c_double]
    L.cw_set_potential_time.restype = C.c_int
    L.cw_count_nodes.argtypes = [vp, C.POINTER(CwOpts), C.POINTER(CwBatch), vp, vp]
    L.cw_count_nodes.restype = C.c_int
    L.cw_integrate.argtypes = [vp, C.POINTER(CwOpts), C.POINTER(CwBatch), C.POINTER(CwResult)]
    L.cw_integrate.restype = C.c_int
    for name in ("cw_gradient", "cw_potential_value", "cw_circular_velocity"):
        f = getattr(L, name)
        f.argtypes = [vp, C.c_int, C.c_int64, vp, vp]
        f.restype = C.c_int
    L.cw_pointwise_t.argtypes = [vp, C.c_int, C.c_int, C.c_int64, vp, vp, vp]
    L.cw_pointwise_t.restype = C.c_int
    L.cw_strerror.argtypes = [C.c_int]
    L.cw_strerror.restype = C.c_char_p
    L.cw_last_error.argtypes = [vp]
    L.cw_last_error.restype = C.c_char_p
    L.cw_abi_version.restype = C.c_int
    L.cw_last_launch_count.argtypes = [vp]
    L.cw_last_launch_count.restype = C.c_int64
    L.cw_last_kernel_ms.argtypes = [vp, C.c_int]
    L.cw_last_kernel_ms.restype = C.c_double
    L.cw_last_total_kernel_ms.argtypes = [vp, C.c_int]
    L.cw_last_total_kernel_ms.restype = C.c_double
    L.cw_measure_fp64_peak.argtypes = [vp, C.c_int, C.c_double]
    L.cw_measure_fp64_peak.restype = C.c_double
    L.cw_measure_scatter_write.argtypes = [vp, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int]
    L.cw_measure_scatter_write.restype = C.c_double
    L.cw_selftest_math.argtypes = [vp, C.c_int64, C.POINTER(C.c_double)]
    L.cw_selftest_math.restype = C.c_int
    _LIB = L
    return L


def ptr(a) -> Optional[int]:
    """Address of a numpy array / torch tensor / raw int pointer (None stays None)."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return a.data_ptr()
    raise TypeError(f"cannot take the address of {type(a)!r}")


class Context:
    """Owns a cw_ctx (streams + scratch on each device).  One call at a time per context."""

    def __init__(self, devices=None):
        self._L = load()
        self._h = C.c_void_p()
        if devices is None:
            rc = self._L.cw_create(C.byref(self._h), 0, None)
        else:
            ids = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self._L.cw_create(C.byref(self._h), len(devices), ids)
        if rc != 0:
            raise NativeLibraryError(f"cw_create failed: {self._L.cw_strerror(rc).decode()}")
        self.potential = None

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.cw_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed: {self._L.cw_strerror(rc).decode()} -- "
                               f"{self._L.cw_last_error(self._h).decode()}")

    @property
    def n_devices(self) -> int:
        return self._L.cw_device_count(self._h)

    # -- potential --------------------------------------------------------------------------
    def set_potential(self, pot):
        types = np.ascontiguousarray(pot.type_array(), dtype=np.int32)
        params = np.ascontiguousarray(pot.param_array(), dtype=np.float64)
        stride = params.shape[1] if params.ndim == 2 and params.shape[0] else 3
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        rot = pot.rotation_array() if hasattr(pot, "rotation_array") else None
        knots = pot.knot_arrays() if hasattr(pot, "knot_arrays") else None
        if rot is None and knots is None:
            rc = self._L.cw_set_potential_ex(self._h, float(pot.G), len(types), types.ctypes.data_as(ip),
                                             params.ctypes.data_as(dp), int(stride))
            self._check(rc, "cw_set_potential_ex")
        else:       # rotated components / gp.TimeInterpolatedPotential: the general kernels
            k_off, k_t, k_amp = knots if knots is not None else (None, None, None)
            rc = self._L.cw_set_potential_general(
                self._h, float(pot.G), len(types), types.ctypes.data_as(ip), params.ctypes.data_as(dp), int(stride),
                None if rot is None else rot.ctypes.data_as(dp),
                None if k_off is None else k_off.ctypes.data_as(ip),
                None if k_t is None else k_t.ctypes.data_as(dp),
                None if k_amp is None else k_amp.ctypes.data_as(dp),
                1 if getattr(pot, "interp", "cubic") == "cubic" else 0)
            self._check(rc, "cw_set_potential_general")
        self._check(self._L.cw_set_potential_time(self._h, float(getattr(pot, "t_eval", 0.0))), "cw_set_potential_time")
        self.potential = pot

    # -- raw entry points (host or device pointers) ------------------------------------------
    def _opts(self, integrator, mem_space, atol, rtol, nmax, stream, no_sync):
        return CwOpts(int(integrator), int(mem_space), float(atol), float(rtol), int(nmax),
                      stream, int(bool(no_sync)), 0)

    def integrate_raw(self, n, w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv, w_final, status,
                      n_nodes=None, ev_step=None, traj_off=None, traj_pos=None, traj_vel=None,
                      traj_t=None, stats=None, integrator=LEAPFROG, mem_space=MEM_HOST, atol=1e-10,
                      rtol=1e-10, nmax=0, stream=None, no_sync=False):
        o = self._opts(integrator, mem_space, atol, rtol, nmax, stream, no_sync)
        b = CwBatch(int(n), ptr(w0), ptr(t1), ptr(t2), ptr(dt), ptr(to_myr), ptr(flags), ptr(ev_off),
                    ptr(ev_time), ptr(ev_dv))
        r = CwResult(ptr(w_final), ptr(status), ptr(n_nodes), ptr(ev_step), ptr(traj_off), ptr(traj_pos),
                     ptr(traj_vel), ptr(traj_t), ptr(stats))
        self._check(self._L.cw_integrate(self._h, C.byref(o), C.byref(b), C.byref(r)), "cw_integrate")

    def count_nodes_raw(self, n, w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv, n_nodes,
                        ev_step=None, mem_space=MEM_HOST):
        o = self._opts(LEAPFROG, mem_space, 0, 0, 0, None, False)
        b = CwBatch(int(n), ptr(w0), ptr(t1), ptr(t2), ptr(dt), ptr(to_myr), ptr(flags), ptr(ev_off),
                    ptr(ev_time), ptr(ev_dv))
        self._check(self._L.cw_count_nodes(self._h, C.byref(o), C.byref(b), ptr(n_nodes), ptr(ev_step)),
                    "cw_count_nodes")

    # -- numpy convenience ---------------------------------------------------------------------
    @staticmethod
    def _prep(w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv):
        w0 = np.ascontiguousarray(w0, dtype=np.float64)
        if w0.ndim != 2 or w0.shape[0] != 6:
            raise ValueError("w0 must have shape (6, n_orbits)")
        n = w0.shape[1]
        f64 = lambda a: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (n,)))
        t1, t2, dt, to_myr = f64(t1), f64(t2), f64(dt), f64(to_myr)
        flags = np.ascontiguousarray(np.broadcast_to(np.asarray(flags, dtype=np.int32), (n,)))
        K = 0
        if ev_off is not None:
            ev_off = np.ascontiguousarray(ev_off, dtype=np.int64)
            if ev_off.shape != (n + 1,):
                raise ValueError("ev_off must have n_orbits + 1 entries")
            K = int(ev_off[-1])
        if K > 0:
            ev_time = np.ascontiguousarray(ev_time, dtype=np.float64)
            ev_dv = np.ascontiguousarray(ev_dv, dtype=np.float64).reshape(-1, 3)
            if ev_time.shape[0] != K or ev_dv.shape[0] != K:
                raise ValueError("ev_time / ev_dv must have ev_off[-1] rows")
        else:
            ev_off = ev_time = ev_dv = None
        return n, K, w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv

    def count_nodes(self, w0, t1, t2, dt, to_myr=1.0, flags=0, ev_off=None, ev_time=None, ev_dv=None):
        n, K, w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv = self._prep(
            w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv)
        n_nodes = np.zeros(n, dtype=np.int32)
        ev_step = np.full(max(K, 1), -1, dtype=np.int32)
        self.count_nodes_raw(n, w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv, n_nodes, ev_step)
        return n_nodes, ev_step[:K]

    def integrate(self, w0, t1, t2, dt, to_myr=1.0, flags=0, ev_off=None, ev_time=None, ev_dv=None,
                  integrator=LEAPFROG, store_all=False, atol=1e-10, rtol=1e-10, nmax=0, with_t=True):
        """Integrate a batch held in numpy arrays.  Returns a dict with w_final (6, n), status,
        n_nodes, ev_step, stats and -- when store_all -- traj_off, traj_pos, traj_vel, traj_t."""
        if self.potential is None:
            raise RuntimeError("set_potential() has not been called")
        n, K, w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv = self._prep(
            w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv)
        out = dict(w_final=np.empty((6, n)), status=np.zeros(n, dtype=np.int32),
                   n_nodes=np.zeros(n, dtype=np.int32), ev_step=np.full(max(K, 1), -1, dtype=np.int32),
                   stats=np.zeros(4, dtype=np.int64))
        traj_off = traj_pos = traj_vel = traj_t = None
        if store_all:
            nn, _ = self.count_nodes(w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv)
            traj_off = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(np.maximum(nn, 0), out=traj_off[1:])
            tot = int(traj_off[-1])
            traj_pos = np.full((3, tot), np.nan)
            traj_vel = np.full((3, tot), np.nan)
            traj_t = np.full(tot, np.nan) if with_t else None
        self.integrate_raw(n, w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv, out["w_final"],
                           out["status"], out["n_nodes"], out["ev_step"], traj_off, traj_pos, traj_vel,
                           traj_t, out["stats"], integrator=integrator, mem_space=MEM_HOST, atol=atol,
                           rtol=rtol, nmax=nmax)
        out["ev_step"] = out["ev_step"][:K]
        out.update(traj_off=traj_off, traj_pos=traj_pos, traj_vel=traj_vel, traj_t=traj_t)
        return out

    # -- pointwise potential -------------------------------------------------------------------
    def _pointwise(self, fn, q, n_out_rows):
        q = np.ascontiguousarray(q, dtype=np.float64)
        if q.ndim == 1:
            q = q.reshape(3, 1)
        n = q.shape[1]
        out = np.empty((n_out_rows, n)) if n_out_rows > 1 else np.empty(n)
        self._check(fn(self._h, MEM_HOST, n, ptr(q), ptr(out)), fn.__name__)
        return out

    def _pointwise_t(self, what, q, t, n_out_rows):
        q = np.ascontiguousarray(q, dtype=np.float64)
        if q.ndim == 1:
            q = q.reshape(3, 1)
        n = q.shape[1]
        t = np.ascontiguousarray(np.broadcast_to(np.asarray(t, dtype=np.float64), (n,)))
        out = np.empty((n_out_rows, n)) if n_out_rows > 1 else np.empty(n)
        self._check(self._L.cw_pointwise_t(self._h, what, MEM_HOST, n, ptr(q), ptr(t), ptr(out)), "cw_pointwise_t")
        return out

    def gradient(self, q, t=None):
        """grad Phi at q (3, n) [kpc]; t [Myr], scalar or per point, matters for time-dependent potentials only."""
        return self._pointwise(self._L.cw_gradient, q, 3) if t is None else self._pointwise_t(0, q, t, 3)

    def potential_value(self, q, t=None):
        return self._pointwise(self._L.cw_potential_value, q, 1) if t is None else self._pointwise_t(1, q, t, 1)

    def circular_velocity(self, q, t=None):
        """sqrt(r |dPhi/dr|) [kpc/Myr] -- `potential.circular_velocity(q, t=...)` of pop.py:1001-1004."""
        return self._pointwise(self._L.cw_circular_velocity, q, 1) if t is None else self._pointwise_t(2, q, t, 1)

    # -- diagnostics ----------------------------------------------------------------------------
    @property
    def last_launch_count(self) -> int:
        return int(self._L.cw_last_launch_count(self._h))

    def last_kernel_ms(self, dev=0) -> float:
        return float(self._L.cw_last_kernel_ms(self._h, dev))

    def last_total_kernel_ms(self, dev=0) -> float:
        return float(self._L.cw_last_total_kernel_ms(self._h, dev))

    def measure_fp64_peak(self, dev=0, ms_target=50.0) -> float:
        return float(self._L.cw_measure_fp64_peak(self._h, dev, ms_target))

    def measure_scatter_write(self, streams, stream_bytes, run_bytes, resident_threads_per_sm=2048, dev=0) -> float:
        return float(self._L.cw_measure_scatter_write(self._h, dev, streams, stream_bytes, run_bytes,
                                                      resident_threads_per_sm))

    def selftest_math(self, n=1 << 22):
        out = (C.c_double * 8)()
        self._check(self._L.cw_selftest_math(self._h, n, out), "cw_selftest_math")
        return dict(rsqrt=out[0], rcp=out[1], log=out[2], log_table=out[3], dfma_latency_cycles=out[4],
                    rsqrt_dadd_latency_cycles=out[5], dmul_latency_cycles=out[6])


_DEFAULT: Optional[Context] = None


def default_context() -> Context:
    """Process-wide context on every visible device (lazily created)."""
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = Context()
    return _DEFAULT
