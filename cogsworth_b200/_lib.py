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
           "cw_set_potential_time", "cw_count_nodes", "cw_host_alloc", "cw_host_alloc_cached", "cw_host_free", "cw_host_trim", "cw_host_is_pinned",
           "cw_host_rotate_kicks",
           "cw_integrate", "cw_gradient", "cw_potential_value", "cw_circular_velocity", "cw_pointwise_t", "cw_strerror",
           "cw_last_error", "cw_abi_version", "cw_last_launch_count", "cw_last_max_attempts", "cw_last_kernel_ms",
           "cw_last_total_kernel_ms", "cw_measure_fp64_peak", "cw_measure_scatter_write", "cw_selftest_math"]


class CwOpts(C.Structure):
    _fields_ = [("integrator", C.c_int32), ("mem_space", C.c_int32), ("atol", C.c_double),
                ("rtol", C.c_double), ("nmax", C.c_int64), ("stream", C.c_void_p),
                ("no_sync", C.c_int32), ("reserved", C.c_int32)]


class CwBatch(C.Structure):
    _fields_ = [("n_orbits", C.c_int64), ("w0", C.c_void_p), ("t1", C.c_void_p), ("t2", C.c_void_p),
                ("dt", C.c_void_p), ("to_myr", C.c_void_p), ("grid_flags", C.c_void_p),
                ("ev_off", C.c_void_p), ("ev_time", C.c_void_p), ("ev_dv", C.c_void_p),
                # ABI 2: compact host forms (include/cogsworth_b200.h)
                ("n_w0", C.c_int64), ("w0_index", C.c_void_p), ("grid_class", C.c_void_p),
                ("classes", C.c_void_p), ("n_classes", C.c_int32), ("reserved", C.c_int32),
                ("ev_off32", C.c_void_p)]


#: numpy mirror of ``cw_grid_class``
GRID_CLASS_DTYPE = np.dtype([("t2", "<f8"), ("dt", "<f8"), ("to_myr", "<f8"), ("grid_flags", "<i4"),
                             ("reserved", "<i4")])


class CwResult(C.Structure):
    _fields_ = [("w_final", C.c_void_p), ("status", C.c_void_p), ("n_nodes", C.c_void_p),
                ("ev_step", C.c_void_p), ("traj_off", C.c_void_p), ("traj_pos", C.c_void_p),
                ("traj_vel", C.c_void_p), ("traj_t", C.c_void_p), ("stats", C.c_void_p), ("n_failed", C.c_void_p)]


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
    L.cw_host_alloc.argtypes = [C.c_size_t]
    L.cw_host_alloc.restype = vp
    L.cw_host_alloc_cached.argtypes = [C.c_size_t]
    L.cw_host_alloc_cached.restype = vp
    L.cw_host_trim.argtypes = []
    L.cw_host_trim.restype = None
    L.cw_host_free.argtypes = [vp]
    L.cw_host_free.restype = None
    L.cw_host_is_pinned.argtypes = [vp]
    L.cw_host_is_pinned.restype = C.c_int
    L.cw_host_rotate_kicks.argtypes = [C.c_int64, vp, vp, vp, C.c_double, vp]
    L.cw_host_rotate_kicks.restype = None
    L.cw_strerror.argtypes = [C.c_int]
    L.cw_strerror.restype = C.c_char_p
    L.cw_last_error.argtypes = [vp]
    L.cw_last_error.restype = C.c_char_p
    L.cw_abi_version.restype = C.c_int
    L.cw_last_launch_count.argtypes = [vp]
    L.cw_last_launch_count.restype = C.c_int64
    L.cw_last_max_attempts.argtypes = [vp]
    L.cw_last_max_attempts.restype = C.c_int64
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


class _PinnedBlock:
    """A page-locked block from the library's cache; returns to the cache when the last array viewing it dies."""

    def __init__(self, nbytes, cached_only=False):
        self._L = load()
        self.nbytes = int(nbytes)
        self.addr = (self._L.cw_host_alloc_cached if cached_only else self._L.cw_host_alloc)(max(self.nbytes, 1))
        if not self.addr:
            raise MemoryError("cw_host_alloc failed")
        self.__array_interface__ = {"shape": (max(self.nbytes, 1),), "typestr": "|u1", "data": (self.addr, False),
                                    "version": 3}

    def __del__(self):
        try:
            if getattr(self, "addr", None):
                self._L.cw_host_free(self.addr)
                self.addr = None
        except Exception:
            pass


#: arrays larger than this are not taken from page-locked memory (host RAM that cannot be swapped is a shared resource)
PINNED_LIMIT_BYTES = int(os.environ.get("CW_PINNED_LIMIT_MB", "16384")) << 20
_PINNED_OK = None
_PINNED_POLICY = "pin"


class pinned_policy:
    """``with pinned_policy("cached"):`` -- inside, :func:`pinned_empty` takes page-locked memory only when the library's
    cache already holds a block that fits and returns ordinary memory otherwise (the library stages that through its
    own ring).  Pinning costs 0.45 ms per MB, a 10^7-orbit batch and its results are 1.4 GB: a caller that runs the
    stage once would pay 0.6 s to save 3 ms.  The default policy ``"pin"`` is for callers in a loop."""

    def __init__(self, policy):
        if policy not in ("pin", "cached"):
            raise ValueError("policy must be 'pin' or 'cached'")
        self.policy = policy

    def __enter__(self):
        global _PINNED_POLICY
        self.previous, _PINNED_POLICY = _PINNED_POLICY, self.policy
        return self

    def __exit__(self, *exc):
        global _PINNED_POLICY
        _PINNED_POLICY = self.previous
        return False


def pinned_empty(shape, dtype=np.float64):
    """``np.empty`` in page-locked host memory (DMA-ed by ``cw_integrate`` without staging); plain ``np.empty`` when
    the library or a CUDA device is missing, the array is larger than ``PINNED_LIMIT_BYTES``, or the policy in force
    (:class:`pinned_policy`) is ``"cached"`` and no cached block fits."""
    global _PINNED_OK
    dtype = np.dtype(dtype)
    shape = (int(shape),) if np.isscalar(shape) else tuple(int(x) for x in shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    if _PINNED_OK is False or nbytes == 0 or nbytes > PINNED_LIMIT_BYTES:
        return np.empty(shape, dtype=dtype)
    cached_only = _PINNED_POLICY == "cached"
    try:
        blk = _PinnedBlock(nbytes, cached_only=cached_only)
    except MemoryError:
        if not cached_only:
            _PINNED_OK = False
        return np.empty(shape, dtype=dtype)
    except (NativeLibraryError, OSError):
        _PINNED_OK = False
        return np.empty(shape, dtype=dtype)
    _PINNED_OK = True
    return np.asarray(blk)[:nbytes].view(dtype).reshape(shape)


def pinned_trim():
    """Release every cached (free) page-locked block back to the driver."""
    load().cw_host_trim()


def is_pinned(a) -> bool:
    """True if the array's memory is page-locked (or device) memory known to the CUDA runtime."""
    return bool(load().cw_host_is_pinned(ptr(a)))


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

    @staticmethod
    def _batch(n, w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv, n_w0=0, w0_index=None, grid_class=None,
               classes=None, ev_off32=None):
        return CwBatch(int(n), ptr(w0), ptr(t1), ptr(t2), ptr(dt), ptr(to_myr), ptr(flags), ptr(ev_off),
                       ptr(ev_time), ptr(ev_dv), int(n_w0), ptr(w0_index), ptr(grid_class), ptr(classes),
                       0 if classes is None else len(classes), 0, ptr(ev_off32))

    def integrate_raw(self, n, w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv, w_final, status,
                      n_nodes=None, ev_step=None, traj_off=None, traj_pos=None, traj_vel=None,
                      traj_t=None, stats=None, integrator=LEAPFROG, mem_space=MEM_HOST, atol=1e-10,
                      rtol=1e-10, nmax=0, stream=None, no_sync=False, n_w0=0, w0_index=None, grid_class=None,
                      classes=None, ev_off32=None, n_failed=None):
        """``cw_integrate`` on raw buffers (numpy / torch / int addresses).  The keyword arguments after ``no_sync``
        are the compact host forms of ABI 2 (``cw_batch`` in include/cogsworth_b200.h)."""
        o = self._opts(integrator, mem_space, atol, rtol, nmax, stream, no_sync)
        b = self._batch(n, w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv, n_w0, w0_index, grid_class,
                        classes, ev_off32)
        r = CwResult(ptr(w_final), ptr(status), ptr(n_nodes), ptr(ev_step), ptr(traj_off), ptr(traj_pos),
                     ptr(traj_vel), ptr(traj_t), ptr(stats), ptr(n_failed))
        self._check(self._L.cw_integrate(self._h, C.byref(o), C.byref(b), C.byref(r)), "cw_integrate")

    def count_nodes_raw(self, n, w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv, n_nodes,
                        ev_step=None, mem_space=MEM_HOST, grid_class=None, classes=None, ev_off32=None):
        o = self._opts(LEAPFROG, mem_space, 0, 0, 0, None, False)
        b = self._batch(n, w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv, 0, None, grid_class, classes,
                        ev_off32)
        self._check(self._L.cw_count_nodes(self._h, C.byref(o), C.byref(b), ptr(n_nodes), ptr(ev_step)),
                    "cw_count_nodes")

    # -- numpy convenience ---------------------------------------------------------------------
    @staticmethod
    def _prep(w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv):
        if not (isinstance(w0, np.ndarray) and w0.strides[-1] == 0):      # (a broadcast placeholder stays one)
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
        """Node count of every orbit and the node index of every event (the grid pre-pass alone); ``w0`` is not
        read and may be ``None``."""
        no_w0 = w0 is None
        if no_w0:
            w0 = np.broadcast_to(np.zeros((6, 1)), (6, np.asarray(t1).shape[0]))
        n, K, w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv = self._prep(
            w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv)
        if no_w0:
            w0 = None
        n_nodes = np.zeros(n, dtype=np.int32)
        ev_step = np.full(max(K, 1), -1, dtype=np.int32)
        self.count_nodes_raw(n, w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv, n_nodes, ev_step)
        return n_nodes, ev_step[:K]

    def integrate(self, w0, t1, t2, dt, to_myr=1.0, flags=0, ev_off=None, ev_time=None, ev_dv=None,
                  integrator=LEAPFROG, store_all=False, atol=1e-10, rtol=1e-10, nmax=0, with_t=True):
        """Integrate a batch held in numpy arrays.  Returns a dict with w_final (6, n), status,
        n_nodes, ev_step, stats and -- when store_all -- traj_off, traj_pos, traj_vel, traj_t."""
        n, K, w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv = self._prep(
            w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv)
        return self._integrate(n, K, dict(w0=w0, t1=t1, t2=t2, dt=dt, to_myr=to_myr, flags=flags, ev_off=ev_off,
                                          ev_time=ev_time, ev_dv=ev_dv), integrator, store_all, atol, rtol, nmax, with_t)

    def integrate_batch(self, batch, integrator=LEAPFROG, store_all=False, atol=1e-10, rtol=1e-10, nmax=0,
                        with_t=True):
        """The same for an :class:`~cogsworth_b200.batch.OrbitBatch`, in its compact form when it has one (shared
        initial conditions of disrupted secondaries, one-byte grid classes, 32-bit event offsets)."""
        a = batch.abi_arrays()
        if store_all and a.get("w0_index") is not None:
            a["t1_full"] = batch.t1          # the node-count pass takes one start time per orbit
        return self._integrate(batch.n_orbits, batch.n_events, a, integrator, store_all, atol, rtol, nmax, with_t)

    def _integrate(self, n, K, a, integrator, store_all, atol, rtol, nmax, with_t):
        if self.potential is None:
            raise RuntimeError("set_potential() has not been called")
        get = a.get
        # results live in page-locked memory from the library's cache: the D2H copies land in them directly
        out = dict(w_final=pinned_empty((6, n)), status=pinned_empty(n, np.int32), n_nodes=pinned_empty(n, np.int32),
                   ev_step=pinned_empty(max(K, 1), np.int32), stats=np.zeros(4, dtype=np.int64),
                   n_failed=np.zeros(1, dtype=np.int64))
        compact = dict(n_w0=get("n_w0", 0), w0_index=get("w0_index"), grid_class=get("grid_class"),
                       classes=get("classes"), ev_off32=get("ev_off32"))
        traj_off = traj_pos = traj_vel = traj_t = None
        if store_all:
            # node counts first (the grid pre-pass alone: start times and grid columns, no phase space, no events)
            nn = pinned_empty(n, np.int32)
            t1_full = get("t1_full", get("t1"))
            self.count_nodes_raw(n, None, t1_full, get("t2"), get("dt"), get("to_myr"), get("flags"), None, None, None,
                                 nn, None, grid_class=compact["grid_class"], classes=compact["classes"])
            traj_off = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(np.maximum(nn, 0), out=traj_off[1:])
            tot = int(traj_off[-1])
            traj_pos, traj_vel = pinned_empty((3, tot)), pinned_empty((3, tot))
            traj_t = pinned_empty(tot) if with_t else None
        self.integrate_raw(n, get("w0"), get("t1"), get("t2"), get("dt"), get("to_myr"), get("flags"), get("ev_off"),
                           get("ev_time"), get("ev_dv"), out["w_final"], out["status"], out["n_nodes"], out["ev_step"],
                           traj_off, traj_pos, traj_vel, traj_t, out["stats"], integrator=integrator,
                           mem_space=MEM_HOST, atol=atol, rtol=rtol, nmax=nmax, n_failed=out["n_failed"], **compact)
        out["n_failed"] = int(out["n_failed"][0])
        if K == 0:
            out["ev_step"][:] = -1
        out["ev_step"] = out["ev_step"][:K]
        if store_all:
            # the kernels write nothing for orbits the pre-pass rejected, and stop writing where an orbit fails:
            # those columns (few) are marked, not the whole buffer
            bad = np.flatnonzero(out["status"] != 0) if out["n_failed"] else ()
            for i in bad:
                sl = slice(int(traj_off[i]), int(traj_off[i + 1]))
                traj_pos[:, sl] = np.nan
                traj_vel[:, sl] = np.nan
                if traj_t is not None:
                    traj_t[sl] = np.nan
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

    @property
    def last_max_attempts(self) -> int:
        return int(self._L.cw_last_max_attempts(self._h))

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
    """Process-wide context (lazily created): every visible device -- orbits are sharded by index, one host thread per
    device -- unless this process is one rank of a one-process-per-GPU job (``LOCAL_RANK`` set, or a
    ``torch.distributed`` group initialised), in which case it owns its rank's device only.  A ``Population`` can be
    given its own through the ``_cw_context`` attribute."""
    global _DEFAULT
    if _DEFAULT is None:
        devices = None
        lr = os.environ.get("LOCAL_RANK")
        if lr is None:
            try:
                import sys
                td = sys.modules.get("torch.distributed")
                if td is not None and td.is_available() and td.is_initialized():
                    import torch
                    lr = td.get_rank() % max(torch.cuda.device_count(), 1)
            except Exception:
                lr = None
        if lr is not None:
            devices = [int(lr)]
        _DEFAULT = Context(devices=devices)
    return _DEFAULT
