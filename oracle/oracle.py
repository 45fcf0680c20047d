"""ctypes front-end of oracle/libcw_oracle.so (see cw_oracle.c for what it restates).

TEST INFRASTRUCTURE ONLY -- never imported by cogsworth_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MAX_COMP = 16
LEAPFROG, DOP853, DOP853_DENSE = 0, 1, 2
FLAG_APPEND_T2 = 1

__all__ = ["build", "lib", "make_potential", "gradient", "gradient_t", "value", "circular_velocity", "time_grid",
           "leapfrog", "dop853", "dop853_dense", "integrate_orbit", "integrate_batch", "max_threads", "set_rounding_noise",
           "LEAPFROG", "DOP853", "DOP853_DENSE", "FLAG_APPEND_T2"]


MAX_KNOTS = 32


class _Pot(C.Structure):
    _fields_ = [("G", C.c_double), ("n", C.c_int32), ("type", C.c_int32 * MAX_COMP),
                ("p", (C.c_double * 6) * MAX_COMP),
                ("has_R", C.c_int32 * MAX_COMP), ("R", (C.c_double * 9) * MAX_COMP),
                ("n_knots", C.c_int32 * MAX_COMP), ("interp", C.c_int32),
                ("knot_t", (C.c_double * MAX_KNOTS) * MAX_COMP), ("knot_y", (C.c_double * MAX_KNOTS) * MAX_COMP),
                ("knot_c", (C.c_double * MAX_KNOTS) * MAX_COMP), ("t_eval", C.c_double)]


class _Stats(C.Structure):
    _fields_ = [("n_accepted", C.c_int64), ("n_rejected", C.c_int64), ("n_fcn", C.c_int64)]


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed Makefile; returns the .so path."""
    so = os.path.join(_HERE, "libcw_oracle.so")
    src = os.path.join(_HERE, "cw_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libcw_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libcw_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.cwo_gradient.argtypes = [C.POINTER(_Pot), dp, dp]
        L.cwo_gradient_t.argtypes = [C.POINTER(_Pot), C.c_double, dp, dp]
        L.cwo_gradient_t.restype = None
        L.cwo_prepare.argtypes = [C.POINTER(_Pot)]
        L.cwo_prepare.restype = None
        L.cwo_value.argtypes = [C.POINTER(_Pot), dp]; L.cwo_value.restype = C.c_double
        L.cwo_circular_velocity.argtypes = [C.POINTER(_Pot), dp]; L.cwo_circular_velocity.restype = C.c_double
        L.cwo_time_grid.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, dp, C.c_int64]
        L.cwo_time_grid.restype = C.c_int64
        L.cwo_leapfrog.argtypes = [C.POINTER(_Pot), dp, dp, C.c_int64, dp, C.c_int64]
        L.cwo_leapfrog.restype = None
        L.cwo_dop853.argtypes = [C.POINTER(_Pot), C.c_double, dp, C.c_double, C.c_double, C.c_double,
                                 C.c_double, C.c_int64, C.POINTER(_Stats)]
        L.cwo_dop853.restype = C.c_int
        L.cwo_dop853_dense.argtypes = [C.POINTER(_Pot), C.c_double, dp, C.c_double, C.c_double, C.c_double,
                                       C.c_double, C.c_int64, C.POINTER(_Stats), dp, C.c_int64, dp, C.c_int64]
        L.cwo_dop853_dense.restype = C.c_int
        L.cwo_dop853_segment.argtypes = [C.POINTER(_Pot), dp, dp, C.c_int64, C.c_double, C.c_double,
                                         C.c_int64, dp, C.c_int64, C.POINTER(_Stats)]
        L.cwo_dop853_segment.restype = C.c_int
        L.cwo_dop853_dense_segment.argtypes = L.cwo_dop853_segment.argtypes
        L.cwo_dop853_dense_segment.restype = C.c_int
        L.cwo_integrate_orbit.argtypes = [
            C.POINTER(_Pot), C.c_int, C.c_double, C.c_double, C.c_int64,
            dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
            C.c_int64, dp, dp, C.c_int, dp, C.c_int64, dp, C.c_int64, dp, ip, lp, C.POINTER(_Stats)]
        L.cwo_integrate_orbit.restype = C.c_int
        L.cwo_integrate_batch.argtypes = [
            C.POINTER(_Pot), C.c_int, C.c_double, C.c_double, C.c_int64, C.c_int64,
            dp, dp, dp, dp, dp, ip, lp, dp, dp, dp, ip, ip, ip, lp, dp, dp, dp, lp, C.c_int]
        L.cwo_integrate_batch.restype = C.c_int
        L.cwo_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _lp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int64))


def make_potential(pot) -> _Pot:
    """pot: anything with .G, .types, .params (cogsworth_b200.potential.CompositePotential)."""
    P = _Pot()
    P.G = float(pot.G)
    P.n = len(pot.types)
    assert P.n <= MAX_COMP
    for i, (t, p) in enumerate(zip(pot.types, pot.params)):
        P.type[i] = int(t)
        for k in range(6):
            P.p[i][k] = float(p[k]) if k < len(p) else 0.0
    rotations = getattr(pot, "rotations", None) or [None] * P.n
    for i, Rm in enumerate(rotations):
        if Rm is not None:
            P.has_R[i] = 1
            flat = np.asarray(Rm, dtype=np.float64).reshape(9)
            for k in range(9):
                P.R[i][k] = float(flat[k])
    knots = getattr(pot, "knots", None) or [None] * P.n
    P.interp = 1 if getattr(pot, "interp", "cubic") == "cubic" else 0
    for i, kn in enumerate(knots):
        if kn is not None:
            t, y = np.asarray(kn[0], dtype=np.float64), np.asarray(kn[1], dtype=np.float64)
            assert len(t) == len(y) <= MAX_KNOTS
            P.n_knots[i] = len(t)
            for k in range(len(t)):
                P.knot_t[i][k] = float(t[k]); P.knot_y[i][k] = float(y[k])
    P.t_eval = float(getattr(pot, "t_eval", 0.0))
    lib().cwo_prepare(C.byref(P))
    return P


def _as_pot(pot):
    return pot if isinstance(pot, _Pot) else make_potential(pot)


def gradient(pot, q):
    P = _as_pot(pot)
    q = np.ascontiguousarray(q, dtype=np.float64)
    g = np.zeros(3)
    lib().cwo_gradient(C.byref(P), _dp(q), _dp(g))
    return g


def gradient_t(pot, t, q):
    P = _as_pot(pot)
    q = np.ascontiguousarray(q, dtype=np.float64)
    g = np.zeros(3)
    lib().cwo_gradient_t(C.byref(P), float(t), _dp(q), _dp(g))
    return g


def value(pot, q):
    P = _as_pot(pot)
    q = np.ascontiguousarray(q, dtype=np.float64)
    return lib().cwo_value(C.byref(P), _dp(q))


def circular_velocity(pot, q):
    P = _as_pot(pot)
    q = np.ascontiguousarray(q, dtype=np.float64)
    return lib().cwo_circular_velocity(C.byref(P), _dp(q))


def time_grid(t1, t2, dt, append_t2):
    n = lib().cwo_time_grid(t1, t2, dt, int(append_t2), None, 0)
    T = np.empty(max(n, 0))
    lib().cwo_time_grid(t1, t2, dt, int(append_t2), _dp(T), n)
    return T


def leapfrog(pot, w0, t_myr, store_all=True):
    P = _as_pot(pot)
    w = np.array(w0, dtype=np.float64)
    t = np.ascontiguousarray(t_myr, dtype=np.float64)
    out = np.empty((6, len(t))) if store_all else None
    lib().cwo_leapfrog(C.byref(P), _dp(w), _dp(t), len(t), _dp(out), len(t))
    return (out if store_all else w)


def dop853(pot, x, y, xend, rtol=1e-10, atol=1e-10, h=None, nmax=0):
    """One Hairer dop853() call; returns (res, y_end, (accepted, rejected, nfcn))."""
    P = _as_pot(pot)
    y = np.array(y, dtype=np.float64)
    st = _Stats()
    res = lib().cwo_dop853(C.byref(P), x, _dp(y), xend, rtol, atol, (xend - x) if h is None else h,
                           nmax, C.byref(st))
    return res, y, (st.n_accepted, st.n_rejected, st.n_fcn)


def dop853_dense(pot, x, y, xend, tout, rtol=1e-10, atol=1e-10, h=None, nmax=0):
    """One dop853() call with dense output (contd8) at ``tout`` (inside (x, xend), ascending).
    Returns (res, y_end, out (6, len(tout)), (accepted, rejected, nfcn))."""
    P = _as_pot(pot)
    y = np.array(y, dtype=np.float64)
    tout = np.ascontiguousarray(tout, dtype=np.float64)
    out = np.full((6, max(len(tout), 1)), np.nan)
    st = _Stats()
    res = lib().cwo_dop853_dense(C.byref(P), x, _dp(y), xend, rtol, atol, (xend - x) if h is None else h,
                                 nmax, C.byref(st), _dp(tout), len(tout), _dp(out), out.shape[1])
    return res, y, out[:, :len(tout)], (st.n_accepted, st.n_rejected, st.n_fcn)


def integrate_orbit(pot, w0, t1, t2, dt, to_myr=1.0, flags=0, ev_time=None, ev_dv=None,
                    integrator=LEAPFROG, rtol=1e-10, atol=1e-10, nmax=0, store_all=False):
    """One orbit through the kicks.py driver.  Returns dict(status, w_final, ev_step, n_nodes,
    traj (6, n) or None, t (n,) Myr or None, stats)."""
    P = _as_pot(pot)
    w0 = np.ascontiguousarray(w0, dtype=np.float64)
    n_ev = 0 if ev_time is None else len(ev_time)
    ev_time_a = np.ascontiguousarray(ev_time, dtype=np.float64) if n_ev else None
    ev_dv_a = np.ascontiguousarray(ev_dv, dtype=np.float64).reshape(-1, 3) if n_ev else None
    n = lib().cwo_time_grid(t1, t2, dt, flags & 1, None, 0)
    traj = np.full((6, n + 1), np.nan) if store_all else None
    tt = np.empty(n) if store_all else None
    wf = np.empty(6)
    ev_step = np.full(max(n_ev, 1), -1, dtype=np.int32)
    nn = C.c_int64(0)
    st = _Stats()
    s = lib().cwo_integrate_orbit(C.byref(P), integrator, rtol, atol, nmax, _dp(w0), t1, t2, dt, to_myr,
                                  flags, n_ev, _dp(ev_time_a), _dp(ev_dv_a), int(store_all), _dp(traj),
                                  n + 1, _dp(tt), n, _dp(wf), _ip(ev_step), C.byref(nn), C.byref(st))
    return dict(status=s, w_final=wf, ev_step=ev_step[:n_ev], n_nodes=nn.value,
                traj=None if traj is None else traj[:, :n], t=tt,
                stats=(st.n_accepted, st.n_rejected, st.n_fcn))


def integrate_batch(pot, w0, t1, t2, dt, to_myr, flags, ev_off=None, ev_time=None, ev_dv=None,
                    integrator=LEAPFROG, rtol=1e-10, atol=1e-10, nmax=0, traj_off=None, n_threads=0):
    """Batch over orbits (OpenMP).  w0: (6, n).  Returns dict like the product's cw_integrate."""
    P = _as_pot(pot)
    w0 = np.ascontiguousarray(w0, dtype=np.float64)
    n = w0.shape[1]
    f64 = lambda a: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (n,)))
    t1, t2, dt, to_myr = f64(t1), f64(t2), f64(dt), f64(to_myr)
    flags = np.ascontiguousarray(np.broadcast_to(np.asarray(flags, dtype=np.int32), (n,)))
    if ev_off is None:
        ev_off = np.zeros(n + 1, dtype=np.int64)
    ev_off = np.ascontiguousarray(ev_off, dtype=np.int64)
    K = int(ev_off[-1])
    ev_time = np.ascontiguousarray(ev_time, dtype=np.float64) if K else np.zeros(1)
    ev_dv = np.ascontiguousarray(ev_dv, dtype=np.float64).reshape(-1, 3) if K else np.zeros((1, 3))
    wf = np.empty((6, n))
    status = np.empty(n, dtype=np.int32)
    n_nodes = np.empty(n, dtype=np.int32)
    ev_step = np.full(max(K, 1), -1, dtype=np.int32)
    stats = np.zeros(3, dtype=np.int64)
    tp = tv = tt = None
    if traj_off is not None:
        traj_off = np.ascontiguousarray(traj_off, dtype=np.int64)
        tot = int(traj_off[-1])
        tp, tv, tt = np.full((3, tot), np.nan), np.full((3, tot), np.nan), np.full(tot, np.nan)
    lib().cwo_integrate_batch(C.byref(P), integrator, rtol, atol, nmax, n, _dp(w0), _dp(t1), _dp(t2),
                              _dp(dt), _dp(to_myr), _ip(flags), _lp(ev_off), _dp(ev_time), _dp(ev_dv),
                              _dp(wf), _ip(status), _ip(n_nodes), _ip(ev_step), _lp(traj_off),
                              _dp(tp), _dp(tv), _dp(tt), _lp(stats), int(n_threads))
    return dict(w_final=wf, status=status, n_nodes=n_nodes, ev_step=ev_step[:K], traj_pos=tp,
                traj_vel=tv, traj_t=tt, stats=tuple(int(s) for s in stats))


def set_rounding_noise(eps: float = 0.0, seed: int = 0) -> None:
    """Multiply every gradient component by 1 + eps * {-1, 0, +1} (hash of the position and ``seed``): the last-bit
    differences any equivalent implementation has.  ``eps = 0`` (default) switches it off.  Process-wide."""
    L = lib()
    L.cwo_set_rounding_noise.argtypes = [C.c_double, C.c_uint64]
    L.cwo_set_rounding_noise.restype = None
    L.cwo_set_rounding_noise(float(eps), int(seed))


def max_threads() -> int:
    return int(lib().cwo_max_threads())
