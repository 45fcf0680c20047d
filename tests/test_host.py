"""CPU tests of the host-side mirror of the reference interface (events / kicks / pop / batch /
orbits) and of the C-ABI library's export surface (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest

import cogsworth_b200 as cb
from cogsworth_b200 import constants as K
from cogsworth_b200 import _lib
from cogsworth_b200.batch import EventTable, build_orbit_batch, rotate_kicks
from cogsworth_b200.events import event_table_from_frame, identify_event_tables, identify_events
from cogsworth_b200.kicks import get_kick_differential, resolve_integrator
from cogsworth_b200.orbits import OrbitArray

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fake_population():
    """The hand-built tables of the reference's tests/test_events.py:17-50 (four binaries: no event;
    bound + kick; disrupted at SN1; disrupted at SN2)."""
    bpp = pd.DataFrame({
        "mass_1": [1, 10, 20, 1.4, 20, 20, 8], "mass_2": [1, 10, 20, 20, 20, 20, 1.4],
        "tphys": [0, 0, 0, 1, 0, 1, 2], "sep": [10, 10, 10, -1.0, 10, 10, -1.0],
        "ecc": [0, 0, 0, -1.0, 0, 0.1, -1.0], "evol_type": [-1, 15, 15, 11, 15, 16, 11],
        "bin_num": [0, 1, 2, 2, 3, 3, 3]})
    bpp.index = bpp["bin_num"].values
    kick_info = pd.DataFrame({
        "star": [0, 1, 1, 1, 2], "disrupted": [0, 0, 1, 0, 1],
        "delta_vsysx_1": [0, 1., 2., 3., 4.], "delta_vsysy_1": [0, 0, 0, 0, 0], "delta_vsysz_1": [0, 0, 0, 0, 0],
        "delta_vsysx_2": [0, 10., 20., 30., 40.], "delta_vsysy_2": [0, 0, 0, 0, 0], "delta_vsysz_2": [0, 0, 0, 0, 0],
        "bin_num": [0, 1, 2, 3, 3]})
    kick_info.index = kick_info["bin_num"].values
    initC = pd.DataFrame({"metallicity": [1e-2] * 4, "bin_num": [0, 1, 2, 3]})
    initC.index = initC["bin_num"].values
    ig = cb.InitialGalaxy(tau=[1.0, 2.0, 3.0, 4.0], x=[8, 7, 6, 5.], y=[0, 1, 2, 3.], z=[0, 0.1, 0.2, 0.3],
                          v_x=[0, 10, 20, 30.], v_y=[220, 210, 200, 190.], v_z=[0, 1, 2, 3.])
    return cb.B200Population(ig, bpp, kick_info, initC, store_entire_orbits=False, integrator="leapfrog")


def test_identify_events_reference_fixture():
    """Same assertions as the reference's tests/test_events.py:52-69."""
    np.random.seed(3)
    p = fake_population()
    primary, secondary = identify_events(p)
    assert len(primary) == 4
    assert len(secondary) == 3
    assert 0 not in primary.index and 0 not in secondary.index
    assert 1 not in secondary.index
    assert len(secondary.loc[[3]]) > len(secondary.loc[[2]])
    assert list(primary.columns) == ["tphys", "delta_vsys_x", "delta_vsys_y", "delta_vsys_z", "inc", "phase"]
    # primaries use delta_vsys*_1 at every SN; secondaries *_2 only at the disrupting SN (events.py:47-67)
    assert list(primary["delta_vsys_x"]) == [1., 2., 3., 4.]
    assert list(secondary.loc[[2]]["delta_vsys_x"]) == [20.]
    assert list(secondary.loc[[3]]["delta_vsys_x"]) == [3., 40.]
    # angles were drawn once and persisted (pop.py:1229-1234)
    for col in ("phase_sn_1", "phase_sn_2", "inc_sn_1", "inc_sn_2"):
        assert col in p.initial_binaries
    again_p, again_s = identify_events(p)
    assert again_p.equals(primary) and again_s.equals(secondary)
    # star 1 rows carry the *_sn_1 angles, star 2 rows the *_sn_2 ones
    assert secondary.loc[[3]]["phase"].iloc[1] == p.initial_binaries.loc[3, "phase_sn_2"]
    assert primary.loc[[1]]["inc"].iloc[0] == p.initial_binaries.loc[1, "inc_sn_1"]


def test_disrupted_mask_and_ordering():
    p = fake_population()
    assert list(p.disrupted) == [False, False, True, True]
    prim, sec = identify_event_tables(p)
    assert list(prim.owner) == [1, 2, 3, 3] and list(sec.owner) == [2, 3, 3]
    ig = p.initial_galaxy
    batch = build_orbit_batch(np.stack([ig.x, ig.y, ig.z]), np.stack([ig.v_x, ig.v_y, ig.v_z]), ig.tau, 12.0, 1.0,
                              p.disrupted, prim, sec)
    # Q7: primaries in bin_num order, then disrupted secondaries
    assert batch.n_orbits == 6 and batch.n_primary == 4
    assert list(batch.secondary_of) == [2, 3]
    assert np.array_equal(batch.w0[:, 4], batch.w0[:, 2]) and np.array_equal(batch.w0[:, 5], batch.w0[:, 3])
    assert list(batch.ev_off) == [0, 0, 1, 2, 4, 5, 7]
    # un-kicked orbit 0: gala's Myr grid, no appended node; kicked: seconds grid + append
    assert batch.flags[0] == 0 and batch.to_myr[0] == 1.0 and batch.t1[0] == 11000.0 and batch.dt[0] == 1.0
    assert batch.flags[1] == 1 and batch.to_myr[1] == K.MYR_PER_SEC and batch.dt[1] == K.SEC_PER_MYR
    assert batch.t1[1] == 10.0 * K.SEC_PER_GYR and batch.t2[1] == 12.0 * K.SEC_PER_GYR
    # event time: (t1_Gyr + tphys*1e-3) * s/Gyr      (kicks.py:134 in astropy's order)
    assert batch.ev_time[3] == (8.0 + 1 * K.GYR_PER_MYR) * K.SEC_PER_GYR and batch.ev_time[2] == 8.0 * K.SEC_PER_GYR
    assert np.allclose(batch.w0[3:, 1], np.array([10, 210, 1.]) / K.KMS_PER_KPC_MYR, rtol=1e-15)


def test_constants_are_astropy_floats():
    assert K.SEC_PER_GYR == 1e9 * 365.25 * 86400 and K.SEC_PER_MYR == 1e6 * 365.25 * 86400
    assert K.GYR_PER_MYR == 0.001 and K.MYR_PER_GYR == 1000.0
    assert abs(K.KMS_PER_KPC_MYR - 3.0856775814913673e16 / 3.15576e13) < 1e-12
    # G = 6.6743e-11 m^3/kg/s^2 in kpc^3/Msun/Myr^2
    G = 6.6743e-11 * 1.988409870698051e30 * (3.15576e13) ** 2 / (3.0856775814913673e19) ** 3
    assert abs(G / K.G_GALACTIC - 1) < 1e-14


def test_kick_rotation_matches_reference_formula():
    rng = np.random.default_rng(0)
    dv = rng.normal(0, 100, (50, 3))
    ph, inc = rng.uniform(0, 2 * np.pi, 50), np.arccos(2 * rng.random(50) - 1)
    rot = rotate_kicks(dv, ph, inc)
    for k in range(50):
        one = get_kick_differential(dv[k, 0], dv[k, 1], dv[k, 2], phase=ph[k], inclination=inc[k])
        assert np.array_equal(one, rot[k])
    # it is a rotation: speed is preserved
    assert np.allclose(np.linalg.norm(rot, axis=1), np.linalg.norm(dv, axis=1), rtol=1e-13)
    # phase = inc = 0 is the identity
    assert np.allclose(get_kick_differential(1., 2., 3., 0.0, 0.0), [1, 2, 3])
    # random angles come from the global numpy RNG like the reference
    np.random.seed(1)
    a = get_kick_differential(1., 2., 3.)
    np.random.seed(1)
    assert np.array_equal(a, get_kick_differential(1., 2., 3.))


def test_resolve_integrator():
    class LeapfrogIntegrator: pass
    class DOPRI853Integrator: pass
    class Ruth4Integrator: pass
    assert resolve_integrator(LeapfrogIntegrator) == _lib.LEAPFROG
    assert resolve_integrator(DOPRI853Integrator) == _lib.DOP853
    assert resolve_integrator(None) == _lib.DOP853          # cogsworth default (pop.py:130)
    assert resolve_integrator("leapfrog") == _lib.LEAPFROG
    # the dense-output form of DOPRI853 is opt-in per call or per process
    assert resolve_integrator("dop853_dense") == _lib.DOP853_DENSE
    assert resolve_integrator(DOPRI853Integrator, {"dense_output": True}) == _lib.DOP853_DENSE
    assert resolve_integrator(LeapfrogIntegrator, {"dense_output": True}) == _lib.LEAPFROG
    import cogsworth_b200.kicks as kk
    kk.DOP853_MODE = "dense"
    try:
        assert resolve_integrator(None) == _lib.DOP853_DENSE
        assert resolve_integrator(DOPRI853Integrator, {"dense_output": False}) == _lib.DOP853
    finally:
        kk.DOP853_MODE = "restart"
    with pytest.raises(NotImplementedError):
        resolve_integrator(Ruth4Integrator)


def test_orbit_array_container():
    off = np.array([0, 3, 4, 6])
    pos = np.arange(18, dtype=float).reshape(3, 6)
    vel = pos / 100.0
    t = np.arange(6, dtype=float)
    oa = OrbitArray(off, pos, vel, t, ok=[True, False, True])
    assert len(oa) == 3 and oa[1] is None
    o0 = oa[0]
    assert o0.shape == (3,) and np.array_equal(o0.pos.xyz, pos[:, :3]) and np.array_equal(o0.t, t[:3])
    assert o0[-1].shape == (1,) and o0[-1:].pos.xyz.shape == (3, 1)
    assert np.array_equal(o0[np.array([True, False, True])].x, pos[0, [0, 2]])
    fp, fv = oa.final_coords()
    assert np.array_equal(fp[0], pos[:, 2]) and np.isinf(fp[1]).all() and np.array_equal(fp[2], pos[:, 5])
    assert np.allclose(fv[2], vel[:, 5] * K.KMS_PER_KPC_MYR)
    sub = oa[np.array([True, False, True])]
    assert len(sub) == 2 and np.array_equal(sub[1].t, t[4:6])
    assert len(oa[[2, 0]]) == 2 and np.array_equal(oa[[2, 0]][0].t, t[4:6])
    obj = np.array(oa)
    assert obj.dtype == object and obj.shape == (3,) and obj[1] is None
    cat = OrbitArray.concatenate([oa[[0]], oa[[2]]])
    assert len(cat) == 2 and list(cat.offsets) == [0, 3, 5] and np.array_equal(cat[1].pos.xyz, pos[:, 4:6])
    off2, p2, v2, t2 = oa[[2]].compact()
    assert list(off2) == [0, 2] and np.allclose(v2, vel[:, 4:6] * K.KMS_PER_KPC_MYR)


def test_event_table_from_frame_roundtrip():
    p = fake_population()
    np.random.seed(0)
    prim_df, sec_df = identify_events(p)
    prim_t, sec_t = identify_event_tables(p)
    a = event_table_from_frame(prim_df, p.bin_nums)
    assert np.array_equal(a.owner, prim_t.owner) and np.array_equal(a.dv, prim_t.dv)
    b = event_table_from_frame(sec_df, p.bin_nums)
    assert np.array_equal(b.owner, sec_t.owner) and np.array_equal(b.tphys, sec_t.tphys)


def test_abi_exports_every_declared_symbol(built):
    """Every function declared in include/cogsworth_b200.h is exported by the built library."""
    hdr = open(os.path.join(ROOT, "include", "cogsworth_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(cw_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(declared) == sorted(_lib.EXPORTS)
    L = _lib.load()
    assert L.cw_abi_version() == 2
    assert L.cw_strerror(-4).decode().startswith("no CUDA device")


def test_no_cpu_fallback_without_device(built):
    """Without a GPU the product refuses to run (cw_create -> CW_E_NO_DEVICE); it never routes to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.NativeLibraryError):
        cb.Context()
    with pytest.raises(_lib.NativeLibraryError):
        cb.integrate_orbit_with_events([8, 0, 0, 0, 220, 0], 11.9, 12.0, 1.0, events=None)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under cogsworth_b200/ may mention it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cogsworth_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", txt, flags=re.M), f
                assert "libcw_oracle" not in txt and "cwo_" not in txt, f


def test_light_orbit_surface():
    """Attribute surface of the gala-less Orbit stand-in."""
    from cogsworth_b200.orbits import Orbit
    pos = np.array([[3.0, 0.0, -1.0], [4.0, 2.0, 0.0], [0.5, -0.5, 0.1]])
    o = Orbit(pos, pos / 10, np.arange(3.0))
    assert np.allclose(o.cylindrical.rho, [5.0, 2.0, 1.0]) and np.array_equal(o.cylindrical.z, pos[2])
    assert np.allclose(o.represent_as("cylindrical").phi, np.arctan2(pos[1], pos[0]))
    assert o[-1].shape == (1,) and o[-1:].shape == (1,) and o[np.array([True, False, True])].shape == (2,)
    assert len(o.pos) == 3 and np.array_equal(o.data.xyz, pos) and np.array_equal(o.vel.d_xyz, pos / 10)
    assert np.array_equal(o.x, pos[0]) and o.ndim == 3
    with pytest.raises(NotImplementedError):
        o.represent_as("spherical")


def test_orbit_wire_format_roundtrip(tmp_path):
    """orbits/{offsets,pos,vel,t} (pop.py:1853-1876): pos kpc, vel km/s, t Myr; load is lazy and the final
    coordinates come from the last column of every orbit (pop.py:715-718)."""
    from cogsworth_b200 import constants as K
    from cogsworth_b200.orbit_io import final_coords_from_file, load_orbits, save_orbits
    from cogsworth_b200.orbits import OrbitArray
    rng = np.random.default_rng(0)
    lens = np.array([5, 1, 9, 3])
    off = np.concatenate([[0], np.cumsum(lens)])
    pos, vel, t = rng.normal(size=(3, off[-1])), rng.normal(size=(3, off[-1])) * 0.2, np.arange(off[-1], dtype=float)
    oa = OrbitArray(off, pos, vel, t)
    f = str(tmp_path / "pop.npz")
    np.savez(f, numeric_params=np.arange(3))            # an existing population file keeps its other keys
    save_orbits(f, oa)
    with np.load(f) as z:
        assert set(z.files) == {"numeric_params", "orbits/offsets", "orbits/pos", "orbits/vel", "orbits/t"}
        assert np.array_equal(z["orbits/offsets"], off) and z["orbits/pos"].shape == (3, off[-1])
        assert np.allclose(z["orbits/vel"], vel * K.KMS_PER_KPC_MYR, rtol=1e-15)
    back = load_orbits(f)
    assert len(back) == 4 and np.array_equal(back.lengths(), lens)
    assert np.array_equal(back.pos, pos) and np.allclose(back.vel, vel, rtol=1e-15) and np.array_equal(back.t, t)
    o2 = back[2]
    assert o2.shape == (9,) and np.array_equal(o2.pos.xyz, pos[:, off[2]:off[3]])
    fp, fv = final_coords_from_file(f)
    fp2, fv2 = oa.final_coords()
    assert np.array_equal(fp, fp2) and np.allclose(fv, fv2, rtol=1e-15)
    # a masked selection is written compacted, in order (primary_orbits / secondary_orbits style slicing)
    save_orbits(f, oa[np.array([True, False, True, False])])
    assert np.array_equal(load_orbits(f).lengths(), [5, 9])
    # failed orbits cannot be represented (the reference drops them before saving)
    bad = OrbitArray(off, pos, vel, t, ok=np.array([True, False, True, True]))
    with pytest.raises(ValueError):
        save_orbits(f, bad)
    with pytest.raises(ValueError):
        load_orbits(str(tmp_path / "pop.npz").replace("pop", "other")) if False else save_orbits(str(tmp_path / "x.bin"), oa)


def test_orbit_wire_format_hdf5(tmp_path):
    """The reference's own container (pop.py:1853-1876: an HDF5 group `orbits` with `offsets`, `pos`, `vel`, `t`)
    without h5py: cogsworth_b200.h5mini writes and reads the subset of the format h5py's defaults use.  Its reader is
    pinned on the one real HDF5 file in the image -- scipy's MATLAB v7.3 fixture (a 512-byte user block, a libhdf5 1.8
    writer), whose `testdouble` is pi/4 * arange(9) by scipy's own test -- and the writer against that reader."""
    from cogsworth_b200 import constants as K, h5mini
    from cogsworth_b200.orbit_io import final_coords_from_file, load_orbits, save_orbits
    from cogsworth_b200.orbits import OrbitArray
    import scipy.io
    real = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if os.path.exists(real):
        assert list(h5mini.list_datasets(real, "/")) == ["testdouble"]
        td = h5mini.read_datasets(real, "/")["testdouble"]
        assert td.shape == (9, 1) and np.array_equal(td.ravel(), np.pi / 4 * np.arange(9))
    rng = np.random.default_rng(1)
    lens = np.array([7, 2, 1, 11, 4])
    off = np.concatenate([[0], np.cumsum(lens)])
    pos, vel, t = rng.normal(size=(3, off[-1])), rng.normal(size=(3, off[-1])) * 0.2, np.arange(off[-1], dtype=float)
    oa = OrbitArray(off, pos, vel, t)
    f = str(tmp_path / "pop.h5")
    save_orbits(f, oa)
    raw = open(f, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0                 # signature, superblock version 0
    assert set(h5mini.list_datasets(f, "orbits")) == {"offsets", "pos", "vel", "t"}
    d = h5mini.read_datasets(f, "orbits")
    assert d["offsets"].dtype == np.int64 and np.array_equal(d["offsets"], off)
    assert d["pos"].shape == (3, off[-1]) and np.array_equal(d["pos"], pos)
    assert np.allclose(d["vel"], vel * K.KMS_PER_KPC_MYR, rtol=1e-15) and np.array_equal(d["t"], t)   # km/s on disk
    back = load_orbits(f)
    assert np.array_equal(back.lengths(), lens) and np.array_equal(back.pos, pos) and np.array_equal(back.t, t)
    fp, fv = final_coords_from_file(f)
    assert np.array_equal(fp, oa.final_coords()[0])
    try:
        import h5py
    except ImportError:
        with pytest.raises(FileExistsError):                              # appending needs the real library
            save_orbits(f, oa)
        with pytest.raises(ValueError):
            h5mini.write_datasets(str(tmp_path / "other.h5"), "x", {"a": np.zeros(3)})
            load_orbits(str(tmp_path / "other.h5"))
    else:                                                                 # and h5py reads what h5mini wrote
        with h5py.File(f, "r") as hf:
            assert np.array_equal(hf["orbits"]["offsets"][...], off) and np.array_equal(hf["orbits"]["pos"][...], pos)


def test_get_rel_speed_expression():
    """classify._get_rel_speed (classify.py:117-155) restated on arrays: the host-side part (cylindrical components, the
    reference's own combination) against a direct evaluation; the circular velocity itself is the GPU kernel's
    (tests/test_gpu_api.py::test_steps_either_side_of_the_path)."""
    import cogsworth_b200.classify as C
    from cogsworth_b200.orbits import OrbitArray
    rng = np.random.default_rng(3)
    n = 50
    w = np.vstack([rng.normal(0, 6, (3, n)), rng.normal(0, 0.2, (3, n))])

    class FakeCtx:
        def circular_velocity(self, q):
            return 0.22 + 0.01 * q[0]

    import cogsworth_b200.kicks as kk
    orig = kk._context
    kk._context = lambda potential, context=None: FakeCtx()
    try:
        got = C.get_rel_speed(OrbitArray.from_final(w, np.zeros(n)), potential=None)
    finally:
        kk._context = orig
    from cogsworth_b200 import constants as K
    phi = np.arctan2(w[1], w[0])
    vx, vy, vz = w[3:] * K.KMS_PER_KPC_MYR
    v_R, v_T = vx * np.cos(phi) + vy * np.sin(phi), -vx * np.sin(phi) + vy * np.cos(phi)
    vc = (0.22 + 0.01 * w[0]) * K.KMS_PER_KPC_MYR
    assert np.allclose(got, np.sqrt((v_R - vc) ** 2 + v_T ** 2 + vz ** 2), rtol=1e-13)


def test_from_gala_translates_rotations_and_time_interpolation():
    """`from_gala` on stand-ins shaped like gala's classes (class name, .parameters, .R, .items()): the halo angle phi
    and the bar angle alpha become rotations, an MN3 disk inside a TimeInterpolatedPotential becomes three
    Miyamoto-Nagai terms whose masses follow the knots (evolving_potentials.ipynb cell 13)."""
    from cogsworth_b200 import potential as P

    def fake(cls_name, params, **attrs):
        obj = type(cls_name, (), {})()
        obj.parameters = params
        for k, v in attrs.items():
            setattr(obj, k, v)
        return obj

    class Composite(dict):
        G = P.G_GALACTIC

    lm10 = Composite(disk=fake("MiyamotoNagaiPotential", dict(m=1e11, a=6.5, b=0.26)),
                     bulge=fake("HernquistPotential", dict(m=3.4e10, c=0.7)),
                     halo=fake("LogarithmicPotential", dict(v_c=0.1762, r_h=12.0, q1=1.38, q2=1.0, q3=1.36,
                                                            phi=np.deg2rad(97.0))))
    pot = P.from_gala(lm10)
    assert pot.types == [P.COMP_MIYAMOTO_NAGAI, P.COMP_HERNQUIST, P.COMP_LOG_TRIAXIAL]
    assert pot.rotations[0] is None and np.allclose(pot.rotations[2], P.LM10Potential().rotations[2])
    bar = P.from_gala(fake("LongMuraliBarPotential", dict(m=1e10, a=10.0, b=3.0, c=1.0, alpha=0.3)))
    assert bar.types == [P.COMP_LONG_MURALI_BAR] and np.allclose(bar.rotations[0], P.rotation_z(0.3))
    assert bar.params[0] == (1e10, 10.0, 3.0, 1.0)

    knots = np.linspace(0, 12000, 5)
    frac = np.linspace(0.1, 1.0, 5)
    MN3 = type("MN3ExponentialDiskPotential", (), {})
    NFW = type("NFWPotential", (), {})
    evolving = Composite(
        disk=fake("TimeInterpolatedPotential", dict(m=6.8e10 * frac, h_R=2.6, h_z=0.3), potential_cls=MN3, time_knots=knots),
        halo=fake("TimeInterpolatedPotential", dict(m=5.4e11 * frac, r_s=15.63), potential_cls=NFW, time_knots=knots))
    pot = P.from_gala(evolving)
    want = P.evolving_milky_way(knots, frac)
    assert pot.types == [1, 1, 1, 3] and pot.interp == "cubic"
    off, kt, ka = pot.knot_arrays()
    assert list(off) == [0, 5, 10, 15, 20] and np.array_equal(kt[:5], knots)
    assert np.allclose(ka[:15], want.knot_arrays()[2][:15], rtol=1e-14)        # the three MN terms of the same disk
    assert np.allclose(ka[15:], 5.4e11 * frac)
    with pytest.raises(NotImplementedError):                                     # a time-dependent scale length
        P.from_gala(fake("TimeInterpolatedPotential", dict(m=1e12, r_s=np.linspace(10, 20, 5)), potential_cls=NFW,
                         time_knots=knots))


def test_from_gala_decomposes_quantities_into_galactic_units():
    """a parameter kept in the user's units (v_c in km/s, a scale radius in pc) must be decomposed into the
    potential's unit system, not read by `.value` -- checked with astropy when it is installed."""
    u = pytest.importorskip("astropy.units")
    from cogsworth_b200 import potential as P
    galactic = [u.kpc, u.Myr, u.Msun, u.radian]

    class Units(list):
        pass

    class Hernquist:
        def __init__(self):
            self.parameters = {"m": 5e9 * u.Msun, "c": 1000.0 * u.pc}
            self.units = Units(galactic)
    Hernquist.__name__ = "HernquistPotential"

    class Log:
        def __init__(self):
            self.parameters = {"v_c": 200.0 * u.km / u.s, "r_h": 12.0 * u.kpc, "q1": 1.0, "q2": 1.0, "q3": 0.8, "phi": 0.0}
            self.units = Units(galactic)
    Log.__name__ = "LogarithmicPotential"

    class Comp(dict):
        units = Units(galactic)
        G = 4.498502151469553e-12
    pot = P.from_gala(Comp(bulge=Hernquist(), halo=Log()))
    assert pot.params[0][:2] == (5e9, 1.0)                                   # 1000 pc -> 1 kpc
    assert abs(pot.params[1][0] - 200.0 / 977.7922216807891) < 1e-15          # km/s -> kpc/Myr
    with pytest.raises(NotImplementedError):
        bad = Comp(bulge=Hernquist())
        bad.units = Units([u.pc, u.yr, u.Msun, u.radian])
        P.from_gala(bad)
