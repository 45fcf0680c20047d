#!/usr/bin/env python
"""bench.py -- FP64 orbit-steps/s of the galactic-orbit hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...     (one rank per GPU)

A "step" is one pass of the hot path over the whole batch of synthetic orbits.

Workloads (synthetic inputs seeded with numpy default_rng(20221101 + cfg_id)):
  config4 (default)  10^7 orbits (primaries + disrupted secondaries), MilkyWayPotential2022, leapfrog,
                     1 Gyr at dt = 1 Myr (1000 steps each), supernova kicks on, final state only.
                     N GPUs: the 10^7 orbits are sharded by index over the ranks (strong scaling; no
                     collective on the data path); --scaling weak gives every rank its own 10^7.
  config2            10^5 binaries with Wagg2022 horizons (<= 12000 steps), final state, kicks
  config3            10^6 orbits x 1000 stored steps (store_all, HBM-write bound), no kicks
  config5            10^6 orbits DOPRI853 rtol = atol = 1e-8, kicks, 1 % plunging orbits

Output: ONE JSON line on rank 0 (see README / task contract): value = device-resident throughput,
e2e = through the Python API (kicks.integrate_orbits) from HOST memory, H2D + D2H inside the timed region
(e2e.c_abi: cw_integrate itself; e2e.api_numpy: plain pageable numpy arrays), roofline (FP64 pipe, measured DFMA
peak), cpu_baseline (the C oracle on the host cores, bounded sample); the default one-GPU run adds other_workloads
(configs 2, 3, 5) and population_e2e (B200Population.perform_galactic_evolution from DataFrames).
--impl reference times the CPU implementation only (the gala-equivalent C oracle: gala itself is not
installable in this image) with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_STEP = {"leapfrog_v2": 114.0, "leapfrog_v1": 78.0, "dop853_v2": 2152.0,
                 # LM10 by the same counting: shared radii 6 + MN 18 + Hernquist 10 + triaxial logarithmic 13
                 # + its rotation there and back 30 + leapfrog 18
                 "leapfrog_lm10": 95.0}
BYTES_PER_STORED_STEP = 48.0

WORKLOADS = {
    "config4": dict(cfg_id=4, n_orbits=10_000_000, horizon_gyr=1.0, integrator="leapfrog", store_all=False,
                    f_sn=1.0, desc="1e7 orbits, MilkyWayPotential2022, leapfrog 1 Gyr dt=1 Myr, kicks, final state"),
    "config2": dict(cfg_id=2, n_binaries=100_000, horizon_gyr=None, integrator="leapfrog", store_all=False,
                    f_sn=1.0, desc="1e5 binaries, Wagg2022 horizons, MilkyWayPotential2022, leapfrog, kicks, final state"),
    "config2_dop853": dict(cfg_id=2, n_binaries=100_000, horizon_gyr=None, integrator="dop853", store_all=False,
                           f_sn=1.0, desc="1e5 binaries, Wagg2022 horizons, MilkyWayPotential2022, DOPRI853 1e-10, kicks"),
    "config3": dict(cfg_id=3, n_orbits=1_000_000, horizon_gyr=1.0, integrator="leapfrog", store_all=True,
                    f_sn=0.0, desc="1e6 orbits x 1000 stored steps (48 GB), MilkyWayPotential2022, leapfrog, no kicks"),
    "config3k": dict(cfg_id=3, n_orbits=1_000_000, horizon_gyr=1.0, integrator="leapfrog", store_all=True,
                     f_sn=1.0, desc="1e6 orbits x 1001 stored steps (48 GB), MilkyWayPotential2022, leapfrog, kicks on"),
    "config2_dop853_dense": dict(cfg_id=2, n_binaries=100_000, horizon_gyr=None, integrator="dop853_dense", store_all=False,
                                 f_sn=1.0, desc="1e5 binaries, Wagg2022 horizons, MilkyWayPotential2022, DOPRI853 1e-10 dense-output form, kicks"),
    "config5_dense": dict(cfg_id=5, n_orbits=1_000_000, horizon_gyr=1.0, integrator="dop853_dense", store_all=False,
                          f_sn=1.0, plunging_fraction=0.01, rtol=1e-8, atol=1e-8,
                          desc="1e6 orbits DOPRI853 rtol=atol=1e-8 (one call per segment, dense-output form), 1 Gyr, kicks, 1% plunging"),
    "config5": dict(cfg_id=5, n_orbits=1_000_000, horizon_gyr=1.0, integrator="dop853", store_all=False,
                    f_sn=1.0, plunging_fraction=0.01, rtol=1e-8, atol=1e-8,
                    desc="1e6 orbits DOPRI853 rtol=atol=1e-8, 1 Gyr, kicks, 1% plunging"),
    # the tutorial's evolving Milky Way (evolving_potentials.ipynb cell 13) through the general kernels: not a BASELINE
    # config, the cost of a time-dependent potential beside config 4
    "config4_evolving": dict(cfg_id=4, n_orbits=2_000_000, horizon_gyr=1.0, integrator="leapfrog", store_all=False,
                             f_sn=1.0, potential="evolving_mw",
                             desc="2e6 orbits, evolving Milky Way (TimeInterpolatedPotential masses, 5 knots), leapfrog 1 Gyr dt=1 Myr, kicks, final state"),
    "config5_evolving": dict(cfg_id=5, n_orbits=200_000, horizon_gyr=1.0, integrator="dop853", store_all=False,
                             f_sn=1.0, potential="evolving_mw",
                             desc="2e5 orbits, evolving Milky Way, DOPRI853 1e-10 (the tutorial's own integrator), 1 Gyr, kicks, final state"),
    "config4_lm10": dict(cfg_id=4, n_orbits=1_000_000, horizon_gyr=1.0, integrator="leapfrog", store_all=False,
                         f_sn=1.0, potential="lm10",
                         desc="1e6 orbits, LM10Potential (MN disk + Hernquist bulge + rotated triaxial logarithmic halo), leapfrog 1 Gyr dt=1 Myr, kicks, final state"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config4", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only; marks the line)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true", help="skip the other_workloads / population_e2e blocks")
    ap.add_argument("--no-inprocess", action="store_true", help="N > 1: skip the one-process-all-GPUs end-to-end leg")
    ap.add_argument("--cpu-sample-orbits", type=int, default=200_000)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def build_shard(wl, rank, world, scaling, scale):
    """This rank's orbits.  Strong scaling: rank r generates 1/world of the population with its own
    seed (the union over ranks is the workload); weak: every rank generates the full size."""
    from cogsworth_b200.synthetic import make_population, population_with_n_orbits
    share = 1.0 if scaling == "weak" else 1.0 / world
    seed = 20221101 + wl["cfg_id"] + 1000 * world + rank if world > 1 else None
    kw = dict(cfg_id=wl["cfg_id"], horizon_gyr=wl["horizon_gyr"], f_sn=wl["f_sn"],
              plunging_fraction=wl.get("plunging_fraction", 0.0), seed=seed)
    if "n_orbits" in wl:
        n = max(int(round(wl["n_orbits"] * share * scale)), 64)
        return population_with_n_orbits(n, **kw)
    n = max(int(round(wl["n_binaries"] * share * scale)), 64)
    return make_population(n, **kw)


def apply_workload_potential(wl, pop):
    """Workloads that integrate in another potential than the one the birth velocities were drawn in."""
    if wl.get("potential") == "evolving_mw":
        from cogsworth_b200.potential import evolving_milky_way
        pop.potential = evolving_milky_way(np.linspace(0.0, 12000.0, 5), np.linspace(0.1, 1.0, 5))
        return "evolving Milky Way: MN3 disk + 2 Hernquist + NFW, masses 10 % -> 100 % over 12 Gyr (general kernels)"
    if wl.get("potential") == "lm10":
        from cogsworth_b200.potential import LM10Potential
        pop.potential = LM10Potential()
        return "LM10Potential: MN disk + Hernquist bulge + logarithmic halo (q1 = 1.38, q3 = 1.36, phi = 97 deg) (general kernels)"
    return "MilkyWayPotential2022 (MN3 disk + 2 Hernquist + NFW)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] if pw else sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's host threads (and therefore its pinned buffers, first-touch) to the NUMA node its GPU hangs
    off: with 8 ranks on a two-socket box the H2D/D2H legs of the end-to-end figure otherwise cross the socket link.
    Best effort: silently does nothing when the topology cannot be read."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, scale, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    ncu --set full capture (profiles/ncu_traffic.json).  The capture of a store_all workload was taken at a
    reduced size (ncu replays need the 48 GB output twice); both kernels' traffic is linear in the number
    of orbits, so it is scaled to this run's size and the scaling is stated."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p))[workload]
    except Exception:
        return None, None, None
    f = (scale / world) / d["scale"]
    note = d["source"] + ("" if f == 1.0 else f"; captured at x{d['scale']} of the workload, scaled x{f:g} (traffic is linear in orbits)")
    return d["dram_bytes_per_launch"] * f, note, d.get("fp64_pipe_busy_pct")


def cpu_baseline(pop, wl, n_sample, threads=0, repeats=1):
    """The gala-equivalent C oracle on the host cores over a bounded sample of the same workload."""
    import oracle as O
    b = pop.batch.take(np.arange(min(n_sample, pop.batch.n_orbits)))
    integ = {"leapfrog": O.LEAPFROG, "dop853": O.DOP853, "dop853_dense": O.DOP853_DENSE}[wl["integrator"]]
    threads = threads or O.max_threads()
    best, steps = None, 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        r = O.integrate_batch(pop.potential, b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv,
                              integrator=integ, rtol=wl.get("rtol", 1e-10), atol=wl.get("atol", 1e-10), n_threads=threads)
        dt = time.perf_counter() - t0
        steps = int((r["n_nodes"].astype(np.int64) - 1).clip(min=0).sum())
        best = dt if best is None else min(best, dt)
    return {"value": steps / best, "unit": "orbit-steps/s", "cores": threads, "kind": "port",
            "sample": f"{b.n_orbits} orbits of the workload ({steps} orbit-steps, {best:.2f} s wall, "
                      "gala-equivalent C oracle -- not gala)"}, steps, best


# ------------------------------------------------------------------------------------------------
REFERENCE_ARM_NOTE = ("--impl reference: the CPU implementation of the path (gala-equivalent C oracle port on all host threads; "
                      "gala / astropy / COSMIC are not installable offline) integrates a {n}-orbit SAMPLE of this workload per "
                      "step and reports the sample's orbit-steps/s")


def config_block(args, wl, world, pot_label, numa_note=""):
    """The static description of the run: the same keys and values in both arms (--impl ours / reference)."""
    size = wl.get("n_orbits", None)
    return {"workload": f"{args.workload}: {wl['desc']}" + ("" if args.scale == 1.0 else f" [SCALED x{args.scale}: not a valid bench line]"),
            "n_orbits": size if size is not None else f"{wl['n_binaries']} binaries + their disrupted secondaries (~1.8x)",
            "integrator": wl["integrator"], "potential": pot_label,
            "parallelism": f"orbits sharded by index over {world} GPU(s), no collective" + numa_note,
            "cache": "inputs + outputs (>= 1 GB per pass) exceed the 126 MB L2; nothing is reused between steps",
            "reference_sample": REFERENCE_ARM_NOTE.format(n=args.cpu_sample_orbits)}


def run_reference(args, wl, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    size = wl.get("n_orbits", int(1.8 * wl.get("n_binaries", 0)))
    pop = build_shard(wl, 0, 1, "strong", min(args.scale, 1.0, 2.0 * args.cpu_sample_orbits / size))
    pot_label = apply_workload_potential(wl, pop)
    import oracle as O
    threads = O.max_threads()
    n_sample = min(args.cpu_sample_orbits, pop.batch.n_orbits)
    for _ in range(args.warmup):
        cpu_baseline(pop, wl, max(n_sample // 8, 64), threads)
    tot_steps, tot_t = 0, 0.0
    for _ in range(args.steps):
        cb, steps, dt = cpu_baseline(pop, wl, n_sample, threads)
        tot_steps += steps
        tot_t += dt
    value = tot_steps / tot_t
    cb["value"] = value
    line = {"impl": "reference", "metric": "FP64 orbit-steps/sec", "value": value, "unit": "orbit-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_block(args, wl, world, pot_label),
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "orbit-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class Workload:
    """One workload on one rank: device-resident copies for `value`, the OrbitBatch for the end-to-end legs."""

    def __init__(self, name, wl, pop, ctx, dev, torch, _lib):
        self.name, self.wl, self.pop, self.ctx, self.dev, self.torch, self._lib = name, wl, pop, ctx, dev, torch, _lib
        b = pop.batch
        self.b, self.n, self.K = b, b.n_orbits, b.n_events
        self.integ = {"leapfrog": _lib.LEAPFROG, "dop853": _lib.DOP853, "dop853_dense": _lib.DOP853_DENSE}[wl["integrator"]]
        self.tol = dict(atol=wl.get("atol", 1e-10), rtol=wl.get("rtol", 1e-10))
        f64, i32, i64 = torch.float64, torch.int32, torch.int64
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.D = dict(w0=to(b.w0), t1=to(b.t1), t2=to(b.t2), dt=to(b.dt), to_myr=to(b.to_myr), flags=to(b.flags))
        if self.K:
            self.D.update(ev_off=to(b.ev_off), ev_time=to(b.ev_time), ev_dv=to(b.ev_dv))
        else:
            self.D.update(ev_off=None, ev_time=None, ev_dv=None)
        self.traj_off_d = self.traj_pos_d = self.traj_vel_d = None
        if wl["store_all"]:
            nn, _ = ctx.count_nodes(None, b.t1, b.t2, b.dt, b.to_myr, b.flags)
            off = np.zeros(self.n + 1, dtype=np.int64)
            np.cumsum(nn, out=off[1:])
            self.traj_off_d = to(off)
            self.traj_pos_d = torch.empty((3, int(off[-1])), dtype=f64, device=dev)
            self.traj_vel_d = torch.empty((3, int(off[-1])), dtype=f64, device=dev)
        self.wf_d = torch.empty((6, self.n), dtype=f64, device=dev)
        self.st_d = torch.empty(self.n, dtype=i32, device=dev)
        self.stats_d = torch.zeros(4, dtype=i64, device=dev)
        self.stream = torch.cuda.current_stream(dev)
        # the plain layout in ordinary (pageable) numpy memory: what a caller without this package's batch builder has
        self.plain = None

    def step_device(self):
        D = self.D
        self.ctx.integrate_raw(self.n, D["w0"], D["t1"], D["t2"], D["dt"], D["to_myr"], D["flags"], D["ev_off"], D["ev_time"],
                               D["ev_dv"], self.wf_d, self.st_d, None, None, self.traj_off_d, self.traj_pos_d, self.traj_vel_d,
                               None, self.stats_d, integrator=self.integ, mem_space=self._lib.MEM_DEVICE,
                               stream=self.stream.cuda_stream, **self.tol)

    def step_api(self, batch=None):
        """The call a user makes: kicks.integrate_orbits on an OrbitBatch held in HOST memory (H2D of the inputs, the
        kernels, D2H of final states / statuses, the retry bookkeeping) -> OrbitArray."""
        from cogsworth_b200 import kicks
        orbits, info = kicks.integrate_orbits(self.b if batch is None else batch, potential=self.pop.potential,
                                              store_all=False, integrator=self.wl["integrator"],
                                              integrator_kwargs=dict(self.tol), context=self.ctx)
        return orbits, info

    def step_abi(self, out):
        """cw_integrate(CW_MEM_HOST) itself on the same page-locked arrays, results into page-locked `out`."""
        a = self.b.abi_arrays()
        self.ctx.integrate_raw(self.n, a["w0"], a["t1"], a.get("t2"), a.get("dt"), a.get("to_myr"), a.get("flags"),
                               a.get("ev_off"), a.get("ev_time"), a.get("ev_dv"), out[0], out[1],
                               integrator=self.integ, mem_space=self._lib.MEM_HOST, n_w0=a.get("n_w0", 0),
                               w0_index=a.get("w0_index"), grid_class=a.get("grid_class"), classes=a.get("classes"),
                               ev_off32=a.get("ev_off32"), **self.tol)

    def h2d_bytes(self):
        a = self.b.abi_arrays()
        return int(sum(v.nbytes for v in a.values() if isinstance(v, np.ndarray)))

    def free(self):
        for k in ("D", "traj_off_d", "traj_pos_d", "traj_vel_d", "wf_d", "st_d"):
            setattr(self, k, None)
        self.torch.cuda.empty_cache()


def measure(W, steps, warmup, barrier, max_over_ranks, sum_over_ranks, peak_fp64, world, args, sampler_index=None,
            with_api=True, with_variants=False):
    """value (device-resident), e2e (through the Python API from host memory) and the roofline of one workload."""
    torch, _lib, ctx, wl = W.torch, W._lib, W.ctx, W.wl
    for _ in range(warmup):
        W.step_device()
    barrier()
    steps_per_pass = int(W.stats_d.cpu()[3])
    sampler = ClockSampler(sampler_index) if sampler_index is not None else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, total_kernel_ms, launches = [], [], 0
    e0.record(W.stream)
    for _ in range(steps):
        W.step_device()
        kernel_ms.append(ctx.last_kernel_ms(0))
        total_kernel_ms.append(ctx.last_total_kernel_ms(0))
        launches += ctx.last_launch_count
    e1.record(W.stream)
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    stats_run = W.stats_d.cpu().numpy().copy()
    total_steps = sum_over_ranks(float(steps_per_pass))
    value = total_steps * steps / (dev_ms * 1e-3)
    n_failed = int((W.st_d != 0).sum().item())

    # ---- end to end from host memory -------------------------------------------------------------------------
    e2e = None
    if with_api and not wl["store_all"]:
        def timed(fn, nwarm=max(1, min(warmup, 2))):
            for _ in range(nwarm):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                fn()
            barrier()
            return max_over_ranks(time.perf_counter() - t0)

        res = {}
        s_api = timed(lambda: res.__setitem__("o", W.step_api()))
        orbits, info = res["o"]
        okd = (W.st_d == 0).cpu().numpy()          # (failed orbits are re-integrated with a finer grid by the API)
        assert np.array_equal(np.asarray(info["w_final"])[:, okd], W.wf_d.cpu().numpy()[:, okd]), \
            "host-buffer path and device path disagree"
        h2d = W.h2d_bytes()
        d2h = W.n * (48 + 4 + 4) + 4 * W.K + 40
        e2e = {"value": total_steps * steps / s_api, "unit": "orbit-steps/s",
               "h2d_bytes_per_step": int(sum_over_ranks(float(h2d))), "d2h_bytes_per_step": int(sum_over_ranks(float(d2h))),
               "ms_per_step": 1e3 * s_api / steps,
               "how": "cogsworth_b200.kicks.integrate_orbits(batch, store_all=False) -- the batch form of "
                      "integrate_orbit_with_events that perform_galactic_evolution calls -- on the OrbitBatch "
                      "build_orbit_batch makes (host memory, page-locked, compact cw_batch forms): H2D of initial conditions / "
                      "start times / grid classes / events, grid pre-pass + integrator kernels, D2H of final states, statuses, "
                      "node counts and kick nodes into page-locked arrays, retry bookkeeping, OrbitArray; host wall clock, "
                      "max over ranks"}
        if with_variants:
            out = (_lib.pinned_empty((6, W.n)), _lib.pinned_empty(W.n, np.int32))
            s_abi = timed(lambda: W.step_abi(out))
            e2e["c_abi"] = {"value": total_steps * steps / s_abi, "ms_per_step": 1e3 * s_abi / steps,
                            "how": "cw_integrate(CW_MEM_HOST) called directly on the same page-locked arrays"}
            from cogsworth_b200.batch import OrbitBatch
            b = W.b
            cp = lambda x: None if x is None else np.array(x, copy=True)
            plain = OrbitBatch(cp(b.w0), cp(b.t1), cp(b.t2), cp(b.dt), cp(b.to_myr), cp(b.flags), cp(b.ev_off),
                               cp(b.ev_time), cp(b.ev_dv))
            s_np = timed(lambda: W.step_api(plain))
            e2e["api_numpy"] = {"value": total_steps * steps / s_np, "ms_per_step": 1e3 * s_np / steps,
                                "h2d_bytes_per_step": int(sum_over_ranks(float(sum(
                                    v.nbytes for v in plain.abi_arrays().values() if isinstance(v, np.ndarray))))),
                                "how": "the same call on the PLAIN layout in ordinary (pageable) numpy arrays: staged through "
                                       "the library's page-locked ring by its host thread pool, chunk by chunk"}
            del plain

    if with_api and wl["store_all"]:
        # Whole trajectories through the Python API (store_entire_orbits=True, the reference's default): a bounded
        # sample of the workload, because every orbit-step returns 56 bytes (6 coordinates + the time stamp) over the
        # host link and 10^6 orbits would need 56 GB of host memory.  The figure is the link's, not the kernel's.
        from cogsworth_b200 import kicks
        ns = int(min(W.n, 50_000))
        sub = W.b.take(np.arange(ns))
        run = lambda: kicks.integrate_orbits(sub, potential=W.pop.potential, store_all=True, integrator=wl["integrator"],
                                             integrator_kwargs=dict(W.tol), context=ctx)
        for _ in range(2):
            orbits, info = run()
        t0 = time.perf_counter()
        for _ in range(2):
            orbits, info = run()
        s_tr = (time.perf_counter() - t0) / 2
        st = float(info["stats"][3])
        nbytes = 56.0 * float(np.asarray(info["n_nodes"], dtype=np.int64).sum())
        e2e = {"value": st / s_tr, "unit": "orbit-steps/s", "ms_per_step": 1e3 * s_tr,
               "h2d_bytes_per_step": int(sum(v.nbytes for v in sub.abi_arrays().values() if isinstance(v, np.ndarray))),
               "d2h_bytes_per_step": int(nbytes) + ns * 56,
               "d2h_gb_per_s": nbytes / s_tr / 1e9,
               "sample": f"{ns} orbits of the workload ({int(st)} orbit-steps, {nbytes / 1e9:.2f} GB of trajectories per call)",
               "how": "cogsworth_b200.kicks.integrate_orbits(batch, store_all=True) on a bounded sample: kernels write the "
                      "trajectories chunk by chunk into bounded device buffers, D2H into page-locked host arrays overlaps the "
                      "next chunk's kernel; bound by the host link"}
        del orbits, info, sub

    # ---- roofline of the dominant kernel -----------------------------------------------------------------------
    k_ms = float(np.mean(kernel_ms))
    hbm_peak, hbm_src = measured_peaks()
    traffic, traffic_src, pipe_pct = ncu_traffic(W.name, args.scale, world if args.scaling == "strong" else 1)
    if wl["store_all"]:
        ach = BYTES_PER_STORED_STEP * steps_per_pass / (k_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": hbm_src,
                "kernel": "leapfrog_kernel<MW_V2,STORE>",
                "kernel_ms": k_ms, "algorithmic_bytes_per_orbit_step": BYTES_PER_STORED_STEP,
                "algorithmic_bytes_per_launch": BYTES_PER_STORED_STEP * steps_per_pass}
    else:
        if W.integ == _lib.LEAPFROG:
            per = FLOP_PER_STEP["leapfrog_lm10" if wl.get("potential") == "lm10" else "leapfrog_v2"]
            flops = per * steps_per_pass
            kname = {"evolving_mw": "leapfrog_kernel<MW_V2_T>", "lm10": "leapfrog_kernel<LM10>"}.get(wl.get("potential"), "leapfrog_kernel<MW_V2>")
        else:
            # 12 gradients + 1000 flop of stage algebra per accepted step, per launch
            flops = FLOP_PER_STEP["dop853_v2"] * float(stats_run[0])
            kname = "dop853_kernel<MW_V2%s%s>" % ("_T" if wl.get("potential") == "evolving_mw" else "",
                                                  ",DENSE" if W.integ == _lib.DOP853_DENSE else "")
            per = FLOP_PER_STEP["dop853_v2"]
        ach = flops / (k_ms * 1e-3) / 1e12
        roof = {"bound": "fp64", "achieved": ach, "peak": peak_fp64, "unit": "TFLOP/s", "frac": ach / peak_fp64,
                "traffic": traffic if W.integ == _lib.LEAPFROG else None, "traffic_source": traffic_src if W.integ == _lib.LEAPFROG else None,
                "algorithmic_bytes_per_launch": float(108 * W.n + 28 * W.K),
                "fp64_pipe_busy_ncu": (pipe_pct / 100.0) if (pipe_pct is not None and W.integ == _lib.LEAPFROG) else None,
                "peak_source": "measured live: DFMA micro-benchmark cw_measure_fp64_peak (MEASURED_PEAKS.json has no "
                               "FP64 entry; nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2)",
                "kernel": kname, "kernel_ms": k_ms, "prepass_sort_pilot_ms": float(np.mean(total_kernel_ms)) - k_ms,
                "algorithmic_flop_per_unit": per,
                "kernel_share_of_step": k_ms / (dev_ms / steps)}
    run = {"n_orbits_total": int(sum_over_ranks(float(W.n))), "n_kick_events_rank0": W.K,
           "orbit_steps_per_step": int(total_steps), "failed_orbits_rank0": n_failed}
    if W.integ != _lib.LEAPFROG:
        run["dop853"] = {"accepted_steps": int(stats_run[0]), "rejected_steps": int(stats_run[1]),
                         "gradient_evals": int(stats_run[2]), "intervals": int(stats_run[3]),
                         "gradient_evals_per_s": float(stats_run[2]) / (k_ms * 1e-3),
                         "max_attempts_one_orbit": ctx.last_max_attempts,
                         "note": "an orbit is sequential: max_attempts_one_orbit x the latency of one 12-stage attempt "
                                 "bounds the launch from below, whatever the other lanes do"}
    return dict(value=value, ms_per_step=dev_ms / steps, e2e=e2e, roofline=roof, run=run, launches=int(launches),
                clocks=clocks, total_steps=total_steps)


def population_e2e(size_label, n_binaries, horizon_gyr, cfg_id, ctx, reps=1):
    """B200Population.perform_galactic_evolution() from synthetic bpp / kick_info / initial_binaries DataFrames: the whole
    galactic stage as a Population user runs it, with the host pipeline itemised."""
    from cogsworth_b200.batch import build_orbit_batch
    from cogsworth_b200.events import identify_event_tables
    from cogsworth_b200.pop import B200Population
    from cogsworth_b200.synthetic import make_population, population_tables
    pop = make_population(n_binaries, cfg_id=cfg_id, horizon_gyr=horizon_gyr, f_sn=1.0)
    galaxy, bpp, kick_info, initial_binaries = population_tables(pop)
    P = B200Population(galaxy, bpp, kick_info, initial_binaries, galactic_potential=pop.potential, max_ev_time=12.0,
                       timestep_size=1.0, store_entire_orbits=False, integrator="leapfrog")
    P._cw_context = ctx
    P.disrupted                                   # cached by the reference as well (pop.py:861-874)
    best, first = None, None
    for k in range(reps + 1):                     # the first pass also pins the context's staging ring (once per context)
        t0 = time.perf_counter()
        P.perform_galactic_evolution(progress_bar=False)
        fp = P.final_pos
        dt = time.perf_counter() - t0
        if k == 0:
            first = dt
        if best is None or dt < best:
            best = dt
    # itemise the host side by running its stages on their own (memory policy as inside the stage)
    from cogsworth_b200 import kicks
    from cogsworth_b200._lib import pinned_policy
    with pinned_policy("cached"):
        t0 = time.perf_counter(); pe, se = identify_event_tables(P); t_ev = time.perf_counter() - t0
        t0 = time.perf_counter()
        b = build_orbit_batch(pop.pos_kpc, pop.vel_kms, pop.tau_gyr, 12.0, 1.0, P.disrupted, pe, se)
        t_b = time.perf_counter() - t0
        t0 = time.perf_counter()
        kicks.integrate_orbits(b, potential=pop.potential, store_all=False, integrator="leapfrog", context=ctx)
        t_i = time.perf_counter() - t0
    steps = float(P._last_integration_info["stats"][3])
    assert np.isfinite(np.asarray(fp)).all()
    return {"size": size_label, "n_binaries": int(n_binaries), "n_orbits": int(b.n_orbits), "orbit_steps": steps,
            "seconds": best, "first_call_seconds": first, "value": steps / best, "unit": "orbit-steps/s",
            "seconds_itemised": {"identify_event_tables (numpy joins of bpp / kick_info / initial_binaries)": t_ev,
                                 "build_orbit_batch (w0, start times, grid classes, kick rotation, CSR)": t_b,
                                 "integrate_orbits (H2D + kernels + D2H + OrbitArray)": t_i,
                                 "rest (Quantity -> array, stacking, final_pos)": max(best - t_ev - t_b - t_i, 0.0)},
            "how": "B200Population(initial_galaxy, bpp, kick_info, initial_binaries).perform_galactic_evolution(); "
                   "final_pos read back; leapfrog, final state only; batch and result arrays in ordinary memory (the "
                   "stage pins nothing but the library's staging ring, once per context: first_call_seconds includes it)"}


def main():
    args = parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist
    import cogsworth_b200 as cb
    from cogsworth_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from cogsworth_b200 import dist as cwdist
    max_over_ranks = lambda x: cwdist.reduce_scalar(x, "max", device=dev)
    sum_over_ranks = lambda x: cwdist.reduce_scalar(x, "sum", device=dev)

    ctx = cb.Context(devices=[local_rank])
    peak = ctx.measure_fp64_peak(0, 200.0)
    pop = build_shard(wl, rank, world, args.scaling, args.scale)
    pot_label = apply_workload_potential(wl, pop)
    ctx.set_potential(pop.potential)
    W = Workload(args.workload, wl, pop, ctx, dev, torch, _lib)
    main_res = measure(W, args.steps, args.warmup, barrier, max_over_ranks, sum_over_ranks, peak, world, args,
                       sampler_index=local_rank if rank == 0 else None, with_variants=(world == 1))
    cbase = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cbase, _, _ = cpu_baseline(pop, wl, args.cpu_sample_orbits)
    W.free()
    del W

    # ---- the other BASELINE configs and the Population-level end-to-end figure (one GPU, rank 0) -----------------
    others, pop_e2e = None, None
    single = (lambda x: x)
    if rank == 0 and world == 1 and args.workload == "config4" and args.scale == 1.0 and not args.no_other_workloads:
        others = {}
        for name in ("config2", "config2_dop853", "config5", "config3"):
            owl = WORKLOADS[name]
            opop = build_shard(owl, 0, 1, "strong", 1.0)
            apply_workload_potential(owl, opop)
            ctx.set_potential(opop.potential)
            OW = Workload(name, owl, opop, ctx, dev, torch, _lib)
            r = measure(OW, 2, 2, torch.cuda.synchronize, single, single, peak, 1, args)
            others[name] = {"workload": owl["desc"], "value": r["value"], "unit": "orbit-steps/s",
                            "ms_per_step": r["ms_per_step"],
                            "e2e": None if r["e2e"] is None else {k: r["e2e"][k] for k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step", "d2h_gb_per_s", "sample") if k in r["e2e"]},
                            "roofline": {k: r["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "kernel", "kernel_ms")},
                            "run": r["run"], "steps": 2, "warmup": 2}
            OW.free()
            del OW, opop
        ctx.set_potential(pop.potential)
        pop_e2e = [population_e2e("config 2 (1e5 binaries, Wagg2022 horizons)", 100_000, None, 2, ctx),
                   population_e2e("config 4 (1e7 orbits = 5.56e6 binaries, 1 Gyr)", int(round(10_000_000 / 1.8)), 1.0, 4, ctx)]

    # ---- N > 1: the same 10^7 orbits through ONE process driving every GPU (what a Population user gets) --------------
    inproc = None
    pg_alive = world > 1
    if world > 1 and not args.no_inprocess:
        # A Population user runs ONE process on the node.  The other ranks therefore leave first -- group destroyed,
        # context closed, process gone, so that rank 0 is alone on every GPU as that user would be -- and rank 0 learns
        # it from the rendezvous store.  (With the other ranks merely parked the leg measured 69 ms per pass on 8 GPUs
        # against 10.3 ms for the same call in a process of its own, tools/gpu_inproc_probe.py.)
        import datetime
        from torch.distributed.distributed_c10d import _get_default_store
        store = _get_default_store()
        barrier()
        ctx.close()
        dist.destroy_process_group()
        pg_alive = False
        if rank != 0:
            store.set(f"cw_rank_gone_{rank}", "1")
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)            # no interpreter teardown: the CUDA context goes with the process, at once
        try:
            store.wait([f"cw_rank_gone_{r}" for r in range(1, world)], datetime.timedelta(seconds=300))
            time.sleep(2.0)        # the driver tears the other processes' contexts down asynchronously
            full = build_shard(wl, 0, 1, "strong", args.scale)
            apply_workload_potential(wl, full)
            with cb.Context(devices=list(range(world))) as call:
                call.set_potential(full.potential)
                from cogsworth_b200 import kicks
                run1 = lambda: kicks.integrate_orbits(full.batch, potential=full.potential, store_all=False,
                                                      integrator=wl["integrator"], context=call)
                for _ in range(3):        # (results are kept as in the timed loop: two generations of page-locked
                    _, info = run1()      # output buffers are then in the library's cache)
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    _, info = run1()
                dt = time.perf_counter() - t0
                st = float(info["stats"][3])
                inproc = {"value": st * args.steps / dt, "unit": "orbit-steps/s", "ms_per_step": 1e3 * dt / args.steps,
                          "how": f"rank 0 alone on the node (the other ranks have exited), one cw_ctx on all {world} GPUs "
                                 "(one host thread per device, orbits sharded by index), kicks.integrate_orbits on the "
                                 "whole workload in page-locked host memory; host wall clock"}
        except Exception as exc:          # the headline line must still be printed
            inproc = {"value": None, "error": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        numa_note = f"; ranks bound to their GPU's NUMA node (rank 0: node {numa_node})" if numa_node is not None else ""
        line = {"metric": "FP64 orbit-steps/sec", "value": main_res["value"], "unit": "orbit-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"],
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config_block(args, wl, world, pot_label, numa_note),
                "e2e": main_res["e2e"], "gpu_launches": main_res["launches"], "clocks": main_res["clocks"],
                "roofline": main_res["roofline"], "cpu_baseline": cbase, "run": main_res["run"]}
        if others is not None:
            line["other_workloads"] = others
        if pop_e2e is not None:
            line["population_e2e"] = pop_e2e
        if inproc is not None:
            line["e2e"]["in_process_all_gpus"] = inproc
        print(json.dumps(line), flush=True)
    if pg_alive:
        ctx.close()
        dist.barrier()
        dist.destroy_process_group()
    elif world == 1:
        ctx.close()


if __name__ == "__main__":
    main()
