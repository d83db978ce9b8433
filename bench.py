#!/usr/bin/env python
"""Benchmark of the MicMec force + integration hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--grid G] [--ensemble npt|nvt|nve] [--impl ours|reference]

Metric (BASELINE.json): MD node-steps/s (force + Verlet, fp64).  A "step" is one VerletIntegrator.propagate of the
synthetic G^3-cell periodic fcu grid (default 256^3, NPT = NHC thermostat + MTK barostat = 3 force evaluations per
step, BASELINE.json configs[3]).  One JSON line is printed by rank 0.

`value`        device-resident loop (mm_md_run), state already in HBM, CUDA events on the launching stream
`e2e`          the same step driven through the C ABI with HOST buffers: pinned-host pos+vel uploaded, one step,
               pos+vel+scalars read back, every step
`roofline`     dominant kernel (the fused kick-drift-force-kick launch k_march2<STEP>): algorithmic bytes / measured launch
               time vs the measured HBM copy bandwidth in MEASURED_PEAKS.json; `fp64` = the same launch against the measured
               FP64 issue peak, `binding` = whichever roofline is the slower one, `traffic` = DRAM bytes of that kernel
               from the committed ncu capture (profiles/ncu_traffic.json)
`cpu_baseline` the CPU oracle port (oracle/micmec_oracle.c, OpenMP) timed on the box's host cores on a bounded sample
`check`        scalars after the run (identical initial state at every N: they agree across N to ~1e-15) and the drift of
               the conserved quantity over the timed steps relative to the kinetic energy
`--impl reference` times that CPU port alone, with all host threads, on the same metric (its input is built with NumPy only).
`--config5 R`  BASELINE.json configs[4] instead: geometry optimisation of R 27-node replicas, sharded over the GPUs.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_NODE_STEP = 108.0   # pos r/w 48 + vel r/w 48 + mass 8 + cell type 4, per force-evaluation-step
BYTES_PER_NODE_EVAL = 52.0    # force-only evaluation: pos 24 + type 4 + gpos 24
FLOPS_PER_NODE_STEP = 600.0   # `original` model, one state: 0.55-0.65 kflop per cell and force evaluation (FMA = 2)
BOLTZMANN = 3.1668154051341965e-06  # molmod.boltzmann (micmec_b200/units.py)
# unmodified Python reference, single core, measured in the build container: NOT this host
REFERENCE_PYTHON = {"value": 2.7e3, "unit": "node-steps/s", "cores": 1, "where": "build box, not this host",
                    "what": "unmodified micmec (Python/NumPy) NPT NHC+MTK MD on data/4x4x4_fcu_micmec.chk, 50 steps: 42.5 steps/s "
                            "(BASELINE.md section 2); NVE 3x3x3_test 1.03e4, NVT 5x5x5_fcu_hollow 3.2e4"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--ensemble", default="npt", choices=["nve", "nvt", "npt"])
    ap.add_argument("--model", default="original", choices=["original", "default"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-grid", type=int, default=48, help="edge of the bounded CPU sample grid")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--chunk", type=int, default=0, help="tuning: planes per block along z")
    ap.add_argument("--wrap", type=int, default=-1, help="tuning: 0 = k_march2 reads the periodic images from the ghost nodes")
    ap.add_argument("--plan", type=int, default=-1, help="tuning: 0 = uniform chunk length for every tile (no two-class block schedule)")
    ap.add_argument("--tail-in-kernel", type=int, default=-1, help="tuning: 1 = the last block of a marching launch runs the tail")
    ap.add_argument("--tail", type=int, default=-1, help="tuning: 0 = reduction / exchange / scalar algebra in their own launches")
    ap.add_argument("--march2", type=int, default=-1, help="tuning: 0 = keep the general kernel k_march for the one-type grids too")
    ap.add_argument("--generic", action="store_true", help="force the indexed-topology kernels (no structured path)")
    ap.add_argument("--config5", type=int, default=0, metavar="R",
                    help="BASELINE.json configs[4] instead of the MD step: geometry optimisation of R perturbed 27-node "
                         "replicas of the 3x3x3 defect configurations (device-resident lockstep QN), sharded over the GPUs")
    return ap.parse_args()


FORCE_EVALS = {"nve": 1, "nvt": 1, "npt": 3}


def global_fields(grid, seed=0, amp=0.1):
    """Displacements and velocities of ALL G^3 nodes in reference order (id = (k*G + l)*G + m), NumPy only: `amp` bohr
    Gaussian displacements and Maxwell-Boltzmann velocities at 300 K (the recipe of sampling/utils.get_random_vel) with
    the centre-of-mass motion removed, from ONE seeded stream.  Every rank of a multi-GPU run draws the same global
    arrays and keeps its z-slab, so N = 1, 2, 4 and 8 integrate the bit-identical initial state."""
    from micmec_b200.celltypes import TYPE_FCU  # constants only (data/4x4x4_fcu_micmec.chk:365-398), pure Python

    n = grid ** 3
    rng = np.random.default_rng(seed)
    dpos = amp * rng.standard_normal((n, 3))
    vel = rng.standard_normal((n, 3)) * np.sqrt(BOLTZMANN * 300.0 / float(TYPE_FCU["mass"]))
    vel -= vel.mean(axis=0)
    return dpos, vel


def make_state(grid, seed=0, amp=0.1, explicit=False):
    """Synthetic G^3 fcu grid (one GPU): rest lattice + the global displacement / velocity fields."""
    from micmec_b200.system import System
    from micmec_b200.celltypes import TYPE_FCU

    system = System.periodic_grid((grid,) * 3, TYPE_FCU, explicit=explicit)
    dpos, vel = global_fields(grid, seed, amp)
    system.pos += dpos
    return system, vel


def md_params(ensemble):
    from micmec_b200.units import femtosecond, pascal

    p = dict(timestep=10.0 * femtosecond, temp=300.0, press=1e6 * pascal, timecon_thermo=100.0 * femtosecond,
             # 1e5 fs: the class default (1000 fs) is numerically unstable for these stiff cells at dt = 10 fs
             # (the unmodified reference collapses the cell in two steps, see tests/golden/make_golden.py)
             timecon_baro=1.0e5 * femtosecond, chain_vel0=np.array([1e-4, -2e-4, 5e-5]),
             vel_press0=1e-8 * np.array([[1.0, 0.2, -0.1], [0.2, -0.5, 0.3], [-0.1, 0.3, 0.8]]))
    p["thermo"] = ensemble in ("nvt", "npt")
    p["baro"] = ensemble == "npt"
    return p


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (pynvml).

    NVML is initialised in the constructor, BEFORE the timed region: nvmlInit attaches to every GPU of the node and took
    up to 200 ms on a busy box, during which this process's own graph launches stalled (profiles/r02, calls ag / ah: the
    same binary at 1.36 and 1.53-1.80 ms per step).  Inside the region only the two cheap queries run, every 20 ms.
    """

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self._nv = self._h = None
        self._names = {}
        try:
            import pynvml as nv

            nv.nvmlInit()
            self._h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
            self._names = {
                getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            self._nv = nv
        except Exception as exc:  # pragma: no cover
            self.reasons.add("unavailable: %s" % exc)

    def sample(self):
        nv = self._nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
        try:
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
            for bit, name in self._names.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        if self._nv is None:
            return
        try:
            while not self._stop_evt.is_set():
                self.sample()
                self._stop_evt.wait(0.02)
        except Exception as exc:  # pragma: no cover
            self.reasons.add("unavailable: %s" % exc)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_oracle_rate(grid, ensemble, model, steps, nthreads, target_seconds=None):
    """node-steps/s of the CPU oracle port (test infrastructure; here it is only the thing being TIMED as the
    reported CPU baseline, never the product path).  Its input is built with NumPy alone: nothing of the product
    library is loaded by this leg.  With target_seconds the step count is chosen from a two-step pilot so that the
    timed run lasts about that long."""
    from oracle import oracle as orc
    from micmec_b200.celltypes import TYPE_FCU  # pure-Python constants

    arrays, pos, masses, rvecs = orc.periodic_grid_system((grid,) * 3, TYPE_FCU)
    dpos, vel = global_fields(grid)
    pos = pos + dpos
    p = md_params(ensemble)
    o = orc.Oracle(model=model, nthreads=nthreads, **arrays)
    thermo = dict(temp=p["temp"], timecon=p["timecon_thermo"], chain_vel0=p["chain_vel0"], chain_pos0=np.zeros(3)) if p["thermo"] else None
    baro = dict(temp=p["temp"], press=p["press"], timecon=p["timecon_baro"], vel_press0=p["vel_press0"]) if p["baro"] else None
    md = o.md(pos, vel, masses, rvecs, p["timestep"], thermo=thermo, baro=baro)
    md.run(1)
    if target_seconds:
        t0 = time.perf_counter()
        md.run(2)
        steps = int(min(2000, max(steps, np.ceil(target_seconds / max((time.perf_counter() - t0) / 2, 1e-6)))))
    t0 = time.perf_counter()
    md.run(steps)
    dt = time.perf_counter() - t0
    return grid ** 3 * steps / dt, dt, steps


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path = the pinned oracle port (oracle/micmec_oracle.c), all host threads,
    on a bounded sample grid of the same workload (the line's config names the grid it really ran)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rate, dt, _ = cpu_oracle_rate(args.cpu_grid, args.ensemble, args.model, args.steps, cores)
    sample = "%d^3-cell fcu grid, %s, %d steps, OpenMP x%d (bounded sample of the %d^3 workload)" % (
        args.cpu_grid, args.ensemble.upper(), args.steps, cores, args.grid)
    config = workload_config(args, world)
    config["workload"] += "; THIS ARM RAN the bounded %d^3-cell sample of it on the host cores" % args.cpu_grid
    config["sample_grid"] = args.cpu_grid
    config["nodes_total"] = config["nodes_per_gpu"] = args.cpu_grid ** 3
    config["parallelism"] = "CPU arm: OpenMP x%d, no GPU" % cores
    config["cache"] = "host memory"
    line = {
        "impl": "reference", "metric": "MD node-steps/s (force+Verlet, fp64)", "value": rate, "unit": "node-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": rate, "unit": "node-steps/s", "cores": cores, "kind": "port", "sample": sample,
                         "reference_python": REFERENCE_PYTHON},
        "e2e": {"value": rate, "unit": "node-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


COMM_MODES = {
    0: "NCCL send/recv halo planes + NCCL all-reduce of 16 doubles",
    1: "halo planes stored into the neighbours' inboxes over NVLink peer memory + one-kernel mailbox all-reduce",
    2: "fused halo: k_march stores its boundary planes into the neighbours' halo planes over NVLink peer memory + one-kernel mailbox all-reduce",
}


def workload_config(args, world, comm_mode=None, tiling=None):
    return {
        "workload": "synthetic %dx%dx%d-cell fcu grid %s MD (%s), dt 10 fs" % (
            args.grid, args.grid, args.grid, args.ensemble.upper(),
            {"nve": "velocity Verlet", "nvt": "NHC thermostat", "npt": "NHC thermostat + MTK barostat"}[args.ensemble]),
        "nodes_total": args.grid ** 3, "nodes_per_gpu": args.grid ** 3 // world, "force_evals_per_step": FORCE_EVALS[args.ensemble], "model": args.model,
        "parallelism": "single GPU" if world == 1 else "%d z-slabs (one per GPU); %s" % (world, COMM_MODES.get(comm_mode, "CPU arm: no exchange")),
        "cache": "inputs larger than L2 (pos/vel/gpos %.0f MB each)" % (24.0 * args.grid ** 3 / 1e6),
        "kernel_tiling": tiling,
        # deviations from SURVEY.md 8(d), same work per step: MTK time constant 1e5 fs instead of the class default 1000 fs
        # (the unmodified reference collapses the cell of these stiff systems within two steps at 1000 fs,
        # tests/golden/make_golden.py), displacement amplitude 0.1 bohr instead of 0.05 h0
        "initial_state": "global_fields(seed 0): 0.1 bohr Gaussian displacements, 300 K Maxwell-Boltzmann velocities, identical for every N",
        "timecon_baro_fs": 1.0e5, "timecon_thermo_fs": 100.0,
    }


def run_config5(args, rank, world, local_rank):
    """configs[4]: R = args.config5 replicas (3 defect topologies x perturbations of a 27-node 3x3x3 grid: the type maps of
    data/3x3x3_conf{0,3,9}_micmec.chk are not shipped to the GPU box, so the defects are drawn here - 2 types per system as
    in the reference's fixtures), Cartesian geometry optimisation to gpos_rms 1e-7 / dpos_rms 1e-5, replicas only: every
    rank optimises its contiguous share, no communication."""
    import torch
    import torch.distributed as dist
    from micmec_b200.system import System
    from micmec_b200.celltypes import TYPE_FCU, TYPE_TEST
    from micmec_b200.replicas import ReplicaBatch
    from micmec_b200.sampling.batchopt import DeviceReplicaQNOptimizer, shard

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    rng = np.random.default_rng(5)
    base = System.periodic_grid((3, 3, 3), TYPE_FCU, explicit=True)
    params = dict(base.params)
    for key in ("cell", "elasticity", "free_energy", "effective_temp", "mass"):
        params["type2/" + key] = TYPE_FCU[key] if key in ("cell", "mass") else TYPE_TEST[key] if key == "elasticity" else TYPE_FCU[key]
    topologies = [np.where(rng.random(27) < 0.15, 2, 1) for _ in range(10)]  # conf0..conf9: a few softer defect cells each
    systems = []
    for r in range(args.config5):
        types = topologies[r % 10]
        s = System(base.pos + 0.5 * rng.standard_normal(base.pos.shape), base.masses, np.array(base.domain.rvecs), base.surrounding_cells,
                   base.surrounding_nodes, grid=types.reshape(3, 3, 3), types=types, params=params)
        systems.append(s)
    mine = shard(systems, rank, world)
    pos0 = np.stack([s.pos for s in mine])
    rvecs0 = np.stack([np.array(s.domain.rvecs) for s in mine])
    batch = ReplicaBatch(mine, device=local_rank)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    opt = DeviceReplicaQNOptimizer(batch, pos0, rvecs0, dof="cartesian", gpos_rms=1e-7, dpos_rms=1e-5)
    sweeps = opt.run(2000)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = opt._fetch()
    stats = torch.tensor([dt, float(st["converged"].sum()), float(st["failed"].sum()), float(sweeps)], device="cuda", dtype=torch.float64)
    if world > 1:
        tmax = stats.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        dt, sweeps = float(tmax[0]), int(tmax[3])
    if rank == 0:
        print(json.dumps({
            "metric": "geometry optimisations/s (config 5: replicas x lockstep QN, fp64)", "value": args.config5 / dt, "unit": "replicas/s",
            "n_gpus": world, "steps": sweeps, "warmup": 0, "ms_per_step": 1e3 * dt / max(sweeps, 1), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "ensemble of %d perturbed 27-node 3x3x3 defect configurations, Cartesian geometry optimisation "
                                   "(QNOptimizer semantics), replicas only" % args.config5,
                       "replicas_per_gpu": len(mine), "sweeps": sweeps, "kernels": "mm_qn (k_qn_refresh, k_batched_eigh, k_qn_step, k_cells + k_gather, k_qn_accept)"},
            "check": {"converged": int(stats[1]), "failed": int(stats[2]), "mean_iterations": float(st["iterations"].mean()),
                      "mean_energy": float(st["f"].mean())},
            "gpu_launches": int(batch.launches), "wall_s": dt}))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.config5:
        return run_config5(args, rank, world, local_rank)

    import torch
    import torch.distributed as dist

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    from micmec_b200 import _lib
    from micmec_b200.pes.mmff import MicMecForceField, ForcePartMechanical
    from micmec_b200.sampling.verlet import VerletIntegrator
    from micmec_b200.sampling.nvt import NHCThermostat
    from micmec_b200.sampling.npt import MTKBarostat, TBCombination

    lib = _lib.load()
    p = md_params(args.ensemble)
    ndof = None
    if world == 1:
        system, vel0 = make_state(args.grid)
        nnodes = nglobal = system.nnodes
        part = ForcePartMechanical(system, model=args.model, device=local_rank, structured=False if args.generic else None)
    else:
        # strong scaling: the SAME G^3 grid, the SAME initial state (global_fields), cut into `world` z-slabs, one per GPU
        from micmec_b200 import slab as slabmod
        from micmec_b200.celltypes import TYPE_FCU

        layout = slabmod.SlabLayout((args.grid,) * 3, rank, world)
        system = slabmod.local_system(layout, TYPE_FCU)
        dpos, gvel = global_fields(args.grid)
        ids = layout.global_ids()
        system.pos += dpos[ids]
        vel0 = np.ascontiguousarray(gvel[ids])
        del dpos, gvel, ids
        nnodes, nglobal = system.nnodes, layout.nnodes_global
        ndof = 3 * nglobal - (3 if (p["thermo"] or p["baro"]) else 0)
        part = ForcePartMechanical(system, model=args.model, device=local_rank, slab=layout.slab_arg())
    if args.march2 >= 0:
        _lib.check(lib.mm_set_option(part.handle, b"march2", args.march2))
    for name, val in (("wrap_on_load", args.wrap), ("tail", args.tail), ("tail_in_kernel", args.tail_in_kernel), ("plan", args.plan)):
        if val >= 0:
            _lib.check(lib.mm_set_option(part.handle, name.encode(), val))
    if args.chunk:
        _lib.check(lib.mm_set_option(part.handle, b"chunk", args.chunk))
    mmf = MicMecForceField(system, [part])
    stream = torch.cuda.Stream(device=local_rank)
    _lib.check(lib.mm_set_stream(part.handle, ctypes.c_void_p(stream.cuda_stream)))
    comm_mode = None
    if world > 1:
        slabmod.init_comm(part, layout)
        comm_mode = int(lib.mm_comm_mode(part.handle))
    hooks = []
    thermo = baro = None
    if p["thermo"]:
        thermo = NHCThermostat(p["temp"], timecon=p["timecon_thermo"], chain_vel0=p["chain_vel0"], chain_pos0=np.zeros(3), restart=True)
    if p["baro"]:
        baro = MTKBarostat(mmf, p["temp"], p["press"], timecon=p["timecon_baro"], vel_press0=p["vel_press0"], restart=True)
    if thermo is not None and baro is not None:
        hooks.append(TBCombination(thermo, baro))
    elif thermo is not None:
        hooks.append(thermo)
    verlet = VerletIntegrator(mmf, timestep=p["timestep"], hooks=hooks, vel0=vel0, ndof=ndof)
    assert verlet.device_mode
    opt = lambda name: int(lib.mm_get_option(part.handle, name))
    tiling = {"kernel": "k_march2" if opt(b"march2") == 1 else "k_march", "rows_per_thread": opt(b"rows_per_thread"),
              "warps": opt(b"tile_rows"), "shortest_chunk_planes": opt(b"chunk"), "blocks": opt(b"blocks"),
              "schedule_efficiency": opt(b"plan_efficiency_permille") / 1000.0,
              "images_on_load": opt(b"wrap_on_load"), "tail": {0: "separate reduction / scalar / halo launches", 1: "one tail launch", 2: "in the marching kernel"}[opt(b"tail")]}
    md = verlet._md

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident value -------------------------------------------------------------------------------
    # Warm-up in two calls.  The library replays pairs of lean steps as CUDA graphs from the second mm_md_run call on and
    # captures a pair the first time it meets a buffer parity: the last six warm-up steps are a call of their own that starts
    # at the parity the timed call starts at, so stream capture + cudaGraphInstantiate (5-90 ms of host time, depending on
    # the box: profiles/r02 calls ag - ai) happen here and not inside the timed region.
    nwarm = max(args.warmup, 9)
    _lib.check(lib.mm_md_run(md, nwarm - 6))
    _lib.check(lib.mm_md_run(md, 6))
    barrier()
    scal1 = np.zeros(_lib.S_COUNT)
    _lib.check(lib.mm_md_scalars(md, _lib.ptr(scal1)))
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = part.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    _lib.check(lib.mm_md_run(md, args.steps))
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = part.launches - launches0
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = nglobal * args.steps / (ms * 1e-3)
    scal = np.zeros(_lib.S_COUNT)
    _lib.check(lib.mm_md_scalars(md, _lib.ptr(scal)))  # raises on NaN: a diverged run is not a measurement
    rv_end = np.zeros(9)
    _lib.check(lib.mm_md_get_state(md, None, None, None, _lib.MM_HOST, _lib.ptr(rv_end), None, None, None))

    # ---- roofline of the dominant kernel (event-bracketed launches, separate short run) ----------------------------
    _lib.check(lib.mm_set_option(part.handle, b"profile", 1))
    nprof = max(3, min(args.steps, 10))
    _lib.check(lib.mm_md_run(md, nprof))
    nl, tot = (ctypes.c_int64 * 2)(), (ctypes.c_double * 2)()
    _lib.check(lib.mm_profile(part.handle, nl, tot))
    _lib.check(lib.mm_set_option(part.handle, b"profile", 0))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # measured FP64 FMA issue rate of a B200 SM (profiles/microbench/fp64_mio_probe_b200.txt): one warp-wide DFMA per
    # 2.168 cycles and scheduler -> 148 SMs x 4 schedulers x 32 lanes x 2 flop / 2.168 x 1.965 GHz = 34.3 TFLOP/s
    fp64_peak = 148 * 4 * 32 * 2 / 2.168 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
    # dominant kernel: the fused kick-drift-force-kick launch when the structured path is active (kind 1), else the
    # per-cell force kernel of the indexed path (kind 0).  Algorithmic bytes: SURVEY.md 8(d), per node and launch.
    fused = nl[1] > 0
    if fused:
        kernel_ms, name, bpn = tot[1] / nl[1], "k_march<STEP> (fused kick-drift-force-kick, structured grid)", BYTES_PER_NODE_STEP
    else:
        kernel_ms, name, bpn = tot[0] / max(nl[0], 1), "k_cells (per-cell force kernel, indexed topology)", BYTES_PER_NODE_EVAL
    achieved = bpn * nnodes / (kernel_ms * 1e-3) / 1e9
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture of this grid and ensemble
    # (single GPU), if any
    traffic = traffic_src = None
    caps = {}
    try:
        caps = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        cap = caps.get("%s_%s@%d" % ("k_march_step" if fused else "k_cells", args.ensemble, args.grid))
        if cap and world == 1 and args.model == "original":
            traffic = float(cap["dram_bytes_read"]) + float(cap["dram_bytes_write"])
            traffic_src = cap.get("kernel")
    except Exception:
        pass
    flops = FLOPS_PER_NODE_STEP * nnodes
    fp64_achieved = flops / (kernel_ms * 1e-3) / 1e12
    t_hbm, t_fp64 = bpn * nnodes / (peak * 1e9), flops / (fp64_peak * 1e12)
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "traffic_kernel": traffic_src,
        "kernel": name, "bytes_per_node": bpn, "kernel_ms": kernel_ms, "launches_timed": int(nl[1] if fused else nl[0]),
        "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s",
        # the path sits on the FP64 ridge - report against min(HBM, FP64) with the MEASURED FMA peak
        "fp64": {"peak_tflops": fp64_peak, "achieved_tflops": fp64_achieved, "frac": fp64_achieved / fp64_peak,
                 "flops_per_node": FLOPS_PER_NODE_STEP,
                 "peak_source": "measured DFMA issue rate, profiles/microbench/fp64_mio_probe_b200.txt"},
        "binding": "fp64" if t_fp64 > t_hbm else "hbm",
        "frac_of_binding_roofline": max(t_hbm, t_fp64) / (kernel_ms * 1e-3),
        # whole step against the survey's convention of 108 B per node and force evaluation
        "step_frac_of_108B_roofline": (BYTES_PER_NODE_STEP * FORCE_EVALS[args.ensemble] * nnodes * args.steps / (ms * 1e-3) / 1e9) / peak,
    }
    if fused and nl[0] > 0:  # the force-only launches of the barostat (52 B/node algorithmic)
        fms = tot[0] / nl[0]
        roofline["force_only_kernel"] = {"kernel": "k_march<FORCE>", "kernel_ms": fms, "bytes_per_node": BYTES_PER_NODE_EVAL,
                                         "achieved": BYTES_PER_NODE_EVAL * nnodes / (fms * 1e-3) / 1e9,
                                         "frac": BYTES_PER_NODE_EVAL * nnodes / (fms * 1e-3) / 1e9 / peak,
                                         "fp64_frac": FLOPS_PER_NODE_STEP * nnodes / (fms * 1e-3) / 1e12 / fp64_peak,
                                         "launches_timed": int(nl[0])}
        cap = caps.get("k_march_force_%s@%d" % (args.ensemble, args.grid))
        if cap and world == 1 and args.model == "original":
            roofline["force_only_kernel"]["traffic"] = float(cap["dram_bytes_read"]) + float(cap["dram_bytes_write"])

    # ---- e2e: host buffers in, one step, host buffers out, every step ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        hpos = torch.empty((nnodes, 3), dtype=torch.float64).pin_memory()
        hvel = torch.empty((nnodes, 3), dtype=torch.float64).pin_memory()
        hp, hv = ctypes.c_void_p(hpos.data_ptr()), ctypes.c_void_p(hvel.data_ptr())
        _lib.check(lib.mm_md_get_state(md, hp, hv, None, _lib.MM_HOST, None, None, None, None))
        esteps = max(2, min(args.steps, 5))
        sc = np.zeros(_lib.S_COUNT)
        barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            _lib.check(lib.mm_md_set_state(md, hp, hv, _lib.MM_HOST))
            _lib.check(lib.mm_md_run(md, 1))
            _lib.check(lib.mm_md_get_state(md, hp, hv, None, _lib.MM_HOST, None, None, None, None))
            _lib.check(lib.mm_md_scalars(md, _lib.ptr(sc)))
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": nglobal * esteps / dt, "unit": "node-steps/s", "h2d_bytes_per_step": 48 * nnodes,
               "d2h_bytes_per_step": 48 * nnodes + 8 * _lib.S_COUNT, "steps": esteps,
               "api": "mm_md_set_state(host) + mm_md_run(1) + mm_md_get_state(host) + mm_md_scalars per step"}

    # ---- CPU baseline (rank 0, N = 1 only) -------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        rate, dt, csteps = cpu_oracle_rate(args.cpu_grid, args.ensemble, args.model, 3, cores, target_seconds=12.0)
        cpu = {"value": rate, "unit": "node-steps/s", "cores": cores, "kind": "port", "reference_python": REFERENCE_PYTHON,
               "sample": "%d^3-cell fcu grid, %s, %d steps, OpenMP x%d, %.1f s" % (args.cpu_grid, args.ensemble.upper(), csteps, cores, dt)}

    if rank == 0:
        line = {
            "metric": "MD node-steps/s (force+Verlet, fp64)", "value": value, "unit": "node-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": nwarm, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world, comm_mode, tiling), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks,
            "force_evals_per_s": value * FORCE_EVALS[args.ensemble],
            # identical initial state at every N (global_fields): these agree across N = 1, 2, 4, 8 to ~1e-10.
            # econs_drift = |econs(after the timed steps) - econs(after the warm-up)| / ekin
            "check": {"temp_K": scal[_lib.S_TEMP], "epot": scal[_lib.S_EPOT], "ekin": scal[_lib.S_EKIN],
                      "econs": scal[_lib.S_ECONS], "rvecs": [float(x) for x in rv_end],
                      "econs_drift": abs(scal[_lib.S_ECONS] - scal1[_lib.S_ECONS]) / scal[_lib.S_EKIN],
                      "steps_total": int(scal[_lib.S_COUNTER])},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
