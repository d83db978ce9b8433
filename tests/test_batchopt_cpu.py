"""Host logic of the lockstep replica optimiser (micmec_b200/sampling/batchopt.py): every replica of a batch must
reach what the single-system QNOptimizer (itself pinned to the reference, test_opt_cpu.py) reaches for that replica.
Forces come from the CPU oracle here (test infrastructure); tests/test_batchopt_gpu.py uses ReplicaBatch."""
import numpy as np
import pytest

import goldenio as gio
from oraclepart import OracleForcePart
from test_force_gpu import make_system


class OracleReplicaEvaluator(object):
    """``ReplicaBatch.compute`` signature on top of the oracle, one replica after the other."""

    def __init__(self, systems):
        from oracle import oracle as orc

        self.oracles = [orc.Oracle(s) for s in systems]

    def compute(self, pos, rvecs, gpos=True, vtens=True):
        res = [o.compute(pos[r], rvecs[r], gpos=gpos, vtens=vtens) for r, o in enumerate(self.oracles)]
        return (np.array([e for e, _, _ in res]), np.stack([g for _, g, _ in res]) if gpos else None,
                np.stack([v for _, _, v in res]) if vtens else None)


def replicas(fixtures, nper, amp, strain=None):
    systems = []
    for name in fixtures:
        d = gio.load("force_" + name)
        for k in range(nper):
            system = make_system(d)
            rng = np.random.default_rng(100 + k)
            if strain is not None:
                a = np.eye(3) + strain * rng.uniform(-1, 1, (3, 3))
                a = 0.5 * (a + a.T)
                system.domain.update_rvecs(np.ascontiguousarray(np.array(system.domain.rvecs) @ a))
                system.pos[:] = system.pos @ a
            system.pos[:] = system.pos + amp * rng.standard_normal(system.pos.shape)
            systems.append(system)
    return systems


def single_runs(systems, kind, kwargs):
    from micmec_b200.pes.mmff import MicMecForceField
    from micmec_b200.sampling.dof import CartesianDOF, StrainCellDOF
    from micmec_b200.sampling.opt import QNOptimizer

    out = []
    for system in systems:
        mmf = MicMecForceField(system, [OracleForcePart(system)])
        dof = (CartesianDOF if kind == "cartesian" else StrainCellDOF)(mmf, **kwargs)
        opt = QNOptimizer(dof)
        opt.run(200)
        assert dof.converged
        out.append((opt.f, system.pos.copy(), np.array(system.domain.rvecs), opt.counter))
    return out


@pytest.mark.parametrize("kind", ["cartesian", "strain"])
def test_replica_optimiser_matches_single_system_runs(kind):
    from micmec_b200.sampling.batchopt import ReplicaQNOptimizer

    if kind == "cartesian":
        kwargs = dict(gpos_rms=1e-7, dpos_rms=1e-5)
        systems = replicas(["3x3x3_conf0", "3x3x3_conf3", "3x3x3_conf9"], 2, 0.5)
    else:
        kwargs = dict(gpos_rms=1e-8, dpos_rms=1e-6, grvecs_rms=1e-8, drvecs_rms=1e-6)
        systems = replicas(["3x3x3_test", "3x3x3_conf0"], 2, 0.3, strain=0.02)
    pos0 = np.stack([s.pos for s in systems])
    rvecs0 = np.stack([np.array(s.domain.rvecs) for s in systems])
    opt = ReplicaQNOptimizer(OracleReplicaEvaluator(systems), pos0, rvecs0, dof=kind, eigh="lapack", **kwargs)
    sweeps = opt.run(300)
    assert opt.converged.all() and not opt.failed.any()
    assert opt.evaluations == sweeps + 1  # one batched force call per sweep
    ref = single_runs(replicas(*({"cartesian": (["3x3x3_conf0", "3x3x3_conf3", "3x3x3_conf9"], 2, 0.5),
                                  "strain": (["3x3x3_test", "3x3x3_conf0"], 2, 0.3, 0.02)}[kind])), kind, kwargs)
    pos, rvecs = opt.pos, opt.rvecs
    for r, (f, p, rv, count) in enumerate(ref):
        # accepted steps = propagate calls + the step taken inside initialize
        assert abs(int(opt.iterations[r]) - (count + 1)) <= 6, (r, opt.iterations[r], count)  # the last steps run at the rounding floor of the energy differences
        scale = np.sqrt(np.mean((p - pos0[r]) ** 2))
        assert np.max(np.abs(pos[r] - p)) <= 1e-4 * scale, r
        assert np.max(np.abs(rvecs[r] - rv)) <= 1e-6 * np.sqrt(np.mean(rv ** 2)), r
        assert abs(opt.f[r] - f) <= 1e-8 * abs(float(opt.f_old.max())) + 1e-12, r


def test_solve_trust_radius_batch_matches_scalar_version():
    from micmec_b200.sampling.batchopt import solve_trust_radius_batch
    from micmec_b200.sampling.opt import solve_trust_radius

    rng = np.random.default_rng(4)
    evals = rng.normal(0.5, 1.0, (50, 15))
    evals[:10] = abs(evals[:10]) + 0.1
    grad = rng.normal(0.0, 1.0, (50, 15))
    radius = 10.0 ** rng.uniform(-3, 1.0, 50)
    steps = solve_trust_radius_batch(grad, evals, radius)
    for r in range(50):
        ref = solve_trust_radius(grad[r], evals[r], radius[r])
        assert np.allclose(steps[r], ref, rtol=1e-9, atol=1e-12), r


def test_shard_covers_everything_once():
    from micmec_b200.sampling.batchopt import shard

    items = list(range(10243))
    parts = [shard(items, r, 8) for r in range(8)]
    assert sum(parts, []) == items and max(map(len, parts)) - min(map(len, parts)) <= 1


def _replica_worker(rank, world, port, out_dir):
    """One rank of a 'replicas only' run: optimise the own share, gather the energies on rank 0 (gloo)."""
    import os

    import torch.distributed as dist
    from micmec_b200.sampling.batchopt import ReplicaQNOptimizer, shard

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    systems = replicas(["3x3x3_conf0", "3x3x3_conf3", "3x3x3_conf9"], 2, 0.4)
    mine = shard(systems, rank, world)
    pos0 = np.stack([s.pos for s in mine])
    rvecs0 = np.stack([np.array(s.domain.rvecs) for s in mine])
    opt = ReplicaQNOptimizer(OracleReplicaEvaluator(mine), pos0, rvecs0, dof="cartesian", eigh="lapack", gpos_rms=1e-7, dpos_rms=1e-5)
    opt.run(300)
    parts = [None] * world
    dist.all_gather_object(parts, (opt.f.tolist(), opt.converged.tolist()))
    if rank == 0:
        np.save(os.path.join(out_dir, "f.npy"), np.array(sum((p[0] for p in parts), [])))
        np.save(os.path.join(out_dir, "conv.npy"), np.array(sum((p[1] for p in parts), [])))
    dist.barrier()
    dist.destroy_process_group()


def test_replicas_spread_over_two_ranks_match_the_single_process_run(tmp_path):
    import socket

    import torch.multiprocessing as mp
    from micmec_b200.sampling.batchopt import ReplicaQNOptimizer

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_replica_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    systems = replicas(["3x3x3_conf0", "3x3x3_conf3", "3x3x3_conf9"], 2, 0.4)
    pos0 = np.stack([s.pos for s in systems])
    rvecs0 = np.stack([np.array(s.domain.rvecs) for s in systems])
    one = ReplicaQNOptimizer(OracleReplicaEvaluator(systems), pos0, rvecs0, dof="cartesian", eigh="lapack", gpos_rms=1e-7, dpos_rms=1e-5)
    one.run(300)
    f = np.load(str(tmp_path / "f.npy"))
    assert np.load(str(tmp_path / "conv.npy")).all() and one.converged.all()
    # replicas never interact: the split run reproduces the single-process energies (same arithmetic per replica)
    assert np.max(np.abs(f - one.f)) <= 1e-12 * np.max(np.abs(one.f_old)) + 1e-14
