"""Host-side logic of the z-slab decomposition on CPU: layout maps, slab systems, gather over gloo (world size 2)."""
import os
import socket

import numpy as np
import pytest


def test_layout_partitions_the_grid():
    from micmec_b200.slab import SlabLayout

    shape = (5, 4, 12)
    seen = np.zeros(np.prod(shape), dtype=int)
    for count in (1, 2, 3, 4, 6):
        seen[:] = 0
        for rank in range(count):
            lay = SlabLayout(shape, rank, count)
            ids = lay.global_ids()
            assert len(ids) == lay.nnodes_local == 5 * 4 * 12 // count
            seen[ids] += 1
            # local order = reference order of a (nx, ny, nzl) grid
            k, l, m = np.unravel_index(np.arange(lay.nnodes_local), lay.local_shape)
            assert np.array_equal(ids, (k * 4 + l) * 12 + (m + lay.m0))
            assert lay.up == (rank + 1) % count and lay.down == (rank - 1) % count
        assert np.all(seen == 1)
    with pytest.raises(ValueError):
        SlabLayout(shape, 0, 5)
    with pytest.raises(ValueError):
        SlabLayout(shape, 3, 3)


def test_local_system_matches_cut_of_the_global_grid():
    from micmec_b200.slab import SlabLayout, local_system
    from micmec_b200.system import System
    from micmec_b200.celltypes import TYPE_FCU

    shape = (4, 3, 8)
    full = System.periodic_grid(shape, TYPE_FCU, explicit=False)
    for rank in range(4):
        lay = SlabLayout(shape, rank, 4)
        loc = local_system(lay, TYPE_FCU)
        assert np.allclose(loc.pos, lay.take(full.pos), rtol=0, atol=1e-12)
        assert np.array_equal(np.array(loc.domain.rvecs), np.array(full.domain.rvecs))  # GLOBAL domain vectors
        assert loc.structured_shape == (4, 3, 2) and loc.nnodes == 24
        cut = local_system(lay, TYPE_FCU, pos=full.pos + 1.0, masses=full.masses * 2)
        assert np.array_equal(cut.pos, lay.take(full.pos + 1.0)) and np.all(cut.masses == 2 * TYPE_FCU["mass"])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shape, out_dir):
    import torch.distributed as dist
    from micmec_b200.slab import SlabLayout, gather_nodes

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lay = SlabLayout(shape, rank, world)
    glob = np.arange(np.prod(shape) * 3, dtype=float).reshape(-1, 3)
    gathered = gather_nodes(lay, lay.take(glob))
    scal = gather_nodes(lay, lay.take(glob[:, 0]))
    if rank == 0:
        np.save(os.path.join(out_dir, "ok.npy"), np.array([np.array_equal(gathered, glob), np.array_equal(scal, glob[:, 0])]))
    else:
        assert gathered is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_over_gloo_world_size_2(tmp_path):
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, _free_port(), (3, 4, 6), str(tmp_path)), nprocs=2, join=True)
    assert np.load(os.path.join(str(tmp_path), "ok.npy")).all()


def test_bench_initial_state_is_identical_for_every_decomposition():
    """bench.py draws ONE global displacement / velocity field and every rank keeps its z-slab of it, so N = 1, 2, 4 and 8
    integrate the bit-identical initial state (multi-GPU parity must be visible in the driver's records)."""
    import bench
    from micmec_b200.slab import SlabLayout

    grid = 16
    dpos, vel = bench.global_fields(grid)
    assert abs(vel.mean(axis=0)).max() < 1e-18 + 1e-12 * abs(vel).max()
    for world in (2, 4, 8):
        got_p, got_v = np.zeros_like(dpos), np.zeros_like(vel)
        for rank in range(world):
            layout = SlabLayout((grid,) * 3, rank, world)
            ids = layout.global_ids()
            d2, v2 = bench.global_fields(grid)  # what that rank would draw
            got_p[ids], got_v[ids] = d2[ids], v2[ids]
        assert np.array_equal(got_p, dpos) and np.array_equal(got_v, vel)


def test_oracle_grid_builder_matches_the_package_generator():
    """oracle.periodic_grid_system (NumPy only, used by the CPU arm of bench.py) against System.periodic_grid and the
    reference-derived image table, including 2-wide axes."""
    from oracle import oracle as orc
    from micmec_b200.celltypes import TYPE_FCU
    from micmec_b200.system import System

    for shape in [(4, 3, 5), (2, 2, 2), (3, 2, 4)]:
        arrays, pos, masses, rvecs = orc.periodic_grid_system(shape, TYPE_FCU)
        system = System.periodic_grid(shape, TYPE_FCU, explicit=True)
        assert np.array_equal(arrays["surrounding_nodes"], system.surrounding_nodes)
        assert np.array_equal(arrays["surrounding_cells"], system.surrounding_cells)
        assert np.array_equal(arrays["shift"], orc.cell_shifts(system.grid, system.surrounding_nodes, True))
        assert np.array_equal(pos, system.pos) and np.array_equal(masses, system.masses)
        assert np.array_equal(rvecs, np.array(system.domain.rvecs))


def test_bench_clock_sampler_degrades_without_nvml():
    """bench.py samples SM clocks through NVML during the timed region; NVML is initialised in the constructor (outside
    the region).  Where there is no driver (this container) the sampler must neither raise nor hang, and say why it
    has no samples."""
    import bench

    sampler = bench.ClockSampler(0)
    sampler.start()
    out = sampler.stop()
    assert set(out) == {"sm_mhz", "sm_max_mhz", "reasons", "samples"}
    if out["samples"] == 0:
        assert out["sm_mhz"] is None
        assert any(r.startswith("unavailable") for r in out["reasons"])
