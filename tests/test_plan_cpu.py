"""Block schedule of the marching kernel (sg_plan_schedule, csrc/mm_structured.cu) - host logic, no device needed.

The schedule replaces the uniform (tiles x chunks) launch grid: every block reads one work item (tile, first plane, one
past the last).  Whatever the planner picks, the items must cover every (tile, owned plane) exactly once - a missed plane
would leave nodes without forces, a doubled one would double-count the energy - and must never do worse than the uniform
schedule it replaces.
"""

import ctypes

import numpy as np
import pytest

from micmec_b200 import _lib


def plan(ntx, nty, planes, nsm=148, images=0, uniform=0):
    lib = _lib.load()
    n = ctypes.c_int64(0)
    cost, ideal = ctypes.c_double(0.0), ctypes.c_double(0.0)
    _lib.check(lib.mm_plan_schedule(ntx, nty, planes, nsm, images, uniform, None, 0, ctypes.byref(n), ctypes.byref(cost), ctypes.byref(ideal)))
    items = np.zeros((n.value, 4), dtype=np.int32)
    _lib.check(lib.mm_plan_schedule(ntx, nty, planes, nsm, images, uniform, items.ctypes.data_as(ctypes.c_void_p), n.value,
                                    ctypes.byref(n), ctypes.byref(cost), ctypes.byref(ideal)))
    return items, cost.value, ideal.value


def coverage(items, ntx, nty, planes):
    cover = np.zeros((nty, ntx, planes + 2), dtype=np.int64)
    for bx, by, lo, hi in items:
        assert 0 <= bx < ntx and 0 <= by < nty
        assert 1 <= lo < hi <= planes + 1
        cover[by, bx, lo:hi] += 1
    return cover


SHAPES = [(9, 19, 256), (9, 19, 128), (9, 19, 64), (9, 19, 32), (9, 43, 32), (1, 1, 2), (1, 1, 5), (2, 3, 7), (3, 5, 16),
          (5, 10, 128), (3, 5, 64), (2, 2, 64), (12, 12, 3), (40, 40, 2), (9, 19, 255), (7, 11, 97)]


@pytest.mark.parametrize("images", [0, 1])
@pytest.mark.parametrize("shape", SHAPES)
def test_every_plane_of_every_tile_once(shape, images):
    ntx, nty, planes = shape
    items, cost, ideal = plan(ntx, nty, planes, images=images)
    cover = coverage(items, ntx, nty, planes)
    assert (cover[:, :, 1 : planes + 1] == 1).all()
    assert (cover[:, :, 0] == 0).all() and (cover[:, :, planes + 1] == 0).all()
    assert cost >= ideal > 0.0


@pytest.mark.parametrize("shape", SHAPES)
def test_not_worse_than_uniform_chunks(shape):
    ntx, nty, planes = shape
    _, cost, _ = plan(ntx, nty, planes)
    for chunk in sorted({planes, max(2, planes // 2), max(2, planes // 4), max(2, planes // 6), 8, 16, 43}):
        if chunk > planes:
            continue
        nchunks = -(-planes // chunk)
        if planes // nchunks < 2:  # the planner itself never cuts below two planes per block
            continue
        items, ucost, _ = plan(ntx, nty, planes, uniform=chunk)
        assert (coverage(items, ntx, nty, planes)[:, :, 1 : planes + 1] == 1).all()
        assert cost <= ucost * (1.0 + 1e-9), (chunk, cost, ucost)


def test_headline_grids():
    """256^3 nodes, tiles of 30 x 14 owned nodes (9 x 19 tiles): one GPU, and the slabs of 2 / 4 / 8 GPUs."""
    eff = {}
    for planes, images in ((256, 0), (128, 1), (64, 1), (32, 1)):
        items, cost, ideal = plan(9, 19, planes, images=images)
        eff[planes] = ideal / cost
        assert len(items) < 4096
    assert eff[256] > 0.95 and eff[128] > 0.93
    assert eff[64] > 0.85 and eff[32] > 0.80, eff


def test_edge_tiles_last_with_images_on_load():
    items, _, _ = plan(9, 19, 128, images=1)
    first = items[:50]
    edge = (first[:, 0] == 0) | (first[:, 0] == 8) | (first[:, 1] == 0) | (first[:, 1] == 18)
    assert not edge.any()
    items0, _, _ = plan(9, 19, 128, images=0)
    assert tuple(items0[0][:2]) == (0, 0) and tuple(items0[1][:2]) == (1, 0)  # natural order, x fastest


def test_rejects_bad_arguments():
    lib = _lib.load()
    n = ctypes.c_int64(0)
    assert lib.mm_plan_schedule(0, 1, 4, 148, 0, 0, None, 0, ctypes.byref(n), None, None) != 0
    items = np.zeros((1, 4), dtype=np.int32)
    assert lib.mm_plan_schedule(9, 19, 64, 148, 0, 0, items.ctypes.data_as(ctypes.c_void_p), 1, ctypes.byref(n), None, None) != 0
