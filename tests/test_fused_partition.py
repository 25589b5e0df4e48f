"""Host logic of the fused kernels' static work partition (gdm_fused_partition; no GPU needed).

Every (tile, plane) of the output window must be covered exactly once, no CTA may be empty, the number of
CTAs must not exceed the SM slots, and the most expensive CTA (planes + 2p ramp planes per segment) must stay
close to the ideal share -- this is what removes the tail of the chunked launch (profiles/r1)."""
import ctypes as C

import numpy as np
import pytest


def partition(lib, aligned, tx, ty, z0, z1, slots, p):
    cap_p, cap_s = 4096, 16384
    ptr = (C.c_int32 * cap_p)()
    segs = (C.c_int32 * (4 * cap_s))()
    nc, ns = C.c_int32(), C.c_int32()
    rc = lib.gdm_fused_partition(aligned, tx, ty, z0, z1, slots, p, ptr, cap_p, segs, cap_s, C.byref(nc), C.byref(ns))
    assert rc == 0, lib.gdm_last_error()
    ptr = np.array(ptr[: nc.value + 1])
    segs = np.array(segs[: 4 * ns.value]).reshape(-1, 4)
    return ptr, segs


CASES = [
    # aligned, tiles_x, tiles_y, z0, z1, slots, p
    (1, 8, 8, 1, 256, 296, 3),     # BASELINE: 256^3 cells, 2 CTAs per SM
    (0, 8, 8, 1, 256, 296, 3),
    (1, 8, 8, 1, 256, 148, 3),
    (1, 8, 4, 0, 257, 148, 3),
    (1, 1, 1, 1, 12, 296, 3),      # tiny grids of the parity tests
    (1, 1, 2, 0, 9, 296, 1),
    (1, 2, 2, 3, 6, 296, 3),       # slab face window of the multi-GPU overlap (p planes)
    (1, 8, 8, 4, 252, 296, 3),     # interior window of the multi-GPU overlap
    (1, 16, 16, 1, 512, 296, 5),
    (0, 16, 16, 1, 512, 296, 5),
    (1, 3, 5, 0, 40, 7, 3),        # fewer slots than tiles
    (1, 8, 8, 1, 256, 592, 1),
]


@pytest.mark.parametrize("aligned,tx,ty,z0,z1,slots,p", CASES)
def test_partition_covers_every_plane_once(lib, aligned, tx, ty, z0, z1, slots, p):
    ptr, segs = partition(lib, aligned, tx, ty, z0, z1, slots, p)
    n_ctas = len(ptr) - 1
    assert 1 <= n_ctas <= slots
    assert ptr[0] == 0 and ptr[-1] == len(segs) and np.all(np.diff(ptr) >= 1)
    cover = np.zeros((ty, tx, z1 - z0), dtype=np.int32)
    for sx, sy, a, b in segs:
        assert 0 <= sx < tx and 0 <= sy < ty and z0 <= a < b <= z1
        cover[sy, sx, a - z0:b - z0] += 1
    assert cover.min() == 1 and cover.max() == 1


@pytest.mark.parametrize("aligned,tx,ty,z0,z1,slots,p", [c for c in CASES if c[3] + 8 * c[6] <= c[4]])
def test_partition_is_balanced(lib, aligned, tx, ty, z0, z1, slots, p):
    ptr, segs = partition(lib, aligned, tx, ty, z0, z1, slots, p)
    cost = np.array([sum(b - a + 2 * p for _, _, a, b in segs[ptr[i]:ptr[i + 1]]) for i in range(len(ptr) - 1)])
    work = tx * ty * (z1 - z0)
    n = min(slots, max(1, work // (2 * p)))
    ideal = work / n + 2 * p
    # the longest CTA decides the kernel time: within 35 % of the ideal share (ramps of split segments included)
    assert cost.max() <= 1.35 * ideal + 2 * p, (cost.max(), ideal)


def test_aligned_partition_cuts_all_tile_columns_at_the_same_planes(lib):
    ptr, segs = partition(lib, 1, 8, 8, 1, 256, 296, 3)
    full = segs[: 4 * 64]          # m = 296 // 64 = 4 full segments per tile column come first
    cuts = {(a, b) for _, _, a, b in full}
    assert len(cuts) == 4
    for a, b in cuts:
        assert sum(1 for s in full if s[2] == a and s[3] == b) == 64


def test_partition_rejects_small_buffers(lib):
    ptr = (C.c_int32 * 2)()
    segs = (C.c_int32 * 4)()
    nc, ns = C.c_int32(), C.c_int32()
    rc = lib.gdm_fused_partition(1, 8, 8, 1, 256, 296, 3, ptr, 2, segs, 1, C.byref(nc), C.byref(ns))
    assert rc != 0 and nc.value > 1 and ns.value > 1
