"""Host logic of the persistent fused kernel's work partition (gdm_pers_partition; no GPU needed).

Every (tile, input plane) must be covered exactly once; a job that hands partial planes to the job below it must be
at least 2p planes long; seams must pair the job ending at plane b with the job of the same tile starting at b; a
share may only wait for a share with a larger index (the ticket order of the kernel makes that deadlock free); the
longest share must stay close to the ideal one (no tail, no ramp: DESIGN.md section 5)."""
import ctypes as C

import numpy as np
import pytest

MAXJ = 8  # not a limit of the kernel (jobs are read from global memory); keeps shares compact


def partition(lib, tx, ty, k0, k1, slots, min_len, aligned, weights=None):
    cap_p, cap_j = 8192, 16384
    w = None if weights is None else (C.c_int32 * len(weights))(*[int(v) for v in weights])
    ptr = (C.c_int32 * cap_p)()
    jobs = (C.c_int32 * (6 * cap_j))()
    ns, nj = C.c_int32(), C.c_int32()
    rc = lib.gdm_pers_partition(tx, ty, k0, k1, slots, min_len, aligned, w, ptr, cap_p, jobs, cap_j, C.byref(ns), C.byref(nj))
    assert rc == 0, lib.gdm_last_error()
    return np.array(ptr[: ns.value + 1]), np.array(jobs[: 6 * nj.value]).reshape(-1, 6)


CASES = [
    # tiles_x, tiles_y, k0, k1, slots, p, aligned
    (8, 8, 1, 256, 296, 3, 1),     # BASELINE: 256^3 cells, 2 CTAs per SM
    (8, 8, 1, 256, 296, 3, 0),
    (8, 8, 1, 256, 148, 3, 1),
    (8, 4, 0, 257, 148, 3, 1),
    (1, 1, 1, 12, 296, 3, 1),      # tiny grids of the parity tests
    (1, 2, 0, 9, 296, 1, 1),
    (2, 2, 0, 9, 296, 3, 1),       # slab face window of the multi-GPU overlap (3p input planes)
    (8, 8, 1, 255, 296, 3, 1),     # interior window of the multi-GPU overlap
    (16, 16, 0, 513, 296, 5, 1),
    (16, 16, 0, 513, 296, 5, 0),
    (3, 5, 0, 40, 7, 3, 1),        # fewer slots than tiles
    (8, 8, 1, 256, 592, 1, 1),
    (8, 8, 0, 68, 296, 3, 1),      # 8-GPU slab
    (17, 17, 0, 131, 148, 5, 0),
    # guided levels (mode 2): one share per (tile, level), short levels at the bottom are handed out last
    (8, 8, 1, 256, 296, 3, 2),
    (8, 11, 1, 256, 296, 3, 2),
    (16, 16, 0, 513, 148, 5, 2),
    (2, 2, 0, 9, 296, 3, 2),
    (3, 5, 0, 40, 7, 3, 2),
    (8, 8, 0, 68, 296, 1, 2),
]


def edge_weights(tx, ty, wx=1390, wy=1330, wxy=1480):
    w = np.full((ty, tx), 1000)
    w[:, [0, -1]] = wx
    w[[0, -1], :] = wy
    for j in (0, -1):
        for i in (0, -1):
            w[j, i] = wxy
    return w.reshape(-1)


@pytest.mark.parametrize("tx,ty,k0,k1,slots,p,aligned", CASES)
@pytest.mark.parametrize("weighted", [False, True])
def test_partition(lib, tx, ty, k0, k1, slots, p, aligned, weighted):
    min_len = 2 * p
    wts = edge_weights(tx, ty) if weighted else np.full(tx * ty, 1000)
    ptr, jobs = partition(lib, tx, ty, k0, k1, slots, min_len, aligned, wts if weighted else None)
    n_shares = len(ptr) - 1
    assert 1 <= n_shares and (n_shares <= slots or aligned == 2)  # guided: more shares than CTAs (self-scheduling)
    assert ptr[0] == 0 and ptr[-1] == len(jobs) and np.all(np.diff(ptr) >= 1) and np.all(np.diff(ptr) <= MAXJ + 8)
    cover = np.zeros((tx * ty, k1 - k0), dtype=int)
    share_of = np.repeat(np.arange(n_shares), np.diff(ptr))
    start = {}
    for j, (jx, jy, a, b, lo, hi) in enumerate(jobs):
        assert 0 <= jx < tx and 0 <= jy < ty and k0 <= a < b <= k1
        cover[jy * tx + jx, a - k0:b - k0] += 1
        start[(jy * tx + jx, a)] = j
    assert np.all(cover == 1)
    for j, (jx, jy, a, b, lo, hi) in enumerate(jobs):
        t = jy * tx + jx
        if a > k0:
            assert lo == j and b - a >= 2 * p          # hands 2p partial planes down
        else:
            assert lo == -1
        if b < k1:
            assert hi == start[(t, b)]
            assert share_of[hi] > share_of[j]           # only waits for a share that took its ticket earlier
        else:
            assert hi == -1
    cost = np.array([sum((jobs[j][3] - jobs[j][2]) * wts[jobs[j][1] * tx + jobs[j][0]] for j in range(ptr[w], ptr[w + 1]))
                     for w in range(n_shares)]) / 1000.0
    ideal = wts.sum() / 1000.0 * (k1 - k0) / min(slots, max(1, tx * ty * (k1 - k0) // (2 * min_len)))
    if k1 - k0 >= 4 * min_len and slots >= tx * ty and aligned != 2:
        assert cost.max() <= 1.06 * ideal + min_len, (cost.max(), ideal)


def test_guided_levels_shrink_towards_the_bottom(lib):
    """Guided partition of the BASELINE grid: every column is cut at the same planes, the levels get shorter from the top
    to the bottom (the bottom ones are dispensed last), none shorter than 8 planes, the top one about one ideal share."""
    ptr, jobs = partition(lib, 8, 8, 1, 256, 296, 6, 2)
    lv = sorted({(a, b) for (_, _, a, b, _, _) in jobs})
    lens = [b - a for a, b in lv]
    assert lv[0][0] == 1 and lv[-1][1] == 256 and all(lv[i][1] == lv[i + 1][0] for i in range(len(lv) - 1))
    assert all(x <= y for x, y in zip(lens, lens[1:])) and min(lens) >= 8
    assert 45 <= lens[-1] <= 62 and len(jobs) == 64 * len(lv)
