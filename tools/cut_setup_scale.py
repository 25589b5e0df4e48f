"""Host-side scaling of the cut-cell set-up (gdm_cut_poisson_create) towards BASELINE configuration 5:
unit sphere in [-1.21, 1.21]^3, p = 3, ghost penalty, rows of ONE of n_ranks slabs.  No GPU needed.
  python tools/cut_setup_scale.py 48 96 128 [--ranks 8]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gdm_b200 as g  # noqa: E402


def main():
    argv = sys.argv[1:]
    n_ranks = 1
    if "--ranks" in argv:
        i = argv.index("--ranks")
        n_ranks = int(argv[i + 1])
        del argv[i:i + 2]
    args = argv
    for n1 in [int(a) for a in args] or [48]:
        ax = -1.21 + np.arange(n1 + 1) * (2.42 / n1)
        z, y, x = np.meshgrid(ax, ax, ax, indexing="ij")
        ls = (np.sqrt(x * x + y * y + z * z) - 1.0).ravel()
        plane = (n1 + 1) ** 2
        stride = -(-n1 // n_ranks)
        rank = n_ranks // 2  # a middle slab: the one with the most surface
        b = 0 if rank == 0 else (stride * rank + 1) * plane
        e = min((stride * (rank + 1) + 1), n1 + 1) * plane
        t = time.time()
        c = g.CutPoisson(3, 3, [n1] * 3, [-1.21] * 3, [1.21] * 3, ls, ghost_penalty=True,
                         row_range=(b, e) if n_ranks > 1 else None)
        dt = time.time() - t
        n_rows, nnz, n_id, cells = c.sizes()
        print(json.dumps({"n_subdivisions": n1, "n_ranks": n_ranks, "rank": rank, "seconds": round(dt, 2),
                          "rows": n_rows, "identity_rows": n_id, "nnz": nnz, "cells_inside_outside_cut": cells,
                          "overlay_bytes": 12 * nnz + 16 * n_rows, "threads": os.cpu_count()}), flush=True)


if __name__ == "__main__":
    main()
