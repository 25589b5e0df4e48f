"""End-to-end apply through host buffers only (the `e2e` leg of bench.py): ms per gdm_operator_vmult_host call."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import gdm_b200 as g

n, p = 256, 3
ctx = g.default_context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
s = g.System(3, p, 1, context=ctx)
s.subdivided_hyper_cube(n)
c = g.AffineConstraints()
s.make_zero_boundary_constraints(c)
c.close()
A = g.SparseMatrix()
g.MatrixCreator.create_laplace_matrix(g.MappingQ1(), s, g.QGauss(p + 1), A, c)
nd = s.n_dofs()
hx = torch.empty(nd, dtype=torch.float64).pin_memory()
hy = torch.empty(nd, dtype=torch.float64).pin_memory()
hx.copy_(torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, nd)))
a, b = hx.numpy(), hy.numpy()
for _ in range(3):
    A.vmult_host(b, a)
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 20
for _ in range(reps):
    A.vmult_host(b, a)
    a, b = b, a
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / reps * 1e3
print(json.dumps({"chunks": os.environ.get("GDM_HOST_CHUNKS", "default"), "ms_per_call": ms, "gdofs": nd / ms / 1e6}))
