#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/cg_only.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import gdm_b200 as g
ctx = g.default_context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
n, p = 256, 3
s = g.System(3, p, 1); s.subdivided_hyper_cube(n)
c = g.AffineConstraints(); s.make_zero_boundary_constraints(c); c.close()
A = g.SparseMatrix(); g.MatrixCreator.create_laplace_matrix(g.MappingQ1(), s, g.QGauss(p + 1), A, c)
b = g.Vector(s); b.set(1.0); c.set_zero(b); u = g.Vector(s)
for pre in (g.PreconditionIdentity(),):
    ctl = g.ReductionControl(12, 1e-30, 1e-30)
    try: g.SolverCG(ctl).solve(A, u, b, pre)
    except g.NoConvergence: pass
torch.cuda.synchronize()
PY
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cg.csv python /tmp/cg_only.py > gpurun_out/ncu_cg.log 2>&1
wc -l gpurun_out/launches_cg.csv
