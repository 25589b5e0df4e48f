#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/gen_only.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import gdm_b200 as g
ctx = g.default_context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
n, p = 256, 3
s = g.System(3, p, 1); s.subdivided_hyper_cube(n)
c = g.AffineConstraints(); s.make_zero_boundary_constraints(c); c.close()
x, y = g.Vector(s, np.random.default_rng(0).uniform(-1, 1, s.n_dofs())), g.Vector(s)
for kind in ("mass", "stiffness"):
    A = g.SparseMatrix()
    if kind == "mass":
        g.MatrixCreator.create_mass_matrix(g.MappingQ1(), s, g.QGauss(p + 1), A, c, kernel=g.capi.KERNEL_GENERIC)
    else:
        g.MatrixCreator.create_laplace_matrix(g.MappingQ1(), s, g.QGauss(p + 1), A, c, kernel=g.capi.KERNEL_GENERIC)
    for _ in range(3):
        A.vmult(y, x)
torch.cuda.synchronize()
PY
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct -k regex:band --clock-control none -c 40 --csv --log-file gpurun_out/launches_generic.csv python /tmp/gen_only.py > gpurun_out/ncu_gen.log 2>&1
wc -l gpurun_out/launches_generic.csv
