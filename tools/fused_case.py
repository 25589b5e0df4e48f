"""Run one fused-kernel parity case in its own process (a device fault kills the CUDA context).
usage: python tools/fused_case.py P NX NY NZ BC KIND [LZ]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np

p, nx, ny, nz, bc, kind = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], sys.argv[6]
if len(sys.argv) > 7:
    os.environ["GDM_FUSED_LZ"] = sys.argv[7]
import gdm_b200 as g
from helpers import make_pair, make_operator, oracle_operator, rel_err

gs, gc, os_, oc = make_pair(3, p, 1, [nx, ny, nz], bc)
b = [1.0, 0.15, -0.05]
A = make_operator(gs, gc, kind, b=b, kernel=g.capi.KERNEL_FUSED)
Ao = oracle_operator(os_, oc, kind, b=b)
xh = np.random.default_rng(0).uniform(-1, 1, gs.n_dofs())
x, y = g.Vector(gs, xh), g.Vector(gs)
y.set(7.0)
A.vmult(y, x)
ref = Ao @ xh
e1 = rel_err(y.numpy(), ref)
y2 = g.Vector(gs, xh)
A.vmult_add(y2, x)
e2 = rel_err(y2.numpy(), ref + xh)
print("CASE", *sys.argv[1:], "err", e1, e2, "OK" if max(e1, e2) <= 1e-12 else "FAIL")
