"""Operator-apply throughput table (GDoF/s and fraction of the HBM roofline) for several operators.
usage: python tools/bench_ops.py [--cells 256] [--steps 50]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import gdm_b200 as g

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=256)
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--p", type=int, nargs="*", default=[1, 3, 5])
args = ap.parse_args()
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
ctx = g.default_context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
rows = []
for p in args.p:
    n = args.cells
    s = g.System(3, p, 1)
    s.subdivided_hyper_cube(n)
    c = g.AffineConstraints()
    s.make_zero_boundary_constraints(c)
    c.close()
    x, y = g.Vector(s, np.random.default_rng(0).uniform(-1, 1, s.n_dofs())), g.Vector(s)
    for kind in ("mass", "stiffness", "advection"):
        for kernel, kname in ((g.capi.KERNEL_FUSED, "fused"), (g.capi.KERNEL_GENERIC, "generic")):
            A = g.SparseMatrix()
            m, q = g.MappingQ1(), g.QGauss(p + 1)
            if kind == "mass":
                g.MatrixCreator.create_mass_matrix(m, s, q, A, c, kernel=kernel)
            elif kind == "stiffness":
                g.MatrixCreator.create_laplace_matrix(m, s, q, A, c, kernel=kernel)
            else:
                g.MatrixCreator.create_advection_matrix(m, s, q, A, c, [1.0, 0.15, -0.05], kernel=kernel)
            for _ in range(5):
                A.vmult(y, x)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                A.vmult(y, x)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            gd = s.n_dofs() / ms / 1e6
            rows.append({"p": p, "op": kind, "kernel": kname, "cells": n, "ms": round(ms, 4), "gdofs": round(gd, 1),
                         "frac_hbm": round(16 * gd / peak, 3)})
            print(json.dumps(rows[-1]), flush=True)
