#!/usr/bin/env python
"""Benchmark of the GDM hot path: matrix-free stiffness apply (FP64, 3D, p=3) on 256^3 cells.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one operator application y = K x over the whole grid (BASELINE.json config 2:
"3D Poisson GDM p=3 on 256^3 cells, matrix-free stiffness apply + CG, 1 B200").  Vectors ping-pong
(x -> y -> x ...), 2 x 136 MB per GPU, i.e. larger than L2.  N > 1 (torchrun, one rank per GPU):
weak scaling, every rank owns a 256 x 256 x 256-cell slab of a 256 x 256 x (256 N) grid, ghost planes
exchanged over NCCL before each apply.  `value` = owned DoFs of all ranks x K / max-over-ranks device
time.  `e2e` = the same apply through the host-buffer entry point gdm_operator_vmult_host (pinned
host memory -> H2D -> apply -> D2H).  `cpu_baseline` / `--impl reference`: the reference's CPU path
(assembled CSR + SpMV, OpenMP over all host cores) restated in oracle/csr_baseline.c, timed on a
bounded sample (96^3 cells) because the CSR of 256^3 would need ~70 GB.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gdm_stiffness_apply_3d_p3_fp64"
UNIT = "GDoF/s"
ALG_BYTES_PER_DOF = 16.0  # read x + write y


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled during the timed region (NVML; nvidia-smi as fallback)."""

    def __init__(self, index=0, interval=0.005):
        super().__init__(daemon=True)
        self.index, self.interval, self.samples, self._stop_evt = index, interval, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def _sample(self):
        if self.nvml is not None:
            n = self.nvml
            sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
            try:
                r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
            except Exception:
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            names = []
            for bit, name in ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap")):
                if r & bit:
                    names.append(name)
            return float(sm), float(self.max_sm), names
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                      text=True, timeout=5)
        t = [x.strip() for x in out.strip().split(",")]
        names = [nm for nm, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], t[2:6])
                 if v.lower().startswith("active")]
        return float(t[0]), float(t[1]), names

    def run(self):
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self._sample())
            except Exception:
                pass
            self._stop_evt.wait(self.interval)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm = sorted(s[0] for s in self.samples)
        reasons = sorted({r for s in self.samples for r in s[2]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((s[1] for s in self.samples), default=None),
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def cpu_csr_sample(n_cells, p, steps, warmup):
    """The reference's CPU path on a bounded sample: assembled CSR SpMV, all host cores."""
    import numpy as np
    import oracle as O
    from oracle.csr_baseline import CsrOperator, num_threads, set_num_threads
    set_num_threads(os.cpu_count() or 1)  # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    s = O.System(3, p)
    s.subdivided_hyper_cube(n_cells)
    c = O.Constraints()
    s.make_zero_boundary_constraints(c)
    c.close()
    A = CsrOperator(s, c, "stiffness")
    x = np.random.default_rng(0).uniform(-1, 1, A.n_rows)
    y = np.zeros_like(x)
    for _ in range(warmup):
        A.vmult(y, x)
        x, y = y, x
    t0 = time.perf_counter()
    for _ in range(steps):
        A.vmult(y, x)
        x, y = y, x
    dt = time.perf_counter() - t0
    return {"value": A.n_rows * steps / dt / 1e9, "unit": UNIT, "cores": num_threads(), "kind": "port",
            "sample": f"{steps} CSR SpMV on {n_cells}^3 cells p={p} ({A.n_rows} DoFs, {A.nnz} nnz, int32 col + fp64 val), "
                      f"OpenMP static row slabs; nproc={os.cpu_count()}",
            "ms_per_step": dt / steps * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(args.steps, 1), max(args.warmup, 1)
    base = cpu_csr_sample(args.cpu_cells, args.p, min(steps, 50), min(warmup, 5))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": min(steps, 50), "warmup": min(warmup, 5), "ms_per_step": base["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"3D Poisson GDM p={args.p} stiffness apply; reference CPU path (assembled CSR SpMV) "
                                   f"on a {args.cpu_cells}^3-cell sample of the {args.cells}^3-cell grid"},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def run_ours(args):
    import numpy as np
    import torch
    import gdm_b200 as g

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    ctx = g.init_distributed(local) if world > 1 else g.default_context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)

    n, p = args.cells, args.p
    sys_ = g.System(3, p, 1, comm="world" if world > 1 else None, context=ctx)
    sys_.subdivided_hyper_rectangle([n, n, n * world], [0.0, 0.0, 0.0], [1.0, 1.0, float(world)])
    con = g.AffineConstraints()
    sys_.make_zero_boundary_constraints(con)
    con.close()
    A = g.SparseMatrix()
    g.MatrixCreator.create_laplace_matrix(g.MappingQ1(), sys_, g.QGauss(p + 1), A, con)
    n_owned = sys_.n_locally_owned_dofs()
    rng = np.random.default_rng(rank)
    xh = rng.uniform(-1, 1, n_owned)
    # two (input, output) pairs used alternately: 4 x 136 MB per GPU cycle through HBM, the inputs
    # are never L2 resident and the values stay bounded (no repeated application to the same data)
    x, y = g.Vector(sys_, xh), g.Vector(sys_)
    x2, y2 = g.Vector(sys_, xh[::-1].copy()), g.Vector(sys_)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    pairs = [(x, y), (x2, y2)]

    def step(i):
        a, b = pairs[i & 1]
        A.vmult(b, a)

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(n_owned)], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(tot, op=torch.distributed.ReduceOp.SUM)
    ms, total_dofs = float(t.item()), float(tot.item())
    value = total_dofs * args.steps / (ms * 1e-3) / 1e9

    if args.quick:
        if rank == 0:
            emit({"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": ms / args.steps, "quick": True,
                  "frac_of_hbm_roofline": ALG_BYTES_PER_DOF * value / world / peaks()[0], "n_gpus": world})
        return
    # ---- end to end through the host-buffer entry point (pinned host memory)
    e2e_steps = max(3, min(args.steps, 10))
    hx = torch.empty(n_owned, dtype=torch.float64).pin_memory()
    hy = torch.empty(n_owned, dtype=torch.float64).pin_memory()
    hx.copy_(torch.from_numpy(xh))
    hxn, hyn = hx.numpy(), hy.numpy()
    A.vmult_host(hyn, hxn)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(e2e_steps):
        A.vmult_host(hyn, hxn)
        hxn, hyn = hyn, hxn
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    t = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_e2e = float(t.item())
    e2e_value = total_dofs * e2e_steps / (ms_e2e * 1e-3) / 1e9

    # ---- CG solve time on the same grid (the second half of BASELINE.json's metric)
    cg = None
    if args.cg_steps > 0:
        b = g.Vector(sys_)
        b.set(1.0)
        con.set_zero(b)
        u = g.Vector(sys_)
        # untimed warm-up solve: the first ncclAllReduce sets up its channels lazily (milliseconds)
        wctl = g.ReductionControl(3, 1e-30, 1e-30)
        try:
            g.SolverCG(wctl).solve(A, u, b, g.PreconditionIdentity())
        except g.NoConvergence:
            pass
        u.set(0.0)
        ctl = g.ReductionControl(args.cg_steps, 1e-30, 1e-30)
        barrier()
        e0.record()
        try:
            g.SolverCG(ctl).solve(A, u, b, g.PreconditionIdentity())
        except g.NoConvergence:
            pass
        e1.record()
        barrier()
        ms_cg = e0.elapsed_time(e1)
        cg = {"iterations": ctl.last_step(), "ms_per_iteration": ms_cg / max(ctl.last_step(), 1),
              "gdof_iterations_per_s": total_dofs * ctl.last_step() / (ms_cg * 1e-3) / 1e9}
        # ---- the converged solve of SURVEY 8d: -Laplace u = 1, u = 0 on the boundary, x0 = 0,
        # ReductionControl(10000, 1e-10, 1e-8), identity and Jacobi (reference: tests/poisson_02_gdm.cc:213-215).
        # b_i = int phi_i = (M 1)_i (partition of unity), zero on constrained rows.  On the BASELINE grid (1 GPU) the
        # iteration counts are checked against the oracle's converged solves (tests/golden/cg_poisson3d.json).
        free = g.AffineConstraints()
        free.close()
        M = g.SparseMatrix()
        g.MatrixCreator.create_mass_matrix(g.MappingQ1(), sys_, g.QGauss(p + 1), M, free)
        ones = g.Vector(sys_)
        ones.set(1.0)
        M.vmult(b, ones)
        con.set_zero(b)
        gold = {}
        gpath = os.path.join(ROOT, "tests", "golden", "cg_poisson3d.json")
        if world == 1 and os.path.exists(gpath):
            gold = {r["precondition"]: r for r in json.load(open(gpath)) if r["N"] == n and r["p"] == p}
        cg["solve"] = {}
        for pre in ("identity", "jacobi"):
            if pre == "identity":
                P = g.PreconditionIdentity()
            else:
                P = g.PreconditionJacobi()
                P.initialize(A)
            sctl = g.ReductionControl(10000, 1e-10, 1e-8)
            u.set(0.0)
            barrier()
            e0.record()
            g.SolverCG(sctl).solve(A, u, b, P)
            e1.record()
            barrier()
            sec = e0.elapsed_time(e1) * 1e-3
            t = torch.tensor([sec], dtype=torch.float64, device="cuda")
            if world > 1:
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            rec = {"iterations": sctl.last_step(), "seconds": float(t.item()), "initial_residual": sctl.initial_value(),
                   "final_residual": sctl.last_value(), "control": "ReductionControl(10000, 1e-10, 1e-8)"}
            if pre in gold:
                rec["oracle_iterations"] = gold[pre]["iterations"]
                rec["oracle_seconds_cpu"] = gold[pre]["oracle_seconds"]
                assert abs(rec["iterations"] - gold[pre]["iterations"]) <= 1, (pre, rec["iterations"], gold[pre]["iterations"])
            cg["solve"][pre] = rec

    if rank != 0:
        return
    peak, which = peaks()
    achieved = ALG_BYTES_PER_DOF * (total_dofs / world) / (ms / args.steps * 1e-3) / 1e9
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this workload and this
    # kernel configuration (profiles/r2/traffic.json; only valid for the BASELINE grid on one GPU)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r2", "traffic.json")
    if os.path.exists(tpath) and world == 1 and n == 256 and p == 3:
        tj = json.load(open(tpath))
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    # second roof: FP64 pipe.  43 FP64 instructions per DoF in sum-factorised form (x 11, y 18, z 14);
    # peak = measured DFMA issue rate (profiles/microbench/microbench2.json: 33.8 TFLOP/s = 16.9 T instr-lanes/s)
    fp64_ops_per_dof, fp64_peak = 43.0, 16.9e12
    fp64 = {"bound": "fp64", "achieved": fp64_ops_per_dof * (total_dofs / world) / (ms / args.steps * 1e-3) / 1e12,
            "peak": fp64_peak / 1e12, "unit": "T FP64 instr-lanes/s", "ops_per_dof": fp64_ops_per_dof}
    fp64["frac"] = fp64["achieved"] / fp64["peak"]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": f"{which} (MEASURED_PEAKS.json hbm_gbs)",
                "kernel": "kron3d_pers_kernel (persistent fused TMA-staged tensor-product apply)" if A.kernel_used() == 2 else "generic band passes",
                "algorithmic_bytes_per_launch": ALG_BYTES_PER_DOF * total_dofs / world,
                "second_roof": fp64,
                "note": "duration = CUDA-event time of the timed region / steps (includes the constrained-face kernels)"}
    base = cpu_csr_sample(args.cpu_cells, p, 10, 2) if world == 1 or True else None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"3D Poisson GDM p={p}, {n}x{n}x{n * world} cells ({int(total_dofs)} DoFs), zero Dirichlet, "
                                   f"matrix-free stiffness apply y=Kx; x~U(-1,1) seed rank; ping-pong two alternating (x,y) pairs "
                                   f"(4 x {n_owned * 8 / 1e6:.0f} MB per GPU > L2, no L2 flush needed)",
                       "parallelism": f"slab{world}", "kernel": roofline["kernel"]},
            "roofline": roofline, "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_owned * 8, "d2h_bytes_per_step": n_owned * 8,
                    "steps": e2e_steps, "api": "gdm_operator_vmult_host (pinned host buffers)"},
            "gpu_launches": int(launches), "clocks": clocks, "cg": cg}
    emit(line)


def run_app(args):
    """BASELINE configs 3 and 4: explicit RK4 time stepping with a Jacobi-CG mass solve in every stage.

    advection_rk4 (config 3, strong scaling): periodic [0,1]^3, `--cells` (512) cells per direction in total, p = `--p`
    (5), b = (1, 0.15, -0.05), u0 = sin(2 pi x) cos(2 pi y) cos(2 pi z), dt = 0.5 h
    (prototypes/advection_01_gdm.cc:33-52,79,144-224,268-281), slabs over the ranks.
    wave_rk4 (config 4, weak scaling): u_tt = Laplace u on `--cells`^3 cells per GPU, zero Dirichlet, first-order system
    [u; v]' = [v; M^-1(-K u)] (applications/wave/include/gdm/wave/problem.h:294-345), dt = 0.3 h.
    Mass solves: ReductionControl(1000, 1e-20, 1e-14) (applications/wave/include/gdm/wave/parameters.h:41-44).
    One "step" = one RK4 step (4 operator applies + 4 CG solves + stage updates).
    """
    import numpy as np
    import torch
    import gdm_b200 as g

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    ctx = g.init_distributed(local) if world > 1 else g.default_context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    adv = args.workload == "advection_rk4"
    n, p = args.cells, args.p
    reps = [n, n, n] if adv else [n, n, n * world]
    hi = [1.0, 1.0, 1.0] if adv else [1.0, 1.0, float(world)]
    sys_ = g.System(3, p, 1, comm="world" if world > 1 else None, context=ctx)
    sys_.subdivided_hyper_rectangle(reps, [0.0, 0.0, 0.0], hi)
    con = g.AffineConstraints()
    if adv:
        for d in range(3):
            sys_.make_periodicity_constraints(d, con)
    else:
        sys_.make_zero_boundary_constraints(con)
    con.close()
    mp, q = g.MappingQ1(), g.QGauss(p + 1)
    M = g.SparseMatrix()
    g.MatrixCreator.create_mass_matrix(mp, sys_, q, M, con)
    R = g.SparseMatrix()
    bvec = [1.0, 0.15, -0.05]
    if adv:
        g.MatrixCreator.create_advection_matrix(mp, sys_, q, R, con, bvec, -1.0)
    else:
        g.MatrixCreator.create_laplace_matrix(mp, sys_, q, R, con)
    own = sys_.locally_owned_dofs()
    nx, ny = reps[0] + 1, reps[1] + 1
    z0, z1 = own.start // (nx * ny), own.stop // (nx * ny)
    h = 1.0 / n
    X = np.arange(nx) * h
    Y = np.arange(ny) * h
    Z = np.arange(z0, z1) * h

    def exact(t):
        if adv:
            return (np.cos(2 * np.pi * (Z - t * bvec[2]))[:, None, None] * np.cos(2 * np.pi * (Y - t * bvec[1]))[None, :, None] *
                    np.sin(2 * np.pi * (X - t * bvec[0]))[None, None, :]).reshape(-1)
        w = np.pi * np.sqrt(2.0 + 1.0 / world ** 2)  # standing wave of the box [0,1]^2 x [0,world]
        return (np.sin(np.pi * Z / world)[:, None, None] * np.sin(np.pi * Y)[None, :, None] * np.sin(np.pi * X)[None, None, :]).reshape(-1) * np.cos(w * t)

    u = g.Vector(sys_, exact(0.0))
    v = g.Vector(sys_)
    pre = g.PreconditionJacobi()
    pre.initialize(M)
    tmp0, tmp1 = g.Vector(sys_), g.Vector(sys_)
    iters = []
    max_it, tol, red = 1000, 1e-20, args.cg_reduce

    direct = args.mass_solver == "direct"  # Kronecker-direct mass inverse instead of the reference's CG

    def mass_solve(out, rhs):
        if direct:
            M.mass_inverse(out, rhs)
            return
        out.set(0.0)
        ctl = g.ReductionControl(max_it, tol, red)
        g.SolverCG(ctl).solve(M, out, rhs, pre)
        iters.append(ctl.last_step())

    def f_adv(t, y, out):
        tmp0.equ(y)
        con.distribute(tmp0)
        R.vmult(tmp1, tmp0)
        mass_solve(out, tmp1)

    def f_wave(t, y, out):
        out[0].equ(y[1])
        R.vmult(tmp1, y[0])
        tmp1.scale(-1.0)
        con.set_zero(tmp1)
        mass_solve(out[1], tmp1)

    rk = g.TimeStepping.ExplicitRungeKutta(g.TimeStepping.RK_CLASSIC_FOURTH_ORDER)
    dt = (0.5 if adv else 0.3) * h
    state = u if adv else [u, v]
    f = f_adv if adv else f_wave

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    t = 0.0
    for _ in range(max(args.warmup, 1) if args.warmup < 3 else 3):
        t = rk.evolve_one_time_step(f, t, dt, state)
        if adv:
            con.distribute(u)
    iters.clear()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        t = rk.evolve_one_time_step(f, t, dt, state)
        if adv:
            con.distribute(u)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    n_owned = sys_.n_locally_owned_dofs()
    err = float(np.abs(u.numpy() - exact(t)).max())
    tt = torch.tensor([ms, err], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(n_owned)], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(tot, op=torch.distributed.ReduceOp.SUM)
    ms, err, total_dofs = float(tt[0].item()), float(tt[1].item()), float(tot.item())
    if rank != 0:
        return
    peak, which = peaks()
    n_it = float(np.mean(iters)) if iters else 0.0
    # algorithmic bytes per RK step and DoF: 4 stages x (apply 16 + Jacobi-CG iterations x 96) + stage combinations
    # (lincomb: read y and up to 4 k, write 1) ~ 4 x 24 + 48; the wave system adds the copy u' = v and the scale/zero passes
    # direct mass inverse: one prepare pass (24) + forward and backward sweep per direction (32, +16 for the border
    # correction of a periodic direction)
    mass_bytes = (24 + 3 * (32 + (16 if adv else 0))) if direct else n_it * 96
    bytes_per_dof = 4 * (16 + mass_bytes) + 4 * 24 + 48 + (0 if adv else 4 * (16 + 16))
    achieved = bytes_per_dof * (total_dofs / world) / (ms / args.steps * 1e-3) / 1e9
    value = total_dofs * args.steps / (ms * 1e-3) / 1e9
    solver_txt = ("Kronecker-direct mass inverse (banded line solves, gdm_operator_mass_inverse) per stage" if direct else
                  f"Jacobi-CG mass solve ReductionControl({max_it},{tol:g},{red:g}) per stage")
    line = {"metric": f"gdm_{args.workload}_3d_p{p}_fp64", "value": value, "unit": "GDoF-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": 3, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if adv else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": (f"advection u_t + b.grad u = 0, periodic [0,1]^3, {n}^3 cells p={p} ({int(total_dofs)} DoFs), RK4, dt=0.5h, "
                                    f"{solver_txt}" if adv else
                                    f"wave u_tt = Laplace u as [u;v] block system, {n}x{n}x{n * world} cells p={p} ({int(total_dofs)} DoFs per block), "
                                    f"zero Dirichlet, RK4, dt=0.3h, {solver_txt}"),
                       "parallelism": f"slab{world}", "kernel": "fused" if (M.kernel_used() == 2 and R.kernel_used() == 2) else "generic"},
            "cg_iterations_per_stage": n_it, "cg_iterations_min_max": [int(min(iters)), int(max(iters))] if iters else None,
            "nodal_linf_error_vs_exact": err, "t_end": t,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "algorithmic_bytes_per_dof_per_step": bytes_per_dof, "peak_source": f"{which} (MEASURED_PEAKS.json hbm_gbs)"},
            "gpu_launches": int(launches), "clocks": clocks}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """Print the one JSON line on the process's original stdout."""
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # native libraries (NCCL's version banner) write to fd 1: keep fd 1 for the JSON line only
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=256, help="cells per direction per GPU")
    ap.add_argument("--p", type=int, default=3)
    ap.add_argument("--cpu-cells", type=int, default=96, help="grid of the bounded CPU sample")
    ap.add_argument("--cg-steps", type=int, default=50)
    ap.add_argument("--quick", action="store_true", help="timed applies only (for ncu captures)")
    ap.add_argument("--workload", default="apply", choices=["apply", "advection_rk4", "wave_rk4"],
                    help="apply (default, BASELINE config 2), advection_rk4 (config 3), wave_rk4 (config 4)")
    ap.add_argument("--cg-reduce", type=float, default=1e-14, help="reduction of the mass solves of the RK workloads")
    ap.add_argument("--mass-solver", default="cg", choices=["cg", "direct"],
                    help="RK workloads: the reference's Jacobi-CG mass solve (default) or the Kronecker-direct inverse (1 GPU)")
    args = ap.parse_args()
    if args.workload != "apply":
        if "--cells" not in sys.argv:
            args.cells = 512 if args.workload == "advection_rk4" else 256
        if "--p" not in sys.argv:
            args.p = 5 if args.workload == "advection_rk4" else 3
        if "--steps" not in sys.argv:
            args.steps = 10
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "apply":
        run_app(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
