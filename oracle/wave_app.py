"""The wave application's explicit run restated (oracle, test-only): `applications/wave`, simulation "wave".

Follows `applications/wave/include/gdm/wave/problem.h:280-345` (`wave-rk`, not composite):

    M = cut mass matrix                      `wave/mass.h:47-249`   (inside mass + gamma_M h^3 ghost penalty, zero diagonal -> 1)
    f(t, [u; v]) = [v; M^-1 rhs(u, t)]       `problem.h:302-320`
    rhs(u, t) = -(grad v, grad u) - Nitsche(u) - gamma_A h ghost penalty(u) + <gamma_D/h v - dv/dn, g(t)>
                                             `wave/stiffness.h:42-407` (matrix-free residual, linear in u)
    RK4 with `DiscreteTime(start, end, cfl * dx^cfl_pow)`, `postprocess` after every step (`problem.h:322-345`)
    postprocess: L2, L1, Linf of u_h - u over the inside part (`problem.h:531-615`)

with the preset of `applications/wave/wave-app.cc:222-284` (p = 3, 40 cells on [-1.21, 1.21], gamma_M = sqrt(3)/4,
gamma_A = sqrt(3)/2, gamma_D = 5 p, cfl 0.3, end time 2, unit-sphere level set, exact solution cos(k r) cos(k t) with
k = 3 pi / 2 in 1D).  The preset interpolates the level set with FE_Q(p); in 1D |x| - 1 is linear on every cut cell, so
the Q1 level set of `oracle.cut` is the same function there.  The mass solve is exact here (the reference's AMG-CG to
1e-14 is one, its `[L] solved in k` lines are not reproducible); the printed error columns are.
"""
import numpy as np
import scipy.sparse.linalg as sla

from . import cut
from .solvers import DiscreteTime, ExplicitRungeKutta4
from .system import System
from .vector_tools import interpolate


def wave_preset(dim=1):
    p = 3
    k = {1: 1.5 * np.pi}[dim]  # 2D needs J0 (boost::math::cyl_bessel_j) and a Q3 level set: not restated
    return dict(dim=dim, fe_degree=p, n_subdivisions=40, left=-1.21, right=1.21,
                ghost_parameter_M=0.25 * np.sqrt(3.0), ghost_parameter_A=0.5 * np.sqrt(3.0), nitsche_parameter=5.0 * p,
                start_t=0.0, end_t=2.0, cfl=0.3, cfl_pow=1.0,
                exact=lambda pts, t: np.cos(k * np.linalg.norm(pts, axis=1)) * np.cos(k * t))


def heat_preset(dim=1):
    """`applications/wave/wave-app.cc:62-150`, "heat-rk": u_t = u_xx + f, exact solution x^9 exp(-t)."""
    p = 3
    return dict(dim=dim, fe_degree=p, n_subdivisions=40, left=-1.21, right=1.21,
                ghost_parameter_M=0.75, ghost_parameter_A=1.5, nitsche_parameter=5.0 * p,
                start_t=0.0, end_t=0.1, cfl=0.3 / p / p, cfl_pow=2.0,
                exact=lambda pts, t: pts[:, 0] ** 9 * np.exp(-t),
                rhs=lambda pts, t: -pts[:, 0] ** 7 * np.exp(-t) * (pts[:, 0] ** 2 + 72))


def wave_operators(params):
    """System, level set, cell locations, the mass matrix M, the matrix A of the residual's linear part (rows of
    untouched DoFs empty) and the volume / surface load functionals."""
    dim, p = params["dim"], params["fe_degree"]
    s = System(dim, p, 1, add_ghost_layer=True)
    s.subdivided_hyper_cube(params["n_subdivisions"], params["left"], params["right"])
    ls = cut.interpolate_level_set(s, cut.sphere_level_set([0.0] * dim, 1.0))
    if params.get("level_set_degree", 1) > 1:  # the 2D presets: FE_Q(fe_degree) level set (`wave-app.cc:277`)
        from .cut_q import LevelSetQ
        ls = LevelSetQ(s, cut.sphere_level_set([0.0] * dim, 1.0), params["level_set_degree"])
    M, _, loc = cut.assemble_cut_poisson(s, ls, True, params["ghost_parameter_M"], rhs_value=0.0, gp_h_power=3, kind="mass")
    A, _, _ = cut.assemble_cut_poisson(s, ls, True, params["ghost_parameter_A"], params["nitsche_parameter"],
                                       rhs_value=0.0, boundary_value=0.0, gp_h_power=1, outside_diagonal=0.0)
    volume, surface = cut.load_functionals(s, ls, params["nitsche_parameter"], loc)
    return s, ls, loc, M, A, volume, surface


def explicit_run(params, second_order, max_steps=None):
    """Rows (step, time, L2, L1, Linf) as `wave-app` prints them; `second_order`: the wave system [u; v]
    (`problem.h:280-345`), else the heat equation u' = M^-1 rhs(u, t) (`problem.h:72-127`)."""
    s, ls, loc, M, A, volume, surface = wave_operators(params)
    exact, source = params["exact"], params.get("rhs")
    dx = (params["right"] - params["left"]) / params["n_subdivisions"]
    dt = params["cfl"] * dx ** params["cfl_pow"]
    solve = sla.factorized(M.tocsc())
    n = s.n_dofs()
    u0 = interpolate(s, lambda pts, c: exact(pts, params["start_t"]))  # `problem.h:295-297`, `:86-88`
    y = np.concatenate([u0, np.zeros(n)]) if second_order else u0

    def residual(t, u):  # `wave/stiffness.h:42-407`
        r = -(A @ u) + cut.apply_load(n, surface, lambda pts: exact(pts, t))
        if source is not None:
            r += cut.apply_load(n, volume, lambda pts: source(pts, t))
        return solve(r)

    def f(t, yy):
        if second_order:
            return np.concatenate([yy[n:], residual(t, yy[:n])])
        return residual(t, yy)

    rows = []

    def postprocess(t):
        rows.append((len(rows), t) + cut.error_norms_inside(s, ls, y[:n], lambda pts: exact(pts, t), loc))

    time = DiscreteTime(params["start_t"], params["end_t"], dt)
    rk = ExplicitRungeKutta4()
    postprocess(0.0)
    while not time.is_at_end() and (max_steps is None or len(rows) <= max_steps):
        _, y = rk.evolve_one_time_step(f, time.get_current_time(), time.get_next_step_size(), y)
        postprocess(time.get_current_time() + time.get_next_step_size())
        time.advance_time()
    return rows


def wave_rk_run(params=None, max_steps=None):
    return explicit_run(params or wave_preset(1), True, max_steps)


def heat_rk_run(params=None, max_steps=None):
    return explicit_run(params or heat_preset(1), False, max_steps)


def heat_impl_run(params=None):
    """`problem.h:210-279` ("heat-impl"): implicit Euler  (M + dt S) u+ = M u + dt data(t + dt)  with the ASSEMBLED stiffness
    matrix S of `wave/stiffness.h:589-799`, whose ghost penalty carries h^3 (`:760-765`) where the matrix-free residual
    has h (`:386-392`), zero diagonal -> 1 (`:796-798`); preset cfl 0.3, cfl_pow 1 (`wave-app.cc:137-141`).  Outside the
    explicit hot path; restated because the golden `heat_0.output` pins the assembled-matrix form of the cut rows."""
    params = dict(params or heat_preset(1), cfl=0.3, cfl_pow=1.0)
    s, ls, loc, M, _, volume, surface = wave_operators(params)
    S, _, _ = cut.assemble_cut_poisson(s, ls, True, params["ghost_parameter_A"], params["nitsche_parameter"],
                                       rhs_value=0.0, boundary_value=0.0, gp_h_power=3)
    exact, source = params["exact"], params["rhs"]
    dx = (params["right"] - params["left"]) / params["n_subdivisions"]
    n = s.n_dofs()
    u = interpolate(s, lambda pts, c: exact(pts, params["start_t"]))
    rows = []

    def postprocess(t):
        rows.append((len(rows), t) + cut.error_norms_inside(s, ls, u, lambda pts: exact(pts, t), loc))

    time = DiscreteTime(params["start_t"], params["end_t"], params["cfl"] * dx ** params["cfl_pow"])
    postprocess(0.0)
    while not time.is_at_end():
        dt, t1 = time.get_next_step_size(), time.get_current_time() + time.get_next_step_size()
        data = cut.apply_load(n, surface, lambda pts: exact(pts, t1)) + cut.apply_load(n, volume, lambda pts: source(pts, t1))
        u = sla.spsolve((M + dt * S).tocsc(), M @ u + dt * data)
        postprocess(t1)
        time.advance_time()
    return rows


def composite_run(params, second_order):
    """The two-domain runs (`problem.h:129-209` heat, `:347-437` wave; presets `wave-app.cc:152-220,286-336`): one field
    on {level set < 0}, one on {level set > 0}, each with its own cut mass matrix and ghost penalty, Nitsche on the box
    boundary for the outer field (`wave/stiffness.h:262-340`), coupled on the surface (`:441-574`).  The outer domain's
    operators are the inner domain's with the level set negated.  Rows alternate inside / outside like the output."""
    dim, p = params["dim"], params["fe_degree"]
    s = System(dim, p, 1, add_ghost_layer=True)
    s.subdivided_hyper_cube(params["n_subdivisions"], params["left"], params["right"])
    ls = cut.interpolate_level_set(s, cut.sphere_level_set([0.0] * dim, 1.0))
    gd, hmin, n = params["nitsche_parameter"], min(s.h), s.n_dofs()
    exact, source = params["exact"], params.get("rhs")
    fields = []
    for sign in (1.0, -1.0):
        lsd = sign * ls
        M, _, loc = cut.assemble_cut_poisson(s, lsd, True, params["ghost_parameter_M"], rhs_value=0.0, gp_h_power=3, kind="mass")
        A, _, _ = cut.assemble_cut_poisson(s, lsd, True, params["ghost_parameter_A"], gd, rhs_value=0.0, boundary_value=0.0,
                                           gp_h_power=1, outside_diagonal=0.0, surface_terms=False)
        B, bload = cut.domain_boundary_terms(s, lsd, gd)
        volume, _ = cut.load_functionals(s, lsd, gd, loc)
        fields.append(dict(ls=lsd, loc=loc, solve=sla.factorized(M.tocsc()), A=(A + B).tocsr(), bload=bload, volume=volume))
    P, Q = cut.coupling_matrices(s, ls)
    tau = 0.5 * gd / hmin

    def residuals(t, u0, u1):
        jump, total = u0 - u1, u0 + u1
        c_sym = -0.5 * (P @ jump)
        c_avg = 0.5 * (P.T @ total)
        c_pen = tau * (Q @ jump)
        out = []
        for k, (f, u) in enumerate(zip(fields, (u0, u1))):
            r = -(f["A"] @ u) + cut.apply_load(n, f["bload"], lambda pts: exact(pts, t))
            if source is not None:
                r += cut.apply_load(n, f["volume"], lambda pts: source(pts, t))
            r -= (c_sym - c_avg + c_pen) if k == 0 else (c_sym + c_avg - c_pen)
            out.append(f["solve"](r))
        return out

    u_init = interpolate(s, lambda pts, c: exact(pts, params["start_t"]))
    y = np.concatenate([u_init, u_init] + ([np.zeros(n), np.zeros(n)] if second_order else []))

    def f(t, yy):
        r0, r1 = residuals(t, yy[:n], yy[n:2 * n])
        return np.concatenate([yy[2 * n:], r0, r1]) if second_order else np.concatenate([r0, r1])

    rows, counter = [], [0, 0]

    def postprocess(t):
        for k, fl in enumerate(fields):
            rows.append((counter[k], t) + cut.error_norms_inside(s, fl["ls"], y[k * n:(k + 1) * n], lambda pts: exact(pts, t), fl["loc"]))
            counter[k] += 1

    dx = (params["right"] - params["left"]) / params["n_subdivisions"]
    time = DiscreteTime(params["start_t"], params["end_t"], params["cfl"] * dx ** params["cfl_pow"])
    rk = ExplicitRungeKutta4()
    postprocess(0.0)
    while not time.is_at_end():
        _, y = rk.evolve_one_time_step(f, time.get_current_time(), time.get_next_step_size(), y)
        postprocess(time.get_current_time() + time.get_next_step_size())
        time.advance_time()
    return rows


def step85_preset(dim=2):
    """`applications/wave/wave-app.cc:13-61` ("step85"): Poisson, f = 4, g = 1, exact 1 - 2/dim (|x|^2 - 1)."""
    p = 3
    return dict(dim=dim, fe_degree=p, n_subdivisions=40, left=-1.21, right=1.21, ghost_parameter_A=0.5,
                nitsche_parameter=5.0 * p, level_set_degree=p,
                exact=lambda pts, t: 1.0 - 2.0 / dim * ((pts ** 2).sum(axis=1) - 1.0))


def poisson_run(params=None):
    """`problem.h:46-70` ("poisson"): the assembled stiffness matrix of `wave/stiffness.h:589-799` (ghost penalty with
    h^3, zero diagonal -> 1), right-hand side `compute_rhs(., 0, false, 0)` = (v, f) + <gamma_D / h v - d_n v, g>, one
    solve, one printed line."""
    params = params or step85_preset(2)
    dim, p = params["dim"], params["fe_degree"]
    s = System(dim, p, 1, add_ghost_layer=True)
    s.subdivided_hyper_cube(params["n_subdivisions"], params["left"], params["right"])
    ls = cut.interpolate_level_set(s, cut.sphere_level_set([0.0] * dim, 1.0))
    if params.get("level_set_degree", 1) > 1:
        from .cut_q import LevelSetQ
        ls = LevelSetQ(s, cut.sphere_level_set([0.0] * dim, 1.0), params["level_set_degree"])
    S, rhs, loc = cut.assemble_cut_poisson(s, ls, True, params["ghost_parameter_A"], params["nitsche_parameter"],
                                           rhs_value=4.0, boundary_value=1.0, gp_h_power=3)
    u = sla.spsolve(S.tocsc(), rhs)
    return [(0, 0.0) + cut.error_norms_inside(s, ls, u, lambda pts: params["exact"](pts, 0.0), loc)]
