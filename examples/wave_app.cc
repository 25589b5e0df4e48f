// The explicit runs of the reference's applications/wave/wave-app.cc written against include/gdm of this repository:
//   ./wave_app 1 wave      (wave-app.cc:222-284, wave/problem.h:280-345: RK4 on [u; v], cut mass solves)
//   ./wave_app 1 heat-rk   (wave-app.cc:62-150,  wave/problem.h:72-127)
//   ./wave_app 2 wave      (the 2D preset: Bessel solution, level set interpolated with FE_Q(3))
//   ./wave_app 2 step85    (wave-app.cc:13-61, wave/problem.h:46-70: Poisson with the assembled cut matrix)
// Cut-cell set-up on the host (GDM::CutCellSetup), mass / stiffness operators with the cut rows attached, Jacobi-CG
// mass solves and the Runge-Kutta stages on the GPU.  stdout has the reference's format (wave/problem.h:609-615 and
// the " [L] solved in k" lines of :498); tests/test_gpu_zz_cut.py diffs the error columns against
// applications/wave/tests/{wave_0,heat_1,wave_1}.output.
#include <gdm/system.h>
#include <gdm/matrix_creator.h>
#include <gdm/vector_tools.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>

using namespace dealii;

template <int dim>
struct TimeFunction : public Function<dim>
{
  std::function<double(double, const Point<dim> &)> fn;
  double                                            time = 0.0;
  double value(const Point<dim> &p, const unsigned int = 0) const override { return fn(time, p); }
};

template <int dim>
struct SignedDistanceSphere : public Function<dim>
{
  double value(const Point<dim> &p, const unsigned int = 0) const override
  {
    double r2 = 0;
    for (int d = 0; d < dim; ++d)
      r2 += p[d] * p[d];
    return std::sqrt(r2) - 1.0;
  }
};

template <int dim>
struct Parameters // applications/wave/include/gdm/wave/parameters.h
{
  bool         poisson = false;      // "poisson": one solve with the assembled matrix (wave/problem.h:46-70)
  bool         second_order = false; // "wave-rk" ([u; v]) or "heat-rk"
  unsigned int fe_degree = 3, n_subdivisions_1D = 40;
  double       geometry_left = -1.21, geometry_right = 1.21;
  double       ghost_parameter_M = -1, ghost_parameter_A = -1, nitsche_parameter = -1;
  TimeFunction<dim> exact_solution, function_rhs;
  bool              has_rhs = false;
  int               level_set_degree = 1;
  double            start_t = 0, end_t = 0, cfl = 0, cfl_pow = 1;
};

template <int dim>
void fill_parameters(Parameters<dim> &params, const std::string &simulation_name)
{
  const double pi = 3.14159265358979323846;
  if (simulation_name == "wave")
    {
      params.second_order      = true;
      params.ghost_parameter_M = 0.25 * std::sqrt(3.0);
      params.ghost_parameter_A = 0.50 * std::sqrt(3.0);
      params.nitsche_parameter = 5.0 * params.fe_degree;
      params.exact_solution.fn = [pi](const double t, const Point<dim> &p) {
        double r2 = 0;
        for (int d = 0; d < dim; ++d)
          r2 += p[d] * p[d];
        const double r = std::sqrt(r2);
        if (dim == 1)
          return std::cos(1.5 * pi * r) * std::cos(1.5 * pi * t);
        return std::cyl_bessel_j(0.0, 3.0 * pi * r) * std::cos(3.0 * pi * t); // dim == 2 (wave-app.cc:256-261)
      };
      if (dim > 2)
        throw ExcNotImplemented("wave preset: dim 1 and 2");
      params.level_set_degree = dim == 1 ? 1 : params.fe_degree; // wave-app.cc:277 (in 1D |x| - 1 is linear on the cut cells)
      params.end_t = 2.0;
      params.cfl   = 0.3;
    }
  else if (simulation_name == "step85")
    {
      // wave-app.cc:13-61: Poisson, f = 4, g = 1 on the unit circle / sphere, exact solution 1 - 2/dim (|x|^2 - 1)
      params.poisson           = true;
      params.ghost_parameter_A = 0.5;
      params.nitsche_parameter = 5.0 * params.fe_degree;
      params.exact_solution.fn = [](const double, const Point<dim> &p) {
        double r2 = 0;
        for (int d = 0; d < dim; ++d)
          r2 += p[d] * p[d];
        return 1. - 2. / dim * (r2 - 1.);
      };
      if (dim != 2)
        throw ExcNotImplemented("step85 preset: dim 2");
      params.level_set_degree = params.fe_degree;
    }
  else if (simulation_name == "heat-rk")
    {
      params.ghost_parameter_M = 0.75;
      params.ghost_parameter_A = 1.5;
      params.nitsche_parameter = 5.0 * params.fe_degree;
      if (dim != 1)
        throw ExcNotImplemented("heat preset: only the 1D form is set up here");
      params.exact_solution.fn = [](const double t, const Point<dim> &p) { return std::pow(p[0], 9.0) * std::exp(-t); };
      params.function_rhs.fn   = [](const double t, const Point<dim> &p) {
        return -std::pow(p[0], 7.0) * std::exp(-t) * (std::pow(p[0], 2.0) + 72);
      };
      params.has_rhs = true;
      params.end_t   = 0.1;
      params.cfl     = 0.3 / params.fe_degree / params.fe_degree;
      params.cfl_pow = 2.0;
    }
  else
    throw ExcNotImplemented("simulation " + simulation_name);
}

template <int dim>
void run(Parameters<dim> &params)
{
  using VectorType = Vector<double>;
  const unsigned int p = params.fe_degree, n = params.n_subdivisions_1D;
  const double       dx = (params.geometry_right - params.geometry_left) / n;

  GDM::System<dim> system(p, 1, true);
  system.subdivided_hyper_cube(n, params.geometry_left, params.geometry_right);
  hp::MappingCollection<dim> mapping;
  mapping.push_back(MappingQ1<dim>());
  hp::QCollection<dim> quadrature;
  quadrature.push_back(QGauss<dim>(p + 1));
  AffineConstraints<double> constraints;
  constraints.close();
  system.categorize();

  if (params.poisson)
    {
      // assembled stiffness matrix (wave/stiffness.h:589-799: ghost penalty with h^3, zero diagonal -> 1) and the right-hand
      // side compute_rhs(., 0, false, 0) = (v, 4) + <gamma_D / h v - dv/dn, 1>
      typename GDM::CutCellSetup<dim>::Parameters ps;
      ps.ghost_parameter   = params.ghost_parameter_A;
      ps.gp_h_power        = 3;
      ps.nitsche_parameter = params.nitsche_parameter;
      ps.rhs_value         = 4.0;
      ps.boundary_value    = 1.0;
      ps.level_set_degree  = params.level_set_degree;
      GDM::CutCellSetup<dim> cut_s(p, n, params.geometry_left, params.geometry_right, SignedDistanceSphere<dim>(), ps);
      SparseMatrix<double> matrix;
      GDM::MatrixCreator::create_laplace_matrix(mapping, system, quadrature, matrix, constraints);
      cut_s.attach_to(matrix);
      VectorType solution(system), rhs(system);
      rhs.from_host(cut_s.rhs());
      solution = 0.0;
      PreconditionJacobi<SparseMatrix<double>> preconditioner;
      preconditioner.initialize(matrix);
      ReductionControl     solver_control(1000, 1.e-20, 1.e-14);
      SolverCG<VectorType> solver(solver_control);
      solver.solve(matrix, solution, rhs, preconditioner);
      printf(" [L] solved in %u\n", solver_control.last_step());
      const auto e = cut_s.error_norms_inside(solution.to_host(), params.exact_solution);
      printf("%5d %8.5f %14.8e %14.8e %14.8e\n", 0, 0.0, e[0], e[1], e[2]);
      return;
    }

  // cut mass matrix (wave/mass.h:47-249) and the linear part of the residual (wave/stiffness.h:42-407)
  typename GDM::CutCellSetup<dim>::Parameters pm, pa;
  pm.kind_mass = true;
  pm.ghost_parameter = params.ghost_parameter_M;
  pm.gp_h_power = 3;
  pm.rhs_value = 0.0;
  pa.ghost_parameter = params.ghost_parameter_A;
  pa.gp_h_power = 1;
  pa.nitsche_parameter = params.nitsche_parameter;
  pa.rhs_value = pa.boundary_value = 0.0;
  pa.outside_diagonal = 0.0;
  pm.level_set_degree = pa.level_set_degree = params.level_set_degree;
  const SignedDistanceSphere<dim> level_set;
  GDM::CutCellSetup<dim> cut_m(p, n, params.geometry_left, params.geometry_right, level_set, pm);
  GDM::CutCellSetup<dim> cut_a(p, n, params.geometry_left, params.geometry_right, level_set, pa);

  SparseMatrix<double> mass_matrix, stiffness_matrix;
  GDM::MatrixCreator::create_mass_matrix(mapping, system, quadrature, mass_matrix, constraints);
  GDM::MatrixCreator::create_laplace_matrix(mapping, system, quadrature, stiffness_matrix, constraints);
  cut_m.attach_to(mass_matrix);
  cut_a.attach_to(stiffness_matrix);

  VectorType u(system), v(system), vec_rhs(system), load(system);
  params.exact_solution.time = params.start_t;
  GDM::VectorTools::interpolate(mapping, system, params.exact_solution, u);
  v = 0.0;

  PreconditionJacobi<SparseMatrix<double>> preconditioner;
  preconditioner.initialize(mass_matrix);

  // result = M^-1 rhs(u, t)
  const auto residual = [&](const double time, const VectorType &solution, VectorType &result) {
    stiffness_matrix.vmult(vec_rhs, solution);
    vec_rhs *= -1.0;
    params.exact_solution.time = time;
    params.function_rhs.time   = time;
    load.from_host(cut_a.load_vector(params.has_rhs ? &params.function_rhs : nullptr, &params.exact_solution));
    vec_rhs.add(1.0, load);
    result = 0.0;
    ReductionControl     solver_control(1000, 1.e-20, 1.e-14); // wave/parameters.h:41-44
    SolverCG<VectorType> solver(solver_control);
    solver.solve(mass_matrix, result, vec_rhs, preconditioner);
    printf(" [L] solved in %u\n", solver_control.last_step());
  };

  unsigned int counter = 0;
  const auto postprocess = [&](const double time, const VectorType &solution) {
    params.exact_solution.time = time;
    const auto e = cut_m.error_norms_inside(solution.to_host(), params.exact_solution);
    printf("%5d %8.5f %14.8e %14.8e %14.8e\n", counter++, time, e[0], e[1], e[2]);
  };

  const double delta_t = params.cfl * std::pow(dx, params.cfl_pow);
  DiscreteTime time(params.start_t, params.end_t, delta_t);
  postprocess(0.0, u);

  if (!params.second_order)
    {
      TimeStepping::ExplicitRungeKutta<VectorType> rk;
      rk.initialize(TimeStepping::RK_CLASSIC_FOURTH_ORDER);
      const std::function<VectorType(const double, const VectorType &)> fu_rhs = [&](const double t, const VectorType &y) {
        VectorType result;
        result.reinit(y);
        residual(t, y, result);
        return result;
      };
      while (!time.is_at_end())
        {
          rk.evolve_one_time_step(fu_rhs, time.get_current_time(), time.get_next_step_size(), u);
          postprocess(time.get_current_time() + time.get_next_step_size(), u);
          time.advance_time();
        }
      return;
    }

  // block system [u; v] (wave/problem.h:294-320) through the C ABI's block Runge-Kutta
  struct Ctx
  {
    const decltype(residual) *rhs_fn;
    gdm_system_t              sys;
    std::size_t               n;
    std::string               error;
  } ctx{&residual, u.system_handle(), u.size(), {}};
  gdm_rk_t rk = nullptr;
  dealii::internal::check(gdm_rk_create(ctx.sys, GDM_RK_CLASSIC_FOURTH_ORDER, 2, &rk));
  const auto tramp = [](double t, const gdm_vector_t *y, gdm_vector_t *out, void *user) -> int {
    Ctx *c = static_cast<Ctx *>(user);
    try
      {
        const VectorType yu = VectorType::view(y[0], c->sys, c->n), yv = VectorType::view(y[1], c->sys, c->n);
        VectorType       ou = VectorType::view(out[0], c->sys, c->n), ov = VectorType::view(out[1], c->sys, c->n);
        ou = yv;                      // du/dt = v
        (*c->rhs_fn)(t, yu, ov);      // dv/dt = M^-1 rhs(u, t)
        return GDM_OK;
      }
    catch (const std::exception &e)
      {
        c->error = e.what();
        return GDM_ERR_INTERNAL;
      }
  };
  while (!time.is_at_end())
    {
      gdm_vector_t blocks[2] = {u.handle(), v.handle()};
      double       t_new     = 0;
      const int    rc = gdm_rk_evolve_one_time_step(rk, tramp, &ctx, time.get_current_time(), time.get_next_step_size(),
                                                 blocks, &t_new);
      if (rc != GDM_OK && !ctx.error.empty())
        {
          gdm_rk_destroy(rk);
          throw ExcMessage(ctx.error);
        }
      dealii::internal::check(rc);
      postprocess(time.get_current_time() + time.get_next_step_size(), u);
      time.advance_time();
    }
  gdm_rk_destroy(rk);
}

int main(int argc, char **argv)
{
  if (argc != 3)
    {
      std::cout << "Usage: ./wave_app dim simulation" << std::endl << std::endl;
      std::cout << "dim         number of dimensions (1, 2)" << std::endl;
      std::cout << "simulation  name of simulation (wave, heat-rk, step85)" << std::endl;
      return 1;
    }
  try
    {
      const int dim = std::atoi(argv[1]);
      if (dim == 1)
        {
          Parameters<1> params;
          fill_parameters(params, argv[2]);
          run(params);
        }
      else if (dim == 2)
        {
          Parameters<2> params;
          fill_parameters(params, argv[2]);
          run(params);
        }
      else
        throw ExcNotImplemented("dim must be 1 or 2");
    }
  catch (const std::exception &e)
    {
      std::cerr << e.what() << std::endl;
      return 2;
    }
  return 0;
}
