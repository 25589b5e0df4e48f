// Periodic advection, RK4 with consistent-mass CG inversion per stage -- the time loop of the reference's
// prototypes/advection_01_gdm.cc:144-281 against include/gdm of this repository (its golden output is stale
// upstream, so the check is against the oracle: tests/test_gpu_examples.py).
#include <gdm/system.h>
#include <gdm/matrix_creator.h>
#include <gdm/vector_tools.h>

#include <cstdlib>
#include <iostream>

using namespace dealii;

template <int dim>
class ExactSolution : public Function<dim>
{
public:
  ExactSolution(const double time = 0.) : Function<dim>(1, time)
  {
    advection[0] = 1.0;
    if (dim > 1) advection[1] = 0.15;
    if (dim > 2) advection[2] = -0.05;
  }
  virtual double value(const Point<dim> &p, const unsigned int = 1) const override
  {
    const double t = this->get_time();
    double       r = std::sin(2. * (p[0] - t * advection[0]) * M_PI);
    for (unsigned int d = 1; d < dim; ++d)
      r *= std::cos(2. * (p[d] - t * advection[d]) * M_PI);
    return r;
  }
  std::array<double, dim> advection{};
};

template <int dim>
void test(const unsigned int fe_degree, const unsigned int n)
{
  using VectorType = Vector<double>;
  const double delta_t = 1.0 / n * 0.5, start_t = 0.0, end_t = 0.1;
  ExactSolution<dim> exact_solution;

  GDM::System<dim> system(fe_degree, 1);
  system.subdivided_hyper_cube(n);
  hp::MappingCollection<dim> mapping;
  mapping.push_back(MappingQ1<dim>());
  hp::QCollection<dim> quadrature;
  quadrature.push_back(QGauss<dim>(fe_degree + 1));
  AffineConstraints<double> constraints;
  for (unsigned int d = 0; d < dim; ++d)
    system.make_periodicity_constraints(d, constraints);
  constraints.close();
  system.categorize();

  SparseMatrix<double> mass, rhs_operator;
  GDM::MatrixCreator::create_mass_matrix(mapping, system, quadrature, mass, constraints);
  GDM::MatrixCreator::create_advection_matrix(mapping, system, quadrature, rhs_operator, constraints, exact_solution.advection, -1.0);

  VectorType solution(system);
  GDM::VectorTools::interpolate(mapping, system, exact_solution, solution);

  const auto fu_rhs = [&](const double, const VectorType &y) {
    VectorType vec_0(y), vec_1, vec_2;
    vec_1.reinit(y);
    vec_2.reinit(y);
    constraints.distribute(vec_0);
    rhs_operator.vmult(vec_1, vec_0);
    PreconditionJacobi<SparseMatrix<double>> preconditioner;
    preconditioner.initialize(mass);
    ReductionControl     solver_control(100, 1.e-10, 1.e-8);
    SolverCG<VectorType> solver(solver_control);
    solver.solve(mass, vec_2, vec_1, preconditioner);
    return vec_2;
  };
  const auto fu_postprocessing = [&](const double time) {
    exact_solution.set_time(time);
    std::vector<double> cell_wise_error;
    GDM::VectorTools::integrate_difference(mapping, system, solution, exact_solution, cell_wise_error, quadrature, GDM::VectorTools::L2_norm);
    std::cout << time << " " << GDM::VectorTools::compute_global_error(cell_wise_error) << std::endl;
  };

  DiscreteTime                                 time(start_t, end_t, delta_t);
  TimeStepping::ExplicitRungeKutta<VectorType> rk;
  rk.initialize(TimeStepping::RK_CLASSIC_FOURTH_ORDER);
  fu_postprocessing(0.0);
  while (time.is_at_end() == false)
    {
      rk.evolve_one_time_step(fu_rhs, time.get_current_time(), time.get_next_step_size(), solution);
      constraints.distribute(solution);
      fu_postprocessing(time.get_current_time() + time.get_next_step_size());
      time.advance_time();
    }
}

int main(int argc, char **argv)
{
  const unsigned int dim = argc > 1 ? std::atoi(argv[1]) : 2;
  const unsigned int p   = argc > 2 ? std::atoi(argv[2]) : 5;
  const unsigned int n   = argc > 3 ? std::atoi(argv[3]) : 40;
  if (dim == 2)
    test<2>(p, n);
  else
    test<3>(p, n);
}
