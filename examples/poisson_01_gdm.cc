// Poisson problem: serial, 1D -- the driver of the reference's tests/poisson_01_gdm.cc written
// against include/gdm of this repository (GPU path).  stdout is diffed against the reference's
// committed golden tests/golden/poisson_01_gdm.output by tests/test_gpu_examples.py.
#include <gdm/system.h>
#include <gdm/matrix_creator.h>
#include <gdm/vector_tools.h>

#include <cstdio>
#include <iostream>

using namespace dealii;

template <int dim, typename Number = double>
class RightHandSideFunction : public dealii::Function<dim, Number>
{
public:
  virtual double value(const dealii::Point<dim> &, const unsigned int = 1) const override { return 1.0; }
};

template <int dim, typename Number = double>
class ExactSolution : public dealii::Function<dim, Number>
{
public:
  virtual double value(const dealii::Point<dim> &p, const unsigned int = 1) const override
  {
    return 0.125 - 0.5 * (p[0] - 0.5) * (p[0] - 0.5);
  }
};

template <int dim>
void test(const unsigned int fe_degree)
{
  const unsigned int n_subdivisions = 10;
  const unsigned int n_components   = 1;
  using Number     = double;
  using VectorType = Vector<Number>;

  ExactSolution<dim> exact_solution;

  GDM::System<dim> system(fe_degree, n_components);
  system.subdivided_hyper_cube(n_subdivisions);

  hp::MappingCollection<dim> mapping;
  mapping.push_back(MappingQ1<dim>());
  hp::QCollection<dim> quadrature;
  quadrature.push_back(QGauss<dim>(fe_degree + 1));

  AffineConstraints<Number> constraints;
  system.make_zero_boundary_constraints(constraints);
  constraints.close();
  system.categorize();

  DynamicSparsityPattern dsp(system.n_dofs());
  system.create_sparsity_pattern(constraints, dsp);
  SparsityPattern sparsity_pattern;
  sparsity_pattern.copy_from(dsp);
  SparseMatrix<Number> sparse_matrix;
  sparse_matrix.reinit(sparsity_pattern);

  VectorType rhs(system), solution(system);

  // matrix: (grad phi_i, grad phi_j); right-hand side (1, phi_i) = M 1 with the constraints applied
  GDM::MatrixCreator::create_laplace_matrix(mapping, system, quadrature, sparse_matrix, constraints);
  {
    SparseMatrix<Number>      mass;
    AffineConstraints<Number> none;
    none.close();
    GDM::MatrixCreator::create_mass_matrix(mapping, system, quadrature, mass, none);
    VectorType ones(system);
    ones = 1.0;
    mass.vmult(rhs, ones);
    constraints.set_zero(rhs);
  }

  PreconditionIdentity preconditioner;
  ReductionControl     solver_control(100, 1.e-10, 1.e-4);
  SolverCG<VectorType> solver(solver_control);
  solver.solve(sparse_matrix, solution, rhs, preconditioner);
  std::cout << solver_control.last_step() << std::endl << std::endl;

  for (const auto &value : solution.to_host())
    std::cout << (std::abs(value) < 1e-300 ? 0.0 : value) << std::endl;

  std::vector<double> cell_wise_error;
  GDM::VectorTools::integrate_difference(mapping, system, solution, exact_solution, cell_wise_error, quadrature,
                                         GDM::VectorTools::L2_norm);
  const auto error = GDM::VectorTools::compute_global_error(cell_wise_error, GDM::VectorTools::L2_norm);
  printf("%8.5f %14.8f\n\n", 0.0, error);
}

int main()
{
  for (const unsigned int fe_degree : {1, 3, 5, 7, 9})
    test<1>(fe_degree);
}
