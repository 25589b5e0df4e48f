// L2 projection with the consistent mass matrix, Jacobi-CG -- the driver of the reference's
// tests/mass_01_gdm.cc against include/gdm of this repository.  Golden: tests/golden/mass_01_gdm.output.
#include <gdm/system.h>
#include <gdm/matrix_creator.h>
#include <gdm/vector_tools.h>

#include <iostream>

using namespace dealii;

template <int dim>
class Linear : public Function<dim>
{
public:
  virtual double value(const Point<dim> &p, const unsigned int c = 0) const override { return p[0] + c; }
};

template <int dim>
void test()
{
  const unsigned int n_subdivisions = 40, fe_degree = 3, n_components = 1;
  using VectorType = Vector<double>;
  Linear<dim> function;

  GDM::System<dim> system(fe_degree, n_components);
  system.subdivided_hyper_cube(n_subdivisions);
  hp::MappingCollection<dim> mapping;
  mapping.push_back(MappingQ1<dim>());
  hp::QCollection<dim> quadrature;
  quadrature.push_back(QGauss<dim>(fe_degree + 1));
  AffineConstraints<double> constraints;
  constraints.close();
  system.categorize();

  SparseMatrix<double> sparse_matrix;
  GDM::MatrixCreator::create_mass_matrix(mapping, system, quadrature, sparse_matrix, constraints);

  // rhs_i = (f, phi_i): f = x is in the GDM space, so rhs = M * interpolate(f) exactly
  VectorType rhs(system), solution(system), nodal(system);
  GDM::VectorTools::interpolate(mapping, system, function, nodal);
  sparse_matrix.vmult(rhs, nodal);

  PreconditionJacobi<SparseMatrix<double>> preconditioner;
  preconditioner.initialize(sparse_matrix);
  ReductionControl     solver_control(100, 1.e-10, 1.e-8);
  SolverCG<VectorType> solver(solver_control);
  solver.solve(sparse_matrix, solution, rhs, preconditioner);

  std::vector<double> cell_wise_error;
  GDM::VectorTools::integrate_difference(mapping, system, solution, function, cell_wise_error, quadrature, GDM::VectorTools::L2_norm);
  std::cout << "error: " << GDM::VectorTools::compute_global_error(cell_wise_error) << std::endl;
}

int main()
{
  test<2>();
}
