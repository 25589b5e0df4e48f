// Cut Poisson problem: serial, GDM -- the driver of the reference's prototypes/cut_poisson_01_gdm.cc written against
// include/gdm of this repository: cut-cell set-up on the host (GDM::CutCellSetup), tensor-product stiffness apply with
// the cut / ghost-penalty rows attached, CG on the GPU.  stdout has the reference's table format
// (prototypes/cut_poisson_01_gdm.output).
#include <gdm/system.h>
#include <gdm/matrix_creator.h>
#include <gdm/vector_tools.h>

#include <cmath>
#include <cstdio>
#include <iostream>

using namespace dealii;

template <int dim>
class AnalyticalSolution : public Function<dim>
{
public:
  double value(const Point<dim> &point, const unsigned int = 0) const override
  {
    double r2 = 0;
    for (int d = 0; d < dim; ++d)
      r2 += point[d] * point[d];
    return 1. - 2. / dim * (r2 - 1.);
  }
};

template <int dim>
class SignedDistanceSphere : public Function<dim> // Functions::SignedDistance::Sphere: unit sphere at the origin
{
public:
  double value(const Point<dim> &point, const unsigned int = 0) const override
  {
    double r2 = 0;
    for (int d = 0; d < dim; ++d)
      r2 += point[d] * point[d];
    return std::sqrt(r2) - 1.0;
  }
};

template <int dim>
void test(const bool do_ghost_penalty, const unsigned int n_subdivisions)
{
  const unsigned int n_components = 1;
  const unsigned int fe_degree    = 3;
  using VectorType                = Vector<double>;

  GDM::System<dim> system(fe_degree, n_components, do_ghost_penalty);
  system.subdivided_hyper_cube(n_subdivisions, -1.21, 1.21);

  hp::MappingCollection<dim> mapping;
  mapping.push_back(MappingQ1<dim>());
  hp::QCollection<dim> quadrature;
  quadrature.push_back(QGauss<dim>(fe_degree + 1));

  AffineConstraints<double> constraints;
  constraints.close();
  system.categorize();

  // level set, classification, cut quadrature, assembly of the rows around the surface
  typename GDM::CutCellSetup<dim>::Parameters prm;
  prm.ghost_penalty     = do_ghost_penalty;
  prm.ghost_parameter   = 0.5;
  prm.nitsche_parameter = 5 * (fe_degree + 1) * fe_degree;
  GDM::CutCellSetup<dim> cut(fe_degree, n_subdivisions, -1.21, 1.21, SignedDistanceSphere<dim>(), prm);

  SparseMatrix<double> stiffness_matrix;
  GDM::MatrixCreator::create_laplace_matrix(mapping, system, quadrature, stiffness_matrix, constraints);
  cut.attach_to(stiffness_matrix);

  VectorType solution(system), rhs(system);
  rhs.from_host(cut.rhs());

  ReductionControl     solver_control(system.n_dofs(), 1.e-10, 1.e-6);
  SolverCG<VectorType> solver(solver_control);
  solver.solve(stiffness_matrix, solution, rhs, PreconditionIdentity());

  const double error_L2 = cut.l2_error_inside(solution.to_host(), AnalyticalSolution<dim>());
  std::cout << std::endl << "Mesh size  L2-Error  " << std::endl;
  printf("   %.4f %.4e \n\n", 2.42 / n_subdivisions, error_L2);
}

int main(int argc, char **argv)
{
  const unsigned int n = argc > 1 ? std::atoi(argv[1]) : 64;
  test<2>(false, n);
  test<2>(true, n);
}
