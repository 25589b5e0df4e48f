// gdm/gdm.h -- C++ front end of the B200-native GDM hot path (header only, over the C ABI in
// gdm/cuda/gdm_c_api.h).
//
// It keeps the names and call shapes of the reference's include/gdm API so that drivers written
// like its tests / prototypes compile against it without deal.II:
//
//   reference (deal.II)                                   here
//   ---------------------------------------------------   -------------------------------------------
//   GDM::System<dim>                   system.h:339-827    GDM::System<dim>
//   GDM::generate_polynomials_1D       fe.h:55-336         GDM::generate_polynomials_1D
//   AffineConstraints<double>          (deal.II)           dealii::AffineConstraints<double>
//   Vector<double> / LA::distributed::Vector<double>       dealii::Vector<double>  (device resident)
//   SparseMatrix<double> / Trilinos SparseMatrix           dealii::SparseMatrix<double> (matrix-free operator)
//   GDM::MatrixCreator::create_mass_matrix   matrix_creator.h:9-62    same name and argument order
//   GDM::MatrixCreator::create_lumped_mass_matrix   :64-117           same
//   (inline stiffness loop, tests/poisson_02_gdm.cc:160-206)          GDM::MatrixCreator::create_laplace_matrix
//   GDM::VectorTools::interpolate / integrate_difference  vector_tools.h:11-86   same
//   SolverCG, ReductionControl, PreconditionIdentity/Jacobi, DiagonalMatrix      same names
//   TimeStepping::ExplicitRungeKutta, DiscreteTime                               same names
//
// Errors: where deal.II throws ExcNotImplemented / SolverControl::NoConvergence these classes throw
// dealii::ExcNotImplemented / dealii::SolverControl::NoConvergence (std::exception based).
// Everything computes on the GPU through libgdm_b200.so; there is no host fallback.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "cuda/gdm_c_api.h"

namespace dealii
{
  struct ExcNotImplemented : std::runtime_error
  {
    using std::runtime_error::runtime_error;
  };
  struct ExcMessage : std::runtime_error
  {
    using std::runtime_error::runtime_error;
  };
  namespace SolverControlNS
  {
    struct NoConvergence : std::runtime_error
    {
      unsigned int last_step;
      double       last_residual;
      NoConvergence(unsigned int s, double r, const std::string &m)
        : std::runtime_error(m)
        , last_step(s)
        , last_residual(r)
      {}
    };
  } // namespace SolverControlNS

  namespace internal
  {
    inline void check(int rc)
    {
      if (rc == GDM_OK)
        return;
      const std::string msg = gdm_last_error();
      if (rc == GDM_ERR_NOT_IMPLEMENTED)
        throw ExcNotImplemented(msg);
      throw ExcMessage("[gdm status " + std::to_string(rc) + "] " + msg);
    }

    // one context per process / device
    inline gdm_context_t default_context(int device = 0)
    {
      static gdm_context_t ctx = nullptr;
      if (!ctx)
        check(gdm_context_create(device, nullptr, &ctx));
      return ctx;
    }
  } // namespace internal

  template <int dim>
  struct Point
  {
    double       x[dim > 0 ? dim : 1] = {};
    double      &operator[](int i) { return x[i]; }
    const double &operator[](int i) const { return x[i]; }
    double       operator()(int i) const { return x[i]; }
  };

  template <int dim, typename Number = double>
  class Function
  {
  public:
    explicit Function(unsigned int n_components = 1, double time = 0.0)
      : n_components(n_components)
      , time(time)
    {}
    virtual ~Function() = default;
    virtual double value(const Point<dim> &p, const unsigned int component = 0) const = 0;
    void   set_time(double t) { time = t; }
    double get_time() const { return time; }
    const unsigned int n_components;

  private:
    double time;
  };

  template <int dim>
  struct MappingQ1
  {};
  template <int dim>
  struct QGauss
  {
    explicit QGauss(unsigned int n)
      : n_points(n)
    {}
    unsigned int n_points;
  };
  namespace hp
  {
    template <int dim>
    struct MappingCollection
    {
      void push_back(const MappingQ1<dim> &) { ++n; }
      int  n = 0;
    };
    template <int dim>
    struct QCollection
    {
      void         push_back(const QGauss<dim> &q) { n_points = q.n_points; }
      unsigned int n_points = 0;
    };
  } // namespace hp
} // namespace dealii

namespace GDM
{
  template <int dim>
  class System;

  // fe.h:55-336 -- [variant][basis function] -> monomial coefficients, lowest power first
  inline std::vector<std::vector<std::vector<double>>> generate_polynomials_1D(const unsigned int fe_degree)
  {
    const unsigned int  p = fe_degree;
    std::vector<double> c((size_t)p * (p + 1) * (p + 1));
    dealii::internal::check(gdm_polynomials_1d((int)p, c.data()));
    std::vector<std::vector<std::vector<double>>> out(p, std::vector<std::vector<double>>(p + 1, std::vector<double>(p + 1)));
    for (unsigned v = 0; v < p; ++v)
      for (unsigned k = 0; k <= p; ++k)
        for (unsigned i = 0; i <= p; ++i)
          out[v][k][i] = c[(v * (p + 1) + k) * (p + 1) + i];
    return out;
  }

  // fe.h:339-397 -- lexicographic index <-> multi-index, x fastest: the layout of the global vector
  template <int dim>
  std::array<unsigned int, dim> index_to_indices(const unsigned int index, const std::array<unsigned int, dim> Ns)
  {
    std::array<unsigned int, dim> indices{};
    unsigned int                  r = index;
    for (int d = 0; d < dim; ++d)
      {
        indices[d] = (d + 1 < dim) ? r % Ns[d] : r;
        r /= Ns[d];
      }
    return indices;
  }
  template <int dim>
  std::array<unsigned int, dim> index_to_indices(const unsigned int index, const unsigned int N)
  {
    std::array<unsigned int, dim> Ns;
    Ns.fill(N);
    return index_to_indices<dim>(index, Ns);
  }
  template <int dim>
  unsigned int indices_to_index(const std::array<unsigned int, dim> indices, const std::array<unsigned int, dim> Ns)
  {
    unsigned int index = 0, stride = 1;
    for (int d = 0; d < dim; ++d)
      {
        index += indices[d] * stride;
        stride *= Ns[d];
      }
    return index;
  }
  template <int dim>
  unsigned int indices_to_index(const std::array<unsigned int, dim> index, const unsigned int N)
  {
    std::array<unsigned int, dim> Ns;
    Ns.fill(N);
    return indices_to_index<dim>(index, Ns);
  }
} // namespace GDM

namespace dealii
{
  // ---------------------------------------------------------------------------- AffineConstraints
  template <typename Number = double>
  class AffineConstraints
  {
  public:
    AffineConstraints() = default;
    AffineConstraints(const AffineConstraints &) = delete;
    ~AffineConstraints()
    {
      if (h)
        gdm_constraints_destroy(h);
    }
    void close()
    {
      closed = true;
      if (h)
        internal::check(gdm_constraints_close(h));
    }
    bool is_constrained(unsigned long long i) const { return h && gdm_constraints_is_constrained(h, i); }
    unsigned long long n_constraints() const { return h ? gdm_constraints_n_constraints(h) : 0; }
    template <class VectorType>
    void distribute(VectorType &v) const
    {
      if (h)
        internal::check(gdm_constraints_distribute(h, v.handle()));
    }
    template <class VectorType>
    void set_zero(VectorType &v) const
    {
      if (h)
        internal::check(gdm_constraints_set_zero(h, v.handle()));
    }
    // the right-hand side part of distribute_local_to_global(cell_matrix, cell_rhs, dofs, A, rhs) with inhomogeneous
    // constraints (tests/poisson_02_gdm.cc:201): b_i -= sum_j A_ij g_j on free rows, b_j = diag_j g_j on constrained rows
    template <class MatrixType, class VectorType>
    void condense_rhs(const MatrixType &A, VectorType &rhs) const
    {
      if (h)
        internal::check(gdm_constraints_condense_rhs(h, A.handle(), rhs.handle()));
    }
    // used by GDM::System
    void bind(gdm_system_t sys) const
    {
      if (!h)
        internal::check(gdm_constraints_create(sys, &h));
    }
    gdm_constraints_t handle() const { return h; }
    bool              is_closed() const { return closed; }

  private:
    mutable gdm_constraints_t h = nullptr;
    bool                      closed = false;
  };

  // ---------------------------------------------------------------------------- Vector
  template <typename Number = double>
  class Vector
  {
    static_assert(std::is_same<Number, double>::value, "the GPU path is FP64");

  public:
    using value_type = double;
    Vector() = default;
    template <int dim>
    explicit Vector(const GDM::System<dim> &system)
    {
      reinit(system);
    }
    Vector(const Vector &o) { *this = o; }
    Vector &operator=(const Vector &o)
    {
      if (this != &o && o.h)
        {
          if (!h || sys != o.sys)
            reinit_raw(o.sys, o.n);
          internal::check(gdm_vector_copy(h, o.h));
        }
      return *this;
    }
    ~Vector() { clear(); }
    template <int dim>
    void reinit(const GDM::System<dim> &system)
    {
      reinit_raw(system.handle(), system.n_locally_owned_dofs());
    }
    void reinit(const Vector &o) { reinit_raw(o.sys, o.n); }
    std::size_t size() const { return n; }
    Vector &operator=(const double s)
    {
      internal::check(gdm_vector_set(h, s));
      return *this;
    }
    void add(const double a, const Vector &x) { internal::check(gdm_vector_add(h, a, x.h)); }
    void sadd(const double s, const double a, const Vector &x) { internal::check(gdm_vector_sadd(h, s, a, x.h)); }
    void equ(const double a, const Vector &x)
    {
      *this = x;
      if (a != 1.0)
        internal::check(gdm_vector_scale(h, a));
    }
    Vector &operator*=(const double a)
    {
      internal::check(gdm_vector_scale(h, a));
      return *this;
    }
    void   scale(const Vector &d) { internal::check(gdm_vector_scale_by(h, d.h)); }
    double operator*(const Vector &o) const
    {
      double r;
      internal::check(gdm_vector_dot(h, o.h, &r));
      return r;
    }
    double l2_norm() const
    {
      double r;
      internal::check(gdm_vector_l2_norm(h, &r));
      return r;
    }
    double linfty_norm() const
    {
      double r;
      internal::check(gdm_vector_linfty_norm(h, &r));
      return r;
    }
    void update_ghost_values() const { internal::check(gdm_vector_update_ghost_values(h)); }
    // host transfer (locally owned DoFs, lexicographic)
    std::vector<double> to_host() const
    {
      std::vector<double> v(n);
      internal::check(gdm_vector_download(h, v.data()));
      return v;
    }
    void from_host(const std::vector<double> &v) { internal::check(gdm_vector_upload(h, v.data())); }
    gdm_vector_t handle() const { return h; }
    gdm_system_t system_handle() const { return sys; }
    // non-owning view of a library-owned vector (Runge-Kutta stage vectors)
    static Vector view(gdm_vector_t handle, gdm_system_t sys, std::size_t n)
    {
      Vector v;
      v.h = handle;
      v.sys = sys;
      v.n = n;
      v.owns = false;
      return v;
    }

  private:
    void clear()
    {
      if (h && owns)
        gdm_vector_destroy(h);
      h = nullptr;
    }
    void reinit_raw(gdm_system_t s, std::size_t n_)
    {
      clear();
      sys  = s;
      n    = n_;
      owns = true;
      internal::check(gdm_vector_create(sys, &h));
    }
    gdm_vector_t h = nullptr;
    gdm_system_t sys = nullptr;
    std::size_t  n = 0;
    bool         owns = true;
  };

  // ---------------------------------------------------------------------------- SparseMatrix (operator)
  // The operator is matrix free: a sparsity pattern is a *view* whose rows are generated on demand by the C ABI
  // (gdm_system_sparsity_row; System::create_sparsity_pattern / create_flux_sparsity_pattern, system.h:586-630).  At the
  // BASELINE size the stored pattern would need 70 GB; row_length / column_number / n_nonzero_elements work at any size.
  struct DynamicSparsityPattern
  {
    explicit DynamicSparsityPattern(std::size_t = 0) {}
    template <class IndexSet>
    explicit DynamicSparsityPattern(const IndexSet &)
    {}
    void bind(gdm_system_t s, bool flux_, std::size_t n)
    {
      sys    = s;
      flux   = flux_;
      n_dofs = n;
    }
    std::size_t n_rows() const { return n_dofs; }
    std::size_t n_cols() const { return n_dofs; }
    unsigned int row_length(const std::size_t row) const
    {
      uint64_t n = 0;
      internal::check(gdm_system_sparsity_row(sys, flux ? 1 : 0, row, nullptr, 0, &n));
      return (unsigned int)n;
    }
    std::vector<unsigned long long> row(const std::size_t r) const
    {
      uint64_t n = row_length(r);
      std::vector<uint64_t> c(n);
      internal::check(gdm_system_sparsity_row(sys, flux ? 1 : 0, r, c.data(), n, &n));
      return std::vector<unsigned long long>(c.begin(), c.end());
    }
    unsigned long long column_number(const std::size_t r, const unsigned int k) const { return row(r)[k]; }
    bool exists(const std::size_t r, const std::size_t c) const
    {
      const auto cols = row(r);
      return std::binary_search(cols.begin(), cols.end(), (unsigned long long)c);
    }
    unsigned long long n_nonzero_elements() const
    {
      unsigned long long n = 0;
      for (std::size_t r = 0; r < n_dofs; ++r)
        n += row_length(r);
      return n;
    }
    gdm_system_t sys = nullptr;
    bool         flux = false;
    std::size_t  n_dofs = 0;
  };
  struct SparsityPattern : DynamicSparsityPattern
  {
    void copy_from(const DynamicSparsityPattern &d) { static_cast<DynamicSparsityPattern &>(*this) = d; }
  };

  template <typename Number = double>
  class SparseMatrix
  {
  public:
    using value_type = double;
    SparseMatrix() = default;
    SparseMatrix(const SparseMatrix &) = delete;
    ~SparseMatrix()
    {
      if (h)
        gdm_operator_destroy(h);
    }
    void reinit(const SparsityPattern &) {} // the operator is matrix free: nothing to allocate
    void create(gdm_system_t sys, gdm_constraints_t c, const gdm_operator_desc &d)
    {
      if (h)
        gdm_operator_destroy(h);
      h = nullptr;
      internal::check(gdm_operator_create(sys, c, &d, &h));
    }
    void vmult(Vector<double> &dst, const Vector<double> &src) const
    {
      internal::check(gdm_operator_vmult(h, dst.handle(), src.handle()));
    }
    void vmult_add(Vector<double> &dst, const Vector<double> &src) const
    {
      internal::check(gdm_operator_vmult_add(h, dst.handle(), src.handle()));
    }
    void Tvmult(Vector<double> &dst, const Vector<double> &src) const
    {
      internal::check(gdm_operator_tvmult(h, dst.handle(), src.handle()));
    }
    // dst = M^-1 src for a mass matrix on a Cartesian grid (Kronecker-direct banded line solves; replaces the
    // preconditioned CG of applications/advection/include/gdm/advection/problem.h:236-267)
    void mass_inverse_vmult(Vector<double> &dst, const Vector<double> &src) const
    {
      internal::check(gdm_operator_mass_inverse(h, dst.handle(), src.handle()));
    }
    // irregular rows (cut cells, ghost penalty, Nitsche) that replace the tensor-product rows: the part of the
    // assembled matrices of prototypes/cut_poisson_01_gdm.cc:196-329 / wave/{mass,stiffness}.h that is not Kronecker
    void attach_irregular_rows(const std::vector<uint64_t> &row_ids, const std::vector<uint64_t> &rowptr,
                               const std::vector<uint64_t> &col, const std::vector<double> &val)
    {
      internal::check(gdm_operator_attach_csr(h, row_ids.size(), row_ids.data(), rowptr.data(), col.data(), val.data()));
    }
    unsigned long long m() const { return gdm_operator_m(h); }
    unsigned long long n() const { return gdm_operator_m(h); }
    gdm_operator_t     handle() const { return h; }

  private:
    gdm_operator_t h = nullptr;
  };

  // ---------------------------------------------------------------------------- solvers
  class ReductionControl
  {
  public:
    ReductionControl(unsigned int n = 100, double tol = 1e-10, double reduce = 1e-2)
    {
      c.max_steps = n;
      c.tolerance = tol;
      c.reduce = reduce;
      c.last_step = 0;
      c.last_value = c.initial_value = 0;
    }
    unsigned int           last_step() const { return c.last_step; }
    double                 last_value() const { return c.last_value; }
    double                 initial_value() const { return c.initial_value; }
    gdm_reduction_control &raw() { return c; }

  private:
    gdm_reduction_control c;
  };
  struct SolverControl
  {
    using NoConvergence = SolverControlNS::NoConvergence;
  };

  struct PreconditionIdentity
  {
    int          kind() const { return GDM_PRECONDITION_IDENTITY; }
    gdm_vector_t vec() const { return nullptr; }
  };
  template <class MatrixType = SparseMatrix<double>>
  struct PreconditionJacobi
  {
    void         initialize(const MatrixType &) {}
    int          kind() const { return GDM_PRECONDITION_JACOBI; }
    gdm_vector_t vec() const { return nullptr; }
  };
  template <class VectorType = Vector<double>>
  class DiagonalMatrix
  {
  public:
    VectorType  &get_vector() { return diag; }
    void         vmult(VectorType &dst, const VectorType &src) const
    {
      dst = src;
      dst.scale(diag);
    }
    int          kind() const { return GDM_PRECONDITION_DIAGONAL; }
    gdm_vector_t vec() const { return diag.handle(); }

  private:
    VectorType diag;
  };

  template <class VectorType = Vector<double>>
  class SolverCG
  {
  public:
    explicit SolverCG(ReductionControl &c)
      : control(c)
    {}
    template <class MatrixType, class Preconditioner>
    void solve(const MatrixType &A, VectorType &x, const VectorType &b, const Preconditioner &P)
    {
      const int rc = gdm_solver_cg(A.handle(), x.handle(), b.handle(), P.kind(), P.vec(), &control.raw());
      if (rc == GDM_ERR_NO_CONVERGENCE)
        throw SolverControl::NoConvergence(control.last_step(), control.last_value(), gdm_last_error());
      internal::check(rc);
    }

  private:
    ReductionControl &control;
  };

  // ---------------------------------------------------------------------------- time stepping
  class DiscreteTime
  {
  public:
    DiscreteTime(double start, double end, double step)
      : end(end)
      , desired(step)
      , current(start)
      , next(calc(start))
    {}
    bool   is_at_end() const { return current == end; }
    double get_current_time() const { return current; }
    double get_next_time() const { return next; }
    double get_next_step_size() const { return next - current; }
    unsigned int get_step_number() const { return n; }
    void   advance_time()
    {
      ++n;
      current = next;
      next    = calc(current);
    }

  private:
    double calc(double t) const
    {
      double nx = t + desired;
      if (nx > end - 0.05 * desired)
        nx = end;
      return nx;
    }
    double       end, desired, current, next;
    unsigned int n = 0;
  };

  namespace TimeStepping
  {
    enum runge_kutta_method
    {
      FORWARD_EULER = GDM_RK_FORWARD_EULER,
      RK_THIRD_ORDER = GDM_RK_THIRD_ORDER,
      RK_CLASSIC_FOURTH_ORDER = GDM_RK_CLASSIC_FOURTH_ORDER
    };

    template <class VectorType = Vector<double>>
    class ExplicitRungeKutta
    {
    public:
      ExplicitRungeKutta() = default;
      ~ExplicitRungeKutta()
      {
        if (h)
          gdm_rk_destroy(h);
      }
      void initialize(runge_kutta_method m) { method = m; }
      // f(t, y) returns f by value as in the reference (wave/problem.h:302-320); the copy into the stage
      // vector is one device-to-device transfer.  f(t, y, out) avoids it.
      double evolve_one_time_step(const std::function<VectorType(const double, const VectorType &)> &f, double t,
                                  double dt, VectorType &y)
      {
        struct Ctx
        {
          const std::function<VectorType(const double, const VectorType &)> *f;
          gdm_system_t                                                         sys;
          std::size_t                                                          n;
          std::string                                                          error;
        } ctx{&f, y.system_handle(), y.size(), {}};
        if (!h)
          internal::check(gdm_rk_create(ctx.sys, method, 1, &h));
        auto tramp = [](double tt, const gdm_vector_t *yy, gdm_vector_t *out, void *user) -> int {
          Ctx *c = static_cast<Ctx *>(user);
          try
            {
              const VectorType yin = VectorType::view(yy[0], c->sys, c->n);
              VectorType       o   = VectorType::view(out[0], c->sys, c->n);
              const VectorType r   = (*c->f)(tt, yin);
              o                    = r;
              return GDM_OK;
            }
          catch (const std::exception &e)
            {
              c->error = e.what();
              return GDM_ERR_INTERNAL;
            }
        };
        gdm_vector_t yv = y.handle();
        double       t_new = t;
        const int    rc = gdm_rk_evolve_one_time_step(h, tramp, &ctx, t, dt, &yv, &t_new);
        if (rc != GDM_OK && !ctx.error.empty())
          throw ExcMessage(ctx.error);
        internal::check(rc);
        return t_new;
      }

    private:
      gdm_rk_t            h = nullptr;
      runge_kutta_method  method = RK_CLASSIC_FOURTH_ORDER;
    };
  } // namespace TimeStepping
} // namespace dealii

namespace GDM
{
  using namespace dealii;

  // ------------------------------------------------------------------------------ System (system.h:339-827)
  template <int dim>
  class System
  {
  public:
    System(const unsigned int fe_degree, const unsigned int n_components, const bool add_ghost_layer = false)
      : fe_degree(fe_degree)
      , n_components(n_components)
      , add_ghost_layer(add_ghost_layer)
    {}
    System(const System &) = delete;
    ~System()
    {
      if (h)
        gdm_system_destroy(h);
    }
    // multi-GPU: call before creating the grid (the reference passes MPI_COMM_WORLD to the constructor)
    void set_partition(int rank_, int n_ranks_)
    {
      rank    = rank_;
      n_ranks = n_ranks_;
    }
    void subdivided_hyper_cube(const unsigned int n, const double left = 0.0, const double right = 1.0)
    {
      std::vector<unsigned int> reps(dim, n);
      Point<dim>                p1, p2;
      for (int d = 0; d < dim; ++d)
        {
          p1[d] = left;
          p2[d] = right;
        }
      subdivided_hyper_rectangle(reps, p1, p2);
    }
    void subdivided_hyper_rectangle(const std::vector<unsigned int> &repetitions, const Point<dim> &p1, const Point<dim> &p2)
    {
      gdm_system_desc d{};
      d.dim          = dim;
      d.fe_degree    = (int)fe_degree;
      d.n_components = (int)n_components;
      for (int i = 0; i < 3; ++i)
        {
          d.n_subdivisions[i] = i < dim ? repetitions[i] : 0;
          d.lo[i]             = i < dim ? p1[i] : 0.0;
          d.hi[i]             = i < dim ? p2[i] : 1.0;
        }
      d.rank            = rank;
      d.n_ranks         = n_ranks;
      d.add_ghost_layer = add_ghost_layer;
      internal::check(gdm_system_create(internal::default_context(), &d, &h));
    }
    void categorize() {} // categories are implicit in the band tables (system.h:404-424)
    unsigned long long n_dofs() const { return gdm_system_n_dofs(h); }
    unsigned long long n_cells() const { return gdm_system_n_cells(h); }
    std::size_t        n_locally_owned_dofs() const
    {
      uint64_t b, e;
      internal::check(gdm_system_locally_owned_range(h, &b, &e));
      return (std::size_t)(e - b);
    }
    unsigned int get_fe_degree() const { return fe_degree; }
    void make_zero_boundary_constraints(AffineConstraints<double> &c) const
    {
      c.bind(h);
      internal::check(gdm_constraints_make_zero_boundary(c.handle(), -1));
    }
    void make_zero_boundary_constraints(const unsigned int surface, AffineConstraints<double> &c) const
    {
      c.bind(h);
      internal::check(gdm_constraints_make_zero_boundary(c.handle(), (int)surface));
    }
    void make_periodicity_constraints(const unsigned int d, AffineConstraints<double> &c) const
    {
      c.bind(h);
      internal::check(gdm_constraints_make_periodicity(c.handle(), (int)d));
    }
    // system.h:511-547: boundary nodes (boundary id 0 = every face of the hyper rectangle) take the value of fu
    void interpolate_boundary_values(const hp::MappingCollection<dim> &, const unsigned int bid, const Function<dim> &fu,
                                     AffineConstraints<double> &c) const
    {
      c.bind(h);
      const Function<dim> *f = &fu;
      internal::check(gdm_constraints_interpolate_boundary_values(
        c.handle(), (int)bid,
        [](const double *pt, int comp, void *user) -> double {
          Point<dim> p;
          for (int d = 0; d < dim; ++d)
            p[d] = pt[d];
          return (*static_cast<const Function<dim> **>(user))->value(p, comp);
        },
        &f));
    }
    // system.h:586-599 / 602-630: the pattern object becomes a view of this system (rows on demand)
    template <class SP>
    void create_sparsity_pattern(const AffineConstraints<double> &, SP &dsp) const
    {
      dsp.bind(h, false, n_dofs());
    }
    template <class SP>
    void create_flux_sparsity_pattern(const AffineConstraints<double> &, SP &dsp) const
    {
      if (!add_ghost_layer)
        throw ExcNotImplemented("create_flux_sparsity_pattern needs add_ghost_layer (system.h:605)");
      dsp.bind(h, true, n_dofs());
    }
    void get_dof_indices(unsigned long long cell, std::vector<unsigned long long> &out) const
    {
      out.resize(gdm_system_dofs_per_cell(h));
      std::vector<uint64_t> t(out.size());
      internal::check(gdm_system_get_dof_indices(h, cell, t.data()));
      out.assign(t.begin(), t.end());
    }
    gdm_system_t handle() const { return h; }

  private:
    const unsigned int fe_degree, n_components;
    const bool         add_ghost_layer;
    int                rank = 0, n_ranks = 1;
    gdm_system_t       h = nullptr;
  };

  namespace internal
  {
    template <int dim>
    inline void check_setup(const System<dim> &system, const hp::QCollection<dim> &q)
    {
      if (q.n_points != 0 && q.n_points != system.get_fe_degree() + 1)
        throw ExcNotImplemented("only QGauss(fe_degree + 1) is implemented");
    }
    inline gdm_constraints_t handle_of(const AffineConstraints<double> &c)
    {
      if (c.handle() && !c.is_closed())
        throw ExcMessage("constraints must be closed");
      return c.handle();
    }
  } // namespace internal

  // ------------------------------------------------------------------------------ MatrixCreator
  namespace MatrixCreator
  {
    // matrix_creator.h:9-62
    template <int dim, typename SparseMatrixType>
    void create_mass_matrix(const hp::MappingCollection<dim> &, const System<dim> &system, const hp::QCollection<dim> &quadrature,
                            SparseMatrixType &sparse_matrix, const AffineConstraints<double> &constraints)
    {
      GDM::internal::check_setup(system, quadrature);
      gdm_operator_desc d{};
      d.kind = GDM_OP_MASS;
      d.scale = 1.0;
      d.constrained_diagonal = GDM_DIAG_ASSEMBLED;
      sparse_matrix.create(system.handle(), GDM::internal::handle_of(constraints), d);
    }
    // the assembly loop of tests/poisson_02_gdm.cc:160-206
    template <int dim, typename SparseMatrixType>
    void create_laplace_matrix(const hp::MappingCollection<dim> &, const System<dim> &system, const hp::QCollection<dim> &quadrature,
                               SparseMatrixType &sparse_matrix, const AffineConstraints<double> &constraints)
    {
      GDM::internal::check_setup(system, quadrature);
      gdm_operator_desc d{};
      d.kind = GDM_OP_STIFFNESS;
      d.scale = 1.0;
      d.constrained_diagonal = GDM_DIAG_ASSEMBLED;
      sparse_matrix.create(system.handle(), GDM::internal::handle_of(constraints), d);
    }
    // scale * (phi_i, b . grad phi_j): prototypes/advection_01_gdm.cc:164-206 uses scale = -1
    template <int dim, typename SparseMatrixType, typename Velocity>
    void create_advection_matrix(const hp::MappingCollection<dim> &, const System<dim> &system, const hp::QCollection<dim> &quadrature,
                                 SparseMatrixType &sparse_matrix, const AffineConstraints<double> &constraints,
                                 const Velocity &velocity, const double scale = 1.0, const bool transpose = false)
    {
      GDM::internal::check_setup(system, quadrature);
      gdm_operator_desc d{};
      d.kind = transpose ? GDM_OP_ADVECTION_T : GDM_OP_ADVECTION;
      d.scale = scale;
      for (int i = 0; i < dim; ++i)
        d.b[i] = velocity[i];
      d.constrained_diagonal = GDM_DIAG_ZERO;
      sparse_matrix.create(system.handle(), GDM::internal::handle_of(constraints), d);
    }
    // matrix_creator.h:64-117: vector <- 1 / lumped mass
    template <int dim, typename VectorType>
    void create_lumped_mass_matrix(const hp::MappingCollection<dim> &mapping, const System<dim> &system,
                                   const hp::QCollection<dim> &quadrature, VectorType &vector,
                                   const AffineConstraints<double> &constraints)
    {
      SparseMatrix<double> M;
      create_mass_matrix(mapping, system, quadrature, M, constraints);
      dealii::internal::check(gdm_operator_lumped_mass_inverse(M.handle(), vector.handle()));
    }
  } // namespace MatrixCreator

  // ------------------------------------------------------------------------------ VectorTools (vector_tools.h)
  namespace VectorTools
  {
    enum NormType
    {
      L2_norm
    };
    template <int dim>
    struct FnCtx
    {
      const Function<dim> *f;
    };
    template <int dim>
    inline double fn_trampoline(const double *pt, int comp, void *user)
    {
      Point<dim> p;
      for (int d = 0; d < dim; ++d)
        p[d] = pt[d];
      return static_cast<FnCtx<dim> *>(user)->f->value(p, comp);
    }
    // vector_tools.h:11-23
    template <typename VectorType, int dim>
    void interpolate(const hp::MappingCollection<dim> &, const System<dim> &system, const Function<dim> &function, VectorType &vec)
    {
      FnCtx<dim> c{&function};
      dealii::internal::check(gdm_interpolate(system.handle(), &fn_trampoline<dim>, &c, vec.handle()));
    }
    // vector_tools.h:25-86 (L2 only, :35); difference = cell-wise errors
    template <int dim, class VectorType>
    void integrate_difference(const hp::MappingCollection<dim> &, const System<dim> &system, const VectorType &fe_function,
                              const Function<dim> &exact_solution, std::vector<double> &difference,
                              const hp::QCollection<dim> &quadrature, const NormType &norm)
    {
      if (norm != L2_norm)
        throw ExcNotImplemented("only L2_norm (vector_tools.h:35)");
      GDM::internal::check_setup(system, quadrature);
      difference.assign(system.n_cells(), 0.0);
      FnCtx<dim> c{&exact_solution};
      double     g = 0;
      dealii::internal::check(gdm_integrate_difference(system.handle(), fe_function.handle(), &fn_trampoline<dim>, &c,
                                                       difference.data(), &g));
    }
    // dealii::VectorTools::compute_global_error for the L2 norm
    inline double compute_global_error(const std::vector<double> &cellwise, const NormType & = L2_norm)
    {
      double s = 0;
      for (double v : cellwise)
        s += v * v;
      return std::sqrt(s);
    }
  } // namespace VectorTools

  // ----------------------------------------------------------------------- cut-cell set-up
  // What prototypes/cut_poisson_01_gdm.cc does with NonMatching::{MeshClassifier, FEValues} and FEInterfaceValues
  // before the solve (:105-121 classification of a Q1 level set, :176-190 cut quadrature, :196-329 assembly,
  // :349-398 error), as one host-side object over gdm_cut_* whose rows go to SparseMatrix::attach_irregular_rows.
  // kind_mass = true gives the cut mass matrix of applications/wave/include/gdm/wave/mass.h:47-249.
  template <int dim>
  class CutCellSetup
  {
  public:
    struct Parameters
    {
      bool   ghost_penalty     = true;
      int    gp_h_power        = 1;    // 1: the prototype; 3: the wave application's matrices
      double ghost_parameter   = 0.5;
      double nitsche_parameter = -1.0; // default 5 (p+1) p
      double rhs_value         = 4.0;
      double boundary_value    = 1.0;
      bool   kind_mass         = false;
      double outside_diagonal  = 1.0;
      bool   surface_terms         = true;  // false: no Nitsche on the cut surface (two-domain runs couple there)
      bool   domain_boundary_terms = false; // Nitsche on the box boundary (function_domain_dbc, wave/stiffness.h:262-340)
      bool   negate_level_set      = false; // the domain is {level set > 0}: the outer field of the two-domain runs
      unsigned long long row_begin = 0, row_end = 0; // locally owned rows of this rank; 0, 0 = all
      int    level_set_degree      = 1;     // > 1 (2D): FE_Q(q) level set as in the 2D presets of applications/wave
    };
    CutCellSetup(const unsigned int fe_degree, const unsigned int n_subdivisions, const double left, const double right,
                 const Function<dim> &level_set_function, const Parameters &prm = Parameters())
    {
      gdm_cut_desc d{};
      d.dim       = dim;
      d.fe_degree = (int)fe_degree;
      std::size_t n_nodes = 1;
      for (int e = 0; e < dim; ++e)
        {
          d.n_subdivisions[e] = n_subdivisions;
          d.lo[e]             = left;
          d.hi[e]             = right;
          n_nodes *= n_subdivisions + 1;
        }
      d.ghost_penalty     = prm.ghost_penalty;
      d.gp_h_power        = prm.gp_h_power;
      d.ghost_parameter   = prm.ghost_parameter;
      d.nitsche_parameter = prm.nitsche_parameter >= 0 ? prm.nitsche_parameter : 5.0 * (fe_degree + 1) * fe_degree;
      d.rhs_value         = prm.rhs_value;
      d.boundary_value    = prm.boundary_value;
      d.kind              = prm.kind_mass ? 1 : 0;
      d.outside_diagonal  = prm.outside_diagonal;
      d.no_surface_terms      = !prm.surface_terms;
      d.domain_boundary_terms = prm.domain_boundary_terms;
      d.row_begin             = prm.row_begin;
      d.row_end               = prm.row_end;
      d.level_set_degree      = prm.level_set_degree;
      // VectorTools::interpolate of the level set into FE_Q(level_set_degree): values at its support points, x fastest
      uint64_t n_points = 0;
      dealii::internal::check(gdm_cut_level_set_points(&d, &n_points, nullptr));
      std::vector<double> points(n_points * dim), level_set(n_points);
      dealii::internal::check(gdm_cut_level_set_points(&d, &n_points, points.data()));
      for (std::size_t i = 0; i < n_points; ++i)
        {
          Point<dim> x;
          for (int e = 0; e < dim; ++e)
            x[e] = points[i * dim + e];
          level_set[i] = (prm.negate_level_set ? -1.0 : 1.0) * level_set_function.value(x, 0);
        }
      n_dofs = n_nodes;
      dealii::internal::check(gdm_cut_poisson_create(&d, level_set.data(), &cut));
    }
    CutCellSetup(const CutCellSetup &) = delete;
    ~CutCellSetup()
    {
      if (cut)
        gdm_cut_destroy(cut);
    }
    void attach_to(SparseMatrix<double> &matrix) const
    {
      uint64_t n_rows = 0, nnz = 0;
      dealii::internal::check(gdm_cut_sizes(cut, &n_rows, &nnz, nullptr, nullptr));
      std::vector<uint64_t> row_ids(n_rows), rowptr(n_rows + 1), col(nnz);
      std::vector<double>   val(nnz);
      dealii::internal::check(gdm_cut_rows(cut, row_ids.data(), rowptr.data(), col.data(), val.data()));
      matrix.attach_irregular_rows(row_ids, rowptr, col, val);
    }
    std::vector<double> rhs() const
    {
      std::vector<double> b(n_dofs);
      dealii::internal::check(gdm_cut_rhs(cut, b.data()));
      return b;
    }
    // (v, f) over the inside part + <gamma_D / h v - dv/dn, g> on the surface (wave/stiffness.h:186-260); either may be null
    std::vector<double> load_vector(const Function<dim> *f, const Function<dim> *g) const
    {
      std::vector<double>      b(n_dofs);
      VectorTools::FnCtx<dim>  cf{f}, cg{g};
      dealii::internal::check(gdm_cut_load_vector(cut, f ? &VectorTools::fn_trampoline<dim> : nullptr, &cf,
                                                  g ? &VectorTools::fn_trampoline<dim> : nullptr, &cg, b.data()));
      return b;
    }
    double l2_error_inside(const std::vector<double> &solution, const Function<dim> &exact) const
    {
      VectorTools::FnCtx<dim> c{&exact};
      double                  e = 0;
      dealii::internal::check(gdm_cut_l2_error_inside(cut, solution.data(), &VectorTools::fn_trampoline<dim>, &c, &e));
      return e;
    }
    // <gamma_D / h v - dv/dn, g> on the box boundary (the load that belongs to domain_boundary_terms)
    std::vector<double> boundary_load_vector(const Function<dim> &g) const
    {
      std::vector<double>     b(n_dofs);
      VectorTools::FnCtx<dim> c{&g};
      dealii::internal::check(gdm_cut_boundary_load_vector(cut, &VectorTools::fn_trampoline<dim>, &c, b.data()));
      return b;
    }
    // interface coupling of the two-domain runs (wave/stiffness.h:441-574) as a CSR-only operator: `matrix` must have
    // been created with scale 0; which = 0: P_ij = <n . grad phi_i, phi_j>, 1: P^T, 2: Q_ij = <phi_i, phi_j>
    void attach_coupling_to(SparseMatrix<double> &matrix, const int which) const
    {
      uint64_t n_rows = 0, nnz = 0;
      dealii::internal::check(gdm_cut_coupling_rows(cut, which, 0, 0, &n_rows, &nnz, nullptr, nullptr, nullptr, nullptr));
      std::vector<uint64_t> row_ids(n_rows), rowptr(n_rows + 1), col(nnz);
      std::vector<double>   val(nnz);
      dealii::internal::check(gdm_cut_coupling_rows(cut, which, n_rows, nnz, &n_rows, &nnz, row_ids.data(), rowptr.data(),
                                                    col.data(), val.data()));
      matrix.attach_irregular_rows(row_ids, rowptr, col, val);
    }
    // L2, L1, Linf over the inside part (wave/problem.h:531-615)
    std::array<double, 3> error_norms_inside(const std::vector<double> &solution, const Function<dim> &exact) const
    {
      VectorTools::FnCtx<dim> c{&exact};
      std::array<double, 3>   e{};
      dealii::internal::check(gdm_cut_error_norms_inside(cut, solution.data(), &VectorTools::fn_trampoline<dim>, &c, e.data()));
      return e;
    }
    gdm_cut_t handle() const { return cut; }

  private:
    gdm_cut_t   cut    = nullptr;
    std::size_t n_dofs = 0;
  };
} // namespace GDM
