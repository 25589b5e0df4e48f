// Internal structures of libgdm_b200 (not part of the ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gdm/cuda/gdm_c_api.h"

namespace gdm
{
  // ---------------------------------------------------------------- errors
  struct Error : std::runtime_error
  {
    int code;
    Error(int code, const std::string &msg)
      : std::runtime_error(msg)
      , code(code)
    {}
  };

  void set_last_error(const std::string &msg);

#define GDM_CUDA_CHECK(expr)                                                                  \
  do                                                                                          \
    {                                                                                         \
      cudaError_t err__ = (expr);                                                             \
      if (err__ != cudaSuccess)                                                               \
        throw gdm::Error(GDM_ERR_CUDA,                                                        \
                         std::string(#expr) + ": " + cudaGetErrorString(err__) + " (" +       \
                           __FILE__ + ":" + std::to_string(__LINE__) + ")");                  \
    }                                                                                         \
  while (0)

#define GDM_REQUIRE(cond, code, msg)                                        \
  do                                                                        \
    {                                                                       \
      if (!(cond))                                                          \
        throw gdm::Error(code, std::string(msg) + " [" #cond "]");          \
    }                                                                       \
  while (0)

  // ------------------------------------------------------------ host math
  // basis.cpp
  constexpr int MAX_DEGREE = 9;
  void gauss_legendre_01(int n, std::vector<long double> &x, std::vector<long double> &w);
  // values[k], derivs[k] of the p+1 Lagrange functions of variant v at x (cell = [0,1])
  void lagrange_eval(int p, int v, long double x, long double *values, long double *derivs);
  // exact monomial coefficients, lowest power first: coeffs[k*(p+1) + power]
  void lagrange_monomials(int p, int v, double *coeffs);
  // Reference-cell matrices of variant v: M[a*(p+1)+b] = int phi_a phi_b, K = int phi_a' phi_b',
  // C = int phi_a phi_b', f[a] = int phi_a
  struct CellMatrices1D
  {
    std::vector<double> M, K, C, f;
  };
  CellMatrices1D cell_matrices_1d(int p, int v);
  // window offset / variant of cell c on N cells (system.h:209-216, 415-420)
  inline int window_offset(int p, int N, int c)
  {
    return (c < p / 2) ? 0 : (std::min(N, c + p / 2 + 1) - p);
  }
  inline int cell_variant(int p, int N, int c)
  {
    return (c < p / 2) ? c : ((c < N - p / 2) ? (p / 2) : (p + c - N));
  }
  // Band table of the assembled 1D matrix on N cells, unit spacing:
  // band[row*(2p+1) + tap], column = row + tap - p; kind 0 M, 1 K, 2 C; also load vector f.
  void band_matrix_1d(int p, int N, int kind, std::vector<double> &band);
  void load_vector_1d(int p, int N, std::vector<double> &f);

  // ----------------------------------------------------------------- comm
  struct Comm; // comm.cpp (NCCL through dlopen)

  // -------------------------------------------------------------- context
  struct Context
  {
    int          device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t comm_stream = nullptr;
    cudaStream_t face_stream = nullptr; // lowest priority (constrained-row kernel beside the persistent tile kernel)
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr; // host-buffer pipeline (created on first use)
    cudaEvent_t  ev_pipe[2][32] = {};
    cudaEvent_t  ev_a = nullptr, ev_b = nullptr;
    cudaEvent_t  ev_c = nullptr; // face kernel done (kron3d.cu, several ranks)
    int          sm_count = 148;
    uint64_t     launches = 0;
    // reduction scratch
    double  *d_partials = nullptr; // [n_slots][max_blocks]
    double  *d_sums = nullptr;     // small array of device scalars
    void    *d_cg_status = nullptr; // status record of the CG solves (blas1.cu)
    unsigned *d_counters = nullptr;
    double  *h_pinned = nullptr;   // pinned host mirror for scalars
    size_t   partial_capacity = 0;
    Comm    *comm = nullptr;      // owned; destroyed by comm_destroy
    int      rank = 0, n_ranks = 1;
    // scratch vectors for the generic multi-pass apply (doubles)
    double *scratch[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t  scratch_size = 0;

    // pooled work vectors (CG / RK temporaries): avoids cudaMalloc/cudaFree per solve
    std::vector<std::pair<size_t, double *>> pool_free;
    std::vector<std::pair<size_t, double *>> pool_used;
    double *acquire(size_t n_doubles); // zero initialised
    void    release(double *p);

    void ensure_scratch(size_t n);
    ~Context();
  };

  // --------------------------------------------------------------- layout
  struct Layout
  {
    int      dim = 1, p = 1, nc = 1;
    int      N[3] = {0, 0, 0};  // cells per direction (0 in unused directions)
    int      nn[3] = {1, 1, 1}; // global nodes per direction
    double   lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1}, h[3] = {1, 1, 1};
    int      rank = 0, n_ranks = 1;
    int      pdim = 0;          // partitioned direction = dim-1
    int      ghost = 0;         // ghost depth (planes)
    int      own0 = 0, own1 = 0;   // owned node range in pdim (global)
    int      loc0 = 0, loc1 = 0;   // stored node range in pdim (global), includes ghosts
    int      ln[3] = {1, 1, 1};    // stored nodes per direction
    int64_t  pitch = 0, plane = 0, size = 0; // doubles
    int64_t  stride[3] = {0, 0, 0};          // element stride of one node step in direction d
    int64_t  own_off = 0, own_len = 0;       // owned contiguous storage block
    int64_t  n_dofs_global = 0, n_owned = 0; // true DoF counts (no padding)
  };

  struct System
  {
    Context        *ctx;
    gdm_system_desc desc;
    Layout          L;
  };

  struct Constraints
  {
    System *sys;
    bool    dirichlet[3][2] = {{false, false}, {false, false}, {false, false}};
    bool    periodic[3] = {false, false, false};
    bool    closed = false;
    // inhomogeneous Dirichlet values (System::interpolate_boundary_values, system.h:511-547): a vector that holds g at
    // the constrained boundary nodes and zero elsewhere (nullptr: all constraints homogeneous)
    double *d_inhom = nullptr;
    ~Constraints();
  };

  struct Vector
  {
    System *sys;
    double *d = nullptr;
    bool    owns = true;
    ~Vector();
  };

  // Operator = scale * sum_d  B_d (x) prod_{e != d} A_e   (has_B)   or   scale * prod_d A_d
  struct CsrOverlay
  {
    int64_t  n_rows = 0, nnz = 0;
    int64_t *d_row_off = nullptr; // storage offset of each irregular row
    int64_t *d_rowptr = nullptr;
    int32_t *d_col_rel = nullptr; // storage offset of a column relative to its row (12 B per nonzero with the value)
    double  *d_val = nullptr;
    double  *d_diag = nullptr;    // diagonal entry of each irregular row (Jacobi)
    ~CsrOverlay();
  };

  struct Operator
  {
    System *sys;
    gdm_operator_desc desc;
    bool    has_B = false;
    int     b_symmetry = +1;          // +1: B symmetric (stiffness), -1: antisymmetric (advection)
    bool    dirichlet[3][2] = {{false, false}, {false, false}, {false, false}};
    bool    periodic[3] = {false, false, false};
    // host band tables (local rows in pdim): [d] -> ln[d]*(2p+1)
    std::vector<double> hA[3], hB[3];
    std::vector<double> hdiagA[3], hdiagB[3]; // unconstrained 1D diagonals (constrained-row value)
    // periodic directions: the unfolded tables (plain one-sided rows on N+1 nodes); the fused kernel applies
    // C^T A C as "duplicate node 0 into node N, apply A, add row N to row 0" 
    std::vector<double> hAu[3], hBu[3];
    double *dA[3] = {nullptr, nullptr, nullptr};
    double *dB[3] = {nullptr, nullptr, nullptr};
    double *ddiagA[3] = {nullptr, nullptr, nullptr};
    double *ddiagB[3] = {nullptr, nullptr, nullptr};
    int     kernel_used = GDM_KERNEL_GENERIC;
    std::unique_ptr<CsrOverlay> csr;
    void   *fused = nullptr;          // FusedPlan* (kron3d.cu)
    void   *massinv = nullptr;        // MassInvPlan* (massinv.cu), created on first use
    void   *unconstrained = nullptr;  // gdm_operator_s* of the same operator without constraints (lifting of inhomogeneous
                                      // Dirichlet values into the right-hand side), created on first use
    void   *transposed = nullptr;     // gdm_operator_s* of the transposed advection operator (Tvmult), created on first use
    double *tmp = nullptr;            // vmult_add with CSR overlay
    double *host_src = nullptr, *host_dst = nullptr; // staging of vmult_host (padded layout)
    double *stage_src = nullptr, *stage_dst = nullptr; // contiguous staging (host order) for 1D PCIe copies
    Operator() = default;
    Operator(const Operator &) = delete;
    ~Operator();
  };

  // ------------------------------------------------------ kernels (host API)
  // generic.cu
  struct BandPassArgs
  {
    const double *src1 = nullptr, *tab1 = nullptr;
    const double *src2 = nullptr, *tab2 = nullptr;
    double       *dst = nullptr;
    int           dir = 0;
    bool          accumulate = false;
    double        scale = 1.0;
    bool          owned_only = false; // restrict to owned planes of pdim
  };
  void launch_band_pass(Context &ctx, const Layout &L, const bool periodic[3], const BandPassArgs &a);
  // returns the number of thread blocks launched; dot_partials != nullptr: block b also writes the sum of
  // src * dst over its rows to dot_partials[b] (fused dot product of the apply)
  int  launch_constrained_rows(Context &ctx, const Layout &L, const Operator &op, double *dst,
                               const double *src, bool accumulate, int plane_lo = -1, int plane_hi = -1,
                               double *dot_partials = nullptr, cudaStream_t stream = nullptr /* nullptr: ctx.stream */);
  // upper bound of the blocks launch_constrained_rows uses (to size partial buffers)
  int  constrained_rows_max_blocks(const Layout &L);
  void launch_csr_overlay(Context &ctx, const CsrOverlay &csr, double *dst, const double *src,
                          bool accumulate);
  // diag[row] = diagonal entry of the irregular rows (after launch_diagonal of the tensor-product part)
  void launch_csr_diagonal(Context &ctx, const CsrOverlay &csr, double *diag);
  void launch_periodic_copy(Context &ctx, const Layout &L, const bool periodic[3], double *v);
  void launch_set_constrained(Context &ctx, const Layout &L, const bool dirichlet[3][2],
                              const bool periodic[3], double *v, double value);
  void launch_diagonal(Context &ctx, const Layout &L, const Operator &op, double *diag);
  void generic_apply(Operator &op, double *dst, const double *src, bool accumulate);
  void launch_repack(Context &ctx, const Layout &L, double *padded, double *compact, int p0, int p1, bool to_padded);

  // kron3d.cu -- fused tensor-product kernel (dim == 3)
  bool fused_supported(const Operator &op);
  void fused_plan_create(Operator &op);
  void fused_plan_destroy(Operator &op);
  // exchange_ghosts: import the ghost planes of src inside the call, overlapped with the interior planes
  // dot_slot >= 0: additionally leaves <src, A src> over the owned DoFs of this rank in ctx.d_sums[dot_slot]
  // (fused into the store epilogue of the tile kernels; deterministic two-stage sum).  Check fused_supports_dot.
  void fused_apply(Operator &op, double *dst, const double *src, bool accumulate, bool exchange_ghosts = false,
                   int dot_slot = -1);
  bool fused_supports_dot(const Operator &op);
  // output planes [z0, z1) only (local plane indices; no ghost import): building block of the pipelined host-buffer apply
  void fused_apply_window(Operator &op, double *dst, const double *src, int z0, int z1);
  // kron3d_pers.cu -- persistent ramp-free fused kernel (default for dim == 3, scalar, non-periodic)
  bool  pers_supported(const Operator &op);
  void *pers_plan_create(Operator &op); // nullptr: not applicable to this operator
  void  pers_plan_destroy(void *plan);
  int   pers_max_grid(const Operator &op, const void *plan);
  int   pers_max_partials(const Operator &op, const void *plan);
  // massinv.cu: Kronecker-direct inverse of a MASS operator (banded line solves per direction)
  bool  massinv_supported(const Operator &op);
  void  massinv_apply(Operator &op, double *dst, const double *src);
  void  massinv_destroy(Operator &op);
  void  pers_tune(Operator &op, void *plan); // plan search at operator creation (measured; kron3d_pers.cu)
  void  pers_window(const void *plan, int &cz0, int &cz1); // output planes of the whole slab (local indices)
  // output planes [oz0, oz1) (local indices); dot_partials != nullptr: CTA w leaves its share of <dot_src, A src> in
  // dot_partials[w]; returns the number of CTAs launched
  // slots_limit > 0: use at most that many CTAs (launches that share the GPU: slab faces beside the interior planes)
  int   pers_launch(Operator &op, void *plan, double *dst, const double *src, bool accumulate, int oz0, int oz1, cudaStream_t stream,
                    const double *dot_src, double *dot_partials, int slots_limit = 0);
  int   pers_error_flag(void *plan);
  // periodic directions of the persistent path (C^T A C = fold . A . duplicate): pre patches src in place (node N := node 0,
  // old values saved), post folds dst (row 0 += row N) and restores src.  No-ops without periodic directions.
  bool  pers_has_periodic(const void *plan);
  void  pers_periodic_pre(Operator &op, void *plan, double *src, cudaStream_t stream);
  void  pers_periodic_post(Operator &op, void *plan, double *dst, double *src, cudaStream_t stream);
  // mode 0: tile-major sweep, 1: aligned chunks + sweep (one share per CTA, weighted), 2: guided levels (self-scheduling)
  void  pers_partition_host(int tiles_x, int tiles_y, int k0, int k1, int slots, int min_len, int mode, const int *weights,
                            std::vector<int> &job_ptr, std::vector<int> &jobs6, double guide_k = 1.0, int guide_min = 8);

  // blas1.cu
  enum SumSlot
  {
    SUM_DOT = 0,
    SUM_PQ = 1,
    SUM_TMP = 2,
    SUM_CG_PAIR0 = 4, // (r.r, r.z) of even iterations: slots 4,5 ; odd iterations: 6,7
    N_SUM_SLOTS = 16
  };
  inline int cg_rr_slot(unsigned it) { return SUM_CG_PAIR0 + 2 * (int)(it & 1u); }
  inline int cg_rz_slot(unsigned it) { return SUM_CG_PAIR0 + 2 * (int)(it & 1u) + 1; }
  void blas_set(Context &ctx, double *v, int64_t n, double value);
  void blas_set_strided(Context &ctx, const Layout &L, double *v, double value);
  void blas_copy(Context &ctx, double *dst, const double *src, int64_t n);
  void blas_scale(Context &ctx, double *v, int64_t n, double a);
  void blas_sadd(Context &ctx, double *v, double s, double a, const double *x, int64_t n);
  void blas_mul(Context &ctx, double *v, const double *d, int64_t n);
  void blas_invert(Context &ctx, double *v, int64_t n);
  // result lands in ctx.d_sums[slot] (device); max variant for linfty
  void blas_dot(Context &ctx, const double *a, const double *b, int64_t n, int slot);
  // ctx.d_sums[slot] = sum of partials[0..n) in a fixed order (one block)
  void blas_sum_partials(Context &ctx, const double *partials, int n, int slot);
  void blas_absmax(Context &ctx, const double *a, int64_t n, int slot);
  double read_sum(Context &ctx, int slot, bool allreduce_sum, bool is_max = false);

  // CG building blocks (blas1.cu)
  void *cg_status_alloc(Context &ctx);
  void  cg_status_free(void *p);
  void  cg_status_read(Context &ctx, void *d_status, int &done, unsigned &last_step, double &last_value,
                       double &initial);
  const int *cg_status_done_flag(void *d_status);
  void cg_launch_init(Context &ctx, double *r, double *p, const double *dinv, int64_t n, void *status,
                      int rr_new, int rz_new);
  void cg_launch_check0(Context &ctx, void *status, int rr_slot, double tol, double reduce, unsigned max_steps);
  void cg_launch_update(Context &ctx, double *x, double *r, const double *p, const double *q,
                        const double *dinv, int64_t n, void *status, int rz_cur, int rr_new, int rz_new);
  void cg_launch_direction(Context &ctx, const double *r, double *p, const double *dinv, int64_t n,
                           void *status, int rz_cur, int rr_new, int rz_new, unsigned it,
                           unsigned max_steps, double tol);

  // cg.cu
  int cg_solve(Operator &A, Vector &x, Vector &b, int precondition, Vector *pvec,
               gdm_reduction_control &ctl);

  // comm.cpp
  void comm_unique_id(void *id128);
  void comm_init(Context &ctx, const void *id128, int rank, int n_ranks);
  void comm_allreduce_sum(Context &ctx, double *d_buf, int count, bool max_op = false);
  struct HaloPlan // stored-plane indices (local) of one ghost import
  {
    int prev = -1, next = -1; // neighbouring non-empty ranks (-1: none)
    int send_lo_plane = 0, send_lo_count = 0, recv_lo_plane = 0, recv_lo_count = 0;
    int send_hi_plane = 0, send_hi_count = 0, recv_hi_plane = 0, recv_hi_count = 0;
  };
  HaloPlan halo_plan(const Layout &L);
  void comm_halo_exchange(Context &ctx, const Layout &L, double *v, cudaStream_t stream = nullptr); // nullptr: ctx.stream
  int  comm_plane_owner(const Layout &L, int plane);
  void comm_send(Context &ctx, const double *buf, int64_t count, int peer, cudaStream_t stream);
  void comm_recv(Context &ctx, double *buf, int64_t count, int peer, cudaStream_t stream);
  void comm_destroy(Context &ctx);

  // vector helpers
  void vector_update_ghosts(Vector &v);
  inline int64_t round_up(int64_t a, int64_t b)
  {
    return (a + b - 1) / b * b;
  }
} // namespace gdm

struct gdm_context_s
{
  gdm::Context impl;
};
struct gdm_system_s
{
  gdm::System impl;
};
struct gdm_constraints_s
{
  gdm::Constraints impl;
};
struct gdm_vector_s
{
  gdm::Vector impl;
};
struct gdm_operator_s
{
  gdm::Operator impl;
};
