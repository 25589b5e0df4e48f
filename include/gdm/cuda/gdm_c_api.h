/* gdm_c_api.h -- C ABI of the B200-native GDM hot path (libgdm_b200.so).
 *
 * What this boundary replaces.  The reference has no FFI layer: its hot path
 * sits behind deal.II's operator/vector concept and a handful of GDM entry
 * points.  Each group of functions below names the reference interface it
 * stands in for (paths relative to the reference repository root):
 *
 *   gdm_system_*        GDM::System<dim>                 include/gdm/system.h:339-827
 *   gdm_polynomials_1d  GDM::generate_polynomials_1D     include/gdm/fe.h:55-336
 *   gdm_constraints_*   dealii::AffineConstraints as filled by
 *                       System::make_zero_boundary_constraints / make_periodicity_constraints
 *                                                        include/gdm/system.h:427-508
 *   gdm_operator_create(MASS)       GDM::MatrixCreator::create_mass_matrix   include/gdm/matrix_creator.h:9-62
 *   gdm_operator_lumped_mass_inverse  ...::create_lumped_mass_matrix          include/gdm/matrix_creator.h:64-117
 *   gdm_operator_create(STIFFNESS)  assembly loop        tests/poisson_02_gdm.cc:160-206
 *   gdm_operator_create(ADVECTION)  residual loop        prototypes/advection_01_gdm.cc:164-206,
 *                                                        applications/advection/include/gdm/advection/stiffness.h:373-418
 *   gdm_operator_vmult[_add]        SparseMatrix::vmult at every solver.solve site
 *                                                        tests/poisson_02_gdm.cc:215, tests/mass_01_gdm.cc:131, ...
 *   gdm_operator_attach_csr         cut-cell / ghost-penalty rows
 *                                                        applications/wave/include/gdm/wave/stiffness.h:589-799
 *   gdm_vector_*        LinearAlgebra::distributed::Vector<double> (deal.II)
 *   gdm_solver_cg       SolverCG + ReductionControl      tests/poisson_01_gdm.cc:164-170
 *   gdm_rk_*            TimeStepping::ExplicitRungeKutta prototypes/advection_01_gdm.cc:259-281
 *   gdm_interpolate / gdm_integrate_difference
 *                       GDM::VectorTools                 include/gdm/vector_tools.h:11-86
 *
 * Conventions: C linkage, opaque handles, plain pointers and sizes.  Every
 * function returns 0 on success or a gdm_status; the message is available from
 * gdm_last_error() (thread local).  Nothing throws across the ABI.  All device
 * work is enqueued on the context's stream (caller supplied cudaStream_t or the
 * legacy default stream); functions that return host data synchronise that
 * stream.  One context per GPU; handles are not thread safe, distinct contexts
 * are.  There is no CPU fallback: if no CUDA device is usable the compute
 * entry points fail with GDM_ERR_CUDA.
 */
#ifndef GDM_C_API_H
#define GDM_C_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDM_API_VERSION 1

typedef enum gdm_status {
  GDM_OK                 = 0,
  GDM_ERR_INVALID        = 1, /* bad argument (deal.II: ExcMessage / ExcDimensionMismatch) */
  GDM_ERR_NOT_IMPLEMENTED= 2, /* deal.II: ExcNotImplemented */
  GDM_ERR_CUDA           = 3, /* CUDA runtime / driver failure */
  GDM_ERR_NO_CONVERGENCE = 4, /* deal.II: SolverControl::NoConvergence */
  GDM_ERR_COMM           = 5, /* NCCL failure */
  GDM_ERR_INTERNAL       = 6
} gdm_status;

typedef struct gdm_context_s     *gdm_context_t;
typedef struct gdm_system_s      *gdm_system_t;
typedef struct gdm_constraints_s *gdm_constraints_t;
typedef struct gdm_operator_s    *gdm_operator_t;
/* scalar function of a point (dealii::Function<dim>::value): point[dim], component */
typedef double (*gdm_function_fn)(const double *point, int component, void *user);
typedef struct gdm_vector_s      *gdm_vector_t;
typedef struct gdm_rk_s          *gdm_rk_t;

const char *gdm_last_error(void);
int         gdm_api_version(void);

/* ---------------------------------------------------------------- context */
/* device: CUDA ordinal; stream: cudaStream_t (may be NULL = default stream). */
int gdm_context_create(int device, void *stream, gdm_context_t *ctx);
int gdm_context_destroy(gdm_context_t ctx);
int gdm_context_set_stream(gdm_context_t ctx, void *stream);
int gdm_context_synchronize(gdm_context_t ctx);
/* Number of kernels this library launched on the context since creation. */
int gdm_context_launch_count(gdm_context_t ctx, uint64_t *count);
/* Multi-GPU (one process per GPU).  The unique id is 128 bytes (ncclUniqueId);
 * rank 0 creates it, the caller broadcasts it (e.g. torch.distributed), every
 * rank calls gdm_context_comm_init.  Collectives used: ncclSend/ncclRecv for the
 * ghost planes, ncclAllReduce for CG scalars. */
int gdm_comm_unique_id(void *id128);
int gdm_context_comm_init(gdm_context_t ctx, const void *id128, int rank, int n_ranks);

/* ----------------------------------------------------------------- basis */
/* Monomial coefficients of all p variants, lowest power first:
 * coeffs[(v*(p+1) + k)*(p+1) + power], v = variant, k = basis function. */
int gdm_polynomials_1d(int fe_degree, double *coeffs);

/* ---------------------------------------------------------------- system */
typedef struct gdm_system_desc {
  int      dim;               /* 1, 2, 3 */
  int      fe_degree;         /* odd: 1,3,5,7,9 */
  int      n_components;      /* interleaved: dof = node*n_components + comp */
  uint32_t n_subdivisions[3]; /* cells per direction (>= fe_degree) */
  double   lo[3], hi[3];      /* subdivided_hyper_rectangle(p1, p2) */
  int      rank, n_ranks;     /* slab partition along the last direction (system.h:720-757) */
  int      add_ghost_layer;   /* widens the ghost zone by one plane (flux sparsity) */
} gdm_system_desc;

int      gdm_system_create(gdm_context_t ctx, const gdm_system_desc *desc, gdm_system_t *sys);
int      gdm_system_destroy(gdm_system_t sys);
uint64_t gdm_system_n_dofs(gdm_system_t sys);             /* global, all components */
uint64_t gdm_system_n_cells(gdm_system_t sys);
int      gdm_system_locally_owned_range(gdm_system_t sys, uint64_t *begin, uint64_t *end);
int      gdm_system_dofs_per_cell(gdm_system_t sys);
/* system.h:195-246; cell = lexicographic cell index; out has dofs_per_cell entries,
 * component-major (FESystem numbering), lexicographic inside a component. */
int      gdm_system_get_dof_indices(gdm_system_t sys, uint64_t cell, uint64_t *out);
/* system.h:404-424 */
int      gdm_system_active_fe_index(gdm_system_t sys, uint64_t cell, uint32_t *fe_index);
/* One row of System::create_sparsity_pattern (flux = 0, system.h:586-599) or create_flux_sparsity_pattern (flux != 0,
 * system.h:602-630): the global DoF indices the row couples to, ascending.  The pattern is never stored (343 entries per
 * row at p = 3 in 3D): rows are generated on demand from the per-direction window rules.  cols may be NULL to query the
 * length only; constrained rows/columns are kept (AffineConstraints::add_entries_local_to_global default); periodic
 * redirections are not included. */
int gdm_system_sparsity_row(gdm_system_t sys, int flux, uint64_t row, uint64_t *cols, uint64_t cap, uint64_t *n_cols);

/* Physical 1D band matrix of direction d without constraints:
 * kind 0 = mass (h*M1), 1 = stiffness (K1/h), 2 = convection (C1, row = test fn).
 * band[(row*(2p+1)) + tap], column = row + tap - p, rows 0..N_d. */
int      gdm_system_matrix_1d(gdm_system_t sys, int d, int kind, double *band);
/* Device storage geometry of vectors of this system (doubles). */
typedef struct gdm_layout_info {
  uint64_t pitch, plane, size;   /* row pitch, plane stride, total storage */
  uint64_t owned_offset, owned_size; /* contiguous owned block inside the storage */
  uint32_t local_nodes[3];       /* stored nodes per direction (incl. ghost planes) */
  uint32_t owned_begin, owned_end; /* owned node range in the partitioned direction (global) */
  uint32_t stored_begin, stored_end;
} gdm_layout_info;
int      gdm_system_layout(gdm_system_t sys, gdm_layout_info *info);
/* The ghost import of update_ghost_values as plane ranges of the local storage (units of the last
 * direction): plan10 = {prev_rank, next_rank, send_lo_plane, send_lo_count, recv_lo_plane, recv_lo_count,
 * send_hi_plane, send_hi_count, recv_hi_plane, recv_hi_count}; ranks are -1 where there is no neighbour. */
int      gdm_system_halo_plan(gdm_system_t sys, int32_t *plan10);
/* Work partition of the persistent fused kernel (kron3d_pers.cu; host logic, no device needed): the input planes
 * [k0, k1) of a tiles_x x tiles_y tile grid are split into shares; share w runs jobs [job_ptr[w], job_ptr[w+1]) of
 * jobs6 = {tile x, tile y, k_begin, k_end, seam_lo, seam_hi} per job.  A job with seam_lo >= 0 hands its first 2p
 * partial output planes to the job below it through scratch slot seam_lo; seam_hi is the slot a job reads when it
 * flushes.  No job is shorter than min_len planes unless it is a whole column; a job's upper neighbour lies in a share
 * with a larger index (the kernel hands shares out in descending order: deadlock-free seams).
 * mode 0: tile-major sweep over at most `slots` shares; 1: equal-cost chunks cut at the same planes + sweep of the rest,
 * weighted by `weights` (cost of a plane per tile, per mille; NULL = equal), at most `slots` shares; 2: guided levels
 * (one share per tile and level, levels shrink towards the bottom, more shares than slots: self-scheduling).
 * Replaces the per-rank slab loop of the reference's cell iteration (system.h:703-761) inside one GPU. */
int      gdm_pers_partition(int tiles_x, int tiles_y, int k0, int k1, int slots, int min_len, int mode,
                            const int32_t *weights, int32_t *job_ptr, int32_t cap_ptr, int32_t *jobs6,
                            int32_t cap_jobs, int32_t *n_shares, int32_t *n_jobs);

/* ----------------------------------------------------------- constraints */
int gdm_constraints_create(gdm_system_t sys, gdm_constraints_t *c);
int gdm_constraints_destroy(gdm_constraints_t c);
/* surface = 2*d + side, or -1 for all 2*dim faces (system.h:466-508) */
int gdm_constraints_make_zero_boundary(gdm_constraints_t c, int surface);
/* node N_d == node 0 in direction d, weight 1 (system.h:427-463) */
int gdm_constraints_make_periodicity(gdm_constraints_t c, int d);
/* System::interpolate_boundary_values (system.h:511-547): constrains every boundary node (boundary id 0) to f there;
 * DoFs that are already constrained keep their constraint.  Call before gdm_constraints_close. */
int gdm_constraints_interpolate_boundary_values(gdm_constraints_t c, int boundary_id, gdm_function_fn f, void *user);
int gdm_constraints_close(gdm_constraints_t c);
/* Right-hand side part of AffineConstraints::distribute_local_to_global with inhomogeneous constraints
 * (tests/poisson_02_gdm.cc:201): free rows b_i -= sum_j A_ij g_j, constrained rows b_j = diag_j g_j.  No-op for
 * homogeneous constraints.  After the solve, gdm_constraints_distribute writes g into the constrained DoFs. */
int gdm_constraints_condense_rhs(gdm_constraints_t c, gdm_operator_t op, gdm_vector_t rhs);
uint64_t gdm_constraints_n_constraints(gdm_constraints_t c);
int gdm_constraints_is_constrained(gdm_constraints_t c, uint64_t dof);
/* AffineConstraints::distribute / set_zero on a device vector */
int gdm_constraints_distribute(gdm_constraints_t c, gdm_vector_t v);
int gdm_constraints_set_zero(gdm_constraints_t c, gdm_vector_t v);

/* --------------------------------------------------------------- vectors */
int gdm_vector_create(gdm_system_t sys, gdm_vector_t *v);          /* zero initialised */
int gdm_vector_destroy(gdm_vector_t v);
/* host buffers hold the locally owned DoFs, compact, lexicographic (x fastest) */
int gdm_vector_upload(gdm_vector_t v, const double *host);
int gdm_vector_download(gdm_vector_t v, double *host);
void *gdm_vector_device_ptr(gdm_vector_t v);                       /* start of storage */
int gdm_vector_set(gdm_vector_t v, double value);                  /* owned entries (pads stay 0) */
int gdm_vector_copy(gdm_vector_t dst, gdm_vector_t src);           /* dst = src */
int gdm_vector_scale(gdm_vector_t v, double a);                    /* v *= a */
int gdm_vector_add(gdm_vector_t v, double a, gdm_vector_t x);      /* v += a x          (Vector::add) */
int gdm_vector_sadd(gdm_vector_t v, double s, double a, gdm_vector_t x); /* v = s v + a x (Vector::sadd) */
int gdm_vector_scale_by(gdm_vector_t v, gdm_vector_t d);           /* v *= d entrywise (Vector::scale) */
int gdm_vector_dot(gdm_vector_t a, gdm_vector_t b, double *result); /* all ranks; synchronises */
int gdm_vector_l2_norm(gdm_vector_t a, double *result);
int gdm_vector_linfty_norm(gdm_vector_t a, double *result);
int gdm_vector_update_ghost_values(gdm_vector_t v);                /* import p ghost planes per neighbour */

/* -------------------------------------------------------------- operators */
typedef enum gdm_operator_kind {
  GDM_OP_MASS        = 0, /* (phi_i, phi_j) */
  GDM_OP_STIFFNESS   = 1, /* (grad phi_i, grad phi_j) */
  GDM_OP_ADVECTION   = 2, /* (phi_i, b . grad phi_j)      prototypes/advection_01_gdm.cc:164-206 (times scale=-1) */
  GDM_OP_ADVECTION_T = 3  /* (b . grad phi_i, phi_j)      advection/stiffness.h:373-418 (alpha = 0 form) */
} gdm_operator_kind;

typedef enum gdm_constrained_diagonal {
  GDM_DIAG_ZERO      = 0, /* constrained rows give 0: vector assembly (residual) semantics */
  GDM_DIAG_ASSEMBLED = 1  /* sum_cells |cell_matrix(i,i)|: AffineConstraints::distribute_local_to_global on a matrix */
} gdm_constrained_diagonal;

enum { GDM_KERNEL_AUTO = 0, GDM_KERNEL_GENERIC = 1, GDM_KERNEL_FUSED = 2 };

typedef struct gdm_operator_desc {
  int    kind;                 /* gdm_operator_kind */
  double scale;                /* y = scale * A x */
  double b[3];                 /* advection velocity (constant) */
  int    constrained_diagonal; /* gdm_constrained_diagonal */
  int    kernel;               /* GDM_KERNEL_* : AUTO picks the fused sm_100a kernel when it covers the case */
} gdm_operator_desc;

int gdm_operator_create(gdm_system_t sys, gdm_constraints_t c /* may be NULL */,
                        const gdm_operator_desc *desc, gdm_operator_t *op);
/* write_matrix_to_file of the reference's eigenvalue tool (applications/wave/wave-ev.cc:93-127): the assembled operator
 * (kind / scale / constrained_diagonal of desc, constraints c or NULL) as "row column value" text lines or binary
 * (uint32, uint32, double) records, in the iteration order of a deal.II SparseMatrix (per row: diagonal first, then
 * ascending columns; pattern of create_sparsity_pattern with constrained rows and columns kept).  Host only (also on a
 * description-only context), one rank, no periodic directions.  n_entries may be NULL. */
int gdm_system_write_matrix(gdm_system_t sys, gdm_constraints_t c, const gdm_operator_desc *desc, const char *file_name,
                            int write_binary_file, uint64_t *n_entries);

int gdm_operator_destroy(gdm_operator_t op);
/* Irregular rows (cut cells, ghost penalty, Nitsche): CSR rows that REPLACE the
 * tensor-product result in those rows.  row_ids/col are global DoF indices;
 * rows must be locally owned, columns within the ghost zone. */
int gdm_operator_attach_csr(gdm_operator_t op, uint64_t n_rows, const uint64_t *row_ids,
                            const uint64_t *rowptr, const uint64_t *col, const double *val);
int gdm_operator_vmult(gdm_operator_t op, gdm_vector_t dst, gdm_vector_t src);     /* imports ghosts of src */
int gdm_operator_vmult_add(gdm_operator_t op, gdm_vector_t dst, gdm_vector_t src);
/* SparseMatrix::Tvmult: dst = A^T src.  Equal to vmult for mass and stiffness; for the advection kinds the transposed
 * operator is created on first use (same velocity, scale, constraints). */
int gdm_operator_tvmult(gdm_operator_t op, gdm_vector_t dst, gdm_vector_t src);
/* Kronecker-direct inverse of a GDM_OP_MASS operator: dst = M^-1 src by banded line solves per direction (Cholesky of
 * the 1D mass matrices; periodic directions through a p-node border), constrained rows: src_i / M_ii.  Replaces the
 * CG / ILU / AMG mass solves of every Runge-Kutta stage (applications/advection/include/gdm/advection/problem.h:236-267,
 * applications/wave/include/gdm/wave/problem.h:471-502, prototypes/advection_01_gdm.cc:208-216) where the grid is
 * Cartesian.  One rank, no irregular rows (GDM_ERR_NOT_IMPLEMENTED otherwise); dst may alias src. */
int gdm_operator_mass_inverse(gdm_operator_t op, gdm_vector_t dst, gdm_vector_t src);
/* Host-buffer form of vmult (the call a deal.II user makes with host vectors):
 * copies src to the device, applies, copies dst back; synchronises. */
/* dst = A src and *src_dot_dst = <src, dst> over all ranks (the q = A p, p.q pair of SolverCG): on the fused path the
 * dot product is accumulated in the store epilogue of the tile kernel, so p and q are not read again. */
int gdm_operator_vmult_dot(gdm_operator_t op, gdm_vector_t dst, gdm_vector_t src, double *src_dot_dst);
int gdm_operator_vmult_host(gdm_operator_t op, double *dst_host, const double *src_host);
int gdm_operator_diagonal(gdm_operator_t op, gdm_vector_t diag);   /* matrix diagonal (Jacobi) */
/* 1 / (row sums of the mass operator): matrix_creator.h:64-117; op must be MASS */
int gdm_operator_lumped_mass_inverse(gdm_operator_t op, gdm_vector_t inv);
int gdm_operator_kernel_used(gdm_operator_t op);                   /* GDM_KERNEL_GENERIC or _FUSED */
uint64_t gdm_operator_m(gdm_operator_t op);

/* ----------------------------------------------------------------- solver */
typedef struct gdm_reduction_control {
  uint32_t max_steps;  /* ReductionControl(n, tol, reduce) */
  double   tolerance;
  double   reduce;
  /* results */
  uint32_t last_step;
  double   last_value;
  double   initial_value;
} gdm_reduction_control;

typedef enum gdm_precondition {
  GDM_PRECONDITION_IDENTITY = 0,
  GDM_PRECONDITION_JACOBI   = 1, /* PreconditionJacobi, relaxation 1 */
  GDM_PRECONDITION_DIAGONAL = 2  /* DiagonalMatrix: z = d .* r with caller supplied d */
} gdm_precondition;

/* deal.II SolverCG::solve(A, x, b, P).  Returns GDM_ERR_NO_CONVERGENCE where
 * deal.II throws SolverControl::NoConvergence (control still filled in). */
int gdm_solver_cg(gdm_operator_t A, gdm_vector_t x, gdm_vector_t b, int precondition,
                  gdm_vector_t precondition_vector /* DIAGONAL only */, gdm_reduction_control *control);

/* ------------------------------------------------------------ time stepping */
typedef enum gdm_rk_method { GDM_RK_FORWARD_EULER = 0, GDM_RK_THIRD_ORDER = 1, GDM_RK_CLASSIC_FOURTH_ORDER = 2 } gdm_rk_method;
/* f(t, y[0..n_blocks), out[0..n_blocks)) enqueues out = f(t, y) on the context stream. */
typedef int (*gdm_rk_rhs_fn)(double t, const gdm_vector_t *y, gdm_vector_t *out, void *user);
int gdm_rk_create(gdm_system_t sys, int method, int n_blocks, gdm_rk_t *rk);
int gdm_rk_destroy(gdm_rk_t rk);
/* ExplicitRungeKutta::evolve_one_time_step; returns t + dt in *t_new */
int gdm_rk_evolve_one_time_step(gdm_rk_t rk, gdm_rk_rhs_fn f, void *user, double t, double dt,
                                gdm_vector_t *y, double *t_new);

/* -------------------------------------------------------------- vector tools */
int gdm_interpolate(gdm_system_t sys, gdm_function_fn f, void *user, gdm_vector_t v);
/* cellwise L2 error (length n_cells of the locally owned cells' global numbering;
 * cells of other ranks are left 0) and its global l2 sum over all ranks */
int gdm_integrate_difference(gdm_system_t sys, gdm_vector_t v, gdm_function_fn exact, void *user,
                             double *cellwise /* may be NULL */, double *global_l2);

/* VTU file (ASCII UnstructuredGrid; point data on the grid cells, one array per component) of a nodal field: the stand-in
 * for GDM::DataOut (include/gdm/data_out.h; prototypes/advection_01_gdm.cc:248-255) for visual checks.  Host only;
 * values = all DoFs in the global numbering (one rank). */
int gdm_system_write_vtu(gdm_system_t sys, const double *values, const char *label, const char *file_name);

/* ------------------------------------------------------- cut-cell set-up (host) */
/* The step before the hot path in a CutFEM run: level-set classification, cut quadrature and the
 * assembly of the rows that differ from the plain stiffness operator, ready for gdm_operator_attach_csr.  Host only,
 * no context needed, scalar field, no constraints; every rank assembles its own row range (row_begin, row_end).
 *   classification   NonMatching::MeshClassifier for a Q1 level set (prototypes/cut_poisson_01_gdm.cc:105-121)
 *   quadrature       NonMatching::FEValues with QGauss<1>(p+1) (prototypes/cut_poisson_01_gdm.cc:176-190)
 *   assembly         prototypes/cut_poisson_01_gdm.cc:196-329 (volume + Nitsche + ghost penalty, zero diagonal -> 1)
 *   error            prototypes/cut_poisson_01_gdm.cc:349-398 */
typedef struct gdm_cut_desc {
  int      dim;               /* 1, 2, 3 */
  int      fe_degree;         /* odd */
  uint32_t n_subdivisions[3];
  double   lo[3], hi[3];
  int      ghost_penalty;     /* face_has_ghost_penalty terms on / off (cut_poisson_01_gdm.cc:123-146) */
  int      gp_h_power;        /* 1: the prototype's scaling; 3: the wave application's matrix (wave/stiffness.h:760-765) */
  double   ghost_parameter;   /* 0.5 */
  double   nitsche_parameter; /* 5 (p+1) p */
  double   rhs_value;         /* constant right-hand side f (4) */
  double   boundary_value;    /* constant Dirichlet value g on the surface (1) */
  int      kind;              /* 0: stiffness + Nitsche (the Poisson matrix / the linear part of the residual
                                 wave/stiffness.h:42-407); 1: the cut mass matrix wave/mass.h:47-249 (no surface terms;
                                 use ghost_parameter = gamma_M, gp_h_power = 3) */
  double   outside_diagonal;  /* diagonal of the rows no active cell touches: 1 for a matrix that is solved with
                                 (cut_poisson_01_gdm.cc:324-329, wave/mass.h:246-248), 0 for the matrix-free residual */
  int      no_surface_terms;      /* 1: no Nitsche terms on the cut surface (function_interface_dbc unset: the two-domain
                                     runs couple there instead, wave/stiffness.h:441-574) */
  int      domain_boundary_terms; /* 1: Nitsche terms on the box boundary for the part of it inside the domain
                                     (function_domain_dbc, wave/stiffness.h:262-340) */
  int      level_set_degree;      /* 0 or 1: Q1 level set, nodal values at the grid nodes.  q > 1 (dim = 2 only): FE_Q(q)
                                     level set as in the 2D presets of applications/wave (wave-app.cc:277); `level_set`
                                     then holds the values on the Gauss-Lobatto refined grid, (q N_e + 1) points per
                                     direction, x fastest (point q c + k of a direction = node k of FE_Q(q) in cell c) */
  uint64_t row_begin, row_end; /* rows (global DoFs) to assemble: the locally owned range of a rank
                                  (gdm_system_locally_owned_range); 0, 0 = all.  With a range the right-hand side is
                                  complete in that range only. */
} gdm_cut_desc;
typedef struct gdm_cut_s *gdm_cut_t;
/* level_set: nodal values of the Q1 level set at the grid nodes, DoF order (x fastest); negative = inside.  A level set
 * that vanishes exactly on a whole grid plane is degenerate (the cells on its positive side classify as intersected with
 * an empty inside part, as with deal.II's classifier, and the surface rule on that plane is empty): shift it. */
/* where the level set is to be sampled for this description: points[n_points * dim] (x fastest; the grid nodes for a Q1
 * level set, the support points of FE_Q(level_set_degree) of every cell otherwise).  points may be NULL (count only). */
int gdm_cut_level_set_points(const gdm_cut_desc *desc, uint64_t *n_points, double *points);
int gdm_cut_poisson_create(const gdm_cut_desc *desc, const double *level_set, gdm_cut_t *cut);
int gdm_cut_destroy(gdm_cut_t cut);
/* n_rows = rows to attach (band rows around the surface + identity rows of DoFs no active cell touches);
 * n_cells_by_location[3] = inside, outside, intersected.  Any pointer may be NULL. */
int gdm_cut_sizes(gdm_cut_t cut, uint64_t *n_rows, uint64_t *nnz, uint64_t *n_identity_rows,
                  uint64_t *n_cells_by_location);
/* the arguments of gdm_operator_attach_csr: row_ids[n_rows] ascending, rowptr[n_rows+1], col/val[nnz] (columns ascending) */
int gdm_cut_rows(gdm_cut_t cut, uint64_t *row_ids, uint64_t *rowptr, uint64_t *col, double *val);
int gdm_cut_rhs(gdm_cut_t cut, double *rhs /* n_dofs */);
/* out[n_dofs] = (v, f) over the inside part + <gamma_D / h v - dv/dn, g> on the surface, for functions of the point: the
 * data-dependent part of the residual (wave/stiffness.h:186-260; time enters through `user`).  f or g may be NULL. */
int gdm_cut_load_vector(gdm_cut_t cut, gdm_function_fn f, void *f_user, gdm_function_fn g, void *g_user, double *out);
/* out[n_dofs] = <gamma_D / h v - dv/dn, g> on the box boundary (the load of domain_boundary_terms) */
int gdm_cut_boundary_load_vector(gdm_cut_t cut, gdm_function_fn g, void *user, double *out);
/* Interface coupling of the two-domain runs (wave/stiffness.h:441-574) as sparse matrices over the cut surface, n = normal
 * of the level set:  which = 0: P_ij = <n . grad phi_i, phi_j>, 1: P^T, 2: Q_ij = <phi_i, phi_j>.  With [u] = u0 - u1 and
 * tau = gamma_D / 2 the residuals get  r0 -= -1/2 P [u] - 1/2 P^T (u0 + u1) + tau / h Q [u],
 * r1 -= -1/2 P [u] + 1/2 P^T (u0 + u1) - tau / h Q [u].  Always returns the sizes; fills the arrays (CSR rows in the
 * form of gdm_operator_attach_csr, to be laid over an operator with scale 0) when they are given and large enough. */
int gdm_cut_coupling_rows(gdm_cut_t cut, int which, uint64_t capacity_rows, uint64_t capacity_nnz, uint64_t *n_rows,
                          uint64_t *nnz, uint64_t *row_ids, uint64_t *rowptr, uint64_t *col, double *val);
int gdm_cut_locations(gdm_cut_t cut, uint8_t *location /* n_cells: 0 inside, 1 outside, 2 intersected */);
/* sqrt( sum over non-outside cells of the integral over the inside part of (u_h - exact)^2 ); u = all DoFs */
int gdm_cut_l2_error_inside(gdm_cut_t cut, const double *u, gdm_function_fn exact, void *user, double *error);
/* norms[3] = L2, L1, Linf (maximum over the quadrature points) of u_h - exact over the inside part: the three columns
 * applications/wave prints after every step (wave/problem.h:531-615) */
int gdm_cut_error_norms_inside(gdm_cut_t cut, const double *u, gdm_function_fn exact, void *user, double *norms);
/* The quadrature generator on its own: unit cell, vertex_values[2^dim] (bit e of the index = upper end in direction e).
 * Fills at most `capacity` points per rule (points [q*dim + e]) and returns the full counts; arrays may be NULL. */
int gdm_cut_quadrature(int dim, const double *vertex_values, int n_gauss, uint64_t capacity, uint64_t *n_inside,
                       double *inside_points, double *inside_weights, uint64_t *n_surface, double *surface_points,
                       double *surface_weights, double *surface_normals);

#ifdef __cplusplus
}
#endif
#endif /* GDM_C_API_H */
