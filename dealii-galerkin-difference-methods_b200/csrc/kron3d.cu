// Fused matrix-free GDM operator apply for dim == 3: orchestration around the persistent tile kernel (kron3d_pers.cu).
//
//     y = scale * ( B_x (x) A_y (x) A_z  +  A_x (x) B_y (x) A_z  +  A_x (x) A_y (x) B_z ) x      (stiffness / advection)
//     y = scale * ( A_x (x) A_y (x) A_z ) x                                                      (mass)
//
// Replaces the assembled SparseMatrix::vmult of the reference (343 nnz/row at p=3: tests/poisson_02_gdm.cc:215,
// applications/wave/include/gdm/wave/problem.h:486-488).  This file decides what is launched where:
//   * one GPU: the constrained-row (Dirichlet face) kernel runs beside the tile kernel on the second stream;
//   * several GPUs (slab partition, include/gdm/system.h:720-757): the ghost import (NCCL send/recv of p contiguous
//     planes per neighbour) runs first, then one launch over all owned planes; an overlapped schedule (slab faces
//     behind the exchange, interior planes at once) is kept behind GDM_FUSED_OVERLAP=1, it measured slower;
//   * periodic directions: C^T A C x = fold(A(dup x)): the input is patched in place (node N := node 0),
//     the tile kernel applies the plain one-sided operator, rows N are folded into rows 0 and the input is restored;
//   * fused dot product <src, A src> (CG: p . A p): per-CTA partial sums of the tile launches + the face kernel, summed
//     in a fixed order (bitwise reproducible).
// The round-1 tile kernels (v3 ... v7: chunked launches with z ramps, profiles/r1) were removed in round 2: the
// persistent kernel supersedes them for every configuration they covered.
#include <algorithm>
#include <cstdlib>

#include "gdm_internal.h"

namespace gdm
{
  namespace
  {
    struct FusedPlan
    {
      void   *pers = nullptr; // PersPlan (kron3d_pers.cu)
      int     cz0 = 0, cz1 = 0;
      // fused dot product: per-CTA partial sums of the launches of one apply (tile kernels, then the face kernel)
      double *d_dot      = nullptr;
      size_t  dot_cap    = 0;
      int     dot_cursor = -1; // >= 0 while an apply with a fused dot is being enqueued
      const double *dot_src = nullptr;
      ~FusedPlan()
      {
        if (pers)
          pers_plan_destroy(pers);
        cudaFree(d_dot);
      }
    };

    // one launch of the persistent kernel over output planes [z0, z1) (local indices)
    void launch_tiles(Operator &op, FusedPlan &plan, double *dst, const double *src, bool accumulate, int z0, int z1,
                      cudaStream_t stream, int slots_limit)
    {
      double *dp = nullptr;
      if (plan.dot_cursor >= 0)
        {
          GDM_REQUIRE(!accumulate && (size_t)(plan.dot_cursor + pers_max_partials(op, plan.pers)) <= plan.dot_cap, GDM_ERR_INTERNAL,
                      "fused dot: partial buffer too small");
          dp = plan.d_dot + plan.dot_cursor;
        }
      const int grid = pers_launch(op, plan.pers, dst, src, accumulate, z0, z1, stream, dp ? plan.dot_src : nullptr, dp, slots_limit);
      if (dp)
        plan.dot_cursor += grid;
    }
  } // namespace

  bool fused_supported(const Operator &op)
  {
    const char *env = std::getenv("GDM_DISABLE_FUSED");
    if (env && env[0] == '1')
      return false;
    return pers_supported(op);
  }

  void fused_plan_create(Operator &op)
  {
    auto *plan = new FusedPlan;
    op.fused   = plan;
    plan->pers = pers_plan_create(op);
    GDM_REQUIRE(plan->pers != nullptr, GDM_ERR_INTERNAL, "fused kernel: plan creation failed for a supported operator");
    pers_window(plan->pers, plan->cz0, plan->cz1);
    pers_tune(op, plan->pers);
  }

  void fused_plan_destroy(Operator &op)
  {
    delete static_cast<FusedPlan *>(op.fused);
    op.fused = nullptr;
  }

  bool fused_supports_dot(const Operator &op)
  {
    return op.fused != nullptr;
  }

  void fused_apply(Operator &op, double *dst, const double *src, bool accumulate, bool exchange_ghosts, int dot_slot)
  {
    FusedPlan    &plan = *static_cast<FusedPlan *>(op.fused);
    Context      &ctx  = *op.sys->ctx;
    const Layout &L    = op.sys->L;
    const int     P    = L.p;
    const bool    want_dot = dot_slot >= 0;
    // periodic directions: src is patched in place and restored
    const bool periodic = pers_has_periodic(plan.pers);
    if (periodic)
      {
        GDM_REQUIRE(!accumulate, GDM_ERR_INTERNAL, "fused periodic apply cannot accumulate (use a temporary)");
        pers_periodic_pre(op, plan.pers, const_cast<double *>(src), ctx.stream);
      }
    double *face_partials = nullptr; // where the face kernel puts its partial sums
    if (want_dot)
      {
        GDM_REQUIRE(!accumulate, GDM_ERR_INTERNAL, "fused dot product with accumulation");
        // partial sums: one per share of the (up to three) tile launches, then one per block of the face kernel
        const size_t need = (size_t)3 * pers_max_partials(op, plan.pers) + (size_t)constrained_rows_max_blocks(L) + 64;
        if (need > plan.dot_cap)
          {
            GDM_CUDA_CHECK(cudaDeviceSynchronize());
            cudaFree(plan.d_dot);
            plan.d_dot = nullptr;
            GDM_CUDA_CHECK(cudaMalloc(&plan.d_dot, need * sizeof(double)));
            plan.dot_cap = need;
          }
        plan.dot_src    = src;
        plan.dot_cursor = 0;
        // the face kernel may run concurrently with the tile kernels: give it the tail of the buffer
        face_partials = plan.d_dot + (plan.dot_cap - (size_t)constrained_rows_max_blocks(L));
      }
    int  face_blocks = 0;
    auto finish_dot  = [&]() {
      if (!want_dot)
        return;
      // compact: [tile partials | face partials] are summed in a fixed order into the slot
      const int n_tile = plan.dot_cursor;
      plan.dot_cursor  = -1;
      plan.dot_src     = nullptr;
      if (face_blocks > 0) // move the face partials behind the tile partials (device-side copy, same stream)
        GDM_CUDA_CHECK(cudaMemcpyAsync(plan.d_dot + n_tile, face_partials, (size_t)face_blocks * sizeof(double), cudaMemcpyDeviceToDevice,
                                       ctx.stream));
      blas_sum_partials(ctx, plan.d_dot, n_tile + face_blocks, dot_slot);
    };
    if (L.n_ranks > 1 && exchange_ghosts)
      {
        // overlap the ghost import (NCCL on the comm stream) with the planes that do not need it
        const int  lo = L.own0 - L.loc0, hi = L.own1 - L.loc0; // owned planes (local indices)
        // Default: ghost import first, then ONE launch over all planes (0.155 ms per apply on 2 B200 at 257^3 per GPU).
        // GDM_FUSED_OVERLAP=1 selects the overlapped schedule below (slab faces behind the exchange on the communication
        // stream, interior planes on the main stream).  With a persistent kernel that fills every CTA slot it is SLOWER
        // (0.30 ms, profiles/r2/session_v_2gpu.txt): the face launches redo 2p ramp planes for p output planes on a
        // handful of slots, and NCCL's copy kernels compete with the resident CTAs for an SM.
        const char *env_ov = std::getenv("GDM_FUSED_OVERLAP");
        const bool  thick  = (hi - lo) > 4 * P && (env_ov && env_ov[0] == '1');
        GDM_CUDA_CHECK(cudaEventRecord(ctx.ev_a, ctx.stream));
        GDM_CUDA_CHECK(cudaStreamWaitEvent(ctx.comm_stream, ctx.ev_a, 0));
        comm_halo_exchange(ctx, L, const_cast<double *>(src), ctx.comm_stream);
        if (thick)
          {
            // The interior launch leaves a share of the CTA slots free (the face planes' share of the work, a little more
            // because they start late by the latency of the exchange), so the face launches find room the moment the
            // ghost planes arrive instead of queueing behind the persistent interior CTAs.
            const int all_slots  = pers_max_grid(op, plan.pers);
            int       face_slots = std::max(4, (int)(1.15 * all_slots * (3.0 * P) / (double)((hi - lo) + 4 * P) + 0.5));
            if (const char *env = std::getenv("GDM_PERS_FACE_SLOTS"))
              face_slots = std::max(1, atoi(env));
            launch_tiles(op, plan, dst, src, accumulate, lo, lo + P, ctx.comm_stream, face_slots);
            launch_tiles(op, plan, dst, src, accumulate, hi - P, hi, ctx.comm_stream, face_slots);
            GDM_CUDA_CHECK(cudaEventRecord(ctx.ev_b, ctx.comm_stream));
            launch_tiles(op, plan, dst, src, accumulate, lo + P, hi - P, ctx.stream, std::max(1, all_slots - 2 * face_slots));
            GDM_CUDA_CHECK(cudaStreamWaitEvent(ctx.stream, ctx.ev_b, 0));
          }
        else
          {
            GDM_CUDA_CHECK(cudaEventRecord(ctx.ev_b, ctx.comm_stream));
            if (!periodic)
              {
                // the constrained rows only read owned values: their kernel runs while the ghost planes are in flight
                GDM_CUDA_CHECK(cudaStreamWaitEvent(ctx.face_stream, ctx.ev_a, 0));
                face_blocks = launch_constrained_rows(ctx, L, op, dst, src, accumulate, -1, -1, face_partials, ctx.face_stream);
                GDM_CUDA_CHECK(cudaEventRecord(ctx.ev_c, ctx.face_stream));
              }
            GDM_CUDA_CHECK(cudaStreamWaitEvent(ctx.stream, ctx.ev_b, 0));
            launch_tiles(op, plan, dst, src, accumulate, plan.cz0, plan.cz1, ctx.stream, 0);
            if (!periodic)
              {
                GDM_CUDA_CHECK(cudaStreamWaitEvent(ctx.stream, ctx.ev_c, 0));
                finish_dot();
                return;
              }
          }
      }
    else if (periodic)
      launch_tiles(op, plan, dst, src, accumulate, plan.cz0, plan.cz1, ctx.stream, 0);
    else
      {
        // Dirichlet faces (skipped by the tiles) and deal.II's constrained diagonal: disjoint outputs, so the small face
        // kernel runs beside the tile kernel.  The tile kernel is launched FIRST and the face kernel goes to a stream of
        // the lowest priority: the persistent CTAs take every slot at once and the face blocks run as the first CTAs
        // retire (in the tail of the launch).  Launched first on a high-priority stream, as in round 1, the face blocks
        // delayed a few persistent CTAs and with them the whole launch (0.128 vs 0.120 ms, profiles/r2/launches_bench.csv).
        const char *env_fo = std::getenv("GDM_FACE_ORDER");
        const bool  first  = env_fo && env_fo[0] == 'f'; // diagnostic: the round-1 order
        GDM_CUDA_CHECK(cudaEventRecord(ctx.ev_a, ctx.stream));
        cudaStream_t fs = first ? ctx.comm_stream : ctx.face_stream;
        if (!first)
          launch_tiles(op, plan, dst, src, accumulate, plan.cz0, plan.cz1, ctx.stream, 0);
        GDM_CUDA_CHECK(cudaStreamWaitEvent(fs, ctx.ev_a, 0));
        face_blocks = launch_constrained_rows(ctx, L, op, dst, src, accumulate, -1, -1, face_partials, fs);
        GDM_CUDA_CHECK(cudaEventRecord(ctx.ev_b, fs));
        if (first)
          launch_tiles(op, plan, dst, src, accumulate, plan.cz0, plan.cz1, ctx.stream, 0);
        GDM_CUDA_CHECK(cudaStreamWaitEvent(ctx.stream, ctx.ev_b, 0));
        finish_dot();
        return;
      }
    if (periodic)
      pers_periodic_post(op, plan.pers, dst, const_cast<double *>(src), ctx.stream);
    // Dirichlet faces (skipped by the tiles), duplicate nodes of periodic directions: deal.II's constrained diagonal
    face_blocks = launch_constrained_rows(ctx, L, op, dst, src, accumulate, -1, -1, face_partials);
    finish_dot();
  }

  void fused_apply_window(Operator &op, double *dst, const double *src, int z0, int z1)
  {
    FusedPlan    &plan = *static_cast<FusedPlan *>(op.fused);
    Context      &ctx  = *op.sys->ctx;
    const Layout &L    = op.sys->L;
    launch_tiles(op, plan, dst, src, false, z0, z1, ctx.stream, 0);
    launch_constrained_rows(ctx, L, op, dst, src, false, z0, z1);
  }
} // namespace gdm
