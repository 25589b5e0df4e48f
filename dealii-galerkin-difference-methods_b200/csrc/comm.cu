// Multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//
// Stands in for the MPI traffic the reference generates through deal.II
// (LinearAlgebra::distributed::Vector::update_ghost_values -> Partitioner Isend/Irecv,
// MPI_Allreduce inside SolverCG; call sites listed in SURVEY.md section 2b, e.g.
// tests/poisson_02_gdm.cc:225,236 and applications/wave/include/gdm/wave/stiffness.h:149).
// The slab partition is the reference's own (include/gdm/system.h:720-757): ghost planes are
// contiguous in memory, so the halo exchange needs no pack kernel -- ncclSend/ncclRecv straight
// from/to the vector storage.  NCCL is resolved with dlopen at comm-init time so that the
// single-GPU path has no NCCL dependency (and picks up the libnccl torch already loaded).
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include "gdm_internal.h"

namespace gdm
{
  namespace
  {
    typedef struct ncclComm *ncclComm_t;
    struct ncclUniqueId
    {
      char internal[128];
    };
    enum
    {
      ncclSuccess = 0
    };
    enum
    {
      ncclFloat64 = 8
    };
    enum
    {
      ncclSum = 0,
      ncclMax = 2
    };

    struct Nccl
    {
      void *lib = nullptr;
      int (*GetUniqueId)(ncclUniqueId *)                                                     = nullptr;
      int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int)                              = nullptr;
      int (*CommDestroy)(ncclComm_t)                                                         = nullptr;
      int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t)     = nullptr;
      int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t)                  = nullptr;
      int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t)                        = nullptr;
      int (*GroupStart)()                                                                    = nullptr;
      int (*GroupEnd)()                                                                      = nullptr;
      const char *(*GetErrorString)(int)                                                     = nullptr;
    };

    Nccl &nccl()
    {
      static Nccl n;
      if (n.lib)
        return n;
      const char *names[] = {"libnccl.so.2", "libnccl.so"};
      for (const char *nm : names)
        {
          n.lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
          if (n.lib)
            break;
        }
      for (const char *nm : names)
        {
          if (n.lib)
            break;
          n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        }
      if (!n.lib)
        throw Error(GDM_ERR_COMM, std::string("cannot load libnccl: ") + dlerror());
#define GDM_NCCL_SYM(field, name)                                          \
  n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.lib, name));       \
  if (!n.field)                                                            \
    throw Error(GDM_ERR_COMM, std::string("libnccl lacks symbol ") + name);
      GDM_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
      GDM_NCCL_SYM(CommInitRank, "ncclCommInitRank");
      GDM_NCCL_SYM(CommDestroy, "ncclCommDestroy");
      GDM_NCCL_SYM(AllReduce, "ncclAllReduce");
      GDM_NCCL_SYM(Send, "ncclSend");
      GDM_NCCL_SYM(Recv, "ncclRecv");
      GDM_NCCL_SYM(GroupStart, "ncclGroupStart");
      GDM_NCCL_SYM(GroupEnd, "ncclGroupEnd");
      GDM_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef GDM_NCCL_SYM
      return n;
    }

    void check(int rc, const char *what)
    {
      if (rc != ncclSuccess)
        throw Error(GDM_ERR_COMM, std::string(what) + ": " + nccl().GetErrorString(rc));
    }
  } // namespace

  struct Comm
  {
    ncclComm_t comm = nullptr;
    ~Comm()
    {
      if (comm)
        nccl().CommDestroy(comm);
    }
  };

  void comm_unique_id(void *id128)
  {
    ncclUniqueId id;
    check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(id128, &id, sizeof(id));
  }

  void comm_init(Context &ctx, const void *id128, int rank, int n_ranks)
  {
    GDM_REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, GDM_ERR_INVALID, "bad rank");
    GDM_CUDA_CHECK(cudaSetDevice(ctx.device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    std::unique_ptr<Comm> c(new Comm);
    check(nccl().CommInitRank(&c->comm, n_ranks, id, rank), "ncclCommInitRank");
    comm_destroy(ctx);
    ctx.comm    = c.release();
    ctx.rank    = rank;
    ctx.n_ranks = n_ranks;
  }

  void comm_destroy(Context &ctx)
  {
    delete ctx.comm;
    ctx.comm = nullptr;
  }

  void comm_allreduce_sum(Context &ctx, double *d_buf, int count, bool max_op)
  {
    GDM_REQUIRE(ctx.comm && ctx.comm->comm, GDM_ERR_COMM, "communicator not initialised");
    check(nccl().AllReduce(d_buf, d_buf, (size_t)count, ncclFloat64, max_op ? ncclMax : ncclSum, ctx.comm->comm,
                           ctx.stream),
          "ncclAllReduce");
  }

  namespace
  {
    void owned_range(const Layout &L, int rank, int &o0, int &o1)
    {
      const int n_last = L.nn[L.pdim];
      const int stride = (L.N[L.pdim] + L.n_ranks - 1) / L.n_ranks;
      o0               = std::min((rank == 0) ? 0 : (stride * rank + 1), n_last);
      o1               = std::min(stride * (rank + 1) + 1, n_last);
    }
  } // namespace

  // Which planes move in a ghost import (pure host logic; also exported for the CPU-side tests).
  HaloPlan halo_plan(const Layout &L)
  {
    HaloPlan h;
    if (L.n_ranks == 1 || L.own1 <= L.own0)
      return h;
    int o0, o1;
    for (int r = L.rank - 1; r >= 0 && h.prev < 0; --r)
      {
        owned_range(L, r, o0, o1);
        if (o1 > o0)
          h.prev = r;
      }
    for (int r = L.rank + 1; r < L.n_ranks && h.next < 0; ++r)
      {
        owned_range(L, r, o0, o1);
        if (o1 > o0)
          h.next = r;
      }
    const int own = L.own1 - L.own0;
    // what I receive
    h.recv_lo_plane = 0;
    h.recv_lo_count = (h.prev >= 0) ? L.own0 - L.loc0 : 0;
    h.recv_hi_plane = L.own1 - L.loc0;
    h.recv_hi_count = (h.next >= 0) ? L.loc1 - L.own1 : 0;
    // what the neighbours expect from me: prev's upper ghost zone, next's lower ghost zone
    h.send_lo_count = (h.prev >= 0) ? std::min(L.ghost, L.nn[L.pdim] - L.own0) : 0;
    h.send_lo_plane = L.own0 - L.loc0;
    h.send_hi_count = (h.next >= 0) ? std::min(L.ghost, L.own1) : 0;
    h.send_hi_plane = L.own0 - L.loc0 + own - h.send_hi_count;
    GDM_REQUIRE(h.send_lo_count <= own && h.send_hi_count <= own, GDM_ERR_NOT_IMPLEMENTED,
                "slab thinner than the ghost zone: use fewer ranks or a larger grid");
    return h;
  }

  // rank that owns node plane `plane` (global index) of the partitioned direction
  int comm_plane_owner(const Layout &L, int plane)
  {
    for (int r = 0; r < L.n_ranks; ++r)
      {
        int o0, o1;
        owned_range(L, r, o0, o1);
        if (plane >= o0 && plane < o1)
          return r;
      }
    return -1;
  }

  // point-to-point transfers of `count` doubles (periodic wrap of the partitioned direction)
  void comm_send(Context &ctx, const double *buf, int64_t count, int peer, cudaStream_t stream)
  {
    GDM_REQUIRE(ctx.comm && ctx.comm->comm, GDM_ERR_COMM, "communicator not initialised");
    check(nccl().Send(buf, (size_t)count, ncclFloat64, peer, ctx.comm->comm, stream), "ncclSend");
  }
  void comm_recv(Context &ctx, double *buf, int64_t count, int peer, cudaStream_t stream)
  {
    GDM_REQUIRE(ctx.comm && ctx.comm->comm, GDM_ERR_COMM, "communicator not initialised");
    check(nccl().Recv(buf, (size_t)count, ncclFloat64, peer, ctx.comm->comm, stream), "ncclRecv");
  }

  // Import L.ghost planes from each neighbouring slab into the ghost zones of v.
  void comm_halo_exchange(Context &ctx, const Layout &L, double *v, cudaStream_t stream)
  {
    if (stream == nullptr)
      stream = ctx.stream;
    const HaloPlan h = halo_plan(L);
    if (h.prev < 0 && h.next < 0)
      return;
    GDM_REQUIRE(ctx.comm && ctx.comm->comm, GDM_ERR_COMM, "communicator not initialised");
    const int64_t unit = L.stride[L.pdim];
    Nccl         &n    = nccl();
    check(n.GroupStart(), "ncclGroupStart");
    if (h.prev >= 0)
      {
        check(n.Send(v + h.send_lo_plane * unit, (size_t)h.send_lo_count * unit, ncclFloat64, h.prev, ctx.comm->comm, stream), "ncclSend");
        check(n.Recv(v + h.recv_lo_plane * unit, (size_t)h.recv_lo_count * unit, ncclFloat64, h.prev, ctx.comm->comm, stream), "ncclRecv");
      }
    if (h.next >= 0)
      {
        check(n.Send(v + h.send_hi_plane * unit, (size_t)h.send_hi_count * unit, ncclFloat64, h.next, ctx.comm->comm, stream), "ncclSend");
        check(n.Recv(v + h.recv_hi_plane * unit, (size_t)h.recv_hi_count * unit, ncclFloat64, h.next, ctx.comm->comm, stream), "ncclRecv");
      }
    check(n.GroupEnd(), "ncclGroupEnd");
  }
} // namespace gdm
