// Microbenchmark: how fast can an SM stage (TX+2P) x (TY+2P) x BZ FP64 tiles of a 257^3 (pitch 260) field into shared
// memory?  Compares TMA box loads (cp.async.bulk.tensor.3d) for several box shapes / stage counts / CTAs per SM with
// cp.async (LDGSTS, 16 B) staging of the same tiles.  No arithmetic: this is the "skeleton" of kron3d.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_box tma_box.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                              \
  do                                                                                       \
    {                                                                                      \
      cudaError_t e = (x);                                                                 \
      if (e != cudaSuccess)                                                                \
        {                                                                                  \
          printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);   \
          exit(1);                                                                         \
        }                                                                                  \
    }                                                                                      \
  while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Args
{
  int tiles_x, tiles_y, tx, ty, lz, nz, halo; // tile steps, z chunk length, planes, halo width
  int bz;                                     // planes per box
  int stages;
  unsigned stage_bytes;
  int      x_shift; // start coordinate offset in x (alignment experiments)
  int      consume; // 1: every thread reads one double per 8 of the stage before releasing it
  double  *sink;
};

__global__ void __launch_bounds__(256) tma_kernel(const __grid_constant__ CUtensorMap tmap, const Args a)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t sb   = smem_u32(smem_raw);
  const uint32_t bar0 = sb + a.stages * ((a.stage_bytes + 127) / 128 * 128);
  const unsigned sstr = (a.stage_bytes + 127) / 128 * 128;
  int            b    = blockIdx.x;
  const int      tx   = b % a.tiles_x;
  b /= a.tiles_x;
  const int ty    = b % a.tiles_y;
  const int chunk = b / a.tiles_y;
  const int x0 = tx * a.tx - a.halo + a.x_shift, y0 = ty * a.ty - a.halo;
  const int kbeg = chunk * a.lz - a.halo, kend = min(chunk * a.lz + a.lz, a.nz) + a.halo;
  const int nbox = (kend - kbeg + a.bz - 1) / a.bz;
  if (threadIdx.x == 0)
    {
      for (int s = 0; s < a.stages; ++s)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8 * s), "r"(1) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  __syncthreads();
  auto issue = [&](int i) {
    const int s = i % a.stages;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * s), "r"(a.stage_bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                   sb + s * sstr),
                 "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(bar0 + 8 * s), "r"(x0), "r"(y0), "r"(kbeg + i * a.bz)
                 : "memory");
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < a.stages && i < nbox; ++i)
      issue(i);
  double acc = 0.0;
  for (int i = 0; i < nbox; ++i)
    {
      const int      s  = i % a.stages;
      const unsigned ph = (i / a.stages) & 1;
      asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar0 + 8 * s),
        "r"(ph)
        : "memory");
      if (a.consume)
        {
          const double *p = reinterpret_cast<const double *>(smem_raw + s * sstr);
          for (unsigned e = threadIdx.x * 8; e < a.stage_bytes / 8; e += 256 * 8)
            acc += p[e];
        }
      __syncthreads();
      if (threadIdx.x == 0 && i + a.stages < nbox)
        issue(i + a.stages);
    }
  if (acc == 123.456)
    a.sink[0] = acc;
}

// cp.async staging of the same tiles: 16-byte chunks, one commit group per plane, `stages` groups in flight
struct LArgs
{
  const double *src;
  int64_t       pitch, plane;
  int           nx, ny, nzt; // array extents
  int           tiles_x, tiles_y, tx, ty, lz, nz, halo, stages;
  int           row_chunks, rows; // 16-byte chunks per tile row, rows per tile
  unsigned      stage_bytes;
  int           consume;
  double       *sink;
};

template <int STAGES>
__global__ void __launch_bounds__(256) ldgsts_kernel(const LArgs a)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t sb   = smem_u32(smem_raw);
  const unsigned sstr = (a.stage_bytes + 127) / 128 * 128;
  int            b    = blockIdx.x;
  const int      tx   = b % a.tiles_x;
  b /= a.tiles_x;
  const int ty    = b % a.tiles_y;
  const int chunk = b / a.tiles_y;
  const int x0 = tx * a.tx - a.halo + 1, y0 = ty * a.ty - a.halo; // +1: x0 even -> 16-byte aligned (pitch is a multiple of 4)
  const int kbeg = chunk * a.lz - a.halo, kend = min(chunk * a.lz + a.lz, a.nz) + a.halo;
  const int n    = kend - kbeg;
  const int nchunk = a.row_chunks * a.rows;
  auto issue = [&](int i) {
    const int k = kbeg + i;
    const int s = i % STAGES;
    for (int c = threadIdx.x; c < nchunk; c += 256)
      {
        const int r = c / a.row_chunks, q = c % a.row_chunks;
        const int gx = x0 + 2 * q, gy = y0 + r;
        const bool ok = (gx >= 0 && gx + 1 < a.nx && gy >= 0 && gy < a.ny && k >= 0 && k < a.nzt);
        const double *g = a.src + (ok ? ((int64_t)k * a.plane + (int64_t)gy * a.pitch + gx) : 0);
        const unsigned sz = ok ? 16u : 0u;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sb + s * sstr + c * 16), "l"(g), "r"(sz) : "memory");
      }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int i = 0; i < STAGES - 1; ++i)
    {
      if (i < n)
        issue(i);
      else
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
  double acc = 0.0;
  for (int i = 0; i < n; ++i)
    {
      if (i + STAGES - 1 < n)
        issue(i + STAGES - 1);
      else
        asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
      __syncthreads();
      if (a.consume)
        {
          const double *p = reinterpret_cast<const double *>(smem_raw + (i % STAGES) * sstr);
          for (unsigned e = threadIdx.x * 8; e < a.stage_bytes / 8; e += 256 * 8)
            acc += p[e];
        }
      __syncthreads();
    }
  if (acc == 123.456)
    a.sink[0] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
  const int     N = 257, pitch = 260;
  const int64_t plane = (int64_t)pitch * N, total = plane * N;
  double       *d, *d2, *sink;
  CK(cudaMalloc(&d, total * 8));
  CK(cudaMalloc(&d2, total * 8));
  CK(cudaMalloc(&sink, 8));
  CK(cudaMemset(d, 0, total * 8));
  CK(cudaMemset(d2, 0, total * 8));
  void                           *fp = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres));
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(fp);
  cudaEvent_t   e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int halo = 3;
  printf("{\"gpu\": \"B200\", \"field\": \"257^3 fp64, pitch 260\", \"results\": [\n");
  struct Case
  {
    int tx, ty, bz, stages, ctas_per_sm, lz, x_shift, consume, promo;
  };
  std::vector<Case> cases = {
    // the kron3d shape (38 x 38 x 1), 3 stages, 2 CTAs/SM, lz 29 and 64
    {32, 32, 1, 3, 2, 29, 1, 0, 2}, {32, 32, 1, 3, 2, 64, 1, 0, 2},   {32, 32, 1, 3, 2, 64, 1, 1, 2},
    {32, 32, 1, 6, 2, 64, 1, 0, 2}, {32, 32, 1, 3, 4, 64, 1, 0, 2},   {32, 32, 1, 6, 4, 64, 1, 0, 2},
    {32, 32, 1, 3, 2, 64, 0, 0, 2}, // unaligned (odd) box start
    {32, 32, 1, 3, 2, 64, 1, 0, 0}, // no L2 promotion
    {32, 32, 1, 3, 2, 64, 1, 0, 3}, // 256 B promotion
    {32, 32, 2, 3, 2, 64, 1, 0, 2}, // two planes per box
    {32, 32, 4, 2, 2, 64, 1, 0, 2}, // four planes per box
    {64, 32, 1, 3, 2, 64, 1, 0, 2}, // 70 x 38
    {64, 32, 1, 3, 1, 128, 1, 0, 2}, {128, 16, 1, 3, 2, 64, 1, 0, 2}, // 134 x 22
    {256, 8, 1, 3, 2, 64, 1, 0, 2},                                   // 262 x 14 (full rows)
    {32, 64, 1, 3, 1, 128, 1, 0, 2},                                  // 38 x 70
    {32, 32, 1, 3, 1, 128, 1, 0, 2},
  };
  int dev_sms = 148;
  for (const Case &c : cases)
    {
      const int bx = c.tx + 2 * halo, by = c.ty + 2 * halo;
      CUtensorMap m;
      cuuint64_t  dims[3]    = {(cuuint64_t)N, (cuuint64_t)N, (cuuint64_t)N};
      cuuint64_t  strides[2] = {(cuuint64_t)pitch * 8, (cuuint64_t)plane * 8};
      cuuint32_t  box[3]     = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)c.bz};
      cuuint32_t  estr[3]    = {1, 1, 1};
      CUresult    rc = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)c.promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rc != CUDA_SUCCESS)
        {
          printf(" {\"error\": \"encode %d\"},\n", (int)rc);
          continue;
        }
      Args a;
      a.tx = c.tx, a.ty = c.ty, a.halo = halo, a.bz = c.bz, a.stages = c.stages, a.lz = c.lz, a.nz = N;
      a.tiles_x     = (N - 2 + c.tx - 1) / c.tx;
      a.tiles_y     = (N - 2 + c.ty - 1) / c.ty;
      a.stage_bytes = (unsigned)bx * by * c.bz * 8;
      a.x_shift     = c.x_shift;
      a.consume     = c.consume;
      a.sink        = sink;
      const int    chunks = (N + c.lz - 1) / c.lz;
      const int    grid   = a.tiles_x * a.tiles_y * chunks;
      const size_t smem   = (size_t)c.stages * ((a.stage_bytes + 127) / 128 * 128) + 64;
      // limit residency to ctas_per_sm through the shared memory request
      const size_t smem_req = std::max(smem, (size_t)(227 * 1024 / c.ctas_per_sm - 2048));
      CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_req));
      float best = 1e30f;
      for (int rep = 0; rep < 6; ++rep)
        {
          CK(cudaMemsetAsync(d2, 0, total * 8)); // flush L2 (136 MB)
          CK(cudaEventRecord(e0));
          tma_kernel<<<grid, 256, smem_req>>>(m, a);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          CK(cudaGetLastError());
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          if (rep > 0)
            best = std::min(best, ms);
        }
      double staged = 0; // bytes requested from L2
      for (int ch = 0; ch < chunks; ++ch)
        {
          const int kb = ch * c.lz - halo, ke = std::min(ch * c.lz + c.lz, N) + halo;
          staged += (double)((ke - kb + c.bz - 1) / c.bz) * a.stage_bytes * a.tiles_x * a.tiles_y;
        }
      printf(" {\"kind\": \"tma\", \"box\": [%d, %d, %d], \"stages\": %d, \"ctas_per_sm\": %d, \"lz\": %d, \"x_shift\": %d, "
             "\"consume\": %d, \"l2promo\": %d, \"grid\": %d, \"ms\": %.4f, \"field_GBs\": %.0f, \"staged_GBs\": %.0f, "
             "\"equiv_GDoFs_at_16B\": %.0f},\n",
             bx, by, c.bz, c.stages, c.ctas_per_sm, c.lz, c.x_shift, c.consume, c.promo, grid, best, total * 8 / best / 1e6,
             staged / best / 1e6, (double)N * N * N / best / 1e6);
      fflush(stdout);
    }
  // ---- cp.async staging
  for (int cps : {2, 4})
    for (int consume : {0, 1})
      for (int lz : {29, 64})
        {
          LArgs a;
          a.src = d, a.pitch = pitch, a.plane = plane, a.nx = pitch, a.ny = N, a.nzt = N;
          a.tx = 32, a.ty = 32, a.halo = halo, a.lz = lz, a.nz = N, a.stages = 3;
          a.tiles_x = 8, a.tiles_y = 8;
          a.row_chunks  = 19;
          a.rows        = 38;
          a.stage_bytes = 19 * 16 * 38;
          a.consume     = consume;
          a.sink        = sink;
          const int    chunks   = (N + lz - 1) / lz;
          const int    grid     = 64 * chunks;
          const size_t smem     = 3 * ((a.stage_bytes + 127) / 128 * 128) + 64;
          const size_t smem_req = std::max(smem, (size_t)(227 * 1024 / cps - 2048));
          CK(cudaFuncSetAttribute(ldgsts_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_req));
          float best = 1e30f;
          for (int rep = 0; rep < 6; ++rep)
            {
              CK(cudaMemsetAsync(d2, 0, total * 8));
              CK(cudaEventRecord(e0));
              ldgsts_kernel<3><<<grid, 256, smem_req>>>(a);
              CK(cudaEventRecord(e1));
              CK(cudaEventSynchronize(e1));
              CK(cudaGetLastError());
              float ms;
              CK(cudaEventElapsedTime(&ms, e0, e1));
              if (rep > 0)
                best = std::min(best, ms);
            }
          printf(" {\"kind\": \"ldgsts\", \"box\": [38, 38, 1], \"stages\": 3, \"ctas_per_sm\": %d, \"lz\": %d, \"consume\": %d, "
                 "\"grid\": %d, \"ms\": %.4f, \"field_GBs\": %.0f, \"equiv_GDoFs_at_16B\": %.0f},\n",
                 cps, lz, consume, grid, best, total * 8 / best / 1e6, (double)N * N * N / best / 1e6);
          fflush(stdout);
        }
  (void)dev_sms;
  printf(" {\"done\": true}]}\n");
  return 0;
}
