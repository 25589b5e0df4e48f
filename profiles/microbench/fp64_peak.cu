// Microbenchmarks that fix the roofline denominators DESIGN.md quotes:
//   1. FP64 FMA (DFMA) throughput, 2. FP64 tensor (DMMA) throughput alone and interleaved with DFMA,
//   3. HBM copy / read bandwidth with 128-bit accesses, 4. shared-memory LDS.64 / LDS.128 bandwidth.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu ; run on a B200.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int ILP>
__global__ void dfma_kernel(double *out, int iters, double a, double b)
{
  double v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i];
  if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int ILP, int FMA_PER_MMA>
__global__ void dmma_kernel(double *out, int iters, double a, double b)
{
  double c[ILP][2];
  double v[8];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = i + threadIdx.x * 1e-3;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      {
        dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int j = 0; j < FMA_PER_MMA; ++j) v[(i * FMA_PER_MMA + j) & 7] = fma(v[(i * FMA_PER_MMA + j) & 7], a, b);
      }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  if (s == 123.456) out[0] = s;
}

__global__ void copy_kernel(const double2 *__restrict__ a, double2 *__restrict__ b, size_t n)
{
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) b[i] = a[i];
}
__global__ void read_kernel(const double2 *__restrict__ a, double *out, size_t n)
{
  size_t stride = (size_t)gridDim.x * blockDim.x;
  double s = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) { double2 v = a[i]; s += v.x + v.y; }
  if (s == 123.456) out[0] = s;
}

template <int VEC>
__global__ void lds_kernel(double *out, int iters)
{
  __shared__ double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double s = 0;
  int idx = threadIdx.x * VEC;
  for (int it = 0; it < iters; ++it)
    {
#pragma unroll
      for (int u = 0; u < 8; ++u)
        {
          if (VEC == 1) s += sm[(idx + u * 512) & 4095];
          else { double2 v = *reinterpret_cast<double2 *>(&sm[(idx + u * 512) & 4095]); s += v.x + v.y; }
        }
      idx = (idx + 64) & 4095;
    }
  if (s == 123.456) out[0] = s;
}

// dependent-chain latency and single-warp throughput as a function of ILP
template <int ILP>
__global__ void dfma_latency_kernel(double *out, long long *cycles, int iters, double a, double b)
{
  double v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i];
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int ILP>
void latency_line(double *out, int warps)
{
  long long *cyc; CK(cudaMalloc(&cyc, 8));
  const int iters = 2000;
  dfma_latency_kernel<ILP><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9);
  dfma_latency_kernel<ILP><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9);
  long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf(" \"dfma_cycles_per_instr_ilp%d_warps%d\": %.2f,\n", ILP, warps, (double)h / (iters * ILP));
  cudaFree(cyc);
}

template <typename F>
float time_ms(F f, int reps = 5)
{
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r)
    {
      CK(cudaEventRecord(e0));
      f();
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
  return best;
}

int main()
{
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", prop.name, sms, prop.clockRate);
  double *out; CK(cudaMalloc(&out, 1024));
  const int iters = 4096;
  // one CTA on one SM: cycles per DFMA warp-instruction seen by warp 0 (ILP 1 = dependent-issue latency)
  latency_line<1>(out, 1); latency_line<2>(out, 1); latency_line<4>(out, 1); latency_line<8>(out, 1); latency_line<16>(out, 1);
  latency_line<1>(out, 4); latency_line<1>(out, 8); latency_line<1>(out, 16); latency_line<2>(out, 16); latency_line<4>(out, 16); latency_line<8>(out, 8);
  {
    const int threads = 512, blocks = sms * 4, ILP = 8;
    float ms = time_ms([&] { dfma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double flops = 2.0 * ILP * iters * (double)threads * blocks;
    printf(" \"dfma_tflops\": %.2f,\n", flops / ms / 1e9);
  }
  {
    const int threads = 512, blocks = sms * 4, ILP = 8;
    float ms = time_ms([&] { dmma_kernel<ILP, 0><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double flops = 512.0 * ILP * iters * (double)(threads / 32) * blocks;
    printf(" \"dmma884_tflops\": %.2f,\n", flops / ms / 1e9);
    ms = time_ms([&] { dmma_kernel<ILP, 1><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double f2 = (512.0 + 64.0 * 1) * ILP * iters * (double)(threads / 32) * blocks;
    printf(" \"dmma884_plus_1dfma_tflops\": %.2f,\n", f2 / ms / 1e9);
    ms = time_ms([&] { dmma_kernel<ILP, 4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    f2 = (512.0 + 64.0 * 4) * ILP * iters * (double)(threads / 32) * blocks;
    printf(" \"dmma884_plus_4dfma_tflops\": %.2f,\n", f2 / ms / 1e9);
    ms = time_ms([&] { dmma_kernel<ILP, 8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    f2 = (512.0 + 64.0 * 8) * ILP * iters * (double)(threads / 32) * blocks;
    printf(" \"dmma884_plus_8dfma_tflops\": %.2f,\n", f2 / ms / 1e9);
  }
  {
    size_t n = (size_t)1 << 29; // 512 Mi doubles = 4 GiB per buffer
    double2 *a, *b; CK(cudaMalloc(&a, n * 8)); CK(cudaMalloc(&b, n * 8));
    CK(cudaMemset(a, 1, n * 8)); CK(cudaMemset(b, 0, n * 8));
    float ms = time_ms([&] { copy_kernel<<<sms * 16, 512>>>(a, b, n / 2); });
    printf(" \"hbm_copy_gbs\": %.1f,\n", 2.0 * n * 8 / ms / 1e6);
    ms = time_ms([&] { read_kernel<<<sms * 16, 512>>>(a, out, n / 2); });
    printf(" \"hbm_read_gbs\": %.1f,\n", 1.0 * n * 8 / ms / 1e6);
    ms = time_ms([&] { CK(cudaMemcpyAsync(b, a, n * 8, cudaMemcpyDeviceToDevice)); });
    printf(" \"hbm_memcpy_gbs\": %.1f,\n", 2.0 * n * 8 / ms / 1e6);
    cudaFree(a); cudaFree(b);
  }
  {
    const int threads = 512, blocks = sms * 4, it2 = 20000;
    float ms = time_ms([&] { lds_kernel<1><<<blocks, threads>>>(out, it2); });
    printf(" \"lds64_bytes_per_clk_per_sm_at_1.9GHz\": %.1f,\n", 8.0 * 8 * it2 * (double)threads * blocks / (ms * 1e-3) / sms / 1.9e9);
    ms = time_ms([&] { lds_kernel<2><<<blocks, threads>>>(out, it2); });
    printf(" \"lds128_bytes_per_clk_per_sm_at_1.9GHz\": %.1f,\n", 16.0 * 8 * it2 * (double)threads * blocks / (ms * 1e-3) / sms / 1.9e9);
  }
  printf(" \"done\": true}\n");
  return 0;
}
