// FP64 issue cost with uniform-register (UR) coefficient operands and with the stencil instruction mix of kron3d
// (pair sums + FMA chains).  Reports cycles per warp-instruction per SMSP for 1/2/4 warps per scheduler.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ILP = 8;
struct K { double c[32]; };

template <int MODE>
__global__ void kern(double *out, const double *in, long long *cycles, int iters, int sel, const __grid_constant__ K k)
{
  double a[ILP], c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a[i] = in[threadIdx.x + 32 * i]; c[i] = in[threadIdx.x + 3 * i + 1]; }
  double v[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) v[i] = in[threadIdx.x + 5 * i + 2];
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
    {
      if (MODE == 0)
        {
#pragma unroll
          for (int i = 0; i < ILP; ++i) c[i] = fma(a[i], k.c[i], c[i]); // constant-bank operand
        }
      if (MODE == 1)
        {
#pragma unroll
          for (int i = 0; i < ILP; ++i) c[i] = fma(a[i], k.c[sel + i], c[i]); // dynamically indexed (uniform) coefficient
        }
      if (MODE == 2)
        {
          // x-pass mix: 4 outputs from a 10-wide window, 2 fields: 3 DADD + (1 DMUL + 3 DFMA) + (1 DMUL + 2 DFMA) per output = 10 ops
          double o[4], q[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            {
              const double s1 = v[j + 2] + v[j + 4], s2 = v[j + 1] + v[j + 5], s3 = v[j] + v[j + 6];
              o[j] = fma(k.c[3], s3, fma(k.c[2], s2, fma(k.c[1], s1, k.c[0] * v[j + 3])));
              q[j] = fma(k.c[6], s2, fma(k.c[5], s1, k.c[4] * v[j + 3]));
            }
#pragma unroll
          for (int j = 0; j < 4; ++j) { v[j + 3] = o[j]; v[j + 7] = q[j]; }
        }
      if (MODE == 3)
        {
          // z-pass mix: 6 accumulators per point, 2 points: acc[j-1] = zA[j]*ua + zB[j]*p + acc[j]
#pragma unroll
          for (int p = 0; p < 2; ++p)
            {
              const double ua = fma(k.c[20], a[p], a[p + 2]);
              double       t  = fma(k.c[7], ua, v[6 * p]);
              a[p + 4] += t;
#pragma unroll
              for (int j = 1; j < 6; ++j) v[6 * p + j - 1] = fma(k.c[14 + j], a[p], fma(k.c[7 + j], ua, v[6 * p + j]));
              v[6 * p + 5] = k.c[13] * ua;
            }
        }
    }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i] + a[i];
#pragma unroll
  for (int i = 0; i < 12; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int MODE>
void run(const char *name, int ops, double *out, double *in, long long *cyc, int warps_per_smsp, int sms)
{
  K k; for (int i = 0; i < 32; ++i) k.c[i] = 0.5 + 1e-3 * i;
  const int iters = 4000, threads = 128 * warps_per_smsp;
  kern<MODE><<<sms, threads>>>(out, in, cyc, iters, 0, k);
  kern<MODE><<<sms, threads>>>(out, in, cyc, iters, 0, k);
  long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf(" \"%s_wps%d_cycles_per_fp64_instr_per_smsp\": %.2f,\n", name, warps_per_smsp, (double)h / ((double)iters * ops * warps_per_smsp));
}

int main()
{
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  double *out, *in; long long *cyc;
  CK(cudaMalloc(&out, 1 << 24)); CK(cudaMalloc(&in, 1 << 16)); CK(cudaMalloc(&cyc, 8));
  CK(cudaMemset(in, 0, 1 << 16));
  printf("{\"gpu\": \"%s\",\n", prop.name);
  for (int w : {1, 2, 4, 8})
    {
      run<0>("dfma_const", ILP, out, in, cyc, w, prop.multiProcessorCount);
      run<1>("dfma_dynidx", ILP, out, in, cyc, w, prop.multiProcessorCount);
      run<2>("xpass_mix", 40, out, in, cyc, w, prop.multiProcessorCount);
      run<3>("zpass_mix", 28, out, in, cyc, w, prop.multiProcessorCount);
    }
  printf(" \"done\": true}\n");
  return 0;
}
