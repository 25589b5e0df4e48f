// FP64 issue cost as a function of operand kinds (register vs uniform/constant) and instruction kind.
// One CTA per SM x 4 SMSPs x WPS warps; reports cycles per warp-instruction per SMSP.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ILP = 8;
struct K { double c[8]; };

template <int MODE>
__global__ void kern(double *out, const double *in, long long *cycles, int iters, const K k)
{
  double a[ILP], b[ILP], c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 32 * i + 7]; c[i] = in[threadIdx.x + 3 * i + 1]; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
    {
#pragma unroll
      for (int i = 0; i < ILP; ++i)
        {
          if (MODE == 0) c[i] = fma(a[i], b[i], c[i]);            // DFMA R,R,R  (3 distinct register pairs)
          if (MODE == 1) c[i] = fma(a[i], k.c[i & 7], c[i]);      // DFMA R,c[],R (uniform / constant coefficient)
          if (MODE == 2) c[i] = a[i] + c[i];                      // DADD R,R
          if (MODE == 3) c[i] = c[i] * k.c[i & 7];                // DMUL R,c[]
          if (MODE == 4) c[i] = fma(c[i], k.c[i & 7], k.c[(i + 1) & 7]); // DFMA R,c[],c[]... (one register)
          if (MODE == 5) c[i] = fma(a[i], a[i], c[i]);            // DFMA R,R(same),R
          if (MODE == 6) c[i] = fma(c[i], b[i], c[i]);            // DFMA with a == c
        }
    }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i] + a[i] + b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int MODE>
void run(const char *name, double *out, double *in, long long *cyc, int warps_per_smsp, int sms)
{
  K k; for (int i = 0; i < 8; ++i) k.c[i] = 1.0 + 1e-9 * i;
  const int iters = 4000, threads = 128 * warps_per_smsp;
  kern<MODE><<<sms, threads>>>(out, in, cyc, iters, k);
  kern<MODE><<<sms, threads>>>(out, in, cyc, iters, k);
  long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  // per SMSP: warps_per_smsp warps each issuing iters*ILP instructions in h cycles
  printf(" \"%s_wps%d_cycles_per_instr_per_smsp\": %.2f,\n", name, warps_per_smsp, (double)h / ((double)iters * ILP * warps_per_smsp));
}

int main()
{
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  double *out, *in; long long *cyc;
  CK(cudaMalloc(&out, 1 << 24)); CK(cudaMalloc(&in, 1 << 16)); CK(cudaMalloc(&cyc, 8));
  CK(cudaMemset(in, 0, 1 << 16));
  printf("{\"gpu\": \"%s\",\n", prop.name);
  for (int w : {1, 2, 4})
    {
      run<0>("dfma_rrr", out, in, cyc, w, prop.multiProcessorCount);
      run<1>("dfma_rcr", out, in, cyc, w, prop.multiProcessorCount);
      run<2>("dadd_rr", out, in, cyc, w, prop.multiProcessorCount);
      run<3>("dmul_rc", out, in, cyc, w, prop.multiProcessorCount);
      run<4>("dfma_rcc", out, in, cyc, w, prop.multiProcessorCount);
      run<5>("dfma_raar", out, in, cyc, w, prop.multiProcessorCount);
      run<6>("dfma_cbc", out, in, cyc, w, prop.multiProcessorCount);
    }
  printf(" \"done\": true}\n");
  return 0;
}
