// FP64 pipe micro-benchmarks for sm_100a (B200): dependent-issue latency of DFMA / MUFU.RCP64H / LDS, and single-SMSP
// DFMA throughput as a function of (warps per SMSP) x (independent chains per thread).  Numbers go to DESIGN.md.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_chain(double* out, long long* cyc, int iters, double m, double b) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  if (s == 1234.5) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void rcp_chain(double* out, long long* cyc, int iters) {
  double a = 1.5 + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      double r;
      asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
      a = r;
    }
  }
  long long t1 = clock64();
  if (a == 1234.5) out[0] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void lds_chain(double* out, long long* cyc, int iters) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i + 32) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) p = idx[p];
  }
  long long t1 = clock64();
  if (p == -1) out[0] = p;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP>
static void run_dfma(int warps_per_block, double* d, long long* c) {
  const int iters = 2048;
  dfma_chain<ILP><<<1, 32 * warps_per_block>>>(d, c, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  dfma_chain<ILP><<<1, 32 * warps_per_block>>>(d, c, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  const double n = (double)iters * 16 * ILP;   // DFMA per thread
  // one block -> one SM, warps spread over 4 SMSPs
  printf("DFMA ilp=%d warps/SM=%2d: %.2f cycles per dependent DFMA step; %.2f DFMA/clk/SM (thread-level)\n", ILP,
         warps_per_block, (double)h / (iters * 16.0), n * 32 * warps_per_block / (double)h);
}

int main() {
  double* d;
  long long* c;
  cudaMalloc(&d, 8);
  cudaMalloc(&c, 8);
  for (int w : {1, 4, 8, 16, 32}) {
    run_dfma<1>(w, d, c);
    run_dfma<2>(w, d, c);
    run_dfma<4>(w, d, c);
    run_dfma<8>(w, d, c);
  }
  long long h;
  rcp_chain<<<1, 32>>>(d, c, 1024);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("MUFU.RCP64H dependent: %.2f cycles\n", (double)h / (1024 * 16.0));
  lds_chain<<<1, 32>>>(d, c, 1024);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("LDS dependent: %.2f cycles\n", (double)h / (1024 * 16.0));
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
