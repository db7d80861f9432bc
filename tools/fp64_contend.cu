// How do scalar FP64 ops behave while other warps of the same SM sub-partition stream DMMAs?
// Warps 0-3 (one per SMSP) run a dependent DFMA chain (or 16 interleaved chains) and time it with
// clock64; the remaining warps run DMMA streams (or nothing).  Prints cycles per scalar FP64 op.
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void contend(long long* out, double* sink, int iters, int dmma_warps_per_smsp, double a0) {
  const int warp = threadIdx.x >> 5;
  if (warp < 4) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a0 + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(1.0000001), "d"(1e-9));
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (threadIdx.x % 32 == 0 && blockIdx.x == 0) out[warp] = t1 - t0;
    if (s == 1.2345) sink[0] = s;
  } else {
    __syncthreads();
    if ((warp - 4) / 4 < dmma_warps_per_smsp) {
      double c[8][2] = {};
      double a = a0, b = 1e-3;
      // run long enough to cover the timed warps
      for (int it = 0; it < iters * 4; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
      }
      double s = 0;
      for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
      if (s == 1.2345) sink[1] = s;
    }
  }
}

int main() {
  long long* out; double* sink;
  cudaMallocManaged(&out, 64); cudaMalloc(&sink, 64);
  const int iters = 2000;
  for (int nd = 0; nd <= 3; ++nd) {
    contend<1><<<148, 32 * 16>>>(out, sink, iters, nd, 1.0); cudaDeviceSynchronize();
    double lat1 = (double)out[0] / (iters * 1);
    contend<16><<<148, 32 * 16>>>(out, sink, iters, nd, 1.0); cudaDeviceSynchronize();
    double lat16 = (double)out[0] / (iters * 16);
    printf("{\"dmma_warps_per_smsp\": %d, \"dfma_dependent_cycles_per_op\": %.1f, \"dfma_ilp16_cycles_per_op\": %.2f}\n", nd, lat1, lat16);
  }
  return 0;
}
