// Cost of the per-k epilogue's scalar FP64 ops under different placements, 16 warps per SM, free running.
//  mode 0: 160 DMMA only            mode 1: [160 DMMA][16 x (DMUL, dependent DFMA)] in a block
//  mode 2: same 32 scalar ops spread evenly between the DMMAs (1 per 5 DMMA)
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void k(double* sink, int iters, double a0) {
  double c[8][2] = {}, v[16];
  unsigned cnt[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int st = 0; st < 20; ++st) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dmma(c[i][0], c[i][1], a, b);
        if (MODE == 2) {
          const int n = st * 8 + i;
          if (n % 5 == 0) {
            const int o = (n / 5) % 16;
            if ((n / 5) < 16) { double d; asm volatile("mul.rn.f64 %0, %1, %1;" : "=d"(d) : "d"(v[o])); v[o] = d; }
            else asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(v[o]) : "d"(1.0000001), "d"(1e-9));
          }
        }
      }
    }
    if (MODE == 3) {
      // block epilogue with the scan kernel's integer/select volume: per output DMUL, DFMA, 64-bit compare, 2 selects, counter add
#pragma unroll
      for (int o = 0; o < 16; ++o) {
        double d; asm volatile("mul.rn.f64 %0, %1, %1;" : "=d"(d) : "d"(c[o & 7][o >> 3]));
        double vv; asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(vv) : "d"(d), "d"(-1.0000001), "d"(1.5));
        const bool better = __double_as_longlong(vv) < __double_as_longlong(v[o]);
        v[o] = better ? vv : v[o];
        cnt[o >> 2] += better ? (1u << ((o & 3) * 8)) : 0u;
        c[o & 7][o >> 3] = 0.0;
      }
    }
    if (MODE == 1) {
#pragma unroll
      for (int o = 0; o < 16; ++o) { double d; asm volatile("mul.rn.f64 %0, %1, %1;" : "=d"(d) : "d"(v[o])); v[o] = d; }
#pragma unroll
      for (int o = 0; o < 16; ++o) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(v[o]) : "d"(1.0000001), "d"(1e-9));
    }
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  for (int i = 0; i < 16; ++i) s += v[i];
  s += cnt[0] + cnt[1] + cnt[2] + cnt[3];
  if (s == 1.2345) sink[0] = s;
}
template <int MODE> float run(double* sink, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, 512>>>(sink, iters, 1.0); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148, 512>>>(sink, iters, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double* sink; cudaMalloc(&sink, 64);
  const int iters = 2000;
  float t0 = run<0>(sink, iters), t1 = run<1>(sink, iters), t2 = run<2>(sink, iters), t3 = run<3>(sink, iters);
  // per SMSP per iteration: 4 warps x 32 scalar ops = 128
  double clk = 1.965e6;  // cycles per ms
  printf("{\"full_epilogue_ms\": %.3f, \"cycles_per_iter_full_epilogue\": %.1f}\n", t3, (t3 - t0) * clk / iters);
  printf("{\"dmma_only_ms\": %.3f, \"block_epilogue_ms\": %.3f, \"interleaved_ms\": %.3f, \"cycles_per_scalar_op_block\": %.2f, \"cycles_per_scalar_op_interleaved\": %.2f}\n",
         t0, t1, t2, (t1 - t0) * clk / iters / 128.0, (t2 - t0) * clk / iters / 128.0);
  return 0;
}
