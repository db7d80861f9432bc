// FP64 roofline denominators for B200, measured (MEASURED_PEAKS.json holds only HBM and bf16):
//   (a) DMMA.8x8x4 issue peak from registers (mma.sync m8n8k4 f64), by warps per SM
//   (b) DFMA peak from registers
//   (c) cuBLAS DGEMM 8192^3: best of 10 and 4 s sustained
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu -lcublas
// Prints one JSON object.
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int NACC>
__global__ void dmma_kernel(double* out, int iters, double a0, double b0) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void dfma_kernel(double* out, int iters, double a0, double b0) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) asm volatile("fma.rn.f64 %0, %1, %0, %2;" : "+d"(c[i]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  if (s == 12345.678) out[0] = s;
}

// mixed: R DFMA per DMMA, to see whether the two share an issue pipe
template <int R>
__global__ void mixed_kernel(double* out, int iters, double a0, double b0) {
  double c[8][2], f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = c[i][1] = 0.0; f[i] = i; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dmma(c[i][0], c[i][1], a, b);
#pragma unroll
      for (int r = 0; r < R; ++r) asm volatile("fma.rn.f64 %0, %1, %0, %2;" : "+d"(f[(i + r) & 7]) : "d"(a), "d"(b));
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i];
  if (s == 12345.678) out[0] = s;
}

// realistic inner loop: 4x4 (or 4x2) atoms per warp, fragments re-loaded from shared memory every k-step
// with the scan kernel's [row][KC=20] layout; no epilogue, no barriers.
template <int BT>
__global__ void dmma_tile_kernel(double* out, int iters) {
  __shared__ double sa[64 * 20], sb[128 * 20];
  for (int i = threadIdx.x; i < 64 * 20; i += blockDim.x) sa[i] = 1e-3 * i;
  for (int i = threadIdx.x; i < 128 * 20; i += blockDim.x) sb[i] = 1e-4 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const double* ap = sa + ((warp & 1) * 32 + g) * 20 + t;
  const double* bp = sb + (((warp >> 1) * 8 * BT) % 128 + g) * 20 + t;
  double c[4][BT][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < BT; ++b) c[a][b][0] = c[a][b][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int st = 0; st < 5; ++st) {
      double af[4], bf[BT];
#pragma unroll
      for (int a = 0; a < 4; ++a) af[a] = ap[a * 160 + st * 4];
#pragma unroll
      for (int b = 0; b < BT; ++b) bf[b] = bp[b * 160 + st * 4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < BT; ++b)
          asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
              : "+d"(c[a][b][0]), "+d"(c[a][b][1]) : "d"(af[a]), "d"(bf[b]));
    }
  }
  double s = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < BT; ++b) s += c[a][b][0] + c[a][b][1];
  if (s == 12345.678) out[0] = s;
}

template <typename F>
float time_ms(F&& launch, int reps = 5) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double* out;
  cudaMalloc(&out, 64);
  const int iters = 20000;
  printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
  // (a) DMMA: 8x8x4 = 256 FMA = 512 flop per warp instruction
  for (int warps : {4, 8, 16, 32}) {
    float ms = time_ms([&] { dmma_kernel<8><<<sms, warps * 32>>>(out, iters, 1.0, 1e-3); });
    double tf = (double)sms * warps * iters * 8 * 512.0 / (ms * 1e-3) / 1e12;
    printf(", \"dmma_tflops_w%d\": %.3f", warps, tf);
  }
  {
    float ms = time_ms([&] { dmma_kernel<16><<<sms, 8 * 32>>>(out, iters, 1.0, 1e-3); });
    printf(", \"dmma_tflops_w8_acc16\": %.3f", (double)sms * 8 * iters * 16 * 512.0 / (ms * 1e-3) / 1e12);
  }
  {
    const int it2 = 4000;
    float ms = time_ms([&] { dmma_tile_kernel<4><<<sms, 8 * 32>>>(out, it2); });
    printf(", \"dmma_tile_4x4_w8_tflops\": %.3f", (double)sms * 8 * it2 * 5 * 16 * 512.0 / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { dmma_tile_kernel<2><<<sms, 16 * 32>>>(out, it2); });
    printf(", \"dmma_tile_4x2_w16_tflops\": %.3f", (double)sms * 16 * it2 * 5 * 8 * 512.0 / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { dmma_tile_kernel<4><<<sms, 4 * 32>>>(out, it2); });
    printf(", \"dmma_tile_4x4_w4_tflops\": %.3f", (double)sms * 4 * it2 * 5 * 16 * 512.0 / (ms * 1e-3) / 1e12);
  }
  // (b) DFMA: 32 FMA = 64 flop per warp instruction
  for (int warps : {8, 16, 32}) {
    float ms = time_ms([&] { dfma_kernel<8><<<sms, warps * 32>>>(out, iters, 1.0000001, 1e-3); });
    double tf = (double)sms * warps * iters * 8 * 64.0 / (ms * 1e-3) / 1e12;
    printf(", \"dfma_tflops_w%d\": %.3f", warps, tf);
  }
  // (c) mixed, 8 warps/SM: time relative to DMMA-only tells whether DFMA issues in the shadow of DMMA
  {
    float base = time_ms([&] { dmma_kernel<8><<<sms, 8 * 32>>>(out, iters, 1.0, 1e-3); });
    float m1 = time_ms([&] { mixed_kernel<1><<<sms, 8 * 32>>>(out, iters, 1.0000001, 1e-3); });
    float m2 = time_ms([&] { mixed_kernel<2><<<sms, 8 * 32>>>(out, iters, 1.0000001, 1e-3); });
    float m4 = time_ms([&] { mixed_kernel<4><<<sms, 8 * 32>>>(out, iters, 1.0000001, 1e-3); });
    printf(", \"mixed_ms\": {\"dmma_only\": %.3f, \"plus1_dfma\": %.3f, \"plus2_dfma\": %.3f, \"plus4_dfma\": %.3f}", base, m1,
           m2, m4);
  }
  // (d) cuBLAS DGEMM
  {
    const int N = 8192;
    double *A, *B, *C;
    cudaMalloc(&A, sizeof(double) * N * N);
    cudaMalloc(&B, sizeof(double) * N * N);
    cudaMalloc(&C, sizeof(double) * N * N);
    cudaMemset(A, 0, sizeof(double) * N * N);
    cudaMemset(B, 0, sizeof(double) * N * N);
    std::vector<double> h((size_t)N * N);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (double)(rand() % 1000) / 1000.0 - 0.5;
    cudaMemcpy(A, h.data(), sizeof(double) * N * N, cudaMemcpyHostToDevice);
    cudaMemcpy(B, h.data(), sizeof(double) * N * N, cudaMemcpyHostToDevice);
    cublasHandle_t hd;
    cublasCreate(&hd);
    const double one = 1.0, zero = 0.0;
    auto gemm = [&] { cublasDgemm(hd, CUBLAS_OP_T, CUBLAS_OP_N, N, N, N, &one, A, N, B, N, &zero, C, N); };
    float best = time_ms(gemm, 10);
    printf(", \"cublas_dgemm_tflops_burst\": %.3f", 2.0 * N * N * (double)N / (best * 1e-3) / 1e12);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int reps = (int)(4000.0 / best) + 1;
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) gemm();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf(", \"cublas_dgemm_tflops_sustained\": %.3f", 2.0 * N * N * (double)N * reps / (ms * 1e-3) / 1e12);
  }
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf(", \"clock_khz_max\": %d}\n", clk);
  return 0;
}
