// mbarrier-token alternation between two 2-warp groups per sub-partition; and bar.sync variant with ids 8..15
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ bool mb_try(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bsync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void barrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int MODE>
__global__ void k(long long* out, int iters) {
  __shared__ uint64_t turn[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / 8, sm = warp & 3;
  if (threadIdx.x < 8) mb_init(&turn[threadIdx.x], 2);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint64_t* mine = &turn[sm * 2 + wm];
  uint64_t* theirs = &turn[sm * 2 + (wm ^ 1)];
  const int bmine = 8 + (MODE == 2 ? (sm & 1) * 2 + wm : sm * 2 + wm), btheirs = 8 + (MODE == 2 ? (sm & 1) * 2 + (wm ^ 1) : sm * 2 + (wm ^ 1));
  const int cnt = MODE == 2 ? 256 : 128;
  if (MODE == 0) { if (wm == 1 && lane == 0) mb_arrive(theirs); }
  else { if (wm == 1) barrive(btheirs, cnt); }
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) { while (!mb_try(mine, it & 1)) {} }
    else bsync(bmine, cnt);
    long long t0 = clock64();
    while (clock64() - t0 < 2000) {}
    long long t1 = clock64();
    if (MODE == 0) { __syncwarp(); if (lane == 0) mb_arrive(theirs); }
    else barrive(btheirs, cnt);
    while (clock64() - t1 < 300) {}
    if (lane == 0 && it < 8) { out[(warp * 8 + it) * 2] = t0; out[(warp * 8 + it) * 2 + 1] = t1; }
  }
  if (MODE != 0 && wm == 0) bsync(bmine, cnt);
}
template <int MODE>
void run(long long* d) {
  k<MODE><<<1, 512>>>(d, 8);
  long long h[16 * 8 * 2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("mode %d err %d\n", MODE, (int)cudaGetLastError());
  long long base = h[0];
  for (int w : {0, 4, 8, 12, 1, 9}) { printf("warp %2d:", w); for (int it = 0; it < 5; ++it) printf(" [%lld,%lld]", h[(w * 8 + it) * 2] - base, h[(w * 8 + it) * 2 + 1] - base); printf("\n"); }
}
int main() {
  long long* d; cudaMalloc(&d, 16 * 8 * 2 * 8);
  run<0>(d);  // mbarrier tokens
  run<1>(d);  // bar ids 8..15, 128 threads
  run<2>(d);  // bar ids 8..11, 256 threads (pairs of sub-partitions)
  return 0;
}
