// PREVIOUS GENERATION (v3) of the fused scan kernel, kept for A/B timing against blmm_scan.cu
// (BLMM_SCAN_KERNEL=v3).
// The fused marker x trait scan kernel (see ScanParams in blmm_kernels.cuh for the arithmetic).
//
// Structure (sm_100a): persistent CTAs, one per SM, 8 warps (2 per SM sub-partition, so every
// thread may hold up to 255 registers: accumulators + running minima + counters + double-buffered
// fragments stay in registers).  A CTA owns a contiguous range of (trait tile, marker tile) units.
// The trait tile (128 traits x whole K) stays resident in shared memory; for every unit the
// k-list's marker tiles (64 markers x whole K, plus that k's per-trait scalars e/et) stream
// through a 2-3 stage ring filled with 1-D bulk asynchronous copies (cp.async.bulk -> UBLKCP, the
// TMA engine; mbarrier transaction counts), issued by one elected thread NS-1 iterations ahead.
// Each warp owns a 32 marker x 32 trait block: FP64 tensor-core mma.sync m8n8k4 (DMMA.8x8x4)
// over the K-chunked operands, then the per-k epilogue in registers (v = e - d^2*et, running
// min, tmax! counter), and ONE logarithm per output after the last k.  LOD / h2 panels are
// written once with streaming stores; nothing per-grid-point touches HBM.
#include <math.h>

#include "blmm_kernels.cuh"

namespace blmm {

namespace {

constexpr int TT = SCAN_TT;  // traits per CTA tile
constexpr int MT = SCAN_MT;  // markers per CTA tile
constexpr int SMEM_LIMIT = 227 * 1024;
#ifndef BLMM_SCAN_BT
#define BLMM_SCAN_BT 2
#endif
constexpr int GRID_MAX = 256;

struct SmemPlan {
  int nstage;
  size_t top_doubles;    // nq*TT*KC
  size_t stage_doubles;  // nq*MT*KC + 2*TT
  size_t bytes;
};

constexpr size_t FIXED_SMEM = (size_t)LOGTAB_N * 16 + GRID_MAX * 8 + 128;

__host__ __device__ inline SmemPlan plan_smem(int nq) {
  SmemPlan s;
  s.top_doubles = (size_t)nq * TT * KC;
  s.stage_doubles = (size_t)nq * MT * KC + 2 * TT;
  s.nstage = 3;
  s.bytes = (s.top_doubles + 3 * s.stage_doubles) * 8 + FIXED_SMEM;
  if (s.bytes > SMEM_LIMIT) {
    s.nstage = 2;
    s.bytes = (s.top_doubles + 2 * s.stage_doubles) * 8 + FIXED_SMEM;
  }
  return s;
}

// first k-step of a tile: C = 0 (no separate zeroing of the accumulators)
__device__ __forceinline__ void dmma884_zero(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%4};"
      : "=d"(c0), "=d"(c1)
      : "d"(a), "d"(b), "d"(0.0));
}

// BT = 8-trait atoms per warp: BT = 4 -> 8 warps of 32 markers x 32 traits, BT = 2 -> 16 warps of
// 32 markers x 16 traits (4 warps per SM sub-partition, <= 128 registers per thread).
template <int NQ, bool ARGMAX, int BT, bool HAS_E>
__global__ void __launch_bounds__(32 * 2 * (TT / (8 * BT)), 1) scan_kernel(const ScanParams P) {
  constexpr int WT_WARPS = TT / (8 * BT);  // warps along the trait dimension
  constexpr int NWARPS = 2 * WT_WARPS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const SmemPlan plan = plan_smem(NQ);
  const int NS = plan.nstage;
  double* top = reinterpret_cast<double*>(smem_raw);
  double* stages = top + plan.top_doubles;
  double2* logtab = reinterpret_cast<double2*>(stages + NS * plan.stage_doubles);
  double* grid_s = reinterpret_cast<double*>(logtab + LOGTAB_N);
  uint64_t* bars = reinterpret_cast<uint64_t*>(grid_s + GRID_MAX);
  uint64_t* full = bars;  // [NS]
  uint64_t* top_full = bars + 3;
  uint64_t* top_empty = bars + 4;
  int* rel_cnt = reinterpret_cast<int*>(bars + 5);  // [NS] consumers done with a stage

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  // The marker operand (all k) is re-read by every trait tile: ask L2 to keep it while the LOD /
  // h2 panels stream through (they are written with evict-first stores).
  const uint64_t keep_policy = l2_evict_last_policy();
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      rel_cnt[s] = 0;
    }
    mbar_init(top_full, 1);
    mbar_init(top_empty, NWARPS);
    mbar_fence_init();
  }
  if (tid < LOGTAB_N) logtab[tid] = reinterpret_cast<const double2*>(P.logtab)[tid];
  if (P.grid && tid < P.ngrid) grid_s[tid] = P.grid[tid];
  __syncthreads();

  const int n_tiles = P.n_tiles_dev ? *P.n_tiles_dev : P.n_tiles_t;
  const int n_mt = P.p_pad / MT;
  const int64_t units = (int64_t)n_tiles * n_mt;  // < 2^31 (checked by the launcher)
  const int u0 = (int)(units * blockIdx.x / gridDim.x);
  const int u1 = (int)(units * (blockIdx.x + 1) / gridDim.x);
  const int nk = P.nk;
  const int total_it = (u1 - u0) * nk;
  constexpr uint32_t marker_chunk_bytes = MT * KC * 8;
  constexpr uint32_t trait_chunk_bytes = TT * KC * 8;
  constexpr uint32_t stage_bytes = (uint32_t)NQ * marker_chunk_bytes + TT * 8 + (HAS_E ? TT * 8 : 0);

  // Fill stage s with the operands of iteration (tt, mt, kk): the k-th marker tile and that k's
  // per-trait scalars.  Called by one thread; completion is counted on full[s].
  auto issue_stage = [&](int s, int tt, int mt, int kk) {
    const int k0 = P.tile_k0 ? P.tile_k0[tt] : 0;
    double* st = stages + (size_t)s * plan.stage_doubles;
    mbar_arrive_expect_tx(&full[s], stage_bytes);
    const double* src = P.Mop + (((size_t)(k0 + kk) * NQ) * P.p_pad + (size_t)mt * MT) * KC;
#pragma unroll
    for (int q = 0; q < NQ; ++q)
      bulk_g2s_hint(st + (size_t)q * MT * KC, src + (size_t)q * P.p_pad * KC, marker_chunk_bytes, &full[s], keep_policy);
    double* sc = st + (size_t)NQ * MT * KC;
    bulk_g2s(sc + TT, P.et + (size_t)kk * P.tcol_pad + (size_t)tt * TT, TT * 8, &full[s]);
    if (HAS_E) bulk_g2s(sc, P.e + (size_t)kk * P.tcol_pad + (size_t)tt * TT, TT * 8, &full[s]);
  };
  // (tt, mt, kk) advanced by `steps` iterations
  auto advance = [&](int& tt, int& mt, int& kk, int steps) {
    kk += steps;
    while (kk >= nk) {
      kk -= nk;
      if (++mt == n_mt) {
        mt = 0;
        ++tt;
      }
    }
  };

  int tt = u0 / n_mt, mt = u0 % n_mt;
  if (tid == 0) {
    int ptt = tt, pmt = mt, pkk = 0;
    for (int i = 0; i < NS && i < total_it; ++i) {
      issue_stage(i, ptt, pmt, pkk);
      advance(ptt, pmt, pkk, 1);
    }
  }

  const int g = lane >> 2, t = lane & 3;
  const int wm = warp / WT_WARPS;  // marker sub-block (0..1)
  const int wt = warp % WT_WARPS;  // trait sub-block
  const int aoff = (wm * 32 + g) * KC + t;
  const int boff = (wt * (8 * BT) + g) * KC + t;

  int it = 0, s = 0;
  uint32_t sphase = 0;  // parity of the ring round
  int ntop = 0, cur_tt = -1;
  for (int u = u0; u < u1; ++u) {
    if (tt != cur_tt) {
      if (tid == 0) {
        if (ntop > 0) mbar_wait(top_empty, (ntop - 1) & 1);
        mbar_arrive_expect_tx(top_full, (uint32_t)NQ * trait_chunk_bytes);
#pragma unroll
        for (int q = 0; q < NQ; ++q)
          bulk_g2s(top + (size_t)q * TT * KC, P.Top + ((size_t)q * P.tcol_pad + (size_t)tt * TT) * KC,
                   trait_chunk_bytes, top_full);
      }
      mbar_wait(top_full, ntop & 1);
      ++ntop;
      cur_tt = tt;
    }
    const bool last_of_tt = (u + 1 == u1) || (mt + 1 == n_mt);

    double acc[4][BT][2];
    double vmin[4][BT][2];
    uint32_t cnt[2 * BT];
#pragma unroll
    for (int i = 0; i < 2 * BT; ++i) cnt[i] = 0u;

    for (int kk = 0; kk < nk; ++kk, ++it) {
      mbar_wait(&full[s], sphase);
      const double* ms = stages + (size_t)s * plan.stage_doubles;
      {
        // K loop, fully unrolled, fragments double-buffered in registers
        const double* ap = ms + aoff;
        const double* bp = top + boff;
        double af[2][4], bf[2][BT];
#pragma unroll
        for (int a = 0; a < 4; ++a) af[0][a] = ap[a * 8 * KC];
#pragma unroll
        for (int b = 0; b < BT; ++b) bf[0][b] = bp[b * 8 * KC];
#pragma unroll
        for (int st = 0; st < NQ * (KC / 4); ++st) {
          const int cur = st & 1;
          if (st + 1 < NQ * (KC / 4)) {
            const int q1 = (st + 1) / (KC / 4), s1 = (st + 1) % (KC / 4);
#pragma unroll
            for (int a = 0; a < 4; ++a) af[cur ^ 1][a] = ap[q1 * (MT * KC) + a * 8 * KC + s1 * 4];
#pragma unroll
            for (int b = 0; b < BT; ++b) bf[cur ^ 1][b] = bp[q1 * (TT * KC) + b * 8 * KC + s1 * 4];
          }
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < BT; ++b) {
              if (st == 0)
                dmma884_zero(acc[a][b][0], acc[a][b][1], af[cur][a], bf[cur][b]);
              else
                dmma884(acc[a][b][0], acc[a][b][1], af[cur][a], bf[cur][b]);
            }
        }
      }
      // per-k trait scalars for this lane's 2*BT trait columns
      double ek[BT][2], etk[BT][2];
      {
        const double* sc = ms + (size_t)NQ * MT * KC + wt * (8 * BT) + 2 * t;
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          if (HAS_E) {
            const double2 v = *reinterpret_cast<const double2*>(sc + b * 8);
            ek[b][0] = v.x; ek[b][1] = v.y;
          } else {
            ek[b][0] = ek[b][1] = 1.0;
          }
          const double2 w = *reinterpret_cast<const double2*>(sc + TT + b * 8);
          etk[b][0] = w.x; etk[b][1] = w.y;
        }
      }
      // Release the stage.  The last of the NWARPS consumers refills it at once with the operands
      // of iteration it + NS, so the copy is in flight as early as the ring allows.
      __syncwarp();
      if (lane == 0) {
        if (last_of_tt && kk == nk - 1) mbar_arrive(top_empty);
        __threadfence_block();
        if (atomicAdd(&rel_cnt[s], 1) == NWARPS - 1) {
          rel_cnt[s] = 0;
          if (it + NS < total_it) {
            int ntt = tt, nmt = mt, nkk = kk;
            advance(ntt, nmt, nkk, NS);
            issue_stage(s, ntt, nmt, nkk);
          }
        }
      }
      if (++s == NS) {
        s = 0;
        sphase ^= 1u;
      }

      const bool first = (kk == 0);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < BT; ++b)
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const double d = acc[a][b][cc];
            const double v = fma(-(d * d), etk[b][cc], ek[b][cc]);
            // strict `<` as `max .< to_compare` in tmax!, on the bit patterns (integer pipe; the
            // FP64 pipe is the bottleneck).  Equivalent for the non-negative v that occur; a
            // negative v (r^2 > 1 by rounding) orders below every positive one, as it should.
            const bool better = __double_as_longlong(v) < __double_as_longlong(vmin[a][b][cc]);
            const bool upd = first || better;
            vmin[a][b][cc] = upd ? v : vmin[a][b][cc];
            const int o = (a * BT + b) * 2 + cc;
            const int sh = (o & 3) * 8;
            if (ARGMAX) {
              if (upd) cnt[o >> 2] = (cnt[o >> 2] & ~(0xFFu << sh)) | ((uint32_t)kk << sh);
            } else {
              if (better && !first) cnt[o >> 2] += (1u << sh);
            }
          }
    }

    // final epilogue: one logarithm per output, streaming stores
    bool special = false;
    double lod[4][BT][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < BT; ++b)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) lod[a][b][cc] = fast_log10(vmin[a][b][cc], logtab, special);
    if (__any_sync(0xffffffffu, special)) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < BT; ++b)
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) lod[a][b][cc] = fix_log10(vmin[a][b][cc], lod[a][b][cc]);
    }
    const int k0 = P.tile_k0 ? P.tile_k0[tt] : 0;
    const int kbase = ARGMAX ? k0 : 0;
    const int i_base = mt * MT + wm * 32 + g;
#pragma unroll
    for (int b = 0; b < BT; ++b)
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int64_t pos = (int64_t)tt * TT + wt * (8 * BT) + b * 8 + 2 * t + cc;
        const int64_t col = P.col_map ? (int64_t)P.col_map[pos] : (pos < P.m ? pos : -1);
        // output column pointers (column 0 may be split off, see ScanParams::L0)
        double* Lc = nullptr;
        double* Hc = nullptr;
        double* Mc = nullptr;
        if (col >= 0) {
          if (P.L0) {
            if (col == 0) {
              Lc = P.L0;
            } else {
              if (P.L) Lc = P.L + (col - 1) * P.ldL;
              if (P.colmax) Mc = P.colmax + (col - 1);
            }
          } else {
            if (P.L) Lc = P.L + col * P.ldL;
            if (P.colmax) Mc = P.colmax + col;
          }
          if (P.H2) Hc = P.H2 + col * P.ldL;
        }
        double cmax = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int i = i_base + a * 8;
          const double l = -P.half_n * lod[a][b][cc];
          if (i < P.p) {
            if (Lc) st_stream(Lc + i, l);
            if (Hc) {
              const int o = (a * BT + b) * 2 + cc;
              const int cv = (int)((cnt[o >> 2] >> ((o & 3) * 8)) & 0xFFu);
              st_stream(Hc + i, grid_s[kbase + cv]);
            }
            cmax = fmax(cmax, l);
          }
        }
        if (P.colmax) {
          cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, 4));
          cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, 8));
          cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, 16));
          if (g == 0 && Mc) atomic_max_nonneg(Mc, cmax + 0.0);
        }
      }
    if (++mt == n_mt) {
      mt = 0;
      ++tt;
    }
  }
}

constexpr int SCAN_BT = BLMM_SCAN_BT;

template <int NQ, bool ARGMAX, bool HAS_E>
void launch_one(const ScanParams& P, int sm_count, cudaStream_t stream) {
  const SmemPlan plan = plan_smem(NQ);
  constexpr int threads = 32 * 2 * (TT / (8 * SCAN_BT));
  cudaFuncSetAttribute(scan_kernel<NQ, ARGMAX, SCAN_BT, HAS_E>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
  scan_kernel<NQ, ARGMAX, SCAN_BT, HAS_E><<<sm_count, threads, plan.bytes, stream>>>(P);
}

template <int NQ>
void launch_nq(const ScanParams& P, int sm_count, cudaStream_t stream) {
  if (P.e) {
    if (P.argmax_mode)
      launch_one<NQ, true, true>(P, sm_count, stream);
    else
      launch_one<NQ, false, true>(P, sm_count, stream);
  } else {
    // one-element k-lists (null-grid bins, permutations): the h2 panel is not produced
    launch_one<NQ, false, false>(P, sm_count, stream);
  }
}

}  // namespace

int launch_scan_v3(const ScanParams& P, int sm_count, cudaStream_t stream) {
  switch (P.nq) {
    case 1: launch_nq<1>(P, sm_count, stream); break;
    case 2: launch_nq<2>(P, sm_count, stream); break;
    case 3: launch_nq<3>(P, sm_count, stream); break;
    case 4: launch_nq<4>(P, sm_count, stream); break;
    case 5: launch_nq<5>(P, sm_count, stream); break;
    default: return 0;
  }
  return 1;
}

}  // namespace blmm
