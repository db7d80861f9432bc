// The fused marker x trait scan kernel (see ScanParams in blmm_kernels.cuh for the arithmetic).
//
// Structure (sm_100a): persistent CTAs, one per SM.  A CTA owns a contiguous range of
// (trait tile, marker tile) units.  The trait tile (128 traits x whole K) stays resident in shared
// memory; for every unit the k-list's marker tiles (64 markers x whole K, plus that k's per-trait
// scalars e/et) stream through a 2-3 stage ring filled by one producer warp with 1-D bulk
// asynchronous copies (cp.async.bulk -> UBLKCP, the TMA engine; mbarrier transaction counts).
// Eight consumer warps each own a 32 marker x 32 trait block: FP64 tensor-core mma.sync m8n8k4
// (DMMA.8x8x4) over the K-chunked operands, then the per-k epilogue in registers
// (v = e - d^2*et, running min, tmax! counter), and ONE log10 per output after the last k.
// LOD / h2 panels are written once with streaming stores; nothing per-grid-point touches HBM.
#include <math.h>

#include "blmm_kernels.cuh"

namespace blmm {

namespace {

constexpr int TT = SCAN_TT;  // traits per CTA tile
constexpr int MT = SCAN_MT;  // markers per CTA tile
constexpr int CONSUMER_WARPS = 8;
constexpr int SCAN_THREADS = 32 * (CONSUMER_WARPS + 1);
constexpr int SMEM_LIMIT = 227 * 1024;

struct SmemPlan {
  int nstage;
  size_t top_doubles;    // nq*TT*KC
  size_t stage_doubles;  // nq*MT*KC + 2*TT
  size_t bytes;
};

__host__ __device__ inline SmemPlan plan_smem(int nq) {
  SmemPlan s;
  s.top_doubles = (size_t)nq * TT * KC;
  s.stage_doubles = (size_t)nq * MT * KC + 2 * TT;
  s.nstage = 3;
  s.bytes = (s.top_doubles + 3 * s.stage_doubles) * 8 + 64;
  if (s.bytes > SMEM_LIMIT) {
    s.nstage = 2;
    s.bytes = (s.top_doubles + 2 * s.stage_doubles) * 8 + 64;
  }
  return s;
}

__global__ void __launch_bounds__(SCAN_THREADS, 1) scan_kernel(const ScanParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nq = P.nq;
  const SmemPlan plan = plan_smem(nq);
  const int NS = plan.nstage;
  double* top = reinterpret_cast<double*>(smem_raw);
  double* stages = top + plan.top_doubles;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + NS * plan.stage_doubles);
  uint64_t* full = bars;           // [NS]
  uint64_t* empty = bars + 3;      // [NS]
  uint64_t* top_full = bars + 6;
  uint64_t* top_empty = bars + 7;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CONSUMER_WARPS);
    }
    mbar_init(top_full, 1);
    mbar_init(top_empty, CONSUMER_WARPS);
    mbar_fence_init();
  }
  __syncthreads();

  const int n_tiles = P.n_tiles_dev ? *P.n_tiles_dev : P.n_tiles_t;
  const int n_mt = P.p_pad / MT;
  const int64_t units = (int64_t)n_tiles * n_mt;
  const int64_t u0 = units * blockIdx.x / gridDim.x;
  const int64_t u1 = units * (blockIdx.x + 1) / gridDim.x;
  const int nk = P.nk;
  const uint32_t marker_chunk_bytes = MT * KC * 8;
  const uint32_t trait_chunk_bytes = TT * KC * 8;

  if (warp == CONSUMER_WARPS) {
    // ------------------------------- producer ------------------------------------------------
    if (lane == 0) {
      int64_t it = 0;
      int ntop = 0, cur_tt = -1;
      for (int64_t u = u0; u < u1; ++u) {
        const int tt = (int)(u / n_mt), mt = (int)(u % n_mt);
        if (tt != cur_tt) {
          if (ntop > 0) mbar_wait(top_empty, (ntop - 1) & 1);
          mbar_arrive_expect_tx(top_full, (uint32_t)nq * trait_chunk_bytes);
          for (int q = 0; q < nq; ++q)
            bulk_g2s(top + (size_t)q * TT * KC, P.Top + ((size_t)q * P.tcol_pad + (size_t)tt * TT) * KC,
                     trait_chunk_bytes, top_full);
          ++ntop;
          cur_tt = tt;
        }
        const int k0 = P.tile_k0 ? P.tile_k0[tt] : 0;
        for (int kk = 0; kk < nk; ++kk, ++it) {
          const int s = (int)(it % NS);
          if (it >= NS) mbar_wait(&empty[s], (uint32_t)((it / NS) - 1) & 1);
          double* st = stages + (size_t)s * plan.stage_doubles;
          uint32_t bytes = (uint32_t)nq * marker_chunk_bytes;
          if (P.e) bytes += TT * 8;
          if (P.et) bytes += TT * 8;
          mbar_arrive_expect_tx(&full[s], bytes);
          const double* src = P.Mop + (((size_t)(k0 + kk) * nq) * P.p_pad + (size_t)mt * MT) * KC;
          for (int q = 0; q < nq; ++q)
            bulk_g2s(st + (size_t)q * MT * KC, src + (size_t)q * P.p_pad * KC, marker_chunk_bytes, &full[s]);
          double* sc = st + (size_t)nq * MT * KC;
          if (P.e) bulk_g2s(sc, P.e + (size_t)kk * P.tcol_pad + (size_t)tt * TT, TT * 8, &full[s]);
          if (P.et) bulk_g2s(sc + TT, P.et + (size_t)kk * P.tcol_pad + (size_t)tt * TT, TT * 8, &full[s]);
        }
      }
    }
    return;
  }

  // --------------------------------- consumers -------------------------------------------------
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2;  // marker sub-block (0..1)
  const int wt = warp & 3;   // trait sub-block (0..3)
  const int aoff = (wm * 32 + g) * KC + t;
  const int boff = (wt * 32 + g) * KC + t;
  const bool has_e = P.e != nullptr, has_et = P.et != nullptr;

  int64_t it = 0;
  int ntop = 0, cur_tt = -1;
  for (int64_t u = u0; u < u1; ++u) {
    const int tt = (int)(u / n_mt), mt = (int)(u % n_mt);
    if (tt != cur_tt) {
      mbar_wait(top_full, ntop & 1);
      ++ntop;
      cur_tt = tt;
    }
    const bool last_of_tt = (u + 1 == u1) || ((int)((u + 1) / n_mt) != tt);

    double acc[4][4][2];
    double vmin[4][4][2];
    uint32_t cnt[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cnt[i] = 0u;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        acc[a][b][0] = acc[a][b][1] = 0.0;
        vmin[a][b][0] = vmin[a][b][1] = 0.0;
      }

    for (int kk = 0; kk < nk; ++kk, ++it) {
      const int s = (int)(it % NS);
      mbar_wait(&full[s], (uint32_t)(it / NS) & 1);
      const double* ms = stages + (size_t)s * plan.stage_doubles;
      for (int q = 0; q < nq; ++q) {
        const double* ap = ms + q * (MT * KC) + aoff;
        const double* bp = top + q * (TT * KC) + boff;
#pragma unroll
        for (int st = 0; st < KC / 4; ++st) {
          double af[4], bf[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) af[a] = ap[a * 8 * KC + st * 4];
#pragma unroll
          for (int b = 0; b < 4; ++b) bf[b] = bp[b * 8 * KC + st * 4];
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
      }
      // per-k trait scalars for this lane's 8 trait columns
      double ek[4][2], etk[4][2];
      {
        const double* sc = ms + (size_t)nq * MT * KC + wt * 32 + 2 * t;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (has_e) {
            const double2 v = *reinterpret_cast<const double2*>(sc + b * 8);
            ek[b][0] = v.x; ek[b][1] = v.y;
          } else {
            ek[b][0] = ek[b][1] = 1.0;
          }
          if (has_et) {
            const double2 v = *reinterpret_cast<const double2*>(sc + TT + b * 8);
            etk[b][0] = v.x; etk[b][1] = v.y;
          } else {
            etk[b][0] = etk[b][1] = 1.0;
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&empty[s]);
        if (last_of_tt && kk == nk - 1) mbar_arrive(top_empty);
      }
      const bool first = (kk == 0);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const double d = acc[a][b][cc];
            const double v = fma(-(d * d), etk[b][cc], ek[b][cc]);
            acc[a][b][cc] = 0.0;
            const bool better = v < vmin[a][b][cc];  // strict, as `max .< to_compare` in tmax!
            const bool upd = first || better;
            vmin[a][b][cc] = upd ? v : vmin[a][b][cc];
            const int o = (a * 4 + b) * 2 + cc;
            const int sh = (o & 3) * 8;
            if (P.argmax_mode) {
              if (upd) cnt[o >> 2] = (cnt[o >> 2] & ~(0xFFu << sh)) | ((uint32_t)kk << sh);
            } else {
              if (better && !first) cnt[o >> 2] += (1u << sh);
            }
          }
    }

    // final epilogue: one logarithm per output, streaming stores
    const int k0 = P.tile_k0 ? P.tile_k0[tt] : 0;
    const int kbase = P.argmax_mode ? k0 : 0;
    const int i_base = mt * MT + wm * 32 + g;
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int64_t pos = (int64_t)tt * TT + wt * 32 + b * 8 + 2 * t + cc;
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
          const double lod = -P.half_n * log10(vmin[a][b][cc]);
          if (i < P.p) {
            if (Lc) st_stream(Lc + i, lod);
            if (Hc) {
              const int o = (a * 4 + b) * 2 + cc;
              const int cv = (int)((cnt[o >> 2] >> ((o & 3) * 8)) & 0xFFu);
              st_stream(Hc + i, P.grid[kbase + cv]);
            }
            cmax = fmax(cmax, lod);
          }
        }
        if (P.colmax) {
          cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, 4));
          cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, 8));
          cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, 16));
          if (g == 0 && Mc) atomic_max_nonneg(Mc, cmax + 0.0);
        }
      }
  }
}

}  // namespace

int scan_max_nq(int nk) {
  (void)nk;
  int nq = 1;
  while (plan_smem(nq + 1).bytes <= (size_t)SMEM_LIMIT) ++nq;
  return nq;
}

int launch_scan(const ScanParams& P, int sm_count, cudaStream_t stream) {
  const SmemPlan plan = plan_smem(P.nq);
  cudaFuncSetAttribute(scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
  scan_kernel<<<sm_count, SCAN_THREADS, plan.bytes, stream>>>(P);
  return 1;
}

}  // namespace blmm
