// The fused marker x trait scan kernel (interface).  See blmm_scan_kernel.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace blmm {

// One launch computes, for every (marker i, trait column s) of the packed operands,
//     v_k   = e[k][s] - (sum_l Mop[k][i][l] * Top[s][l])^2 * tinv[k][s]        k in the tile's k-list
//     L     = -(n/2) * log10( min_k v_k )
//     h2    = grid[ #strict improvements of the running min ]   (tmax! counter semantics)  or
//             grid[ argmin_k v_k ]                               (argmax mode)
// which is bulkscan_alt_grid (src/bulkscan.jl:445-526 + tmax! src/bulkscan_helpers.jl:330-350) in
// its single-logarithm form, and with a one-element k-list and e = 1 is weighted_liteqtl + r2lod
// (src/bulkscan_helpers.jl:175-201, 22-24) for null-grid bins, null-exact and permutations.
struct ScanParams {
  const double* Top;      // trait operand   [nq][tcol_pad][KC]
  const double* Mop;      // marker operand  [nk_total][nq][p_pad][KC]
  const double* e;        // [nk_tile][tcol_pad] or nullptr (=> 1)
  const double* tinv;     // [nk_tile][tcol_pad] or nullptr (=> 1)
  const int* tile_k0;     // per trait tile: first k (index into Mop) ; nullptr => 0
  const int* tile_nk;     // per trait tile: number of k              ; nullptr => nk_all
  const int* n_tiles_dev; // device scalar: number of trait tiles in use; nullptr => n_tiles_t
  const int* col_map;     // [tcol_pad] packed column -> output column (or -1 = padding); nullptr => identity (< m)
  const double* grid;     // device copy of the h2 grid (for the h2 panel)
  double* L;              // p x m output, ld = ldL (nullptr => not stored)
  double* H2;             // p x m h2 panel, ld = ldL (nullptr => not stored)
  double* colmax;         // [m] running max over markers per output column (nullptr => off); caller zero-fills
  int64_t ldL;
  int nq;                 // K-chunks (ceil(n / KC))
  int p;                  // markers
  int p_pad;              // padded marker count (multiple of the marker tile)
  int m;                  // output columns
  int tcol_pad;           // padded packed-trait count (multiple of the trait tile)
  int n_tiles_t;          // trait tiles (upper bound when n_tiles_dev is given)
  int nk_all;             // k-list length when tile_nk == nullptr
  int argmax_mode;        // 0 = tmax! counter semantics, 1 = arg-max index
  double half_n;          // n / 2
};

// Tile configuration chosen from n; returns the trait-tile size (256/128/64) that keeps the whole-K
// trait tile resident in shared memory, or 0 if n is too large for the resident kernel.
int scan_trait_tile(int64_t n);
int scan_marker_tile(int trait_tile);

// Launch on `stream`; sm_count CTAs (persistent).  Returns cudaGetLastError().
cudaError_t launch_scan(const ScanParams& P, int trait_tile, int sm_count, cudaStream_t stream);

}  // namespace blmm
