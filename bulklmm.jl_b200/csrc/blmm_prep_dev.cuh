// Device helpers shared by blmm_prep.cu and the fused single-trait kernel of blmm_fit.cu (each translation unit gets
// its own copy: anonymous namespace).
#pragma once
#include <float.h>
#include <math.h>

#include "blmm_kernels.cuh"

namespace blmm {
namespace {

__device__ __forceinline__ double block_sum_128(double v, double* red) {
  // blockDim.x == 128; returns the sum to every thread
  v = warp_sum(v);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  return (red[0] + red[1]) + (red[2] + red[3]);
}

// One weight vector's constants, by a block of 128 threads (the body of weight_consts_kernel; also called by the fused
// single-trait kernel of blmm_fit.cu): w, sw, the CholQR^2 basis Q of sw .* C0, sum log w, log det (C0' W C0).
// Ends with a __syncthreads(): the outputs are visible to the whole block on return.
struct WcShared {
  double red[4];
  double Linv[MAXC][MAXC];
  double ldsum;
};
__device__ __forceinline__ void weight_consts_block(bool ols, double h2, const double* __restrict__ lambda,
                                                    const double* __restrict__ C0, int n, int n_pad, int c,
                                                    double* __restrict__ w, double* __restrict__ sw,
                                                    double* __restrict__ Q, double* slw_out, double* lds_out,
                                                    int* flags, WcShared& sh) {
  double* red = sh.red;
  double(*Linv)[MAXC] = sh.Linv;
  double& ldsum = sh.ldsum;
  const int tid = threadIdx.x;
  const double delta = h2 / (1.0 - h2);

  double slw = 0.0;
  for (int l = tid; l < n_pad; l += 128) {
    double wv = 0.0;
    if (l < n) {
      wv = ols ? 1.0 : 1.0 / (delta * lambda[l] + 1.0);
      if (!(wv > 0.0)) atomicExch(&flags[FLAG_WEIGHTS], 1);
      slw += log(wv);
    }
    const double s = sqrt(wv);
    w[l] = wv;
    sw[l] = s;
    for (int a = 0; a < c; ++a) Q[(int64_t)a * n_pad + l] = s * C0[(int64_t)a * n_pad + l];
  }
  slw = block_sum_128(slw, red);
  if (tid == 0) {
    *slw_out = slw;
    ldsum = 0.0;
  }
  // Cholesky-QR, applied twice: Q <- Q * inv(chol(Q'Q))'.  The second pass restores orthonormality
  // to rounding level; log det(C0' W C0) accumulates over the passes.
  for (int pass = 0; pass < 2; ++pass) {
    double S[MAXC][MAXC];
    for (int a = 0; a < c; ++a)
      for (int b = 0; b <= a; ++b) {
        double s = 0.0;
        for (int l = tid; l < n_pad; l += 128) s = fma(Q[(int64_t)a * n_pad + l], Q[(int64_t)b * n_pad + l], s);
        S[a][b] = block_sum_128(s, red);
      }
    if (tid == 0) {
      // in-place lower Cholesky, then its inverse
      bool ok = true;
      double Lc[MAXC][MAXC];
      for (int a = 0; a < c; ++a) {
        for (int b = 0; b <= a; ++b) {
          double s = S[a][b];
          for (int t = 0; t < b; ++t) s -= Lc[a][t] * Lc[b][t];
          if (a == b) {
            if (!(s > 0.0)) ok = false;
            Lc[a][a] = sqrt(s);
          } else {
            Lc[a][b] = s / Lc[b][b];
          }
        }
      }
      if (!ok) atomicExch(&flags[FLAG_NOT_SPD], 1);
      double ld = 0.0;
      for (int a = 0; a < c; ++a) ld += log(Lc[a][a]);
      ldsum += 2.0 * ld;
      for (int a = 0; a < c; ++a) {
        for (int b = 0; b < c; ++b) Linv[a][b] = 0.0;
        Linv[a][a] = 1.0 / Lc[a][a];
        for (int b = 0; b < a; ++b) {
          double s = 0.0;
          for (int t = b; t < a; ++t) s -= Lc[a][t] * Linv[t][b];
          Linv[a][b] = s / Lc[a][a];
        }
      }
    }
    __syncthreads();
    for (int l = tid; l < n_pad; l += 128) {
      for (int a = c - 1; a >= 0; --a) {
        double s = 0.0;
        for (int b = 0; b <= a; ++b) s = fma(Linv[a][b], Q[(int64_t)b * n_pad + l], s);
        Q[(int64_t)a * n_pad + l] = s;
      }
    }
    __syncthreads();
  }
  if (tid == 0) *lds_out = ldsum;
  __syncthreads();
}


// z = P_k (sw_k .* x): returns ||z||^2; the caller re-evaluates z_l through proj_elem.
__device__ __forceinline__ void proj_coefs(const double* __restrict__ xb, const double* __restrict__ sw,
                                           const double* __restrict__ Q, int n_pad, int c, int lane,
                                           double coef[MAXC]) {
#pragma unroll
  for (int a = 0; a < MAXC; ++a) {
    if (a < c) {
      double s = 0.0;
      for (int l = lane; l < n_pad; l += 32) s = fma(Q[(int64_t)a * n_pad + l], sw[l] * xb[l], s);
      coef[a] = warp_sum(s);
    }
  }
}
__device__ __forceinline__ double proj_elem(const double* __restrict__ xb, const double* __restrict__ sw,
                                            const double* __restrict__ Q, int n_pad, int c, int l,
                                            const double coef[MAXC]) {
  double z = sw[l] * xb[l];
#pragma unroll
  for (int a = 0; a < MAXC; ++a)
    if (a < c) z = fma(-Q[(int64_t)a * n_pad + l], coef[a], z);
  return z;
}

}  // namespace
}  // namespace blmm
