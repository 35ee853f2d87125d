// chol.cuh — CTA-cooperative blocked Cholesky used by the chi² gate (k_feature.cu), the EKF update (k_ekf.cu) and the
// normal-equations compression (k_gram.cu).

#include <math.h>

// 1/sqrt(d) from a float seed refined to full double precision; the library sqrt/divide pair costs several
// hundred cycles on a single dependent chain, and every Cholesky pivot sits on the critical path of the whole CTA.
__device__ __forceinline__ double fast_rsqrt_old(double d) {
  if (d > 1e-30 && d < 1e30) {
    // Halley step (cubic: 23 -> ~66 bits) followed by one Newton polish on the residual, two dependent rounds shorter
    // than three Newton steps
    double y = (double)rsqrtf((float)d);
    double e = 1.0 - d * y * y;
    y = y + y * e * (0.5 + 0.375 * e);
    e = 1.0 - d * y * y;
    return y + 0.5 * y * e;
  }
  return 1.0 / sqrt(d);
}

// 8x8 diagonal block of the blocked Cholesky, factored in registers by one warp: lane i (< nbk) holds row i and the
// pivots are broadcast by shuffle. Writes L_kk back, the reciprocal pivots to invd[0..8) (and inv_out[kb..]).
__device__ __forceinline__ void chol_diag8_old(double *S, int ld, int kb, int nbk, int *flag, double *invd, const double *diag0, double psd_tol,
                                           double *inv_out) {
  const int NBK = 8;
  const int lane = threadIdx.x & 31;
  double x[NBK];
  const int li = min(lane, nbk - 1);
#pragma unroll
  for (int c = 0; c < NBK; c++)
    x[c] = (c < nbk && c <= li) ? S[(kb + li) * ld + kb + c] : 0.0;
#pragma unroll
  for (int j = 0; j < NBK; j++) {
    const double d = __shfl_sync(0xffffffffu, x[j], j);
    double inv = 0.0, ljj = 0.0;
    if (j < nbk) {
      bool ok = d > 0.0;
      if (diag0 != nullptr) {
        ok = d > psd_tol * diag0[kb + j];
      } else if (!ok && lane == 0) {
        *flag = 1;
      }
      if (ok) {
        inv = fast_rsqrt_old(d);
        ljj = d * inv;
      }
    }
    if (lane == j)
      x[j] = ljj;
    else if (lane > j)
      x[j] *= inv;
#pragma unroll
    for (int c = 0; c < NBK; c++) {
      if (c > j) {
        const double lcj = __shfl_sync(0xffffffffu, x[j], c);
        if (lane >= c)
          x[c] -= x[j] * lcj;
      }
    }
    if (lane == 0) {
      invd[j] = inv;
      if (inv_out != nullptr && j < nbk)
        inv_out[kb + j] = inv; // reciprocal pivots for later triangular solves
    }
  }
  if (lane < nbk) {
#pragma unroll
    for (int c = 0; c < NBK; c++)
      if (c <= lane)
        S[(kb + lane) * ld + kb + c] = x[c];
  }
}

// Blocked in-place Cholesky of the n x n matrix at S (lower triangle used), with `extra` right-hand-side rows stored as
// rows n..n+extra-1 (they receive rhs * L^-T, i.e. (L^-1 rhs')').
//   diag0 == nullptr : strict mode, a pivot <= 0 (or NaN) raises *flag and the result must be discarded.
//   diag0 != nullptr : semidefinite mode, a pivot <= psd_tol * diag0[j] (roundoff-level: the direction carries no
//                      information) zeroes column j of L instead of failing.
// invd: 16 doubles of shared memory scratch. inv_out (optional, n doubles): receives 1/L[j][j]. Returns true when no
// strict-mode pivot failed.
// Look-ahead: during the trailing update of block step k, warp 0 updates only the next 8x8 diagonal block and factors
// it straight away, so the serial pivot chain (8 dependent rsqrt/shuffle rounds) hides behind the other warps' update.
template <int THREADS>
__device__ bool chol_lower_block_old(double *S, int ld, int n, int extra, int *flag, double *invd, const double *diag0 = nullptr,
                                 double psd_tol = 0.0, double *inv_out = nullptr) {
  const int NWARPS = THREADS / 32;
  const int NBK = 8;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (n <= 0)
    return *flag == 0;
  if (wid == 0)
    chol_diag8_old(S, ld, 0, min(NBK, n), flag, invd, diag0, psd_tol, inv_out);
  __syncthreads();
  int par = 0; // invd is double-buffered: the look-ahead writes the next step's pivots while nothing reads this step's
  for (int kb = 0; kb < n; kb += NBK) {
    const int nbk = min(NBK, n - kb);
    const double *invk = invd + 8 * par;
    // ---- panel rows below: x L_kk' = S[i][kb..kb+nbk), using the stored reciprocal pivots (no divisions)
    for (int i = kb + nbk + tid; i < n + extra; i += THREADS) {
      double x[NBK];
#pragma unroll
      for (int c = 0; c < NBK; c++) {
        if (c < nbk) {
          double v = S[i * ld + kb + c];
#pragma unroll
          for (int t = 0; t < NBK; t++)
            if (t < c)
              v -= x[t] * S[(kb + c) * ld + kb + t];
          x[c] = v * invk[c];
        }
      }
#pragma unroll
      for (int c = 0; c < NBK; c++)
        if (c < nbk)
          S[i * ld + kb + c] = x[c];
    }
    __syncthreads();
    // ---- trailing update: S[i][j] -= sum_t S[i][kb+t] S[j][kb+t], kb+nbk <= j <= min(i, n-1)
    const int first = kb + nbk;
    const int nb2 = min(NBK, n - first); // rows of the next diagonal block (<= 0 when this was the last step)
    if (wid == 0 && nb2 > 0) {
      for (int e = lane; e < NBK * NBK; e += 32) {
        const int r = e >> 3, c = e & 7;
        if (r < nb2 && c <= r) {
          const int i = first + r, j = first + c;
          double pr[NBK];
#pragma unroll
          for (int t = 0; t < NBK; t++)
            pr[t] = (t < nbk) ? S[i * ld + kb + t] * S[j * ld + kb + t] : 0.0;
          S[i * ld + j] -= ((pr[0] + pr[1]) + (pr[2] + pr[3])) + ((pr[4] + pr[5]) + (pr[6] + pr[7]));
        }
      }
      __syncwarp();
      chol_diag8_old(S, ld, first, nb2, flag, invd + 8 * (par ^ 1), diag0, psd_tol, inv_out);
    }
    {
      // remaining rows over warps 1.. (all warps when there is no look-ahead work or only one warp)
      const bool la = (nb2 > 0) && (NWARPS > 1);
      const int w0 = la ? wid - 1 : wid, nw = la ? NWARPS - 1 : NWARPS;
      const int istart = first + max(nb2, 0);
      if (w0 >= 0) {
        for (int i = istart + w0; i < n + extra; i += nw) {
          double li[NBK];
#pragma unroll
          for (int t = 0; t < NBK; t++)
            li[t] = (t < nbk) ? S[i * ld + kb + t] : 0.0;
          const int jmax = min(i, n - 1);
          for (int j = first + lane; j <= jmax; j += 32) {
            // eight products summed as a tree (3 dependent adds instead of an 8-long FMA chain: FP64 latency is ~19 cycles)
            double pr[NBK];
#pragma unroll
            for (int t = 0; t < NBK; t++)
              pr[t] = (t < nbk) ? li[t] * S[j * ld + kb + t] : 0.0;
            S[i * ld + j] -= ((pr[0] + pr[1]) + (pr[2] + pr[3])) + ((pr[4] + pr[5]) + (pr[6] + pr[7]));
          }
        }
      }
    }
    __syncthreads();
    par ^= 1;
  }
  return *flag == 0;
}
