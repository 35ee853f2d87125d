// chol.cuh — CTA-cooperative blocked Cholesky used by the chi² gate (k_feature.cu), the EKF update (k_ekf.cu) and the
// normal-equations compression (k_gram.cu).
#pragma once
#include <math.h>

// Stage the lower triangle of rows [0, nrows) (columns j <= min(i, ncols-1)) of a row-major global matrix into shared
// memory with 8-byte cp.async: every load of the CTA is in flight at once instead of one L2 round trip per loop trip.
// Call from all threads; ends with wait + __syncthreads().
template <int THREADS>
__device__ __forceinline__ void stage_lower_async(double *dst, int ldd, const double *__restrict__ src, size_t lds, int nrows, int ncols) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = wid; i < nrows; i += THREADS / 32) {
    const int jmax = min(i, ncols - 1);
    for (int j = lane; j <= jmax; j += 32) {
      const unsigned d = (unsigned)__cvta_generic_to_shared(dst + (size_t)i * ldd + j);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src + (size_t)i * lds + j) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
}

#define EKC_THREADS 512 // threads of the single-CTA Cholesky kernels (16 warps: cheaper barriers, 128 registers per thread)

// 1/sqrt(d) from the hardware's double-precision seed (MUFU.RSQ64H, ~2^-20 relative, no float round trip) refined by
// one third-order (Halley) step to full double precision: ~80 cycles on the dependent chain instead of ~170 for the
// float-seeded variant and several hundred for the library sqrt/divide pair (tools/ubench/cholqr_bench.cu). Every
// Cholesky pivot sits on the critical path of its whole CTA.
__device__ __forceinline__ double fast_rsqrt(double d) {
  // branch free (a taken branch on the pivot chain costs more than the arithmetic): valid for normal positive d; the
  // callers select the result away for d <= 0, and pivots are >= sigma^2 or a shift, far from the subnormal range
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = 1.0 - d * y * y;
  return y + y * (e * (0.5 + 0.375 * e));
}

// 8x8 diagonal block of the blocked Cholesky, factored in registers by one warp: lane i (< nbk) holds row i and the
// pivots are broadcast by shuffle (a square-root-free LDL' pivot chain with the scalings batched at the end was
// measured and is NOT faster here: tools/ubench/chol_bench.cu). Writes L_kk back, the reciprocal pivots 1/L_jj to invd[0..8) (and inv_out[kb..]).
__device__ __forceinline__ void chol_diag8(double *S, int ld, int kb, int nbk, int *flag, double *invd, const double *diag0, double psd_tol,
                                           double *inv_out) {
  const int NBK = 8;
  const int lane = threadIdx.x & 31;
  double x[NBK];
  const int li = min(lane, nbk - 1);
#pragma unroll
  for (int c = 0; c < NBK; c++)
    x[c] = (c < nbk && c <= li) ? S[(kb + li) * ld + kb + c] : 0.0;
  {
#pragma unroll
    for (int j = 0; j < NBK; j++) {
      const double d = __shfl_sync(0xffffffffu, x[j], j);
      bool ok = false;
      if (j < nbk) {
        ok = d > 0.0;
        if (diag0 != nullptr) {
          ok = d > psd_tol * diag0[kb + j];
        } else if (!ok && lane == 0) {
          *flag = 1;
        }
      }
      const double inv = ok ? fast_rsqrt(d) : 0.0;
      if (lane == j)
        x[j] = d * inv;
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
          inv_out[kb + j] = inv;
      }
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
// it straight away, so the serial pivot chain hides behind the other warps' update. RB rows per warp are updated
// together (the panel rows L[j][kb..] are loaded once per RB rows and the RB FMA chains interleave).
template <int THREADS, int RB>
__device__ bool chol_lower_block(double *S, int ld, int n, int extra, int *flag, double *invd, const double *diag0 = nullptr,
                                 double psd_tol = 0.0, double *inv_out = nullptr) {
  const int NWARPS = THREADS / 32;
  const int NBK = 8;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (n <= 0)
    return *flag == 0;
  if (wid == 0)
    chol_diag8(S, ld, 0, min(NBK, n), flag, invd, diag0, psd_tol, inv_out);
  __syncthreads();
#ifdef CHOL_PROBE
  long long pt0 = 0, pt1 = 0, pt2 = 0, pt3 = 0, pt4 = 0;
#endif
  int par = 0; // invd is double-buffered: the look-ahead writes the next step's pivots while nothing reads this step's
  for (int kb = 0; kb < n; kb += NBK) {
    const int nbk = min(NBK, n - kb);
    const double *invk = invd + 8 * par;
#ifdef CHOL_PROBE
    pt0 = clock64();
#endif
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
#ifdef CHOL_PROBE
    pt1 = clock64();
#endif
    __syncthreads();
#ifdef CHOL_PROBE
    pt2 = clock64();
#endif
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
      chol_diag8(S, ld, first, nb2, flag, invd + 8 * (par ^ 1), diag0, psd_tol, inv_out);
    }
    {
      // remaining rows over warps 1.. (all warps when there is no look-ahead work or only one warp), RB rows at a time
      const bool la = (nb2 > 0) && (NWARPS > 1);
      const int w0 = la ? wid - 1 : wid, nw = la ? NWARPS - 1 : NWARPS;
      const int istart = first + max(nb2, 0);
      const int iend = n + extra;
      if (w0 >= 0) {
        for (int ib = istart + RB * w0; ib < iend; ib += RB * nw) {
          double li[RB][NBK];
#pragma unroll
          for (int r = 0; r < RB; r++)
#pragma unroll
            for (int t = 0; t < NBK; t++)
              li[r][t] = (t < nbk && ib + r < iend) ? S[(ib + r) * ld + kb + t] : 0.0;
          const int jtop = min(ib + RB - 1, n - 1);
          for (int j = first + lane; j <= jtop; j += 32) {
            double lj[NBK];
#pragma unroll
            for (int t = 0; t < NBK; t++)
              lj[t] = (t < nbk) ? S[j * ld + kb + t] : 0.0;
#pragma unroll
            for (int r = 0; r < RB; r++) {
              const int i = ib + r;
              if (i < iend && j <= min(i, n - 1)) {
                // four short chains per row (depth 4 with or without FMA contraction); the RB rows are independent
                double a0 = li[r][0] * lj[0], a1 = li[r][1] * lj[1], a2 = li[r][2] * lj[2], a3 = li[r][3] * lj[3];
                a0 += li[r][4] * lj[4];
                a1 += li[r][5] * lj[5];
                a2 += li[r][6] * lj[6];
                a3 += li[r][7] * lj[7];
                a0 += a2;
                a1 += a3;
                S[i * ld + j] -= a0 + a1;
              }
            }
          }
        }
      }
    }
#ifdef CHOL_PROBE
    pt3 = clock64();
#endif
    __syncthreads();
#ifdef CHOL_PROBE
    pt4 = clock64();
    if ((kb == 0 || kb == 40) && (tid == 0 || tid == 32 || tid == THREADS - 1))
      printf("kb=%d tid=%d panel %lld bar %lld trail/diag %lld bar %lld\n", kb, tid, pt1 - pt0, pt2 - pt1, pt3 - pt2, pt4 - pt3);
#endif
    par ^= 1;
  }
  return *flag == 0;
}
