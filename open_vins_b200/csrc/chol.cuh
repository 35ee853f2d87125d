// chol.cuh — CTA-cooperative blocked Cholesky used by the chi² gate (k_feature.cu) and the EKF update (k_ekf.cu).
#pragma once
#include <math.h>

// ---- blocked in-place Cholesky of the n x n matrix at S (lower triangle used), with `extra` right-hand-side rows
// stored as rows n..n+extra-1 (they receive rhs * L^-T, i.e. (L^-1 rhs')'). Returns false when a pivot is not positive.
template <int THREADS>
__device__ bool chol_lower_block(double *S, int ld, int n, int extra, int *flag) {
  const int NWARPS = THREADS / 32;
  const int NBK = 8;
  int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int kb = 0; kb < n; kb += NBK) {
    int nbk = min(NBK, n - kb);
    if (wid == 0) {
      for (int j = 0; j < nbk; j++) {
        double d = S[(kb + j) * ld + kb + j];
        if (!(d > 0.0)) {
          if (lane == 0)
            *flag = 1;
          d = 1.0; // keep going with a harmless value; the caller discards the result
        }
        d = sqrt(d);
        __syncwarp();
        if (lane == 0)
          S[(kb + j) * ld + kb + j] = d;
        if (lane > j && lane < nbk)
          S[(kb + lane) * ld + kb + j] /= d;
        __syncwarp();
        // trailing update inside the diagonal block
        for (int e = lane; e < nbk * nbk; e += 32) {
          int i = e / nbk, c = e % nbk;
          if (c > j && i >= c)
            S[(kb + i) * ld + kb + c] -= S[(kb + i) * ld + kb + j] * S[(kb + c) * ld + kb + j];
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // panel rows below: x L_kk' = S[i][kb..kb+nbk)
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
          x[c] = v / S[(kb + c) * ld + kb + c];
        }
      }
#pragma unroll
      for (int c = 0; c < NBK; c++)
        if (c < nbk)
          S[i * ld + kb + c] = x[c];
    }
    __syncthreads();
    // trailing update: S[i][j] -= sum_t S[i][kb+t] S[j][kb+t], kb+nbk <= j <= min(i, n-1)
    int first = kb + nbk;
    for (int i = first + wid; i < n + extra; i += NWARPS) {
      double li[NBK];
#pragma unroll
      for (int t = 0; t < NBK; t++)
        li[t] = (t < nbk) ? S[i * ld + kb + t] : 0.0;
      int jmax = min(i, n - 1);
      for (int j = first + lane; j <= jmax; j += 32) {
        double acc = 0.0;
#pragma unroll
        for (int t = 0; t < NBK; t++)
          if (t < nbk)
            acc += li[t] * S[j * ld + kb + t];
        S[i * ld + j] -= acc;
      }
    }
    __syncthreads();
  }
  return *flag == 0;
}

