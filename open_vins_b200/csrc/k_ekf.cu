// k_ekf.cu — dense covariance algebra on the device-resident P.
// Replaces StateHelper::EKFUpdate (ov_msckf/src/state/StateHelper.cpp:116-197), StateHelper::EKFPropagation (:36-114),
// StateHelper::clone / augment_clone's time-offset term (:341-391, :604-615), StateHelper::marginalize (:271-339) and
// get_marginal_covariance (:226-254).
//
// EKFUpdate on the device:   M = P[:,cols] H'   S = H M[cols,:] + R   S = L L'   Y = M L^-T   w = L^-1 res
//                            P <- sym_U(P - Y Y')      dx = Y w
// (K M' = M S^-1 M' = Y Y' and K res = Y w, so neither S^-1 nor K is formed; the reference forms both.)
// P is row-major with a fixed leading dimension so clone() grows it in place.
#include "chol.cuh"
#include "ovb_internal.cuh"

#define EK_T 32

// C[x][y] = sum_j Aop(x,j) * Bop(j,y), 32x32 tile per CTA, 256 threads (4 outputs each), K tiles of 32.
//  mode 0 (M = P[:,cols] H'):  Aop(a,j) = P[cs[j]*ldP + a]      Bop(j,i) = H[i*ldH + j]      C = M [N x r]
//  mode 1 (S = H M[cols,:]+R): Aop(i,j) = H[i*ldH + j]          Bop(j,k) = M[cs[j]*ldM + k]  C = S [r x r], lower tiles only
__global__ void __launch_bounds__(256) k_ekf_gemm(int mode, const double *__restrict__ P, int ldP, const double *__restrict__ H, int ldH,
                                                  const double *__restrict__ Min, int ldM, const DevUpdateInfo *__restrict__ info, int X, int Y,
                                                  int K, double *__restrict__ Cout, int ldC, double sigma2, const double *__restrict__ Rdiag) {
  OVB_PDL_ENTER();
  __shared__ double As[EK_T][EK_T + 1]; // [j][x]
  __shared__ double Bs[EK_T][EK_T + 1]; // [j][y]
  const int x0 = blockIdx.x * EK_T, y0 = blockIdx.y * EK_T;
  if (mode == 1 && y0 > x0 + EK_T - 1)
    return; // strictly upper tile of S
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5; // ty 0..7
  double acc[4] = {0, 0, 0, 0};
  const int *cs = info->col_state;
  for (int j0 = 0; j0 < K; j0 += EK_T) {
    // load tiles
    for (int e = tid; e < EK_T * EK_T; e += 256) {
      int jj = e >> 5, q = e & 31; // jj = k index in tile, q = x or y
      int j = j0 + jj;
      double av = 0.0, bv = 0.0;
      if (j < K) {
        if (mode == 0) {
          int a = x0 + q;
          if (a < X)
            av = P[(size_t)cs[j] * ldP + a];
        } else {
          // Aop(i,j) = H[i][j]: transpose load (lanes over q=i → strided); small matrices, acceptable
          int i = x0 + q;
          if (i < X)
            av = H[(size_t)i * ldH + j];
        }
        if (mode == 0) {
          int i = y0 + q;
          if (i < Y)
            bv = H[(size_t)i * ldH + j];
        } else {
          int k = y0 + q;
          if (k < Y)
            bv = Min[(size_t)cs[j] * ldM + k];
        }
      }
      As[jj][q] = av;
      Bs[jj][q] = bv;
    }
    __syncthreads();
#pragma unroll 8
    for (int jj = 0; jj < EK_T; jj++) {
      double b = Bs[jj][tx];
#pragma unroll
      for (int u = 0; u < 4; u++)
        acc[u] += As[jj][ty + 8 * u] * b;
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; u++) {
    int x = x0 + ty + 8 * u, y = y0 + tx;
    if (x < X && y < Y) {
      double v = acc[u];
      if (mode == 1 && x == y)
        v += Rdiag ? Rdiag[x] : sigma2;
      Cout[(size_t)x * ldC + y] = v;
    }
  }
}

// single CTA: S (r x r, lower) plus the residual as row r  ->  L and w = L^-1 res in row r. Works in shared memory when
// it fits, else in place in global memory (L2).
__global__ void __launch_bounds__(EKC_THREADS) k_ekf_chol(double *__restrict__ S, int ldS, int r, const double *__restrict__ res, double *__restrict__ w,
                                                   double *__restrict__ invdiag, DevUpdateInfo *__restrict__ info, int use_smem) {
  OVB_PDL_ENTER();
  extern __shared__ __align__(16) double chol_sm[];
  __shared__ int flag;
  __shared__ double invd_sh[16];
  const int tid = threadIdx.x;
  if (tid == 0)
    flag = 0;
  for (int j = tid; j < r; j += EKC_THREADS)
    S[(size_t)r * ldS + j] = res[j];
  __syncthreads();
  double *W = S;
  int ld = ldS;
  const int lane = tid & 31, wid = tid >> 5;
  if (use_smem) {
    ld = r | 1;
    W = chol_sm;
    stage_lower_async<EKC_THREADS>(W, ld, S, (size_t)ldS, r + 1, r);
  }
  chol_lower_block<EKC_THREADS, 4>(W, ld, r, 1, &flag, invd_sh, nullptr, 0.0, invdiag);
  __syncthreads();
  if (use_smem) {
    for (int i = wid; i < r; i += EKC_THREADS / 32) {
      for (int j = lane; j <= i; j += 32)
        S[(size_t)i * ldS + j] = W[i * ld + j];
    }
  }
  for (int j = tid; j < r; j += EKC_THREADS)
    w[j] = W[(size_t)r * ld + j];
  if (tid == 0 && flag)
    info->not_spd = 1;
}

// Y = M L^-T : one WARP per row of M (independent forward substitutions y = L^-1 m), blocked by 8 columns. L is staged
// once per CTA in shared memory (the substitution is a serial chain: an L2 round trip per step would dominate); the eight
// dot products against the solved part run as independent chains, one butterfly reduces them, then every lane solves
// the 8x8 triangle redundantly with the reciprocal pivots written by the Cholesky kernel.
#define TR_ROWS 8 // rows of M (warps) per CTA
__global__ void __launch_bounds__(32 * TR_ROWS) k_ekf_trsm(const double *__restrict__ M, int ldM, const double *__restrict__ L, int ldL,
                                                           const double *__restrict__ invdiag, int N, int r, double *__restrict__ Yout, int ldY,
                                                           int L_in_smem, const DevUpdateInfo *__restrict__ info) {
  OVB_PDL_ENTER();
  extern __shared__ __align__(16) double tsm[]; // [TR_ROWS][r] y rows, then r reciprocal pivots, then L (r x ldl) when it fits
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (info->not_spd)
    return; // failed factor: leave P untouched (the status code is then recoverable, include/ovb200.h)
  const int ldl = r | 1;
  double *inv_s = tsm + (size_t)TR_ROWS * r;
  double *Ls = inv_s + r;
  for (int e = threadIdx.x; e < r; e += 32 * TR_ROWS)
    inv_s[e] = invdiag[e];
  if (L_in_smem)
    stage_lower_async<32 * TR_ROWS>(Ls, ldl, L, (size_t)ldL, r, r);
  else
    __syncthreads();
  const double *Lp = L_in_smem ? Ls : L;
  const int lp = L_in_smem ? ldl : ldL;
  const int a = blockIdx.x * TR_ROWS + wid;
  if (a >= N)
    return;
  double *y = tsm + (size_t)wid * r;
  for (int t = lane; t < r; t += 32)
    y[t] = M[(size_t)a * ldM + t];
  __syncwarp();
  for (int kb = 0; kb < r; kb += 8) {
    const int nbk = min(8, r - kb);
    double s[8];
#pragma unroll
    for (int c = 0; c < 8; c++)
      s[c] = 0.0;
    for (int t = lane; t < kb; t += 32) {
      const double yt = y[t];
#pragma unroll
      for (int c = 0; c < 8; c++)
        if (c < nbk)
          s[c] += Lp[(size_t)(kb + c) * lp + t] * yt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < 8; c++)
        s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
    }
    double yn[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
      if (c < nbk) {
        const double *Lr = Lp + (size_t)(kb + c) * lp + kb;
        // everything that does not depend on yn[c-1] first: the serial chain is one FMA + one multiply per column
        double v = y[kb + c] - s[c];
#pragma unroll
        for (int t = 0; t < 8; t++)
          if (t + 1 < c)
            v -= Lr[t] * yn[t];
        if (c > 0)
          v -= Lr[c - 1] * yn[c - 1];
        yn[c] = v * inv_s[kb + c];
      }
    }
    __syncwarp();
    if (lane < nbk) {
      double v = yn[0];
#pragma unroll
      for (int c = 1; c < 8; c++)
        v = (lane == c) ? yn[c] : v;
      y[kb + lane] = v;
    }
    __syncwarp();
  }
  for (int t = lane; t < r; t += 32)
    Yout[(size_t)a * ldY + t] = y[t];
}

// P <- sym_U(P - Y Y'): upper tiles only, mirrored on write; negative-diagonal check; dx = Y w on the diagonal tiles' rows
__global__ void __launch_bounds__(256) k_ekf_downdate(double *__restrict__ P, int ldP, const double *__restrict__ Yin, int ldY, int N, int r,
                                                      const double *__restrict__ w, double *__restrict__ dx, DevUpdateInfo *__restrict__ info) {
  OVB_PDL_ENTER();
  __shared__ double As[EK_T][EK_T + 1]; // [k][a]
  __shared__ double Bs[EK_T][EK_T + 1]; // [k][b]
  const int a0 = blockIdx.x * EK_T, b0 = blockIdx.y * EK_T;
  if (b0 < a0)
    return;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  if (info->not_spd) { // failed factor: P stays as it was, no correction
    if (a0 == b0 && tid < EK_T && a0 + tid < N)
      dx[a0 + tid] = 0.0;
    return;
  }
  double acc[4] = {0, 0, 0, 0};
  double dxa = 0.0; // dx partial for row a0 + tid (diagonal tiles, tid < 32)
  for (int k0 = 0; k0 < r; k0 += EK_T) {
    for (int e = tid; e < EK_T * EK_T; e += 256) {
      int q = e >> 5, kk = e & 31; // lanes over k: coalesced rows of Y
      int k = k0 + kk;
      As[kk][q] = (k < r && a0 + q < N) ? Yin[(size_t)(a0 + q) * ldY + k] : 0.0;
      Bs[kk][q] = (k < r && b0 + q < N) ? Yin[(size_t)(b0 + q) * ldY + k] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < EK_T; kk++) {
      double b = Bs[kk][tx];
#pragma unroll
      for (int u = 0; u < 4; u++)
        acc[u] += As[kk][ty + 8 * u] * b;
    }
    if (a0 == b0 && tid < EK_T) {
      for (int kk = 0; kk < EK_T; kk++)
        if (k0 + kk < r)
          dxa += As[kk][tid] * w[k0 + kk];
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; u++) {
    int a = a0 + ty + 8 * u, b = b0 + tx;
    if (a < N && b < N && b >= a) {
      double v = P[(size_t)a * ldP + b] - acc[u];
      P[(size_t)a * ldP + b] = v;
      P[(size_t)b * ldP + a] = v;
      if (a == b && v < 0.0)
        atomicMin(&info->neg_diag_index, a);
      if (!isfinite(v))
        info->nonfinite = 1;
    }
  }
  if (a0 == b0 && tid < EK_T && a0 + tid < N)
    dx[a0 + tid] = dxa;
}

// ---- one-shot variants for K <= EK1_KMAX (the whole inner dimension staged at once: a CTA pays ONE L2 round trip instead
// of one per 32-wide K tile; at config-2 sizes these products are latency-, not flop-bound). Same contract as k_ekf_gemm.
#define EK1_KMAX 160
__device__ __forceinline__ void ek_cpa8(double *dst_smem, const double *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__global__ void __launch_bounds__(256) k_ekf_gemm1(int mode, const double *__restrict__ P, int ldP, const double *__restrict__ H, int ldH,
                                                   const double *__restrict__ Min, int ldM, const DevUpdateInfo *__restrict__ info, int X, int Y,
                                                   int K, double *__restrict__ Cout, int ldC, double sigma2, const double *__restrict__ Rdiag) {
  OVB_PDL_ENTER();
  extern __shared__ __align__(16) double g1[];
  const int x0 = blockIdx.x * EK_T, y0 = blockIdx.y * EK_T;
  if (mode == 1 && y0 > x0 + EK_T - 1)
    return; // strictly upper tile of S
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int *cs = info->col_state;
  const int pk = K | 1;
  // mode 0: Gs[j][a] = P[cs[j]][x0+a] (K x 32, pitch 33)   Hs[i][j] = H[y0+i][j] (32 x K, pitch pk)      M[a][i]
  // mode 1: Hs[i][j] = H[x0+i][j]      (32 x K, pitch pk)   Gs[j][k] = M[cs[j]][y0+k] (K x 32, pitch 33)   S[i][k]
  double *Gs = g1, *Hs = g1 + (size_t)K * 33;
  const double *Gsrc = (mode == 0) ? P : Min;
  const int ldG = (mode == 0) ? ldP : ldM;
  const int g0 = (mode == 0) ? x0 : y0, gmax = (mode == 0) ? X : Y;
  const int h0 = (mode == 0) ? y0 : x0, hmax = (mode == 0) ? Y : X;
  // the column map first (the gathered rows' addresses depend on it), then every element of both operands as an 8-byte
  // cp.async: one L2 round trip for the whole staging instead of a dependent index -> element pair per loop trip
  __shared__ int cs_s[EK1_KMAX];
  for (int j = tid; j < K; j += 256)
    cs_s[j] = cs[j];
  __syncthreads();
  for (int e = tid; e < K * 32; e += 256) {
    const int j = e >> 5, a = e & 31;
    if (g0 + a < gmax)
      ek_cpa8(&Gs[j * 33 + a], &Gsrc[(size_t)cs_s[j] * ldG + g0 + a]);
    else
      Gs[j * 33 + a] = 0.0;
  }
  for (int i = ty; i < 32; i += 8)
    for (int j = tx; j < K; j += 32) {
      if (h0 + i < hmax)
        ek_cpa8(&Hs[i * pk + j], &H[(size_t)(h0 + i) * ldH + j]);
      else
        Hs[i * pk + j] = 0.0;
    }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  double acc[4] = {0, 0, 0, 0};
  if (mode == 0) {
    // acc[u] = sum_j Gs[j][tx] * Hs[ty + 8u][j]   -> M[x0 + tx][y0 + ty + 8u]
    for (int j = 0; j < K; j++) {
      const double gv = Gs[j * 33 + tx];
#pragma unroll
      for (int u = 0; u < 4; u++)
        acc[u] += gv * Hs[(ty + 8 * u) * pk + j];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int a = x0 + tx, i = y0 + ty + 8 * u;
      if (a < X && i < Y)
        Cout[(size_t)a * ldC + i] = acc[u];
    }
  } else {
    // acc[u] = sum_j Hs[ty + 8u][j] * Gs[j][tx]   -> S[x0 + ty + 8u][y0 + tx]
    for (int j = 0; j < K; j++) {
      const double gv = Gs[j * 33 + tx];
#pragma unroll
      for (int u = 0; u < 4; u++)
        acc[u] += Hs[(ty + 8 * u) * pk + j] * gv;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = x0 + ty + 8 * u, k = y0 + tx;
      if (i < X && k < Y) {
        double v = acc[u];
        if (i == k)
          v += Rdiag ? Rdiag[i] : sigma2;
        Cout[(size_t)i * ldC + k] = v;
      }
    }
  }
}

// P <- sym_U(P - Y Y'), dx = Y w, r <= EK1_KMAX: both 32-row strips of Y staged at once (same contract as k_ekf_downdate)
__global__ void __launch_bounds__(256) k_ekf_downdate1(double *__restrict__ P, int ldP, const double *__restrict__ Yin, int ldY, int N, int r,
                                                       const double *__restrict__ w, double *__restrict__ dx, DevUpdateInfo *__restrict__ info) {
  OVB_PDL_ENTER();
  extern __shared__ __align__(16) double d1[];
  const int a0 = blockIdx.x * EK_T, b0 = blockIdx.y * EK_T;
  if (b0 < a0)
    return;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  if (info->not_spd) { // failed factor: P stays as it was, no correction
    if (a0 == b0 && tid < EK_T && a0 + tid < N)
      dx[a0 + tid] = 0.0;
    return;
  }
  const int pk = r | 1;
  double *As = d1, *Bs = d1 + (size_t)32 * pk, *ws = Bs + (size_t)32 * pk;
  for (int i = ty; i < 32; i += 8)
    for (int k = tx; k < r; k += 32) {
      if (a0 + i < N)
        ek_cpa8(&As[i * pk + k], &Yin[(size_t)(a0 + i) * ldY + k]);
      else
        As[i * pk + k] = 0.0;
      if (b0 + i < N)
        ek_cpa8(&Bs[i * pk + k], &Yin[(size_t)(b0 + i) * ldY + k]);
      else
        Bs[i * pk + k] = 0.0;
    }
  for (int k = tid; k < r; k += 256)
    ek_cpa8(&ws[k], &w[k]);
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  double acc[4] = {0, 0, 0, 0};
  for (int k = 0; k < r; k++) {
    const double bv = Bs[tx * pk + k];
#pragma unroll
    for (int u = 0; u < 4; u++)
      acc[u] += As[(ty + 8 * u) * pk + k] * bv;
  }
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const int a = a0 + ty + 8 * u, b = b0 + tx;
    if (a < N && b < N && b >= a) {
      const double v = P[(size_t)a * ldP + b] - acc[u];
      P[(size_t)a * ldP + b] = v;
      P[(size_t)b * ldP + a] = v;
      if (a == b && v < 0.0)
        atomicMin(&info->neg_diag_index, a);
      if (!isfinite(v))
        info->nonfinite = 1;
    }
  }
  if (a0 == b0 && tid < EK_T && a0 + tid < N) {
    double s0 = 0.0, s1 = 0.0;
    int k = 0;
    for (; k + 1 < r; k += 2) {
      s0 += As[tid * pk + k] * ws[k];
      s1 += As[tid * pk + k + 1] * ws[k + 1];
    }
    if (k < r)
      s0 += As[tid * pk + k] * ws[k];
    dx[a0 + tid] = s0 + s1;
  }
}

__global__ void k_ekf_prep(DevUpdateInfo *info) {
  OVB_PDL_ENTER();
  info->neg_diag_index = 0x7fffffff;
  info->not_spd = 0;
  info->nonfinite = 0;
}

// H: r x n (row-major, ldHm), r <= n <= N (callers compress first when r > n); column j of H is state column
// d_info->col_state[j]. Everything is enqueued on the context stream; flags land in d_info.
// gate_only: stop after the Cholesky — d_w then holds w = L^-1 res (|w|^2 = res' S^-1 res) and P is untouched
// (the Mahalanobis test of StateHelper::initialize, StateHelper.cpp:458-470).
void launch_ekf_update(ovb_ctx *ctx, const double *H, int ldHm, int r, int n, bool gate_only, double sigma2, const double *Rdiag_dev) {
  const int N = ctx->N;
  const int ld = ctx->ldP;
  double *P = ctx->P[ctx->cur];
  ovb_launch(ctx, k_ekf_prep, dim3(1), dim3(1), (size_t)(0), ctx->d_info);
  if (r <= 0 || n <= 0)
    return;
  if (!ctx->attr_done[2]) { // function attributes are per device: one flag per context
    cudaFuncSetAttribute(k_ekf_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_ekf_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_ekf_gemm1, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_ekf_downdate1, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    ctx->attr_done[2] = 1;
  }
  // small inner dimensions (config 2: n = r = 154): one-shot staged products, DMMA Cholesky + register solve of k_cholqr.cu
  const bool one_shot = ctx->ekf_chol_dmma && n <= EK1_KMAX && r <= EK1_KMAX;
  dim3 g0((N + EK_T - 1) / EK_T, (r + EK_T - 1) / EK_T);
  dim3 g1((r + EK_T - 1) / EK_T, (r + EK_T - 1) / EK_T);
  if (one_shot) {
    const size_t sm01 = sizeof(double) * ((size_t)n * 33 + (size_t)32 * (n | 1));
    ovb_launch(ctx, k_ekf_gemm1, dim3(g0), dim3(256), sm01, 0, P, ld, H, ldHm, nullptr, 0, ctx->d_info, N, r, n, ctx->d_M, ld, 0.0, nullptr);
    ovb_launch(ctx, k_ekf_gemm1, dim3(g1), dim3(256), sm01, 1, P, ld, H, ldHm, ctx->d_M, ld, ctx->d_info, r, r, n, ctx->d_S, ld, sigma2, Rdiag_dev);
  } else {
    ovb_launch(ctx, k_ekf_gemm, dim3(g0), dim3(256), (size_t)(0), 0, P, ld, H, ldHm, nullptr, 0, ctx->d_info, N, r, n, ctx->d_M, ld, 0.0, nullptr);
    ovb_launch(ctx, k_ekf_gemm, dim3(g1), dim3(256), (size_t)(0), 1, P, ld, H, ldHm, ctx->d_M, ld, ctx->d_info, r, r, n, ctx->d_S, ld, sigma2, Rdiag_dev);
  }
  size_t chol_bytes = sizeof(double) * (size_t)(r + 1) * (size_t)(r | 1);
  int use_smem = chol_bytes <= 200 * 1024;
  // the residual vector is column n of H's row (TSQR output) or a separate buffer: callers stage it in d_w
  double *invdiag = ctx->d_w + ctx->cfg.max_state; // d_w holds 4 x max_state doubles: [w | 1/diag(L) | ...]
  double *Lpk = nullptr; // packed factor for the register solve (only written by the DMMA Cholesky)
  const bool dmma_chol = ctx->ekf_chol_dmma && launch_chol_ekf_dmma(ctx, ctx->d_S, ld, r, ctx->d_w, ctx->d_w, invdiag, &Lpk);
  // wider than one CTA's Cholesky: blocked DMMA factorisation + blocked solve (Y = M L^-T in place over M)
  const bool wide = !dmma_chol && ctx->ekf_chol_dmma && r < ld && launch_chol_solve_wide(ctx, ctx->d_S, ld, r, ctx->d_w, invdiag, ctx->d_M, ld, N, gate_only);
  if (wide) {
    if (gate_only)
      return;
    dim3 g2w((N + EK_T - 1) / EK_T, (N + EK_T - 1) / EK_T);
    ovb_launch(ctx, k_ekf_downdate, dim3(g2w), dim3(256), (size_t)(0), P, ld, (const double *)ctx->d_M, ld, N, r, ctx->d_w, ctx->d_dx, ctx->d_info);
    return;
  }
  if (!dmma_chol)
    ovb_launch(ctx, k_ekf_chol, dim3(1), dim3(EKC_THREADS), (size_t)(use_smem ? chol_bytes : 0), ctx->d_S, ld, r, ctx->d_w, ctx->d_w, invdiag, ctx->d_info, use_smem);
  if (gate_only)
    return;
  const double *Yd = ctx->d_Y;
  if (dmma_chol && Lpk != nullptr && launch_trsm_rows(ctx, ctx->d_M, ld, N, r, Lpk)) {
    Yd = ctx->d_M; // Y = M L^-T in place
  } else {
    size_t trsm_small = sizeof(double) * ((size_t)TR_ROWS * r + r);
    size_t trsm_full = trsm_small + sizeof(double) * (size_t)r * (size_t)(r | 1);
    int L_in_smem = trsm_full <= 200 * 1024;
    ovb_launch(ctx, k_ekf_trsm, dim3((N + TR_ROWS - 1) / TR_ROWS), dim3(32 * TR_ROWS), (size_t)(L_in_smem ? trsm_full : trsm_small), ctx->d_M, ld, ctx->d_S, ld, invdiag, N,
               r, ctx->d_Y, ld, L_in_smem, ctx->d_info);
  }
  dim3 g2((N + EK_T - 1) / EK_T, (N + EK_T - 1) / EK_T);
  if (one_shot) {
    const size_t smd = sizeof(double) * ((size_t)64 * (r | 1) + r);
    ovb_launch(ctx, k_ekf_downdate1, dim3(g2), dim3(256), smd, P, ld, Yd, ld, N, r, ctx->d_w, ctx->d_dx, ctx->d_info);
  } else {
    ovb_launch(ctx, k_ekf_downdate, dim3(g2), dim3(256), (size_t)(0), P, ld, Yd, ld, N, r, ctx->d_w, ctx->d_dx, ctx->d_info);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// covariance structure operations

// StateHelper::initialize_invertible (StateHelper.cpp:484-577): grow P by the k-wide new variable.
//   m = P[:, cols] Hx'            (N x k)     Hx = invertible part's state Jacobian (k x n), cols from info->col_state
//   M = Hx P[cols, cols] Hx' + s2 I (k x k, upper triangle mirrored like selfadjointView<Upper>)
//   P[0:N, N:N+k] = -m Hinv',  P[N:N+k, 0:N] = its transpose,  P[N:N+k, N:N+k] = Hinv M Hinv'
// Single CTA (N <= a few hundred rows, k <= 3): the step is a serial point in the reference as well.
__global__ void k_cov_init_augment(double *__restrict__ P, int ld, int N, int k, int n, const DevUpdateInfo *__restrict__ info,
                                   const double *__restrict__ Hx, const double *__restrict__ Hinv, double sigma2) {
  extern __shared__ double ism[]; // m[N][k], then M[k][k]
  double *m = ism, *M = ism + (size_t)N * k;
  const int tid = threadIdx.x;
  for (int a = tid; a < N; a += blockDim.x) {
    for (int i = 0; i < k; i++) {
      double acc = 0.0;
      for (int j = 0; j < n; j++)
        acc += P[(size_t)a * ld + info->col_state[j]] * Hx[i * n + j];
      m[a * k + i] = acc;
    }
  }
  __syncthreads();
  if (tid < k * k) {
    const int i = tid / k, i2 = tid % k;
    double acc = 0.0;
    for (int j = 0; j < n; j++)
      acc += Hx[i * n + j] * m[info->col_state[j] * k + i2];
    M[i * k + i2] = acc + (i == i2 ? sigma2 : 0.0);
  }
  __syncthreads();
  if (tid < k * k) {
    const int i = tid / k, i2 = tid % k;
    if (i2 < i)
      M[i * k + i2] = M[i2 * k + i]; // selfadjointView<Upper>
  }
  __syncthreads();
  for (int a = tid; a < N; a += blockDim.x) {
    for (int q = 0; q < k; q++) {
      double acc = 0.0;
      for (int i = 0; i < k; i++)
        acc += m[a * k + i] * Hinv[q * k + i];
      P[(size_t)a * ld + N + q] = -acc;
      P[(size_t)(N + q) * ld + a] = -acc;
    }
  }
  if (tid < k * k) {
    const int q = tid / k, q2 = tid % k;
    double acc = 0.0;
    for (int i = 0; i < k; i++)
      for (int i2 = 0; i2 < k; i2++)
        acc += Hinv[q * k + i] * M[i * k + i2] * Hinv[q2 * k + i2];
    // the reference's dense product leaves P_LL symmetric only to rounding; write the upper entry to both places so the
    // resident covariance stays exactly symmetric
    if (q <= q2) {
      P[(size_t)(N + q) * ld + N + q2] = acc;
      P[(size_t)(N + q2) * ld + N + q] = acc;
    }
  }
}

bool launch_cov_init_augment(ovb_ctx *ctx, int k, int n, const double *Hx_dev, const double *Hinv_dev, double sigma2) {
  const int N = ctx->N;
  size_t smem = sizeof(double) * ((size_t)N * k + (size_t)k * k);
  if (smem > 200 * 1024)
    return false; // N*k doubles must fit one CTA's shared memory (N <= ~8500 for a 3-wide landmark)
  if (!ctx->attr_done[5]) {
    cudaFuncSetAttribute(k_cov_init_augment, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    ctx->attr_done[5] = 1;
  }
  k_cov_init_augment<<<1, 256, smem, ctx->stream>>>(ctx->P[ctx->cur], ctx->ldP, N, k, n, ctx->d_info, Hx_dev, Hinv_dev, sigma2);
  return cudaGetLastError() == cudaSuccess;
}

// StateHelper::clone: append a copy of the `size`-wide variable at old_off (StateHelper.cpp:371-373)
__global__ void k_cov_clone(double *P, int ld, int N, int old_off, int size) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int N2 = N + size;
  if (idx >= N2 * size)
    return;
  int i = idx / size, j = idx % size; // element (i, N+j) and its mirror (N+j, i)
  double v;
  if (i < N)
    v = P[(size_t)i * ld + old_off + j];
  else
    v = P[(size_t)(old_off + (i - N)) * ld + old_off + j];
  double vr = (i < N) ? P[(size_t)(old_off + j) * ld + i] : v;
  P[(size_t)i * ld + N + j] = v;
  if (i < N)
    P[(size_t)(N + j) * ld + i] = vr;
}
// augment_clone time-offset term, step 1: P[:, N..N+size) += P[:, dt] dnc'   (StateHelper.cpp:611-612)
__global__ void k_cov_dt_cols(double *P, int ld, int N2, int new_off, int size, int dt_off, const double *dnc) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N2 * size)
    return;
  int i = idx / size, j = idx % size;
  P[(size_t)i * ld + new_off + j] += P[(size_t)i * ld + dt_off] * dnc[j];
}
// step 2: P[N..N+size, :] += dnc P[dt, :]   (StateHelper.cpp:613-614) — reads the row written by step 1
__global__ void k_cov_dt_rows(double *P, int ld, int N2, int new_off, int size, int dt_off, const double *dnc) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N2 * size)
    return;
  int i = idx / N2, j = idx % N2;
  P[(size_t)(new_off + i) * ld + j] += dnc[i] * P[(size_t)dt_off * ld + j];
}

void launch_cov_clone(ovb_ctx *ctx, int old_off, int size, const double *dnc_dt_dev, int dt_off) {
  double *P = ctx->P[ctx->cur];
  int N = ctx->N, N2 = N + size;
  int tot = N2 * size;
  k_cov_clone<<<(tot + 255) / 256, 256, 0, ctx->stream>>>(P, ctx->ldP, N, old_off, size);
  if (dnc_dt_dev) {
    k_cov_dt_cols<<<(tot + 255) / 256, 256, 0, ctx->stream>>>(P, ctx->ldP, N2, N, size, dt_off, dnc_dt_dev);
    k_cov_dt_rows<<<(tot + 255) / 256, 256, 0, ctx->stream>>>(P, ctx->ldP, N2, N, size, dt_off, dnc_dt_dev);
  }
}

// StateHelper::marginalize (StateHelper.cpp:293-313): out of place into the other buffer; the x2-x1 block is the
// transpose of the copied x1-x2 block exactly as the reference builds it.
__global__ void k_cov_marg(const double *Pin, double *Pout, int ld, int N, int off, int size) {
  int N2 = N - size;
  int i = blockIdx.y * blockDim.y + threadIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N2 || j >= N2)
    return;
  int si = i < off ? i : i + size, sj = j < off ? j : j + size;
  double v = (i >= off && j < off) ? Pin[(size_t)sj * ld + si] : Pin[(size_t)si * ld + sj];
  Pout[(size_t)i * ld + j] = v;
}

void launch_cov_marginalize(ovb_ctx *ctx, int off, int size) {
  int N2 = ctx->N - size;
  dim3 b(32, 8), g((N2 + 31) / 32, (N2 + 7) / 8);
  k_cov_marg<<<g, b, 0, ctx->stream>>>(ctx->P[ctx->cur], ctx->P[ctx->cur ^ 1], ctx->ldP, ctx->N, off, size);
}

// EKFPropagation (StateHelper.cpp:80-100). old_idx[k] = covariance index of Phi's column k (q of them).
//  C[a][j]   = sum_k P[a][old_idx[k]] Phi[j][k]                       (Cov_PhiT, N x p)  -> Cbuf
//  PCP[i][j] = Qsym[i][j] + sum_k Phi[i][k] C[old_idx[k]][j]          (p x p)            -> Sbuf
__global__ void k_prop_C(const double *P, int ld, int N, int p, int q, const int *old_idx, const double *Phi, double *Cbuf, int ldC) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * p)
    return;
  int a = idx / p, j = idx % p;
  double acc = 0.0;
  for (int k = 0; k < q; k++)
    acc += P[(size_t)a * ld + old_idx[k]] * Phi[(size_t)j * q + k];
  Cbuf[(size_t)a * ldC + j] = acc;
}
__global__ void k_prop_PCP(const double *Cbuf, int ldC, int p, int q, const int *old_idx, const double *Phi, const double *Q, double *Sbuf,
                           int ldS) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p * p)
    return;
  int i = idx / p, j = idx % p;
  double acc = (i <= j) ? Q[(size_t)i * p + j] : Q[(size_t)j * p + i];
  for (int k = 0; k < q; k++)
    acc += Phi[(size_t)i * q + k] * Cbuf[(size_t)old_idx[k] * ldC + j];
  Sbuf[(size_t)i * ldS + j] = acc;
}
__global__ void k_prop_write(double *P, int ld, int N, int new_off, int p, const double *Cbuf, int ldC, const double *Sbuf, int ldS,
                             DevUpdateInfo *info) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * p)
    return;
  int a = idx / p, j = idx % p;
  if (a >= new_off && a < new_off + p) {
    double v = Sbuf[(size_t)(a - new_off) * ldS + j];
    P[(size_t)a * ld + new_off + j] = v;
    if (a - new_off == j && v < 0.0)
      atomicMin(&info->neg_diag_index, a);
  } else {
    double v = Cbuf[(size_t)a * ldC + j];
    P[(size_t)a * ld + new_off + j] = v;
    P[(size_t)(new_off + j) * ld + a] = v;
  }
}

void launch_cov_propagate(ovb_ctx *ctx, int new_off, int p, int q, const int *old_idx_dev, const double *Phi_dev, const double *Q_dev) {
  double *P = ctx->P[ctx->cur];
  int N = ctx->N, ld = ctx->ldP;
  ovb_launch(ctx, k_ekf_prep, dim3(1), dim3(1), (size_t)(0), ctx->d_info);
  k_prop_C<<<(N * p + 255) / 256, 256, 0, ctx->stream>>>(P, ld, N, p, q, old_idx_dev, Phi_dev, ctx->d_M, ld);
  k_prop_PCP<<<(p * p + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_M, ld, p, q, old_idx_dev, Phi_dev, Q_dev, ctx->d_S, ld);
  k_prop_write<<<(N * p + 255) / 256, 256, 0, ctx->stream>>>(P, ld, N, new_off, p, ctx->d_M, ld, ctx->d_S, ld, ctx->d_info);
}
