// geom.cuh — device-side fixed-size geometry for the per-feature stage.
// Files including this header are compiled with -fmad=false: every expression below is evaluated with separate
// IEEE roundings, in the association order of the reference's Eigen expressions (left to right, k ascending in
// 3x3 products), so that the only differences against a sequential CPU evaluation come from warp-tree reductions.
#pragma once
#include "ovb_internal.cuh"
#include <math.h>

struct dv3 {
  double x, y, z;
};
struct dm3 { // row-major
  double m[9];
};

__device__ __forceinline__ dv3 mk3(double a, double b, double c) { return dv3{a, b, c}; }
__device__ __forceinline__ dv3 sub3(dv3 a, dv3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ dv3 add3(dv3 a, dv3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ double dot3(dv3 a, dv3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ double norm3(dv3 a) { return sqrt(dot3(a, a)); }
__device__ __forceinline__ dm3 ld_m3(const double *p) {
  dm3 r;
#pragma unroll
  for (int i = 0; i < 9; i++)
    r.m[i] = p[i];
  return r;
}
__device__ __forceinline__ dv3 ld_v3(const double *p) { return mk3(p[0], p[1], p[2]); }
// A*B, coefficient (i,j) = (a_i0 b_0j + a_i1 b_1j) + a_i2 b_2j
__device__ __forceinline__ dm3 mul33(const dm3 &a, const dm3 &b) {
  dm3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      r.m[3 * i + j] = (a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j]) + a.m[3 * i + 2] * b.m[6 + j];
  return r;
}
// A*B'
__device__ __forceinline__ dm3 mul33T(const dm3 &a, const dm3 &b) {
  dm3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      r.m[3 * i + j] = (a.m[3 * i] * b.m[3 * j] + a.m[3 * i + 1] * b.m[3 * j + 1]) + a.m[3 * i + 2] * b.m[3 * j + 2];
  return r;
}
// A'*B
__device__ __forceinline__ dm3 mulT33(const dm3 &a, const dm3 &b) {
  dm3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      r.m[3 * i + j] = (a.m[i] * b.m[j] + a.m[3 + i] * b.m[3 + j]) + a.m[6 + i] * b.m[6 + j];
  return r;
}
__device__ __forceinline__ dv3 mv3(const dm3 &a, dv3 v) {
  return mk3((a.m[0] * v.x + a.m[1] * v.y) + a.m[2] * v.z, (a.m[3] * v.x + a.m[4] * v.y) + a.m[5] * v.z,
             (a.m[6] * v.x + a.m[7] * v.y) + a.m[8] * v.z);
}
// A'*v
__device__ __forceinline__ dv3 mTv3(const dm3 &a, dv3 v) {
  return mk3((a.m[0] * v.x + a.m[3] * v.y) + a.m[6] * v.z, (a.m[1] * v.x + a.m[4] * v.y) + a.m[7] * v.z,
             (a.m[2] * v.x + a.m[5] * v.y) + a.m[8] * v.z);
}
// (-A)*v
__device__ __forceinline__ dv3 negmv3(const dm3 &a, dv3 v) {
  return mk3(((-a.m[0]) * v.x + (-a.m[1]) * v.y) + (-a.m[2]) * v.z, ((-a.m[3]) * v.x + (-a.m[4]) * v.y) + (-a.m[5]) * v.z,
             ((-a.m[6]) * v.x + (-a.m[7]) * v.y) + (-a.m[8]) * v.z);
}
__device__ __forceinline__ dm3 skew3(dv3 w) {
  dm3 r;
  r.m[0] = 0;
  r.m[1] = -w.z;
  r.m[2] = w.y;
  r.m[3] = w.z;
  r.m[4] = 0;
  r.m[5] = -w.x;
  r.m[6] = -w.y;
  r.m[7] = w.x;
  r.m[8] = 0;
  return r;
}

// ---- 3x3 column-pivoted Householder solve (Eigen::ColPivHouseholderQR<Matrix3d>::solve restated;
// feat/FeatureInitializer.cpp:88,294). Uniform across the warp: every lane solves the same system.
__device__ inline dv3 colpiv_solve3(const double Ain[9], dv3 bin) {
  double A[3][3];
  double b[3] = {bin.x, bin.y, bin.z};
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      A[i][j] = Ain[3 * i + j];
  int perm[3] = {0, 1, 2};
  double maxpivot = 0.0;
  double diag[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    int best = k;
    double bestn = -1.0;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      if (j < k)
        continue;
      double n2 = 0.0;
#pragma unroll
      for (int i = 0; i < 3; i++)
        if (i >= k)
          n2 += A[i][j] * A[i][j];
      if (n2 > bestn) {
        bestn = n2;
        best = j;
      }
    }
    if (best != k) {
#pragma unroll
      for (int i = 0; i < 3; i++) {
        double t = A[i][k];
#pragma unroll
        for (int j = 0; j < 3; j++)
          if (j == best) {
            A[i][k] = A[i][j];
            A[i][j] = t;
          }
      }
      int tp = perm[k];
#pragma unroll
      for (int j = 0; j < 3; j++)
        if (j == best) {
          perm[k] = perm[j];
          perm[j] = tp;
        }
    }
    double c0 = A[k][k];
    double tail2 = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++)
      if (i > k)
        tail2 += A[i][k] * A[i][k];
    double tau, beta;
    double ess[3] = {0, 0, 0};
    if (tail2 <= 2.2250738585072014e-308) {
      tau = 0.0;
      beta = c0;
    } else {
      beta = sqrt(c0 * c0 + tail2);
      if (c0 >= 0.0)
        beta = -beta;
#pragma unroll
      for (int i = 0; i < 3; i++)
        if (i > k)
          ess[i] = A[i][k] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    A[k][k] = beta;
#pragma unroll
    for (int i = 0; i < 3; i++)
      if (i > k)
        A[i][k] = 0.0;
    diag[k] = beta;
    if (fabs(beta) > maxpivot)
      maxpivot = fabs(beta);
    if (tau != 0.0) {
#pragma unroll
      for (int j = 0; j < 3; j++) {
        if (j <= k)
          continue;
        double w = A[k][j];
#pragma unroll
        for (int i = 0; i < 3; i++)
          if (i > k)
            w += ess[i] * A[i][j];
        A[k][j] -= tau * w;
#pragma unroll
        for (int i = 0; i < 3; i++)
          if (i > k)
            A[i][j] -= tau * w * ess[i];
      }
      double w = b[k];
#pragma unroll
      for (int i = 0; i < 3; i++)
        if (i > k)
          w += ess[i] * b[i];
      b[k] -= tau * w;
#pragma unroll
      for (int i = 0; i < 3; i++)
        if (i > k)
          b[i] -= tau * w * ess[i];
    }
  }
  double thr = maxpivot * (2.220446049250313e-16 * 3.0);
  int rank = 0;
#pragma unroll
  for (int k = 0; k < 3; k++)
    if (fabs(diag[k]) > thr)
      rank++;
  double y[3] = {0, 0, 0};
#pragma unroll
  for (int k = 2; k >= 0; k--) {
    if (k >= rank)
      continue;
    double s = b[k];
#pragma unroll
    for (int j = 0; j < 3; j++)
      if (j > k && j < rank)
        s -= A[k][j] * y[j];
    y[k] = s / A[k][k];
  }
  double x[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < 3; k++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      if (perm[k] == j)
        x[j] = y[k];
  return mk3(x[0], x[1], x[2]);
}

// ---- condition number of the symmetric PSD 3x3 normal matrix via cyclic Jacobi (JacobiSVD singular values of a
// symmetric PSD matrix are its eigenvalues; feat/FeatureInitializer.cpp:91-95)
__device__ inline double cond_sym3(const double Ain[9]) {
  double a[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      a[i][j] = 0.5 * (Ain[3 * i + j] + Ain[3 * j + i]);
  for (int sweep = 0; sweep < 30; sweep++) {
    double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    double dg = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
    if (off <= 1e-300 || off <= 1e-17 * dg)
      break;
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
      for (int q = 1; q < 3; q++) {
        if (q <= p)
          continue;
        if (a[p][q] == 0.0)
          continue;
        double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
      }
  }
  double e0 = fabs(a[0][0]), e1 = fabs(a[1][1]), e2 = fabs(a[2][2]);
  double mx = fmax(e0, fmax(e1, e2));
  double mn = fmin(e0, fmin(e1, e2));
  return mx / mn;
}

// ---- the two columns of Q (HouseholderQR of the 3-vector p) orthogonal to p; feat/FeatureInitializer.cpp:338-354
__device__ inline void householder_tangent3(dv3 p, dv3 &q1, dv3 &q2) {
  double c0 = p.x;
  double tail2 = p.y * p.y + p.z * p.z;
  double tau, e1, e2;
  if (tail2 <= 2.2250738585072014e-308) {
    tau = 0.0;
    e1 = e2 = 0.0;
  } else {
    double beta = sqrt(c0 * c0 + tail2);
    if (c0 >= 0.0)
      beta = -beta;
    e1 = p.y / (c0 - beta);
    e2 = p.z / (c0 - beta);
    tau = (beta - c0) / beta;
  }
  q1 = mk3(0.0 - tau * 1.0 * e1, 1.0 - tau * e1 * e1, 0.0 - tau * e2 * e1);
  q2 = mk3(0.0 - tau * 1.0 * e2, 0.0 - tau * e1 * e2, 1.0 - tau * e2 * e2);
}

// ---- camera models. distort with the float round trip of CamBase::distort_d (cam/CamBase.h:130-135,
// cam/CamRadtan.h:127-146, cam/CamEqui.h:136-158)
__device__ inline void cam_distort_d(int model, const double *cam_d, double xn_d, double yn_d, double &u, double &v) {
  float xf = (float)xn_d, yf = (float)yn_d;
  double x = (double)xf, y = (double)yf;
  float r2f = __fadd_rn(__fmul_rn(xf, xf), __fmul_rn(yf, yf));
  double r = (double)__fsqrt_rn(r2f);
  if (model == OVB_CAM_RADTAN) {
    double r_2 = r * r;
    double r_4 = r_2 * r_2;
    float two_xx = __fmul_rn(__fmul_rn(2.0f, xf), xf);
    float two_yy = __fmul_rn(__fmul_rn(2.0f, yf), yf);
    double x1 = x * (1 + cam_d[4] * r_2 + cam_d[5] * r_4) + 2 * cam_d[6] * x * y + cam_d[7] * (r_2 + (double)two_xx);
    double y1 = y * (1 + cam_d[4] * r_2 + cam_d[5] * r_4) + cam_d[6] * (r_2 + (double)two_yy) + 2 * cam_d[7] * x * y;
    u = (double)(float)(cam_d[0] * x1 + cam_d[2]);
    v = (double)(float)(cam_d[1] * y1 + cam_d[3]);
  } else {
    double theta = atan(r);
    double t2 = theta * theta;
    double t3 = t2 * theta, t5 = t3 * t2, t7 = t5 * t2, t9 = t7 * t2;
    double theta_d = theta + cam_d[4] * t3 + cam_d[5] * t5 + cam_d[6] * t7 + cam_d[7] * t9;
    double inv_r = (r > 1e-8) ? 1.0 / r : 1.0;
    double cdist = (r > 1e-8) ? theta_d * inv_r : 1.0;
    double x1 = x * cdist;
    double y1 = y * cdist;
    u = (double)(float)(cam_d[0] * x1 + cam_d[2]);
    v = (double)(float)(cam_d[1] * y1 + cam_d[3]);
  }
}

// compute_distort_jacobian (cam/CamRadtan.h:154-199, cam/CamEqui.h:166-234). dzn 2x2 row-major, dzeta 2x8 row-major.
__device__ inline void cam_distort_jacobian(int model, const double *cam_d, double x, double y, double dzn[4], double dzeta[16]) {
#pragma unroll
  for (int i = 0; i < 16; i++)
    dzeta[i] = 0.0;
  double r = sqrt(x * x + y * y);
  if (model == OVB_CAM_RADTAN) {
    double r_2 = r * r;
    double r_4 = r_2 * r_2;
    double x_2 = x * x, y_2 = y * y, x_y = x * y;
    dzn[0] = cam_d[0] * ((1 + cam_d[4] * r_2 + cam_d[5] * r_4) + (2 * cam_d[4] * x_2 + 4 * cam_d[5] * x_2 * r_2) + 2 * cam_d[6] * y +
                         (2 * cam_d[7] * x + 4 * cam_d[7] * x));
    dzn[1] = cam_d[0] * (2 * cam_d[4] * x_y + 4 * cam_d[5] * x_y * r_2 + 2 * cam_d[6] * x + 2 * cam_d[7] * y);
    dzn[2] = cam_d[1] * (2 * cam_d[4] * x_y + 4 * cam_d[5] * x_y * r_2 + 2 * cam_d[6] * x + 2 * cam_d[7] * y);
    dzn[3] = cam_d[1] * ((1 + cam_d[4] * r_2 + cam_d[5] * r_4) + (2 * cam_d[4] * y_2 + 4 * cam_d[5] * y_2 * r_2) + 2 * cam_d[7] * x +
                         (2 * cam_d[6] * y + 4 * cam_d[6] * y));
    double x1 = x * (1 + cam_d[4] * r_2 + cam_d[5] * r_4) + 2 * cam_d[6] * x * y + cam_d[7] * (r_2 + 2 * x * x);
    double y1 = y * (1 + cam_d[4] * r_2 + cam_d[5] * r_4) + cam_d[6] * (r_2 + 2 * y * y) + 2 * cam_d[7] * x * y;
    dzeta[0] = x1;
    dzeta[2] = 1;
    dzeta[4] = cam_d[0] * x * r_2;
    dzeta[5] = cam_d[0] * x * r_4;
    dzeta[6] = 2 * cam_d[0] * x * y;
    dzeta[7] = cam_d[0] * (r_2 + 2 * x * x);
    dzeta[8 + 1] = y1;
    dzeta[8 + 3] = 1;
    dzeta[8 + 4] = cam_d[1] * y * r_2;
    dzeta[8 + 5] = cam_d[1] * y * r_4;
    dzeta[8 + 6] = cam_d[1] * (r_2 + 2 * y * y);
    dzeta[8 + 7] = 2 * cam_d[1] * x * y;
  } else {
    double theta = atan(r);
    double t2 = theta * theta;
    double t3 = t2 * theta, t4 = t2 * t2, t5 = t3 * t2, t6 = t3 * t3, t7 = t5 * t2, t8 = t4 * t4, t9 = t7 * t2;
    double theta_d = theta + cam_d[4] * t3 + cam_d[5] * t5 + cam_d[6] * t7 + cam_d[7] * t9;
    double inv_r = (r > 1e-8) ? 1.0 / r : 1.0;
    double cdist = (r > 1e-8) ? theta_d * inv_r : 1.0;
    double dxy_dxyn = theta_d * inv_r;
    double dxy_dr0 = -x * theta_d * inv_r * inv_r, dxy_dr1 = -y * theta_d * inv_r * inv_r;
    double dr0 = x * inv_r, dr1 = y * inv_r;
    double dthd_dth = 1 + 3 * cam_d[4] * t2 + 5 * cam_d[5] * t4 + 7 * cam_d[6] * t6 + 9 * cam_d[7] * t8;
    double dth_dr = 1 / (r * r + 1);
    double c0 = dxy_dr0 + dr0 * dthd_dth * dth_dr, c1 = dxy_dr1 + dr1 * dthd_dth * dth_dr;
    double i0 = dxy_dxyn + c0 * dr0, i1 = 0.0 + c0 * dr1, i2 = 0.0 + c1 * dr0, i3 = dxy_dxyn + c1 * dr1;
    dzn[0] = cam_d[0] * i0;
    dzn[1] = cam_d[0] * i1;
    dzn[2] = cam_d[1] * i2;
    dzn[3] = cam_d[1] * i3;
    double x1 = x * cdist, y1 = y * cdist;
    dzeta[0] = x1;
    dzeta[2] = 1;
    dzeta[4] = cam_d[0] * x * inv_r * t3;
    dzeta[5] = cam_d[0] * x * inv_r * t5;
    dzeta[6] = cam_d[0] * x * inv_r * t7;
    dzeta[7] = cam_d[0] * x * inv_r * t9;
    dzeta[8 + 1] = y1;
    dzeta[8 + 3] = 1;
    dzeta[8 + 4] = cam_d[1] * y * inv_r * t3;
    dzeta[8 + 5] = cam_d[1] * y * inv_r * t5;
    dzeta[8 + 6] = cam_d[1] * y * inv_r * t7;
    dzeta[8 + 7] = cam_d[1] * y * inv_r * t9;
  }
}

// ---- warp helpers
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
