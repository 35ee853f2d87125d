// oracle/ovo_math.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Small dense algebra for the CPU restatement ("oracle") of the OpenVINS MSCKF update path.
// Dependency-free C++17 standing in for the Eigen calls the reference makes
// (Eigen 3.x is an un-vendored, unpinned system dependency that is absent from
// this image, so its published algorithms are restated here: JacobiRotation::makeGivens,
// ColPivHouseholderQR::solve, JacobiSVD (3x3 symmetric case), HouseholderQR of a 3-vector, LLT).
// PARITY UNPINNED: the reference ships no golden vectors for this path; this
// restatement is cross-checked by oracle/np_twin.py and algebraic invariants in tests/.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may use oracle/.
// Compile with -ffp-contract=off: no FMA contraction, so that the device code (built with
// -fmad=false for the per-feature stage) can follow the same rounding sequence.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <limits>
#include <vector>

namespace ovo {

// ---------------------------------------------------------------- fixed 3-vectors / 3x3 (row-major)
struct V3 {
  double v[3];
  double &operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
};
struct M3 {
  double m[3][3];
  double &operator()(int r, int c) { return m[r][c]; }
  double operator()(int r, int c) const { return m[r][c]; }
};

inline V3 v3(double a, double b, double c) { return V3{{a, b, c}}; }
inline V3 vsub(const V3 &a, const V3 &b) { return v3(a(0) - b(0), a(1) - b(1), a(2) - b(2)); }
inline V3 vadd(const V3 &a, const V3 &b) { return v3(a(0) + b(0), a(1) + b(1), a(2) + b(2)); }
inline V3 vneg(const V3 &a) { return v3(-a(0), -a(1), -a(2)); }
inline V3 vscale(double s, const V3 &a) { return v3(s * a(0), s * a(1), s * a(2)); }
inline double vdot(const V3 &a, const V3 &b) { return (a(0) * b(0) + a(1) * b(1)) + a(2) * b(2); }
inline double vnorm(const V3 &a) { return std::sqrt(vdot(a, a)); }
inline M3 m3zero() {
  M3 r;
  std::memset(&r, 0, sizeof(r));
  return r;
}
inline M3 m3eye() {
  M3 r = m3zero();
  r(0, 0) = r(1, 1) = r(2, 2) = 1.0;
  return r;
}
inline M3 m3T(const M3 &a) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      r(i, j) = a(j, i);
  return r;
}
inline M3 m3neg(const M3 &a) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      r(i, j) = -a(i, j);
  return r;
}
// coefficient-wise lazy product, k ascending: ((a0*b0 + a1*b1) + a2*b2)
inline M3 m3mul(const M3 &a, const M3 &b) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      r(i, j) = (a(i, 0) * b(0, j) + a(i, 1) * b(1, j)) + a(i, 2) * b(2, j);
  return r;
}
inline M3 m3mulT(const M3 &a, const M3 &b) { return m3mul(a, m3T(b)); } // a * b'
inline M3 m3Tmul(const M3 &a, const M3 &b) { return m3mul(m3T(a), b); } // a' * b
inline V3 m3v(const M3 &a, const V3 &x) {
  V3 r;
  for (int i = 0; i < 3; i++)
    r(i) = (a(i, 0) * x(0) + a(i, 1) * x(1)) + a(i, 2) * x(2);
  return r;
}
inline V3 m3Tv(const M3 &a, const V3 &x) { return m3v(m3T(a), x); }
// ov_core/src/utils/quat_ops.h:135-139 skew_x
inline M3 skew(const V3 &w) {
  M3 r;
  r(0, 0) = 0;
  r(0, 1) = -w(2);
  r(0, 2) = w(1);
  r(1, 0) = w(2);
  r(1, 1) = 0;
  r(1, 2) = -w(0);
  r(2, 0) = -w(1);
  r(2, 1) = w(0);
  r(2, 2) = 0;
  return r;
}

// ---------------------------------------------------------------- Eigen::JacobiRotation (real scalars)
// makeGivens(p,q): G' * [p;q] = [r;0], r = |.| >= 0 (Eigen/src/Jacobi/Jacobi.h, real branch)
struct Givens {
  double c, s;
};
inline Givens make_givens(double p, double q) {
  Givens g;
  if (q == 0.0) {
    g.c = p < 0.0 ? -1.0 : 1.0;
    g.s = 0.0;
  } else if (p == 0.0) {
    g.c = 0.0;
    g.s = q < 0.0 ? 1.0 : -1.0;
  } else if (std::fabs(p) > std::fabs(q)) {
    double t = q / p;
    double u = std::sqrt(1.0 + t * t);
    if (p < 0.0)
      u = -u;
    g.c = 1.0 / u;
    g.s = -t * g.c;
  } else {
    double t = p / q;
    double u = std::sqrt(1.0 + t * t);
    if (q < 0.0)
      u = -u;
    g.s = -1.0 / u;
    g.c = -t * g.s;
  }
  return g;
}
// applyOnTheLeft(0,1,G.adjoint()) on the row pair (x,y): x' = c x - s y ; y' = s x + c y
inline void apply_givens(const Givens &g, double &x, double &y) {
  double xi = x, yi = y;
  x = g.c * xi - g.s * yi;
  y = g.s * xi + g.c * yi;
}

// ---------------------------------------------------------------- 3x3 column-pivoted Householder solve
// Restates Eigen::ColPivHouseholderQR<Matrix3d>::solve: at step k pivot the remaining column with the largest
// squared norm (of rows k..2), Householder-reflect, then c = Q'b, back-substitute on the rank-revealed R and
// undo the permutation. Eigen's norm down-dating only changes the pivot on near-ties.
inline V3 colpiv_qr_solve3(const M3 &Ain, const V3 &bin) {
  double A[3][3];
  double b[3] = {bin(0), bin(1), bin(2)};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      A[i][j] = Ain(i, j);
  int perm[3] = {0, 1, 2};
  double maxpivot = 0.0;
  int rank = 0;
  double diag[3] = {0, 0, 0};
  for (int k = 0; k < 3; k++) {
    // pick pivot column
    int best = k;
    double bestn = -1.0;
    for (int j = k; j < 3; j++) {
      double n2 = 0.0;
      for (int i = k; i < 3; i++)
        n2 += A[i][j] * A[i][j];
      if (n2 > bestn) {
        bestn = n2;
        best = j;
      }
    }
    if (best != k) {
      for (int i = 0; i < 3; i++)
        std::swap(A[i][k], A[i][best]);
      std::swap(perm[k], perm[best]);
    }
    // Householder on A[k..2][k]  (Eigen makeHouseholderInPlace convention: beta = -sign(c0)*norm)
    double c0 = A[k][k];
    double tail2 = 0.0;
    for (int i = k + 1; i < 3; i++)
      tail2 += A[i][k] * A[i][k];
    double tau, beta;
    double ess[3] = {0, 0, 0};
    if (tail2 <= std::numeric_limits<double>::min()) {
      tau = 0.0;
      beta = c0;
    } else {
      beta = std::sqrt(c0 * c0 + tail2);
      if (c0 >= 0.0)
        beta = -beta;
      for (int i = k + 1; i < 3; i++)
        ess[i] = A[i][k] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    A[k][k] = beta;
    for (int i = k + 1; i < 3; i++)
      A[i][k] = 0.0;
    diag[k] = beta;
    if (std::fabs(beta) > maxpivot)
      maxpivot = std::fabs(beta);
    // apply H = I - tau v v', v = [1; ess] to remaining columns and to b
    if (tau != 0.0) {
      for (int j = k + 1; j < 3; j++) {
        double w = A[k][j];
        for (int i = k + 1; i < 3; i++)
          w += ess[i] * A[i][j];
        A[k][j] -= tau * w;
        for (int i = k + 1; i < 3; i++)
          A[i][j] -= tau * w * ess[i];
      }
      double w = b[k];
      for (int i = k + 1; i < 3; i++)
        w += ess[i] * b[i];
      b[k] -= tau * w;
      for (int i = k + 1; i < 3; i++)
        b[i] -= tau * w * ess[i];
    }
  }
  // rank by Eigen's default threshold: |pivot| > eps*size * maxpivot
  double thr = maxpivot * (std::numeric_limits<double>::epsilon() * 3.0);
  for (int k = 0; k < 3; k++)
    if (std::fabs(diag[k]) > thr)
      rank++;
  double y[3] = {0, 0, 0};
  for (int k = rank - 1; k >= 0; k--) {
    double s = b[k];
    for (int j = k + 1; j < rank; j++)
      s -= A[k][j] * y[j];
    y[k] = s / A[k][k];
  }
  V3 x = v3(0, 0, 0);
  for (int k = 0; k < 3; k++)
    x(perm[k]) = y[k];
  return x;
}

// ---------------------------------------------------------------- cond(A) for the symmetric 3x3 normal matrix
// The reference takes JacobiSVD singular values of A = sum (I - b b') (feat/FeatureInitializer.cpp:91-95). A is
// symmetric PSD, so singular values == eigenvalues; cyclic Jacobi sweeps give them to full precision.
inline void sym_eig3(const M3 &Ain, double ev[3]) {
  double a[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      a[i][j] = 0.5 * (Ain(i, j) + Ain(j, i));
  for (int sweep = 0; sweep < 30; sweep++) {
    double off = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
    double dg = std::fabs(a[0][0]) + std::fabs(a[1][1]) + std::fabs(a[2][2]);
    if (off <= 1e-300 || off <= 1e-17 * dg)
      break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (a[p][q] == 0.0)
          continue;
        double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; k++) { // columns
          double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) { // rows
          double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
      }
  }
  ev[0] = a[0][0];
  ev[1] = a[1][1];
  ev[2] = a[2][2];
}
inline double cond_sym3(const M3 &A) {
  double ev[3];
  sym_eig3(A, ev);
  double mx = std::max(std::fabs(ev[0]), std::max(std::fabs(ev[1]), std::fabs(ev[2])));
  double mn = std::min(std::fabs(ev[0]), std::min(std::fabs(ev[1]), std::fabs(ev[2])));
  return mx / mn;
}

// ---------------------------------------------------------------- Q(:,1:2) of HouseholderQR of a 3-vector
// feat/FeatureInitializer.cpp:338-339,354: the two columns of Q orthogonal to p. H = I - tau v v', v = [1; ess].
inline void householder_tangent3(const V3 &p, V3 &q1, V3 &q2) {
  double c0 = p(0);
  double tail2 = p(1) * p(1) + p(2) * p(2);
  double tau, ess1, ess2;
  if (tail2 <= std::numeric_limits<double>::min()) {
    tau = 0.0;
    ess1 = ess2 = 0.0;
  } else {
    double beta = std::sqrt(c0 * c0 + tail2);
    if (c0 >= 0.0)
      beta = -beta;
    ess1 = p(1) / (c0 - beta);
    ess2 = p(2) / (c0 - beta);
    tau = (beta - c0) / beta;
  }
  double v[3] = {1.0, ess1, ess2};
  for (int i = 0; i < 3; i++) {
    q1(i) = (i == 1 ? 1.0 : 0.0) - tau * v[i] * v[1];
    q2(i) = (i == 2 ? 1.0 : 0.0) - tau * v[i] * v[2];
  }
}

// ---------------------------------------------------------------- dynamic dense matrix, column-major like Eigen::MatrixXd
struct Mat {
  int r = 0, c = 0;
  std::vector<double> d;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), d((size_t)r_ * c_, 0.0) {}
  double &operator()(int i, int j) { return d[(size_t)j * r + i]; }
  double operator()(int i, int j) const { return d[(size_t)j * r + i]; }
  void resize_zero(int r_, int c_) {
    r = r_;
    c = c_;
    d.assign((size_t)r_ * c_, 0.0);
  }
};

// C = A * B        (plain triple loop, j-k-i order so the inner loop is contiguous)
inline Mat matmul(const Mat &A, const Mat &B) {
  Mat C(A.r, B.c);
  for (int j = 0; j < B.c; j++)
    for (int k = 0; k < A.c; k++) {
      double b = B(k, j);
      if (b == 0.0)
        continue;
      const double *a = &A.d[(size_t)k * A.r];
      double *cc = &C.d[(size_t)j * C.r];
      for (int i = 0; i < A.r; i++)
        cc[i] += a[i] * b;
    }
  return C;
}
// C = A * B'
inline Mat matmul_nt(const Mat &A, const Mat &B) {
  Mat C(A.r, B.r);
  for (int k = 0; k < A.c; k++)
    for (int j = 0; j < B.r; j++) {
      double b = B(j, k);
      if (b == 0.0)
        continue;
      const double *a = &A.d[(size_t)k * A.r];
      double *cc = &C.d[(size_t)j * C.r];
      for (int i = 0; i < A.r; i++)
        cc[i] += a[i] * b;
    }
  return C;
}

// In-place lower Cholesky of the symmetric matrix whose UPPER triangle is valid (selfadjointView<Upper>().llt()).
// Returns false if a pivot is not positive. L is left in the lower triangle (upper untouched).
inline bool llt_from_upper(Mat &S) {
  int n = S.r;
  for (int j = 0; j < n; j++) {
    double d = S(j, j);
    for (int k = 0; k < j; k++)
      d -= S(j, k) * S(j, k);
    if (!(d > 0.0))
      return false;
    d = std::sqrt(d);
    S(j, j) = d;
    for (int i = j + 1; i < n; i++) {
      double v = S(j, i); // upper entry (j,i) == symmetric (i,j)
      for (int k = 0; k < j; k++)
        v -= S(i, k) * S(j, k);
      S(i, j) = v / d;
    }
  }
  return true;
}
// solve L L' x = b in place given L in the lower triangle
inline void llt_solve_vec(const Mat &L, double *x) {
  int n = L.r;
  for (int i = 0; i < n; i++) {
    double v = x[i];
    for (int k = 0; k < i; k++)
      v -= L(i, k) * x[k];
    x[i] = v / L(i, i);
  }
  for (int i = n - 1; i >= 0; i--) {
    double v = x[i];
    for (int k = i + 1; k < n; k++)
      v -= L(k, i) * x[k];
    x[i] = v / L(i, i);
  }
}

} // namespace ovo
