// oracle/ovo_core.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (see ovo_math.hpp).
//
// CPU restatement of the OpenVINS MSCKF-update hot path, one function per reference function,
// following the arithmetic (operation order, float32 casts, FEJ choices) of the cited lines.
// All file:line citations are relative to the reference root (/root/reference at authoring time).
//
// Data enters through the plain-C structs of include/ovb200.h (the same ones the product's C ABI takes),
// so tests can hand byte-identical inputs to both sides.
#pragma once
#include "../include/ovb200.h"
#include "ovo_math.hpp"
#include <chrono>
#include <cstdint>
#include <map>

namespace ovo {

// ======================================================================= cameras
// ov_core/src/cam/CamRadtan.h:127-146, CamEqui.h:136-158 via CamBase::distort_d (CamBase.h:130-135):
// the normalized point is cast to float, the model is evaluated in double FROM THOSE FLOATS, the pixel is cast
// to float and promoted back to double.
inline void distort_d(int model, const double *cam_d, double xn_d, double yn_d, double &u, double &v) {
  float xf = (float)xn_d, yf = (float)yn_d; // uv_norm.cast<float>()
  double x = (double)xf, y = (double)yf;
  if (model == OVB_CAM_RADTAN) {
    // NOTE the reference forms r from FLOAT products: uv_norm(0)*uv_norm(0) is float*float (Vector2f) summed in float,
    // then std::sqrt(float) -> float, assigned to double r.
    float r2f = xf * xf + yf * yf;
    double r = (double)std::sqrt(r2f);
    double r_2 = r * r;
    double r_4 = r_2 * r_2;
    // uv_norm(0) * (double expr): float promoted to double, so the rest is double arithmetic.
    // 2 * cam_d(6) * uv_norm(0) * uv_norm(1): ((2*d6)*x)*y in double
    // (r_2 + 2 * uv_norm(0) * uv_norm(0)): 2*uv_norm(0) is int*float -> float; times uv_norm(0) float -> float;
    // added to double r_2.
    float two_xx = (2 * xf) * xf;
    float two_yy = (2 * yf) * yf;
    double x1 = x * (1 + cam_d[4] * r_2 + cam_d[5] * r_4) + 2 * cam_d[6] * x * y + cam_d[7] * (r_2 + (double)two_xx);
    double y1 = y * (1 + cam_d[4] * r_2 + cam_d[5] * r_4) + cam_d[6] * (r_2 + (double)two_yy) + 2 * cam_d[7] * x * y;
    u = (double)(float)(cam_d[0] * x1 + cam_d[2]);
    v = (double)(float)(cam_d[1] * y1 + cam_d[3]);
  } else {
    float r2f = xf * xf + yf * yf;
    double r = (double)std::sqrt(r2f);
    double theta = std::atan(r);
    double theta_d = theta + cam_d[4] * std::pow(theta, 3) + cam_d[5] * std::pow(theta, 5) + cam_d[6] * std::pow(theta, 7) +
                     cam_d[7] * std::pow(theta, 9);
    double inv_r = (r > 1e-8) ? 1.0 / r : 1.0;
    double cdist = (r > 1e-8) ? theta_d * inv_r : 1.0;
    double x1 = x * cdist;
    double y1 = y * cdist;
    u = (double)(float)(cam_d[0] * x1 + cam_d[2]);
    v = (double)(float)(cam_d[1] * y1 + cam_d[3]);
  }
}

// CamRadtan.h:154-199 / CamEqui.h:166-234 compute_distort_jacobian (all double, at the double uv_norm).
// dzn is 2x2 row-major, dzeta is 2x8 row-major.
inline void distort_jacobian(int model, const double *cam_d, double x, double y, double dzn[4], double dzeta[16]) {
  for (int i = 0; i < 16; i++)
    dzeta[i] = 0.0;
  if (model == OVB_CAM_RADTAN) {
    double r = std::sqrt(x * x + y * y);
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
    double r = std::sqrt(x * x + y * y);
    double theta = std::atan(r);
    double theta_d = theta + cam_d[4] * std::pow(theta, 3) + cam_d[5] * std::pow(theta, 5) + cam_d[6] * std::pow(theta, 7) +
                     cam_d[7] * std::pow(theta, 9);
    double inv_r = (r > 1e-8) ? 1.0 / r : 1.0;
    double cdist = (r > 1e-8) ? theta_d * inv_r : 1.0;
    double dxy_dxyn = theta_d * inv_r;
    double dxy_dr[2] = {-x * theta_d * inv_r * inv_r, -y * theta_d * inv_r * inv_r};
    double dr_dxyn[2] = {x * inv_r, y * inv_r};
    double dxy_dthd[2] = {x * inv_r, y * inv_r};
    double dthd_dth = 1 + 3 * cam_d[4] * std::pow(theta, 2) + 5 * cam_d[5] * std::pow(theta, 4) + 7 * cam_d[6] * std::pow(theta, 6) +
                      9 * cam_d[7] * std::pow(theta, 8);
    double dth_dr = 1 / (r * r + 1);
    // duv_dxy * (dxy_dxyn + (dxy_dr + dxy_dthd * dthd_dth * dth_dr) * dr_dxyn)
    double col[2] = {dxy_dr[0] + dxy_dthd[0] * dthd_dth * dth_dr, dxy_dr[1] + dxy_dthd[1] * dthd_dth * dth_dr};
    double inner[4] = {dxy_dxyn + col[0] * dr_dxyn[0], 0.0 + col[0] * dr_dxyn[1], 0.0 + col[1] * dr_dxyn[0], dxy_dxyn + col[1] * dr_dxyn[1]};
    dzn[0] = cam_d[0] * inner[0] + 0.0 * inner[2];
    dzn[1] = cam_d[0] * inner[1] + 0.0 * inner[3];
    dzn[2] = 0.0 * inner[0] + cam_d[1] * inner[2];
    dzn[3] = 0.0 * inner[1] + cam_d[1] * inner[3];
    double x1 = x * cdist, y1 = y * cdist;
    dzeta[0] = x1;
    dzeta[2] = 1;
    dzeta[4] = cam_d[0] * x * inv_r * std::pow(theta, 3);
    dzeta[5] = cam_d[0] * x * inv_r * std::pow(theta, 5);
    dzeta[6] = cam_d[0] * x * inv_r * std::pow(theta, 7);
    dzeta[7] = cam_d[0] * x * inv_r * std::pow(theta, 9);
    dzeta[8 + 1] = y1;
    dzeta[8 + 3] = 1;
    dzeta[8 + 4] = cam_d[1] * y * inv_r * std::pow(theta, 3);
    dzeta[8 + 5] = cam_d[1] * y * inv_r * std::pow(theta, 5);
    dzeta[8 + 6] = cam_d[1] * y * inv_r * std::pow(theta, 7);
    dzeta[8 + 7] = cam_d[1] * y * inv_r * std::pow(theta, 9);
  }
}

// ======================================================================= frame helpers
inline M3 load_m3(const double *p) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      r(i, j) = p[i * 3 + j];
  return r;
}
inline V3 load_v3(const double *p) { return v3(p[0], p[1], p[2]); }

// FeatureInitializer::ClonePose of camera k at clone c: update/UpdaterMSCKF.cpp:98-115
struct CamPose {
  M3 R; // R_GtoCi
  V3 p; // p_CiinG
};
inline CamPose cam_clone_pose(const ovb_frame &fr, int cam, int clone) {
  M3 R_ItoC = load_m3(fr.cam_R + 9 * cam);
  V3 p_IinC = load_v3(fr.cam_p + 3 * cam);
  M3 R_GtoI = load_m3(fr.clone_R + 9 * clone);
  V3 p_IinG = load_v3(fr.clone_p + 3 * clone);
  CamPose cp;
  cp.R = m3mul(R_ItoC, R_GtoI);
  cp.p = vsub(p_IinG, m3Tv(cp.R, p_IinC));
  return cp;
}

// per-feature view of the batch
struct FeatView {
  int m0, m1;            // measurement range
  std::vector<int> keys; // camera keys in visit order (may include cameras with no measurements)
};
inline FeatView feat_view(const ovb_feat_batch &fb, int f) {
  FeatView fv;
  fv.m0 = fb.meas_off[f];
  fv.m1 = fb.meas_off[f + 1];
  if (fb.cam_keys_off && fb.cam_keys) {
    for (int i = fb.cam_keys_off[f]; i < fb.cam_keys_off[f + 1]; i++)
      fv.keys.push_back(fb.cam_keys[i]);
  } else {
    for (int i = fv.m0; i < fv.m1; i++)
      if (fv.keys.empty() || fv.keys.back() != (int)fb.cam[i])
        fv.keys.push_back(fb.cam[i]);
  }
  return fv;
}

// anchor selection: feat/FeatureInitializer.cpp:35-46 — first visited camera with the strictly largest count,
// its LAST measurement.
inline bool pick_anchor(const ovb_feat_batch &fb, const FeatView &fv, int &anchor_cam, int &anchor_meas) {
  size_t most = 0;
  anchor_cam = 0;
  bool any = false;
  for (int key : fv.keys) {
    size_t cnt = 0;
    for (int i = fv.m0; i < fv.m1; i++)
      if (fb.cam[i] == key)
        cnt++;
    if (cnt > most) {
      anchor_cam = key;
      most = cnt;
      any = true;
    }
  }
  anchor_meas = -1;
  for (int i = fv.m0; i < fv.m1; i++)
    if (fb.cam[i] == anchor_cam)
      anchor_meas = i;
  return any && anchor_meas >= 0;
}

// ======================================================================= triangulation
// feat/FeatureInitializer.cpp:30-112
inline int single_triangulation(const ovb_frame &fr, const ovb_feat_batch &fb, const ovb_opts &op, int f, V3 &p_FinA, V3 &p_FinG,
                                int &anchor_cam, int &anchor_clone) {
  FeatView fv = feat_view(fb, f);
  int ameas;
  pick_anchor(fb, fv, anchor_cam, ameas);
  anchor_clone = fb.clone[ameas];
  M3 A = m3zero();
  V3 b = v3(0, 0, 0);
  CamPose anc = cam_clone_pose(fr, anchor_cam, anchor_clone);
  const M3 &R_GtoA = anc.R;
  const V3 &p_AinG = anc.p;
  for (int i = fv.m0; i < fv.m1; i++) {
    CamPose ci = cam_clone_pose(fr, fb.cam[i], fb.clone[i]);
    M3 R_AtoCi = m3mulT(ci.R, R_GtoA);
    V3 p_CiinA = m3v(R_GtoA, vsub(ci.p, p_AinG));
    V3 b_i = v3((double)fb.uvn[2 * i], (double)fb.uvn[2 * i + 1], 1.0);
    b_i = m3Tv(R_AtoCi, b_i);
    double nb = vnorm(b_i);
    b_i = v3(b_i(0) / nb, b_i(1) / nb, b_i(2) / nb);
    M3 Bperp = skew(b_i);
    M3 Ai = m3Tmul(Bperp, Bperp);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++)
        A(r, c) += Ai(r, c);
    V3 Aip = m3v(Ai, p_CiinA);
    b = vadd(b, Aip);
  }
  V3 p_f = colpiv_qr_solve3(A, b);
  double condA = cond_sym3(A);
  if (std::fabs(condA) > op.max_cond_number)
    return OVB_FEAT_TRI_COND;
  if (p_f(2) < op.min_dist || p_f(2) > op.max_dist)
    return OVB_FEAT_TRI_DEPTH;
  if (std::isnan(vnorm(p_f)))
    return OVB_FEAT_TRI_NAN;
  p_FinA = p_f;
  p_FinG = vadd(m3Tv(R_GtoA, p_FinA), p_AinG);
  return OVB_FEAT_OK;
}

// feat/FeatureInitializer.cpp:114-195
inline int single_triangulation_1d(const ovb_frame &fr, const ovb_feat_batch &fb, const ovb_opts &op, int f, V3 &p_FinA, V3 &p_FinG,
                                   int &anchor_cam, int &anchor_clone) {
  FeatView fv = feat_view(fb, f);
  int ameas;
  pick_anchor(fb, fv, anchor_cam, ameas);
  anchor_clone = fb.clone[ameas];
  double A = 0.0, b = 0.0;
  CamPose anc = cam_clone_pose(fr, anchor_cam, anchor_clone);
  const M3 &R_GtoA = anc.R;
  const V3 &p_AinG = anc.p;
  V3 bearing_inA = v3((double)fb.uvn[2 * ameas], (double)fb.uvn[2 * ameas + 1], 1.0);
  double nba = vnorm(bearing_inA);
  bearing_inA = v3(bearing_inA(0) / nba, bearing_inA(1) / nba, bearing_inA(2) / nba);
  for (int i = fv.m0; i < fv.m1; i++) {
    if (i == ameas)
      continue;
    CamPose ci = cam_clone_pose(fr, fb.cam[i], fb.clone[i]);
    M3 R_AtoCi = m3mulT(ci.R, R_GtoA);
    V3 p_CiinA = m3v(R_GtoA, vsub(ci.p, p_AinG));
    V3 b_i = v3((double)fb.uvn[2 * i], (double)fb.uvn[2 * i + 1], 1.0);
    b_i = m3Tv(R_AtoCi, b_i);
    double nb = vnorm(b_i);
    b_i = v3(b_i(0) / nb, b_i(1) / nb, b_i(2) / nb);
    M3 Bperp = skew(b_i);
    V3 BperpBanchor = m3v(Bperp, bearing_inA);
    A += vdot(BperpBanchor, BperpBanchor);
    b += vdot(BperpBanchor, m3v(Bperp, p_CiinA));
  }
  double depth = b / A;
  V3 p_f = vscale(depth, bearing_inA);
  if (p_f(2) < op.min_dist || p_f(2) > op.max_dist)
    return OVB_FEAT_TRI_DEPTH;
  if (std::isnan(vnorm(p_f)))
    return OVB_FEAT_TRI_NAN;
  p_FinA = p_f;
  p_FinG = vadd(m3Tv(R_GtoA, p_FinA), p_AinG);
  return OVB_FEAT_OK;
}

// per-measurement geometry relative to the anchor, recomputed identically on every pass by the reference
// (feat/FeatureInitializer.cpp:245-254, 396-404)
struct RelPose {
  M3 R_AtoCi;
  V3 p_CiinA;
  V3 p_AinCi;
};
inline RelPose rel_pose(const CamPose &ci, const M3 &R_GtoA, const V3 &p_AinG) {
  RelPose rp;
  rp.R_AtoCi = m3mulT(ci.R, R_GtoA);
  rp.p_CiinA = m3v(R_GtoA, vsub(ci.p, p_AinG));
  rp.p_AinCi = m3v(m3neg(rp.R_AtoCi), rp.p_CiinA);
  return rp;
}

// feat/FeatureInitializer.cpp:377-423
inline double compute_error(const std::vector<RelPose> &rel, const float *uvn, double alpha, double beta, double rho) {
  double err = 0;
  for (size_t i = 0; i < rel.size(); i++) {
    const M3 &R = rel[i].R_AtoCi;
    const V3 &q = rel[i].p_AinCi;
    double hi1 = R(0, 0) * alpha + R(0, 1) * beta + R(0, 2) + rho * q(0);
    double hi2 = R(1, 0) * alpha + R(1, 1) * beta + R(1, 2) + rho * q(1);
    double hi3 = R(2, 0) * alpha + R(2, 1) * beta + R(2, 2) + rho * q(2);
    float z0 = (float)(hi1 / hi3), z1 = (float)(hi2 / hi3);
    float r0 = uvn[2 * i] - z0, r1 = uvn[2 * i + 1] - z1;
    float nrm = std::sqrt(r0 * r0 + r1 * r1); // res.norm() in float
    err += std::pow((double)nrm, 2);
  }
  return err;
}

struct GnTrace {
  int runs = 0;      // accepted steps
  int solves = 0;    // LM solves attempted
  double lam = 0;    // final lambda
  double cost = 0;   // final cost
};

// feat/FeatureInitializer.cpp:197-375
inline int single_gaussnewton(const ovb_frame &fr, const ovb_feat_batch &fb, const ovb_opts &op, int f, int anchor_cam, int anchor_clone,
                              V3 &p_FinA, V3 &p_FinG, GnTrace *trace) {
  FeatView fv = feat_view(fb, f);
  double rho = 1 / p_FinA(2);
  double alpha = p_FinA(0) / p_FinA(2);
  double beta = p_FinA(1) / p_FinA(2);
  double lam = op.init_lamda;
  double eps = 10000;
  int runs = 0;
  bool recompute = true;
  double Hess[3][3] = {{0}}, grad[3] = {0};
  CamPose anc = cam_clone_pose(fr, anchor_cam, anchor_clone);
  const M3 &R_GtoA = anc.R;
  const V3 &p_AinG = anc.p;
  std::vector<RelPose> rel;
  for (int i = fv.m0; i < fv.m1; i++)
    rel.push_back(rel_pose(cam_clone_pose(fr, fb.cam[i], fb.clone[i]), R_GtoA, p_AinG));
  const float *uvn = fb.uvn + 2 * fv.m0;
  double cost_old = compute_error(rel, uvn, alpha, beta, rho);
  int solves = 0;
  while (runs < op.max_runs && lam < op.max_lamda && eps > op.min_dx) {
    if (recompute) {
      for (int r = 0; r < 3; r++) {
        grad[r] = 0;
        for (int c = 0; c < 3; c++)
          Hess[r][c] = 0;
      }
      for (size_t i = 0; i < rel.size(); i++) {
        const M3 &R = rel[i].R_AtoCi;
        const V3 &q = rel[i].p_AinCi;
        double hi1 = R(0, 0) * alpha + R(0, 1) * beta + R(0, 2) + rho * q(0);
        double hi2 = R(1, 0) * alpha + R(1, 1) * beta + R(1, 2) + rho * q(1);
        double hi3 = R(2, 0) * alpha + R(2, 1) * beta + R(2, 2) + rho * q(2);
        double h3sq = std::pow(hi3, 2);
        double H[2][3];
        H[0][0] = (R(0, 0) * hi3 - hi1 * R(2, 0)) / h3sq;
        H[0][1] = (R(0, 1) * hi3 - hi1 * R(2, 1)) / h3sq;
        H[0][2] = (q(0) * hi3 - hi1 * q(2)) / h3sq;
        H[1][0] = (R(1, 0) * hi3 - hi2 * R(2, 0)) / h3sq;
        H[1][1] = (R(1, 1) * hi3 - hi2 * R(2, 1)) / h3sq;
        H[1][2] = (q(1) * hi3 - hi2 * q(2)) / h3sq;
        float z0 = (float)(hi1 / hi3), z1 = (float)(hi2 / hi3);
        float r0 = uvn[2 * i] - z0, r1 = uvn[2 * i + 1] - z1;
        double rd0 = (double)r0, rd1 = (double)r1;
        for (int a = 0; a < 3; a++) {
          grad[a] += H[0][a] * rd0 + H[1][a] * rd1;
          for (int c = 0; c < 3; c++)
            Hess[a][c] += H[0][a] * H[0][c] + H[1][a] * H[1][c];
        }
      }
    }
    M3 Hl;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++)
        Hl(r, c) = Hess[r][c];
    for (int r = 0; r < 3; r++)
      Hl(r, r) *= (1.0 + lam);
    V3 dx = colpiv_qr_solve3(Hl, v3(grad[0], grad[1], grad[2]));
    solves++;
    double cost = compute_error(rel, uvn, alpha + dx(0), beta + dx(1), rho + dx(2));
    if (cost <= cost_old && (cost_old - cost) / cost_old < op.min_dcost) {
      alpha += dx(0);
      beta += dx(1);
      rho += dx(2);
      eps = 0;
      cost_old = cost;
      break;
    }
    if (cost <= cost_old) {
      recompute = true;
      cost_old = cost;
      alpha += dx(0);
      beta += dx(1);
      rho += dx(2);
      runs++;
      lam = lam / op.lam_mult;
      eps = vnorm(dx);
    } else {
      recompute = false;
      lam = lam * op.lam_mult;
      continue;
    }
  }
  if (trace) {
    trace->runs = runs;
    trace->solves = solves;
    trace->lam = lam;
    trace->cost = cost_old;
  }
  p_FinA = v3(alpha / rho, beta / rho, 1 / rho);
  V3 q1, q2;
  householder_tangent3(p_FinA, q1, q2);
  double base_line_max = 0.0;
  for (size_t i = 0; i < rel.size(); i++) {
    const V3 &t = rel[i].p_CiinA;
    double a0 = vdot(q1, t), a1 = vdot(q2, t);
    double base_line = std::sqrt(a0 * a0 + a1 * a1);
    if (base_line > base_line_max)
      base_line_max = base_line;
  }
  if (p_FinA(2) < op.min_dist || p_FinA(2) > op.max_dist)
    return OVB_FEAT_GN_DEPTH;
  if ((vnorm(p_FinA) / base_line_max) > op.max_baseline)
    return OVB_FEAT_GN_BASELINE;
  if (std::isnan(vnorm(p_FinA)))
    return OVB_FEAT_GN_NAN;
  p_FinG = vadd(m3Tv(R_GtoA, p_FinA), p_AinG);
  return OVB_FEAT_OK;
}

// ======================================================================= per-feature Jacobian
// A state variable a feature's Jacobian touches: (covariance offset, size). Identity = offset.
struct Var {
  int off, size;
};

struct FeatJac {
  int rows = 0;           // 2*M before projection, 2*M-3 (or -1) after
  int nf = 3;             // columns of H_f
  std::vector<Var> order; // Hx_order (local column blocks in order)
  Mat Hf, Hx;             // rows x nf, rows x total_hx
  std::vector<double> res;
};

inline int find_var(const std::vector<Var> &order, int off) {
  int col = 0;
  for (const Var &v : order) {
    if (v.off == off)
      return col;
    col += v.size;
  }
  return -1;
}

// update/UpdaterHelper.cpp:32-190 — dpfg_dlambda (3 x nf) and the anchor-pose / anchor-extrinsics terms.
inline void jacobian_representation(const ovb_frame &fr, const ovb_opts &op, int rep, const V3 &p_FinG_in, const V3 &p_FinG_fej_in,
                                    const V3 &p_FinA_in, int anchor_cam, int anchor_clone, Mat &H_f, std::vector<Mat> &H_x,
                                    std::vector<Var> &x_order) {
  if (rep == OVB_REP_GLOBAL_3D) {
    H_f.resize_zero(3, 3);
    H_f(0, 0) = H_f(1, 1) = H_f(2, 2) = 1.0;
    return;
  }
  auto invdepth_jac = [](const V3 &p, Mat &J) {
    double rho = 1 / vnorm(p);
    double phi = std::acos(rho * p(2));
    double theta = std::atan2(p(1), p(0));
    double sin_th = std::sin(theta), cos_th = std::cos(theta), sin_phi = std::sin(phi), cos_phi = std::cos(phi);
    J.resize_zero(3, 3);
    J(0, 0) = -(1.0 / rho) * sin_th * sin_phi;
    J(0, 1) = (1.0 / rho) * cos_th * cos_phi;
    J(0, 2) = -(1.0 / (rho * rho)) * cos_th * sin_phi;
    J(1, 0) = (1.0 / rho) * cos_th * sin_phi;
    J(1, 1) = (1.0 / rho) * sin_th * cos_phi;
    J(1, 2) = -(1.0 / (rho * rho)) * sin_th * sin_phi;
    J(2, 0) = 0.0;
    J(2, 1) = -(1.0 / rho) * sin_phi;
    J(2, 2) = -(1.0 / (rho * rho)) * cos_phi;
  };
  if (rep == OVB_REP_GLOBAL_FULL_INVERSE_DEPTH) {
    V3 p = op.do_fej ? p_FinG_fej_in : p_FinG_in;
    invdepth_jac(p, H_f);
    return;
  }
  (void)p_FinG_in;
  // anchored
  M3 R_ItoC = load_m3(fr.cam_R + 9 * anchor_cam);
  V3 p_IinC = load_v3(fr.cam_p + 3 * anchor_cam);
  M3 R_GtoI = load_m3(fr.clone_R + 9 * anchor_clone);
  V3 p_IinG = load_v3(fr.clone_p + 3 * anchor_clone);
  V3 p_FinA = p_FinA_in;
  if (op.do_fej) {
    // p_FinG_best = R_GtoI' * R_ItoC' * (p_FinA - p_IinC) + p_IinG  (left-to-right: (R_GtoI' * R_ItoC') * v)
    M3 RtRt = m3mul(m3T(R_GtoI), m3T(R_ItoC));
    V3 p_FinG_best = vadd(m3v(RtRt, vsub(p_FinA_in, p_IinC)), p_IinG);
    R_GtoI = load_m3(fr.clone_R_fej + 9 * anchor_clone);
    p_IinG = load_v3(fr.clone_p_fej + 3 * anchor_clone);
    M3 RR = m3T(m3mul(m3T(R_GtoI), m3T(R_ItoC)));
    p_FinA = vadd(m3v(RR, vsub(p_FinG_best, p_IinG)), p_IinC);
  }
  M3 R_CtoG = m3mul(m3T(R_GtoI), m3T(R_ItoC));
  Mat H_anc(3, 6);
  {
    M3 blk = m3mul(m3neg(m3T(R_GtoI)), skew(m3Tv(R_ItoC, vsub(p_FinA, p_IinC))));
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        H_anc(r, c) = blk(r, c);
        H_anc(r, 3 + c) = (r == c) ? 1.0 : 0.0;
      }
  }
  x_order.push_back(Var{fr.clone_off[anchor_clone], 6});
  H_x.push_back(H_anc);
  if (op.do_calib_camera_pose) {
    Mat H_calib(3, 6);
    M3 blk = m3mul(m3neg(R_CtoG), skew(vsub(p_FinA, p_IinC)));
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        H_calib(r, c) = blk(r, c);
        H_calib(r, 3 + c) = -R_CtoG(r, c);
      }
    x_order.push_back(Var{fr.cam_ext_off[anchor_cam], 6});
    H_x.push_back(H_calib);
  }
  auto set_from_m3 = [](Mat &M, const M3 &a) {
    M.resize_zero(3, 3);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++)
        M(r, c) = a(r, c);
  };
  if (rep == OVB_REP_ANCHORED_3D) {
    set_from_m3(H_f, R_CtoG);
    return;
  }
  if (rep == OVB_REP_ANCHORED_FULL_INVERSE_DEPTH) {
    Mat d;
    invdepth_jac(p_FinA, d);
    M3 dd;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++)
        dd(r, c) = d(r, c);
    set_from_m3(H_f, m3mul(R_CtoG, dd));
    return;
  }
  if (rep == OVB_REP_ANCHORED_MSCKF_INVERSE_DEPTH) {
    double alpha = p_FinA(0) / p_FinA(2);
    double beta = p_FinA(1) / p_FinA(2);
    double rho = 1 / p_FinA(2);
    M3 dd = m3zero();
    dd(0, 0) = (1.0 / rho);
    dd(0, 2) = -(1.0 / (rho * rho)) * alpha;
    dd(1, 1) = (1.0 / rho);
    dd(1, 2) = -(1.0 / (rho * rho)) * beta;
    dd(2, 2) = -(1.0 / (rho * rho));
    set_from_m3(H_f, m3mul(R_CtoG, dd));
    return;
  }
  // ANCHORED_INVERSE_DEPTH_SINGLE
  {
    double rho = 1.0 / p_FinA(2);
    V3 bearing = vscale(rho, p_FinA);
    V3 d = vscale(-(1.0 / (rho * rho)), bearing);
    V3 hf = m3v(R_CtoG, d);
    H_f.resize_zero(3, 1);
    for (int r = 0; r < 3; r++)
      H_f(r, 0) = hf(r);
  }
}

inline bool is_relative(int rep) {
  return rep == OVB_REP_ANCHORED_3D || rep == OVB_REP_ANCHORED_FULL_INVERSE_DEPTH || rep == OVB_REP_ANCHORED_MSCKF_INVERSE_DEPTH ||
         rep == OVB_REP_ANCHORED_INVERSE_DEPTH_SINGLE;
}

// update/UpdaterHelper.cpp:192-424 get_feature_jacobian_full
inline void feature_jacobian_full(const ovb_frame &fr, const ovb_feat_batch &fb, const ovb_opts &op, int f, int rep, const V3 &p_FinG_in,
                                  const V3 &p_FinG_fej_in, const V3 &p_FinA, int anchor_cam, int anchor_clone, FeatJac &J) {
  FeatView fv = feat_view(fb, f);
  int total_meas = fv.m1 - fv.m0;
  // ---- column map (:201-261)
  int total_hx = 0;
  J.order.clear();
  for (int key : fv.keys) {
    if (op.do_calib_camera_pose) {
      if (find_var(J.order, fr.cam_ext_off[key]) < 0) {
        J.order.push_back(Var{fr.cam_ext_off[key], 6});
        total_hx += 6;
      }
    }
    if (op.do_calib_camera_intrinsics) {
      if (find_var(J.order, fr.cam_intr_off[key]) < 0) {
        J.order.push_back(Var{fr.cam_intr_off[key], 8});
        total_hx += 8;
      }
    }
    for (int i = fv.m0; i < fv.m1; i++) {
      if (fb.cam[i] != key)
        continue;
      int off = fr.clone_off[fb.clone[i]];
      if (find_var(J.order, off) < 0) {
        J.order.push_back(Var{off, 6});
        total_hx += 6;
      }
    }
  }
  if (is_relative(rep)) {
    int off = fr.clone_off[anchor_clone];
    if (find_var(J.order, off) < 0) {
      J.order.push_back(Var{off, 6});
      total_hx += 6;
    }
    if (op.do_calib_camera_pose) {
      int eo = fr.cam_ext_off[anchor_cam];
      if (find_var(J.order, eo) < 0) {
        J.order.push_back(Var{eo, 6});
        total_hx += 6;
      }
    }
  }
  // ---- feature position (:266-287)
  V3 p_FinG = p_FinG_in;
  if (is_relative(rep)) {
    M3 R_ItoC = load_m3(fr.cam_R + 9 * anchor_cam);
    V3 p_IinC = load_v3(fr.cam_p + 3 * anchor_cam);
    M3 R_GtoI = load_m3(fr.clone_R + 9 * anchor_clone);
    V3 p_IinG = load_v3(fr.clone_p + 3 * anchor_clone);
    M3 RtRt = m3mul(m3T(R_GtoI), m3T(R_ItoC));
    p_FinG = vadd(m3v(RtRt, vsub(p_FinA, p_IinC)), p_IinG);
  }
  V3 p_FinG_fej = p_FinG_fej_in;
  if (is_relative(rep))
    p_FinG_fej = p_FinG;
  // ---- allocate (:293-297)
  int jacobsize = (rep != OVB_REP_ANCHORED_INVERSE_DEPTH_SINGLE) ? 3 : 1;
  J.rows = 2 * total_meas;
  J.nf = jacobsize;
  J.res.assign(J.rows, 0.0);
  J.Hf.resize_zero(J.rows, jacobsize);
  J.Hx.resize_zero(J.rows, total_hx);
  Mat dpfg_dlambda;
  std::vector<Mat> dpfg_dx;
  std::vector<Var> dpfg_dx_order;
  jacobian_representation(fr, op, rep, p_FinG, p_FinG_fej, p_FinA, anchor_cam, anchor_clone, dpfg_dlambda, dpfg_dx, dpfg_dx_order);
  // ---- measurements (:313-423)
  int c = 0;
  for (int i = fv.m0; i < fv.m1; i++, c++) {
    int cam = fb.cam[i], cl = fb.clone[i];
    M3 R_ItoC = load_m3(fr.cam_R + 9 * cam);
    V3 p_IinC = load_v3(fr.cam_p + 3 * cam);
    M3 R_GtoIi = load_m3(fr.clone_R + 9 * cl);
    V3 p_IiinG = load_v3(fr.clone_p + 3 * cl);
    V3 p_FinIi = m3v(R_GtoIi, vsub(p_FinG, p_IiinG));
    V3 p_FinCi = vadd(m3v(R_ItoC, p_FinIi), p_IinC);
    double un = p_FinCi(0) / p_FinCi(2), vn = p_FinCi(1) / p_FinCi(2);
    double ud, vd;
    distort_d(fr.cam_model[cam], fr.cam_intr + 8 * cam, un, vn, ud, vd);
    J.res[2 * c] = (double)fb.uv[2 * i] - ud;
    J.res[2 * c + 1] = (double)fb.uv[2 * i + 1] - vd;
    if (op.do_fej) {
      R_GtoIi = load_m3(fr.clone_R_fej + 9 * cl);
      p_IiinG = load_v3(fr.clone_p_fej + 3 * cl);
      p_FinIi = m3v(R_GtoIi, vsub(p_FinG_fej, p_IiinG));
      p_FinCi = vadd(m3v(R_ItoC, p_FinIi), p_IinC);
    }
    double dz_dzn[4], dz_dzeta[16];
    distort_jacobian(fr.cam_model[cam], fr.cam_intr + 8 * cam, un, vn, dz_dzn, dz_dzeta);
    double dzn_dpfc[2][3] = {{1 / p_FinCi(2), 0, -p_FinCi(0) / (p_FinCi(2) * p_FinCi(2))},
                             {0, 1 / p_FinCi(2), -p_FinCi(1) / (p_FinCi(2) * p_FinCi(2))}};
    M3 dpfc_dpfg = m3mul(R_ItoC, R_GtoIi);
    M3 dpfc_dth = m3mul(R_ItoC, skew(p_FinIi));
    double dpfc_dclone[3][6];
    for (int r = 0; r < 3; r++)
      for (int k = 0; k < 3; k++) {
        dpfc_dclone[r][k] = dpfc_dth(r, k);
        dpfc_dclone[r][3 + k] = -dpfc_dpfg(r, k);
      }
    double dz_dpfc[2][3], dz_dpfg[2][3];
    for (int r = 0; r < 2; r++)
      for (int k = 0; k < 3; k++)
        dz_dpfc[r][k] = dz_dzn[2 * r + 0] * dzn_dpfc[0][k] + dz_dzn[2 * r + 1] * dzn_dpfc[1][k];
    for (int r = 0; r < 2; r++)
      for (int k = 0; k < 3; k++)
        dz_dpfg[r][k] = (dz_dpfc[r][0] * dpfc_dpfg(0, k) + dz_dpfc[r][1] * dpfc_dpfg(1, k)) + dz_dpfc[r][2] * dpfc_dpfg(2, k);
    for (int r = 0; r < 2; r++)
      for (int k = 0; k < jacobsize; k++)
        J.Hf(2 * c + r, k) = (dz_dpfg[r][0] * dpfg_dlambda(0, k) + dz_dpfg[r][1] * dpfg_dlambda(1, k)) + dz_dpfg[r][2] * dpfg_dlambda(2, k);
    int ccol = find_var(J.order, fr.clone_off[cl]);
    for (int r = 0; r < 2; r++)
      for (int k = 0; k < 6; k++)
        J.Hx(2 * c + r, ccol + k) = (dz_dpfc[r][0] * dpfc_dclone[0][k] + dz_dpfc[r][1] * dpfc_dclone[1][k]) + dz_dpfc[r][2] * dpfc_dclone[2][k];
    for (size_t e = 0; e < dpfg_dx_order.size(); e++) {
      int ecol = find_var(J.order, dpfg_dx_order[e].off);
      for (int r = 0; r < 2; r++)
        for (int k = 0; k < dpfg_dx_order[e].size; k++)
          J.Hx(2 * c + r, ecol + k) +=
              (dz_dpfg[r][0] * dpfg_dx[e](0, k) + dz_dpfg[r][1] * dpfg_dx[e](1, k)) + dz_dpfg[r][2] * dpfg_dx[e](2, k);
    }
    if (op.do_calib_camera_pose) {
      M3 sk = skew(vsub(p_FinCi, p_IinC));
      double dpfc_dcalib[3][6];
      for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) {
          dpfc_dcalib[r][k] = sk(r, k);
          dpfc_dcalib[r][3 + k] = (r == k) ? 1.0 : 0.0;
        }
      int ecol = find_var(J.order, fr.cam_ext_off[cam]);
      for (int r = 0; r < 2; r++)
        for (int k = 0; k < 6; k++)
          J.Hx(2 * c + r, ecol + k) += (dz_dpfc[r][0] * dpfc_dcalib[0][k] + dz_dpfc[r][1] * dpfc_dcalib[1][k]) + dz_dpfc[r][2] * dpfc_dcalib[2][k];
    }
    if (op.do_calib_camera_intrinsics) {
      int icol = find_var(J.order, fr.cam_intr_off[cam]);
      for (int r = 0; r < 2; r++)
        for (int k = 0; k < 8; k++)
          J.Hx(2 * c + r, icol + k) = dz_dzeta[8 * r + k];
    }
  }
}

// update/UpdaterHelper.cpp:426-454
inline void nullspace_project_inplace(FeatJac &J) {
  int rows = J.rows, nf = J.nf, w = J.Hx.c;
  for (int n = 0; n < nf; ++n) {
    for (int m = rows - 1; m > n; m--) {
      Givens g = make_givens(J.Hf(m - 1, n), J.Hf(m, n));
      for (int k = n; k < nf; k++)
        apply_givens(g, J.Hf(m - 1, k), J.Hf(m, k));
      for (int k = 0; k < w; k++)
        apply_givens(g, J.Hx(m - 1, k), J.Hx(m, k));
      apply_givens(g, J.res[m - 1], J.res[m]);
    }
  }
  Mat Hx2(rows - nf, w);
  for (int k = 0; k < w; k++)
    for (int i = nf; i < rows; i++)
      Hx2(i - nf, k) = J.Hx(i, k);
  J.Hx = Hx2;
  J.res.erase(J.res.begin(), J.res.begin() + nf);
  J.rows = rows - nf;
}

// state/StateHelper.cpp:226-254
inline Mat get_marginal_covariance(const double *P, int N, const std::vector<Var> &vars) {
  int sz = 0;
  for (const Var &v : vars)
    sz += v.size;
  Mat S(sz, sz);
  int ii = 0;
  for (const Var &vi : vars) {
    int kk = 0;
    for (const Var &vk : vars) {
      for (int r = 0; r < vi.size; r++)
        for (int c = 0; c < vk.size; c++)
          S(ii + r, kk + c) = P[(size_t)(vi.off + r) * N + (vk.off + c)];
      kk += vk.size;
    }
    ii += vi.size;
  }
  return S;
}

// chi² of a projected feature: update/UpdaterMSCKF.cpp:209-212
inline double feature_chi2(const double *P, int N, const FeatJac &J, double sigma_pix_sq, bool *spd) {
  Mat P_marg = get_marginal_covariance(P, N, J.order);
  Mat HP = matmul(J.Hx, P_marg);
  Mat S = matmul_nt(HP, J.Hx);
  for (int i = 0; i < S.r; i++)
    S(i, i) += sigma_pix_sq * 1.0;
  // S.llt(): Eigen's LLT reads the lower triangle by default; S is symmetric up to rounding. Use the lower one.
  Mat L = S;
  int n = L.r;
  bool ok = true;
  for (int j = 0; j < n; j++) {
    double d = L(j, j);
    for (int k = 0; k < j; k++)
      d -= L(j, k) * L(j, k);
    if (!(d > 0.0)) {
      ok = false;
      break;
    }
    d = std::sqrt(d);
    L(j, j) = d;
    for (int i = j + 1; i < n; i++) {
      double v = L(i, j);
      for (int k = 0; k < j; k++)
        v -= L(i, k) * L(j, k);
      L(i, j) = v / d;
    }
  }
  if (spd)
    *spd = ok;
  if (!ok)
    return std::numeric_limits<double>::quiet_NaN();
  std::vector<double> x = J.res;
  llt_solve_vec(L, x.data());
  double chi2 = 0;
  for (int i = 0; i < n; i++)
    chi2 += J.res[i] * x[i];
  return chi2;
}

// ======================================================================= compression + EKF
// update/UpdaterHelper.cpp:456-487 — column-major Givens sweep exactly as the reference (this is the CPU hot loop)
inline void measurement_compress_inplace(Mat &H, std::vector<double> &res) {
  if (H.r <= H.c)
    return;
  int rows = H.r, cols = H.c;
  for (int n = 0; n < cols; n++) {
    for (int m = rows - 1; m > n; m--) {
      Givens g = make_givens(H(m - 1, n), H(m, n));
      for (int k = n; k < cols; k++)
        apply_givens(g, H(m - 1, k), H(m, k));
      apply_givens(g, res[m - 1], res[m]);
    }
  }
  int r = std::min(rows, cols);
  Mat H2(r, cols);
  for (int k = 0; k < cols; k++)
    for (int i = 0; i < r; i++)
      H2(i, k) = H(i, k);
  H = H2;
  res.resize(r);
}

// state/StateHelper.cpp:116-197. P is N x N row-major (symmetric). Rdiag: the diagonal of R (R is diagonal at every
// call site of the path: UpdaterMSCKF.cpp:282, UpdaterSLAM.cpp:444). Returns ovb_status; dx has length N.
inline int ekf_update(double *P, int N, const std::vector<Var> &H_order, const Mat &H, const std::vector<double> &res,
                      const std::vector<double> &Rdiag, double *dx, int *neg_index) {
  int r = H.r;
  Mat M_a(N, r);
  std::vector<int> H_id;
  int cur = 0;
  for (const Var &v : H_order) {
    H_id.push_back(cur);
    cur += v.size;
  }
  // M_a = P[:, cols] * H'  accumulated per measuring variable (:137-146)
  for (size_t i = 0; i < H_order.size(); i++) {
    const Var &mv = H_order[i];
    for (int k = 0; k < mv.size; k++) {
      for (int j = 0; j < r; j++) {
        double h = H(j, H_id[i] + k);
        if (h == 0.0)
          continue;
        double *col = &M_a.d[(size_t)j * N];
        const double *pc = P + (size_t)(mv.off + k) * N; // P symmetric: column == row
        for (int a = 0; a < N; a++)
          col[a] += pc[a] * h;
      }
    }
  }
  Mat P_small = get_marginal_covariance(P, N, H_order);
  Mat HP = matmul(H, P_small);
  Mat S = matmul_nt(HP, H);
  for (int i = 0; i < r; i++)
    S(i, i) += Rdiag[i];
  // LLT of selfadjointView<Upper>, Sinv = S^-1 by solving against I (:160-161)
  Mat L = S;
  if (!llt_from_upper(L))
    return OVB_ERR_NOT_SPD;
  Mat Sinv(r, r);
  std::vector<double> e(r);
  for (int j = 0; j < r; j++) {
    std::fill(e.begin(), e.end(), 0.0);
    e[j] = 1.0;
    llt_solve_vec(L, e.data());
    for (int i = 0; i < r; i++)
      Sinv(i, j) = e[i];
  }
  // K = M_a * Sinv.selfadjointView<Upper>()
  Mat SinvU(r, r);
  for (int i = 0; i < r; i++)
    for (int j = 0; j < r; j++)
      SinvU(i, j) = (i <= j) ? Sinv(i, j) : Sinv(j, i);
  Mat K = matmul(M_a, SinvU);
  // P.upper -= K * M_a' ; mirror (:166-167)
  Mat KMt = matmul_nt(K, M_a);
  for (int i = 0; i < N; i++)
    for (int j = i; j < N; j++) {
      double v = P[(size_t)i * N + j] - KMt(i, j);
      P[(size_t)i * N + j] = v;
      P[(size_t)j * N + i] = v;
    }
  if (neg_index)
    *neg_index = -1;
  int status = OVB_OK;
  for (int i = 0; i < N; i++)
    if (P[(size_t)i * N + i] < 0.0) {
      if (neg_index && *neg_index < 0)
        *neg_index = i;
      status = OVB_ERR_NEG_DIAG;
    }
  for (int a = 0; a < N; a++) {
    double s = 0;
    for (int j = 0; j < r; j++)
      s += K(a, j) * res[j];
    dx[a] = s;
  }
  return status;
}

// state/StateHelper.cpp:36-114. Phi p x q row-major, Q p x p row-major (upper triangle used).
inline int ekf_propagation(double *P, int N, int new_off, int p, const std::vector<Var> &order_old, const double *Phi, const double *Q) {
  int q = 0;
  std::vector<int> Phi_id;
  for (const Var &v : order_old) {
    Phi_id.push_back(q);
    q += v.size;
  }
  Mat Cov_PhiT(N, p);
  for (size_t i = 0; i < order_old.size(); i++) {
    const Var &v = order_old[i];
    for (int k = 0; k < v.size; k++)
      for (int j = 0; j < p; j++) {
        double ph = Phi[(size_t)j * q + Phi_id[i] + k];
        for (int a = 0; a < N; a++)
          Cov_PhiT(a, j) += P[(size_t)a * N + v.off + k] * ph;
      }
  }
  Mat PCP(p, p);
  for (int i = 0; i < p; i++)
    for (int j = 0; j < p; j++)
      PCP(i, j) = (i <= j) ? Q[(size_t)i * p + j] : Q[(size_t)j * p + i];
  for (size_t i = 0; i < order_old.size(); i++) {
    const Var &v = order_old[i];
    for (int k = 0; k < v.size; k++)
      for (int j = 0; j < p; j++) {
        double c = Cov_PhiT(v.off + k, j);
        for (int a = 0; a < p; a++)
          PCP(a, j) += Phi[(size_t)a * q + Phi_id[i] + k] * c;
      }
  }
  for (int a = 0; a < N; a++)
    for (int j = 0; j < p; j++)
      P[(size_t)(new_off + j) * N + a] = Cov_PhiT(a, j);
  for (int a = 0; a < N; a++)
    for (int j = 0; j < p; j++)
      P[(size_t)a * N + new_off + j] = Cov_PhiT(a, j);
  for (int i = 0; i < p; i++)
    for (int j = 0; j < p; j++)
      P[(size_t)(new_off + i) * N + new_off + j] = PCP(i, j);
  for (int i = 0; i < N; i++)
    if (P[(size_t)i * N + i] < 0.0)
      return OVB_ERR_NEG_DIAG;
  return OVB_OK;
}

// state/StateHelper.cpp:341-391 (+ :604-615). Pin is N x N, Pout is (N+size) x (N+size), both row-major.
inline void cov_clone(const double *Pin, int N, int old_off, int size, const double *dnc_dt, int dt_off, double *Pout) {
  int N2 = N + size;
  for (int i = 0; i < N2 * N2; i++)
    Pout[i] = 0.0;
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++)
      Pout[(size_t)i * N2 + j] = Pin[(size_t)i * N + j];
  for (int i = 0; i < size; i++)
    for (int j = 0; j < size; j++)
      Pout[(size_t)(N + i) * N2 + N + j] = Pin[(size_t)(old_off + i) * N + old_off + j];
  for (int i = 0; i < N; i++)
    for (int j = 0; j < size; j++) {
      Pout[(size_t)i * N2 + N + j] = Pin[(size_t)i * N + old_off + j];
      Pout[(size_t)(N + j) * N2 + i] = Pin[(size_t)(old_off + j) * N + i];
    }
  if (dnc_dt) {
    // _Cov.block(0, pose.id, rows, 6) += _Cov.block(0, dt.id, rows, 1) * dnc_dt'
    for (int i = 0; i < N2; i++)
      for (int j = 0; j < size; j++)
        Pout[(size_t)i * N2 + N + j] += Pout[(size_t)i * N2 + dt_off] * dnc_dt[j];
    // _Cov.block(pose.id, 0, 6, rows) += dnc_dt * _Cov.block(dt.id, 0, 1, rows)
    for (int i = 0; i < size; i++)
      for (int j = 0; j < N2; j++)
        Pout[(size_t)(N + i) * N2 + j] += dnc_dt[i] * Pout[(size_t)dt_off * N2 + j];
  }
}

// state/StateHelper.cpp:271-339
inline void cov_marginalize(const double *Pin, int N, int off, int size, double *Pout) {
  int N2 = N - size;
  auto src = [&](int i) { return i < off ? i : i + size; };
  for (int i = 0; i < N2; i++)
    for (int j = 0; j < N2; j++) {
      // the reference copies P(x1,x2) and mirrors it into P(x2,x1) (:303-306)
      int si = src(i), sj = src(j);
      if (i >= off && j < off)
        Pout[(size_t)i * N2 + j] = Pin[(size_t)sj * N + si];
      else
        Pout[(size_t)i * N2 + j] = Pin[(size_t)si * N + sj];
    }
}

// ======================================================================= the whole update
struct UpdateDump {   // optional stage outputs for parity tests
  std::vector<Var> order_big;       // Hx_order_big
  Mat H_big;                        // stacked, before compression (ct_meas x ct_jacob)
  std::vector<double> res_big;
  Mat H_cmp;                        // after compression
  std::vector<double> res_cmp;
  double t_tri = 0, t_sys = 0, t_cmp = 0, t_upd = 0; // seconds, same stopwatch points as UpdaterMSCKF.cpp:65-294
};

inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Steps 2-6 of UpdaterMSCKF::update (update/UpdaterMSCKF.cpp:98-285). P (N x N row-major) is updated in place.
inline int msckf_update(const ovb_frame &fr, const ovb_feat_batch &fb, const ovb_opts &op, const double *chi2_table, double *P, int N,
                        ovb_feat_out *out, double *dx, ovb_stats *stats, UpdateDump *dump) {
  int F = fb.n_feats;
  double sigma_pix_sq = std::pow(op.sigma_pix, 2);
  std::vector<int> status(F, OVB_FEAT_OK);
  std::vector<V3> pA(F), pG(F);
  std::vector<int> acam(F, -1), aclone(F, -1);
  double T0 = now_s();
  // 1/3. count + triangulate (:75-142)
  for (int f = 0; f < F; f++) {
    pA[f] = pG[f] = v3(std::nan(""), std::nan(""), std::nan(""));
    if (fb.meas_off[f + 1] - fb.meas_off[f] < 2) {
      status[f] = OVB_FEAT_FEW_MEAS;
      continue;
    }
    int st = op.triangulate_1d ? single_triangulation_1d(fr, fb, op, f, pA[f], pG[f], acam[f], aclone[f])
                               : single_triangulation(fr, fb, op, f, pA[f], pG[f], acam[f], aclone[f]);
    // NOTE the reference still runs single_gaussnewton on a failed triangulation's stale p_FinA and then drops the
    // feature (:121-140); the outcome (dropped) is the same, so it is skipped here.
    if (st == OVB_FEAT_OK && op.refine_features)
      st = single_gaussnewton(fr, fb, op, f, acam[f], aclone[f], pA[f], pG[f], nullptr);
    status[f] = st;
  }
  double T1 = now_s();
  // 4. per-feature system (:145-256)
  size_t max_meas_size = 0;
  for (int f = 0; f < F; f++)
    if (status[f] == OVB_FEAT_OK)
      max_meas_size += 2 * (size_t)(fb.meas_off[f + 1] - fb.meas_off[f]);
  int max_hx_size = N;
  std::vector<double> res_big(max_meas_size, 0.0);
  Mat Hx_big((int)max_meas_size, max_hx_size);
  std::vector<Var> Hx_order_big;
  int ct_jacob = 0, ct_meas = 0, used = 0;
  int rep = op.feat_rep;
  if (rep == OVB_REP_ANCHORED_INVERSE_DEPTH_SINGLE)
    rep = OVB_REP_ANCHORED_MSCKF_INVERSE_DEPTH; // :180-183
  std::vector<double> chi2s(F, std::nan(""));
  for (int f = 0; f < F; f++) {
    if (status[f] != OVB_FEAT_OK)
      continue;
    FeatJac J;
    feature_jacobian_full(fr, fb, op, f, rep, pG[f], pG[f], pA[f], acam[f], aclone[f], J);
    nullspace_project_inplace(J);
    bool spd = true;
    double chi2 = feature_chi2(P, N, J, sigma_pix_sq, &spd);
    chi2s[f] = chi2;
    double chi2_check = chi2_table[std::min(J.rows, OVB_CHI2_TABLE_LEN - 1)];
    if (!(chi2 <= op.chi2_multipler * chi2_check)) { // reference: reject if chi2 > thr (NaN passes there; treat as reject)
      status[f] = OVB_FEAT_CHI2;
      continue;
    }
    int ct_hx = 0;
    for (const Var &var : J.order) {
      int col = find_var(Hx_order_big, var.off);
      if (col < 0) {
        col = ct_jacob;
        Hx_order_big.push_back(var);
        ct_jacob += var.size;
      }
      for (int k = 0; k < var.size; k++)
        for (int i = 0; i < J.rows; i++)
          Hx_big(ct_meas + i, col + k) = J.Hx(i, ct_hx + k);
      ct_hx += var.size;
    }
    for (int i = 0; i < J.rows; i++)
      res_big[ct_meas + i] = J.res[i];
    ct_meas += J.rows;
    used++;
  }
  double T2 = now_s();
  if (out) {
    for (int f = 0; f < F; f++) {
      if (out->status)
        out->status[f] = status[f];
      for (int k = 0; k < 3; k++) {
        if (out->p_FinA)
          out->p_FinA[3 * f + k] = pA[f](k);
        if (out->p_FinG)
          out->p_FinG[3 * f + k] = pG[f](k);
      }
      if (out->anchor_cam)
        out->anchor_cam[f] = acam[f];
      if (out->anchor_clone)
        out->anchor_clone[f] = aclone[f];
      if (out->chi2)
        out->chi2[f] = chi2s[f];
    }
  }
  if (stats) {
    stats->n_feats_in = F;
    stats->n_feats_used = used;
    stats->rows_stacked = ct_meas;
    stats->cols_stacked = ct_jacob;
    stats->rows_update = 0;
    stats->neg_diag_index = -1;
    stats->ms_total = 0;
  }
  for (int i = 0; i < N; i++)
    dx[i] = 0.0;
  if (ct_meas < 1)
    return OVB_OK;
  // conservativeResize (:271-272)
  Mat H(ct_meas, ct_jacob);
  for (int k = 0; k < ct_jacob; k++)
    for (int i = 0; i < ct_meas; i++)
      H(i, k) = Hx_big(i, k);
  res_big.resize(ct_meas);
  if (dump) {
    dump->order_big = Hx_order_big;
    dump->H_big = H;
    dump->res_big = res_big;
  }
  // 5. compress (:275)
  measurement_compress_inplace(H, res_big);
  double T3 = now_s();
  if (dump) {
    dump->H_cmp = H;
    dump->res_cmp = res_big;
  }
  // 6. update (:282-285)
  std::vector<double> Rdiag(H.r, sigma_pix_sq);
  int neg = -1;
  int st = ekf_update(P, N, Hx_order_big, H, res_big, Rdiag, dx, &neg);
  double T4 = now_s();
  if (stats) {
    stats->rows_update = H.r;
    stats->neg_diag_index = neg;
    stats->ms_total = (float)((T4 - T0) * 1e3);
  }
  if (dump) {
    dump->t_tri = T1 - T0;
    dump->t_sys = T2 - T1;
    dump->t_cmp = T3 - T2;
    dump->t_upd = T4 - T3;
  }
  return st;
}

// Steps 4-5 of UpdaterSLAM::update (update/UpdaterSLAM.cpp:310-470). Landmarks live in the state: value / FEJ value
// come from `lm` (Landmark::get_xyz), H_xf = [H_x, H_f], no nullspace projection — except for ANCHORED_INVERSE_DEPTH_SINGLE,
// whose landmark is the 1-wide depth: H_xf = [H_x, dz/drho] and the two bearing columns are projected out (:344-353) —,
// no compression, ONE EKFUpdate with the block-scaled identity R_big (:444). P updated in place.
inline int slam_update(const ovb_frame &fr, const ovb_feat_batch &fb, const ovb_landmarks &lm, const ovb_opts &op, const double *chi2_table,
                       double *P, int N, ovb_feat_out *out, double *dx, ovb_stats *stats, UpdateDump *dump) {
  const int F = fb.n_feats;
  int rep = op.feat_rep;
  const bool single = (rep == OVB_REP_ANCHORED_INVERSE_DEPTH_SINGLE);
  if (single)
    rep = OVB_REP_ANCHORED_MSCKF_INVERSE_DEPTH; // :327-329
  const int lm_size = single ? 1 : 3;
  std::vector<int> status(F, OVB_FEAT_OK);
  std::vector<double> chi2s(F, std::nan(""));
  size_t max_meas_size = 0;
  for (int f = 0; f < F; f++)
    max_meas_size += 2 * (size_t)(fb.meas_off[f + 1] - fb.meas_off[f]);
  std::vector<double> res_big(max_meas_size, 0.0), R_big(max_meas_size, 1.0);
  Mat Hx_big((int)max_meas_size, N);
  std::vector<Var> Hx_order_big;
  int ct_jacob = 0, ct_meas = 0, used = 0;
  double T0 = now_s();
  for (int f = 0; f < F; f++) {
    if (fb.meas_off[f + 1] - fb.meas_off[f] < (single ? 2 : 1)) { // :278-290 (too few measurements: dropped before the update)
      status[f] = OVB_FEAT_FEW_MEAS;
      continue;
    }
    // :333-341 — the landmark's value and FEJ value in the frame its representation lives in
    V3 val = v3(lm.value[3 * f], lm.value[3 * f + 1], lm.value[3 * f + 2]);
    V3 val_fej = v3(lm.value_fej[3 * f], lm.value_fej[3 * f + 1], lm.value_fej[3 * f + 2]);
    V3 p_FinG = val, p_FinG_fej = val_fej, p_FinA = val;
    int acam = -1, aclone = -1;
    if (is_relative(rep)) {
      acam = lm.anchor_cam[f];
      aclone = lm.anchor_clone[f];
    }
    FeatJac J;
    feature_jacobian_full(fr, fb, op, f, rep, p_FinG, p_FinG_fej, p_FinA, acam, aclone, J);
    // :354-361 — H_xf = [H_x, H_f], order += landmark
    FeatJac Jxf;
    Jxf.rows = J.rows;
    Jxf.order = J.order;
    Jxf.order.push_back(Var{lm.lm_off[f], lm_size});
    Jxf.res = J.res;
    Jxf.Hx.resize_zero(J.rows, J.Hx.c + lm_size);
    for (int i = 0; i < J.rows; i++) {
      for (int k = 0; k < J.Hx.c; k++)
        Jxf.Hx(i, k) = J.Hx(i, k);
      for (int k = 0; k < lm_size; k++)
        Jxf.Hx(i, J.Hx.c + k) = J.Hf(i, single ? 2 : k);
    }
    if (single) { // :344-353 — project the bearing portion (the first two columns of H_f) out of [H_x, dz/drho] and res
      Jxf.nf = 2;
      Jxf.Hf.resize_zero(J.rows, 2);
      for (int i = 0; i < J.rows; i++)
        for (int k = 0; k < 2; k++)
          Jxf.Hf(i, k) = J.Hf(i, k);
      nullspace_project_inplace(Jxf);
    }
    // :389-420 — chi² gate with the per-class noise and multiplier
    const double sigma_pix = lm.sigma_pix ? lm.sigma_pix[f] : op.sigma_pix;
    const double sigma_pix_sq = std::pow(sigma_pix, 2);
    const double mult = lm.chi2_multipler ? lm.chi2_multipler[f] : op.chi2_multipler;
    bool spd = true;
    double chi2 = feature_chi2(P, N, Jxf, sigma_pix_sq, &spd);
    chi2s[f] = chi2;
    double chi2_check = chi2_table[std::min(Jxf.rows, OVB_CHI2_TABLE_LEN - 1)];
    if (!(chi2 <= mult * chi2_check)) {
      status[f] = OVB_FEAT_CHI2;
      continue;
    }
    // :424-447 — append with the first-seen column map
    int ct_hx = 0;
    for (const Var &var : Jxf.order) {
      int col = find_var(Hx_order_big, var.off);
      if (col < 0) {
        col = ct_jacob;
        Hx_order_big.push_back(var);
        ct_jacob += var.size;
      }
      for (int k = 0; k < var.size; k++)
        for (int i = 0; i < Jxf.rows; i++)
          Hx_big(ct_meas + i, col + k) = Jxf.Hx(i, ct_hx + k);
      ct_hx += var.size;
    }
    for (int i = 0; i < Jxf.rows; i++) {
      res_big[ct_meas + i] = Jxf.res[i];
      R_big[ct_meas + i] = sigma_pix_sq;
    }
    ct_meas += Jxf.rows;
    used++;
  }
  double T1 = now_s();
  if (out) {
    for (int f = 0; f < F; f++) {
      if (out->status)
        out->status[f] = status[f];
      if (out->chi2)
        out->chi2[f] = chi2s[f];
    }
  }
  if (stats) {
    stats->n_feats_in = F;
    stats->n_feats_used = used;
    stats->rows_stacked = ct_meas;
    stats->cols_stacked = ct_jacob;
    stats->rows_update = ct_meas;
    stats->neg_diag_index = -1;
    stats->ms_total = 0;
  }
  for (int i = 0; i < N; i++)
    dx[i] = 0.0;
  if (ct_meas < 1)
    return OVB_OK;
  Mat H(ct_meas, ct_jacob);
  for (int k = 0; k < ct_jacob; k++)
    for (int i = 0; i < ct_meas; i++)
      H(i, k) = Hx_big(i, k);
  res_big.resize(ct_meas);
  R_big.resize(ct_meas);
  if (dump) {
    dump->order_big = Hx_order_big;
    dump->H_big = H;
    dump->res_big = res_big;
    dump->H_cmp = H;
    dump->res_cmp = R_big; // the SLAM path never compresses: this slot carries diag(R_big) instead
    dump->t_sys = T1 - T0;
  }
  int neg = -1;
  int st = ekf_update(P, N, Hx_order_big, H, res_big, R_big, dx, &neg);
  if (stats) {
    stats->neg_diag_index = neg;
    stats->ms_total = (float)((now_s() - T0) * 1e3);
  }
  if (dump)
    dump->t_upd = now_s() - T1;
  return st;
}

// StateHelper::initialize + initialize_invertible (state/StateHelper.cpp:393-577): add a new k-wide variable (k = H_L.c)
// at the END of the covariance from  res = H_R dx + H_L dx_new + n,  n ~ N(0, sigma2 I).
// Pout must hold (N+k)^2 doubles (row-major, leading dimension N+k); it is written only when *accepted = 1.
inline int cov_initialize(const double *P, int N, const std::vector<Var> &H_order, Mat H_R, Mat H_L, std::vector<double> res, double sigma2,
                          double chi2_mult, const double *chi2_table, double *Pout, int *accepted, double *dx_new, double *dx) {
  const int r = H_L.r, k = H_L.c, n = H_R.c;
  *accepted = 0;
  // Givens split (:429-440): top k rows depend on the new variable, the rest does not
  for (int c = 0; c < k; ++c) {
    for (int m = r - 1; m > c; m--) {
      Givens g = make_givens(H_L(m - 1, c), H_L(m, c));
      for (int j = c; j < k; j++)
        apply_givens(g, H_L(m - 1, j), H_L(m, j));
      apply_givens(g, res[m - 1], res[m]);
      for (int j = 0; j < n; j++)
        apply_givens(g, H_R(m - 1, j), H_R(m, j));
    }
  }
  Mat Hxinit(k, n), H_finit(k, k), Hup(r - k, n);
  std::vector<double> resinit(k), resup(r - k);
  for (int i = 0; i < k; i++) {
    for (int j = 0; j < n; j++)
      Hxinit(i, j) = H_R(i, j);
    for (int j = 0; j < k; j++)
      H_finit(i, j) = H_L(i, j);
    resinit[i] = res[i];
  }
  for (int i = k; i < r; i++) {
    for (int j = 0; j < n; j++)
      Hup(i - k, j) = H_R(i, j);
    resup[i - k] = res[i];
  }
  // Mahalanobis gate on the update portion (:458-470); threshold from chi2(res.rows())
  if (r - k > 0) {
    FeatJac J;
    J.rows = r - k;
    J.order = H_order;
    J.Hx = Hup;
    J.res = resup;
    bool spd = true;
    double chi2 = feature_chi2(P, N, J, sigma2, &spd);
    double chi2_check = chi2_table[std::min(r, OVB_CHI2_TABLE_LEN - 1)];
    if (!(chi2 <= chi2_mult * chi2_check))
      return OVB_OK;
  }
  // ---- initialize_invertible (:484-577)
  std::vector<int> cols;
  for (const Var &v : H_order)
    for (int q = 0; q < v.size; q++)
      cols.push_back(v.off + q);
  Mat M_a(N, k); // P[:, cols] * Hxinit'
  for (int a = 0; a < N; a++)
    for (int i = 0; i < k; i++) {
      double acc = 0.0;
      for (int j = 0; j < n; j++)
        acc += P[(size_t)a * N + cols[j]] * Hxinit(i, j);
      M_a(a, i) = acc;
    }
  Mat M(k, k); // Hxinit P_small Hxinit' + R
  for (int i = 0; i < k; i++)
    for (int i2 = 0; i2 < k; i2++) {
      double acc = 0.0;
      for (int j = 0; j < n; j++)
        acc += Hxinit(i, j) * M_a(cols[j], i2);
      M(i, i2) = acc + (i == i2 ? sigma2 : 0.0);
    }
  for (int i = 0; i < k; i++) // selfadjointView<Upper>
    for (int i2 = 0; i2 < i; i2++)
      M(i, i2) = M(i2, i);
  // H_L^-1 by Gauss-Jordan with partial pivoting (Eigen: PartialPivLU for dynamic sizes)
  Mat A = H_finit, Inv(k, k);
  for (int i = 0; i < k; i++)
    Inv(i, i) = 1.0;
  for (int c = 0; c < k; c++) {
    int piv = c;
    for (int i = c + 1; i < k; i++)
      if (std::fabs(A(i, c)) > std::fabs(A(piv, c)))
        piv = i;
    if (piv != c)
      for (int j = 0; j < k; j++) {
        std::swap(A(c, j), A(piv, j));
        std::swap(Inv(c, j), Inv(piv, j));
      }
    double d = A(c, c);
    for (int j = 0; j < k; j++) {
      A(c, j) /= d;
      Inv(c, j) /= d;
    }
    for (int i = 0; i < k; i++) {
      if (i == c)
        continue;
      double f = A(i, c);
      for (int j = 0; j < k; j++) {
        A(i, j) -= f * A(c, j);
        Inv(i, j) -= f * Inv(c, j);
      }
    }
  }
  const int N2 = N + k;
  for (int a = 0; a < N; a++)
    for (int b = 0; b < N; b++)
      Pout[(size_t)a * N2 + b] = P[(size_t)a * N + b];
  for (int a = 0; a < N; a++)
    for (int q = 0; q < k; q++) {
      double acc = 0.0;
      for (int i = 0; i < k; i++)
        acc += M_a(a, i) * Inv(q, i); // -M_a * H_Linv'
      Pout[(size_t)a * N2 + N + q] = -acc;
      Pout[(size_t)(N + q) * N2 + a] = -acc;
    }
  for (int q = 0; q < k; q++)
    for (int q2 = 0; q2 < k; q2++) {
      double acc = 0.0;
      for (int i = 0; i < k; i++)
        for (int i2 = 0; i2 < k; i2++)
          acc += Inv(q, i) * M(i, i2) * Inv(q2, i2);
      Pout[(size_t)(N + q) * N2 + N + q2] = acc;
    }
  for (int q = 0; q < k; q++) {
    double acc = 0.0;
    for (int i = 0; i < k; i++)
      acc += Inv(q, i) * resinit[i];
    dx_new[q] = acc; // new_variable->update(H_Linv * res)
  }
  *accepted = 1;
  for (int i = 0; i < N2; i++)
    dx[i] = 0.0;
  // update with the nullspace-projected portion (:476-479)
  if (r - k > 0) {
    std::vector<double> Rdiag(r - k, sigma2);
    int neg = -1;
    return ekf_update(Pout, N2, H_order, Hup, resup, Rdiag, dx, &neg);
  }
  return OVB_OK;
}

// UpdaterSLAM::perform_anchor_change (update/UpdaterSLAM.cpp:506-647), the host math in front of its EKFPropagation:
// the landmark re-expressed in the new anchor (value and FEJ value), the variable order phi_order_OLD and Phi.
struct AnchorChange {
  V3 value, value_fej;
  std::vector<Var> order; // phi_order_OLD: x_order_old (+ new ones from x_order_new), then the landmark
  Mat Phi;                // phisize x sum(order sizes)
};
inline AnchorChange anchor_change(const ovb_frame &fr, const ovb_opts &op, int rep, int lm_off, const V3 &value, const V3 &value_fej,
                                  int old_cam, int old_clone, int new_cam, int new_clone) {
  AnchorChange out;
  Mat Hf_old, Hf_new;
  std::vector<Mat> Hx_old, Hx_new;
  std::vector<Var> xo_old, xo_new;
  jacobian_representation(fr, op, rep, value, value_fej, value, old_cam, old_clone, Hf_old, Hx_old, xo_old);
  auto cam_pose = [&](int cam, int cl, bool fej, M3 &R_GtoC, V3 &p_CinG) {
    M3 R_GtoI = load_m3((fej ? fr.clone_R_fej : fr.clone_R) + 9 * cl);
    V3 p_IinG = load_v3((fej ? fr.clone_p_fej : fr.clone_p) + 3 * cl);
    R_GtoC = m3mul(load_m3(fr.cam_R + 9 * cam), R_GtoI);
    p_CinG = vsub(p_IinG, m3Tv(R_GtoC, load_v3(fr.cam_p + 3 * cam)));
  };
  auto transfer = [&](bool fej, const V3 &p_old) {
    M3 R_GtoOLD, R_GtoNEW;
    V3 p_OLDinG, p_NEWinG;
    cam_pose(old_cam, old_clone, fej, R_GtoOLD, p_OLDinG);
    cam_pose(new_cam, new_clone, fej, R_GtoNEW, p_NEWinG);
    M3 R_OLDtoNEW = m3mul(R_GtoNEW, m3T(R_GtoOLD));
    V3 p_OLDinNEW = m3v(R_GtoNEW, vsub(p_OLDinG, p_NEWinG));
    return vadd(m3v(R_OLDtoNEW, p_old), p_OLDinNEW);
  };
  out.value = transfer(false, value);
  out.value_fej = transfer(true, value_fej);
  jacobian_representation(fr, op, rep, out.value, out.value_fej, out.value, new_cam, new_clone, Hf_new, Hx_new, xo_new);
  // phi_order_OLD (:600-617)
  std::vector<int> col_old(xo_old.size()), col_new(xo_new.size());
  int cur = 0;
  auto place = [&](const Var &v) {
    int c = find_var(out.order, v.off);
    if (c < 0) {
      c = cur;
      out.order.push_back(v);
      cur += v.size;
    }
    return c;
  };
  for (size_t i = 0; i < xo_old.size(); i++)
    col_old[i] = place(xo_old[i]);
  for (size_t i = 0; i < xo_new.size(); i++)
    col_new[i] = place(xo_new[i]);
  const int phisize = (rep != OVB_REP_ANCHORED_INVERSE_DEPTH_SINGLE) ? 3 : 1;
  const int col_lm = cur;
  out.order.push_back(Var{lm_off, phisize});
  cur += phisize;
  out.Phi.resize_zero(phisize, cur);
  // H_f_new^-1 (:624-629)
  Mat Inv(phisize, 3);
  if (phisize == 1) {
    double n2 = 0;
    for (int i = 0; i < 3; i++)
      n2 += Hf_new(i, 0) * Hf_new(i, 0);
    for (int i = 0; i < 3; i++)
      Inv(0, i) = 1.0 / n2 * Hf_new(i, 0);
  } else {
    M3 A;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        A(i, j) = Hf_new(i, j);
    for (int c = 0; c < 3; c++) { // colPivHouseholderQr().solve(Identity), column by column
      V3 e = v3(c == 0, c == 1, c == 2);
      V3 x = colpiv_qr_solve3(A, e);
      for (int i = 0; i < 3; i++)
        Inv(i, c) = x(i);
    }
  }
  auto add_block = [&](int col, const Mat &B, double sign) { // Phi[:, col:col+B.c] += sign * Inv * B
    for (int i = 0; i < phisize; i++)
      for (int j = 0; j < B.c; j++) {
        double acc = 0.0;
        for (int k = 0; k < 3; k++)
          acc += Inv(i, k) * B(k, j);
        out.Phi(i, col + j) += sign * acc;
      }
  };
  for (size_t i = 0; i < Hx_old.size(); i++)
    add_block(col_old[i], Hx_old[i], 1.0);
  {
    Mat blk(phisize, phisize);
    for (int i = 0; i < phisize; i++)
      for (int j = 0; j < phisize; j++) {
        double acc = 0.0;
        for (int k = 0; k < 3; k++)
          acc += Inv(i, k) * Hf_old(k, j);
        out.Phi(i, col_lm + j) = acc;
      }
  }
  for (size_t i = 0; i < Hx_new.size(); i++)
    add_block(col_new[i], Hx_new[i], -1.0);
  return out;
}

} // namespace ovo
