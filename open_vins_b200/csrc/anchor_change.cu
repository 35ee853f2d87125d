// anchor_change.cu — host math of UpdaterSLAM::perform_anchor_change (ov_msckf/src/update/UpdaterSLAM.cpp:506-647).
// Re-anchoring a SLAM landmark is a handful of 3x3 products per landmark per marginalised clone: it stays on the host,
// stateless, and hands Phi + the variable order to ovb_cov_propagate (StateHelper::EKFPropagation on the resident P).
// The representation Jacobians follow UpdaterHelper::get_feature_jacobian_representation (UpdaterHelper.cpp:32-190) for
// the anchored representations, including its FEJ rule (the anchor pose is taken at its first estimate, the landmark is
// re-expressed in that FEJ anchor frame from the best global position, :89-96).
#include "../../include/ovb200.h"
#include <cmath>
#include <cstring>

namespace {

struct M3 {
  double a[9];
};
inline M3 mul(const M3 &x, const M3 &y) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      r.a[3 * i + j] = x.a[3 * i] * y.a[j] + x.a[3 * i + 1] * y.a[3 + j] + x.a[3 * i + 2] * y.a[6 + j];
  return r;
}
inline M3 tr(const M3 &x) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      r.a[3 * i + j] = x.a[3 * j + i];
  return r;
}
inline void mv(const M3 &x, const double v[3], double out[3]) {
  for (int i = 0; i < 3; i++)
    out[i] = x.a[3 * i] * v[0] + x.a[3 * i + 1] * v[1] + x.a[3 * i + 2] * v[2];
}
inline M3 skew(const double w[3]) {
  M3 r = {{0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0}};
  return r;
}
inline M3 load(const double *p) {
  M3 r;
  std::memcpy(r.a, p, sizeof(r.a));
  return r;
}

// d p_FinG / d(landmark parameters) [3 x nf], d p_FinG / d(anchor clone) [3 x 6], d p_FinG / d(anchor extrinsics) [3 x 6]
void rep_jacobian(const ovb_frame *fr, const ovb_opts *op, int rep, const double p_FinA_in[3], int acam, int aclone, double Hf[9], int *nf,
                  double Hanc[18], double Hcal[18]) {
  const M3 R_ItoC = load(fr->cam_R + 9 * acam);
  const double *p_IinC = fr->cam_p + 3 * acam;
  M3 R_GtoI = load(fr->clone_R + 9 * aclone);
  double p_IinG[3] = {fr->clone_p[3 * aclone], fr->clone_p[3 * aclone + 1], fr->clone_p[3 * aclone + 2]};
  double p_FinA[3] = {p_FinA_in[0], p_FinA_in[1], p_FinA_in[2]};
  if (op->do_fej) {
    // best global position with the current estimates, then back into the FEJ anchor frame (:89-96)
    const M3 RtRt = mul(tr(R_GtoI), tr(R_ItoC));
    double d[3] = {p_FinA[0] - p_IinC[0], p_FinA[1] - p_IinC[1], p_FinA[2] - p_IinC[2]}, best[3];
    mv(RtRt, d, best);
    for (int i = 0; i < 3; i++)
      best[i] += p_IinG[i];
    R_GtoI = load((fr->clone_R_fej ? fr->clone_R_fej : fr->clone_R) + 9 * aclone);
    const double *pf = (fr->clone_p_fej ? fr->clone_p_fej : fr->clone_p) + 3 * aclone;
    for (int i = 0; i < 3; i++)
      p_IinG[i] = pf[i];
    const M3 RR = tr(mul(tr(R_GtoI), tr(R_ItoC)));
    double e[3] = {best[0] - p_IinG[0], best[1] - p_IinG[1], best[2] - p_IinG[2]};
    mv(RR, e, p_FinA);
    for (int i = 0; i < 3; i++)
      p_FinA[i] += p_IinC[i];
  }
  const M3 R_CtoG = mul(tr(R_GtoI), tr(R_ItoC));
  // H_anc = [-R_GtoI' skew(R_ItoC' (p_FinA - p_IinC)), I] (:100-102)
  {
    double d[3] = {p_FinA[0] - p_IinC[0], p_FinA[1] - p_IinC[1], p_FinA[2] - p_IinC[2]}, v[3];
    mv(tr(R_ItoC), d, v);
    M3 nR = tr(R_GtoI);
    for (double &x : nR.a)
      x = -x;
    const M3 blk = mul(nR, skew(v));
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        Hanc[6 * r + c] = blk.a[3 * r + c];
        Hanc[6 * r + 3 + c] = (r == c) ? 1.0 : 0.0;
      }
  }
  // H_calib = [-R_CtoG skew(p_FinA - p_IinC), -R_CtoG] (:109-115)
  {
    double d[3] = {p_FinA[0] - p_IinC[0], p_FinA[1] - p_IinC[1], p_FinA[2] - p_IinC[2]};
    M3 nR = R_CtoG;
    for (double &x : nR.a)
      x = -x;
    const M3 blk = mul(nR, skew(d));
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        Hcal[6 * r + c] = blk.a[3 * r + c];
        Hcal[6 * r + 3 + c] = -R_CtoG.a[3 * r + c];
      }
  }
  *nf = 3;
  if (rep == OVB_REP_ANCHORED_3D) { // :118-121
    std::memcpy(Hf, R_CtoG.a, sizeof(double) * 9);
    return;
  }
  M3 d = {{0, 0, 0, 0, 0, 0, 0, 0, 0}};
  if (rep == OVB_REP_ANCHORED_FULL_INVERSE_DEPTH) { // :124-150
    const double rho = 1 / std::sqrt(p_FinA[0] * p_FinA[0] + p_FinA[1] * p_FinA[1] + p_FinA[2] * p_FinA[2]);
    const double phi = std::acos(rho * p_FinA[2]);
    const double theta = std::atan2(p_FinA[1], p_FinA[0]);
    const double sin_th = std::sin(theta), cos_th = std::cos(theta), sin_phi = std::sin(phi), cos_phi = std::cos(phi);
    d.a[0] = -(1.0 / rho) * sin_th * sin_phi;
    d.a[1] = (1.0 / rho) * cos_th * cos_phi;
    d.a[2] = -(1.0 / (rho * rho)) * cos_th * sin_phi;
    d.a[3] = (1.0 / rho) * cos_th * sin_phi;
    d.a[4] = (1.0 / rho) * sin_th * cos_phi;
    d.a[5] = -(1.0 / (rho * rho)) * sin_th * sin_phi;
    d.a[6] = 0.0;
    d.a[7] = -(1.0 / rho) * sin_phi;
    d.a[8] = -(1.0 / (rho * rho)) * cos_phi;
  } else if (rep == OVB_REP_ANCHORED_MSCKF_INVERSE_DEPTH) { // :153-172
    const double alpha = p_FinA[0] / p_FinA[2], beta = p_FinA[1] / p_FinA[2], rho = 1 / p_FinA[2];
    d.a[0] = (1.0 / rho);
    d.a[2] = -(1.0 / (rho * rho)) * alpha;
    d.a[4] = (1.0 / rho);
    d.a[5] = -(1.0 / (rho * rho)) * beta;
    d.a[8] = -(1.0 / (rho * rho));
  } else { // ANCHORED_INVERSE_DEPTH_SINGLE (:175-186): only the depth is a parameter
    const double rho = 1.0 / p_FinA[2];
    const double v[3] = {-(1.0 / (rho * rho)) * (rho * p_FinA[0]), -(1.0 / (rho * rho)) * (rho * p_FinA[1]), -(1.0 / (rho * rho)) * (rho * p_FinA[2])};
    double out[3];
    mv(R_CtoG, v, out);
    Hf[0] = out[0];
    Hf[1] = out[1];
    Hf[2] = out[2];
    *nf = 1;
    return;
  }
  const M3 L = mul(R_CtoG, d);
  std::memcpy(Hf, L.a, sizeof(double) * 9);
}

// inverse of a 3x3 by Gauss-Jordan with partial pivoting (the reference solves H_f_new X = I with ColPivHouseholderQR)
bool inv3(const double A_in[9], double Inv[9]) {
  double A[9];
  std::memcpy(A, A_in, sizeof(A));
  for (int i = 0; i < 9; i++)
    Inv[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int c = 0; c < 3; c++) {
    int piv = c;
    for (int i = c + 1; i < 3; i++)
      if (std::fabs(A[3 * i + c]) > std::fabs(A[3 * piv + c]))
        piv = i;
    if (!(std::fabs(A[3 * piv + c]) > 0.0))
      return false;
    if (piv != c)
      for (int j = 0; j < 3; j++) {
        double t = A[3 * c + j];
        A[3 * c + j] = A[3 * piv + j];
        A[3 * piv + j] = t;
        t = Inv[3 * c + j];
        Inv[3 * c + j] = Inv[3 * piv + j];
        Inv[3 * piv + j] = t;
      }
    const double d = A[3 * c + c];
    for (int j = 0; j < 3; j++) {
      A[3 * c + j] /= d;
      Inv[3 * c + j] /= d;
    }
    for (int i = 0; i < 3; i++) {
      if (i == c)
        continue;
      const double f = A[3 * i + c];
      for (int j = 0; j < 3; j++) {
        A[3 * i + j] -= f * A[3 * c + j];
        Inv[3 * i + j] -= f * Inv[3 * c + j];
      }
    }
  }
  return true;
}

void camera_pose(const ovb_frame *fr, int cam, int cl, bool fej, M3 *R_GtoC, double p_CinG[3]) {
  const M3 R_GtoI = load(((fej && fr->clone_R_fej) ? fr->clone_R_fej : fr->clone_R) + 9 * cl);
  const double *p_IinG = ((fej && fr->clone_p_fej) ? fr->clone_p_fej : fr->clone_p) + 3 * cl;
  *R_GtoC = mul(load(fr->cam_R + 9 * cam), R_GtoI);
  double t[3];
  mv(tr(*R_GtoC), fr->cam_p + 3 * cam, t);
  for (int i = 0; i < 3; i++)
    p_CinG[i] = p_IinG[i] - t[i];
}

void transfer(const ovb_frame *fr, int old_cam, int old_clone, int new_cam, int new_clone, bool fej, const double p_old[3], double p_new[3]) {
  M3 R_GtoOLD, R_GtoNEW;
  double p_OLDinG[3], p_NEWinG[3];
  camera_pose(fr, old_cam, old_clone, fej, &R_GtoOLD, p_OLDinG);
  camera_pose(fr, new_cam, new_clone, fej, &R_GtoNEW, p_NEWinG);
  const M3 R_OLDtoNEW = mul(R_GtoNEW, tr(R_GtoOLD));
  const double d[3] = {p_OLDinG[0] - p_NEWinG[0], p_OLDinG[1] - p_NEWinG[1], p_OLDinG[2] - p_NEWinG[2]};
  double p_OLDinNEW[3], r[3];
  mv(R_GtoNEW, d, p_OLDinNEW);
  mv(R_OLDtoNEW, p_old, r);
  for (int i = 0; i < 3; i++)
    p_new[i] = r[i] + p_OLDinNEW[i];
}

} // namespace

extern "C" ovb_status ovb_slam_anchor_change(const ovb_frame *fr, const ovb_opts *op, int lm_off, const double *value, const double *value_fej,
                                             int old_cam, int old_clone, int new_cam, int new_clone, double *new_value, double *new_value_fej,
                                             double *Phi, int32_t *order_off, int32_t *order_sz, int32_t *n_order, int32_t *n_cols) {
  if (!fr || !op || !value || !value_fej || !new_value || !new_value_fej || !Phi || !order_off || !order_sz || !n_order || !n_cols)
    return OVB_ERR_ARG;
  const int rep = op->feat_rep;
  if (rep < OVB_REP_ANCHORED_3D || rep > OVB_REP_ANCHORED_INVERSE_DEPTH_SINGLE) // global representations have no anchor (:493-496)
    return OVB_ERR_ARG;
  if (old_cam < 0 || old_cam >= fr->n_cams || new_cam < 0 || new_cam >= fr->n_cams || old_clone < 0 || old_clone >= fr->n_clones || new_clone < 0 ||
      new_clone >= fr->n_clones)
    return OVB_ERR_ARG;
  const bool ext = op->do_calib_camera_pose != 0;
  if (ext && (!fr->cam_ext_off || fr->cam_ext_off[old_cam] < 0 || fr->cam_ext_off[new_cam] < 0))
    return OVB_ERR_ARG;
  double Hf_old[9], Hf_new[9], Hanc_old[18], Hcal_old[18], Hanc_new[18], Hcal_new[18];
  int nf_old = 3, nf_new = 3;
  rep_jacobian(fr, op, rep, value, old_cam, old_clone, Hf_old, &nf_old, Hanc_old, Hcal_old);
  transfer(fr, old_cam, old_clone, new_cam, new_clone, false, value, new_value);
  transfer(fr, old_cam, old_clone, new_cam, new_clone, true, value_fej, new_value_fej);
  rep_jacobian(fr, op, rep, new_value, new_cam, new_clone, Hf_new, &nf_new, Hanc_new, Hcal_new);
  // phi_order_OLD = unique(x_order_old ++ x_order_new) ++ landmark (:600-617)
  int n = 0, cur = 0;
  auto place = [&](int off) {
    int c = 0;
    for (int i = 0; i < n; i++) {
      if (order_off[i] == off)
        return c;
      c += order_sz[i];
    }
    order_off[n] = off;
    order_sz[n] = 6;
    n++;
    cur += 6;
    return cur - 6;
  };
  const int c_old_anc = place(fr->clone_off[old_clone]);
  const int c_old_cal = ext ? place(fr->cam_ext_off[old_cam]) : -1;
  const int c_new_anc = place(fr->clone_off[new_clone]);
  const int c_new_cal = ext ? place(fr->cam_ext_off[new_cam]) : -1;
  const int phisize = nf_new; // 3, or 1 for the single-depth representation
  const int c_lm = cur;
  order_off[n] = lm_off;
  order_sz[n] = phisize;
  n++;
  cur += phisize;
  *n_order = n;
  *n_cols = cur;
  // H_f_new^-1: phisize x 3 (:624-629)
  double Inv[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (phisize == 1) {
    const double n2 = Hf_new[0] * Hf_new[0] + Hf_new[1] * Hf_new[1] + Hf_new[2] * Hf_new[2];
    for (int i = 0; i < 3; i++)
      Inv[i] = 1.0 / n2 * Hf_new[i];
  } else if (!inv3(Hf_new, Inv)) {
    return OVB_ERR_ARG;
  }
  for (int i = 0; i < phisize * cur; i++)
    Phi[i] = 0.0;
  auto add6 = [&](int col, const double *B, double sign) { // Phi[:, col:col+6] += sign * Inv * B (B is 3 x 6)
    for (int i = 0; i < phisize; i++)
      for (int j = 0; j < 6; j++) {
        double acc = 0.0;
        for (int k = 0; k < 3; k++)
          acc += Inv[3 * i + k] * B[6 * k + j];
        Phi[(size_t)i * cur + col + j] += sign * acc;
      }
  };
  add6(c_old_anc, Hanc_old, 1.0);
  if (ext)
    add6(c_old_cal, Hcal_old, 1.0);
  for (int i = 0; i < phisize; i++)
    for (int j = 0; j < phisize; j++) {
      double acc = 0.0;
      for (int k = 0; k < 3; k++)
        acc += Inv[3 * i + k] * Hf_old[nf_old * k + j];
      Phi[(size_t)i * cur + c_lm + j] = acc;
    }
  add6(c_new_anc, Hanc_new, -1.0);
  if (ext)
    add6(c_new_cal, Hcal_new, -1.0);
  return OVB_OK;
}
