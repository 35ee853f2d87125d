// oracle/ovo_capi.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (see ovo_math.hpp).
//
// extern "C" entry points of the CPU oracle so that tests/ and bench.py's cpu_baseline / --impl reference legs can
// call it through ctypes with the same structs the product ABI takes. Built by oracle/Makefile into
// oracle/libovoracle.so. Nothing under open_vins_b200/ may link or load this library.
#include "ovo_core.hpp"

using namespace ovo;

extern "C" {

int ovo_triangulate(const ovb_frame *fr, const ovb_feat_batch *fb, const ovb_opts *op, ovb_feat_out *out, int32_t *gn_runs,
                    int32_t *gn_solves, double *gn_lambda) {
  for (int f = 0; f < fb->n_feats; f++) {
    V3 pA = v3(std::nan(""), std::nan(""), std::nan("")), pG = pA;
    int acam = -1, aclone = -1, st;
    GnTrace tr;
    if (fb->meas_off[f + 1] - fb->meas_off[f] < 2) {
      st = OVB_FEAT_FEW_MEAS;
    } else {
      st = op->triangulate_1d ? single_triangulation_1d(*fr, *fb, *op, f, pA, pG, acam, aclone)
                              : single_triangulation(*fr, *fb, *op, f, pA, pG, acam, aclone);
      if (st == OVB_FEAT_OK && op->refine_features)
        st = single_gaussnewton(*fr, *fb, *op, f, acam, aclone, pA, pG, &tr);
    }
    out->status[f] = st;
    for (int k = 0; k < 3; k++) {
      out->p_FinA[3 * f + k] = pA(k);
      out->p_FinG[3 * f + k] = pG(k);
    }
    out->anchor_cam[f] = acam;
    out->anchor_clone[f] = aclone;
    if (out->chi2)
      out->chi2[f] = std::nan("");
    if (gn_runs)
      gn_runs[f] = tr.runs;
    if (gn_solves)
      gn_solves[f] = tr.solves;
    if (gn_lambda)
      gn_lambda[f] = tr.lam;
  }
  return OVB_OK;
}

// Per-feature Jacobian dump in the canonical column layout (column j <-> covariance column col_index[j]); see
// ovb_feature_jacobians in include/ovb200.h. P may be NULL when stage 0 (then chi2 is not evaluated).
int ovo_feature_jacobians(const ovb_frame *fr, const ovb_feat_batch *fb, const ovb_opts *op, const double *chi2_table, const double *P,
                          int N, ovb_feat_out *out, int stage, double *Hf_out, double *Hx_out, double *res_out, int32_t *row_off_out,
                          int ncols, const int32_t *col_index, int ld_out) {
  int rep = op->feat_rep;
  if (rep == OVB_REP_ANCHORED_INVERSE_DEPTH_SINGLE)
    rep = OVB_REP_ANCHORED_MSCKF_INVERSE_DEPTH;
  std::vector<int> col_of(N > 0 ? N : 4096, -1);
  for (int j = 0; j < ncols; j++)
    if (col_index[j] >= 0 && col_index[j] < (int)col_of.size())
      col_of[col_index[j]] = j;
  double sigma_pix_sq = std::pow(op->sigma_pix, 2);
  int row = 0;
  for (int f = 0; f < fb->n_feats; f++) {
    row_off_out[f] = row;
    int M = fb->meas_off[f + 1] - fb->meas_off[f];
    int nrows = (stage == 0) ? 2 * M : std::max(0, 2 * M - 3);
    if (out->status[f] != OVB_FEAT_OK) {
      row += nrows;
      continue;
    }
    V3 pG = v3(out->p_FinG[3 * f], out->p_FinG[3 * f + 1], out->p_FinG[3 * f + 2]);
    V3 pA = v3(out->p_FinA[3 * f], out->p_FinA[3 * f + 1], out->p_FinA[3 * f + 2]);
    FeatJac J;
    feature_jacobian_full(*fr, *fb, *op, f, rep, pG, pG, pA, out->anchor_cam[f], out->anchor_clone[f], J);
    if (stage == 1) {
      nullspace_project_inplace(J);
      if (P) {
        bool spd;
        double chi2 = feature_chi2(P, N, J, sigma_pix_sq, &spd);
        if (out->chi2)
          out->chi2[f] = chi2;
        double chi2_check = chi2_table[std::min(J.rows, OVB_CHI2_TABLE_LEN - 1)];
        if (!(chi2 <= op->chi2_multipler * chi2_check)) {
          out->status[f] = OVB_FEAT_CHI2;
          row += nrows;
          continue;
        }
      }
    }
    for (int i = 0; i < J.rows; i++) {
      if (res_out)
        res_out[row + i] = J.res[i];
      if (stage == 0 && Hf_out)
        for (int k = 0; k < 3; k++)
          Hf_out[(size_t)(row + i) * 3 + k] = (k < J.nf) ? J.Hf(i, k) : 0.0;
      int lc = 0;
      for (const Var &v : J.order) {
        for (int k = 0; k < v.size; k++) {
          int gc = col_of[v.off + k];
          if (gc >= 0 && Hx_out)
            Hx_out[(size_t)(row + i) * ld_out + gc] = J.Hx(i, lc + k);
        }
        lc += v.size;
      }
    }
    row += nrows;
  }
  row_off_out[fb->n_feats] = row;
  return OVB_OK;
}

// Whole update. Optional dumps (any may be NULL):
//  order_off/order_sz[<=OVB_MAX_VARS], n_order : Hx_order_big
//  H_big (ld = *cols) / res_big  : stacked system BEFORE compression, row-major, capacity cap_rows rows
//  H_cmp / res_cmp               : after compression, row-major (rows_update x cols)
//  times[4]                      : seconds {triangulate, create system, compress, update}
int ovo_msckf_update(const ovb_frame *fr, const ovb_feat_batch *fb, const ovb_opts *op, const double *chi2_table, double *P, int N,
                     ovb_feat_out *out, double *dx, ovb_stats *stats, int32_t *order_off, int32_t *order_sz, int32_t *n_order,
                     double *H_big, double *res_big, int cap_rows, double *H_cmp, double *res_cmp, double *times) {
  UpdateDump dump;
  ovb_stats st_local;
  if (!stats)
    stats = &st_local;
  int st = msckf_update(*fr, *fb, *op, chi2_table, P, N, out, dx, stats, &dump);
  if (n_order)
    *n_order = (int)dump.order_big.size();
  for (size_t i = 0; i < dump.order_big.size() && i < OVB_MAX_VARS; i++) {
    if (order_off)
      order_off[i] = dump.order_big[i].off;
    if (order_sz)
      order_sz[i] = dump.order_big[i].size;
  }
  int cols = dump.H_big.c;
  if (H_big)
    for (int i = 0; i < dump.H_big.r && i < cap_rows; i++)
      for (int k = 0; k < cols; k++)
        H_big[(size_t)i * cols + k] = dump.H_big(i, k);
  if (res_big)
    for (int i = 0; i < (int)dump.res_big.size() && i < cap_rows; i++)
      res_big[i] = dump.res_big[i];
  if (H_cmp)
    for (int i = 0; i < dump.H_cmp.r; i++)
      for (int k = 0; k < cols; k++)
        H_cmp[(size_t)i * cols + k] = dump.H_cmp(i, k);
  if (res_cmp)
    for (int i = 0; i < (int)dump.res_cmp.size(); i++)
      res_cmp[i] = dump.res_cmp[i];
  if (times) {
    times[0] = dump.t_tri;
    times[1] = dump.t_sys;
    times[2] = dump.t_cmp;
    times[3] = dump.t_upd;
  }
  return st;
}

// UpdaterSLAM::update steps 4-5. H_big/res_big/Rdiag_big: the stacked system handed to EKFUpdate (row-major, ld = *cols).
int ovo_slam_update(const ovb_frame *fr, const ovb_feat_batch *fb, const ovb_landmarks *lm, const ovb_opts *op, const double *chi2_table,
                    double *P, int N, ovb_feat_out *out, double *dx, ovb_stats *stats, int32_t *order_off, int32_t *order_sz, int32_t *n_order,
                    double *H_big, double *res_big, double *Rdiag_big, int cap_rows) {
  UpdateDump dump;
  ovb_stats st_local;
  if (!stats)
    stats = &st_local;
  int st = slam_update(*fr, *fb, *lm, *op, chi2_table, P, N, out, dx, stats, &dump);
  if (n_order)
    *n_order = (int)dump.order_big.size();
  for (size_t i = 0; i < dump.order_big.size() && i < OVB_MAX_VARS; i++) {
    if (order_off)
      order_off[i] = dump.order_big[i].off;
    if (order_sz)
      order_sz[i] = dump.order_big[i].size;
  }
  int cols = dump.H_big.c;
  if (H_big)
    for (int i = 0; i < dump.H_big.r && i < cap_rows; i++)
      for (int k = 0; k < cols; k++)
        H_big[(size_t)i * cols + k] = dump.H_big(i, k);
  for (int i = 0; i < (int)dump.res_big.size() && i < cap_rows; i++) {
    if (res_big)
      res_big[i] = dump.res_big[i];
    if (Rdiag_big)
      Rdiag_big[i] = dump.res_cmp[i];
  }
  return st;
}

// StateHelper::initialize: H_R r x n, H_L r x k (row-major). Pout (N+k)^2.
int ovo_cov_initialize(const double *P, int N, const int *off, const int *sz, int nvar, const double *H_R, const double *H_L, const double *res,
                       int r, int k, double sigma2, double chi2_mult, const double *chi2_table, double *Pout, int *accepted, double *dx_new,
                       double *dx) {
  std::vector<Var> order;
  int n = 0;
  for (int i = 0; i < nvar; i++) {
    order.push_back(Var{off[i], sz[i]});
    n += sz[i];
  }
  Mat HR(r, n), HL(r, k);
  for (int i = 0; i < r; i++) {
    for (int j = 0; j < n; j++)
      HR(i, j) = H_R[(size_t)i * n + j];
    for (int j = 0; j < k; j++)
      HL(i, j) = H_L[(size_t)i * k + j];
  }
  std::vector<double> rv(res, res + r);
  return cov_initialize(P, N, order, HR, HL, rv, sigma2, chi2_mult, chi2_table, Pout, accepted, dx_new, dx);
}

// UpdaterSLAM::perform_anchor_change host math. Phi_out: phisize x 27 capacity (row-major, leading dimension *ncols).
int ovo_anchor_change(const ovb_frame *fr, const ovb_opts *op, int lm_off, const double *value, const double *value_fej, int old_cam, int old_clone,
                      int new_cam, int new_clone, double *new_value, double *new_value_fej, double *Phi_out, int32_t *order_off, int32_t *order_sz,
                      int32_t *n_order, int32_t *ncols) {
  AnchorChange a = anchor_change(*fr, *op, op->feat_rep, lm_off, v3(value[0], value[1], value[2]), v3(value_fej[0], value_fej[1], value_fej[2]),
                                 old_cam, old_clone, new_cam, new_clone);
  for (int k = 0; k < 3; k++) {
    new_value[k] = a.value(k);
    new_value_fej[k] = a.value_fej(k);
  }
  *n_order = (int)a.order.size();
  for (size_t i = 0; i < a.order.size(); i++) {
    order_off[i] = a.order[i].off;
    order_sz[i] = a.order[i].size;
  }
  *ncols = a.Phi.c;
  for (int i = 0; i < a.Phi.r; i++)
    for (int j = 0; j < a.Phi.c; j++)
      Phi_out[(size_t)i * a.Phi.c + j] = a.Phi(i, j);
  return OVB_OK;
}

// measurement_compress_inplace on a row-major H (m x n); outputs R (min(m,n) x n row-major) and z.
int ovo_compress(const double *H, int m, int n, const double *res, double *R_out, double *z_out) {
  Mat Hc(m, n);
  for (int i = 0; i < m; i++)
    for (int k = 0; k < n; k++)
      Hc(i, k) = H[(size_t)i * n + k];
  std::vector<double> r(res, res + m);
  measurement_compress_inplace(Hc, r);
  for (int i = 0; i < Hc.r; i++)
    for (int k = 0; k < n; k++)
      R_out[(size_t)i * n + k] = Hc(i, k);
  for (int i = 0; i < Hc.r; i++)
    z_out[i] = r[i];
  return OVB_OK;
}

int ovo_ekf_update(double *P, int N, const int *off, const int *sz, int nvar, const double *H, int r, const double *res, double sigma2,
                   const double *Rdiag, double *dx, int *neg_index) {
  std::vector<Var> order;
  int n = 0;
  for (int i = 0; i < nvar; i++) {
    order.push_back(Var{off[i], sz[i]});
    n += sz[i];
  }
  Mat Hc(r, n);
  for (int i = 0; i < r; i++)
    for (int k = 0; k < n; k++)
      Hc(i, k) = H[(size_t)i * n + k];
  std::vector<double> rv(res, res + r), Rd(r, sigma2);
  if (Rdiag)
    for (int i = 0; i < r; i++)
      Rd[i] = Rdiag[i];
  return ekf_update(P, N, order, Hc, rv, Rd, dx, neg_index);
}

int ovo_cov_propagate(double *P, int N, int new_off, int p, const int *old_off, const int *old_sz, int nold, const double *Phi,
                      const double *Q) {
  std::vector<Var> order;
  for (int i = 0; i < nold; i++)
    order.push_back(Var{old_off[i], old_sz[i]});
  return ekf_propagation(P, N, new_off, p, order, Phi, Q);
}

int ovo_cov_clone(const double *Pin, int N, int old_off, int size, const double *dnc_dt, int dt_off, double *Pout) {
  cov_clone(Pin, N, old_off, size, dnc_dt, dt_off, Pout);
  return OVB_OK;
}

int ovo_cov_marginalize(const double *Pin, int N, int off, int size, double *Pout) {
  cov_marginalize(Pin, N, off, size, Pout);
  return OVB_OK;
}

int ovo_cov_get_marginal(const double *P, int N, const int *off, const int *sz, int nvar, double *out) {
  std::vector<Var> order;
  for (int i = 0; i < nvar; i++)
    order.push_back(Var{off[i], sz[i]});
  Mat S = get_marginal_covariance(P, N, order);
  for (int i = 0; i < S.r; i++)
    for (int j = 0; j < S.c; j++)
      out[(size_t)i * S.c + j] = S(i, j);
  return OVB_OK;
}

// single-function probes used by the numpy twin tests
void ovo_distort_d(int model, const double *cam_d, double x, double y, double *uv) { distort_d(model, cam_d, x, y, uv[0], uv[1]); }
void ovo_distort_jacobian(int model, const double *cam_d, double x, double y, double *dzn, double *dzeta) {
  distort_jacobian(model, cam_d, x, y, dzn, dzeta);
}
void ovo_make_givens(double p, double q, double *cs) {
  Givens g = make_givens(p, q);
  cs[0] = g.c;
  cs[1] = g.s;
}
void ovo_solve3(const double *A, const double *b, double *x) {
  V3 r = colpiv_qr_solve3(load_m3(A), load_v3(b));
  x[0] = r(0);
  x[1] = r(1);
  x[2] = r(2);
}
double ovo_cond3(const double *A) { return cond_sym3(load_m3(A)); }

} // extern "C"
