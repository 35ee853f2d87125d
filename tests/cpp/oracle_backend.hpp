// oracle_backend.hpp — TEST INFRASTRUCTURE: the CPU oracle (oracle/libovoracle.so) behind ovb200::CovBackend, so that the
// host runner of include/ovb200_vio.hpp can be driven with the reference's CPU arithmetic on the very same inputs the CUDA
// engine gets. Only tests/ may include this file (oracle/ is the checker, never the product path).
#pragma once
#include "../../include/ovb200_vio.hpp"

extern "C" {
int ovo_msckf_update(const ovb_frame *fr, const ovb_feat_batch *fb, const ovb_opts *op, const double *chi2_table, double *P, int N, ovb_feat_out *out,
                     double *dx, ovb_stats *stats, int32_t *order_off, int32_t *order_sz, int32_t *n_order, double *H_big, double *res_big, int cap_rows,
                     double *H_cmp, double *res_cmp, double *times);
int ovo_cov_propagate(double *P, int N, int new_off, int p, const int *old_off, const int *old_sz, int nold, const double *Phi, const double *Q);
int ovo_cov_clone(const double *Pin, int N, int old_off, int size, const double *dnc_dt, int dt_off, double *Pout);
int ovo_cov_marginalize(const double *Pin, int N, int off, int size, double *Pout);
int ovo_cov_get_marginal(const double *P, int N, const int *off, const int *sz, int nvar, double *out);
}

namespace ovb200 {
class OracleCov : public CovBackend {
public:
  OracleCov() {
    table_.resize(OVB_CHI2_TABLE_LEN);
    for (int k = 0; k < OVB_CHI2_TABLE_LEN; k++)
      table_[(size_t)k] = ovb_chi2_quantile95(k); // the table the product embeds (host function of libovb200.so)
  }
  int dim() override { return N_; }
  void set(const std::vector<double> &P, int N) override {
    P_ = P;
    N_ = N;
  }
  std::vector<double> get() override { return P_; }
  std::vector<double> get_marginal(const std::vector<int> &off, const std::vector<int> &sz) override {
    int n = 0;
    for (int s : sz)
      n += s;
    std::vector<double> out((size_t)n * n);
    ovo_cov_get_marginal(P_.data(), N_, off.data(), sz.data(), (int)off.size(), out.data());
    return out;
  }
  void clone(int old_off, int size, const double *dnc_dt, int dt_off) override {
    std::vector<double> Pn((size_t)(N_ + size) * (N_ + size));
    ovo_cov_clone(P_.data(), N_, old_off, size, dnc_dt, dt_off, Pn.data());
    P_.swap(Pn);
    N_ += size;
  }
  void marginalize(int off, int size) override {
    std::vector<double> Pn((size_t)(N_ - size) * (N_ - size));
    ovo_cov_marginalize(P_.data(), N_, off, size, Pn.data());
    P_.swap(Pn);
    N_ -= size;
  }
  void propagate(int new_off, int p, const std::vector<int> &old_off, const std::vector<int> &old_sz, const std::vector<double> &Phi,
                 const std::vector<double> &Q) override {
    const int st = ovo_cov_propagate(P_.data(), N_, new_off, p, old_off.data(), old_sz.data(), (int)old_off.size(), Phi.data(), Q.data());
    if (st != OVB_OK)
      throw Error((ovb_status)st, "oracle EKFPropagation failed");
  }
  int msckf_update(const ovb_frame *frame, const ovb_feat_batch *feats, const ovb_opts *opts, ovb_feat_out *out, double *dx, ovb_stats *stats) override {
    const int st = ovo_msckf_update(frame, feats, opts, table_.data(), P_.data(), N_, out, dx, stats, nullptr, nullptr, nullptr, nullptr, nullptr, 0,
                                    nullptr, nullptr, nullptr);
    if (st != OVB_OK)
      throw Error((ovb_status)st, "oracle msckf_update failed");
    return st;
  }

private:
  std::vector<double> P_, table_;
  int N_ = 0;
};
} // namespace ovb200
