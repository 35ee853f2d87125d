// ovb200_vio.hpp — the host-side callers either side of the MSCKF update, header-only C++17:
//   ov_msckf::Propagator                 state/Propagator.cpp:33-138 (propagate_and_clone), :269-393 (select_imu_readings),
//                                        :395-480 (predict_and_compute), :482-598 (mean: discrete / RK4), :600-681 (Xi sums,
//                                        analytic mean), :683-828 (analytic F, G), :830-950 (discrete F, G), :952-1015 (H_Dw/Da/Tg)
//   ov_msckf::State (mean + ids)         state/State.cpp:28-166, state/State.h:66-135
//   StateHelper::augment_clone           state/StateHelper.cpp:579-616; marginalize_old_clone :618-629; EKFUpdate's mean update :185-196
//   ov_core::FeatureDatabase             feat/FeatureDatabase.cpp:59-321        ov_core::TrackSIM   track/TrackSIM.cpp:30-79
//   VioManager                           core/VioManager.cpp:166-254 (feed_*), :323-644 (do_feature_propagate_update: feature
//                                        selection, sort, cap, update, cleanup, marginalize, timing CSV), VioManagerHelper.cpp:40-76
//   run_simulation                       ov_msckf/src/run_simulation.cpp:117-176 (init from ground truth, 1-frame delay buffer)
//   ov_eval ATE                          ov_eval/src/calc/ResultTrajectory.cpp:82-109 (alignment "none": the filter starts from truth)
// The covariance is behind CovBackend: the product backend is the CUDA engine (EngineCov, libovb200.so through the C ABI);
// tests plug the CPU oracle behind the same interface (tests/cpp/oracle_backend.hpp), so the two runs consume byte-identical
// inputs. Scope: MSCKF features only (max_slam = 0, no ZUPT, no ArUco), radtan cameras, KALIBR IMU model.
#ifndef OVB200_VIO_HPP
#define OVB200_VIO_HPP

#include "ovb200_host.hpp"
#include "ovb200_sim.hpp"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>

namespace ovb200 {

struct ImuData {
  double timestamp = 0;
  Vec3 wm{0, 0, 0}, am{0, 0, 0};
};

// ---------------------------------------------------------------------------------------------------------------------
// covariance residency + the update arithmetic, i.e. everything the reference does on State::_Cov
struct CovBackend {
  virtual ~CovBackend() {}
  virtual int dim() = 0;
  virtual void set(const std::vector<double> &P, int N) = 0;                                              // StateHelper::set_initial_covariance
  virtual std::vector<double> get() = 0;                                                                  // get_full_covariance
  virtual std::vector<double> get_marginal(const std::vector<int> &off, const std::vector<int> &sz) = 0;  // get_marginal_covariance
  virtual void clone(int old_off, int size, const double *dnc_dt, int dt_off) = 0;                        // clone + augment_clone's dt term
  virtual void marginalize(int off, int size) = 0;                                                        // marginalize
  virtual void propagate(int new_off, int p, const std::vector<int> &old_off, const std::vector<int> &old_sz, const std::vector<double> &Phi,
                         const std::vector<double> &Q) = 0;                                               // EKFPropagation
  virtual int msckf_update(const ovb_frame *frame, const ovb_feat_batch *feats, const ovb_opts *opts, ovb_feat_out *out, double *dx,
                           ovb_stats *stats) = 0;                                                         // UpdaterMSCKF::update steps 2-6
};

// product backend: the device-resident covariance of the CUDA engine
class EngineCov : public CovBackend {
public:
  explicit EngineCov(const ovb_config &cfg) {
    const ovb_status st = ovb_create(&cfg, &ctx_);
    if (st != OVB_OK)
      throw Error(st, std::string("ovb_create: ") + (ctx_ ? ovb_last_error(ctx_) : "no context (is a B200 visible?)"));
  }
  ~EngineCov() override {
    if (ctx_)
      ovb_destroy(ctx_);
  }
  ovb_ctx *ctx() const { return ctx_; }
  int dim() override { return ovb_cov_dim(ctx_); }
  void set(const std::vector<double> &P, int N) override { check(ovb_cov_set(ctx_, P.data(), N), "cov_set"); }
  std::vector<double> get() override {
    const int N = dim();
    std::vector<double> P((size_t)N * N);
    check(ovb_cov_get(ctx_, P.data(), N), "cov_get");
    return P;
  }
  std::vector<double> get_marginal(const std::vector<int> &off, const std::vector<int> &sz) override {
    int n = 0;
    for (int s : sz)
      n += s;
    std::vector<double> out((size_t)n * n);
    check(ovb_cov_get_marginal(ctx_, off.data(), sz.data(), (int)off.size(), out.data()), "cov_get_marginal");
    return out;
  }
  void clone(int old_off, int size, const double *dnc_dt, int dt_off) override { check(ovb_cov_clone(ctx_, old_off, size, dnc_dt, dt_off), "cov_clone"); }
  void marginalize(int off, int size) override { check(ovb_cov_marginalize(ctx_, off, size), "cov_marginalize"); }
  void propagate(int new_off, int p, const std::vector<int> &old_off, const std::vector<int> &old_sz, const std::vector<double> &Phi,
                 const std::vector<double> &Q) override {
    check(ovb_cov_propagate(ctx_, new_off, p, old_off.data(), old_sz.data(), (int)old_off.size(), Phi.data(), Q.data()), "cov_propagate");
  }
  int msckf_update(const ovb_frame *frame, const ovb_feat_batch *feats, const ovb_opts *opts, ovb_feat_out *out, double *dx, ovb_stats *stats) override {
    const ovb_status st = ovb_msckf_update(ctx_, frame, feats, opts, out, dx, stats);
    check(st, "msckf_update");
    return st;
  }

private:
  ovb_ctx *ctx_ = nullptr;
  void check(ovb_status st, const char *where) const {
    if (st != OVB_OK)
      throw Error(st, std::string(where) + ": " + ovb_last_error(ctx_));
  }
};

// ---------------------------------------------------------------------------------------------------------------------
enum IntegrationMethod { INTEGRATION_DISCRETE = 0, INTEGRATION_RK4 = 1, INTEGRATION_ANALYTICAL = 2 }; // StateOptions::IntegrationMethod

// StateOptions (state/StateOptions.h:35-176) + the estimator options of VioManagerOptions the runner reads; defaults =
// config/rpng_sim/estimator_config.yaml
struct VioOptions {
  bool do_fej = true;
  int integration_method = INTEGRATION_RK4;
  bool do_calib_camera_pose = true, do_calib_camera_intrinsics = true, do_calib_camera_timeoffset = true;
  bool do_calib_imu_intrinsics = true, do_calib_imu_g_sensitivity = true;
  int max_clone_size = 11;
  int max_msckf_in_update = 10;
  int num_cameras = 2;
  int feat_rep_msckf = OVB_REP_GLOBAL_3D;
  double gravity_mag = 9.81;
  double sigma_w = 1.6968e-04, sigma_wb = 1.9393e-05, sigma_a = 2.0000e-3, sigma_ab = 3.0000e-3; // NoiseManager (utils/NoiseManager.h)
  UpdaterOptions msckf_options{1.0, 1.0}; // up_msckf_chi2_multipler 1, up_msckf_sigma_px 1
  FeatureInitializerOptions featinit_options;
  int col_order = OVB_COLS_CANONICAL;
  int compress = OVB_COMPRESS_CHOLQR2;
  bool record_timing_information = false;
  std::string record_timing_filepath = "/tmp/traj_timing.txt";
};

struct ClonePose { // ov_type::PoseJPL of a clone: value + FEJ
  int id = -1;
  Vec4 q{0, 0, 0, 1}, q_fej{0, 0, 0, 1};
  Vec3 p{0, 0, 0}, p_fej{0, 0, 0};
};

// ov_msckf::State: the mean of every variable and its covariance id (state/State.cpp:28-131 fixes the order:
// IMU 15 | dw 6 | da 6 | tg 9 | R_GYROtoIMU 3 | dt 1 | per camera: extrinsics 6, intrinsics 8 | clones 6 ...)
struct VioState {
  VioOptions opt;
  double timestamp = -1;
  // ov_type::IMU (types/IMU.h): value and fej, [q(4) p(3) v(3) bg(3) ba(3)]
  Vec4 q{0, 0, 0, 1}, q_fej{0, 0, 0, 1};
  Vec3 p{0, 0, 0}, v{0, 0, 0}, bg{0, 0, 0}, ba{0, 0, 0}, p_fej{0, 0, 0}, v_fej{0, 0, 0};
  int imu_id = 0;
  double dw[6] = {1, 0, 0, 1, 0, 1}, da[6] = {1, 0, 0, 1, 0, 1}, tg[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  Vec4 q_GYROtoIMU{0, 0, 0, 1}, q_ACCtoIMU{0, 0, 0, 1};
  int dw_id = -1, da_id = -1, tg_id = -1, gyro_id = -1;
  double dt_CAMtoIMU = 0;
  int dt_id = -1;
  struct Cam {
    Vec4 q_ItoC{0, 0, 0, 1};
    Vec3 p_IinC{0, 0, 0};
    double intr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int ext_id = -1, intr_id = -1;
    CamRadtan model; // State::_cam_intrinsics_cameras: refreshed from intr after every update (StateHelper.cpp:192-196)
  };
  std::vector<Cam> cams;
  std::map<double, ClonePose> clones; // State::_clones_IMU
  int base_size = 0;

  int imu_intrinsic_size() const { // State.h:126-135
    int sz = 0;
    if (opt.do_calib_imu_intrinsics) {
      sz += 15;
      if (opt.do_calib_imu_g_sensitivity)
        sz += 9;
    }
    return sz;
  }
  double margtimestep() const { // State.h:66-75
    double time = std::numeric_limits<double>::infinity();
    for (const auto &c : clones)
      if (c.first < time)
        time = c.first;
    return time;
  }
  Mat3 Rot() const { return quat_2_Rot(q); }
  Mat3 Rot_fej() const { return quat_2_Rot(q_fej); }
};

// ---------------------------------------------------------------------------------------------------------------------
// ov_core::FeatureDatabase (feat/FeatureDatabase.cpp); iteration order of the unordered_map is part of the behaviour
class FeatureDatabase {
public:
  std::unordered_map<size_t, std::shared_ptr<Feature>> features_idlookup;

  std::shared_ptr<Feature> get_feature(size_t id) {
    auto it = features_idlookup.find(id);
    return it == features_idlookup.end() ? nullptr : it->second;
  }
  void update_feature(size_t id, double timestamp, size_t cam_id, float u, float v, float u_n, float v_n) { // :59-85
    auto it = features_idlookup.find(id);
    std::shared_ptr<Feature> feat;
    if (it != features_idlookup.end()) {
      feat = it->second;
    } else {
      feat = std::make_shared<Feature>();
      feat->featid = id;
    }
    feat->uvs[cam_id].push_back({u, v});
    feat->uvs_norm[cam_id].push_back({u_n, v_n});
    feat->timestamps[cam_id].push_back(timestamp);
    if (it == features_idlookup.end())
      features_idlookup[id] = feat;
  }
  std::vector<std::shared_ptr<Feature>> features_not_containing_newer(double timestamp, bool remove = false, bool skip_deleted = false) { // :87-125
    std::vector<std::shared_ptr<Feature>> feats_old;
    for (auto it = features_idlookup.begin(); it != features_idlookup.end();) {
      if (skip_deleted && it->second->to_delete) {
        ++it;
        continue;
      }
      bool has_newer_measurement = false;
      for (auto const &pair : it->second->timestamps) {
        has_newer_measurement = (!pair.second.empty() && pair.second.at(pair.second.size() - 1) >= timestamp);
        if (has_newer_measurement)
          break;
      }
      if (!has_newer_measurement) {
        feats_old.push_back(it->second);
        if (remove)
          it = features_idlookup.erase(it);
        else
          ++it;
      } else {
        ++it;
      }
    }
    return feats_old;
  }
  std::vector<std::shared_ptr<Feature>> features_containing(double timestamp, bool remove = false, bool skip_deleted = false) { // :169-207
    std::vector<std::shared_ptr<Feature>> feats_has_timestamp;
    for (auto it = features_idlookup.begin(); it != features_idlookup.end();) {
      if (skip_deleted && it->second->to_delete) {
        ++it;
        continue;
      }
      bool has_timestamp = false;
      for (auto const &pair : it->second->timestamps) {
        has_timestamp = (std::find(pair.second.begin(), pair.second.end(), timestamp) != pair.second.end());
        if (has_timestamp)
          break;
      }
      if (has_timestamp) {
        feats_has_timestamp.push_back(it->second);
        if (remove)
          it = features_idlookup.erase(it);
        else
          ++it;
      } else {
        ++it;
      }
    }
    return feats_has_timestamp;
  }
  void cleanup() { // :211-221
    for (auto it = features_idlookup.begin(); it != features_idlookup.end();) {
      if (it->second->to_delete)
        it = features_idlookup.erase(it);
      else
        ++it;
    }
  }
  // cleanup_measurements (:223-240) with Feature::clean_older_measurements (feat/Feature.cpp:81-110): drop measurements
  // strictly older than `timestamp`, then features without any
  void cleanup_measurements(double timestamp) {
    for (auto it = features_idlookup.begin(); it != features_idlookup.end();) {
      Feature &f = *it->second;
      int ct_meas = 0;
      for (auto &pair : f.timestamps) {
        auto &ts = pair.second;
        auto &uv = f.uvs[pair.first];
        auto &uvn = f.uvs_norm[pair.first];
        size_t w = 0;
        for (size_t i = 0; i < ts.size(); i++) {
          if (!(ts[i] < timestamp)) {
            ts[w] = ts[i];
            uv[w] = uv[i];
            uvn[w] = uvn[i];
            w++;
          }
        }
        ts.resize(w);
        uv.resize(w);
        uvn.resize(w);
        ct_meas += (int)w;
      }
      if (ct_meas < 1)
        it = features_idlookup.erase(it);
      else
        ++it;
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// ov_msckf::Propagator: IMU buffer, mean propagation and the state-transition / noise matrices; the covariance step goes
// through CovBackend::propagate (StateHelper::EKFPropagation) and CovBackend::clone (augment_clone).
class Propagator {
public:
  explicit Propagator(double gravity_mag) : gravity_{0.0, 0.0, gravity_mag} {}

  void feed_imu(const ImuData &message, double oldest_time = -1) { // Propagator.h:65-74
    imu_data.push_back(message);
    clean_old_imu_measurements(oldest_time - 0.10);
  }
  void clean_old_imu_measurements(double oldest_time) { // Propagator.h:80-92
    if (oldest_time < 0)
      return;
    auto it0 = imu_data.begin();
    while (it0 != imu_data.end()) {
      if (it0->timestamp < oldest_time)
        it0 = imu_data.erase(it0);
      else
        ++it0;
    }
  }

  static ImuData interpolate_data(const ImuData &imu_1, const ImuData &imu_2, double timestamp) { // Propagator.h:154-164
    const double lambda = (timestamp - imu_1.timestamp) / (imu_2.timestamp - imu_1.timestamp);
    ImuData data;
    data.timestamp = timestamp;
    data.am = (1 - lambda) * imu_1.am + lambda * imu_2.am;
    data.wm = (1 - lambda) * imu_1.wm + lambda * imu_2.wm;
    return data;
  }

  static std::vector<ImuData> select_imu_readings(const std::vector<ImuData> &imu_data, double time0, double time1) { // Propagator.cpp:269-393
    std::vector<ImuData> prop_data;
    if (imu_data.empty())
      return prop_data;
    for (size_t i = 0; i + 1 < imu_data.size(); i++) {
      if (imu_data[i + 1].timestamp > time0 && imu_data[i].timestamp < time0) {
        prop_data.push_back(interpolate_data(imu_data[i], imu_data[i + 1], time0));
        continue;
      }
      if (imu_data[i].timestamp >= time0 && imu_data[i + 1].timestamp <= time1) {
        prop_data.push_back(imu_data[i]);
        continue;
      }
      if (imu_data[i + 1].timestamp > time1) {
        if (imu_data[i].timestamp > time1 && i == 0) {
          break;
        } else if (imu_data[i].timestamp > time1) {
          prop_data.push_back(interpolate_data(imu_data[i - 1], imu_data[i], time1));
        } else {
          prop_data.push_back(imu_data[i]);
        }
        if (prop_data.back().timestamp != time1)
          prop_data.push_back(interpolate_data(imu_data[i], imu_data[i + 1], time1));
        break;
      }
    }
    if (prop_data.empty())
      return prop_data;
    if (prop_data.back().timestamp != time1)
      prop_data.push_back(interpolate_data(imu_data[imu_data.size() - 2], imu_data[imu_data.size() - 1], time1));
    for (size_t i = 0; i + 1 < prop_data.size(); i++) {
      if (std::abs(prop_data[i + 1].timestamp - prop_data[i].timestamp) < 1e-12) {
        prop_data.erase(prop_data.begin() + (long)i);
        i--;
      }
    }
    if (prop_data.size() < 2)
      prop_data.clear();
    return prop_data;
  }

  // Propagator::propagate_and_clone (Propagator.cpp:33-138)
  void propagate_and_clone(VioState &state, CovBackend &cov, double timestamp) {
    if (state.timestamp == timestamp)
      throw Error(OVB_ERR_ARG, "Propagator::propagate_and_clone(): propagation called again at the same timestep");
    if (state.timestamp > timestamp)
      throw Error(OVB_ERR_ARG, "Propagator::propagate_and_clone(): propagation called trying to propagate backwards in time");
    if (!have_last_prop_time_offset) {
      last_prop_time_offset = state.dt_CAMtoIMU;
      have_last_prop_time_offset = true;
    }
    const double t_off_new = state.dt_CAMtoIMU;
    const double time0 = state.timestamp + last_prop_time_offset;
    const double time1 = timestamp + t_off_new;
    const std::vector<ImuData> prop_data = select_imu_readings(imu_data, time0, time1);
    const int n = state.imu_intrinsic_size() + 15;
    std::vector<double> Phi_summed((size_t)n * n, 0.0), Qd_summed((size_t)n * n, 0.0), tmp((size_t)n * n), tmp2((size_t)n * n);
    for (int i = 0; i < n; i++)
      Phi_summed[(size_t)i * n + i] = 1.0;
    if (prop_data.size() > 1) {
      for (size_t i = 0; i + 1 < prop_data.size(); i++) {
        std::vector<double> F, Qdi;
        predict_and_compute(state, prop_data[i], prop_data[i + 1], F, Qdi);
        // Phi_summed = F * Phi_summed; Qd_summed = F * Qd_summed * F' + Qdi, symmetrised (:91-99)
        matmul(F, Phi_summed, tmp, n);
        Phi_summed = tmp;
        matmul(F, Qd_summed, tmp, n);
        matmul_bt(tmp, F, tmp2, n);
        for (int a = 0; a < n * n; a++)
          tmp2[(size_t)a] += Qdi[(size_t)a];
        for (int a = 0; a < n; a++)
          for (int b = 0; b < n; b++)
            Qd_summed[(size_t)a * n + b] = 0.5 * (tmp2[(size_t)a * n + b] + tmp2[(size_t)b * n + a]);
      }
    }
    // last angular velocity for the clone's time-offset Jacobian (:104-113)
    Vec3 last_w{0, 0, 0};
    if (!prop_data.empty()) {
      const Mat3 Dw = Simulator::Dm(state.dw), Da = Simulator::Dm(state.da), Tg = Simulator::Tgm(state.tg);
      const Vec3 last_a = quat_2_Rot(state.q_ACCtoIMU) * (Da * (prop_data.back().am - state.ba));
      last_w = quat_2_Rot(state.q_GYROtoIMU) * (Dw * (prop_data.back().wm - state.bg - Tg * last_a));
    }
    // covariance: EKFPropagation over [imu | dw | da | tg | R_GYROtoIMU] (:115-130); all contiguous from the IMU id
    std::vector<int> off{state.imu_id}, sz{15};
    if (state.opt.do_calib_imu_intrinsics) {
      off.push_back(state.dw_id), sz.push_back(6);
      off.push_back(state.da_id), sz.push_back(6);
      if (state.opt.do_calib_imu_g_sensitivity)
        off.push_back(state.tg_id), sz.push_back(9);
      off.push_back(state.gyro_id), sz.push_back(3);
    }
    cov.propagate(state.imu_id, n, off, sz, Phi_summed, Qd_summed);
    state.timestamp = timestamp;
    last_prop_time_offset = t_off_new;
    // StateHelper::augment_clone (StateHelper.cpp:579-616): clone the IMU pose (value and fej), dt Jacobian [last_w; v]
    ClonePose c;
    c.id = cov.dim();
    c.q = state.q, c.p = state.p, c.q_fej = state.q_fej, c.p_fej = state.p_fej;
    if (state.clones.count(state.timestamp))
      throw Error(OVB_ERR_ARG, "augment_clone: tried to insert a clone at the time of an existing clone");
    double dnc_dt[6] = {last_w[0], last_w[1], last_w[2], state.v[0], state.v[1], state.v[2]};
    cov.clone(state.imu_id, 6, state.opt.do_calib_camera_timeoffset ? dnc_dt : nullptr, state.dt_id);
    state.clones[state.timestamp] = c;
  }

  // Propagator::predict_and_compute (Propagator.cpp:395-480). F, Qd are n x n row-major, n = 15 + imu intrinsics.
  void predict_and_compute(VioState &state, const ImuData &data_minus, const ImuData &data_plus, std::vector<double> &F, std::vector<double> &Qd) {
    const double dt = data_plus.timestamp - data_minus.timestamp;
    const Mat3 Dw = Simulator::Dm(state.dw), Da = Simulator::Dm(state.da), Tg = Simulator::Tgm(state.tg);
    Vec3 a_hat1 = data_minus.am - state.ba, a_hat2 = data_plus.am - state.ba;
    Vec3 a_hat_avg = .5 * (a_hat1 + a_hat2);
    const Vec3 a_uncorrected = a_hat_avg;
    const Mat3 R_ACCtoIMU = quat_2_Rot(state.q_ACCtoIMU);
    a_hat1 = R_ACCtoIMU * (Da * a_hat1);
    a_hat2 = R_ACCtoIMU * (Da * a_hat2);
    a_hat_avg = R_ACCtoIMU * (Da * a_hat_avg);
    Vec3 w_hat1 = data_minus.wm - state.bg - Tg * a_hat1, w_hat2 = data_plus.wm - state.bg - Tg * a_hat2;
    Vec3 w_hat_avg = .5 * (w_hat1 + w_hat2);
    const Vec3 w_uncorrected = w_hat_avg;
    const Mat3 R_GYROtoIMU = quat_2_Rot(state.q_GYROtoIMU);
    w_hat1 = R_GYROtoIMU * (Dw * w_hat1);
    w_hat2 = R_GYROtoIMU * (Dw * w_hat2);
    w_hat_avg = R_GYROtoIMU * (Dw * w_hat_avg);
    XiSum Xi;
    const bool analytic = state.opt.integration_method == INTEGRATION_RK4 || state.opt.integration_method == INTEGRATION_ANALYTICAL;
    if (analytic)
      compute_Xi_sum(dt, w_hat_avg, a_hat_avg, Xi);
    Vec4 new_q;
    Vec3 new_v, new_p;
    if (state.opt.integration_method == INTEGRATION_ANALYTICAL)
      predict_mean_analytic(state, dt, a_hat_avg, new_q, new_v, new_p, Xi);
    else if (state.opt.integration_method == INTEGRATION_RK4)
      predict_mean_rk4(state, dt, w_hat1, a_hat1, w_hat2, a_hat2, new_q, new_v, new_p);
    else
      predict_mean_discrete(state, dt, w_hat_avg, a_hat_avg, new_q, new_v, new_p);
    const int n = state.imu_intrinsic_size() + 15;
    F.assign((size_t)n * n, 0.0);
    std::vector<double> G((size_t)n * 12, 0.0);
    compute_F_and_G(state, analytic, dt, w_uncorrected, a_uncorrected, new_q, new_v, new_p, Xi, F, G, n);
    // Qd = G Qc G' with Qc = diag(sigma^2 / dt) (:453-464), symmetrised
    const double qc[4] = {state.opt.sigma_w * state.opt.sigma_w / dt, state.opt.sigma_a * state.opt.sigma_a / dt, state.opt.sigma_wb * state.opt.sigma_wb / dt,
                          state.opt.sigma_ab * state.opt.sigma_ab / dt};
    std::vector<double> Qt((size_t)n * n, 0.0);
    for (int a = 0; a < n; a++)
      for (int b = 0; b < n; b++) {
        double s = 0;
        for (int k = 0; k < 12; k++)
          s += G[(size_t)a * 12 + k] * qc[k / 3] * G[(size_t)b * 12 + k];
        Qt[(size_t)a * n + b] = s;
      }
    Qd.assign((size_t)n * n, 0.0);
    for (int a = 0; a < n; a++)
      for (int b = 0; b < n; b++)
        Qd[(size_t)a * n + b] = 0.5 * (Qt[(size_t)a * n + b] + Qt[(size_t)b * n + a]);
    // replace the IMU estimate and its FEJ with the propagated values (:471-479)
    state.q = state.q_fej = new_q;
    state.p = state.p_fej = new_p;
    state.v = state.v_fej = new_v;
  }

  std::vector<ImuData> imu_data;

private:
  Vec3 gravity_;
  bool have_last_prop_time_offset = false;
  double last_prop_time_offset = 0;
  struct XiSum {
    Mat3 R_ktok1 = eye3(), Xi_1 = zero3(), Xi_2 = zero3(), Jr_ktok1 = eye3(), Xi_3 = zero3(), Xi_4 = zero3();
  };

  static void matmul(const std::vector<double> &A, const std::vector<double> &B, std::vector<double> &C, int n) {
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        double s = 0;
        for (int k = 0; k < n; k++)
          s += A[(size_t)i * n + k] * B[(size_t)k * n + j];
        C[(size_t)i * n + j] = s;
      }
  }
  static void matmul_bt(const std::vector<double> &A, const std::vector<double> &B, std::vector<double> &C, int n) { // C = A B'
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        double s = 0;
        for (int k = 0; k < n; k++)
          s += A[(size_t)i * n + k] * B[(size_t)j * n + k];
        C[(size_t)i * n + j] = s;
      }
  }
  static void put3(std::vector<double> &M, int ld, int r, int c, const Mat3 &B) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        M[(size_t)(r + i) * ld + c + j] = B[(size_t)i * 3 + j];
  }

  void predict_mean_discrete(const VioState &state, double dt, const Vec3 &w_hat, const Vec3 &a_hat, Vec4 &new_q, Vec3 &new_v, Vec3 &new_p) const { // :482-508
    const double w_norm = norm(w_hat);
    const Mat3 R_Gtoi = state.Rot();
    Vec4 bq;
    const Vec4 Oq = Omega_times(w_hat, state.q);
    if (w_norm > 1e-12) {
      const double c = std::cos(0.5 * w_norm * dt), s = 1 / w_norm * std::sin(0.5 * w_norm * dt);
      for (int i = 0; i < 4; i++)
        bq[(size_t)i] = c * state.q[(size_t)i] + s * Oq[(size_t)i];
    } else {
      for (int i = 0; i < 4; i++)
        bq[(size_t)i] = state.q[(size_t)i] + 0.5 * dt * Oq[(size_t)i];
    }
    new_q = quatnorm(bq);
    new_v = state.v + transpose(R_Gtoi) * a_hat * dt - gravity_ * dt;
    new_p = state.p + state.v * dt + 0.5 * (transpose(R_Gtoi) * a_hat) * dt * dt - 0.5 * gravity_ * dt * dt;
  }

  void predict_mean_rk4(const VioState &state, double dt, const Vec3 &w_hat1, const Vec3 &a_hat1, const Vec3 &w_hat2, const Vec3 &a_hat2, Vec4 &new_q,
                        Vec3 &new_v, Vec3 &new_p) const { // :510-598
    Vec3 w_hat = w_hat1, a_hat = a_hat1;
    const Vec3 w_alpha = (1.0 / dt) * (w_hat2 - w_hat1), a_jerk = (1.0 / dt) * (a_hat2 - a_hat1);
    const Vec4 q_0 = state.q;
    const Vec3 p_0 = state.p, v_0 = state.v;
    auto scale4 = [](double s, const Vec4 &a) { return Vec4{s * a[0], s * a[1], s * a[2], s * a[3]}; };
    auto add4 = [](const Vec4 &a, const Vec4 &b) { return Vec4{a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]}; };
    // k1
    const Vec4 dq_0{0, 0, 0, 1};
    const Vec4 q0_dot = scale4(0.5, Omega_times(w_hat, dq_0));
    const Vec3 p0_dot = v_0;
    const Mat3 R_Gto0 = quat_2_Rot(quat_multiply(dq_0, q_0));
    const Vec3 v0_dot = transpose(R_Gto0) * a_hat - gravity_;
    const Vec4 k1_q = scale4(dt, q0_dot);
    const Vec3 k1_p = p0_dot * dt, k1_v = v0_dot * dt;
    // k2
    w_hat += 0.5 * w_alpha * dt;
    a_hat += 0.5 * a_jerk * dt;
    const Vec4 dq_1 = quatnorm(add4(dq_0, scale4(0.5, k1_q)));
    const Vec3 v_1 = v_0 + 0.5 * k1_v;
    const Vec4 q1_dot = scale4(0.5, Omega_times(w_hat, dq_1));
    const Vec3 p1_dot = v_1;
    const Mat3 R_Gto1 = quat_2_Rot(quat_multiply(dq_1, q_0));
    const Vec3 v1_dot = transpose(R_Gto1) * a_hat - gravity_;
    const Vec4 k2_q = scale4(dt, q1_dot);
    const Vec3 k2_p = p1_dot * dt, k2_v = v1_dot * dt;
    // k3
    const Vec4 dq_2 = quatnorm(add4(dq_0, scale4(0.5, k2_q)));
    const Vec3 v_2 = v_0 + 0.5 * k2_v;
    const Vec4 q2_dot = scale4(0.5, Omega_times(w_hat, dq_2));
    const Vec3 p2_dot = v_2;
    const Mat3 R_Gto2 = quat_2_Rot(quat_multiply(dq_2, q_0));
    const Vec3 v2_dot = transpose(R_Gto2) * a_hat - gravity_;
    const Vec4 k3_q = scale4(dt, q2_dot);
    const Vec3 k3_p = p2_dot * dt, k3_v = v2_dot * dt;
    // k4
    w_hat += 0.5 * w_alpha * dt;
    a_hat += 0.5 * a_jerk * dt;
    const Vec4 dq_3 = quatnorm(add4(dq_0, k3_q));
    const Vec3 v_3 = v_0 + k3_v;
    const Vec4 q3_dot = scale4(0.5, Omega_times(w_hat, dq_3));
    const Vec3 p3_dot = v_3;
    const Mat3 R_Gto3 = quat_2_Rot(quat_multiply(dq_3, q_0));
    const Vec3 v3_dot = transpose(R_Gto3) * a_hat - gravity_;
    const Vec4 k4_q = scale4(dt, q3_dot);
    const Vec3 k4_p = p3_dot * dt, k4_v = v3_dot * dt;
    // y+dt
    const Vec4 dq = quatnorm(add4(add4(add4(add4(dq_0, scale4(1.0 / 6.0, k1_q)), scale4(1.0 / 3.0, k2_q)), scale4(1.0 / 3.0, k3_q)), scale4(1.0 / 6.0, k4_q)));
    new_q = quat_multiply(dq, q_0);
    new_p = p_0 + (1.0 / 6.0) * k1_p + (1.0 / 3.0) * k2_p + (1.0 / 3.0) * k3_p + (1.0 / 6.0) * k4_p;
    new_v = v_0 + (1.0 / 6.0) * k1_v + (1.0 / 3.0) * k2_v + (1.0 / 3.0) * k3_v + (1.0 / 6.0) * k4_v;
  }

  static void compute_Xi_sum(double dt, const Vec3 &w_hat, const Vec3 &a_hat, XiSum &X) { // :600-667
    const double w_norm = norm(w_hat), d_th = w_norm * dt;
    Vec3 k_hat{0, 0, 0};
    if (w_norm > 1e-12)
      k_hat = (1.0 / w_norm) * w_hat;
    const Mat3 I = eye3();
    const double d_t2 = std::pow(dt, 2), d_t3 = std::pow(dt, 3), w_norm2 = std::pow(w_norm, 2), w_norm3 = std::pow(w_norm, 3);
    const double cos_dth = std::cos(d_th), sin_dth = std::sin(d_th), d_th2 = std::pow(d_th, 2), d_th3 = std::pow(d_th, 3);
    const Mat3 sK = skew_x(k_hat), sK2 = sK * sK, sA = skew_x(a_hat);
    X.R_ktok1 = exp_so3(-(w_hat * dt));
    X.Jr_ktok1 = Jr_so3(-(w_hat * dt));
    const bool small_w = (w_norm < 1.0 / 180 * M_PI / 2);
    const double ka = dot(k_hat, a_hat);
    if (!small_w) {
      X.Xi_1 = dt * I + ((1.0 - cos_dth) / w_norm) * sK + (dt - sin_dth / w_norm) * sK2;
      X.Xi_2 = (1.0 / 2 * d_t2) * I + ((d_th - sin_dth) / w_norm2) * sK + (1.0 / 2 * d_t2 - (1.0 - cos_dth) / w_norm2) * sK2;
      X.Xi_3 = (1.0 / 2 * d_t2) * sA + ((sin_dth - d_th) / w_norm2) * (sA * sK) + ((sin_dth - d_th * cos_dth) / w_norm2) * (sK * sA) +
               (1.0 / 2 * d_t2 - (1.0 - cos_dth) / w_norm2) * (sA * sK2) +
               (1.0 / 2 * d_t2 + (1.0 - cos_dth - d_th * sin_dth) / w_norm2) * (sK2 * sA + ka * sK) -
               ((3 * sin_dth - 2 * d_th - d_th * cos_dth) / w_norm2 * ka) * sK2;
      X.Xi_4 = (1.0 / 6 * d_t3) * sA + ((2 * (1.0 - cos_dth) - d_th2) / (2 * w_norm3)) * (sA * sK) +
               ((2 * (1.0 - cos_dth) - d_th * sin_dth) / w_norm3) * (sK * sA) + ((sin_dth - d_th) / w_norm3 + d_t3 / 6) * (sA * sK2) +
               ((d_th - 2 * sin_dth + 1.0 / 6 * d_th3 + d_th * cos_dth) / w_norm3) * (sK2 * sA + ka * sK) +
               ((4 * cos_dth - 4 + d_th2 + d_th * sin_dth) / w_norm3 * ka) * sK2;
    } else {
      X.Xi_1 = dt * (I + sin_dth * sK + (1.0 - cos_dth) * sK2);
      X.Xi_2 = (1.0 / 2 * dt) * X.Xi_1;
      X.Xi_3 = (1.0 / 2 * d_t2) * (sA + sin_dth * (-(sA * sK) + sK * sA + ka * sK2) + (1.0 - cos_dth) * (sA * sK2 + sK2 * sA + ka * sK));
      X.Xi_4 = (1.0 / 3 * dt) * X.Xi_3;
    }
  }

  void predict_mean_analytic(const VioState &state, double dt, const Vec3 &a_hat, Vec4 &new_q, Vec3 &new_v, Vec3 &new_p, const XiSum &X) const { // :669-681
    const Mat3 R_Gtok = state.Rot();
    const Vec4 q_ktok1 = rot_2_quat(X.R_ktok1);
    new_q = quat_multiply(q_ktok1, state.q);
    new_v = state.v + transpose(R_Gtok) * (X.Xi_1 * a_hat) - gravity_ * dt;
    new_p = state.p + state.v * dt + transpose(R_Gtok) * (X.Xi_2 * a_hat) - 0.5 * gravity_ * dt * dt;
  }

  // compute_F_and_G_analytic (:683-828) / compute_F_and_G_discrete (:830-950); KALIBR model (th_wtoI block present)
  void compute_F_and_G(const VioState &state, bool analytic, double dt, const Vec3 &w_uncorrected, const Vec3 &a_uncorrected, const Vec4 &new_q,
                       const Vec3 &new_v, const Vec3 &new_p, const XiSum &X, std::vector<double> &F, std::vector<double> &G, int n) const {
    const int th_id = 0, p_id = 3, v_id = 6, bg_id = 9, ba_id = 12;
    int Dw_id = -1, Da_id = -1, Tg_id = -1, th_wtoI_id = -1, local = 15;
    if (state.opt.do_calib_imu_intrinsics) {
      Dw_id = local, local += 6;
      Da_id = local, local += 6;
      if (state.opt.do_calib_imu_g_sensitivity)
        Tg_id = local, local += 9;
      th_wtoI_id = local, local += 3;
    }
    Mat3 R_k = state.Rot();
    Vec3 v_k = state.v, p_k = state.p;
    if (state.opt.do_fej) {
      R_k = state.Rot_fej();
      v_k = state.v_fej;
      p_k = state.p_fej;
    }
    const Mat3 dR_ktok1 = quat_2_Rot(new_q) * transpose(R_k);
    const Mat3 Dw = Simulator::Dm(state.dw), Da = Simulator::Dm(state.da), Tg = Simulator::Tgm(state.tg);
    const Mat3 R_atoI = quat_2_Rot(state.q_ACCtoIMU), R_wtoI = quat_2_Rot(state.q_GYROtoIMU);
    const Vec3 a_k = R_atoI * (Da * a_uncorrected);
    const Vec3 w_k = R_wtoI * (Dw * w_uncorrected);
    const Mat3 Rkt = transpose(R_k);
    const Mat3 Jr = analytic ? X.Jr_ktok1 : Jr_so3(log_so3(dR_ktok1));
    const Mat3 dRJdt = dt * (dR_ktok1 * Jr);
    const Mat3 RwDw = R_wtoI * Dw, RaDa = R_atoI * Da;
    put3(F, n, th_id, th_id, dR_ktok1);
    put3(F, n, p_id, th_id, -(skew_x(new_p - p_k - v_k * dt + 0.5 * gravity_ * dt * dt) * Rkt));
    put3(F, n, v_id, th_id, -(skew_x(new_v - v_k + gravity_ * dt) * Rkt));
    put3(F, n, p_id, p_id, eye3());
    put3(F, n, p_id, v_id, dt * eye3());
    put3(F, n, v_id, v_id, eye3());
    put3(F, n, bg_id, bg_id, eye3());
    put3(F, n, ba_id, ba_id, eye3());
    put3(F, n, th_id, bg_id, -(dRJdt * RwDw));
    put3(F, n, th_id, ba_id, dRJdt * RwDw * Tg * RaDa);
    Mat3 P_a, V_a, P_w, V_w; // position / velocity sensitivities to (corrected) acceleration and angular-rate perturbations
    if (analytic) {
      P_w = Rkt * X.Xi_4, V_w = Rkt * X.Xi_3;
      P_a = Rkt * (X.Xi_2 + X.Xi_4 * RwDw * Tg), V_a = Rkt * (X.Xi_1 + X.Xi_3 * RwDw * Tg);
      put3(F, n, p_id, bg_id, P_w * RwDw);
      put3(F, n, v_id, bg_id, V_w * RwDw);
    } else {
      P_w = zero3(), V_w = zero3();
      P_a = (0.5 * dt * dt) * Rkt, V_a = dt * Rkt;
    }
    put3(F, n, p_id, ba_id, -(P_a * RaDa));
    put3(F, n, v_id, ba_id, -(V_a * RaDa));
    auto put3xk = [&](int r, int c, const Mat3 &L, const double *H, int k, double sign) { // F[r.., c..] = sign * L (3x3) * H (3 x k)
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < k; j++)
          F[(size_t)(r + i) * n + c + j] = sign * (L[(size_t)i * 3] * H[j] + L[(size_t)i * 3 + 1] * H[k + j] + L[(size_t)i * 3 + 2] * H[2 * k + j]);
    };
    if (Dw_id != -1) { // compute_H_Dw (:952-971), KALIBR: [w1 I, w2 e2, w2 e3, w3 e3]
      const double w1 = w_uncorrected[0], w2 = w_uncorrected[1], w3 = w_uncorrected[2];
      const double H[18] = {w1, 0, 0, 0, 0, 0, 0, w1, 0, w2, 0, 0, 0, 0, w1, 0, w2, w3};
      put3xk(th_id, Dw_id, dRJdt * R_wtoI, H, 6, 1.0);
      if (analytic) {
        put3xk(p_id, Dw_id, P_w * R_wtoI, H, 6, -1.0);
        put3xk(v_id, Dw_id, V_w * R_wtoI, H, 6, -1.0);
      }
      for (int i = 0; i < 6; i++)
        F[(size_t)(Dw_id + i) * n + Dw_id + i] = 1.0;
    }
    if (Da_id != -1) { // compute_H_Da (:973-992)
      const double a1 = a_uncorrected[0], a2 = a_uncorrected[1], a3 = a_uncorrected[2];
      const double H[18] = {a1, 0, 0, 0, 0, 0, 0, a1, 0, a2, 0, 0, 0, 0, a1, 0, a2, a3};
      // the discrete variant omits Dw in the orientation block (:905): -dR Jr dt R_wtoI Tg R_atoI H_Da
      put3xk(th_id, Da_id, analytic ? dRJdt * RwDw * Tg * R_atoI : dRJdt * R_wtoI * Tg * R_atoI, H, 6, -1.0);
      put3xk(p_id, Da_id, P_a * R_atoI, H, 6, 1.0);
      put3xk(v_id, Da_id, V_a * R_atoI, H, 6, 1.0);
      for (int i = 0; i < 6; i++)
        F[(size_t)(Da_id + i) * n + Da_id + i] = 1.0;
    }
    if (Tg_id != -1) { // compute_H_Tg (:994-1015): [a1 I, a2 I, a3 I]
      const double H[27] = {a_k[0], 0, 0, a_k[1], 0, 0, a_k[2], 0, 0, 0, a_k[0], 0, 0, a_k[1], 0, 0, a_k[2], 0, 0, 0, a_k[0], 0, 0, a_k[1], 0, 0, a_k[2]};
      put3xk(th_id, Tg_id, dRJdt * RwDw, H, 9, -1.0);
      if (analytic) {
        put3xk(p_id, Tg_id, P_w * RwDw, H, 9, 1.0);
        put3xk(v_id, Tg_id, V_w * RwDw, H, 9, 1.0);
      }
      for (int i = 0; i < 9; i++)
        F[(size_t)(Tg_id + i) * n + Tg_id + i] = 1.0;
    }
    if (th_wtoI_id != -1) {
      put3(F, n, th_id, th_wtoI_id, dRJdt * skew_x(w_k));
      if (analytic) {
        put3(F, n, p_id, th_wtoI_id, -(P_w * skew_x(w_k)));
        put3(F, n, v_id, th_wtoI_id, -(V_w * skew_x(w_k)));
      }
      put3(F, n, th_wtoI_id, th_wtoI_id, eye3());
    }
    // G: columns [n_w | n_a | n_wb | n_ab]
    put3(G, 12, th_id, 0, -(dRJdt * RwDw));
    put3(G, 12, th_id, 3, dRJdt * RwDw * Tg * RaDa);
    if (analytic) {
      put3(G, 12, p_id, 0, P_w * RwDw);
      put3(G, 12, v_id, 0, V_w * RwDw);
    }
    put3(G, 12, p_id, 3, -(P_a * RaDa));
    put3(G, 12, v_id, 3, -(V_a * RaDa));
    put3(G, 12, bg_id, 6, dt * eye3());
    put3(G, 12, ba_id, 9, dt * eye3());
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// timing row of VioManager (core/VioManager.cpp:604-644) and trajectory sample for the evaluation
struct FrameTiming {
  double timestamp_inI = 0, time_track = 0, time_prop = 0, time_msckf = 0, time_marg = 0, time_total = 0;
  int feats_in = 0, feats_used = 0, rows = 0, cols = 0;
};
struct TrajSample {
  double t = 0;
  Vec4 q{0, 0, 0, 1};
  Vec3 p{0, 0, 0};
};

// VioManager reduced to the rpng_sim path (TrackSIM front-end, ground-truth initialisation, MSCKF updates)
class VioManager {
public:
  VioState state;
  std::shared_ptr<CovBackend> cov;
  Propagator propagator;
  FeatureDatabase database;
  std::vector<FrameTiming> timing;
  std::vector<TrajSample> trajectory_est;
  ovb_stats last_stats{};
  // optional hook: called with the marshalled inputs of every MSCKF update before it runs (capture of update cases)
  std::function<void(const ovb_frame &, const ovb_feat_batch &, const ovb_opts &, int frame_index)> on_update;
  int frames_done = 0;
  long status_hist[16] = {0}; // ovb_feat_status histogram over all updates (diagnostics)

  VioManager(const VioOptions &opt, const SimParams &calib, std::shared_ptr<CovBackend> backend)
      : cov(std::move(backend)), propagator(opt.gravity_mag) {
    // State::State (state/State.cpp:28-131): variable order and ids
    state.opt = opt;
    int id = 0;
    state.imu_id = id, id += 15;
    if (opt.do_calib_imu_intrinsics) {
      state.dw_id = id, id += 6;
      state.da_id = id, id += 6;
      if (opt.do_calib_imu_g_sensitivity)
        state.tg_id = id, id += 9;
      state.gyro_id = id, id += 3; // KALIBR: R_GYROtoIMU
    }
    if (opt.do_calib_camera_timeoffset)
      state.dt_id = id, id += 1;
    state.cams.resize((size_t)opt.num_cameras);
    for (int i = 0; i < opt.num_cameras; i++) {
      auto &c = state.cams[(size_t)i];
      if (opt.do_calib_camera_pose)
        c.ext_id = id, id += 6;
      if (opt.do_calib_camera_intrinsics)
        c.intr_id = id, id += 8;
      // VioManager::VioManager loads the calibration into the state (core/VioManager.cpp:69-90)
      c.q_ItoC = calib.camera_extrinsics[(size_t)i].first;
      c.p_IinC = calib.camera_extrinsics[(size_t)i].second;
      std::memcpy(c.intr, calib.camera_intrinsics[(size_t)i].d, sizeof(c.intr));
      c.model = calib.camera_intrinsics[(size_t)i];
    }
    state.dt_CAMtoIMU = calib.calib_camimu_dt;
    std::memcpy(state.dw, calib.vec_dw, sizeof(state.dw));
    std::memcpy(state.da, calib.vec_da, sizeof(state.da));
    std::memcpy(state.tg, calib.vec_tg, sizeof(state.tg));
    state.q_GYROtoIMU = calib.q_GYROtoIMU;
    state.q_ACCtoIMU = calib.q_ACCtoIMU;
    state.base_size = id;
    // initial covariance (State.cpp:134-165)
    const int N = id;
    std::vector<double> P((size_t)N * N, 0.0);
    auto diag = [&](int off, int n, double sigma) {
      for (int k = 0; k < n; k++)
        P[(size_t)(off + k) * N + off + k] = sigma * sigma;
    };
    diag(0, N, 1e-3);
    if (opt.do_calib_imu_intrinsics) {
      diag(state.dw_id, 6, 0.005);
      diag(state.da_id, 6, 0.008);
      if (opt.do_calib_imu_g_sensitivity)
        diag(state.tg_id, 9, 0.005);
      diag(state.gyro_id, 3, 0.005);
    }
    if (opt.do_calib_camera_timeoffset)
      diag(state.dt_id, 1, 0.01);
    for (auto &c : state.cams) {
      if (opt.do_calib_camera_pose) {
        diag(c.ext_id, 3, 0.005);
        diag(c.ext_id + 3, 3, 0.015);
      }
      if (opt.do_calib_camera_intrinsics) {
        diag(c.intr_id, 4, 1.0);
        diag(c.intr_id + 4, 4, 0.005);
      }
    }
    P0_ = P;
  }

  // VioManager::initialize_with_gt (core/VioManagerHelper.cpp:40-76): imustate = [t q p v bg ba]
  void initialize_with_gt(const std::array<double, 17> &imustate) {
    state.q = state.q_fej = {imustate[1], imustate[2], imustate[3], imustate[4]};
    state.p = state.p_fej = {imustate[5], imustate[6], imustate[7]};
    state.v = state.v_fej = {imustate[8], imustate[9], imustate[10]};
    state.bg = {imustate[11], imustate[12], imustate[13]};
    state.ba = {imustate[14], imustate[15], imustate[16]};
    const int N = state.base_size;
    std::vector<double> P = P0_;
    for (int k = 0; k < 15; k++) {
      const double s = k < 3 ? 0.017 : (k < 6 ? 0.05 : (k < 9 ? 0.01 : 0.02));
      for (int j = 0; j < 15; j++)
        P[(size_t)k * N + j] = P[(size_t)j * N + k] = 0.0;
      P[(size_t)k * N + k] = s * s;
    }
    cov->set(P, N);
    state.timestamp = imustate[0];
    startup_time = imustate[0];
    is_initialized_vio = true;
    database.cleanup_measurements(state.timestamp);
  }

  void feed_measurement_imu(const ImuData &message) { // core/VioManager.cpp:166-189
    double oldest_time = state.margtimestep();
    if (oldest_time > state.timestamp)
      oldest_time = -1;
    propagator.feed_imu(message, oldest_time);
  }

  // VioManager::feed_measurement_simulation (:191-254) with TrackSIM::feed_measurement_simulation (track/TrackSIM.cpp:30-79)
  void feed_measurement_simulation(double timestamp, const std::vector<int> &camids, const std::vector<std::vector<SimFeat>> &feats) {
    const auto rT1 = clock_now();
    for (size_t i = 0; i < camids.size(); i++) {
      const int cam_id = camids[i];
      for (const auto &feat : feats[i]) {
        float xn, yn;
        state.cams[(size_t)cam_id].model.undistort_f(feat.u, feat.v, xn, yn); // camera_calib.at(cam_id)->undistort_cv
        database.update_feature(feat.id, timestamp, (size_t)cam_id, feat.u, feat.v, xn, yn);
      }
    }
    const auto rT2 = clock_now();
    if (!is_initialized_vio)
      throw Error(OVB_ERR_ARG, "[SIM]: your vio system should already be initialized before simulating features");
    do_feature_propagate_update(timestamp, camids, rT1, rT2);
  }

  // ov_eval: position ATE RMSE with alignment "none" (ResultTrajectory.cpp:82-109 after AlignTrajectory "none")
  static void calculate_ate(const std::vector<TrajSample> &est, const std::vector<TrajSample> &gt, double &rmse_ori_deg, double &rmse_pos) {
    double so = 0, sp = 0;
    const size_t n = std::min(est.size(), gt.size());
    for (size_t i = 0; i < n; i++) {
      const Mat3 e_R = transpose(quat_2_Rot(est[i].q)) * quat_2_Rot(gt[i].q);
      const double ori_err = 180.0 / M_PI * norm(log_so3(e_R));
      const double pos_err = norm(gt[i].p - est[i].p);
      so += ori_err * ori_err;
      sp += pos_err * pos_err;
    }
    rmse_ori_deg = n ? std::sqrt(so / (double)n) : 0.0;
    rmse_pos = n ? std::sqrt(sp / (double)n) : 0.0;
  }

  // timing file in the reference's format (core/VioManager.cpp:117-121 header, :631-644 rows; no SLAM columns: max_slam = 0)
  void write_timing_csv(const std::string &path) const {
    FILE *f = std::fopen(path.c_str(), "w");
    if (!f)
      return;
    std::fprintf(f, "# timestamp (sec),tracking,propagation,msckf update,marginalization,total\n");
    for (const auto &t : timing)
      std::fprintf(f, "%.15f,%.5f,%.5f,%.5f,%.5f,%.5f\n", t.timestamp_inI, t.time_track, t.time_prop, t.time_msckf, t.time_marg, t.time_total);
    std::fclose(f);
  }

private:
  std::vector<double> P0_;
  bool is_initialized_vio = false;
  double startup_time = -1;
  using clk = std::chrono::steady_clock;
  static clk::time_point clock_now() { return clk::now(); }
  static double secs(clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); }

  // VioManager::do_feature_propagate_update (:323-644), MSCKF branch
  void do_feature_propagate_update(double timestamp, const std::vector<int> &sensor_ids, clk::time_point rT1, clk::time_point rT2) {
    if (state.timestamp > timestamp)
      return; // image received out of order
    if (state.timestamp != timestamp)
      propagator.propagate_and_clone(state, *cov, timestamp);
    const auto rT3 = clock_now();
    if ((int)state.clones.size() < std::min(state.opt.max_clone_size, 5))
      return;
    if (state.timestamp != timestamp)
      return;
    // ---- feature selection (:368-400, :509-524)
    std::vector<std::shared_ptr<Feature>> feats_lost, feats_marg;
    feats_lost = database.features_not_containing_newer(state.timestamp, false, true);
    if ((int)state.clones.size() > state.opt.max_clone_size || (int)state.clones.size() > 5)
      feats_marg = database.features_containing(state.margtimestep(), false, true);
    for (auto it1 = feats_lost.begin(); it1 != feats_lost.end();) { // keep features seen from a camera of this message
      bool found = false;
      for (const auto &camuvpair : (*it1)->uvs)
        if (std::find(sensor_ids.begin(), sensor_ids.end(), (int)camuvpair.first) != sensor_ids.end()) {
          found = true;
          break;
        }
      it1 = found ? it1 + 1 : feats_lost.erase(it1);
    }
    for (auto it1 = feats_lost.begin(); it1 != feats_lost.end();) // no duplicates with the marg list
      it1 = (std::find(feats_marg.begin(), feats_marg.end(), *it1) != feats_marg.end()) ? feats_lost.erase(it1) : it1 + 1;
    std::vector<std::shared_ptr<Feature>> feats_maxtracks;
    for (auto it2 = feats_marg.begin(); it2 != feats_marg.end();) {
      bool reached_max = false;
      for (const auto &cams : (*it2)->timestamps)
        if ((int)cams.second.size() > state.opt.max_clone_size) {
          reached_max = true;
          break;
        }
      if (reached_max) {
        feats_maxtracks.push_back(*it2);
        it2 = feats_marg.erase(it2);
      } else {
        ++it2;
      }
    }
    std::vector<std::shared_ptr<Feature>> featsup_MSCKF = feats_lost;
    featsup_MSCKF.insert(featsup_MSCKF.end(), feats_marg.begin(), feats_marg.end());
    featsup_MSCKF.insert(featsup_MSCKF.end(), feats_maxtracks.begin(), feats_maxtracks.end());
    // the reference sorts by track length with std::sort (unstable) on an unordered_map-ordered list: ties are
    // implementation-defined there; a stable sort keyed (length, featid) gives both backends one total order
    auto nmeas = [](const std::shared_ptr<Feature> &a) {
      size_t s = 0;
      for (const auto &pair : a->timestamps)
        s += pair.second.size();
      return s;
    };
    std::stable_sort(featsup_MSCKF.begin(), featsup_MSCKF.end(), [&](const std::shared_ptr<Feature> &a, const std::shared_ptr<Feature> &b) {
      const size_t na = nmeas(a), nb = nmeas(b);
      return na != nb ? na < nb : a->featid < b->featid;
    });
    if ((int)featsup_MSCKF.size() > state.opt.max_msckf_in_update)
      featsup_MSCKF.erase(featsup_MSCKF.begin(), featsup_MSCKF.end() - state.opt.max_msckf_in_update);
    FrameTiming ft;
    ft.feats_in = (int)featsup_MSCKF.size();
    msckf_update(featsup_MSCKF);
    ft.feats_used = last_stats.n_feats_used, ft.rows = last_stats.rows_stacked, ft.cols = last_stats.cols_stacked;
    const auto rT4 = clock_now();
    for (auto const &feat : featsup_MSCKF)
      feat->to_delete = true;
    database.cleanup();
    if ((int)state.clones.size() > state.opt.max_clone_size)
      database.cleanup_measurements(state.margtimestep());
    // StateHelper::marginalize_old_clone (StateHelper.cpp:618-629) + the id shift of marginalize (:318-326)
    if ((int)state.clones.size() > state.opt.max_clone_size) {
      const double marginal_time = state.margtimestep();
      const ClonePose marg = state.clones.at(marginal_time);
      cov->marginalize(marg.id, 6);
      state.clones.erase(marginal_time);
      for (auto &c : state.clones)
        if (c.second.id > marg.id)
          c.second.id -= 6;
    }
    const auto rT7 = clock_now();
    ft.timestamp_inI = state.timestamp + state.dt_CAMtoIMU;
    ft.time_track = secs(rT1, rT2), ft.time_prop = secs(rT2, rT3), ft.time_msckf = secs(rT3, rT4), ft.time_marg = secs(rT4, rT7), ft.time_total = secs(rT1, rT7);
    timing.push_back(ft);
    trajectory_est.push_back({state.timestamp, state.q, state.p});
    frames_done++;
  }

  // UpdaterMSCKF::update (update/UpdaterMSCKF.cpp:58-295): host steps 0-1, marshalling, device pipeline, mean update
  void msckf_update(std::vector<std::shared_ptr<Feature>> &feature_vec) {
    last_stats = ovb_stats{};
    if (feature_vec.empty())
      return;
    std::vector<double> clonetimes;
    for (const auto &c : state.clones)
      clonetimes.push_back(c.first);
    for (auto it = feature_vec.begin(); it != feature_vec.end();) { // :75-94
      (*it)->clean_old_measurements(clonetimes);
      int ct_meas = 0;
      for (const auto &pair : (*it)->timestamps)
        ct_meas += (int)pair.second.size();
      if (ct_meas < 2) {
        (*it)->to_delete = true;
        it = feature_vec.erase(it);
      } else {
        ++it;
      }
    }
    if (feature_vec.empty())
      return;
    const int C = (int)state.clones.size(), K = (int)state.cams.size(), N = cov->dim();
    std::vector<double> cR((size_t)9 * C), cp((size_t)3 * C), cRf((size_t)9 * C), cpf((size_t)3 * C), kR((size_t)9 * K), kp((size_t)3 * K), kin((size_t)8 * K);
    std::vector<int> coff((size_t)C), kmodel((size_t)K, OVB_CAM_RADTAN), kext((size_t)K), kintr((size_t)K);
    {
      int c = 0;
      for (const auto &cl : state.clones) {
        const Mat3 R = quat_2_Rot(cl.second.q), Rf = quat_2_Rot(cl.second.q_fej);
        std::copy(R.begin(), R.end(), cR.begin() + 9 * c);
        std::copy(Rf.begin(), Rf.end(), cRf.begin() + 9 * c);
        std::copy(cl.second.p.begin(), cl.second.p.end(), cp.begin() + 3 * c);
        std::copy(cl.second.p_fej.begin(), cl.second.p_fej.end(), cpf.begin() + 3 * c);
        coff[(size_t)c] = cl.second.id;
        c++;
      }
      for (int k = 0; k < K; k++) {
        const auto &cam = state.cams[(size_t)k];
        const Mat3 R = quat_2_Rot(cam.q_ItoC);
        std::copy(R.begin(), R.end(), kR.begin() + 9 * k);
        std::copy(cam.p_IinC.begin(), cam.p_IinC.end(), kp.begin() + 3 * k);
        std::copy(cam.intr, cam.intr + 8, kin.begin() + 8 * k);
        kext[(size_t)k] = state.opt.do_calib_camera_pose ? cam.ext_id : -1;
        kintr[(size_t)k] = state.opt.do_calib_camera_intrinsics ? cam.intr_id : -1;
      }
    }
    ovb_frame frame{C, K, cR.data(), cp.data(), cRf.data(), cpf.data(), coff.data(), kR.data(), kp.data(), kin.data(), kmodel.data(), kext.data(), kintr.data()};
    const int F = (int)feature_vec.size();
    std::vector<int32_t> meas_off(1, 0), keys_off(1, 0);
    std::vector<uint8_t> cam, keys;
    std::vector<uint16_t> clone;
    std::vector<float> uv, uvn;
    for (const auto &feat : feature_vec) {
      for (const auto &pair : feat->timestamps) { // the unordered_map's visit order
        keys.push_back((uint8_t)pair.first);
        const auto &fuv = feat->uvs.at(pair.first);
        const auto &fuvn = feat->uvs_norm.at(pair.first);
        for (size_t m = 0; m < pair.second.size(); m++) {
          const int ci = (int)(std::lower_bound(clonetimes.begin(), clonetimes.end(), pair.second[m]) - clonetimes.begin());
          cam.push_back((uint8_t)pair.first);
          clone.push_back((uint16_t)ci);
          uv.push_back(fuv[m][0]), uv.push_back(fuv[m][1]);
          uvn.push_back(fuvn[m][0]), uvn.push_back(fuvn[m][1]);
        }
      }
      meas_off.push_back((int32_t)cam.size());
      keys_off.push_back((int32_t)keys.size());
    }
    ovb_feat_batch batch{F, (int)cam.size(), meas_off.data(), cam.data(), clone.data(), uv.data(), uvn.data(), keys_off.data(), keys.data()};
    ovb_opts o;
    ovb_opts_default(&o);
    const auto &fi = state.opt.featinit_options;
    o.triangulate_1d = fi.triangulate_1d, o.refine_features = fi.refine_features, o.max_runs = fi.max_runs;
    o.init_lamda = fi.init_lamda, o.max_lamda = fi.max_lamda, o.min_dx = fi.min_dx, o.min_dcost = fi.min_dcost, o.lam_mult = fi.lam_mult;
    o.min_dist = fi.min_dist, o.max_dist = fi.max_dist, o.max_baseline = fi.max_baseline, o.max_cond_number = fi.max_cond_number;
    o.sigma_pix = state.opt.msckf_options.sigma_pix;
    o.chi2_multipler = state.opt.msckf_options.chi2_multipler;
    o.do_fej = state.opt.do_fej;
    o.feat_rep = state.opt.feat_rep_msckf;
    o.do_calib_camera_pose = state.opt.do_calib_camera_pose;
    o.do_calib_camera_intrinsics = state.opt.do_calib_camera_intrinsics;
    o.col_order = state.opt.col_order;
    o.compress = state.opt.compress;
    if (on_update)
      on_update(frame, batch, o, frames_done);
    std::vector<int32_t> status((size_t)F), acam((size_t)F), aclone((size_t)F);
    std::vector<double> pA((size_t)3 * F), pG((size_t)3 * F), chi2((size_t)F), dx((size_t)N, 0.0);
    ovb_feat_out out{status.data(), pA.data(), pG.data(), acam.data(), aclone.data(), chi2.data()};
    cov->msckf_update(&frame, &batch, &o, &out, dx.data(), &last_stats);
    for (int f = 0; f < F; f++) {
      Feature &feat = *feature_vec[(size_t)f];
      feat.last_status = status[(size_t)f];
      feat.last_chi2 = chi2[(size_t)f];
      feat.to_delete = true;
      status_hist[status[(size_t)f] & 15]++;
    }
    apply_dx(dx);
  }

  // the mean side of StateHelper::EKFUpdate (state/StateHelper.cpp:185-196): Type::update of every variable
  void apply_dx(const std::vector<double> &dx) {
    auto qupdate = [](Vec4 &q, const double *d) { // JPLQuat::update (types/JPLQuat.h:114-126)
      const Vec4 dq = quatnorm({.5 * d[0], .5 * d[1], .5 * d[2], 1.0});
      q = quat_multiply(dq, q);
    };
    const double *d = dx.data() + state.imu_id; // IMU::update (types/IMU.h:78-96)
    qupdate(state.q, d);
    for (int k = 0; k < 3; k++) {
      state.p[(size_t)k] += d[3 + k];
      state.v[(size_t)k] += d[6 + k];
      state.bg[(size_t)k] += d[9 + k];
      state.ba[(size_t)k] += d[12 + k];
    }
    if (state.opt.do_calib_imu_intrinsics) {
      for (int k = 0; k < 6; k++) {
        state.dw[k] += dx[(size_t)(state.dw_id + k)];
        state.da[k] += dx[(size_t)(state.da_id + k)];
      }
      if (state.opt.do_calib_imu_g_sensitivity)
        for (int k = 0; k < 9; k++)
          state.tg[k] += dx[(size_t)(state.tg_id + k)];
      qupdate(state.q_GYROtoIMU, dx.data() + state.gyro_id);
    }
    if (state.opt.do_calib_camera_timeoffset)
      state.dt_CAMtoIMU += dx[(size_t)state.dt_id];
    for (auto &c : state.cams) {
      if (state.opt.do_calib_camera_pose) { // PoseJPL::update (types/PoseJPL.h:74-91)
        qupdate(c.q_ItoC, dx.data() + c.ext_id);
        for (int k = 0; k < 3; k++)
          c.p_IinC[(size_t)k] += dx[(size_t)(c.ext_id + 3 + k)];
      }
      if (state.opt.do_calib_camera_intrinsics) {
        for (int k = 0; k < 8; k++)
          c.intr[k] += dx[(size_t)(c.intr_id + k)];
        std::memcpy(c.model.d, c.intr, sizeof(c.intr)); // StateHelper.cpp:192-196
      }
    }
    for (auto &cl : state.clones) {
      qupdate(cl.second.q, dx.data() + cl.second.id);
      for (int k = 0; k < 3; k++)
        cl.second.p[(size_t)k] += dx[(size_t)(cl.second.id + 3 + k)];
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// run_simulation main loop (ov_msckf/src/run_simulation.cpp:117-176): initialise from the simulator's ground truth, feed
// IMU at sim_freq_imu and camera frames with the reference's one-frame delay buffer. Stops after max_frames camera updates
// (0 = whole trajectory). Ground truth samples are taken at the estimate's timestamps (+dt) from the simulator's spline.
struct SimRunResult {
  std::vector<TrajSample> est, gt;
  double ate_ori_deg = 0, ate_pos = 0;
  int frames = 0;
};
inline SimRunResult run_simulation(Simulator &sim, VioManager &sys, int max_frames = 0) {
  const double next_imu_time = sim.current_timestamp() + 1.0 / sim.params.sim_freq_imu;
  std::array<double, 17> imustate;
  if (!sim.get_state(next_imu_time, imustate))
    throw Error(OVB_ERR_ARG, "[SIM]: could not initialize the filter to the first state");
  imustate[0] -= sim.params.calib_camimu_dt;
  sys.initialize_with_gt(imustate);
  double buffer_timecam = -1;
  std::vector<int> buffer_camids;
  std::vector<std::vector<SimFeat>> buffer_feats;
  SimRunResult res;
  while (sim.ok()) {
    ImuData message_imu;
    if (sim.get_next_imu(message_imu.timestamp, message_imu.wm, message_imu.am))
      sys.feed_measurement_imu(message_imu);
    double time_cam;
    std::vector<int> camids;
    std::vector<std::vector<SimFeat>> feats;
    if (sim.get_next_cam(time_cam, camids, feats)) {
      if (buffer_timecam != -1) {
        const size_t before = sys.trajectory_est.size();
        sys.feed_measurement_simulation(buffer_timecam, buffer_camids, buffer_feats);
        if (sys.trajectory_est.size() > before) {
          std::array<double, 17> gt;
          const TrajSample &e = sys.trajectory_est.back();
          if (sim.get_state(e.t + sim.params.calib_camimu_dt, gt))
            res.gt.push_back({e.t, {gt[1], gt[2], gt[3], gt[4]}, {gt[5], gt[6], gt[7]}});
          else
            res.gt.push_back(e);
        }
        if (max_frames > 0 && sys.frames_done >= max_frames)
          break;
      }
      buffer_timecam = time_cam;
      buffer_camids = camids;
      buffer_feats = feats;
    }
  }
  res.est = sys.trajectory_est;
  res.frames = sys.frames_done;
  VioManager::calculate_ate(res.est, res.gt, res.ate_ori_deg, res.ate_pos);
  return res;
}

} // namespace ovb200
#endif
