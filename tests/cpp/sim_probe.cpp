// sim_probe.cpp — test helper: prints values of the host-side simulator / propagator functions for tests/test_sim_cpu.py.
//   sim_probe undistort u v            -> xn yn            (cam0 of rpng_sim)
//   sim_probe distort xn yn            -> u v
//   sim_probe spline TRAJ t            -> R(9) p(3) w(3) v(3) alpha(3) a(3) start_time
//   sim_probe propfd METHOD            -> max |F_analytic - F_numeric| over the 15+24 state, and |F| for scale
//   sim_probe propmean METHOD K        -> R_GtoI before / R_GtoI p v after K steps of mean propagation with constant readings
//   sim_probe propq METHOD             -> n dt sigma_w sigma_a sigma_wb sigma_ab / diag(Q_d)[0..15) of one IMU step
//   sim_probe simproj TRAJ             -> noise-free simulator frames: state, calibration, map points and their pixels
#include "../../include/ovb200_vio.hpp"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
using namespace ovb200;

// error state of b relative to a (JPL left error for quaternions: q_b = dq(dtheta) ⊗ q_a), additive elsewhere
static Vec3 qerr(const Vec4 &qb, const Vec4 &qa) {
  const Vec4 ainv{-qa[0], -qa[1], -qa[2], qa[3]};
  const Vec4 d = quat_multiply(qb, ainv);
  return {2 * d[0], 2 * d[1], 2 * d[2]};
}

int main(int argc, char **argv) {
  if (argc < 2)
    return 2;
  const std::string cmd = argv[1];
  SimParams sp;
  rpng_sim_cameras(2, sp);
  if (cmd == "undistort" && argc >= 4) {
    float x, y;
    sp.camera_intrinsics[0].undistort_f((float)std::atof(argv[2]), (float)std::atof(argv[3]), x, y);
    std::printf("%.9g %.9g\n", x, y);
    return 0;
  }
  if (cmd == "distort" && argc >= 4) {
    float u, v;
    sp.camera_intrinsics[0].distort_f((float)std::atof(argv[2]), (float)std::atof(argv[3]), u, v);
    std::printf("%.9g %.9g\n", u, v);
    return 0;
  }
  if (cmd == "spline" && argc >= 4) {
    const std::string traj = argv[2];
    auto data = traj.substr(traj.size() - 4) == ".bin" ? load_trajectory_bin(traj) : load_simulated_trajectory(traj);
    BsplineSE3 s;
    s.feed_trajectory(data);
    for (int i = 3; i < argc; i++) {
      Mat3 R;
      Vec3 p, w, v, al, a;
      const double t = s.get_start_time() + std::atof(argv[i]);
      const bool ok = s.get_acceleration(t, R, p, w, v, al, a);
      std::printf("%d", (int)ok);
      for (double x : R) std::printf(" %.17g", x);
      for (const Vec3 *q : {&p, &w, &v, &al, &a})
        for (double x : *q) std::printf(" %.17g", x);
      std::printf(" %.17g\n", s.get_start_time());
    }
    return 0;
  }
  if (cmd == "simproj" && argc >= 3) {
    // first camera frames of a noise-free simulator: ground-truth IMU state, extrinsics, intrinsics and, per camera, the
    // map points with their simulated pixels — for a projection with OpenCV on the Python side
    const std::string traj = argv[2];
    auto data = traj.substr(traj.size() - 4) == ".bin" ? load_trajectory_bin(traj) : load_simulated_trajectory(traj);
    sp.sigma_pix = 0.0;
    sp.num_pts = 60;
    Simulator sim(sp, data);
    int frames = 0;
    bool pending = false;
    double tc_pending = 0;
    std::vector<int> camids_p;
    std::vector<std::vector<SimFeat>> feats_p;
    while (sim.ok() && frames < 3) {
      double t;
      Vec3 wm, am;
      sim.get_next_imu(t, wm, am);
      std::array<double, 17> st;
      // the true-bias history (which get_state interpolates) trails the camera time by an IMU sample: ask again after the next one
      if (pending && sim.get_state(tc_pending + sp.calib_camimu_dt, st)) {
        pending = false;
        frames++;
        for (size_t c = 0; c < camids_p.size(); c++) {
          const int cam = camids_p[c];
          std::printf("FRAME %d %zu", cam, feats_p[c].size());
          for (int k = 1; k < 8; k++) std::printf(" %.17g", st[(size_t)k]); // q_GtoI (JPL xyzw), p_IinG
          const Vec4 &qe = sp.camera_extrinsics[(size_t)cam].first;
          const Vec3 &pe = sp.camera_extrinsics[(size_t)cam].second;
          std::printf(" %.17g %.17g %.17g %.17g %.17g %.17g %.17g", qe[0], qe[1], qe[2], qe[3], pe[0], pe[1], pe[2]);
          for (int k = 0; k < 8; k++) std::printf(" %.17g", sp.camera_intrinsics[(size_t)cam].d[k]);
          std::printf("\n");
          for (const SimFeat &f : feats_p[c]) {
            const size_t id = sp.use_stereo ? f.id : f.id - (size_t)cam * sim.featmap.size();
            const Vec3 &P = sim.featmap.at(id);
            std::printf("%.17g %.17g %.17g %.9g %.9g\n", P[0], P[1], P[2], (double)f.u, (double)f.v);
          }
        }
      }
      double tc;
      std::vector<int> camids;
      std::vector<std::vector<SimFeat>> feats;
      if (!pending && sim.get_next_cam(tc, camids, feats)) {
        pending = true;
        tc_pending = tc;
        camids_p = camids;
        feats_p = feats;
      }
    }
    return 0;
  }
  if (cmd == "propq" && argc >= 3) {
    // diagonal of the discrete process noise Q_d of ONE IMU step (identity intrinsics, zero biases), and dt
    const std::string m = argv[2];
    VioOptions vo;
    vo.integration_method = m == "discrete" ? INTEGRATION_DISCRETE : (m == "analytical" ? INTEGRATION_ANALYTICAL : INTEGRATION_RK4);
    struct NullCov3 : CovBackend {
      int dim() override { return 0; }
      void set(const std::vector<double> &, int) override {}
      std::vector<double> get() override { return {}; }
      std::vector<double> get_marginal(const std::vector<int> &, const std::vector<int> &) override { return {}; }
      void clone(int, int, const double *, int) override {}
      void marginalize(int, int) override {}
      void propagate(int, int, const std::vector<int> &, const std::vector<int> &, const std::vector<double> &, const std::vector<double> &) override {}
      int msckf_update(const ovb_frame *, const ovb_feat_batch *, const ovb_opts *, ovb_feat_out *, double *, ovb_stats *) override { return 0; }
    };
    VioManager sys(vo, sp, std::make_shared<NullCov3>());
    VioState st = sys.state;
    st.q = st.q_fej = quatnorm({0.1, -0.2, 0.3, 0.9});
    st.p = st.p_fej = {1, 2, 3};
    st.v = st.v_fej = {0.5, -0.3, 0.2};
    ImuData d0, d1;
    d0.timestamp = 0, d1.timestamp = 0.0025;
    d0.wm = d1.wm = {0.3, -0.2, 0.5};
    d0.am = d1.am = {0.5, 9.6, 1.0};
    Propagator prop(9.81);
    std::vector<double> F, Qd;
    prop.predict_and_compute(st, d0, d1, F, Qd);
    const int n = (int)std::lround(std::sqrt((double)Qd.size()));
    std::printf("%d %.17g %.17g %.17g %.17g %.17g\n", n, d1.timestamp - d0.timestamp, vo.sigma_w, vo.sigma_a, vo.sigma_wb, vo.sigma_ab);
    for (int i = 0; i < 15; i++) std::printf("%.17g ", Qd[(size_t)i * n + i]);
    std::printf("\n");
    return 0;
  }
  if (cmd == "propmean" && argc >= 4) {
    // K steps of the mean propagation (predict_mean_*, state/Propagator.cpp:482-681) with constant raw IMU readings, identity
    // intrinsics, zero biases: prints R_GtoI (row-major), p_IinG, v_IinG for comparison with an ODE solver
    const std::string m = argv[2];
    const int K = std::atoi(argv[3]);
    VioOptions vo;
    vo.integration_method = m == "discrete" ? INTEGRATION_DISCRETE : (m == "analytical" ? INTEGRATION_ANALYTICAL : INTEGRATION_RK4);
    struct NullCov2 : CovBackend {
      int dim() override { return 0; }
      void set(const std::vector<double> &, int) override {}
      std::vector<double> get() override { return {}; }
      std::vector<double> get_marginal(const std::vector<int> &, const std::vector<int> &) override { return {}; }
      void clone(int, int, const double *, int) override {}
      void marginalize(int, int) override {}
      void propagate(int, int, const std::vector<int> &, const std::vector<int> &, const std::vector<double> &, const std::vector<double> &) override {}
      int msckf_update(const ovb_frame *, const ovb_feat_batch *, const ovb_opts *, ovb_feat_out *, double *, ovb_stats *) override { return 0; }
    };
    VioManager sys(vo, sp, std::make_shared<NullCov2>());
    VioState st = sys.state;
    st.q = st.q_fej = quatnorm({0.1, -0.2, 0.3, 0.9});
    st.p = st.p_fej = {1, 2, 3};
    st.v = st.v_fej = {0.5, -0.3, 0.2};
    st.bg = {0, 0, 0};
    st.ba = {0, 0, 0};
    ImuData d0, d1;
    d0.wm = d1.wm = {0.3, -0.2, 0.5};
    d0.am = d1.am = {0.5, 9.6, 1.0};
    Propagator prop(9.81);
    const double dt = 0.0025;
    const Mat3 R0 = quat_2_Rot(st.q);
    std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", R0[0], R0[1], R0[2], R0[3], R0[4], R0[5], R0[6], R0[7], R0[8]);
    for (int k = 0; k < K; k++) {
      d0.timestamp = k * dt, d1.timestamp = (k + 1) * dt;
      std::vector<double> F, Qd;
      prop.predict_and_compute(st, d0, d1, F, Qd);
    }
    const Mat3 R = quat_2_Rot(st.q);
    std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", R[0], R[1], R[2], R[3], R[4], R[5], R[6], R[7],
                R[8], st.p[0], st.p[1], st.p[2], st.v[0], st.v[1], st.v[2]);
    return 0;
  }
  if (cmd == "propfd" && argc >= 3) {
    const std::string m = argv[2];
    VioOptions vo;
    vo.do_fej = false; // numeric Jacobian around the current estimate
    vo.integration_method = m == "discrete" ? INTEGRATION_DISCRETE : (m == "analytical" ? INTEGRATION_ANALYTICAL : INTEGRATION_RK4);
    struct NullCov : CovBackend {
      int dim() override { return 0; }
      void set(const std::vector<double> &, int) override {}
      std::vector<double> get() override { return {}; }
      std::vector<double> get_marginal(const std::vector<int> &, const std::vector<int> &) override { return {}; }
      void clone(int, int, const double *, int) override {}
      void marginalize(int, int) override {}
      void propagate(int, int, const std::vector<int> &, const std::vector<int> &, const std::vector<double> &, const std::vector<double> &) override {}
      int msckf_update(const ovb_frame *, const ovb_feat_batch *, const ovb_opts *, ovb_feat_out *, double *, ovb_stats *) override { return 0; }
    };
    VioManager sys(vo, sp, std::make_shared<NullCov>());
    VioState s0 = sys.state;
    s0.q = s0.q_fej = quatnorm({0.1, -0.2, 0.3, 0.9});
    s0.p = s0.p_fej = {1, 2, 3};
    s0.v = s0.v_fej = {0.5, -0.3, 0.2};
    s0.bg = {0.01, -0.02, 0.005};
    s0.ba = {0.05, 0.02, -0.03};
    const double dwv[6] = {1.01, 0.002, -0.003, 0.99, 0.004, 1.02}, dav[6] = {0.98, -0.001, 0.002, 1.01, 0.003, 0.99};
    const double tgv[9] = {1e-3, -2e-3, 5e-4, 3e-4, 1e-3, -1e-3, 2e-4, -4e-4, 6e-4};
    std::copy(dwv, dwv + 6, s0.dw);
    std::copy(dav, dav + 6, s0.da);
    std::copy(tgv, tgv + 9, s0.tg);
    s0.q_GYROtoIMU = quatnorm({0.01, -0.02, 0.015, 1.0});
    ImuData d0, d1;
    d0.timestamp = 0, d1.timestamp = 0.0025;
    d0.wm = {0.3, -0.2, 0.5}, d1.wm = {0.31, -0.19, 0.52};
    d0.am = {0.5, 9.6, 1.0}, d1.am = {0.55, 9.62, 0.98};
    Propagator prop(9.81);
    const int n = 39;
    std::vector<double> F, Qd;
    VioState nominal = s0;
    prop.predict_and_compute(nominal, d0, d1, F, Qd);
    double maxdiff = 0, maxF = 0;
    const double eps = 1e-6;
    for (int j = 0; j < n; j++) {
      VioState sp2 = s0;
      auto rot = [&](Vec4 &q) { q = quat_multiply(quatnorm({0.5 * eps * (j % 3 == 0), 0.5 * eps * (j % 3 == 1), 0.5 * eps * (j % 3 == 2), 1.0}), q); };
      if (j < 3) rot(sp2.q);
      else if (j < 6) sp2.p[(size_t)(j - 3)] += eps;
      else if (j < 9) sp2.v[(size_t)(j - 6)] += eps;
      else if (j < 12) sp2.bg[(size_t)(j - 9)] += eps;
      else if (j < 15) sp2.ba[(size_t)(j - 12)] += eps;
      else if (j < 21) sp2.dw[j - 15] += eps;
      else if (j < 27) sp2.da[j - 21] += eps;
      else if (j < 36) sp2.tg[j - 27] += eps;
      else rot(sp2.q_GYROtoIMU);
      sp2.q_fej = sp2.q, sp2.p_fej = sp2.p, sp2.v_fej = sp2.v;
      std::vector<double> F2, Q2;
      prop.predict_and_compute(sp2, d0, d1, F2, Q2);
      double col[15];
      const Vec3 dth = qerr(sp2.q, nominal.q);
      for (int k = 0; k < 3; k++) {
        col[k] = dth[(size_t)k] / eps;
        col[3 + k] = (sp2.p[(size_t)k] - nominal.p[(size_t)k]) / eps;
        col[6 + k] = (sp2.v[(size_t)k] - nominal.v[(size_t)k]) / eps;
        col[9 + k] = (sp2.bg[(size_t)k] - nominal.bg[(size_t)k]) / eps;
        col[12 + k] = (sp2.ba[(size_t)k] - nominal.ba[(size_t)k]) / eps;
      }
      for (int i = 0; i < 15; i++) {
        maxdiff = std::max(maxdiff, std::abs(col[i] - F[(size_t)i * n + j]));
        maxF = std::max(maxF, std::abs(F[(size_t)i * n + j]));
      }
    }
    double asym = 0, mind = 1e300;
    for (int i = 0; i < n; i++) {
      for (int j = 0; j < n; j++)
        asym = std::max(asym, std::abs(Qd[(size_t)i * n + j] - Qd[(size_t)j * n + i]));
      if (i < 15)
        mind = std::min(mind, Qd[(size_t)i * n + i]);
    }
    std::printf("%.6g %.6g %.6g %.6g\n", maxdiff, maxF, asym, mind);
    return 0;
  }
  return 2;
}
