// run_simulation.cpp — rpng_sim runner (ov_msckf/src/run_simulation.cpp) on the host layer of include/ovb200_vio.hpp.
//   ovb_run_simulation   (this file, -DOVB_SIM_ENGINE): covariance and MSCKF updates on the CUDA engine (libovb200.so)
//   tests/cpp/run_simulation_oracle (same file, -DOVB_SIM_ORACLE, test infrastructure): the CPU oracle behind the same interface
// Usage: <exe> --traj FILE(.txt|.bin) [--cams K] [--clones C] [--msckf M] [--pts P] [--frames F] [--calib 0|1]
//              [--est OUT.txt] [--timing OUT.csv] [--capture FRAME PREFIX] [--integration discrete|rk4|analytical]
// Prints one JSON line: frames, ATE (alignment none), mean per-stage host times.
#ifdef OVB_SIM_ORACLE
#include "oracle_backend.hpp"
#else
#include "../include/ovb200_vio.hpp"
#endif
#include <cstdio>
#include <cstring>
#include <string>

using namespace ovb200;

static void write_blob(FILE *f, const void *p, size_t bytes) { std::fwrite(p, 1, bytes, f); }

int main(int argc, char **argv) {
  std::string traj, est_path, timing_path, capture_prefix, integration = "rk4", compress = "cholqr2";
  int cams = 2, clones = 11, msckf = 10, pts = 250, frames = 0, calib = 1, capture_frame = -1;
  for (int i = 1; i < argc; i++) {
    auto next = [&]() { return std::string(i + 1 < argc ? argv[++i] : ""); };
    const std::string a = argv[i];
    if (a == "--traj") traj = next();
    else if (a == "--cams") cams = std::stoi(next());
    else if (a == "--clones") clones = std::stoi(next());
    else if (a == "--msckf") msckf = std::stoi(next());
    else if (a == "--pts") pts = std::stoi(next());
    else if (a == "--frames") frames = std::stoi(next());
    else if (a == "--calib") calib = std::stoi(next());
    else if (a == "--est") est_path = next();
    else if (a == "--timing") timing_path = next();
    else if (a == "--integration") integration = next();
    else if (a == "--compress") compress = next();
    else if (a == "--capture") { capture_frame = std::stoi(next()); capture_prefix = next(); }
  }
  std::vector<std::array<double, 8>> traj_data =
      traj.size() > 4 && traj.substr(traj.size() - 4) == ".bin" ? load_trajectory_bin(traj) : load_simulated_trajectory(traj);
  if (traj_data.size() < 4) {
    std::fprintf(stderr, "could not load the trajectory '%s'\n", traj.c_str());
    return 2;
  }
  SimParams sp;
  rpng_sim_cameras(cams, sp);
  sp.use_stereo = cams > 1;
  sp.num_pts = pts;
  VioOptions vo;
  vo.num_cameras = cams;
  vo.max_clone_size = clones;
  vo.max_msckf_in_update = msckf;
  vo.do_calib_camera_pose = vo.do_calib_camera_intrinsics = vo.do_calib_camera_timeoffset = vo.do_calib_imu_intrinsics = vo.do_calib_imu_g_sensitivity = calib != 0;
  vo.compress = compress == "tsqr" ? OVB_COMPRESS_HOUSEHOLDER_TSQR : (compress == "gram" ? OVB_COMPRESS_NORMAL_EQUATIONS : OVB_COMPRESS_CHOLQR2);
  vo.integration_method = integration == "discrete" ? INTEGRATION_DISCRETE : (integration == "analytical" ? INTEGRATION_ANALYTICAL : INTEGRATION_RK4);
  try {
    Simulator sim(sp, traj_data);
#ifdef OVB_SIM_ORACLE
    auto backend = std::make_shared<OracleCov>();
    const char *backend_name = "oracle";
#else
    ovb_config cfg{0, 640, std::max(1024, msckf), std::max(1024, msckf) * 2 * (clones + 1) * cams / 2 + 1024, 0};
    auto backend = std::make_shared<EngineCov>(cfg);
    const char *backend_name = "engine";
#endif
    VioManager sys(vo, sp, backend);
    if (capture_frame >= 0) {
      // dump the marshalled inputs of ONE update (the golden "update case" wire format of tests/golden_io.py:
      // little-endian, a text header line with the array shapes followed by the raw arrays) and the prior covariance
      sys.on_update = [&](const ovb_frame &fr, const ovb_feat_batch &fb, const ovb_opts &op, int frame_index) {
        if (frame_index != capture_frame)
          return;
        FILE *f = std::fopen((capture_prefix + ".case").c_str(), "wb");
        if (!f)
          return;
        const int N = backend->dim();
        const std::vector<double> P = backend->get();
        int nkeys = fb.cam_keys_off ? fb.cam_keys_off[fb.n_feats] : 0;
        std::fprintf(f, "OVBCASE1 n_clones=%d n_cams=%d n_feats=%d n_meas=%d n_keys=%d N=%d opts=%zu\n", fr.n_clones, fr.n_cams, fb.n_feats, fb.n_meas, nkeys, N,
                     sizeof(ovb_opts));
        write_blob(f, fr.clone_R, sizeof(double) * 9 * fr.n_clones);
        write_blob(f, fr.clone_p, sizeof(double) * 3 * fr.n_clones);
        write_blob(f, fr.clone_R_fej, sizeof(double) * 9 * fr.n_clones);
        write_blob(f, fr.clone_p_fej, sizeof(double) * 3 * fr.n_clones);
        write_blob(f, fr.clone_off, sizeof(int) * fr.n_clones);
        write_blob(f, fr.cam_R, sizeof(double) * 9 * fr.n_cams);
        write_blob(f, fr.cam_p, sizeof(double) * 3 * fr.n_cams);
        write_blob(f, fr.cam_intr, sizeof(double) * 8 * fr.n_cams);
        write_blob(f, fr.cam_model, sizeof(int) * fr.n_cams);
        write_blob(f, fr.cam_ext_off, sizeof(int) * fr.n_cams);
        write_blob(f, fr.cam_intr_off, sizeof(int) * fr.n_cams);
        write_blob(f, fb.meas_off, sizeof(int32_t) * (fb.n_feats + 1));
        write_blob(f, fb.cam, fb.n_meas);
        write_blob(f, fb.clone, sizeof(uint16_t) * fb.n_meas);
        write_blob(f, fb.uv, sizeof(float) * 2 * fb.n_meas);
        write_blob(f, fb.uvn, sizeof(float) * 2 * fb.n_meas);
        write_blob(f, fb.cam_keys_off, sizeof(int32_t) * (fb.n_feats + 1));
        write_blob(f, fb.cam_keys, (size_t)nkeys);
        write_blob(f, &op, sizeof(ovb_opts));
        write_blob(f, P.data(), sizeof(double) * P.size());
        std::fclose(f);
      };
    }
    SimRunResult res = run_simulation(sim, sys, frames);
    if (!est_path.empty()) {
      FILE *f = std::fopen(est_path.c_str(), "w");
      if (f) {
        std::fprintf(f, "# timestamp(s) tx ty tz qx qy qz qw | gt: tx ty tz qx qy qz qw\n");
        for (size_t i = 0; i < res.est.size(); i++) {
          const auto &e = res.est[i];
          const auto &g = res.gt[i];
          std::fprintf(f, "%.9f %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", e.t, e.p[0], e.p[1], e.p[2], e.q[0],
                       e.q[1], e.q[2], e.q[3], g.p[0], g.p[1], g.p[2], g.q[0], g.q[1], g.q[2], g.q[3]);
        }
        std::fclose(f);
      }
    }
    if (!timing_path.empty())
      sys.write_timing_csv(timing_path);
    double t_prop = 0, t_msckf = 0, t_total = 0, feats = 0, used = 0, rows = 0;
    for (const auto &t : sys.timing) {
      t_prop += t.time_prop, t_msckf += t.time_msckf, t_total += t.time_total;
      feats += t.feats_in, used += t.feats_used, rows += t.rows;
    }
    const double n = sys.timing.empty() ? 1.0 : (double)sys.timing.size();
    std::printf("{\"backend\": \"%s\", \"frames\": %d, \"cams\": %d, \"max_clones\": %d, \"max_msckf_in_update\": %d, \"num_pts\": %d, \"calib\": %d, "
                "\"state_dim\": %d, \"ate_pos_m\": %.12g, \"ate_ori_deg\": %.12g, \"mean_feats_in\": %.2f, \"mean_feats_used\": %.2f, \"mean_rows\": %.1f, "
                "\"mean_ms_propagation\": %.4f, \"mean_ms_msckf_update\": %.4f, \"mean_ms_total\": %.4f, \"map_points\": %zu, \"status_hist\": [%ld, %ld, %ld, %ld, %ld, %ld, %ld, %ld, %ld]}\n",
                backend_name, res.frames, cams, clones, msckf, pts, calib, backend->dim(), res.ate_pos, res.ate_ori_deg, feats / n, used / n, rows / n,
                1e3 * t_prop / n, 1e3 * t_msckf / n, 1e3 * t_total / n, sim.featmap.size(), sys.status_hist[0], sys.status_hist[1], sys.status_hist[2],
                sys.status_hist[3], sys.status_hist[4], sys.status_hist[5], sys.status_hist[6], sys.status_hist[7], sys.status_hist[8]);
  } catch (const std::exception &e) {
    std::fprintf(stderr, "run_simulation failed: %s\n", e.what());
    return 1;
  }
  return 0;
}
