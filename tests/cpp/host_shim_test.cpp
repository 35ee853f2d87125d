// Drives one MSCKF update through the C++ host mirror (include/ovb200_host.hpp), the way reference code would call
// UpdaterMSCKF::update(state, feature_vec): reads a flat little-endian case file written by tests/test_gpu_host_shim.py,
// rebuilds State / Feature objects, runs the update, writes results + the marshalled SoA back.
// build: g++ -std=c++17 -O2 -I include tests/cpp/host_shim_test.cpp -L open_vins_b200 -lovb200 -Wl,-rpath,... -o host_shim_test
#include "ovb200_host.hpp"

#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>

using namespace ovb200;

template <class T> static std::vector<T> rd(std::ifstream &f, size_t n) {
  std::vector<T> v(n);
  f.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(n * sizeof(T)));
  if (!f)
    throw std::runtime_error("short read");
  return v;
}
template <class T> static void wr(std::ofstream &f, const std::vector<T> &v) { f.write(reinterpret_cast<const char *>(v.data()), (std::streamsize)(v.size() * sizeof(T))); }

int main(int argc, char **argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s case.bin out.bin\n", argv[0]);
    return 2;
  }
  try {
    std::ifstream in(argv[1], std::ios::binary);
    auto hdr = rd<int32_t>(in, 16);
    if (hdr[0] != 0x0b200)
      throw std::runtime_error("bad magic");
    const int C = hdr[1], K = hdr[2], N = hdr[3], F = hdr[4], M = hdr[5];
    auto dpar = rd<double>(in, 2);
    auto clone_times = rd<double>(in, C);
    auto clone_R = rd<double>(in, 9 * C), clone_p = rd<double>(in, 3 * C), clone_Rf = rd<double>(in, 9 * C), clone_pf = rd<double>(in, 3 * C);
    auto clone_off = rd<int32_t>(in, C);
    auto cam_R = rd<double>(in, 9 * K), cam_p = rd<double>(in, 3 * K), cam_intr = rd<double>(in, 8 * K);
    auto cam_model = rd<int32_t>(in, K), cam_ext = rd<int32_t>(in, K), cam_in = rd<int32_t>(in, K);
    auto P = rd<double>(in, (size_t)N * N);
    auto meas_off = rd<int32_t>(in, F + 1);
    auto mcam = rd<uint8_t>(in, M);
    auto mclone = rd<uint16_t>(in, M);
    auto uv = rd<float>(in, 2 * (size_t)M), uvn = rd<float>(in, 2 * (size_t)M);

    // ---- State
    StateOptions so;
    so.do_fej = hdr[6];
    so.feat_rep_msckf = hdr[7];
    so.do_calib_camera_pose = hdr[8];
    so.do_calib_camera_intrinsics = hdr[9];
    so.num_cameras = K;
    so.max_clone_size = C;
    ovb_config cfg{0, 256, 1024, 1024 * 48, 0};
    State state(so, cfg);
    for (int c = 0; c < C; c++) {
      auto pose = std::make_shared<PoseJPL>();
      pose->id = clone_off[c];
      std::copy(clone_R.begin() + 9 * c, clone_R.begin() + 9 * c + 9, pose->Rot);
      std::copy(clone_p.begin() + 3 * c, clone_p.begin() + 3 * c + 3, pose->pos);
      std::copy(clone_Rf.begin() + 9 * c, clone_Rf.begin() + 9 * c + 9, pose->Rot_fej);
      std::copy(clone_pf.begin() + 3 * c, clone_pf.begin() + 3 * c + 3, pose->pos_fej);
      state._clones_IMU[clone_times[c]] = pose;
    }
    for (int k = 0; k < K; k++) {
      Camera &cam = state._cameras[k];
      cam.calib_id = cam_ext[k];
      cam.intrinsics_id = cam_in[k];
      std::copy(cam_R.begin() + 9 * k, cam_R.begin() + 9 * k + 9, cam.R_ItoC);
      std::copy(cam_p.begin() + 3 * k, cam_p.begin() + 3 * k + 3, cam.p_IinC);
      std::copy(cam_intr.begin() + 8 * k, cam_intr.begin() + 8 * k + 8, cam.intrinsics);
      cam.model = cam_model[k];
    }
    StateHelper::set_initial_covariance(state, P, N);

    // ---- features, reference style (per-camera maps); every third one also carries a stale measurement at a time that
    // is not a clone time, and two extra features have fewer than two measurements: update() must clean/drop them
    std::vector<std::shared_ptr<Feature>> all, feature_vec;
    for (int f = 0; f < F; f++) {
      auto feat = std::make_shared<Feature>();
      feat->featid = (size_t)f;
      for (int i = meas_off[f]; i < meas_off[f + 1]; i++) {
        const size_t cam = mcam[i];
        feat->uvs[cam].push_back({uv[2 * i], uv[2 * i + 1]});
        feat->uvs_norm[cam].push_back({uvn[2 * i], uvn[2 * i + 1]});
        feat->timestamps[cam].push_back(clone_times[mclone[i]]);
      }
      if (f % 3 == 0) {
        const size_t cam = mcam[meas_off[f]];
        feat->uvs[cam].insert(feat->uvs[cam].begin(), {1.0f, 2.0f});
        feat->uvs_norm[cam].insert(feat->uvs_norm[cam].begin(), {0.1f, 0.2f});
        feat->timestamps[cam].insert(feat->timestamps[cam].begin(), clone_times[0] - 5.0);
      }
      all.push_back(feat);
      feature_vec.push_back(feat);
    }
    for (int e = 0; e < 2; e++) {
      auto feat = std::make_shared<Feature>();
      feat->featid = (size_t)(F + e);
      feat->uvs[0].push_back({3.0f, 4.0f});
      feat->uvs_norm[0].push_back({0.0f, 0.0f});
      feat->timestamps[0].push_back(e == 0 ? clone_times[0] : clone_times[0] - 1.0);
      all.push_back(feat);
      feature_vec.insert(feature_vec.begin() + (e == 0 ? 0 : (long)feature_vec.size() / 2), feat);
    }

    UpdaterOptions uo;
    uo.sigma_pix = dpar[0];
    uo.chi2_multipler = dpar[1];
    FeatureInitializerOptions fo;
    UpdaterMSCKF updater(uo, fo);
    updater.col_order = hdr[10];
    updater.compress = hdr[11];
    std::vector<double> dx = updater.update(state, feature_vec);
    std::vector<double> Ppost = StateHelper::get_full_covariance(state);

    // ---- results per ORIGINAL feature
    std::ofstream out(argv[2], std::ios::binary);
    std::vector<int32_t> st(F), used_flag(F, 0), del(F + 2);
    std::vector<double> pG(3 * (size_t)F), anchor_t(F);
    for (int f = 0; f < F; f++) {
      st[f] = all[f]->last_status;
      std::copy(all[f]->p_FinG, all[f]->p_FinG + 3, pG.begin() + 3 * f);
      anchor_t[f] = all[f]->anchor_clone_timestamp;
    }
    for (auto &feat : feature_vec)
      if (feat->featid < (size_t)F)
        used_flag[feat->featid] = 1;
    for (int f = 0; f < F + 2; f++)
      del[f] = all[f]->to_delete ? 1 : 0;
    std::vector<int32_t> oh = {F, N, (int32_t)feature_vec.size(), updater.last_stats.n_feats_used, (int32_t)updater.cam.size(), (int32_t)updater.keys.size()};
    wr(out, oh);
    wr(out, st);
    wr(out, used_flag);
    wr(out, del);
    wr(out, pG);
    wr(out, anchor_t);
    wr(out, dx);
    wr(out, Ppost);
    // the SoA batch update() marshalled (camera visit order of std::unordered_map included)
    wr(out, updater.meas_off);
    wr(out, updater.keys_off);
    wr(out, updater.cam);
    wr(out, updater.clone);
    wr(out, updater.uv);
    wr(out, updater.uvn);
    wr(out, updater.keys);
    // ---- StateHelper::initialize -> marginalize_old_clone: every variable behind the removed clone moves up, the SLAM
    // landmarks at the end of the covariance included (state/StateHelper.cpp:318-326)
    {
      const int N0 = state.max_covariance_size();
      auto lm = std::make_shared<Landmark>();
      lm->_featid = 777;
      Var newvar(-1, 3);
      const Var oldest(state._clones_IMU.begin()->second->id, 6), newest(state._clones_IMU.rbegin()->second->id, 6);
      std::vector<Var> H_order = {oldest, newest};
      const int r = 6;
      std::vector<double> H_R((size_t)r * 12), H_L((size_t)r * 3), res(r);
      for (int i = 0; i < r; i++) {
        for (int j = 0; j < 12; j++)
          H_R[(size_t)i * 12 + j] = 0.3 * std::sin(1.0 + 0.7 * i + 1.3 * j);
        for (int j = 0; j < 3; j++)
          H_L[(size_t)i * 3 + j] = (i % 3 == j ? 2.0 : 0.0) + 0.2 * std::cos(0.5 * i + j);
        res[i] = 0.01 * (i - 2);
      }
      std::vector<double> dx_new, dx2;
      if (!StateHelper::initialize(state, newvar, H_order, H_R, H_L, res, 1.0, 1e6, dx_new, dx2))
        throw std::runtime_error("initialize rejected the synthetic landmark");
      lm->id = newvar.first;
      state._features_SLAM[lm->_featid] = lm;
      if (lm->id != N0 || state.max_covariance_size() != N0 + 3)
        throw std::runtime_error("initialize: landmark not appended at the end of the covariance");
      const std::vector<double> blk_before = StateHelper::get_marginal_covariance(state, {Var(lm->id, 3)});
      state._options.max_clone_size = (int)state._clones_IMU.size() - 1;
      StateHelper::marginalize_old_clone(state);
      if (lm->id != N0 - 6 || state.max_covariance_size() != N0 - 3)
        throw std::runtime_error("marginalize_old_clone: landmark id not shifted");
      const std::vector<double> blk_after = StateHelper::get_marginal_covariance(state, {Var(lm->id, 3)});
      for (size_t i = 0; i < 9; i++)
        if (blk_before[i] != blk_after[i])
          throw std::runtime_error("marginalize_old_clone: landmark block moved to the wrong place");
      std::printf("host_shim_test: marginalize bookkeeping ok (landmark id %d -> %d)\n", N0, lm->id);
    }
    std::printf("host_shim_test: %d features in, %d used, dx norm2 %.6e\n", F, (int)feature_vec.size(), [&] {
      double s = 0;
      for (double v : dx)
        s += v * v;
      return s;
    }());
    return 0;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "host_shim_test: %s\n", e.what());
    return 1;
  }
}
