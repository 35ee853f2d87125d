"""A closed-loop sliding-window filter around the update path (test infrastructure, not product code).

What is under test is the update path in its real role: frame after frame the SAME state (mean and covariance) is
propagated, cloned, updated with the MSCKF tracks that end in that frame, and marginalised — the loop of
VioManager::do_feature_propagate_update (ov_msckf/src/core/VioManager.cpp:341-596). The IMU propagator of the reference
(`Propagator`) is NOT restated: the motion model here is a relative-pose ("odometry") increment with
white noise, which is enough to close the loop: if a Jacobian sign, an FEJ rule, a column map or the dx convention were
wrong, the estimate would drift away from the truth within a few frames instead of staying inside its 3-sigma band.

Two interchangeable backends run the identical host logic:
  OracleBackend  CPU restatement (oracle/) — the covariance is a numpy array
  EngineBackend  libovb200.so through capi.Engine — the covariance lives on the GPU for the whole run
State layout: [current IMU pose 6][per camera: extrinsics 6, intrinsics 8 when calibrated][clone poses 6 each, oldest first].
Error-state convention (types/JPLQuat.h, PoseJPL.h): R_true = exp(-[dtheta]x) R_est, p_true = p_est + dp; the correction
dx of an update is applied as R <- exp(-[dx_theta]x) R, p <- p + dx_p (Type::update, StateHelper.cpp:185-188).
"""
from __future__ import annotations

import numpy as np

from open_vins_b200 import capi, sim
from open_vins_b200.capi import FeatArrays, FrameArrays


class OracleBackend:
    def __init__(self, oracle, P):
        self.o = oracle
        self.P = np.array(P, dtype=np.float64)

    def dim(self):
        return self.P.shape[0]

    def propagate(self, Phi, Q):
        st, self.P = self.o.cov_propagate(self.P, 0, Phi, Q, [0], [6])
        assert st == 0

    def clone(self):
        self.P = self.o.cov_clone(self.P, 0, 6, None, -1)

    def marginalize(self, off):
        self.P = self.o.cov_marginalize(self.P, off, 6)

    def update(self, frame, feats, opts):
        r = self.o.msckf_update(frame, feats, opts, self.P, dumps=False)
        assert r["status"] == 0
        self.P = r["P"]
        return r["out"].status.copy(), r["dx"], int(r["stats"].n_feats_used)

    def cov(self):
        return self.P.copy()


class EngineBackend:
    def __init__(self, P, max_state=256):
        self.e = capi.Engine(max_state=max_state, max_feats=512, max_meas=512 * 48)
        self.e.cov_set(np.array(P, dtype=np.float64))

    def dim(self):
        return self.e.cov_dim()

    def propagate(self, Phi, Q):
        assert self.e.cov_propagate(0, Phi, Q, [0], [6]) == 0

    def clone(self):
        self.e.cov_clone(0, 6)

    def marginalize(self, off):
        self.e.cov_marginalize(off, 6)

    def update(self, frame, feats, opts):
        st, out, dx, stats = self.e.msckf_update(frame, feats, opts)
        assert st == 0
        return out.status.copy(), dx, int(stats.n_feats_used)

    def cov(self):
        return self.e.cov_get()

    def close(self):
        self.e.close()


def _apply_pose(R, p, d):
    return sim._orthonormalize(sim.exp_so3(-d[0:3]) @ R), p + d[3:6]


def run(backend_factory, n_frames=40, window=8, n_cams=2, feats_per_frame=30, calib=True, seed=0, dt=0.1, track_len=6):
    """Runs the loop; returns dict(p_est [n_frames][3], p_true, sigma_p [n_frames] (1-sigma of the current position from P),
    used (features accepted per frame), P_final)."""
    rng = np.random.default_rng(seed)
    K = n_cams
    # ---- truth: trajectory + rig
    camR_true = np.array([np.array(sim._T_IMU_CAM[k])[:, :3].T for k in range(K)])
    camp_true = np.array([-camR_true[k] @ np.array(sim._T_IMU_CAM[k])[:, 3] for k in range(K)])
    intr_true = np.array([sim._INTR[k] for k in range(K)], dtype=np.float64)
    t0 = 3.0
    Rt, pt = sim._trajectory(t0)
    # ---- initial estimate and covariance
    sig_th, sig_p = 1e-3, 5e-3
    ext_off = [6 + 14 * k if calib else -1 for k in range(K)]
    intr_off = [6 + 14 * k + 6 if calib else -1 for k in range(K)]
    N0 = 6 + (14 * K if calib else 0)
    s0 = np.zeros(N0)
    s0[0:3], s0[3:6] = sig_th, sig_p
    if calib:
        for k in range(K):
            s0[ext_off[k]:ext_off[k] + 3], s0[ext_off[k] + 3:ext_off[k] + 6] = 1e-3, 2e-3
            s0[intr_off[k]:intr_off[k] + 4], s0[intr_off[k] + 4:intr_off[k] + 8] = 0.3, 5e-4
    P0 = np.diag(s0 ** 2)
    e0 = s0 * rng.standard_normal(N0)
    R_est, p_est = sim._orthonormalize(sim.exp_so3(e0[0:3]) @ Rt), pt - e0[3:6]
    camR_est, camp_est, intr_est = camR_true.copy(), camp_true.copy(), intr_true.copy()
    if calib:
        for k in range(K):
            camR_est[k] = sim._orthonormalize(sim.exp_so3(e0[ext_off[k]:ext_off[k] + 3]) @ camR_true[k])
            camp_est[k] = camp_true[k] - e0[ext_off[k] + 3:ext_off[k] + 6]
            intr_est[k] = intr_true[k] - e0[intr_off[k]:intr_off[k] + 8]
    be = backend_factory(P0)
    opts = capi.default_opts(do_calib_camera_pose=int(calib), do_calib_camera_intrinsics=int(calib), col_order=capi.COLS_CANONICAL)
    clones = []      # dicts: R, p, R_fej, p_fej, t, R_true, p_true
    tracks = []      # live feature tracks: dict(pf, obs=[(frame_idx, cam, uv(2,f32))...], last)
    out = dict(p_est=[], p_true=[], sigma_p=[], used=[])
    sig_odo_th, sig_odo_p = 2e-4, 2e-3
    for fidx in range(n_frames):
        t = t0 + dt * (fidx + 1)
        Rt_new, pt_new = sim._trajectory(t)
        # ---- 1. propagate mean + covariance with a noisy relative-pose increment (stand-in for Propagator::propagate_and_clone)
        dR_true = Rt_new @ Rt.T                      # R_new = dR R_old
        dp_true = Rt @ (pt_new - pt)                 # increment expressed in the old IMU frame
        dR_meas = sim.exp_so3(sig_odo_th * rng.standard_normal(3)) @ dR_true
        dp_meas = dp_true + sig_odo_p * rng.standard_normal(3)
        R_old, p_old = R_est, p_est
        R_est = sim._orthonormalize(dR_meas @ R_old)
        p_est = p_old + R_old.T @ dp_meas
        # error propagation: dtheta' = dR dtheta (+ noise), dp' = dp - R_old' [dp_meas]x dtheta (- R_old' noise)
        Phi = np.zeros((6, 6))
        Phi[0:3, 0:3] = dR_meas
        Phi[3:6, 3:6] = np.eye(3)
        Phi[3:6, 0:3] = -R_old.T @ sim.skew(dp_meas)
        Q = np.zeros((6, 6))
        Q[0:3, 0:3] = sig_odo_th ** 2 * np.eye(3)
        Q[3:6, 3:6] = sig_odo_p ** 2 * np.eye(3)
        be.propagate(Phi, Q)
        Rt, pt = Rt_new, pt_new
        # ---- 2. clone the current pose (StateHelper::augment_clone); its FEJ is its value at cloning time
        be.clone()
        clones.append(dict(R=R_est.copy(), p=p_est.copy(), R_fej=R_est.copy(), p_fej=p_est.copy(), R_true=Rt.copy(), p_true=pt.copy()))
        n_cl = len(clones)
        clone_off = [N0 + 6 * c for c in range(n_cl)]
        # ---- 3. observe: new tracks start here, live tracks get a measurement per camera
        for _ in range(feats_per_frame):
            z = rng.uniform(5.0, 7.0)
            un, vn = rng.uniform(-0.45, 0.45), rng.uniform(-0.30, 0.30)
            R_GtoC = camR_true[0] @ Rt
            pf = R_GtoC.T @ (np.array([un * z, vn * z, z]) - camp_true[0]) + pt
            tracks.append(dict(pf=pf, obs=[], born=fidx, life=int(rng.integers(3, track_len + 1))))
        for tr in tracks:
            for cam in range(K):
                pc = camR_true[cam] @ (Rt @ (tr["pf"] - pt)) + camp_true[cam]
                if pc[2] < 0.2:
                    continue
                u, v = sim.distort(0, intr_true[cam], pc[0] / pc[2], pc[1] / pc[2])
                if not (0 <= u < 752 and 0 <= v < 480):
                    continue
                uv = np.array([u + rng.standard_normal(), v + rng.standard_normal()], dtype=np.float32)
                tr["obs"].append((fidx, cam, uv))
        # ---- 4. tracks that end now (or whose oldest clone is about to leave the window) feed the MSCKF update
        first_frame = fidx - n_cl + 1
        ending = [tr for tr in tracks if fidx - tr["born"] + 1 >= tr["life"]]
        tracks = [tr for tr in tracks if fidx - tr["born"] + 1 < tr["life"]]
        meas_off, cam_l, clone_l, uv_l, uvn_l = [0], [], [], [], []
        for tr in ending:
            cnt = 0
            for cam in range(K)[::-1]:  # camera visit order of the GCC-built reference
                for (fi, c, uv) in tr["obs"]:
                    if c != cam or fi < first_frame:
                        continue
                    xn, yn = sim.undistort(0, intr_est[cam], float(uv[0]), float(uv[1]))
                    cam_l.append(cam)
                    clone_l.append(fi - first_frame)
                    uv_l.append(uv)
                    uvn_l.append((np.float32(xn), np.float32(yn)))
                    cnt += 1
            meas_off.append(meas_off[-1] + cnt)
        n_used = 0
        if len(ending) and meas_off[-1] > 0:
            frame = FrameArrays(np.array([c["R"] for c in clones]), np.array([c["p"] for c in clones]),
                                np.array([c["R_fej"] for c in clones]), np.array([c["p_fej"] for c in clones]), clone_off,
                                camR_est, camp_est, intr_est, [0] * K, ext_off, intr_off)
            feats = FeatArrays(meas_off, cam_l, clone_l, np.array(uv_l, dtype=np.float32).reshape(-1, 2),
                               np.array(uvn_l, dtype=np.float32).reshape(-1, 2))
            status, dx, n_used = be.update(frame, feats, opts)
            # ---- apply the correction to every variable (StateHelper.cpp:185-188)
            R_est, p_est = _apply_pose(R_est, p_est, dx[0:6])
            if calib:
                for k in range(K):
                    camR_est[k], camp_est[k] = _apply_pose(camR_est[k], camp_est[k], dx[ext_off[k]:ext_off[k] + 6])
                    intr_est[k] = intr_est[k] + dx[intr_off[k]:intr_off[k] + 8]
            for c, off in zip(clones, clone_off):
                c["R"], c["p"] = _apply_pose(c["R"], c["p"], dx[off:off + 6])
        # ---- 5. marginalise the oldest clone once the window is full (StateHelper::marginalize_old_clone)
        if len(clones) > window:
            be.marginalize(N0)
            clones.pop(0)
        Pc = be.cov()
        out["p_est"].append(p_est.copy())
        out["p_true"].append(pt.copy())
        out["sigma_p"].append(float(np.sqrt(np.trace(Pc[3:6, 3:6]))))
        out["used"].append(n_used)
    out["P_final"] = be.cov()
    out["p_est"], out["p_true"] = np.array(out["p_est"]), np.array(out["p_true"])
    out["intr_err"] = float(np.abs(intr_est[:, :4] - intr_true[:, :4]).max()) if calib else 0.0
    if hasattr(be, "close"):
        be.close()
    return out
