"""CPU tests of the host-side rpng_sim pipeline (include/ovb200_sim.hpp, ovb200_vio.hpp): simulator, B-spline, Propagator,
runner + ATE, with the CPU oracle as the update backend (tests/cpp/run_simulation_oracle).
Reference: ov_msckf/src/sim/Simulator.cpp, ov_core/src/sim/BsplineSE3.cpp, ov_msckf/src/state/Propagator.cpp,
ov_msckf/src/test_sim_repeat.cpp:134-160 (determinism), ov_eval/src/calc/ResultTrajectory.cpp:82-109 (ATE)."""
import os
import subprocess

import numpy as np
import pytest

from open_vins_b200 import build as b
from open_vins_b200 import simrun

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRAJ = simrun.TRAJ_FIXTURE


@pytest.fixture(scope="module")
def exes(oracle):
    eng, orc = b.build_sim_tools(), oracle.build_sim_runner()
    probe = os.path.join(ROOT, "tests", "cpp", "sim_probe")
    src = os.path.join(ROOT, "tests", "cpp", "sim_probe.cpp")
    if not os.path.exists(probe) or os.path.getmtime(probe) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "include", "ovb200_vio.hpp"))):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), src, "-L", os.path.join(ROOT, "open_vins_b200"),
                               "-lovb200", "-Wl,-rpath,$ORIGIN/../../open_vins_b200", "-o", probe])
    return {"engine": eng, "oracle": orc, "probe": probe}


def _probe(exes, *args):
    out = subprocess.run([exes["probe"]] + [str(a) for a in args], check=True, capture_output=True, text=True).stdout
    return [np.array([float(x) for x in line.split()]) for line in out.strip().splitlines()]


def test_undistort_matches_opencv(exes):
    """CamRadtan::undistort_f = cv::undistortPoints (cam/CamRadtan.h:95-114): the restated 5-step fixed-point iteration
    against the OpenCV of this image, bit for bit in float32, over the image."""
    cv2 = pytest.importorskip("cv2")
    K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1.0]])
    D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])
    rng = np.random.default_rng(0)
    for u, v in np.column_stack([rng.uniform(0, 752, 40), rng.uniform(0, 480, 40)]).astype(np.float32):
        ref = cv2.undistortPoints(np.array([[[u, v]]], dtype=np.float32), K, D).ravel()
        got = _probe(exes, "undistort", repr(float(u)), repr(float(v)))[0].astype(np.float32)
        assert np.array_equal(got, ref.astype(np.float32)), (u, v, got, ref)
        # distort(undistort(uv)) returns to the pixel within the 5-iteration accuracy
        back = _probe(exes, "distort", repr(float(got[0])), repr(float(got[1])))[0]
        assert np.abs(back - [u, v]).max() < 0.5  # 5 iterations leave a fraction of a pixel in the image corners (OpenCV behaves the same)


def test_bspline_derivatives(exes):
    """BsplineSE3::get_velocity / get_acceleration (sim/BsplineSE3.cpp:122-233) against central differences of get_pose."""
    h = 2.0**-10  # exactly representable next to timestamps of ~1.5e9 s (ulp 2.4e-7)
    ts = [5.015625, 12.265625, 40.015625]  # off the spline knots (multiples of 0.05 s), exact on the 2^-22 s grid
    for t in ts:
        lo, mid, hi = _probe(exes, "spline", TRAJ, t - h, t, t + h)
        assert lo[0] == mid[0] == hi[0] == 1
        p = lambda r: r[10:13]
        v, a = mid[16:19], mid[22:25]
        assert np.allclose((p(hi) - p(lo)) / (2 * h), v, atol=1e-4)
        assert np.allclose((p(hi) - 2 * p(mid) + p(lo)) / h**2, a, atol=1e-4)
        R = lambda r: r[1:10].reshape(3, 3)  # R_GtoI
        W = R(mid) @ ((R(hi).T - R(lo).T) / (2 * h))  # R_ItoG' d/dt R_ItoG = skew(w_IinI)
        w = np.array([W[2, 1], W[0, 2], W[1, 0]])
        assert np.allclose(w, mid[13:16], atol=2e-4)


@pytest.mark.parametrize("method", ["discrete", "rk4", "analytical"])
def test_propagator_F_matches_finite_differences(exes, method):
    """compute_F_and_G_analytic / _discrete (state/Propagator.cpp:683-950) incl. the 24 IMU-intrinsic columns: the
    state-transition matrix equals the numerical Jacobian of the mean propagation (predict_mean_*, :482-681)."""
    maxdiff, maxF, asym, mind = _probe(exes, "propfd", method)[0]
    assert maxF == pytest.approx(1.0)
    assert maxdiff < 2e-6  # O(dt^2) consistency between the mean integrator and the linearisation, dt = 2.5 ms
    assert asym == 0.0 and mind > 0


def test_simulation_is_repeatable_and_consistent(exes, tmp_path):
    """ov_msckf/src/test_sim_repeat.cpp: two runs from the same seeds are bit-identical; the filter stays consistent
    (position ATE of the 200-frame mono run well below 0.2 m, the initial 1-sigma being 5 cm)."""
    e1, e2 = str(tmp_path / "a.txt"), str(tmp_path / "b.txt")
    kw = dict(exe=exes["oracle"], traj=TRAJ, cams=1, clones=11, msckf=50, pts=200, frames=200)
    r1 = simrun.run(est=e1, timing=str(tmp_path / "t.csv"), **kw)
    r2 = simrun.run(est=e2, **kw)
    assert open(e1).read() == open(e2).read()
    assert r1["frames"] == 200 and r1["state_dim"] == 15 + 24 + 1 + 14 + 6 * 11
    t, p, q, pg, qg = simrun.load_estimate(e1)
    assert simrun.ate_rmse(p, pg) == pytest.approx(r1["ate_pos_m"], rel=1e-9)
    assert r1["ate_pos_m"] < 0.2 and r1["ate_ori_deg"] < 1.0
    # chi² gate at the 95 % quantile: about 5 % of the features that reach it are rejected
    h = r1["status_hist"]
    assert 0.01 < h[8] / (h[0] + h[8]) < 0.12
    # timing file in the reference's CSV layout (core/VioManager.cpp:117-121)
    rows = open(tmp_path / "t.csv").read().strip().splitlines()
    assert rows[0].startswith("# timestamp (sec),tracking,propagation,msckf update") and len(rows) == 201
    assert len(rows[1].split(",")) == 6


def test_stereo_calib_off_runs(exes):
    r = simrun.run(exe=exes["oracle"], traj=TRAJ, cams=2, clones=8, msckf=20, pts=100, frames=60, calib=0)
    assert r["state_dim"] == 15 + 6 * 8 and r["ate_pos_m"] < 0.2


@pytest.mark.parametrize("method,tol_p,tol_v,tol_R", [("analytical", 1e-10, 1e-10, 1e-11), ("rk4", 1e-8, 1e-8, 1e-10), ("discrete", 2e-2, 2e-2, 1e-11)])
def test_mean_propagation_against_ode_solver(exes, method, tol_p, tol_v, tol_R):
    """Third-party pin of predict_mean_{analytic,rk4,discrete} (state/Propagator.cpp:482-681): one second of IMU kinematics with
    constant body-frame readings, against scipy's adaptive ODE solver on  R' = R [w]x,  v' = R a - g e_z,  p' = v  (R = R_ItoG).
    The analytical integrator is exact for constant readings, RK4 fourth order; the discrete one holds the attitude over each
    step for the velocity / position update (first order: its bound is that error, not a tolerance on the others)."""
    from scipy.integrate import solve_ivp
    K, dt = 400, 0.0025
    first, last = _probe(exes, "propmean", method, K)
    R0 = first.reshape(3, 3).T  # the probe prints R_GtoI
    w, a, g = np.array([0.3, -0.2, 0.5]), np.array([0.5, 9.6, 1.0]), np.array([0.0, 0.0, 9.81])

    def skew(x):
        return np.array([[0, -x[2], x[1]], [x[2], 0, -x[0]], [-x[1], x[0], 0]])

    def rhs(t, y):
        R = y[:9].reshape(3, 3)
        return np.concatenate([(R @ skew(w)).ravel(), y[12:15], R @ a - g])
    y0 = np.concatenate([R0.ravel(), [1, 2, 3], [0.5, -0.3, 0.2]])
    sol = solve_ivp(rhs, [0, K * dt], y0, method="DOP853", rtol=1e-13, atol=1e-14)
    yT = sol.y[:, -1]
    R_ref, p_ref, v_ref = yT[:9].reshape(3, 3), yT[9:12], yT[12:15]
    R_got, p_got, v_got = last[:9].reshape(3, 3).T, last[9:12], last[12:15]
    assert np.abs(R_got - R_ref).max() <= tol_R
    assert np.abs(p_got - p_ref).max() <= tol_p
    assert np.abs(v_got - v_ref).max() <= tol_v


def test_simulated_pixels_against_opencv_projection(exes):
    """Third-party pin of the simulator's measurement geometry (Simulator::project_pointcloud, sim/Simulator.cpp:455-499) and of
    the conventions it rests on: the ground-truth JPL quaternion q_GtoI and the extrinsics (q_ItoC, p_IinC) are turned into
    OpenCV's world-to-camera pose with scipy's (Hamilton) quaternions — R_JPL(q) = R_Hamilton(q)' — and the map points go
    through cv2.projectPoints with the camera's plumb-bob intrinsics. The simulator's noise-free float32 pixels agree to
    float32 rounding."""
    cv2 = pytest.importorskip("cv2")
    from scipy.spatial.transform import Rotation
    out = subprocess.run([exes["probe"], "simproj", TRAJ], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    i, frames, worst = 0, 0, 0.0
    while i < len(out):
        head = out[i].split()
        assert head[0] == "FRAME"
        n = int(head[2])
        v = np.array([float(x) for x in head[3:]])
        q_GtoI, p_IinG, q_ItoC, p_IinC, intr = v[0:4], v[4:7], v[7:11], v[11:14], v[14:22]
        pts = np.array([[float(x) for x in line.split()] for line in out[i + 1:i + 1 + n]])
        i += 1 + n
        if n == 0:
            continue
        R_GtoI = Rotation.from_quat(q_GtoI).as_matrix().T
        R_ItoC = Rotation.from_quat(q_ItoC).as_matrix().T
        R_GtoC = R_ItoC @ R_GtoI
        t = R_ItoC @ (-R_GtoI @ p_IinG) + p_IinC
        K = np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1.0]])
        img, _ = cv2.projectPoints(pts[:, :3].reshape(-1, 1, 3), cv2.Rodrigues(R_GtoC)[0], t, K, intr[4:8])
        err = np.abs(img.reshape(-1, 2) - pts[:, 3:5]).max()
        worst = max(worst, float(err))
        frames += 1
        # every simulated point is in front of the camera, inside the image
        pc = (R_GtoC @ pts[:, :3].T).T + t
        assert np.all(pc[:, 2] > 0.1) and np.all((pts[:, 3] >= 0) & (pts[:, 3] <= 752) & (pts[:, 4] >= 0) & (pts[:, 4] <= 480))
    assert frames >= 4
    assert worst <= 5e-4  # pixels (measured 5.6e-5): float32 normalised coordinates times a 458 px focal length


@pytest.mark.parametrize("method", ["discrete", "rk4", "analytical"])
def test_process_noise_scaling(exes, method):
    """Textbook pin of the discrete process noise (state/Propagator.cpp:510-530, 900-1016): over one IMU step of length dt the
    orientation / velocity errors integrate white noise of density sigma_w / sigma_a (variance sigma^2 dt) and the biases are
    random walks of density sigma_wb / sigma_ab (variance sigma^2 dt) — to first order in dt whatever the integrator; the
    position variance is third order (sigma_a^2 dt^3 / 3 up to the integrator's constant)."""
    head, diag = _probe(exes, "propq", method)
    n, dt, sw, sa, swb, sab = head
    assert int(n) == 39
    th, p, v, bg, ba = diag[0:3], diag[3:6], diag[6:9], diag[9:12], diag[12:15]
    assert np.allclose(th, sw**2 * dt, rtol=5e-3)
    assert np.allclose(v, sa**2 * dt, rtol=5e-3)
    assert np.allclose(bg, swb**2 * dt, rtol=1e-9) and np.allclose(ba, sab**2 * dt, rtol=1e-9)
    assert np.all(p > 0) and np.all(p < sa**2 * dt**3)  # between dt^3/4 and dt^3/3 times sigma_a^2 for every integrator
    assert np.all(p > 0.2 * sa**2 * dt**3)
