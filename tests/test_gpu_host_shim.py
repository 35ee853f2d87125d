"""The C++ host mirror (include/ovb200_host.hpp: State / StateHelper / UpdaterMSCKF with the reference's names) drives the
same update as the ctypes path: tests/cpp/host_shim_test.cpp rebuilds reference-style Feature objects (per-camera
unordered_maps, stale measurements, too-short tracks), calls UpdaterMSCKF::update and writes the results back."""
import os
import struct
import subprocess

import numpy as np
import pytest

from tests import golden_io
from open_vins_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_case(path, d, frame, feats, opts):
    C, K, N = frame.n_clones, frame.n_cams, d["P"].shape[0]
    hdr = np.zeros(16, dtype=np.int32)
    hdr[:12] = [0x0b200, C, K, N, feats.n_feats, feats.n_meas, opts.do_fej, opts.feat_rep, opts.do_calib_camera_pose,
                opts.do_calib_camera_intrinsics, opts.col_order, opts.compress]
    with open(path, "wb") as f:
        f.write(hdr.tobytes())
        f.write(np.array([opts.sigma_pix, opts.chi2_multipler], dtype=np.float64).tobytes())
        f.write((10.0 + 0.1 * np.arange(C)).astype(np.float64).tobytes())  # clone timestamps (oldest first)
        for a in (frame.clone_R, frame.clone_p, frame.clone_R_fej, frame.clone_p_fej):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
        f.write(frame.clone_off.astype(np.int32).tobytes())
        for a in (frame.cam_R, frame.cam_p, frame.cam_intr):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
        for a in (frame.cam_model, frame.cam_ext_off, frame.cam_intr_off):
            f.write(a.astype(np.int32).tobytes())
        f.write(np.ascontiguousarray(d["P"], dtype=np.float64).tobytes())
        f.write(feats.meas_off.astype(np.int32).tobytes())
        f.write(feats.cam.astype(np.uint8).tobytes())
        f.write(feats.clone.astype(np.uint16).tobytes())
        f.write(feats.uv.astype(np.float32).tobytes())
        f.write(feats.uvn.astype(np.float32).tobytes())


def _read_out(path):
    b = open(path, "rb").read()
    o = 0

    def take(dtype, n):
        nonlocal o
        a = np.frombuffer(b, dtype=dtype, count=n, offset=o)
        o += a.nbytes
        return a.copy()

    F, N, n_left, n_used, M2, nkeys = take(np.int32, 6)
    r = dict(F=int(F), N=int(N), n_left=int(n_left), n_used=int(n_used))
    r["status"], r["used"], r["to_delete"] = take(np.int32, F), take(np.int32, F), take(np.int32, F + 2)
    r["p_FinG"], r["anchor_t"] = take(np.float64, 3 * F).reshape(F, 3), take(np.float64, F)
    r["dx"], r["P"] = take(np.float64, N), take(np.float64, N * N).reshape(N, N)
    r["meas_off"], r["keys_off"] = take(np.int32, F + 1), take(np.int32, F + 1)
    r["cam"], r["clone"] = take(np.uint8, M2), take(np.uint16, M2)
    r["uv"], r["uvn"], r["keys"] = take(np.float32, 2 * M2), take(np.float32, 2 * M2), take(np.uint8, nkeys)
    assert o == len(b)
    return r


@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_io.names())
def test_cpp_host_mirror_runs_the_update(oracle, name, tmp_path):
    exe = build.build_host_test()
    d, frame, feats, opts = golden_io.load(name)
    case_path, out_path = str(tmp_path / "case.bin"), str(tmp_path / "out.bin")
    _write_case(case_path, d, frame, feats, opts)
    res = subprocess.run([exe, case_path, out_path], capture_output=True, text=True, timeout=300)
    # initialize -> marginalize_old_clone must shift the SLAM landmark ids too (checked inside the driver)
    assert "marginalize bookkeeping ok" in res.stdout, res.stdout + res.stderr
    assert res.returncode == 0, res.stderr
    r = _read_out(out_path)
    # every feature that entered is flagged for deletion, the two malformed extras included (UpdaterMSCKF.cpp:90-94, :276-279)
    assert r["to_delete"].all()
    # stale measurements were cleaned, too-short tracks dropped: the batch that reached the engine has the golden sizes
    assert r["F"] == feats.n_feats and len(r["cam"]) == feats.n_meas
    # The C++ layer visits a feature's cameras in std::unordered_map order, like the reference; the
    # golden files were generated with ascending camera ids. The anchor camera (most measurements, first visited wins
    # ties) and with it the triangulated point depend on that order, so the checker is the oracle on the SAME batch.
    feats2 = capi.FeatArrays(r["meas_off"], r["cam"], r["clone"], r["uv"], r["uvn"], r["keys_off"], r["keys"])
    ref = oracle.msckf_update(frame, feats2, opts, d["P"], dumps=False)
    assert np.array_equal(r["status"], ref["out"].status)
    assert np.array_equal(r["used"], (ref["out"].status == 0).astype(np.int32))
    assert r["n_left"] == r["n_used"] == int((ref["out"].status == 0).sum())
    tri = ~np.isnan(ref["out"].p_FinG[:, 0])
    # 1e-12 relative to the point's norm (BASELINE.json bar for triangulated points)
    assert (np.linalg.norm(r["p_FinG"][tri] - ref["out"].p_FinG[tri], axis=1) <= 1e-12 * np.linalg.norm(ref["out"].p_FinG[tri], axis=1)).all()
    assert np.linalg.norm(r["dx"] - ref["dx"]) <= 1e-9 * np.linalg.norm(ref["dx"])
    assert np.linalg.norm(r["P"] - ref["P"]) <= 1e-9 * np.linalg.norm(ref["P"])
    if frame.n_cams == 1:  # one camera: no order ambiguity, the golden numbers apply directly
        assert np.array_equal(r["status"], d["out_status"])
        assert np.linalg.norm(r["P"] - d["P_post"]) <= 1e-9 * np.linalg.norm(d["P_post"])
    # the ctypes path on the batch the C++ layer marshalled gives bit-identical results (same library, same inputs)
    eng = capi.Engine(max_state=256, max_feats=1024, max_meas=1024 * 48)
    eng.cov_set(d["P"])
    st, out, dx, stats = eng.msckf_update(frame, feats2, opts)
    assert st == 0
    assert np.array_equal(out.status, r["status"])
    assert np.array_equal(dx, r["dx"])
    assert np.array_equal(eng.cov_get(), r["P"])
    eng.close()


def test_host_header_compiles_standalone():
    """CPU-only: the C++ mirror is plain C++17 over the C ABI (no CUDA, no Eigen needed to compile against it)."""
    src = "#include \"ovb200_host.hpp\"\nint main() { ovb200::UpdaterOptions o; ovb200::FeatureInitializerOptions fo; ovb200::UpdaterSLAM us(o, o, fo); ovb200::UpdaterMSCKF um(o, fo); (void)us; (void)um; return o.sigma_pix > 0 ? 0 : 1; }\n"
    res = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-x", "c++", "-"],
                         input=src, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
