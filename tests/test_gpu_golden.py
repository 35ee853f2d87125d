"""GPU engine against the committed golden update cases (no oracle in the loop)."""
import numpy as np
import pytest

from tests import golden_io
from open_vins_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_io.names())
def test_engine_matches_golden(name):
    d, frame, feats, opts = golden_io.load(name)
    eng = capi.Engine(max_state=256, max_feats=256, max_meas=8192)
    eng.cov_set(d["P"])
    st, out, dx, stats = eng.msckf_update(frame, feats, opts)
    assert st == 0
    assert np.array_equal(out.status, d["out_status"])
    ok = d["out_status"] == 0
    assert stats.n_feats_used == int(d["n_used"])
    tri = ~np.isnan(d["out_p_FinG"][:, 0])
    # BASELINE.json bars: triangulated points 1e-12 relative, post-update state/covariance 1e-9 relative Frobenius
    assert (np.linalg.norm(out.p_FinG[tri] - d["out_p_FinG"][tri], axis=1) <= 1e-12 * np.linalg.norm(d["out_p_FinG"][tri], axis=1)).all()
    np.testing.assert_allclose(out.chi2[ok], d["out_chi2"][ok], rtol=1e-8)
    assert np.linalg.norm(dx - d["dx"]) <= 1e-9 * np.linalg.norm(d["dx"])
    Pg = eng.cov_get()
    assert np.linalg.norm(Pg - d["P_post"]) <= 1e-9 * np.linalg.norm(d["P_post"])
    eng.close()


@pytest.mark.parametrize("name", golden_io.slam_names())
def test_engine_matches_slam_golden(name):
    d, frame, feats, lms, opts = golden_io.load_slam(name)
    eng = capi.Engine(max_state=256, max_feats=256, max_meas=8192)
    eng.cov_set(d["P"])
    st, out, dx, stats = eng.slam_update(frame, feats, lms, opts)
    assert st == 0 and np.array_equal(out.status, d["out_status"]) and stats.n_feats_used == int(d["n_used"])
    ok = d["out_status"] == 0
    np.testing.assert_allclose(out.chi2[ok], d["out_chi2"][ok], rtol=1e-8)
    assert np.linalg.norm(dx - d["dx"]) <= 1e-9 * np.linalg.norm(d["dx"])
    assert np.linalg.norm(eng.cov_get() - d["P_post"]) <= 1e-9 * np.linalg.norm(d["P_post"])
    eng.close()
