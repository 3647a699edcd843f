"""
Pins the CPU oracle (oracle/neuralfoil_oracle.py) against
  (a) the reference's own known-answer vectors (reference tests/test_golden_values.py:22-37),
  (b) fixtures produced by the unmodified reference in the build container
      (tests/golden/make_golden.py), all 8 sizes x 198 outputs,
  (c) the reference's structural invariants (reference tests/test_invariants.py,
      tests/test_robustness.py), restated against the oracle.
"""

import numpy as np
import pytest

from neuralfoil_b200 import synthetic
from oracle import neuralfoil_oracle as orc

KULFAN = synthetic.FIXED_KULFAN
SCALARS = ["analysis_confidence", "CL", "CD", "CM", "Top_Xtr", "Bot_Xtr"]

# reference tests/test_golden_values.py:29-37
GOLDEN_MEDIUM = {
    "analysis_confidence": 0.955711837783,
    "CL": 1.10332809679,
    "CD": 0.00919882438456,
    "CM": -0.110598030451,
    "Top_Xtr": 0.250543496781,
    "Bot_Xtr": 0.964878409058,
}


def test_reference_known_answer_medium(oracle):
    aero = oracle.get_aero_from_kulfan_parameters(KULFAN, alpha=5.0, Re=1e6, model_size="medium")
    for key, want in GOLDEN_MEDIUM.items():
        assert float(aero[key][0]) == pytest.approx(want, rel=1e-6, abs=1e-9), key


@pytest.mark.parametrize("size", orc.MODEL_SIZES)
def test_matches_reference_fixture_all_outputs(oracle, golden_dir, size):
    g = np.load(golden_dir / "golden_all_sizes.npz")
    assert list(g["keys"]) == orc.output_keys()
    got = oracle.evaluate_raw(g["raw"], size)
    want = g[f"out_{size}"]
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-300)


def test_matches_reference_fixture_config1(oracle, golden_dir):
    g = np.load(golden_dir / "golden_config1_xlarge.npz")
    aero = oracle.get_aero_from_kulfan_parameters(KULFAN, alpha=g["alpha"], Re=float(g["Re"]), model_size="xlarge")
    full = np.stack(list(aero.values()))
    np.testing.assert_allclose(full[:6], g["scalars"], rtol=1e-11)
    np.testing.assert_allclose(full[:, ::20], g["full_every_20"], rtol=1e-11)


def test_output_contract(oracle):
    """reference tests/test_invariants.py:93-118"""
    aero = oracle.get_aero_from_kulfan_parameters(KULFAN, alpha=3.0, Re=1e6, model_size="large")
    assert list(aero) == orc.output_keys() and len(aero) == 198
    for k, v in aero.items():
        assert v.shape == (1,) and v.dtype == np.float64 and np.isfinite(v).all(), k
    assert 0 <= aero["analysis_confidence"][0] <= 1 and aero["CD"][0] > 0
    for i in range(32):
        assert aero[f"upper_bl_theta_{i}"][0] > 0 and aero[f"lower_bl_H_{i}"][0] > 1


def test_symmetric_airfoil_zero_alpha(oracle):
    """reference tests/test_invariants.py:35-52"""
    w = np.array([0.15, 0.20, 0.15, 0.20, 0.15, 0.20, 0.15, 0.15])
    sym = dict(upper_weights=w, lower_weights=-w, leading_edge_weight=0.0, TE_thickness=0.002)
    for size in ("medium", "xlarge"):
        aero = oracle.get_aero_from_kulfan_parameters(sym, alpha=0.0, Re=1e6, model_size=size)
        assert abs(aero["CL"][0]) < 1e-14 and abs(aero["CM"][0]) < 1e-14
        assert aero["Top_Xtr"][0] == pytest.approx(aero["Bot_Xtr"][0], abs=1e-14)


def test_mirror_identity(oracle, golden_dir):
    """reference tests/test_invariants.py:55-90, plus equality with the reference's numbers"""
    g = np.load(golden_dir / "golden_mirror_large.npz")
    alpha = g["alpha"]
    kw = dict(Re=5e5, n_crit=8.0, model_size="large")
    mirror = dict(
        upper_weights=-KULFAN["lower_weights"],
        lower_weights=-KULFAN["upper_weights"],
        leading_edge_weight=-KULFAN["leading_edge_weight"],
        TE_thickness=KULFAN["TE_thickness"],
    )
    a = oracle.get_aero_from_kulfan_parameters(KULFAN, alpha=alpha, xtr_upper=0.7, xtr_lower=0.4, **kw)
    m = oracle.get_aero_from_kulfan_parameters(mirror, alpha=-alpha, xtr_upper=0.4, xtr_lower=0.7, **kw)
    np.testing.assert_allclose(np.stack(list(a.values())), g["out"], rtol=1e-11)
    np.testing.assert_allclose(np.stack(list(m.values())), g["out_mirror"], rtol=1e-11)
    tol = dict(rtol=0, atol=1e-12)
    np.testing.assert_allclose(m["CL"], -a["CL"], **tol)
    np.testing.assert_allclose(m["CD"], a["CD"], **tol)
    np.testing.assert_allclose(m["Top_Xtr"], a["Bot_Xtr"], **tol)
    for i in range(32):
        np.testing.assert_allclose(m[f"upper_bl_theta_{i}"], a[f"lower_bl_theta_{i}"], **tol)
        np.testing.assert_allclose(m[f"upper_bl_ue/vinf_{i}"], -a[f"lower_bl_ue/vinf_{i}"], **tol)


def test_vectorised_matches_scalar_loop(oracle):
    """reference tests/test_invariants.py:121-136"""
    alphas = np.array([-10.0, -2.0, 3.0, 8.0, 15.0])
    Res = np.array([5e4, 2e5, 1e6, 5e6, 2e7])
    batch = oracle.get_aero_from_kulfan_parameters(KULFAN, alpha=alphas, Re=Res, model_size="medium")
    for i in range(5):
        one = oracle.get_aero_from_kulfan_parameters(KULFAN, alpha=alphas[i], Re=Res[i], model_size="medium")
        for k in SCALARS:
            np.testing.assert_allclose(batch[k][i], one[k][0], rtol=1e-9, atol=1e-12)


def test_error_paths(oracle):
    """reference tests/test_robustness.py:60-83"""
    with pytest.raises(ValueError, match="model_size"):
        oracle.get_aero_from_kulfan_parameters(KULFAN, alpha=5.0, Re=1e6, model_size="gigantic")
    with pytest.raises(ValueError, match="same length"):
        oracle.get_aero_from_kulfan_parameters(KULFAN, alpha=np.linspace(0, 5, 3), Re=np.full(4, 1e6))
    bad = dict(KULFAN, upper_weights=np.full(7, 0.2), lower_weights=np.full(7, -0.1))
    with pytest.raises(ValueError, match="8 CST weights per side"):
        oracle.get_aero_from_kulfan_parameters(bad, alpha=5.0, Re=1e6)


def test_bl_x_points():
    """reference tests/test_invariants.py:139-144"""
    x = orc.bl_x_points
    assert len(x) == 32 and np.all(np.diff(x) > 0) and np.all((x > 0) & (x < 1))


def test_synthetic_xxxlarge_is_deterministic():
    a = synthetic.synthetic_xxxlarge_params()
    b = synthetic.synthetic_xxxlarge_params()
    assert [a[k].shape for k in a] == [b[k].shape for k in b]
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert sum(v.size for v in a.values()) == 1_428_166  # reference README.md:246


def test_jacobian_matches_finite_differences():
    """The oracle's analytic Jacobian against central differences of the pinned value oracle."""
    o = orc.Oracle()
    raw = synthetic.sample_training_distribution(24, seed=5)
    for size in ("xxsmall", "large"):
        out, jac = o.jacobian_raw(raw, size)
        ref = o.evaluate_raw(raw, size)
        np.testing.assert_allclose(out, ref, rtol=1e-11, atol=0)
        inside = (ref[4:6] > 1e-3) & (ref[4:6] < 1 - 1e-3) | (ref[4:6] == 0) | (ref[4:6] == 1)   # away from the Xtr clip
        for t in range(23):
            s_t = np.maximum(np.abs(raw[t]), 1e-2)
            h = 1e-6 * s_t
            rp, rm = raw.copy(), raw.copy()
            rp[t] += h
            rm[t] -= h
            fd = (o.evaluate_raw(rp, size) - o.evaluate_raw(rm, size)) / (2 * h)
            err = np.abs(jac[:, t, :] - fd) * s_t / (np.abs(fd) * s_t + 1e-3 * (np.abs(ref) + 1e-6))
            err[4:6] = np.where(inside, err[4:6], 0.0)
            assert err.max() < 2e-4, (size, t, err.max())
