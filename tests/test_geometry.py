"""
The geometry front doors (reference main.py:335-468), checked on the CPU:
`neuralfoil_b200.geometry` + the oracle must reproduce the reference's own golden numbers for the full
NACA 4412 -> xlarge pipeline (reference tests/test_golden_values.py:48-55, 70-79), which pins the NACA generator,
the normalisation and the Kulfan fit against AeroSandbox's without needing AeroSandbox.
"""

import numpy as np
import pytest

from neuralfoil_b200 import geometry

# reference tests/test_golden_values.py:48-55
GOLDEN_NACA4412_XLARGE = {
    "analysis_confidence": 0.968356999039,
    "CL": 1.02508809111,
    "CD": 0.00791137686241,
    "CM": -0.0992982374935,
    "Top_Xtr": 0.402283708087,
    "Bot_Xtr": 1.0,
}


def pipeline(oracle, coordinates, alpha, Re, model_size="xlarge"):
    """main.py:360-387 with the oracle standing in for the learned core."""
    nrm = geometry.normalize(coordinates)
    kulfan = geometry.kulfan_parameters_from_coordinates(nrm["coordinates"])
    da, sc = nrm["rotation_angle"], nrm["scale_factor"]
    aero = oracle.get_aero_from_kulfan_parameters(kulfan, alpha=alpha + da, Re=Re / sc, model_size=model_size)
    x_qc = -nrm["x_translation"] + 0.25 * (1.0 / sc * np.cos(np.radians(da))) - 0.25
    y_qc = -nrm["y_translation"] + 0.25 * (1.0 / sc * np.sin(np.radians(-da)))
    aero["CM"] = aero["CM"] + (-aero["CL"] * x_qc + aero["CD"] * y_qc)
    return aero


def test_reference_golden_airfoil_pipeline(oracle):
    aero = pipeline(oracle, geometry.Airfoil("naca4412").coordinates, 5.0, 1e6)
    for key, want in GOLDEN_NACA4412_XLARGE.items():
        assert float(aero[key][0]) == pytest.approx(want, rel=1e-6, abs=1e-9), key  # the reference's own tolerance


def test_naca_generator_shape():
    c = geometry.naca4_coordinates("naca0012")
    assert c.shape == (399, 2) and c[0, 0] == 1.0 and c[-1, 0] == 1.0 and c[199, 0] == 0.0
    np.testing.assert_allclose(c[:200, 1], -c[199:, 1][::-1], atol=1e-15)  # symmetric section
    assert abs(np.max(c[:, 1]) - 0.06) < 1e-3


def test_normalize_undoes_a_similarity_transform():
    c = geometry.naca4_coordinates("naca2412")
    base = geometry.normalize(c)
    ang = np.radians(7.0)
    rot = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
    moved = (c * 2.5) @ rot.T + np.array([0.3, -0.2])
    nrm = geometry.normalize(moved)
    np.testing.assert_allclose(nrm["coordinates"], base["coordinates"], atol=1e-12)
    assert nrm["scale_factor"] == pytest.approx(base["scale_factor"] / 2.5, rel=1e-12)
    assert nrm["rotation_angle"] == pytest.approx(base["rotation_angle"] - 7.0, abs=1e-9)
    te = (nrm["coordinates"][0] + nrm["coordinates"][-1]) / 2
    np.testing.assert_allclose(te, [1.0, 0.0], atol=1e-12)
    np.testing.assert_allclose(nrm["coordinates"][geometry.le_index(nrm["coordinates"])], [0.0, 0.0], atol=1e-15)


def test_kulfan_fit_recovers_a_cst_airfoil():
    """Coordinates generated from known CST parameters are fitted back exactly."""
    from math import comb

    rng = np.random.default_rng(0)
    up, lo, le, te = 0.2 + 0.05 * rng.random(8), -0.1 - 0.05 * rng.random(8), 0.1, 0.004
    x = geometry.cosspace(0, 1, 150)

    def surf(w, sign):
        s = sum(w[i] * comb(7, i) * x**i * (1 - x) ** (7 - i) for i in range(8))
        return x**0.5 * (1 - x) * s + le * x * (1 - x) ** 8.5 + sign * te * x / 2

    c = np.stack([np.concatenate([x[::-1], x[1:]]), np.concatenate([surf(up, +1)[::-1], surf(lo, -1)[1:]])], axis=1)
    fit = geometry.kulfan_parameters_from_coordinates(c)
    np.testing.assert_allclose(fit["upper_weights"], up, atol=1e-9)
    np.testing.assert_allclose(fit["lower_weights"], lo, atol=1e-9)
    assert fit["leading_edge_weight"] == pytest.approx(le, abs=1e-9) and fit["TE_thickness"] == pytest.approx(te, abs=1e-9)


def test_entry_points_agree(oracle, tmp_path):
    """reference tests/test_entry_points.py:22-48: airfoil vs coordinates vs .dat round trip (geometry part)."""
    af = geometry.Airfoil("naca4412")
    a = pipeline(oracle, af.coordinates, 4.0, 2e6)
    dat = tmp_path / "naca4412.dat"
    dat.write_text(af.name + "\n" + "\n".join(f"{x:.16f} {y:.16f}" for x, y in af.coordinates))
    coords = geometry.coordinates_from_raw_dat(dat.read_text().splitlines())
    assert coords.shape == af.coordinates.shape
    b = pipeline(oracle, coords, 4.0, 2e6)
    for key in ("analysis_confidence", "CL", "CD", "CM", "Top_Xtr", "Bot_Xtr"):
        np.testing.assert_allclose(a[key], b[key], rtol=1e-6, err_msg=key)
    with pytest.raises(ValueError):
        geometry.Airfoil("e387")  # database lookups are not part of this package


def test_kulfan_airfoil_objects_skip_the_fit():
    """reference main.py:336: `airfoil` may be a KulfanAirfoil; its 8 + 8 parameters are used as they are."""
    coords = geometry.normalize(geometry.Airfoil("naca4412").coordinates)["coordinates"]
    params = geometry.kulfan_parameters_from_coordinates(coords)
    kaf = geometry.KulfanAirfoil("fit", **params)
    got = geometry.kulfan_parameters_of(kaf)
    assert got is not None and all(np.array_equal(got[k], params[k]) for k in params)
    assert geometry.kulfan_parameters_of(geometry.Airfoil("naca0012")) is None
    # the generated coordinates are the forward map of the fit: refitting them returns the parameters
    back = geometry.kulfan_parameters_from_coordinates(kaf.coordinates)
    for k in params:
        np.testing.assert_allclose(back[k], params[k], atol=1e-9)
    # and they are already normalised: leading edge at the origin, unit chord along x
    nrm = geometry.normalize(kaf.coordinates)
    assert abs(nrm["rotation_angle"]) < 1e-9 and abs(nrm["scale_factor"] - 1) < 1e-12 and nrm["x_translation"] == 0
    coarse = geometry.KulfanAirfoil("coarse", lower_weights=-np.ones(6) * 0.1, upper_weights=np.ones(6) * 0.1)
    assert geometry.kulfan_parameters_of(coarse) is None  # another resolution: refitted from its coordinates

    class ParamsOnly:
        kulfan_parameters = {"lower_weights": np.zeros(5), "upper_weights": np.zeros(5), "leading_edge_weight": 0.0, "TE_thickness": 0.0}

    with pytest.raises(ValueError, match="8 CST weights per side"):
        geometry.kulfan_parameters_of(ParamsOnly())
