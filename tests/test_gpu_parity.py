"""
Parity of the CUDA path (through the C ABI) against
  * fixtures written by the unmodified reference (tests/golden/*.npz, tests/golden/make_golden.py),
  * the reference's own known-answer vector (reference tests/test_golden_values.py:22-37),
  * the CPU oracle on seeded inputs, every model size, ragged and tiny batches,
  * the reference's structural invariants (reference tests/test_invariants.py, test_robustness.py),
  * size-independent properties at BASELINE.json's batch sizes (mirror identity, batch-position
    independence, a subsample against the oracle).
The tolerance is defined and justified in tests/parity.py.
"""

import numpy as np
import pytest

from neuralfoil_b200 import synthetic
from neuralfoil_b200.main import MODEL_SIZES, output_keys

import parity

pytestmark = pytest.mark.gpu

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


def stack(aero: dict) -> np.ndarray:
    return np.stack([aero[k] for k in output_keys()])


def mirrored(raw: np.ndarray) -> np.ndarray:
    """The airfoil mirrored about its chord, at -alpha, transition locations swapped (test_invariants.py:55-90)."""
    m = raw.copy()
    m[0:8] = -raw[8:16]
    m[8:16] = -raw[0:8]
    m[16] = -raw[16]
    m[18] = -raw[18]
    m[21] = raw[22]
    m[22] = raw[21]
    return m


def unmirror_outputs(out: np.ndarray) -> np.ndarray:
    """What the mirrored airfoil's outputs must look like in terms of the original's."""
    n = 32
    u = out.copy()
    u[1] = -out[1]
    u[3] = -out[3]
    u[4], u[5] = out[5], out[4]
    u[6 : 6 + 2 * n], u[6 + 3 * n : 6 + 5 * n] = out[6 + 3 * n : 6 + 5 * n], out[6 : 6 + 2 * n]
    u[6 + 2 * n : 6 + 3 * n] = -out[6 + 5 * n : 6 + 6 * n]
    u[6 + 5 * n : 6 + 6 * n] = -out[6 + 2 * n : 6 + 3 * n]
    return u


# ---------------------------------------------------------------------------------------------
# Known answers
# ---------------------------------------------------------------------------------------------
def test_reference_known_answer_medium(evaluator):
    aero = evaluator.get_aero_from_kulfan_parameters(KULFAN, alpha=5.0, Re=1e6, model_size="medium")
    for key, want in GOLDEN_MEDIUM.items():
        # reference tolerance is rel 1e-6 for its own float64 path; fp32 bound from tests/parity.py
        assert float(aero[key][0]) == pytest.approx(want, rel=1e-5, abs=6e-5), key


@pytest.mark.parametrize("size", MODEL_SIZES)
def test_matches_reference_fixture_all_outputs(evaluator, golden_dir, size):
    g = np.load(golden_dir / "golden_all_sizes.npz")
    raw = g["raw"]
    got = evaluator.evaluate_host(raw, size, out_dtype=np.float64)
    assert got.shape == (198, raw.shape[1]) and got.dtype == np.float64
    parity.assert_parity(got, g[f"out_{size}"], raw[19], label=size)


def test_matches_reference_fixture_config1(evaluator, golden_dir):
    """BASELINE.json configs[0]: alpha sweep of 1000 points, Re = 5e6, xlarge, through the dict API."""
    g = np.load(golden_dir / "golden_config1_xlarge.npz")
    aero = evaluator.get_aero_from_kulfan_parameters(KULFAN, alpha=g["alpha"], Re=float(g["Re"]), model_size="xlarge")
    full = stack(aero)
    a = parity.LATENT_ATOL  # the latent-space bound pushed through each scalar's decode (tests/parity.py)
    for i, tol in enumerate([a / 4, a / 2, None, a / 20, a, a]):
        if tol is None:
            np.testing.assert_allclose(full[i], g["scalars"][i], rtol=2 * a + 1e-5)
        else:
            np.testing.assert_allclose(full[i], g["scalars"][i], rtol=1e-5, atol=tol)
    parity.assert_parity(full[:, ::20], g["full_every_20"], np.full(50, float(g["Re"])))


# ---------------------------------------------------------------------------------------------
# Against the oracle, every size, both I/O precisions, ragged sizes
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", MODEL_SIZES)
def test_matches_oracle_random_batch(evaluator, oracle, size):
    raw = synthetic.sample_training_distribution(5003, seed=11)  # not a multiple of any tile size
    want = oracle.evaluate_raw(raw, size)
    got64 = evaluator.evaluate_host(raw, size, out_dtype=np.float64)
    rep = parity.assert_parity(got64, want, raw[19], label=f"{size} f64 io")
    # fp32 inputs and outputs: the oracle sees the same rounded inputs
    raw32 = raw.astype(np.float32)
    got32 = evaluator.evaluate_host(raw32, size, out_dtype=np.float32)
    assert got32.dtype == np.float32
    want32 = oracle.evaluate_raw(raw32.astype(np.float64), size)
    parity.assert_parity(got32, want32, raw32[19].astype(np.float64), label=f"{size} f32 io")
    # the CUDA path must be no worse than a plain fp32 evaluation by more than a small factor
    yard = parity.error_report(oracle.evaluate_raw(raw, size, fp32_mlp=True), want, raw[19])
    assert rep["median_rel"] <= 3 * yard["median_rel"] + 1e-7, (rep, yard)
    assert rep["worst_ratio"] <= 4 * yard["worst_ratio"] + 0.05, (rep, yard)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 31, 32, 33, 63, 65, 255, 257, 1000])
@pytest.mark.parametrize("size", ["xxsmall", "large", "xxxlarge"])
def test_ragged_batch_sizes(evaluator, oracle, size, n):
    raw = synthetic.sample_training_distribution(n, seed=100 + n)
    got = evaluator.evaluate_host(raw, size, out_dtype=np.float64)
    assert got.shape == (198, n)
    parity.assert_parity(got, oracle.evaluate_raw(raw, size), raw[19], label=f"{size} n={n}")


def test_polar_grid_broadcast_rows(evaluator, oracle):
    """BASELINE.json configs[1] shape: one airfoil (18 broadcast scalars), 3 varying flow columns."""
    raw = synthetic.polar_grid(12, 11, 10)
    kulfan, flow = synthetic.kulfan_dict_from_raw(raw[:, :1])
    aero = evaluator.get_aero_from_kulfan_parameters(
        {k: (v[:, 0] if v.ndim == 2 else float(v[0])) for k, v in kulfan.items()},
        alpha=raw[18], Re=raw[19], n_crit=raw[20], model_size="large",
    )
    parity.assert_parity(stack(aero), oracle.evaluate_raw(raw, "large"), raw[19])


# ---------------------------------------------------------------------------------------------
# The reference's invariants (exact in the reference; must stay exact in fp32)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", ["medium", "xlarge", "xxxlarge"])
def test_symmetric_airfoil_zero_alpha(evaluator, size):
    """reference tests/test_invariants.py:35-52"""
    w = np.array([0.15, 0.20, 0.15, 0.20, 0.15, 0.20, 0.15, 0.15])
    sym = dict(upper_weights=w, lower_weights=-w, leading_edge_weight=0.0, TE_thickness=0.002)
    aero = evaluator.get_aero_from_kulfan_parameters(sym, alpha=0.0, Re=1e6, model_size=size)
    assert abs(aero["CL"][0]) < 1e-14 and abs(aero["CM"][0]) < 1e-14
    assert aero["Top_Xtr"][0] == pytest.approx(aero["Bot_Xtr"][0], abs=1e-14)


def test_mirror_identity_reference_case(evaluator, golden_dir):
    """reference tests/test_invariants.py:55-90, plus agreement with the reference's numbers"""
    g = np.load(golden_dir / "golden_mirror_large.npz")
    alpha = g["alpha"]
    kw = dict(Re=5e5, n_crit=8.0, model_size="large")
    mirror = dict(
        upper_weights=-KULFAN["lower_weights"],
        lower_weights=-KULFAN["upper_weights"],
        leading_edge_weight=-KULFAN["leading_edge_weight"],
        TE_thickness=KULFAN["TE_thickness"],
    )
    a = evaluator.get_aero_from_kulfan_parameters(KULFAN, alpha=alpha, xtr_upper=0.7, xtr_lower=0.4, **kw)
    m = evaluator.get_aero_from_kulfan_parameters(mirror, alpha=-alpha, xtr_upper=0.4, xtr_lower=0.7, **kw)
    re = np.full(alpha.size, 5e5)
    parity.assert_parity(stack(a), g["out"], re)
    parity.assert_parity(stack(m), g["out_mirror"], re)
    tol = dict(rtol=0, atol=1e-12)  # the reference's own tolerance
    np.testing.assert_allclose(m["CL"], -a["CL"], **tol)
    np.testing.assert_allclose(m["CM"], -a["CM"], **tol)
    np.testing.assert_allclose(m["CD"], a["CD"], **tol)
    np.testing.assert_allclose(m["analysis_confidence"], a["analysis_confidence"], **tol)
    np.testing.assert_allclose(m["Top_Xtr"], a["Bot_Xtr"], **tol)
    np.testing.assert_allclose(m["Bot_Xtr"], a["Top_Xtr"], **tol)
    for i in range(32):
        for q in ("theta", "H"):
            np.testing.assert_allclose(m[f"upper_bl_{q}_{i}"], a[f"lower_bl_{q}_{i}"], **tol)
            np.testing.assert_allclose(m[f"lower_bl_{q}_{i}"], a[f"upper_bl_{q}_{i}"], **tol)
        np.testing.assert_allclose(m[f"upper_bl_ue/vinf_{i}"], -a[f"lower_bl_ue/vinf_{i}"], **tol)
        np.testing.assert_allclose(m[f"lower_bl_ue/vinf_{i}"], -a[f"upper_bl_ue/vinf_{i}"], **tol)


def test_output_contract(evaluator):
    """reference tests/test_invariants.py:93-118"""
    aero = evaluator.get_aero_from_kulfan_parameters(KULFAN, alpha=3.0, Re=1e6, model_size="large")
    assert list(aero) == output_keys() and len(aero) == 198
    for k, v in aero.items():
        assert isinstance(v, np.ndarray) and v.shape == (1,) and v.dtype == np.float64, k
        assert np.isfinite(v).all() and v.flags["C_CONTIGUOUS"] and v.flags["WRITEABLE"], k
    assert 0 <= aero["analysis_confidence"][0] <= 1 and aero["CD"][0] > 0
    assert 0 <= aero["Top_Xtr"][0] <= 1 and 0 <= aero["Bot_Xtr"][0] <= 1
    for i in range(32):
        for side in ("upper", "lower"):
            assert aero[f"{side}_bl_theta_{i}"][0] > 0 and aero[f"{side}_bl_H_{i}"][0] > 1


def test_vectorised_matches_scalar_loop(evaluator):
    """reference tests/test_invariants.py:121-136 -- here bit for bit: a query's result may not depend on its position"""
    alphas = np.array([-10.0, -2.0, 3.0, 8.0, 15.0])
    Res = np.array([5e4, 2e5, 1e6, 5e6, 2e7])
    batch = evaluator.get_aero_from_kulfan_parameters(KULFAN, alpha=alphas, Re=Res, model_size="medium")
    for i in range(5):
        one = evaluator.get_aero_from_kulfan_parameters(KULFAN, alpha=alphas[i], Re=Res[i], model_size="medium")
        for k in output_keys():
            assert batch[k][i] == one[k][0], (k, i)


@pytest.mark.parametrize("size", [s for s in MODEL_SIZES])
def test_all_model_sizes_load_and_run(evaluator, size):
    """reference tests/test_robustness.py:31-40 (xxxlarge: seeded synthetic weights, physics ranges not asserted)"""
    aero = evaluator.get_aero_from_kulfan_parameters(KULFAN, alpha=5.0, Re=1e6, model_size=size)
    for k in SCALARS:
        assert np.isfinite(aero[k]).all(), k
    if size != "xxxlarge" or "xxxlarge" not in evaluator.model_dims or not _is_synthetic(evaluator):
        assert 0 < aero["CD"][0] < 1 and -1 < aero["CL"][0] < 3


def _is_synthetic(evaluator) -> bool:
    return not (evaluator.weights_dir / "nn-xxxlarge.npz").exists()


@pytest.mark.parametrize("alpha", [-180.0, -90.0, 90.0, 180.0])
@pytest.mark.parametrize("Re", [1e2, 1e10])
def test_always_returns_an_answer(evaluator, oracle, alpha, Re):
    """reference tests/test_robustness.py:43-57"""
    aero = evaluator.get_aero_from_kulfan_parameters(KULFAN, alpha=alpha, Re=Re, model_size="large")
    for k in SCALARS:
        assert np.isfinite(aero[k]).all(), k
    assert 0 <= aero["analysis_confidence"][0] <= 1 and aero["CD"][0] > 0
    want = oracle.get_aero_from_kulfan_parameters(KULFAN, alpha=alpha, Re=Re, model_size="large")
    parity.assert_parity(stack(aero), np.stack(list(want.values())), np.array([Re]))


def test_error_paths(evaluator):
    """reference tests/test_robustness.py:60-83"""
    with pytest.raises(ValueError, match="model_size"):
        evaluator.get_aero_from_kulfan_parameters(KULFAN, alpha=5.0, Re=1e6, model_size="gigantic")
    with pytest.raises(ValueError, match="same length"):
        evaluator.get_aero_from_kulfan_parameters(KULFAN, alpha=np.linspace(0, 5, 3), Re=np.full(4, 1e6))
    bad = dict(KULFAN, upper_weights=np.full(7, 0.2), lower_weights=np.full(7, -0.1))
    with pytest.raises(ValueError, match="8 CST weights per side"):
        evaluator.get_aero_from_kulfan_parameters(bad, alpha=5.0, Re=1e6)


def test_nan_inputs_propagate(evaluator):
    """NaN input -> NaN outputs, silently; Re <= 0 -> NaN."""
    aero = evaluator.get_aero_from_kulfan_parameters(KULFAN, alpha=np.array([np.nan, 2.0]), Re=1e6, model_size="medium")
    assert np.isnan(aero["CL"][0]) and np.isnan(aero["Top_Xtr"][0]) and np.isfinite(aero["CL"][1])
    aero = evaluator.get_aero_from_kulfan_parameters(KULFAN, alpha=2.0, Re=-1.0, model_size="medium")
    assert np.isnan(aero["CL"][0])


def test_kulfan_weights_as_8_by_n(evaluator, oracle):
    """(8, N) weight arrays, the layout of studies/naca_airfoil_extrapolation_study.py:44-53"""
    raw = synthetic.sample_training_distribution(77, seed=3)
    kulfan, flow = synthetic.kulfan_dict_from_raw(raw)
    aero = evaluator.get_aero_from_kulfan_parameters(kulfan, model_size="small", **flow)
    parity.assert_parity(stack(aero), oracle.evaluate_raw(raw, "small"), raw[19])


# ---------------------------------------------------------------------------------------------
# Device-resident API and BASELINE-sized batches
# ---------------------------------------------------------------------------------------------
def test_device_api_equals_host_api(evaluator):
    import torch

    raw = synthetic.sample_training_distribution(4099, seed=21).astype(np.float32)
    host = evaluator.evaluate_host(raw, "xlarge", out_dtype=np.float32)
    dev = evaluator.evaluate_device(torch.from_numpy(raw).cuda(), "xlarge")
    torch.cuda.synchronize()
    assert dev.shape == (198, 4099) and dev.dtype == torch.float32
    np.testing.assert_array_equal(dev.cpu().numpy(), host)
    # unaligned leading dimension -> scalar-store path, same numbers
    out = torch.empty((198, 4101), dtype=torch.float32, device="cuda")
    dev2 = evaluator.evaluate_device(torch.from_numpy(raw).cuda(), "xlarge", out=out)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dev2.cpu().numpy(), host)


def test_host_chunking_is_invisible(evaluator):
    raw = synthetic.sample_training_distribution(50_000, seed=8).astype(np.float32)
    one = evaluator.evaluate_host(raw, "small", out_dtype=np.float32)
    many = evaluator.evaluate_host(raw, "small", out_dtype=np.float32, chunk=7000)
    np.testing.assert_array_equal(one, many)


def test_pageable_host_arrays_are_staged_without_changing_a_bit(evaluator):
    """nfb_eval_host with pageable (NumPy) memory goes through the library's pinned staging buffers and host threads, and
    the GPU writes float32 even for a float64 result; with pinned memory the copies are direct.  Same bits either way."""
    import torch

    n = 300_007  # three staged chunks of 128 Ki queries, the last one ragged
    raw = synthetic.sample_training_distribution(n, seed=31)
    dev = evaluator.evaluate_device(torch.from_numpy(raw.astype(np.float32)).cuda(), "medium")
    torch.cuda.synchronize()
    want32 = dev.cpu().numpy()
    got32 = evaluator.evaluate_host(raw.astype(np.float32), "medium", out_dtype=np.float32)
    np.testing.assert_array_equal(got32, want32)
    got64 = evaluator.evaluate_host(raw.astype(np.float32), "medium", out_dtype=np.float64)
    assert got64.dtype == np.float64
    np.testing.assert_array_equal(got64, want32.astype(np.float64))
    np.testing.assert_array_equal(evaluator.evaluate_host(raw.astype(np.float32), "medium", out_dtype=np.float64, chunk=50_000), got64)
    sc = evaluator.evaluate_host(raw.astype(np.float32), "medium", out_dtype=np.float64, outputs="scalars")
    np.testing.assert_array_equal(sc, got64[:6])
    # float64 inputs, broadcast rows (the drop-in call): one airfoil, varying alpha / Re
    kul = dict(upper_weights=raw[0:8, 0], lower_weights=raw[8:16, 0], leading_edge_weight=raw[16, 0], TE_thickness=raw[17, 0])
    aero = evaluator.get_aero_from_kulfan_parameters(kul, alpha=raw[18], Re=raw[19], n_crit=7.0, model_size="medium")
    rows = raw.copy()
    rows[0:18] = raw[0:18, :1]
    rows[20], rows[21], rows[22] = 7.0, 1.0, 1.0
    ref = evaluator.evaluate_device(torch.from_numpy(rows).cuda(), "medium", out_dtype=torch.float64)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(np.stack([aero[k] for k in output_keys()]), ref.cpu().numpy())


@pytest.mark.parametrize("size,n", [("xxxlarge", 1_000_000), ("xlarge", 1_000_000), ("xxsmall", 2_000_000)])
def test_large_batch_properties(evaluator, oracle, size, n):
    """
    BASELINE-sized batches, checked through properties that do not need the oracle at full size:
    (a) mirror identity, bit for bit, over the whole batch; (b) batch-position independence: a
    permuted batch gives permuted results, bit for bit; (c) a strided subsample against the oracle.
    """
    import torch

    raw = synthetic.sample_training_distribution(n, seed=2).astype(np.float32)
    d_raw = torch.from_numpy(raw).cuda()
    out = evaluator.evaluate_device(d_raw, size)
    out_m = evaluator.evaluate_device(torch.from_numpy(mirrored(raw)).cuda(), size)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(0)).cuda()
    out_p = evaluator.evaluate_device(d_raw[:, perm].contiguous(), size)
    torch.cuda.synchronize()
    assert torch.isfinite(out[:6]).all()
    assert torch.equal(out_p, out[:, perm])
    o, m = out.cpu().numpy(), out_m.cpu().numpy()
    np.testing.assert_array_equal(m, unmirror_outputs(o))
    idx = np.arange(0, n, n // 2000)
    want = oracle.evaluate_raw(raw[:, idx].astype(np.float64), size)
    parity.assert_parity(o[:, idx], want, raw[19, idx].astype(np.float64), label=f"{size} n={n}")


# ---------------------------------------------------------------------------------------------
# Tensor-core variant (tcgen05, fp16 operands): stated looser tolerance, same exact invariants
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", ["xxlarge", "xxxlarge"])  # xxlarge: the reference's TRAINED weights on the tensor path
def test_tensor_variant_matches_oracle(evaluator, oracle, size):
    raw = synthetic.sample_training_distribution(20011, seed=11)
    want = oracle.evaluate_raw(raw, size)
    got = evaluator.evaluate_host(raw, size, out_dtype=np.float64, variant="tensor")
    rep = parity.assert_parity(got, want, raw[19], label=f"tensor {size}", variant="tensor")
    # and it must be a real low-precision evaluation of the same network, not something else: close to the fp32 path
    ref = evaluator.evaluate_host(raw, size, out_dtype=np.float64)
    assert np.nanmax(np.abs(got[1] - ref[1])) <= 7.5e-2 and rep["median_rel"] > 1e-6  # CL = z / 2, worst-case latent 0.15


@pytest.mark.parametrize("n", [1, 3, 63, 64, 65, 127, 129, 1000])
@pytest.mark.parametrize("size", ["xxlarge", "xxxlarge"])
def test_tensor_variant_ragged_batch_sizes(evaluator, oracle, size, n):
    raw = synthetic.sample_training_distribution(n, seed=300 + n)
    got = evaluator.evaluate_host(raw, size, out_dtype=np.float64, variant="tensor")
    assert got.shape == (198, n)
    parity.assert_parity(got, oracle.evaluate_raw(raw, size), raw[19], variant="tensor")


def test_tensor_variant_golden_and_far_ood(evaluator, golden_dir):
    """The reference-written fixture (incl. alpha = +-90, +-180, Re = 1e2, 1e10) with the tensor tolerance; finite everywhere."""
    g = np.load(golden_dir / "golden_all_sizes.npz")
    raw = g["raw"]
    for size in ("xxlarge", "xxxlarge"):
        got = evaluator.evaluate_host(raw, size, out_dtype=np.float64, variant="tensor")
        assert np.isfinite(got[:6]).all(), size
        parity.assert_parity(got, g[f"out_{size}"], raw[19], variant="tensor", label=size)


def test_tensor_variant_invariants_are_exact(evaluator):
    """Symmetric airfoil, mirror identity and batch-position independence hold bit for bit on the tensor path too."""
    import torch

    w = np.array([0.15, 0.20, 0.15, 0.20, 0.15, 0.20, 0.15, 0.15])
    sym = dict(upper_weights=w, lower_weights=-w, leading_edge_weight=0.0, TE_thickness=0.002)
    n = 200_000
    raw = synthetic.sample_training_distribution(n, seed=4).astype(np.float32)
    d_raw = torch.from_numpy(raw).cuda()
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(1)).cuda()
    for size in ("xxlarge", "xxxlarge"):
        aero = evaluator.get_aero_from_kulfan_parameters(sym, alpha=0.0, Re=1e6, model_size=size, variant="tensor")
        assert aero["CL"][0] == 0.0 and aero["CM"][0] == 0.0 and aero["Top_Xtr"][0] == aero["Bot_Xtr"][0]
        out = evaluator.evaluate_device(d_raw, size, variant="tensor")
        out_m = evaluator.evaluate_device(torch.from_numpy(mirrored(raw)).cuda(), size, variant="tensor")
        out_p = evaluator.evaluate_device(d_raw[:, perm].contiguous(), size, variant="tensor")
        torch.cuda.synchronize()
        assert torch.equal(out_p, out[:, perm]), size
        np.testing.assert_array_equal(out_m.cpu().numpy(), unmirror_outputs(out.cpu().numpy()))


def test_tensor_variant_refuses_what_it_cannot_run(evaluator):
    from neuralfoil_b200 import Evaluator

    raw = synthetic.sample_training_distribution(8, seed=0)
    with pytest.raises(ValueError, match="variant"):
        evaluator.evaluate_host(raw, "xxxlarge", variant="bf16")
    deep = Evaluator(device=0, extra_models={"odd_deep": _random_network([25] + [33] * 15 + [198], 5)})  # 16 affine layers
    try:
        with pytest.raises(ValueError, match="tensor-core variants"):
            deep.evaluate_host(raw, "odd_deep", variant="tensor")
        assert deep.evaluate_host(raw, "odd_deep").shape == (198, 8)  # the fp32 path takes it
    finally:
        deep.close()


# ---------------------------------------------------------------------------------------------
# Tensor-core variants of the NARROW models (nfb_tcn.cuh): 128-wide (large, xlarge = the reference's default model) and
# 64-wide (xxsmall .. medium), trained weights: same tolerances, same exact invariants as the wide models' kernels
# ---------------------------------------------------------------------------------------------
NARROW = ["xxsmall", "xsmall", "small", "medium", "large", "xlarge"]


@pytest.mark.parametrize("variant", ["tensor", "tensor_fp32"])
@pytest.mark.parametrize("size", NARROW)
def test_narrow_tensor_variants_match_oracle(evaluator, oracle, size, variant):
    raw = synthetic.sample_training_distribution(20011, seed=11)
    want = oracle.evaluate_raw(raw, size)
    got = evaluator.evaluate_host(raw, size, out_dtype=np.float64, variant=variant)
    parity.assert_parity(got, want, raw[19], label=f"{variant} {size}", variant=variant, narrow48=size in ("xxsmall", "xsmall"))
    ref = evaluator.evaluate_host(raw, size, out_dtype=np.float64)
    assert not np.array_equal(got, ref)  # really another kernel ...
    assert np.nanmax(np.abs(got[1] - ref[1])) <= 7.5e-2  # ... of the same network


@pytest.mark.parametrize("variant", ["tensor", "tensor_fp32"])
@pytest.mark.parametrize("n", [1, 3, 63, 64, 65, 127, 129, 1000, 9473])
@pytest.mark.parametrize("size", ["medium", "xlarge"])
def test_narrow_tensor_variants_ragged_batch_sizes(evaluator, oracle, size, n, variant):
    raw = synthetic.sample_training_distribution(n, seed=300 + n)
    got = evaluator.evaluate_host(raw, size, out_dtype=np.float64, variant=variant)
    assert got.shape == (198, n)
    parity.assert_parity(got, oracle.evaluate_raw(raw, size), raw[19], variant=variant, label=f"{variant} {size} n={n}")
    scal = evaluator.evaluate_host(raw, size, out_dtype=np.float64, variant=variant, outputs="scalars")
    assert np.array_equal(scal, got[:6], equal_nan=True)


@pytest.mark.parametrize("variant", ["tensor", "tensor_fp32"])
def test_narrow_tensor_variants_golden_far_ood_and_invariants(evaluator, golden_dir, variant):
    import torch

    g = np.load(golden_dir / "golden_all_sizes.npz")
    raw48 = g["raw"]
    w = np.array([0.15, 0.20, 0.15, 0.20, 0.15, 0.20, 0.15, 0.15])
    sym = dict(upper_weights=w, lower_weights=-w, leading_edge_weight=0.0, TE_thickness=0.002)
    n = 100_000
    raw = synthetic.sample_training_distribution(n, seed=4).astype(np.float32)
    d_raw = torch.from_numpy(raw).cuda()
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(1)).cuda()
    for size in ("medium", "large", "xlarge"):
        got = evaluator.evaluate_host(raw48, size, out_dtype=np.float64, variant=variant)  # incl. alpha = +-90, +-180, Re = 1e2, 1e10
        assert np.isfinite(got[:6]).all(), size
        parity.assert_parity(got, g[f"out_{size}"], raw48[19], variant=variant, label=f"{variant} {size} fixture")
        aero = evaluator.get_aero_from_kulfan_parameters(sym, alpha=0.0, Re=1e6, model_size=size, variant=variant)
        assert aero["CL"][0] == 0.0 and aero["CM"][0] == 0.0 and aero["Top_Xtr"][0] == aero["Bot_Xtr"][0]
        out = evaluator.evaluate_device(d_raw, size, variant=variant)
        out_m = evaluator.evaluate_device(torch.from_numpy(mirrored(raw)).cuda(), size, variant=variant)
        out_p = evaluator.evaluate_device(d_raw[:, perm].contiguous(), size, variant=variant)
        torch.cuda.synchronize()
        assert torch.equal(out_p, out[:, perm]), size
        np.testing.assert_array_equal(out_m.cpu().numpy(), unmirror_outputs(out.cpu().numpy()))
    bad = raw[:, :5000].copy()
    bad[18, 7] = np.nan
    out_bad = evaluator.evaluate_host(bad, "xlarge", variant=variant)
    out_ok = evaluator.evaluate_host(raw[:, :5000], "xlarge", variant=variant)
    assert np.isnan(out_bad[1, 7]) and np.array_equal(np.delete(out_bad, 7, axis=1), np.delete(out_ok, 7, axis=1))


# ---------------------------------------------------------------------------------------------
# Small-batch (latency) kernel
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", MODEL_SIZES)
def test_small_and_large_kernels_agree(evaluator, size):
    """A query gets bit-identical results from the 4-queries-per-CTA kernel (small batches) and the tiled kernel."""
    raw = synthetic.sample_training_distribution(6000, seed=17).astype(np.float32)
    big = evaluator.evaluate_host(raw, size, out_dtype=np.float32)           # 6000 > every small-batch limit
    for n in (1, 5, 37, 256):
        small = evaluator.evaluate_host(np.ascontiguousarray(raw[:, :n]), size, out_dtype=np.float32)
        np.testing.assert_array_equal(small, big[:, :n])


# ---------------------------------------------------------------------------------------------
# Geometry front doors (reference main.py:335-468) on the CUDA path
# ---------------------------------------------------------------------------------------------
def test_golden_airfoil_pipeline_on_gpu(evaluator):
    """reference tests/test_golden_values.py:48-55, 70-79 (rel 1e-6 there, for float64; fp32 bound of tests/parity.py here)"""
    from neuralfoil_b200 import Airfoil

    golden = {"analysis_confidence": 0.968356999039, "CL": 1.02508809111, "CD": 0.00791137686241,
              "CM": -0.0992982374935, "Top_Xtr": 0.402283708087, "Bot_Xtr": 1.0}
    aero = evaluator.get_aero_from_airfoil(Airfoil("naca4412"), alpha=5.0, Re=1e6, model_size="xlarge")
    assert list(aero) == output_keys()
    for key, want in golden.items():
        assert float(aero[key][0]) == pytest.approx(want, rel=1e-5, abs=6e-5), key


def test_entry_points_agree_on_gpu(evaluator, tmp_path):
    """reference tests/test_entry_points.py:22-48"""
    from neuralfoil_b200 import Airfoil

    af = Airfoil("naca4412")
    a = evaluator.get_aero_from_airfoil(af, alpha=4.0, Re=2e6)
    c = evaluator.get_aero_from_coordinates(af.coordinates, alpha=4.0, Re=2e6)
    dat = tmp_path / "naca4412.dat"
    dat.write_text(af.name + "\n" + "\n".join(f"{x:.16f} {y:.16f}" for x, y in af.coordinates))
    d = evaluator.get_aero_from_dat_file(dat, alpha=4.0, Re=2e6)
    for key in SCALARS:
        np.testing.assert_allclose(a[key], c[key], rtol=1e-10, err_msg=key)
        np.testing.assert_allclose(a[key], d[key], rtol=1e-5, atol=1e-6, err_msg=key)
    sweep = evaluator.get_aero_from_airfoil(af, alpha=np.linspace(-5, 10, 31), Re=2e6)
    assert sweep["CL"].shape == (31,) and np.all(np.diff(sweep["CL"][:20]) > 0)


# ---------------------------------------------------------------------------------------------
# Three-term tensor variant ("tensor_fp32"): the fp32 worst-case tolerance, its own distribution bound
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", ["xxlarge", "xxxlarge"])
def test_tensor_fp32_variant_matches_oracle(evaluator, oracle, size):
    raw = synthetic.sample_training_distribution(20011, seed=11)
    want = oracle.evaluate_raw(raw, size)
    got = evaluator.evaluate_host(raw, size, out_dtype=np.float64, variant="tensor_fp32")
    rep = parity.assert_parity(got, want, raw[19], label=f"tensor_fp32 {size}", variant="tensor_fp32")
    x1 = parity.error_report(evaluator.evaluate_host(raw, size, out_dtype=np.float64, variant="tensor"), want, raw[19])
    assert rep["median_rel"] < 0.1 * x1["median_rel"]  # two orders of magnitude tighter than the one-term variant


@pytest.mark.parametrize("n", [1, 3, 31, 32, 33, 63, 64, 65, 129, 1000])
@pytest.mark.parametrize("size", ["xxlarge", "xxxlarge"])
def test_tensor_fp32_variant_ragged_batch_sizes(evaluator, oracle, size, n):
    raw = synthetic.sample_training_distribution(n, seed=500 + n)
    got = evaluator.evaluate_host(raw, size, out_dtype=np.float64, variant="tensor_fp32")
    assert got.shape == (198, n)
    parity.assert_parity(got, oracle.evaluate_raw(raw, size), raw[19], variant="tensor_fp32")


def test_tensor_fp32_variant_golden_and_invariants(evaluator, golden_dir):
    import torch

    g = np.load(golden_dir / "golden_all_sizes.npz")
    raw = g["raw"]
    w = np.array([0.15, 0.20, 0.15, 0.20, 0.15, 0.20, 0.15, 0.15])
    sym = dict(upper_weights=w, lower_weights=-w, leading_edge_weight=0.0, TE_thickness=0.002)
    n = 100_000
    big = synthetic.sample_training_distribution(n, seed=6).astype(np.float32)
    d_big = torch.from_numpy(big).cuda()
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(2)).cuda()
    for size in ("xxlarge", "xxxlarge"):
        got = evaluator.evaluate_host(raw, size, out_dtype=np.float64, variant="tensor_fp32")
        parity.assert_parity(got, g[f"out_{size}"], raw[19], variant="tensor_fp32", label=size)
        aero = evaluator.get_aero_from_kulfan_parameters(sym, alpha=0.0, Re=1e6, model_size=size, variant="tensor_fp32")
        assert aero["CL"][0] == 0.0 and aero["CM"][0] == 0.0 and aero["Top_Xtr"][0] == aero["Bot_Xtr"][0]
        out = evaluator.evaluate_device(d_big, size, variant="tensor_fp32")
        out_m = evaluator.evaluate_device(torch.from_numpy(mirrored(big)).cuda(), size, variant="tensor_fp32")
        out_p = evaluator.evaluate_device(d_big[:, perm].contiguous(), size, variant="tensor_fp32")
        torch.cuda.synchronize()
        assert torch.equal(out_p, out[:, perm]), size
        np.testing.assert_array_equal(out_m.cpu().numpy(), unmirror_outputs(out.cpu().numpy()))


# ---------------------------------------------------------------------------------------------
# The tensor variants run on CTA pairs (nfb_tc2.cuh); the one-SM kernels (nfb_tc.cuh) are a second, independent
# implementation of the same pipeline, selected by the NFB_DEBUG_TC_ONE_SM flag of the C ABI (variant "...:one_sm")
# ---------------------------------------------------------------------------------------------
def test_one_sm_tensor_kernels_still_match_oracle(evaluator, oracle):
    for size, n in (("xxlarge", 4001), ("xxxlarge", 1003), ("xxxlarge", 65)):
        raw = synthetic.sample_training_distribution(n, seed=77 + n)
        want = oracle.evaluate_raw(raw, size)
        for variant in ("tensor", "tensor_fp32"):
            got = evaluator.evaluate_host(raw, size, out_dtype=np.float64, variant=variant + ":one_sm")
            parity.assert_parity(got, want, raw[19], check_distribution=n >= 1000, variant=variant, label=f"{variant}:one_sm {size} n={n}")


@pytest.mark.parametrize("variant", ["tensor", "tensor_fp32"])
def test_tensor_variants_fp32_output_unaligned_and_fp64_input(evaluator, oracle, variant):
    """Device API on the pair kernels: fp64 strided inputs, fp32 output whose leading dimension is not a multiple of 4."""
    import torch

    n = 1237
    raw = synthetic.sample_training_distribution(n, seed=91)
    d_raw = torch.from_numpy(raw).cuda()                     # float64 rows
    for size in ("xxlarge", "xxxlarge"):
        want = oracle.evaluate_raw(raw, size)
        out = torch.full((198, n + 1), float("nan"), dtype=torch.float32, device="cuda")[:, :n]   # ld = n + 1: odd
        res = evaluator.evaluate_device(d_raw, size, out=out, variant=variant)
        torch.cuda.synchronize()
        parity.assert_parity(res.cpu().numpy().astype(np.float64), want, raw[19], variant=variant, label=size)


# ---------------------------------------------------------------------------------------------
# Jacobian kernel: values bit-identical to the fp32 path, derivatives against the oracle's analytic Jacobian
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", MODEL_SIZES)
def test_jacobian_matches_oracle(evaluator, oracle, size):
    n = 400 if size in ("xxlarge", "xxxlarge") else 1000
    raw = synthetic.sample_training_distribution(n, seed=31)
    want_v, want_j = oracle.jacobian_raw(raw, size)
    got_v, got_j = evaluator.evaluate_jacobian_host(raw, size)
    assert got_v.shape == (198, n) and got_j.shape == (n, 198, 23)
    np.testing.assert_array_equal(got_v, evaluator.evaluate_host(raw, size, out_dtype=np.float64))  # same arithmetic as the fp32 kernels
    parity.assert_jacobian_parity(got_j, want_j, got_v, want_v, raw, label=size)


@pytest.mark.parametrize("n", [1, 3, 17])
def test_jacobian_small_batches_and_device_api(evaluator, oracle, n):
    import torch

    raw = synthetic.sample_training_distribution(n, seed=40 + n)
    want_v, want_j = oracle.jacobian_raw(raw, "xlarge")
    got_v, got_j = evaluator.evaluate_jacobian_host(raw, "xlarge", out_dtype=np.float32)
    assert got_v.dtype == np.float32 and got_j.dtype == np.float32
    parity.assert_jacobian_parity(got_j, want_j, got_v.astype(np.float64), want_v, raw)
    dv, dj = evaluator.evaluate_jacobian_device(torch.from_numpy(raw.astype(np.float32)).cuda(), "xlarge")
    torch.cuda.synchronize()
    assert dv.shape == (198, n) and dj.shape == (n, 198, 23) and dj.dtype == torch.float32
    # float32 inputs differ from the float64 ones by rounding only
    np.testing.assert_allclose(dj.cpu().numpy()[:, 1, 18], got_j[:, 1, 18], rtol=2e-3, atol=1e-6)  # dCL/dalpha


def test_jacobian_dict_api_and_symmetry(evaluator):
    """Reference-shaped front door; for a symmetric airfoil at alpha = 0 the odd outputs have even inputs' derivatives = 0."""
    import neuralfoil_b200 as nf

    w = np.array([0.15, 0.20, 0.15, 0.20, 0.15, 0.20, 0.15, 0.15])
    sym = dict(upper_weights=w, lower_weights=-w, leading_edge_weight=0.0, TE_thickness=0.002)
    aero, jac = evaluator.get_aero_and_jacobian_from_kulfan_parameters(sym, alpha=0.0, Re=1e6, model_size="large")
    ref = evaluator.get_aero_from_kulfan_parameters(sym, alpha=0.0, Re=1e6, model_size="large")
    assert list(aero) == list(ref) and all(np.array_equal(aero[k], ref[k]) for k in ref)
    assert set(jac) == set(ref) and jac["CL"].shape == (1, 23) and len(nf.JACOBIAN_INPUTS) == 23
    i_alpha, i_re, i_te = nf.JACOBIAN_INPUTS.index("alpha"), nf.JACOBIAN_INPUTS.index("Re"), nf.JACOBIAN_INPUTS.index("TE_thickness")
    assert jac["CL"][0, i_alpha] > 0.05                      # lift-curve slope, per degree
    assert jac["CL"][0, i_re] == 0.0 and jac["CL"][0, i_te] == 0.0 and jac["CM"][0, i_re] == 0.0   # exact by symmetry
    assert jac["CD"][0, i_alpha] == 0.0
    # mirrored pairs of Kulfan weights: dCL/d upper_i == dCL/d lower_i at a symmetric point (both raise camber)
    np.testing.assert_array_equal(jac["CL"][0, 0:8], jac["CL"][0, 8:16])
    # finite-difference sanity check of the slope through the value path (fp32 values: loose)
    a1 = evaluator.get_aero_from_kulfan_parameters(sym, alpha=0.5, Re=1e6, model_size="large")["CL"][0]
    a0 = evaluator.get_aero_from_kulfan_parameters(sym, alpha=-0.5, Re=1e6, model_size="large")["CL"][0]
    assert abs((a1 - a0) - jac["CL"][0, i_alpha]) < 2e-3
    with pytest.raises(ValueError):
        evaluator.evaluate_jacobian_host(np.zeros((22, 4)), "large")


# ---------------------------------------------------------------------------------------------
# outputs="scalars" (NFB_OUTPUT_SCALARS): the six scalar rows only, bit-identical to rows 0..5 of the full block
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 5, 700, 4099, 70001])
def test_scalars_only_equals_first_six_rows(evaluator, n):
    import torch

    raw = synthetic.sample_training_distribution(n, seed=600 + n)
    cases = [(size, "fp32") for size in ("xxsmall", "medium", "xlarge")] + [(s, v) for s in ("xxlarge", "xxxlarge") for v in ("fp32", "tensor", "tensor_fp32")]
    for size, variant in cases:
        full = evaluator.evaluate_host(raw, size, out_dtype=np.float64, variant=variant)
        six = evaluator.evaluate_host(raw, size, out_dtype=np.float64, variant=variant, outputs="scalars")
        assert six.shape == (6, n)
        np.testing.assert_array_equal(six, full[:6], err_msg=f"{size} {variant}")
    d = torch.from_numpy(raw.astype(np.float32)).cuda()
    out6 = evaluator.evaluate_device(d, "xxxlarge", variant="tensor_fp32", outputs="scalars")
    out198 = evaluator.evaluate_device(d, "xxxlarge", variant="tensor_fp32")
    torch.cuda.synchronize()
    assert out6.shape == (6, n) and torch.equal(out6, out198[:6])


def test_scalars_only_dict_api_and_errors(evaluator):
    aero = evaluator.get_aero_from_kulfan_parameters(synthetic.FIXED_KULFAN, alpha=np.linspace(-5, 10, 31), Re=1e6, model_size="large", outputs="scalars")
    ref = evaluator.get_aero_from_kulfan_parameters(synthetic.FIXED_KULFAN, alpha=np.linspace(-5, 10, 31), Re=1e6, model_size="large")
    assert list(aero) == ["analysis_confidence", "CL", "CD", "CM", "Top_Xtr", "Bot_Xtr"]
    for k in aero:
        np.testing.assert_array_equal(aero[k], ref[k])
    with pytest.raises(ValueError):
        evaluator.evaluate_host(np.zeros((23, 4)), "large", outputs="some")
    with pytest.raises(ValueError):
        evaluator.evaluate_host_into(np.zeros((23, 4)), "large", np.zeros((198, 4)), outputs="scalars")


@pytest.mark.parametrize("variant", ["tensor", "tensor_fp32"])
def test_nan_inputs_propagate_on_tensor_variants(evaluator, variant):
    """NaN input / Re <= 0 -> NaN outputs for that query only, also through the fp16-operand kernels,
    whose conversions saturate instead of overflowing: the neighbours in the same tile must be untouched."""
    n = 300
    raw = synthetic.sample_training_distribution(n, seed=77)
    clean = evaluator.evaluate_host(raw, "xxxlarge", out_dtype=np.float64, variant=variant)
    bad = raw.copy()
    bad[18, 5] = np.nan      # alpha
    bad[19, 70] = -1.0       # Re <= 0 -> log -> NaN
    bad[3, 131] = np.nan     # a Kulfan weight
    got = evaluator.evaluate_host(bad, "xxxlarge", out_dtype=np.float64, variant=variant)
    for q in (5, 70, 131):
        assert np.isnan(got[:4, q]).all() and np.isnan(got[6:, q]).all(), (variant, q)
    keep = np.ones(n, dtype=bool)
    keep[[5, 70, 131]] = False
    np.testing.assert_array_equal(got[:, keep], clean[:, keep])


def test_jacobian_nan_inputs_propagate(evaluator):
    raw = synthetic.sample_training_distribution(6, seed=78)
    raw[18, 2] = np.nan
    values, jac = evaluator.evaluate_jacobian_host(raw, "large")
    assert np.isnan(values[1, 2]) and np.isnan(jac[2, 1]).any() and np.isfinite(jac[[0, 1, 3, 4, 5]]).all()


@pytest.mark.parametrize("size", ["xxsmall", "xlarge", "xxxlarge"])
def test_vjp_matches_oracle_and_jacobian(evaluator, oracle, size):
    """grad = J^T cotangent from the VJP mode: against the oracle's analytic Jacobian, and against the library's own Jacobian."""
    n = 257
    raw = synthetic.sample_training_distribution(n, seed=55)
    rng = np.random.default_rng(5)
    cot = np.zeros((198, n))
    cot[1], cot[2], cot[3] = rng.normal(size=n), 50 * rng.normal(size=n), rng.normal(size=n)   # an objective of CL, CD, CM
    cot[6 + 64 : 6 + 96] = 0.01 * rng.normal(size=(32, n))                                    # and of the upper edge velocities
    values, grad = evaluator.evaluate_vjp_host(raw, cot, size)
    assert values.shape == (198, n) and grad.shape == (23, n)
    np.testing.assert_array_equal(values, evaluator.evaluate_host(raw, size, out_dtype=np.float64))
    got_v, got_j = evaluator.evaluate_jacobian_host(raw, size)
    mine = np.einsum("qrt,rq->tq", got_j, cot)
    scale = parity.JAC_INPUT_SCALE.copy()[:, None] * np.ones((1, n))
    scale[19] = raw[19]
    np.testing.assert_allclose(grad * scale, mine * scale, rtol=2e-4, atol=2e-5)   # same derivatives, summed in fp32 vs fp64
    want_v, want_j = oracle.jacobian_raw(raw, size)
    parity.assert_vjp_parity(grad, want_j, want_v, cot, raw, label=f"forward-mode kernel: {size}")


# ---------------------------------------------------------------------------------------------
# The fp32 path runs the column-group kernels (nfb_ffma_groups.cuh); the CTA-synchronous tiled kernels (nfb_ffma.cuh)
# are the earlier implementation of the same arithmetic, selected by the NFB_DEBUG_FFMA_TILED flag of the C ABI
# (variant "fp32:tiled").  Both must give the same bits: same bias + fma chain per output, same SiLU, same decode.
# ---------------------------------------------------------------------------------------------
_GROUP_CASES = [("xxsmall", 20_011), ("xsmall", 9_999), ("small", 20_000), ("medium", 33_333), ("large", 20_001),
                ("xlarge", 50_021), ("xxlarge", 20_003), ("xxxlarge", 9_001)]


@pytest.mark.parametrize("size,n", _GROUP_CASES)
def test_column_group_kernels_match_tiled_kernels_bit_for_bit(evaluator, size, n):
    import torch

    raw = torch.from_numpy(synthetic.sample_training_distribution(n, seed=500 + n, dtype=np.float32)).cuda()
    launches0 = evaluator.launch_count()
    groups = evaluator.evaluate_device(raw, size).cpu().numpy()
    tiled = evaluator.evaluate_device(raw, size, variant="fp32:tiled").cpu().numpy()
    assert evaluator.launch_count() == launches0 + 2
    assert np.array_equal(groups, tiled, equal_nan=True)
    for variant in ("fp32", "fp32:tiled"):
        scalars = evaluator.evaluate_device(raw, size, variant=variant, outputs="scalars").cpu().numpy()
        assert scalars.shape[0] == 6 and np.array_equal(scalars, groups[:6], equal_nan=True), variant
    small = evaluator.evaluate_device(raw[:, :37].contiguous(), size, variant="fp32:tiled").cpu().numpy()  # the flag also bypasses the small-batch kernel
    assert np.array_equal(small, groups[:, :37], equal_nan=True)


# ---------------------------------------------------------------------------------------------
# User-supplied networks: any widths up to 512, uneven from layer to layer, 2..16 affine layers.
# Exercises the zero padding of rows, the K trimming of the column-group kernels and every padded width.
# ---------------------------------------------------------------------------------------------
_ODD_ARCHITECTURES = {
    "odd40": [25, 40, 40, 198],                 # -> 64-wide kernel, K trimmed to 48
    "odd1layer": [25, 200, 198],                # one hidden layer -> 256-wide kernel
    "odd_uneven": [25, 100, 72, 100, 198],      # uneven widths -> 128-wide kernel, K trimmed to 112
    "odd300": [25, 300, 300, 300, 198],         # -> 512-wide kernel, K trimmed to 304
    "odd_deep": [25] + [33] * 15 + [198],       # 16 affine layers (NFB_MAX_LAYERS)
}


def _random_network(dims, seed):
    rng = np.random.default_rng(seed)
    params = {}
    for j in range(len(dims) - 1):
        gain = 1.1 if j == len(dims) - 2 else 1.7
        params[f"net.{2 * j}.weight"] = (rng.standard_normal((dims[j + 1], dims[j])) * gain / np.sqrt(dims[j])).astype(np.float32)
        params[f"net.{2 * j}.bias"] = (rng.standard_normal(dims[j + 1]) * 0.1).astype(np.float32)
    return params


def test_user_supplied_architectures_match_oracle():
    import torch
    from neuralfoil_b200 import Evaluator
    from oracle.neuralfoil_oracle import Oracle

    extra = {name: _random_network(dims, 900 + i) for i, (name, dims) in enumerate(_ODD_ARCHITECTURES.items())}
    ev, orc = Evaluator(device=0, extra_models=extra), Oracle(extra_models=extra)
    try:
        raw = synthetic.sample_training_distribution(9_001, seed=91)  # above every small-batch limit, ragged last tile
        for name in extra:
            want = orc.evaluate_raw(raw, name)
            got = ev.evaluate_host(raw, name, out_dtype=np.float64)
            parity.assert_parity(got, want, raw[19], label=name)
            small = ev.evaluate_host(np.ascontiguousarray(raw[:, :37]), name, out_dtype=np.float64)  # the 4-queries-per-CTA kernel
            np.testing.assert_array_equal(small, got[:, :37])
            sc = ev.evaluate_device(torch.from_numpy(raw).cuda(), name, outputs="scalars", out_dtype=torch.float64)
            torch.cuda.synchronize()
            np.testing.assert_array_equal(sc.cpu().numpy(), got[:6])
    finally:
        ev.close()


# ---------------------------------------------------------------------------------------------
# Reverse-mode gradient of the six scalar outputs (nfb_grad.cuh): against the oracle's analytic float64 Jacobian, against
# the forward-mode VJP kernel, values against the value kernel; ragged tiles, float32 / float64, broadcast-free device API.
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", list(MODEL_SIZES) + ["xxxlarge"])
def test_scalar_gradient_matches_oracle_and_forward_mode(evaluator, oracle, size):
    n = 1_501 if size != "xxxlarge" else 403
    raw = synthetic.sample_training_distribution(n, seed=58)
    rng = np.random.default_rng(6)
    cot = rng.normal(size=(6, n)) * np.array([1.0, 1.0, 50.0, 1.0, 1.0, 1.0])[:, None]
    values, grad = evaluator.evaluate_scalar_grad_host(raw, cot, size)
    assert values.shape == (6, n) and grad.shape == (23, n) and np.isfinite(grad).all()
    np.testing.assert_array_equal(values, evaluator.evaluate_host(raw, size, out_dtype=np.float64, outputs="scalars"))
    scale = parity.JAC_INPUT_SCALE.copy()[:, None] * np.ones((1, n))
    scale[19] = raw[19]
    want_v, want_j = oracle.jacobian_raw(raw, size)
    parity.assert_vjp_parity(grad, want_j[:6], want_v[:6], cot, raw, label=f"scalar-gradient kernel: {size}")
    cot198 = np.zeros((198, n))
    cot198[:6] = cot
    _, forward_mode = evaluator.evaluate_vjp_host(raw, cot198, size)
    np.testing.assert_allclose(grad * scale, forward_mode * scale, rtol=2e-3, atol=2e-4)


def test_scalar_gradient_float32_device_api_and_symmetry(evaluator):
    import torch

    n = 70_003  # many tiles per CTA on every width, ragged last tile
    raw = torch.from_numpy(synthetic.sample_training_distribution(n, seed=59).astype(np.float32)).cuda()
    cot = torch.zeros((6, n), dtype=torch.float32, device="cuda")
    cot[1] = 1.0  # f = CL
    for size in ("xxsmall", "xlarge", "xxlarge"):
        values, grad = evaluator.evaluate_scalar_grad_device(raw, cot, size)
        none, grad2 = evaluator.evaluate_scalar_grad_device(raw, cot, size, with_values=False)
        torch.cuda.synchronize()
        assert none is None and torch.equal(grad, grad2) and grad.dtype == torch.float32
        assert torch.equal(values, evaluator.evaluate_device(raw, size, outputs="scalars"))
        perm = torch.randperm(n, generator=torch.Generator().manual_seed(3)).cuda()
        _, grad_p = evaluator.evaluate_scalar_grad_device(raw[:, perm].contiguous(), cot, size, with_values=False)
        torch.cuda.synchronize()
        assert torch.equal(grad_p, grad[:, perm]), size  # position independence
    # a symmetric airfoil at alpha = 0: CL is odd in alpha, so d CL / d Re = d CL / d n_crit = 0 exactly, and
    # d CL / d upper_i = d CL / d lower_i (the twin's contribution mirrors the query's)
    w = np.array([0.15, 0.20, 0.15, 0.20, 0.15, 0.20, 0.15, 0.15])
    sym = np.zeros((23, 8), dtype=np.float64)
    sym[0:8], sym[8:16], sym[17], sym[19], sym[20], sym[21], sym[22] = w[:, None], -w[:, None], 0.002, 1e6, 9.0, 1.0, 1.0
    c = np.zeros((6, 8)); c[1] = 1.0
    _, g = evaluator.evaluate_scalar_grad_host(sym, c, "large")
    assert np.all(g[19] == 0.0) and np.all(g[20] == 0.0) and np.all(g[17] == 0.0)
    np.testing.assert_array_equal(g[0:8], g[8:16])


@pytest.mark.parametrize("size", list(MODEL_SIZES) + ["xxxlarge"])
def test_reverse_vjp_all_outputs_matches_forward_mode_and_oracle(evaluator, oracle, size):
    """nfb_eval_vjp_reverse_device (fused reverse pass, cotangents of all 198 outputs) against the forward-mode VJP kernel and
    against the oracle's analytic float64 Jacobian; the 198 values are the value kernel's bits."""
    n = 1_001 if size != "xxxlarge" else 403
    raw = synthetic.sample_training_distribution(n, seed=61)
    rng = np.random.default_rng(8)
    cot = np.zeros((198, n))
    cot[:6] = rng.normal(size=(6, n)) * np.array([1.0, 1.0, 50.0, 1.0, 1.0, 1.0])[:, None]
    cot[parity.ROW_THETA] = 10.0 * rng.normal(size=(64, n))
    cot[parity.ROW_H] = 0.1 * rng.normal(size=(64, n))
    cot[parity.ROW_UE] = 0.1 * rng.normal(size=(64, n))
    values, grad = evaluator.evaluate_vjp_host(raw, cot, size, reverse=True)
    assert values.shape == (198, n) and grad.shape == (23, n)
    np.testing.assert_array_equal(values, evaluator.evaluate_host(raw, size, out_dtype=np.float64))
    _, forward_mode = evaluator.evaluate_vjp_host(raw, cot, size, reverse=False)
    want_v, want_j = oracle.jacobian_raw(raw, size)
    # theta's derivative diverges where the edge velocity crosses zero: the elementwise bound is stated on the queries whose
    # 64 stations all have |ue| >= VJP_UE_WELL (tests/parity.py) ...
    well = np.abs(want_v[parity.ROW_UE]).min(axis=0) >= parity.VJP_UE_WELL
    assert well.mean() > (0.5 if size != "xxxlarge" else 0.05)  # (the synthetic xxxlarge's edge velocities cross zero often: 13 %)
    rep = parity.assert_vjp_parity(grad, want_j, want_v, cot, raw, label=f"reverse kernel, all outputs: {size}", well=well)
    # ... and on ALL queries the reverse pass may not do worse than the forward-mode kernel, the other fp32 evaluation of the
    # same derivative, and the two must agree with each other
    rep_fwd = parity.vjp_report(forward_mode, want_j, want_v, cot, raw)
    assert rep["ok"].mean() >= rep_fwd["ok"].mean() - 1e-3, (rep["ok"].mean(), rep_fwd["ok"].mean())
    between = np.abs(grad - forward_mode) * rep["scale"]
    assert (between <= 5e-3 * rep["ref"] + 1e-6).mean() > 0.998 and np.median(between / np.maximum(rep["ref"], 1e-30)) < 1e-5
    parity._record("vjp reverse kernel, all queries", {"frac_outside": float(1 - rep["ok"].mean()), "frac_outside_fwd": float(1 - rep_fwd["ok"].mean())},
                   ("frac_outside", "frac_outside_fwd"))


def test_reverse_vjp_device_api_ragged_and_position_independent(evaluator):
    import torch

    n = 40_013
    raw = torch.from_numpy(synthetic.sample_training_distribution(n, seed=62).astype(np.float32)).cuda()
    cot = torch.randn((198, n), dtype=torch.float32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(5)).cuda()
    for size in ("xsmall", "large", "xxlarge"):
        values, grad = evaluator.evaluate_vjp_device(raw, cot, size)   # n >= 256: the reverse kernel
        v_p, g_p = evaluator.evaluate_vjp_device(raw[:, perm].contiguous(), cot[:, perm].contiguous(), size, reverse=True)
        torch.cuda.synchronize()
        assert torch.equal(values, evaluator.evaluate_device(raw, size))
        assert torch.equal(v_p, values[:, perm]) and torch.equal(g_p, grad[:, perm]), size
        # cotangents of the six scalars only: the scalar-gradient kernel must agree to fp32 summation order
        cot6 = torch.zeros_like(cot)
        cot6[:6] = cot[:6]
        _, g_full = evaluator.evaluate_vjp_device(raw, cot6, size, reverse=True)
        _, g_scal = evaluator.evaluate_scalar_grad_device(raw, cot[:6].contiguous(), size, with_values=False)
        torch.cuda.synchronize()
        denom = g_scal.abs().amax(dim=1, keepdim=True) + 1e-30
        assert float(((g_full - g_scal).abs() / denom).max()) < 1e-3, size


def test_reverse_kernels_contain_nan_queries(evaluator):
    """A NaN / Re <= 0 query poisons its own gradient only: the columns of a tile are independent in the reverse pass too."""
    import torch

    n = 5_000
    raw = synthetic.sample_training_distribution(n, seed=63).astype(np.float32)
    bad = raw.copy()
    bad[18, 5] = np.nan      # alpha
    bad[19, 9] = -1.0        # Re <= 0 -> log(Re) = NaN in the reference, too
    bad[3, 4000] = np.nan    # a Kulfan weight
    poisoned = [5, 9, 4000]
    cot = torch.randn((198, n), dtype=torch.float32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(9))
    for size in ("medium", "xlarge"):
        _, g_ref = evaluator.evaluate_vjp_device(torch.from_numpy(raw).cuda(), cot, size, reverse=True)
        v_bad, g_bad = evaluator.evaluate_vjp_device(torch.from_numpy(bad).cuda(), cot, size, reverse=True)
        _, s_ref = evaluator.evaluate_scalar_grad_device(torch.from_numpy(raw).cuda(), cot[:6].contiguous(), size, with_values=False)
        _, s_bad = evaluator.evaluate_scalar_grad_device(torch.from_numpy(bad).cuda(), cot[:6].contiguous(), size, with_values=False)
        torch.cuda.synchronize()
        keep = torch.ones(n, dtype=torch.bool, device="cuda")
        keep[poisoned] = False
        assert torch.equal(g_bad[:, keep], g_ref[:, keep]) and torch.equal(s_bad[:, keep], s_ref[:, keep]), size
        assert torch.isfinite(g_ref).all() and torch.isfinite(s_ref).all()
        for q in poisoned:
            assert not torch.isfinite(g_bad[:, q]).all() and not torch.isfinite(s_bad[:, q]).all(), (size, q)
            assert torch.isnan(v_bad[1, q])


def test_pinned_host_blocks_default_chunk_is_invisible(evaluator):
    """Pinned caller memory takes the direct-copy path of nfb_eval_host, whose default chunk is a sixteenth of the batch
    (64 Ki .. 512 Ki queries): a shard of 1.25 M queries -- one of eight GPUs' share of BASELINE configs[2] -- runs as 16
    ragged chunks on two streams.  Same bits as one device-resident launch and as 512 Ki chunks."""
    import torch

    n = 1_250_003
    raw = torch.from_numpy(synthetic.sample_training_distribution(n, seed=77).astype(np.float32)).pin_memory()
    out_a = torch.empty((198, (n + 3) & ~3), dtype=torch.float32).pin_memory()
    out_b = torch.empty_like(out_a).pin_memory()
    dev = evaluator.evaluate_device(raw.cuda(), "small")
    torch.cuda.synchronize()
    want = dev.cpu().numpy()
    got = evaluator.evaluate_host_into(raw.numpy(), "small", out_a.numpy())
    np.testing.assert_array_equal(got, want)
    got_b = evaluator.evaluate_host_into(raw.numpy(), "small", out_b.numpy(), chunk=1 << 19)
    np.testing.assert_array_equal(got_b, want)
    six = torch.empty((6, (n + 3) & ~3), dtype=torch.float32).pin_memory()
    np.testing.assert_array_equal(evaluator.evaluate_host_into(raw.numpy(), "small", six.numpy(), outputs="scalars"), want[:6])


def test_result_pool_never_hands_out_memory_that_is_still_referenced(evaluator):
    """The drop-in call's large result blocks are recycled (main._ResultPool).  A dict that is still alive keeps its memory;
    memory comes back only after every array of a result is gone, and the values are the same either way."""
    import gc

    from neuralfoil_b200.main import _result_pool

    n = 50_000  # 198 x 50,000 float64 = 79 MB: pooled
    K = synthetic.FIXED_KULFAN
    alpha = np.linspace(-8.0, 12.0, n)
    _result_pool.clear()
    first = evaluator.get_aero_from_kulfan_parameters(K, alpha=alpha, Re=2e6, model_size="small")
    cl_first = first["CL"].copy()
    addr_first = first["CL"].__array_interface__["data"][0]
    second = evaluator.get_aero_from_kulfan_parameters(K, alpha=alpha[::-1].copy(), Re=2e6, model_size="small")
    assert second["CL"].__array_interface__["data"][0] != addr_first  # `first` is alive: fresh memory
    np.testing.assert_array_equal(first["CL"], cl_first)               # ... and untouched by the second call
    np.testing.assert_array_equal(second["CL"], cl_first[::-1])
    keep = first["CD"]  # one surviving view keeps the whole block
    del first
    gc.collect()
    third = evaluator.get_aero_from_kulfan_parameters(K, alpha=alpha, Re=2e6, model_size="small")
    assert third["CL"].__array_interface__["data"][0] != addr_first
    cd = keep.copy()
    del keep, third
    gc.collect()
    fourth = evaluator.get_aero_from_kulfan_parameters(K, alpha=alpha, Re=2e6, model_size="small")
    np.testing.assert_array_equal(fourth["CL"], cl_first)  # recycled memory, same values
    np.testing.assert_array_equal(fourth["CD"], cd)
    assert len(fourth) == 198 and all(v.shape == (n,) and v.dtype == np.float64 for v in fourth.values())
