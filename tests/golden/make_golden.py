"""
Regenerates tests/golden/*.npz by running the UNMODIFIED reference implementation
(/root/reference/neuralfoil/main.py) in the build container.  Not run by the test-suite and not
runnable on the GPU box (/root/reference does not exist there); the fixtures it writes are
committed.

The reference imports `aerosandbox`, which is not installed here.  On the hot path it only uses
NumPy re-exports plus four helpers (main.py:161,176-178,239): `length`, `sind`, `cosd`, `swish`.
A throw-away stand-in package providing exactly those is created in a temp dir.
With it the reference reproduces its own golden numbers for `medium`
(tests/test_golden_values.py:29-37), which this script asserts before writing anything.

Usage:  python tests/golden/make_golden.py
"""

from __future__ import annotations

import sys
import tempfile
import textwrap
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
REFERENCE = Path("/root/reference")
OUT_DIR = Path(__file__).resolve().parent

sys.path.insert(0, str(REPO))
from neuralfoil_b200 import synthetic  # noqa: E402


def import_reference():
    shim_root = Path(tempfile.mkdtemp(prefix="asb_shim_"))
    pkg = shim_root / "aerosandbox"
    (pkg / "numpy").mkdir(parents=True)
    (pkg / "__init__.py").write_text(
        textwrap.dedent(
            """
            class Airfoil: ...
            class KulfanAirfoil: ...
            """
        )
    )
    (pkg / "numpy" / "__init__.py").write_text(
        textwrap.dedent(
            """
            from numpy import *
            import numpy as _np

            def length(x):
                try:
                    return len(x)
                except TypeError:
                    return 1

            def sind(x):
                return _np.sin(x * _np.pi / 180)

            def cosd(x):
                return _np.cos(x * _np.pi / 180)

            def swish(x, beta=1.0):
                return x / (1 + _np.exp(-beta * x))
            """
        )
    )
    sys.path.insert(0, str(shim_root))
    sys.path.insert(0, str(REFERENCE))
    import neuralfoil  # the reference package
    from neuralfoil import main as ref_main

    assert Path(neuralfoil.__file__).is_relative_to(REFERENCE)
    return neuralfoil, ref_main


def ref_eval_raw(nf, raw: np.ndarray, model_size: str) -> np.ndarray:
    kulfan, flow = synthetic.kulfan_dict_from_raw(raw)
    aero = nf.get_aero_from_kulfan_parameters(kulfan, model_size=model_size, **flow)
    return np.stack([np.asarray(v, dtype=np.float64) for v in aero.values()], axis=0), list(aero.keys())


def edge_case_columns() -> np.ndarray:
    """Far-out-of-distribution queries the reference tests exercise (tests/test_robustness.py:43-57)."""
    cols = []
    k = synthetic.FIXED_KULFAN
    base = np.concatenate(
        [k["upper_weights"], k["lower_weights"], [k["leading_edge_weight"], k["TE_thickness"]]]
    )
    for alpha in (-180.0, -90.0, 90.0, 180.0):
        for Re in (1e2, 1e10):
            cols.append(np.concatenate([base, [alpha, Re, 9.0, 1.0, 1.0]]))
    # symmetric airfoil at alpha = 0 (tests/test_invariants.py:35-52)
    w = np.array([0.15, 0.20, 0.15, 0.20, 0.15, 0.20, 0.15, 0.15])
    cols.append(np.concatenate([w, -w, [0.0, 0.002], [0.0, 1e6, 9.0, 1.0, 1.0]]))
    # the golden test point (tests/test_golden_values.py:58-67)
    cols.append(np.concatenate([base, [5.0, 1e6, 9.0, 1.0, 1.0]]))
    return np.stack(cols, axis=1)


def main():
    nf, ref_main = import_reference()

    # Gate: the shimmed reference reproduces its own known answers (tests/test_golden_values.py:29-37).
    golden_medium = {
        "analysis_confidence": 0.955711837783,
        "CL": 1.10332809679,
        "CD": 0.00919882438456,
        "CM": -0.110598030451,
        "Top_Xtr": 0.250543496781,
        "Bot_Xtr": 0.964878409058,
    }
    aero = nf.get_aero_from_kulfan_parameters(synthetic.FIXED_KULFAN, alpha=5.0, Re=1e6, model_size="medium")
    for key, want in golden_medium.items():
        got = float(aero[key][0])
        assert abs(got - want) <= 1e-6 * abs(want) + 1e-9, (key, got, want)
    print("reference reproduces its medium golden values")

    # Synthetic xxxlarge is injected at run time; the reference tree is read-only.
    if "xxxlarge" not in ref_main._allowable_model_sizes:
        ref_main._nn_parameters["xxxlarge"] = synthetic.synthetic_xxxlarge_params()
        ref_main._allowable_model_sizes.add("xxxlarge")
        xxxlarge_kind = "synthetic"
    else:
        xxxlarge_kind = "real"

    # 1. Random in-distribution queries + edge cases, every size, all 198 outputs.
    raw = np.concatenate([synthetic.sample_training_distribution(36, seed=1234), edge_case_columns()], axis=1)
    blob = {"raw": raw, "xxxlarge_kind": np.array(xxxlarge_kind)}
    keys = None
    for size in ("xxsmall", "xsmall", "small", "medium", "large", "xlarge", "xxlarge", "xxxlarge"):
        out, keys = ref_eval_raw(nf, raw, size)
        blob[f"out_{size}"] = out
        print(size, out.shape, "finite:", bool(np.isfinite(out).all()))
    blob["keys"] = np.array(keys)
    np.savez_compressed(OUT_DIR / "golden_all_sizes.npz", **blob)

    # 2. Config 1 of BASELINE.json: README example shape (alpha sweep, Re = 5e6, xlarge).
    alpha = np.linspace(-25, 25, 1000)
    aero = nf.get_aero_from_kulfan_parameters(synthetic.FIXED_KULFAN, alpha=alpha, Re=5e6, model_size="xlarge")
    full = np.stack([np.asarray(v) for v in aero.values()], axis=0)
    np.savez_compressed(
        OUT_DIR / "golden_config1_xlarge.npz",
        alpha=alpha,
        Re=np.array(5e6),
        scalars=full[:6],  # all 1000 points
        full_every_20=full[:, ::20],  # all 198 outputs at every 20th point
    )

    # 3. Mirror identity inputs of tests/test_invariants.py:55-90 (large), both sides.
    a5 = np.linspace(-8, 12, 5)
    kw = dict(Re=5e5, n_crit=8.0, model_size="large")
    k = synthetic.FIXED_KULFAN
    mirror = dict(
        upper_weights=-k["lower_weights"],
        lower_weights=-k["upper_weights"],
        leading_edge_weight=-k["leading_edge_weight"],
        TE_thickness=k["TE_thickness"],
    )
    a1 = nf.get_aero_from_kulfan_parameters(k, alpha=a5, xtr_upper=0.7, xtr_lower=0.4, **kw)
    a2 = nf.get_aero_from_kulfan_parameters(mirror, alpha=-a5, xtr_upper=0.4, xtr_lower=0.7, **kw)
    np.savez_compressed(
        OUT_DIR / "golden_mirror_large.npz",
        alpha=a5,
        out=np.stack([np.asarray(v) for v in a1.values()], axis=0),
        out_mirror=np.stack([np.asarray(v) for v in a2.values()], axis=0),
    )
    print("wrote fixtures to", OUT_DIR)


if __name__ == "__main__":
    main()
