"""
Seeded synthetic data for the path: stand-in xxxlarge weights and the benchmark query samplers.

* `synthetic_xxxlarge_params()`: the reference checkout ships no `nn-xxxlarge.npz`
  (/root/reference/.MISSING_LARGE_BLOBS:1).  Its architecture is known
  (25 -> 512, 5 x (512 -> 512), 512 -> 198; reference training/train.py:16-18,32-46), so for
  benchmarking and parity we generate random weights of that shape with the per-layer
  statistics of the shipped xxlarge model (recipe: SURVEY.md section 8(d)).  Every report that
  uses them says "synthetic".  A real file dropped into the weights directory takes precedence.
* `sample_training_distribution()`: Config 3 of BASELINE.json -- random Kulfan airfoils from the
  Gaussian fitted to the training inputs plus the flow-condition sampler of
  reference training/training_data/generate_xfoil_data.py:142-157.
* `polar_grid()`: Config 2 -- one airfoil, alpha x Re x n_crit grid.

All arrays are "raw rows": (23, N), row order upper[0:8], lower[8:16], LE, TE, alpha, Re,
n_crit, xtr_upper, xtr_lower.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np

SYNTHETIC_XXXLARGE_SEED = 20261019
XXXLARGE_DIMS = (25, 512, 512, 512, 512, 512, 512, 198)
_GAINS = (1.76, 1.72, 1.72, 1.72, 1.72, 1.72, 1.14)
_BIAS_STD = (0.31, 0.075, 0.075, 0.075, 0.075, 0.075, 0.063)

_WEIGHTS_DIR = Path(__file__).resolve().parent / "nn_weights_and_biases"

# The fixed test airfoil of the reference's golden tests (tests/test_golden_values.py:22-27).
FIXED_KULFAN = dict(
    upper_weights=np.array([0.20, 0.25, 0.20, 0.27, 0.20, 0.25, 0.20, 0.20]),
    lower_weights=np.array([-0.10, -0.05, -0.10, 0.00, 0.05, 0.05, 0.10, 0.10]),
    leading_edge_weight=0.15,
    TE_thickness=0.002,
)


def synthetic_xxxlarge_params(seed: int = SYNTHETIC_XXXLARGE_SEED) -> dict[str, np.ndarray]:
    """`net.{2j}.weight/bias` float32 dict with the xxxlarge shape; deterministic in `seed`."""
    rng = np.random.default_rng(seed)
    params: dict[str, np.ndarray] = {}
    for j in range(len(XXXLARGE_DIMS) - 1):
        fan_in, fan_out = XXXLARGE_DIMS[j], XXXLARGE_DIMS[j + 1]
        w = rng.standard_normal((fan_out, fan_in)) * (_GAINS[j] / np.sqrt(fan_in))
        b = rng.standard_normal(fan_out) * _BIAS_STD[j]
        params[f"net.{2 * j}.weight"] = w.astype(np.float32)
        params[f"net.{2 * j}.bias"] = b.astype(np.float32)
    return params


def sample_training_distribution(n: int, seed: int = 0, dtype=np.float64) -> np.ndarray:
    """Config 3: (23, n) raw rows sampled over the training distribution."""
    rng = np.random.default_rng(seed)
    with np.load(_WEIGHTS_DIR / "scaled_input_distribution.npz") as f:
        mean = np.asarray(f["mean_inputs_scaled"], dtype=np.float64)[:18]
        cov = np.asarray(f["cov_inputs_scaled"], dtype=np.float64)[:18, :18]
    chol = np.linalg.cholesky(cov)
    raw = np.empty((23, n), dtype=np.float64)
    # Generate in chunks so that 10M-query batches do not need a (18, n) float64 temporary twice.
    step = 1 << 20
    for s in range(0, n, step):
        e = min(n, s + step)
        k = mean[:, None] + chol @ rng.standard_normal((18, e - s))
        raw[0:17, s:e] = k[0:17]
        raw[17, s:e] = np.maximum(k[17], 0.0) / 50.0
        raw[18, s:e] = (
            rng.choice(np.linspace(-15, 15, 7), size=e - s)
            + rng.uniform(-2.5, 2.5, size=e - s)
            + 2.5 * rng.standard_normal(e - s)
        )
        raw[19, s:e] = 10.0 ** (5.5 + 1.5 * rng.standard_normal(e - s))
        raw[20, s:e] = rng.uniform(0, 18, size=e - s)
        for row in (21, 22):
            free = rng.random(e - s) < 0.8
            raw[row, s:e] = np.where(free, 1.0, rng.uniform(0, 1, size=e - s))
    return raw.astype(dtype, copy=False)


def polar_grid(n_alpha: int = 100, n_re: int = 100, n_ncrit: int = 100, dtype=np.float64) -> np.ndarray:
    """Config 2: the fixed test airfoil over alpha x Re x n_crit; (23, n_alpha*n_re*n_ncrit)."""
    a, r, c = np.meshgrid(
        np.linspace(-20, 20, n_alpha), np.logspace(4, 8, n_re), np.linspace(1, 17, n_ncrit), indexing="ij"
    )
    n = a.size
    raw = np.empty((23, n), dtype=np.float64)
    raw[0:8] = FIXED_KULFAN["upper_weights"][:, None]
    raw[8:16] = FIXED_KULFAN["lower_weights"][:, None]
    raw[16] = FIXED_KULFAN["leading_edge_weight"]
    raw[17] = FIXED_KULFAN["TE_thickness"]
    raw[18] = a.ravel()
    raw[19] = r.ravel()
    raw[20] = c.ravel()
    raw[21] = 1.0
    raw[22] = 1.0
    return raw.astype(dtype, copy=False)


def kulfan_dict_from_raw(raw: np.ndarray) -> tuple[dict, dict]:
    """Split raw rows into the reference's `kulfan_parameters` dict ((8, N) weights) and flow kwargs."""
    kulfan = dict(
        upper_weights=raw[0:8],
        lower_weights=raw[8:16],
        leading_edge_weight=raw[16],
        TE_thickness=raw[17],
    )
    flow = dict(alpha=raw[18], Re=raw[19], n_crit=raw[20], xtr_upper=raw[21], xtr_lower=raw[22])
    return kulfan, flow
