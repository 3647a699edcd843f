"""
CPU oracle for the NeuralFoil learned-core path (TEST INFRASTRUCTURE, NOT PRODUCT).

This file restates, in plain NumPy float64, the algorithm of the reference's
`get_aero_from_kulfan_parameters` (/root/reference/neuralfoil/main.py:64-332).  It exists only
so that the CUDA path in `neuralfoil_b200/` has something to be checked against on a box where
`/root/reference` is not present.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product package never
does: `neuralfoil_b200` raises if its CUDA library is missing instead of falling back to this.

Parity pins (see tests/test_oracle_golden.py):
  * the reference's own known-answer vector for `medium`
    (reference tests/test_golden_values.py:22-37), rel 1e-6;
  * fixtures under tests/golden/*.npz that were produced by importing the UNMODIFIED reference
    `main.py` in the build container (tests/golden/make_golden.py), all 7 shipped sizes plus
    the seeded synthetic xxxlarge, all 198 outputs, rel 1e-11.
  * the real xxxlarge known-answer vector (tests/test_golden_values.py:38-45) is UNPINNED:
    nn-xxxlarge.npz is absent from the reference checkout (.MISSING_LARGE_BLOBS:1).

Layout convention used throughout this file: "SoA" = one row per quantity, one column per
query, i.e. arrays shaped (rows, N).  That is the reference's effective output layout too
(main.py:240 returns the transpose of a C-ordered (198, N) block).
"""

from __future__ import annotations

from pathlib import Path

import numpy as np

N_BL = 32  # stations per surface; reference neuralfoil/_basic_data_type.py:51
N_INPUTS = 25  # latent inputs; main.py:171-184
N_RAW = 23  # raw scalars per query: 8 + 8 + LE + TE + alpha + Re + n_crit + xtr_u + xtr_l
N_OUTPUTS = 6 + 6 * N_BL  # 198; _basic_data_type.py:386-401

MODEL_SIZES = ("xxsmall", "xsmall", "small", "medium", "large", "xlarge", "xxlarge", "xxxlarge")

# Reference _basic_data_type.py:23-38,52 -- midpoints of a uniform 33-point grid on [0, 1].
bl_x_points = (np.arange(N_BL) + 0.5) / N_BL

# main.py:15-16 -- the clip bound used by the reference's logistic function.
LN_EPS = float(np.log(10.0 / np.finfo(np.float64).max))  # -707.48...

DEFAULT_WEIGHTS_DIR = Path(__file__).resolve().parent.parent / "neuralfoil_b200" / "nn_weights_and_biases"


# ------------------------------------------------------------------------------------------
# Output naming (main.py:318-331)
# ------------------------------------------------------------------------------------------
def output_keys() -> list[str]:
    """The 198 dict keys in the reference's insertion order (main.py:318-331)."""
    keys = ["analysis_confidence", "CL", "CD", "CM", "Top_Xtr", "Bot_Xtr"]
    for side in ("upper", "lower"):
        for quantity in ("theta", "H", "ue/vinf"):
            keys += [f"{side}_bl_{quantity}_{i}" for i in range(N_BL)]
    return keys


# ------------------------------------------------------------------------------------------
# Model data (main.py:25-44)
# ------------------------------------------------------------------------------------------
def load_distribution(weights_dir: Path | str = DEFAULT_WEIGHTS_DIR) -> dict[str, np.ndarray]:
    """Mean / inverse covariance of the scaled inputs (main.py:27-32)."""
    with np.load(Path(weights_dir) / "scaled_input_distribution.npz") as f:
        return {
            "mean": np.asarray(f["mean_inputs_scaled"], dtype=np.float64),
            "cov": np.asarray(f["cov_inputs_scaled"], dtype=np.float64),
            "inv_cov": np.asarray(f["inv_cov_inputs_scaled"], dtype=np.float64),
        }


def layers_from_npz_dict(params: dict[str, np.ndarray]) -> list[tuple[np.ndarray, np.ndarray]]:
    """
    Ordered [(W, b), ...] from a `net.{i}.weight` / `net.{i}.bias` dictionary.
    Ordering rule and error behaviour follow main.py:206-215.
    """
    try:
        ids = sorted({int(name.split(".")[1]) for name in params})
    except (TypeError, ValueError, IndexError) as e:
        raise ValueError(
            "Got an unexpected neural network file format: keys must look like "
            f"'net.0.weight', 'net.0.bias', ...; got {list(params)}"
        ) from e
    return [(np.asarray(params[f"net.{i}.weight"]), np.asarray(params[f"net.{i}.bias"])) for i in ids]


def load_layers(model_size: str, weights_dir: Path | str = DEFAULT_WEIGHTS_DIR):
    with np.load(Path(weights_dir) / f"nn-{model_size}.npz") as f:
        return layers_from_npz_dict({k: f[k] for k in f.files})


# ------------------------------------------------------------------------------------------
# The path, step by step
# ------------------------------------------------------------------------------------------
def featurise(raw: np.ndarray) -> np.ndarray:
    """
    raw (23, N) float64 -> latent inputs (25, N) float64.  main.py:171-184.
    Raw row order: upper[0:8], lower[8:16], LE 16, TE 17, alpha 18, Re 19, n_crit 20,
    xtr_upper 21, xtr_lower 22.
    """
    raw = np.asarray(raw, dtype=np.float64)
    n = raw.shape[1]
    x = np.empty((N_INPUTS, n))
    x[0:17] = raw[0:17]
    x[17] = raw[17] * 50
    alpha = raw[18]
    # aerosandbox.numpy.sind / cosd: argument converted with x * pi / 180.
    x[18] = np.sin(2 * alpha * np.pi / 180)
    c = np.cos(alpha * np.pi / 180)
    x[19] = c
    x[20] = 1 - c**2
    x[21] = (np.log(raw[19]) - 12.5) / 3.5
    x[22] = (raw[20] - 9) / 4.5
    x[23] = raw[21]
    x[24] = raw[22]
    return x


def mlp(layers, x: np.ndarray) -> np.ndarray:
    """(25, N) -> (198, N).  Affine layers with SiLU between them; main.py:218-241."""
    h = x
    last = len(layers) - 1
    for j, (w, b) in enumerate(layers):
        h = w @ h + b[:, None]
        if j != last:
            h = h / (1 + np.exp(-h))  # aerosandbox.numpy.swish, beta = 1
    return h


def mlp_fp32(layers, x: np.ndarray) -> np.ndarray:
    """
    The same network evaluated entirely in float32 (inputs rounded to float32, float32 GEMM,
    float32 SiLU).  Not part of the reference: it is the "plain fp32 reference" that tells the tests
    how much of the CUDA path's deviation from float64 is inherent to fp32 arithmetic.
    """
    h = np.asarray(x, dtype=np.float32)
    last = len(layers) - 1
    for j, (w, b) in enumerate(layers):
        h = np.asarray(w, dtype=np.float32) @ h + np.asarray(b, dtype=np.float32)[:, None]
        if j != last:
            with np.errstate(over="ignore"):
                h = (h / (np.float32(1) + np.exp(-h))).astype(np.float32)
    return h.astype(np.float64)


def squared_mahalanobis(dist, x: np.ndarray) -> np.ndarray:
    """(25, N) -> (N,).  main.py:47-61."""
    d = x - dist["mean"][:, None]
    return np.einsum("in,ij,jn->n", d, dist["inv_cov"], d, optimize=True)


def flip_inputs(x: np.ndarray) -> np.ndarray:
    """Latent inputs of the airfoil mirrored about its chord at -alpha; main.py:253-265."""
    xf = x.copy()
    xf[0:8] = -x[8:16]
    xf[8:16] = -x[0:8]
    xf[16] = -x[16]
    xf[18] = -x[18]
    xf[23] = x[24]
    xf[24] = x[23]
    return xf


def unflip_outputs(f: np.ndarray) -> np.ndarray:
    """Map the mirrored airfoil's latent outputs back; main.py:274-289."""
    n = N_BL
    u = f.copy()
    u[1] = -f[1]
    u[3] = -f[3]
    u[4] = f[5]
    u[5] = f[4]
    u[6 : 6 + 2 * n] = f[6 + 3 * n : 6 + 5 * n]
    u[6 + 3 * n : 6 + 5 * n] = f[6 : 6 + 2 * n]
    u[6 + 2 * n : 6 + 3 * n] = -f[6 + 5 * n : 6 + 6 * n]
    u[6 + 5 * n : 6 + 6 * n] = -f[6 + 2 * n : 6 + 3 * n]
    return u


def latent_twins(layers, dist, x: np.ndarray, net=mlp) -> tuple[np.ndarray, np.ndarray]:
    """
    Both network evaluations with the Mahalanobis logit correction applied to each
    (main.py:243-246, 267-270).  Returns (y, y_flipped), each (198, N), latent space.
    """
    y = net(layers, x)
    y[0] = y[0] - squared_mahalanobis(dist, x) / (2 * N_INPUTS)
    xf = flip_inputs(x)
    yf = net(layers, xf)
    yf[0] = yf[0] - squared_mahalanobis(dist, xf) / (2 * N_INPUTS)
    return y, yf


def fuse(y: np.ndarray, yf: np.ndarray) -> np.ndarray:
    """Average the twins, squash confidence, clip Xtr; main.py:292-295, 19-22."""
    z = (y + unflip_outputs(yf)) / 2
    z[0] = 1 / (1 + np.exp(-np.clip(z[0], LN_EPS, -LN_EPS)))
    z[4] = np.clip(z[4], 0, 1)
    z[5] = np.clip(z[5], 0, 1)
    return z


def decode(z: np.ndarray, Re: np.ndarray) -> np.ndarray:
    """
    Fused latent (198, N) -> physical outputs (198, N) in the dict's key order.
    main.py:298-316.  Note the latent row order (theta, H, ue per side) already equals the
    dict's key order, so decode is row-wise.
    """
    n = N_BL
    out = np.empty_like(z)
    out[0] = z[0]
    out[1] = z[1] / 2
    out[2] = np.exp((z[2] - 2) * 2)
    out[3] = z[3] / 20
    out[4] = z[4]
    out[5] = z[5]
    for o in (6, 6 + 3 * n):
        ue = z[o + 2 * n : o + 3 * n]
        out[o : o + n] = (10 ** z[o : o + n] - 0.1) / (np.abs(ue) * Re[None, :])
        out[o + n : o + 2 * n] = 2.6 * np.exp(z[o + n : o + 2 * n])
        out[o + 2 * n : o + 3 * n] = ue
    return out


def evaluate_raw(layers, dist, raw: np.ndarray, return_latent: bool = False, fp32_mlp: bool = False):
    """
    raw (23, N) float64 -> (198, N) float64 physical outputs (and optionally the fused latent).
    `fp32_mlp=True` swaps in `mlp_fp32` (everything else stays float64): the fp32 yardstick.
    """
    raw = np.asarray(raw, dtype=np.float64)
    x = featurise(raw)
    y, yf = latent_twins(layers, dist, x, net=mlp_fp32 if fp32_mlp else mlp)
    z = fuse(y, yf)
    out = decode(z, raw[19])
    return (out, z) if return_latent else out


# ------------------------------------------------------------------------------------------
# Jacobian of the path.  The reference has no gradient code of its own: its users
# differentiate main.py by tracing it through CasADi (studies/sequential_multifidelity_optimization.py:21-56).
# This is the forward-mode derivative of the functions above, term by term, in float64; it is pinned
# by central differences of evaluate_raw (tests/test_oracle_golden.py::test_jacobian_matches_finite_differences).
# Tangent arrays carry the 23 raw inputs on axis 1: (rows, 23, N).
# ------------------------------------------------------------------------------------------
def featurise_with_tangents(raw: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    raw = np.asarray(raw, dtype=np.float64)
    n = raw.shape[1]
    x = featurise(raw)
    dx = np.zeros((N_INPUTS, N_RAW, n))
    for i in range(17):
        dx[i, i] = 1.0
    dx[17, 17] = 50.0
    k = np.pi / 180
    a = raw[18] * k
    dx[18, 18] = 2 * k * np.cos(2 * a)
    dx[19, 18] = -k * np.sin(a)
    dx[20, 18] = 2 * k * np.cos(a) * np.sin(a)
    dx[21, 19] = 1.0 / (3.5 * raw[19])
    dx[22, 20] = 1.0 / 4.5
    dx[23, 21] = 1.0
    dx[24, 22] = 1.0
    return x, dx


def mlp_with_tangents(layers, x: np.ndarray, dx: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    h, dh = x, dx
    last = len(layers) - 1
    for j, (w, b) in enumerate(layers):
        w = np.asarray(w, dtype=np.float64)
        h = w @ h + np.asarray(b, dtype=np.float64)[:, None]
        dh = np.einsum("oi,itn->otn", w, dh, optimize=True)
        if j != last:
            s = 1 / (1 + np.exp(-h))
            dh = (s * (1 + h * (1 - s)))[:, None, :] * dh  # d/dz [z sigmoid(z)]
            h = h * s
    return h, dh


def jacobian_raw(layers, dist, raw: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """raw (23, N) -> outputs (198, N) and d outputs / d raw inputs (198, 23, N), float64."""
    raw = np.asarray(raw, dtype=np.float64)
    n_bl = N_BL
    x, dx = featurise_with_tangents(raw)
    sym = dist["inv_cov"] + dist["inv_cov"].T

    def twin(xx, dxx):
        y, dy = mlp_with_tangents(layers, xx, dxx)
        d = xx - dist["mean"][:, None]
        y[0] = y[0] - squared_mahalanobis(dist, xx) / (2 * N_INPUTS)
        dy[0] = dy[0] - np.einsum("in,ij,jtn->tn", d, sym, dxx, optimize=True) / (2 * N_INPUTS)
        return y, dy

    y, dy = twin(x, dx)
    yf, dyf = twin(flip_inputs(x), flip_inputs(dx))  # the flip is a signed permutation of axis 0: linear
    z = (y + unflip_outputs(yf)) / 2
    dz = (dy + unflip_outputs(dyf)) / 2
    out = np.empty_like(z)
    jac = np.empty_like(dz)
    conf = 1 / (1 + np.exp(-np.clip(z[0], LN_EPS, -LN_EPS)))
    out[0], jac[0] = conf, (conf * (1 - conf))[None, :] * dz[0]
    out[1], jac[1] = z[1] / 2, dz[1] / 2
    out[2] = np.exp((z[2] - 2) * 2)
    jac[2] = 2 * out[2][None, :] * dz[2]
    out[3], jac[3] = z[3] / 20, dz[3] / 20
    for r in (4, 5):
        out[r] = np.clip(z[r], 0, 1)
        jac[r] = dz[r] * ((z[r] > 0) & (z[r] < 1))[None, :]
    re = raw[19]
    d_re = np.zeros((N_RAW, raw.shape[1]))
    d_re[19] = 1.0
    for o in (6, 6 + 3 * n_bl):
        zt, dzt = z[o : o + n_bl], dz[o : o + n_bl]
        ue, due = z[o + 2 * n_bl : o + 3 * n_bl], dz[o + 2 * n_bl : o + 3 * n_bl]
        p10 = 10**zt
        theta = (p10 - 0.1) / (np.abs(ue) * re[None, :])
        out[o : o + n_bl] = theta
        jac[o : o + n_bl] = (
            (np.log(10) * p10 / (np.abs(ue) * re[None, :]))[:, None, :] * dzt
            - (theta / ue)[:, None, :] * due  # d|ue| / |ue| = due / ue
            - theta[:, None, :] * (d_re / re[None, :])[None, :, :]
        )
        hh = 2.6 * np.exp(z[o + n_bl : o + 2 * n_bl])
        out[o + n_bl : o + 2 * n_bl] = hh
        jac[o + n_bl : o + 2 * n_bl] = hh[:, None, :] * dz[o + n_bl : o + 2 * n_bl]
        out[o + 2 * n_bl : o + 3 * n_bl] = ue
        jac[o + 2 * n_bl : o + 3 * n_bl] = due
    return out, jac


# ------------------------------------------------------------------------------------------
# Reference-shaped front door (validation + broadcast: main.py:154-199)
# ------------------------------------------------------------------------------------------
def _length(v) -> int:
    """aerosandbox.numpy.length: len() for sized things, 1 for scalars."""
    try:
        return len(v)
    except TypeError:
        return 1


def raw_rows_from_arguments(kulfan_parameters, alpha, Re, n_crit, xtr_upper, xtr_lower) -> np.ndarray:
    """
    Validation + broadcast of the user-facing arguments to a (23, N) float64 block.
    Error conditions and messages follow main.py:160-168 and :187-196.
    """
    for key in ("upper_weights", "lower_weights"):
        n_w = _length(kulfan_parameters[key])
        if n_w != 8:
            raise ValueError(
                f"NeuralFoil's neural networks expect exactly 8 CST weights per side, but "
                f"`kulfan_parameters[{key!r}]` has {n_w}."
            )
    rows = [kulfan_parameters["upper_weights"][i] for i in range(8)]
    rows += [kulfan_parameters["lower_weights"][i] for i in range(8)]
    rows += [
        kulfan_parameters["leading_edge_weight"],
        kulfan_parameters["TE_thickness"],
        alpha,
        Re,
        n_crit,
        xtr_upper,
        xtr_lower,
    ]
    n_cases = 1
    for r in rows:
        ln = _length(r)
        if ln > 1:
            if n_cases == 1:
                n_cases = ln
            elif ln != n_cases:
                raise ValueError(
                    "The inputs to the neural network must all have the same length. "
                    f"(Conflicting lengths: {n_cases} and {ln})"
                )
    return np.stack([np.ones(n_cases) * np.asarray(r, dtype=np.float64) for r in rows], axis=0)


class Oracle:
    """Holds the model data; `get_aero_from_kulfan_parameters` mirrors the reference signature."""

    def __init__(self, weights_dir: Path | str = DEFAULT_WEIGHTS_DIR, extra_models: dict | None = None):
        self.weights_dir = Path(weights_dir)
        self.dist = load_distribution(self.weights_dir)
        self.models: dict[str, list] = {}
        for path in sorted(self.weights_dir.glob("nn-*.npz")):  # main.py:35-44
            size = path.stem.removeprefix("nn-")
            self.models[size] = load_layers(size, self.weights_dir)
        for size, params in (extra_models or {}).items():
            self.models[size] = layers_from_npz_dict(params)

    def evaluate_raw(self, raw, model_size="xlarge", return_latent=False, fp32_mlp=False):
        return evaluate_raw(self.models[model_size], self.dist, raw, return_latent, fp32_mlp)

    def jacobian_raw(self, raw, model_size="xlarge"):
        return jacobian_raw(self.models[model_size], self.dist, raw)

    def get_aero_from_kulfan_parameters(
        self, kulfan_parameters, alpha, Re, n_crit=9.0, xtr_upper=1.0, xtr_lower=1.0, model_size="xlarge"
    ) -> dict[str, np.ndarray]:
        if model_size not in self.models:
            raise ValueError(f"Invalid {model_size=}. Must be one of {set(self.models)}.")
        raw = raw_rows_from_arguments(kulfan_parameters, alpha, Re, n_crit, xtr_upper, xtr_lower)
        out = self.evaluate_raw(raw, model_size)
        return {k: out[i] for i, k in enumerate(output_keys())}
