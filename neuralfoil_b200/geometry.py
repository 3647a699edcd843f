"""
Geometry side of the reference's higher-level entry points (reference neuralfoil/main.py:335-468):
coordinates -> normalised airfoil -> 8 + 8 + 1 + 1 Kulfan (CST) parameters.

In the reference this work is delegated to AeroSandbox (`Airfoil.normalize`, `Airfoil.to_kulfan_airfoil`,
`get_coordinates_from_raw_dat`, `Airfoil("naca4412")`), a third-party package that is neither vendored in the
reference repository nor installable here.  This module restates the published algorithms in NumPy so that
`get_aero_from_airfoil / _coordinates / _dat_file` exist without that dependency.  It runs once per airfoil on the
host; the per-query work stays in the CUDA path.

Pin: with these routines the NACA 4412 -> xlarge pipeline reproduces the reference's own golden numbers
(reference tests/test_golden_values.py:48-55) -- checked against the CPU oracle in tests/test_geometry.py.
"""

from __future__ import annotations

import re
from math import comb
from pathlib import Path

import numpy as np

N_WEIGHTS_PER_SIDE = 8  # what the networks were trained on (main.py:160-168)


def cosspace(lo: float, hi: float, n: int) -> np.ndarray:
    """Cosine-spaced points, denser at both ends (AeroSandbox `np.cosspace`)."""
    mean, amp = (hi + lo) / 2, (hi - lo) / 2
    s = mean + amp * np.cos(np.linspace(np.pi, 0, n))
    s[0], s[-1] = lo, hi
    return s


def naca4_coordinates(name: str, n_points_per_side: int = 200) -> np.ndarray:
    """Selig-ordered coordinates of a NACA 4-digit airfoil, blunt trailing edge (the classic 0.1015 coefficient)."""
    digits = re.sub(r"[^0-9]", "", name)
    if len(digits) != 4:
        raise ValueError(f"{name!r} is not a NACA 4-digit designation")
    m, p, t = int(digits[0]) / 100, int(digits[1]) / 10, int(digits[2:]) / 100
    x = cosspace(0, 1, n_points_per_side)
    yt = 5 * t * (0.2969 * x**0.5 - 0.1260 * x - 0.3516 * x**2 + 0.2843 * x**3 - 0.1015 * x**4)
    if p == 0:
        p = 0.5
    yc = np.where(x <= p, m / p**2 * (2 * p * x - x**2), m / (1 - p) ** 2 * ((1 - 2 * p) + 2 * p * x - x**2))
    dyc = np.where(x <= p, 2 * m / p**2 * (p - x), 2 * m / (1 - p) ** 2 * (p - x))
    th = np.arctan(dyc)
    xu, yu = x - yt * np.sin(th), yc + yt * np.cos(th)
    xl, yl = x + yt * np.sin(th), yc - yt * np.cos(th)
    return np.stack([np.concatenate([xu[::-1], xl[1:]]), np.concatenate([yu[::-1], yl[1:]])], axis=1)


def coordinates_from_raw_dat(raw_text) -> np.ndarray:
    """Parse the lines of a Selig-format .dat file: every line that is exactly two numbers is a point."""
    pts = []
    for line in raw_text:
        parts = re.split(r"[\s,]+", line.strip())
        if len(parts) != 2:
            continue
        try:
            pts.append((float(parts[0]), float(parts[1])))
        except ValueError:
            continue
    if len(pts) < 3:
        raise ValueError("no coordinates found in the .dat text")
    return np.array(pts, dtype=np.float64)


def le_index(coordinates: np.ndarray) -> int:
    """Index of the leading edge: the point farthest from the trailing edge (mid-point of first and last points)."""
    te = (coordinates[0] + coordinates[-1]) / 2
    return int(np.argmax(np.sum((coordinates - te) ** 2, axis=1)))


def normalize(coordinates: np.ndarray) -> dict:
    """
    Translate the leading edge to (0, 0), scale the chord to 1, rotate the trailing edge onto (1, 0).
    Returns the transformed coordinates and the transformation, named as the reference consumes them
    (main.py:360-373): x_translation, y_translation, scale_factor, rotation_angle [degrees].
    """
    c = np.asarray(coordinates, dtype=np.float64)
    if c.ndim != 2 or c.shape[1] != 2 or c.shape[0] < 3:
        raise ValueError("coordinates must have shape (N_points, 2)")
    i_le = le_index(c)
    x_translation, y_translation = -c[i_le, 0], -c[i_le, 1]
    c = c + np.array([x_translation, y_translation])
    te = (c[0] + c[-1]) / 2
    scale = 1.0 / float(np.hypot(te[0], te[1]))
    c = c * scale
    te = (c[0] + c[-1]) / 2
    angle = -float(np.arctan2(te[1], te[0]))
    ca, sa = np.cos(angle), np.sin(angle)
    c = c @ np.array([[ca, sa], [-sa, ca]])
    return {
        "coordinates": c,
        "x_translation": x_translation,
        "y_translation": y_translation,
        "scale_factor": scale,
        "rotation_angle": float(np.degrees(angle)),
    }


def kulfan_parameters_from_coordinates(coordinates: np.ndarray, n_weights_per_side: int = N_WEIGHTS_PER_SIDE,
                                       N1: float = 0.5, N2: float = 1.0) -> dict:
    """
    Least-squares CST (Kulfan) fit with the leading-edge-modification term, of already-normalised coordinates
    (Kulfan, "Universal parametric geometry representation method", J. Aircraft 45(1), 2008; LEM term:
    Kulfan, "Modification of CST airfoil representation methodology", 2020).

        y(x) = C(x) * sum_i w_i * B_i(x)  +  w_LE * x * (1 - x)^(n + 0.5)  +-  TE_thickness * x / 2
        C(x) = x^N1 (1 - x)^N2,  B_i = Bernstein polynomials of degree n - 1

    One linear system over both surfaces: unknowns = lower weights, upper weights, w_LE, TE thickness.
    A negative fitted TE thickness is re-fitted with the thickness pinned to zero.
    """
    c = np.asarray(coordinates, dtype=np.float64)
    x, y = c[:, 0], c[:, 1]
    i_le = le_index(c)
    is_upper = np.arange(len(x)) <= i_le
    n = n_weights_per_side
    # No clipping of x to [0, 1]: AeroSandbox does not clip, and the reference's golden numbers
    # (tests/test_golden_values.py:48-55) are only reproduced (to 2e-12) without it.  The maximum() only keeps a
    # slightly negative x from turning x**0.5 into NaN; it changes nothing for x >= 0.
    xc = x
    cls = np.maximum(xc, 0.0) ** N1 * (1 - xc) ** N2
    deg = n - 1
    k = np.array([comb(deg, i) for i in range(n)], dtype=np.float64)
    i_arr = np.arange(n)
    bern = k[None, :] * xc[:, None] ** i_arr[None, :] * (1 - xc[:, None]) ** (deg - i_arr[None, :])
    shape = cls[:, None] * bern
    le_col = xc * np.maximum(1 - xc, 0.0) ** (n + 0.5)
    te_col = np.where(is_upper, xc / 2, -xc / 2)

    def solve(with_te: bool):
        cols = [np.where(is_upper[:, None], 0.0, shape), np.where(is_upper[:, None], shape, 0.0), le_col[:, None]]
        if with_te:
            cols.append(te_col[:, None])
        sol, *_ = np.linalg.lstsq(np.concatenate(cols, axis=1), y, rcond=None)
        return sol

    sol = solve(True)
    te_thickness = float(sol[-1])
    if te_thickness < 0:
        sol = np.append(solve(False), 0.0)
        te_thickness = 0.0
    return {
        "lower_weights": sol[:n].copy(),
        "upper_weights": sol[n : 2 * n].copy(),
        "leading_edge_weight": float(sol[2 * n]),
        "TE_thickness": te_thickness,
    }


class Airfoil:
    """
    The little of `aerosandbox.Airfoil` the reference's entry points use: a name and Selig-ordered coordinates.
    `Airfoil("naca4412")` generates NACA 4-digit shapes; anything else needs explicit coordinates.
    """

    def __init__(self, name: str = "Untitled", coordinates=None):
        self.name = name
        if coordinates is None:
            if re.fullmatch(r"\s*naca\s*\d{4}\s*", name, flags=re.I):
                coordinates = naca4_coordinates(name)
            else:
                raise ValueError(
                    f"No coordinates given for airfoil {name!r}; only NACA 4-digit names can be generated here "
                    "(the UIUC database lookup of AeroSandbox is not part of this package)."
                )
        elif isinstance(coordinates, (str, Path)):
            coordinates = coordinates_from_raw_dat(Path(coordinates).read_text().splitlines())
        self.coordinates = np.asarray(coordinates, dtype=np.float64)
        if self.coordinates.ndim != 2 or self.coordinates.shape[1] != 2:
            raise ValueError("coordinates must have shape (N_points, 2)")


def coordinates_from_kulfan_parameters(kulfan_parameters: dict, n_points_per_side: int = 200,
                                       N1: float = 0.5, N2: float = 1.0) -> np.ndarray:
    """Selig-ordered coordinates of a CST airfoil: the forward map of `kulfan_parameters_from_coordinates`."""
    lower = np.asarray(kulfan_parameters["lower_weights"], dtype=np.float64)
    upper = np.asarray(kulfan_parameters["upper_weights"], dtype=np.float64)
    w_le = float(kulfan_parameters["leading_edge_weight"])
    te = float(kulfan_parameters["TE_thickness"])
    x = cosspace(0, 1, n_points_per_side)

    def side(w, sign):
        n = len(w)
        deg = n - 1
        k = np.array([comb(deg, i) for i in range(n)], dtype=np.float64)
        i_arr = np.arange(n)
        bern = k[None, :] * x[:, None] ** i_arr[None, :] * (1 - x[:, None]) ** (deg - i_arr[None, :])
        return x**N1 * (1 - x) ** N2 * (bern @ w) + w_le * x * (1 - x) ** (n + 0.5) + sign * te * x / 2

    yu, yl = side(upper, +1.0), side(lower, -1.0)
    return np.stack([np.concatenate([x[::-1], x[1:]]), np.concatenate([yu[::-1], yl[1:]])], axis=1)


class KulfanAirfoil:
    """
    The little of `aerosandbox.KulfanAirfoil` the reference's `get_aero_from_airfoil` uses (main.py:336,361-364): an
    airfoil defined by its CST parameters.  `coordinates` are generated on demand.
    """

    def __init__(self, name: str = "Untitled", lower_weights=None, upper_weights=None, leading_edge_weight: float = 0.0,
                 TE_thickness: float = 0.0):
        self.name = name
        self.lower_weights = np.asarray(lower_weights, dtype=np.float64)
        self.upper_weights = np.asarray(upper_weights, dtype=np.float64)
        self.leading_edge_weight = float(leading_edge_weight)
        self.TE_thickness = float(TE_thickness)

    @property
    def kulfan_parameters(self) -> dict:
        return {"lower_weights": self.lower_weights, "upper_weights": self.upper_weights,
                "leading_edge_weight": self.leading_edge_weight, "TE_thickness": self.TE_thickness}

    @property
    def coordinates(self) -> np.ndarray:
        return coordinates_from_kulfan_parameters(self.kulfan_parameters)


def kulfan_parameters_of(airfoil):
    """
    The CST parameters of a Kulfan-parameterised airfoil object if they can be used as they are -- a `kulfan_parameters`
    mapping with 8 weights per side -- else None (the caller then fits the object's coordinates).
    """
    params = getattr(airfoil, "kulfan_parameters", None)
    if params is None:
        return None
    try:
        ok = len(params["upper_weights"]) == N_WEIGHTS_PER_SIDE and len(params["lower_weights"]) == N_WEIGHTS_PER_SIDE
    except (KeyError, TypeError):
        ok = False
    if not ok:
        if not hasattr(airfoil, "coordinates"):
            raise ValueError(
                "NeuralFoil's neural networks expect exactly 8 CST weights per side; this airfoil has another resolution "
                "and no coordinates to refit."
            )
        return None
    return {k: params[k] for k in ("lower_weights", "upper_weights", "leading_edge_weight", "TE_thickness")}
