"""
`neuralfoil` -- the reference's import name (reference neuralfoil/__init__.py:1-15), served by neuralfoil_b200.

Third-party callers import this name (AeroSandbox's `Airfoil.get_aero_from_neuralfoil` does `import neuralfoil as nf`):
with this repository on the path ahead of the reference package they reach the B200 evaluator without any edit.  Same five
public names, same signatures, same return values; `neuralfoil.main` exists too, because the reference's own tests and
studies import from it (`from neuralfoil.main import get_aero_from_kulfan_parameters`).
"""

from neuralfoil_b200 import (  # noqa: F401
    bl_x_points,
    get_aero_from_airfoil,
    get_aero_from_coordinates,
    get_aero_from_dat_file,
    get_aero_from_kulfan_parameters,
)
from neuralfoil_b200 import __version__ as _b200_version

__all__ = [
    "get_aero_from_kulfan_parameters",
    "get_aero_from_airfoil",
    "get_aero_from_coordinates",
    "get_aero_from_dat_file",
    "bl_x_points",
]

__version__ = f"0.3.3+b200.{_b200_version}"  # the reference release this surface mirrors + this package's version
