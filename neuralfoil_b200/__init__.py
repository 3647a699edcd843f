"""
neuralfoil_b200 -- B200-native (sm_100a) evaluation of NeuralFoil's learned core.

Drop-in surface (reference neuralfoil/__init__.py:1-15): `get_aero_from_kulfan_parameters` -- the hot path --
plus its three geometry front doors and `bl_x_points`.  Importing the package needs no GPU; the first evaluation creates the CUDA context
and raises if the library or a B200 is missing (there is no CPU fallback).
"""

from .geometry import Airfoil, KulfanAirfoil  # noqa: F401
from .main import (  # noqa: F401
    JACOBIAN_INPUTS,
    MODEL_SIZES,
    Evaluator,
    bf16_bits_to_float32,
    bl_x_points,
    default_evaluator,
    default_loss_weights,
    get_aero_from_airfoil,
    get_aero_from_coordinates,
    get_aero_from_dat_file,
    get_aero_and_jacobian_from_kulfan_parameters,
    get_aero_from_kulfan_parameters,
    output_keys,
)

__version__ = "0.5.0"  # == NFB_VERSION of include/neuralfoil_b200.h (tests/test_cabi_symbols.py)
