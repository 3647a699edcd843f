"""`neuralfoil.main` of the reference (neuralfoil/main.py): the names its tests, studies and users import from it."""

from neuralfoil_b200.main import (  # noqa: F401
    bl_x_points,
    get_aero_from_airfoil,
    get_aero_from_coordinates,
    get_aero_from_dat_file,
    get_aero_from_kulfan_parameters,
    nn_weights_dir,
)
