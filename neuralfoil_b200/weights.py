"""
Weight-file tooling either side of the path.

The reference keeps two on-disk forms of a network: the training checkpoint `nn-<size>.pth`
(`{"model_state_dict", "optimizer_state_dict"}`, reference training/train.py:324-330) and the inference file
`nn-<size>.npz` made from it by training/convert_all_pth_files_to_npz.py:16-19 -- `net.{2j}.weight` (out, in) and
`net.{2j}.bias` (out,) as float32.  This module converts the former to the latter, validates an `.npz` against what
the kernels support, and installs a user-supplied file (e.g. the real `nn-xxxlarge.npz`, absent from the reference
checkout) into the package's weights directory so that `model_size="xxxlarge"` becomes allowable.

    python -m neuralfoil_b200.weights convert  nn-xxxlarge.pth  nn-xxxlarge.npz
    python -m neuralfoil_b200.weights check    nn-xxxlarge.npz
    python -m neuralfoil_b200.weights install  nn-xxxlarge.npz
"""

from __future__ import annotations

import shutil
import sys
from pathlib import Path

import numpy as np

from .main import layers_from_params, nn_weights_dir

SUPPORTED_MAX_WIDTH = 512  # widest hidden layer the kernels are built for (include/neuralfoil_b200.h)
SUPPORTED_MAX_LAYERS = 16


def state_dict_to_params(state_dict) -> dict[str, np.ndarray]:
    """A torch `model_state_dict` (or any mapping of array-likes) -> the reference's `.npz` dictionary."""
    params = {}
    for key, value in state_dict.items():
        if hasattr(value, "detach"):
            value = value.detach().cpu().numpy()
        params[str(key)] = np.asarray(value, dtype=np.float32)
    return params


def convert_pth_to_npz(pth_file: str | Path, npz_file: str | Path | None = None) -> Path:
    """training/convert_all_pth_files_to_npz.py:16-19 for one file.  Needs torch only to unpickle the checkpoint."""
    import torch

    pth_file = Path(pth_file)
    npz_file = Path(npz_file) if npz_file is not None else pth_file.with_suffix(".npz")
    checkpoint = torch.load(pth_file, map_location="cpu", weights_only=True)
    state_dict = checkpoint["model_state_dict"] if "model_state_dict" in checkpoint else checkpoint
    params = state_dict_to_params(state_dict)
    describe(params)  # validates
    np.savez_compressed(npz_file, **params)
    return npz_file


def describe(params: dict) -> dict:
    """Validate a network against what the library can run; returns its dimensions.  Raises ValueError otherwise."""
    layers = layers_from_params(params)
    dims = [layers[0][0].shape[1]] + [w.shape[0] for w, _ in layers]
    if dims[0] != 25 or dims[-1] != 198:
        raise ValueError(f"The network must map 25 latent inputs to 198 outputs, got {dims[0]} -> {dims[-1]}.")
    if not 2 <= len(layers) <= SUPPORTED_MAX_LAYERS:
        raise ValueError(f"{len(layers)} affine layers; supported: 2..{SUPPORTED_MAX_LAYERS}.")
    if max(dims[1:-1]) > SUPPORTED_MAX_WIDTH:
        raise ValueError(f"Hidden width {max(dims[1:-1])} exceeds the widest kernel ({SUPPORTED_MAX_WIDTH}).")
    for w, b in layers:
        if not (np.isfinite(w).all() and np.isfinite(b).all()):
            raise ValueError("The network contains non-finite weights.")
    width = max(dims[1:-1])
    return {
        "dims": dims,
        "n_parameters": int(sum(w.size + b.size for w, b in layers)),
        "padded_width": 64 if width <= 64 else 128 if width <= 128 else 256 if width <= 256 else 512,
        "tensor_core_variant": width > 128 and len(layers) <= 7,
    }


def install(npz_file: str | Path, model_size: str | None = None, weights_dir: str | Path | None = None) -> Path:
    """Copy a validated `.npz` into the weights directory as `nn-<model_size>.npz` (name taken from the file by default)."""
    npz_file = Path(npz_file)
    with np.load(npz_file) as f:
        describe({k: f[k] for k in f.files})
    if model_size is None:
        model_size = npz_file.stem.removeprefix("nn-")
    dst = Path(weights_dir or nn_weights_dir) / f"nn-{model_size}.npz"
    shutil.copyfile(npz_file, dst)
    return dst


def _main(argv: list[str]) -> int:
    if len(argv) >= 2 and argv[0] == "convert":
        print(convert_pth_to_npz(argv[1], argv[2] if len(argv) > 2 else None))
    elif len(argv) == 2 and argv[0] == "check":
        with np.load(argv[1]) as f:
            print(describe({k: f[k] for k in f.files}))
    elif len(argv) >= 2 and argv[0] == "install":
        print(install(argv[1], argv[2] if len(argv) > 2 else None))
    else:
        print(__doc__)
        return 2
    return 0


if __name__ == "__main__":
    raise SystemExit(_main(sys.argv[1:]))
