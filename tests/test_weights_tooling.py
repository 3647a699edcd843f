"""
Weight-file tooling: .pth -> .npz conversion as the reference's
training/convert_all_pth_files_to_npz.py:16-19 does it, validation, installation of a user-supplied network.
"""

import numpy as np
import pytest

from neuralfoil_b200 import synthetic, weights
from neuralfoil_b200.main import nn_weights_dir


def test_shipped_files_validate():
    widths = {"xxsmall": 48, "xsmall": 48, "small": 64, "medium": 64, "large": 128, "xlarge": 128, "xxlarge": 256}
    for size, width in widths.items():
        with np.load(nn_weights_dir / f"nn-{size}.npz") as f:
            info = weights.describe({k: f[k] for k in f.files})
        assert info["dims"][0] == 25 and info["dims"][-1] == 198 and max(info["dims"][1:-1]) == width
        assert info["tensor_core_variant"] == (width > 128)


def test_pth_round_trip(tmp_path):
    """A checkpoint written the way training/train.py:324-330 writes it converts to the reference's .npz layout."""
    torch = pytest.importorskip("torch")
    layers = [torch.nn.Linear(25, 48), torch.nn.SiLU(), torch.nn.Linear(48, 48), torch.nn.SiLU(), torch.nn.Linear(48, 198)]

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.net = torch.nn.Sequential(*layers)

    model = Net()
    pth = tmp_path / "nn-tiny.pth"
    torch.save({"model_state_dict": model.state_dict(), "optimizer_state_dict": {}}, pth)
    npz = weights.convert_pth_to_npz(pth)
    with np.load(npz) as f:
        assert sorted(f.files) == sorted(f"net.{i}.{p}" for i in (0, 2, 4) for p in ("weight", "bias"))
        np.testing.assert_array_equal(f["net.2.weight"], model.net[2].weight.detach().numpy())
        assert f["net.4.bias"].dtype == np.float32
    (tmp_path / "installed").mkdir()
    dst = weights.install(npz, weights_dir=tmp_path / "installed")
    assert dst.name == "nn-tiny.npz" and dst.exists()


def test_validation_errors(tmp_path):
    good = synthetic.synthetic_xxxlarge_params()
    assert weights.describe(good)["n_parameters"] == 1_428_166 and weights.describe(good)["padded_width"] == 512
    with pytest.raises(ValueError, match="25 latent inputs"):
        weights.describe({"net.0.weight": np.zeros((198, 24), np.float32), "net.0.bias": np.zeros(198, np.float32),
                          "net.2.weight": np.zeros((198, 198), np.float32), "net.2.bias": np.zeros(198, np.float32)})
    wide = {"net.0.weight": np.zeros((1024, 25), np.float32), "net.0.bias": np.zeros(1024, np.float32),
            "net.2.weight": np.zeros((198, 1024), np.float32), "net.2.bias": np.zeros(198, np.float32)}
    with pytest.raises(ValueError, match="widest kernel"):
        weights.describe(wide)
    with pytest.raises(ValueError, match="unexpected neural network file format"):
        weights.describe({"weights": np.zeros(3)})
