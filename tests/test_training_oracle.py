"""
Pins of the training-step oracle (oracle/training_oracle.py) on the CPU:
  * torch float64 autograd of a line-by-line restatement of the reference's `Net.forward` and `loss_function`
    (/root/reference/training/train.py:56-114, 194-236; the module itself loads its training data at import time and
    cannot be imported), including NaN targets in every column -- loss components and every parameter gradient to 1e-10;
  * central finite differences of the oracle's own loss;
  * the loss weights of train.py:180-190.
"""

import numpy as np
import pytest

from neuralfoil_b200.main import default_loss_weights
from oracle import training_oracle as T
from oracle.neuralfoil_oracle import load_distribution


def make_case(dims, batch, seed, nan_fraction=0.2):
    rng = np.random.default_rng(seed)
    layers = [(rng.standard_normal((dims[i + 1], dims[i])) * (1.5 / np.sqrt(dims[i])), rng.standard_normal(dims[i + 1]) * 0.1)
              for i in range(len(dims) - 1)]
    dist = load_distribution()
    x = rng.standard_normal((batch, 25)) * 0.3 + dist["mean"]
    fused, _, _, _ = T.forward(layers, dist["mean"], dist["inv_cov"], x)
    y = fused + rng.standard_normal(fused.shape) * 0.04        # residuals on both sides of the Huber delta
    y[:, 0] = (rng.random(batch) < 0.7).astype(float)           # analysis_confidence is a binary label
    y[:, 4:6] = np.clip(y[:, 4:6] + 0.5, -0.2, 1.2)             # transition locations around and beyond the clip
    y[rng.random(y.shape) < nan_fraction] = np.nan
    return layers, dist, x, y


def torch_reference(layers, dist, x, y, loss_weights, delta):
    """train.py:56-114 and :194-236 in torch float64, statement by statement."""
    import torch

    tl = [(torch.tensor(w, requires_grad=True), torch.tensor(b, requires_grad=True)) for w, b in layers]
    mean, inv_cov = torch.tensor(dist["mean"]), torch.tensor(dist["inv_cov"])

    def net(a):
        for l, (w, b) in enumerate(tl):
            a = torch.nn.functional.linear(a, w, b)
            if l < len(tl) - 1:
                a = torch.nn.functional.silu(a)
        return a

    def d2(v):
        return torch.sum((v - mean) @ inv_cov * (v - mean), dim=1)

    xt = torch.tensor(x)
    yy = net(xt)
    yy[:, 0] = yy[:, 0] - d2(xt) / (2 * 25)
    xf = xt.clone()
    xf[:, :8] = -1 * xt[:, 8:16]
    xf[:, 8:16] = -1 * xt[:, :8]
    xf[:, 16] = -1 * xt[:, 16]
    xf[:, 18] = -1 * xt[:, 18]
    xf[:, 23] = xt[:, 24]
    xf[:, 24] = xt[:, 23]
    yf = net(xf)
    yf[:, 0] = yf[:, 0] - d2(xf) / (2 * 25)
    yu = yf.clone()
    yu[:, 1] = yf[:, 1] * -1
    yu[:, 3] = yf[:, 3] * -1
    yu[:, 4] = yf[:, 5]
    yu[:, 5] = yf[:, 4]
    yu[:, 6 + 32 * 0 : 6 + 32 * 2] = yf[:, 6 + 32 * 3 : 6 + 32 * 5]
    yu[:, 6 + 32 * 2 : 6 + 32 * 3] = yf[:, 6 + 32 * 5 : 6 + 32 * 6] * -1
    yu[:, 6 + 32 * 3 : 6 + 32 * 5] = yf[:, 6 + 32 * 0 : 6 + 32 * 2]
    yu[:, 6 + 32 * 5 : 6 + 32 * 6] = yf[:, 6 + 32 * 2 : 6 + 32 * 3] * -1
    fused = (yy + yu) / 2
    fused[:, 4] = torch.clip(fused[:, 4].clone(), 0, 1)
    fused[:, 5] = torch.clip(fused[:, 5].clone(), 0, 1)
    y_data = torch.tensor(y)
    y_data = torch.where(torch.isnan(y_data), fused, y_data)
    conf = torch.mean(torch.nn.functional.binary_cross_entropy_with_logits(input=fused[:, 0:1], target=y_data[:, 0:1], reduction="none"), dim=0)
    other = torch.mean(torch.nn.functional.huber_loss(fused[:, 1:], y_data[:, 1:], reduction="none", delta=delta), dim=0)
    weighted = torch.concatenate([conf, other], dim=0) * torch.tensor(loss_weights)
    torch.sum(weighted).backward()
    return weighted.detach().numpy(), [(w.grad.numpy(), b.grad.numpy()) for w, b in tl]


@pytest.mark.parametrize("dims,batch", [([25, 40, 40, 198], 37), ([25, 64, 198], 5), ([25, 48, 48, 48, 48, 198], 64)])
def test_oracle_equals_torch_autograd_of_the_reference_net(dims, batch):
    layers, dist, x, y = make_case(dims, batch, seed=len(dims) + batch)
    w = T.default_loss_weights()
    comps, grads = T.train_step(layers, dist["mean"], dist["inv_cov"], x, y, w, 0.05)
    want_c, want_g = torch_reference(layers, dist, x, y, w, 0.05)
    np.testing.assert_allclose(comps, want_c, rtol=1e-10, atol=1e-14)
    for (gw, gb), (tw, tb) in zip(grads, want_g):
        np.testing.assert_allclose(gw, tw, rtol=1e-9, atol=1e-12 * np.abs(tw).max())
        np.testing.assert_allclose(gb, tb, rtol=1e-9, atol=1e-12 * np.abs(tb).max())


def test_oracle_gradient_matches_finite_differences():
    layers, dist, x, y = make_case([25, 24, 24, 198], 16, seed=3, nan_fraction=0.0)  # (a NaN target makes the loss depend on the
    comps, grads = T.train_step(layers, dist["mean"], dist["inv_cov"], x, y)          #  prediction through the target too: see the torch pin)
    rng = np.random.default_rng(1)
    for l in range(len(layers)):
        for _ in range(6):
            i, j = rng.integers(layers[l][0].shape[0]), rng.integers(layers[l][0].shape[1])
            h = 1e-6
            vals = []
            for sgn in (+1, -1):
                pert = [(w.copy(), b.copy()) for w, b in layers]
                pert[l][0][i, j] += sgn * h
                vals.append(T.train_step(pert, dist["mean"], dist["inv_cov"], x, y)[0].sum())
            fd = (vals[0] - vals[1]) / (2 * h)
            # (the loss is only piecewise smooth -- Huber's delta, the Xtr clip -- so this is a coarse check; the torch pin is the exact one)
            assert abs(fd - grads[l][0][i, j]) <= 2e-4 * max(1.0, abs(fd)), (l, i, j, fd, grads[l][0][i, j])


def test_loss_weights_are_the_references():
    w = T.default_loss_weights()
    assert w.shape == (198,) and abs(w.sum() - 1000) < 1e-9
    assert w[2] / w[1] == pytest.approx(3) and w[0] / w[1] == pytest.approx(0.005) and w[100] / w[1] == pytest.approx(5e-3)
    np.testing.assert_array_equal(w, default_loss_weights())  # the product's copy
