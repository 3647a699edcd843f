"""
CPU oracle for one training step of the learned core (TEST INFRASTRUCTURE, NOT PRODUCT) -- SURVEY 8(f-4).

NumPy float64 restatement of what the reference's training loop computes between `net(x)` and `loss.backward()`:
  * `Net.forward`                 /root/reference/training/train.py:56-114
  * `loss_function`, loss_weights /root/reference/training/train.py:180-236
and the analytic gradient of that loss with respect to every `net.{2l}.weight` / `net.{2l}.bias`.

Pins (tests/test_training_oracle.py): torch float64 autograd of a line-by-line restatement of the reference's `Net` and
`loss_function` (the reference module itself cannot be imported: it loads its training data at import time,
train.py:2-9), to 1e-10; and central finite differences of this file's own loss.  Only tests/, smoke() and bench.py's CPU
legs may import this file.
"""

from __future__ import annotations

import numpy as np

N_INPUTS, N_OUTPUTS, N_BL = 25, 198, 32


def default_loss_weights() -> np.ndarray:
    """train.py:180-190"""
    w = np.ones(N_OUTPUTS)
    w[0] *= 0.005  # analysis confidence
    w[1] *= 1      # CL
    w[2] *= 3      # ln(CD)
    w[3] *= 0.25   # CM
    w[4] *= 0.25   # Top Xtr
    w[5] *= 0.25   # Bot Xtr
    w[6:] *= 5e-3  # boundary-layer outputs
    return w / w.sum() * 1000


def flip_inputs(x: np.ndarray) -> np.ndarray:
    """train.py:66-79 (the same signed permutation as main.py:253-265)"""
    f = x.copy()
    f[:, :8] = -x[:, 8:16]
    f[:, 8:16] = -x[:, :8]
    f[:, 16] = -x[:, 16]
    f[:, 18] = -x[:, 18]
    f[:, 23] = x[:, 24]
    f[:, 24] = x[:, 23]
    return f


def unflip_outputs(y: np.ndarray) -> np.ndarray:
    """train.py:87-104.  The map is a signed permutation and its own inverse, hence also its own transpose."""
    u = y.copy()
    u[:, 1] = -y[:, 1]
    u[:, 3] = -y[:, 3]
    u[:, 4] = y[:, 5]
    u[:, 5] = y[:, 4]
    n = N_BL
    u[:, 6 : 6 + 2 * n] = y[:, 6 + 3 * n : 6 + 5 * n]
    u[:, 6 + 2 * n : 6 + 3 * n] = -y[:, 6 + 5 * n : 6 + 6 * n]
    u[:, 6 + 3 * n : 6 + 5 * n] = y[:, 6 : 6 + 2 * n]
    u[:, 6 + 5 * n : 6 + 6 * n] = -y[:, 6 + 2 * n : 6 + 3 * n]
    return u


def _sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def _net(layers, x):
    """train.py:33-46: Linear, SiLU, ..., Linear.  Returns the output and what the backward pass needs."""
    a, inputs, pre = x, [], []
    for l, (w, b) in enumerate(layers):
        inputs.append(a)
        z = a @ w.T + b
        pre.append(z)
        a = z * _sigmoid(z) if l < len(layers) - 1 else z
    return a, inputs, pre


def _maha(x, mean, inv_cov):
    d = x - mean
    return np.sum(d @ inv_cov * d, axis=1)  # train.py:48-54


def forward(layers, mean, inv_cov, x):
    """y_fused of train.py:56-114 (latent space: no sigmoid, no decode) plus the intermediates of both passes."""
    x = np.asarray(x, dtype=np.float64)
    xf = flip_inputs(x)
    y, in_a, pre_a = _net(layers, x)
    y = y.copy()
    y[:, 0] -= _maha(x, mean, inv_cov) / (2 * N_INPUTS)
    yf, in_b, pre_b = _net(layers, xf)
    yf = yf.copy()
    yf[:, 0] -= _maha(xf, mean, inv_cov) / (2 * N_INPUTS)
    fused = (y + unflip_outputs(yf)) / 2
    unclipped = fused[:, 4:6].copy()
    fused[:, 4:6] = np.clip(unclipped, 0, 1)
    return fused, unclipped, (in_a, pre_a), (in_b, pre_b)


def loss_terms(pred, target, delta):
    """Element-wise loss and d loss / d pred of train.py:194-216, NaN targets replaced by the (attached) prediction."""
    nan = np.isnan(target)
    t = np.where(nan, pred, target)
    loss, dl = np.empty_like(pred), np.empty_like(pred)
    p0, t0 = pred[:, 0], t[:, 0]
    loss[:, 0] = np.maximum(p0, 0) - p0 * t0 + np.log1p(np.exp(-np.abs(p0)))  # BCE with logits
    dl[:, 0] = _sigmoid(p0) - t0
    dl[:, 0] = np.where(nan[:, 0], dl[:, 0] - p0, dl[:, 0])  # the target path of torch.where(isnan, y_pred, y_data): d BCE / d t = -p
    d = pred[:, 1:] - t[:, 1:]
    ad = np.abs(d)
    loss[:, 1:] = np.where(ad < delta, 0.5 * d * d, delta * (ad - 0.5 * delta))  # Huber
    dl[:, 1:] = np.clip(d, -delta, delta)
    return loss, dl


def train_step(layers, mean, inv_cov, x, y_data, loss_weights=None, delta: float = 0.05):
    """
    -> (loss_components (198,), [(dW_l, db_l), ...]): the weighted mean loss of every output (their sum is the loss the
    reference back-propagates) and its gradient with respect to every layer's parameters.
    """
    layers = [(np.asarray(w, dtype=np.float64), np.asarray(b, dtype=np.float64)) for w, b in layers]
    w_loss = default_loss_weights() if loss_weights is None else np.asarray(loss_weights, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    y_data = np.asarray(y_data, dtype=np.float64)
    batch = x.shape[0]
    fused, unclipped, pass_a, pass_b = forward(layers, mean, inv_cov, x)
    loss, dl = loss_terms(fused, y_data, delta)
    components = loss.mean(axis=0) * w_loss
    g = dl * (w_loss / batch)[None, :]
    g[:, 4:6] *= (unclipped >= 0) & (unclipped <= 1)  # torch.clip passes the gradient on the closed interval
    grads = [(np.zeros_like(w), np.zeros_like(b)) for w, b in layers]
    for (inputs, pre), gy in ((pass_a, g / 2), (pass_b, unflip_outputs(g / 2))):
        dz = gy
        for l in range(len(layers) - 1, -1, -1):
            grads[l][0][...] += dz.T @ inputs[l]
            grads[l][1][...] += dz.sum(axis=0)
            if l > 0:
                z = pre[l - 1]
                s = _sigmoid(z)
                dz = (dz @ layers[l][0]) * (s * (1 + z * (1 - s)))
    return components, grads
