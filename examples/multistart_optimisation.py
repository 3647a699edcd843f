"""
Multi-start shape optimisation on the reverse-mode kernel: 100,000 gradient-ascent runs at once.

Every run maximises CL / CD of its own airfoil (17 Kulfan parameters) at its own operating point (alpha, Re), keeping the
network's analysis confidence above a floor.  One iteration of ALL runs is one call of `evaluate_scalar_grad_device`
(nfb_grad.cuh: a fused forward + reverse pass; values and d objective / d 23 inputs per query), everything stays on the GPU.
The reference's users get such gradients by tracing main.py through CasADi, one design at a time
(studies/sequential_multifidelity_optimization.py:21-56).  Needs a B200.

    python examples/multistart_optimisation.py [model_size] [runs]
"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from neuralfoil_b200 import Evaluator, synthetic  # noqa: E402

model_size = sys.argv[1] if len(sys.argv) > 1 else "xlarge"
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
ev = Evaluator(device=0)
raw = torch.from_numpy(synthetic.sample_training_distribution(runs, seed=7, dtype=np.float32)).cuda()  # starting designs
raw[18].clamp_(1.0, 6.0)        # alpha [deg]
raw[19].clamp_(3e5, 1e7)        # Re
raw[20], raw[21], raw[22] = 9.0, 1.0, 1.0
step = torch.full((runs,), 2e-5, device="cuda")


def objective(r):
    """f = CL / CD - 2000 max(0, 0.9 - confidence)^2 and its gradient w.r.t. the 17 shape parameters, for every run."""
    # two launches: the values first (to form d f / d outputs), then values + gradient; a custom cotangent needs the values
    cot = torch.zeros((6, r.shape[1]), device="cuda")
    vals, _ = ev.evaluate_scalar_grad_device(r, cot, model_size)
    conf, cl, cd = vals[0], vals[1], vals[2]
    pen = torch.clamp(0.9 - conf, min=0.0)
    f = cl / cd - 2000.0 * pen**2
    cot[0], cot[1], cot[2] = 4000.0 * pen, 1.0 / cd, -cl / cd**2
    _, grad = ev.evaluate_scalar_grad_device(r, cot, model_size, with_values=False)
    return f, grad[:17], (cl, cd, conf)


f, g, _ = objective(raw)
f0 = f.clone()
torch.cuda.synchronize()
t0 = time.perf_counter()
iters = 60
for it in range(iters):
    trial = raw.clone()
    trial[:17] += step * g
    trial[0:8] = torch.maximum(trial[0:8], trial[8:16] + 0.05)  # keep some thickness everywhere
    f_new, g_new, _ = objective(trial)
    better = f_new > f
    raw = torch.where(better, trial, raw)
    g = torch.where(better, g_new, g)
    f = torch.where(better, f_new, f)
    step = torch.where(better, step * 1.3, step * 0.5)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
_, _, (cl, cd, conf) = objective(raw)
best = int(torch.argmax(torch.where(conf > 0.9, cl / cd, torch.zeros_like(cl))))
print(f"{runs} runs x {iters} iterations on {model_size}: {dt:.2f} s = {2 * runs * iters / dt / 1e6:.1f} M kernel queries/s "
      f"({runs * iters / dt / 1e6:.1f} M objective gradients/s)")
print(f"median objective {float(f0.median()):.1f} -> {float(f.median()):.1f}; runs improved: {float((f > f0).float().mean()) * 100:.1f} %")
print(f"best trusted design: L/D {float(cl[best] / cd[best]):.1f} at alpha {float(raw[18, best]):.2f} deg, Re {float(raw[19, best]):.3g}, "
      f"confidence {float(conf[best]):.3f}")
print("upper:", np.round(raw[0:8, best].cpu().numpy(), 4), " lower:", np.round(raw[8:16, best].cpu().numpy(), 4))
