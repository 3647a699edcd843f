"""Diagnostic run of the tensor-core variant on the GPU box: errors against the oracle and timing."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import numpy as np
from neuralfoil_b200 import Evaluator, synthetic
from oracle.neuralfoil_oracle import Oracle

extra = {"xxxlarge": synthetic.synthetic_xxxlarge_params()}
ev = Evaluator(device=0, extra_models=extra)
orc = Oracle(extra_models=extra)
for n in (64, 1003):
    raw = synthetic.sample_training_distribution(n, seed=42)
    want, zw = orc.evaluate_raw(raw, "xxxlarge", return_latent=True)
    ref32 = ev.evaluate_host(raw, "xxxlarge", out_dtype=np.float64)
    got = ev.evaluate_host(raw, "xxxlarge", out_dtype=np.float64, variant="tensor")
    d = np.abs(got - want)
    print(f"n={n}: finite {np.isfinite(got).all()}  fp32-path max|d| scalars {[float(np.abs(ref32[i]-want[i]).max()) for i in range(6)]}")
    print("   tensor max|d| scalars", [float(d[i].max()) for i in range(6)])
    print("   tensor p99.9 |d|: conf %.2e CL %.2e CDrel %.2e CM %.2e  ue max %.2e  H rel max %.2e" % (
        np.quantile(d[0], .999), np.quantile(d[1], .999), np.quantile(d[2] / want[2], .999), np.quantile(d[3], .999),
        d[70:102].max(), (d[38:70] / want[38:70]).max()))
    if n == 64:
        print("   sample CL got/want", got[1, :4], want[1, :4])
        print("   sample row 6", got[6, :3], want[6, :3], " row 197", got[197, :3], want[197, :3])
import torch
for variant in ("tensor", "fp32"):
    n = 4_000_000 if variant == "tensor" else 1_000_000
    r = torch.from_numpy(synthetic.sample_training_distribution(n, seed=1, dtype=np.float32)).cuda()
    out = torch.empty((198, n), dtype=torch.float32, device="cuda")
    for _ in range(2):
        ev.evaluate_device(r, "xxxlarge", out=out, variant=variant)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ev.evaluate_device(r, "xxxlarge", out=out, variant=variant)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{variant}: n={n} {ms:.2f} ms -> {n / ms / 1e3:.2f} M q/s", flush=True)
