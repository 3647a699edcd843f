"""Throughput of the reverse-mode kernels (all 198 cotangents, and the six scalars) for one or more sizes; used with the
NFB_GRAD_EXP builds of nfb_grad.cuh to see what each phase costs.   python scripts/vjp_time.py [size ...]"""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import numpy as np, torch
from neuralfoil_b200 import Evaluator, synthetic

ev = Evaluator(device=0, extra_models={"xxxlarge": synthetic.synthetic_xxxlarge_params()})
for size in (sys.argv[1:] or ["xlarge"]):
    nn = 500_000 if size == "xxxlarge" else 2_000_000
    r = torch.from_numpy(synthetic.sample_training_distribution(nn, seed=1, dtype=np.float32)).cuda()
    c = torch.randn((198, nn), device="cuda")
    out = {}
    for name, f in (("full", lambda: ev.evaluate_vjp_device(r, c, size, reverse=True)),
                    ("scalars", lambda: ev.evaluate_scalar_grad_device(r, c[:6].contiguous(), size))):
        for _ in range(2): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): f()
        e1.record(); torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / 3
    print(f"{size}: full {out['full']:.2f} ms = {nn / out['full'] / 1e3:.1f} M VJP/s | scalars {out['scalars']:.2f} ms = {nn / out['scalars'] / 1e3:.1f} M grad/s", flush=True)
