"""Reverse-mode scalar-gradient kernel (nfb_grad.cuh) on the GPU box: error against the oracle's analytic float64 Jacobian and
against the forward-mode VJP kernel, values against the value kernel, throughput at large batches."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import numpy as np, torch
import parity
from neuralfoil_b200 import Evaluator, synthetic
from oracle.neuralfoil_oracle import Oracle

extra = {"xxxlarge": synthetic.synthetic_xxxlarge_params()}
ev, orc = Evaluator(device=0, extra_models=extra), Oracle(extra_models=extra)
sizes = sys.argv[1:] or ["xxsmall", "xsmall", "small", "medium", "large", "xlarge", "xxlarge", "xxxlarge"]
n = 1003
raw = synthetic.sample_training_distribution(n, seed=55)
rng = np.random.default_rng(5)
cot = rng.normal(size=(6, n)) * np.array([1.0, 1.0, 50.0, 1.0, 1.0, 1.0])[:, None]
scale = parity.JAC_INPUT_SCALE.copy()[:, None] * np.ones((1, n)); scale[19] = raw[19]
for size in sizes:
    values, grad = ev.evaluate_scalar_grad_host(raw, cot, size)
    ref_values = ev.evaluate_host(raw, size, out_dtype=np.float64, outputs="scalars")
    want_v, want_j = orc.jacobian_raw(raw, size)
    want = np.einsum("rtq,rq->tq", want_j[:6], cot)
    cot198 = np.zeros((198, n)); cot198[:6] = cot
    _, fwd = ev.evaluate_vjp_host(raw, cot198, size)
    err = np.abs(grad - want) * scale
    ref = np.abs(want) * scale + np.einsum("rq,rq->q", np.abs(cot), np.abs(want_v[:6]))[None, :] * 1e-1
    ok = (err <= 5e-3 * ref + 1e-6)
    print(size, "values equal:", bool(np.array_equal(values, ref_values)), "frac ok", float(ok.mean()), "median rel", float(np.median(err / np.maximum(ref, 1e-30))),
          "max err/ref", float((err / np.maximum(ref, 1e-30)).max()), "| vs forward-mode kernel: max scaled diff", float((np.abs(grad - fwd) * scale).max()),
          "rows worst", np.argsort(-(err / np.maximum(ref, 1e-30)).max(axis=1))[:4].tolist(), flush=True)
FLOP = {"xxsmall": 52032, "xlarge": 376320, "xxxlarge": 5699584}
for size, nn in (("xxsmall", 4_000_000), ("xlarge", 4_000_000), ("xxxlarge", 1_000_000)):
    if size not in sizes: continue
    r = torch.from_numpy(synthetic.sample_training_distribution(nn, seed=1, dtype=np.float32)).cuda()
    c = torch.randn((6, nn), device="cuda")
    for _ in range(2): ev.evaluate_scalar_grad_device(r, c, size)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): ev.evaluate_scalar_grad_device(r, c, size)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{size}: {nn / ms / 1e3:.2f} M gradients/s ({ms:.2f} ms per {nn})", flush=True)
