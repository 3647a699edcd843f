"""Full reverse-mode VJP (nfb_grad.cuh, FULL) on the GPU box: cotangents of all 198 outputs; error against the oracle's analytic
float64 Jacobian and against the forward-mode kernel, values against the value kernel, throughput."""
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
cot = np.zeros((198, n))
cot[:6] = rng.normal(size=(6, n)) * np.array([1.0, 1.0, 50.0, 1.0, 1.0, 1.0])[:, None]
cot[parity.ROW_THETA] = 100.0 * rng.normal(size=(64, n))
cot[parity.ROW_H] = 0.1 * rng.normal(size=(64, n))
cot[parity.ROW_UE] = 0.1 * rng.normal(size=(64, n))
scale = parity.JAC_INPUT_SCALE.copy()[:, None] * np.ones((1, n)); scale[19] = raw[19]
for size in sizes:
    values, grad = ev.evaluate_vjp_host(raw, cot, size, reverse=True)
    ref_values = ev.evaluate_host(raw, size, out_dtype=np.float64)
    want_v, want_j = orc.jacobian_raw(raw, size)
    want = np.einsum("rtq,rq->tq", want_j, cot)
    _, fwd = ev.evaluate_vjp_host(raw, cot, size, reverse=False)
    err = np.abs(grad - want) * scale
    ref = np.abs(want) * scale + np.einsum("rq,rq->q", np.abs(cot), np.abs(want_v))[None, :] * 1e-1
    ok = (err <= 5e-3 * ref + 1e-6)
    efw = np.abs(fwd - want) * scale
    print(size, "values equal:", bool(np.array_equal(values, ref_values)), "finite", bool(np.isfinite(grad).all()), "frac ok", float(ok.mean()),
          "median rel", float(np.median(err / np.maximum(ref, 1e-30))), "max", float((err / np.maximum(ref, 1e-30)).max()),
          "| forward-mode: frac ok", float((efw <= 5e-3 * ref + 1e-6).mean()), "max", float((efw / np.maximum(ref, 1e-30)).max()),
          "| worst rows", np.argsort(-(err / np.maximum(ref, 1e-30)).max(axis=1))[:4].tolist(), flush=True)
for size, nn in (("xxsmall", 2_000_000), ("xlarge", 2_000_000), ("xxxlarge", 500_000)):
    if size not in sizes: continue
    r = torch.from_numpy(synthetic.sample_training_distribution(nn, seed=1, dtype=np.float32)).cuda()
    c = torch.randn((198, nn), device="cuda")
    for _ in range(2): ev.evaluate_vjp_device(r, c, size, reverse=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): ev.evaluate_vjp_device(r, c, size, reverse=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{size}: {nn / ms / 1e3:.2f} M VJPs/s ({ms:.2f} ms per {nn}; values + gradient, 198 cotangent rows in)", flush=True)
