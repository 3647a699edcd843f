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
for size, n in (("xxxlarge", 64), ("xxxlarge", 1003), ("xxlarge", 64), ("xxlarge", 20000)):
    raw = synthetic.sample_training_distribution(n, seed=42)
    want, zw = orc.evaluate_raw(raw, size, return_latent=True)
    ref32 = ev.evaluate_host(raw, size, out_dtype=np.float64)
    got = ev.evaluate_host(raw, size, out_dtype=np.float64, variant="tensor")
    print(size, end=" ")
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
for size, variant in (("xxlarge", "tensor"), ("xxlarge", "fp32")):
    n = 4_000_000
    r = torch.from_numpy(synthetic.sample_training_distribution(n, seed=1, dtype=np.float32)).cuda()
    out = torch.empty((198, n), dtype=torch.float32, device="cuda")
    for _ in range(2):
        ev.evaluate_device(r, size, out=out, variant=variant)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ev.evaluate_device(r, size, out=out, variant=variant)
    e1.record(); torch.cuda.synchronize()
    print(f"{size} {variant}: {n / (e0.elapsed_time(e1) / 3) / 1e3:.2f} M q/s", flush=True)
    del r, out
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

# phase stamps of CTA 0's second tile
import ctypes as C
buf = (C.c_longlong * 64)()
ev._lib.nfb_debug_tc_stamps(ev._ctx, 1, None)
r = torch.from_numpy(synthetic.sample_training_distribution(64 * 148 * 4, seed=1, dtype=np.float32)).cuda()
ev.evaluate_device(r, "xxxlarge", variant="tensor"); torch.cuda.synchronize()
ev._lib.nfb_debug_tc_stamps(ev._ctx, 0, buf)
st = list(buf)
t0 = st[0]
print("epilogue-warp stamps (cycles from tile start): featurise done", st[1] - t0)
for l in range(6):
    print(f"  layer {l}: D ready at {st[2 + 2 * l] - t0}, epilogue done at {st[3 + 2 * l] - t0}  (epilogue {st[3 + 2 * l] - st[2 + 2 * l]})")
print("  last layer D ready", st[20] - t0, "staged", st[21] - t0, "decoded+stored", st[22] - t0)
for l in range(7):
    print(f"  MMA layer {l}: start {st[32 + 2 * l] - t0} issue-end {st[33 + 2 * l] - t0} (span {st[33 + 2 * l] - st[32 + 2 * l]})")
