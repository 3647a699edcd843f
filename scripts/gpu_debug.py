"""Diagnostic run on the GPU box: per-size error report against the oracle, plus quick timings."""
import sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import numpy as np
import parity
from neuralfoil_b200 import Evaluator, synthetic
from oracle.neuralfoil_oracle import Oracle

extra = {"xxxlarge": synthetic.synthetic_xxxlarge_params()}
ev = Evaluator(device=0, extra_models=extra)
print(ev.device_info(), flush=True)
orc = Oracle(extra_models=extra)
raw = synthetic.sample_training_distribution(1003, seed=42)
sizes = sys.argv[1:] or ["xxsmall", "xsmall", "small", "medium", "large", "xlarge", "xxlarge", "xxxlarge"]
for size in sizes:
    got = ev.evaluate_host(raw, size, out_dtype=np.float64)
    want = orc.evaluate_raw(raw, size)
    rep = parity.error_report(got, want, raw[19])
    print(size, rep, flush=True)
    if rep["n_bad"]:
        err = np.abs(got - want)
        rows = np.argsort(-np.nanmax(err, axis=1))[:8]
        for r in rows:
            c = int(np.nanargmax(err[r]))
            print(f"   row {r} col {c}: got {got[r, c]:.8g} want {want[r, c]:.8g}")
print("fp32 peak TFLOP/s:", ev.measure_fp32_peak_tflops(5), flush=True)
import torch
for size, n in [("xxxlarge", 1_000_000), ("xxlarge", 2_000_000), ("xlarge", 4_000_000), ("large", 4_000_000), ("medium", 4_000_000), ("xxsmall", 4_000_000)]:
    r = torch.from_numpy(synthetic.sample_training_distribution(n, seed=1, dtype=np.float32)).cuda()
    out = torch.empty((198, n), dtype=torch.float32, device="cuda")
    for _ in range(2):
        ev.evaluate_device(r, size, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ev.evaluate_device(r, size, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{size}: n={n} {ms:.2f} ms -> {n / ms / 1e3:.2f} M q/s", flush=True)
    del r, out
