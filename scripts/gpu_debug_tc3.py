"""Diagnostic run of the three-term (fp32-class) tensor variant: fp32-tolerance report against the oracle, and timing."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import numpy as np, torch
import parity
from neuralfoil_b200 import Evaluator, synthetic
from oracle.neuralfoil_oracle import Oracle

extra = {"xxxlarge": synthetic.synthetic_xxxlarge_params()}
ev = Evaluator(device=0, extra_models=extra)
orc = Oracle(extra_models=extra)
for size, n in (("xxlarge", 64), ("xxlarge", 20011), ("xxxlarge", 32), ("xxxlarge", 5003)):
    raw = synthetic.sample_training_distribution(n, seed=11)
    want = orc.evaluate_raw(raw, size)
    for variant in ("fp32", "tensor_fp32"):
        got = ev.evaluate_host(raw, size, out_dtype=np.float64, variant=variant)
        rep = parity.error_report(got, want, raw[19])
        print(f"{size} n={n} {variant:12s} n_bad {rep['n_bad']} worst {rep['worst_ratio']:.3f} median_rel {rep['median_rel']:.2e} frac<=1e-5 {rep['frac_rel_le_1e-5']:.4f} max|d| scalars {[f'{v:.1e}' for v in rep['max_abs_scalars']]}", flush=True)
for size, variant, n in (("xxlarge", "tensor_fp32", 4_000_000), ("xxxlarge", "tensor_fp32", 4_000_000), ("xxxlarge", "tensor", 4_000_000)):
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
import ctypes as C
buf = (C.c_longlong * 64)()
ev._lib.nfb_debug_tc_stamps(ev._ctx, 1, None)
r = torch.from_numpy(synthetic.sample_training_distribution(32 * 148 * 4, seed=1, dtype=np.float32)).cuda()
ev.evaluate_device(r, "xxxlarge", variant="tensor_fp32"); torch.cuda.synchronize()
ev._lib.nfb_debug_tc_stamps(ev._ctx, 0, buf)
st = list(buf); t0 = st[0]
print("x3 xxxlarge stamps: featurise done", st[1] - t0)
for l in range(6):
    print(f"  layer {l}: D0 ready {st[2 + 2 * l] - t0}, both halves written {st[3 + 2 * l] - t0}")
print("  last: D ready", st[20] - t0, "staged", st[21] - t0, "decoded", st[22] - t0)
for l in range(7):
    print(f"  MMA layer {l}: start {st[32 + 2 * l] - t0} end {st[33 + 2 * l] - t0} (span {st[33 + 2 * l] - st[32 + 2 * l]})")
print(f"  MMA warp waits over the tile: weight slots {st[48]}, a0 {st[49]}, a1 {st[50]} cycles")
