"""Phase stamps of a narrow tensor kernel (nfb_tcn.cuh; tile 1 of CTA 0) and its throughput.
    python scripts/gpu_debug_tcn_stamps.py [tensor|tensor_fp32] [xlarge|large|medium|...]"""
import sys, ctypes as C
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import numpy as np, torch
from neuralfoil_b200 import Evaluator, synthetic

variant = sys.argv[1] if len(sys.argv) > 1 else "tensor"
size = sys.argv[2] if len(sys.argv) > 2 else "xlarge"
ev = Evaluator(device=0)
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
ms = e0.elapsed_time(e1) / 3
tiles_per_cta = n / 64 / 148
print(f"{size} {variant}: {n / ms / 1e3:.2f} M q/s ({ms:.3f} ms for {n} queries; {ms * 1e-3 * 1.965e9 / tiles_per_cta:.0f} cycles per tile per CTA)", flush=True)
buf = (C.c_longlong * 64)()
ev._lib.nfb_debug_tc_stamps(ev._ctx, 1, None)
ev.evaluate_device(r, size, out=out, variant=variant); torch.cuda.synchronize()
ev._lib.nfb_debug_tc_stamps(ev._ctx, 0, buf)
st = list(buf); t0 = st[0]
L = len(ev.model_dims[size]) - 1
for l in range(L - 1):
    print(f"  layer {l}: MMA start {st[32 + 2 * l] - t0} end {st[33 + 2 * l] - t0} | accumulators ready {st[2 + 2 * l] - t0}, activations published {st[3 + 2 * l] - t0}")
print(f"  last layer: MMA start {st[32 + 2 * (L - 1)] - t0} end {st[33 + 2 * (L - 1)] - t0} | ready {st[20] - t0}, accumulators drained {st[21] - t0}, decoded {st[22] - t0}")
print(f"  featurise-ahead (tile 2): {st[47] - st[46]} cycles from start to published")
