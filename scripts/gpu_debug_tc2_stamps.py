"""Phase stamps and MMA-warp wait accounting of a CTA-pair tensor kernel (tile 1 of CTA 0), and its throughput.
    python scripts/gpu_debug_tc2_stamps.py [tensor|tensor_fp32] [xxxlarge|xxlarge]"""
import sys, ctypes as C
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import numpy as np, torch
from neuralfoil_b200 import Evaluator, synthetic

variant = sys.argv[1] if len(sys.argv) > 1 else "tensor_fp32"
size = sys.argv[2] if len(sys.argv) > 2 else "xxxlarge"
ev = Evaluator(device=0, extra_models={"xxxlarge": synthetic.synthetic_xxxlarge_params()})
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
print(f"{size} {variant}: {n / ms / 1e3:.2f} M q/s ({ms:.3f} ms for {n} queries)", flush=True)
buf = (C.c_longlong * 64)()
ev._lib.nfb_debug_tc_stamps(ev._ctx, 1, None)
ev.evaluate_device(r, size, out=out, variant=variant); torch.cuda.synchronize()
ev._lib.nfb_debug_tc_stamps(ev._ctx, 0, buf)
st = list(buf); t0 = st[0]
print("stamps (cycles from tile start): featurise done", st[1] - t0)
L = 7 if size == "xxxlarge" else 6
for l in range(L - 1):
    print(f"  layer {l}: D0 ready {st[2 + 2 * l] - t0}, both halves written {st[3 + 2 * l] - t0}")
print("  last: D ready", st[20] - t0, "staged", st[21] - t0, "decoded", st[22] - t0)
for l in range(L):
    print(f"  MMA layer {l}: start {st[32 + 2 * l] - t0} end {st[33 + 2 * l] - t0} (span {st[33 + 2 * l] - st[32 + 2 * l]})")
print(f"  MMA warp waits over the tile: weight slots {st[48]}, a0 {st[49]}, a1 {st[50]} cycles")
print(f"  producer waits for a free slot over the tile: leader {st[51]}, peer {st[52]} cycles")
print(f"  layer 2 epilogue: D0 ready {st[6] - t0}, half 0 activated {st[53] - t0}, K region 0 free {st[54] - t0}, half 0 published {st[55] - t0}, D1 ready {st[56] - t0}, half 1 published {st[7] - t0}")
