"""Device-resident throughput of every model size (fp32 kernel), quick."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import numpy as np, torch
from neuralfoil_b200 import Evaluator, synthetic
ev = Evaluator(device=0, extra_models={"xxxlarge": synthetic.synthetic_xxxlarge_params()})
sizes = sys.argv[1:] or ["xxsmall", "xsmall", "small", "medium", "large", "xlarge", "xxlarge", "xxxlarge"]
FLOP = {"xxsmall": 52032, "xsmall": 61248, "small": 89856, "medium": 106240, "large": 310784, "xlarge": 376320, "xxlarge": 1276928, "xxxlarge": 5699584}
for size in sizes:
    n = 1_000_000 if size == "xxxlarge" else 4_000_000
    r = torch.from_numpy(synthetic.sample_training_distribution(n, seed=1, dtype=np.float32)).cuda()
    out = torch.empty((198, n), dtype=torch.float32, device="cuda")
    for _ in range(2): ev.evaluate_device(r, size, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): ev.evaluate_device(r, size, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{size}: {n / ms / 1e3:.1f} M q/s  {FLOP[size] * n / ms / 1e9:.1f} TFLOP/s  checksum {float(out.double().nan_to_num(posinf=0.0, neginf=0.0).sum()):.12e}", flush=True)
    del r, out
