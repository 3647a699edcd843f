"""Device-resident throughput of the tensor-core variants (4 M queries, fp32 in / fp32 out), both wide models.
NFB_TC_ONE_SM=1 selects the one-SM kernels (nfb_tc.cuh) instead of the CTA-pair kernels (nfb_tc2.cuh)."""
import os, sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import numpy as np, torch
from neuralfoil_b200 import Evaluator, synthetic

ev = Evaluator(device=0, extra_models={"xxxlarge": synthetic.synthetic_xxxlarge_params()})
n = 4_000_000
r = torch.from_numpy(synthetic.sample_training_distribution(n, seed=1, dtype=np.float32)).cuda()
out = torch.empty((198, n), dtype=torch.float32, device="cuda")
print("kernels:", "one-SM" if os.environ.get("NFB_TC_ONE_SM") else "CTA pair")
for size in ("xxlarge", "xxxlarge"):
    for variant in ("tensor", "tensor_fp32"):
        for _ in range(2):
            ev.evaluate_device(r, size, out=out, variant=variant)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ev.evaluate_device(r, size, out=out, variant=variant)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"  {size:9s} {variant:12s} {n / ms / 1e3:8.2f} M q/s", flush=True)
