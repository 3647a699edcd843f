"""ncu target: one warm-up and one measured launch of each tensor variant on xxxlarge at the bench's size (10 M queries).
    ncu --set full --clock-control none --import-source on -k regex:nfb_tc2 -o gpurun_out/tc2 python scripts/ncu_target_tensor.py"""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import numpy as np, torch
from neuralfoil_b200 import Evaluator, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ev = Evaluator(device=0, extra_models={"xxxlarge": synthetic.synthetic_xxxlarge_params()})
r = torch.from_numpy(synthetic.sample_training_distribution(n, seed=1, dtype=np.float32)).cuda()
out = torch.empty((198, n), dtype=torch.float32, device="cuda")
for variant in ("tensor", "tensor_fp32"):
    for _ in range(2):
        ev.evaluate_device(r, "xxxlarge", out=out, variant=variant)
    torch.cuda.synchronize()
print("done")
