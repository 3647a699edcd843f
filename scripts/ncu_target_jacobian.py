"""ncu target: one warm-up and one measured launch of the Jacobian kernel, xlarge and xxxlarge, 10,000 queries.
    ncu --set full --clock-control none --import-source on -k regex:nfb_jac -o gpurun_out/jac python scripts/ncu_target_jacobian.py"""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import numpy as np, torch
from neuralfoil_b200 import Evaluator, synthetic

ev = Evaluator(device=0, extra_models={"xxxlarge": synthetic.synthetic_xxxlarge_params()})
r = torch.from_numpy(synthetic.sample_training_distribution(10_000, seed=1, dtype=np.float32)).cuda()
for size in ("xlarge", "xxxlarge"):
    for _ in range(2):
        ev.evaluate_jacobian_device(r, size)
    torch.cuda.synchronize()
print("done")
