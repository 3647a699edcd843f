"""ncu target: the fused reverse pass (all 198 cotangents) for one model size.
    ncu --set full --clock-control none --import-source on -k regex:nfb_grad_kernel -s 1 -c 1 -o gpurun_out/grad python scripts/ncu_target_grad.py xlarge"""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import numpy as np, torch
from neuralfoil_b200 import Evaluator, synthetic

size = sys.argv[1] if len(sys.argv) > 1 else "xlarge"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
ev = Evaluator(device=0, extra_models={"xxxlarge": synthetic.synthetic_xxxlarge_params()})
r = torch.from_numpy(synthetic.sample_training_distribution(n, seed=1, dtype=np.float32)).cuda()
c = torch.randn((198, n), device="cuda")
for _ in range(2):
    ev.evaluate_vjp_device(r, c, size, reverse=True)
torch.cuda.synchronize()
print("done")
