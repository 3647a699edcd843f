"""ncu target: two launches of the fp32 kernel for one model size (default xlarge, 4 M queries).
    ncu --set full --clock-control none --import-source on -k regex:nfb_ffma_kernel -s 1 -c 1 -o gpurun_out/ffma python scripts/ncu_target_ffma.py xlarge"""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import numpy as np, torch
from neuralfoil_b200 import Evaluator, synthetic

size = sys.argv[1] if len(sys.argv) > 1 else "xlarge"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
ev = Evaluator(device=0, extra_models={"xxxlarge": synthetic.synthetic_xxxlarge_params()})
r = torch.from_numpy(synthetic.sample_training_distribution(n, seed=1, dtype=np.float32)).cuda()
out = torch.empty((198, n), dtype=torch.float32, device="cuda")
for _ in range(2):
    ev.evaluate_device(r, size, out=out)
torch.cuda.synchronize()
print("done", float(out.double().nan_to_num().sum()))
