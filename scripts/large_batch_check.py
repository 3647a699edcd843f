"""Size-independent check at a BASELINE-size batch: every variant must give, for the same queries, bit-identical results
whether they sit at the start of a small batch or deep inside a 30 M-query one (different tiles, CTAs, pipeline phases)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import numpy as np, torch
from neuralfoil_b200 import Evaluator, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30_000_000
ev = Evaluator(device=0, extra_models={"xxxlarge": synthetic.synthetic_xxxlarge_params()})
base = torch.from_numpy(synthetic.sample_training_distribution(1_000_000, seed=9, dtype=np.float32)).cuda()
raw = base.repeat(1, n // 1_000_000)                      # (23, n): the same million queries over and over
out = torch.empty((198, n), dtype=torch.float32, device="cuda")
for size, variant in (("xxxlarge", "tensor"), ("xxxlarge", "tensor_fp32"), ("xxlarge", "tensor"), ("xxlarge", "tensor_fp32"), ("xxxlarge", "fp32")):
    if variant == "fp32" and n > 10_000_000:
        m = 10_000_000
    else:
        m = n
    ev.evaluate_device(raw[:, :m], size, out=out[:, :m], variant=variant)
    small = ev.evaluate_device(base[:, :4097].contiguous(), size, variant=variant)
    torch.cuda.synchronize()
    ok_first = torch.equal(out[:, :4097], small)
    k = (m // 1_000_000 - 1) * 1_000_000
    ok_last = torch.equal(out[:, k : k + 4097], small)
    finite = bool(torch.isfinite(out[:6, :m]).all())
    print(f"{size} {variant}: n = {m}: first block identical {ok_first}, block at {k} identical {ok_last}, scalars finite {finite}", flush=True)
    assert ok_first and ok_last and finite
print("LARGE_BATCH_OK")
