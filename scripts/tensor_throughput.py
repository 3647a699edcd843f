"""Device-resident throughput of the tensor-core variants, every model size that has them (4 M queries, fp32 in / fp32
out; full outputs and outputs="scalars"), beside the fp32 kernel.  `one_sm` as first argument selects the one-SM kernels
(nfb_tc.cuh) for the wide models instead of the CTA-pair kernels (nfb_tc2.cuh)."""
import json, sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import numpy as np, torch
from neuralfoil_b200 import Evaluator, synthetic

FLOP = {"xxsmall": 52032, "xsmall": 61248, "small": 89856, "medium": 106240, "large": 310784, "xlarge": 376320, "xxlarge": 1276928, "xxxlarge": 5699584}
suffix = ":one_sm" if "one_sm" in sys.argv[1:] else ""
sizes = [a for a in sys.argv[1:] if a in FLOP] or list(FLOP)
ev = Evaluator(device=0, extra_models={"xxxlarge": synthetic.synthetic_xxxlarge_params()})
n = 4_000_000
r = torch.from_numpy(synthetic.sample_training_distribution(n, seed=1, dtype=np.float32)).cuda()
out = torch.empty((198, n), dtype=torch.float32, device="cuda")
out6 = torch.empty((6, n), dtype=torch.float32, device="cuda")
for size in sizes:
    row = {"size": size}
    for variant in ("fp32", "tensor", "tensor_fp32"):
        for outputs, o in (("all", out), ("scalars", out6)):
            v = variant + (suffix if variant != "fp32" and size in ("xxlarge", "xxxlarge") else "")
            for _ in range(2):
                ev.evaluate_device(r, size, out=o, variant=v, outputs=outputs)
            torch.cuda.synchronize()
            reps = 3 if size == "xxxlarge" and variant == "fp32" else 5
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                ev.evaluate_device(r, size, out=o, variant=v, outputs=outputs)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            key = variant + ("" if outputs == "all" else "_scalars")
            row[key + "_Mqps"] = round(n / ms / 1e3, 1)
            if outputs == "all":
                row[key + "_tflops"] = round((3 if variant == "tensor_fp32" else 1) * FLOP[size] * n / ms / 1e9, 1)
    print(json.dumps(row), flush=True)
