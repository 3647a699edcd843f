"""Training-step throughput (nfb_train_step_device): samples/s and fp32 TFLOP/s per model size and batch size.
    python scripts/train_throughput.py [size ...]"""
import json, sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import numpy as np, torch  # noqa: E402
from neuralfoil_b200 import Evaluator, synthetic  # noqa: E402

FLOP = {"xxsmall": 52_032, "xsmall": 61_248, "small": 89_856, "medium": 106_240, "large": 310_784, "xlarge": 376_320,
        "xxlarge": 1_276_928, "xxxlarge": 5_699_584}  # forward, both twins (bench.py)
sizes = sys.argv[1:] or ["xxsmall", "xlarge", "xxxlarge"]
ev = Evaluator(device=0, extra_models={"xxxlarge": synthetic.synthetic_xxxlarge_params()})
peak = ev.measure_fp32_peak_tflops(3)
for size in sizes:
    for batch in (256, 8192, 65536, 262144):
        if size == "xxxlarge" and batch > 65536:
            continue
        x = torch.randn((batch, 25), device="cuda") * 0.3
        y = torch.randn((batch, 198), device="cuda") * 0.1
        y[:, 0] = (torch.rand(batch, device="cuda") < 0.8).float()
        for _ in range(3):
            ev.train_step_device(x, y, size)
        torch.cuda.synchronize()
        reps = 20 if batch <= 8192 else 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ev.train_step_device(x, y, size)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tf = 3 * FLOP[size] * batch / (ms * 1e-3) / 1e12
        print(json.dumps({"size": size, "batch": batch, "ms_per_step": round(ms, 4), "samples_per_s": round(batch / (ms * 1e-3)),
                          "tflops_fp32": round(tf, 2), "frac_of_measured_fp32_peak": round(tf / peak, 3), "fp32_peak": round(peak, 1)}), flush=True)
