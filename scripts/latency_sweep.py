"""BASELINE.json configs[3]: small-batch latency of the drop-in dict API (host arrays in, dict out), median of N calls."""
import json, sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import numpy as np
from neuralfoil_b200 import Evaluator, synthetic
from oracle.neuralfoil_oracle import Oracle  # CPU comparison only

ev = Evaluator(device=0)
orc = Oracle()
K = synthetic.FIXED_KULFAN
rows = []
for size in ("xxsmall", "xlarge"):
    for n in (1, 10, 100, 1000, 10000):
        alpha = np.linspace(-5, 10, n) if n > 1 else 3.0
        for _ in range(20):
            ev.get_aero_from_kulfan_parameters(K, alpha=alpha, Re=1e6, model_size=size)
        ts = []
        for _ in range(300):
            t0 = time.perf_counter()
            ev.get_aero_from_kulfan_parameters(K, alpha=alpha, Re=1e6, model_size=size)
            ts.append(time.perf_counter() - t0)
        tc = []
        for _ in range(5 if n >= 1000 else 30):
            t0 = time.perf_counter()
            orc.get_aero_from_kulfan_parameters(K, alpha=alpha, Re=1e6, model_size=size)
            tc.append(time.perf_counter() - t0)
        row = {"model_size": size, "batch": n, "gpu_dict_api_us_median": 1e6 * float(np.median(ts)), "gpu_dict_api_us_p10": 1e6 * float(np.quantile(ts, 0.1)),
               "cpu_oracle_us_median": 1e6 * float(np.median(tc))}
        rows.append(row)
        print(json.dumps(row), flush=True)
