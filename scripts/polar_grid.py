"""BASELINE.json configs[1]: one airfoil, 1,000,000 (alpha x Re x n_crit) points, every model size, one GPU.

Three timings per size (median of 5 after 2 warm-up calls):
  dict    the drop-in call: host arrays in, the reference's 198-key dict of float64 (N,) arrays out
  host32  Evaluator.evaluate_host: the same 23 rows, one (198, N) float32 host block out
  device  Evaluator.evaluate_device: broadcast rows (stride 0) and the three varying rows resident in HBM
and a parity check of 2,000 random points of the grid against the oracle.
"""
import json, sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import numpy as np, torch
import parity
from neuralfoil_b200 import Evaluator, synthetic
from neuralfoil_b200.main import output_keys
from oracle.neuralfoil_oracle import Oracle

extra = {"xxxlarge": synthetic.synthetic_xxxlarge_params()}
ev, orc = Evaluator(device=0, extra_models=extra), Oracle(extra_models=extra)
kulfan = dict(  # the reference's fixed test airfoil (tests/test_golden_values.py:22-27)
    upper_weights=np.array([0.18, 0.24, 0.20, 0.22, 0.18, 0.20, 0.16, 0.15]),
    lower_weights=np.array([-0.12, -0.06, -0.10, -0.04, -0.06, -0.02, 0.02, 0.05]),
    leading_edge_weight=0.1, TE_thickness=0.002)
A, R, Ncr = np.meshgrid(np.linspace(-20, 20, 100), np.logspace(4, 8, 100), np.linspace(1, 17, 100), indexing="ij")
alpha, Re, n_crit = A.ravel(), R.ravel(), Ncr.ravel()
n = alpha.size
sizes = sys.argv[1:] or ["xxsmall", "xsmall", "small", "medium", "large", "xlarge", "xxlarge", "xxxlarge"]


def med(f, reps=5, warm=2):
    for _ in range(warm): f()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


raw = np.empty((23, n))
raw[0:8] = kulfan["upper_weights"][:, None]; raw[8:16] = kulfan["lower_weights"][:, None]
raw[16], raw[17], raw[18], raw[19], raw[20], raw[21], raw[22] = 0.1, 0.002, alpha, Re, n_crit, 1.0, 1.0
pick = np.random.default_rng(3).choice(n, 2000, replace=False)
dev_rows = [torch.tensor([raw[i, 0]], dtype=torch.float32, device="cuda") for i in range(23)]
for i in (18, 19, 20):
    dev_rows[i] = torch.from_numpy(raw[i].astype(np.float32)).cuda()
for size in sizes:
    aero = ev.get_aero_from_kulfan_parameters(kulfan, alpha, Re, n_crit, model_size=size)
    got = np.stack([aero[k][pick] for k in output_keys()])
    rep = parity.assert_parity(got, orc.evaluate_raw(raw[:, pick], size), Re[pick], check_distribution=False, label=size)
    t_dict = med(lambda: ev.get_aero_from_kulfan_parameters(kulfan, alpha, Re, n_crit, model_size=size))
    t_host = med(lambda: ev.evaluate_host(raw.astype(np.float32), size))
    out = torch.empty((198, n), dtype=torch.float32, device="cuda")
    t_dev = med(lambda: ev.evaluate_device(dev_rows, size, out=out))
    print(json.dumps({"size": size, "points": n, "dict_ms": round(1e3 * t_dict, 2), "host_f32_ms": round(1e3 * t_host, 2),
                      "device_ms": round(1e3 * t_dev, 3), "device_Mqps": round(n / t_dev / 1e6, 1),
                      "dict_Mqps": round(n / t_dict / 1e6, 2), "worst_err_over_tol": round(rep["worst_ratio"], 3)}), flush=True)
