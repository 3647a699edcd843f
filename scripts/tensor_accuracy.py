"""Error statistics of the one-term tensor variant against the oracle (tests/parity.py 'tensor' tolerance quantities).
NFB_TC_ONE_SM=1 / NFB_LIB_PATH select other builds for comparison."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import numpy as np
import parity
from neuralfoil_b200 import Evaluator, synthetic
from oracle.neuralfoil_oracle import Oracle

extra = {"xxxlarge": synthetic.synthetic_xxxlarge_params()}
ev = Evaluator(device=0, extra_models=extra)
orc = Oracle(extra_models=extra)
for size, n in (("xxlarge", 20000), ("xxxlarge", 5003)):
    raw = synthetic.sample_training_distribution(n, seed=42)
    want = orc.evaluate_raw(raw, size)
    got = ev.evaluate_host(raw, size, out_dtype=np.float64, variant="tensor")
    rep = parity.error_report(got, want, raw[19], tier="tensor") if "tier" in parity.error_report.__code__.co_varnames else parity.error_report(got, want, raw[19])
    d = np.abs(got - want)
    print(f"{size} n={n}: median_rel {rep['median_rel']:.2e} worst_ratio {rep['worst_ratio']:.3f}  p99.9 |d|: conf {np.quantile(d[0], .999):.2e} CL {np.quantile(d[1], .999):.2e} "
          f"CDrel {np.quantile(d[2] / want[2], .999):.2e} CM {np.quantile(d[3], .999):.2e}  max|d| CL {d[1].max():.2e} conf {d[0].max():.2e}", flush=True)
