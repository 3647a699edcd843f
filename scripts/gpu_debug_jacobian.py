"""Diagnostic run of the Jacobian kernel: errors against the oracle's analytic float64 Jacobian, bit-equality of the
values with the fp32 path, and timing."""
import sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import numpy as np, torch
from neuralfoil_b200 import Evaluator, synthetic
from oracle.neuralfoil_oracle import Oracle

extra = {"xxxlarge": synthetic.synthetic_xxxlarge_params()}
ev = Evaluator(device=0, extra_models=extra)
orc = Oracle(extra_models=extra)
# typical size of a change in each input (the scale a derivative is multiplied by in practice)
SCALE = np.array([0.1] * 16 + [0.1, 0.002, 1.0, 0.0, 1.0, 0.1, 0.1])
for size, n in (("xxsmall", 500), ("medium", 500), ("xlarge", 500), ("xxlarge", 300), ("xxxlarge", 200)):
    raw = synthetic.sample_training_distribution(n, seed=21)
    want_v, want_j = orc.jacobian_raw(raw, size)          # (198, n), (198, 23, n)
    got_v, got_j = ev.evaluate_jacobian_host(raw, size)   # (198, n), (n, 198, 23)
    ref_v = ev.evaluate_host(raw, size, out_dtype=np.float64)
    got_j = np.transpose(got_j, (1, 2, 0))
    scale = SCALE[None, :, None] * np.ones_like(want_j)
    scale[:, 19, :] = raw[19][None, :]                    # Re: a relative change
    dw, dg = want_j * scale, got_j * scale                # change of each output for a typical change of each input
    err = np.abs(dg - dw)
    # reference magnitude per output row: the output itself (theta scaled like the parity tests do is overkill here)
    mag = np.abs(want_v)[:, None, :] + np.abs(dw)
    rel = err / np.maximum(mag, 1e-30)
    print(f"{size} n={n}: values bit-identical to fp32 path: {np.array_equal(got_v, ref_v)}  finite {np.isfinite(got_j).all()}"
          f"  |dJ|/(|out|+|J|) median {np.median(rel):.2e} p99.9 {np.quantile(rel, .999):.2e} max {rel.max():.2e}", flush=True)
    for r, name in ((1, "CL"), (2, "CD"), (3, "CM"), (0, "conf")):
        k = np.unravel_index(np.argmax(rel[r]), rel[r].shape)
        print(f"    {name}: max rel {rel[r].max():.2e} at input {k[0]} (got {got_j[r][k]:.6g} want {want_j[r][k]:.6g}); dCL/dalpha sample {got_j[1, 18, 0]:.6f} vs {want_j[1, 18, 0]:.6f}" if r == 1 else f"    {name}: max rel {rel[r].max():.2e}")
for size, n in (("xlarge", 1), ("xlarge", 100), ("xlarge", 10000), ("xxxlarge", 100), ("xxxlarge", 10000)):
    r = torch.from_numpy(synthetic.sample_training_distribution(n, seed=1, dtype=np.float32)).cuda()
    for _ in range(2):
        ev.evaluate_jacobian_device(r, size)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ev.evaluate_jacobian_device(r, size)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{size} n={n}: {ms * 1e3:.1f} us per call = {n / ms / 1e3:.3f} M Jacobians/s (device-resident)", flush=True)
kp = dict(upper_weights=np.array([0.18, 0.21, 0.19, 0.22, 0.18, 0.21, 0.17, 0.16]), lower_weights=-np.array([0.10, 0.08, 0.12, 0.05, 0.10, 0.06, 0.08, 0.05]),
          leading_edge_weight=0.1, TE_thickness=0.002)
t0 = time.perf_counter()
for _ in range(50):
    aero, jac = ev.get_aero_and_jacobian_from_kulfan_parameters(kp, alpha=4.0, Re=2e6, model_size="xlarge")
print(f"dict API, batch 1, xlarge: {(time.perf_counter() - t0) / 50 * 1e6:.0f} us per call; dCL/dalpha = {jac['CL'][0, 18]:.5f} per degree")
