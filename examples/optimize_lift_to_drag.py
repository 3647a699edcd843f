"""
Gradient-based shape optimisation with the Jacobian kernel (the use case of the reference's
studies/sequential_multifidelity_optimization.py, which obtains its gradients by tracing main.py through CasADi).

Maximise CL / CD of an airfoil at fixed alpha and Re over its 18 Kulfan parameters, by projected gradient ascent with
the derivatives that `get_aero_and_jacobian_from_kulfan_parameters` returns (one fused CUDA launch per iterate), keeping
the analysis confidence of the network above a floor so that the optimiser does not wander out of the training
distribution.  Needs a B200.

    python examples/optimize_lift_to_drag.py [model_size]
"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import neuralfoil_b200 as nf  # noqa: E402

model_size = sys.argv[1] if len(sys.argv) > 1 else "xlarge"
alpha, Re = 3.0, 2e6
x = np.concatenate([[0.17, 0.16, 0.15, 0.14, 0.13, 0.13, 0.12, 0.12], [-0.13, -0.10, -0.08, -0.07, -0.06, -0.05, -0.05, -0.04], [0.05]])
names = nf.JACOBIAN_INPUTS


def unpack(v):
    return dict(upper_weights=v[0:8], lower_weights=v[8:16], leading_edge_weight=v[16], TE_thickness=0.002)


def objective(v):
    aero, jac = nf.get_aero_and_jacobian_from_kulfan_parameters(unpack(v), alpha=alpha, Re=Re, model_size=model_size)
    cl, cd, conf = aero["CL"][0], aero["CD"][0], aero["analysis_confidence"][0]
    dcl, dcd, dconf = jac["CL"][0, :17], jac["CD"][0, :17], jac["analysis_confidence"][0, :17]
    ld = cl / cd
    grad = dcl / cd - cl / cd**2 * dcd
    penalty = max(0.0, 0.90 - conf)  # stay where the network trusts itself
    return ld - 2000.0 * penalty**2, grad + 4000.0 * penalty * dconf, (cl, cd, conf)


objective(x)  # creates the CUDA context and loads the networks
t0 = time.perf_counter()
f, g, (cl, cd, conf) = objective(x)
print(f"start : L/D {cl / cd:7.2f}  CL {cl:.4f}  CD {cd:.5f}  confidence {conf:.3f}")
step = 2e-5
for it in range(200):
    trial = x + step * g
    trial[0:8] = np.maximum(trial[0:8], trial[8:16] + 0.05)  # keep some thickness everywhere
    f_new, g_new, aux = objective(trial)
    if f_new > f:
        x, f, g, (cl, cd, conf) = trial, f_new, g_new, aux
        step *= 1.3
    else:
        step *= 0.5
    if it % 40 == 39:
        print(f"it {it + 1:3d}: L/D {cl / cd:7.2f}  CL {cl:.4f}  CD {cd:.5f}  confidence {conf:.3f}  step {step:.1e}")
dt = time.perf_counter() - t0
print(f"final : L/D {cl / cd:7.2f}  after 201 evaluations with Jacobians in {dt:.2f} s ({dt / 201 * 1e6:.0f} us each), model {model_size}")
print("upper:", np.round(x[0:8], 4), " lower:", np.round(x[8:16], 4), " LE:", round(float(x[16]), 4))
