"""
compute-sanitizer target: small batches through every kernel family, each checked against the oracle.

    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python scripts/sanitizer_target.py [family ...]

Families: ffma_groups ffma_tiled small jac grad tc2 tc (default: all).  The batches are tiny (the tools slow a kernel
down 10 - 1000 x) but cover a ragged last tile, several tiles per CTA where the kernel is persistent, and every padded
width.  `small_batch_max=0` routes even these batches to the large-batch kernels.
"""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import numpy as np, torch  # noqa: E402
import parity  # noqa: E402
from neuralfoil_b200 import Evaluator, synthetic  # noqa: E402
from oracle.neuralfoil_oracle import Oracle  # noqa: E402

families = sys.argv[1:] or ["ffma_groups", "ffma_tiled", "small", "jac", "grad", "tc2", "tc"]
extra = {"xxxlarge": synthetic.synthetic_xxxlarge_params()}
ev, orc = Evaluator(device=0, extra_models=extra), Oracle(extra_models=extra)
SIZES = ("xxsmall", "large", "xxlarge", "xxxlarge")  # padded widths 64, 128, 256, 512


def check(label, got, raw, size, variant="fp32"):
    rep = parity.assert_parity(got, orc.evaluate_raw(raw, size), raw[19], check_distribution=False, variant=variant, label=label)
    print(f"{label}: {raw.shape[1]} queries ok, worst error/tolerance {rep['worst_ratio']:.3f}", flush=True)


for fam in families:
    if fam in ("ffma_groups", "ffma_tiled"):
        ev.set_tuning(small_batch_max=0)
        for size in SIZES:
            n = {"xxsmall": 148 * 256 + 77, "large": 148 * 128 + 77, "xxlarge": 148 * 64 + 33, "xxxlarge": 148 * 32 + 21}[size]
            n = n if fam == "ffma_groups" else n // 8  # more than one tile per CTA for the persistent column-group kernel
            raw = synthetic.sample_training_distribution(n, seed=3)
            check(f"{fam} {size}", ev.evaluate_host(raw, size, out_dtype=np.float64, variant="fp32" if fam == "ffma_groups" else "fp32:tiled"), raw, size)
        ev.set_tuning(small_batch_max=-1)
    elif fam == "small":
        for size in SIZES:
            raw = synthetic.sample_training_distribution(9, seed=4)
            check(f"small {size}", ev.evaluate_host(raw, size, out_dtype=np.float64), raw, size)
    elif fam == "jac":
        for size in SIZES:
            raw = synthetic.sample_training_distribution(3, seed=5)
            want_v, want_j = orc.jacobian_raw(raw, size)
            got_v, got_j = ev.evaluate_jacobian_host(raw, size)
            parity.assert_jacobian_parity(got_j, want_j, got_v, want_v, raw, label=f"jac {size}")
            cot = np.ones((198, 3)) * 0.01
            _, g = ev.evaluate_vjp_host(raw, cot, size, reverse=False)
            assert np.isfinite(g).all()
            print(f"jac {size}: ok", flush=True)
    elif fam == "grad":
        for size in SIZES:
            n = 300
            raw = synthetic.sample_training_distribution(n, seed=6)
            cot = np.zeros((198, n)); cot[1], cot[2], cot[3] = 1.0, -50.0, 0.5; cot[6 + 64: 6 + 96] = 0.01
            _, g_fwd = ev.evaluate_vjp_host(raw, cot, size, reverse=False)
            _, g_rev = ev.evaluate_vjp_host(raw, cot, size, reverse=True)
            _, g_sc = ev.evaluate_scalar_grad_host(raw, cot[:6].copy(), size)
            _, g_fwd6 = ev.evaluate_vjp_host(raw, np.concatenate([cot[:6], np.zeros((192, n))]), size, reverse=False)
            scale = parity.JAC_INPUT_SCALE.copy()[:, None] * np.ones((1, n)); scale[19] = raw[19]
            for name, a, b in (("reverse vjp", g_rev, g_fwd), ("scalar grad", g_sc, g_fwd6)):
                d = np.abs(a - b) * scale
                ref = np.abs(b) * scale + 1e-3
                assert np.median(d / ref) < 1e-4 and (d <= 2e-2 * ref).mean() > 0.99, (size, name, np.median(d / ref))
            print(f"grad {size}: reverse kernels agree with forward mode", flush=True)
    elif fam in ("tc2", "tc"):
        for size in ("xxlarge", "xxxlarge"):
            for variant in ("tensor", "tensor_fp32"):
                n = 2 * 148 * 32 + 37 if fam == "tc2" else 1000
                raw = synthetic.sample_training_distribution(n, seed=7)
                v = variant + (":one_sm" if fam == "tc" else "")
                check(f"{fam} {size} {variant}", ev.evaluate_host(raw, size, out_dtype=np.float64, variant=v), raw, size, variant)
torch.cuda.synchronize()
ev.close()
print("SANITIZER_TARGET_OK", flush=True)
