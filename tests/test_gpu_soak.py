"""
Soak test of the CTA-pair tensor kernels (nfb_tc2.cuh), whose cross-CTA hand-offs use default (CTA-scope) mbarrier
semantics on purpose ("barrier protocol").  A query's arithmetic does not depend on which CTA pair
evaluates its tile, on how many pairs run, or on what else the memory system is doing -- so any run of the same batch must
give the same BITS.  A hand-off that let a tensor core read activations or weights before they had landed would show up
as a difference in some tile of some run.  More than 1e9 queries pass through the pair kernels here, on full and on odd
grid sizes (NFB_TUNE_TC2_MAX_GRID), with concurrent host<->device copy traffic and a second stream hammering HBM; every
output element of every run is compared with the reference run, and the reference run with the one-SM kernels
(nfb_tc.cuh: another K order, hence equal to the variant's tolerance, not bit for bit) and the oracle.
"""

import numpy as np
import pytest

from neuralfoil_b200 import synthetic

import parity

pytestmark = pytest.mark.gpu

CASES = [  # (model, variant, queries per launch, launches) -> 4 x 256 M = 1.02e9 queries through the pair kernels
    ("xxlarge", "tensor", 4_000_000, 64),
    ("xxlarge", "tensor_fp32", 4_000_000, 64),
    ("xxxlarge", "tensor", 2_000_000, 128),
    ("xxxlarge", "tensor_fp32", 2_000_000, 128),
]
GRIDS = [0, 146, 74, 38, 6, 2, 100, 148]  # 0 = every SM; the kernel rounds down to whole pairs


@pytest.mark.parametrize("size,variant,n,launches", CASES)
def test_pair_kernels_are_bit_reproducible_under_load(evaluator, oracle, size, variant, n, launches):
    import torch

    raw = torch.from_numpy(synthetic.sample_training_distribution(n, seed=2026, dtype=np.float32)).cuda()
    ref = evaluator.evaluate_device(raw, size, variant=variant).clone()
    torch.cuda.synchronize()
    # the reference run itself: against the one-SM kernels and the oracle on a subsample
    idx = torch.arange(0, n, n // 4000, device="cuda")[:4000]
    one_sm = evaluator.evaluate_device(raw[:, idx].contiguous(), size, variant=variant + ":one_sm")
    want = oracle.evaluate_raw(raw[:, idx].double().cpu().numpy(), size)
    re = raw[19, idx].double().cpu().numpy()
    parity.assert_parity(ref[:, idx].double().cpu().numpy(), want, re, variant=variant, label=f"soak {size} {variant}")
    parity.assert_parity(one_sm.double().cpu().numpy(), want, re, variant=variant, label=f"soak one-SM {size} {variant}")

    out = torch.empty_like(ref)
    side = torch.cuda.Stream()
    pinned = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    dev_buf = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    big_a = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    big_b = torch.empty_like(big_a)
    mismatches = 0
    try:
        for i in range(launches):
            evaluator.set_tuning(tc2_max_grid=GRIDS[i % len(GRIDS)])
            with torch.cuda.stream(side):  # copy traffic in both directions and an HBM-bound kernel beside the launch
                dev_buf.copy_(pinned, non_blocking=True)
                pinned.copy_(dev_buf, non_blocking=True)
                big_b.copy_(big_a, non_blocking=True)
            evaluator.evaluate_device(raw, size, out=out, variant=variant)
            mismatches += int((out.view(torch.int32) != ref.view(torch.int32)).sum().item())  # bits, NaN-safe
            if mismatches:
                break
    finally:
        evaluator.set_tuning(tc2_max_grid=0)
        torch.cuda.synchronize()
    assert mismatches == 0, f"{size} {variant}: {mismatches} elements differ from the reference run (launch {i}, grid {GRIDS[i % len(GRIDS)]})"
