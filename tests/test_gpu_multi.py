"""
The product's multi-GPU path and its concurrency contract, on real devices (through the C ABI):
  * nfb_multi_eval_host == nfb_eval_host bit for bit -- with every visible device, with an explicit device list, and with
    one device (the same sharding code path, runnable on a one-GPU box);
  * `devices="auto"` brings the other devices in for large batches only;
  * SharedHostBlock: ranks of a one-process-per-GPU job land their shards in one page-locked shared block;
  * two contexts on one device keep their own distribution statistics (they travel in the launch parameters);
  * Python threads sharing one Evaluator, and reverse-mode launches on two streams of one context, give the serial results.
"""

import threading

import numpy as np
import pytest

from neuralfoil_b200 import Evaluator, sharding, synthetic

import parity

pytestmark = pytest.mark.gpu


def _device_count() -> int:
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("n", [1, 3, 4, 5, 1000, 70_001, 300_000])
def test_multi_device_entry_equals_single_device_bit_for_bit(evaluator, n):
    raw = synthetic.sample_training_distribution(n, seed=700 + n, dtype=np.float32)
    want = evaluator.evaluate_host(raw, "large")
    specs = [[0], "all"] + ([[1, 0]] if _device_count() >= 2 else [])
    for devices in specs:
        ev = Evaluator(devices=devices, multi_threshold=0)  # every batch goes through nfb_multi_eval_host
        try:
            assert ev.n_devices == (len(devices) if isinstance(devices, list) else _device_count())
            l0 = ev.launch_count()
            got = ev.evaluate_host(raw, "large")
            assert np.array_equal(got, want, equal_nan=True), (devices, n)
            # every device that owns a non-empty shard launched exactly once (small batches: one kernel per shard)
            shards = [sharding.shard_range(n, r, ev.n_devices) for r in range(ev.n_devices)]
            if n <= 16384:
                assert ev.launch_count() - l0 == sum(1 for s, e in shards if e > s)
            got64 = ev.evaluate_host(raw.astype(np.float64), "large", out_dtype=np.float64, outputs="scalars")
            assert np.array_equal(got64, want[:6].astype(np.float64), equal_nan=True)
        finally:
            ev.close()


def test_multi_device_dict_api_pageable_and_pinned(evaluator, oracle):
    """The drop-in call on every device: NumPy (pageable) arrays in, the reference's dict out; and a caller-owned block."""
    n = 400_000
    raw = synthetic.sample_training_distribution(n, seed=31)
    kulfan = {"upper_weights": raw[0:8], "lower_weights": raw[8:16], "leading_edge_weight": raw[16], "TE_thickness": raw[17]}
    ev = Evaluator(devices="all", multi_threshold=1 << 16)
    try:
        aero = ev.get_aero_from_kulfan_parameters(kulfan, raw[18], raw[19], raw[20], raw[21], raw[22], model_size="medium")
        one = evaluator.get_aero_from_kulfan_parameters(kulfan, raw[18], raw[19], raw[20], raw[21], raw[22], model_size="medium")
        assert list(aero) == list(one)
        for k in aero:
            assert aero[k].shape == (n,) and aero[k].dtype == np.float64
            assert np.array_equal(aero[k], one[k], equal_nan=True), k
        idx = np.random.default_rng(0).choice(n, 2000, replace=False)
        got = np.stack([aero[k][idx] for k in aero])
        parity.assert_parity(got, oracle.evaluate_raw(raw[:, idx], "medium"), raw[19, idx], label="multi-device dict API")
        out = np.zeros((198, n + 8), dtype=np.float32)
        view = ev.evaluate_host_into(raw.astype(np.float32), "medium", out)
        assert view.shape == (198, n) and np.array_equal(view, evaluator.evaluate_host(raw.astype(np.float32), "medium"))
        assert not out[:, n:].any()  # nothing written beyond the batch
    finally:
        ev.close()


def test_auto_devices_join_for_large_batches_only():
    ev = Evaluator(device=0, devices="auto", multi_threshold=50_000)
    try:
        small = synthetic.sample_training_distribution(1000, seed=1, dtype=np.float32)
        big = synthetic.sample_training_distribution(120_000, seed=2, dtype=np.float32)
        a = ev.evaluate_host(small, "small")
        assert ev._auto_multi is None
        b = ev.evaluate_host(big, "small")
        if _device_count() >= 2:
            assert ev._auto_multi is not None and ev._auto_multi.n_devices == _device_count()
        else:
            assert ev._auto_multi is None and ev._auto is False  # a one-GPU box: nothing to bring in
        single = Evaluator(device=0)
        try:
            assert np.array_equal(a, single.evaluate_host(small, "small")) and np.array_equal(b, single.evaluate_host(big, "small"))
        finally:
            single.close()
    finally:
        ev.close()


def test_shared_host_block_receives_every_ranks_shard(evaluator):
    """What the ranks of a torchrun job do, here from one process: evaluate shard r straight into the shared pinned block."""
    n, world = 100_003, 4
    raw = synthetic.sample_training_distribution(n, seed=9, dtype=np.float32)
    want = evaluator.evaluate_host(raw, "xlarge")
    ld = (n + 3) & ~3
    owner = sharding.SharedHostBlock(198, ld, np.float32, create=True)
    try:
        for rank in range(world):
            peer = sharding.SharedHostBlock(198, ld, np.float32, name=owner.name)  # a second mapping, as another process has
            try:
                s, e = sharding.shard_range(n, rank, world)
                evaluator.evaluate_host_into(np.ascontiguousarray(raw[:, s:e]), "xlarge", peer.shard(rank, world, n))
            finally:
                peer.close()
        assert np.array_equal(owner.array[:, :n], want)
    finally:
        owner.close()
        owner.unlink()


def test_two_contexts_on_one_device_keep_their_own_distribution(tmp_path, evaluator, oracle):
    import shutil

    from neuralfoil_b200.main import nn_weights_dir

    alt = tmp_path / "weights"
    alt.mkdir()
    shutil.copy(nn_weights_dir / "nn-small.npz", alt / "nn-small.npz")
    with np.load(nn_weights_dir / "scaled_input_distribution.npz") as f:
        mean, inv_cov = f["mean_inputs_scaled"].copy(), f["inv_cov_inputs_scaled"].copy()
    np.savez(alt / "scaled_input_distribution.npz", mean_inputs_scaled=mean + 0.05, inv_cov_inputs_scaled=inv_cov * 2.0)
    raw = synthetic.sample_training_distribution(3000, seed=4)
    before = evaluator.evaluate_host(raw, "small", out_dtype=np.float64)
    other = Evaluator(device=0, weights_dir=alt)
    try:
        shifted = other.evaluate_host(raw, "small", out_dtype=np.float64)
        after = evaluator.evaluate_host(raw, "small", out_dtype=np.float64)
        assert np.array_equal(before, after), "the second context's distribution leaked into the first"
        assert np.array_equal(shifted[1:], before[1:])              # only analysis_confidence uses the distribution
        assert np.abs(shifted[0] - before[0]).max() > 1e-3
    finally:
        other.close()


def test_python_threads_may_share_one_evaluator(evaluator):
    cases = [(size, synthetic.sample_training_distribution(n, seed=40 + i))
             for i, (size, n) in enumerate([("xlarge", 1), ("xxsmall", 700), ("large", 20_000), ("medium", 64), ("xlarge", 5000)] * 2)]
    serial = [evaluator.evaluate_host(raw, size, out_dtype=np.float64) for size, raw in cases]
    results, errors = [None] * len(cases), []

    def work(i):
        try:
            for _ in range(5):
                size, raw = cases[i]
                results[i] = evaluator.evaluate_host(raw, size, out_dtype=np.float64)
                jv, _ = evaluator.evaluate_jacobian_host(raw[:, :3], size)
                assert np.array_equal(jv, serial[i][:, :3]) or raw.shape[1] < 3
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for got, want in zip(results, serial):
        assert np.array_equal(got, want, equal_nan=True)


def test_reverse_mode_launches_on_two_streams_do_not_share_scratch(evaluator):
    import torch

    n = 60_000
    raws = [torch.from_numpy(synthetic.sample_training_distribution(n, seed=s, dtype=np.float32)).cuda() for s in (1, 2)]
    cots = [torch.randn((6, n), device="cuda", generator=torch.Generator(device="cuda").manual_seed(s)) for s in (3, 4)]
    serial = []
    for r, c in zip(raws, cots):
        serial.append(evaluator.evaluate_scalar_grad_device(r, c, "xlarge")[1].clone())
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for _ in range(4):
        outs = []
        for st, r, c in zip(streams, raws, cots):
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                outs.append(evaluator.evaluate_scalar_grad_device(r, c, "xlarge")[1])
        torch.cuda.synchronize()
        for got, want in zip(outs, serial):
            assert torch.equal(got, want)


def test_forward_and_reverse_vjp_ignore_rows_with_zero_cotangent(evaluator):
    """Re so small that theta overflows to inf: with zero cotangents on the theta rows both VJP kernels stay finite and agree."""
    import torch

    n = 300  # >= 256: evaluate_vjp_device would pick the reverse kernel by itself
    raw = synthetic.sample_training_distribution(n, seed=77)
    raw[19] = 1e-40
    cot = np.zeros((198, n))
    cot[1], cot[2] = 1.0, -30.0
    cot[parity.ROW_UE] = 0.05
    r, c = torch.from_numpy(raw).cuda(), torch.from_numpy(cot).cuda()
    v_f, g_f = evaluator.evaluate_vjp_device(r, c, "xlarge", reverse=False)
    v_r, g_r = evaluator.evaluate_vjp_device(r, c, "xlarge", reverse=True)
    torch.cuda.synchronize()
    assert torch.isinf(v_f[6:38]).any(), "the case is meant to overflow theta"
    rows = [t for t in range(23) if t != 19]  # d/dRe itself scales with 1/Re = 1e40
    assert torch.isfinite(g_f[rows]).all() and torch.isfinite(g_r[rows]).all()
    scale = torch.from_numpy(parity.JAC_INPUT_SCALE[rows]).cuda()[:, None]
    ref = (g_f[rows].abs() * scale).amax() + 1e-3
    assert ((g_f[rows] - g_r[rows]).abs() * scale).max() <= 5e-3 * ref


def test_out_argument_is_validated(evaluator):
    import torch

    raw = torch.from_numpy(synthetic.sample_training_distribution(64, seed=1, dtype=np.float32)).cuda()
    for bad in (torch.empty((198, 64)), torch.empty((198, 64), dtype=torch.float16, device="cuda"),
                torch.empty((197, 64), device="cuda")):
        with pytest.raises(ValueError):
            evaluator.evaluate_device(raw, "small", out=bad)
    with pytest.raises(ValueError):
        evaluator.evaluate_jacobian_device(raw.cpu(), "small")
    with pytest.raises(ValueError):
        evaluator.evaluate_scalar_grad_device(raw, torch.zeros((6, 64)), "small")
    if _device_count() >= 2:
        with pytest.raises(ValueError):
            evaluator.evaluate_device(raw, "small", out=torch.empty((198, 64), device="cuda:1"))


@pytest.mark.parametrize("n", [1, 5, 64, 1000, 1024])
def test_zero_copy_small_batches_equal_copied_ones(evaluator, n):
    """nfb_eval_host runs the smallest batches zero-copy (the kernel reads / writes the pinned staging block): same bits."""
    raw = synthetic.sample_training_distribution(n, seed=900 + n)
    try:
        evaluator.set_tuning(mapped_max=0)
        copied = [evaluator.evaluate_host(raw, size, out_dtype=dt) for size in ("xxsmall", "xlarge") for dt in (np.float32, np.float64)]
        evaluator.set_tuning(mapped_max=4096)
        mapped = [evaluator.evaluate_host(raw, size, out_dtype=dt) for size in ("xxsmall", "xlarge") for dt in (np.float32, np.float64)]
    finally:
        evaluator.set_tuning(mapped_max=-1)
    for a, b in zip(copied, mapped):
        assert a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("size,variant", [("xxsmall", "fp32"), ("xlarge", "fp32"), ("xlarge", "tensor"), ("xxlarge", "tensor_fp32"),
                                          ("xxxlarge", "tensor"), ("xxxlarge", "fp32")])
def test_mixed_precision_outputs_are_the_bf16_rounding_of_the_fp32_ones(evaluator, size, variant):
    """nfb_eval_*_mixed: scalars bit-identical, boundary-layer rows = the float32 results rounded to nearest bfloat16."""
    import torch

    from neuralfoil_b200 import bf16_bits_to_float32

    for n in (37, 3000, 70_001):  # small-batch kernel, one chunk, ragged
        raw = synthetic.sample_training_distribution(n, seed=n, dtype=np.float32)
        d_raw = torch.from_numpy(raw).cuda()
        full = evaluator.evaluate_device(d_raw, size, variant=variant)
        scal, bl = evaluator.evaluate_device_mixed(d_raw, size, variant=variant)
        torch.cuda.synchronize()
        assert scal.dtype == torch.float32 and bl.dtype == torch.bfloat16 and bl.shape == (192, n)
        assert torch.equal(scal, full[:6])
        assert torch.equal(bl.view(torch.int16), full[6:].to(torch.bfloat16).view(torch.int16))  # torch rounds to nearest even too
        ld = (n + 3) & ~3
        h_scal, h_bl = np.zeros((6, ld), dtype=np.float32), np.zeros((192, ld), dtype=np.uint16)
        a, b = evaluator.evaluate_host_mixed_into(raw, size, h_scal, h_bl, variant=variant, chunk=16384)
        assert np.array_equal(a, scal.cpu().numpy(), equal_nan=True)
        want = full[6:].cpu().numpy()
        got = bf16_bits_to_float32(b)
        assert np.array_equal(got, bl.float().cpu().numpy(), equal_nan=True)
        finite = np.isfinite(want) & (np.abs(want) > 1e-30)
        assert np.max(np.abs(got[finite] - want[finite]) / np.abs(want[finite])) <= 2.0 ** -8
    with pytest.raises(ValueError):
        evaluator.evaluate_host_mixed_into(raw, size, h_scal, h_bl.astype(np.float32), variant=variant)
