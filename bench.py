#!/usr/bin/env python
"""
bench.py -- throughput of the learned-core evaluation (BASELINE.json metric: airfoil-queries/sec at
xxxlarge) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--model-size xxxlarge] [--queries-per-gpu 10000000]

One "step" = one pass of the hot path over the whole per-GPU batch (BASELINE.json configs[2]: random
Kulfan airfoils sampled over the training distribution, full boundary-layer outputs).  Weak scaling:
every rank evaluates `queries_per_gpu` queries of its own (different seed per rank); there is no
collective on the data path, only the barrier around the timed region.

Printed JSON (rank 0, one line):
  value        whole-job queries/s, inputs and outputs resident in HBM, CUDA-event timed, max over ranks
  e2e          same metric through the C ABI's host entry point (nfb_eval_host): pinned HOST inputs and
               outputs, H2D + kernel + D2H inside the timed region
  roofline     the dominant (only) kernel against the fp32 FFMA peak measured on this GPU by
               nfb_measure_fp32_peak (MEASURED_PEAKS.json has no fp32 entry), HBM fraction beside it
  cpu_baseline the oracle (NumPy port of the reference path) on this box's host cores, bounded sample

`--impl reference` times that CPU path alone (rank 0 only under torchrun).
Nothing here reads /root/reference.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

from neuralfoil_b200 import synthetic  # noqa: E402

# Algorithmic work per query.
BYTES_PER_QUERY = 23 * 4 + 198 * 4  # fp32 rows in, fp32 rows out
FLOP_PER_QUERY = {  # 2 twins x 2 x sum(in * out) over the layers, unpadded
    "xxsmall": 52_032, "xsmall": 61_248, "small": 89_856, "medium": 106_240,
    "large": 310_784, "xlarge": 376_320, "xxlarge": 1_276_928, "xxxlarge": 5_699_584,
}
# DRAM traffic per query of the two xxxlarge kernels, from one `ncu --set full` capture each of this file's own
# command at 10 M queries (profiles/r1_ncu_full_xxxlarge_ffma_groups.txt / _tensor_pair_kernels.txt: dram__bytes_read.sum + dram__bytes_write.sum
# = 8.900 GB and 8.805 GB per launch; algorithmic 8.840 GB).  Reported as roofline.traffic, scaled to the run's n.
NCU_TRAFFIC_BYTES_PER_QUERY = {("xxxlarge", "fp32"): 890.0, ("xxxlarge", "tensor"): 880.4, ("xxxlarge", "tensor_fp32"): 882.9}
METRIC = "airfoil_queries_per_sec"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--model-size", default="xxxlarge")
    ap.add_argument("--variant", choices=("fp32", "tensor", "tensor_fp32"), default="fp32",
                    help="kernel behind value/e2e: fp32 FFMA (reference accuracy), tcgen05 one-term fp16 (TF32 class) "
                         "or tcgen05 three-term fp16 split (fp32 class); the tensor kernels exist for xxlarge / xxxlarge")
    ap.add_argument("--no-tensor-extra", action="store_true", help="skip the extra tensor-variant measurement")
    ap.add_argument("--queries-per-gpu", type=int, default=10_000_000)
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def extra_models():
    if (REPO / "neuralfoil_b200" / "nn_weights_and_biases" / "nn-xxxlarge.npz").exists():
        return {}, "real nn-xxxlarge.npz"
    return {"xxxlarge": synthetic.synthetic_xxxlarge_params()}, "synthetic xxxlarge weights (seed 20261019; reference ships none)"


def workload_config(args, weights_note: str) -> dict:
    return {
        "workload": "BASELINE.json configs[2]: random Kulfan airfoils over the training distribution, full BL outputs",
        "model_size": args.model_size,
        "variant": getattr(args, "variant", "fp32"),
        "queries_per_gpu": args.queries_per_gpu,
        "weights": weights_note if args.model_size == "xxxlarge" else "shipped .npz",
        "layout": "SoA fp32: 23 input rows, 198 output rows",
        "l2_note": "inputs (92 B/query) and outputs (792 B/query) per step exceed the 126 MB L2; no flush needed",
    }


# ---------------------------------------------------------------------------------------------
# CPU reference arm (the oracle port of the reference's NumPy path)
# ---------------------------------------------------------------------------------------------
def cpu_reference_qps(model_size: str, sample: int, steps: int, warmup: int):
    from oracle.neuralfoil_oracle import Oracle  # bench.py's cpu_baseline / reference arm only

    orc = Oracle(extra_models=extra_models()[0])
    raw = synthetic.sample_training_distribution(sample, seed=12345)
    for _ in range(warmup):
        orc.evaluate_raw(raw[:, : max(1, sample // 10)], model_size)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        orc.evaluate_raw(raw, model_size)
        times.append(time.perf_counter() - t0)
    return sample / (sum(times) / len(times)), sum(times) / len(times)


def auto_cpu_sample(model_size: str) -> int:
    # about 10 s of CPU work per pass at the survey's measured rates
    approx_qps = {"xxsmall": 116e3, "xsmall": 101e3, "small": 87e3, "medium": 77e3, "large": 48e3,
                  "xlarge": 45e3, "xxlarge": 24e3, "xxxlarge": 8.3e3}[model_size]
    return int(min(400_000, max(20_000, 8 * approx_qps)))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_sample or auto_cpu_sample(args.model_size) // 2
    qps, sec = cpu_reference_qps(args.model_size, sample, args.steps, min(args.warmup, 1))
    cores = os.cpu_count()
    line = {
        "impl": "reference",
        "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, extra_models()[1]),
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} queries of the same workload per step (NumPy float64, BLAS threads = all cores)"},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, sm_max, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); sm_max.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(sm_max), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    from neuralfoil_b200 import Evaluator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1 (one process per GPU)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    models, weights_note = extra_models()
    ev = Evaluator(device=local_rank, extra_models=models)
    n = args.queries_per_gpu
    size = args.model_size

    # ---- inputs: sampled on the host once, pinned (for e2e) and resident on the device (for value) ----
    raw_host = torch.empty((23, n), dtype=torch.float32).pin_memory()
    raw_host.numpy()[...] = synthetic.sample_training_distribution(n, seed=1000 + rank, dtype=np.float32)
    raw_dev = raw_host.cuda(non_blocking=True)
    ld = (n + 3) & ~3
    out_dev = torch.empty((198, ld), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()

    fp32_peak = ev.measure_fp32_peak_tflops(5) if rank == 0 else 0.0

    # ---- device-resident throughput ----
    variant = args.variant
    for _ in range(max(args.warmup, 3)):
        ev.evaluate_device(raw_dev, size, out=out_dev, variant=variant)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ev.launch_count()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_begin.record()
    for i in range(args.steps):
        starts[i].record()
        ev.evaluate_device(raw_dev, size, out=out_dev, variant=variant)
        stops[i].record()
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    launches = ev.launch_count() - launches0
    total_ms = max_over_ranks(t_begin.elapsed_time(t_end))
    kernel_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3)
    checksum = float(out_dev[:6, :1000].double().sum().item())  # touch the result; also a sanity value

    # ---- end to end through the C ABI's host entry point ----
    e2e = None
    if not args.no_e2e:
        out_host = torch.empty((198, ld), dtype=torch.float32).pin_memory()
        for _ in range(1):
            ev.evaluate_host_into(raw_host.numpy(), size, out_host.numpy(), variant=variant)
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            ev.evaluate_host_into(raw_host.numpy(), size, out_host.numpy(), variant=variant)
            _ = float(out_host[1, 0])  # read a result on the host
        torch.cuda.synchronize()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        e2e = {"value": world * n / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "h2d_bytes_per_step": 23 * 4 * n * world, "d2h_bytes_per_step": 198 * 4 * n * world,
               "api": "nfb_eval_host (C ABI), pinned host buffers, chunked H2D/kernel/D2H overlap"}
        host_check = float(out_host[:6, :1000].double().sum().item())
        assert abs(host_check - checksum) <= 1e-6 * max(1.0, abs(checksum)), (host_check, checksum)

    # ---- the other variant, briefly (xxxlarge only): reported beside the headline, never mixed into it ----
    def measure_other(other: str) -> dict:
        for _ in range(3):
            ev.evaluate_device(raw_dev, size, out=out_dev, variant=other)
        barrier()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        for _ in range(args.steps):
            ev.evaluate_device(raw_dev, size, out=out_dev, variant=other)
        t1e.record()
        barrier()
        t_ms = max_over_ranks(t0e.elapsed_time(t1e)) / args.steps
        extra = {"ms_per_step": t_ms, "value": world * n / (t_ms * 1e-3)}
        if not args.no_e2e:
            ev.evaluate_host_into(raw_host.numpy(), size, out_host.numpy(), variant=other)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                ev.evaluate_host_into(raw_host.numpy(), size, out_host.numpy(), variant=other)
                _ = float(out_host[1, 0])
            torch.cuda.synchronize()
            te = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
            extra["e2e"] = {"value": world * n / te, "unit": UNIT, "ms_per_step": te * 1e3,
                            "note": "same nfb_eval_host path (pinned host in/out); PCIe device->host moves 792 B/query"}
        return extra

    # ---- outputs="scalars" (NFB_OUTPUT_SCALARS), briefly: only the six scalar rows travel back over PCIe ----
    def measure_scalars(v: str) -> dict:
        out6_dev = torch.empty((6, ld), dtype=torch.float32, device="cuda")
        for _ in range(2):
            ev.evaluate_device(raw_dev, size, out=out6_dev, variant=v, outputs="scalars")
        barrier()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        for _ in range(args.steps):
            ev.evaluate_device(raw_dev, size, out=out6_dev, variant=v, outputs="scalars")
        t1e.record()
        barrier()
        t_ms = max_over_ranks(t0e.elapsed_time(t1e)) / args.steps
        res = {"value": world * n / (t_ms * 1e-3), "unit": UNIT, "ms_per_step": t_ms, "d2h_bytes_per_query": 24}
        if not args.no_e2e:
            out6_host = torch.empty((6, ld), dtype=torch.float32).pin_memory()
            ev.evaluate_host_into(raw_host.numpy(), size, out6_host.numpy(), variant=v, outputs="scalars")
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                ev.evaluate_host_into(raw_host.numpy(), size, out6_host.numpy(), variant=v, outputs="scalars")
                _ = float(out6_host[1, 0])
            torch.cuda.synchronize()
            te = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
            res["e2e"] = {"value": world * n / te, "unit": UNIT, "ms_per_step": te * 1e3}
        return res

    tensor_extra = tensor_fp32_extra = None
    if size in ("xxxlarge", "xxlarge") and variant == "fp32" and not args.no_tensor_extra:
        tensor_fp32_extra = measure_other("tensor_fp32")
        tensor_extra = measure_other("tensor")
        if not args.no_e2e:
            tensor_fp32_extra["scalars_only"] = measure_scalars("tensor_fp32")
            tensor_extra["scalars_only"] = measure_scalars("tensor")

    # ---- the Jacobian kernel, briefly: outputs + 198 x 23 derivatives per query, device-resident ----
    jacobian_extra = None
    if variant == "fp32" and not args.no_tensor_extra:
        nj = min(n, 20_000)
        rj = raw_dev[:, :nj].contiguous()
        for _ in range(2):
            ev.evaluate_jacobian_device(rj, size)
        barrier()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        for _ in range(3):
            ev.evaluate_jacobian_device(rj, size)
        t1e.record()
        barrier()
        t_ms = max_over_ranks(t0e.elapsed_time(t1e)) / 3
        jf = 24 * FLOP_PER_QUERY[size] * nj / (t_ms * 1e-3) / 1e12
        jacobian_extra = {"value": world * nj / (t_ms * 1e-3), "unit": "jacobians/s", "queries_per_launch": nj, "ms_per_launch": t_ms,
                          "kernel": "nfb_jac_kernel<H> (forward mode, fp32 FFMA: the query, its twin and 23 tangents each)",
                          "achieved_tflops": jf, "note": "24 x the value path's FLOP per query; one CTA per query"}

    # ---- the reverse-mode kernel: gradients of an objective of the six scalars, device-resident ----
    scalar_grad_extra = None
    if variant == "fp32" and not args.no_tensor_extra:
        try:  # an extra: it must never cost the headline line
            ng = min(n, 2_000_000)
            rg = raw_dev[:, :ng].contiguous()
            cot = torch.randn((6, ng), dtype=torch.float32, device="cuda")
            for _ in range(2):
                ev.evaluate_scalar_grad_device(rg, cot, size)
            barrier()
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0e.record()
            for _ in range(3):
                ev.evaluate_scalar_grad_device(rg, cot, size)
            t1e.record()
            barrier()
            t_ms = max_over_ranks(t0e.elapsed_time(t1e)) / 3
            scalar_grad_extra = {"value": world * ng / (t_ms * 1e-3), "unit": "gradients/s", "queries_per_launch": ng, "ms_per_launch": t_ms,
                                 "kernel": "nfb_grad_kernel<., false> (fused forward + reverse pass, fp32 FFMA2; six scalar values + d objective / d 23 raw inputs per query)"}
            del rg, cot
        except Exception as e:  # noqa: BLE001
            scalar_grad_extra = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        peaks_path = REPO / "MEASURED_PEAKS.json"
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        if peaks_path.exists():
            hbm_peak, hbm_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        k_ms = float(np.mean(kernel_ms))
        achieved_tflops = FLOP_PER_QUERY[size] * n / (k_ms * 1e-3) / 1e12
        achieved_gbs = BYTES_PER_QUERY * n / (k_ms * 1e-3) / 1e9
        bf16_peak, bf16_src = 1590.0, "fallback (B200_PROFILING.md)"
        if peaks_path.exists():
            bf16_peak, bf16_src = float(json.loads(peaks_path.read_text())["bf16_tflops"]), "measured burst (MEASURED_PEAKS.json)"
        if tensor_extra is not None:
            tf = FLOP_PER_QUERY[size] * n / (tensor_extra["ms_per_step"] * 1e-3) / 1e12  # per GPU: n is per rank
            tensor_extra.update({
                "unit": UNIT, "kernel": "nfb_tc2_kernel<H, 1> (tcgen05 cta_group::2 kind::f16 on CTA pairs, fp16 operands, fp32 accumulate in TMEM)",
                "accuracy": "TF32-class: median relative error 3e-4; tests/parity.py 'tensor' tolerance",
                "roofline": {"bound": "tensor", "achieved": tf, "peak": bf16_peak, "unit": "TFLOP/s", "frac": tf / bf16_peak,
                             "peak_source": bf16_src}})
        if tensor_fp32_extra is not None:
            tf = 3 * FLOP_PER_QUERY[size] * n / (tensor_fp32_extra["ms_per_step"] * 1e-3) / 1e12
            tensor_fp32_extra.update({
                "unit": UNIT, "kernel": "nfb_tc2_kernel<H, 3> (tcgen05 cta_group::2 on CTA pairs; three fp16 MMAs per product: a_hi w_hi + a_lo w_hi + a_hi w_lo)",
                "accuracy": "fp32-class worst case (fp32 tolerance), median relative error 2e-6 .. 4e-6; tests/parity.py 'tensor_fp32'",
                "roofline": {"bound": "tensor", "achieved": tf, "peak": bf16_peak, "unit": "TFLOP/s", "frac": tf / bf16_peak,
                             "peak_source": bf16_src,
                             "note": "achieved counts the 3 tensor-core MACs issued per algorithmic MAC; the 512-wide tile is 64 rows per CTA, "
                                     "run at full rate as half of a pair-wide M = 128 MMA"}})
        roofline = {
            "bound": "fp32",
            "kernel": "nfb_ffma_groups_kernel",
            "achieved": achieved_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": achieved_tflops / fp32_peak if fp32_peak else None,
            "peak_source": "measured on this GPU by nfb_measure_fp32_peak (register-resident FFMA loop); "
                           "MEASURED_PEAKS.json has no fp32 figure; nominal 148 SM x 128 x 2 x 1.965 GHz = 74.4",
            "frac_of_nominal_74.4": achieved_tflops / 74.4,
            "kernel_ms_per_launch": k_ms,
            "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                    "peak_source": hbm_src},
            "traffic": None,
        }
        per_q = NCU_TRAFFIC_BYTES_PER_QUERY.get((size, variant))
        if per_q is not None:
            roofline["traffic"] = per_q * n
            roofline["traffic_note"] = ("bytes per launch = ncu-measured DRAM bytes per query (10 M-query capture, profiles/) x this run's "
                                        f"queries per launch; algorithmic {BYTES_PER_QUERY * n}")
        if variant != "fp32":
            mult = 3 if variant == "tensor_fp32" else 1
            roofline.update({"bound": "tensor", "kernel": "nfb_tc2_kernel", "peak": bf16_peak, "peak_source": bf16_src,
                             "achieved": mult * achieved_tflops, "frac": mult * achieved_tflops / bf16_peak,
                             "tensor_macs_per_algorithmic_mac": mult})
            roofline.pop("frac_of_nominal_74.4", None)
        cpu_baseline = None
        if not args.no_cpu_baseline and world == 1:
            sample = args.cpu_sample or auto_cpu_sample(size)
            qps, sec = cpu_reference_qps(size, sample, 1, 1)
            cpu_baseline = {"value": qps, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                            "sample": f"{sample} queries of the same workload, one pass, {sec:.1f} s "
                                      "(oracle: NumPy float64 port of the reference path, BLAS threads = all cores)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32", "tensor": "f16 operands / f32 accumulate",
                                                       "tensor_fp32": "2 x f16 split operands / f32 accumulate"}[variant],
            "data": "synthetic", "config": workload_config(args, weights_note),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "tensor_variant": tensor_extra,
            "tensor_fp32_variant": tensor_fp32_extra, "jacobian": jacobian_extra, "scalar_gradient": scalar_grad_extra,
            "checksum_first_1000": checksum,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
