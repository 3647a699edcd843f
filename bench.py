#!/usr/bin/env python
"""
bench.py -- throughput of the learned-core evaluation (BASELINE.json metric: airfoil-queries/sec at
xxxlarge) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--scaling strong|weak] [--queries 10000000] [--model-size xxxlarge] [--variant fp32]

One "step" = one pass of the hot path over the whole batch (BASELINE.json configs[2]: 10 M random Kulfan airfoils
sampled over the training distribution, xxxlarge, full boundary-layer outputs, on 1/2/4/8 B200).

Scaling.  `strong` (default; configs[2] as written): ONE batch of `--queries` queries, generated once by rank 0 into a
page-locked shared host block that every rank of the box maps; rank r owns the contiguous range
`sharding.shard_range(n, r, N)` of it.  `weak`: `--queries` queries per rank (the batch grows with N).  There is no
collective on the data path either way; NCCL carries only the barriers and the max-over-ranks of the timings.

Printed JSON (rank 0, one line):
  value        whole-job queries/s, each rank's shard resident in its GPU's HBM, CUDA-event timed, max over ranks
  e2e          the same batch through the product's host entry point (nfb_eval_host behind `Evaluator.evaluate_host_into`):
               every rank reads its input range from the shared pinned block and its device->host copies land in its
               column range of ONE shared pinned (198, n) result block -- the gather of SURVEY 8(e) -- then a barrier,
               then rank 0 reads a value of every shard.  H2D + kernel + D2H + gather inside the timed region.
  roofline     the dominant (only) kernel against the fp32 FFMA peak measured on this GPU by nfb_measure_fp32_peak
               (MEASURED_PEAKS.json has no fp32 entry), HBM fraction beside it; `traffic` from an ncu side-run of the same
               kernel at the same size after the timed region (N = 1; `--no-ncu` skips it)
  cpu_baseline the oracle (NumPy port of the reference path) on this box's host cores, bounded sample

`--impl reference` times that CPU path alone (rank 0 only under torchrun), with every host core.
`python bench.py --gpus N` WITHOUT torchrun runs the single-process product path (`Evaluator(devices=[0..N-1])`,
nfb_multi_eval_host: one host thread per device).  Nothing here reads /root/reference.
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

# The CPU legs (reference arm, cpu_baseline) use every host core whatever the launcher exported: torchrun sets
# OMP_NUM_THREADS=1 for its children, which NumPy's BLAS reads once, at import.
if "reference" in sys.argv or int(os.environ.get("WORLD_SIZE", "1")) == 1:
    for _var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        if "reference" in sys.argv or _var not in os.environ:
            os.environ[_var] = str(os.cpu_count())

import numpy as np  # noqa: E402

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

from neuralfoil_b200 import sharding, synthetic  # noqa: E402

# Algorithmic work per query.
BYTES_PER_QUERY = 23 * 4 + 198 * 4  # fp32 rows in, fp32 rows out
FLOP_PER_QUERY = {  # 2 twins x 2 x sum(in * out) over the layers, unpadded
    "xxsmall": 52_032, "xsmall": 61_248, "small": 89_856, "medium": 106_240,
    "large": 310_784, "xlarge": 376_320, "xxlarge": 1_276_928, "xxxlarge": 5_699_584,
}
# DRAM traffic per query of the xxxlarge kernels from the round-1 `ncu --set full` captures (profiles/r1_ncu_full_*):
# the fallback for roofline.traffic when the ncu side-run is skipped or unavailable.
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
                         "or tcgen05 three-term fp16 split (fp32 class)")
    ap.add_argument("--scaling", choices=("strong", "weak"), default="strong",
                    help="strong: one batch of --queries split over the GPUs (BASELINE configs[2]); weak: --queries per GPU")
    ap.add_argument("--queries", type=int, default=10_000_000)
    ap.add_argument("--queries-per-gpu", type=int, default=0, help="shorthand for --scaling weak --queries Q")
    ap.add_argument("--no-extras", "--no-tensor-extra", dest="no_extras", action="store_true",
                    help="skip the extra measurements (tensor variants, Jacobian, gradients)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ncu", action="store_true", help="do not measure roofline.traffic with an ncu side-run")
    args = ap.parse_args()
    if args.queries_per_gpu:
        args.scaling, args.queries = "weak", args.queries_per_gpu
    return args


def extra_models():
    if (REPO / "neuralfoil_b200" / "nn_weights_and_biases" / "nn-xxxlarge.npz").exists():
        return {}, "real nn-xxxlarge.npz"
    return {"xxxlarge": synthetic.synthetic_xxxlarge_params()}, "synthetic xxxlarge weights (seed 20261019; reference ships none)"


def workload_config(args, weights_note: str, n_total: int, world: int) -> dict:
    return {
        "workload": "BASELINE.json configs[2]: random Kulfan airfoils over the training distribution, full BL outputs",
        "model_size": args.model_size,
        "variant": getattr(args, "variant", "fp32"),
        "queries_total": n_total,
        "queries_per_gpu": -(-n_total // world),
        "sharding": "contiguous index ranges of one batch (sharding.shard_range == nfb_shard_range), no collective",
        "weights": weights_note if args.model_size == "xxxlarge" else "shipped .npz",
        "layout": "SoA fp32: 23 input rows, 198 output rows",
        "l2_note": "inputs (92 B/query) and outputs (792 B/query) per step exceed the 126 MB L2; no flush needed",
    }


# ---------------------------------------------------------------------------------------------
# CPU reference arm (the oracle port of the reference's NumPy path)
# ---------------------------------------------------------------------------------------------
def blas_threads() -> int:
    """Threads NumPy's BLAS will really use (threadpoolctl asks the loaded library, not the environment)."""
    try:
        from threadpoolctl import threadpool_info

        pools = [p.get("num_threads", 0) for p in threadpool_info() if p.get("user_api") == "blas"]
        if pools:
            return int(max(pools))
    except Exception:  # noqa: BLE001
        pass
    return int(os.environ.get("OMP_NUM_THREADS", os.cpu_count()))


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
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    sample = args.cpu_sample or auto_cpu_sample(args.model_size) // 2
    qps, sec = cpu_reference_qps(args.model_size, sample, args.steps, min(args.warmup, 1))
    threads = blas_threads()
    n_total = args.queries * (world if args.scaling == "weak" else 1)
    config = workload_config(args, extra_models()[1], n_total, world)
    config["reference_arm_sample"] = (f"each step evaluates a bounded sample of {sample} queries of this workload, not the "
                                      f"{n_total}; the path is O(N), so queries/s is representative")
    line = {
        "impl": "reference",
        "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": threads, "kind": "port", "host_cpus": os.cpu_count(),
                         "sample": f"{sample} queries of the same workload per step (oracle: NumPy float64 port of the "
                                   f"reference path; BLAS threads actually used: {threads}, set explicitly -- torchrun's "
                                   "OMP_NUM_THREADS=1 is overridden)"},
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
# roofline.traffic: an ncu side-run of the same kernel on the same number of queries, after the timed region
# ---------------------------------------------------------------------------------------------
_NCU_TARGET = r"""
import sys
sys.path.insert(0, {repo!r})
import numpy as np, torch
from neuralfoil_b200 import Evaluator, synthetic
ev = Evaluator(device=0, extra_models={{"xxxlarge": synthetic.synthetic_xxxlarge_params()}} if {size!r} == "xxxlarge" else None)
n = {n}
base = torch.from_numpy(synthetic.sample_training_distribution(min(n, 1 << 20), seed=1000, dtype=np.float32)).cuda()
raw = base.repeat(1, -(-n // base.shape[1]))[:, :n].contiguous()  # traffic does not depend on the values
out = torch.empty((198, (n + 3) & ~3), dtype=torch.float32, device="cuda")
ev.evaluate_device(raw, {size!r}, out=out, variant={variant!r})
torch.cuda.synchronize()
"""


def measure_traffic_with_ncu(size: str, variant: str, n: int, timeout_s: int = 240):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the bench's kernel at the bench's size, or None."""
    import shutil

    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not Path(ncu).exists():
        return None, "ncu not found"
    code = _NCU_TARGET.format(repo=str(REPO), size=size, variant=variant, n=n)
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--csv",
           "-k", "regex:nfb_(ffma_groups|tc2|tc_h)", "-c", "1", sys.executable, "-c", code]
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s)
    except (OSError, subprocess.TimeoutExpired) as e:
        return None, f"ncu side-run failed: {type(e).__name__}"
    total, unit_scale = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    seen = 0
    for line in res.stdout.splitlines():
        if "dram__bytes_" not in line:
            continue
        cells = [c.strip('"') for c in line.split('","')]
        try:
            total += float(cells[-1].replace(",", "")) * unit_scale.get(cells[-2], 1.0)
            seen += 1
        except (ValueError, IndexError):
            continue
    if seen != 2:
        return None, f"ncu side-run gave no counters (exit {res.returncode}): {res.stderr[-200:]}"
    return total, "measured by an ncu side-run (dram__bytes_read.sum + dram__bytes_write.sum, one launch, same size) after the timed region"


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
    threads_mode = world == 1 and args.gpus > 1   # one process, one host thread per device (nfb_multi_eval_host)
    n_dev = args.gpus if threads_mode else world
    if not threads_mode and world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        if threads_mode:
            for d in range(n_dev):
                torch.cuda.synchronize(d)
        else:
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    models, weights_note = extra_models()
    size, variant = args.model_size, args.variant
    n_total = args.queries * (n_dev if args.scaling == "weak" else 1)
    ld_total = (n_total + 3) & ~3

    # ---- ONE batch in page-locked shared host memory: rank 0 samples it, every rank maps it ----
    names = [None, None, None]
    if rank == 0:
        raw_block = sharding.SharedHostBlock(23, ld_total, np.float32, create=True)
        out_block = sharding.SharedHostBlock(198, ld_total, np.float32, create=True)
        bl16_block = sharding.SharedHostBlock(192, ld_total, np.uint16, create=True)  # bfloat16 boundary-layer rows (mixed outputs)
        step = 1 << 21
        for s in range(0, n_total, step):  # the same sampler as ever, chunked so that 80 M queries need no 15 GB temporary
            e = min(n_total, s + step)
            raw_block.array[:, s:e] = synthetic.sample_training_distribution(e - s, seed=1000 + s // step, dtype=np.float32)
        names = [raw_block.name, out_block.name, bl16_block.name]
    if world > 1:
        dist.broadcast_object_list(names, src=0)
        if rank != 0:
            raw_block = sharding.SharedHostBlock(23, ld_total, np.float32, name=names[0])
            out_block = sharding.SharedHostBlock(198, ld_total, np.float32, name=names[1])
            bl16_block = sharding.SharedHostBlock(192, ld_total, np.uint16, name=names[2])
        dist.barrier()

    # the shards this process drives: one (its own rank's) under torchrun, all of them in threads mode
    my_shards = list(range(n_dev)) if threads_mode else [rank]
    evs = {d: Evaluator(device=(d if threads_mode else local_rank), extra_models=models) for d in my_shards}
    ev = evs[my_shards[0]]
    ranges = {d: sharding.shard_range(n_total, d, n_dev) for d in my_shards}
    raw_dev, out_dev = {}, {}
    for d in my_shards:
        s, e = ranges[d]
        dev = f"cuda:{d if threads_mode else local_rank}"
        raw_dev[d] = torch.from_numpy(raw_block.array[:, s:e]).to(dev)
        out_dev[d] = torch.empty((198, (e - s + 3) & ~3), dtype=torch.float32, device=dev)
    n_local = sum(e - s for s, e in ranges.values())
    barrier()

    fp32_peak = ev.measure_fp32_peak_tflops(5) if rank == 0 else 0.0

    def launch_all(v, outputs="all", outs=None):
        for d in my_shards:
            with torch.cuda.device(d if threads_mode else local_rank):
                evs[d].evaluate_device(raw_dev[d], size, out=(outs or out_dev)[d], variant=v, outputs=outputs)

    def device_timed(v, steps, warm, outputs="all", outs=None):
        """ms per step, CUDA events on every device this process drives, max over devices and ranks."""
        for _ in range(warm):
            launch_all(v, outputs, outs)
        barrier()
        ev0 = {d: torch.cuda.Event(enable_timing=True) for d in my_shards}
        ev1 = {d: torch.cuda.Event(enable_timing=True) for d in my_shards}
        for d in my_shards:
            with torch.cuda.device(d if threads_mode else local_rank):
                ev0[d].record()
        for _ in range(steps):
            launch_all(v, outputs, outs)
        for d in my_shards:
            with torch.cuda.device(d if threads_mode else local_rank):
                ev1[d].record()
        barrier()
        return max_over_ranks(max(ev0[d].elapsed_time(ev1[d]) for d in my_shards)) / steps

    # ---- device-resident throughput (the shards already in HBM) ----
    warm = max(args.warmup, 3)
    for _ in range(warm):
        launch_all(variant)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = sum(e_.launch_count() for e_ in evs.values())
    d0 = my_shards[0]
    with torch.cuda.device(d0 if threads_mode else local_rank):
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    t_begin = {d: torch.cuda.Event(enable_timing=True) for d in my_shards}
    t_end = {d: torch.cuda.Event(enable_timing=True) for d in my_shards}
    for d in my_shards:
        with torch.cuda.device(d if threads_mode else local_rank):
            t_begin[d].record()
    for i in range(args.steps):
        with torch.cuda.device(d0 if threads_mode else local_rank):
            starts[i].record()
        launch_all(variant)
        with torch.cuda.device(d0 if threads_mode else local_rank):
            stops[i].record()
    for d in my_shards:
        with torch.cuda.device(d if threads_mode else local_rank):
            t_end[d].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    launches = sum(e_.launch_count() for e_ in evs.values()) - launches0
    total_ms = max_over_ranks(max(t_begin[d].elapsed_time(t_end[d]) for d in my_shards))
    kernel_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]  # this rank's first device: one launch per step
    ms_per_step = total_ms / args.steps
    value = n_total / (ms_per_step * 1e-3)
    checksum = float(out_dev[d0][:6, :1000].double().sum().item())  # touch the result; also a sanity value

    # ---- end to end through the product's host entry point, outputs gathered in ONE host block ----
    e2e = None
    e2e_steps = max(1, min(args.steps, 3))
    multi_ev = Evaluator(devices=list(range(n_dev)), extra_models=models, multi_threshold=0) if threads_mode else None

    def e2e_timed(v, outputs="all"):
        rows = 198 if outputs == "all" else 6  # ("scalars" and "mixed": the six float32 rows)
        view_out = out_block.array[:rows]

        def one_pass():
            if outputs == "mixed":  # six float32 scalars + 192 bfloat16 boundary-layer rows: 408 B per query device->host
                s, e = ranges[rank]
                if e > s:
                    ev.evaluate_host_mixed_into(raw_block.array[:, s:e], size, view_out[:, s:e], bl16_block.array[:, s:e], variant=v)
            elif threads_mode:  # the single-process product call: shards, threads and landing inside nfb_multi_eval_host
                multi_ev.evaluate_host_into(raw_block.array[:, :n_total], size, view_out, variant=v, outputs=outputs)
            else:
                s, e = ranges[rank]
                if e > s:
                    ev.evaluate_host_into(raw_block.array[:, s:e], size, view_out[:, s:e], variant=v, outputs=outputs)
            if world > 1:
                dist.barrier()  # every rank's device->host copies have landed in the shared block: the gather is done
            if rank == 0:  # rank 0 reads a result of every shard from its own memory
                return float(sum(view_out[1, sharding.shard_range(n_total, r, n_dev)[0]] for r in range(n_dev)
                                 if sharding.shard_range(n_total, r, n_dev)[1] > sharding.shard_range(n_total, r, n_dev)[0]))
            return 0.0

        one_pass()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            one_pass()
        return max_over_ranks((time.perf_counter() - t0) / e2e_steps)

    if not args.no_e2e:
        e2e_s = e2e_timed(variant)
        e2e = {"value": n_total / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "h2d_bytes_per_step": 23 * 4 * n_total, "d2h_bytes_per_step": 198 * 4 * n_total,
               "gather": "inside the timed region: every rank's device->host copies land in its column range of one "
                         "page-locked shared (198, n) block owned by rank 0; barrier; rank 0 reads a value of every shard",
               "api": ("Evaluator(devices=[...]).evaluate_host_into -> nfb_multi_eval_host (one process, one host thread per device)"
                       if threads_mode else "Evaluator.evaluate_host_into -> nfb_eval_host (C ABI), one process per GPU")}
        if rank == 0:
            host_check = float(out_block.array[:6, :1000].astype(np.float64).sum())
            assert abs(host_check - checksum) <= 1e-6 * max(1.0, abs(checksum)), (host_check, checksum)

    # ---- the other variants, briefly: reported beside the headline, never mixed into it ----
    def measure_other(other: str) -> dict:
        t_ms = device_timed(other, args.steps, 3)
        extra = {"ms_per_step": t_ms, "value": n_total / (t_ms * 1e-3)}
        if not args.no_e2e:
            te = e2e_timed(other)
            extra["e2e"] = {"value": n_total / te, "unit": UNIT, "ms_per_step": te * 1e3,
                            "note": "same host path and gather as the headline's e2e; PCIe device->host moves 792 B/query"}
            if not threads_mode:
                tm = e2e_timed(other, outputs="mixed")
                extra["e2e_bf16_boundary_layer"] = {"value": n_total / tm, "unit": UNIT, "ms_per_step": tm * 1e3, "d2h_bytes_per_query": 408,
                                                    "note": "nfb_eval_host_mixed: six float32 scalars + 192 bfloat16 boundary-layer rows "
                                                            "(the float32 results rounded to nearest), same shared pinned blocks and gather"}
        return extra

    def measure_scalars(v: str) -> dict:  # outputs="scalars": only the six scalar rows travel back over PCIe
        out6 = {d: torch.empty((6, out_dev[d].shape[1]), dtype=torch.float32, device=out_dev[d].device) for d in my_shards}
        t_ms = device_timed(v, args.steps, 2, outputs="scalars", outs=out6)
        res = {"value": n_total / (t_ms * 1e-3), "unit": UNIT, "ms_per_step": t_ms, "d2h_bytes_per_query": 24}
        if not args.no_e2e:
            te = e2e_timed(v, outputs="scalars")
            res["e2e"] = {"value": n_total / te, "unit": UNIT, "ms_per_step": te * 1e3}
        return res

    tensor_extra = tensor_fp32_extra = None
    if size in ("xxxlarge", "xxlarge") and variant == "fp32" and not args.no_extras:
        tensor_fp32_extra = measure_other("tensor_fp32")
        tensor_extra = measure_other("tensor")
        if not args.no_e2e:
            tensor_fp32_extra["scalars_only"] = measure_scalars("tensor_fp32")
            tensor_extra["scalars_only"] = measure_scalars("tensor")

    # ---- weak scaling beside the strong-scaling headline (N > 1): 10 M queries on every GPU, device-resident ----
    weak_extra = None
    if args.scaling == "strong" and n_dev > 1 and not args.no_extras and not threads_mode:
        try:
            nw = args.queries
            s, e = ranges[rank]
            reps = -(-nw // max(1, e - s))
            raw_w = raw_dev[rank].repeat(1, reps)[:, :nw].contiguous()
            out_w = torch.empty((198, (nw + 3) & ~3), dtype=torch.float32, device="cuda")
            for _ in range(2):
                ev.evaluate_device(raw_w, size, out=out_w, variant=variant)
            barrier()
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0e.record()
            for _ in range(2):
                ev.evaluate_device(raw_w, size, out=out_w, variant=variant)
            t1e.record()
            barrier()
            t_ms = max_over_ranks(t0e.elapsed_time(t1e)) / 2
            weak_extra = {"value": n_dev * nw / (t_ms * 1e-3), "unit": UNIT, "ms_per_step": t_ms, "queries_per_gpu": nw,
                          "note": "weak scaling, device-resident: every GPU evaluates as many queries as the whole strong-scaling batch"}
            del raw_w, out_w
        except Exception as e:  # noqa: BLE001
            weak_extra = {"error": f"{type(e).__name__}: {e}"}

    # ---- the Jacobian kernel, briefly: outputs + 198 x 23 derivatives per query, device-resident ----
    jacobian_extra = scalar_grad_extra = None
    if variant == "fp32" and not args.no_extras and not threads_mode:
        try:
            nj = min(n_local, 20_000)
            rj = raw_dev[rank][:, :nj].contiguous()
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
            # the reverse-mode kernel: gradients of an objective of the six scalars
            ng = min(n_local, 2_000_000)
            rg = raw_dev[rank][:, :ng].contiguous()
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
        except Exception as e:  # noqa: BLE001  (an extra must never cost the headline line)
            scalar_grad_extra = {"error": f"{type(e).__name__}: {e}"}

    # ---- the training step (forward + loss + backward to every layer's weight gradients), device-resident ----
    training_extra = None
    if variant == "fp32" and not args.no_extras and not threads_mode:
        try:
            nb = 65536
            xt = torch.randn((nb, 25), device="cuda") * 0.3
            yt = torch.randn((nb, 198), device="cuda") * 0.1
            yt[:, 0] = (torch.rand(nb, device="cuda") < 0.8).float()
            for _ in range(2):
                ev.train_step_device(xt, yt, size)
            barrier()
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0e.record()
            for _ in range(3):
                ev.train_step_device(xt, yt, size)
            t1e.record()
            barrier()
            t_ms = max_over_ranks(t0e.elapsed_time(t1e)) / 3
            tf = 3 * FLOP_PER_QUERY[size] * nb / (t_ms * 1e-3) / 1e12
            training_extra = {"value": world * nb / (t_ms * 1e-3), "unit": "samples/s", "batch": nb, "ms_per_step": t_ms,
                              "achieved_tflops": tf, "flop_model": "3 x the forward pass (forward, d a = W^T d z, d W = d z a^T), both twins",
                              "kernels": "nfb_train_fwdbwd_kernel + nfb_train_wgrad_kernel + nfb_train_reduce_kernel (fp32 FFMA2)",
                              "reference": "training/train.py:56-114, 194-236 (forward, loss, backward; the optimiser step stays with the caller)"}
            del xt, yt
        except Exception as e:  # noqa: BLE001
            training_extra = {"error": f"{type(e).__name__}: {e}"}

    # ---- the reference's DEFAULT model (xlarge, main.py:71) on the three kernels, device-resident, briefly ----
    default_model_extra = None
    if size != "xlarge" and not args.no_extras and not threads_mode:
        try:
            nd = min(n_local, 4_000_000)
            rd = raw_dev[rank][:, :nd].contiguous()
            od = torch.empty((198, (nd + 3) & ~3), dtype=torch.float32, device="cuda")
            default_model_extra = {"model_size": "xlarge", "queries_per_launch": nd, "unit": UNIT}
            for v in ("fp32", "tensor", "tensor_fp32"):
                for _ in range(2):
                    ev.evaluate_device(rd, "xlarge", out=od, variant=v)
                barrier()
                t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0e.record()
                for _ in range(5):
                    ev.evaluate_device(rd, "xlarge", out=od, variant=v)
                t1e.record()
                barrier()
                t_ms = max_over_ranks(t0e.elapsed_time(t1e)) / 5
                default_model_extra[v] = world * nd / (t_ms * 1e-3)
            del rd, od
        except Exception as e:  # noqa: BLE001
            default_model_extra = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        peaks_path = REPO / "MEASURED_PEAKS.json"
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        bf16_peak, bf16_src = 1590.0, "fallback (B200_PROFILING.md)"
        if peaks_path.exists():
            peaks = json.loads(peaks_path.read_text())
            hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
            # every tensor figure here is timed inside a process that keeps the GPU busy for tens of seconds: sustained
            bf16_peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
            bf16_src = "measured sustained (MEASURED_PEAKS.json: bf16_tflops_sustained)" if "bf16_tflops_sustained" in peaks \
                else "measured burst (MEASURED_PEAKS.json)"
        k_ms = float(np.mean(kernel_ms))
        n_k = ranges[d0][1] - ranges[d0][0]  # queries of the launch kernel_ms times (this rank's first device)
        achieved_tflops = FLOP_PER_QUERY[size] * n_k / (k_ms * 1e-3) / 1e12
        achieved_gbs = BYTES_PER_QUERY * n_k / (k_ms * 1e-3) / 1e9
        n_per_gpu = -(-n_total // n_dev)
        if tensor_extra is not None:
            tf = FLOP_PER_QUERY[size] * n_per_gpu / (tensor_extra["ms_per_step"] * 1e-3) / 1e12  # per GPU
            tensor_extra.update({
                "unit": UNIT, "kernel": "nfb_tc2_kernel<H, 1> (tcgen05 cta_group::2 kind::f16 on CTA pairs, fp16 operands, fp32 accumulate in TMEM)",
                "accuracy": "TF32-class: median relative error 3e-4; tests/parity.py 'tensor' tolerance",
                "roofline": {"bound": "tensor", "achieved": tf, "peak": bf16_peak, "unit": "TFLOP/s", "frac": tf / bf16_peak,
                             "peak_source": bf16_src}})
        if tensor_fp32_extra is not None:
            tf = 3 * FLOP_PER_QUERY[size] * n_per_gpu / (tensor_fp32_extra["ms_per_step"] * 1e-3) / 1e12
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
            "kernel_ms_per_launch": k_ms, "queries_per_launch": n_k,
            "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                    "peak_source": hbm_src},
            "traffic": None, "algorithmic_bytes_per_launch": BYTES_PER_QUERY * n_k,
        }
        if not args.no_ncu and n_dev == 1:
            traffic, how = measure_traffic_with_ncu(size, variant, n_k)
            if traffic is not None:
                roofline["traffic"], roofline["traffic_note"] = traffic, how
            else:
                roofline["traffic_note"] = how
        if roofline["traffic"] is None:
            per_q = NCU_TRAFFIC_BYTES_PER_QUERY.get((size, variant))
            if per_q is not None:
                roofline["traffic"] = per_q * n_k
                roofline["traffic_note"] = (roofline.get("traffic_note", "") + "; fallback: DRAM bytes per query of the round-1 ncu capture "
                                            "(10 M queries, profiles/r1_ncu_full_*) x this launch's queries").lstrip("; ")
        if variant != "fp32":
            mult = 3 if variant == "tensor_fp32" else 1
            roofline.update({"bound": "tensor", "kernel": "nfb_tc2_kernel", "peak": bf16_peak, "peak_source": bf16_src,
                             "achieved": mult * achieved_tflops, "frac": mult * achieved_tflops / bf16_peak,
                             "tensor_macs_per_algorithmic_mac": mult})
            roofline.pop("frac_of_nominal_74.4", None)
        if training_extra and "achieved_tflops" in training_extra:
            training_extra["roofline"] = {"bound": "fp32", "achieved": training_extra["achieved_tflops"], "peak": fp32_peak, "unit": "TFLOP/s",
                                          "frac": training_extra["achieved_tflops"] / fp32_peak if fp32_peak else None}
        cpu_baseline = None
        if not args.no_cpu_baseline and n_dev == 1:
            sample = args.cpu_sample or auto_cpu_sample(size)
            qps, sec = cpu_reference_qps(size, sample, 1, 1)
            threads = blas_threads()
            cpu_baseline = {"value": qps, "unit": UNIT, "cores": threads, "kind": "port", "host_cpus": os.cpu_count(),
                            "sample": f"{sample} queries of the same workload, one pass, {sec:.1f} s "
                                      f"(oracle: NumPy float64 port of the reference path, BLAS threads actually used: {threads})"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_dev, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": {"fp32": "f32", "tensor": "f16 operands / f32 accumulate",
                                                                   "tensor_fp32": "2 x f16 split operands / f32 accumulate"}[variant],
            "data": "synthetic", "config": workload_config(args, weights_note, n_total, n_dev),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "tensor_variant": tensor_extra,
            "tensor_fp32_variant": tensor_fp32_extra, "weak_scaling": weak_extra, "jacobian": jacobian_extra,
            "scalar_gradient": scalar_grad_extra, "training_step": training_extra, "default_model_xlarge": default_model_extra,
            "checksum_first_1000": checksum,
            "process_model": "one process, one host thread per device" if threads_mode else "one process per GPU (torchrun)" if world > 1 else "one process, one GPU",
        }
        print(json.dumps(line), flush=True)

    # ---- tidy up: unmap the shared blocks everywhere, unlink on rank 0 ----
    for e_ in evs.values():
        e_.close()
    if multi_ev is not None:
        multi_ev.close()
    raw_dev.clear(); out_dev.clear()
    if world > 1:
        dist.barrier()
    raw_block.close(); out_block.close(); bl16_block.close()
    if rank == 0:
        raw_block.unlink(); out_block.unlink(); bl16_block.unlink()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
