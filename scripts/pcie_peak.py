"""The box's PCIe ceiling for this path: plain cudaMemcpyAsync between pinned host memory and the device, both directions,
alone and (with --all) on every GPU at once -- what `e2e` with full outputs (792 B per query device->host) is bounded by.
    python scripts/pcie_peak.py [--all]"""
import json, sys, threading, time
import torch

def measure(dev: int, gib: float = 1.0, reps: int = 5):
    torch.cuda.set_device(dev)
    n = int(gib * (1 << 30))
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    devb = torch.empty(n, dtype=torch.uint8, device=f"cuda:{dev}")
    out = {}
    for name, dst, src in (("h2d", devb, host), ("d2h", host, devb)):
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize(dev)
        out[name + "_GBps"] = round(n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
    return out

if __name__ == "__main__":
    print(json.dumps({"gpu": 0, "alone": True, **measure(0)}), flush=True)
    if "--all" in sys.argv and torch.cuda.device_count() > 1:
        # Concurrent: buffers pinned first, then every GPU copies at once and the WALL clock over all of them is what counts.
        # (Round 2's first version timed each GPU with its own events while the threads were still pinning their buffers one
        # after the other -- the copies hardly overlapped and the "sum" read 306 GB/s on a box whose ceiling is 92 GB/s.)
        G = torch.cuda.device_count()
        n = 1 << 30
        host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(G)]
        devb = [torch.empty(n, dtype=torch.uint8, device=f"cuda:{d}") for d in range(G)]
        out = {"gpus": G, "concurrent": True, "timing": "wall clock over all GPUs"}
        for name in ("h2d", "d2h"):
            def work(d, reps):
                torch.cuda.set_device(d)
                for _ in range(reps):
                    (devb[d] if name == "h2d" else host[d]).copy_(host[d] if name == "h2d" else devb[d], non_blocking=True)
                torch.cuda.synchronize(d)
            for reps in (1, 6):
                ts = [threading.Thread(target=work, args=(d, reps)) for d in range(G)]
                t0 = time.perf_counter()
                for t in ts: t.start()
                for t in ts: t.join()
                dt = time.perf_counter() - t0
            out[f"sum_{name}_GBps"] = round(G * 6 * n / dt / 1e9, 1)
        print(json.dumps(out), flush=True)
