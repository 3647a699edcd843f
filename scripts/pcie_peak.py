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
        res = {}
        def work(d):
            res[d] = measure(d, reps=8)
        ts = [threading.Thread(target=work, args=(d,)) for d in range(torch.cuda.device_count())]
        t0 = time.perf_counter()
        for t in ts: t.start()
        for t in ts: t.join()
        print(json.dumps({"gpus": torch.cuda.device_count(), "concurrent": True, "per_gpu": res,
                          "sum_d2h_GBps": round(sum(r["d2h_GBps"] for r in res.values()), 1),
                          "sum_h2d_GBps": round(sum(r["h2d_GBps"] for r in res.values()), 1)}), flush=True)
