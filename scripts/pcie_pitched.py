"""Device->host copies of full outputs from every GPU at once, by destination pattern (cudaMemcpy2DAsync, one thread per GPU):
  contiguous   one cudaMemcpyAsync of the shard's bytes into the GPU's own pinned buffer                      (the PCIe ceiling)
  own_block    2-D copy, 198 rows x W, into the GPU's OWN (198, W) pinned block                               (pitched source only)
  shared_rows  2-D copy, 198 rows x W, into the GPU's column range of ONE (198, n) pinned block, row pitch 4 n (the product's landing)
  shared_rows/chunked  the same in chunks of 128 Ki columns alternating over two streams                     (nfb_eval_host)
    python scripts/pcie_pitched.py [n_total]"""
import ctypes as C, json, sys, threading, time
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
G = torch.cuda.device_count()
ROWS = 198
per = ((-(-n // G)) + 3) & ~3
ld = per * G
torch.cuda.init()
rt = None
for line in open("/proc/self/maps"):
    if "libcudart" in line:
        rt = C.CDLL(line.split()[-1]); break
assert rt is not None, "libcudart not loaded"
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
D2H = 2
shared = torch.empty((ROWS, ld), dtype=torch.float32).pin_memory()
own = [torch.empty((ROWS, per), dtype=torch.float32).pin_memory() for _ in range(G)]
dev = [torch.empty((ROWS, per), dtype=torch.float32, device=f"cuda:{g}") for g in range(G)]
streams = [[torch.cuda.Stream(device=g), torch.cuda.Stream(device=g)] for g in range(G)]

def run(mode, g, reps=4):
    torch.cuda.set_device(g)
    s0, s1 = streams[g]
    for _ in range(reps):
        if mode == "contiguous":
            rt.cudaMemcpyAsync(own[g].data_ptr(), dev[g].data_ptr(), ROWS * per * 4, D2H, s0.cuda_stream)
        elif mode == "own_block":
            rt.cudaMemcpy2DAsync(own[g].data_ptr(), per * 4, dev[g].data_ptr(), per * 4, per * 4, ROWS, D2H, s0.cuda_stream)
        elif mode == "shared_rows":
            rt.cudaMemcpy2DAsync(shared.data_ptr() + g * per * 4, ld * 4, dev[g].data_ptr(), per * 4, per * 4, ROWS, D2H, s0.cuda_stream)
        else:
            ck = 1 << 17
            for i, c0 in enumerate(range(0, per, ck)):
                w = min(ck, per - c0)
                rt.cudaMemcpy2DAsync(shared.data_ptr() + (g * per + c0) * 4, ld * 4, dev[g].data_ptr() + c0 * 4, per * 4, w * 4, ROWS, D2H,
                                     (s0 if i % 2 == 0 else s1).cuda_stream)
    s0.synchronize(); s1.synchronize()

for mode in ("contiguous", "own_block", "shared_rows", "shared_rows/chunked"):
    for gpus in ([0], list(range(G))):
        for g in gpus: run(mode, g, 1)
        ts = [threading.Thread(target=run, args=(mode, g)) for g in gpus]
        t0 = time.perf_counter()
        for t in ts: t.start()
        for t in ts: t.join()
        dt = time.perf_counter() - t0
        print(json.dumps({"mode": mode, "gpus": len(gpus), "bytes_per_gpu": ROWS * per * 4, "aggregate_GBps": round(4 * len(gpus) * ROWS * per * 4 / dt / 1e9, 1)}), flush=True)
