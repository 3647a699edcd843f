"""Which host memory can eight GPUs write full outputs into at PCIe rate?  One process, one host thread per device
(nfb_multi_eval_host), ONE (198, n) result block of three kinds:
  shm       POSIX shared memory + cudaHostRegister (sharding.SharedHostBlock: what one-process-per-GPU jobs must use)
  hostalloc cudaHostAlloc (torch pin_memory)
  anon      anonymous mmap (NumPy) + cudaHostRegister
and the chunk size of nfb_eval_host.  Prints e2e queries/s and the aggregate device->host GB/s.
    python scripts/e2e_host_memory_ab.py [n] [variant]"""
import json, sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import ctypes as C
import numpy as np, torch
from neuralfoil_b200 import Evaluator, synthetic, sharding, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["tensor", "fp32"]
size = "xxxlarge"
G = torch.cuda.device_count()
ld = (n + 3) & ~3
ev = Evaluator(devices=list(range(G)), extra_models={size: synthetic.synthetic_xxxlarge_params()}, multi_threshold=0)
lib = _lib.load()
raw_t = torch.empty((23, ld), dtype=torch.float32).pin_memory()
raw = raw_t.numpy()
step = 1 << 21
for s in range(0, n, step):
    e = min(n, s + step)
    raw[:, s:e] = synthetic.sample_training_distribution(e - s, seed=1000 + s // step, dtype=np.float32)

def blocks():
    b = sharding.SharedHostBlock(198, ld, np.float32, create=True)
    yield "shm+register", b.array, lambda: (b.close(), b.unlink())
    t = torch.empty((198, ld), dtype=torch.float32).pin_memory()
    yield "hostalloc", t.numpy(), lambda: None
    a = np.empty((198, ld), dtype=np.float32)
    a[:] = 0
    _lib.check(lib.nfb_host_register(C.c_void_p(a.ctypes.data), a.nbytes))
    yield "anon+register", a, lambda: lib.nfb_host_unregister(C.c_void_p(a.ctypes.data))

for kind, out, done in blocks():
    for v in variants:
        for chunk in (0, 1 << 17):
            ev.evaluate_host_into(raw[:, :n], size, out, variant=v, chunk=chunk)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                ev.evaluate_host_into(raw[:, :n], size, out, variant=v, chunk=chunk)
            dt = (time.perf_counter() - t0) / reps
            print(json.dumps({"gpus": G, "n": n, "host_block": kind, "variant": v, "chunk": chunk or "default (512 Ki)",
                              "e2e_Mqps": round(n / dt / 1e6, 2), "ms": round(dt * 1e3, 2),
                              "d2h_GBps_aggregate": round(n * 792 / dt / 1e9, 1)}), flush=True)
    done()
