"""Where a batch-1 dict-API call spends its time (host side vs library call)."""
import sys, time, ctypes as C
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import numpy as np
from neuralfoil_b200 import Evaluator, synthetic, main as nfm, _lib
ev = Evaluator(device=0)
K = synthetic.FIXED_KULFAN
def med(f, n=500):
    for _ in range(20): f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(ts))
for size in ("xxsmall", "xlarge", "xxxlarge") if False else ("xxsmall", "xlarge"):
    mid = ev._model_id(size)
    rows, n = nfm.prepare_raw_rows(K, 3.0, 1e6, 9.0, 1.0, 1.0)
    print(size, "prepare_raw_rows", med(lambda: nfm.prepare_raw_rows(K, 3.0, 1e6, 9.0, 1.0, 1.0)))
    print(size, "_evaluate_host_rows (alloc + ctypes + library)", med(lambda: ev._evaluate_host_rows(mid, rows, n, np.float64)))
    block = np.empty((198, 4)); ptrs = (C.c_void_p * 23)(*[r.ctypes.data for r in rows]); strides = (C.c_int64 * 23)(*([1] * 23))
    f = lambda: ev._lib.nfb_eval_host(ev._ctx, mid, 0, 1, ptrs, strides, 1, block.ctypes.data_as(C.c_void_p), 4, 1, 0)
    print(size, "nfb_eval_host alone", med(f))
    print(size, "dict build", med(lambda: dict(zip(nfm._OUTPUT_KEYS, block[:, :1]))))
    print(size, "full call", med(lambda: ev.get_aero_from_kulfan_parameters(K, alpha=3.0, Re=1e6, model_size=size)))
import torch
for size in ("xxsmall", "xlarge"):
    r = torch.from_numpy(np.ascontiguousarray(synthetic.sample_training_distribution(1, seed=1, dtype=np.float32))).cuda()
    out = torch.empty((198, 4), dtype=torch.float32, device="cuda")
    for n in (1,):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(5): ev.evaluate_device(r, size, out=out)
        torch.cuda.synchronize(); e0.record()
        for _ in range(50): ev.evaluate_device(r, size, out=out)
        e1.record(); torch.cuda.synchronize()
        print(size, "kernel time per launch at n=1 (us, back-to-back launches)", e0.elapsed_time(e1) / 50 * 1e3)
# zero-copy (mapped pinned staging block) against explicit copies, library call alone
for size in ("xxsmall", "xlarge"):
    mid = ev._model_id(size)
    rows, n = nfm.prepare_raw_rows(K, 3.0, 1e6, 9.0, 1.0, 1.0)
    for nq in (1, 64, 1024):
        raw = np.ascontiguousarray(synthetic.sample_training_distribution(nq, seed=2))
        block = np.empty((198, (nq + 3) & ~3)); ptrs = (C.c_void_p * 23)(*[raw[i].ctypes.data for i in range(23)]); strides = (C.c_int64 * 23)(*([1] * 23))
        f = lambda: ev._lib.nfb_eval_host(ev._ctx, mid, 0, nq, ptrs, strides, 1, block.ctypes.data_as(C.c_void_p), block.shape[1], 1, 0)
        ev.set_tuning(mapped_max=0)
        t_copy = med(f)
        ev.set_tuning(mapped_max=-1)
        t_map = med(f)
        print(size, f"nfb_eval_host n={nq}: explicit copies {t_copy:.1f} us, zero-copy {t_map:.1f} us")
