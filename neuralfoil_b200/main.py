"""
Host side of the B200 evaluator: the reference's `get_aero_from_kulfan_parameters`
(/root/reference/neuralfoil/main.py:64-332) with the same signature, return value and errors, whose
numerical body -- featurisation, both network passes, Mahalanobis correction, un-flip, averaging,
squashing and decode -- runs in one fused CUDA kernel behind the C ABI of
include/neuralfoil_b200.h.

What stays on the host is what the reference does before `x = np.stack(...)` (main.py:154-199):
argument validation and the decision of how many cases N the call describes.  The 23 raw rows are
handed to the library as they are (a pointer and a stride each; scalars are stride-0 rows), so no
(N, 25) matrix is ever built on the host.

Two entry levels:
  * `get_aero_from_kulfan_parameters(...)`  NumPy in, dict of 198 float64 arrays out (drop-in).
  * `Evaluator.evaluate_device(...)`        torch CUDA tensors in and out, nothing crosses PCIe.
"""

from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

import numpy as np

from . import _lib

N_BL = 32  # reference neuralfoil/_basic_data_type.py:51
bl_x_points = (np.arange(N_BL) + 0.5) / N_BL  # _basic_data_type.py:23-38,52; re-exported like main.py:11

MODEL_SIZES = ("xxsmall", "xsmall", "small", "medium", "large", "xlarge", "xxlarge", "xxxlarge")
nn_weights_dir = Path(__file__).resolve().parent / "nn_weights_and_biases"


def output_keys() -> list[str]:
    """The 198 keys of the returned dict, in the reference's insertion order (main.py:318-331)."""
    keys = ["analysis_confidence", "CL", "CD", "CM", "Top_Xtr", "Bot_Xtr"]
    for side in ("upper", "lower"):
        for quantity in ("theta", "H", "ue/vinf"):
            keys += [f"{side}_bl_{quantity}_{i}" for i in range(N_BL)]
    return keys


_OUTPUT_KEYS = output_keys()
_ONES_23 = (C.c_int64 * 23)(*([1] * 23))

# The 23 scalar inputs of the path in the row order of the C ABI: the columns of every Jacobian this package returns.
JACOBIAN_INPUTS = tuple(
    [f"upper_weights_{i}" for i in range(8)] + [f"lower_weights_{i}" for i in range(8)]
    + ["leading_edge_weight", "TE_thickness", "alpha", "Re", "n_crit", "xtr_upper", "xtr_lower"]
)


def default_loss_weights() -> np.ndarray:
    """The reference's loss weights (training/train.py:180-190): 198 float64, normalised to sum to 1000."""
    w = np.ones(198)
    w[0] *= 0.005  # analysis confidence
    w[2] *= 3      # ln(CD)
    w[3:6] *= 0.25  # CM, Top_Xtr, Bot_Xtr
    w[6:] *= 5e-3  # boundary-layer outputs
    return w / w.sum() * 1000


def bf16_bits_to_float32(bits: np.ndarray) -> np.ndarray:
    """uint16 bfloat16 bit patterns (the `out_bl` block of `evaluate_host_mixed_into`) -> float32, exactly."""
    return (np.asarray(bits, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def _length(x) -> int:
    """aerosandbox.numpy.length as used at main.py:161,189-193: len() of sized things, 1 for scalars."""
    try:
        return len(x)
    except TypeError:
        return 1


def layers_from_params(nn_params: dict) -> list[tuple[np.ndarray, np.ndarray]]:
    """[(W, b), ...] in layer order from `net.{i}.weight/bias` keys; rule and error of main.py:206-215."""
    try:
        layer_indices = sorted({int(key.split(".")[1]) for key in nn_params.keys()})
    except (TypeError, ValueError, IndexError) as e:
        raise ValueError(
            "Got an unexpected neural network file format.\n"
            "Dictionary keys should be strings of the form 'net.0.weight', 'net.0.bias', 'net.2.weight', etc.'.\n"
            f"Instead, got keys of the form {nn_params.keys()}.\n"
        ) from e
    layers = []
    for i in layer_indices:
        w = np.ascontiguousarray(nn_params[f"net.{i}.weight"], dtype=np.float32)
        b = np.ascontiguousarray(nn_params[f"net.{i}.bias"], dtype=np.float32)
        if w.ndim != 2 or b.shape != (w.shape[0],):
            raise ValueError(f"Layer net.{i} has inconsistent shapes {w.shape} / {b.shape}.")
        layers.append((w, b))
    for (w0, _), (w1, _) in zip(layers[:-1], layers[1:]):
        if w1.shape[1] != w0.shape[0]:
            raise ValueError("Consecutive layers of the network do not chain.")
    return layers


def prepare_raw_rows(kulfan_parameters, alpha, Re, n_crit, xtr_upper, xtr_lower):
    """
    Validation and case counting of main.py:160-199.  Returns (rows, n_cases): 23 float64 1-D arrays,
    each of size n_cases or 1 (a size-1 row is broadcast by the kernel, like `np.ones(N) * row`).

    Differences from the reference, both supersets: Python lists are accepted for the flow inputs
    (the reference's `2 * alpha` duplicates a list and then fails the length check), and the raw
    inputs, not the featurised ones, are what gets counted.
    """
    for key in ("upper_weights", "lower_weights"):
        n_weights = _length(kulfan_parameters[key])
        if n_weights != 8:
            raise ValueError(
                f"NeuralFoil's neural networks expect exactly 8 CST weights per side, but "
                f"`kulfan_parameters[{key!r}]` has {n_weights}. If your airfoil is parameterized "
                f"with a different CST resolution, refit it with 8 weights per side (e.g., via "
                f"`aerosandbox.Airfoil.to_kulfan_airfoil(n_weights_per_side=8)`)."
            )
    rows = [kulfan_parameters["upper_weights"][i] for i in range(8)]
    rows += [kulfan_parameters["lower_weights"][i] for i in range(8)]
    rows += [
        kulfan_parameters["leading_edge_weight"],
        kulfan_parameters["TE_thickness"],
        alpha,
        Re,
        n_crit,
        xtr_upper,
        xtr_lower,
    ]
    n_cases = 1
    arrays = []
    for row in rows:
        a = np.asarray(row, dtype=np.float64)
        if a.ndim > 1:
            raise ValueError("The inputs to the neural network must be scalars or 1-D arrays.")
        a = np.ascontiguousarray(a.reshape(-1))
        if a.size == 0:
            raise ValueError("The inputs to the neural network must not be empty.")
        if a.size > 1:
            if n_cases == 1:
                n_cases = a.size
            elif a.size != n_cases:
                raise ValueError(
                    "The inputs to the neural network must all have the same length. "
                    f"(Conflicting lengths: {n_cases} and {a.size})"
                )
        arrays.append(a)
    return arrays, n_cases



class _PooledBlock:
    """Owns the memory of one large result block and hands it back to the pool when the last NumPy view of it dies.
    NumPy keeps the object exposing `__array_interface__` alive as the base of every array built on it."""

    __slots__ = ("_mem", "_pool", "__array_interface__")

    def __init__(self, mem: np.ndarray, shape, dtype, pool):
        self._mem, self._pool = mem, pool
        self.__array_interface__ = {"shape": tuple(shape), "typestr": np.dtype(dtype).str,
                                    "data": (mem.__array_interface__["data"][0], False), "version": 3}

    def __del__(self):
        try:
            self._pool._give_back(self._mem)
        except Exception:  # interpreter shutdown
            pass


class _ResultPool:
    """
    Recycles the host memory of large result blocks (the 198 x N float64 block behind the returned dict).

    A fresh 1.6 GB NumPy array costs the kernel a zero-filled page fault for every page the first time it is written --
    for a 1 M-point call that was a third of the drop-in call's time (profiles/r2_polar_grid_result_pool.jsonl).  Blocks of
    at least `MIN_BYTES` are therefore carved out of byte buffers that return here when the caller drops the last array
    that points into them, and the next call of the same size finds its pages already mapped.  A caller that rebinds
    `aero = get_aero_from_kulfan_parameters(...)` in a loop alternates between two buffers.  At most `cap_bytes` are kept
    (`$NEURALFOIL_B200_RESULT_POOL_MB`, default 4096; 0 switches the pool off); the oldest buffers go first.
    """

    MIN_BYTES = 8 << 20

    def __init__(self, cap_bytes: int):
        self.cap_bytes = int(cap_bytes)
        self._lock = threading.Lock()
        self._free: list[np.ndarray] = []  # oldest first

    def empty(self, shape, dtype) -> np.ndarray:
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        if self.cap_bytes <= 0 or nbytes < self.MIN_BYTES:
            return np.empty(shape, dtype=dtype)
        mem = None
        with self._lock:
            for i in range(len(self._free) - 1, -1, -1):
                if self._free[i].nbytes == nbytes:
                    mem = self._free.pop(i)
                    break
        if mem is None:
            mem = np.empty(nbytes, dtype=np.uint8)
        return np.asarray(_PooledBlock(mem, shape, dtype, self))

    def _give_back(self, mem: np.ndarray) -> None:
        with self._lock:
            self._free.append(mem)
            while self._free and sum(m.nbytes for m in self._free) > self.cap_bytes:
                self._free.pop(0)

    def clear(self) -> None:
        with self._lock:
            self._free.clear()


_result_pool = _ResultPool(int(float(os.environ.get("NEURALFOIL_B200_RESULT_POOL_MB", "4096")) * (1 << 20)))


class Evaluator:
    """
    The distribution statistics and every network, repacked and resident on one GPU -- or replicated on several GPUs
    of one box.  Mirrors the reference's import-time globals (main.py:25-44).

    `device`   one CUDA ordinal (default: $NEURALFOIL_B200_DEVICE, else $LOCAL_RANK, else 0).
    `devices`  "all" or a list of ordinals: the host entry points (`get_aero_from_kulfan_parameters`, `evaluate_host`,
               `evaluate_host_into`) split batches of at least `multi_threshold` queries into contiguous shards, one per
               GPU, and every GPU copies its results straight into its column range of the one result block
               (include/neuralfoil_b200.h: nfb_multi_eval_host).  The values are bit-identical to the
               single-GPU ones.  Device-resident calls (`evaluate_device`, ...) run on the first device of the list.
               "auto" = one device now; the others join the first time a batch is large enough to need them.

    Threading: one Evaluator may be shared by any number of Python threads -- the library serialises the calls into a
    context ("Threading"); use one Evaluator per thread for concurrency.
    """

    def __init__(self, device: int | None = None, weights_dir: Path | str | None = None, extra_models=None,
                 devices=None, multi_threshold: int = 1 << 18):
        self._lib = _lib.load()
        self.weights_dir = Path(weights_dir) if weights_dir is not None else nn_weights_dir
        self.multi_threshold = int(multi_threshold)
        self._multi = None          # nfb_multi handle when this evaluator spans several devices
        self._auto_multi = None     # devices="auto": the lazily created multi-device twin of this evaluator
        self._auto = devices == "auto"
        self._extra_models = dict(extra_models or {})
        self._lock = threading.Lock()
        self._tls = threading.local()
        handle = C.c_void_p()
        if devices is not None and not self._auto:
            if device is not None:
                raise ValueError("give either `device` or `devices`")
            if isinstance(devices, str):
                if devices != "all":
                    raise ValueError(f"Invalid {devices=}. Must be 'all', 'auto' or a list of CUDA ordinals.")
                _lib.check(self._lib.nfb_multi_create(None, 0, C.byref(handle)))
            else:
                ids = [int(d) for d in devices]
                _lib.check(self._lib.nfb_multi_create((C.c_int * len(ids))(*ids), len(ids), C.byref(handle)))
            self._multi = handle
            self.n_devices = int(self._lib.nfb_multi_device_count(self._multi))
            self._ctx = C.c_void_p(self._lib.nfb_multi_context(self._multi, 0))
            self.device = self.device_info()["device"]
        else:
            if device is None:
                device = int(os.environ.get("NEURALFOIL_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
            self.device = int(device)
            _lib.check(self._lib.nfb_create(self.device, C.byref(handle)))
            self._ctx = handle
            self.n_devices = 1
        self._model_ids: dict[str, int] = {}
        self.model_dims: dict[str, list[int]] = {}

        with np.load(self.weights_dir / "scaled_input_distribution.npz") as f:  # main.py:27-32
            mean = np.ascontiguousarray(f["mean_inputs_scaled"], dtype=np.float64)
            inv_cov = np.ascontiguousarray(f["inv_cov_inputs_scaled"], dtype=np.float64)
        if mean.shape != (25,) or inv_cov.shape != (25, 25):
            raise ValueError("scaled_input_distribution.npz does not describe 25 inputs.")
        set_distribution = self._lib.nfb_multi_set_distribution if self._multi else self._lib.nfb_set_distribution
        _lib.check(
            set_distribution(
                self._multi or self._ctx, mean.ctypes.data_as(C.POINTER(C.c_double)),
                inv_cov.ctypes.data_as(C.POINTER(C.c_double)),
            )
        )
        for path in sorted(self.weights_dir.glob("nn-*.npz")):  # main.py:35-44
            with np.load(path) as f:
                self.load_model(path.stem.removeprefix("nn-"), {k: f[k] for k in f.files})
        for name, params in (extra_models or {}).items():
            self.load_model(name, params)

    # -- model management ------------------------------------------------------------------------
    @property
    def allowable_model_sizes(self) -> set[str]:
        return set(self._model_ids)

    def load_model(self, name: str, nn_params: dict) -> None:
        """Upload a network given as the reference's `.npz` dictionary (`net.{2j}.weight/bias`)."""
        layers = layers_from_params(nn_params)
        dims = [layers[0][0].shape[1]] + [w.shape[0] for w, _ in layers]
        model_id = self._model_ids.get(name, len(self._model_ids))
        n = len(layers)
        w_ptrs = (C.c_void_p * n)(*[w.ctypes.data for w, _ in layers])
        b_ptrs = (C.c_void_p * n)(*[b.ctypes.data for _, b in layers])
        c_dims = (C.c_int * (n + 1))(*dims)
        load = self._lib.nfb_multi_load_model if self._multi else self._lib.nfb_load_model
        _lib.check(load(self._multi or self._ctx, model_id, n, c_dims, w_ptrs, b_ptrs))
        self._model_ids[name] = model_id
        self.model_dims[name] = dims
        if name not in self._extra_models and not (self.weights_dir / f"nn-{name}.npz").exists():
            self._extra_models[name] = nn_params  # so that a lazily created multi-device twin gets it too
        if self._auto_multi is not None:
            self._auto_multi.load_model(name, nn_params)

    @staticmethod
    def _variant(variant: str) -> int:
        """
        "fp32" = FFMA kernel (reference accuracy); "tensor" / "tensor_fp32" = the tcgen05 kernels (one fp16 MMA per
        product / three-term fp16 split).  A suffix selects a cross-check kernel (tests and A/B only): "fp32:tiled" the
        CTA-synchronous fp32 kernel, "tensor:one_sm" / "tensor_fp32:one_sm" the one-SM tensor kernels.
        """
        name, _, debug = variant.partition(":")
        try:
            code = _lib.VARIANTS[name]
            if debug:
                code |= {"tiled": _lib.DEBUG_FFMA_TILED, "one_sm": _lib.DEBUG_TC_ONE_SM}[debug]
            return code
        except KeyError:
            raise ValueError(f"Invalid {variant=}. Must be one of {set(_lib.VARIANTS)}.") from None

    def set_tuning(self, small_batch_max: int | None = None, tc2_max_grid: int | None = None,
                   host_threads: int | None = None, mapped_max: int | None = None) -> None:
        """Experiment knobs of include/neuralfoil_b200.h (nfb_set_tuning); the defaults are the measured optima."""
        for key, value in ((_lib.TUNE_SMALL_BATCH_MAX, small_batch_max), (_lib.TUNE_TC2_MAX_GRID, tc2_max_grid),
                           (_lib.TUNE_HOST_THREADS, host_threads), (_lib.TUNE_MAPPED_MAX, mapped_max)):
            if value is not None:
                _lib.check(self._lib.nfb_set_tuning(self._ctx, key, int(value)))

    def _host_target(self, n: int):
        """(entry point, handle) for a host-side evaluation of n queries: several devices when there are enough queries."""
        if n >= self.multi_threshold:
            if self._multi is not None and self.n_devices > 1:
                return self._lib.nfb_multi_eval_host, self._multi
            if self._auto:
                twin = self._multi_twin()
                if twin is not None:
                    return self._lib.nfb_multi_eval_host, twin._multi
        return self._lib.nfb_eval_host, self._ctx

    def _multi_twin(self):
        """devices="auto": the all-devices evaluator, created the first time a batch is large enough (None on a one-GPU box)."""
        with self._lock:
            if self._auto_multi is None:
                twin = Evaluator(devices="all", weights_dir=self.weights_dir, extra_models=self._extra_models,
                                 multi_threshold=self.multi_threshold)
                if twin.n_devices < 2:
                    twin.close()
                    self._auto = False
                    return None
                self._auto_multi = twin
            return self._auto_multi

    @staticmethod
    def _rows_and_flag(outputs: str) -> tuple[int, int]:
        """`outputs="all"`: the 198 rows; `"scalars"`: analysis_confidence, CL, CD, CM, Top_Xtr, Bot_Xtr only (rows 0..5)."""
        if outputs == "all":
            return _lib.N_OUTPUT_ROWS, 0
        if outputs == "scalars":
            return _lib.N_SCALAR_ROWS, _lib.OUTPUT_SCALARS
        raise ValueError(f"Invalid {outputs=}. Must be 'all' or 'scalars'.")

    def _model_id(self, model_size: str) -> int:
        if model_size not in self._model_ids:
            raise ValueError(f"Invalid {model_size=}. Must be one of {self.allowable_model_sizes}.")
        return self._model_ids[model_size]

    # -- the drop-in call --------------------------------------------------------------------------
    def get_aero_from_kulfan_parameters(
        self, kulfan_parameters, alpha, Re, n_crit=9.0, xtr_upper=1.0, xtr_lower=1.0, model_size="xlarge",
        variant: str = "fp32", outputs: str = "all",
    ) -> dict[str, np.ndarray]:
        model_id = self._model_id(model_size)  # main.py:154-157
        block = self._evaluate_single_case(model_id, kulfan_parameters, alpha, Re, n_crit, xtr_upper, xtr_lower, variant, outputs)
        if block is not None:  # one case given as scalars: an optimiser's per-iteration call (BASELINE configs[3])
            return dict(zip(_OUTPUT_KEYS, block[:, :1]))
        rows, n_cases = prepare_raw_rows(kulfan_parameters, alpha, Re, n_crit, xtr_upper, xtr_lower)
        block = self._evaluate_host_rows(model_id, rows, n_cases, np.float64, variant=variant, outputs=outputs)
        # Row views of one (198, ld) block: each value is a contiguous (N,) float64 array, as in the
        # reference, where the keys are views of a shared Fortran-ordered block.
        return dict(zip(_OUTPUT_KEYS, block[:, :n_cases]))  # iterating a 2-D array yields its row views

    def evaluate_host(self, raw: np.ndarray, model_size: str = "xlarge", out_dtype=np.float32, chunk: int = 0,
                      variant: str = "fp32", outputs: str = "all"):
        """raw (23, N) float32/float64 host array -> (198, N) host array (rows are contiguous); (6, N) with outputs="scalars"."""
        raw = np.asarray(raw)
        if raw.ndim != 2 or raw.shape[0] != _lib.N_RAW_ROWS:
            raise ValueError(f"raw must have shape (23, N), got {raw.shape}")
        if raw.dtype not in (np.float32, np.float64):
            raw = raw.astype(np.float64)
        rows = [np.ascontiguousarray(raw[i]) for i in range(_lib.N_RAW_ROWS)]
        block = self._evaluate_host_rows(self._model_id(model_size), rows, raw.shape[1], out_dtype, chunk, variant, outputs)
        return block[:, : raw.shape[1]]

    # -- derivatives -----------------------------------------------------------------
    def evaluate_jacobian_host(self, raw: np.ndarray, model_size: str = "xlarge", out_dtype=np.float64):
        """
        raw (23, N) host array -> (values (198, N), jac (N, 198, 23)): the outputs and their derivatives with
        respect to the 23 raw inputs in row order (`JACOBIAN_INPUTS`; alpha per degree).  Forward mode in fp32 on
        the GPU; `values` is bit-identical to `evaluate_host(..., variant="fp32")`.
        """
        raw = np.ascontiguousarray(raw, dtype=np.float64)
        if raw.ndim != 2 or raw.shape[0] != _lib.N_RAW_ROWS:
            raise ValueError(f"raw must have shape (23, N), got {raw.shape}")
        n = raw.shape[1]
        ld = (n + 3) & ~3
        values = np.empty((_lib.N_OUTPUT_ROWS, ld), dtype=out_dtype)
        jac = np.empty((n, _lib.N_OUTPUT_ROWS, _lib.N_RAW_ROWS), dtype=out_dtype)
        ptrs = (C.c_void_p * _lib.N_RAW_ROWS)(*[raw[i].ctypes.data for i in range(_lib.N_RAW_ROWS)])
        _lib.check(
            self._lib.nfb_eval_jacobian_host(
                self._ctx, self._model_id(model_size), n, ptrs, _ONES_23, _lib.F64,
                values.ctypes.data_as(C.c_void_p), ld, jac.ctypes.data_as(C.c_void_p),
                _lib.F64 if np.dtype(out_dtype) == np.float64 else _lib.F32,
            )
        )
        return values[:, :n], jac

    def evaluate_jacobian_device(self, raw, model_size: str = "xlarge", out_dtype=None):
        """raw: (23, N) torch CUDA tensor (float32/float64, contiguous rows) -> (values (198, N), jac (N, 198, 23)) CUDA tensors.
        Enqueues on torch's current stream; does not synchronise."""
        import torch

        if (not isinstance(raw, torch.Tensor) or raw.dim() != 2 or raw.shape[0] != _lib.N_RAW_ROWS or not raw.is_cuda
                or raw.device.index != self.device):
            raise ValueError(f"raw must be a (23, N) tensor on cuda:{self.device}")
        if raw.dtype not in (torch.float32, torch.float64) or raw.stride(1) != 1:
            raise ValueError("raw must be float32 or float64 with contiguous rows")
        n = int(raw.shape[1])
        out_dtype = out_dtype or torch.float32
        ld = (n + 3) & ~3
        dev = f"cuda:{self.device}"
        values = torch.empty((_lib.N_OUTPUT_ROWS, ld), dtype=out_dtype, device=dev)
        jac = torch.empty((n, _lib.N_OUTPUT_ROWS, _lib.N_RAW_ROWS), dtype=out_dtype, device=dev)
        ptrs = (C.c_void_p * _lib.N_RAW_ROWS)(*[raw[i].data_ptr() for i in range(_lib.N_RAW_ROWS)])
        _lib.check(
            self._lib.nfb_eval_jacobian_device(
                self._ctx, self._model_id(model_size), n, ptrs, _ONES_23,
                _lib.F64 if raw.dtype == torch.float64 else _lib.F32,
                C.c_void_p(values.data_ptr()), ld, C.c_void_p(jac.data_ptr()),
                _lib.F64 if out_dtype == torch.float64 else _lib.F32,
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream),
            )
        )
        return values[:, :n], jac

    def evaluate_vjp_device(self, raw, cotangent, model_size: str = "xlarge", reverse: bool | None = None):
        """
        Vector-Jacobian product: raw (23, N) and cotangent (198, N) CUDA tensors (same float dtype for the outputs;
        contiguous rows) -> (values (198, N), grad (23, N)), grad[t] = sum_row cotangent[row] * d output_row / d raw_t:
        the gradient with respect to the raw inputs of any scalar objective of the outputs, without materialising the
        198 x 23 Jacobian (18 KB per query).  Enqueues on torch's current stream.
        `reverse`: True = the fused reverse pass (nfb_grad.cuh; ~3 forward passes per query), False = the forward-mode
        kernel (24 passes per query, but the lower latency for a handful of queries); None = reverse from 256 queries up.
        """
        import torch

        for name, t, rows in (("raw", raw, _lib.N_RAW_ROWS), ("cotangent", cotangent, _lib.N_OUTPUT_ROWS)):
            if (not isinstance(t, torch.Tensor) or t.dim() != 2 or t.shape[0] != rows or not t.is_cuda
                    or t.device.index != self.device or t.stride(1) != 1):
                raise ValueError(f"{name} must be a ({rows}, N) tensor on cuda:{self.device} with contiguous rows")
            if t.dtype not in (torch.float32, torch.float64):
                raise ValueError(f"{name} must be float32 or float64")
        n = int(raw.shape[1])
        if cotangent.shape[1] != n:
            raise ValueError("raw and cotangent must have the same number of columns")
        ld = (n + 3) & ~3
        dev = f"cuda:{self.device}"
        values = torch.empty((_lib.N_OUTPUT_ROWS, ld), dtype=cotangent.dtype, device=dev)
        grad = torch.empty((_lib.N_RAW_ROWS, ld), dtype=cotangent.dtype, device=dev)
        ptrs = (C.c_void_p * _lib.N_RAW_ROWS)(*[raw[i].data_ptr() for i in range(_lib.N_RAW_ROWS)])
        if reverse is None:
            reverse = n >= 256
        entry = self._lib.nfb_eval_vjp_reverse_device if reverse else self._lib.nfb_eval_vjp_device
        _lib.check(
            entry(
                self._ctx, self._model_id(model_size), n, ptrs, _ONES_23,
                _lib.F64 if raw.dtype == torch.float64 else _lib.F32,
                C.c_void_p(cotangent.data_ptr()), int(cotangent.stride(0)),
                C.c_void_p(values.data_ptr()), ld, C.c_void_p(grad.data_ptr()), ld,
                _lib.F64 if cotangent.dtype == torch.float64 else _lib.F32,
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream),
            )
        )
        return values[:, :n], grad[:, :n]

    def evaluate_scalar_grad_device(self, raw, cotangent, model_size: str = "xlarge", with_values: bool = True):
        """
        Reverse-mode gradient of the six scalar outputs for large batches: raw (23, N) and cotangent (6, N) CUDA tensors
        (rows: analysis_confidence, CL, CD, CM, Top_Xtr, Bot_Xtr; contiguous rows) -> (values (6, N) or None, grad (23, N)),
        grad[t] = sum_{i<6} cotangent[i] * d output_i / d raw_t.  About 1.6 forward passes per query (the forward-mode
        `evaluate_vjp_device`, which also takes cotangents of the boundary-layer rows, costs 24).  Enqueues on torch's
        current stream.
        """
        import torch

        for name, t, rows in (("raw", raw, _lib.N_RAW_ROWS), ("cotangent", cotangent, _lib.N_SCALAR_ROWS)):
            if (not isinstance(t, torch.Tensor) or t.dim() != 2 or t.shape[0] != rows or not t.is_cuda
                    or t.device.index != self.device or t.stride(1) != 1):
                raise ValueError(f"{name} must be a ({rows}, N) tensor on cuda:{self.device} with contiguous rows")
            if t.dtype not in (torch.float32, torch.float64):
                raise ValueError(f"{name} must be float32 or float64")
        n = int(raw.shape[1])
        if cotangent.shape[1] != n:
            raise ValueError("raw and cotangent must have the same number of columns")
        ld = (n + 3) & ~3
        dev = f"cuda:{self.device}"
        values = torch.empty((_lib.N_SCALAR_ROWS, ld), dtype=cotangent.dtype, device=dev) if with_values else None
        grad = torch.empty((_lib.N_RAW_ROWS, ld), dtype=cotangent.dtype, device=dev)
        ptrs = (C.c_void_p * _lib.N_RAW_ROWS)(*[raw[i].data_ptr() for i in range(_lib.N_RAW_ROWS)])
        _lib.check(
            self._lib.nfb_eval_scalar_grad_device(
                self._ctx, self._model_id(model_size), n, ptrs, _ONES_23,
                _lib.F64 if raw.dtype == torch.float64 else _lib.F32,
                C.c_void_p(cotangent.data_ptr()), int(cotangent.stride(0)),
                C.c_void_p(values.data_ptr()) if with_values else None, ld, C.c_void_p(grad.data_ptr()), ld,
                _lib.F64 if cotangent.dtype == torch.float64 else _lib.F32,
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream),
            )
        )
        return (values[:, :n] if with_values else None), grad[:, :n]

    def evaluate_scalar_grad_host(self, raw: np.ndarray, cotangent: np.ndarray, model_size: str = "xlarge"):
        """NumPy front door of `evaluate_scalar_grad_device` (moves the arrays with torch): -> (values (6, N), grad (23, N)) float64."""
        import torch

        dev = f"cuda:{self.device}"
        r = torch.from_numpy(np.ascontiguousarray(raw, dtype=np.float64)).to(dev)
        c = torch.from_numpy(np.ascontiguousarray(cotangent, dtype=np.float64)).to(dev)
        values, grad = self.evaluate_scalar_grad_device(r, c, model_size)
        torch.cuda.synchronize(self.device)
        return values.cpu().numpy(), grad.cpu().numpy()

    def evaluate_vjp_host(self, raw: np.ndarray, cotangent: np.ndarray, model_size: str = "xlarge", reverse: bool | None = None):
        """NumPy front door of `evaluate_vjp_device` (moves the arrays with torch): -> (values (198, N), grad (23, N)) float64."""
        import torch

        dev = f"cuda:{self.device}"
        r = torch.from_numpy(np.ascontiguousarray(raw, dtype=np.float64)).to(dev)
        c = torch.from_numpy(np.ascontiguousarray(cotangent, dtype=np.float64)).to(dev)
        values, grad = self.evaluate_vjp_device(r, c, model_size, reverse=reverse)
        return values.cpu().numpy(), grad.cpu().numpy()

    def get_aero_and_jacobian_from_kulfan_parameters(
        self, kulfan_parameters, alpha, Re, n_crit=9.0, xtr_upper=1.0, xtr_lower=1.0, model_size="xlarge"
    ):
        """
        The reference's dict (as `get_aero_from_kulfan_parameters`) plus, under the same keys, the derivatives
        of every output with respect to the 23 scalar inputs: `jac[key]` has shape (N_cases, 23), columns named by
        `JACOBIAN_INPUTS`.  What an optimiser obtains from the reference by tracing main.py through CasADi.
        """
        model_id = self._model_id(model_size)
        rows, n_cases = prepare_raw_rows(kulfan_parameters, alpha, Re, n_crit, xtr_upper, xtr_lower)
        raw = np.empty((_lib.N_RAW_ROWS, n_cases))
        for i, r in enumerate(rows):
            raw[i] = r
        del model_id
        values, jac = self.evaluate_jacobian_host(raw, model_size)
        return dict(zip(_OUTPUT_KEYS, values)), {k: jac[:, i, :] for i, k in enumerate(_OUTPUT_KEYS)}

    def evaluate_host_into(self, raw: np.ndarray, model_size: str, out: np.ndarray, chunk: int = 0,
                           variant: str = "fp32", outputs: str = "all") -> np.ndarray:
        """
        Like `evaluate_host`, but writes into a caller-owned (ideally pinned) block `out` of shape
        (198, ld >= N) with contiguous rows, and reads `raw` (23, N) in place: no host allocation
        or copy happens on this side of the C ABI.  Returns the (198, N) view of `out`.
        """
        if raw.ndim != 2 or raw.shape[0] != _lib.N_RAW_ROWS or raw.dtype not in (np.float32, np.float64):
            raise ValueError("raw must be a (23, N) float32/float64 array")
        if raw.strides[1] != raw.itemsize:
            raise ValueError("raw rows must be contiguous")
        n = raw.shape[1]
        n_rows, flag = self._rows_and_flag(outputs)
        if (out.ndim != 2 or out.shape[0] != n_rows or out.shape[1] < n
                or out.strides[1] != out.itemsize or out.dtype not in (np.float32, np.float64)):
            raise ValueError(f"out must be a ({n_rows}, >=N) float32/float64 array with contiguous rows")
        ptrs = (C.c_void_p * _lib.N_RAW_ROWS)(*[raw[i].ctypes.data for i in range(_lib.N_RAW_ROWS)])
        strides = (C.c_int64 * _lib.N_RAW_ROWS)(*([1] * _lib.N_RAW_ROWS))
        entry, handle = self._host_target(n)
        _lib.check(
            entry(
                handle, self._model_id(model_size), self._variant(variant) | flag, n, ptrs, strides,
                _lib.F64 if raw.dtype == np.float64 else _lib.F32,
                out.ctypes.data_as(C.c_void_p), out.strides[0] // out.itemsize,
                _lib.F64 if out.dtype == np.float64 else _lib.F32, chunk,
            )
        )
        return out[:, :n]

    def _evaluate_single_case(self, model_id, kulfan_parameters, alpha, Re, n_crit, xtr_upper, xtr_lower, variant, outputs):
        """
        Fast path of the drop-in call for ONE case whose flow inputs are scalars and whose Kulfan weights are plain length-8
        vectors: the 23 numbers go straight into a per-thread (23, 1) block whose pointer table is built once.  Returns the
        (198, 4) result block, or None when the arguments are anything else (the general path then validates them and raises
        the reference's errors).  Same library call, same bits.
        """
        try:
            upper, lower = kulfan_parameters["upper_weights"], kulfan_parameters["lower_weights"]
            le, te = kulfan_parameters["leading_edge_weight"], kulfan_parameters["TE_thickness"]
            if len(upper) != 8 or len(lower) != 8:
                return None
            if (np.ndim(alpha) or np.ndim(Re) or np.ndim(n_crit) or np.ndim(xtr_upper) or np.ndim(xtr_lower) or np.ndim(le)
                    or np.ndim(te) or np.ndim(upper) != 1 or np.ndim(lower) != 1):
                return None
            tls = self._tls
            buf = getattr(tls, "buf", None)
            if buf is None:
                buf = tls.buf = np.empty((_lib.N_RAW_ROWS, 1), dtype=np.float64)
                base = buf.__array_interface__["data"][0]
                tls.ptrs = (C.c_void_p * _lib.N_RAW_ROWS)(*range(base, base + _lib.N_RAW_ROWS * 8, 8))
            buf[0:8, 0] = upper
            buf[8:16, 0] = lower
            buf[16:23, 0] = (le, te, alpha, Re, n_crit, xtr_upper, xtr_lower)
        except (TypeError, ValueError, KeyError, IndexError):
            return None
        n_rows, flag = self._rows_and_flag(outputs)
        block = np.empty((n_rows, 4), dtype=np.float64)
        _lib.check(
            self._lib.nfb_eval_host(
                self._ctx, model_id, self._variant(variant) | flag, 1, tls.ptrs, _ONES_23, _lib.F64,
                block.__array_interface__["data"][0], 4, _lib.F64, 0,
            )
        )
        return block

    def _evaluate_host_rows(self, model_id: int, rows, n: int, out_dtype, chunk: int = 0, variant: str = "fp32",
                            outputs: str = "all"):
        ld = (n + 3) & ~3
        n_rows, flag = self._rows_and_flag(outputs)
        block = _result_pool.empty((n_rows, ld), out_dtype)
        if n <= 4096:
            # Small batches (an optimiser's per-iteration call): gather the rows into one (23, n) float64 block so
            # that the pointer table is plain integer arithmetic -- numpy's `.ctypes` accessors cost ~1 us apiece.
            buf = np.empty((_lib.N_RAW_ROWS, n), dtype=np.float64)
            for i, r in enumerate(rows):
                buf[i] = r
            base, step = buf.__array_interface__["data"][0], n * 8
            ptrs = (C.c_void_p * _lib.N_RAW_ROWS)(*range(base, base + _lib.N_RAW_ROWS * step, step))
            strides, in_code = _ONES_23, _lib.F64
        else:
            in_dtype = rows[0].dtype
            if any(r.dtype != in_dtype for r in rows):
                rows = [r.astype(np.float64) for r in rows]
                in_dtype = np.dtype(np.float64)
            buf = rows  # keeps the arrays alive across the call
            ptrs = (C.c_void_p * _lib.N_RAW_ROWS)(*[r.__array_interface__["data"][0] for r in rows])
            strides = (C.c_int64 * _lib.N_RAW_ROWS)(*[0 if r.size == 1 and n > 1 else 1 for r in rows])
            in_code = _lib.F64 if in_dtype == np.float64 else _lib.F32
        entry, handle = self._host_target(n)
        _lib.check(
            entry(
                handle, model_id, self._variant(variant) | flag, n, ptrs, strides, in_code,
                block.__array_interface__["data"][0], ld,
                _lib.F64 if block.dtype == np.float64 else _lib.F32, chunk,
            )
        )
        del buf
        return block

    # -- device-resident call ----------------------------------------------------------------------
    def evaluate_device(self, raw, model_size: str = "xlarge", out=None, out_dtype=None, variant: str = "fp32",
                        outputs: str = "all"):
        """
        raw: torch CUDA tensor (23, N), float32 or float64, rows contiguous -- or a list of 23 1-D CUDA
        tensors of length N or 1 (length-1 rows are broadcast).  Returns a (198, N) CUDA tensor view
        (row stride padded to a multiple of 4).  Enqueues on torch's current stream; does not synchronise.
        """
        import torch

        model_id = self._model_id(model_size)
        if isinstance(raw, torch.Tensor):
            if raw.dim() != 2 or raw.shape[0] != _lib.N_RAW_ROWS:
                raise ValueError(f"raw must have shape (23, N), got {tuple(raw.shape)}")
            rows = list(raw.unbind(0))
        else:
            rows = list(raw)
            if len(rows) != _lib.N_RAW_ROWS:
                raise ValueError(f"expected 23 raw rows, got {len(rows)}")
        n = max(int(r.numel()) for r in rows)
        dtype = rows[0].dtype
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("raw rows must be float32 or float64")
        strides = []
        for r in rows:
            if not r.is_cuda or r.device.index != self.device or r.dtype != dtype or r.dim() != 1:
                raise ValueError(f"every raw row must be a 1-D {dtype} tensor on cuda:{self.device}")
            if r.numel() == n and n > 1:
                strides.append(int(r.stride(0)))
            elif r.numel() == 1:
                strides.append(0)
            else:
                raise ValueError("The inputs to the neural network must all have the same length.")
        if out_dtype is None:
            out_dtype = out.dtype if out is not None else torch.float32
        ld = (n + 3) & ~3
        n_rows, flag = self._rows_and_flag(outputs)
        if out is None:
            out = torch.empty((n_rows, ld), dtype=out_dtype, device=f"cuda:{self.device}")
        else:
            if (not isinstance(out, torch.Tensor) or not out.is_cuda or out.device.index != self.device
                    or out.dtype not in (torch.float32, torch.float64)):
                raise ValueError(f"out must be a float32 or float64 tensor on cuda:{self.device}")
            if out.dim() != 2 or out.shape[0] != n_rows or out.shape[1] < n or out.stride(1) != 1:
                raise ValueError(f"out must be a ({n_rows}, >=N) tensor with contiguous rows")
            ld = int(out.stride(0))
        ptrs = (C.c_void_p * _lib.N_RAW_ROWS)(*[r.data_ptr() for r in rows])
        c_strides = (C.c_int64 * _lib.N_RAW_ROWS)(*strides)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(
            self._lib.nfb_eval_device(
                self._ctx, model_id, self._variant(variant) | flag, n, ptrs, c_strides,
                _lib.F64 if dtype == torch.float64 else _lib.F32,
                C.c_void_p(out.data_ptr()), ld, _lib.F64 if out.dtype == torch.float64 else _lib.F32,
                C.c_void_p(stream),
            )
        )
        return out[:, :n]

    # -- mixed-precision outputs: six float32 scalars + 192 bfloat16 boundary-layer rows (408 instead of 792 B per query) --------
    def evaluate_device_mixed(self, raw, model_size: str = "xlarge", variant: str = "fp32"):
        """
        raw: (23, N) torch CUDA tensor, float32 or float64, contiguous rows -> (scalars (6, N) float32, bl (192, N) bfloat16):
        the six scalars bit-identical to `evaluate_device`, the boundary-layer rows (output rows 6..197) its float32 values
        rounded to nearest bfloat16.  Enqueues on torch's current stream.
        """
        import torch

        if (not isinstance(raw, torch.Tensor) or raw.dim() != 2 or raw.shape[0] != _lib.N_RAW_ROWS or not raw.is_cuda
                or raw.device.index != self.device or raw.dtype not in (torch.float32, torch.float64) or raw.stride(1) != 1):
            raise ValueError(f"raw must be a (23, N) float32 / float64 tensor on cuda:{self.device} with contiguous rows")
        n = int(raw.shape[1])
        ld = (n + 3) & ~3
        dev = f"cuda:{self.device}"
        scalars = torch.empty((_lib.N_SCALAR_ROWS, ld), dtype=torch.float32, device=dev)
        bl = torch.empty((_lib.N_OUTPUT_ROWS - _lib.N_SCALAR_ROWS, ld), dtype=torch.bfloat16, device=dev)
        ptrs = (C.c_void_p * _lib.N_RAW_ROWS)(*[raw[i].data_ptr() for i in range(_lib.N_RAW_ROWS)])
        _lib.check(
            self._lib.nfb_eval_device_mixed(
                self._ctx, self._model_id(model_size), self._variant(variant), n, ptrs, _ONES_23,
                _lib.F64 if raw.dtype == torch.float64 else _lib.F32, C.c_void_p(scalars.data_ptr()), C.c_void_p(bl.data_ptr()), ld,
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream),
            )
        )
        return scalars[:, :n], bl[:, :n]

    def evaluate_host_mixed_into(self, raw: np.ndarray, model_size: str, out_scalars: np.ndarray, out_bl: np.ndarray,
                                 chunk: int = 0, variant: str = "fp32"):
        """
        Host version: raw (23, N) float32/float64 -> out_scalars (6, >=N) float32 and out_bl (192, >=N) uint16 holding bfloat16
        bit patterns (`bf16_bits_to_float32` widens them), both with the same row stride and contiguous rows, ideally pinned.
        Returns their (., N) views.
        """
        if raw.ndim != 2 or raw.shape[0] != _lib.N_RAW_ROWS or raw.dtype not in (np.float32, np.float64) or raw.strides[1] != raw.itemsize:
            raise ValueError("raw must be a (23, N) float32/float64 array with contiguous rows")
        n = raw.shape[1]
        n_bl = _lib.N_OUTPUT_ROWS - _lib.N_SCALAR_ROWS
        if (out_scalars.dtype != np.float32 or out_scalars.ndim != 2 or out_scalars.shape[0] != _lib.N_SCALAR_ROWS or out_scalars.shape[1] < n
                or out_scalars.strides[1] != 4):
            raise ValueError("out_scalars must be a (6, >=N) float32 array with contiguous rows")
        if (out_bl.dtype != np.uint16 or out_bl.ndim != 2 or out_bl.shape[0] != n_bl or out_bl.shape[1] < n or out_bl.strides[1] != 2
                or out_bl.strides[0] // 2 != out_scalars.strides[0] // 4):
            raise ValueError("out_bl must be a (192, >=N) uint16 array with contiguous rows and the row stride (in elements) of out_scalars")
        ptrs = (C.c_void_p * _lib.N_RAW_ROWS)(*[raw[i].ctypes.data for i in range(_lib.N_RAW_ROWS)])
        _lib.check(
            self._lib.nfb_eval_host_mixed(
                self._ctx, self._model_id(model_size), self._variant(variant), n, ptrs, _ONES_23,
                _lib.F64 if raw.dtype == np.float64 else _lib.F32, out_scalars.ctypes.data_as(C.c_void_p),
                out_bl.ctypes.data_as(C.c_void_p), out_scalars.strides[0] // 4, chunk,
            )
        )
        return out_scalars[:, :n], out_bl[:, :n]

    # -- training step ----------------------------------------------------------------
    def train_step_device(self, x, y_data, model_size: str, loss_weights=None, huber_delta: float = 0.05):
        """
        Forward, loss and backward of one iteration of the reference's training loop (training/train.py:56-114, 194-236,
        251-255) for the network loaded under `model_size`, on this evaluator's GPU:
        x (batch, 25) scaled inputs and y_data (batch, 198) targets (NaN = no data), float32 CUDA tensors with contiguous
        rows -> (loss_components (198,) float64, [(dW_l, db_l), ...] float32 in the reference's `net.{2l}.weight/bias`
        shapes).  `loss_components.sum()` is the loss the reference back-propagates.  `loss_weights` defaults to
        `default_loss_weights()`.  Enqueues on torch's current stream; the optimiser step (and `load_model` of the updated
        parameters) stays with the caller.
        """
        import torch

        model_id = self._model_id(model_size)
        dev = f"cuda:{self.device}"
        for name, t, cols in (("x", x, _lib.N_LATENT_INPUTS), ("y_data", y_data, _lib.N_OUTPUT_ROWS)):
            if (not isinstance(t, torch.Tensor) or t.dim() != 2 or t.shape[1] != cols or not t.is_cuda
                    or t.device.index != self.device or t.dtype != torch.float32 or t.stride(1) != 1):
                raise ValueError(f"{name} must be a (batch, {cols}) float32 tensor on {dev} with contiguous rows")
        batch = int(x.shape[0])
        if batch < 1 or y_data.shape[0] != batch:
            raise ValueError("x and y_data must have the same, non-zero number of rows")
        if loss_weights is None:
            loss_weights = default_loss_weights()
        w = torch.as_tensor(np.asarray(loss_weights, dtype=np.float32) if not isinstance(loss_weights, torch.Tensor) else loss_weights,
                            dtype=torch.float32, device=dev).contiguous()
        if w.shape != (_lib.N_OUTPUT_ROWS,):
            raise ValueError("loss_weights must have 198 entries")
        dims = self.model_dims[model_size]
        grads = [(torch.empty((dims[l + 1], dims[l]), dtype=torch.float32, device=dev),
                  torch.empty((dims[l + 1],), dtype=torch.float32, device=dev)) for l in range(len(dims) - 1)]
        loss = torch.empty((_lib.N_OUTPUT_ROWS,), dtype=torch.float64, device=dev)
        n = len(grads)
        gw = (C.c_void_p * n)(*[g.data_ptr() for g, _ in grads])
        gb = (C.c_void_p * n)(*[b.data_ptr() for _, b in grads])
        _lib.check(
            self._lib.nfb_train_step_device(
                self._ctx, model_id, batch, C.c_void_p(x.data_ptr()), int(x.stride(0)), C.c_void_p(y_data.data_ptr()),
                int(y_data.stride(0)), C.c_void_p(w.data_ptr()), float(huber_delta), C.c_void_p(loss.data_ptr()), gw, gb,
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream),
            )
        )
        return loss, grads

    # -- measurement helpers -----------------------------------------------------------------------
    def launch_count(self) -> int:
        """Kernel launches so far, on every device of this evaluator (and of its lazily created multi-device twin)."""
        own = self._lib.nfb_multi_launch_count(self._multi) if self._multi else self._lib.nfb_launch_count(self._ctx)
        return int(own) + (self._auto_multi.launch_count() if self._auto_multi is not None else 0)

    def device_info(self) -> dict:
        vals = [C.c_int() for _ in range(4)]
        _lib.check(self._lib.nfb_device_info(self._ctx, *[C.byref(v) for v in vals]))
        return dict(zip(("device", "sm_count", "cc_major", "cc_minor"), (v.value for v in vals)))

    def measure_fp32_peak_tflops(self, iters: int = 5) -> float:
        out = C.c_double()
        _lib.check(self._lib.nfb_measure_fp32_peak(self._ctx, iters, C.byref(out)))
        return out.value

    def close(self) -> None:
        if getattr(self, "_auto_multi", None) is not None:
            self._auto_multi.close()
            self._auto_multi = None
        if getattr(self, "_multi", None):
            self._lib.nfb_multi_destroy(self._multi)  # owns its contexts
            self._multi = None
            self._ctx = None
        elif getattr(self, "_ctx", None):
            self._lib.nfb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


    # -- the reference's higher-level entry points (main.py:335-468) -----------------------------------
    def get_aero_from_airfoil(self, airfoil, alpha, Re, n_crit=9.0, xtr_upper=1.0, xtr_lower=1.0,
                              model_size="xlarge", variant: str = "fp32") -> dict[str, np.ndarray]:
        """
        `airfoil`: anything with a `.coordinates` (N_points, 2) array in Selig order -- `geometry.Airfoil`, or an
        `aerosandbox.Airfoil` if the caller has that package.  Normalise, fit 8 + 8 + 1 + 1 Kulfan parameters, call
        the learned core at (alpha + rotation, Re / scale), move the moment reference back (main.py:360-387).
        """
        from . import geometry

        kulfan = geometry.kulfan_parameters_of(airfoil)
        if kulfan is not None:
            # A Kulfan-parameterised airfoil (`geometry.KulfanAirfoil`, `aerosandbox.KulfanAirfoil`: main.py:336) with the
            # networks' 8 + 8 weights: by construction its leading edge is at (0, 0) and its chord is the unit x axis, so the
            # reference's normalize() is the identity (rotation 0, scale 1, no translation) and its refit returns the same
            # parameters up to least-squares noise -- the parameters are used as they are.
            nrm = {"x_translation": 0.0, "y_translation": 0.0, "scale_factor": 1.0, "rotation_angle": 0.0}
        else:
            nrm = geometry.normalize(np.asarray(airfoil.coordinates, dtype=np.float64))
            kulfan = geometry.kulfan_parameters_from_coordinates(nrm["coordinates"])
        delta_alpha, scale = nrm["rotation_angle"], nrm["scale_factor"]
        x_qc = -nrm["x_translation"] + 0.25 * (1.0 / scale * np.cos(np.radians(delta_alpha))) - 0.25
        y_qc = -nrm["y_translation"] + 0.25 * (1.0 / scale * np.sin(np.radians(-delta_alpha)))
        aero = self.get_aero_from_kulfan_parameters(
            kulfan, alpha=np.asarray(alpha, dtype=np.float64) + delta_alpha, Re=np.asarray(Re, dtype=np.float64) / scale,
            n_crit=n_crit, xtr_upper=xtr_upper, xtr_lower=xtr_lower, model_size=model_size, variant=variant,
        )
        aero["CM"] = aero["CM"] + (-aero["CL"] * x_qc + aero["CD"] * y_qc)  # main.py:386-387
        return aero

    def get_aero_from_coordinates(self, coordinates, alpha, Re, n_crit=9.0, xtr_upper=1.0, xtr_lower=1.0,
                                  model_size="xlarge", variant: str = "fp32") -> dict[str, np.ndarray]:
        """main.py:392-426"""
        from . import geometry

        return self.get_aero_from_airfoil(geometry.Airfoil(coordinates=coordinates), alpha, Re, n_crit, xtr_upper,
                                          xtr_lower, model_size, variant)

    def get_aero_from_dat_file(self, filename, alpha, Re, n_crit=9.0, xtr_upper=1.0, xtr_lower=1.0,
                               model_size="xlarge", variant: str = "fp32") -> dict[str, np.ndarray]:
        """main.py:429-468"""
        from . import geometry

        with open(filename, "r") as f:
            coordinates = geometry.coordinates_from_raw_dat(f.readlines())
        return self.get_aero_from_coordinates(coordinates, alpha, Re, n_crit, xtr_upper, xtr_lower, model_size, variant)


_default: Evaluator | None = None
_default_lock = threading.Lock()


def default_evaluator() -> Evaluator:
    """
    The process-wide evaluator the module-level functions use (created on first call; safe to call from any thread).
    $NEURALFOIL_B200_DEVICES = "all" or "0,1,..." replicates it on several GPUs from the start.  Otherwise it lives on one
    GPU -- $NEURALFOIL_B200_DEVICE, else $LOCAL_RANK (one process per GPU under torchrun) -- and, when neither is set,
    brings in the box's other GPUs the first time a call carries at least 262,144 queries (`devices="auto"`).
    """
    global _default
    with _default_lock:
        if _default is None:
            spec = os.environ.get("NEURALFOIL_B200_DEVICES")
            if spec:
                _default = Evaluator(devices="all" if spec == "all" else [int(d) for d in spec.split(",")])
            elif "NEURALFOIL_B200_DEVICE" in os.environ or "LOCAL_RANK" in os.environ:
                _default = Evaluator()
            else:
                _default = Evaluator(device=0, devices="auto")
        return _default


def get_aero_from_kulfan_parameters(
    kulfan_parameters, alpha, Re, n_crit=9.0, xtr_upper=1.0, xtr_lower=1.0, model_size="xlarge"
) -> dict[str, np.ndarray]:
    """
    Drop-in for the reference's `neuralfoil.get_aero_from_kulfan_parameters` (main.py:64-332): same
    arguments, same 198-key dict of float64 arrays of shape (N_cases,), same ValueErrors.  The values
    come from the fp32 CUDA path and agree with the reference's float64 NumPy path to the tolerance
    stated in DESIGN.md.
    """
    return default_evaluator().get_aero_from_kulfan_parameters(
        kulfan_parameters, alpha, Re, n_crit=n_crit, xtr_upper=xtr_upper, xtr_lower=xtr_lower, model_size=model_size
    )


def get_aero_and_jacobian_from_kulfan_parameters(
    kulfan_parameters, alpha, Re, n_crit=9.0, xtr_upper=1.0, xtr_lower=1.0, model_size="xlarge"
):
    """`get_aero_from_kulfan_parameters` plus d(every output)/d(every scalar input): see the `Evaluator` method."""
    return default_evaluator().get_aero_and_jacobian_from_kulfan_parameters(
        kulfan_parameters, alpha, Re, n_crit=n_crit, xtr_upper=xtr_upper, xtr_lower=xtr_lower, model_size=model_size
    )


def get_aero_from_airfoil(airfoil, alpha, Re, n_crit=9.0, xtr_upper=1.0, xtr_lower=1.0, model_size="xlarge"):
    """Drop-in for the reference's `get_aero_from_airfoil` (main.py:335-389); geometry by `neuralfoil_b200.geometry`."""
    return default_evaluator().get_aero_from_airfoil(airfoil, alpha, Re, n_crit, xtr_upper, xtr_lower, model_size)


def get_aero_from_coordinates(coordinates, alpha, Re, n_crit=9.0, xtr_upper=1.0, xtr_lower=1.0, model_size="xlarge"):
    """Drop-in for the reference's `get_aero_from_coordinates` (main.py:392-426)."""
    return default_evaluator().get_aero_from_coordinates(coordinates, alpha, Re, n_crit, xtr_upper, xtr_lower, model_size)


def get_aero_from_dat_file(filename, alpha, Re, n_crit=9.0, xtr_upper=1.0, xtr_lower=1.0, model_size="xlarge"):
    """Drop-in for the reference's `get_aero_from_dat_file` (main.py:429-468)."""
    return default_evaluator().get_aero_from_dat_file(filename, alpha, Re, n_crit, xtr_upper, xtr_lower, model_size)
