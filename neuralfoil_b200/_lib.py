"""
ctypes binding of include/neuralfoil_b200.h.  This is the only place the package touches the
shared library; there is no fallback of any kind -- if the library is missing or no B200 is
present, the calls raise.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libneuralfoil_b200.so"
if os.environ.get("NFB_LIB_PATH"):  # A/B experiments only: another build of the same library
    LIB_PATH = Path(os.environ["NFB_LIB_PATH"])

N_RAW_ROWS = 23
N_LATENT_INPUTS = 25
N_OUTPUT_ROWS = 198
F32, F64 = 0, 1
VARIANT_FP32_FFMA = 0
VARIANT_TENSOR_FP16 = 1
VARIANT_TENSOR_FP16X3 = 2
OUTPUT_SCALARS = 0x100  # NFB_OUTPUT_SCALARS: OR-ed into the variant, only the six scalar rows are written
DEBUG_FFMA_TILED = 0x1000  # cross-check kernels (tests / A-B only), OR-ed into the variant
DEBUG_TC_ONE_SM = 0x2000
DEBUG_FLAGS = {"ffma_tiled": DEBUG_FFMA_TILED, "tc_one_sm": DEBUG_TC_ONE_SM}
TUNE_SMALL_BATCH_MAX, TUNE_TC2_MAX_GRID, TUNE_HOST_THREADS, TUNE_MAPPED_MAX = 0, 1, 2, 3
N_SCALAR_ROWS = 6
VARIANTS = {"fp32": VARIANT_FP32_FFMA, "tensor": VARIANT_TENSOR_FP16, "tensor_fp32": VARIANT_TENSOR_FP16X3}

OK = 0
ERR_INVALID_ARGUMENT = -1
ERR_UNSUPPORTED_MODEL = -2
ERR_NOT_LOADED = -3
ERR_CUDA = -4
ERR_NO_DEVICE = -5

# Every symbol include/neuralfoil_b200.h declares (tests/test_cabi_symbols.py checks the two agree).
_SIGNATURES = {
    "nfb_version": (C.c_char_p, []),
    "nfb_last_error": (C.c_char_p, []),
    "nfb_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "nfb_destroy": (None, [C.c_void_p]),
    "nfb_device_info": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int)] * 4),
    "nfb_set_distribution": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "nfb_load_model": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)],
    ),
    "nfb_eval_device": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int,
         C.c_void_p, C.c_int64, C.c_int, C.c_void_p],
    ),
    "nfb_eval_host": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int,
         C.c_void_p, C.c_int64, C.c_int, C.c_int64],
    ),
    "nfb_eval_device_mixed": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int,
         C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p],
    ),
    "nfb_eval_host_mixed": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int,
         C.c_void_p, C.c_void_p, C.c_int64, C.c_int64],
    ),
    "nfb_eval_jacobian_device": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int,
         C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p],
    ),
    "nfb_eval_vjp_device": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int,
         C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p],
    ),
    "nfb_eval_scalar_grad_device": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int,
         C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p],
    ),
    "nfb_eval_vjp_reverse_device": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int,
         C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p],
    ),
    "nfb_eval_jacobian_host": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int,
         C.c_void_p, C.c_int64, C.c_void_p, C.c_int],
    ),
    "nfb_train_step_device": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_float,
         C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p],
    ),
    "nfb_set_tuning": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "nfb_multi_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "nfb_multi_destroy": (None, [C.c_void_p]),
    "nfb_multi_device_count": (C.c_int, [C.c_void_p]),
    "nfb_multi_context": (C.c_void_p, [C.c_void_p, C.c_int]),
    "nfb_multi_set_distribution": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "nfb_multi_load_model": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)],
    ),
    "nfb_multi_eval_host": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int,
         C.c_void_p, C.c_int64, C.c_int, C.c_int64],
    ),
    "nfb_shard_range": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "nfb_multi_launch_count": (C.c_int64, [C.c_void_p]),
    "nfb_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "nfb_host_free": (C.c_int, [C.c_void_p]),
    "nfb_host_register": (C.c_int, [C.c_void_p, C.c_int64]),
    "nfb_host_unregister": (C.c_int, [C.c_void_p]),
    "nfb_launch_count": (C.c_int64, [C.c_void_p]),
    "nfb_debug_tc_stamps": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]),
    "nfb_measure_fp32_peak": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
}

_lib = None


class NeuralFoilB200Error(RuntimeError):
    """A failure inside the CUDA library that is not a caller mistake."""


def load() -> C.CDLL:
    """dlopen the library (once) and declare the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise NeuralFoilB200Error(
            f"{LIB_PATH} is missing: build it with `python -m neuralfoil_b200._build` "
            "(neuralfoil_b200 has no CPU or PyTorch fallback)."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Map a status code to the exception the reference's callers expect."""
    if rc == OK:
        return
    msg = load().nfb_last_error().decode("utf-8", "replace")
    if rc in (ERR_INVALID_ARGUMENT, ERR_UNSUPPORTED_MODEL, ERR_NOT_LOADED):
        raise ValueError(msg)
    raise NeuralFoilB200Error(msg)
