"""
CPU-side checks of the C ABI: the library builds and loads, it exports exactly the entry points
include/neuralfoil_b200.h declares, the ctypes table agrees with the header, and -- with no GPU in
this container -- the library refuses to run instead of falling back to anything.
No compute entry point is called here.
"""

import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]
HEADER = REPO / "include" / "neuralfoil_b200.h"


def declared_functions() -> list[str]:
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(nfb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("nfb_create", "nfb_destroy", "nfb_load_model", "nfb_set_distribution", "nfb_eval_device",
                 "nfb_eval_host", "nfb_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(library):
    from neuralfoil_b200 import _lib

    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True)
    exported = {line.split()[-1] for line in out.stdout.splitlines() if " T " in line}
    for name in declared_functions():
        assert name in exported, f"{name} declared in the header but not exported"
        assert hasattr(library, name)
    extra = {s for s in exported if s.startswith("nfb_")} - set(declared_functions())
    assert not extra, f"exported but undeclared: {extra}"


def test_ctypes_table_matches_header(library):
    from neuralfoil_b200 import _lib

    assert sorted(_lib._SIGNATURES) == declared_functions()


def test_library_is_sm100a_only(library):
    from neuralfoil_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_version_and_error_strings(library):
    import neuralfoil_b200

    assert b"sm_100a" in library.nfb_version()
    assert f"neuralfoil_b200 {neuralfoil_b200.__version__} ".encode() in library.nfb_version()  # one version, two places
    assert isinstance(library.nfb_last_error(), bytes)


def test_no_cpu_fallback_without_a_device(library):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is exercised in the CPU container")
    from neuralfoil_b200 import _lib, Evaluator

    handle = C.c_void_p()
    rc = library.nfb_create(0, C.byref(handle))
    assert rc == _lib.ERR_NO_DEVICE and not handle.value
    assert b"no CPU fallback" in library.nfb_last_error()
    with pytest.raises(_lib.NeuralFoilB200Error):
        Evaluator(device=0)


def test_argument_errors_need_no_device(library):
    from neuralfoil_b200 import _lib

    assert library.nfb_create(0, None) == _lib.ERR_INVALID_ARGUMENT
    assert library.nfb_eval_host(None, 0, 0, 1, None, None, 0, None, 1, 0, 0) == _lib.ERR_INVALID_ARGUMENT
    assert library.nfb_launch_count(None) == 0
    assert library.nfb_multi_eval_host(None, 0, 0, 1, None, None, 0, None, 1, 0, 0) == _lib.ERR_INVALID_ARGUMENT
    assert library.nfb_multi_device_count(None) == 0 and library.nfb_multi_launch_count(None) == 0
    assert library.nfb_set_tuning(None, 0, 0) == _lib.ERR_INVALID_ARGUMENT
    handle = C.c_void_p()
    assert library.nfb_multi_create(None, 0, C.byref(handle)) in (_lib.ERR_NO_DEVICE, _lib.OK)  # OK only on a GPU box
    if handle.value:
        library.nfb_multi_destroy(handle)
