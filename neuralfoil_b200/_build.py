"""
Builds the CUDA library in-tree:  neuralfoil_b200/lib/libneuralfoil_b200.so  (sm_100a only).

nvcc cross-compiles without a GPU, so this runs in the CPU build container; the resulting `.so` is
git-ignored but travels to the GPU box with the repository snapshot.

    python -m neuralfoil_b200._build [--force] [--verbose]
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libneuralfoil_b200.so"

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "--no-compress",  # the SASS pass (_sass.py) edits the code bytes in place
    "--shared",
    "-Xcompiler",
    "-fPIC",
]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _dependencies() -> list[Path]:
    return _sources() + sorted(CSRC.glob("*.cuh")) + sorted(INCLUDE.glob("*.h")) + [Path(__file__), PKG_DIR / "_sass.py"]


def is_stale() -> bool:
    if not LIB_PATH.exists():
        return True
    built = LIB_PATH.stat().st_mtime
    return any(p.stat().st_mtime > built for p in _dependencies())


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise RuntimeError("nvcc not found; neuralfoil_b200 needs the CUDA toolkit to build its library")
    return nvcc


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ into one shared library.  Returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    LIB_DIR.mkdir(exist_ok=True)
    tmp = LIB_PATH.with_suffix(f".so.tmp{os.getpid()}")
    extra = os.environ.get("NFB_NVCC_EXTRA", "").split()  # tuning experiments only, e.g. "-DNFB_TN=8"
    cmd = [find_nvcc(), *NVCC_FLAGS, *extra, "-I", str(INCLUDE), "-o", str(tmp), *map(str, _sources())]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
        print(" ".join(cmd), flush=True)
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        tmp.unlink(missing_ok=True)
        raise RuntimeError(f"nvcc failed with exit code {proc.returncode}")
    if os.environ.get("NFB_NO_SASS_PASS", "") != "1":  # A/B only: the library as ptxas left it
        # operand-reuse flags of the FFMA2 blocks of the 8-warp fp32 kernels (see _sass.py): same instructions, same order,
        # bit-identical results.  The pass edits a copy and is verified (only scheduling-hint bits of FFMA2s may differ) before
        # the copy replaces the linked library; any failure leaves the library as ptxas wrote it -- slower by 2 %, never wrong.
        from . import _sass

        patched = tmp.with_suffix(tmp.suffix + ".sass")
        try:
            shutil.copyfile(tmp, patched)
            stats = _sass.patch_file(str(patched), verbose=verbose, select=_sass.DEFAULT_SELECT)
            _sass.verify_hint_bits_only(str(tmp), str(patched))
            os.replace(patched, tmp)
            os.chmod(tmp, 0o755)
            if verbose:
                print(f"sass pass: {stats}", flush=True)
        except Exception as exc:  # cuobjdump missing, unexpected encoding, ...
            patched.unlink(missing_ok=True)
            sys.stderr.write(f"neuralfoil_b200: SASS pass skipped ({type(exc).__name__}: {exc}); the library is unpatched\n")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
