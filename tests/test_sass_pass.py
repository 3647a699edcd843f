"""The SASS post-pass of the build (neuralfoil_b200/_sass.py): what it may touch, and that it is a fixed point.  CPU only."""
import os
import shutil
import struct

import pytest

from neuralfoil_b200 import _build, _sass


@pytest.fixture(scope="module")
def lib_path():
    path = _build.build()
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    return path


def test_default_selection_is_the_eight_warp_value_kernels():
    yes = [
        "_ZN3nfb22nfb_ffma_groups_kernelINS_8GroupCfgILi512ELi8ELi5ELi16ELi16ELi0EEEEEvNS_10EvalParamsE",
        "_ZN3nfb22nfb_ffma_groups_kernelINS_8GroupCfgILi256ELi8ELi8ELi16ELi32ELi0EEEEEvNS_10EvalParamsE",
        "_ZN3nfb15nfb_ffma_kernelINS_7FfmaCfgILi512ELi8ELi4ELi16EEEEEvNS_10EvalParamsE",
    ]
    no = [
        "_ZN3nfb22nfb_ffma_groups_kernelINS_8GroupCfgILi128ELi8ELi16ELi8ELi32ELi0EEEEEvNS_10EvalParamsE",  # 16 warps: loses
        "_ZN3nfb22nfb_ffma_groups_kernelINS_8GroupCfgILi64ELi8ELi16ELi8ELi32ELi0EEEEEvNS_10EvalParamsE",
        "_ZN3nfb15nfb_grad_kernelINS_8GroupCfgILi512ELi8ELi5ELi8ELi16ELi0EEELb1EEEvNS_10GradParamsE",
        "_ZN3nfb23nfb_train_fwdbwd_kernelINS_8GroupCfgILi256ELi8ELi8ELi8ELi32ELi0EEEEEvNS_11TrainParamsE",
        "_ZN3nfb14nfb_jac_kernelILi512EEEvNS_9JacParamsE",
    ]
    assert all(_sass.DEFAULT_SELECT.search(n) for n in yes)
    assert not any(_sass.DEFAULT_SELECT.search(n) for n in no)


def test_pass_is_a_fixed_point_and_touches_hint_bits_only(lib_path, tmp_path):
    if os.environ.get("NFB_NO_SASS_PASS", "") == "1":
        pytest.skip("the library was built without the pass (A/B build)")
    copy = tmp_path / "lib.so"
    shutil.copyfile(lib_path, copy)
    stats = _sass.patch_file(str(copy), verbose=False, select=_sass.DEFAULT_SELECT)
    assert stats["functions"] >= 2 and stats["ffma2"] >= 1500 and stats["moved"] == 0
    # the shipped library has been through the pass already: a second application changes nothing
    assert copy.read_bytes() == lib_path.read_bytes()
    assert stats["chain_starts_after"] == stats["chain_starts_before"]
    assert _sass.verify_hint_bits_only(str(lib_path), str(copy)) == 0

    # the selected kernels really carry the flags: most FFMA2s find a source in the reuse cache, none yields
    for name, body in _sass.functions(str(lib_path)):
        if "nfb_ffma_groups_kernelINS_8GroupCfgILi512" not in name:
            continue
        ffma2 = [(t, hi) for _, t, _, hi in body if t.startswith("FFMA2 ")]
        assert len(ffma2) == 768
        assert all(hi & _sass.YIELD_BIT for _, hi in ffma2)
        assert sum(1 for t, _ in ffma2 if ".reuse" in t) >= 500
        break
    else:
        pytest.fail("xxxlarge kernel not found in the library")


def test_verifier_rejects_any_other_change(lib_path, tmp_path):
    data = bytearray(lib_path.read_bytes())
    name, body = next((n, b) for n, b in _sass.functions(str(lib_path)) if "nfb_ffma_groups_kernelINS_8GroupCfgILi512" in n)
    head = b"".join(struct.pack("<QQ", lo, hi) for _, _, lo, hi in body)
    base = data.find(head)
    assert base >= 0
    k = next(i for i, (_, t, _, _) in enumerate(body) if t.startswith("FFMA2 "))
    for what, off, bit in (("a register field of an FFMA2", base + 16 * k, 17), ("the stall count of an FFMA2", base + 16 * k + 8, 41),
                           ("a hint bit of another instruction", base + 8, 45)):
        bad = bytearray(data)
        word = struct.unpack_from("<Q", bad, off)[0] ^ (1 << bit)
        struct.pack_into("<Q", bad, off, word)
        f = tmp_path / "bad.so"
        f.write_bytes(bad)
        with pytest.raises(ValueError):
            _sass.verify_hint_bits_only(str(lib_path), str(f))
    ok = bytearray(data)
    word = struct.unpack_from("<Q", ok, base + 16 * k + 8)[0] ^ (1 << 58)
    struct.pack_into("<Q", ok, base + 16 * k + 8, word)
    f = tmp_path / "ok.so"
    f.write_bytes(ok)
    assert _sass.verify_hint_bits_only(str(lib_path), str(f)) == 1
