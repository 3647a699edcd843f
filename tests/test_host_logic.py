"""
Host-side logic of the drop-in call, checked on the CPU: validation and case counting
(reference main.py:154-199), layer discovery (main.py:206-215), output naming (main.py:318-331),
the shard partition rule, and the product package's independence from the oracle.
"""

import ast
from pathlib import Path

import numpy as np
import pytest

from neuralfoil_b200 import main as nfb_main
from neuralfoil_b200 import sharding, synthetic
from oracle import neuralfoil_oracle as orc

REPO = Path(__file__).resolve().parents[1]
KULFAN = synthetic.FIXED_KULFAN


def test_output_keys_match_oracle_and_reference_order():
    assert nfb_main.output_keys() == orc.output_keys()
    keys = nfb_main.output_keys()
    assert keys[:6] == ["analysis_confidence", "CL", "CD", "CM", "Top_Xtr", "Bot_Xtr"]
    assert keys[6] == "upper_bl_theta_0" and keys[38] == "upper_bl_H_0" and keys[70] == "upper_bl_ue/vinf_0"
    assert keys[102] == "lower_bl_theta_0" and keys[197] == "lower_bl_ue/vinf_31" and len(keys) == 198


def test_bl_x_points_contract():
    """reference tests/test_invariants.py:139-144"""
    x = nfb_main.bl_x_points
    assert len(x) == 32 and np.all(np.diff(x) > 0) and np.all((x > 0) & (x < 1))
    np.testing.assert_array_equal(x, orc.bl_x_points)


def test_prepare_rows_scalars_and_vectors():
    rows, n = nfb_main.prepare_raw_rows(KULFAN, 5.0, 1e6, 9.0, 1.0, 1.0)
    assert n == 1 and len(rows) == 23 and all(r.shape == (1,) and r.dtype == np.float64 for r in rows)
    rows, n = nfb_main.prepare_raw_rows(KULFAN, np.linspace(0, 1, 7), 1e6, 9, np.array([0.5]), 1)
    assert n == 7 and rows[18].shape == (7,) and rows[19].shape == (1,) and rows[21].shape == (1,)
    want = orc.raw_rows_from_arguments(KULFAN, np.linspace(0, 1, 7), 1e6, 9, np.array([0.5]), 1)
    got = np.stack([np.broadcast_to(r, (7,)) for r in rows])
    np.testing.assert_array_equal(got, want)


def test_prepare_rows_8_by_n_weights():
    raw = synthetic.sample_training_distribution(9, seed=0)
    kulfan, flow = synthetic.kulfan_dict_from_raw(raw)
    rows, n = nfb_main.prepare_raw_rows(kulfan, flow["alpha"], flow["Re"], flow["n_crit"], flow["xtr_upper"], flow["xtr_lower"])
    assert n == 9
    np.testing.assert_array_equal(np.stack(rows), raw)
    assert all(r.flags["C_CONTIGUOUS"] for r in rows)


def test_prepare_rows_errors():
    """reference tests/test_robustness.py:60-83"""
    with pytest.raises(ValueError, match="same length"):
        nfb_main.prepare_raw_rows(KULFAN, np.linspace(0, 5, 3), np.full(4, 1e6), 9.0, 1.0, 1.0)
    bad = dict(KULFAN, upper_weights=np.full(7, 0.2), lower_weights=np.full(7, -0.1))
    with pytest.raises(ValueError, match="8 CST weights per side"):
        nfb_main.prepare_raw_rows(bad, 5.0, 1e6, 9.0, 1.0, 1.0)
    n_by_8 = dict(KULFAN, upper_weights=np.zeros((5, 8)), lower_weights=np.zeros((5, 8)))
    with pytest.raises(ValueError, match="8 CST weights per side"):
        nfb_main.prepare_raw_rows(n_by_8, 5.0, 1e6, 9.0, 1.0, 1.0)
    with pytest.raises(ValueError):
        nfb_main.prepare_raw_rows(KULFAN, np.zeros((3, 2)), 1e6, 9.0, 1.0, 1.0)
    with pytest.raises(ValueError):
        nfb_main.prepare_raw_rows(KULFAN, np.zeros(0), 1e6, 9.0, 1.0, 1.0)


def test_layer_discovery_and_file_format_error():
    params = synthetic.synthetic_xxxlarge_params()
    layers = nfb_main.layers_from_params(params)
    assert [w.shape for w, _ in layers] == [(512, 25)] + [(512, 512)] * 5 + [(198, 512)]
    assert all(w.dtype == np.float32 and w.flags["C_CONTIGUOUS"] for w, _ in layers)
    with pytest.raises(ValueError, match="unexpected neural network file format"):
        nfb_main.layers_from_params({"weights": np.zeros(3)})


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 8, 9, 1000, 10_000_000):
        for world in (1, 2, 3, 4, 8):
            ranges = [sharding.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [e - s for s, e in ranges]
            assert max(sizes) <= ((-(-n // world) + 3) & ~3) and all(s % 4 == 0 or s == n for s, _ in ranges)
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure; a product path routed through it would void parity."""
    for path in (REPO / "neuralfoil_b200").rglob("*.py"):
        tree = ast.parse(path.read_text())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n.split(".")[0] == "oracle" for n in names), f"{path} imports the oracle"
    for path in (REPO / "neuralfoil_b200" / "csrc").glob("*"):
        assert "oracle" not in path.read_text()


def test_reference_import_name_is_served():
    """reference neuralfoil/__init__.py:1-15: third-party callers `import neuralfoil`; the alias package re-exports the B200 path."""
    import neuralfoil
    import neuralfoil.main as ref_style_main
    import neuralfoil_b200

    assert neuralfoil.__all__ == ["get_aero_from_kulfan_parameters", "get_aero_from_airfoil", "get_aero_from_coordinates",
                                  "get_aero_from_dat_file", "bl_x_points"]
    for name in neuralfoil.__all__:
        assert getattr(neuralfoil, name) is getattr(neuralfoil_b200, name)
        assert getattr(ref_style_main, name) is getattr(neuralfoil_b200, name)
    assert neuralfoil.__version__.startswith("0.3.3+b200.")


def test_evaluator_rejects_conflicting_device_arguments(library):
    with pytest.raises(ValueError, match="either"):
        nfb_main.Evaluator(device=0, devices=[0, 1])
    with pytest.raises(ValueError, match="Invalid devices"):
        nfb_main.Evaluator(devices="every")
    with pytest.raises(ValueError, match="Invalid variant"):
        nfb_main.Evaluator._variant("fp32:fast")
    assert nfb_main.Evaluator._variant("fp32:tiled") == 0x1000 and nfb_main.Evaluator._variant("tensor_fp32:one_sm") == 0x2002


def test_result_pool_recycles_only_released_blocks():
    """Large result blocks come from a pool (main._ResultPool): memory returns to it when the LAST view dies, never before."""
    import gc

    from neuralfoil_b200.main import _ResultPool

    pool = _ResultPool(64 << 20)
    a = pool.empty((198, 20000), np.float64)  # 31.7 MB >= MIN_BYTES
    assert a.shape == (198, 20000) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.flags["WRITEABLE"]
    a[:] = 3.0
    ptr = a.__array_interface__["data"][0]
    views = [a[i, :19999] for i in range(198)]  # what the returned dict holds
    del a
    gc.collect()
    assert len(pool._free) == 0, "a block still referenced by a row view must not be recycled"
    b = pool.empty((198, 20000), np.float64)
    assert b.__array_interface__["data"][0] != ptr
    assert float(views[7][5]) == 3.0
    del views
    gc.collect()
    assert len(pool._free) == 1
    c = pool.empty((198, 20000), np.float64)
    assert c.__array_interface__["data"][0] == ptr and len(pool._free) == 0  # same size: reused
    d = pool.empty((198, 10000), np.float64)
    assert d.__array_interface__["data"][0] != ptr  # other size: fresh
    del b, c, d
    gc.collect()
    assert sum(m.nbytes for m in pool._free) <= pool.cap_bytes  # 31.7 + 31.7 + 15.8 MB > 64 MB: the oldest went
    small = pool.empty((198, 100), np.float64)
    assert small.base is None  # below MIN_BYTES: a plain NumPy array
    off = _ResultPool(0)
    assert off.empty((198, 20000), np.float64).base is None
