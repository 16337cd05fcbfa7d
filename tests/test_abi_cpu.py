"""The C-ABI library loads on a machine without a GPU and exports every symbol include/epgx.h
declares; host-only entry points (plan creation / validation / configuration) work; compute entry
points fail loudly instead of falling back to the CPU."""

import ctypes
import os
import re

import numpy as np
import pytest

import cases
from util import product_namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "epgx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(epgx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from epgpy_b200 import engine

    L = engine.lib()
    syms = declared_symbols()
    assert set(syms) == set(engine.EXPORTS), (syms, engine.EXPORTS)
    for s in syms:
        assert getattr(L, s) is not None
    assert L.epgx_version() == 109


def test_struct_sizes_match_header():
    from epgpy_b200 import engine, lowering

    assert lowering.OP_DTYPE.itemsize == 32 and lowering.SEG_DTYPE.itemsize == 32
    # epgx_tape: 8 + 64 + 8 + 64*8*4 + 64*4 + 6*8 + 8 + 4 + 4 + 4*4 + 3*4, padded to 8
    assert ctypes.sizeof(engine._Tape) == 2504
    assert ctypes.sizeof(engine.Config) == 64


def test_plan_create_and_config_host_only():
    from epgpy_b200 import engine, lowering

    epg = product_namespace()
    case = cases.fisp(epg, 40)
    low = lowering.lower(case["seq"])
    plan = engine.Plan(low)
    cfg = plan.config()
    assert cfg["ring"] == low.max_order + 1 and cfg["var_tiles"] == 1 and cfg["threads_per_cta"] <= 256
    assert cfg["flops_per_atom"] > 0 and cfg["updates_per_atom"] > 0
    assert plan.workspace_bytes() >= low.coef.nbytes
    plan.set_variant(lanes_per_atom=8, atoms_per_cta=4)
    assert plan.config()["lanes_per_atom"] == 8 and plan.config()["atoms_per_cta"] == 4
    # float32 plan of the same tape
    low32 = lowering.lower(case["seq"], dtype="f32")
    assert engine.Plan(low32).workspace_bytes() < plan.workspace_bytes()


def test_plan_validation_rejects_bad_tapes():
    from epgpy_b200 import engine, lowering

    epg = product_namespace()
    low = lowering.lower(cases.readme_mse(epg)["seq"])
    low.ops = low.ops.copy()
    low.ops["off"][3][0] = 10 ** 9  # coefficient block far outside the table
    with pytest.raises(engine.EpgxError, match="out of range"):
        engine.Plan(low)
    low = lowering.lower(cases.readme_mse(epg)["seq"])
    low.segs = low.segs.copy()
    low.segs["shift"][0] = 3  # only unit shifts are valid
    with pytest.raises(engine.EpgxError, match="bad segment"):
        engine.Plan(low)


def test_capacity_error_is_loud():
    """an atom whose state cannot live on chip is refused, not silently spilled"""
    from epgpy_b200 import engine, lowering

    epg = product_namespace()
    seq = [epg.T(30, 0), epg.E(5, 1000.0, 50.0), epg.ADC, epg.S(1)] * 6000
    low = lowering.lower(seq, prune_unobservable=False)
    with pytest.raises(MemoryError):
        engine.Plan(low)


def test_no_cpu_fallback():
    """without a CUDA device the compute path raises (no CPU fallback)"""
    import torch

    from epgpy_b200 import engine

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    epg = product_namespace()
    with pytest.raises(engine.EpgxError, match="no CPU fallback"):
        epg.simulate(cases.readme_mse(epg)["seq"])
    sig = np.zeros((20, 3, 1), dtype=np.complex128)
    from epgpy_b200 import lowering

    plan = engine.Plan(lowering.lower(cases.readme_mse(epg)["seq"]))
    with pytest.raises(engine.EpgxError):
        plan.run_host(0, 0, 3, sig)


def test_expand_real_rows_on_the_host():
    """epgx_expand_real: rows of reals -> complex rows with zero imaginary parts, pitched source and destination, any
    alignment, several threads (pure host marshalling: runs without a GPU)"""
    import numpy as np

    from epgpy_b200 import engine

    L = engine.lib()
    rng = np.random.RandomState(0)
    for dt, cdt, code in ((np.float64, np.complex128, 0), (np.float32, np.complex64, 1)):
        for rows, cols, sp, dp, off, nth in ((7, 1003, 1100, 5000, 3, 4), (1, 5, 5, 5, 0, 8), (33, 64, 64, 200, 1, 1)):
            src = rng.rand(rows, sp).astype(dt)
            dst = np.full((rows, dp), 1 + 1j, dtype=cdt)
            assert L.epgx_expand_real(code, src.ctypes.data, sp, dst.ctypes.data + off * dst.itemsize, dp, rows, cols, nth) == 0
            want = np.full((rows, dp), 1 + 1j, dtype=cdt)
            want[:, off:off + cols] = src[:, :cols]
            assert np.array_equal(dst, want)
    assert L.epgx_expand_real(0, None, 1, None, 1, 1, 1, 1) < 0
