"""The C-ABI library loads on a CPU-only box, exports every symbol include/tpsb200.h declares, and fails
loudly (no CPU fallback) when a compute entry point is used without a CUDA device."""
import os
import re

import numpy as np
import pytest

import tps_b200
from tps_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported(lib_built):
    hdr = open(os.path.join(ROOT, "include", "tpsb200.h")).read()
    declared = set(re.findall(r"\b(tpsb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(capi.EXPORTS)
    for name in declared:
        assert hasattr(lib_built, name), name


def test_no_cpu_fallback(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(tps_b200.TpsbError, match="no CPU fallback"):
        tps_b200.RhsOperator(tps_b200.cartesian_hex_mesh(3, 3, 3))


def test_reference_tables_are_consistent(lib_built):
    for p in (1, 2, 3):
        T = capi.ref_tables(p)
        npn = T["np"]
        assert T["nq"] == npn + 1
        assert abs(T["wn"].sum() - 1) < 1e-15 and abs(T["wq"].sum() - 1) < 1e-15
        # differentiation matrix kills constants, differentiates x exactly
        assert np.abs(T["D"].sum(1)).max() < 1e-13
        assert np.abs(T["D"] @ T["xn"] - 1).max() < 1e-13
        # interpolation/extrapolation rows are partitions of unity
        assert np.abs(T["P"].sum(1) - 1).max() < 1e-14 and np.abs(T["lb"].sum(1) - 1).max() < 1e-14
        # exactness that licenses the collapsed gradient face term: (p+2)-pt rule of l_a * l_b = w_a delta_ab
        M = T["P"].T @ np.diag(T["wq"]) @ T["P"]
        assert np.abs(M - np.diag(T["wn"])).max() < 1e-15
