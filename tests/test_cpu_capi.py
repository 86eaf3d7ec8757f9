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


def test_fused_swizzle_is_conflict_free():
    """The shared-memory node swizzle of elem_fused_kernel (rhs_fused.cuh: fused_swz) sends every quarter-warp access
    pattern of the kernel -- node order, DMMA B fragments and D-fragment stores along x / y / z, line tasks along
    x / y / z -- to eight distinct 16-byte bank groups, and is a permutation of the 64 nodes."""
    src = open(os.path.join(ROOT, "tps_b200", "csrc", "rhs_fused.cuh")).read()
    assert "return (n & 0x38) | ((n & 7) ^ ((n >> 2) & 2) ^ (((n >> 4) & 1) * 5) ^ ((n >> 4) & 2));" in src

    def swz(n):
        return (n & 0x38) | ((n & 7) ^ ((n >> 2) & 2) ^ (((n >> 4) & 1) * 5) ^ ((n >> 4) & 2))

    assert sorted(swz(n) for n in range(64)) == list(range(64))
    stride = [1, 4, 16]

    def base(d, L):
        return 4 * L if d == 0 else ((L & 3) + 16 * (L >> 2) if d == 1 else L)

    pats = [[8 * q + t for t in range(8)] for q in range(8)]  # one thread per node
    for d in range(3):
        for h in range(2):
            for t in range(4):  # B fragment: lanes (fr, fk) = (line 8h + fr, position fk), a quarter warp = two lines
                pats.append([base(d, 8 * h + fr) + fk * stride[d] for fr in (2 * t, 2 * t + 1) for fk in range(4)])
            for t in range(2):  # D-fragment stores of rows 0-3: lanes (row, fk) -> lines 8h + 2fk (+1)
                for s in range(2):
                    pats.append([base(d, 8 * h + 2 * fk + s) + row * stride[d] for row in (2 * t, 2 * t + 1) for fk in range(4)])
        for m in range(4):  # line tasks: lane = line, all at position m
            for t in range(2):
                pats.append([base(d, L) + m * stride[d] for L in range(8 * t, 8 * t + 8)])
    for nodes in pats:
        assert len({swz(n) % 8 for n in nodes}) == 8, nodes
