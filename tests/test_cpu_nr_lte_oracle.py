"""CPU checks of the oracle's two round-2 additions.
Non-reflecting inlets / outlets (restated from inletBC.cpp / outletBC.cpp over the per-point physics): the restatement over
the port physics equals the one over the reference's object code; a uniform state that already satisfies the target is a
fixed point (no boundary-state drift, boundary flux == interior flux); the patch mean is the plain mean over the boundary
quadrature points; the state is advanced by exactly dt times the characteristic derivative.
LTE fluid (the reference's own LteMixture / LteTransport object code): with tables that describe a calorically perfect gas
(e = R T / (gamma - 1), constant R, c = sqrt(gamma R T)) it must reproduce the dry-air operator - a cross-check of the table
logic (Newton inversion, p = rho R T, tabulated sound speed, boundary-state construction) against the pinned dry-air path."""
import os

import numpy as np
import pytest

import axisym_cases as ac
import oracle_api
import tps_b200
from common import rel_l2

HAVE_REF = os.path.exists(os.path.join(oracle_api.ORACLE_DIR, "_ref", "liboracle_ref.so")) or os.path.isdir("/root/reference/src")
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref (reference object code) not built")
REF_LEN = 0.7


def _nr(data, tangent=(1.0, 0.0, 0.0)):
    d = list(data) + [0.0] * (8 - len(data))
    return tuple(d + [REF_LEN] + list(tangent))


def _oracle(m, specs, kind="port", eq=1, order=2):
    orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(eq, 3e4, 0.2), kind=kind, basis_type=1, int_rule=1, neq=4, nvel=2)
    orc.set_bcs(m["face_attr"], [oracle_api.make_bc(*b) for b in specs], True)
    return orc


SPECS = [(1, 2, 0, ()), (2, 2, 0, ()), (3, 0, 6, _nr((1.25, 3.0, 12.0, 0.0))), (4, 1, 2, _nr((101000.0,)))]


@needs_ref
def test_port_and_reference_back_ends_agree(oracle_built):
    m = ac.box(warp=0.05)
    a, b = _oracle(m, SPECS, "port"), _oracle(m, SPECS, "ref")
    U = ac.dry_state(a.node_coords(), 2)
    for o in (a, b):
        o.set_bc_time_step(2e-6)
    for it in range(3):
        ya, yb = a.mult(U), b.mult(U)
        assert rel_l2(ya, yb) < 1e-12, it
        for attr in (3, 4):
            (ma, ba), (mb, bb) = a.bc_state(attr), b.bc_state(attr)
            assert rel_l2(ma, mb) < 1e-14 and rel_l2(ba, bb) < 1e-13
        U = U + 1e-6 * ya


def test_uniform_state_at_the_target_is_a_fixed_point(oracle_built):
    """density / velocity of the inlet and pressure of the outlet equal to the uniform interior state: every characteristic
    amplitude vanishes, the boundary states do not move and the boundary flux is the interior flux (dU/dt = 0)"""
    m = ac.box(warp=0.0)
    rho, u, v, p = 1.2, 0.0, 30.0, 101300.0
    specs = [(1, 2, 0, ()), (2, 2, 0, ()), (3, 0, 6, _nr((rho, u, v, 0.0))), (4, 1, 2, _nr((p,)))]
    orc = _oracle(m, specs, eq=0)
    N = orc.N
    U = np.concatenate([np.full(N, rho), np.full(N, rho * u), np.full(N, rho * v), np.full(N, p / 0.4 + 0.5 * rho * (u * u + v * v))])
    orc.set_bc_time_step(1e-4)
    for _ in range(3):
        y = orc.mult(U)
        assert np.abs(y[:N]).max() < 1e-9 * rho * v and np.abs(y[3 * N:]).max() < 1e-9 * (p / 0.4) * v
    for attr in (3, 4):
        mean, bu = orc.bc_state(attr)
        assert np.allclose(mean, [rho, u, v, p / (rho * 287.058)], rtol=1e-13, atol=1e-12)
        assert np.allclose(bu, [rho, rho * u, rho * v, p / 0.4 + 0.5 * rho * v * v], rtol=1e-12, atol=1e-9)


def test_patch_mean_and_first_evaluation(oracle_built):
    """meanUp = plain mean of the primitives interpolated to the patch's face quadrature points; the first evaluation
    initialises the boundary states with the conserved form of those primitives; dt = 0 leaves them where they are"""
    m = ac.box(warp=0.04)
    orc = _oracle(m, SPECS)
    U = ac.dry_state(orc.node_coords(), 2)
    orc.set_bc_time_step(0.0)
    orc.mult(U)
    mean, bu = orc.bc_state(4)
    prim = np.array([orc.pt("prim", row[None, :])[0] for row in bu])     # back to primitives, point by point
    assert np.allclose(prim.mean(axis=0), mean, rtol=1e-12)
    bu0 = bu.copy()
    orc.mult(U)
    assert rel_l2(orc.bc_state(4)[1], bu0) < 1e-14
    # ... and a time step moves every boundary state linearly in dt
    orc.set_bc_time_step(1e-6)
    orc.mult(U)
    d1 = orc.bc_state(4)[1] - bu0
    orc2 = _oracle(m, SPECS)
    orc2.set_bc_time_step(0.0)
    orc2.mult(U)
    orc2.set_bc_time_step(2e-6)
    orc2.mult(U)
    d2 = orc2.bc_state(4)[1] - bu0
    assert rel_l2(d2, 2 * d1) < 1e-9 and np.abs(d1).max() > 0


def ideal_gas_tables(gamma=1.4, R=287.058, Tlo=200.0, Thi=450.0, n=1000):
    T = np.linspace(Tlo, Thi, n)
    Tt = np.linspace(Tlo, Thi, 64)
    return tps_b200.LteTables.make(T, R / (gamma - 1.0) * T, np.full(n, R), np.sqrt(gamma * R * T), Tt, 1.8e-5 + 0 * Tt,
                                   0.026 + 0 * Tt, 1.0 + 0 * Tt)


@needs_ref
@pytest.mark.parametrize("bc", [None, "c4"])
def test_lte_with_ideal_gas_tables_is_dry_air(oracle_built, bc):
    """Euler: the LTE operator over ideal-gas tables equals the dry-air operator (the tabulated sqrt of the sound speed is
    the only interpolation error: 1000 points over 250 K: relative 2e-8 in the wave speed)"""
    m = ac.box(warp=0.05) if bc else tps_b200.cartesian_quad_mesh(5, 4, lo=(-1, -1), hi=(1, 1))
    t = ideal_gas_tables()
    common = dict(basis_type=1, int_rule=1, neq=4, nvel=2)
    lte = oracle_api.Oracle(2, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.lte_params(t, 0), kind="ref", **common)
    dry = oracle_api.Oracle(2, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(0), kind="ref", **common)
    if bc:
        for o in (lte, dry):
            o.set_bcs(m["face_attr"], [oracle_api.make_bc(*b) for b in ac.bcs(bc, 2)], False)
    U = ac.dry_state(dry.node_coords() * 2.0, 2)
    assert rel_l2(lte.primitives(U), dry.primitives(U)) < 1e-13       # Newton on a linear e(T) lands exactly
    yl, yd = lte.mult(U), dry.mult(U)
    N = dry.N
    for k in range(4):
        assert rel_l2(yl[k * N:(k + 1) * N], yd[k * N:(k + 1) * N]) < 2e-8, k
    assert abs(lte.max_char_speed / dry.max_char_speed - 1) < 3e-8   # interpolation error of sqrt(gamma R T): (dT / T)^2 / 32
