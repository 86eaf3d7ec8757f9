"""Generic path through the C ABI against the CPU oracle: boundary conditions on quadrilateral meshes (BCintegrator /
WallBC / InletBC / OutletBC for dry air and mixtures), axisymmetric runs (nvel = 3 on a 2-D mesh: r-weighted operators,
Me_inv_rad, axisymmetric viscous terms, AxisymmetricSource) and the six-species two-temperature argon mixture of
BASELINE config C4 (test/inputs/plasma.ini physics on a restated quad mesh)."""
import os

import numpy as np
import pytest

import axisym_cases as ac
import oracle_api
from common import rel_l2

pytestmark = pytest.mark.gpu
HAVE_REF = os.path.exists(os.path.join(oracle_api.ORACLE_DIR, "_ref", "liboracle_ref.so"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref (reference object code) not built")


def _compare(op, orc, U, tol_y=1e-10):
    import torch
    N, neq = orc.N, op.neq
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo, go = orc.mult(U, want_grad=True)
    up, g = op.fields()
    assert rel_l2(up.cpu().numpy(), orc.primitives(U)) < 1e-13
    assert rel_l2(g.cpu().numpy(), go) < 1e-11
    for k in range(neq):
        ref = yo[k * N:(k + 1) * N]
        assert np.linalg.norm(y[k * N:(k + 1) * N] - ref) <= tol_y * max(np.linalg.norm(ref), 1e-30) + 1e-9, k
    assert abs(op.max_char_speed() / orc.max_char_speed - 1) < 1e-13


@pytest.mark.parametrize("order,bt,ir", [(2, 1, 1), (3, 0, 0), (1, 0, 0)])
@pytest.mark.parametrize("eq,bc,ubg", [(1, "c4", True), (1, "c4", False), (1, "adiabatic", False), (0, "inviscid", False),
                                       (1, "slip", False)])
def test_quad_boundary_conditions_dry_air(lib_built, oracle_built, order, bt, ir, eq, bc, ubg):
    m = ac.box(warp=0.06)
    op, orc = ac.make_pair(m, order, eq, bt, ir, 2, bc, ubg)
    _compare(op, orc, ac.dry_state(orc.node_coords(), 2))


@pytest.mark.parametrize("order,bt,ir", [(3, 0, 0), (2, 1, 1)])
@pytest.mark.parametrize("eq,bc,ubg", [(1, "c4", True), (1, "adiabatic", False), (0, "inviscid", False), (1, None, False)])
def test_axisymmetric_dry_air(lib_built, oracle_built, order, bt, ir, eq, bc, ubg):
    m = ac.box(warp=0.06)
    op, orc = ac.make_pair(m, order, eq, bt, ir, 3, bc, ubg)
    assert op.neq == 5
    _compare(op, orc, ac.dry_state(orc.node_coords(), 3))


def test_axisymmetric_vessel_at_rest(lib_built, oracle_built):
    m = ac.box(n=(4, 3), warp=0.07)
    op, orc = ac.make_pair(m, 3, 1, 0, 0, 3, "inviscid", False)
    import torch
    N = orc.N
    U = np.concatenate([np.full(N, 1.2), np.zeros(3 * N), np.full(N, 101300.0 / 0.4)])
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    assert np.abs(y).max() < 1e-10 * 101300.0 / 0.5


@needs_ref
@pytest.mark.parametrize("nvel,bc,ubg", [(3, "c4", True), (3, "adiabatic", False), (2, "c4", True), (2, "c4", False)])
def test_plasma_axisym_configuration_c4(lib_built, oracle_built, nvel, bc, ubg):
    """Six-species non-ambipolar two-temperature argon (neq = nvel + 8: 11 when axisymmetric), p = 3 GL/GL, inviscid +
    isothermal walls, subsonic inlet with species, pressure outlet, useBCinGrad."""
    m = ac.box(n=(4, 3), warp=0.05)
    op, orc = ac.make_pair(m, 3, 1, 0, 0, nvel, bc, ubg, mixture=ac.argon6_dict())
    assert op.neq == nvel + 8
    up = ac.argon6_primitives(orc.node_coords(), nvel)
    U = np.ascontiguousarray(orc.pt("cons", up).T).reshape(-1)
    _compare(op, orc, U)


@needs_ref
def test_plasma_axisym_c4_with_argon_mixture_transport(lib_built, oracle_built):
    """Config C4 with the collision-integral transport variant (transport_model = argon_mixture) instead of constants."""
    m = ac.box(n=(4, 3), warp=0.05)
    d = ac.argon6_dict()
    d.update(transport_model="argon_mixture", third_order_k_electron=True)
    op, orc = ac.make_pair(m, 3, 1, 0, 0, 3, "c4", True, mixture=d)
    up = ac.argon6_primitives(orc.node_coords(), 3)
    up[:, 4] *= 8.0   # T_h ~ 7000 K
    up[:, 10] *= 3.0  # T_e ~ 9000 K
    up[:, 9] = up[:, 5]  # quasi-neutral
    U = np.ascontiguousarray(orc.pt("cons", up).T).reshape(-1)
    _compare(op, orc, U)


@needs_ref
def test_ternary_mixture_with_walls(lib_built, oracle_built):
    import plasma_cases
    import tps_b200
    m = ac.box(n=(4, 4), lo=(-1.0, -1.0), hi=(1.0, 1.0))
    pm = plasma_cases.ternary_models()
    specs = [(1, 2, 3, (450.0,)), (2, 2, 2, ()), (3, 0, 2, (1.2, 10.0, 2.0, 0.0, 0.3 * (plasma_cases.MW_AR - plasma_cases.MW_E))),
             (4, 1, 0, (120000.0,))]
    for ubg in (True, False):
        op = tps_b200.RhsOperator(m, order=2, physics=tps_b200.Physics.plasma_mixture(pm, 1), basis_type=1, int_rule_type=1,
                                  face_attr=m["face_attr"], use_bc_in_grad=ubg, bcs=[tps_b200.BcDesc.make(*b) for b in specs])
        orc = oracle_api.Oracle(2, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                                phys=oracle_api.mixture_params(pm, 1), kind="ref", basis_type=1, int_rule=1, neq=6, nvel=2)
        orc.set_bcs(m["face_attr"], [oracle_api.make_bc(*b) for b in specs], ubg)
        up = plasma_cases.smooth_primitives(orc.node_coords() * np.pi)
        U = np.ascontiguousarray(orc.pt("cons", up).T).reshape(-1)
        _compare(op, orc, U)


@pytest.mark.parametrize("bc", [None, "inviscid", "c4"])
def test_roe_riemann_solver_2d(lib_built, oracle_built, bc):
    """flow/useRoe = 1 (RiemannSolverTPS::Eval_Roe, 2-D dry air) on interior faces and inviscid walls."""
    import tps_b200
    m = ac.box(n=(5, 4), warp=0.05) if bc else tps_b200.cartesian_quad_mesh(5, 4, lo=(-np.pi, -np.pi), hi=(np.pi, np.pi))
    specs = ac.bcs(bc) if bc else []
    orc = oracle_api.Oracle(2, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(1, 3e4, 0.2, use_roe=True), basis_type=1, int_rule=1)
    if specs:
        orc.set_bcs(m["face_attr"], [oracle_api.make_bc(*b) for b in specs], True)
    op = tps_b200.RhsOperator(m, order=2, physics=tps_b200.Physics.dry_air(1, 3e4, 0.2, use_roe=True), basis_type=1,
                              int_rule_type=1, face_attr=m.get("face_attr") if specs else None, use_bc_in_grad=True,
                              bcs=[tps_b200.BcDesc.make(*b) for b in specs] if specs else None)
    _compare(op, orc, ac.dry_state(orc.node_coords(), 2))
    with pytest.raises(tps_b200.TpsbError):  # the reference's Roe solver has no 3-D form
        tps_b200.RhsOperator(tps_b200.cartesian_hex_mesh(2, 2, 2), order=1, physics=tps_b200.Physics.dry_air(1, 1.0, 0.0, use_roe=True))


@needs_ref
@pytest.mark.parametrize("hvy,elec,two_t", [(1, 1, True), (0, 0, True), (1, 2, True), (0, 2, True), (1, 0, False), (0, 2, False),
                                            (1, 2, False)])
def test_general_wall_and_sheath(lib_built, oracle_built, hvy, elec, two_t):
    """WallType VISC_GNRL (wallBC.cpp:112-147, 512-543): no-slip wall with independent heavy-species and electron
    thermal conditions (ADIAB / ISOTH / SHTH), PerfectMixture::computeSheathBdrFlux for the sheath.  (An isothermal
    electron condition on a single-temperature mixture is not a valid input of the reference either: it writes T_e
    into primitive num_equation - 1, the last species, wallBC.cpp:133-134.)"""
    import plasma_cases
    import tps_b200
    m = ac.box(n=(4, 4), lo=(-1.0, -1.0), hi=(1.0, 1.0), warp=0.04)
    pm = plasma_cases.ternary_models(two_temperature=two_t)
    wall = (float(hvy), float(elec), 420.0, 1800.0)
    specs = [(1, 2, 4, wall), (2, 2, 4, wall), (3, 2, 4, wall), (4, 1, 0, (120000.0,))]
    neq = 6 if two_t else 5
    op = tps_b200.RhsOperator(m, order=2, physics=tps_b200.Physics.plasma_mixture(pm, 1), basis_type=1, int_rule_type=1,
                              face_attr=m["face_attr"], use_bc_in_grad=True, bcs=[tps_b200.BcDesc.make(*b) for b in specs])
    orc = oracle_api.Oracle(2, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.mixture_params(pm, 1), kind="ref", basis_type=1, int_rule=1, neq=neq, nvel=2)
    orc.set_bcs(m["face_attr"], [oracle_api.make_bc(*b) for b in specs], True)
    up = plasma_cases.smooth_primitives(orc.node_coords() * np.pi)[:, :neq]
    U = np.ascontiguousarray(orc.pt("cons", up).T).reshape(-1)
    _compare(op, orc, U)


def test_general_wall_dry_air(lib_built, oracle_built):
    import tps_b200
    m = ac.box(warp=0.05)
    for wall in ((1.0, 0.0, 330.0, 0.0), (0.0, 0.0, 0.0, 0.0), (1.0, 1.0, 330.0, 360.0)):
        specs = [(1, 2, 4, wall), (2, 2, 4, wall), (3, 0, 2, (1.25, 12.0, 3.0, 1.5)), (4, 1, 0, (101000.0,))]
        op = tps_b200.RhsOperator(m, order=3, physics=tps_b200.Physics.dry_air(1, 3e4, 0.2), face_attr=m["face_attr"],
                                  bcs=[tps_b200.BcDesc.make(*b) for b in specs], basis_type=1, int_rule_type=1)
        orc = oracle_api.Oracle(3, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                                phys=oracle_api.dry_air_params(1, 3e4, 0.2), basis_type=1, int_rule=1)
        orc.set_bcs(m["face_attr"], [oracle_api.make_bc(*b) for b in specs], False)
        _compare(op, orc, ac.dry_state(orc.node_coords(), 2))
