"""Polynomial orders 4 and 5 (the reference takes any flow/order MFEM supports; its regression inputs use up to 3).  They run
on the generic path (dense reference-element tables; the largest rule, the 3-D Gauss-Lobatto face rule at p = 5, has 8 points
per direction) and are held to the same bar as orders 1-3: primitives 1e-14, gradients 1e-11, dU/dt 1e-10 per equation
against the CPU oracle, on affine and warped meshes, with boundary conditions, for both node / rule families."""
import numpy as np
import pytest

import oracle_api
import tps_b200
from common import rel_l2, tgv_state, warp_mesh
from test_gpu_generic_parity import _state2d, _warp2d

pytestmark = pytest.mark.gpu
PI = np.pi


@pytest.mark.parametrize("warp", [0.0, 0.1])
@pytest.mark.parametrize("order,bt,ir,eq", [(4, 0, 0, 1), (5, 0, 0, 1), (4, 1, 1, 1), (5, 1, 1, 0), (5, 1, 1, 1)])
def test_quadrilaterals_order_4_5(lib_built, oracle_built, order, bt, ir, eq, warp):
    import torch
    lo, hi = (-PI, -PI), (PI, PI)
    m = tps_b200.cartesian_quad_mesh(5, 4, lo=lo, hi=hi)
    if warp:
        m["elem_xyz"] = _warp2d(m["elem_xyz"], warp, lo, hi)
    op = tps_b200.RhsOperator(m, order=order, physics=tps_b200.Physics.dry_air(eq, 3e4, 0.2), basis_type=bt, int_rule_type=ir)
    assert op.path() == "generic"
    orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(eq, 3e4, 0.2), basis_type=bt, int_rule=ir)
    U = _state2d(orc.node_coords())
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo, go = orc.mult(U, want_grad=True)
    up, g = op.fields()
    N = orc.N
    assert op.N == N == 20 * (order + 1) ** 2
    assert rel_l2(up.cpu().numpy(), orc.primitives(U)) < 1e-14
    assert rel_l2(g.cpu().numpy(), go) < 1e-11
    for k in range(4):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k
    assert abs(op.max_char_speed() / orc.max_char_speed - 1) < 1e-13


@pytest.mark.parametrize("order,bt,ir,warp", [(4, 0, 0, 0.0), (4, 0, 0, 0.08), (5, 0, 0, 0.08), (4, 1, 1, 0.08), (5, 1, 1, 0.0)])
def test_hexahedra_order_4_5(lib_built, oracle_built, order, bt, ir, warp):
    import torch
    m = tps_b200.cartesian_hex_mesh(3, 3, 3, lo=(-PI,) * 3, hi=(PI,) * 3)
    if warp:
        m = warp_mesh(m, amp=warp)
    op = tps_b200.RhsOperator(m, order=order, physics=tps_b200.Physics.dry_air(1, 3e4, 0.2), basis_type=bt, int_rule_type=ir)
    assert op.path() == "generic"
    orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(1, 3e4, 0.2), basis_type=bt, int_rule=ir)
    U = tgv_state(orc.node_coords())
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo, go = orc.mult(U, want_grad=True)
    N = orc.N
    assert N == 27 * (order + 1) ** 3
    assert rel_l2(op.fields()[1].cpu().numpy(), go) < 1e-11
    for k in range(5):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k


def test_order_4_time_stepping(lib_built, oracle_built):
    """20 RK4 steps at p = 4 on a warped 3-D mesh stay within 1e-10 of the oracle's."""
    import torch
    m = warp_mesh(tps_b200.cartesian_hex_mesh(3, 3, 3, lo=(-PI,) * 3, hi=(PI,) * 3), amp=0.05)
    op = tps_b200.RhsOperator(m, order=4, physics=tps_b200.Physics.dry_air(1, 3e4, 0.2))
    orc = oracle_api.Oracle(4, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(1, 3e4, 0.2))
    U = tgv_state(orc.node_coords())
    x = torch.from_numpy(U.copy()).cuda()
    op.ode_step(x, 2e-5, scheme=4, nsteps=20)
    assert rel_l2(x.cpu().numpy(), orc.rk4(U, 2e-5, 20)) < 1e-10


def test_order_6_is_refused(lib_built):
    m = tps_b200.cartesian_quad_mesh(3, 3)
    with pytest.raises(RuntimeError, match="order must be 1..5"):
        tps_b200.RhsOperator(m, order=6, physics=tps_b200.Physics.dry_air(0), basis_type=1, int_rule_type=1)


@pytest.mark.parametrize("order,bt,ir,nvel,bc,ubg", [(4, 0, 0, 2, "c4", True), (5, 1, 1, 2, "adiabatic", False),
                                                    (4, 0, 0, 3, "c4", True), (4, 1, 1, 3, "inviscid", False)])
def test_boundary_conditions_and_axisymmetry_order_4_5(lib_built, oracle_built, order, bt, ir, nvel, bc, ubg):
    """Inlet / outlet / isothermal and adiabatic walls (useBCinGrad) on warped quadrilaterals, planar and axisymmetric."""
    import axisym_cases as ac
    from test_gpu_generic_bc_axisym import _compare
    m = ac.box(warp=0.06)
    op, orc = ac.make_pair(m, order, 0 if bc == "inviscid" else 1, bt, ir, nvel, bc, ubg)
    assert op.path() == "generic"
    _compare(op, orc, ac.dry_state(orc.node_coords(), nvel))


def test_six_species_argon_order_4(lib_built, oracle_built):
    """Config C4's physics (six-species two-temperature argon, axisymmetric, its boundary set) at p = 4."""
    import axisym_cases as ac
    from test_gpu_generic_bc_axisym import HAVE_REF, _compare
    if not HAVE_REF:
        pytest.skip("oracle/_ref (reference object code) not built")
    m = ac.box(n=(3, 3), warp=0.05)
    op, orc = ac.make_pair(m, 4, 1, 0, 0, 3, "c4", True, mixture=ac.argon6_dict())
    up = ac.argon6_primitives(orc.node_coords(), 3)
    U = np.ascontiguousarray(orc.pt("cons", up).T).reshape(-1)
    _compare(op, orc, U)
