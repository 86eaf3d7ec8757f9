"""The generic tensor-product path (rhs_generic.cuh) through the C ABI against the CPU oracle: 2-D quadrilaterals
(the reference's mms.euler_2d configuration: p = 2, Gauss-Lobatto nodes and rules) and Gauss-Lobatto /
mixed node-rule choices in 3-D, plus cross-checks against the specialised 3-D kernels."""
import numpy as np
import pytest

import oracle_api
import tps_b200
from common import rel_l2, tgv_state, warp_mesh

pytestmark = pytest.mark.gpu
PI = np.pi


def _warp2d(xyz, amp, lo, hi):
    L = np.asarray(hi) - np.asarray(lo)
    t = 2 * PI * (xyz - np.asarray(lo)) / L
    d = np.empty_like(xyz)
    d[..., 0] = np.sin(t[..., 0]) * np.cos(t[..., 1])
    d[..., 1] = np.cos(2 * t[..., 0]) * np.sin(t[..., 1])
    return np.ascontiguousarray(xyz + amp * L / (2 * PI) * d)


def _state2d(xy, seed=20261018):
    x, y = xy[:, 0], xy[:, 1]
    rho = 1.2 + 0.1 * np.sin(x) * np.cos(y)
    u = 30 * np.sin(x) * np.cos(y) + 10
    v = -30 * np.cos(x) * np.sin(y) + 4
    p = 101300 + 500 * (np.cos(2 * x) + np.cos(2 * y))
    U = np.concatenate([rho, rho * u, rho * v, p / 0.4 + 0.5 * rho * (u * u + v * v)])
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(U * (1 + 0.01 * rng.uniform(-1, 1, U.shape)))


@pytest.mark.parametrize("warp", [0.0, 0.12])
@pytest.mark.parametrize("order,bt,ir,eq", [(2, 1, 1, 0), (2, 1, 1, 1), (2, 0, 0, 1), (3, 0, 0, 1), (1, 1, 1, 1),
                                            (3, 1, 1, 1), (2, 0, 1, 1), (2, 1, 0, 1)])
def test_quadrilateral_parity(lib_built, oracle_built, order, bt, ir, eq, warp):
    import torch
    lo, hi = (-PI, -PI), (PI, PI)
    m = tps_b200.cartesian_quad_mesh(7, 6, lo=lo, hi=hi)
    if warp:
        m["elem_xyz"] = _warp2d(m["elem_xyz"], warp, lo, hi)
    op = tps_b200.RhsOperator(m, order=order, physics=tps_b200.Physics.dry_air(eq, 3e4, 0.2), basis_type=bt,
                              int_rule_type=ir)
    orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(eq, 3e4, 0.2), basis_type=bt, int_rule=ir)
    U = _state2d(orc.node_coords())
    x = torch.from_numpy(U).cuda()
    y = op.Mult(x).cpu().numpy()
    yo, go = orc.mult(U, want_grad=True)
    up, g = op.fields()
    N = orc.N
    assert op.N == N and op.neq == 4
    assert rel_l2(up.cpu().numpy(), orc.primitives(U)) < 1e-14
    assert rel_l2(g.cpu().numpy(), go) < 1e-11
    for k in range(4):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k
    assert abs(op.max_char_speed() / orc.max_char_speed - 1) < 1e-13


def test_mms_euler_2d_configuration(lib_built, oracle_built):
    """BASELINE config C1 (mms.euler_2d: p = 2, basisType = 1, integrationRule = 1, periodic quads from
    utils/beam_mesh.cpp -nx 1 -nt 5 -a 3.02 -b 3.02 refined): parity on a 40 x 40 restatement, and 10 RK4 steps."""
    import torch
    m = tps_b200.cartesian_quad_mesh(40, 40, lo=(0, 0), hi=(3.02, 3.02))
    op = tps_b200.RhsOperator(m, order=2, physics=tps_b200.Physics.dry_air(0), basis_type=1, int_rule_type=1)
    orc = oracle_api.Oracle(2, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(0), basis_type=1, int_rule=1)
    U = _state2d(orc.node_coords() * (2 * PI / 3.02))
    N = orc.N
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo = orc.mult(U)
    for k in range(4):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k
    dt = 1e-5
    x = torch.from_numpy(U.copy()).cuda()
    op.ode_step(x, dt, scheme=4, nsteps=10)
    ref = orc.rk4(U, dt, 10)
    assert rel_l2(x.cpu().numpy(), ref) < 1e-12


@pytest.mark.parametrize("order,bt,ir", [(2, 1, 1), (3, 1, 1), (2, 0, 1)])
def test_hexahedral_gauss_lobatto_parity(lib_built, oracle_built, order, bt, ir):
    import torch
    m = warp_mesh(tps_b200.cartesian_hex_mesh(4, 3, 3, lo=(-PI,) * 3, hi=(PI,) * 3), amp=0.1)
    op = tps_b200.RhsOperator(m, order=order, physics=tps_b200.Physics.dry_air(1, 3e4, 0.2), basis_type=bt, int_rule_type=ir)
    orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(1, 3e4, 0.2), basis_type=bt, int_rule=ir)
    U = tgv_state(orc.node_coords())
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo = orc.mult(U)
    N = orc.N
    for k in range(5):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k


def test_generic_path_agrees_with_specialised_kernels(lib_built, oracle_built, monkeypatch):
    """Three independent implementations of the same operator (fast, general trilinear, generic) on one mesh."""
    import torch
    m = tps_b200.cartesian_hex_mesh(4, 5, 3, lo=(-PI,) * 3, hi=(PI,) * 3)
    phys = tps_b200.Physics.dry_air(1, 2e4, 0.1)
    op = tps_b200.RhsOperator(m, order=3, physics=phys)
    xyz = oracle_api.Oracle(1, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"]).node_coords()
    from common import node_coords_from_mesh
    U = tgv_state(node_coords_from_mesh(m["elem_xyz"], 3))
    y_fast = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    monkeypatch.setenv("TPSB_PATH", "generic")
    op2 = tps_b200.RhsOperator(m, order=3, physics=phys)
    y_gen = op2.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    assert rel_l2(y_gen, y_fast) < 1e-12
