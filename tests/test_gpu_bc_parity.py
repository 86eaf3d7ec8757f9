"""Boundary conditions through the C ABI against the CPU oracle: inlet / outlet / inviscid, adiabatic and
isothermal walls (BCintegrator, WallBC, InletBC, OutletBC; BASELINE config C2 'cyl3d' uses inlet + outlet +
isothermal wall), on parallelepiped (fast path) and warped trilinear (general path) boxes."""
import numpy as np
import pytest

import oracle_api
import tps_b200
from common import box_face_attrs, rel_l2, tgv_state, warp_mesh

pytestmark = pytest.mark.gpu
LO, HI = (0.0, 0.0, 0.0), (2.0, 1.2, 1.0)
# attr: 1 x- inlet, 2 x+ outlet, 3 y- isothermal wall, 4 y+ adiabatic wall, 5 z- inviscid wall, 6 z+ isothermal wall
BC_SPECS = [(1, 0, 2, (1.2, 25.0, 1.0, -2.0)), (2, 1, 0, (101300.0,)), (3, 2, 3, (310.0,)), (4, 2, 2, ()),
            (5, 2, 0, ()), (6, 2, 3, (290.0,))]


def _channel(order, eq, warp, use_bc_in_grad, visc_mult=4e3, n=(4, 3, 3)):
    import torch
    m = tps_b200.cartesian_hex_mesh(*n, lo=LO, hi=HI, periodic=(0, 0, 0))
    attr = box_face_attrs(m, LO, HI)
    if warp:
        m = warp_mesh(m, amp=0.1, lo=LO, hi=HI)
    phys = tps_b200.Physics.dry_air(eq, visc_mult, 0.2)
    op = tps_b200.RhsOperator(m, order=order, physics=phys, face_attr=attr, use_bc_in_grad=use_bc_in_grad,
                              bcs=[tps_b200.BcDesc.make(*b) for b in BC_SPECS])
    orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(eq, visc_mult, 0.2))
    orc.set_bcs(attr, [oracle_api.make_bc(*b) for b in BC_SPECS], use_bc_in_grad)
    xyz = orc.node_coords()
    U = tgv_state(xyz * np.pi)  # smooth, non-trivial state; +-1 % seeded perturbation
    return torch, op, orc, U


@pytest.mark.parametrize("warp", [False, True])
@pytest.mark.parametrize("order,eq,use", [(3, 1, False), (3, 1, True), (2, 1, True), (1, 1, False), (3, 0, False)])
def test_bc_rhs_parity(lib_built, oracle_built, order, eq, use, warp):
    torch, op, orc, U = _channel(order, eq, warp, use)
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo, go = orc.mult(U, want_grad=True)
    g = op.fields()[1].cpu().numpy()
    assert rel_l2(g, go) < 1e-11
    N = orc.N
    for k in range(5):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k


def test_bc_rhs_parity_against_reference_object_code(lib_built, oracle_built):
    import os
    if not os.path.exists(os.path.join(oracle_api.ORACLE_DIR, "_ref", "liboracle_ref.so")):
        pytest.skip("oracle/_ref not built")
    import torch
    m = tps_b200.cartesian_hex_mesh(3, 3, 3, lo=LO, hi=HI, periodic=(0, 0, 0))
    attr = box_face_attrs(m, LO, HI)
    m = warp_mesh(m, amp=0.1, lo=LO, hi=HI)
    op = tps_b200.RhsOperator(m, order=3, physics=tps_b200.Physics.dry_air(1, 4e3, 0.2), face_attr=attr,
                              use_bc_in_grad=True, bcs=[tps_b200.BcDesc.make(*b) for b in BC_SPECS])
    orc = oracle_api.Oracle(3, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(1, 4e3, 0.2), kind="ref")
    orc.set_bcs(attr, [oracle_api.make_bc(*b) for b in BC_SPECS], True)
    U = tgv_state(orc.node_coords() * np.pi)
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    assert rel_l2(y, orc.mult(U)) < 1e-10


def test_uniform_flow_is_preserved(lib_built):
    """Inlet / outlet / inviscid walls around a uniform axial flow on a warped mesh: dU/dt = 0."""
    import torch
    m = tps_b200.cartesian_hex_mesh(6, 4, 4, lo=LO, hi=HI, periodic=(0, 0, 0))
    attr = box_face_attrs(m, LO, HI)
    m = warp_mesh(m, amp=0.1, lo=LO, hi=HI)
    rho, u, p = 1.2, 30.0, 101300.0
    bcs = [tps_b200.BcDesc.make(1, 0, 2, (rho, u, 0, 0)), tps_b200.BcDesc.make(2, 1, 0, (p,))]
    bcs += [tps_b200.BcDesc.make(a, 2, 0) for a in (3, 4, 5, 6)]
    op = tps_b200.RhsOperator(m, order=3, physics=tps_b200.Physics.dry_air(1, 1.0), face_attr=attr, bcs=bcs)
    N = op.N
    U = torch.empty(5 * N, dtype=torch.float64, device="cuda")
    for k, v in enumerate((rho, rho * u, 0.0, 0.0, p / 0.4 + 0.5 * rho * u * u)):
        U[k * N:(k + 1) * N] = v
    y = op.Mult(U)
    assert float(y[:N].abs().max()) < 1e-9 * rho * u and float(y[4 * N:].abs().max()) < 1e-9 * (p / 0.4) * u


def test_unsupported_bc_and_missing_bc_are_reported(lib_built):
    m = tps_b200.cartesian_hex_mesh(3, 3, 3, lo=LO, hi=HI, periodic=(0, 0, 0))
    attr = box_face_attrs(m, LO, HI)
    with pytest.raises(tps_b200.TpsbError, match="no boundary condition"):
        tps_b200.RhsOperator(m, order=2, face_attr=attr, bcs=[tps_b200.BcDesc.make(1, 2, 0)])
    with pytest.raises(tps_b200.TpsbError, match="not built"):
        tps_b200.RhsOperator(m, order=2, face_attr=attr, bcs=[tps_b200.BcDesc.make(a, 1, 1) for a in range(1, 7)])  # RESIST_IN


def test_cylinder_ogrid_parity(lib_built, oracle_built):
    """BASELINE config C2 restated on hexahedra: O-grid around a unit-diameter cylinder, p = 3,
    inlet rho = 1.2, u = (20, 0, 0), outlet p = 101300, isothermal wall 300 K with useBCinGrad
    (test/inputs/input.4iters.cyl.ini), 2560 trilinear elements."""
    import torch
    m = tps_b200.cylinder_ogrid_mesh(10, 32, 8)
    specs = [(1, 2, 3, (300.0,)), (2, 0, 2, (1.2, 20.0, 0.0, 0.0)), (3, 1, 0, (101300.0,))]
    phys = tps_b200.Physics.dry_air(1, 50.0, 0.0)
    op = tps_b200.RhsOperator(m, order=3, physics=phys, face_attr=m["face_attr"], use_bc_in_grad=True,
                              bcs=[tps_b200.BcDesc.make(*b) for b in specs])
    orc = oracle_api.Oracle(3, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(1, 50.0, 0.0))
    orc.set_bcs(m["face_attr"], [oracle_api.make_bc(*b) for b in specs], True)
    N = orc.N
    xyz = orc.node_coords()
    rng = np.random.default_rng(20261018)
    rho, p = 1.2, 102300.0
    r = np.hypot(xyz[:, 0], xyz[:, 1])
    u = 20.0 * (1.0 - (0.5 / r) ** 2)  # potential-flow-like, vanishing at the wall
    U = np.concatenate([np.full(N, rho), rho * u, np.zeros(N), np.zeros(N), p / 0.4 + 0.5 * rho * u * u])
    U *= 1.0 + 0.01 * rng.uniform(-1, 1, size=U.shape)
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo = orc.mult(U)
    for k in range(5):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k
    assert abs(op.max_char_speed() / orc.max_char_speed - 1) < 1e-13


@pytest.mark.parametrize("warp", [False, True])
def test_slip_wall_3d(lib_built, oracle_built, warp):
    """WallType SLIP (wallBC.cpp:326-428) on the z and y sides of the channel, fast and general 3-D paths."""
    import torch
    specs = [(1, 0, 2, (1.2, 25.0, 1.0, -2.0)), (2, 1, 0, (101300.0,)), (3, 2, 1, ()), (4, 2, 1, ()), (5, 2, 1, ()),
             (6, 2, 3, (290.0,))]
    m = tps_b200.cartesian_hex_mesh(4, 3, 3, lo=LO, hi=HI, periodic=(0, 0, 0))
    attr = box_face_attrs(m, LO, HI)
    if warp:
        m = warp_mesh(m, amp=0.1, lo=LO, hi=HI)
    op = tps_b200.RhsOperator(m, order=3, physics=tps_b200.Physics.dry_air(1, 4e3, 0.2), face_attr=attr,
                              bcs=[tps_b200.BcDesc.make(*b) for b in specs])
    orc = oracle_api.Oracle(3, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(1, 4e3, 0.2))
    orc.set_bcs(attr, [oracle_api.make_bc(*b) for b in specs], False)
    U = tgv_state(orc.node_coords() * np.pi)
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo = orc.mult(U)
    N = orc.N
    for k in range(5):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k


@pytest.mark.parametrize("warp,chunks", [(False, 4), (True, 6)])
def test_chunked_host_pipeline_with_boundary_conditions(lib_built, monkeypatch, warp, chunks):
    """tpsb_rhs_mult_host on a channel with every boundary type: the chunked copy / compute pipeline (boundary faces in
    the face op of their element's chunk; fast path on the parallelepiped mesh, general path on the warped one) is
    bit-identical to the device entry point."""
    import torch
    monkeypatch.setenv("TPSB_HOST_CHUNKS", str(chunks))
    m = tps_b200.cartesian_hex_mesh(4, 3, 5, lo=LO, hi=HI, periodic=(0, 0, 0))
    attr = box_face_attrs(m, LO, HI)
    if warp:
        m = warp_mesh(m, amp=0.1, lo=LO, hi=HI)
    op = tps_b200.RhsOperator(m, order=3, physics=tps_b200.Physics.dry_air(1, 4e3, 0.2), face_attr=attr, use_bc_in_grad=True,
                              bcs=[tps_b200.BcDesc.make(*b) for b in BC_SPECS])
    from common import node_coords_from_mesh
    U = tgv_state(node_coords_from_mesh(m["elem_xyz"], 3) * np.pi)
    y_dev = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    hx = torch.from_numpy(U).pin_memory()
    hy = torch.zeros_like(hx).pin_memory()
    op.mult_host(hx, hy)
    assert np.array_equal(hy.numpy(), y_dev)
    assert np.isfinite(y_dev).all() and np.abs(y_dev).max() > 0


def test_general_wall_in_3d_runs_on_the_generic_path(lib_built, oracle_built):
    """WallType VISC_GNRL (independent heavy-species / electron thermal conditions, wallBC.cpp:512-543) is built on the
    generic path: a 3-D dry-air Gauss-Legendre run that asks for it is routed there instead of being refused."""
    import torch
    m = tps_b200.cartesian_hex_mesh(3, 3, 3, lo=LO, hi=HI, periodic=(0, 0, 0))
    attr = box_face_attrs(m, LO, HI)
    m = warp_mesh(m, amp=0.1, lo=LO, hi=HI)
    specs = [(1, 0, 2, (1.2, 25.0, 1.0, -2.0)), (2, 1, 0, (101300.0,)), (3, 2, 4, (1.0, 0.0, 310.0, 0.0)), (4, 2, 4, (0.0, 0.0, 0.0, 0.0)),
             (5, 2, 0, ()), (6, 2, 3, (290.0,))]
    op = tps_b200.RhsOperator(m, order=2, physics=tps_b200.Physics.dry_air(1, 4e3, 0.2), face_attr=attr,
                              bcs=[tps_b200.BcDesc.make(*b) for b in specs])
    orc = oracle_api.Oracle(2, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(1, 4e3, 0.2))
    orc.set_bcs(attr, [oracle_api.make_bc(*b) for b in specs], False)
    U = tgv_state(orc.node_coords() * np.pi)
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo = orc.mult(U)
    N = orc.N
    for k in range(5):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k
