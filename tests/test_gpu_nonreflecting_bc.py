"""Non-reflecting and mass-flow inlets / outlets (InletBC::subsonicNonReflectingDensityVelocity src/inletBC.cpp:576-727,
OutletBC::subsonicNonReflectingPressure / subsonicNonRefMassFlow / subsonicNonRefPWMassFlow src/outletBC.cpp:573-1027, with
updateMean src/inletBC.cpp:482-564 / src/outletBC.cpp:470-561) through the C ABI against the CPU oracle.  They are stateful:
every evaluation refreshes the patch mean of the primitives and advances a conserved boundary state per face quadrature point
with the current time step, so the tests compare SEQUENCES of evaluations - dU/dt, the patch mean and the boundary states
after each call - and whole RK4 steps."""
import numpy as np
import pytest

import axisym_cases as ac
import oracle_api
import tps_b200
from common import box_face_attrs, rel_l2, tgv_state, warp_mesh

pytestmark = pytest.mark.gpu
REF_LEN = 0.7


def _nr(data, tangent=(0.0, 0.0, 0.0)):
    d = list(data) + [0.0] * (8 - len(data))
    return tuple(d + [REF_LEN] + list(tangent))


def _quad_specs(inlet_type, outlet_type, tangent_in=(0, 0, 0), tangent_out=(0, 0, 0)):
    """attr 1 x = lo inviscid wall, 2 x = hi isothermal wall, 3 y = lo inlet, 4 y = hi outlet (axisym_cases.box)"""
    inlet = (1.25, 3.0, 12.0, 0.0)
    specs = [(1, 2, 0, ()), (2, 2, 3, (298.15,))]
    specs.append((3, 0, inlet_type, _nr(inlet, tangent_in) if inlet_type in (6, 7) else inlet))
    out_data = (101000.0,) if outlet_type in (0, 2) else (1.25 * 12.0 * 1.2,)  # pressure, or mass flow rho v A
    specs.append((4, 1, outlet_type, _nr(out_data, tangent_out) if outlet_type >= 2 else out_data))
    return specs


def _pair(m, order, eq, bt, ir, specs, dim):
    phys_o, phys_g = oracle_api.dry_air_params(eq, 3e4, 0.2), tps_b200.Physics.dry_air(eq, 3e4, 0.2)
    orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"], phys=phys_o,
                            basis_type=bt, int_rule=ir, neq=dim + 2, nvel=dim)
    orc.set_bcs(m["face_attr"], [oracle_api.make_bc(*b) for b in specs], True)
    op = tps_b200.RhsOperator(m, order=order, physics=phys_g, basis_type=bt, int_rule_type=ir, nvel=dim,
                              face_attr=m["face_attr"], use_bc_in_grad=True, bcs=[tps_b200.BcDesc.make(*b) for b in specs])
    assert op.path() == "generic"
    return op, orc


def _sequence(op, orc, U0, attrs, dt=2e-6, calls=3):
    """calls evaluations on a drifting state: dU/dt, patch means and boundary states must agree after every call"""
    import torch
    op.set_time_step(dt)
    orc.set_bc_time_step(dt)
    N, neq = orc.N, op.neq
    U = U0.copy()
    for it in range(calls):
        y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
        yo = orc.mult(U)
        for k in range(neq):
            ref = yo[k * N:(k + 1) * N]
            assert np.linalg.norm(y[k * N:(k + 1) * N] - ref) <= 1e-10 * np.linalg.norm(ref), (it, k)
        for a in attrs:
            mg, bg = op.bc_state(a)
            mo, bo = orc.bc_state(a)
            assert bg.shape == bo.shape and bg.shape[0] > 0
            assert np.abs(mg - mo).max() <= 1e-13 * np.abs(mo).max(), (it, a)
            assert rel_l2(bg, bo) < 1e-13, (it, a)
        U = U + 0.3 * dt * yo  # next call sees another state (and the boundary states have moved)
    return U


@pytest.mark.parametrize("order,bt,ir,eq", [(3, 0, 0, 1), (2, 1, 1, 1), (2, 1, 1, 0), (4, 0, 0, 1)])
@pytest.mark.parametrize("inlet_type,outlet_type", [(2, 2), (2, 3), (2, 4), (6, 0), (7, 0), (6, 3)])
def test_quadrilateral_channel(lib_built, oracle_built, order, bt, ir, eq, inlet_type, outlet_type):
    m = ac.box(warp=0.06)
    specs = _quad_specs(inlet_type, outlet_type)
    op, orc = _pair(m, order, eq, bt, ir, specs, 2)
    attrs = ([3] if inlet_type in (6, 7) else []) + ([4] if outlet_type >= 2 else [])
    _sequence(op, orc, ac.dry_state(orc.node_coords(), 2), attrs)


def test_given_tangent_equals_derived_tangent(lib_built, oracle_built):
    """the tangent of a y = const patch of the box is +-e_x: handing it over explicitly must change nothing"""
    import torch
    m = ac.box(warp=0.0)
    U = None
    res = []
    for tang in ((0.0, 0.0, 0.0), (1.0, 0.0, 0.0)):
        op, orc = _pair(m, 3, 1, 0, 0, _quad_specs(6, 2, tang, tang), 2)
        if U is None:
            U = ac.dry_state(orc.node_coords(), 2)
        op.set_time_step(1e-6)
        op.Mult(torch.from_numpy(U).cuda())
        res.append((op.Mult(torch.from_numpy(U).cuda()).cpu().numpy(), op.bc_state(3)[1], op.bc_state(4)[1]))
    for a, b in zip(res[0], res[1]):
        assert rel_l2(a, b) < 1e-14


@pytest.mark.parametrize("inlet_type,outlet_type,warp", [(6, 2, 0.0), (2, 3, 0.08), (7, 4, 0.08)])
def test_hexahedral_channel(lib_built, oracle_built, inlet_type, outlet_type, warp):
    """3-D dry-air channel: attr 1 x- inlet, 2 x+ outlet, walls elsewhere; a run with these conditions uses the generic path"""
    lo, hi = (0.0, 0.0, 0.0), (2.0, 1.2, 1.0)
    m = tps_b200.cartesian_hex_mesh(4, 3, 3, lo=lo, hi=hi, periodic=(0, 0, 0))
    m["face_attr"] = box_face_attrs(m, lo, hi)
    if warp:
        m = dict(warp_mesh(m, amp=warp, lo=lo, hi=hi), face_attr=m["face_attr"])
    inlet = (1.2, 25.0, 1.0, -2.0)
    area = 1.2 * 1.0
    specs = [(1, 0, inlet_type, _nr(inlet) if inlet_type in (6, 7) else inlet),
             (2, 1, outlet_type, _nr((101300.0,) if outlet_type == 2 else (1.2 * 25.0 * area,))),
             (3, 2, 3, (310.0,)), (4, 2, 2, ()), (5, 2, 0, ()), (6, 2, 3, (290.0,))]
    op, orc = _pair(m, 3, 1, 0, 0, specs, 3)
    U = tgv_state(orc.node_coords() * np.pi)
    _sequence(op, orc, U, ([1] if inlet_type in (6, 7) else []) + [2], dt=1e-6)


def test_rk4_steps_with_nonreflecting_outlet(lib_built, oracle_built):
    """tpsb_ode_step hands its step size to the boundary conditions (BoundaryCondition::dt is M2ulPhyS::dt): 5 RK4 steps = 20
    stateful evaluations"""
    import torch
    m = ac.box(warp=0.05)
    op, orc = _pair(m, 3, 1, 0, 0, _quad_specs(6, 2), 2)
    U = ac.dry_state(orc.node_coords(), 2)
    x = torch.from_numpy(U.copy()).cuda()
    op.ode_step(x, 2e-6, scheme=4, nsteps=5)
    ref = orc.rk4(U, 2e-6, 5)
    assert rel_l2(x.cpu().numpy(), ref) < 1e-12
    for a in (3, 4):
        assert rel_l2(op.bc_state(a)[1], orc.bc_state(a)[1]) < 1e-12


def test_refusals(lib_built):
    m = ac.box()
    with pytest.raises(tps_b200.TpsbError, match="refLength"):
        tps_b200.RhsOperator(m, order=2, nvel=2, face_attr=m["face_attr"],
                             bcs=[tps_b200.BcDesc.make(*b) for b in [(1, 2, 0, ()), (2, 2, 0, ()), (3, 2, 0, ()), (4, 1, 2, (1e5,))]])
    with pytest.raises(tps_b200.TpsbError, match="planar 2-D or 3-D only"):
        tps_b200.RhsOperator(m, order=2, nvel=3, face_attr=m["face_attr"],
                             bcs=[tps_b200.BcDesc.make(*b) for b in [(1, 2, 0, ()), (2, 2, 0, ()), (3, 2, 0, ()),
                                                                    (4, 1, 2, _nr((1e5,)))]])
