"""The LTE working fluid with 1-D look-up tables (LteMixture src/lte_mixture.cpp, LteTransport
src/lte_transport_properties.cpp, LinearTable src/table.cpp; what M2ulPhyS builds for flow/lte/table_dim = 1,
src/M2ulPhyS.cpp:175-258) through the C ABI against the oracle, whose LTE back end IS the reference's LteMixture /
LteTransport / Fluxes / RiemannSolverTPS / WallBC object code (oracle/_ref).  The reference's table files are LFS pointers, so
the tables here are synthetic (smooth, strictly increasing energy; the logic under test - Newton inversion of e(T), p = rho R(T) T,
tabulated speed of sound and transport, boundary-state construction, radiative sink - does not depend on their values)."""
import os

import numpy as np
import pytest

import axisym_cases as ac
import oracle_api
import tps_b200
from common import box_face_attrs, rel_l2, warp_mesh

pytestmark = pytest.mark.gpu
HAVE_REF = os.path.exists(os.path.join(oracle_api.ORACLE_DIR, "_ref", "liboracle_ref.so"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref (reference object code) not built")


def make_tables(nec=False, n=80):
    """argon-like equilibrium tables on a non-uniform temperature grid"""
    T = 250.0 * np.exp(np.linspace(0.0, np.log(100.0), n))           # 250 K .. 25 000 K
    ion = 0.5 * (1 + np.tanh((T - 11000.0) / 2500.0))                 # ionisation fraction
    R = 208.13 * (1 + ion)
    energy = 1.5 * R * T + ion * 3.8e7 + 40.0 * T                     # strictly increasing in T
    c = np.sqrt((5.0 / 3.0 - 0.35 * ion * (1 - ion) * 4) * R * T)
    Tt = np.linspace(200.0, 26000.0, 57)
    mu = 2.2e-5 * (Tt / 300.0) ** 0.72 * (1 - 0.6 * 0.5 * (1 + np.tanh((Tt - 12000.0) / 2000.0)))
    kappa = 0.018 * (Tt / 300.0) ** 0.8 + 1.5 * np.exp(-((Tt - 14000.0) / 3000.0) ** 2)
    sigma = 1e4 * 0.5 * (1 + np.tanh((Tt - 9000.0) / 1500.0))
    necv = None
    if nec:
        Tn = np.linspace(3000.0, 25000.0, 23)
        necv = (Tn, 2.0e2 * np.exp((Tn - 9000.0) / 1800.0), 0, 1)   # log scale of the coefficient, as the reference's NEC table
    return tps_b200.LteTables.make(T, energy, R, c, Tt, mu, kappa, sigma, necv)


def primitives(xy, nvel, seed=20261018):
    """smooth hot-gas primitives [rho, u (nvel), T] per node, T between 4000 K and 13 000 K (across the ionisation ramp)"""
    x, y = xy[:, 0], xy[:, 1]
    z = xy[:, 2] if xy.shape[1] > 2 else 0 * x
    rho = 0.08 + 0.02 * np.sin(2 * x + z) * np.cos(3 * y)
    vel = [40 * np.sin(2 * x) * np.cos(y) + 15, -25 * np.cos(x) * np.sin(2 * y + z) + 5, 12 * np.cos(x + y - z) + 2][:nvel]
    T = 8500.0 + 4500.0 * np.sin(1.5 * x + 0.7) * np.cos(2 * y - z)
    rng = np.random.default_rng(seed)
    up = np.stack([rho] + vel + [T], axis=1)
    return np.ascontiguousarray(up * (1 + 0.005 * rng.uniform(-1, 1, up.shape)))


def _pair(m, order, eq, bt, ir, nvel, specs, ubg, tables, dim):
    orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.lte_params(tables, eq), kind="ref", basis_type=bt, int_rule=ir, neq=nvel + 2, nvel=nvel)
    if specs:
        orc.set_bcs(m["face_attr"], [oracle_api.make_bc(*b) for b in specs], ubg)
    op = tps_b200.RhsOperator(m, order=order, physics=tps_b200.Physics.lte_fluid(tables, eq), basis_type=bt, int_rule_type=ir, nvel=nvel,
                              face_attr=m["face_attr"] if specs else None, use_bc_in_grad=ubg,
                              bcs=[tps_b200.BcDesc.make(*b) for b in specs] if specs else None)
    assert op.path() == "generic" and op.neq == nvel + 2
    return op, orc


def _compare(op, orc, up):
    import torch
    U = np.ascontiguousarray(orc.pt("cons", up).T).reshape(-1)
    N, neq = orc.N, op.neq
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo, go = orc.mult(U, want_grad=True)
    upd, g = op.fields()
    assert rel_l2(upd.cpu().numpy(), orc.primitives(U)) < 1e-13      # Newton inversion of the energy table
    assert rel_l2(upd.cpu().numpy().reshape(neq, -1)[-1], up[:, -1]) < 1e-12   # ... recovers the temperature the state was built from
    assert rel_l2(g.cpu().numpy(), go) < 1e-11
    for k in range(neq):
        ref = yo[k * N:(k + 1) * N]
        assert np.linalg.norm(y[k * N:(k + 1) * N] - ref) <= 1e-10 * np.linalg.norm(ref), k
    assert abs(op.max_char_speed() / orc.max_char_speed - 1) < 1e-13
    return U


# attr 1 x = lo inviscid wall, 2 x = hi isothermal wall, 3 y = lo inlet, 4 y = hi outlet
C4 = [(1, 2, 0, ()), (2, 2, 3, (3000.0,)), (3, 0, 2, (0.09, 10.0, 45.0, 3.0)), (4, 1, 0, (18000.0,))]
WALLS = [(1, 2, 2, ()), (2, 2, 2, ()), (3, 2, 1, ()), (4, 2, 3, (5000.0,))]


@needs_ref
@pytest.mark.parametrize("order,bt,ir", [(3, 0, 0), (2, 1, 1), (4, 0, 0)])
@pytest.mark.parametrize("nvel,eq,specs,ubg,nec", [(2, 1, C4, True, False), (2, 1, C4, False, True), (2, 1, WALLS, False, False),
                                                  (2, 0, C4, False, False), (3, 1, C4, True, True), (3, 1, WALLS, False, False)])
def test_quadrilaterals(lib_built, oracle_built, order, bt, ir, nvel, eq, specs, ubg, nec):
    """planar (nvel = 2) and axisymmetric (nvel = 3) runs with inlet / outlet / inviscid, slip, adiabatic and isothermal walls"""
    m = ac.box(warp=0.06)
    op, orc = _pair(m, order, eq, bt, ir, nvel, specs, ubg, make_tables(nec), 2)
    _compare(op, orc, primitives(orc.node_coords(), nvel))


@needs_ref
@pytest.mark.parametrize("periodic,warp,nec", [(True, 0.08, True), (False, 0.08, False)])
def test_hexahedra(lib_built, oracle_built, periodic, warp, nec):
    lo, hi = (0.0, 0.0, 0.0), (2.0, 1.2, 1.0)
    per = (1, 1, 1) if periodic else (0, 0, 0)
    m = tps_b200.cartesian_hex_mesh(3, 3, 3, lo=lo, hi=hi, periodic=per)
    specs = None
    if not periodic:
        m["face_attr"] = box_face_attrs(m, lo, hi)
        specs = [(1, 0, 2, (0.09, 45.0, 3.0, -2.0)), (2, 1, 0, (18000.0,)), (3, 2, 3, (4000.0,)), (4, 2, 2, ()), (5, 2, 0, ()),
                 (6, 2, 1, ())]
        m = dict(warp_mesh(m, amp=warp, lo=lo, hi=hi), face_attr=m["face_attr"])
    else:
        m = warp_mesh(m, amp=warp, lo=lo, hi=hi)
    op, orc = _pair(m, 3, 1, 0, 0, 3, specs, True, make_tables(nec), 3)
    _compare(op, orc, primitives(orc.node_coords() * np.pi, 3))


@needs_ref
def test_rk4_steps(lib_built, oracle_built):
    import torch
    m = ac.box(warp=0.05)
    op, orc = _pair(m, 3, 1, 0, 0, 2, C4, True, make_tables(True), 2)
    U = np.ascontiguousarray(orc.pt("cons", primitives(orc.node_coords(), 2)).T).reshape(-1)
    x = torch.from_numpy(U.copy()).cuda()
    op.ode_step(x, 1e-7, scheme=4, nsteps=10)
    assert rel_l2(x.cpu().numpy(), orc.rk4(U, 1e-7, 10)) < 1e-12


def test_refusals(lib_built):
    m = ac.box()
    t = make_tables()
    with pytest.raises(tps_b200.TpsbError, match="mixing length"):
        tps_b200.RhsOperator(m, order=2, physics=tps_b200.Physics.lte_fluid(t).with_mixing_length(0.1), nvel=2, face_attr=m["face_attr"],
                             bcs=[tps_b200.BcDesc.make(*b) for b in WALLS])
    bad = tps_b200.LteTables.make([300.0, 200.0], [1.0, 2.0], [1, 1], [1, 1], [1, 2], [1, 1], [1, 1], [1, 1])
    with pytest.raises(tps_b200.TpsbError, match="increase strictly"):
        tps_b200.RhsOperator(m, order=2, physics=tps_b200.Physics.lte_fluid(bad), nvel=2, face_attr=m["face_attr"],
                             bcs=[tps_b200.BcDesc.make(*b) for b in WALLS])


def test_ideal_gas_tables_reproduce_the_dry_air_operator_on_the_device(lib_built):
    """no oracle involved: the device's LTE path over tables that describe a calorically perfect gas against the device's own
    dry-air path (Euler, so the only interpolation error is the tabulated sound speed: 1000 points over 250 K, 2e-8)"""
    import torch
    from test_cpu_nr_lte_oracle import ideal_gas_tables
    m = ac.box(warp=0.05)
    specs = ac.bcs("c4", 2)
    kw = dict(order=3, nvel=2, face_attr=m["face_attr"], use_bc_in_grad=False, bcs=[tps_b200.BcDesc.make(*b) for b in specs])
    lte = tps_b200.RhsOperator(m, physics=tps_b200.Physics.lte_fluid(ideal_gas_tables(), 0), **kw)
    dry = tps_b200.RhsOperator(m, physics=tps_b200.Physics.dry_air(0), **kw)
    assert lte.path() == "generic" and dry.path() == "generic"
    from common import node_coords_from_mesh  # noqa: F401
    N = dry.N
    rng = np.random.default_rng(5)
    rho = 1.2 + 0.05 * rng.uniform(-1, 1, N)
    u, v = 20 + 5 * rng.uniform(-1, 1, N), 8 + 5 * rng.uniform(-1, 1, N)
    p = 101300 + 800 * rng.uniform(-1, 1, N)
    U = np.concatenate([rho, rho * u, rho * v, p / 0.4 + 0.5 * rho * (u * u + v * v)])
    x = torch.from_numpy(U).cuda()
    yl, yd = lte.Mult(x).cpu().numpy(), dry.Mult(x).cpu().numpy()
    assert rel_l2(lte.fields()[0].cpu().numpy(), dry.fields()[0].cpu().numpy()) < 1e-13
    for k in range(4):
        assert rel_l2(yl[k * N:(k + 1) * N], yd[k * N:(k + 1) * N]) < 2e-8, k
    assert abs(lte.max_char_speed() / dry.max_char_speed() - 1) < 3e-8
