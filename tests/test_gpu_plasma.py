"""Plasma models on the CUDA path against the reference's OWN compiled classes (PerfectMixture, ConstantTransport,
Chemistry, Fluxes, RiemannSolverTPS in oracle/_ref): point-wise physics, then the full DG right-hand side of the
mms.ternary_2d configuration (BASELINE config C3: 2-D quads, p = 2, Gauss-Lobatto, ambipolar two-temperature argon
ternary mixture, constant transport, ionisation chemistry with detailed balance, SourceTerm)."""
import os

import numpy as np
import pytest

import oracle_api
import plasma_cases
import tps_b200
from common import rel_l2

pytestmark = pytest.mark.gpu
PI = np.pi
HAVE_REF = os.path.exists(os.path.join(oracle_api.ORACLE_DIR, "_ref", "liboracle_ref.so"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref (reference object code) not built")


def _pair(order=2, n=(5, 4), eq=1, **kw):
    pm = plasma_cases.ternary_models(**kw)
    m = tps_b200.cartesian_quad_mesh(*n, lo=(-PI, -PI), hi=(PI, PI))
    op = tps_b200.RhsOperator(m, order=order, physics=tps_b200.Physics.plasma_mixture(pm, eq), basis_type=1, int_rule_type=1)
    orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.mixture_params(pm, eq), kind="ref", basis_type=1, int_rule=1, neq=op.neq, nvel=2)
    return op, orc


@needs_ref
def test_pointwise_mixture_physics_equals_reference_classes(lib_built, oracle_built):
    import torch
    op, orc = _pair()
    assert op.neq == 6
    up = plasma_cases.random_primitives(500)
    U = orc.pt("cons", up)
    g = np.random.default_rng(1).normal(size=(500, 12)) * np.array([0.05, 20, 20, 50, 0.2, 300] * 2)
    Ud, gd = torch.from_numpy(U).cuda(), torch.from_numpy(g).cuda()
    for what, args, tol in (("prim", (U,), 1e-13), ("max_char_speed", (U,), 1e-13), ("conv_flux", (U,), 1e-13),
                            ("visc_flux", (U, g), 1e-12), ("source", (U, orc.pt("prim", U), g), 1e-12)):
        ref = orc.pt(what, *args)
        got = op.point_eval(what, Ud, gd if what in ("visc_flux", "source") else None).cpu().numpy()
        scale = np.abs(ref).max(axis=0) if ref.ndim > 1 else np.abs(ref).max()
        assert (np.abs(got - ref) <= tol * scale + 1e-300).all(), (what, np.abs(got - ref).max(axis=0) / scale)


@needs_ref
@pytest.mark.parametrize("two_t", [True, False])
def test_ternary_2d_rhs_parity(lib_built, oracle_built, two_t):
    import torch
    op, orc = _pair(n=(6, 5), two_temperature=two_t)
    xy = orc.node_coords()
    up = plasma_cases.smooth_primitives(xy)
    if not two_t:
        up = up[:, :5]
    U = np.ascontiguousarray(orc.pt("cons", up).T).reshape(-1)  # byNODES
    N = orc.N
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo, go = orc.mult(U, want_grad=True)
    assert rel_l2(op.fields()[0].cpu().numpy(), orc.primitives(U)) < 1e-13
    assert rel_l2(op.fields()[1].cpu().numpy(), go) < 1e-11
    for k in range(op.neq):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k
    assert abs(op.max_char_speed() / orc.max_char_speed - 1) < 1e-13


@needs_ref
def test_source_term_reads_the_solution_vector_not_the_stage_vector(lib_built, oracle_built):
    """Parity trap 1: SourceTerm takes the conserved state from the solution grid function U_
    (src/source_term.cpp:66,77,121) while Up / gradUp come from the stage vector."""
    import torch
    op, orc = _pair(n=(4, 4))
    xy = orc.node_coords()
    U = np.ascontiguousarray(orc.pt("cons", plasma_cases.smooth_primitives(xy)).T).reshape(-1)
    Usol = np.ascontiguousarray(orc.pt("cons", plasma_cases.smooth_primitives(xy, seed=5)).T).reshape(-1)
    sol = torch.from_numpy(Usol).cuda()
    op.set_solution_view(sol)
    orc.set_solution_view(Usol)
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo = orc.mult(U)
    N = orc.N
    for k in range(op.neq):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k
    op.set_solution_view(None)
    orc.set_solution_view(None)
    y2 = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    assert rel_l2(y2[5 * N:], y[5 * N:]) > 1e-6  # the electron-energy source really depends on U_
    assert rel_l2(y2, orc.mult(U)) < 1e-10


def _pair_models(pm, order=2, n=(5, 4), eq=1):
    m = tps_b200.cartesian_quad_mesh(*n, lo=(-PI, -PI), hi=(PI, PI))
    op = tps_b200.RhsOperator(m, order=order, physics=tps_b200.Physics.plasma_mixture(pm, eq), basis_type=1, int_rule_type=1)
    orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.mixture_params(pm, eq), kind="ref", basis_type=1, int_rule=1, neq=op.neq, nvel=2)
    return op, orc


MULT = dict(viscosity=1.3, bulk_viscosity=1.0, heavy_thermal_conductivity=0.7, electron_thermal_conductivity=1.9,
            momentum_transfer_frequency=2.5, diffusivity=0.6, mobility=1.4)


@needs_ref
@pytest.mark.parametrize("third,mult,two_t", [(False, None, True), (True, None, True), (True, MULT, True), (False, None, False)])
def test_argon_minimal_transport_equals_reference_classes(lib_built, oracle_built, third, mult, two_t):
    """GasMinimalTransport (Chapman-Enskog transport from the collision-integral fits, Debye-length-scaled Coulomb
    integrals, Curtiss-Hirschfelder diffusion, third-order electron conductivity, artificial multipliers) point-wise
    against the reference's own gas_transport.cpp / collision_integrals.cpp object code."""
    import torch
    pm = tps_b200.PlasmaModels.from_dict(plasma_cases.argon_minimal_dict(third, mult, two_t))
    op, orc = _pair_models(pm)
    rng = np.random.default_rng(7)
    n = 400
    up = np.zeros((n, op.neq))
    up[:, 0] = rng.uniform(0.03, 0.08, n)
    up[:, 1:3] = rng.uniform(-300, 300, (n, 2))
    up[:, 3] = rng.uniform(3000, 12000, n)
    up[:, 4] = rng.uniform(1e-4, 0.05, n)
    if two_t:
        up[:, 5] = rng.uniform(5000, 15000, n)
    U = orc.pt("cons", up)
    g = rng.normal(size=(n, 2 * op.neq)) * np.array(([0.005, 50, 50, 800, 0.01] + ([900] if two_t else [])) * 2)
    Ud, gd = torch.from_numpy(U).cuda(), torch.from_numpy(g).cuda()
    # the third-order conductivity divides by L11 - L12^2 / L22, a difference of nearly equal numbers: the 1-2 ulp
    # differences between device and host pow / log are amplified ~1e4 times there
    t = 5e-10 if third else 1e-11
    for what, args, tol in (("visc_flux", (U, g), t), ("source", (U, orc.pt("prim", U), g), 1e-11)):
        ref = orc.pt(what, *args)
        got = op.point_eval(what, Ud, gd).cpu().numpy()
        scale = np.abs(ref).max(axis=0)
        assert (np.abs(got - ref) <= tol * scale + 1e-300).all(), (what, np.abs(got - ref).max(axis=0) / scale)


@needs_ref
@pytest.mark.parametrize("third", [False, True])
def test_argon_minimal_ternary_rhs_parity(lib_built, oracle_built, third):
    import torch
    pm = tps_b200.PlasmaModels.from_dict(plasma_cases.argon_minimal_dict(third))
    op, orc = _pair_models(pm, n=(5, 4))
    up = plasma_cases.hot_primitives(orc.node_coords())
    U = np.ascontiguousarray(orc.pt("cons", up).T).reshape(-1)
    N = orc.N
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo, go = orc.mult(U, want_grad=True)
    assert rel_l2(op.fields()[1].cpu().numpy(), go) < 1e-11
    for k in range(op.neq):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k


def _tab_models(kind):
    d = plasma_cases.ternary_dict()
    T = np.geomspace(9.0e3, 1.3e4, 57)  # narrower than the electron temperatures of the test state
    kf = 4.7 * T ** 1.2 * np.exp(-6.49e4 / 8.3144598 / T) + 1e-30
    rx = dict(d["reactions"][0])
    if kind == "table_loglog":
        rx.update(model=2, table=(T, kf, True, True))
    elif kind == "table_lin":
        rx.update(model=2, table=(T, kf, False, False))
    else:
        rx.update(model=3, component=1)
    d["reactions"] = [rx]
    return tps_b200.PlasmaModels.from_dict(d)


@needs_ref
@pytest.mark.parametrize("kind", ["table_loglog", "table_lin", "gridfunction"])
def test_tabulated_and_grid_function_reaction_rates(lib_built, oracle_built, kind):
    """Reaction models TABULATED_RXN (LinearTable, binary search, log/linear scales; out-of-range temperatures
    extrapolate from the end intervals) and GRIDFUNCTION_RXN (rate coefficient per node from an external field)."""
    import torch
    pm = _tab_models(kind)
    op, orc = _pair_models(pm, n=(4, 4))
    N = orc.N
    up = plasma_cases.smooth_primitives(orc.node_coords())
    up[:, 5] *= 8.0  # electron temperatures 8000 .. 16000 K: inside and on both sides outside the table
    U = np.ascontiguousarray(orc.pt("cons", up).T).reshape(-1)
    if kind == "gridfunction":
        rates = np.abs(np.random.default_rng(3).normal(size=(2, N))) * 1e-3
        orc.set_rates(rates)
        op.set_reaction_rate_field(torch.from_numpy(rates).cuda())
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo = orc.mult(U)
    for k in range(op.neq):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k
    if kind == "gridfunction":  # without a field the reaction is switched off (reaction.cpp:113-116)
        op.set_reaction_rate_field(None)
        y0 = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
        assert rel_l2(y0[4 * N:5 * N], y[4 * N:5 * N]) > 1e-8


def test_mean_time_derivatives(lib_built):
    import torch
    m = tps_b200.cartesian_hex_mesh(3, 3, 3)
    op = tps_b200.RhsOperator(m, order=2)
    y = torch.from_numpy(np.random.default_rng(0).normal(size=5 * op.N)).cuda()
    got = op.mean_time_derivatives(y)
    ref = np.abs(y.cpu().numpy()).reshape(5, -1).mean(axis=1)
    assert np.allclose(got, ref, rtol=1e-13)


@needs_ref
def test_net_emission_radiation_sink(lib_built, oracle_built):
    """RadiationInput NET_EMISSION / TABULATED_NEC: -4 pi eps_N(T_h) on the total-energy equation."""
    import torch
    d = plasma_cases.ternary_dict()
    T = np.linspace(300.0, 600.0, 31)
    d["nec_table"] = (T, 1e3 * (T / 300.0) ** 4, False, False)
    pm = tps_b200.PlasmaModels.from_dict(d)
    op, orc = _pair_models(pm, n=(4, 4))
    op0, _ = _pair_models(plasma_cases.ternary_models(), n=(4, 4))
    N = orc.N
    U = np.ascontiguousarray(orc.pt("cons", plasma_cases.smooth_primitives(orc.node_coords())).T).reshape(-1)
    x = torch.from_numpy(U).cuda()
    y = op.Mult(x).cpu().numpy()
    yo = orc.mult(U)
    for k in range(op.neq):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k
    y0 = op0.Mult(x).cpu().numpy()
    Th = op.fields()[0].cpu().numpy()[3 * N:4 * N]
    sink = -4 * np.pi * np.interp(Th, T, 1e3 * (T / 300.0) ** 4)
    inside = (Th > 300) & (Th < 600)
    assert inside.any()
    assert np.allclose((y - y0)[3 * N:4 * N][inside], sink[inside], rtol=1e-9, atol=1e-6 * np.abs(y0[3 * N:4 * N]).max())


@needs_ref
@pytest.mark.parametrize("third,mult", [(False, None), (True, None), (True, MULT)])
def test_argon_mixture_transport_equals_reference_classes(lib_built, oracle_built, third, mult):
    """GasMixtureTransport (generic collision-integral transport over the species-pair collision types: screened
    Coulomb attractive / repulsive, Ar-Ar, Ar-Ar.+1, Ar-E) for the six-species argon mixture of config C4, point-wise
    against the reference's own gas_transport.cpp object code."""
    import axisym_cases as ac
    import torch
    d = ac.argon6_dict()
    d.update(transport_model="argon_mixture", third_order_k_electron=third)
    if mult:
        d["multipliers"] = mult
    pm = tps_b200.PlasmaModels.from_dict(d)
    op, orc = _pair_models(pm, order=2, n=(3, 3))
    assert op.neq == 10
    rng = np.random.default_rng(11)
    n = 300
    up = np.zeros((n, 10))
    up[:, 0] = rng.uniform(0.04, 0.08, n)
    up[:, 1:3] = rng.uniform(-300, 300, (n, 2))
    up[:, 3] = rng.uniform(3000, 12000, n)
    up[:, 4:9] = rng.uniform(1e-4, 0.03, (n, 5))
    up[:, 8] = up[:, 4]  # quasi-neutral: n_e = n_ion
    up[:, 9] = rng.uniform(5000, 15000, n)
    U = orc.pt("cons", up)
    g = rng.normal(size=(n, 20)) * np.array(([0.005, 50, 50, 800] + [0.01] * 5 + [900]) * 2)
    Ud, gd = torch.from_numpy(U).cuda(), torch.from_numpy(g).cuda()
    t = 5e-10 if third else 1e-11
    for what, args, tol in (("visc_flux", (U, g), t), ("source", (U, orc.pt("prim", U), g), 1e-11)):
        ref = orc.pt(what, *args)
        got = op.point_eval(what, Ud, gd).cpu().numpy()
        scale = np.abs(ref).max(axis=0)
        assert (np.abs(got - ref) <= tol * scale + 1e-300).all(), (what, np.abs(got - ref).max(axis=0) / (scale + 1e-300))


def nitrogen6_dict():
    """Six-species nitrogen mixture in mixture order [Ni.+1, Ni_e, N2_e, Ni, E, N2] with the species data of
    test/inputs/input.reactNitrogen.ini (formation energies, molar heat capacities); 'kind' = GasSpcs."""
    MW_N, MW_E = 14.0067e-3, 5.48579908782496e-7
    names = ["NI1P", "NI", "N2", "NI", "E", "N2"]
    mw = [MW_N - MW_E, MW_N, 2 * MW_N, MW_N, MW_E, 2 * MW_N]
    ch = [1.0, 0.0, 0.0, 0.0, -1.0, 0.0]
    fe = [1873823.43223, 758424.8665, 812408.331926, 470723.5922, 0.0, 0.0]
    cv = [1.5, 1.5, 2.5, 1.5, 1.5, 2.5]
    sp = [dict(mw=mw[i], charge=ch[i], formation_energy=fe[i], molar_cv=cv[i], diffusivity=1e-3, mt_freq=1e3, kind=names[i])
          for i in range(6)]
    return dict(ambipolar=False, two_temperature=True, viscosity=2.2e-3, bulk_viscosity=4.0e-4, thermal_conductivity=0.12,
                electron_thermal_conductivity=0.3, species=sp, reactions=[], ion_index=0, neutral_index=3)


@needs_ref
@pytest.mark.parametrize("third", [False, True])
def test_nitrogen_mixture_transport_equals_reference_classes(lib_built, oracle_built, third):
    """GasMixtureTransport over the nitrogen collision types (N-N, N2-N2, N2-N, N-N.+1, N2-N.+1, e-N, e-N2 curve fits of
    src/collision_integrals.cpp:210-625), point-wise against the reference's own object code."""
    import torch
    d = nitrogen6_dict()
    d.update(transport_model="argon_mixture", third_order_k_electron=third)
    pm = tps_b200.PlasmaModels.from_dict(d)
    ns = 6
    got_types = {pm.collision_index[i + j * ns] for i in range(ns) for j in range(i, ns)}
    assert got_types == {0, 1, 6, 7, 8, 10, 11, 12, 13}
    op, orc = _pair_models(pm, order=2, n=(3, 3))
    assert op.neq == 10
    rng = np.random.default_rng(12)
    n = 300
    up = np.zeros((n, 10))
    up[:, 0] = rng.uniform(0.04, 0.08, n)
    up[:, 1:3] = rng.uniform(-300, 300, (n, 2))
    up[:, 3] = rng.uniform(3000, 12000, n)
    up[:, 4:9] = rng.uniform(1e-4, 0.03, (n, 5))
    up[:, 8] = up[:, 4]  # quasi-neutral: n_e = n_ion
    up[:, 9] = rng.uniform(5000, 15000, n)
    # density consistent with a positive background (N2) number density
    MW_N, MW_E = 14.0067e-3, 5.48579908782496e-7
    mw = np.array([MW_N - MW_E, MW_N, 2 * MW_N, MW_N, MW_E])
    up[:, 0] = (up[:, 4:9] * mw).sum(axis=1) + rng.uniform(0.5, 2.0, n) * 2 * MW_N
    U = orc.pt("cons", up)
    g = rng.normal(size=(n, 20)) * np.array(([0.005, 50, 50, 800] + [0.01] * 5 + [900]) * 2)
    Ud, gd = torch.from_numpy(U).cuda(), torch.from_numpy(g).cuda()
    # the e-N / e-N2 fits are sums of terms ~1e4 cancelling to ~-45 (round-off ~1e-12 relative in each integral); the
    # third-order electron conductivity L11 - L12^2 / L22 cancels another 3-4 digits (measured 1.9e-9 on the energy flux)
    t = 5e-9 if third else 1e-10
    for what, args, tol in (("visc_flux", (U, g), t), ("source", (U, orc.pt("prim", U), g), 1e-10)):
        ref = orc.pt(what, *args)
        got = op.point_eval(what, Ud, gd).cpu().numpy()
        assert np.isfinite(ref).all()
        scale = np.abs(ref).max(axis=0)
        assert (np.abs(got - ref) <= tol * scale + 1e-300).all(), (what, np.abs(got - ref).max(axis=0) / (scale + 1e-300))
