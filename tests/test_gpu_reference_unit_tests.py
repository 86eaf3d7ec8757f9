"""The reference's own unit tests for this path, re-expressed against the CUDA kernels (through the C ABI) and the
oracle's reference object code: test/test_speed_of_sound.cpp + inputs/perfectGas.air.ini,
test/test_perfect_mixture.cpp + inputs/perfectGas.argon.ini (state round trips in the four ambipolar x two-temperature
combinations), test/test_gradient.cpp (gradient of a sine field converges), test/test_table.cpp (linear tables)."""
import os

import numpy as np
import pytest

import oracle_api
import tps_b200
from common import rel_l2

pytestmark = pytest.mark.gpu
PI = np.pi
RU = 8.3144598
HAVE_REF = os.path.exists(os.path.join(oracle_api.ORACLE_DIR, "_ref", "liboracle_ref.so"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref (reference object code) not built")


def _pair(d, neq, n=(3, 3), order=1):
    pm = tps_b200.PlasmaModels.from_dict(d)
    m = tps_b200.cartesian_quad_mesh(*n, lo=(-PI, -PI), hi=(PI, PI))
    op = tps_b200.RhsOperator(m, order=order, physics=tps_b200.Physics.plasma_mixture(pm, 1))
    orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.mixture_params(pm, 1), kind="ref", neq=neq, nvel=2)
    assert op.neq == neq
    return op, orc


def _species(mw, charge, cv, fe=None):
    fe = fe or [0.0] * len(mw)
    return [dict(mw=mw[i], charge=charge[i], formation_energy=fe[i], molar_cv=cv[i], diffusivity=1e-3, mt_freq=1.0)
            for i in range(len(mw))]


@needs_ref
def test_speed_of_sound_of_dry_air_composition(lib_built, oracle_built):
    """test/test_speed_of_sound.cpp with test/inputs/perfectGas.air.ini: [CO2, Ar, O2, E, N2] (mixture order),
    non-ambipolar, single temperature, at 293.15 K and 1.2041 kg/m^3: c within 1e-4 of sqrt(1.4 R_u / M_air T)."""
    import torch
    C, O, AR, N, E = 12.011e-3, 15.999e-3, 39.948e-3, 14.0067e-3, 5.4858e-07
    d = dict(ambipolar=False, two_temperature=False, viscosity=1e-5, bulk_viscosity=0.0, thermal_conductivity=1e-2,
             electron_thermal_conductivity=0.0,
             species=_species([C + 2 * O, AR, 2 * O, E, 2 * N], [0, 0, 0, -1, 0], [3.4230, 1.5, 2.5257, 1.5, 2.5017]))
    op, orc = _pair(d, neq=2 + 2 + 4)
    airMW, rho, T_h = 28.964e-3, 1.2041, 293.15
    X = np.array([0.0407e-2, 0.934e-2, 20.946e-2, 0.0])
    prim = np.zeros((1, 8))
    prim[0, 0], prim[0, 3] = rho, T_h
    prim[0, 4:8] = rho / airMW * X
    U = orc.pt("cons", prim)
    ref = np.sqrt(1.4 * RU / airMW * T_h)
    c_orc = orc.pt("max_char_speed", U)[0]  # u = 0: |u| + c = c
    c_dev = op.point_eval("max_char_speed", torch.from_numpy(U).cuda()).cpu().numpy()[0]
    assert abs(c_orc - ref) / ref < 1e-4
    assert abs(c_dev - ref) / ref < 1e-4
    assert abs(c_dev - c_orc) / c_orc < 1e-14
    assert np.allclose(op.point_eval("prim", torch.from_numpy(U).cuda()).cpu().numpy(), prim, rtol=1e-13, atol=1e-300)


@needs_ref
@pytest.mark.parametrize("ambipolar", [False, True])
@pytest.mark.parametrize("two_t", [False, True])
def test_perfect_mixture_state_round_trips(lib_built, oracle_built, ambipolar, two_t):
    """test/test_perfect_mixture.cpp with test/inputs/perfectGas.argon.ini: [Ar.+1, Ar2.+1, E, Ar] (the test's electron
    mass 1.182592e-2), primitives -> conserved (reference) -> primitives (device) to 1e-13, and the pressure /
    characteristic speed of the device equal to the reference's."""
    import torch
    AR, E = 39.948e-3, 1.182592e-2
    d = dict(ambipolar=ambipolar, two_temperature=two_t, viscosity=1e-5, bulk_viscosity=0.0, thermal_conductivity=1e-2,
             electron_thermal_conductivity=1e-2,
             species=_species([AR - E, 2 * AR - E, E, AR], [1, 1, -1, 0], [1.5, 2.5, 1.5, 1.5], [1.521e6, 2.666e6, 0.0, 0.0]))
    nact = 2 if ambipolar else 3
    neq = 4 + nact + (1 if two_t else 0)
    op, orc = _pair(d, neq=neq)
    rng = np.random.default_rng(5)
    n = 200
    prim = np.zeros((n, neq))
    prim[:, 1:3] = rng.uniform(-100, 100, (n, 2))
    prim[:, 3] = rng.uniform(300, 5000, n)
    nI, nI2 = rng.uniform(0.01, 1.0, n), rng.uniform(0.01, 1.0, n)
    nAr = rng.uniform(5.0, 50.0, n)
    ne = nI + nI2  # quasi-neutral (required when ambipolar, harmless otherwise)
    prim[:, 4], prim[:, 5] = nI, nI2
    if not ambipolar:
        prim[:, 6] = ne
    prim[:, 0] = nI * (AR - E) + nI2 * (2 * AR - E) + ne * E + nAr * AR
    if two_t:
        prim[:, neq - 1] = rng.uniform(300, 20000, n)
    U = orc.pt("cons", prim)
    Ud = torch.from_numpy(U).cuda()
    back = op.point_eval("prim", Ud).cpu().numpy()
    assert np.abs(back / prim - 1.0)[:, prim[0] != 0].max() < 1e-13
    assert np.allclose(back, orc.pt("prim", U), rtol=1e-14, atol=1e-300)
    c_dev, c_ref = op.point_eval("max_char_speed", Ud).cpu().numpy(), orc.pt("max_char_speed", U)
    assert np.abs(c_dev / c_ref - 1).max() < 1e-13
    f_dev, f_ref = op.point_eval("conv_flux", Ud).cpu().numpy(), orc.pt("conv_flux", U)
    assert np.abs(f_dev - f_ref).max() <= 1e-13 * np.abs(f_ref).max()


@pytest.mark.parametrize("order", [1, 2, 3])
def test_gradient_of_a_sine_field_converges(lib_built, order):
    """test/test_gradient.cpp: the BR1 gradient of a smooth periodic field converges under mesh refinement."""
    import torch
    errs = []
    for n in (4, 8):
        m = tps_b200.cartesian_hex_mesh(n, n, n, lo=(-PI,) * 3, hi=(PI,) * 3)
        op = tps_b200.RhsOperator(m, order=order, physics=tps_b200.Physics.dry_air(1, 1.0))
        from common import node_coords_from_mesh
        X = node_coords_from_mesh(m["elem_xyz"], order)
        x, y, z = X[:, 0], X[:, 1], X[:, 2]
        rho = 1.2 + 0.2 * np.sin(x) * np.cos(y) * np.sin(z + 0.3)
        u, v, w = 30 * np.sin(y + 0.1), 20 * np.cos(z) * np.sin(x), 10 * np.cos(x + y)
        T = 300 + 40 * np.sin(x) * np.sin(y) * np.cos(z)
        p = rho * 287.058 * T
        U = np.concatenate([rho, rho * u, rho * v, rho * w, p / 0.4 + 0.5 * rho * (u * u + v * v + w * w)])
        op.updateGradients(torch.from_numpy(U).cuda())
        g = op.fields()[1].cpu().numpy().reshape(3, 5, -1)
        exact_dTdx = 40 * np.cos(x) * np.sin(y) * np.cos(z)
        exact_dudy = 30 * np.cos(y + 0.1)
        errs.append((rel_l2(g[0, 4], exact_dTdx), rel_l2(g[1, 1], exact_dudy)))
    for a, b in zip(errs[0], errs[1]):
        assert b < a / 2 ** (order - 0.5), (order, errs)


def test_linear_table_lookup(lib_built):
    """test/test_table.cpp logic through a tabulated reaction rate: values at the knots, linear between them,
    end-interval extrapolation outside (TableInterpolator::findInterval, LinearTable::eval)."""
    import torch
    import plasma_cases
    d = plasma_cases.ternary_dict()
    T = np.array([300.0, 500.0, 900.0, 1700.0, 3300.0])
    kf = np.array([1.0, 3.0, 2.0, 8.0, 5.0])
    rx = dict(d["reactions"][0])
    rx.update(model=2, table=(T, kf, False, False), detailed=False)
    d["reactions"] = [rx]
    pm = tps_b200.PlasmaModels.from_dict(d)
    m = tps_b200.cartesian_quad_mesh(3, 3, lo=(-PI, -PI), hi=(PI, PI))
    op = tps_b200.RhsOperator(m, order=1, physics=tps_b200.Physics.plasma_mixture(pm, 1))
    Te = np.array([300.0, 400.0, 500.0, 700.0, 1300.0, 3300.0, 200.0, 4000.0])
    n = len(Te)
    # conserved states built by hand: rho, momentum 0, n_ion, T_h = 1000 K; source of the ion density = M_ion kf n_Ar n_e
    nI, nAr = 0.2, 20.0
    MW_AR, MW_E = plasma_cases.MW_AR, plasma_cases.MW_E
    rho = nI * (MW_AR - MW_E) + nI * MW_E + nAr * MW_AR
    cv = 1.5 * RU
    U = np.zeros((n, 6))
    U[:, 0] = rho
    U[:, 4] = nI * (MW_AR - MW_E)
    U[:, 5] = nI * cv * Te
    U[:, 3] = (nI + nAr) * cv * 1000.0 + U[:, 5] + nI * 1.521e4
    g = np.zeros((n, 12))
    src = op.point_eval("source", torch.from_numpy(U).cuda(), torch.from_numpy(g).cuda()).cpu().numpy()
    expect = np.interp(Te, T, kf)
    expect[6] = kf[0] + (kf[1] - kf[0]) / (T[1] - T[0]) * (200.0 - T[0])   # extrapolated from the first interval
    expect[7] = kf[3] + (kf[4] - kf[3]) / (T[4] - T[3]) * (4000.0 - T[3])  # ... and from the last
    got = src[:, 4] / ((MW_AR - MW_E) * nAr * nI)
    assert np.allclose(got, expect, rtol=1e-12)
