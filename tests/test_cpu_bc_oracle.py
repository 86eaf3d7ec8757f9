"""Boundary conditions of the CPU oracle: the thin BC logic (restated from wallBC.cpp / inletBC.cpp /
outletBC.cpp) over per-point physics served either by the port or by the reference's own object code."""
import os

import numpy as np
import pytest

import oracle_api
import tps_b200
from common import box_face_attrs, rel_l2, warp_mesh

PI = np.pi
HAVE_REF = os.path.exists(os.path.join(oracle_api.ORACLE_DIR, "_ref", "liboracle_ref.so")) or os.path.isdir("/root/reference/src")

# (kind, type, data): inlet SUB_DENS_VEL, outlet SUB_P, walls INV / VISC_ADIAB / VISC_ISOTH
BCS = [(0, 2, (1.2, 20.0, 1.0, -2.0)), (1, 0, (101300.0,)), (2, 0, ()), (2, 2, ()), (2, 3, (300.0,))]


def _tiny_oracle(kind, eq=1):
    m = tps_b200.cartesian_hex_mesh(1, 1, 1, periodic=(0, 0, 0))
    return oracle_api.Oracle(1, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                             phys=oracle_api.dry_air_params(eq, 5.0, 0.3), nthreads=1, kind=kind)


def _random_point(rng):
    rho = rng.uniform(0.9, 1.4)
    v = rng.uniform(-60, 60, 3)
    p = rng.uniform(0.8e5, 1.2e5)
    U = np.array([rho, *(rho * v), p / 0.4 + 0.5 * rho * (v @ v)])
    g = rng.normal(size=15) * np.array([0.1, 50, 50, 50, 100] * 3)
    n = rng.normal(size=3) * 0.3
    return U, g, n


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not available")
def test_bc_fluxes_port_equals_reference_object_code(oracle_built):
    """Every BC flux of the port physics agrees with the same logic over the reference's compiled
    DryAir / Fluxes::ComputeBdrViscousFluxes / RiemannSolverTPS to round-off."""
    a, b = _tiny_oracle("port"), _tiny_oracle("ref")
    rng = np.random.default_rng(7)
    for kind, typ, data in BCS:
        bc = oracle_api.make_bc(1, kind, typ, data)
        for use in (False, True):
            for _ in range(20):
                U, g, n = _random_point(rng)
                fa, fb = a.bc_flux(bc, n, U, g, use), b.bc_flux(bc, n, U, g, use)
                assert np.allclose(fa, fb, rtol=2e-13, atol=1e-9 * np.abs(fb).max()), (kind, typ, fa, fb)


def test_boundary_viscous_flux_without_prescription_is_viscous_flux_dot_n(oracle_built):
    """test/test_boundary_flux.cpp re-expressed for dry air: with nothing prescribed, the boundary form of
    the viscous flux equals ComputeViscousFluxes . n (tolerance 5e-13 as in the reference test)."""
    lib = oracle_api.load("port")
    o = _tiny_oracle("port")
    rng = np.random.default_rng(3)
    # an adiabatic wall prescribes the heat flux only: momentum rows must equal the interior stress . n
    bc = oracle_api.make_bc(1, 2, 2)
    for _ in range(10):
        U, g, n = _random_point(rng)
        U[1:4] = 0.0  # stagnation state == state: both viscous evaluations see the same state
        f = o.bc_flux(bc, n, U, g)
        # LF flux of identical states = F(U).n = (0, p n, 0); subtract it
        p = 0.4 * U[4]
        visc_n = -(f[1:4] - p * n)  # = 1/2 wall + 1/2 interior stress . n
        F = np.zeros(15)
        lib.orc_phys_init(oracle_api.dry_air_params(1, 5.0, 0.3))
        lib.orc_phys_visc_flux(1, U, g, F)
        ref = F.reshape(3, 5)[:, 1:4].T @ n
        # 5e-13 relative as in the reference test, plus the round-off of removing the O(p |n|) pressure term
        assert np.abs(visc_n - ref).max() <= 5e-13 * np.abs(ref).max() + 8 * 2.3e-16 * p * np.abs(n).max()


def test_outlet_at_interior_pressure_and_inlet_at_interior_state_are_transparent(oracle_built):
    o = _tiny_oracle("port", eq=0)
    rng = np.random.default_rng(5)
    lib = oracle_api.load("port")
    lib.orc_phys_init(oracle_api.dry_air_params(0, 1.0, 0.0))
    for _ in range(10):
        U, g, n = _random_point(rng)
        F = np.zeros(15)
        lib.orc_phys_conv_flux(1, U, F)
        fn = F.reshape(3, 5).T @ n
        p = 0.4 * (U[4] - 0.5 * (U[1:4] @ U[1:4]) / U[0])
        out = o.bc_flux(oracle_api.make_bc(1, 1, 0, (p,)), n, U, g)
        assert np.allclose(out, fn, rtol=1e-12, atol=1e-9 * np.abs(fn).max())
        inl = o.bc_flux(oracle_api.make_bc(1, 0, 2, (U[0], *(U[1:4] / U[0]))), n, U, g)
        assert np.allclose(inl, fn, rtol=1e-12, atol=1e-9 * np.abs(fn).max())


@pytest.mark.parametrize("order", [1, 3])
def test_uniform_flow_through_warped_channel_has_zero_residual(oracle_built, order):
    """Free-stream preservation with inlet (x-), outlet (x+) and inviscid walls on a warped (trilinear) box:
    a uniform axial flow satisfies every BC exactly, so dU/dt = 0 up to the metric identities' round-off."""
    lo, hi = (0.0, 0.0, 0.0), (2.0, 1.0, 1.0)
    m0 = tps_b200.cartesian_hex_mesh(4, 3, 3, lo=lo, hi=hi, periodic=(0, 0, 0))
    attr = box_face_attrs(m0, lo, hi)
    o = oracle_api.Oracle(order, m0["elem_xyz"], m0["face_el1"], m0["face_el2"], m0["face_inf1"], m0["face_inf2"],
                          phys=oracle_api.dry_air_params(1, 1.0, 0.0))
    rho, u, p = 1.2, 30.0, 101300.0
    bcs = [oracle_api.make_bc(1, 0, 2, (rho, u, 0.0, 0.0)), oracle_api.make_bc(2, 1, 0, (p,))]
    bcs += [oracle_api.make_bc(a, 2, 0) for a in (3, 4, 5, 6)]
    o.set_bcs(attr, bcs)
    N = o.N
    U = np.concatenate([np.full(N, rho), np.full(N, rho * u), np.zeros(N), np.zeros(N), np.full(N, p / 0.4 + 0.5 * rho * u * u)])
    y = o.mult(U)
    assert np.abs(y[:N]).max() < 1e-9 * rho * u and np.abs(y[4 * N:]).max() < 1e-9 * (p / 0.4) * u


def test_isothermal_wall_gradient_uses_the_wall_state(oracle_built):
    """useBCinGrad: at an isothermal wall the BR1 jump is 1/2 (Up_bc - Up), Up_bc = (rho, 0, T_wall)
    (src/faceGradientIntegration.cpp:96-115, src/wallBC.cpp:241-266): for a fluid at rest at T_wall the
    gradient is unchanged, for a moving fluid it is not."""
    lo, hi = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
    m = tps_b200.cartesian_hex_mesh(3, 2, 3, lo=lo, hi=hi, periodic=(1, 0, 1))
    attr = box_face_attrs(m, lo, hi)
    Tw, rho = 300.0, 1.2
    bcs = [oracle_api.make_bc(3, 2, 3, (Tw,)), oracle_api.make_bc(4, 2, 3, (Tw,))]
    N = None
    res = {}
    for use in (False, True):
        o = oracle_api.Oracle(2, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                              phys=oracle_api.dry_air_params(1, 1.0, 0.0))
        o.set_bcs(attr, bcs, use)
        N = o.N
        for vel in (0.0, 10.0):
            E = rho * 287.058 * Tw / 0.4 + 0.5 * rho * vel * vel
            U = np.concatenate([np.full(N, rho), np.full(N, rho * vel), np.zeros(N), np.zeros(N), np.full(N, E)])
            res[(use, vel)] = o.gradients(U)
    assert np.abs(res[(False, 0.0)]).max() < 1e-9 and np.abs(res[(True, 0.0)]).max() < 1e-9
    assert np.abs(res[(False, 10.0)]).max() < 1e-9
    assert np.abs(res[(True, 10.0)]).max() > 1.0  # du/dy picks up the no-slip jump
