"""CPU checks of the oracle's axisymmetric operator (config.isAxisymmetric(): 2-D mesh, three velocity components;
src/rhs_operator.cpp:191-205,441-445, src/domain_integrator.cpp:83-89, src/face_integrator.cpp:344-346,
src/BCintegrator.cpp:416-429, src/forcing_terms.cpp:255-380) and of boundary conditions in 2-D."""
import os

import numpy as np
import pytest

import axisym_cases as ac
import oracle_api
from common import rel_l2

HAVE_REF = os.path.exists(os.path.join(oracle_api.ORACLE_DIR, "_ref", "liboracle_ref.so"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref (reference object code) not built")


@pytest.fixture(scope="module")
def lib_cpu():
    import tps_b200
    if not os.path.exists(tps_b200.library_path()):
        pytest.skip("libtpsb200.so (host meshkit) not built")
    return tps_b200.lib()


@pytest.mark.parametrize("order,bt,ir", [(2, 0, 0), (3, 0, 0), (2, 1, 1)])
def test_fluid_at_rest_is_steady_in_a_closed_axisymmetric_vessel(lib_cpu, oracle_built, order, bt, ir):
    """Uniform p, rho, u = 0 between inviscid walls: d(r p)/dr / r - p / r = 0, so the r-weighted weak form, the
    r-weighted wall fluxes, Me_inv_rad and the p / r term of AxisymmetricSource must cancel -- to round-off when
    nodes and rule are collocated (Gauss-Legendre: Me_inv_rad int phi_j p = p / r_j exactly, the nodal source), to
    the accuracy of the r-weighted projection of 1/r otherwise (the reference adds the source as a NODAL value after
    Me^-1, src/rhs_operator.cpp:393-399)."""
    m = ac.box(n=(4, 3), warp=0.07)
    _, orc = ac.make_pair(m, order, 1, bt, ir, 3, "inviscid", False, gpu=False)
    N = orc.N
    U = np.concatenate([np.full(N, 1.2), np.zeros(3 * N), np.full(N, 101300.0 / 0.4)])
    y = orc.mult(U)
    scale = 101300.0 / 0.5  # p / r
    tol = 1e-10 if (bt, ir) == (0, 0) else 2e-2
    assert np.abs(y[N:2 * N]).max() < tol * scale
    assert np.abs(y).max() < tol * scale


def test_rigid_rotation_balances_the_centrifugal_term(lib_cpu, oracle_built):
    """u_theta = Omega r with dp/dr = rho Omega^2 r (p = p0 + rho Omega^2 r^2 / 2, quadratic: exactly representable at
    p >= 2): inviscid steady state, so the r-momentum residual vanishes to round-off (Euler, inviscid walls)."""
    m = ac.box(n=(4, 3))
    _, orc = ac.make_pair(m, 3, 0, 0, 0, 3, "inviscid", False, gpu=False)
    N = orc.N
    r = orc.node_coords()[:, 0]
    rho, Om = 1.2, 40.0
    ut = Om * r
    p = 101300.0 + 0.5 * rho * Om * Om * r * r
    U = np.concatenate([np.full(N, rho), np.zeros(N), np.zeros(N), rho * ut, p / 0.4 + 0.5 * rho * ut * ut])
    y = orc.mult(U)
    # E = p/0.4 + rho ut^2/2 is quadratic in r as well; the Rusanov flux sees no jumps
    assert np.abs(y[N:2 * N]).max() < 1e-9 * (101300.0 / r.min())
    assert np.abs(y[3 * N:4 * N]).max() < 1e-9 * (101300.0 / r.min())


@needs_ref
@pytest.mark.parametrize("nvel,bc,ubg", [(2, "c4", True), (2, "adiabatic", False), (3, "c4", True), (3, "adiabatic", False),
                                         (3, None, False)])
def test_port_physics_equals_reference_object_code_2d_and_axisymmetric(lib_cpu, oracle_built, nvel, bc, ubg):
    m = ac.box(n=(4, 3), warp=0.05)
    _, a = ac.make_pair(m, 2, 1, 0, 0, nvel, bc, ubg, gpu=False, kind="port")
    _, b = ac.make_pair(m, 2, 1, 0, 0, nvel, bc, ubg, gpu=False, kind="ref")
    U = ac.dry_state(a.node_coords(), nvel)
    ya, ga = a.mult(U, want_grad=True)
    yb, gb = b.mult(U, want_grad=True)
    assert rel_l2(ga, gb) < 1e-13
    N = a.N
    for k in range(nvel + 2):
        assert rel_l2(ya[k * N:(k + 1) * N], yb[k * N:(k + 1) * N]) < 1e-12, k


@needs_ref
@pytest.mark.parametrize("bc", [None, "inviscid", "c4"])
def test_roe_port_equals_reference_object_code(lib_cpu, oracle_built, bc):
    """flow/useRoe = 1: the restated Eval_Roe (2-D, gamma - 1 = 0.4 hard-coded) against the reference's own
    RiemannSolverTPS::Eval_Roe; inviscid walls use it too, the other boundary conditions force Lax-Friedrichs."""
    import tps_b200
    m = ac.box(n=(4, 3), warp=0.05) if bc else tps_b200.cartesian_quad_mesh(4, 3, lo=(-np.pi, -np.pi), hi=(np.pi, np.pi))
    out = []
    for kind in ("port", "ref"):
        orc = oracle_api.Oracle(2, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                                phys=oracle_api.dry_air_params(1, 3e4, 0.2, use_roe=True), kind=kind, basis_type=1, int_rule=1)
        if bc:
            orc.set_bcs(m["face_attr"], [oracle_api.make_bc(*b) for b in ac.bcs(bc)], True)
        U = ac.dry_state(orc.node_coords(), 2)
        out.append(orc.mult(U))
    lf = oracle_api.Oracle(2, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                           phys=oracle_api.dry_air_params(1, 3e4, 0.2), kind="port", basis_type=1, int_rule=1)
    if bc:
        lf.set_bcs(m["face_attr"], [oracle_api.make_bc(*b) for b in ac.bcs(bc)], True)
    assert rel_l2(out[0], out[1]) < 1e-12
    assert rel_l2(out[0], lf.mult(U)) > 1e-4  # and it really is a different solver than Lax-Friedrichs
