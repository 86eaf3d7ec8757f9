"""The oracle itself: physics pinned on the reference's object code, operators on MMS convergence."""
import ctypes as C
import os

import numpy as np
import pytest

import meshref
import mms
import oracle_api
from common import rel_l2

HAVE_REF = os.path.exists(os.path.join(oracle_api.ORACLE_DIR, "_ref", "liboracle_ref.so")) or os.path.isdir("/root/reference/src")


def _states(n, seed=7):
    rng = np.random.default_rng(seed)
    rho = rng.uniform(0.3, 3.0, n)
    vel = rng.uniform(-300, 300, (n, 3))
    T = rng.uniform(150, 2500, n)
    E = rho * 287.058 * T / 0.4 + 0.5 * rho * (vel ** 2).sum(1)
    return np.ascontiguousarray(np.column_stack([rho, rho[:, None] * vel, E]))


@pytest.mark.parametrize("kind", ["port"] + (["ref"] if HAVE_REF else []))
def test_known_answer_vector_from_reference_classes(oracle_built, kind):
    """values produced by the reference's own DryAir+DryAirTransport+Fluxes+RiemannSolverTPS."""
    lib = oracle_api.load(kind)
    ph = oracle_api.dry_air_params(1)
    lib.orc_phys_init(C.byref(ph))
    u1 = np.array([[1.2, 24, 1, -2, 253312.5]])
    u2 = np.array([[1.1, 20, 0.5, 1, 250000.0]])
    nor = np.array([[0.3, -0.2, 0.1]])
    F = np.zeros((1, 5))
    lib.orc_phys_riemann(1, u1, u2, nor, F)
    gold = np.array([13.412484842853365, 30576.326666441397, -20076.261818209976, 9844.446666835609, 2191166.9369117124])
    assert np.abs(F[0] / gold - 1).max() < 1e-14
    up = np.zeros((1, 5))
    lib.orc_phys_prim(1, u1, up)
    assert abs(up[0, 4] / 293.86676405310266 - 1) < 1e-15
    lam = np.zeros(1)
    lib.orc_phys_max_char_speed(1, u1, lam)
    assert abs(lam[0] / 363.74273648180633 - 1) < 1e-15


@pytest.mark.skipif(not HAVE_REF, reason="reference object code not built (oracle/_ref)")
@pytest.mark.parametrize("eq,vm,bm", [(1, 1.0, 0.0), (1, 37.5, 0.6), (0, 1.0, 0.0)])
def test_port_physics_equals_reference_object_code(oracle_built, eq, vm, bm):
    n = 2000
    U1, U2 = _states(n, 1), _states(n, 2)
    rng = np.random.default_rng(3)
    nor = np.ascontiguousarray(rng.normal(size=(n, 3)))
    G = np.ascontiguousarray(rng.normal(size=(n, 15)) * 50)
    out = {}
    for kind in ("port", "ref"):
        lib = oracle_api.load(kind)
        ph = oracle_api.dry_air_params(eq, vm, bm)
        lib.orc_phys_init(C.byref(ph))
        up, lam, fc, fv, fr = np.zeros((n, 5)), np.zeros(n), np.zeros((n, 15)), np.zeros((n, 15)), np.zeros((n, 5))
        lib.orc_phys_prim(n, U1, up)
        lib.orc_phys_max_char_speed(n, U1, lam)
        lib.orc_phys_conv_flux(n, U1, fc)
        lib.orc_phys_visc_flux(n, U1, G, fv)
        lib.orc_phys_riemann(n, U1, U2, nor, fr)
        out[kind] = (up, lam, fc, fv, fr)
    for a, b in zip(out["port"], out["ref"]):
        assert np.array_equal(a, b)  # same arithmetic, same compiler: bit-identical


def _mms_errors(n, ph, kind="port"):
    ev, xyz = meshref.cartesian_hex(n, n, n, lo=(-np.pi,) * 3, hi=(np.pi,) * 3)
    el1, el2, i1, i2 = meshref.build_faces(ev)
    o = oracle_api.Oracle(3, xyz, el1, el2, i1, i2, phys=ph, kind=kind)
    U, R, G = mms.manufactured(o.node_coords(), ph)
    Y, Gh = o.mult(U, want_grad=True)
    N = o.N
    return np.array([rel_l2(Y[k * N:(k + 1) * N], R[k * N:(k + 1) * N]) for k in range(5)] + [rel_l2(Gh, G)])


@pytest.mark.parametrize("eq,vm", [(0, 1.0), (1, 3e6)])
def test_oracle_converges_to_exact_rhs(oracle_built, eq, vm):
    """Role of the reference's MMS tests (test/mms.euler.test: rates ~p+1 on the solution, i.e. ~p on the
    RHS): the restated operator converges to the exact -div(F_c - F_v) of a manufactured state."""
    ph = oracle_api.dry_air_params(eq, vm, 0.7)
    e6, e12 = _mms_errors(6, ph), _mms_errors(12, ph)
    rates = np.log2(e6 / e12)
    assert (e12 < 2e-2).all()
    assert (rates[:5] > 1.7).all() and rates[5] > 2.3


def test_uniform_state_has_zero_rhs_and_gradient(oracle_built):
    ev, xyz = meshref.cartesian_hex(3, 4, 3)
    o = oracle_api.Oracle(3, xyz, *meshref.build_faces(ev))
    N = o.N
    U = np.concatenate([np.full(N, v) for v in (1.2, 12.0, -3.0, 5.0, 253000.0)])
    Y, G = o.mult(U, want_grad=True)
    assert np.abs(G).max() < 1e-9
    assert np.abs(Y[:N]).max() < 1e-9 and np.abs(Y[4 * N:]).max() / 253000.0 < 1e-9
