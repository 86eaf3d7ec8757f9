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
