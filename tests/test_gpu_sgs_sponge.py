"""Sub-grid-scale eddy viscosity (Fluxes::sgsSmag / sgsSigma) and the planar viscous sponge (Fluxes::viscSpongePlanar)
inside the viscous fluxes of the 3-D dry-air path, against the oracle running the reference's own Fluxes object code
(src/fluxes.cpp:224-246, 386-407, 513-684).  Reference runs: test/inputs/input.sgsSmag.ini, input.sgsSigma.ini
(periodic box, order 1) and viscosityMultiplierFunction in the cylinder inputs."""
import os

import numpy as np
import pytest

import oracle_api
import tps_b200
from common import box_face_attrs, rel_l2, tgv_state, warp_mesh

pytestmark = pytest.mark.gpu
PI = np.pi
HAVE_REF = os.path.exists(os.path.join(oracle_api.ORACLE_DIR, "_ref", "liboracle_ref.so"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref (reference object code) not built")
SPONGE = ((0.3, 1.0, -0.2), (0.4, 0.1, 0.2), 7.5, 0.8)  # normal (not unit: Fluxes normalises it), point, ratio, width


def _periodic(order, n, warp, sgs, sponge, visc_mult=50.0):
    import torch
    m = tps_b200.cartesian_hex_mesh(*n, lo=(-PI,) * 3, hi=(PI,) * 3)
    if warp:
        m = warp_mesh(m, amp=0.08, lo=(-PI,) * 3, hi=(PI,) * 3)
    op = tps_b200.RhsOperator(m, order=order, physics=tps_b200.Physics.dry_air(1, visc_mult, 0.3, sgs=sgs, sponge=sponge))
    orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(1, visc_mult, 0.3, sgs=sgs, sponge=sponge), kind="ref")
    U = tgv_state(orc.node_coords())
    return torch, op, orc, U


def _check(torch, op, orc, U, tol=1e-10):
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo = orc.mult(U)
    N = orc.N
    for k in range(5):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < tol, k
    return y, yo


@needs_ref
@pytest.mark.parametrize("warp", [False, True])
@pytest.mark.parametrize("order,n", [(1, (6, 5, 4)), (3, (4, 4, 3))])
def test_smagorinsky_parity(lib_built, oracle_built, order, n, warp):
    """flow/sgsModel = smagorinsky with the reference's default constant 0.12 and a floor below the element size."""
    torch, op, orc, U = _periodic(order, n, warp, sgs=(1, 0.12, 0.05), sponge=None)
    y, yo = _check(torch, op, orc, U)
    # the model is active: the same state without it gives a different momentum / energy residual
    torch2, op0, orc0, _ = _periodic(order, n, warp, sgs=None, sponge=None)
    y0 = op0.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    assert rel_l2(y[orc.N:], y0[orc.N:]) > 1e-6


@needs_ref
@pytest.mark.parametrize("warp", [False, True])
def test_sigma_model_parity(lib_built, oracle_built, warp):
    """flow/sgsModel = sigma (closed-form singular values, as a device build of the reference computes them).  The
    eigenvalue formula cancels (acos near +-1, sigma_1 - sigma_2): the eddy viscosity itself agrees to ~1e-9 relative,
    the residual to the usual 1e-10 because mu_sgs is a small part of the flux."""
    torch, op, orc, U = _periodic(3, (4, 3, 4), warp, sgs=(2, 0.135, 0.0), sponge=None)
    _check(torch, op, orc, U, tol=1e-10)


@needs_ref
@pytest.mark.parametrize("sgs", [None, (1, 0.12, 0.0)])
def test_viscous_sponge_parity(lib_built, oracle_built, sgs):
    torch, op, orc, U = _periodic(2, (4, 4, 4), True, sgs=sgs, sponge=SPONGE)
    y, _ = _check(torch, op, orc, U)
    _, op0, _, _ = _periodic(2, (4, 4, 4), True, sgs=sgs, sponge=None)
    y0 = op0.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    assert rel_l2(y[orc.N:], y0[orc.N:]) > 1e-6


@needs_ref
@pytest.mark.parametrize("use_bc_in_grad", [False, True])
def test_sgs_and_sponge_with_boundary_conditions(lib_built, oracle_built, use_bc_in_grad):
    """Inlet / outlet / every wall type with the modified transport in the interior and boundary viscous fluxes
    (ComputeViscousFluxes and ComputeBdrViscousFluxes both carry the SGS / sponge block)."""
    import torch
    lo, hi = (0.0, 0.0, 0.0), (2.0, 1.2, 1.0)
    specs = [(1, 0, 2, (1.2, 25.0, 1.0, -2.0)), (2, 1, 0, (101300.0,)), (3, 2, 3, (310.0,)), (4, 2, 2, ()), (5, 2, 0, ()),
             (6, 2, 3, (290.0,))]
    m = tps_b200.cartesian_hex_mesh(4, 3, 3, lo=lo, hi=hi, periodic=(0, 0, 0))
    attr = box_face_attrs(m, lo, hi)
    m = warp_mesh(m, amp=0.1, lo=lo, hi=hi)
    sgs, sponge = (1, 0.12, 0.01), ((1.0, 0.0, 0.0), (1.5, 0.0, 0.0), 5.0, 0.4)
    op = tps_b200.RhsOperator(m, order=3, physics=tps_b200.Physics.dry_air(1, 4e3, 0.2, sgs=sgs, sponge=sponge),
                              face_attr=attr, use_bc_in_grad=use_bc_in_grad, bcs=[tps_b200.BcDesc.make(*b) for b in specs])
    orc = oracle_api.Oracle(3, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(1, 4e3, 0.2, sgs=sgs, sponge=sponge), kind="ref")
    orc.set_bcs(attr, [oracle_api.make_bc(*b) for b in specs], use_bc_in_grad)
    U = tgv_state(orc.node_coords() * PI)
    _check(torch, op, orc, U)


def test_sgs_rejected_where_not_built(lib_built):
    """2-D / Gauss-Lobatto / mixture runs go through the generic path, which does not carry the SGS block yet: the
    library must say so instead of silently dropping the model."""
    m = tps_b200.cartesian_quad_mesh(3, 3)
    with pytest.raises(tps_b200.TpsbError):
        tps_b200.RhsOperator(m, order=2, physics=tps_b200.Physics.dry_air(1, 1.0, sgs=(1, 0.12, 0.0)), basis_type=1, int_rule_type=1)
