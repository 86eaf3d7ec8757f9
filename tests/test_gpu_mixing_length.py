"""flow/useMixingLength: MixingLengthTransport (src/mixing_length_transport.cpp:62-121) wrapped around the molecular
transport, as test/inputs/plasma.ini (BASELINE config C4) runs it -- eddy viscosity rho l^2 |S| with l = min(0.41 d_wall,
l_max) from a nodal wall-distance field, read at the nodes by GetFlux and interpolated to the face points by each side
(src/rhs_operator.cpp:534-537, src/face_integrator.cpp:304-309, src/BCintegrator.cpp:408-411).  Oracle: the reference's
own MixingLengthTransport object code around its DryAirTransport / ConstantTransport / GasMixtureTransport."""
import os

import numpy as np
import pytest

import axisym_cases as ac
import oracle_api
from common import rel_l2

pytestmark = pytest.mark.gpu
HAVE_REF = os.path.exists(os.path.join(oracle_api.ORACLE_DIR, "_ref", "liboracle_ref.so"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref (reference object code) not built")


def _wall_distance(xy):
    """A smooth positive stand-in for the distance solver's output (the reference reads it as a given grid function)."""
    x, y = xy[:, 0], xy[:, 1]
    return 0.02 + 0.015 * np.sin(1.3 * x + 0.2) ** 2 + 0.03 * (y - y.min()) / (y.max() - y.min() + 1e-30)


def _run(op, orc, U, dist, tol=1e-10):
    import torch
    d = torch.from_numpy(dist).cuda()
    op.set_distance_field(d)
    orc.set_distance(dist)
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo = orc.mult(U)
    N = orc.N
    for k in range(op.neq):
        ref = yo[k * N:(k + 1) * N]
        assert np.linalg.norm(y[k * N:(k + 1) * N] - ref) <= tol * max(np.linalg.norm(ref), 1e-30) + 1e-9, k
    return y


@needs_ref
@pytest.mark.parametrize("nvel,bc,order,bt,ir", [(2, "c4", 3, 0, 0), (3, "c4", 3, 0, 0), (3, "inviscid", 2, 1, 1), (2, None, 2, 1, 1)])
def test_mixing_length_dry_air(lib_built, oracle_built, nvel, bc, order, bt, ir):
    """Planar and axisymmetric (swirl terms of |S|) dry air, with the wall set of config C4, inviscid walls (the only
    wall type that hands the distance to its viscous fluxes, src/wallBC.cpp:309-313) and a box without boundary integrators."""
    m = ac.box(warp=0.06)
    ml = (0.02, 0.9, 0.3)
    op, orc = ac.make_pair(m, order, 1, bt, ir, nvel, bc, bc == "c4", mixing_length=ml)
    U = ac.dry_state(orc.node_coords(), nvel)
    dist = _wall_distance(orc.node_coords())
    y = _run(op, orc, U, dist)
    # the model is active, and switches off with a zero distance field
    op.set_distance_field(None)
    import torch
    y0 = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    assert rel_l2(y[orc.N:], y0[orc.N:]) > 1e-7
    op1, _ = ac.make_pair(m, order, 1, bt, ir, nvel, bc, bc == "c4")
    assert rel_l2(y0, op1.Mult(torch.from_numpy(U).cuda()).cpu().numpy()) < 1e-13


@needs_ref
@pytest.mark.parametrize("transport", ["constant", "argon_mixture"])
def test_mixing_length_plasma_axisym_c4(lib_built, oracle_built, transport):
    """Config C4 as test/inputs/plasma.ini sets it up: six-species two-temperature argon, axisymmetric, mixing length
    with max-mixing-length 0.01 and Pr_ratio 0 (the ini's values) -- and with a non-zero Pr_ratio / bulk multiplier."""
    m = ac.box(n=(4, 3), warp=0.05)
    d = ac.argon6_dict()
    if transport != "constant":
        d.update(transport_model=transport, third_order_k_electron=False)
    for ml in ((0.01, 0.0, 0.0), (0.015, 0.85, 0.5)):
        op, orc = ac.make_pair(m, 3, 1, 0, 0, 3, "c4", True, mixture=d, mixing_length=ml)
        up = ac.argon6_primitives(orc.node_coords(), 3)
        if transport != "constant":
            up[:, 4] *= 8.0
            up[:, 10] *= 3.0
            up[:, 9] = up[:, 5]
        U = np.ascontiguousarray(orc.pt("cons", up).T).reshape(-1)
        _run(op, orc, U, _wall_distance(orc.node_coords()))
