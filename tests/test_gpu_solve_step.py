"""tpsb_solve_step = M2ulPhyS::solveStep without the I/O (src/M2ulPhyS.cpp:2004-2016): one ODE step, Check_NAN,
Check_Undershoot for mixtures, adaptive time step CFL hmin / max_char_speed / dim."""
import numpy as np
import pytest

import axisym_cases as ac
import oracle_api
import tps_b200
from common import node_coords_from_mesh, tgv_state, warp_mesh

pytestmark = pytest.mark.gpu
PI = np.pi


def test_step_nan_count_and_adaptive_dt(lib_built, oracle_built):
    import torch
    m = warp_mesh(tps_b200.cartesian_hex_mesh(4, 3, 3, lo=(-PI,) * 3, hi=(PI,) * 3), amp=0.08)
    op = tps_b200.RhsOperator(m, order=3, physics=tps_b200.Physics.dry_air(1, 2e3, 0.1))
    orc = oracle_api.Oracle(3, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(1, 2e3, 0.1))
    delta = np.zeros(orc.NE)
    orc.lib.orc_elem_size(orc.h, delta)
    hmin = 3 * delta.min()  # GetElementSize(e, 1) = order * delta
    assert abs(op.hmin() / hmin - 1) < 1e-12
    U = tgv_state(orc.node_coords())
    x = torch.from_numpy(U.copy()).cuda()
    x_ref = x.clone()
    dt, cfl = 2e-6, 0.12
    nan, dt_next = op.solve_step(x, dt, scheme=4, cfl=cfl)
    op.ode_step(x_ref, dt, scheme=4, nsteps=1)
    assert nan == 0 and torch.equal(x, x_ref)
    # the characteristic speed is the one of the step's last stage, as in the reference (max_char_speed of the last Mult)
    assert abs(dt_next / (cfl * hmin / op.max_char_speed() / 3.0) - 1) < 1e-14
    ref = orc.rk4(U, dt, 1)
    assert np.abs(x.cpu().numpy() - ref).max() <= 1e-10 * np.abs(ref).max()
    # ... and that speed is the oracle's for the last stage vector y4 = x + dt k3 (RK4Solver)
    f = lambda v: orc.mult(np.ascontiguousarray(v))
    k1 = f(U)
    k2 = f(U + 0.5 * dt * k1)
    k3 = f(U + 0.5 * dt * k2)
    f(U + dt * k3)
    assert abs(dt_next / (cfl * hmin / orc.max_char_speed / 3.0) - 1) < 1e-12
    # constant time step: dt comes back unchanged; NaNs are counted
    x[7] = float("nan")
    x[op.N + 11] = float("nan")
    nan, dt_next = op.solve_step(x, dt, scheme=1, cfl=0.0)
    assert nan >= 2 and dt_next == dt


def test_quadrilateral_hmin(lib_built, oracle_built):
    m = ac.box(n=(5, 4), warp=0.07)
    op, orc = ac.make_pair(m, 2, 0, 1, 1, 2, "inviscid", False)
    delta = np.zeros(orc.NE)
    orc.lib.orc_elem_size(orc.h, delta)
    assert abs(op.hmin() / (2 * delta.min()) - 1) < 1e-12


def test_undershoot_clamp_on_mixtures_only(lib_built):
    """Check_Undershoot: negative active-species densities are set to zero after the step (user-defined fluids)."""
    import os
    import torch
    import golden_cases
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rhs_golden.npz"))
    _, _, op, _ = golden_cases.build("ternary2d_p2", gpu=True)
    N = op.N
    assert op.neq == 6
    U0 = g["ternary2d_p2/U"].copy()
    U0[4 * N + 3] = -1e-9  # an undershoot of the active species (equation nvel + 2 = 4) as a time step can leave it
    U = torch.from_numpy(U0).cuda()
    assert op.check_state(U) == 0
    out = U.cpu().numpy()
    assert out[4 * N + 3] == 0.0
    keep = np.ones(6 * N, bool)
    keep[4 * N + 3] = False
    assert np.array_equal(out[keep], U0[keep])
