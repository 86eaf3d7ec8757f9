"""Parity at the sizes BASELINE.md section 4 names: per-evaluation du/dt at n = 12, the state
after 100 RK4 steps at n = 6 (and on a rotated, boundary-free general-path mesh), and config C1 at its full 25 600
quadrilaterals.  Tolerances are BASELINE.json's: 1e-10 per evaluation, 1e-8 after 100 steps."""
import numpy as np
import pytest

import oracle_api
import tps_b200
from common import rel_l2, rotate_elements, tgv_state, warp_mesh

pytestmark = pytest.mark.gpu
PI = np.pi


def _box(n, visc=1e3, bulk=0.0, warp=0.0, rotate=False):
    m = tps_b200.cartesian_hex_mesh(n, n, n, lo=(-PI,) * 3, hi=(PI,) * 3)
    if warp:
        m = warp_mesh(m, amp=warp)
    if rotate:
        m = rotate_elements(m)
    op = tps_b200.RhsOperator(m, order=3, physics=tps_b200.Physics.dry_air(1, visc, bulk))
    orc = oracle_api.Oracle(3, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(1, visc, bulk))
    return m, op, orc, tgv_state(orc.node_coords())


@pytest.mark.parametrize("warp", [0.0, 0.1])
def test_per_evaluation_parity_n12(lib_built, oracle_built, warp):
    import torch
    m, op, orc, U = _box(12, visc=2e4, bulk=0.3, warp=warp)
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo = orc.mult(U)
    N = orc.N
    for k in range(5):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k
    assert abs(op.max_char_speed() / orc.max_char_speed - 1) < 1e-13


@pytest.mark.parametrize("warp,rotate", [(0.0, False), (0.08, True)])
def test_hundred_rk4_steps_n6(lib_built, oracle_built, warp, rotate):
    import torch
    m, op, orc, U = _box(6, warp=warp, rotate=rotate)
    orc.mult(U)
    h = 2 * PI / 6
    dt = 0.3 * (h / 3) / orc.max_char_speed / 3  # CFL-like, cf. src/M2ulPhyS.cpp:2014
    x = torch.from_numpy(U.copy()).cuda()
    op.ode_step(x, dt, scheme=4, nsteps=100)
    ref = orc.rk4(U, dt, 100)
    got = x.cpu().numpy()
    N = orc.N
    assert np.isfinite(got).all() and rel_l2(ref, U) > 1e-5
    for k in range(5):
        assert rel_l2(got[k * N:(k + 1) * N], ref[k * N:(k + 1) * N]) < 1e-8, k


def test_c1_full_size(lib_built, oracle_built):
    """Config C1 (mms.euler_2d) at the reference size: 160 x 160 = 25 600 quadrilaterals, p = 2, GLL / GLL, Euler."""
    import torch
    m = tps_b200.cartesian_quad_mesh(160, 160, lo=(0, 0), hi=(3.02, 3.02))
    op = tps_b200.RhsOperator(m, order=2, physics=tps_b200.Physics.dry_air(0), basis_type=1, int_rule_type=1)
    orc = oracle_api.Oracle(2, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(0), basis_type=1, int_rule=1)
    xy = orc.node_coords() * (2 * PI / 3.02)
    rho = 1.0 + 0.2 * np.sin(xy[:, 0]) * np.cos(xy[:, 1])
    u, v = 30.0 + 5.0 * np.cos(xy[:, 0]), -10.0 + 4.0 * np.sin(xy[:, 1] + 0.3)
    p = 101300.0 * (1.0 + 0.05 * np.cos(xy[:, 0] - xy[:, 1]))
    rng = np.random.default_rng(20261018)
    U = np.concatenate([rho, rho * u, rho * v, p / 0.4 + 0.5 * rho * (u * u + v * v)])
    U = np.ascontiguousarray(U * (1.0 + 0.01 * rng.uniform(-1, 1, U.shape)))
    N = orc.N
    assert N == 230400
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo = orc.mult(U)
    for k in range(4):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k
