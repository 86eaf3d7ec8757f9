"""test/argon_minimal.binary.test on the device: 1000 RK4 steps of the Ar / Ar+ diffusion wave through tpsb_ode_step
land on the reference's analytic solution within the reference's own tolerance (2e-4 relative on rho Y_Ar+), and on
the oracle's trajectory to 1e-8 (BASELINE.json's bound after many steps)."""
import os

import numpy as np
import pytest

import binary_mixture_case as bm
import oracle_api
import tps_b200
from common import rel_l2

pytestmark = pytest.mark.gpu
REF_SO = os.path.join(oracle_api.ORACLE_DIR, "_ref", "liboracle_ref.so")


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built")
def test_device_reproduces_the_binary_diffusion_benchmark(lib_built, oracle_built):
    import torch
    m = bm.mesh()
    models = tps_b200.PlasmaModels.from_dict(bm.models_dict())
    op = tps_b200.RhsOperator(m, order=3, physics=tps_b200.Physics.plasma_mixture(models), basis_type=1, int_rule_type=1,
                              nvel=2)
    orc = oracle_api.Oracle(3, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.mixture_params(models), kind="ref", basis_type=1, int_rule=1, neq=6, nvel=2)
    xy = orc.node_coords()
    U0 = bm.state(xy)
    N = orc.N
    D = orc.mixture_average_diffusivity(U0[0::N], 3)
    ref, decay = bm.analytic(xy, D[0])
    x = torch.from_numpy(U0.copy()).cuda()
    op.ode_step(x, bm.DT, scheme=4, nsteps=bm.NSTEPS)
    got = x.cpu().numpy()
    rel = np.abs(got[4 * N:5 * N] - ref[4 * N:5 * N]) / np.abs(ref[4 * N:5 * N])
    assert rel.max() < bm.TOL, rel.max()
    traj = orc.rk4(U0, bm.DT, bm.NSTEPS)
    for k in (0, 1, 3, 4):
        assert rel_l2(got[k * N:(k + 1) * N], traj[k * N:(k + 1) * N]) < 1e-8, k
