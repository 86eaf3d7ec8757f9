"""Pin of the oracle's DG operator layer (L3/L4) on a reference-held end-to-end value: the CPU oracle, integrating the
Ar / Ar+ diffusion wave of test/argon_minimal.binary.test with the reference's own transport object code, lands on the
reference's analytic solution within the reference's own tolerance (2e-4 relative on rho Y_Ar+)."""
import os

import numpy as np
import pytest

import binary_mixture_case as bm
import oracle_api
import tps_b200

REF_SO = os.path.join(oracle_api.ORACLE_DIR, "_ref", "liboracle_ref.so")


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_reproduces_the_binary_diffusion_benchmark(lib_built, oracle_built):
    m = bm.mesh()
    models = tps_b200.PlasmaModels.from_dict(bm.models_dict())
    orc = oracle_api.Oracle(3, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.mixture_params(models), kind="ref", basis_type=1, int_rule=1, neq=6, nvel=2)
    xy = orc.node_coords()
    U0 = bm.state(xy)
    N = orc.N
    D = orc.mixture_average_diffusivity(U0[0::N], 3)      # node 0, as binary_mixture_ic.cpp:121-131
    ref, decay = bm.analytic(xy, D[0])
    assert 0.05 < decay < 0.95                              # the wave really decays: the test is sensitive to D_ia
    got = orc.rk4(U0, bm.DT, bm.NSTEPS)
    rel = np.abs(got[4 * N:5 * N] - ref[4 * N:5 * N]) / np.abs(ref[4 * N:5 * N])
    assert rel.max() < bm.TOL, rel.max()
    # without diffusion (pure advection of the initial wave) the benchmark is missed by far: the check has teeth
    adv, _ = bm.analytic(xy, 0.0)
    assert (np.abs(adv[4 * N:5 * N] - ref[4 * N:5 * N]) / np.abs(ref[4 * N:5 * N])).max() > 50 * bm.TOL
