"""Averaging::addSample (src/averaging.cpp:198-420) on the device against a line-by-line numpy restatement: running means of
the primitive family with the pressure in the temperature slot (the GasMixture overload M2ulPhyS uses, src/M2ulPhyS.cpp:634)
and the variances / covariances of the velocity components about the updated mean."""
import numpy as np
import pytest

import tps_b200
from common import node_coords_from_mesh, rel_l2, tgv_state

pytestmark = pytest.mark.gpu
PI = np.pi


def _ref_add_sample(inst, mean, vari, ns_mean, ns_vari, dim, vstart, vcomp, pressure):
    """averaging.cpp:330-420, node by node but vectorised over the nodes."""
    nf = inst.shape[0]
    new = mean.copy()
    for eq in range(nf):
        v = pressure if (pressure is not None and eq == 1 + dim) else inst[eq]
        new[eq] = (ns_mean * mean[eq] + v) / (ns_mean + 1)
    nv = None
    if vari is not None:
        nv = vari.copy()
        vi = 0
        for i in range(vstart, vstart + vcomp):
            d = inst[i] - new[i]
            nv[vi] = (vari[vi] * ns_vari + d * d) / (ns_vari + 1)
            vi += 1
        for i in range(vstart, vstart + vcomp - 1):
            for j in range(i + 1, vstart + vcomp):
                nv[vi] = (vari[vi] * ns_vari + (inst[i] - new[i]) * (inst[j] - new[j])) / (ns_vari + 1)
                vi += 1
    return new, nv


def test_running_mean_and_covariances_of_the_primitive_state(lib_built):
    import torch
    m = tps_b200.cartesian_hex_mesh(4, 3, 3, lo=(-PI,) * 3, hi=(PI,) * 3)
    op = tps_b200.RhsOperator(m, order=3, physics=tps_b200.Physics.dry_air(1, 1e3))
    N = op.N
    xyz = node_coords_from_mesh(m["elem_xyz"], 3)
    mean_d = torch.zeros(5 * N, dtype=torch.float64, device="cuda")
    vari_d = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
    mean, vari = np.zeros((5, N)), np.zeros((6, N))
    for s in range(4):
        U = tgv_state(xyz, perturb=0.05, seed=100 + s)
        x = torch.from_numpy(U).cuda()
        op.Mult(x)                     # an evaluation in between: the fused path keeps Up on chip ...
        op.averaging_add_sample(mean_d, vari_d, s, s)   # ... and the averaging asks for it
        r = U.reshape(5, N)
        up = np.stack([r[0], r[1] / r[0], r[2] / r[0], r[3] / r[0],
                       0.4 / 287.058 * (r[4] - 0.5 * (r[1] ** 2 + r[2] ** 2 + r[3] ** 2) / r[0]) / r[0]])
        mean, vari = _ref_add_sample(up, mean, vari, s, s, 3, 1, 3, up[0] * 287.058 * up[4])
        assert rel_l2(mean_d.cpu().numpy(), mean.reshape(-1)) < 1e-13
        assert rel_l2(vari_d.cpu().numpy(), vari.reshape(-1)) < 1e-11
    # a plain family without the pressure slot and without variances (addSampleInternal(), :250-328)
    fam = torch.rand(3 * N, dtype=torch.float64, device="cuda")
    fm = torch.zeros_like(fam)
    op.averaging_add_sample(fm, None, 0, 0, inst=fam, pressure_slot=False)
    op.averaging_add_sample(fm, None, 1, 1, inst=2 * fam, pressure_slot=False)
    assert rel_l2(fm.cpu().numpy(), 1.5 * fam.cpu().numpy()) < 1e-15


def test_averaging_on_the_generic_path_uses_the_mixture_pressure(lib_built, oracle_built):
    import torch
    import axisym_cases as ac
    import oracle_api
    m = ac.box(n=(5, 4), warp=0.03)
    op, orc = ac.make_pair(m, 2, 1, 1, 1, 3, None, False, mixture=ac.argon6_dict())
    up = ac.argon6_primitives(orc.node_coords(), 3)         # [N, neq]
    U = np.ascontiguousarray(orc.pt("cons", up).T.reshape(-1))
    N, neq = orc.N, orc.neq
    x = torch.from_numpy(U).cuda()
    op.updatePrimitives(x)
    mean_d = torch.zeros(neq * N, dtype=torch.float64, device="cuda")
    vari_d = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
    op.averaging_add_sample(mean_d, vari_d, 0, 0, vari_components=3)
    got = mean_d.cpu().numpy().reshape(neq, N)
    upd = op.fields()[0].cpu().numpy().reshape(neq, N)
    # mixture pressure of the primitive state: sum of the partial pressures (PerfectMixture::ComputePressureFromPrimitives);
    # the six-species state [rho, u(3), T_h, n_sp(5), T_e]: p = sum_heavy n R T_h + n_e R T_e with n_Ar by difference
    R = 8.3144598
    d = ac.argon6_dict()["species"]
    n_act = upd[5:10]
    rho_act = sum(n_act[i] * d[i]["mw"] for i in range(5))
    n_bg = (upd[0] - rho_act) / d[5]["mw"]
    p = (n_act[0] + n_act[1] + n_act[2] + n_act[3] + n_bg) * R * upd[4] + n_act[4] * R * upd[10]
    assert rel_l2(got[3], p) < 1e-12      # slot 1 + dim = 3 (dim = 2: the reference indexes with the MESH dimension)
    for eq in (0, 1, 2, 4, 5, 10):
        assert np.array_equal(got[eq], upd[eq])
