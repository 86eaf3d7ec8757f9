"""The C++ host shim (tps_b200/csrc/host/tpsb_host.hpp: RHSoperator / ODESolver classes with the reference's method
names over the C ABI) through its compute_rhs probe (the analogue of the reference's utils/compute_rhs.cpp): the
stand-alone binary must reproduce what the ctypes binding computes on the same mesh and state."""
import os
import re
import subprocess

import numpy as np
import pytest

import tps_b200
from common import node_coords_from_mesh

pytestmark = pytest.mark.gpu
PI = np.pi
EXE = os.path.join(os.path.dirname(tps_b200.library_path()), "compute_rhs")


@pytest.mark.parametrize("n,steps", [(4, 2), (6, 1)])
def test_compute_rhs_probe_matches_ctypes_path(lib_built, n, steps):
    import torch
    if not os.path.exists(EXE):
        pytest.skip("compute_rhs not built")
    out = subprocess.run([EXE, str(n), str(steps)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    l2 = [float(v) for v in re.findall(r"rhs_l2\[\d\] (\S+)", out.stdout)]
    mcs = float(re.search(r"max_char_speed (\S+)", out.stdout).group(1))
    total = float(re.search(r"sum\(U\) (\S+)", out.stdout).group(1))
    assert len(l2) == 5
    # the same problem through the ctypes binding: TGV state without perturbation, visc_mult 1420, lexicographic order
    m = tps_b200.cartesian_hex_mesh(n, n, n, lo=(-PI,) * 3, hi=(PI,) * 3)
    op = tps_b200.RhsOperator(m, order=3, physics=tps_b200.Physics.dry_air(1, 1420.0, 0.0))
    xyz = node_coords_from_mesh(m["elem_xyz"], 3)
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    rho0, p0, g = 1.2, 101300.0, 1.4
    V0 = 0.1 * np.sqrt(g * p0 / rho0)
    u, v = V0 * np.sin(x) * np.cos(y) * np.cos(z), -V0 * np.cos(x) * np.sin(y) * np.cos(z)
    p = p0 + rho0 * V0 * V0 / 16.0 * (np.cos(2 * x) + np.cos(2 * y)) * (np.cos(2 * z) + 2.0)
    U = np.ascontiguousarray(np.concatenate([np.full_like(x, rho0), rho0 * u, rho0 * v, 0 * x,
                                             p / (g - 1) + 0.5 * rho0 * (u * u + v * v)]))
    xd = torch.from_numpy(U).cuda()
    r = op.Mult(xd).cpu().numpy()
    N = op.N
    ref = [float(np.sqrt((r[k * N:(k + 1) * N] ** 2).mean())) for k in range(5)]
    for k in (1, 2, 3, 4):
        assert abs(l2[k] / ref[k] - 1) < 1e-9, (k, l2[k], ref[k])
    assert abs(l2[0] - ref[0]) < 1e-9 * ref[1]   # d(rho)/dt of the solenoidal field is a cancellation residue
    assert abs(mcs / op.max_char_speed() - 1) < 1e-13
    op.ode_step(xd, 1e-5, scheme=4, nsteps=steps)
    assert abs(total / float(xd.sum().item()) - 1) < 1e-12
