"""Device path against COMMITTED golden du/dt vectors (tests/golden/rhs_golden.npz), produced in the build container by
the oracle running the reference's own physics object code (tests/golden/make_rhs_golden.py).  Nothing under oracle/ is
imported or executed here: this parity check stands on the fixture alone, so it also runs where oracle/_ref is absent."""
import os

import numpy as np
import pytest

import golden_cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "rhs_golden.npz")


@pytest.mark.parametrize("name", golden_cases.CASES)
def test_device_rhs_matches_committed_reference_physics_result(lib_built, name):
    import torch
    g = np.load(GOLD)
    _, _, op, _ = golden_cases.build(name, gpu=True)
    U, y_ref = g[name + "/U"], g[name + "/y"]
    assert U.size == op.N * op.neq
    if name + "/dist" in g:
        dist = torch.from_numpy(g[name + "/dist"]).cuda()
        op.set_distance_field(dist)
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    N = op.N
    for k in range(op.neq):
        ref = y_ref[k * N:(k + 1) * N]
        assert np.linalg.norm(y[k * N:(k + 1) * N] - ref) <= 1e-10 * max(np.linalg.norm(ref), 1e-30) + 1e-9, (name, k)
