"""Golden du/dt vectors produced by the oracle with the REFERENCE'S OWN physics object code (oracle/_ref, compiled from
/root/reference/src) on small seeded cases; committed as tests/golden/rhs_golden.npz so that the device path (and the
oracle's dry-air port) can be checked where /root/reference and oracle/_ref are absent.  The input states
(and the wall-distance field of the mixing-length case) are stored next to the results, so the device test needs no
oracle at all; meshes and model parameters come from the deterministic builders in tests/golden_cases.py.
Run in the build container only:  python tests/golden/make_rhs_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import golden_cases  # noqa: E402

out = {}
for name in golden_cases.CASES:
    orc, U, _, extra = golden_cases.build(name, gpu=False)
    y = orc.mult(U)
    out[name + "/y"] = y
    out[name + "/U"] = U
    for k, v in extra.items():
        out[name + "/" + k] = v
    print(f"{name:28s} N = {orc.N:6d}  |y| = {np.linalg.norm(y):.6e}")
np.savez_compressed(os.path.join(HERE, "rhs_golden.npz"), **out)
print("wrote rhs_golden.npz")
