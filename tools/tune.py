"""Development helper: per-kernel device times for the pre-instantiated CTA shapes (TPSB_TUNE knob)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tps_b200  # noqa: E402
from bench import tgv_visc_mult  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
settings = sys.argv[2:] or ["0,0,0", "1,1,1", "2,2,2", "0,3,0", "0,4,0"]
PI = float(np.pi)
mesh = tps_b200.cartesian_hex_mesh(n, n, n, lo=(-PI,) * 3, hi=(PI,) * 3, order_mode=1)
phys = tps_b200.Physics.dry_air(1, tgv_visc_mult())
N = n ** 3 * 64
U = torch.empty(5 * N, dtype=torch.float64, device="cuda")
g = torch.Generator(device="cuda").manual_seed(1)
for k, (v, a) in enumerate(((1.2, 0.02), (10.0, 20.0), (-5.0, 20.0), (3.0, 20.0), (253000.0, 2000.0))):
    U[k * N:(k + 1) * N] = v + a * (torch.rand(N, dtype=torch.float64, device="cuda", generator=g) - 0.5)
Y = torch.empty_like(U)
for st in settings:
    os.environ["TPSB_TUNE"] = st
    op = tps_b200.RhsOperator(mesh, order=3, physics=phys)
    for _ in range(3):
        op.Mult(U, Y)
    op.set_profiling(True)
    op.kernel_times()
    for _ in range(5):
        op.Mult(U, Y)
    kt = op.kernel_times()
    op.set_profiling(False)
    tot = sum(v[0] for v in kt.values()) / 5
    print(st, {k: round(v[0] / 5, 3) for k, v in kt.items() if v[1]}, "total ms", round(tot, 3),
          "DOF/s %.3e" % (N / (tot * 1e-3)), flush=True)
    op.close()
