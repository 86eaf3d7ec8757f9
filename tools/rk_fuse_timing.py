"""RK4 steps per second at the bench size with the stage update fused into the residual kernel (default) or as separate
axpy sweeps (TPSB_ODE_FUSE=0):  python tools/rk_fuse_timing.py [n]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import tps_b200  # noqa: E402
from common import node_coords_from_mesh, tgv_state  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
PI = np.pi
m = tps_b200.cartesian_hex_mesh(n, n, n, lo=(-PI,) * 3, hi=(PI,) * 3, order_mode=1)
op = tps_b200.RhsOperator(m, order=3, physics=tps_b200.Physics.dry_air(1, 1420.0))
U = tgv_state(node_coords_from_mesh(m["elem_xyz"], 3), perturb=0.0)
x = torch.from_numpy(U).cuda()
op.ode_step(x, 1e-7, scheme=4, nsteps=3)
torch.cuda.synchronize()
t0 = time.perf_counter()
op.ode_step(x, 1e-7, scheme=4, nsteps=10)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"TPSB_ODE_FUSE={os.environ.get('TPSB_ODE_FUSE', '1')}: {n}^3 p=3, RK4 step {dt * 1e3:.2f} ms "
      f"({op.N * 4 / dt:.3e} DOF-evals/s incl. the stage updates), finite = {bool(torch.isfinite(x).all())}")
