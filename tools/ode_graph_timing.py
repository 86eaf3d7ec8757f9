"""RK4 steps per second with and without the CUDA-graph replay of tpsb_ode_step on small configurations
(launch-latency bound): C1-like 2-D Euler quads (generic path) and a small 3-D box (fast path)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import tps_b200  # noqa: E402
from common import node_coords_from_mesh, tgv_state  # noqa: E402


def bench(make, U, nsteps=200):
    out = {}
    for graph in (0, 1):
        os.environ["TPSB_ODE_GRAPH"] = str(graph)
        op = make()
        x = torch.from_numpy(U.copy()).cuda()
        op.ode_step(x, 1e-7, scheme=4, nsteps=10)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        op.ode_step(x, 1e-7, scheme=4, nsteps=nsteps)
        torch.cuda.synchronize()
        out[graph] = nsteps / (time.perf_counter() - t0)
    return out


if __name__ == "__main__":
    PI = np.pi
    m3 = tps_b200.cartesian_hex_mesh(8, 8, 8, lo=(-PI,) * 3, hi=(PI,) * 3)
    U3 = tgv_state(node_coords_from_mesh(m3["elem_xyz"], 3))
    r = bench(lambda: tps_b200.RhsOperator(m3, order=3, physics=tps_b200.Physics.dry_air(1, 1.0)), U3)
    print(f"3-D box 8^3 p=3 (32768 nodes), RK4 steps/s: eager {r[0]:.0f}  graph {r[1]:.0f}  x{r[1] / r[0]:.2f}")
    m2 = tps_b200.cartesian_quad_mesh(40, 40, lo=(-PI, -PI), hi=(PI, PI))
    op = tps_b200.RhsOperator(m2, order=2, physics=tps_b200.Physics.dry_air(0, 1.0), basis_type=1, int_rule_type=1)
    N = op.N
    U2 = np.concatenate([np.full(N, 1.2), np.full(N, 12.0), np.full(N, 3.0), np.full(N, 101300 / 0.4 + 0.5 * 1.2 * 109)])
    r = bench(lambda: tps_b200.RhsOperator(m2, order=2, physics=tps_b200.Physics.dry_air(0, 1.0), basis_type=1, int_rule_type=1), U2)
    print(f"2-D Euler 40x40 quads p=2 GLL ({N} nodes), RK4 steps/s: eager {r[0]:.0f}  graph {r[1]:.0f}  x{r[1] / r[0]:.2f}")
