"""Per-kernel device times of the generic path on the 2-D configurations (C1 mms.euler_2d size, a C4-like plasma case)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import tps_b200  # noqa: E402


def run(name, op, U, reps=20):
    x = torch.from_numpy(U).cuda()
    y = torch.empty_like(x)
    for _ in range(3):
        op.Mult(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        op.Mult(x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    op.set_profiling(True)
    for _ in range(5):
        op.Mult(x, y)
    torch.cuda.synchronize()
    kt = op.kernel_times()
    op.set_profiling(False)
    print(f"{name}: N={op.N} neq={op.neq}  {ms:.3f} ms/eval  {op.N / (ms * 1e-3):.3e} DOF-evals/s  kernels(ms total of 5): {kt}")


if __name__ == "__main__":
    PI = np.pi
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 160
    m2 = tps_b200.cartesian_quad_mesh(n, n, lo=(-PI, -PI), hi=(PI, PI))
    op = tps_b200.RhsOperator(m2, order=2, physics=tps_b200.Physics.dry_air(0, 1.0), basis_type=1, int_rule_type=1)
    N = op.N
    U2 = np.concatenate([np.full(N, 1.2), np.full(N, 12.0), np.full(N, 3.0), np.full(N, 101300 / 0.4 + 0.5 * 1.2 * 109)])
    run(f"C1 Euler {n}x{n} quads p=2 GLL", op, U2)
    op = tps_b200.RhsOperator(m2, order=2, physics=tps_b200.Physics.dry_air(1, 1.0), basis_type=1, int_rule_type=1)
    run(f"NS {n}x{n} quads p=2 GLL", op, U2)
    import axisym_cases as ac
    import oracle_api
    m = ac.box(n=(n // 4, n // 4), warp=0.05)
    d = ac.argon6_dict()
    op, orc = ac.make_pair(m, 3, 1, 0, 0, 3, "c4", True, mixture=d)
    up = ac.argon6_primitives(orc.node_coords(), 3)
    U = np.ascontiguousarray(orc.pt("cons", up).T).reshape(-1)
    run(f"C4 argon6 axisym {n // 4}x{n // 4} quads p=3, constant transport", op, U)
