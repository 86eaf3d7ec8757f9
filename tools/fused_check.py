"""Development check of the fused fast path (rhs_fused.cuh): parity against the oracle and the unfused kernels on a small
box, then per-kernel timings at n^3 for a few launch shapes.  python tools/fused_check.py [n]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch

import oracle_api
import tps_b200
from common import rel_l2, tgv_state

PI = np.pi


def small():
    m = tps_b200.cartesian_hex_mesh(5, 6, 4, lo=(-PI,) * 3, hi=(PI,) * 3)
    phys = tps_b200.Physics.dry_air(1, 2e4, 0.1)
    orc = oracle_api.Oracle(3, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(1, 2e4, 0.1))
    U = tgv_state(orc.node_coords())
    x = torch.from_numpy(U).cuda()
    op = tps_b200.RhsOperator(m, order=3, physics=phys)
    y = op.Mult(x).cpu().numpy()
    yo, go = orc.mult(U, want_grad=True)
    print("fused vs oracle  ", rel_l2(y, yo), " mcs", op.max_char_speed() / orc.max_char_speed - 1)
    g = op.fields()[1].cpu().numpy()
    print("gradUp vs oracle ", rel_l2(g, go))
    os.environ["TPSB_PATH"] = "unfused"
    op2 = tps_b200.RhsOperator(m, order=3, physics=phys)
    del os.environ["TPSB_PATH"]
    y2 = op2.Mult(x).cpu().numpy()
    print("fused vs unfused ", rel_l2(y, y2))
    N = orc.N
    for k in range(5):
        print("  eq", k, rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]))


def timing(n):
    m = tps_b200.cartesian_hex_mesh(n, n, n, lo=(-PI,) * 3, hi=(PI,) * 3, order_mode=1)
    phys = tps_b200.Physics.dry_air(1, 100.0)
    cases = [("fused default", {})] + [(f"fused tune {t}", {"TPSB_TUNE": f"{t},0,0"}) for t in range(1, 6)]
    cases += [(f"lift ctas {q}", {"TPSB_LIFT_CTAS": str(q)}) for q in (1, 3, 4)] + [("unfused", {"TPSB_PATH": "unfused"})]
    for label, env in cases:
        os.environ.update(env)
        op = tps_b200.RhsOperator(m, order=3, physics=phys)
        for k in env:
            del os.environ[k]
        N = op.N
        U = torch.rand(5 * N, dtype=torch.float64, device="cuda") * 0.1
        U[0:N] += 1.2
        U[4 * N:] += 2.5e5
        Y = torch.empty_like(U)
        for _ in range(3):
            op.Mult(U, Y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            op.Mult(U, Y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        op.set_profiling(True)
        op.kernel_times()
        for _ in range(3):
            op.Mult(U, Y)
        kt = op.kernel_times()
        op.set_profiling(False)
        print(f"{label:14s} n={n}: {ms:7.3f} ms/eval  {N / ms / 1e6:7.3f} GDOF/s  " +
              "  ".join(f"{k} {v[0] / 3:.3f}" for k, v in kt.items() if v[1] > 0), flush=True)
        op.close()
        del U, Y


if __name__ == "__main__":
    small()
    timing(int(sys.argv[1]) if len(sys.argv) > 1 else 64)
