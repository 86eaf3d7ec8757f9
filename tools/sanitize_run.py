"""One small evaluation of every kernel set, for compute-sanitizer (the reference's bar: test/cuda-memcheck.test).
  compute-sanitizer --tool memcheck|racecheck|initcheck python tools/sanitize_run.py
Prints one line per case; exits non-zero if a result is not finite."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch

import tps_b200
from common import box_face_attrs, node_coords_from_mesh, rotate_elements, tgv_state, warp_mesh

PI = np.pi
BCS = [(1, 0, 2, (1.2, 25.0, 1.0, -2.0)), (2, 1, 0, (101300.0,)), (3, 2, 3, (310.0,)), (4, 2, 2, ()), (5, 2, 0, ()),
       (6, 2, 3, (290.0,))]


def run(label, op, U, steps=True):
    x = torch.from_numpy(U).cuda()
    y = op.Mult(x)
    ok = bool(torch.isfinite(y).all().item())
    if steps:
        op.ode_step(x, 1e-7, scheme=4, nsteps=3)  # eager step + graph capture + replay
        ok = ok and bool(torch.isfinite(x).all().item())
    hx = torch.from_numpy(U).pin_memory()
    hy = torch.empty_like(hx).pin_memory()
    op.mult_host(hx, hy)
    ok = ok and bool(torch.isfinite(hy).all().item())
    torch.cuda.synchronize()
    print(f"{label:34s} path={op.path():8s} N={op.N:6d} finite={ok}", flush=True)
    op.close()
    return ok


def main():
    ok = True
    box = tps_b200.cartesian_hex_mesh(4, 3, 3, lo=(-PI,) * 3, hi=(PI,) * 3)
    for order in (3, 2, 1):
        m = rotate_elements(box)
        U = tgv_state(node_coords_from_mesh(m["elem_xyz"], order))
        ok &= run(f"affine box p={order}", tps_b200.RhsOperator(m, order=order, physics=tps_b200.Physics.dry_air(1, 2e4, 0.2)), U)
    os.environ["TPSB_PATH"] = "unfused"
    U = tgv_state(node_coords_from_mesh(box["elem_xyz"], 3))
    ok &= run("affine box p=3 (four launches)", tps_b200.RhsOperator(box, order=3, physics=tps_b200.Physics.dry_air(1, 2e4, 0.2)), U)
    del os.environ["TPSB_PATH"]
    lo, hi = (0.0, 0.0, 0.0), (2.0, 1.2, 1.0)
    ch = tps_b200.cartesian_hex_mesh(4, 3, 3, lo=lo, hi=hi, periodic=(0, 0, 0))
    attr = box_face_attrs(ch, lo, hi)
    bcs = [tps_b200.BcDesc.make(*b) for b in BCS]
    for warp in (False, True):
        m = warp_mesh(ch, amp=0.1, lo=lo, hi=hi) if warp else ch
        U = tgv_state(node_coords_from_mesh(m["elem_xyz"], 3) * PI)
        ok &= run(f"channel with BCs warp={warp}", tps_b200.RhsOperator(m, order=3, physics=tps_b200.Physics.dry_air(1, 4e3, 0.2),
                                                                      face_attr=attr, use_bc_in_grad=True, bcs=bcs), U)
    U = tgv_state(node_coords_from_mesh(warp_mesh(ch, amp=0.1, lo=lo, hi=hi)["elem_xyz"], 3) * PI)
    ok &= run("SGS + sponge (general path)", tps_b200.RhsOperator(
        warp_mesh(ch, amp=0.1, lo=lo, hi=hi), order=3, face_attr=attr, bcs=bcs,
        physics=tps_b200.Physics.dry_air(1, 4e3, 0.2, sgs=(2, 0.09, 0.0), sponge=((1.0, 0.0, 0.0), (1.5, 0.0, 0.0), 20.0, 0.3))), U)
    # generic path: 2-D GLL quads (config C1 type) and a ternary plasma mixture with chemistry (config C3 type)
    from plasma_cases import smooth_primitives, ternary_models
    import oracle_api
    q = tps_b200.cartesian_quad_mesh(6, 5, lo=(-PI, -PI), hi=(PI, PI))
    n2 = 6 * 5 * 9
    xy = np.random.default_rng(1).uniform(-PI, PI, (n2, 2))
    rho = 1.0 + 0.2 * np.sin(xy[:, 0]) * np.cos(xy[:, 1])
    u, v, p = 30 + 5 * np.cos(xy[:, 0]), -10 + 4 * np.sin(xy[:, 1]), 101300 * (1 + 0.05 * np.cos(xy[:, 0] - xy[:, 1]))
    U = np.ascontiguousarray(np.concatenate([rho, rho * u, rho * v, p / 0.4 + 0.5 * rho * (u * u + v * v)]))
    ok &= run("2-D GLL quads (generic path)", tps_b200.RhsOperator(q, order=2, physics=tps_b200.Physics.dry_air(1, 1e3), basis_type=1,
                                                                  int_rule_type=1), U)
    models = ternary_models()
    op = tps_b200.RhsOperator(q, order=2, physics=tps_b200.Physics.plasma_mixture(models), basis_type=1, int_rule_type=1, nvel=2)
    up = smooth_primitives(xy)   # [rho, u, v, T_h, n_ion, T_e]
    orc = oracle_api.Oracle(2, q["elem_xyz"], q["face_el1"], q["face_el2"], q["face_inf1"], q["face_inf2"],
                            phys=oracle_api.mixture_params(models), kind="ref", basis_type=1, int_rule=1, neq=6, nvel=2)
    Uc = orc.pt("cons", up)       # conserved state from the primitives (host side, checker)
    ok &= run("ternary plasma (generic path)", op, np.ascontiguousarray(Uc.T.reshape(-1)), steps=False)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
