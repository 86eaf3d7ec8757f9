"""Deterministic small cases shared by tests/golden/make_rhs_golden.py (which evaluates them with the reference's physics
object code) and the tests that compare the device path / the oracle's port with the committed results."""
import numpy as np

import axisym_cases as ac
import oracle_api
import tps_b200
from common import node_coords_from_mesh, tgv_state, warp_mesh

PI = np.pi
CASES = ["tgv3d_ns_p2", "tgv3d_warped_p3", "tgv3d_smagorinsky_p1", "tgv3d_sigma_sponge_p2", "quad_euler_gll_p2", "axisym_dry_c4bcs_p3",
         "ternary2d_p2", "argon6_axisym_mixlen_p2", "nitrogen6_2d_p1"]


def _wall_distance(xy):
    x, y = xy[:, 0], xy[:, 1]
    return 0.02 + 0.015 * np.sin(1.3 * x + 0.2) ** 2 + 0.03 * (y - y.min()) / (y.max() - y.min() + 1e-30)


def _hex_pair(order, n, gpu, warp=False, sgs=None, sponge=None, vm=2e3):
    m = tps_b200.cartesian_hex_mesh(*n, lo=(-PI,) * 3, hi=(PI,) * 3)
    if warp:
        m = warp_mesh(m, amp=0.08, lo=(-PI,) * 3, hi=(PI,) * 3)
    orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(1, vm, 0.3, sgs=sgs, sponge=sponge), kind="ref") if not gpu else None
    op = tps_b200.RhsOperator(m, order=order, physics=tps_b200.Physics.dry_air(1, vm, 0.3, sgs=sgs, sponge=sponge)) if gpu else None
    U = tgv_state(node_coords_from_mesh(m["elem_xyz"], order))
    return orc, U, op, {}


def build(name, gpu):
    """(oracle or None, U or None, RhsOperator or None, extra).  gpu=False builds the reference-physics oracle and the
    input state (generator / CPU tests); gpu=True builds the device operator ONLY -- no oracle is touched, the input state
    and the wall-distance field come from the committed fixture."""
    orc, U, op, extra = _build(name, gpu)
    return orc, U, op, extra


def _build(name, gpu):
    if name == "tgv3d_ns_p2":
        return _hex_pair(2, (3, 3, 4), gpu)
    if name == "tgv3d_warped_p3":
        return _hex_pair(3, (3, 3, 3), gpu, warp=True)
    if name == "tgv3d_smagorinsky_p1":
        return _hex_pair(1, (4, 3, 5), gpu, sgs=(1, 0.12, 0.05), vm=50.0)
    if name == "tgv3d_sigma_sponge_p2":
        return _hex_pair(2, (3, 4, 3), gpu, warp=True, sgs=(2, 0.135, 0.0), sponge=((0.3, 1.0, -0.2), (0.4, 0.1, 0.2), 7.5, 0.8), vm=50.0)
    if name in ("quad_euler_gll_p2", "axisym_dry_c4bcs_p3"):
        m = ac.box(warp=0.06)
        args = (m, 2, 0, 1, 1, 2, "inviscid", False) if name == "quad_euler_gll_p2" else (m, 3, 1, 0, 0, 3, "c4", True)
        op, orc = ac.make_pair(*args, gpu=gpu, kind="ref", want_oracle=not gpu)
        U = ac.dry_state(orc.node_coords(), args[5]) if orc is not None else None
        return orc, U, op, {}
    if name == "ternary2d_p2":
        import plasma_cases
        m = ac.box(n=(4, 3), warp=0.05)
        d = plasma_cases.ternary_dict()
        pm = tps_b200.PlasmaModels.from_dict(d)
        if gpu:
            return None, None, tps_b200.RhsOperator(m, order=2, physics=tps_b200.Physics.plasma_mixture(pm, 1), basis_type=1,
                                                    int_rule_type=1), {}
        orc = oracle_api.Oracle(2, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                                phys=oracle_api.mixture_params(pm, 1), kind="ref", basis_type=1, int_rule=1, neq=6, nvel=2)
        up = plasma_cases.smooth_primitives(orc.node_coords())
        return orc, np.ascontiguousarray(orc.pt("cons", up).T).reshape(-1), None, {}
    if name == "argon6_axisym_mixlen_p2":
        m = ac.box(n=(4, 3), warp=0.05)
        op, orc = ac.make_pair(m, 2, 1, 0, 0, 3, "c4", True, mixture=ac.argon6_dict(), gpu=gpu, mixing_length=(0.015, 0.85, 0.5),
                               want_oracle=not gpu)
        if gpu:
            return None, None, op, {}  # the test hands the fixture's distance field to op.set_distance_field
        up = ac.argon6_primitives(orc.node_coords(), 3)
        dist = _wall_distance(orc.node_coords())
        orc.set_distance(dist)
        return orc, np.ascontiguousarray(orc.pt("cons", up).T).reshape(-1), op, {"dist": dist}
    if name == "nitrogen6_2d_p1":
        from test_gpu_plasma import nitrogen6_dict
        m = ac.box(n=(3, 3), warp=0.05)
        d = nitrogen6_dict()
        d.update(transport_model="argon_mixture", third_order_k_electron=False)
        op, orc = ac.make_pair(m, 1, 1, 0, 0, 2, None, False, mixture=d, gpu=gpu, want_oracle=not gpu)
        if gpu:
            return None, None, op, {}
        xy = orc.node_coords()
        x, y = xy[:, 0], xy[:, 1]
        MW_N, MW_E = 14.0067e-3, 5.48579908782496e-7
        nsp = np.stack([0.01 + 0.004 * np.sin(2 * x + y), 0.02 + 0.01 * np.cos(3 * x), 0.015 + 0.005 * np.sin(2 * y),
                        0.03 + 0.01 * np.cos(x - y), 0.01 + 0.004 * np.sin(2 * x + y)], 1)
        mw = np.array([MW_N - MW_E, MW_N, 2 * MW_N, MW_N, MW_E])
        rho = (nsp * mw).sum(1) + (1.0 + 0.2 * np.sin(x) * np.cos(y)) * 2 * MW_N
        up = np.column_stack([rho, 150 * np.sin(2 * x) * np.cos(y), -100 * np.cos(x) * np.sin(2 * y),
                              7000 + 800 * np.cos(2 * x) * np.cos(y), nsp, 9000 + 900 * np.sin(x + y)])
        return orc, np.ascontiguousarray(orc.pt("cons", up).T).reshape(-1), op, {}
    raise KeyError(name)
