"""The remaining ForcingTerms subclasses (src/forcing_terms.cpp): ConstantPressureGradient, HeatSource, JouleHeating and
SpongeZone (planar / annular, user-defined / mixed-out target) through tpsb_add_forcing on every kernel set, against the
oracle's restatement of the reference's updateTerms bodies.  Bar: per-equation rel-L2 <= 1e-10."""
import numpy as np
import pytest

import oracle_api
import tps_b200
from common import box_face_attrs, node_coords_from_mesh, rel_l2, tgv_state, warp_mesh
from tps_b200.capi import ForcingDesc

pytestmark = pytest.mark.gpu
PI = np.pi


def _desc(kind, **kw):
    d = ForcingDesc()
    d.kind = {"pressure_gradient": 0, "heat_source": 1, "joule_heating": 2, "sponge_zone": 3}[kind]

    def put(dst, src):
        for i, v in enumerate(src):
            dst[i] = float(v)
    if kind == "pressure_gradient":
        put(d.pressure_grad, kw["g"])
    elif kind == "heat_source":
        put(d.hs_point1, kw["point1"]), put(d.hs_point2, kw["point2"])
        d.hs_radius, d.hs_value = kw["radius"], kw["value"]
    elif kind == "joule_heating":
        d.joule_heating = kw["field"].ctypes.data
        d._keep = kw["field"]
    else:
        put(d.sz_normal, kw["normal"]), put(d.sz_point0, kw["point0"]), put(d.sz_point_init, kw["point_init"])
        d.sz_type, d.sz_mixed_out = int(kw.get("type", 0)), int(kw.get("mixed_out", False))
        d.sz_r1, d.sz_r2, d.sz_tol, d.sz_mult = (float(kw.get(k, 0.0)) for k in ("r1", "r2", "tol", "mult"))
        put(d.sz_target, kw.get("target", (0,) * 5))
    return d


# SpongeZone geometry (src/forcing_terms.cpp:566-607): the zone lies between the entry plane through point_init and the end
# plane through point0, and the normal points from the END plane back INTO the zone (distF = n.(x - point0) > 0 and
# distInit = -n.(x - point_init) > 0 inside); sigma grows from the entry plane towards the end plane
FORCINGS = {
    "pressure_gradient": dict(g=(8.0, -3.0, 1.5)),
    "heat_source": dict(point1=(-1.0, -0.5, -0.2), point2=(1.5, 0.8, 0.6), radius=1.1, value=3.0e5),
    "sponge_planar": dict(normal=(-2.0, -0.3, 0.0), point0=(2.8, 0.0, 0.0), point_init=(0.6, 0.0, 0.0), mult=1.7,
                          target=(1.1, 20.0, -4.0, 2.0, 99000.0)),
    "sponge_planar_mixed_out": dict(normal=(-1.0, 0.0, 0.0), point0=(2.9, 0.0, 0.0), point_init=(0.4, 0.0, 0.0), mult=0.8,
                                    mixed_out=True, tol=0.35),
    "sponge_annulus": dict(normal=(0.0, 0.0, -1.0), point0=(0.0, 0.0, 2.5), point_init=(0.0, 0.0, -2.9), type=1, r1=1.2, r2=3.3,
                           mult=1.3, target=(1.15, 3.0, 6.0, 25.0, 100500.0)),
}


def _apply(op, orc, name, N):
    kw = FORCINGS[name] if name in FORCINGS else None
    if name == "joule_heating":
        import torch
        rng = np.random.default_rng(5)
        jh = np.ascontiguousarray(rng.uniform(-1e5, 4e5, N))  # negative entries must be ignored
        op.add_forcing("joule_heating", field=torch.from_numpy(jh).cuda())
        orc.add_forcing(_desc("joule_heating", field=jh))
        return
    kind = "sponge_zone" if name.startswith("sponge") else name
    op.add_forcing(kind, **kw)
    orc.add_forcing(_desc(kind, **kw))


@pytest.mark.parametrize("name", ["pressure_gradient", "heat_source", "joule_heating", "sponge_planar",
                                  "sponge_planar_mixed_out", "sponge_annulus"])
@pytest.mark.parametrize("path", ["", "general", "generic"])
def test_forcing_term_parity_3d(lib_built, oracle_built, monkeypatch, name, path):
    import torch
    m = tps_b200.cartesian_hex_mesh(5, 4, 4, lo=(-PI,) * 3, hi=(PI,) * 3)
    if path:
        monkeypatch.setenv("TPSB_PATH", path)
    op = tps_b200.RhsOperator(m, order=3, physics=tps_b200.Physics.dry_air(1, 2e4, 0.2))
    assert op.path() == (path or "fused")
    orc = oracle_api.Oracle(3, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(1, 2e4, 0.2))
    U = tgv_state(orc.node_coords())
    x = torch.from_numpy(U).cuda()
    y0 = op.Mult(x).cpu().numpy()
    _apply(op, orc, name, orc.N)
    y = op.Mult(x).cpu().numpy()
    yo = orc.mult(U)
    N = orc.N
    for k in range(5):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, (name, k)
    assert rel_l2(y, y0) > 1e-9  # the term is active
    # time stepping applies the term in every stage (the stage update is not fused over it)
    xs = torch.from_numpy(U.copy()).cuda()
    op.ode_step(xs, 1e-6, scheme=4, nsteps=4)
    ref = orc.rk4(U, 1e-6, 4)
    assert rel_l2(xs.cpu().numpy(), ref) < 1e-11  # (the mixed-out target is the root of a cancelling quadratic)
    op.clear_forcings()
    assert np.array_equal(op.Mult(x).cpu().numpy(), y0)


def test_forcing_terms_stack_on_a_trilinear_channel(lib_built, oracle_built):
    """Pressure gradient + heat source + planar sponge together, with boundary conditions, on warped elements."""
    import torch
    lo, hi = (0.0, 0.0, 0.0), (3.0, 1.2, 1.0)
    m0 = tps_b200.cartesian_hex_mesh(6, 3, 3, lo=lo, hi=hi, periodic=(0, 0, 1))
    attr = box_face_attrs(m0, lo, hi)
    m = warp_mesh(m0, amp=0.05, lo=lo, hi=hi)
    bcs = [(1, 0, 2, (1.2, 25.0, 1.0, -2.0)), (2, 1, 0, (101300.0,)), (3, 2, 3, (310.0,)), (4, 2, 2, ())]
    op = tps_b200.RhsOperator(m, order=3, physics=tps_b200.Physics.dry_air(1, 4e3, 0.2), face_attr=attr, use_bc_in_grad=True,
                              bcs=[tps_b200.BcDesc.make(*b) for b in bcs])
    orc = oracle_api.Oracle(3, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(1, 4e3, 0.2))
    orc.set_bcs(attr, [oracle_api.make_bc(*b) for b in bcs], True)
    U = tgv_state(orc.node_coords() * PI)
    for name in ("pressure_gradient", "heat_source", "sponge_planar"):
        _apply(op, orc, name, orc.N)
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo = orc.mult(U)
    N = orc.N
    for k in range(5):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k


def test_forcing_terms_2d_and_mixture(lib_built, oracle_built):
    """Generic path: pressure gradient and planar sponge on Gauss-Lobatto quadrilaterals (dry air); heat source and Joule
    heating on the axisymmetric six-species two-temperature mixture (the electron energy receives the heating too)."""
    import torch
    import axisym_cases as ac
    m = ac.box(n=(6, 5), warp=0.04)
    op, orc = ac.make_pair(m, 2, 1, 1, 1, 2, "c4", True)
    U = ac.dry_state(orc.node_coords(), 2)
    for kind, kw in (("pressure_gradient", dict(g=(5.0, -2.0, 0.0))),
                     ("sponge_zone", dict(normal=(-1.0, -0.2, 0.0), point0=(1.65, 0.0, 0.0), point_init=(1.0, 0.0, 0.0), mult=2.0,
                                          target=(1.1, 10.0, 2.0, 0.0, 100000.0)))):
        op.add_forcing(kind, **kw)
        orc.add_forcing(_desc(kind, **kw))
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo = orc.mult(U)
    N = orc.N
    for k in range(4):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k
    op, orc = ac.make_pair(m, 2, 1, 1, 1, 3, "c4", True, mixture=ac.argon6_dict())
    xy = orc.node_coords()
    U = np.ascontiguousarray(orc.pt("cons", ac.argon6_primitives(xy, 3)).T.reshape(-1))
    N = orc.N
    jh = np.ascontiguousarray(np.random.default_rng(2).uniform(-1e4, 5e4, N))
    op.add_forcing("joule_heating", field=torch.from_numpy(jh).cuda())
    orc.add_forcing(_desc("joule_heating", field=jh))
    kw = dict(point1=(0.6, -0.3, 0.0), point2=(1.6, 0.4, 0.0), radius=0.2, value=2.0e4)
    op.add_forcing("heat_source", **kw)
    orc.add_forcing(_desc("heat_source", **kw))
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo = orc.mult(U)
    for k in range(orc.neq):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k
