"""Face-orientation coverage: every element of the box is relabelled by a random cube rotation,
so every Elem2Inf orientation a valid mesh can produce and all 36 (local face, local face) pairings occur -- what any
unstructured hex mesh produces (src/M2ulPhyS.cpp:937-958).  (Two right-handed hexahedra see their common face with
opposite handedness, so only the four ODD quad orientations 1, 3, 5, 7 exist on a valid mesh; the even ones would need
an inverted element.  The permutation tables of all eight are checked on the CPU in tests/test_cpu_mesh_maps.py.)  The three kernel sets (fused/fast affine, general trilinear, generic) must
reproduce the oracle on it; bar: per-equation rel-L2 <= 1e-10, gradients 1e-11."""
import numpy as np
import pytest

import oracle_api
import tps_b200
from common import box_face_attrs, rel_l2, rotate_elements, tgv_state, warp_mesh

pytestmark = pytest.mark.gpu
PI = np.pi


def _coverage(m):
    two = m["face_el2"] >= 0
    ori = set((m["face_inf2"][two] % 64).tolist())
    pairs = set(zip((m["face_inf1"][two] // 64).tolist(), (m["face_inf2"][two] // 64).tolist()))
    return ori, pairs


def _check(m, order, eq, path, monkeypatch, bcs=None, attr=None, use=False, tol_g=1e-11):
    import torch
    if path:
        monkeypatch.setenv("TPSB_PATH", path)
    kw = {}
    if bcs:
        kw = dict(face_attr=attr, use_bc_in_grad=use, bcs=[tps_b200.BcDesc.make(*b) for b in bcs])
    op = tps_b200.RhsOperator(m, order=order, physics=tps_b200.Physics.dry_air(eq, 2e4, 0.2), **kw)
    orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(eq, 2e4, 0.2))
    if bcs:
        orc.set_bcs(attr, [oracle_api.make_bc(*b) for b in bcs], use)
    U = tgv_state(orc.node_coords())
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo, go = orc.mult(U, want_grad=True)
    N = orc.N
    for k in range(5):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, (path, k)
    assert rel_l2(op.fields()[1].cpu().numpy(), go) < tol_g
    assert abs(op.max_char_speed() / orc.max_char_speed - 1) < 1e-13


def test_rotated_mesh_covers_every_orientation_and_pairing(lib_built):
    m = rotate_elements(tps_b200.cartesian_hex_mesh(5, 4, 6, lo=(-PI,) * 3, hi=(PI,) * 3))
    ori, pairs = _coverage(m)
    assert ori == {1, 3, 5, 7} and len(pairs) == 36


@pytest.mark.parametrize("order,eq", [(3, 1), (3, 0), (2, 1), (1, 1)])
@pytest.mark.parametrize("path", ["", "unfused", "general", "generic"])
def test_rotated_affine_box(lib_built, oracle_built, monkeypatch, order, eq, path):
    """Rotated parallelepipeds stay parallelepipeds: default = fused (p = 3) / fast path, then the other kernel sets."""
    m = rotate_elements(tps_b200.cartesian_hex_mesh(5, 4, 6, lo=(-PI,) * 3, hi=(PI,) * 3))
    _check(m, order, eq, path, monkeypatch)


@pytest.mark.parametrize("order", [3, 2])
@pytest.mark.parametrize("path", ["", "generic"])
def test_rotated_trilinear_box(lib_built, oracle_built, monkeypatch, order, path):
    m = rotate_elements(warp_mesh(tps_b200.cartesian_hex_mesh(5, 4, 6, lo=(-PI,) * 3, hi=(PI,) * 3), amp=0.1), seed=7)
    _check(m, order, 1, path, monkeypatch)


@pytest.mark.parametrize("warp", [False, True])
@pytest.mark.parametrize("use", [False, True])
def test_rotated_channel_with_boundary_conditions(lib_built, oracle_built, monkeypatch, warp, use):
    """Walls / inlet / outlet on rotated elements: the boundary faces sit on arbitrary local faces."""
    lo, hi = (0.0, -1.0, 0.0), (2.0, 1.0, 1.5)
    m0 = tps_b200.cartesian_hex_mesh(5, 4, 3, lo=lo, hi=hi, periodic=(0, 0, 1))
    attr0 = box_face_attrs(m0, lo, hi)
    if warp:
        m0 = warp_mesh(m0, amp=0.06, lo=lo, hi=hi)
    m = rotate_elements(m0, seed=11)
    # boundary attributes follow the face's vertices, not its (renumbered) index
    key0 = {}
    from meshref import HEX_FACE_VERT
    for f in np.nonzero(m0["face_el2"] < 0)[0]:
        v = m0["elem_verts"][m0["face_el1"][f], HEX_FACE_VERT[m0["face_inf1"][f] // 64]]
        key0[tuple(sorted(v.tolist()))] = attr0[f]
    attr = np.zeros(len(m["face_el1"]), np.int32)
    for f in np.nonzero(m["face_el2"] < 0)[0]:
        v = m["elem_verts"][m["face_el1"][f], HEX_FACE_VERT[m["face_inf1"][f] // 64]]
        attr[f] = key0[tuple(sorted(v.tolist()))]
    bcs = [(1, 0, 2, (1.2, 30.0, 0.0, 0.0)), (2, 1, 0, (101300.0,)), (3, 2, 3, (300.0,)), (4, 2, 2, ())]
    _check(m, 3, 1, "", monkeypatch, bcs=bcs, attr=attr, use=use)
