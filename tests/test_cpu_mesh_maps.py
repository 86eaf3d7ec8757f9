"""Index maps: meshkit (product host C++) vs the numpy restatement of MFEM's conventions vs the
reference's own text mesh fixture (tests/golden/periodic_cube.json <- test/meshes/periodic-cube.mesh).
Integer work: bit-exact."""
import json
import os

import numpy as np
import pytest

import meshref
import tps_b200
from tps_b200 import capi

GOLD = os.path.join(os.path.dirname(__file__), "golden", "periodic_cube.json")


def test_periodic_cube_fixture_reproduced(lib_built):
    g = json.load(open(GOLD))
    m = tps_b200.cartesian_hex_mesh(3, 3, 3)
    assert np.array_equal(m["elem_verts"], np.array(g["elements"], dtype=np.int32))
    # the fixture stores 6 decimals
    assert np.abs(m["elem_xyz"] - np.array(g["node_xyz_vertex_order"])).max() < 1e-6
    # every "boundary" quad of the periodic fixture is an interior face of the identified mesh
    el1, el2, inf1, inf2 = (m[k] for k in ("face_el1", "face_el2", "face_inf1", "face_inf2"))
    assert len(el1) == 81 and (el2 >= 0).all()
    keys = {tuple(sorted(m["elem_verts"][e][meshref.HEX_FACE_VERT[i // 64]])) for e, i in zip(el1, inf1)}
    for b in g["boundary"]:
        assert tuple(sorted(b[2:])) in keys


@pytest.mark.parametrize("n", [(3, 3, 3), (4, 3, 5), (6, 6, 6)])
@pytest.mark.parametrize("periodic", [(1, 1, 1), (0, 1, 1), (0, 0, 0)])
def test_meshkit_matches_numpy_restatement(lib_built, n, periodic):
    m = tps_b200.cartesian_hex_mesh(*n, lo=(-1, 0, 2), hi=(1, 3, 2.5), periodic=periodic)
    ev, xyz = meshref.cartesian_hex(*n, lo=(-1, 0, 2), hi=(1, 3, 2.5), periodic=[bool(p) for p in periodic])
    el1, el2, i1, i2 = meshref.build_faces(ev)
    assert np.array_equal(m["elem_verts"], ev)
    assert np.array_equal(m["elem_xyz"], xyz)
    for a, b in ((m["face_el1"], el1), (m["face_el2"], el2), (m["face_inf1"], i1), (m["face_inf2"], i2)):
        assert np.array_equal(a, b)


def test_empty_and_invalid_inputs(lib_built):
    L = tps_b200.lib()
    assert L.tpsb_mk_build_faces(0, np.zeros(8, np.int32).ctypes.data_as(capi.C.POINTER(capi.C.c_int)),
                                 None, None, None, None) == 0
    with pytest.raises(tps_b200.TpsbError):
        tps_b200.cartesian_hex_mesh(2, 3, 3)  # periodic direction with < 3 elements is not a valid MFEM mesh


def test_blocked_order_is_a_permutation(lib_built):
    a = tps_b200.cartesian_hex_mesh(9, 10, 11, order_mode=0)
    b = tps_b200.cartesian_hex_mesh(9, 10, 11, order_mode=1)
    ka = {tuple(r) for r in a["elem_verts"]}
    kb = {tuple(r) for r in b["elem_verts"]}
    assert ka == kb and len(b["face_el1"]) == len(a["face_el1"])


def test_orientation_permutations_match_face_geometry(lib_built, oracle_built):
    """perm[ori] must send a face node to the Elem2-local face node at the same physical point:
    checked against the oracle's independent Loc1/Loc2 maps through the trace of a generic field."""
    import oracle_api
    T = capi.ref_tables(3)
    npn = T["np"]
    m = tps_b200.cartesian_hex_mesh(3, 3, 3)
    o = oracle_api.Oracle(3, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"])
    X = o.node_coords().reshape(-1, npn ** 3, 3)
    L = 2.0  # periodic box length
    seen = set()
    for f in range(len(m["face_el1"])):
        e1, e2 = m["face_el1"][f], m["face_el2"][f]
        lf1, lf2, ori = m["face_inf1"][f] // 64, m["face_inf2"][f] // 64, m["face_inf2"][f] % 64
        seen.add(int(ori))
        for ab in range(npn * npn):
            n1 = T["face_base"][lf1][ab]
            n2 = T["face_base"][lf2][T["perm"][ori][ab]]
            # compare the two tangential coordinates (normal coordinate differs by construction)
            ax = [d for d in range(3) if d != {0: 2, 5: 2, 1: 1, 3: 1, 2: 0, 4: 0}[int(lf1)]]
            d = X[e1, n1, ax] - X[e2, n2, ax]
            d = d - L * np.round(d / L)
            assert np.abs(d).max() < 1e-12
            assert T["iperm"][ori][T["perm"][ori][ab]] == ab
    assert len(seen) >= 2
