"""The oracle against the committed golden du/dt vectors (tests/golden/rhs_golden.npz, made with the reference's physics
object code): the dry-air PORT must reproduce them (it runs where oracle/_ref is absent), and the reference back end must
reproduce its own fixture (guards the fixture against drift of the case builders)."""
import os

import numpy as np
import pytest

import axisym_cases as ac
import golden_cases
import oracle_api
import tps_b200
from common import node_coords_from_mesh, rel_l2, tgv_state, warp_mesh

GOLD = os.path.join(os.path.dirname(__file__), "golden", "rhs_golden.npz")
HAVE_REF = os.path.exists(os.path.join(oracle_api.ORACLE_DIR, "_ref", "liboracle_ref.so"))
PI = np.pi


def test_fixture_holds_every_case():
    g = np.load(GOLD)
    for name in golden_cases.CASES:
        assert name + "/y" in g and name + "/U" in g and np.isfinite(g[name + "/y"]).all()
        assert g[name + "/y"].shape == g[name + "/U"].shape


@pytest.mark.parametrize("name,order,n,warp", [("tgv3d_ns_p2", 2, (3, 3, 4), False), ("tgv3d_warped_p3", 3, (3, 3, 3), True)])
def test_dry_air_port_reproduces_the_reference_physics_fixture(lib_built, oracle_built, name, order, n, warp):
    g = np.load(GOLD)
    m = tps_b200.cartesian_hex_mesh(*n, lo=(-PI,) * 3, hi=(PI,) * 3)
    if warp:
        m = warp_mesh(m, amp=0.08, lo=(-PI,) * 3, hi=(PI,) * 3)
    orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(1, 2e3, 0.3), kind="port")
    U = g[name + "/U"]
    assert np.array_equal(U, tgv_state(node_coords_from_mesh(m["elem_xyz"], order)))  # the builders have not drifted
    assert rel_l2(orc.mult(U), g[name + "/y"]) < 1e-12


def test_quad_euler_port_reproduces_the_fixture(lib_built, oracle_built):
    g = np.load(GOLD)
    m = ac.box(warp=0.06)
    _, orc = ac.make_pair(m, 2, 0, 1, 1, 2, "inviscid", False, gpu=False, kind="port")
    assert rel_l2(orc.mult(g["quad_euler_gll_p2/U"]), g["quad_euler_gll_p2/y"]) < 1e-12


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref (reference object code) not built")
@pytest.mark.parametrize("name", golden_cases.CASES)
def test_reference_back_end_reproduces_its_fixture(lib_built, oracle_built, name):
    g = np.load(GOLD)
    orc, U, _, _ = golden_cases.build(name, gpu=False)
    assert np.array_equal(U, g[name + "/U"])
    assert rel_l2(orc.mult(U), g[name + "/y"]) < 1e-13
