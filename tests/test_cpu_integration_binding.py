"""INTEGRATION.md's reference-side binding is a real file (integration/rhs_operator_b200.hpp).  MFEM and the TPS build tree
are absent here, so it is type-checked against a declarations-only stub of the MFEM / TPS interfaces it touches
(integration/stub/): every tpsb_* call, every POD field and every reference accessor it names must exist and agree in type."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_binding_compiles_against_the_interface_stub():
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-Iintegration/stub", "-Iinclude",
                        "integration/stub/check_binding.cpp"], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_binding_uses_every_block_of_the_abi():
    src = open(os.path.join(ROOT, "integration", "rhs_operator_b200.hpp")).read()
    for name in ("tpsb_mesh_maps", "tpsb_space_desc", "tpsb_physics", "tpsb_plasma_models", "tpsb_bc_set", "tpsb_halo_desc",
                 "tpsb_forcing_desc", "tpsb_create", "tpsb_destroy", "tpsb_rhs_mult", "tpsb_set_solution_view",
                 "tpsb_set_distance_field", "tpsb_add_forcing", "tpsb_get_max_char_speed", "tpsb_averaging_add_sample",
                 "TPSB_BC_INLET", "TPSB_BC_OUTLET", "TPSB_BC_WALL"):
        assert name in src, name
    assert "..." not in src.replace("A &&...base_args", "").replace("class... A", "").replace("std::forward<A>(base_args)...", ""), \
        "no elided blocks in the binding"
