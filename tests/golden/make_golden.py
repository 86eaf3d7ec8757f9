"""Regenerates tests/golden/periodic_cube.json from the reference's own text mesh fixture
test/meshes/periodic-cube.mesh (27 periodic hexes).  Run in the build container only
(/root/reference is absent on the GPU box):  python tests/golden/make_golden.py"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import meshref  # noqa: E402

elems, bdr, xyz = meshref.parse_mfem_mesh("/root/reference/test/meshes/periodic-cube.mesh")
out = {
    "source": "pecos/tps test/meshes/periodic-cube.mesh",
    "elements": elems.tolist(),
    "boundary": bdr.tolist(),
    "node_xyz_vertex_order": xyz.tolist(),
}
with open(os.path.join(os.path.dirname(__file__), "periodic_cube.json"), "w") as f:
    json.dump(out, f)
print("wrote", len(out["elements"]), "elements")
