"""Pure-numpy restatement of the MFEM mesh conventions the reference relies on
(test helper; independent of both the product's C++ meshkit and the C++ oracle).

[MFEM] = third-party behaviour, not under /root/reference:
  * hex vertex order / Geometry::Constants<CUBE>::FaceVert / quad_t::Orient,
  * Mesh::GetElementToFaceTable: faces are numbered by first appearance while looping
    elements then local faces; Mesh::GenerateFaces: the first element seen becomes Elem1
    (Elem1Inf = 64*lf), the second Elem2 with Elem2Inf = 64*lf + GetQuadOrientation(base, test),
  * reference call sites: src/M2ulPhyS.cpp:937-958 (element_to_faces, el1/el2 per face).
"""
import numpy as np

HEX_VERT = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0],
                     [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=np.int64)
HEX_FACE_VERT = np.array([[3, 2, 1, 0], [0, 1, 5, 4], [1, 2, 6, 5],
                          [2, 3, 7, 6], [3, 0, 4, 7], [4, 5, 6, 7]], dtype=np.int64)
QUAD_ORIENT = np.array([[0, 1, 2, 3], [0, 3, 2, 1], [1, 2, 3, 0], [1, 0, 3, 2],
                        [2, 3, 0, 1], [2, 1, 0, 3], [3, 0, 1, 2], [3, 2, 1, 0]], dtype=np.int64)


def cartesian_hex(nx, ny, nz, lo=(-1.0, -1.0, -1.0), hi=(1.0, 1.0, 1.0), periodic=(True, True, True)):
    """Cartesian hex mesh, elements and vertices x-fastest, hex vertex order as in
    test/meshes/periodic-cube.mesh (periodic directions identify vertices).
    Returns elem_verts [NE,8] int32 and elem_xyz [NE,8,3] float64 (un-wrapped coordinates)."""
    n = (nx, ny, nz)
    nv = [n[d] if periodic[d] else n[d] + 1 for d in range(3)]
    ez, ey, ex = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ex, ey, ez = ex.ravel(), ey.ravel(), ez.ravel()
    NE = nx * ny * nz
    elem_verts = np.zeros((NE, 8), dtype=np.int32)
    elem_xyz = np.zeros((NE, 8, 3), dtype=np.float64)
    h = [(hi[d] - lo[d]) / n[d] for d in range(3)]
    for a in range(8):
        ix = ex + HEX_VERT[a, 0]
        iy = ey + HEX_VERT[a, 1]
        iz = ez + HEX_VERT[a, 2]
        elem_xyz[:, a, 0] = lo[0] + ix * h[0]
        elem_xyz[:, a, 1] = lo[1] + iy * h[1]
        elem_xyz[:, a, 2] = lo[2] + iz * h[2]
        elem_verts[:, a] = (ix % nv[0]) + nv[0] * ((iy % nv[1]) + nv[1] * (iz % nv[2]))
    return elem_verts, elem_xyz


def quad_orientation(base, test):
    """[MFEM Mesh::GetQuadOrientation]"""
    i = list(test).index(base[0])
    if test[(i + 1) % 4] == base[1]:
        return 2 * i
    return 2 * i + 1


def build_faces(elem_verts):
    """Face tables in MFEM convention (slow reference implementation, small meshes only)."""
    table = {}
    face_verts = []
    el1, el2, inf1, inf2 = [], [], [], []
    for e in range(elem_verts.shape[0]):
        v = elem_verts[e]
        for lf in range(6):
            fv = [int(v[k]) for k in HEX_FACE_VERT[lf]]
            key = tuple(sorted(fv))
            if key not in table:
                table[key] = len(face_verts)
                face_verts.append(fv)
                el1.append(e)
                el2.append(-1)
                inf1.append(64 * lf)
                inf2.append(-1)
            else:
                f = table[key]
                assert el2[f] == -1, "non-manifold face"
                el2[f] = e
                inf2[f] = 64 * lf + quad_orientation(face_verts[f], fv)
    to32 = lambda a: np.asarray(a, dtype=np.int32)
    return to32(el1), to32(el2), to32(inf1), to32(inf2)


QUAD_VERT = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=np.int64)
QUAD_EDGE_VERT = np.array([[0, 1], [1, 2], [2, 3], [3, 0]], dtype=np.int64)


def cartesian_quad(nx, ny, lo=(-1.0, -1.0), hi=(1.0, 1.0), periodic=(True, True)):
    """2-D counterpart of cartesian_hex ([MFEM] Mesh::MakeCartesian2D + MakePeriodic, utils/beam_mesh.cpp)."""
    n = (nx, ny)
    nv = [n[d] if periodic[d] else n[d] + 1 for d in range(2)]
    ey, ex = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    ex, ey = ex.ravel(), ey.ravel()
    ev = np.zeros((nx * ny, 4), dtype=np.int32)
    xyz = np.zeros((nx * ny, 4, 2))
    h = [(hi[d] - lo[d]) / n[d] for d in range(2)]
    for a in range(4):
        ix, iy = ex + QUAD_VERT[a, 0], ey + QUAD_VERT[a, 1]
        xyz[:, a, 0] = lo[0] + ix * h[0]
        xyz[:, a, 1] = lo[1] + iy * h[1]
        ev[:, a] = (ix % nv[0]) + nv[0] * (iy % nv[1])
    return ev, xyz


def build_faces2d(elem_verts):
    """Edge tables in MFEM convention: orientation 1 = the second element runs along the edge backwards."""
    table, base = {}, []
    el1, el2, inf1, inf2 = [], [], [], []
    for e in range(elem_verts.shape[0]):
        v = elem_verts[e]
        for le in range(4):
            v0, v1 = int(v[QUAD_EDGE_VERT[le, 0]]), int(v[QUAD_EDGE_VERT[le, 1]])
            key = (min(v0, v1), max(v0, v1))
            if key not in table:
                table[key] = len(base)
                base.append((v0, v1))
                el1.append(e), el2.append(-1), inf1.append(64 * le), inf2.append(-1)
            else:
                f = table[key]
                assert el2[f] == -1
                el2[f] = e
                inf2[f] = 64 * le + (0 if base[f] == (v0, v1) else 1)
    to32 = lambda a: np.asarray(a, dtype=np.int32)
    return to32(el1), to32(el2), to32(inf1), to32(inf2)


def element_to_faces(NE, el1, el2):
    """Stride-7 interior-face list per element (src/M2ulPhyS.cpp:878-958)."""
    e2f = np.zeros(7 * NE, dtype=np.int32)
    for f in range(len(el1)):
        if el2[f] < 0:
            continue
        for e in (el1[f], el2[f]):
            nf = e2f[7 * e]
            e2f[7 * e + nf + 1] = f
            e2f[7 * e] = nf + 1
    return e2f


def parse_mfem_mesh(path):
    """Parse the text 'MFEM mesh v1.0' hex fixtures of the reference (elements, boundary, L2 P1 nodes)."""
    lines = [ln.strip() for ln in open(path)]
    lines = [ln for ln in lines if ln and not ln.startswith("#")]
    i = lines.index("elements")
    ne = int(lines[i + 1])
    elems = np.array([[int(t) for t in lines[i + 2 + k].split()[2:]] for k in range(ne)], dtype=np.int32)
    i = lines.index("boundary")
    nb = int(lines[i + 1])
    bdr = np.array([[int(t) for t in lines[i + 2 + k].split()] for k in range(nb)], dtype=np.int32)
    i = lines.index("nodes")
    assert lines[i + 2].endswith("L2_T1_3D_P1") and lines[i + 4] == "Ordering: 1"
    vals = np.array([[float(t) for t in ln.split()] for ln in lines[i + 5:i + 5 + ne * 8]])
    lex = vals.reshape(ne, 8, 3)
    # lexicographic L2 P1 node order -> MFEM hex vertex order
    lex2vert = [0, 1, 3, 2, 4, 5, 7, 6]
    xyz = lex[:, lex2vert, :]
    return elems, bdr, xyz
