"""Shared inputs for the parity tests: the Taylor-Green state of SURVEY.md section 8(d) plus the seeded
+-1 % perturbation that breaks every symmetry (so index-map bugs cannot hide)."""
import numpy as np

SEED = 20261018


def tgv_state(xyz, perturb=0.01, seed=SEED, gamma=1.4):
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    rho0, p0, M0 = 1.2, 101300.0, 0.1
    c0 = np.sqrt(gamma * p0 / rho0)
    V0 = M0 * c0
    rho = np.full_like(x, rho0)
    u = V0 * np.sin(x) * np.cos(y) * np.cos(z)
    v = -V0 * np.cos(x) * np.sin(y) * np.cos(z)
    w = np.zeros_like(x)
    p = p0 + rho0 * V0 * V0 / 16.0 * (np.cos(2 * x) + np.cos(2 * y)) * (np.cos(2 * z) + 2.0)
    U = np.concatenate([rho, rho * u, rho * v, rho * w, p / (gamma - 1.0) + 0.5 * rho * (u * u + v * v + w * w)])
    if perturb:
        rng = np.random.default_rng(seed)
        U = U * (1.0 + perturb * rng.uniform(-1.0, 1.0, size=U.shape))
    return np.ascontiguousarray(U)


def node_coords_from_mesh(elem_xyz, order):
    """Physical coordinates of the GL nodes (lexicographic, x fastest) of every trilinear hex."""
    from numpy.polynomial.legendre import leggauss
    g, _ = leggauss(order + 1)
    xn = 0.5 * (g + 1.0)
    npn = order + 1
    k, j, i = np.meshgrid(np.arange(npn), np.arange(npn), np.arange(npn), indexing="ij")
    xi = np.stack([xn[i.ravel()], xn[j.ravel()], xn[k.ravel()]], axis=1)  # [dof,3]
    hv = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=float)
    shp = np.ones((xi.shape[0], 8))
    for a in range(8):
        for d in range(3):
            shp[:, a] *= np.where(hv[a, d] > 0, xi[:, d], 1.0 - xi[:, d])
    return np.einsum("na,ead->end", shp, elem_xyz).reshape(-1, 3)


def rel_l2(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300)))


def warp_mesh(m, amp=0.08, lo=(-np.pi,) * 3, hi=(np.pi,) * 3):
    """Smooth, box-periodic displacement of every vertex: turns the Cartesian box into a mesh of genuinely
    trilinear (non-parallelepiped) hexahedra while keeping periodic copies of a vertex consistent."""
    xyz = m["elem_xyz"]
    L = np.asarray(hi) - np.asarray(lo)
    t = 2 * np.pi * (xyz - np.asarray(lo)) / L
    d = np.empty_like(xyz)
    d[..., 0] = np.sin(t[..., 0]) * np.cos(t[..., 1]) * np.cos(2 * t[..., 2])
    d[..., 1] = np.cos(2 * t[..., 0]) * np.sin(t[..., 1]) * np.cos(t[..., 2])
    d[..., 2] = np.cos(t[..., 0]) * np.cos(2 * t[..., 1]) * np.sin(t[..., 2])
    out = dict(m)
    out["elem_xyz"] = np.ascontiguousarray(xyz + amp * L / (2 * np.pi) * d)
    return out


HEX_FACE_VERT = np.array([[3, 2, 1, 0], [0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [3, 0, 4, 7], [4, 5, 6, 7]])


def box_face_attrs(m, lo, hi):
    """Boundary attribute per face of a (non-periodic) box mesh: 1/2 = x-/x+, 3/4 = y-/y+, 5/6 = z-/z+;
    0 on interior faces.  Uses the un-warped vertex coordinates of the face (pass the mesh before warp_mesh)."""
    el1, el2, inf1 = m["face_el1"], m["face_el2"], m["face_inf1"]
    attr = np.zeros(len(el1), dtype=np.int32)
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    tol = 1e-9 * np.abs(hi - lo).max()
    for f in np.nonzero(el2 < 0)[0]:
        c = m["elem_xyz"][el1[f], HEX_FACE_VERT[inf1[f] // 64]].mean(axis=0)
        for d in range(3):
            if abs(c[d] - lo[d]) < tol:
                attr[f] = 2 * d + 1
            elif abs(c[d] - hi[d]) < tol:
                attr[f] = 2 * d + 2
        assert attr[f] > 0, f
    return attr


def cube_rotations():
    """The 24 orientation-preserving symmetries of the reference cube as vertex permutations p: vertex a of the
    rotated element is vertex p[a] of the original one (MFEM hex vertex order, tables.hpp HEX_VERT)."""
    import itertools
    hv = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=float)
    perms = []
    for axes in itertools.permutations(range(3)):
        for signs in itertools.product((1, -1), repeat=3):
            R = np.zeros((3, 3))
            for r in range(3):
                R[r, axes[r]] = signs[r]
            if np.linalg.det(R) < 0:
                continue
            old = (hv - 0.5) @ R.T + 0.5
            perms.append([int(np.argmin(np.abs(hv - o).sum(axis=1))) for o in old])
    assert len(perms) == 24 and len({tuple(p) for p in perms}) == 24
    return np.array(perms)


def rotate_elements(m, seed=SEED):
    """Relabel every element of a hex mesh by a random cube rotation (seeded) and rebuild the face tables: the mesh an
    unstructured generator would hand over -- all eight face orientations and arbitrary (local face, local face)
    pairings occur (src/M2ulPhyS.cpp:937-958), unlike on a Cartesian numbering."""
    import tps_b200
    from tps_b200 import capi
    rot = cube_rotations()
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, 24, size=m["elem_verts"].shape[0])
    idx = rot[pick]                                      # [NE, 8]
    ev = np.ascontiguousarray(np.take_along_axis(m["elem_verts"], idx, axis=1), dtype=np.int32)
    xyz = np.ascontiguousarray(np.take_along_axis(m["elem_xyz"], idx[:, :, None], axis=1))
    NE = ev.shape[0]
    f = [np.zeros(6 * NE, np.int32) for _ in range(4)]
    nf = tps_b200.lib().tpsb_mk_build_faces(NE, capi._ip(ev), *(capi._ip(a) for a in f))
    assert nf > 0
    out = dict(m)
    out.update(elem_verts=ev, elem_xyz=xyz, face_el1=f[0][:nf].copy(), face_el2=f[1][:nf].copy(),
               face_inf1=f[2][:nf].copy(), face_inf2=f[3][:nf].copy())
    return out
