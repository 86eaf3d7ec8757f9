// meshkit -- host-side stand-in for the handful of MFEM mesh services the RHS path consumes
// when MFEM itself is not linked (MFEM still does this job in a TPS build).
// It produces tables in MFEM's conventions so that an MFEM-built table and a meshkit-built table
// of the same mesh can be compared bit-for-bit:
//   * hex vertex order and Cartesian numbering of test/meshes/periodic-cube.mesh,
//   * face numbering by first appearance over (element, local face)   [MFEM GetElementToFaceTable],
//   * Elem1 = first element seen, Elem2Inf = 64*lf + quad orientation [MFEM GenerateFaces],
// which is what M2ulPhyS::initIndirectionArrays walks (src/M2ulPhyS.cpp:937-958).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "../../include/tpsb200.h"
#include "tables.hpp"

namespace {

struct Key {
  int a, b, c;  // three smallest vertex ids of the quad
  bool operator==(const Key &o) const { return a == o.a && b == o.b && c == o.c; }
};
struct KeyHash {
  size_t operator()(const Key &k) const {
    uint64_t h = static_cast<uint64_t>(k.a) * 0x9E3779B97F4A7C15ull;
    h ^= (static_cast<uint64_t>(k.b) + 0x7F4A7C159E3779B9ull + (h << 6) + (h >> 2));
    h ^= (static_cast<uint64_t>(k.c) + 0x94D049BB133111EBull + (h << 6) + (h >> 2));
    return static_cast<size_t>(h);
  }
};

// [MFEM Mesh::GetQuadOrientation]
int quad_orientation(const int *base, const int *test) {
  int i;
  for (i = 0; i < 4; i++)
    if (test[i] == base[0]) break;
  if (i == 4) return -1;
  if (test[(i + 1) % 4] == base[1]) return 2 * i;
  return 2 * i + 1;
}

}  // namespace

extern "C" int tpsb_mk_cartesian_hex(int nx, int ny, int nz, const double lo[3], const double hi[3],
                                     const int periodic[3], int order_mode, int *elem_verts, double *elem_xyz) {
  if (nx < 1 || ny < 1 || nz < 1 || !elem_verts || !elem_xyz) return TPSB_EINVAL;
  const int n[3] = {nx, ny, nz};
  int nv[3];
  for (int d = 0; d < 3; d++) {
    if (periodic[d] && n[d] < 3) return TPSB_EINVAL;  // two elements would share two faces
    nv[d] = periodic[d] ? n[d] : n[d] + 1;
  }
  const double h[3] = {(hi[0] - lo[0]) / nx, (hi[1] - lo[1]) / ny, (hi[2] - lo[2]) / nz};
  // element visiting order
  std::vector<int64_t> order;
  order.reserve(static_cast<size_t>(nx) * ny * nz);
  if (order_mode == 0) {
    for (int64_t e = 0; e < static_cast<int64_t>(nx) * ny * nz; e++) order.push_back(e);
  } else {
    // blocked order: tiles, lexicographic inside and across tiles.  8 x 8 in the plane keeps the face neighbours of a
    // tile close in memory (L2 reuse of the neighbour reads); the tile is only 3 elements thick in z so that an element's
    // dependency cone spans few z-layers -- the chunked host-buffer pipeline (tpsb_rhs_mult_host) can then finish all
    // but the first and last few layers while copies are still in flight.  Measured at 96^3 (B200): 8x8x8 tiles 4.28e9
    // DOF-evals/s device-resident / 0.88e9 through host buffers, 8x8x3 4.25e9 / 1.03e9, 8x8x2 4.22e9 / 1.05e9,
    // lexicographic 4.09e9 / 0.98e9.  TPSB_MK_TILE="bx,by,bz" overrides (development).
    int B[3] = {8, 8, 3};
    if (const char *ev = getenv("TPSB_MK_TILE")) sscanf(ev, "%d,%d,%d", &B[0], &B[1], &B[2]);
    for (int d = 0; d < 3; d++) B[d] = std::max(1, B[d]);
    for (int bz = 0; bz < nz; bz += B[2])
      for (int by = 0; by < ny; by += B[1])
        for (int bx = 0; bx < nx; bx += B[0])
          for (int k = bz; k < std::min(bz + B[2], nz); k++)
            for (int j = by; j < std::min(by + B[1], ny); j++)
              for (int i = bx; i < std::min(bx + B[0], nx); i++)
                order.push_back(i + static_cast<int64_t>(nx) * (j + static_cast<int64_t>(ny) * k));
  }
  for (size_t e = 0; e < order.size(); e++) {
    const int64_t c = order[e];
    const int ei[3] = {static_cast<int>(c % nx), static_cast<int>((c / nx) % ny),
                       static_cast<int>(c / (static_cast<int64_t>(nx) * ny))};
    for (int a = 0; a < 8; a++) {
      int iv[3];
      for (int d = 0; d < 3; d++) {
        iv[d] = ei[d] + tpsb::HEX_VERT[a][d];
        elem_xyz[(e * 8 + a) * 3 + d] = lo[d] + iv[d] * h[d];
      }
      elem_verts[e * 8 + a] = (iv[0] % nv[0]) + nv[0] * ((iv[1] % nv[1]) + nv[1] * (iv[2] % nv[2]));
    }
  }
  return TPSB_OK;
}

extern "C" int tpsb_mk_build_faces(int num_elems, const int *elem_verts, int *face_el1, int *face_el2,
                                   int *face_inf1, int *face_inf2) {
  if (num_elems < 0 || !elem_verts) return -1;
  std::unordered_map<Key, int, KeyHash> table;
  table.reserve(static_cast<size_t>(num_elems) * 4);
  std::vector<int> base;  // first-seen vertex order of every face
  base.reserve(static_cast<size_t>(num_elems) * 12);
  int nfaces = 0;
  const bool fill = face_el1 && face_el2 && face_inf1 && face_inf2;
  for (int e = 0; e < num_elems; e++) {
    const int *v = &elem_verts[static_cast<size_t>(e) * 8];
    for (int lf = 0; lf < 6; lf++) {
      int fv[4], s[4];
      for (int k = 0; k < 4; k++) s[k] = fv[k] = v[tpsb::HEX_FACE_VERT[lf][k]];
      std::sort(s, s + 4);
      const Key key{s[0], s[1], s[2]};
      auto it = table.find(key);
      if (it == table.end()) {
        table.emplace(key, nfaces);
        base.insert(base.end(), fv, fv + 4);
        if (fill) {
          face_el1[nfaces] = e;
          face_el2[nfaces] = -1;
          face_inf1[nfaces] = 64 * lf;
          face_inf2[nfaces] = -1;
        }
        nfaces++;
      } else if (fill) {
        const int f = it->second;
        if (face_el2[f] != -1) return -2;  // non-manifold
        const int ori = quad_orientation(&base[static_cast<size_t>(f) * 4], fv);
        if (ori < 0) return -3;
        face_el2[f] = e;
        face_inf2[f] = 64 * lf + ori;
      }
    }
  }
  return nfaces;
}

// ---- 2-D: Cartesian quadrilaterals (Mesh::MakeCartesian2D + MakePeriodic as utils/beam_mesh.cpp uses them) ----
extern "C" int tpsb_mk_cartesian_quad(int nx, int ny, const double lo[2], const double hi[2], const int periodic[2],
                                      int *elem_verts, double *elem_xyz) {
  if (nx < 1 || ny < 1 || !elem_verts || !elem_xyz) return TPSB_EINVAL;
  const int n[2] = {nx, ny};
  int nv[2];
  for (int d = 0; d < 2; d++) {
    if (periodic[d] && n[d] < 3) return TPSB_EINVAL;
    nv[d] = periodic[d] ? n[d] : n[d] + 1;
  }
  static const int QV[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};  // Geometry::Constants<SQUARE> vertex order
  const double h[2] = {(hi[0] - lo[0]) / nx, (hi[1] - lo[1]) / ny};
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx; i++) {
      const size_t e = static_cast<size_t>(i) + static_cast<size_t>(nx) * j;
      for (int a = 0; a < 4; a++) {
        const int iv[2] = {i + QV[a][0], j + QV[a][1]};
        for (int d = 0; d < 2; d++) elem_xyz[(e * 4 + a) * 2 + d] = lo[d] + iv[d] * h[d];
        elem_verts[e * 4 + a] = (iv[0] % nv[0]) + nv[0] * (iv[1] % nv[1]);
      }
    }
  return TPSB_OK;
}

// Edges numbered by first appearance over (element, local edge); Elem2Inf = 64*le + 1 when the second element
// traverses the edge in the opposite direction (always, for consistently oriented quads), else + 0
// [MFEM Mesh::AddSegmentFaceElement / GenerateFaces].
extern "C" int tpsb_mk_build_faces2d(int num_elems, const int *elem_verts, int *face_el1, int *face_el2,
                                     int *face_inf1, int *face_inf2) {
  if (num_elems < 0 || !elem_verts) return -1;
  static const int EV[4][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}};  // Geometry::Constants<SQUARE>::Edges
  std::unordered_map<Key, int, KeyHash> table;
  table.reserve(static_cast<size_t>(num_elems) * 3);
  std::vector<int> base;
  int nfaces = 0;
  const bool fill = face_el1 && face_el2 && face_inf1 && face_inf2;
  for (int e = 0; e < num_elems; e++) {
    const int *v = &elem_verts[static_cast<size_t>(e) * 4];
    for (int le = 0; le < 4; le++) {
      const int v0 = v[EV[le][0]], v1 = v[EV[le][1]];
      const Key key{std::min(v0, v1), std::max(v0, v1), -1};
      auto it = table.find(key);
      if (it == table.end()) {
        table.emplace(key, nfaces);
        base.push_back(v0);
        base.push_back(v1);
        if (fill) {
          face_el1[nfaces] = e;
          face_el2[nfaces] = -1;
          face_inf1[nfaces] = 64 * le;
          face_inf2[nfaces] = -1;
        }
        nfaces++;
      } else if (fill) {
        const int f = it->second;
        if (face_el2[f] != -1) return -2;
        face_el2[f] = e;
        face_inf2[f] = 64 * le + ((base[2 * f] == v0 && base[2 * f + 1] == v1) ? 0 : 1);
      }
    }
  }
  return nfaces;
}

extern "C" int tpsb_mk_partition(const int n[3], const double lo[3], const double hi[3], const int periodic[3],
                                 const int procs[3], int rank, int order_mode, tpsb_mk_part_sizes *sizes,
                                 int *elem_verts, double *elem_xyz, int64_t *elem_gid, int *face_el1, int *face_el2,
                                 int *face_inf1, int *face_inf2, int *nbr_rank, int *send_offset, int *send_elems,
                                 int *recv_offset) {
  if (!n || !lo || !hi || !periodic || !procs || !sizes) return TPSB_EINVAL;
  const int nranks = procs[0] * procs[1] * procs[2];
  if (rank < 0 || rank >= nranks) return TPSB_EINVAL;
  int nv[3];
  for (int d = 0; d < 3; d++) {
    if (procs[d] < 1 || n[d] < procs[d]) return TPSB_EINVAL;
    if (periodic[d] && n[d] < 3) return TPSB_EINVAL;
    nv[d] = periodic[d] ? n[d] : n[d] + 1;
  }
  const int rc[3] = {rank % procs[0], (rank / procs[0]) % procs[1], rank / (procs[0] * procs[1])};
  int b0[3], b1[3];
  for (int d = 0; d < 3; d++) {
    b0[d] = static_cast<int>(static_cast<int64_t>(n[d]) * rc[d] / procs[d]);
    b1[d] = static_cast<int>(static_cast<int64_t>(n[d]) * (rc[d] + 1) / procs[d]);
  }
  auto owner_of = [&](const int e[3]) {
    int r[3];
    for (int d = 0; d < 3; d++) {
      // inverse of the even split above
      int q = static_cast<int>((static_cast<int64_t>(e[d]) * procs[d] + procs[d] - 1) / n[d]);
      while (q > 0 && static_cast<int64_t>(n[d]) * q / procs[d] > e[d]) q--;
      while (q + 1 < procs[d] && static_cast<int64_t>(n[d]) * (q + 1) / procs[d] <= e[d]) q++;
      r[d] = q;
    }
    return r[0] + procs[0] * (r[1] + procs[1] * r[2]);
  };
  auto gid_of = [&](const int e[3]) { return e[0] + static_cast<int64_t>(n[0]) * (e[1] + static_cast<int64_t>(n[1]) * e[2]); };
  // local elements in visiting order
  std::vector<int64_t> loc;
  const int B = order_mode ? 8 : (1 << 30);
  for (int bz = b0[2]; bz < b1[2]; bz += std::min(B, b1[2] - b0[2]))
    for (int by = b0[1]; by < b1[1]; by += std::min(B, b1[1] - b0[1]))
      for (int bx = b0[0]; bx < b1[0]; bx += std::min(B, b1[0] - b0[0]))
        for (int k = bz; k < std::min<int64_t>(static_cast<int64_t>(bz) + B, b1[2]); k++)
          for (int j = by; j < std::min<int64_t>(static_cast<int64_t>(by) + B, b1[1]); j++)
            for (int i = bx; i < std::min<int64_t>(static_cast<int64_t>(bx) + B, b1[0]); i++) {
              const int e[3] = {i, j, k};
              loc.push_back(gid_of(e));
            }
  const int NE = static_cast<int>(loc.size());
  std::unordered_map<int64_t, int> lid;
  lid.reserve(loc.size() * 2);
  for (int e = 0; e < NE; e++) lid[loc[e]] = e;
  // halo elements (owner, gid) and send list (peer, local gid)
  std::vector<std::pair<int, int64_t>> halo, send;
  for (int e = 0; e < NE; e++) {
    const int64_t g = loc[e];
    const int c[3] = {static_cast<int>(g % n[0]), static_cast<int>((g / n[0]) % n[1]),
                      static_cast<int>(g / (static_cast<int64_t>(n[0]) * n[1]))};
    for (int d = 0; d < 3; d++)
      for (int s = -1; s <= 1; s += 2) {
        int q[3] = {c[0], c[1], c[2]};
        q[d] += s;
        if (q[d] < 0 || q[d] >= n[d]) {
          if (!periodic[d]) continue;
          q[d] = (q[d] + n[d]) % n[d];
        }
        const int ow = owner_of(q);
        if (ow == rank) continue;
        halo.emplace_back(ow, gid_of(q));
        send.emplace_back(ow, g);
      }
  }
  std::sort(halo.begin(), halo.end());
  halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
  std::sort(send.begin(), send.end());
  send.erase(std::unique(send.begin(), send.end()), send.end());
  std::vector<int> peers;
  for (auto &h : halo)
    if (peers.empty() || peers.back() != h.first) peers.push_back(h.first);
  const int NEH = static_cast<int>(halo.size());
  sizes->num_elems = NE;
  sizes->num_nbr_elems = NEH;
  sizes->num_nbr_ranks = static_cast<int>(peers.size());
  sizes->num_send = static_cast<int>(send.size());
  // element tables
  std::vector<int> ev(static_cast<size_t>(NE + NEH) * 8);
  std::vector<double> xyz_tmp;
  double *xyz = elem_xyz;
  if (!xyz) {
    xyz_tmp.resize(static_cast<size_t>(NE + NEH) * 24);
    xyz = xyz_tmp.data();
  }
  const double h[3] = {(hi[0] - lo[0]) / n[0], (hi[1] - lo[1]) / n[1], (hi[2] - lo[2]) / n[2]};
  for (int e = 0; e < NE + NEH; e++) {
    const int64_t g = e < NE ? loc[e] : halo[e - NE].second;
    if (elem_gid) elem_gid[e] = g;
    const int c[3] = {static_cast<int>(g % n[0]), static_cast<int>((g / n[0]) % n[1]),
                      static_cast<int>(g / (static_cast<int64_t>(n[0]) * n[1]))};
    for (int a = 0; a < 8; a++) {
      int iv[3];
      for (int d = 0; d < 3; d++) {
        iv[d] = c[d] + tpsb::HEX_VERT[a][d];
        xyz[(static_cast<size_t>(e) * 8 + a) * 3 + d] = lo[d] + iv[d] * h[d];
      }
      ev[static_cast<size_t>(e) * 8 + a] = (iv[0] % nv[0]) + nv[0] * ((iv[1] % nv[1]) + nv[1] * (iv[2] % nv[2]));
    }
  }
  // faces over local + halo elements; keep those whose first element is local
  std::vector<int> f1(static_cast<size_t>(NE + NEH) * 6), f2(f1.size()), i1(f1.size()), i2(f1.size());
  const int nf_all = tpsb_mk_build_faces(NE + NEH, ev.data(), f1.data(), f2.data(), i1.data(), i2.data());
  if (nf_all < 0) return TPSB_EINVAL;
  int nf = 0;
  for (int f = 0; f < nf_all; f++) {
    if (f1[f] >= NE) continue;  // halo-only face
    if (face_el1) {
      face_el1[nf] = f1[f];
      face_el2[nf] = f2[f];
      face_inf1[nf] = i1[f];
      face_inf2[nf] = i2[f];
    }
    nf++;
  }
  sizes->num_faces = nf;
  if (elem_verts) std::copy(ev.begin(), ev.end(), elem_verts);
  if (nbr_rank && send_offset && send_elems && recv_offset) {
    size_t hp = 0, sp = 0;
    for (size_t p = 0; p < peers.size(); p++) {
      nbr_rank[p] = peers[p];
      recv_offset[p] = static_cast<int>(hp);
      send_offset[p] = static_cast<int>(sp);
      while (hp < halo.size() && halo[hp].first == peers[p]) hp++;
      while (sp < send.size() && send[sp].first == peers[p]) {
        send_elems[sp] = lid[send[sp].second];
        sp++;
      }
    }
    recv_offset[peers.size()] = static_cast<int>(hp);
    send_offset[peers.size()] = static_cast<int>(sp);
  }
  return TPSB_OK;
}

// ---- general element partitions (unstructured meshes: the O-grid of config C2, rotated / warped boxes) ----
// METIS 5 k-way partition of the element dual graph: what MFEM's Mesh::GeneratePartitioning(nprocs, 1) calls
// (src/M2ulPhyS.cpp:332,362).  The METIS library is the static one shipped with the CUDA toolkit (cuSOLVER's
// re-ordering dependency; no header is installed): idx_t = int64, real_t = float, probed with a known graph.
extern "C" int METIS_PartGraphKway(int64_t *nvtxs, int64_t *ncon, int64_t *xadj, int64_t *adjncy, int64_t *vwgt,
                                   int64_t *vsize, int64_t *adjwgt, int64_t *nparts, float *tpwgts, float *ubvec,
                                   int64_t *options, int64_t *edgecut, int64_t *part);
extern "C" int METIS_SetDefaultOptions(int64_t *options);

extern "C" int tpsb_mk_partition_metis(int num_elems, int num_faces, const int *face_el1, const int *face_el2, int nparts,
                                       int *elem_rank, int64_t *edge_cut) {
  if (num_elems <= 0 || nparts < 1 || !face_el1 || !face_el2 || !elem_rank) return TPSB_EINVAL;
  if (nparts == 1) {
    std::fill(elem_rank, elem_rank + num_elems, 0);
    if (edge_cut) *edge_cut = 0;
    return TPSB_OK;
  }
  std::vector<int64_t> xadj(static_cast<size_t>(num_elems) + 1, 0), adj;
  for (int f = 0; f < num_faces; f++)
    if (face_el2[f] >= 0 && face_el2[f] < num_elems && face_el2[f] != face_el1[f]) {
      xadj[face_el1[f] + 1]++;
      xadj[face_el2[f] + 1]++;
    }
  for (int e = 0; e < num_elems; e++) xadj[e + 1] += xadj[e];
  adj.resize(static_cast<size_t>(xadj[num_elems]));
  std::vector<int64_t> pos(xadj.begin(), xadj.end() - 1);
  for (int f = 0; f < num_faces; f++)
    if (face_el2[f] >= 0 && face_el2[f] < num_elems && face_el2[f] != face_el1[f]) {
      adj[static_cast<size_t>(pos[face_el1[f]]++)] = face_el2[f];
      adj[static_cast<size_t>(pos[face_el2[f]]++)] = face_el1[f];
    }
  int64_t nv = num_elems, ncon = 1, np = nparts, cut = 0, opts[40];
  METIS_SetDefaultOptions(opts);
  std::vector<int64_t> part(static_cast<size_t>(num_elems), 0);
  const int rc = METIS_PartGraphKway(&nv, &ncon, xadj.data(), adj.data(), nullptr, nullptr, nullptr, &np, nullptr, nullptr,
                                     opts, &cut, part.data());
  if (rc != 1) return TPSB_EINVAL;  // METIS_OK == 1
  for (int e = 0; e < num_elems; e++) elem_rank[e] = static_cast<int>(part[e]);
  if (edge_cut) *edge_cut = cut;
  return TPSB_OK;
}

// Recursive coordinate bisection of the element centroids (nparts a power of two or not: the split is proportional).
extern "C" int tpsb_mk_partition_rcb_dim(int dim, int num_elems, const double *elem_xyz, int nparts, int *elem_rank) {
  if ((dim != 2 && dim != 3) || num_elems <= 0 || nparts < 1 || !elem_xyz || !elem_rank) return TPSB_EINVAL;
  const int nv = 1 << dim;
  std::vector<double> cen(static_cast<size_t>(num_elems) * 3, 0.0);
  for (int e = 0; e < num_elems; e++)
    for (int a = 0; a < nv; a++)
      for (int d = 0; d < dim; d++) cen[static_cast<size_t>(e) * 3 + d] += elem_xyz[(static_cast<size_t>(e) * nv + a) * dim + d] / nv;
  std::vector<int> ids(num_elems);
  for (int e = 0; e < num_elems; e++) ids[e] = e;
  struct Job {
    int b, n, r0, np;
  };
  std::vector<Job> stack{{0, num_elems, 0, nparts}};
  while (!stack.empty()) {
    const Job j = stack.back();
    stack.pop_back();
    if (j.np == 1) {
      for (int t = j.b; t < j.b + j.n; t++) elem_rank[ids[t]] = j.r0;
      continue;
    }
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int t = j.b; t < j.b + j.n; t++)
      for (int d = 0; d < 3; d++) {
        lo[d] = std::min(lo[d], cen[static_cast<size_t>(ids[t]) * 3 + d]);
        hi[d] = std::max(hi[d], cen[static_cast<size_t>(ids[t]) * 3 + d]);
      }
    int ax = 0;
    for (int d = 1; d < 3; d++)
      if (hi[d] - lo[d] > hi[ax] - lo[ax]) ax = d;
    const int npl = j.np / 2, nl = static_cast<int>(static_cast<int64_t>(j.n) * npl / j.np);
    std::nth_element(ids.begin() + j.b, ids.begin() + j.b + nl, ids.begin() + j.b + j.n, [&](int x, int y) {
      const double cx = cen[static_cast<size_t>(x) * 3 + ax], cy = cen[static_cast<size_t>(y) * 3 + ax];
      return cx < cy || (cx == cy && x < y);
    });
    stack.push_back({j.b, nl, j.r0, npl});
    stack.push_back({j.b + nl, j.n - nl, j.r0 + npl, j.np - npl});
  }
  return TPSB_OK;
}
extern "C" int tpsb_mk_partition_rcb(int num_elems, const double *elem_xyz, int nparts, int *elem_rank) {
  return tpsb_mk_partition_rcb_dim(3, num_elems, elem_xyz, nparts, elem_rank);
}

// This rank's piece of a GLOBAL hexahedral mesh under an arbitrary element -> rank map: ParMesh's local numbering
// (local elements in global order first, then the face-neighbour elements grouped by owner and sorted by global id,
// the order in which the owner sends them; src/M2ulPhyS.cpp:421, ExchangeFaceNbrData).  Two calls: sizes, then fill.
//   face_gface[f]  global face of local face f (so the caller can carry boundary attributes over)
extern "C" int tpsb_mk_build_faces2d(int num_elems, const int *elem_verts, int *face_el1, int *face_el2, int *face_inf1,
                                     int *face_inf2);
extern "C" int tpsb_mk_partition_general_dim(int dim, int num_elems, const int *elem_verts, const double *elem_xyz, int num_faces,
                                             const int *gface_el1, const int *gface_el2, const int *elem_rank, int rank,
                                             tpsb_mk_part_sizes *sizes, int *l_elem_verts, double *l_elem_xyz, int64_t *elem_gid,
                                             int *face_el1, int *face_el2, int *face_inf1, int *face_inf2, int *face_gface,
                                             int *nbr_rank, int *send_offset, int *send_elems, int *recv_offset) {
  if ((dim != 2 && dim != 3) || num_elems <= 0 || !elem_verts || !gface_el1 || !gface_el2 || !elem_rank || !sizes) return TPSB_EINVAL;
  const int nv = 1 << dim, nfe = 2 * dim;
  static const int QUAD_EDGE[4][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}};  // Geometry::Constants<SQUARE>::Edges
  auto face_key = [&](const int *v, int lf) {
    int s[4] = {0, 0, 0, 0};
    if (dim == 3) {
      for (int k = 0; k < 4; k++) s[k] = v[tpsb::HEX_FACE_VERT[lf][k]];
      std::sort(s, s + 4);
      return Key{s[0], s[1], s[2]};
    }
    const int a = v[QUAD_EDGE[lf][0]], b = v[QUAD_EDGE[lf][1]];
    return Key{std::min(a, b), std::max(a, b), -1};
  };
  std::vector<int> loc, lid(static_cast<size_t>(num_elems), -1);
  for (int e = 0; e < num_elems; e++)
    if (elem_rank[e] == rank) {
      lid[e] = static_cast<int>(loc.size());
      loc.push_back(e);
    }
  const int NE = static_cast<int>(loc.size());
  if (NE == 0) return TPSB_EINVAL;
  std::vector<std::pair<int, int>> halo, send;  // (owner, global element), (peer, global element)
  for (int f = 0; f < num_faces; f++) {
    const int e1 = gface_el1[f], e2 = gface_el2[f];
    if (e2 < 0) continue;
    const int r1 = elem_rank[e1], r2 = elem_rank[e2];
    if (r1 == r2) continue;
    if (r1 == rank) halo.emplace_back(r2, e2), send.emplace_back(r2, e1);
    if (r2 == rank) halo.emplace_back(r1, e1), send.emplace_back(r1, e2);
  }
  std::sort(halo.begin(), halo.end());
  halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
  std::sort(send.begin(), send.end());
  send.erase(std::unique(send.begin(), send.end()), send.end());
  std::vector<int> peers;
  for (auto &h : halo)
    if (peers.empty() || peers.back() != h.first) peers.push_back(h.first);
  const int NEH = static_cast<int>(halo.size());
  std::vector<int> ev(static_cast<size_t>(NE + NEH) * nv);
  for (int e = 0; e < NE + NEH; e++) {
    const int g = e < NE ? loc[e] : halo[e - NE].second;
    std::copy(&elem_verts[static_cast<size_t>(g) * nv], &elem_verts[static_cast<size_t>(g) * nv] + nv, &ev[static_cast<size_t>(e) * nv]);
    if (elem_gid) elem_gid[e] = g;
    if (l_elem_xyz && elem_xyz)
      std::copy(&elem_xyz[static_cast<size_t>(g) * nv * dim], &elem_xyz[static_cast<size_t>(g) * nv * dim] + nv * dim,
                &l_elem_xyz[static_cast<size_t>(e) * nv * dim]);
  }
  std::vector<int> f1(static_cast<size_t>(NE + NEH) * nfe), f2(f1.size()), i1(f1.size()), i2(f1.size());
  const int nf_all = dim == 3 ? tpsb_mk_build_faces(NE + NEH, ev.data(), f1.data(), f2.data(), i1.data(), i2.data())
                              : tpsb_mk_build_faces2d(NE + NEH, ev.data(), f1.data(), f2.data(), i1.data(), i2.data());
  if (nf_all < 0) return TPSB_EINVAL;
  // global face of (global element, local face): faces are found again through their vertex key
  std::unordered_map<Key, int, KeyHash> gtab;
  if (face_gface) {
    gtab.reserve(static_cast<size_t>(num_faces) * 2);
    for (int e = 0; e < num_elems; e++)
      for (int lf = 0; lf < nfe; lf++) {
        const Key key = face_key(&elem_verts[static_cast<size_t>(e) * nv], lf);
        if (gtab.find(key) == gtab.end()) gtab.emplace(key, static_cast<int>(gtab.size()));  // first appearance == MFEM face number
      }
  }
  int nf = 0;
  for (int f = 0; f < nf_all; f++) {
    if (f1[f] >= NE) continue;  // face between face-neighbour elements only
    if (face_el1) {
      face_el1[nf] = f1[f];
      face_el2[nf] = f2[f];
      face_inf1[nf] = i1[f];
      face_inf2[nf] = i2[f];
      if (face_gface) face_gface[nf] = gtab.at(face_key(&ev[static_cast<size_t>(f1[f]) * nv], i1[f] / 64));
    }
    nf++;
  }
  sizes->num_elems = NE;
  sizes->num_nbr_elems = NEH;
  sizes->num_faces = nf;
  sizes->num_nbr_ranks = static_cast<int>(peers.size());
  sizes->num_send = static_cast<int>(send.size());
  if (l_elem_verts) std::copy(ev.begin(), ev.end(), l_elem_verts);
  if (nbr_rank && send_offset && send_elems && recv_offset) {
    size_t hp = 0, sp = 0;
    for (size_t p = 0; p < peers.size(); p++) {
      nbr_rank[p] = peers[p];
      recv_offset[p] = static_cast<int>(hp);
      send_offset[p] = static_cast<int>(sp);
      while (hp < halo.size() && halo[hp].first == peers[p]) hp++;
      while (sp < send.size() && send[sp].first == peers[p]) {
        send_elems[sp] = lid[send[sp].second];
        sp++;
      }
    }
    recv_offset[peers.size()] = static_cast<int>(hp);
    send_offset[peers.size()] = static_cast<int>(sp);
  }
  return TPSB_OK;
}
extern "C" int tpsb_mk_partition_general(int num_elems, const int *elem_verts, const double *elem_xyz, int num_faces,
                                         const int *gface_el1, const int *gface_el2, const int *elem_rank, int rank,
                                         tpsb_mk_part_sizes *sizes, int *l_elem_verts, double *l_elem_xyz, int64_t *elem_gid,
                                         int *face_el1, int *face_el2, int *face_inf1, int *face_inf2, int *face_gface,
                                         int *nbr_rank, int *send_offset, int *send_elems, int *recv_offset) {
  return tpsb_mk_partition_general_dim(3, num_elems, elem_verts, elem_xyz, num_faces, gface_el1, gface_el2, elem_rank, rank, sizes,
                                       l_elem_verts, l_elem_xyz, elem_gid, face_el1, face_el2, face_inf1, face_inf2, face_gface,
                                       nbr_rank, send_offset, send_elems, recv_offset);
}

// Flattened RefTables for the tests: [np, nq, xn(np), wn(np), D(np*np), lb(2*np), xq(nq), wq(nq), P(nq*np),
// face_base(6*np^2), face_cstride(6), face_side(6), perm(8*np^2), iperm(8*np^2)]
extern "C" int tpsb_get_ref_tables(int order, double *out, int cap) {
  tpsb::RefTables T;
  if (!tpsb::build_ref_tables(order, T)) return -1;
  std::vector<double> v;
  const int np = T.np, nq = T.nq;
  v.push_back(np);
  v.push_back(nq);
  for (int i = 0; i < np; i++) v.push_back(T.xn[i]);
  for (int i = 0; i < np; i++) v.push_back(T.wn[i]);
  for (int i = 0; i < np; i++)
    for (int m = 0; m < np; m++) v.push_back(T.D[i][m]);
  for (int s = 0; s < 2; s++)
    for (int c = 0; c < np; c++) v.push_back(T.lb[s][c]);
  for (int i = 0; i < nq; i++) v.push_back(T.xq[i]);
  for (int i = 0; i < nq; i++) v.push_back(T.wq[i]);
  for (int a = 0; a < nq; a++)
    for (int m = 0; m < np; m++) v.push_back(T.P[a][m]);
  for (int lf = 0; lf < 6; lf++)
    for (int ab = 0; ab < np * np; ab++) v.push_back(T.face_base[lf][ab]);
  for (int lf = 0; lf < 6; lf++) v.push_back(T.face_cstride[lf]);
  for (int lf = 0; lf < 6; lf++) v.push_back(T.face_side[lf]);
  for (int o = 0; o < 8; o++)
    for (int ab = 0; ab < np * np; ab++) v.push_back(T.perm[o][ab]);
  for (int o = 0; o < 8; o++)
    for (int ab = 0; ab < np * np; ab++) v.push_back(T.iperm[o][ab]);
  if (!out || cap < static_cast<int>(v.size())) return -2;
  std::copy(v.begin(), v.end(), out);
  return static_cast<int>(v.size());
}
