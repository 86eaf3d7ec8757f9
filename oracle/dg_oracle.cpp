// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Never linked into the product;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load it.
//
// CPU restatement of the reference's *CPU* branch of RHSoperator::Mult
// (src/rhs_operator.cpp:343-464) for tensor-product DG meshes (quadrilaterals in 2-D, hexahedra in
// 3-D; Gauss-Legendre or Gauss-Lobatto nodes and rules), using the same DENSE per-element operators
// the reference builds (Me_inv, Ke, the (F,grad w) element blocks of Aflux) and the same per-face /
// per-quadrature-point loops -- deliberately NOT sum-factorised, so it is independent of the CUDA
// kernels it checks.
//
// MFEM (third party, >=4.4, absent from /root/reference) supplies the mesh/FE conventions
// the reference relies on; they are restated here from MFEM's documented behaviour
// (SURVEY.md Appendix B):
//   vertex / face-vertex / orientation tables of SQUARE and CUBE, FaceElementTransformations
//   Loc1/Loc2, CalcOrtho, IntegrationRules, L2 tensor basis ordering, RK4Solver tableau.
// PARITY PINNED on a value the reference holds: the run of its regression test
// test/argon_minimal.binary.test (1000 RK4 steps, analytic Ar / Ar+ diffusion wave, tolerance 2e-4) re-expressed
// in tests/binary_mixture_case.py lands at 7.6e-5 with this operator and the reference's physics object code.  Bit-level
// identity of the index maps with a live MFEM build stays unverifiable here (no MFEM, binary goldens are LFS pointers).
// Per-point physics goes through orc::Physics (port or the reference's own object code).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "orc_basis.hpp"
#include "orc_physics.hpp"

namespace orc {

#define SLIPFN static inline
// WallBC::computeSlipWallFlux (wallBC.cpp:326-428): velocity in a wall-aligned basis (outward unit normal, an arbitrary
// tangent built around the dominant normal component, their cross product), normal component mirrored, transformed
// back with the inverse basis matrix.  [MFEM CalcInverse: adjugate / determinant for 2 x 2 and 3 x 3.]
SLIPFN void slip_mirror_velocity(int dim, const double *normal, const double *vel, double *nVel) {
  const double sml = 1.0e-15;
  double unitNorm[3] = {0, 0, 0}, tangent1[3] = {0, 0, 0}, tangent2[3] = {0, 0, 0};
  double normN = 0.;
  for (int d = 0; d < dim; d++) normN += normal[d] * normal[d];
  normN = sqrt(fmax(normN, sml));
  for (int d = 0; d < dim; d++) unitNorm[d] = normal[d] * (1. / normN);
  int dir = 0;
  if (dim == 3) {
    if (fabs(unitNorm[0]) >= fabs(unitNorm[1]) && fabs(unitNorm[0]) >= fabs(unitNorm[2])) dir = 0;
    if (fabs(unitNorm[1]) >= fabs(unitNorm[0]) && fabs(unitNorm[1]) >= fabs(unitNorm[2])) dir = 1;
    if (fabs(unitNorm[2]) >= fabs(unitNorm[0]) && fabs(unitNorm[2]) >= fabs(unitNorm[1])) dir = 2;
  } else {
    if (fabs(unitNorm[0]) >= fabs(unitNorm[1])) dir = 0;
    if (fabs(unitNorm[1]) >= fabs(unitNorm[0])) dir = 1;
  }
  const int next_dir = (dir + 1) % dim, previous_dir = (dir + 2) % dim;
  tangent1[next_dir] = +1.;
  tangent1[previous_dir] = -1.;
  tangent1[dir] = unitNorm[previous_dir] * tangent1[previous_dir] + unitNorm[next_dir] * tangent1[next_dir];
  tangent1[dir] *= -1. / unitNorm[dir];
  double mod = 0.;
  for (int d = 0; d < dim; d++) mod += tangent1[d] * tangent1[d];
  for (int d = 0; d < dim; d++) tangent1[d] *= 1. / fmax(sqrt(mod), sml);
  if (dim == 3) {
    tangent2[0] = +(unitNorm[1] * tangent1[2] - unitNorm[2] * tangent1[1]);
    tangent2[1] = -(unitNorm[0] * tangent1[2] - unitNorm[2] * tangent1[0]);
    tangent2[2] = +(unitNorm[0] * tangent1[1] - unitNorm[1] * tangent1[0]);
    mod = 0.;
    for (int d = 0; d < dim; d++) mod += tangent2[d] * tangent2[d];
    for (int d = 0; d < dim; d++) tangent2[d] *= 1. / fmax(sqrt(mod), sml);
  }
  double M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, inv[3][3], w[3] = {0, 0, 0};
  for (int d = 0; d < dim; d++) {
    M[0][d] = unitNorm[d];
    M[1][d] = tangent1[d];
    if (dim == 3) M[2][d] = tangent2[d];
  }
  for (int r = 0; r < dim; r++)
    for (int d = 0; d < dim; d++) w[r] += M[r][d] * vel[d];
  w[0] = -w[0];  // mirror the normal component
  if (dim == 2) {
    const double t = 1.0 / (M[0][0] * M[1][1] - M[0][1] * M[1][0]);
    inv[0][0] = M[1][1] * t;
    inv[0][1] = -M[0][1] * t;
    inv[1][0] = -M[1][0] * t;
    inv[1][1] = M[0][0] * t;
  } else {
    const double det = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
                       M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
    const double t = 1.0 / det;
    inv[0][0] = (M[1][1] * M[2][2] - M[1][2] * M[2][1]) * t;
    inv[0][1] = (M[0][2] * M[2][1] - M[0][1] * M[2][2]) * t;
    inv[0][2] = (M[0][1] * M[1][2] - M[0][2] * M[1][1]) * t;
    inv[1][0] = (M[1][2] * M[2][0] - M[1][0] * M[2][2]) * t;
    inv[1][1] = (M[0][0] * M[2][2] - M[0][2] * M[2][0]) * t;
    inv[1][2] = (M[0][2] * M[1][0] - M[0][0] * M[1][2]) * t;
    inv[2][0] = (M[1][0] * M[2][1] - M[1][1] * M[2][0]) * t;
    inv[2][1] = (M[0][1] * M[2][0] - M[0][0] * M[2][1]) * t;
    inv[2][2] = (M[0][0] * M[1][1] - M[0][1] * M[1][0]) * t;
  }
  for (int r = 0; r < dim; r++) {
    nVel[r] = 0.;
    for (int d = 0; d < dim; d++) nVel[r] += inv[r][d] * w[d];
  }
}

// ---- MFEM geometry constants [MFEM fem/geom.cpp: Geometry::Constants<CUBE/SQUARE/SEGMENT>] ----
static const double HEX_VERT[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0},
                                      {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
static const int HEX_FACE_VERT[6][4] = {{3, 2, 1, 0}, {0, 1, 5, 4}, {1, 2, 6, 5},
                                        {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
static const int QUAD_ORIENT[8][4] = {{0, 1, 2, 3}, {0, 3, 2, 1}, {1, 2, 3, 0}, {1, 0, 3, 2},
                                      {2, 3, 0, 1}, {2, 1, 0, 3}, {3, 0, 1, 2}, {3, 2, 1, 0}};
static const double QUAD_VERT[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};
static const int QUAD_EDGE_VERT[4][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}};
static const int SEG_ORIENT[2][2] = {{0, 1}, {1, 0}};

static inline void cross3(const double *a, const double *b, double *c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// Dense in-place inverse, Gauss-Jordan with partial pivoting (stands in for
// DenseMatrix::Invert, src/rhs_operator.cpp:187). Row-major n x n.
static void invert_dense(std::vector<double> &a, int n) {
  std::vector<double> inv(static_cast<size_t>(n) * n, 0.0);
  for (int i = 0; i < n; i++) inv[i * n + i] = 1.0;
  for (int c = 0; c < n; c++) {
    int piv = c;
    double best = fabs(a[c * n + c]);
    for (int r = c + 1; r < n; r++)
      if (fabs(a[r * n + c]) > best) {
        best = fabs(a[r * n + c]);
        piv = r;
      }
    if (piv != c) {
      for (int k = 0; k < n; k++) {
        std::swap(a[c * n + k], a[piv * n + k]);
        std::swap(inv[c * n + k], inv[piv * n + k]);
      }
    }
    double d = 1.0 / a[c * n + c];
    for (int k = 0; k < n; k++) {
      a[c * n + k] *= d;
      inv[c * n + k] *= d;
    }
    for (int r = 0; r < n; r++) {
      if (r == c) continue;
      double f = a[r * n + c];
      if (f == 0.0) continue;
      for (int k = 0; k < n; k++) {
        a[r * n + k] -= f * a[c * n + k];
        inv[r * n + k] -= f * inv[c * n + k];
      }
    }
  }
  a.swap(inv);
}

// smallest singular value of a dim x dim matrix, column-major with leading dimension dim
// (Mesh::GetElementSize(e, 1) -> J.CalcSingularvalue(dim-1))
static double min_singular(const double *J, int dim) {
  double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < dim; i++)
    for (int j = 0; j < dim; j++) {
      double s = 0;
      for (int k = 0; k < dim; k++) s += J[k + dim * i] * J[k + dim * j];
      A[i][j] = s;
    }
  // cyclic Jacobi on symmetric A
  for (int sweep = 0; sweep < 50; sweep++) {
    double off = 0;
    for (int p = 0; p < dim; p++)
      for (int q = p + 1; q < dim; q++) off += fabs(A[p][q]);
    if (off < 1e-300) break;
    for (int p = 0; p < dim - 1; p++)
      for (int q = p + 1; q < dim; q++) {
        if (fabs(A[p][q]) < 1e-300) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < dim; k++) {
          double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < dim; k++) {
          double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
      }
  }
  double m = A[0][0];
  for (int i = 1; i < dim; i++) m = std::min(m, A[i][i]);
  return sqrt(std::max(m, 0.0));
}

struct Oracle {
  int dim = 3, p = 0, np = 0, dof = 0, NE = 0, NF = 0, neq = 5, nvel = 3, nthreads = 1;
  int basis_type = 0, int_rule = 0;  // flow/basisType, flow/integrationRule: 0 Gauss-Legendre, 1 Gauss-Lobatto
  int nv = 8;                        // vertices per element
  long N = 0;                        // vfes->GetNDofs()
  OrcPhysParams phys;
  Physics *ph = nullptr;
  std::vector<double> vx;  // [NE][nv][dim]
  std::vector<int> f_el1, f_el2, f_inf1, f_inf2;
  std::vector<std::vector<int>> el_faces;   // interior faces per element, ascending (element_to_faces)
  std::vector<std::vector<int>> el_bfaces;  // boundary faces per element
  std::vector<int> f_attr;                  // boundary attribute per face (0 on interior faces)
  std::vector<OrcBc> bcs;                   // BCintegrator's attribute -> boundary condition maps
  int use_bc_in_grad = 0;                   // boundaryConditions/useBCinGrad (src/M2ulPhyS.cpp:3480)
  std::vector<double> nodes1d;
  int nqv1 = 0, nqf1 = 0, nqv = 0, nqf = 0;
  std::vector<double> qxv, qwv, qxf, qwf;
  std::vector<double> Me_inv, Me_inv_rad, Ke, Kfl;  // per element dense
  bool axisym = false;
  std::vector<double> elSize;           // per element delta = h_min / order
  std::vector<double> distance;         // nodal wall distance (M2ulPhyS::distance_, src/M2ulPhyS.cpp:265-283); empty: 0
  std::vector<double> nodeXYZ;          // [NE][dof][dim]
  std::vector<OrcForcing> forcings;     // ForcingTerms registered by orc_add_forcing, applied in order after Me^-1
  std::map<int, std::vector<double>> shapeTab;  // inf code -> [nqf][dof]
  // work
  std::vector<double> Up, gradUp;
  double max_char_speed = 0;
  const double *sol_view = nullptr;  // the solution grid function U_ the forcing terms read (parity trap 1)

  // ---- element geometry: multilinear map from the 2^dim vertices (mesh nodes of order 1) ----
  double vref(int a, int d) const { return dim == 3 ? HEX_VERT[a][d] : QUAD_VERT[a][d]; }
  void elem_map(int e, const double *xi, double *x, double *J) const {
    const double *v = &vx[static_cast<size_t>(e) * nv * dim];
    double Nn[8], dN[8][3];
    for (int a = 0; a < nv; a++) {
      double f[3], g[3];
      for (int d = 0; d < dim; d++) {
        f[d] = vref(a, d) ? xi[d] : 1 - xi[d];
        g[d] = vref(a, d) ? 1 : -1;
      }
      double prod = 1;
      for (int d = 0; d < dim; d++) prod *= f[d];
      Nn[a] = prod;
      for (int d = 0; d < dim; d++) {
        double q = g[d];
        for (int d2 = 0; d2 < dim; d2++)
          if (d2 != d) q *= f[d2];
        dN[a][d] = q;
      }
    }
    if (x)
      for (int i = 0; i < dim; i++) {
        double s = 0;
        for (int a = 0; a < nv; a++) s += Nn[a] * v[a * dim + i];
        x[i] = s;
      }
    if (J)  // column-major J(i,j) = dx_i/dxi_j at J[i + dim*j]
      for (int i = 0; i < dim; i++)
        for (int j = 0; j < dim; j++) {
          double s = 0;
          for (int a = 0; a < nv; a++) s += dN[a][j] * v[a * dim + i];
          J[i + dim * j] = s;
        }
  }
  double det(const double *J) const {
    if (dim == 2) return J[0] * J[3] - J[2] * J[1];
    return J[0] * (J[4] * J[8] - J[5] * J[7]) - J[3] * (J[1] * J[8] - J[2] * J[7]) + J[6] * (J[1] * J[5] - J[2] * J[4]);
  }
  // adjugate, column-major: adj(J) = det(J) * inv(J)
  void adj(const double *J, double *A) const {
    if (dim == 2) {
      A[0] = J[3];
      A[1] = -J[1];
      A[2] = -J[2];
      A[3] = J[0];
      return;
    }
    A[0 + 3 * 0] = J[4] * J[8] - J[7] * J[5];
    A[0 + 3 * 1] = J[6] * J[5] - J[3] * J[8];
    A[0 + 3 * 2] = J[3] * J[7] - J[6] * J[4];
    A[1 + 3 * 0] = J[7] * J[2] - J[1] * J[8];
    A[1 + 3 * 1] = J[0] * J[8] - J[6] * J[2];
    A[1 + 3 * 2] = J[6] * J[1] - J[0] * J[7];
    A[2 + 3 * 0] = J[1] * J[5] - J[4] * J[2];
    A[2 + 3 * 1] = J[3] * J[2] - J[0] * J[5];
    A[2 + 3 * 2] = J[0] * J[4] - J[3] * J[1];
  }
  // L2 tensor basis, lexicographic x fastest; dshape [dof][dim] row-major or NULL
  void calc_shape(const double *xi, double *shape, double *dshape) const {
    double v[3][16], d[3][16];
    for (int a = 0; a < dim; a++) lagrange(nodes1d, xi[a], v[a], d[a]);
    for (int n = 0; n < dof; n++) {
      int idx[3] = {n % np, (n / np) % np, n / (np * np)};
      double s = 1;
      for (int a = 0; a < dim; a++) s *= v[a][idx[a]];
      shape[n] = s;
      if (dshape)
        for (int a = 0; a < dim; a++) {
          double q = d[a][idx[a]];
          for (int b = 0; b < dim; b++)
            if (b != a) q *= v[b][idx[b]];
          dshape[n * dim + a] = q;
        }
    }
  }
  // [MFEM Mesh::GetLocalQuadToHexTransformation / GetLocalSegToQuadTransformation]: face reference point
  // -> element reference point for info code inf = 64*local_face + orientation; dloc = d(xi)/d(s[,t])
  // (dim x (dim-1), column-major).
  void loc_map(int inf, const double *st, double *xi, double *dloc) const {
    if (dim == 3) {
      const double s = st[0], t = st[1];
      const int *hv = HEX_FACE_VERT[inf / 64];
      const int *qo = QUAD_ORIENT[inf % 64];
      const double Nq[4] = {(1 - s) * (1 - t), s * (1 - t), s * t, (1 - s) * t};
      const double dNs[4] = {-(1 - t), (1 - t), t, -t};
      const double dNt[4] = {-(1 - s), -s, s, (1 - s)};
      for (int i = 0; i < 3; i++) {
        double a = 0, b = 0, c = 0;
        for (int j = 0; j < 4; j++) {
          const double vj = HEX_VERT[hv[qo[j]]][i];
          a += Nq[j] * vj;
          b += dNs[j] * vj;
          c += dNt[j] * vj;
        }
        xi[i] = a;
        if (dloc) {
          dloc[i + 3 * 0] = b;
          dloc[i + 3 * 1] = c;
        }
      }
    } else {
      const double s = st[0];
      const int *ev = QUAD_EDGE_VERT[inf / 64];
      const int *so = SEG_ORIENT[inf % 64];
      for (int i = 0; i < 2; i++) {
        const double v0 = QUAD_VERT[ev[so[0]]][i], v1 = QUAD_VERT[ev[so[1]]][i];
        xi[i] = (1 - s) * v0 + s * v1;
        if (dloc) dloc[i] = v1 - v0;
      }
    }
  }
  void face_point(int q, double *st) const {
    st[0] = qxf[q % nqf1];
    st[1] = dim == 3 ? qxf[q / nqf1] : 0.0;
  }
  double face_weight(int q) const { return dim == 3 ? qwf[q % nqf1] * qwf[q / nqf1] : qwf[q]; }
  const std::vector<double> &shape_table(int inf) {
    auto it = shapeTab.find(inf);
    if (it != shapeTab.end()) return it->second;
    std::vector<double> tab(static_cast<size_t>(nqf) * dof);
    for (int q = 0; q < nqf; q++) {
      double st[2], xi[3];
      face_point(q, st);
      loc_map(inf, st, xi, nullptr);
      calc_shape(xi, &tab[static_cast<size_t>(q) * dof], nullptr);
    }
    return shapeTab.emplace(inf, std::move(tab)).first->second;
  }
  // face geometry at quadrature point q of face f: CalcOrtho(Tr.Jacobian()) and Tr.Transform
  void face_geom(int f, int q, double *nor, double *xyz) const {
    double st[2], xi[3], dloc[6], J[9], Jf[6];
    face_point(q, st);
    loc_map(f_inf1[f], st, xi, dloc);
    elem_map(f_el1[f], xi, xyz, J);
    for (int i = 0; i < dim; i++)
      for (int c = 0; c < dim - 1; c++) {
        double a = 0;
        for (int k = 0; k < dim; k++) a += J[i + dim * k] * dloc[k + dim * c];
        Jf[i + dim * c] = a;
      }
    if (dim == 3) {
      cross3(&Jf[0], &Jf[3], nor);
    } else {  // [MFEM CalcOrtho, 2x1 Jacobian]: n = (dy/ds, -dx/ds)
      nor[0] = Jf[1];
      nor[1] = -Jf[0];
    }
  }

  // ---- boundary conditions (restated from src/BCintegrator.cpp, wallBC.cpp, inletBC.cpp, outletBC.cpp) ----
  const OrcBc *bc_of_face(int f) const {
    for (const OrcBc &b : bcs)
      if (b.attr == f_attr[f]) return &b;
    return nullptr;
  }
  void set_bcs(const int *face_attr, int nbc, const OrcBc *b, int use_in_grad) {
    f_attr.assign(face_attr, face_attr + NF);
    bcs.assign(b, b + nbc);
    use_bc_in_grad = use_in_grad;
    el_bfaces.assign(NE, {});
    for (int f = 0; f < NF; f++)
      if (f_el2[f] < 0) {
        el_bfaces[f_el1[f]].push_back(f);
        shape_table(f_inf1[f]);
      }
    setup_nr();
  }
  // ---- non-reflecting / mass-flow inlets and outlets (dry air): the stateful part of InletBC / OutletBC ----
  // Per patch: boundaryU (a conserved state per boundary quadrature point, advanced by every flux evaluation with the
  // current time step), meanUp (mean of the interpolated primitives over the patch's points, refreshed by updateMean at
  // every Mult: src/rhs_operator.cpp:364), tangent1 (first to second quadrature point of the patch's first face, or given),
  // the patch area (BoundaryCondition::aggregateArea, src/BoundaryCondition.cpp:59-81).
  // OrcBc::data: inlet {rho, u, v, w}, outlet {p | mass flow}; data[8] = refLength, data[9..11] = tangent1 (all zero: derive).
  struct NrState {
    std::vector<int> faces;
    std::map<int, int> face_off;  // face -> first point
    std::vector<double> boundaryU;
    double meanUp[16];
    double tangent1[3];
    double area = 0.0;
    bool init = false;
  };
  mutable std::map<int, NrState> nr;  // by index into bcs
  double bc_dt = 0.0;                 // BoundaryCondition::dt (a reference to M2ulPhyS::dt)
  static bool is_nr(const OrcBc &b) {
    return (b.kind == 0 && (b.type == 6 || b.type == 7)) || (b.kind == 1 && b.type >= 2 && b.type <= 4);
  }
  void setup_nr() {
    nr.clear();
    for (size_t i = 0; i < bcs.size(); i++) {
      if (!is_nr(bcs[i])) continue;
      NrState st;
      int off = 0;
      for (int f = 0; f < NF; f++)
        if (f_el2[f] < 0 && f_attr[f] == bcs[i].attr) {
          st.faces.push_back(f);
          st.face_off[f] = off;
          off += nqf;
        }
      st.boundaryU.assign(static_cast<size_t>(off) * neq, 0.0);
      for (int d = 0; d < 3; d++) st.tangent1[d] = bcs[i].data[9 + d];
      const double tm = st.tangent1[0] * st.tangent1[0] + st.tangent1[1] * st.tangent1[1] + st.tangent1[2] * st.tangent1[2];
      if (tm == 0.0 && !st.faces.empty()) {  // OutletBC constructor (src/outletBC.cpp:161-176): coords of points 0 and 1
        double n0[3], x0[3] = {0, 0, 0}, x1[3] = {0, 0, 0};
        face_geom(st.faces[0], 0, n0, x0);
        face_geom(st.faces[0], 1, n0, x1);
        double m = 0;
        for (int d = 0; d < dim; d++) m += (x1[d] - x0[d]) * (x1[d] - x0[d]);
        for (int d = 0; d < dim; d++) st.tangent1[d] = (x1[d] - x0[d]) * (1. / sqrt(m));
      }
      for (int f : st.faces)
        for (int q = 0; q < nqf; q++) {
          double nor[3], xyz[3], m = 0;
          face_geom(f, q, nor, xyz);
          for (int d = 0; d < dim; d++) m += nor[d] * nor[d];
          st.area += sqrt(m) * face_weight(q);
        }
      nr[static_cast<int>(i)] = st;
    }
  }
  // InletBC::updateMean / OutletBC::updateMean, CPU branch (src/inletBC.cpp:482-564, src/outletBC.cpp:470-561)
  void update_bc_mean() {
    for (auto &kv : nr) {
      NrState &st = kv.second;
      double sum[16] = {0};
      int nb = 0;
      for (int f : st.faces) {
        const int e1 = f_el1[f];
        const std::vector<double> &sh1 = shapeTab.at(f_inf1[f]);
        for (int q = 0; q < nqf; q++) {
          const double *s1 = &sh1[static_cast<size_t>(q) * dof];
          double iUp[16], iState[16];
          for (int eq = 0; eq < neq; eq++) {
            double a = 0.;
            for (int k = 0; k < dof; k++) a += s1[k] * Up[static_cast<size_t>(e1) * dof + k + eq * N];
            sum[eq] += a;
            iUp[eq] = a;
          }
          if (!st.init) {
            ph->cons(iUp, iState);
            for (int eq = 0; eq < neq; eq++) st.boundaryU[static_cast<size_t>(nb) * neq + eq] = iState[eq];
          }
          nb++;
        }
      }
      for (int eq = 0; eq < neq; eq++) st.meanUp[eq] = sum[eq] * (1. / static_cast<double>(nb));
      st.init = true;
    }
  }
  // InletBC::subsonicNonReflectingDensityVelocity (src/inletBC.cpp:576-727; SUB_DENS_VEL_NR, SUB_VEL_CONST_ENT) and
  // OutletBC::subsonicNonReflectingPressure / subsonicNonRefMassFlow / subsonicNonRefPWMassFlow
  // (src/outletBC.cpp:573-729, 739-892, 894-1027): characteristic update of the point's boundary state, then the
  // Lax-Friedrichs flux against the state it had BEFORE the update.
  void nr_flux(const OrcBc &b, NrState &st, int pt, const double *normal, const double *stateIn, const double *gradState,
               double *bdrFlux) const {
    const double gamma = phys.gamma, Rg = phys.R;
    const bool inlet = b.kind == 0;
    double unitNorm[3] = {0, 0, 0}, tangent2[3] = {0, 0, 0};
    const double *tangent1 = st.tangent1, *meanUp = st.meanUp;
    {
      double mod = 0.;
      for (int d = 0; d < dim; d++) mod += normal[d] * normal[d];
      for (int d = 0; d < dim; d++) unitNorm[d] = normal[d] * ((inlet ? -1. : 1.) / sqrt(mod));  // inlet: into the domain
    }
    double meanVel[3] = {0, 0, 0};
    for (int d = 0; d < dim; d++) {
      meanVel[0] += unitNorm[d] * meanUp[d + 1];
      meanVel[1] += tangent1[d] * meanUp[d + 1];
    }
    if (dim == 3) {
      tangent2[0] = unitNorm[1] * tangent1[2] - unitNorm[2] * tangent1[1];
      tangent2[1] = unitNorm[2] * tangent1[0] - unitNorm[0] * tangent1[2];
      tangent2[2] = unitNorm[0] * tangent1[1] - unitNorm[1] * tangent1[0];
      for (int d = 0; d < dim; d++) meanVel[2] += tangent2[d] * meanUp[d + 1];
    }
    double normGrad[16];
    for (int eq = 0; eq < neq; eq++) {
      normGrad[eq] = 0.;
      for (int d = 0; d < dim; d++) normGrad[eq] += unitNorm[d] * gradState[eq + d * neq];
    }
    // DryAir::ComputePressureDerivative(normGrad, stateIn, false) (src/equation_of_state.cpp:350-359)
    const double T = ph->pressure(stateIn) / (Rg * stateIn[0]);
    const double dpdn = Rg * (T * normGrad[0] + stateIn[0] * normGrad[1 + nvel]);
    const double speedSound = sqrt(gamma * Rg * meanUp[1 + nvel]);
    double meanK = 0.;
    for (int d = 0; d < nvel; d++) meanK += meanUp[1 + d] * meanUp[1 + d];
    meanK *= 0.5;
    const double refLength = b.data[8];
    const double sigma = speedSound / refLength;
    double L1, L2, L3, L4 = 0., L5;
    if (inlet) {
      double meanDV[3];
      for (int d = 0; d < nvel; d++) meanDV[d] = meanUp[1 + d] - b.data[1 + d];
      L1 = 0.;
      for (int d = 0; d < dim; d++) L1 += unitNorm[d] * normGrad[1 + d];
      L1 = dpdn - meanUp[0] * speedSound * L1;
      L1 *= meanVel[0] - speedSound;
      L5 = 0.;
      for (int d = 0; d < dim; d++) L5 += meanDV[d] * unitNorm[d];
      L5 *= sigma * 2. * meanUp[0] * speedSound;
      L3 = 0.;
      for (int d = 0; d < dim; d++) L3 += meanDV[d] * tangent1[d];
      L3 *= sigma;
      if (dim == 3) {
        for (int d = 0; d < dim; d++) L4 += meanDV[d] * tangent2[d];
        L4 *= sigma;
      }
      L2 = sigma * speedSound * speedSound * (meanUp[0] - b.data[0]) - 0.5 * L5;
      if (b.type == 7) L2 = 0.;  // SUB_VEL_CONST_ENT
    } else {
      L2 = speedSound * speedSound * normGrad[0] - dpdn;
      L2 *= meanVel[0];
      L3 = 0.;
      for (int d = 0; d < dim; d++) L3 += tangent1[d] * normGrad[1 + d];
      L3 *= meanVel[0];
      if (dim == 3) {
        for (int d = 0; d < dim; d++) L4 += tangent2[d] * normGrad[1 + d];
        L4 *= meanVel[0];
      }
      L5 = 0.;
      for (int d = 0; d < dim; d++) L5 += unitNorm[d] * normGrad[1 + d];
      L5 = dpdn + meanUp[0] * speedSound * L5;
      L5 *= meanVel[0] + speedSound;
      if (b.type == 2) {  // SUB_P_NR
        const double meanP = Rg * meanUp[0] * meanUp[1 + nvel];
        L1 = sigma * (meanP - b.data[0]);
      } else {
        double vn = meanVel[0];  // SUB_MF_NR: the patch mean; SUB_MF_NR_PW: the point's own normal velocity
        if (b.type == 4) {
          vn = 0.;
          for (int d = 0; d < dim; d++) vn += stateIn[1 + d] * unitNorm[d];
          vn /= stateIn[0];
        }
        L1 = -sigma * (vn - b.data[0] / meanUp[0] / st.area);
        L1 *= meanUp[0] * speedSound;
      }
    }
    const double d1 = (L2 + 0.5 * (L5 + L1)) / speedSound / speedSound;
    const double d2 = 0.5 * (L5 - L1) / meanUp[0] / speedSound;
    const double d3 = L3, d4 = L4, d5 = 0.5 * (L5 + L1);
    double dF[16];
    for (int eq = 0; eq < neq; eq++) dF[eq] = 0.;
    dF[0] = d1;
    dF[1] = meanVel[0] * d1 + meanUp[0] * d2;
    dF[2] = meanVel[1] * d1 + meanUp[0] * d3;
    if (dim == 3) dF[3] = meanVel[2] * d1 + meanUp[0] * d4;
    dF[1 + dim] = meanUp[0] * meanVel[0] * d2;
    dF[1 + dim] += meanUp[0] * meanVel[1] * d3;
    if (dim == 3) dF[1 + dim] += meanUp[0] * meanVel[2] * d4;
    dF[1 + dim] += meanK * d1 + d5 / (gamma - 1.);
    double state2[16], stateN[16], newU[16];
    double *bU = &st.boundaryU[static_cast<size_t>(pt) * neq];
    for (int eq = 0; eq < neq; eq++) state2[eq] = bU[eq], stateN[eq] = bU[eq];
    for (int d = 0; d < dim; d++) stateN[1 + d] = 0.;
    for (int d = 0; d < dim; d++) {
      stateN[1] += state2[1 + d] * unitNorm[d];
      stateN[2] += state2[1 + d] * tangent1[d];
      if (dim == 3) stateN[3] += state2[1 + d] * tangent2[d];
    }
    for (int i = 0; i < neq; i++) newU[i] = stateN[i] - bc_dt * dF[i];
    {  // back to Cartesian momentum: inverse of the matrix with rows unitNorm, tangent1, tangent2
      double M[9], invM[9], momX[3] = {0, 0, 0};
      for (int d = 0; d < dim; d++) {
        M[0 + d * dim] = unitNorm[d];
        M[1 + d * dim] = tangent1[d];
        if (dim == 3) M[2 + d * dim] = tangent2[d];
      }
      small_inverse(dim, M, invM);
      for (int i = 0; i < dim; i++)
        for (int j = 0; j < dim; j++) momX[i] += invM[i + j * dim] * newU[1 + j];
      for (int d = 0; d < dim; d++) newU[1 + d] = momX[d];
    }
    for (int eq = 0; eq < neq; eq++) bU[eq] = newU[eq];
    ph->riemann(stateIn, state2, normal, bdrFlux, true);
  }
  // column-major inverse of a 2 x 2 / 3 x 3 matrix (mfem::CalcInverse: adjugate / determinant)
  static void small_inverse(int n, const double *a, double *inv) {
    if (n == 2) {
      const double t = 1.0 / (a[0] * a[3] - a[1] * a[2]);
      inv[0] = a[3] * t, inv[1] = -a[1] * t, inv[2] = -a[2] * t, inv[3] = a[0] * t;
      return;
    }
    const double t = 1.0 / (a[0] * (a[4] * a[8] - a[5] * a[7]) - a[3] * (a[1] * a[8] - a[2] * a[7]) + a[6] * (a[1] * a[5] - a[2] * a[4]));
    inv[0] = (a[4] * a[8] - a[5] * a[7]) * t, inv[3] = (a[5] * a[6] - a[3] * a[8]) * t, inv[6] = (a[3] * a[7] - a[4] * a[6]) * t;
    inv[1] = (a[2] * a[7] - a[1] * a[8]) * t, inv[4] = (a[0] * a[8] - a[2] * a[6]) * t, inv[7] = (a[1] * a[6] - a[0] * a[7]) * t;
    inv[2] = (a[1] * a[5] - a[2] * a[4]) * t, inv[5] = (a[2] * a[3] - a[0] * a[5]) * t, inv[8] = (a[0] * a[4] - a[1] * a[3]) * t;
  }
  // BoundaryCondition::computeBdrPrimitiveStateForGradient (src/BoundaryCondition.cpp:55: copy) and the
  // WallBC override (src/wallBC.cpp:241-266: only VISC_ISOTH changes anything)
  void bc_prim_for_gradient(const OrcBc &b, const double *primIn, double *primBC) const {
    for (int eq = 0; eq < neq; eq++) primBC[eq] = primIn[eq];
    if (b.kind == 2 && b.type == 3) {  // WallType VISC_ISOTH
      for (int i = 0; i < nvel; i++) primBC[1 + i] = 0.0;
      primBC[nvel + 1] = b.data[0];
    }
  }
  // BCintegrator::computeBdrFlux (src/BCintegrator.cpp:228-242) dispatching to the per-type routines
  void bc_flux(const OrcBc &b, const double *normal, const double *stateIn, const double *gradState, double *xyz,
               double delta, double *bdrFlux, double dist = 0.0) const {
    double state2[16], wallState[16], viscF[48], wallViscF[16], unitN[3], primFlux[16];
    bool idx[16];
    for (int i = 0; i < 16; i++) {
      primFlux[i] = 0.0;
      idx[i] = false;
    }
    const int nsp = ph->num_species();
    double normN = 0.;
    for (int d = 0; d < dim; d++) normN += normal[d] * normal[d];
    for (int d = 0; d < dim; d++) unitN[d] = normal[d] * (1. / sqrt(normN));  // "unitNorm *= 1. / sqrt(normN)"
    if (b.kind == 0) {
      if (b.type != 2) return;  // InletType SUB_DENS_VEL only
      // InletBC::subsonicReflectingDensityVelocity (src/inletBC.cpp:729-756)
      const double pr = ph->pressure(stateIn);
      for (int eq = 0; eq < neq; eq++) state2[eq] = stateIn[eq];
      state2[0] = b.data[0];
      state2[1] = b.data[0] * b.data[1];
      state2[2] = b.data[0] * b.data[2];
      if (nvel == 3) state2[3] = b.data[0] * b.data[3];
      // inputState[4 + sp] = rho Y_sp of the active species (src/inletBC.cpp:742-750)
      for (int sp = 0; sp < ph->num_active_species(); sp++) state2[nvel + 2 + sp] = b.data[4 + sp];
      ph->modify_energy_for_pressure(state2, state2, pr, true);
      ph->riemann(stateIn, state2, normal, bdrFlux, true);  // rsolver->Eval(..., true): forced Lax-Friedrichs
    } else if (b.kind == 1) {
      if (b.type != 0) return;  // OutletType SUB_P only
      // OutletBC::subsonicReflectingPressure (src/outletBC.cpp:731-737)
      ph->modify_energy_for_pressure(stateIn, state2, b.data[0], false);
      ph->riemann(stateIn, state2, normal, bdrFlux, true);  // rsolver->Eval(..., true): forced Lax-Friedrichs
    } else if (ph->wall_bc_flux(b.type, b.data, use_bc_in_grad, normal, stateIn, gradState, xyz, delta, dist, bdrFlux)) {
      // reference back end: WallBC::computeBdrFlux itself (wallBC.cpp compiled into oracle/_ref); the branches below are
      // the restatement the "port" back end runs (and are checked against the object code by tests/test_cpu_bc_oracle.py)
    } else if (b.type == 0) {
      // WallBC::computeINVwallFlux (src/wallBC.cpp:277-320)
      double vel[3], un[3];
      for (int d = 0; d < nvel; d++) vel[d] = stateIn[1 + d] / stateIn[0];
      double norm = sqrt(normN);
      for (int d = 0; d < dim; d++) un[d] = normal[d] / norm;
      double vn = 0;
      for (int d = 0; d < dim; d++) vn += vel[d] * un[d];
      for (int eq = 0; eq < neq; eq++) state2[eq] = stateIn[eq];
      state2[1] = stateIn[0] * (vel[0] - 2. * vn * un[0]);
      state2[2] = stateIn[0] * (vel[1] - 2. * vn * un[1]);
      if (dim == 3) state2[3] = stateIn[0] * (vel[2] - 2. * vn * un[2]);
      if ((nvel == 3) && (dim == 2)) state2[3] = stateIn[0] * vel[2];
      ph->riemann(stateIn, state2, normal, bdrFlux);
      double viscFw[48];
      ph->visc_flux(state2, gradState, xyz, delta, dist, viscFw);
      for (int eq = 0; eq < neq; eq++) {
        wallViscF[eq] = 0.;
        for (int d = 0; d < dim; d++) wallViscF[eq] += viscFw[eq + d * neq] * normal[d];
      }
      ph->visc_flux(stateIn, gradState, xyz, delta, dist, viscF);
      for (int eq = 1; eq < neq; eq++) {
        bdrFlux[eq] -= 0.5 * wallViscF[eq];
        for (int d = 0; d < dim; d++) bdrFlux[eq] -= 0.5 * viscF[eq + d * neq] * normal[d];
      }
    } else if (b.type == 1) {
      // WallBC::computeSlipWallFlux (src/wallBC.cpp:326-428): mirror state, Riemann flux only
      double vel[3] = {0, 0, 0}, nVel[3];
      for (int d = 0; d < nvel; d++) vel[d] = stateIn[1 + d] / stateIn[0];
      slip_mirror_velocity(dim, normal, vel, nVel);
      for (int eq = 0; eq < neq; eq++) state2[eq] = stateIn[eq];
      state2[1] = stateIn[0] * nVel[0];
      state2[2] = stateIn[0] * nVel[1];
      if (dim == 3) state2[3] = stateIn[0] * nVel[2];
      ph->riemann(stateIn, state2, normal, bdrFlux);
    } else if (b.type == 2 || b.type == 3) {
      if (b.type == 2) {
        // WallBC::computeAdiabaticWallFlux (src/wallBC.cpp:430-469); bcFlux_: species + heat flux prescribed 0 (:88-96)
        ph->stagnation_state(stateIn, wallState);
        ph->riemann(stateIn, wallState, normal, bdrFlux, true);  // rsolver->Eval(..., true): forced Lax-Friedrichs
        ph->visc_flux(stateIn, gradState, xyz, delta, 0.0, viscF);
        for (int i = 0; i < nsp; i++) idx[i] = true;
        idx[nsp + nvel] = true;
        if (neq - (nvel + 2) - ph->num_active_species() == 1) idx[nsp + nvel + 1] = true;  // two-temperature (:93)
        ph->bdr_visc_flux(wallState, gradState, xyz, delta, 0.0, unitN, primFlux, idx, wallViscF);
      } else {
        // WallBC::computeIsothermalWallFlux (src/wallBC.cpp:471-510); bcFlux_: species flux prescribed 0 (:97-110)
        for (int eq = 0; eq < neq; eq++) wallState[eq] = stateIn[eq];
        if (use_bc_in_grad) {
          for (int i = 0; i < nvel; i++) wallState[i + 1] *= -1.0;
        } else {
          ph->stagnant_state_with_temp(stateIn, b.data[0], wallState);
        }
        ph->riemann(stateIn, wallState, normal, bdrFlux, true);  // rsolver->Eval(..., true): forced Lax-Friedrichs
        ph->stagnant_state_with_temp(stateIn, b.data[0], wallState);
        for (int i = 0; i < nsp; i++) idx[i] = true;
        ph->bdr_visc_flux(wallState, gradState, xyz, delta, 0.0, unitN, primFlux, idx, wallViscF);
        ph->visc_flux(stateIn, gradState, xyz, delta, 0.0, viscF);
      }
      for (int eq = 0; eq < neq; eq++) wallViscF[eq] *= sqrt(normN);
      for (int eq = 1; eq < neq; eq++) {
        bdrFlux[eq] -= 0.5 * wallViscF[eq];
        for (int d = 0; d < dim; d++) bdrFlux[eq] -= 0.5 * viscF[eq + d * neq] * normal[d];
      }
    } else if (b.type == 4) {
      // WallType VISC_GNRL: WallBC constructor (src/wallBC.cpp:112-147) + computeGeneralWallFlux (:512-543).
      // data = {hvyThermalCond, elecThermalCond, Th, Te}; ThermalCondition 0 ADIAB, 1 ISOTH, 2 SHTH
      const int hvy = static_cast<int>(b.data[0]), elec = static_cast<int>(b.data[1]);
      const bool twoT = neq - (nvel + 2) - ph->num_active_species() == 1;
      double bprim[16];
      bool pidx[16];
      for (int i = 0; i < 16; i++) {
        bprim[i] = 0.0;
        pidx[i] = false;
      }
      for (int d = 0; d < nvel; d++) pidx[d + 1] = true;
      for (int i = 0; i < nsp; i++) idx[i] = true;
      if (hvy == 1) {
        bprim[nvel + 1] = b.data[2];
        pidx[nvel + 1] = true;
      } else {
        idx[nsp + nvel] = true;
      }
      if (elec == 1) {  // index num_equation - 1 whether or not there is an electron-energy equation (:133-134)
        bprim[neq - 1] = b.data[3];
        pidx[neq - 1] = true;
      } else if (elec == 0) {
        idx[nsp + nvel + 1] = true;
      } else if (twoT) {
        idx[nsp + nvel + 1] = true;
      }
      ph->modify_state_from_primitive(stateIn, bprim, pidx, wallState);
      ph->riemann(stateIn, wallState, normal, bdrFlux, true);
      if (elec == 2) ph->sheath_bdr_flux(wallState, primFlux);
      ph->bdr_visc_flux(wallState, gradState, xyz, delta, 0.0, unitN, primFlux, idx, wallViscF);
      for (int eq = 0; eq < neq; eq++) wallViscF[eq] *= sqrt(normN);
      ph->visc_flux(stateIn, gradState, xyz, delta, 0.0, viscF);
      for (int eq = 1; eq < neq; eq++) {
        bdrFlux[eq] -= 0.5 * wallViscF[eq];
        for (int d = 0; d < dim; d++) bdrFlux[eq] -= 0.5 * viscF[eq + d * neq] * normal[d];
      }
    }
  }

  void setup() {
    np = p + 1;
    nv = 1 << dim;
    dof = 1;
    for (int d = 0; d < dim; d++) dof *= np;
    N = static_cast<long>(NE) * dof;
    std::vector<double> wtmp;
    // DG_FECollection(order, dim, basisType): GaussLegendre (open) or GaussLobatto (closed) nodes
    if (basis_type == 0)
      gauss_legendre01(np, nodes1d, wtmp);
    else
      gauss_lobatto01(np, nodes1d, wtmp);
    // volume rule: order 2p (src/rhs_operator.cpp:181, src/gradients.cpp:97, src/domain_integrator.cpp:69)
    nqv1 = int_rule == 0 ? gl_npts_for_order(2 * p) : gll_npts_for_order(2 * p);
    // face rule: min(OrderW1,OrderW2) + 2p, OrderW(multilinear map) = dim - 1 (src/face_integrator.cpp:233-243)
    nqf1 = int_rule == 0 ? gl_npts_for_order(dim - 1 + 2 * p) : gll_npts_for_order(dim - 1 + 2 * p);
    if (int_rule == 0) {
      gauss_legendre01(nqv1, qxv, qwv);
      gauss_legendre01(nqf1, qxf, qwf);
    } else {
      gauss_lobatto01(nqv1, qxv, qwv);
      gauss_lobatto01(nqf1, qxf, qwf);
    }
    nqv = 1;
    for (int d = 0; d < dim; d++) nqv *= nqv1;
    nqf = 1;
    for (int d = 0; d < dim - 1; d++) nqf *= nqf1;

    el_faces.assign(NE, {});
    for (int f = 0; f < NF; f++) {
      if (f_el2[f] < 0) continue;  // boundary faces are not in element_to_faces (src/M2ulPhyS.cpp:937-958)
      el_faces[f_el1[f]].push_back(f);
      if (f_el2[f] < NE) el_faces[f_el2[f]].push_back(f);
      shape_table(f_inf1[f]);
      shape_table(f_inf2[f]);
    }

    const size_t d2 = static_cast<size_t>(dof) * dof;
    axisym = (dim == 2 && nvel == 3);  // config.isAxisymmetric(): 2-D mesh (r, z) carrying three velocity components
    Me_inv.assign(NE * d2, 0.0);
    if (axisym) Me_inv_rad.assign(NE * d2, 0.0);
    Ke.assign(NE * d2 * dim, 0.0);
    Kfl.assign(NE * d2 * dim, 0.0);
    elSize.assign(NE, 0.0);
    nodeXYZ.assign(static_cast<size_t>(N) * dim, 0.0);
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int e = 0; e < NE; e++) {
      std::vector<double> shape(dof), dshape(dof * dim), Me(d2, 0.0), MeR(d2, 0.0), physd(dof * dim), dsdx(dof * dim);
      double *ke = &Ke[e * d2 * dim], *kf = &Kfl[e * d2 * dim];
      for (int q = 0; q < nqv; q++) {
        const int qi[3] = {q % nqv1, (q / nqv1) % nqv1, q / (nqv1 * nqv1)};
        double xi[3] = {0, 0, 0}, w = 1;
        for (int d = 0; d < dim; d++) {
          xi[d] = qxv[qi[d]];
          w *= qwv[qi[d]];
        }
        double J[9], A[9], xq[3] = {0, 0, 0};
        elem_map(e, xi, xq, J);
        const double dt = det(J);
        adj(J, A);
        calc_shape(xi, shape.data(), dshape.data());
        // MassIntegrator (src/rhs_operator.cpp:179-185)
        for (int i = 0; i < dof; i++)
          for (int j = 0; j < dof; j++) Me[i * dof + j] += w * dt * shape[i] * shape[j];
        // axisymmetric: M_ij = int r phi_i phi_j (MassIntegrator(radiusFcn), src/rhs_operator.cpp:191-205)
        if (axisym)
          for (int i = 0; i < dof; i++)
            for (int j = 0; j < dof; j++) MeR[i * dof + j] += w * dt * xq[0] * shape[i] * shape[j];
        // CalcPhysDShape = dshape * inv(J); dshapedx = dshape * adj(J)
        for (int k = 0; k < dof; k++)
          for (int d = 0; d < dim; d++) {
            double a = 0;
            for (int r = 0; r < dim; r++) a += dshape[k * dim + r] * A[r + dim * d];
            dsdx[k * dim + d] = a;
            physd[k * dim + d] = a / dt;
          }
        // Ke(j, k + d*dof) += shape(j) * dshape(k,d) * detJac  (src/gradients.cpp:112-120)
        const double detJac = dt * w;
        for (int d = 0; d < dim; d++)
          for (int k = 0; k < dof; k++)
            for (int j = 0; j < dof; j++) ke[j * (dim * dof) + k + d * dof] += shape[j] * physd[k * dim + d] * detJac;
        // elmat(j, k + d*dof) += (shape(k)*w [*radius]) * dshapedx(j,d)  (src/domain_integrator.cpp:71-97)
        const double wk = axisym ? w * xq[0] : w;
        for (int d = 0; d < dim; d++)
          for (int j = 0; j < dof; j++)
            for (int k = 0; k < dof; k++) kf[j * (dim * dof) + k + d * dof] += shape[k] * wk * dsdx[j * dim + d];
      }
      invert_dense(Me, dof);
      std::copy(Me.begin(), Me.end(), &Me_inv[e * d2]);
      if (axisym) {
        invert_dense(MeR, dof);
        std::copy(MeR.begin(), MeR.end(), &Me_inv_rad[e * d2]);
      }
      // elSize: GetElementSize(e,1)/order (src/rhs_operator.cpp:149-156)
      {
        const double c[3] = {0.5, 0.5, 0.5};
        double J[9];
        elem_map(e, c, nullptr, J);
        elSize[e] = min_singular(J, dim) / p;
      }
      // node coordinates (mesh->GetNodes into a byNODES L2 space, src/rhs_operator.cpp:139-142)
      for (int n = 0; n < dof; n++) {
        const double xi[3] = {nodes1d[n % np], nodes1d[(n / np) % np], dim == 3 ? nodes1d[n / (np * np)] : 0.0};
        elem_map(e, xi, &nodeXYZ[(static_cast<size_t>(e) * dof + n) * dim], nullptr);
      }
    }
    Up.assign(static_cast<size_t>(N) * neq, 0.0);
    gradUp.assign(static_cast<size_t>(N) * neq * dim, 0.0);
  }

  // src/rhs_operator.cpp:641-649
  void update_primitives(const double *x) {
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (long i = 0; i < N; i++) {
      double s[16], pr[16];
      for (int eq = 0; eq < neq; eq++) s[eq] = x[i + eq * N];
      ph->prim(s, pr);
      for (int eq = 0; eq < neq; eq++) Up[i + eq * N] = pr[eq];
    }
  }

  // src/gradients.cpp:144-232 with GradFaceIntegrator (src/faceGradientIntegration.cpp:40-140)
  void compute_gradients() {
    const int nd = neq * dim;
    // face contributions, one private buffer per face side, gathered in face order below
    std::vector<double> fc(static_cast<size_t>(NF) * 2 * dof * nd, 0.0);
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 16)
    for (int f = 0; f < NF; f++) {
      const bool bdr = f_el2[f] < 0;
      const OrcBc *bc = bdr && !bcs.empty() ? bc_of_face(f) : nullptr;
      if (bdr && (bcs.empty() || f_attr.empty())) continue;  // no boundary integrator registered
      const int e1 = f_el1[f], e2 = bdr ? e1 : f_el2[f];
      // Elem2No < 0: shape2 = shape1 (src/faceGradientIntegration.cpp:88-90)
      const std::vector<double> &sh1 = shapeTab.at(f_inf1[f]), &sh2 = shapeTab.at(bdr ? f_inf1[f] : f_inf2[f]);
      double *v1 = &fc[(static_cast<size_t>(f) * 2 + 0) * dof * nd];
      double *v2 = &fc[(static_cast<size_t>(f) * 2 + 1) * dof * nd];
      double iUp1[16], iUp2[16], mean[16], du1n[48], du2n[48], nor[3], xyz[3];
      for (int q = 0; q < nqf; q++) {
        const double *s1 = &sh1[static_cast<size_t>(q) * dof], *s2 = &sh2[static_cast<size_t>(q) * dof];
        for (int eq = 0; eq < neq; eq++) {
          double a = 0, b = 0;
          for (int k = 0; k < dof; k++) a += Up[static_cast<size_t>(e1) * dof + k + eq * N] * s1[k];
          if (!bdr)
            for (int k = 0; k < dof; k++) b += Up[static_cast<size_t>(e2) * dof + k + eq * N] * s2[k];
          iUp1[eq] = a;
          iUp2[eq] = b;
        }
        if (bdr) {  // src/faceGradientIntegration.cpp:96-115
          if (use_bc_in_grad && bc) {
            bc_prim_for_gradient(*bc, iUp1, iUp2);
          } else {
            for (int eq = 0; eq < neq; eq++) iUp2[eq] = iUp1[eq];
          }
        }
        for (int eq = 0; eq < neq; eq++) {
          mean[eq] = 0.5 * iUp1[eq];
          mean[eq] += 0.5 * iUp2[eq];
        }
        face_geom(f, q, nor, xyz);
        const double w = face_weight(q);
        for (int d = 0; d < dim; d++) nor[d] *= w;
        for (int d = 0; d < dim; d++)
          for (int eq = 0; eq < neq; eq++) {
            du1n[eq + d * neq] = (mean[eq] - iUp1[eq]) * nor[d];
            du2n[eq + d * neq] = (iUp2[eq] - mean[eq]) * nor[d];
          }
        for (int k = 0; k < dof; k++)
          for (int c = 0; c < nd; c++) {
            v1[k * nd + c] += s1[k] * du1n[c];
            v2[k * nd + c] += s2[k] * du2n[c];
          }
      }
    }
    const size_t d2 = static_cast<size_t>(dof) * dof;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int e = 0; e < NE; e++) {
      std::vector<double> rhs(static_cast<size_t>(dof) * nd, 0.0), fsum(static_cast<size_t>(dof) * nd, 0.0);
      const double *ke = &Ke[e * d2 * dim];
      // volume: elGradUp(j, eq + d*neq) = sum_k Ke(j, k + d*dof) * elUp(k, eq)  (src/gradients.cpp:174-182)
      for (int eq = 0; eq < neq; eq++)
        for (int d = 0; d < dim; d++)
          for (int j = 0; j < dof; j++) {
            double a = 0;
            for (int k = 0; k < dof; k++) a += ke[j * (dim * dof) + k + d * dof] * Up[static_cast<size_t>(e) * dof + k + eq * N];
            rhs[j * nd + eq + d * neq] = a;
          }
      for (int f : el_faces[e]) {
        const int side = (f_el1[f] == e) ? 0 : 1;
        const double *v = &fc[(static_cast<size_t>(f) * 2 + side) * dof * nd];
        for (size_t i = 0; i < fsum.size(); i++) fsum[i] += v[i];
      }
      if (!el_bfaces.empty())
        for (int f : el_bfaces[e]) {  // boundary faces: only Elem1's block is added (MFEM NonlinearForm::Mult)
          const double *v = &fc[(static_cast<size_t>(f) * 2 + 0) * dof * nd];
          for (size_t i = 0; i < fsum.size(); i++) fsum[i] += v[i];
        }
      for (size_t i = 0; i < rhs.size(); i++) rhs[i] += fsum[i];
      // Me_inv (src/gradients.cpp:209-227)
      const double *mi = &Me_inv[e * d2];
      for (int d = 0; d < dim; d++)
        for (int eq = 0; eq < neq; eq++)
          for (int j = 0; j < dof; j++) {
            double a = 0;
            for (int k = 0; k < dof; k++) a += mi[j * dof + k] * rhs[k * nd + eq + d * neq];
            gradUp[static_cast<size_t>(e) * dof + j + eq * N + static_cast<size_t>(d) * neq * N] = a;
          }
    }
  }

  // RHSoperator::Mult, CPU branch (src/rhs_operator.cpp:343-464)
  void mult(const double *x, double *y) {
    max_char_speed = 0.;
    update_primitives(x);
    compute_gradients();
    update_bc_mean();  // bcIntegrator->updateBCMean(Up) (src/rhs_operator.cpp:364)
    const int nact = ph->num_active_species();
    // ---- A->Mult: FaceIntegrator::NonLinearFaceIntegration (src/face_integrator.cpp:194-352)
    std::vector<double> fz(static_cast<size_t>(NF) * 2 * dof * neq, 0.0);
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 16)
    for (int f = 0; f < NF; f++) {
      if (f_el2[f] < 0) continue;
      const int e1 = f_el1[f], e2 = f_el2[f];
      const std::vector<double> &sh1 = shapeTab.at(f_inf1[f]), &sh2 = shapeTab.at(f_inf2[f]);
      double *v1 = &fz[(static_cast<size_t>(f) * 2 + 0) * dof * neq];
      double *v2 = &fz[(static_cast<size_t>(f) * 2 + 1) * dof * neq];
      const double delta1 = elSize[e1], delta2 = elSize[e2];
      double u1[16], u2[16], g1[48], g2[48], nor[3], xyz[3], fluxN[16], vF1[48], vF2[48];
      for (int q = 0; q < nqf; q++) {
        const double *s1 = &sh1[static_cast<size_t>(q) * dof], *s2 = &sh2[static_cast<size_t>(q) * dof];
        for (int eq = 0; eq < neq; eq++) {
          double a = 0, b = 0;
          for (int k = 0; k < dof; k++) a += x[static_cast<size_t>(e1) * dof + k + eq * N] * s1[k];
          for (int k = 0; k < dof; k++) b += x[static_cast<size_t>(e2) * dof + k + eq * N] * s2[k];
          u1[eq] = a;
          u2[eq] = b;
        }
        for (int sp = 0; sp < nact; sp++) {
          const int eq = nvel + 2 + sp;
          u1[eq] = std::max(u1[eq], 0.0);
          u2[eq] = std::max(u2[eq], 0.0);
        }
        for (int eq = 0; eq < neq; eq++)
          for (int d = 0; d < dim; d++) {
            double a = 0, b = 0;
            const double *ga = &gradUp[static_cast<size_t>(e1) * dof + eq * N + static_cast<size_t>(d) * neq * N];
            const double *gb = &gradUp[static_cast<size_t>(e2) * dof + eq * N + static_cast<size_t>(d) * neq * N];
            for (int k = 0; k < dof; k++) a += ga[k] * s1[k];
            for (int k = 0; k < dof; k++) b += gb[k] * s2[k];
            g1[eq + d * neq] = a;
            g2[eq + d * neq] = b;
          }
        face_geom(f, q, nor, xyz);
        ph->riemann(u1, u2, nor, fluxN);
        double d1 = 0.0, d2 = 0.0;  // wall distance, each side's own interpolation (src/face_integrator.cpp:304-309)
        if (!distance.empty()) {
          for (int k = 0; k < dof; k++) d1 += distance[static_cast<size_t>(e1) * dof + k] * s1[k];
          for (int k = 0; k < dof; k++) d2 += distance[static_cast<size_t>(e2) * dof + k] * s2[k];
        }
        ph->visc_flux(u1, g1, xyz, delta1, d1, vF1);
        ph->visc_flux(u2, g2, xyz, delta2, d2, vF2);
        // viscF1 += viscF2; viscF1 *= -0.5; viscF1.AddMult(nor, fluxN); fluxN *= ip.weight
        for (int i = 0; i < neq * dim; i++) {
          vF1[i] += vF2[i];
          vF1[i] *= -0.5;
        }
        for (int eq = 0; eq < neq; eq++) {
          double a = 0;
          for (int d = 0; d < dim; d++) a += vF1[eq + d * neq] * nor[d];
          fluxN[eq] += a;
        }
        const double w = face_weight(q);
        for (int eq = 0; eq < neq; eq++) fluxN[eq] *= w;
        if (axisym)  // src/face_integrator.cpp:344-346
          for (int eq = 0; eq < neq; eq++) fluxN[eq] *= xyz[0];
        for (int k = 0; k < dof; k++)
          for (int eq = 0; eq < neq; eq++) {
            v2[k * neq + eq] += s2[k] * fluxN[eq];
            v1[k * neq + eq] += -1.0 * s1[k] * fluxN[eq];
          }
      }
    }
    // ---- A->Mult boundary faces: BCintegrator::AssembleFaceVector (src/BCintegrator.cpp:295-441)
    if (!bcs.empty())
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 16)
      for (int f = 0; f < NF; f++) {
        if (f_el2[f] >= 0) continue;
        const OrcBc *bc = bc_of_face(f);
        if (!bc) continue;
        const int e1 = f_el1[f];
        const std::vector<double> &sh1 = shapeTab.at(f_inf1[f]);
        double *v1 = &fz[(static_cast<size_t>(f) * 2 + 0) * dof * neq];
        const double delta = elSize[e1];
        double u1[16], g1[48], nor[3], xyz[3], fluxN[16];
        for (int q = 0; q < nqf; q++) {
          const double *s1 = &sh1[static_cast<size_t>(q) * dof];
          for (int eq = 0; eq < neq; eq++) {
            double a = 0;
            for (int k = 0; k < dof; k++) a += x[static_cast<size_t>(e1) * dof + k + eq * N] * s1[k];
            const int sp = eq - nvel - 2;
            u1[eq] = (sp >= 0 && sp < nact) ? std::max(a, 0.0) : a;
            for (int d = 0; d < dim; d++) {
              double b = 0;
              const double *ga = &gradUp[static_cast<size_t>(e1) * dof + eq * N + static_cast<size_t>(d) * neq * N];
              for (int k = 0; k < dof; k++) b += ga[k] * s1[k];
              g1[eq + d * neq] = b;
            }
          }
          face_geom(f, q, nor, xyz);
          for (int eq = 0; eq < neq; eq++) fluxN[eq] = 0.;
          double d1 = 0.0;  // src/BCintegrator.cpp:416-424
          if (!distance.empty())
            for (int k = 0; k < dof; k++) d1 += distance[static_cast<size_t>(e1) * dof + k] * s1[k];
          if (is_nr(*bc)) {
            NrState &st = nr.at(static_cast<int>(bc - bcs.data()));
            nr_flux(*bc, st, st.face_off.at(f) + q, nor, u1, g1, fluxN);
          } else {
            bc_flux(*bc, nor, u1, g1, xyz, delta, fluxN, d1);
          }
          const double w = face_weight(q);
          for (int eq = 0; eq < neq; eq++) fluxN[eq] *= w;
          if (axisym)  // src/BCintegrator.cpp:427-429
            for (int eq = 0; eq < neq; eq++) fluxN[eq] *= xyz[0];
          for (int k = 0; k < dof; k++)
            for (int eq = 0; eq < neq; eq++) v1[k * neq + eq] -= fluxN[eq] * s1[k];
        }
      }
    // ---- GetFlux (src/rhs_operator.cpp:493-559), Aflux->AddMult (:379-391), Me_inv (:432-448)
    const size_t d2 = static_cast<size_t>(dof) * dof;
    std::vector<double> mcs_t(nthreads, 0.0);
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int e = 0; e < NE; e++) {
      int tid = 0;
#ifdef _OPENMP
      tid = omp_get_thread_num();
#endif
      std::vector<double> z(static_cast<size_t>(dof) * neq, 0.0), fl(static_cast<size_t>(dof) * dim * neq);
      for (int f : el_faces[e]) {
        const int side = (f_el1[f] == e) ? 0 : 1;
        const double *v = &fz[(static_cast<size_t>(f) * 2 + side) * dof * neq];
        for (size_t i = 0; i < z.size(); i++) z[i] += v[i];
      }
      if (!el_bfaces.empty())
        for (int f : el_bfaces[e]) {
          const double *v = &fz[(static_cast<size_t>(f) * 2 + 0) * dof * neq];
          for (size_t i = 0; i < z.size(); i++) z[i] += v[i];
        }
      for (int n = 0; n < dof; n++) {
        const size_t i = static_cast<size_t>(e) * dof + n;
        double st[16], g[48], fc[48], fv[48], xyz[3];
        for (int k = 0; k < neq; k++) st[k] = x[i + k * N];
        for (int sp = 0; sp < nact; sp++) st[nvel + 2 + sp] = std::max(st[nvel + 2 + sp], 0.0);
        for (int eq = 0; eq < neq; eq++)
          for (int d = 0; d < dim; d++) g[eq + d * neq] = gradUp[i + eq * N + static_cast<size_t>(d) * neq * N];
        for (int d = 0; d < dim; d++) xyz[d] = nodeXYZ[i * dim + d];
        ph->conv_flux(st, fc);
        if (phys.eq_system != 0) {
          ph->visc_flux(st, g, xyz, elSize[e], distance.empty() ? 0.0 : distance[i], fv);  // rhs_operator.cpp:519-533
          for (int c = 0; c < neq * dim; c++) fc[c] -= fv[c];
        }
        for (int d = 0; d < dim; d++)
          for (int k = 0; k < neq; k++) fl[(n * dim + d) * neq + k] = fc[k + d * neq];
        const double mcs = ph->max_char_speed(st);
        if (mcs > mcs_t[tid]) mcs_t[tid] = mcs;
      }
      const double *kf = &Kfl[e * d2 * dim];
      for (int eq = 0; eq < neq; eq++)
        for (int j = 0; j < dof; j++) {
          double a = 0;
          for (int d = 0; d < dim; d++)
            for (int k = 0; k < dof; k++) a += kf[j * (dim * dof) + k + d * dof] * fl[(k * dim + d) * neq + eq];
          z[j * neq + eq] += a;
        }
      const double *mi = axisym ? &Me_inv_rad[e * d2] : &Me_inv[e * d2];  // src/rhs_operator.cpp:441-445
      for (int eq = 0; eq < neq; eq++)
        for (int j = 0; j < dof; j++) {
          double a = 0;
          for (int k = 0; k < dof; k++) a += mi[j * dof + k] * z[k * neq + eq];
          y[static_cast<size_t>(e) * dof + j + eq * N] = a;
        }
    }
    for (double m : mcs_t) max_char_speed = std::max(max_char_speed, m);
    // ---- forcing terms, added after Me_inv (src/rhs_operator.cpp:451-461): SourceTerm::updateTerms
    if (ph->has_source() && !ph->source_update(sol_view ? sol_view : x, Up.data(), gradUp.data(), N, y)) {
      const double *Usol = sol_view ? sol_view : x;
#pragma omp parallel for num_threads(nthreads) schedule(static)
      for (long n = 0; n < N; n++) {
        double Un[16], upn[16], g[48], src[16];
        for (int eq = 0; eq < neq; eq++) {
          upn[eq] = Up[n + eq * N];
          Un[eq] = Usol[n + eq * N];
          for (int d = 0; d < dim; d++) g[eq + d * neq] = gradUp[n + eq * N + static_cast<size_t>(d) * neq * N];
        }
        ph->source_term(Un, upn, g, static_cast<int>(n), src);
        for (int eq = 0; eq < neq; eq++) y[n + eq * N] += src[eq];
      }
    }
    // AxisymmetricSource::updateTerms (src/forcing_terms.cpp:255-380), registered after SourceTerm
    // (src/rhs_operator.cpp:126-160); reads U_ (the solution grid function), Up and gradUp.
    if (axisym) {
      const double *Usol = sol_view ? sol_view : x;
#pragma omp parallel for num_threads(nthreads) schedule(static)
      for (long n = 0; n < N; n++) {
        double Un[16], upn[16], g[48];
        for (int eq = 0; eq < neq; eq++) {
          Un[eq] = Usol[n + eq * N];
          upn[eq] = Up[n + eq * N];
          for (int d = 0; d < dim; d++) g[eq + d * neq] = gradUp[n + eq * N + static_cast<size_t>(d) * neq * N];
        }
        for (int sp = 0; sp < nact; sp++) {
          const int eq = 3 + 2 + sp;
          Un[eq] = std::max(Un[eq], 0.0);
          upn[eq] = std::max(upn[eq], 0.0);
        }
        const double radius = nodeXYZ[n * dim + 0];
        const double rho = upn[0], ur = upn[1], ut = upn[3];
        const double pressure = ph->pressure_from_primitives(upn);
        const double rurut = rho * ur * ut, rutut = rho * ut * ut;
        double tau_tt, tau_tr;
        if (phys.eq_system == 0) {
          tau_tt = tau_tr = 0.0;
        } else {
          const double ur_r = g[1 + 0 * neq], uz_z = g[2 + 1 * neq], ut_r = g[3 + 0 * neq];
          double visc, bulkVisc, visc_vec[2];
          ph->get_viscosities(Un, upn, g, radius, 0.0, visc_vec);
          visc = visc_vec[0];
          bulkVisc = visc_vec[1];
          bulkVisc -= 2. / 3. * visc;
          double divV = ur_r + uz_z;
          if (radius > 0) divV += ur / radius;
          tau_tt = (radius > 0) ? 2.0 * ur / radius * visc : 0.0;
          tau_tt += bulkVisc * divV;
          tau_tr = ut_r;
          if (radius > 0) tau_tr -= ut / radius;
          tau_tr *= visc;
        }
        y[n + 1 * N] += (pressure + rutut - tau_tt) / radius;
        y[n + 3 * N] += (-rurut + tau_tr) / radius;
      }
    }
    for (const OrcForcing &f : forcings) apply_forcing(f, y);
  }

  // The remaining ForcingTerms subclasses (src/forcing_terms.cpp; the file needs MFEM's mesh / FE-space services in its
  // constructors, so the short updateTerms bodies are restated here), serial like the reference's CPU path.
  void apply_forcing(const OrcForcing &f, double *y) {
    if (f.kind == 0) {  // ConstantPressureGradient::updateTerms, CPU branch (:147-172)
      for (long n = 0; n < N; n++) {
        double primi[16];
        for (int eq = 0; eq < neq; eq++) primi[eq] = Up[n + eq * N];
        const double p = ph->pressure_from_primitives(primi);
        double grad_pV = 0.;
        for (int d = 0; d < dim; d++) {
          const double vel = Up[n + (d + 1) * N];
          y[n + (d + 1) * N] -= f.pressure_grad[d];
          grad_pV -= vel * f.pressure_grad[d];
          grad_pV -= p * gradUp[n + (d + 1) * N + static_cast<size_t>(d) * N * neq];
        }
        y[n + (1 + nvel) * N] += grad_pV;
      }
    } else if (f.kind == 1) {  // HeatSource::HeatSource node search (:941-970) + updateTerms (:972-985)
      double norm[3] = {0, 0, 0}, mod = 0.;
      for (int d = 0; d < dim; d++) norm[d] = f.hs_point2[d] - f.hs_point1[d];
      for (int d = 0; d < dim; d++) mod += norm[d] * norm[d];
      mod = sqrt(mod);
      for (int d = 0; d < dim; d++) norm[d] /= mod;
      for (long n = 0; n < N; n++) {
        double X[3] = {0, 0, 0}, proj = 0, normR = 0;
        for (int d = 0; d < dim; d++) X[d] = nodeXYZ[n * dim + d] - f.hs_point1[d];
        for (int d = 0; d < dim; d++) proj += X[d] * norm[d];
        for (int d = 0; d < dim; d++) {
          const double r = X[d] - proj * norm[d];
          normR += r * r;
        }
        normR = sqrt(normR);
        if (normR < f.hs_radius && proj > 0 && proj < mod) y[n + (dim + 1) * N] += f.hs_value;
      }
    } else if (f.kind == 2) {  // JouleHeating::updateTerms (:443-470)
      const bool twoT = neq - (nvel + 2) - ph->num_active_species() == 1;
      for (long n = 0; n < N; n++) {
        const double heating = f.joule_heating[n];
        if (heating > 0.) {
          y[n + (nvel + 1) * N] += heating;
          if (twoT) y[n + (neq - 1) * N] += heating;
        }
      }
    } else {  // SpongeZone (:472-760), dry air
      double nrm[3] = {0, 0, 0}, mod = 0.;
      for (int d = 0; d < dim; d++) mod += f.sz_normal[d] * f.sz_normal[d];
      mod = sqrt(mod);
      for (int d = 0; d < dim; d++) nrm[d] = f.sz_normal[d] / mod;
      double targetU[16], Upt[16];
      if (!f.sz_mixed_out) {  // USERDEF (:486-520): conserved from (rho, u, v, w, p) via modifyEnergyForPressure
        double conserved[16];
        conserved[0] = f.sz_target[0];
        for (int d = 0; d < nvel; d++) conserved[1 + d] = f.sz_target[0] * f.sz_target[1 + d];
        conserved[1 + nvel] = 0.0;
        ph->modify_energy_for_pressure(conserved, targetU, f.sz_target[4], false);
      } else {  // computeMixedOutValues (:706-742)
        double mean[17];
        for (int i = 0; i <= neq; i++) mean[i] = 0.;
        for (long n = 0; n < N; n++) {
          double distInit = 0.;
          for (int d = 0; d < dim; d++) distInit -= nrm[d] * (nodeXYZ[n * dim + d] - f.sz_point_init[d]);
          bool in;
          if (f.sz_type == 0) {
            in = fabs(distInit) < f.sz_tol;
          } else {
            double R = 0.;
            for (int d = 0; d < dim; d++) {
              const double t = nodeXYZ[n * dim + d] - f.sz_point_init[d] + distInit * nrm[d];
              R += t * t;
            }
            in = fabs(sqrt(R) - f.sz_r1) < f.sz_tol;
          }
          if (!in) continue;
          double up[16], Un[16], fl[48];
          for (int eq = 0; eq < neq; eq++) up[eq] = Up[n + eq * N];
          ph->cons(up, Un);
          ph->conv_flux(Un, fl);
          for (int eq = 0; eq < neq; eq++)
            for (int d = 0; d < dim; d++) mean[eq] += nrm[d] * fl[eq + d * neq];
          mean[neq] += 1.0;
        }
        for (int eq = 0; eq < neq; eq++) mean[eq] /= mean[neq];
        // DryAir::computeConservedStateFromConvectiveFlux (src/equation_of_state.cpp:414-442)
        const double gamma = phys.gamma;
        double temp = 0.;
        for (int d = 0; d < dim; d++) temp += mean[1 + d] * nrm[d];
        const double A = 1. - 2. * gamma / (gamma - 1.), B = 2 * temp / (gamma - 1.);
        double C = -2. * mean[0] * mean[1 + nvel];
        for (int d = 0; d < nvel; d++) C += mean[1 + d] * mean[1 + d];
        const double p = (-B - sqrt(B * B - 4. * A * C)) / (2. * A);
        double up[16];
        up[0] = mean[0] * mean[0] / (temp - p);
        up[1 + nvel] = p / (phys.R * up[0]);
        for (int d = 0; d < nvel; d++) up[1 + d] = d < dim ? (mean[1 + d] - p * nrm[d]) / mean[0] : mean[1 + d] / mean[0];
        ph->cons(up, targetU);
      }
      ph->prim(targetU, Upt);
      const double speedSound = sqrt(phys.gamma * phys.R * Upt[1 + nvel]);  // DryAir::ComputeSpeedOfSound(Up, true)
      for (long n = 0; n < N; n++) {  // sigma (:566-607) and addSpongeZoneForcing (:631-700)
        double distInit = 0., distF = 0., X[3] = {0, 0, 0};
        for (int d = 0; d < dim; d++) X[d] = nodeXYZ[n * dim + d];
        for (int d = 0; d < dim; d++) distInit -= nrm[d] * (X[d] - f.sz_point_init[d]);
        for (int d = 0; d < dim; d++) distF += nrm[d] * (X[d] - f.sz_point0[d]);
        double sgm = 0., ur[3] = {0, 0, 0};
        if (f.sz_type == 0) {
          if (distInit > 0. && distF > 0.) {
            const double planeDistance = distF + distInit;
            sgm = distInit / planeDistance / planeDistance;
          }
        } else {
          double R = 0., tmp[3] = {0, 0, 0};
          for (int d = 0; d < dim; d++) tmp[d] = X[d] - f.sz_point_init[d] + distInit * nrm[d];
          for (int d = 0; d < dim; d++) R += tmp[d] * tmp[d];
          R = sqrt(R);
          if (distInit > 0. && distF > 0. && R - f.sz_r1 > 0.) {
            const double planeDistance = f.sz_r2 - f.sz_r1;
            sgm = (R - f.sz_r1) / planeDistance / planeDistance;
            for (int d = 0; d < dim; d++) ur[d] = tmp[d] / R;
          }
        }
        if (!(sgm > 0.)) continue;
        sgm *= f.sz_mult;
        double up[16], Un[16], targetCyl[16];
        for (int eq = 0; eq < neq; eq++) up[eq] = Up[n + eq * N];
        ph->cons(up, Un);
        for (int eq = 0; eq < neq; eq++) targetCyl[eq] = targetU[eq];
        if (f.sz_type == 1) {  // MM rows (ur, uth, uz); targetCyl(1..3) = MM^-1 targetU(1..3)
          double uth[3], M[3][3], inv[3][3];
          uth[0] = nrm[1] * ur[2] - ur[1] * nrm[2];
          uth[1] = nrm[2] * ur[0] - nrm[0] * ur[2];
          uth[2] = nrm[0] * ur[1] - ur[0] * nrm[1];
          for (int d = 0; d < 3; d++) M[0][d] = ur[d], M[1][d] = uth[d], M[2][d] = nrm[d];
          const double det = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
                             M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
          inv[0][0] = (M[1][1] * M[2][2] - M[1][2] * M[2][1]) / det;
          inv[0][1] = (M[0][2] * M[2][1] - M[0][1] * M[2][2]) / det;
          inv[0][2] = (M[0][1] * M[1][2] - M[0][2] * M[1][1]) / det;
          inv[1][0] = (M[1][2] * M[2][0] - M[1][0] * M[2][2]) / det;
          inv[1][1] = (M[0][0] * M[2][2] - M[0][2] * M[2][0]) / det;
          inv[1][2] = (M[0][2] * M[1][0] - M[0][0] * M[1][2]) / det;
          inv[2][0] = (M[1][0] * M[2][1] - M[1][1] * M[2][0]) / det;
          inv[2][1] = (M[0][1] * M[2][0] - M[0][0] * M[2][1]) / det;
          inv[2][2] = (M[0][0] * M[1][1] - M[0][1] * M[1][0]) / det;
          for (int r = 0; r < 3; r++) targetCyl[1 + r] = inv[r][0] * targetU[1] + inv[r][1] * targetU[2] + inv[r][2] * targetU[3];
        }
        for (int eq = 0; eq < neq; eq++) y[n + eq * N] -= speedSound * sgm * (Un[eq] - targetCyl[eq]);
      }
    }
  }
};

}  // namespace orc

using orc::Oracle;

extern "C" {

const char *orc_physics_kind() {
  OrcPhysParams p;
  memset(&p, 0, sizeof(p));
  p.gamma = 1.4;
  p.R = 287.058;
  p.Pr = 0.71;
  p.S0 = 110.4;
  p.C1 = 1.458e-6;
  p.visc_mult = 1;
  orc::Physics *ph = orc::make_physics(p, 3, 3, 5);
  static char buf[32];
  snprintf(buf, sizeof(buf), "%s", ph ? ph->kind() : "none");
  delete ph;
  return buf;
}

// General constructor.  vx: [NE][2^dim][dim] element vertex coordinates (MFEM vertex order); faces in MFEM
// convention: el1/el2 = Elem1No/Elem2No (-1 boundary), inf = 64*local_face + orientation.
void *orc_create_ex(int dim, int order, int basis_type, int int_rule, int neq, int nvel, int NE, const double *vx,
                    int NF, const int *el1, const int *el2, const int *inf1, const int *inf2, const OrcPhysParams *phys,
                    int nthreads) {
  if (dim != 2 && dim != 3) return nullptr;
  Oracle *o = new Oracle;
  o->dim = dim;
  o->p = order;
  o->basis_type = basis_type;
  o->int_rule = int_rule;
  o->neq = neq;
  o->nvel = nvel;
  o->NE = NE;
  o->NF = NF;
  o->vx.assign(vx, vx + static_cast<size_t>(NE) * (1 << dim) * dim);
  o->f_el1.assign(el1, el1 + NF);
  o->f_el2.assign(el2, el2 + NF);
  o->f_inf1.assign(inf1, inf1 + NF);
  o->f_inf2.assign(inf2, inf2 + NF);
  o->phys = *phys;
  o->nthreads = nthreads > 0 ? nthreads : 1;
  o->ph = orc::make_physics(*phys, dim, nvel, neq);
  if (!o->ph) {
    delete o;
    return nullptr;
  }
  o->setup();
  return o;
}
// 3-D hexahedra, Gauss-Legendre nodes and rules, dry air (5 equations)
void *orc_create(int order, int NE, const double *vx, int NF, const int *el1, const int *el2, const int *inf1,
                 const int *inf2, const OrcPhysParams *phys, int nthreads) {
  return orc_create_ex(3, order, 0, 0, 5, 3, NE, vx, NF, el1, el2, inf1, inf2, phys, nthreads);
}
// boundary attribute per face (0 on interior faces) and BCintegrator's attribute maps
void orc_set_bcs(void *h, const int *face_attr, int nbc, const OrcBc *bcs, int use_bc_in_grad) {
  static_cast<Oracle *>(h)->set_bcs(face_attr, nbc, bcs, use_bc_in_grad);
}
// BoundaryCondition::dt (the time step the non-reflecting conditions advance their boundary state with)
void orc_set_bc_time_step(void *h, double dt) { static_cast<Oracle *>(h)->bc_dt = dt; }
// state of the non-reflecting condition on attribute attr: meanUp[neq], then boundaryU[points][neq]; returns the number of points
int orc_get_bc_state(void *h, int attr, double *mean_up, double *boundary_u, int capacity) {
  Oracle *o = static_cast<Oracle *>(h);
  for (auto &kv : o->nr)
    if (o->bcs[kv.first].attr == attr) {
      const int npts = static_cast<int>(kv.second.boundaryU.size()) / o->neq;
      for (int eq = 0; eq < o->neq; eq++) mean_up[eq] = kv.second.meanUp[eq];
      for (int i = 0; i < std::min(npts, capacity) * o->neq; i++) boundary_u[i] = kv.second.boundaryU[i];
      return npts;
    }
  return -1;
}
// one boundary flux evaluation (test probe): BCintegrator::computeBdrFlux
void orc_bc_flux(void *h, const OrcBc *bc, int use_bc_in_grad, const double *normal, const double *stateIn,
                 const double *gradState, double *flux) {
  Oracle *o = static_cast<Oracle *>(h);
  const int save = o->use_bc_in_grad;
  o->use_bc_in_grad = use_bc_in_grad;
  double xyz[3] = {0, 0, 0};
  for (int eq = 0; eq < o->neq; eq++) flux[eq] = 0.;
  o->bc_flux(*bc, normal, stateIn, gradState, xyz, 0.0, flux);
  o->use_bc_in_grad = save;
}
void orc_set_solution_view(void *h, const double *U) { static_cast<Oracle *>(h)->sol_view = U; }
void orc_add_forcing(void *h, const OrcForcing *f) { static_cast<Oracle *>(h)->forcings.push_back(*f); }
void orc_clear_forcings(void *h) { static_cast<Oracle *>(h)->forcings.clear(); }
void orc_set_rates(void *h, const double *data, int size) { static_cast<Oracle *>(h)->ph->set_rates(data, size); }
void orc_destroy(void *h) {
  Oracle *o = static_cast<Oracle *>(h);
  if (!o) return;
  delete o->ph;
  delete o;
}
long orc_ndofs(void *h) { return static_cast<Oracle *>(h)->N; }
int orc_num_equation(void *h) { return static_cast<Oracle *>(h)->neq; }
int orc_dim(void *h) { return static_cast<Oracle *>(h)->dim; }
void orc_update_primitives(void *h, const double *x, double *Up_out) {
  Oracle *o = static_cast<Oracle *>(h);
  o->update_primitives(x);
  if (Up_out) std::copy(o->Up.begin(), o->Up.end(), Up_out);
}
void orc_compute_gradients(void *h, const double *x, double *gradUp_out) {
  Oracle *o = static_cast<Oracle *>(h);
  o->update_primitives(x);
  o->compute_gradients();
  if (gradUp_out) std::copy(o->gradUp.begin(), o->gradUp.end(), gradUp_out);
}
void orc_rhs_mult(void *h, const double *x, double *y, double *gradUp_out, double *max_char_speed) {
  Oracle *o = static_cast<Oracle *>(h);
  o->mult(x, y);
  if (gradUp_out) std::copy(o->gradUp.begin(), o->gradUp.end(), gradUp_out);
  if (max_char_speed) *max_char_speed = o->max_char_speed;
}
// [MFEM RK4Solver::Step]: classic RK4, stages at t, t+dt/2, t+dt/2, t+dt (src/M2ulPhyS.cpp:721-739, :2005)
void orc_rk4_steps(void *h, double *U, double dt, int nsteps) {
  Oracle *o = static_cast<Oracle *>(h);
  const size_t n = static_cast<size_t>(o->N) * o->neq;
  std::vector<double> k(n), yv(n), z(n);
  if (!o->nr.empty()) o->bc_dt = dt;  // M2ulPhyS::dt, which the boundary conditions hold by reference
  for (int s = 0; s < nsteps; s++) {
    o->mult(U, k.data());
    for (size_t i = 0; i < n; i++) {
      yv[i] = U[i] + (dt / 2) * k[i];
      z[i] = U[i] + (dt / 6) * k[i];
    }
    o->mult(yv.data(), k.data());
    for (size_t i = 0; i < n; i++) {
      yv[i] = U[i] + (dt / 2) * k[i];
      z[i] += (dt / 3) * k[i];
    }
    o->mult(yv.data(), k.data());
    for (size_t i = 0; i < n; i++) {
      yv[i] = U[i] + dt * k[i];
      z[i] += (dt / 3) * k[i];
    }
    o->mult(yv.data(), k.data());
    for (size_t i = 0; i < n; i++) U[i] = z[i] + (dt / 6) * k[i];
  }
}
// geometry / table probes for the unit tests
void orc_node_coords(void *h, double *xyz /*[N][dim]*/) {
  Oracle *o = static_cast<Oracle *>(h);
  std::copy(o->nodeXYZ.begin(), o->nodeXYZ.end(), xyz);
}
int orc_face_nq(void *h) { return static_cast<Oracle *>(h)->nqf; }
void orc_face_geometry(void *h, int f, double *nor /*[nqf][dim]*/, double *xyz /*[nqf][dim]*/, double *w /*[nqf]*/) {
  Oracle *o = static_cast<Oracle *>(h);
  for (int q = 0; q < o->nqf; q++) {
    o->face_geom(f, q, &nor[q * o->dim], &xyz[q * o->dim]);
    w[q] = o->face_weight(q);
  }
}
// wall-distance grid function the mixing-length model reads (nullptr: none)
void orc_set_distance(void *h, const double *dist /*[N]*/) {
  Oracle *o = static_cast<Oracle *>(h);
  if (dist) o->distance.assign(dist, dist + o->N);
  else o->distance.clear();
}
void orc_elem_size(void *h, double *delta /*[NE]*/) {
  Oracle *o = static_cast<Oracle *>(h);
  std::copy(o->elSize.begin(), o->elSize.end(), delta);
}
void orc_dense_ops(void *h, int e, double *Me_inv, double *Ke, double *Kfl) {
  Oracle *o = static_cast<Oracle *>(h);
  const size_t d2 = static_cast<size_t>(o->dof) * o->dof;
  if (Me_inv) std::copy(&o->Me_inv[e * d2], &o->Me_inv[(e + 1) * d2], Me_inv);
  if (Ke) std::copy(&o->Ke[e * d2 * o->dim], &o->Ke[(e + 1) * d2 * o->dim], Ke);
  if (Kfl) std::copy(&o->Kfl[e * d2 * o->dim], &o->Kfl[(e + 1) * d2 * o->dim], Kfl);
}
void orc_gl_rule(int n, double *x, double *w) {
  std::vector<double> xv, wv;
  orc::gauss_legendre01(n, xv, wv);
  std::copy(xv.begin(), xv.end(), x);
  std::copy(wv.begin(), wv.end(), w);
}
void orc_gll_rule(int n, double *x, double *w) {
  std::vector<double> xv, wv;
  orc::gauss_lobatto01(n, xv, wv);
  std::copy(xv.begin(), xv.end(), x);
  std::copy(wv.begin(), wv.end(), w);
}

// ---- point-wise probes of the operator's own physics object (any fluid / dim / neq) ----
void orc_pt_prim(void *h, int n, const double *U, double *Up) {
  Oracle *o = static_cast<Oracle *>(h);
  for (int i = 0; i < n; i++) o->ph->prim(U + o->neq * i, Up + o->neq * i);
}
void orc_pt_cons(void *h, int n, const double *Up, double *U) {
  Oracle *o = static_cast<Oracle *>(h);
  for (int i = 0; i < n; i++) o->ph->cons(Up + o->neq * i, U + o->neq * i);
}
void orc_pt_max_char_speed(void *h, int n, const double *U, double *out) {
  Oracle *o = static_cast<Oracle *>(h);
  for (int i = 0; i < n; i++) out[i] = o->ph->max_char_speed(U + o->neq * i);
}
void orc_pt_conv_flux(void *h, int n, const double *U, double *F) {
  Oracle *o = static_cast<Oracle *>(h);
  for (int i = 0; i < n; i++) o->ph->conv_flux(U + o->neq * i, F + o->neq * o->dim * i);
}
void orc_pt_visc_flux(void *h, int n, const double *U, const double *gradUp, double *F) {
  Oracle *o = static_cast<Oracle *>(h);
  double xyz[3] = {0, 0, 0};
  for (int i = 0; i < n; i++) o->ph->visc_flux(U + o->neq * i, gradUp + o->neq * o->dim * i, xyz, 0.0, 0.0, F + o->neq * o->dim * i);
}
int orc_pt_mix_diffusivity(void *h, const double *U, double *D /*[numSpecies]*/) {
  return static_cast<Oracle *>(h)->ph->mixture_average_diffusivity(U, D) ? 0 : 1;
}
void orc_pt_source(void *h, int n, const double *Un, const double *Up, const double *gradUp, double *S) {
  Oracle *o = static_cast<Oracle *>(h);
  for (int i = 0; i < n; i++) {
    double a[16], b[16];
    for (int eq = 0; eq < o->neq; eq++) a[eq] = Un[o->neq * i + eq], b[eq] = Up[o->neq * i + eq];
    o->ph->source_term(a, b, gradUp + o->neq * o->dim * i, i, S + o->neq * i);
  }
}

// ---- point-wise physics probes (tests compare product device physics and port vs reference) ----
static orc::Physics *g_pp = nullptr;
void orc_phys_init(const OrcPhysParams *p) {
  delete g_pp;
  g_pp = orc::make_physics(*p, 3, 3, 5);
}
void orc_phys_prim(int n, const double *U, double *Up) {
  for (int i = 0; i < n; i++) g_pp->prim(U + 5 * i, Up + 5 * i);
}
void orc_phys_max_char_speed(int n, const double *U, double *out) {
  for (int i = 0; i < n; i++) out[i] = g_pp->max_char_speed(U + 5 * i);
}
void orc_phys_conv_flux(int n, const double *U, double *F) {
  for (int i = 0; i < n; i++) g_pp->conv_flux(U + 5 * i, F + 15 * i);
}
void orc_phys_visc_flux(int n, const double *U, const double *gradUp, double *F) {
  double xyz[3] = {0, 0, 0};
  for (int i = 0; i < n; i++) g_pp->visc_flux(U + 5 * i, gradUp + 15 * i, xyz, 0.0, 0.0, F + 15 * i);
}
void orc_phys_riemann(int n, const double *U1, const double *U2, const double *nor, double *F) {
  for (int i = 0; i < n; i++) g_pp->riemann(U1 + 5 * i, U2 + 5 * i, nor + 3 * i, F + 5 * i);
}
}
