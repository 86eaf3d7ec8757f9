// Reference-element tables shared by meshkit (host) and the CUDA kernels (uploaded to __constant__).
//
// For basisType = 0 / integrationRule = 0 the reference's dense operators (Me_inv, Ke, the Aflux
// blocks; src/rhs_operator.cpp:173-189, src/gradients.cpp:87-129, src/domain_integrator.cpp:44-99)
// are exactly tensor products of the 1-D objects below, because the order-2p volume rule has p+1
// Gauss-Legendre points = the basis nodes ("design licence").
// MFEM conventions restated here (third party, not in /root/reference): hex vertex order,
// Geometry::Constants<CUBE>::FaceVert, quad_t::Orient, GetLocalQuadToHexTransformation.
#pragma once
#include <cmath>
#include <cstring>

namespace tpsb {

static const int HEX_VERT[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0},
                                   {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
static const int HEX_FACE_VERT[6][4] = {{3, 2, 1, 0}, {0, 1, 5, 4}, {1, 2, 6, 5},
                                        {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
static const int QUAD_ORIENT[8][4] = {{0, 1, 2, 3}, {0, 3, 2, 1}, {1, 2, 3, 0}, {1, 0, 3, 2},
                                      {2, 3, 0, 1}, {2, 1, 0, 3}, {3, 0, 1, 2}, {3, 2, 1, 0}};

constexpr int MAX_NP = 4;  // p <= 3
constexpr int MAX_NQ = 5;  // face points per direction for p = 3 (order 2 + 2p -> 5 GL points)

// Everything a kernel needs about the reference element, for one (NP, NQ).
struct RefTables {
  int np, nq;
  double xn[MAX_NP];          // Gauss-Legendre nodes on [0,1] (basis nodes = volume quadrature points)
  double wn[MAX_NP];          // their weights
  double D[MAX_NP][MAX_NP];   // D[i][m] = l_m'(xn[i])
  double lb[2][MAX_NP];       // lb[s][c] = l_c(s), s = 0,1 : trace extrapolation to the two ends
  double xq[MAX_NQ], wq[MAX_NQ];  // face quadrature (1-D factor)
  double P[MAX_NQ][MAX_NP];   // P[alpha][a] = l_a(xq[alpha]) : nodes -> face quadrature points
  // Local face lf in its own (orientation-0) face coordinates (s,t): face node (a,b), a along s.
  // Element node of (a,b,c) is face_base[lf][a + np*b] + c*face_cstride[lf], c = index along the
  // face-normal axis in ELEMENT order; the face sits at end face_side[lf] of that axis.
  int face_base[6][MAX_NP * MAX_NP];
  int face_cstride[6];
  int face_side[6];
  int face_axis[6];
  int face_vert[6][4];         // Geometry::Constants<CUBE>::FaceVert
  // inverse: element node n -> face-node index (a + np*b) of local face lf, and its c
  int node_ab[6][MAX_NP * MAX_NP * MAX_NP];
  int node_c[6][MAX_NP * MAX_NP * MAX_NP];
  // perm[ori][a + np*b] = a' + np*b': face coords -> Elem2-local face coords for orientation ori
  // (FaceElementTransformations Loc2), iperm = inverse permutation.
  int perm[8][MAX_NP * MAX_NP];
  int iperm[8][MAX_NP * MAX_NP];
  // Arithmetic forms of the tables above (what the kernels actually use, no divergent table reads):
  // face_par[lf] = as | at<<2 | an<<4 | ss<<6 | st<<7 | side<<8, with as/at/an the element axes of the
  // face's s / t / normal directions and ss/st = 1 when s / t run in the +axis direction.
  int face_par[6];
  // perm_code[ori] / iperm_code[ori] = swap | flip_a<<1 | flip_b<<2 : (x,y) = swap ? (b,a) : (a,b),
  // a' = flip_a ? np-1-x : x, b' = flip_b ? np-1-y : y.
  int perm_code[8];
  int iperm_code[8];
};

inline void gauss_legendre01(int n, double *x, double *w) {
  for (int i = 0; i < (n + 1) / 2; i++) {
    long double z = cosl(M_PIl * (i + 0.75L) / (n + 0.5L)), pp = 1, p1 = 1;
    for (int it = 0; it < 100; it++) {
      p1 = 1.0L;
      long double p2 = 0.0L;
      for (int j = 1; j <= n; j++) {
        const long double p3 = p2;
        p2 = p1;
        p1 = ((2.0L * j - 1.0L) * z * p2 - (j - 1.0L) * p3) / j;
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0L);
      const long double dz = p1 / pp;
      z -= dz;
      if (fabsl(dz) < 1e-19L) break;
    }
    p1 = 1.0L;
    long double p2 = 0.0L;
    for (int j = 1; j <= n; j++) {
      const long double p3 = p2;
      p2 = p1;
      p1 = ((2.0L * j - 1.0L) * z * p2 - (j - 1.0L) * p3) / j;
    }
    pp = n * (z * p1 - p2) / (z * z - 1.0L);
    const long double wi = 2.0L / ((1.0L - z * z) * pp * pp);
    x[i] = static_cast<double>(0.5L * (1.0L - z));
    x[n - 1 - i] = static_cast<double>(0.5L * (1.0L + z));
    w[i] = w[n - 1 - i] = static_cast<double>(0.5L * wi);
  }
}

inline void lagrange_ld(const double *nodes, int n, long double x, long double *val, long double *der) {
  for (int i = 0; i < n; i++) {
    long double v = 1, denom = 1;
    for (int j = 0; j < n; j++) {
      if (j == i) continue;
      v *= (x - nodes[j]);
      denom *= (static_cast<long double>(nodes[i]) - nodes[j]);
    }
    val[i] = v / denom;
    if (der) {
      long double d = 0;
      for (int m = 0; m < n; m++) {
        if (m == i) continue;
        long double t = 1;
        for (int j = 0; j < n; j++) {
          if (j == i || j == m) continue;
          t *= (x - nodes[j]);
        }
        d += t;
      }
      der[i] = d / denom;
    }
  }
}

// n-point Gauss-Lobatto rule on [0,1] (also the BasisType::GaussLobatto nodes): end points + roots of P'_{n-1},
// found by Newton on the derivative recurrence; w_i = 1 / (n (n-1) P_{n-1}(z_i)^2) on [0,1].
inline void gauss_lobatto01(int n, double *x, double *w) {
  const int m = n - 1;
  auto eval = [&](long double z, long double &pm, long double &dpm, long double &ddpm) {
    // P_m, P_m', P_m'' by upward recurrence
    long double p0 = 1, p1 = z, d0 = 0, d1 = 1, s0 = 0, s1 = 0;
    if (m == 0) {
      pm = 1, dpm = 0, ddpm = 0;
      return;
    }
    for (int j = 2; j <= m; j++) {
      const long double p2 = ((2.0L * j - 1) * z * p1 - (j - 1.0L) * p0) / j;
      const long double d2 = d0 + (2.0L * j - 1) * p1;
      const long double s2 = s0 + (2.0L * j - 1) * d1;
      p0 = p1, p1 = p2, d0 = d1, d1 = d2, s0 = s1, s1 = s2;
    }
    pm = p1, dpm = d1, ddpm = s1;
  };
  for (int i = 0; i < n; i++) {
    long double z = (i == 0) ? -1.0L : (i == n - 1 ? 1.0L : -cosl(M_PIl * i / m));
    long double pm, dpm, ddpm;
    if (i > 0 && i < n - 1)
      for (int it = 0; it < 100; it++) {
        eval(z, pm, dpm, ddpm);
        const long double dz = dpm / ddpm;
        z -= dz;
        if (fabsl(dz) < 1e-19L) break;
      }
    eval(z, pm, dpm, ddpm);
    x[i] = static_cast<double>(0.5L * (1.0L + z));
    w[i] = static_cast<double>(1.0L / (static_cast<long double>(m) * (m + 1) * pm * pm));
  }
}

// order p, Gauss-Legendre basis and rules; trilinear hex mesh (OrderW = 2): face rule order 2 + 2p.
inline bool build_ref_tables(int p, RefTables &T) {
  memset(&T, 0, sizeof(T));
  const int np = p + 1;
  const int nq = ((2 + 2 * p) | 1) / 2 + 1;
  if (np > MAX_NP || nq > MAX_NQ || p < 1) return false;
  T.np = np;
  T.nq = nq;
  gauss_legendre01(np, T.xn, T.wn);
  gauss_legendre01(nq, T.xq, T.wq);
  long double v[MAX_NP], d[MAX_NP];
  for (int i = 0; i < np; i++) {
    lagrange_ld(T.xn, np, T.xn[i], v, d);
    for (int m = 0; m < np; m++) T.D[i][m] = static_cast<double>(d[m]);
  }
  for (int s = 0; s < 2; s++) {
    lagrange_ld(T.xn, np, static_cast<long double>(s), v, nullptr);
    for (int c = 0; c < np; c++) T.lb[s][c] = static_cast<double>(v[c]);
  }
  for (int a = 0; a < nq; a++) {
    lagrange_ld(T.xn, np, T.xq[a], v, nullptr);
    for (int m = 0; m < np; m++) T.P[a][m] = static_cast<double>(v[m]);
  }
  const int stride[3] = {1, np, np * np};
  for (int lf = 0; lf < 6; lf++) {
    const int *hv = HEX_FACE_VERT[lf];
    int o[3], es[3], et[3];
    for (int k = 0; k < 3; k++) {
      o[k] = HEX_VERT[hv[0]][k];
      es[k] = HEX_VERT[hv[1]][k] - o[k];
      et[k] = HEX_VERT[hv[3]][k] - o[k];
    }
    int as = -1, at = -1, an = -1;
    for (int k = 0; k < 3; k++) {
      if (es[k] != 0) as = k;
      if (et[k] != 0) at = k;
    }
    for (int k = 0; k < 3; k++)
      if (k != as && k != at) an = k;
    T.face_axis[lf] = an;
    for (int q = 0; q < 4; q++) T.face_vert[lf][q] = hv[q];
    T.face_side[lf] = o[an];
    T.face_par[lf] = as | (at << 2) | (an << 4) | ((es[as] > 0 ? 1 : 0) << 6) | ((et[at] > 0 ? 1 : 0) << 7) | (o[an] << 8);
    T.face_cstride[lf] = stride[an];
    for (int b = 0; b < np; b++)
      for (int a = 0; a < np; a++) {
        const int ia = es[as] > 0 ? a : np - 1 - a;
        const int ib = et[at] > 0 ? b : np - 1 - b;
        const int base = ia * stride[as] + ib * stride[at];
        T.face_base[lf][a + np * b] = base;
        for (int c = 0; c < np; c++) {
          T.node_ab[lf][base + c * stride[an]] = a + np * b;
          T.node_c[lf][base + c * stride[an]] = c;
        }
      }
  }
  // Loc2 for orientation ori maps face vertex j to Elem2's local face vertex qo[j]; in node-index
  // space the affine map of the unit square sends corner v_j to corner v_{qo[j]}.
  static const int QV[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};
  for (int ori = 0; ori < 8; ori++) {
    const int *qo = QUAD_ORIENT[ori];
    const int o0[2] = {QV[qo[0]][0], QV[qo[0]][1]};
    const int ds[2] = {QV[qo[1]][0] - o0[0], QV[qo[1]][1] - o0[1]};
    const int dt[2] = {QV[qo[3]][0] - o0[0], QV[qo[3]][1] - o0[1]};
    for (int b = 0; b < np; b++)
      for (int a = 0; a < np; a++) {
        // index-space image: coordinate value index i <-> (np-1-i) under reflection
        int img[2];
        for (int k = 0; k < 2; k++) {
          // s' = o0 + ds*s + dt*t with s,t in {index}; a reflected axis maps index i -> np-1-i
          int val;
          if (ds[k] != 0)
            val = ds[k] > 0 ? a : np - 1 - a;
          else
            val = dt[k] > 0 ? b : np - 1 - b;
          img[k] = val;
        }
        const int src = a + np * b, dst = img[0] + np * img[1];
        T.perm[ori][src] = dst;
        T.iperm[ori][dst] = src;
      }
  }
  // encode the orientation permutations arithmetically and verify the encoding against the tables
  for (int ori = 0; ori < 8; ori++)
    for (int inv = 0; inv < 2; inv++) {
      const int *tab = inv ? T.iperm[ori] : T.perm[ori];
      int found = -1;
      for (int code = 0; code < 8 && found < 0; code++) {
        bool ok = true;
        for (int b = 0; b < np && ok; b++)
          for (int a = 0; a < np && ok; a++) {
            const int x = (code & 1) ? b : a, y = (code & 1) ? a : b;
            const int a2 = (code & 2) ? np - 1 - x : x, b2 = (code & 4) ? np - 1 - y : y;
            ok = tab[a + np * b] == a2 + np * b2;
          }
        if (ok) found = code;
      }
      if (found < 0) return false;
      (inv ? T.iperm_code : T.perm_code)[ori] = found;
    }
  return true;
}

}  // namespace tpsb
