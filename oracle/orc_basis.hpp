// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Never linked into the product.
//
// 1-D quadrature and nodal bases the way the reference obtains them from MFEM
// (third-party, not in /root/reference):
//   * IntegrationRules(0, Quadrature1D::GaussLegendre): segment rule of "real order"
//     order|1 with n = real_order/2 + 1 Gauss-Legendre points on [0,1]
//     (call sites: src/M2ulPhyS.cpp:558-562, src/rhs_operator.cpp:181-182,
//      src/gradients.cpp:97-98, src/domain_integrator.cpp:69-70, src/face_integrator.cpp:233-243)
//   * DG_FECollection(order, dim, BasisType::GaussLegendre): Lagrange basis on the
//     p+1 Gauss-Legendre points of [0,1], tensor ordering x fastest
//     (src/M2ulPhyS.cpp:565-572).
#pragma once
#include <cmath>
#include <vector>

namespace orc {

// n-point Gauss-Legendre rule on [0,1] (Newton iteration on P_n, ascending points).
inline void gauss_legendre01(int n, std::vector<double> &x, std::vector<double> &w) {
  x.assign(n, 0.0);
  w.assign(n, 0.0);
  for (int i = 0; i < (n + 1) / 2; i++) {
    long double z = cosl(M_PIl * (i + 0.75L) / (n + 0.5L));
    long double pp = 0, p1 = 0;
    for (int it = 0; it < 100; it++) {
      p1 = 1.0L;
      long double p2 = 0.0L;
      for (int j = 1; j <= n; j++) {
        long double p3 = p2;
        p2 = p1;
        p1 = ((2.0L * j - 1.0L) * z * p2 - (j - 1.0L) * p3) / j;
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0L);
      long double dz = p1 / pp;
      z -= dz;
      if (fabsl(dz) < 1e-19L) break;
    }
    // recompute derivative at converged root
    {
      p1 = 1.0L;
      long double p2 = 0.0L;
      for (int j = 1; j <= n; j++) {
        long double p3 = p2;
        p2 = p1;
        p1 = ((2.0L * j - 1.0L) * z * p2 - (j - 1.0L) * p3) / j;
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0L);
    }
    long double wi = 2.0L / ((1.0L - z * z) * pp * pp);
    // map [-1,1] -> [0,1]; root z is the (i)-th from the right
    x[i] = static_cast<double>(0.5L * (1.0L - z));
    x[n - 1 - i] = static_cast<double>(0.5L * (1.0L + z));
    w[i] = w[n - 1 - i] = static_cast<double>(0.5L * wi);
  }
}

// Number of 1-D Gauss-Legendre points MFEM uses for a requested order.
inline int gl_npts_for_order(int order) { return (order | 1) / 2 + 1; }
// ... and of Gauss-Lobatto points (IntegrationRules(0, Quadrature1D::GaussLobatto): n = order/2 + 2,
// exact to degree 2n-3; used with flow/integrationRule = 1, src/M2ulPhyS.cpp:558-562)
inline int gll_npts_for_order(int order) { return order / 2 + 2; }

// n-point Gauss-Lobatto rule on [0,1]: end points plus the roots of P'_{n-1}, w_i = 2 / (n (n-1) P_{n-1}(x_i)^2)
// on [-1,1].  Also the nodes of BasisType::GaussLobatto (flow/basisType = 1).
inline void gauss_lobatto01(int n, std::vector<double> &x, std::vector<double> &w) {
  x.assign(n, 0.0);
  w.assign(n, 0.0);
  const int m = n - 1;  // polynomial degree
  auto legendre = [&](long double z, long double &pm, long double &dpm) {
    long double p0 = 1.0L, p1 = z;
    if (m == 0) {
      pm = 1.0L;
      dpm = 0.0L;
      return;
    }
    for (int j = 2; j <= m; j++) {
      const long double p2 = ((2.0L * j - 1.0L) * z * p1 - (j - 1.0L) * p0) / j;
      p0 = p1;
      p1 = p2;
    }
    pm = p1;
    dpm = m * (z * p1 - p0) / (z * z - 1.0L);
  };
  for (int i = 0; i < n; i++) {
    long double z;
    if (i == 0) {
      z = -1.0L;
    } else if (i == n - 1) {
      z = 1.0L;
    } else {
      z = -cosl(M_PIl * i / m);  // Chebyshev-Lobatto start
      for (int it = 0; it < 100; it++) {
        // Newton on q(z) = (1 - z^2) P'_m(z): q' = -m (m+1) P_m(z)
        long double pm, dpm;
        legendre(z, pm, dpm);
        const long double q = (1.0L - z * z) * dpm, dq = -static_cast<long double>(m) * (m + 1) * pm;
        const long double dz = q / dq;
        z -= dz;
        if (fabsl(dz) < 1e-19L) break;
      }
    }
    long double pm, dpm;
    if (i == 0 || i == n - 1) {
      pm = (i == 0 && (m % 2)) ? -1.0L : 1.0L;
    } else {
      legendre(z, pm, dpm);
    }
    x[i] = static_cast<double>(0.5L * (1.0L + z));
    w[i] = static_cast<double>(0.5L * 2.0L / (static_cast<long double>(m) * (m + 1) * pm * pm));
  }
}

// Lagrange basis on `nodes` evaluated at x: values and derivatives (plain product form).
inline void lagrange(const std::vector<double> &nodes, double x, double *val, double *der) {
  const int n = static_cast<int>(nodes.size());
  for (int i = 0; i < n; i++) {
    double v = 1.0, denom = 1.0;
    for (int j = 0; j < n; j++) {
      if (j == i) continue;
      v *= (x - nodes[j]);
      denom *= (nodes[i] - nodes[j]);
    }
    val[i] = v / denom;
    if (der) {
      double d = 0.0;
      for (int m = 0; m < n; m++) {
        if (m == i) continue;
        double t = 1.0;
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

}  // namespace orc
