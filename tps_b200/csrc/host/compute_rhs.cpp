// Stand-alone RHS probe in the shape of the reference's utils/compute_rhs.cpp: build the operator,
// call rhsOperator->Mult(state, rhs) once (utils/compute_rhs.cpp:102), report norms per equation, then
// take a few RK4 steps through the ODESolver interface.  Mesh tables come from meshkit (MFEM's job in a
// TPS build); the state is the Taylor-Green field of SURVEY.md 8(d) evaluated at the element vertices'
// trilinear GL nodes.
//   usage: compute_rhs [n=8] [rk4_steps=2]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tpsb_host.hpp"

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 8;
  const int nsteps = argc > 2 ? atoi(argv[2]) : 2;
  const double PI = 3.14159265358979323846;
  const double lo[3] = {-PI, -PI, -PI}, hi[3] = {PI, PI, PI};
  const int per[3] = {1, 1, 1};
  const int NE = n * n * n;
  std::vector<int> ev(8 * (size_t)NE), f1(6 * (size_t)NE), f2(f1.size()), i1(f1.size()), i2(f1.size());
  std::vector<double> xyz(24 * (size_t)NE);
  if (tpsb_mk_cartesian_hex(n, n, n, lo, hi, per, 0, ev.data(), xyz.data()) != TPSB_OK) return 2;
  const int nf = tpsb_mk_build_faces(NE, ev.data(), f1.data(), f2.data(), i1.data(), i2.data());
  if (nf < 0) return 2;
  tpsb_mesh_maps maps = {3, NE, 0, xyz.data(), nf, f1.data(), f2.data(), i1.data(), i2.data(), nullptr};
  tpsb_space_desc space = {3, 0, 0, 5, 3};
  tpsb_physics phys = {TPSB_NS, TPSB_DRY_AIR, 1.4, 287.058, 1420.0, 0.0, 1.458e-6, 110.4, 0.71, nullptr};
  try {
    tpsb_host::RHSoperator rhsOperator(maps, space, phys);
    const int64_t N = tpsb_num_dofs(rhsOperator.context());
    // Taylor-Green state at the GL nodes
    double tab[512];
    tpsb_get_ref_tables(3, tab, 512);
    const double *xn = tab + 2;
    std::vector<double> U(5 * (size_t)N);
    const double rho0 = 1.2, p0 = 101300.0, g = 1.4, V0 = 0.1 * std::sqrt(g * p0 / rho0);
    const double h = 2 * PI / n;
    for (int e = 0; e < NE; e++) {
      const double *v0 = &xyz[24 * (size_t)e];
      for (int nd = 0; nd < 64; nd++) {
        const double x = v0[0] + h * xn[nd % 4], y = v0[1] + h * xn[(nd / 4) % 4], z = v0[2] + h * xn[nd / 16];
        const double u = V0 * sin(x) * cos(y) * cos(z), v = -V0 * cos(x) * sin(y) * cos(z);
        const double p = p0 + rho0 * V0 * V0 / 16.0 * (cos(2 * x) + cos(2 * y)) * (cos(2 * z) + 2.0);
        const size_t i = (size_t)e * 64 + nd;
        U[i] = rho0;
        U[i + N] = rho0 * u;
        U[i + 2 * N] = rho0 * v;
        U[i + 3 * N] = 0.0;
        U[i + 4 * N] = p / (g - 1) + 0.5 * rho0 * (u * u + v * v);
      }
    }
    tpsb_host::Vector src_state, rhs(5 * N);
    src_state.SetFromHost(U);
    rhsOperator.Mult(src_state, rhs);  // utils/compute_rhs.cpp:102
    const std::vector<double> r = rhs.HostCopy();
    printf("numElems %d  dofs %lld\n", NE, (long long)N);
    for (int eq = 0; eq < 5; eq++) {
      double s = 0;
      for (int64_t i = 0; i < N; i++) s += r[i + eq * N] * r[i + eq * N];
      printf("rhs_l2[%d] %.15e\n", eq, std::sqrt(s / N));
    }
    printf("max_char_speed %.15e\n", rhsOperator.getMaxCharSpeed());
    tpsb_host::RK4Solver timeIntegrator;
    timeIntegrator.Init(rhsOperator);
    double t = 0, dt = 1e-5;
    for (int s = 0; s < nsteps; s++) timeIntegrator.Step(src_state, t, dt);
    const std::vector<double> u1 = src_state.HostCopy();
    double s = 0;
    for (size_t i = 0; i < u1.size(); i++) s += u1[i];
    printf("time %.6e  sum(U) %.15e\n", t, s);
  } catch (const std::exception &ex) {
    fprintf(stderr, "compute_rhs: %s\n", ex.what());
    return 1;
  }
  return 0;
}
