// Per-point dry-air physics as plain device functions over a POD parameter block -- the
// B200-side replacement of the reference's virtual GasMixture / TransportProperties / Fluxes /
// RiemannSolverTPS objects (no device vtables, no placement-new kernels; cf. src/gpu_constructor.cpp).
// Each function follows the arithmetic of the reference routine it replaces so that results agree
// to round-off; citations are to /root/reference/src.
#pragma once
#include <cuda_runtime.h>

namespace tpsb {

constexpr int NEQ = 5;   // 3-D dry air: [rho, rho u, rho v, rho w, rho E]
constexpr int DIM = 3;

struct PhysParams {
  int eq_system;  // 0 EULER, 1 NS
  double gamma, R, gm1;
  double visc_mult, bulk_visc_mult, C1, S0, Pr;
  double cp_div_pr;  // gamma R / (Pr (gamma-1))   transport_properties.cpp:220
};

// DryAir::ComputePressure   equation_of_state.hpp:610-617
__device__ __forceinline__ double dry_pressure(const PhysParams &p, const double *s) {
  double den_vel2 = s[1] * s[1] + s[2] * s[2] + s[3] * s[3];
  den_vel2 /= s[0];
  return p.gm1 * (s[4] - 0.5 * den_vel2);
}

// DryAir::GetPrimitivesFromConservatives   equation_of_state.cpp:321-335 (+ ComputeTemperature hpp:621-627)
__device__ __forceinline__ void dry_prim(const PhysParams &p, const double *s, double *up) {
  double den_vel2 = s[1] * s[1] + s[2] * s[2] + s[3] * s[3];
  den_vel2 /= s[0];
  const double T = p.gm1 / p.R * (s[4] - 0.5 * den_vel2) / s[0];
  up[0] = s[0];
  up[1] = s[1] / s[0];
  up[2] = s[2] / s[0];
  up[3] = s[3] / s[0];
  up[4] = T;
}

// DryAir::ComputeMaxCharSpeed   equation_of_state.cpp:279-294
__device__ __forceinline__ double dry_max_char_speed(const PhysParams &p, const double *s) {
  const double den = s[0];
  double den_vel2 = s[1] * s[1] + s[2] * s[2] + s[3] * s[3];
  den_vel2 /= den;
  const double pres = p.gm1 * (s[4] - 0.5 * den_vel2);
  const double sound = sqrt(p.gamma * pres / den);
  const double vel = sqrt(den_vel2 / den);
  return vel + sound;
}

// Fluxes::ComputeConvectiveFluxes   fluxes.cpp:135-170 ; f[eq + d*NEQ]
__device__ __forceinline__ void dry_conv_flux(const PhysParams &p, const double *s, double *f) {
  const double pres = dry_pressure(p, s);
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    f[0 + d * NEQ] = s[d + 1];
#pragma unroll
    for (int i = 0; i < DIM; i++) f[1 + i + d * NEQ] = s[i + 1] * s[d + 1] / s[0];
    f[1 + d + d * NEQ] += pres;
  }
  const double H = (s[4] + pres) / s[0];
#pragma unroll
  for (int d = 0; d < DIM; d++) f[4 + d * NEQ] = s[d + 1] * H;
}

// RiemannSolverTPS::ComputeFluxDotN   riemann_solver.cpp:53-64
__device__ __forceinline__ void dry_flux_dot_n(const PhysParams &p, const double *s, const double *nor, double *fn) {
  double f[NEQ * DIM];
  dry_conv_flux(p, s, f);
#pragma unroll
  for (int eq = 0; eq < NEQ; eq++) {
    double a = 0;
#pragma unroll
    for (int d = 0; d < DIM; d++) a += f[eq + d * NEQ] * nor[d];
    fn[eq] = a;
  }
}

// RiemannSolverTPS::Eval_LF   riemann_solver.cpp:89-114
__device__ __forceinline__ void dry_riemann_lf(const PhysParams &p, const double *s1, const double *s2,
                                               const double *nor, double *flux) {
  const double maxE = fmax(dry_max_char_speed(p, s1), dry_max_char_speed(p, s2));
  double f1[NEQ], f2[NEQ];
  dry_flux_dot_n(p, s1, nor, f1);
  dry_flux_dot_n(p, s2, nor, f2);
  const double normag = sqrt(nor[0] * nor[0] + nor[1] * nor[1] + nor[2] * nor[2]);
#pragma unroll
  for (int i = 0; i < NEQ; i++) flux[i] = 0.5 * (f1[i] + f2[i]) - 0.5 * maxE * (s2[i] - s1[i]) * normag;
}

// Fluxes::ComputeViscousFluxes (fluxes.cpp:178-335) with DryAirTransport::ComputeFluxMolecularTransport
// (transport_properties.cpp:223-234): non-axisymmetric, no SGS, no viscous sponge, single temperature.
// g[eq + d*NEQ] = d(Up_eq)/dx_d ; f[eq + d*NEQ].
__device__ __forceinline__ void dry_visc_flux(const PhysParams &p, const double *s, const double *g, double *f) {
  const double pr = dry_pressure(p, s);
  const double temp = pr / p.R / s[0];
  // pow(temp, 1.5) of the reference evaluated as temp*sqrt(temp) (agrees to an ulp)
  const double visc = (p.C1 * p.visc_mult * (temp * sqrt(temp)) / (temp + p.S0));
  double bulk = p.bulk_visc_mult * visc;
  const double k = p.cp_div_pr * visc;
  bulk -= 2. / 3. * visc;
  double stress[DIM * DIM];
  double divV = 0.;
#pragma unroll
  for (int i = 0; i < DIM; i++) {
#pragma unroll
    for (int j = 0; j < DIM; j++) stress[i + j * DIM] = g[(1 + j) + i * NEQ] + g[(1 + i) + j * NEQ];
    divV += g[(1 + i) + i * NEQ];
  }
#pragma unroll
  for (int i = 0; i < DIM * DIM; i++) stress[i] *= visc;
#pragma unroll
  for (int i = 0; i < DIM; i++) stress[i + i * DIM] += bulk * divV;
  double vel[DIM];
#pragma unroll
  for (int d = 0; d < DIM; d++) vel[d] = s[1 + d] / s[0];
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    f[0 + d * NEQ] = 0.;
#pragma unroll
    for (int i = 0; i < DIM; i++) f[(1 + i) + d * NEQ] = stress[i + d * DIM];
    double vtmp = 0.0;
#pragma unroll
    for (int j = 0; j < DIM; j++) vtmp += stress[d + j * DIM] * vel[j];
    f[4 + d * NEQ] = vtmp + k * g[4 + d * NEQ];
  }
}

// ---- fused / reciprocal forms used by the hot kernels -------------------------------------------
// Same formulas as above with 1/rho formed once and the normal contraction done analytically; they
// differ from the reference's operation order by a few ulp (well inside the 1e-10 parity bound) and
// cut the FP64 divide/sqrt count per quadrature point from ~40 to 8.
struct DryPoint {
  double rinv, vel[DIM], p;
};
__device__ __forceinline__ DryPoint dry_point(const PhysParams &p, const double *s) {
  DryPoint q;
  q.rinv = 1.0 / s[0];
  q.vel[0] = s[1] * q.rinv;
  q.vel[1] = s[2] * q.rinv;
  q.vel[2] = s[3] * q.rinv;
  q.p = p.gm1 * (s[4] - 0.5 * (s[1] * q.vel[0] + s[2] * q.vel[1] + s[3] * q.vel[2]));
  return q;
}
__device__ __forceinline__ double dry_char_speed_pt(const PhysParams &p, const DryPoint &q) {
  return sqrt(q.vel[0] * q.vel[0] + q.vel[1] * q.vel[1] + q.vel[2] * q.vel[2]) + sqrt(p.gamma * q.p * q.rinv);
}
// F_c(s).n
__device__ __forceinline__ void dry_conv_dot_n(const double *s, const DryPoint &q, const double *nor, double *fn) {
  const double vn = q.vel[0] * nor[0] + q.vel[1] * nor[1] + q.vel[2] * nor[2];
  fn[0] = s[1] * nor[0] + s[2] * nor[1] + s[3] * nor[2];
  fn[1] = s[1] * vn + q.p * nor[0];
  fn[2] = s[2] * vn + q.p * nor[1];
  fn[3] = s[3] * vn + q.p * nor[2];
  fn[4] = (s[4] + q.p) * vn;
}
// transport coefficients of DryAirTransport (Sutherland) at a point
__device__ __forceinline__ void dry_transport_pt(const PhysParams &p, const DryPoint &q, double &visc, double &bulk,
                                                 double &k) {
  const double temp = q.p * q.rinv / p.R;
  visc = p.C1 * p.visc_mult * (temp * sqrt(temp)) / (temp + p.S0);
  bulk = (p.bulk_visc_mult - 2. / 3.) * visc;
  k = p.cp_div_pr * visc;
}
// F_v(s, g).n ; gu[i + 3*d] = d u_i / d x_d, gT[d] = dT/dx_d ; fn[0] = 0
__device__ __forceinline__ void dry_visc_dot_n(const PhysParams &p, const DryPoint &q, const double *gu,
                                               const double *gT, const double *nor, double *fn) {
  double visc, bulk, k;
  dry_transport_pt(p, q, visc, bulk, k);
  const double divV = gu[0 + 3 * 0] + gu[1 + 3 * 1] + gu[2 + 3 * 2];
  double tn[DIM];
#pragma unroll
  for (int i = 0; i < DIM; i++) {
    double a = 0;
#pragma unroll
    for (int j = 0; j < DIM; j++) a += (gu[j + 3 * i] + gu[i + 3 * j]) * nor[j];
    tn[i] = visc * a + bulk * divV * nor[i];
  }
  fn[0] = 0.0;
  fn[1] = tn[0];
  fn[2] = tn[1];
  fn[3] = tn[2];
  fn[4] = q.vel[0] * tn[0] + q.vel[1] * tn[1] + q.vel[2] * tn[2] +
          k * (gT[0] * nor[0] + gT[1] * nor[1] + gT[2] * nor[2]);
}

}  // namespace tpsb
