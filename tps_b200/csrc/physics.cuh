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
  // Fluxes' sub-grid-scale model and planar viscous sponge (fluxes.cpp:224-246)
  int sgs_model;  // flow/sgsModel: 0 none, 1 smagorinsky, 2 sigma
  double sgs_const, sgs_floor;
  int sponge;  // viscosityMultiplierFunction/isEnabled
  double sp_n[3], sp_p[3], sp_ratio, sp_width;
};
// what the modified transport needs besides the state: delta = h_min / order of the element the gradient belongs to
// (Mesh::GetElementSize(e, 1) / order, rhs_operator.cpp:149-156, face_integrator.cpp:253-276) and the physical point
struct DryAux {
  double delta, x[3];
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
struct DryAux;
__host__ __device__ __forceinline__ void dry_modify_transport(const PhysParams &p, double rho, const double *gv, int ds, const DryAux &ax,
                                                     double &visc, double &bulk, double &k);
__device__ __forceinline__ void dry_visc_flux(const PhysParams &p, const double *s, const double *g, double *f,
                                              const DryAux *ax = nullptr) {
  const double pr = dry_pressure(p, s);
  const double temp = pr / p.R / s[0];
  // pow(temp, 1.5) of the reference evaluated as temp*sqrt(temp) (agrees to an ulp)
  double visc = (p.C1 * p.visc_mult * (temp * sqrt(temp)) / (temp + p.S0));
  double bulk = p.bulk_visc_mult * visc;
  double k = p.cp_div_pr * visc;
  bulk -= 2. / 3. * visc;
  if (ax && (p.sgs_model | p.sponge)) dry_modify_transport(p, s[0], g + 1, NEQ, *ax, visc, bulk, k);
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

// Fluxes::sgsSmag (fluxes.cpp:513-541): mu_sgs = rho (C_d max(delta - floor, 0))^2 sqrt(2 S_ij S_ij).
// gv[i + ds*d] = d u_i / d x_d
__host__ __device__ __forceinline__ double dry_sgs_smag(const PhysParams &p, double rho, const double *gv, int ds, double delta) {
  const double s3 = 0.5 * (gv[0 + ds] + gv[1]), s4 = 0.5 * (gv[0 + 2 * ds] + gv[2]), s5 = 0.5 * (gv[1 + 2 * ds] + gv[2 + ds]);
  double sm = 0.;
  sm += gv[0] * gv[0];
  sm += gv[1 + ds] * gv[1 + ds];
  sm += gv[2 + 2 * ds] * gv[2 + 2 * ds];
  sm += 2.0 * s3 * s3;
  sm += 2.0 * s4 * s4;
  sm += 2.0 * s5 * s5;
  sm = sqrt(2.0 * sm);
  const double dm = p.sgs_const * fmax(delta - p.sgs_floor, 0.0);
  return rho * dm * dm * sm;
}
// Fluxes::sgsSigma (fluxes.cpp:543-650), the branch without LAPACK (the one a device build and the oracle's
// reference object code run): singular values of grad u from the closed-form eigenvalues of d^4 g^T g
// (Nicoud et al. 2011); the reference's truncated pi and its 1e-12 guards are kept.
__host__ __device__ __noinline__ double dry_sgs_sigma(const PhysParams &p, double rho, const double *gv, int ds, double delta) {
  const double sml = 1.0e-12, pi = 3.14159265359, third = 1. / 3.;
  const double dm = fmax(delta - p.sgs_floor, sml);
  const double d4 = pow(dm, 4);
  double Q[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double a = 0;
      for (int k = 0; k < 3; k++) a += gv[k + ds * i] * gv[k + ds * j];
      Q[i][j] = a * d4;
    }
  const double p1 = Q[0][1] * Q[0][1] + Q[0][2] * Q[0][2] + Q[1][2] * Q[1][2];
  const double q = third * (Q[0][0] + Q[1][1] + Q[2][2]);
  const double p2 = (Q[0][0] - q) * (Q[0][0] - q) + (Q[1][1] - q) * (Q[1][1] - q) + (Q[2][2] - q) * (Q[2][2] - q) + 2.0 * p1;
  const double pp = sqrt(fmax(p2, 0.0) / 6.0);
  double B[3][3];
  const double ip = 1.0 / fmax(pp, sml);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) B[i][j] = (Q[i][j] - (i == j ? q : 0.0)) * ip;
  const double detB = B[0][0] * (B[1][1] * B[2][2] - B[2][1] * B[1][2]) - B[0][1] * (B[1][0] * B[2][2] - B[2][0] * B[1][2]) +
                      B[0][2] * (B[1][0] * B[2][1] - B[2][0] * B[1][1]);
  const double r = 0.5 * detB;
  const double phi = r <= -1.0 ? third * pi : (r >= 1.0 ? 0.0 : third * acos(r));
  const double e0 = q + 2.0 * pp * cos(phi), e2 = q + 2.0 * pp * cos(phi + (2.0 * third * pi));
  const double e1 = 3.0 * q - e0 - e2;
  const double g0 = sqrt(fmax(e0, sml)), g1 = sqrt(fmax(e1, sml)), g2 = sqrt(fmax(e2, sml));
  double mu = g2 * (g0 - g1) * (g1 - g2);
  mu = fmax(mu, 0.0);
  mu /= (g0 * g0);
  mu *= (p.sgs_const * p.sgs_const);
  mu *= rho;
  if (mu != mu) mu = 0.0;
  return mu;
}
// Fluxes::viscSpongePlanar (fluxes.cpp:664-684)
__host__ __device__ __forceinline__ double dry_sponge_weight(const PhysParams &p, const double *x) {
  const double factor = fmax(p.sp_ratio, 1.0);
  double dist = 0.;
#pragma unroll
  for (int d = 0; d < DIM; d++) dist += (x[d] - p.sp_p[d]) * p.sp_n[d];
  double wgt = 0.5 * (tanh(dist / p.sp_width - 2.0) + 1.0);
  wgt *= (factor - 1.0);
  wgt += 1.0;
  return wgt;
}
// the modification block of Fluxes::ComputeViscousFluxes / ComputeBdrViscousFluxes (fluxes.cpp:224-246, 386-407):
// bulk arrives as (bulk_mult - 2/3) visc
__host__ __device__ __forceinline__ void dry_modify_transport(const PhysParams &p, double rho, const double *gv, int ds,
                                                              const DryAux &ax, double &visc, double &bulk, double &k) {
  if (p.sgs_model > 0) {
    const double pr_cp = visc / k;
    const double mu_sgs = p.sgs_model == 1 ? dry_sgs_smag(p, rho, gv, ds, ax.delta) : dry_sgs_sigma(p, rho, gv, ds, ax.delta);
    bulk *= (1.0 + mu_sgs / visc);
    visc += mu_sgs;
    k += (mu_sgs / pr_cp);
  }
  if (p.sponge) {
    const double wgt = dry_sponge_weight(p, ax.x);
    visc *= wgt;
    bulk *= wgt;
    k *= wgt;
  }
}

// branch-free double reciprocal / square root: MUFU seed + Newton steps, ~1 ulp.  The MUFU.RCP64H /
// MUFU.RSQ64H seeds carry only ~9 good bits (the library routines spend 5-6 DFMAs on them as well), so
// one cubic step (error e^3) plus one quadratic step (e^6) are needed.  What is saved against the library
// divide / sqrt is the slow-path branch (BSSY/BSYNC + call) per use; arguments here are positive, O(1)-scaled
// physical quantities (rho, T, p/rho, |v|^2 >= 0).
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, fma(e, e, e), r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}
// sqrt(x) for x >= 0 (returns 0 at 0)
__device__ __forceinline__ double fast_sqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(fmax(x, 1e-300)));
  // r <- r (1 + e/2 + 3 e^2/8), e = 1 - x r^2 (cubic), one Newton step, then Heron on s = x r
  double e = fma(-x * r, r, 1.0);
  r = fma(r * e, fma(0.375, e, 0.5), r);
  e = fma(-x * r, r, 1.0);
  r = fma(0.5 * r, e, r);
  double s = x * r;
  s = fma(fma(-s, s, x), 0.5 * r, s);
  return s;
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
  q.rinv = fast_rcp(s[0]);
  q.vel[0] = s[1] * q.rinv;
  q.vel[1] = s[2] * q.rinv;
  q.vel[2] = s[3] * q.rinv;
  q.p = p.gm1 * (s[4] - 0.5 * (s[1] * q.vel[0] + s[2] * q.vel[1] + s[3] * q.vel[2]));
  return q;
}
__device__ __forceinline__ double dry_char_speed_pt(const PhysParams &p, const DryPoint &q) {
  return fast_sqrt(q.vel[0] * q.vel[0] + q.vel[1] * q.vel[1] + q.vel[2] * q.vel[2]) + fast_sqrt(p.gamma * q.p * q.rinv);
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
  visc = p.C1 * p.visc_mult * (temp * fast_sqrt(temp)) * fast_rcp(temp + p.S0);
  bulk = (p.bulk_visc_mult - 2. / 3.) * visc;
  k = p.cp_div_pr * visc;
}
// transport coefficients with the SGS model / sponge applied when ax is given
__device__ __forceinline__ void dry_visc_coeffs(const PhysParams &p, const DryPoint &q, const double *gu, const DryAux *ax,
                                                double rho, double &visc, double &bulk, double &k) {
  dry_transport_pt(p, q, visc, bulk, k);
  if (ax && (p.sgs_model | p.sponge)) dry_modify_transport(p, rho, gu, 3, *ax, visc, bulk, k);
}
// F_v(s, g).n ; gu[i + 3*d] = d u_i / d x_d, gT[d] = dT/dx_d ; fn[0] = 0
__device__ __forceinline__ void dry_visc_dot_n_c(const DryPoint &q, double visc, double bulk, double k, const double *gu,
                                                 const double *gT, const double *nor, double *fn) {
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
__device__ __forceinline__ void dry_visc_dot_n(const PhysParams &p, const DryPoint &q, const double *gu,
                                               const double *gT, const double *nor, double *fn, const DryAux *ax = nullptr,
                                               double rho = 0.0) {
  double visc, bulk, k;
  dry_visc_coeffs(p, q, gu, ax, rho, visc, bulk, k);
  dry_visc_dot_n_c(q, visc, bulk, k, gu, gT, nor, fn);
}

#define SLIPFN __host__ __device__ __noinline__
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

// ---- boundary conditions (dry air) ----------------------------------------------------------------
constexpr int MAX_BC = 8;
struct BcDev {
  int kind, type;  // kind: 0 inlet, 1 outlet, 2 wall; type: InletType / OutletType / WallType (dataStructures.hpp:168-196)
  double d[4];
};
struct BcTable {
  int nbc, use_bc_in_grad;
  BcDev bc[MAX_BC];
};

// BoundaryCondition::computeBdrPrimitiveStateForGradient (BoundaryCondition.cpp:55) / WallBC override
// (wallBC.cpp:241-266): only an isothermal wall changes the state used by the BR1 jump.
__device__ __forceinline__ void dry_bc_prim_for_gradient(const BcDev &bc, const double *primIn, double *primBC) {
#pragma unroll
  for (int eq = 0; eq < NEQ; eq++) primBC[eq] = primIn[eq];
  if (bc.kind == 2 && bc.type == 3) {
    primBC[1] = 0.0;
    primBC[2] = 0.0;
    primBC[3] = 0.0;
    primBC[4] = bc.d[0];
  }
}

// Fluxes::ComputeBdrViscousFluxes (fluxes.cpp:344-504) for dry air: one species with zero diffusion velocity,
// no species-enthalpy term, single temperature.  nrm = unit normal; heat_prescribed: primFluxIdxs[numSpecies+nvel]
// with value 0 (adiabatic wall).  g[eq + d*NEQ].
__device__ __forceinline__ void dry_bdr_visc_flux(const PhysParams &p, const double *s, const double *g, const double *nrm,
                                                  bool heat_prescribed, double *nf, const DryAux *ax = nullptr) {
#pragma unroll
  for (int eq = 0; eq < NEQ; eq++) nf[eq] = 0.;
  if (p.eq_system == 0) return;
  const double pr = dry_pressure(p, s);
  const double temp = pr / p.R / s[0];
  double visc = (p.C1 * p.visc_mult * (temp * sqrt(temp)) / (temp + p.S0));
  double bulk = p.bulk_visc_mult * visc;
  double k = p.cp_div_pr * visc;
  bulk -= 2. / 3. * visc;
  if (ax && (p.sgs_model | p.sponge)) dry_modify_transport(p, s[0], g + 1, NEQ, *ax, visc, bulk, k);
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
  double sn[DIM], q = 0.;
#pragma unroll
  for (int i = 0; i < DIM; i++) {
    sn[i] = 0.;
#pragma unroll
    for (int j = 0; j < DIM; j++) sn[i] += stress[i + j * DIM] * nrm[j];
  }
#pragma unroll
  for (int d = 0; d < DIM; d++) q -= k * g[4 + d * NEQ] * nrm[d];
  if (heat_prescribed) q = 0.;
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    nf[1 + d] = sn[d];
    nf[4] += sn[d] * (s[1 + d] / s[0]);
  }
  nf[4] -= q;
}

// BCintegrator::computeBdrFlux (BCintegrator.cpp:228-242) -> InletBC::subsonicReflectingDensityVelocity
// (inletBC.cpp:729-756), OutletBC::subsonicReflectingPressure (outletBC.cpp:731-737), WallBC::computeINVwallFlux /
// computeAdiabaticWallFlux / computeIsothermalWallFlux (wallBC.cpp:277-320, 430-510).  g[eq + d*NEQ] interior
// gradients of the primitives; nor = CalcOrtho normal (area weighted, outward).
__device__ __forceinline__ void dry_bc_flux(const PhysParams &p, const BcDev &bc, int use_bc_in_grad, const double *u1,
                                            const double *g, const double *nor, double *fx, const DryAux *ax = nullptr) {
  double s2[NEQ];
  const double normN = nor[0] * nor[0] + nor[1] * nor[1] + nor[2] * nor[2];
  if (bc.kind == 0) {
    const double pr = dry_pressure(p, u1);
    s2[0] = bc.d[0];
    s2[1] = bc.d[0] * bc.d[1];
    s2[2] = bc.d[0] * bc.d[2];
    s2[3] = bc.d[0] * bc.d[3];
    double ke = s2[1] * s2[1] + s2[2] * s2[2] + s2[3] * s2[3];
    ke *= 0.5 / s2[0];
    s2[4] = pr / p.gm1 + ke;  // DryAir::modifyEnergyForPressure (equation_of_state.cpp:402-411)
    dry_riemann_lf(p, u1, s2, nor, fx);
    return;
  }
  if (bc.kind == 1) {
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) s2[eq] = u1[eq];
    double ke = u1[1] * u1[1] + u1[2] * u1[2] + u1[3] * u1[3];
    ke *= 0.5 / u1[0];
    s2[4] = bc.d[0] / p.gm1 + ke;
    dry_riemann_lf(p, u1, s2, nor, fx);
    return;
  }
  if (bc.type == 1) {  // SLIP (wallBC.cpp:326-428): mirror state, Riemann flux only
    const double vel[3] = {u1[1] / u1[0], u1[2] / u1[0], u1[3] / u1[0]};
    double nVel[3];
    slip_mirror_velocity(3, nor, vel, nVel);
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) s2[eq] = u1[eq];
    s2[1] = u1[0] * nVel[0];
    s2[2] = u1[0] * nVel[1];
    s2[3] = u1[0] * nVel[2];
    dry_riemann_lf(p, u1, s2, nor, fx);
    return;
  }
  double viscF[NEQ * DIM], wallViscF[NEQ];
  if (bc.type == 0) {  // INV: mirror state
    const double norm = sqrt(normN);
    const double un[3] = {nor[0] / norm, nor[1] / norm, nor[2] / norm};
    const double vel[3] = {u1[1] / u1[0], u1[2] / u1[0], u1[3] / u1[0]};
    const double vn = vel[0] * un[0] + vel[1] * un[1] + vel[2] * un[2];
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) s2[eq] = u1[eq];
    s2[1] = u1[0] * (vel[0] - 2. * vn * un[0]);
    s2[2] = u1[0] * (vel[1] - 2. * vn * un[1]);
    s2[3] = u1[0] * (vel[2] - 2. * vn * un[2]);
    dry_riemann_lf(p, u1, s2, nor, fx);
    if (p.eq_system == 0) return;
    dry_visc_flux(p, s2, g, viscF, ax);
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++)
      wallViscF[eq] = viscF[eq] * nor[0] + viscF[eq + NEQ] * nor[1] + viscF[eq + 2 * NEQ] * nor[2];
    dry_visc_flux(p, u1, g, viscF, ax);
  } else {
    const double isq = 1. / sqrt(normN);
    const double un[3] = {nor[0] * isq, nor[1] * isq, nor[2] * isq};
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) s2[eq] = u1[eq];
    if (bc.type == 2) {  // VISC_ADIAB: stagnation state (equation_of_state.cpp:365-377)
      const double pr = dry_pressure(p, u1);
      s2[1] = s2[2] = s2[3] = 0.;
      s2[4] = pr / p.gm1;
      dry_riemann_lf(p, u1, s2, nor, fx);
      if (p.eq_system == 0) return;
      dry_bdr_visc_flux(p, s2, g, un, true, wallViscF, ax);
    } else {  // VISC_ISOTH
      if (use_bc_in_grad) {
        s2[1] = -u1[1];
        s2[2] = -u1[2];
        s2[3] = -u1[3];
      } else {
        s2[1] = s2[2] = s2[3] = 0.;
        s2[4] = p.R / p.gm1 * u1[0] * bc.d[0];  // computeStagnantStateWithTemp (equation_of_state.cpp:379-386)
      }
      dry_riemann_lf(p, u1, s2, nor, fx);
      if (p.eq_system == 0) return;
      s2[1] = s2[2] = s2[3] = 0.;
      s2[4] = p.R / p.gm1 * u1[0] * bc.d[0];
      dry_bdr_visc_flux(p, s2, g, un, false, wallViscF, ax);
    }
    const double nm = sqrt(normN);
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) wallViscF[eq] *= nm;
    dry_visc_flux(p, u1, g, viscF, ax);
  }
#pragma unroll
  for (int eq = 1; eq < NEQ; eq++) {
    fx[eq] -= 0.5 * wallViscF[eq];
    fx[eq] -= 0.5 * (viscF[eq] * nor[0] + viscF[eq + NEQ] * nor[1] + viscF[eq + 2 * NEQ] * nor[2]);
  }
}

}  // namespace tpsb
