// Per-point physics with RUN-TIME dimension / equation count for the generic tensor-product path
// (rhs_generic.cuh): 2-D and 3-D, nvel = dim (or 3 when axisymmetric), num_equation <= GEN_MAXEQ.
// Dry air here; the mixture models (mix_physics.cuh) plug in behind the same entry points.
// Formulas and operation order follow the reference routines cited at each function.
#pragma once
#include <cuda_runtime.h>

#include "mix_physics.cuh"
#include "physics.cuh"

namespace tpsb {

constexpr int GEN_MAXEQ = 12;  // gpudata::MAXEQUATIONS of the reference's CUDA build (src/dataStructures.hpp:52)
constexpr int GEN_MAXDIM = 3;

// LteMixture / LteTransport over 1-D linear tables (flow/lte/table_dim = 1: lte_mixture.cpp, lte_transport_properties.cpp,
// LinearTable table.cpp:52-116): one species, one temperature; energy, gas constant, speed of sound, viscosity and
// conductivity are functions of T alone.  Tables live in one device pool: x[n] | a[n-1] | b[n-1] per table.
constexpr int LTE_E = 0, LTE_R = 1, LTE_C = 2, LTE_T = 3, LTE_MU = 4, LTE_KAPPA = 5, LTE_NEC = 6, LTE_NTAB = 7;
struct LteParams {
  int n[LTE_NTAB];  // rows; n[LTE_NEC] == 0: no radiation
  int xlog[LTE_NTAB], flog[LTE_NTAB];
  const double *x[LTE_NTAB];
};
// TableInterpolator::findInterval + LinearTable::eval / eval_x (table.cpp:52-116)
__host__ __device__ __forceinline__ int lte_interval(const LteParams &L, int t, double xEval) {
  const int n = L.n[t];
  const double *x = L.x[t];
  int count = n, first = 0;
  while (count > 0) {  // std::upper_bound
    int it = first;
    const int step = count / 2;
    it += step;
    if (xEval > x[it]) {
      first = ++it;
      count -= step + 1;
    } else {
      count = step;
    }
  }
  first = first < n - 1 ? first : n - 1;
  first = first > 1 ? first : 1;
  return first - 1;
}
__host__ __device__ __forceinline__ double lte_eval(const LteParams &L, int t, double xEval) {
  const int i = lte_interval(L, t, xEval), n = L.n[t];
  const double *ta = L.x[t] + n, *tb = ta + (n - 1);
  const double xt = L.xlog[t] ? log(xEval) : xEval;
  double ft = ta[i] + tb[i] * xt;
  if (L.flog[t]) ft = exp(ft);
  return ft;
}
__host__ __device__ __forceinline__ double lte_eval_x(const LteParams &L, int t, double xEval) {
  const int i = lte_interval(L, t, xEval), n = L.n[t];
  const double *ta = L.x[t] + n, *tb = ta + (n - 1);
  const double xt = L.xlog[t] ? log(xEval) : xEval;
  const double xt_xt = L.xlog[t] ? 1. / xEval : 1.0;
  double ft_x = tb[i] * xt_xt;
  if (L.flog[t]) ft_x *= exp(ta[i] + tb[i] * xt);
  return ft_x;
}
// LteMixture::ComputeTemperatureInternal (lte_mixture.cpp:161-223): Newton on e(T) = e(U) from the T(e) table's guess
__host__ __device__ __forceinline__ double lte_temperature(const LteParams &L, int nvel, const double *state) {
  const double rho = state[0];
  double den_vel2 = 0;
  for (int d = 0; d < nvel; d++) den_vel2 += state[d + 1] * state[d + 1];
  den_vel2 /= rho;
  const double energy = (state[1 + nvel] - 0.5 * den_vel2) / rho;
  double T = lte_eval(L, LTE_T, energy);
  double res = energy - lte_eval(L, LTE_E, T);
  const double res0 = fabs(res);
  const double atol = 1e-18, rtol = 1e-12, dT_atol = 1e-12, dT_rtol = 1e-8;
  bool converged = ((fabs(res) < atol) || (fabs(res) / fabs(res0) < rtol));
  int niter = 0;
  while (!converged && (niter < 20)) {
    const double dedT = lte_eval_x(L, LTE_E, T);
    const double dT = res / dedT;
    T += dT;
    res = energy - lte_eval(L, LTE_E, T);
    converged = ((fabs(res) < atol) || (fabs(res) / res0 < rtol) || (fabs(dT) < dT_atol) || (fabs(dT) / T < dT_rtol));
    niter++;
  }
  return T;
}
// LteMixture::ComputeTemperatureFromDensityPressure (lte_mixture.cpp:241-305): Newton on p = rho R(T) T from T = p / (208 rho)
__host__ __device__ __forceinline__ double lte_temperature_rho_p(const LteParams &L, double rho, double p) {
  double T = p / (rho * 208.);
  double R = lte_eval(L, LTE_R, T);
  double res = p - rho * R * T;
  const double res0 = fabs(res);
  const double atol = 1e-18, rtol = 1e-12, dT_atol = 1e-12, dT_rtol = 1e-8;
  bool converged = ((fabs(res) < atol) || (fabs(res) / fabs(res0) < rtol));
  int niter = 0;
  while (!converged && (niter < 20)) {
    const double R_T = lte_eval_x(L, LTE_R, T);
    const double dpdT = rho * R + rho * R_T * T;
    const double dT = res / dpdT;
    T += dT;
    R = lte_eval(L, LTE_R, T);
    res = p - rho * R * T;
    converged = ((fabs(res) < atol) || (fabs(res) / res0 < rtol) || (fabs(dT) < dT_atol) || (fabs(dT) / T < dT_rtol));
    niter++;
  }
  return T;
}

struct GenPhys {
  int dim, nvel, neq;
  int use_roe;           // flow/useRoe (2-D dry air): Eval_Roe unless a caller forces Lax-Friedrichs
  int axisym;            // config.isAxisymmetric(): dim == 2 with nvel == 3, state [rho, rho u_r, rho u_z, rho u_theta, ...]
  int fluid;             // 0 dry air; 1 user-defined plasma mixture (mix != NULL)
  PhysParams dry;        // gamma, R, Sutherland, multipliers, eq_system
  const MixParams *mix;  // device (or host, for host-side checks) pointer
  // flow/useMixingLength for dry air (mixtures carry it in MixParams): MixingLengthTransport (mixing_length_transport.cpp)
  int ml_on;
  double ml_max, ml_prt, ml_bulk;
  // LTE_FLUID with 1-D tables: fluid stays 0 (one species, the dry-air code structure) and every equation-of-state /
  // transport evaluation below switches to the tables
  const LteParams *lte;
};

// DryAir::ComputePressure (equation_of_state.hpp:610-617)
__host__ __device__ __forceinline__ double dry_gen_pressure(const GenPhys &g, const double *s) {
  if (g.lte) {  // LteMixture::ComputePressure (lte_mixture.cpp:123-135)
    const double T = lte_temperature(*g.lte, g.nvel, s);
    return s[0] * lte_eval(*g.lte, LTE_R, T) * T;
  }
  double den_vel2 = 0;
  for (int d = 0; d < g.nvel; d++) den_vel2 += s[d + 1] * s[d + 1];
  den_vel2 /= s[0];
  return g.dry.gm1 * (s[1 + g.nvel] - 0.5 * den_vel2);
}

// DryAir::GetPrimitivesFromConservatives (equation_of_state.cpp:321-335)
__host__ __device__ __forceinline__ void dry_gen_prim(const GenPhys &g, const double *s, double *up) {
  double den_vel2 = 0;
  for (int d = 0; d < g.nvel; d++) den_vel2 += s[d + 1] * s[d + 1];
  den_vel2 /= s[0];
  const double T = g.lte ? lte_temperature(*g.lte, g.nvel, s)  // LteMixture::GetPrimitivesFromConservatives (:319-330)
                         : g.dry.gm1 / g.dry.R * (s[1 + g.nvel] - 0.5 * den_vel2) / s[0];
  for (int eq = 0; eq < g.neq; eq++) up[eq] = s[eq];
  for (int d = 0; d < g.nvel; d++) up[1 + d] = s[1 + d] / s[0];
  up[1 + g.nvel] = T;
}

// DryAir::ComputeMaxCharSpeed (equation_of_state.cpp:279-294)
__host__ __device__ __forceinline__ double dry_gen_max_char_speed(const GenPhys &g, const double *s) {
  const double den = s[0];
  double den_vel2 = 0;
  for (int d = 0; d < g.nvel; d++) den_vel2 += s[d + 1] * s[d + 1];
  den_vel2 /= den;
  if (g.lte)  // LteMixture::ComputeMaxCharSpeed / ComputeSpeedOfSound (lte_mixture.cpp:366-401)
    return sqrt(den_vel2 / den) + lte_eval(*g.lte, LTE_C, lte_temperature(*g.lte, g.nvel, s));
  const double pres = g.dry.gm1 * (s[1 + g.nvel] - 0.5 * den_vel2);
  return sqrt(den_vel2 / den) + sqrt(g.dry.gamma * pres / den);
}

// Fluxes::ComputeConvectiveFluxes (fluxes.cpp:135-170): f[eq + d*neq]
__host__ __device__ __forceinline__ void dry_gen_conv_flux(const GenPhys &g, const double *s, double *f) {
  const double pres = dry_gen_pressure(g, s);
  const int neq = g.neq;
  for (int d = 0; d < g.dim; d++) {
    f[0 + d * neq] = s[d + 1];
    for (int i = 0; i < g.nvel; i++) f[1 + i + d * neq] = s[i + 1] * s[d + 1] / s[0];
    f[1 + d + d * neq] += pres;
  }
  const double H = (s[1 + g.nvel] + pres) / s[0];
  for (int d = 0; d < g.dim; d++) f[1 + g.nvel + d * neq] = s[d + 1] * H;
}

// DryAirTransport::ComputeFluxMolecularTransport (transport_properties.cpp:223-234: Sutherland viscosity, bulk and Prandtl
// conductivity) or LteTransport::ComputeFluxMolecularTransport (lte_transport_properties.cpp:86-108: tables, no bulk viscosity)
__host__ __device__ __forceinline__ void dry_gen_transport(const GenPhys &g, const double *s, double &visc, double &bulk, double &k) {
  if (g.lte) {
    const double T = lte_temperature(*g.lte, g.nvel, s);
    visc = lte_eval(*g.lte, LTE_MU, T);
    k = lte_eval(*g.lte, LTE_KAPPA, T);
    bulk = 0.0;
    return;
  }
  const double pr = dry_gen_pressure(g, s);
  const double temp = pr / g.dry.R / s[0];
  visc = (g.dry.C1 * g.dry.visc_mult * (temp * sqrt(temp)) / (temp + g.dry.S0));
  bulk = g.dry.bulk_visc_mult * visc;
  k = g.dry.cp_div_pr * visc;
}

// Fluxes::ComputeViscousFluxes (fluxes.cpp:178-335) with DryAirTransport (transport_properties.cpp:223-234);
// no SGS, no sponge.  gr[eq + d*neq] = d Up_eq / d x_d; radius = x[0] of the point (axisymmetric terms only).
// Written with fixed 3x3 register tiles and guards instead of run-time-indexed local arrays (entries beyond
// `dim` are zero, so sums keep the reference's operation order): nvcc 12.9 -O3 produced wrong stores for the
// run-time-indexed form once inlined into gen_resid_kernel (caught by the parity test; the stand-alone
// function was correct, tools/ubench/gen_visc_check.cu).
// ax != NULL: Fluxes' sub-grid-scale model / planar viscous sponge (fluxes.cpp:224-246) with the element size delta and the
// physical point the caller supplies (the SGS models read a 3 x 3 velocity gradient: dim == 3 only, checked at create).
__host__ __device__ __forceinline__ void dry_gen_visc_flux(const GenPhys &g, const double *s, const double *gr, double radius,
                                                           double *f, double distance = 0.0, const DryAux *ax = nullptr) {
  const int neq = g.neq, dim = g.dim;
  for (int i = 0; i < neq * dim; i++) f[i] = 0.;
  if (g.dry.eq_system == 0) return;
  double visc, bulk, k;
  dry_gen_transport(g, s, visc, bulk, k);
  if (g.ml_on) mixlen_add(dim, g.nvel, neq, g.ml_max, g.ml_prt, g.ml_bulk, s, gr, radius, distance, visc, bulk, k);
  bulk -= 2. / 3. * visc;
  if (ax && (g.dry.sgs_model | g.dry.sponge)) dry_modify_transport(g.dry, s[0], gr + 1, neq, *ax, visc, bulk, k);
  double gu[3][3], st[3][3], vel[3], gT[3];  // gu[i][d] = d u_i / d x_d
#pragma unroll
  for (int i = 0; i < 3; i++) {
    vel[i] = i < dim ? s[1 + i] / s[0] : 0.0;
    gT[i] = i < dim ? gr[(1 + g.nvel) + i * neq] : 0.0;
#pragma unroll
    for (int d = 0; d < 3; d++) gu[i][d] = (i < dim && d < dim) ? gr[(1 + i) + d * neq] : 0.0;
  }
  double divV = 0.;
#pragma unroll
  for (int i = 0; i < 3; i++) divV += gu[i][i];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      st[i][j] = gu[j][i] + gu[i][j];
      st[i][j] *= visc;
    }
  const double ur = g.axisym ? s[1] / s[0] : 0.0, ut = g.axisym ? s[3] / s[0] : 0.0;
  if (g.axisym && radius > 0) divV += ur / radius;
#pragma unroll
  for (int i = 0; i < 3; i++) st[i][i] += bulk * divV;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    if (j < dim) {
      double vtmp = 0.0;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        if (i < dim) f[(1 + i) + j * neq] = st[i][j];
        vtmp += st[j][i] * vel[i];
      }
      f[(1 + g.nvel) + j * neq] = vtmp + k * gT[j];
    }
  }
  if (g.axisym) {  // fluxes.cpp:285-297, 320-323
    double tau_tr = gr[3 + 0 * neq];
    if (radius > 0) tau_tr -= ut / radius;
    tau_tr *= visc;
    const double tau_tz = visc * gr[3 + 1 * neq];
    f[(1 + 2) + 0 * neq] = tau_tr;
    f[(1 + 2) + 1 * neq] = tau_tz;
    f[(1 + g.nvel) + 0 * neq] += ut * tau_tr;
    f[(1 + g.nvel) + 1 * neq] += ut * tau_tz;
  }
}

// Fluxes::ComputeBdrViscousFluxes (fluxes.cpp:344-504) for dry air: one species with zero diffusion velocity, no
// species-enthalpy term, single temperature.  nrm = unit normal; heat_prescribed: primFluxIdxs[numSpecies + nvel]
// (prescribed value 0: adiabatic wall).
__host__ __device__ __forceinline__ void dry_gen_bdr_visc_flux(const GenPhys &g, const double *s, const double *gr, double radius,
                                                               const double *nrm, bool heat_prescribed, double *nf,
                                                               const DryAux *ax = nullptr) {
  const int neq = g.neq, dim = g.dim, nvel = g.nvel;
  for (int eq = 0; eq < neq; eq++) nf[eq] = 0.;
  if (g.dry.eq_system == 0) return;
  double visc, bulk, k;
  dry_gen_transport(g, s, visc, bulk, k);
  bulk -= 2. / 3. * visc;
  if (ax && (g.dry.sgs_model | g.dry.sponge)) dry_modify_transport(g.dry, s[0], gr + 1, neq, *ax, visc, bulk, k);  // fluxes.cpp:386-407
  double gu[3][3], st[3][3], nn[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    nn[i] = i < dim ? nrm[i] : 0.0;
#pragma unroll
    for (int d = 0; d < 3; d++) gu[i][d] = (i < dim && d < dim) ? gr[(1 + i) + d * neq] : 0.0;
  }
  double divV = 0.;
#pragma unroll
  for (int i = 0; i < 3; i++) divV += gu[i][i];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      st[i][j] = gu[j][i] + gu[i][j];
      st[i][j] *= visc;
    }
  const double ur = g.axisym ? s[1] / s[0] : 0.0, ut = g.axisym ? s[3] / s[0] : 0.0;
  if (g.axisym && radius > 0) divV += ur / radius;
#pragma unroll
  for (int i = 0; i < 3; i++) st[i][i] += bulk * divV;
  double pf[3] = {0, 0, 0};  // normalPrimFlux[numSpecies + i]
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      if (i < dim && j < dim) pf[i] += st[i][j] * nn[j];
  if (g.axisym) {
    double tau_tr = gr[3 + 0 * neq];
    if (radius > 0) tau_tr -= ut / radius;
    tau_tr *= visc;
    const double tau_tz = visc * gr[3 + 1 * neq];
    pf[2] += tau_tr * nn[0];
    pf[2] += tau_tz * nn[1];
  }
  double q = 0.;
  for (int d = 0; d < dim; d++) q -= k * gr[(1 + nvel) + d * neq] * nrm[d];
  if (heat_prescribed) q = 0.;
  for (int d = 0; d < nvel; d++) nf[d + 1] = pf[d];
  for (int d = 0; d < nvel; d++) nf[nvel + 1] += pf[d] * (s[1 + d] / s[0]);
  nf[nvel + 1] -= q;
}

// ---- dispatch on the working fluid (the reference's virtual GasMixture / TransportProperties calls) ----
__host__ __device__ __forceinline__ void gen_prim(const GenPhys &g, const double *s, double *up) {
  if (g.fluid) mix_prim(*g.mix, s, up); else dry_gen_prim(g, s, up);
}
__host__ __device__ __forceinline__ double gen_max_char_speed(const GenPhys &g, const double *s) {
  return g.fluid ? mix_max_char_speed(*g.mix, s) : dry_gen_max_char_speed(g, s);
}
__host__ __device__ __forceinline__ void gen_conv_flux(const GenPhys &g, const double *s, double *f) {
  if (g.fluid) mix_conv_flux(*g.mix, s, f); else dry_gen_conv_flux(g, s, f);
}
// distance: wall distance at the point (mixing-length model only; the viscous walls pass 0 as the reference does)
__host__ __device__ __forceinline__ void gen_visc_flux(const GenPhys &g, const double *s, const double *gr, double radius,
                                                       double *f, double distance = 0.0, const DryAux *ax = nullptr) {
  if (g.fluid) mix_visc_flux(*g.mix, s, gr, radius, f, distance, &g.dry, ax);
  else dry_gen_visc_flux(g, s, gr, radius, f, distance, ax);
}
__host__ __device__ __forceinline__ void gen_bdr_visc_flux(const GenPhys &g, const double *s, const double *gr, double radius,
                                                           const double *nrm, bool heat_prescribed, double *nf,
                                                           const DryAux *ax = nullptr) {
  if (g.fluid) mix_bdr_visc_flux(*g.mix, s, gr, radius, nrm, heat_prescribed, nf, &g.dry, ax);
  else dry_gen_bdr_visc_flux(g, s, gr, radius, nrm, heat_prescribed, nf, ax);
}
__host__ __device__ __forceinline__ double gen_pressure(const GenPhys &g, const double *s) {
  return g.fluid ? mix_pressure(*g.mix, s, nullptr) : dry_gen_pressure(g, s);
}
// GasMixture::ComputePressureFromPrimitives (equation_of_state.cpp:360-364, 988-1010)
__host__ __device__ __forceinline__ double gen_pressure_from_prim(const GenPhys &g, const double *up) {
  if (g.lte) return up[0] * lte_eval(*g.lte, LTE_R, up[1 + g.nvel]) * up[1 + g.nvel];  // lte_mixture.cpp:142-151
  return g.fluid ? mix_pressure_from_prim(*g.mix, up) : g.dry.R * up[0] * up[1 + g.nvel];
}
// TransportProperties::GetViscosities (transport_properties.hpp:264-269, 305-309): visc[0] shear, visc[1] bulk
__host__ __device__ __forceinline__ void gen_viscosities(const GenPhys &g, const double *U, const double *up, double *visc) {
  if (g.fluid) {
    mix_viscosities(*g.mix, U, up, visc);
  } else {
    const double temp = up[1 + g.nvel];
    if (g.lte) {  // LteTransport::GetViscosities (lte_transport_properties.cpp:125-136)
      visc[0] = lte_eval(*g.lte, LTE_MU, temp);
      visc[1] = 0.;
      return;
    }
    visc[0] = (g.dry.C1 * g.dry.visc_mult * pow(temp, 1.5) / (temp + g.dry.S0));
    visc[1] = g.dry.bulk_visc_mult * visc[0];
  }
}
// GasMixture::GetConservativesFromPrimitives (DryAir equation_of_state.cpp:298-315, PerfectMixture :744-783)
__host__ __device__ __forceinline__ void gen_cons(const GenPhys &g, const double *up, double *U) {
  if (g.fluid) {
    mix_cons(*g.mix, up, U);
  } else {
    for (int eq = 0; eq < g.neq; eq++) U[eq] = up[eq];
    double v2 = 0.;
    for (int d = 0; d < g.nvel; d++) {
      v2 += up[1 + d] * up[1 + d];
      U[1 + d] *= up[0];
    }
    if (g.lte)  // LteMixture::GetConservativesFromPrimitives (lte_mixture.cpp:337-359)
      U[1 + g.nvel] = up[0] * (lte_eval(*g.lte, LTE_E, up[1 + g.nvel]) + 0.5 * v2);
    else
      U[1 + g.nvel] = g.dry.R * up[0] * up[1 + g.nvel] / g.dry.gm1 + 0.5 * up[0] * v2;
  }
}
// GasMixture::computeStagnationState (equation_of_state.cpp:100-116; DryAir :367-377)
__host__ __device__ __forceinline__ void gen_stagnation_state(const GenPhys &g, const double *in, double *out) {
  for (int eq = 0; eq < g.neq; eq++) out[eq] = in[eq];
  if (g.fluid || g.lte) {  // LteMixture keeps the base-class version
    double ke = 0.0;
    for (int d = 0; d < g.nvel; d++) ke += 0.5 * in[1 + d] * in[1 + d] / in[0];
    for (int d = 0; d < g.nvel; d++) out[1 + d] = 0.;
    out[1 + g.nvel] = in[1 + g.nvel] - ke;
  } else {
    const double p = dry_gen_pressure(g, in);
    for (int d = 0; d < g.nvel; d++) out[1 + d] = 0.;
    out[1 + g.nvel] = p / g.dry.gm1;
  }
}
// GasMixture::computeStagnantStateWithTemp (DryAir :380-387, PerfectMixture :1596-1620)
__host__ __device__ __forceinline__ void gen_stagnant_state_with_temp(const GenPhys &g, const double *in, double T, double *out) {
  if (g.fluid) {
    mix_stagnant_state_with_temp(*g.mix, in, T, out);
  } else {
    for (int eq = 0; eq < g.neq; eq++) out[eq] = in[eq];
    for (int d = 0; d < g.nvel; d++) out[1 + d] = 0.;
    out[1 + g.nvel] = g.lte ? in[0] * lte_eval(*g.lte, LTE_E, T)  // lte_mixture.cpp:430-446
                            : g.dry.R / g.dry.gm1 * in[0] * T;
  }
}
// GasMixture::modifyEnergyForPressure (DryAir :402-411, PerfectMixture :1698-1741); in and out may alias
__host__ __device__ __forceinline__ void gen_modify_energy_for_pressure(const GenPhys &g, const double *in, double *out, double p,
                                                                        bool modifyElectronEnergy) {
  if (g.fluid) {
    mix_modify_energy_for_pressure(*g.mix, in, out, p, modifyElectronEnergy);
  } else {
    double ke = 0.;
    for (int d = 0; d < g.nvel; d++) ke += in[1 + d] * in[1 + d];
    ke *= 0.5 / in[0];
    const double rho = in[0];
    for (int eq = 0; eq < g.neq; eq++) out[eq] = in[eq];
    if (g.lte)  // LteMixture::modifyEnergyForPressure (lte_mixture.cpp:453-470)
      out[1 + g.nvel] = rho * lte_eval(*g.lte, LTE_E, lte_temperature_rho_p(*g.lte, rho, p)) + ke;
    else
      out[1 + g.nvel] = p / g.dry.gm1 + ke;
  }
}
__host__ __device__ __forceinline__ int gen_num_active_species(const GenPhys &g) { return g.fluid ? g.mix->numActive : 0; }

// RiemannSolverTPS::Eval_Roe (riemann_solver.cpp:117-206), Roe-Lohner: two velocity components, gamma - 1 = 0.4
// hard-coded (:153), entropy fix |lambda_0| >= 1e-4 -- restated as is (parity trap 9).
__host__ __device__ __noinline__ void dry_gen_riemann_roe(const GenPhys &g, const double *state1, const double *state2,
                                                          const double *nor, double *flux) {
  const int neq = g.neq;
  const double normag = sqrt(nor[0] * nor[0] + nor[1] * nor[1]);
  const double unitN[2] = {nor[0] / normag, nor[1] / normag};
  double f1[GEN_MAXEQ * GEN_MAXDIM], f2[GEN_MAXEQ * GEN_MAXDIM], meanFlux[4];
  dry_gen_conv_flux(g, state1, f1);
  dry_gen_conv_flux(g, state2, f2);
  for (int eq = 0; eq < 4; eq++) {
    meanFlux[eq] = 0.;
    for (int d = 0; d < 2; d++) meanFlux[eq] += (f1[eq + d * neq] + f2[eq + d * neq]) * unitN[d];
  }
  const double r = sqrt(state1[0] * state2[0]);
  double vel[2];
  for (int i = 0; i < 2; i++) {
    vel[i] = state1[i + 1] / sqrt(state1[0]) + state2[i + 1] / sqrt(state2[0]);
    vel[i] /= sqrt(state1[0]) + sqrt(state2[0]);
  }
  const double qk = vel[0] * unitN[0] + vel[1] * unitN[1];
  const double p1 = dry_gen_pressure(g, state1), p2 = dry_gen_pressure(g, state2);
  double H = (state1[3] + p1) / sqrt(state1[0]) + (state2[3] + p2) / sqrt(state2[0]);
  H /= sqrt(state1[0]) + sqrt(state2[0]);
  const double a2 = 0.4 * (H - 0.5 * (vel[0] * vel[0] + vel[1] * vel[1]));
  const double a = sqrt(a2);
  double lamb0 = qk;
  const double lamb1 = qk + a, lamb2 = qk - a;
  if (fabs(lamb0) < 1e-4) lamb0 = 1e-4;
  const double deltaP = p2 - p1;
  const double deltaU = state2[1] / state2[0] - state1[1] / state1[0];
  const double deltaV = state2[2] / state2[0] - state1[2] / state1[0];
  const double deltaQk = deltaU * unitN[0] + deltaV * unitN[1];
  double DF1[4], DF4[4], DF5[4];
  DF1[0] = 1.;
  DF1[1] = vel[0];
  DF1[2] = vel[1];
  DF1[3] = 0.5 * (vel[0] * vel[0] + vel[1] * vel[1]);
  for (int i = 0; i < 4; i++) DF1[i] *= state2[0] - state1[0] - deltaP / a2;
  DF1[1] += r * (deltaU - unitN[0] * deltaQk);
  DF1[2] += r * (deltaV - unitN[1] * deltaQk);
  DF1[3] += r * (vel[0] * deltaU + vel[1] * deltaV - qk * deltaQk);
  for (int i = 0; i < 4; i++) DF1[i] *= fabs(lamb0);
  DF4[0] = 1.;
  DF4[1] = vel[0] + unitN[0] * a;
  DF4[2] = vel[1] + unitN[1] * a;
  DF4[3] = H + qk * a;
  for (int i = 0; i < 4; i++) DF4[i] *= fabs(lamb1) * (deltaP + r * a * deltaQk) * 0.5 / a2;
  DF5[0] = 1.;
  DF5[1] = vel[0] - unitN[0] * a;
  DF5[2] = vel[1] - unitN[1] * a;
  DF5[3] = H - qk * a;
  for (int i = 0; i < 4; i++) DF5[i] *= fabs(lamb2) * (deltaP - r * a * deltaQk) * 0.5 / a2;
  for (int i = 0; i < 4; i++) flux[i] = (meanFlux[i] - (DF1[i] + DF4[i] + DF5[i])) * 0.5 * normag;
}

// RiemannSolverTPS::Eval_LF (riemann_solver.cpp:89-114)
__host__ __device__ __forceinline__ void gen_riemann_lf(const GenPhys &g, const double *s1, const double *s2, const double *nor,
                                               double *flux) {
  const double maxE = fmax(gen_max_char_speed(g, s1), gen_max_char_speed(g, s2));
  double f1[GEN_MAXEQ * GEN_MAXDIM], f2[GEN_MAXEQ * GEN_MAXDIM];
  gen_conv_flux(g, s1, f1);
  gen_conv_flux(g, s2, f2);
  double normag = 0;
  for (int d = 0; d < g.dim; d++) normag += nor[d] * nor[d];
  normag = sqrt(normag);
  for (int eq = 0; eq < g.neq; eq++) {
    double a = 0, b = 0;
    for (int d = 0; d < g.dim; d++) {
      a += f1[eq + d * g.neq] * nor[d];
      b += f2[eq + d * g.neq] * nor[d];
    }
    flux[eq] = 0.5 * (a + b) - 0.5 * maxE * (s2[eq] - s1[eq]) * normag;
  }
}

// RiemannSolverTPS::Eval (riemann_solver.cpp:66-83): Roe when useRoe and the caller does not force Lax-Friedrichs
__host__ __device__ __forceinline__ void gen_riemann(const GenPhys &g, const double *s1, const double *s2, const double *nor,
                                                     double *flux) {
  if (g.use_roe) dry_gen_riemann_roe(g, s1, s2, nor, flux); else gen_riemann_lf(g, s1, s2, nor, flux);
}

}  // namespace tpsb
