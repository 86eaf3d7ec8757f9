// Per-point physics with RUN-TIME dimension / equation count for the generic tensor-product path
// (rhs_generic.cuh): 2-D and 3-D, nvel = dim (or 3 when axisymmetric), num_equation <= GEN_MAXEQ.
// Dry air only so far; the mixture models plug in behind the same five entry points.
// Formulas and operation order follow the reference routines cited at each function.
#pragma once
#include <cuda_runtime.h>

#include "mix_physics.cuh"
#include "physics.cuh"

namespace tpsb {

constexpr int GEN_MAXEQ = 12;  // gpudata::MAXEQUATIONS of the reference's CUDA build (src/dataStructures.hpp:52)
constexpr int GEN_MAXDIM = 3;

struct GenPhys {
  int dim, nvel, neq;
  int fluid;             // 0 dry air; 1 user-defined plasma mixture (mix != NULL)
  PhysParams dry;        // gamma, R, Sutherland, multipliers, eq_system
  const MixParams *mix;  // device (or host, for host-side checks) pointer
};

// DryAir::ComputePressure (equation_of_state.hpp:610-617)
__host__ __device__ __forceinline__ double dry_gen_pressure(const GenPhys &g, const double *s) {
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
  const double T = g.dry.gm1 / g.dry.R * (s[1 + g.nvel] - 0.5 * den_vel2) / s[0];
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

// Fluxes::ComputeViscousFluxes (fluxes.cpp:178-335) with DryAirTransport (transport_properties.cpp:223-234);
// non-axisymmetric, no SGS, no sponge.  gr[eq + d*neq] = d Up_eq / d x_d.
// Written with fixed 3x3 register tiles and guards instead of run-time-indexed local arrays (entries beyond
// `dim` are zero, so sums keep the reference's operation order): nvcc 12.9 -O3 produced wrong stores for the
// run-time-indexed form once inlined into gen_resid_kernel (caught by the parity test; the stand-alone
// function was correct, tools/ubench/gen_visc_check.cu).
__host__ __device__ __forceinline__ void dry_gen_visc_flux(const GenPhys &g, const double *s, const double *gr, double *f) {
  const int neq = g.neq, dim = g.dim;
  for (int i = 0; i < neq * dim; i++) f[i] = 0.;
  if (g.dry.eq_system == 0) return;
  const double pr = dry_gen_pressure(g, s);
  const double temp = pr / g.dry.R / s[0];
  const double visc = (g.dry.C1 * g.dry.visc_mult * (temp * sqrt(temp)) / (temp + g.dry.S0));
  double bulk = g.dry.bulk_visc_mult * visc;
  const double k = g.dry.cp_div_pr * visc;
  bulk -= 2. / 3. * visc;
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
__host__ __device__ __forceinline__ void gen_visc_flux(const GenPhys &g, const double *s, const double *gr, double *f) {
  if (g.fluid) mix_visc_flux(*g.mix, s, gr, f); else dry_gen_visc_flux(g, s, gr, f);
}
__host__ __device__ __forceinline__ int gen_num_active_species(const GenPhys &g) { return g.fluid ? g.mix->numActive : 0; }

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

}  // namespace tpsb
