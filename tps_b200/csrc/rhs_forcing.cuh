// Forcing terms of RHSoperator::Mult (ForcingTerms subclasses, src/forcing_terms.cpp): nodal terms added to dU/dt after
// Me^-1 (src/rhs_operator.cpp:451-461).  One node-wise kernel serves every path: it rebuilds the node's coordinates from
// its element's vertices and its primitives from the stage vector, so no node lists are stored (the reference's
// constructors build them once by looping over all nodes: HeatSource :941-970, SpongeZone :549-610).
#pragma once
#include "rhs_generic.cuh"

namespace tpsb {

constexpr int MAX_FORCING = 8;
struct ForcingDev {
  int kind, sz_type, sz_mixed;
  double g[3];                        // pressure gradient
  double p1[3], axis[3], len, radius, value;  // heat source: segment start, unit axis, length
  const double *field;                // Joule heating
  double n[3], p0[3], pi[3], r1, r2, tol, mult;  // sponge zone (n normalised)
  double targetU[GEN_MAXEQ], sound;   // user-defined target (conserved) and its speed of sound
  double *mix;                        // mixed-out scratch [neq + 1 sums | neq target | 1 sound]
};
struct ForcingArgs {
  int dim, nvel, neq, dof, nv, NE, nf;
  long long N;
  GenPhys phys;
  const double *vx;    // [NE][nv][dim]
  const double *xiN;   // [dof][dim]
  const double *x;     // stage vector: Up = prim(x)
  const double *gradUp;
  double *y;
  ForcingDev f[MAX_FORCING];
};

__device__ __forceinline__ double sponge_sigma(const ForcingDev &f, int dim, const double *X, double *radial) {
  // SpongeZone::SpongeZone (src/forcing_terms.cpp:566-607): distInit < 0 ahead of the entry plane
  double distInit = 0., distF = 0.;
  for (int d = 0; d < dim; d++) distInit -= f.n[d] * (X[d] - f.pi[d]);
  for (int d = 0; d < dim; d++) distF += f.n[d] * (X[d] - f.p0[d]);
  if (f.sz_type == 0) {
    if (distInit > 0. && distF > 0.) {
      const double planeDistance = distF + distInit;
      return distInit / planeDistance / planeDistance;
    }
    return 0.0;
  }
  double R = 0., tmp[3] = {0, 0, 0};
  for (int d = 0; d < dim; d++) tmp[d] = X[d] - f.pi[d] + distInit * f.n[d];
  for (int d = 0; d < dim; d++) R += tmp[d] * tmp[d];
  R = sqrt(R);
  if (distInit > 0. && distF > 0. && R - f.r1 > 0.) {
    const double planeDistance = f.r2 - f.r1;
    for (int d = 0; d < dim; d++) radial[d] = tmp[d] / R;
    return (R - f.r1) / planeDistance / planeDistance;
  }
  return 0.0;
}

// SpongeZone::computeMixedOutValues, first half (src/forcing_terms.cpp:706-733): sums of the normal convective fluxes over the
// nodes within tol of the entry plane (annulus: of the radius r1) and their count
__global__ void forcing_mixed_out_sum_kernel(ForcingArgs a, int which) {
  const long long n = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  const ForcingDev &f = a.f[which];
  const long long e = n / a.dof;
  const int k = static_cast<int>(n % a.dof);
  double X[3];
  gen_point(a.dim, a.vx + e * a.nv * a.dim, a.xiN + k * a.dim, X);
  double distInit = 0.;
  for (int d = 0; d < a.dim; d++) distInit -= f.n[d] * (X[d] - f.pi[d]);
  bool in;
  if (f.sz_type == 0) {
    in = fabs(distInit) < f.tol;
  } else {
    double R = 0.;
    for (int d = 0; d < a.dim; d++) {
      const double t = X[d] - f.pi[d] + distInit * f.n[d];
      R += t * t;
    }
    in = fabs(sqrt(R) - f.r1) < f.tol;
  }
  if (!in) return;
  double s[GEN_MAXEQ], up[GEN_MAXEQ], un[GEN_MAXEQ], fl[GEN_MAXEQ * GEN_MAXDIM];
  for (int eq = 0; eq < a.neq; eq++) s[eq] = a.x[n + eq * a.N];
  gen_prim(a.phys, s, up);       // the reference goes Up -> conserved -> flux
  gen_cons(a.phys, up, un);
  gen_conv_flux(a.phys, un, fl);
  for (int eq = 0; eq < a.neq; eq++) {
    double v = 0;
    for (int d = 0; d < a.dim; d++) v += f.n[d] * fl[eq + d * a.neq];
    atomicAdd(&f.mix[eq], v);
  }
  atomicAdd(&f.mix[a.neq], 1.0);
}
// ... second half (:735-742) + DryAir::computeConservedStateFromConvectiveFlux (src/equation_of_state.cpp:414-442)
__global__ void forcing_mixed_out_target_kernel(ForcingArgs a, int which) {
  const ForcingDev &f = a.f[which];
  const int neq = a.neq, nvel = a.nvel, dim = a.dim;
  double m[GEN_MAXEQ];
  const double cnt = f.mix[neq];
  for (int eq = 0; eq < neq; eq++) m[eq] = f.mix[eq] / cnt;
  const double gamma = a.phys.dry.gamma;
  double temp = 0.;
  for (int d = 0; d < dim; d++) temp += m[1 + d] * f.n[d];
  const double A = 1. - 2. * gamma / (gamma - 1.), B = 2 * temp / (gamma - 1.);
  double C = -2. * m[0] * m[1 + nvel];
  for (int d = 0; d < nvel; d++) C += m[1 + d] * m[1 + d];
  const double p = (-B - sqrt(B * B - 4. * A * C)) / (2. * A);
  double up[GEN_MAXEQ], U[GEN_MAXEQ];
  up[0] = m[0] * m[0] / (temp - p);
  up[1 + nvel] = p / (a.phys.dry.R * up[0]);
  for (int d = 0; d < nvel; d++) up[1 + d] = d < dim ? (m[1 + d] - p * f.n[d]) / m[0] : m[1 + d] / m[0];
  gen_cons(a.phys, up, U);
  for (int eq = 0; eq < neq; eq++) f.mix[neq + 1 + eq] = U[eq];
  f.mix[2 * neq + 1] = sqrt(gamma * a.phys.dry.R * up[1 + nvel]);  // ComputeSpeedOfSound(Up, true) of the target
}

__global__ void forcing_kernel(ForcingArgs a) {
  const long long n = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  const long long e = n / a.dof;
  const int k = static_cast<int>(n % a.dof);
  const int neq = a.neq, nvel = a.nvel, dim = a.dim;
  double X[3];
  gen_point(dim, a.vx + e * a.nv * dim, a.xiN + k * dim, X);
  double s[GEN_MAXEQ], up[GEN_MAXEQ];
  bool have_prim = false;
  auto prim = [&]() {
    if (have_prim) return;
    for (int eq = 0; eq < neq; eq++) s[eq] = a.x[n + eq * a.N];
    gen_prim(a.phys, s, up);
    have_prim = true;
  };
  for (int i = 0; i < a.nf; i++) {
    const ForcingDev &f = a.f[i];
    if (f.kind == 0) {  // ConstantPressureGradient::updateTerms, CPU branch (src/forcing_terms.cpp:147-172)
      prim();
      const double p = gen_pressure_from_prim(a.phys, up);
      double grad_pV = 0.;
      for (int d = 0; d < dim; d++) {
        a.y[n + (d + 1) * a.N] -= f.g[d];
        grad_pV -= up[d + 1] * f.g[d];
        grad_pV -= p * a.gradUp[n + (d + 1) * a.N + d * a.N * neq];
      }
      a.y[n + (1 + nvel) * a.N] += grad_pV;
    } else if (f.kind == 1) {  // HeatSource (:941-1010), "cylinder"
      double proj = 0, r2 = 0;
      for (int d = 0; d < dim; d++) proj += (X[d] - f.p1[d]) * f.axis[d];
      for (int d = 0; d < dim; d++) {
        const double r = (X[d] - f.p1[d]) - proj * f.axis[d];
        r2 += r * r;
      }
      if (sqrt(r2) < f.radius && proj > 0 && proj < f.len) a.y[n + (dim + 1) * a.N] += f.value;
    } else if (f.kind == 2) {  // JouleHeating::updateTerms (:443-470)
      const double heating = f.field[n];
      if (heating > 0.) {
        a.y[n + (nvel + 1) * a.N] += heating;
        if (a.phys.fluid && a.phys.mix->twoTemp) a.y[n + (neq - 1) * a.N] += heating;
      }
    } else {  // SpongeZone::addSpongeZoneForcing (:631-700)
      double ur[3] = {0, 0, 0};
      double sg = sponge_sigma(f, dim, X, ur);
      if (sg > 0.) {
        sg *= f.mult;
        prim();
        double Un[GEN_MAXEQ], tgt[GEN_MAXEQ], tcyl[GEN_MAXEQ];
        gen_cons(a.phys, up, Un);
        const double *T = f.sz_mixed ? f.mix + neq + 1 : f.targetU;
        const double sound = f.sz_mixed ? f.mix[2 * neq + 1] : f.sound;
        for (int eq = 0; eq < neq; eq++) tgt[eq] = tcyl[eq] = T[eq];
        if (f.sz_type == 1 && dim == 3) {  // target momentum from (radial, azimuthal, axial) to Cartesian components (:676-695)
          double uth[3], M[9], inv[9];
          uth[0] = f.n[1] * ur[2] - ur[1] * f.n[2];
          uth[1] = f.n[2] * ur[0] - f.n[0] * ur[2];
          uth[2] = f.n[0] * ur[1] - ur[0] * f.n[1];
          for (int d = 0; d < 3; d++) M[0 + 3 * d] = ur[d], M[1 + 3 * d] = uth[d], M[2 + 3 * d] = f.n[d];
          const double det = det3(M);
          adj3(M, inv);
          for (int r = 0; r < 3; r++) {
            double v = 0;
            for (int c = 0; c < 3; c++) v += inv[r + 3 * c] / det * tgt[1 + c];
            tcyl[1 + r] = v;
          }
        }
        for (int eq = 0; eq < neq; eq++) a.y[n + eq * a.N] -= sound * sg * (Un[eq] - tcyl[eq]);
      }
    }
  }
}

// Averaging::addSampleInternal (src/averaging.cpp:250-328 plain, :330-420 with a GasMixture): one thread per node
__global__ void averaging_kernel(long long N, int nf, int dim, const double *inst, double *mean, double *vari, int vstart, int vcomp,
                                 double ns_mean, double ns_vari, int pressure_slot, GenPhys phys) {
  const long long n = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double nUp[GEN_MAXEQ + 4], mstate[GEN_MAXEQ + 4];
  for (int eq = 0; eq < nf; eq++) nUp[eq] = inst[n + eq * N];
  for (int eq = 0; eq < nf; eq++) {
    const double N_umean = ns_mean * mean[n + eq * N];
    double v = nUp[eq];
    if (pressure_slot && eq == 1 + dim) v = gen_pressure_from_prim(phys, nUp);
    mstate[eq] = (N_umean + v) / (ns_mean + 1);
    mean[n + eq * N] = mstate[eq];
  }
  if (!vari) return;
  int vi = 0;
  for (int i = vstart; i < vstart + vcomp; i++) {
    const double di = nUp[i] - mstate[i];
    vari[n + vi * N] = (vari[n + vi * N] * ns_vari + di * di) / (ns_vari + 1);
    vi++;
  }
  for (int i = vstart; i < vstart + vcomp - 1; i++) {
    const double di = nUp[i] - mstate[i];
    for (int j = i + 1; j < vstart + vcomp; j++) {
      const double dj = nUp[j] - mstate[j];
      vari[n + vi * N] = (vari[n + vi * N] * ns_vari + di * dj) / (ns_vari + 1);
      vi++;
    }
  }
}

}  // namespace tpsb
