// Plasma-model per-point physics for the generic path: PerfectMixture (multi-species perfect gas with ambipolar
// electrons and an optional electron-temperature equation), ConstantTransport, Chemistry (Arrhenius /
// Hoffert-Lien forward rates, detailed balance) and the node-wise SourceTerm.  One POD block (MixParams, the
// union of the reference's PerfectMixtureInput / constantTransportData / ChemistryInput) lives in device
// memory; every function below is a plain __host__ __device__ restatement of the reference routine it cites,
// in the reference's operation order.  Species order is the MIXTURE order of the reference: electron second
// to last, background last (src/M2ulPhyS.cpp:2979-3137, src/equation_of_state.hpp:137-146).
#pragma once
#include <cuda_runtime.h>

#include <cmath>

namespace tpsb {

constexpr int MIX_MAXSP = 8;    // gpudata::MAXSPECIES
constexpr int MIX_MAXRX = 34;   // gpudata::MAXREACTIONS
constexpr int MIX_MAXDIM = 3;
constexpr int MIX_MAXEQ = 12;

constexpr double MIX_RU = 8.3144598;            // UNIVERSALGASCONSTANT (equation_of_state.hpp:55)
constexpr double MIX_NA = 6.0221409e+23;        // AVOGADRONUMBER
constexpr double MIX_KB = MIX_RU / MIX_NA;      // BOLTZMANNCONSTANT
constexpr double MIX_QE = 1.60218e-19;          // ELECTRONCHARGE
constexpr double MIX_QE_OVER_KB = MIX_QE / MIX_KB;
constexpr double MIX_XEPS = 1.0e-30;            // TransportProperties::Xeps_

struct MixParams {
  // PerfectMixture (equation_of_state.cpp:478-574)
  int numSpecies, numActive, ambipolar, twoTemp, iElectron, iBackground;
  int dim, nvel, neq, iTh, iTe, eq_system, axisym;
  double mw[MIX_MAXSP], charge[MIX_MAXSP], formE[MIX_MAXSP], molarCV[MIX_MAXSP], molarCP[MIX_MAXSP];
  // ConstantTransport (transport_properties.cpp:303-330)
  double visc, bulk, kh, ke, diff[MIX_MAXSP], mtFreq[MIX_MAXSP];
  int trElectron;
  // Chemistry (chemistry.cpp:40-113)
  int numReactions, chElectron;
  double minTemp;
  int rxModel[MIX_MAXRX], detailed[MIX_MAXRX];
  double rxA[MIX_MAXRX], rxB[MIX_MAXRX], rxE[MIX_MAXRX], rxEnergy[MIX_MAXRX];
  double eqA[MIX_MAXRX], eqB[MIX_MAXRX], eqE[MIX_MAXRX];
  double reactS[MIX_MAXRX * MIX_MAXSP], prodS[MIX_MAXRX * MIX_MAXSP];  // [sp + r*numSpecies]
};

// The larger routines are deliberately NOT inlined (MIXBIG): one compiled body serves every kernel, so the body the
// point-wise parity test (tpsb_debug_point_eval vs the reference classes) certifies is the body the DG kernels run.
// nvcc 12.9 -O3 produced wrong code for the run-time-indexed small arrays of these routines in some inlining
// contexts (NaN species sources inside gen_source_kernel while the same call in gen_point_eval_kernel was exact).
#define MIXFN __host__ __device__ __forceinline__
#define MIXBIG __host__ __device__ __noinline__

// PerfectMixture::computeAmbipolarElectronNumberDensity (equation_of_state.cpp:607-618)
MIXFN double mix_ambipolar_ne(const MixParams &m, const double *n_sp) {
  double n_e = 0.0;
  for (int sp = 0; sp < m.numActive; sp++) n_e += m.charge[sp] * n_sp[sp];
  if (n_e < 0.0) n_e = 0.0;
  return n_e;
}

// PerfectMixture::computeBackgroundMassDensity (:620-650), electron density already known
MIXFN double mix_background_rho(const MixParams &m, double rho, const double *n_sp, double n_e) {
  double rhoB = rho;
  for (int sp = 0; sp < m.numActive; sp++) rhoB -= m.mw[sp] * n_sp[sp];
  if (m.ambipolar) rhoB -= n_e * m.mw[m.iElectron];
  return rhoB;
}

// PerfectMixture::computeNumberDensities (:947-961)
MIXFN void mix_number_densities(const MixParams &m, const double *U, double *n_sp) {
  for (int sp = 0; sp < m.numSpecies; sp++) n_sp[sp] = 0.0;
  double n_e = 0.0;
  for (int sp = 0; sp < m.numActive; sp++) n_sp[sp] = U[m.nvel + 2 + sp] / m.mw[sp];
  if (m.ambipolar) {
    n_e = mix_ambipolar_ne(m, n_sp);
    n_sp[m.iElectron] = n_e;
  }
  const double rhoB = mix_background_rho(m, U[0], n_sp, n_e);
  n_sp[m.iBackground] = rhoB / m.mw[m.iBackground];
}

// PerfectMixture::computeHeaviesHeatCapacity (:576-584)
MIXFN double mix_heavies_cv(const MixParams &m, const double *n_sp, double nB) {
  double c = 0.0;
  for (int sp = 0; sp < m.numActive; sp++) {
    if (sp == m.iElectron) continue;
    c += n_sp[sp] * m.molarCV[sp];
  }
  c += nB * m.molarCV[m.iBackground];
  return c;
}

// PerfectMixture::computeTemperaturesBase (:1141-1172)
MIXFN void mix_temperatures(const MixParams &m, const double *U, const double *n_sp, double n_e, double n_B, double &T_h,
                            double &T_e) {
  double totalHeatCapacity = mix_heavies_cv(m, n_sp, n_B);
  if (!m.twoTemp) totalHeatCapacity += n_e * m.molarCV[m.iElectron];
  double totalEnergy = U[m.iTh];
  for (int sp = 0; sp < m.numSpecies - 2; sp++) totalEnergy -= n_sp[sp] * m.formE[sp];
  T_h = 0.0;
  for (int d = 0; d < m.nvel; d++) T_h -= U[d + 1] * U[d + 1];
  T_h *= 0.5 / U[0];
  T_h += totalEnergy;
  if (m.twoTemp) T_h -= U[m.iTe];
  T_h /= totalHeatCapacity;
  if (m.twoTemp) {
    T_e = U[m.iTe] / n_e / m.molarCV[m.iElectron];
  } else {
    T_e = T_h;
  }
}

// PerfectMixture::computePressureBase (:1044-1063)
MIXFN double mix_pressure_base(const MixParams &m, const double *n_sp, double n_e, double n_B, double T_h, double T_e) {
  double n_h = 0.0;
  for (int sp = 0; sp < m.numActive; sp++) {
    if (sp == m.iElectron) continue;
    n_h += n_sp[sp];
  }
  n_h += n_B;
  double p = n_h * T_h;
  if (m.twoTemp) {
    p += n_e * T_e;
  } else {
    p += n_e * T_h;
  }
  p *= MIX_RU;
  return p;
}

// PerfectMixture::ComputePressure (:1029-1042)
MIXFN double mix_pressure(const MixParams &m, const double *U, double *Pe) {
  double n_sp[MIX_MAXSP], T_h, T_e;
  mix_number_densities(m, U, n_sp);
  mix_temperatures(m, U, n_sp, n_sp[m.iElectron], n_sp[m.iBackground], T_h, T_e);
  if (Pe) *Pe = n_sp[m.iElectron] * MIX_RU * T_e;
  return mix_pressure_base(m, n_sp, n_sp[m.iElectron], n_sp[m.iBackground], T_h, T_e);
}

// PerfectMixture::GetPrimitivesFromConservatives (:679-700)
MIXBIG void mix_prim(const MixParams &m, const double *U, double *Up) {
  double n_sp[MIX_MAXSP];
  mix_number_densities(m, U, n_sp);
  for (int sp = 0; sp < m.numActive; sp++) Up[m.nvel + 2 + sp] = n_sp[sp];
  Up[0] = U[0];
  for (int d = 0; d < m.nvel; d++) Up[d + 1] = U[d + 1] / U[0];
  double T_h, T_e;
  mix_temperatures(m, U, n_sp, n_sp[m.iElectron], n_sp[m.iBackground], T_h, T_e);
  Up[m.iTh] = T_h;
  if (m.twoTemp) Up[m.iTe] = T_e;
}

// PerfectMixture::ComputeMaxCharSpeed (:1359-1373) with ComputeSpeedOfSound(conserved) (:1424-1434) and the
// heavy-species heat ratio (:1311-1340)
MIXBIG double mix_max_char_speed(const MixParams &m, const double *U) {
  const double den = U[0];
  double den_vel2 = 0;
  for (int d = 0; d < m.nvel; d++) den_vel2 += U[d + 1] * U[d + 1];
  den_vel2 /= den;
  double n_sp[MIX_MAXSP], T_h, T_e;
  mix_number_densities(m, U, n_sp);
  mix_temperatures(m, U, n_sp, n_sp[m.iElectron], n_sp[m.iBackground], T_h, T_e);
  const double p = mix_pressure_base(m, n_sp, n_sp[m.iElectron], n_sp[m.iBackground], T_h, T_e);
  const double n_B = n_sp[m.iBackground];
  double mixtureCV = 0.0, n_h = n_B;
  for (int sp = 0; sp < m.numActive; sp++) {
    if (sp == m.iElectron) continue;
    mixtureCV += n_sp[sp] * m.molarCV[sp];
  }
  mixtureCV += n_B * m.molarCV[m.iBackground];
  for (int sp = 0; sp < m.numActive; sp++) {
    if (sp == m.iElectron) continue;
    n_h += n_sp[sp];
  }
  const double gamma = 1.0 + n_h * MIX_RU / mixtureCV;
  const double sound = sqrt(gamma * p / den);
  const double vel = sqrt(den_vel2 / den);
  return vel + sound;
}

// PerfectMixture::computeSpeciesEnthalpies (:1192-1207)
MIXFN void mix_species_enthalpies(const MixParams &m, const double *U, double *h) {
  double n_sp[MIX_MAXSP], T_h, T_e;
  mix_number_densities(m, U, n_sp);
  mix_temperatures(m, U, n_sp, n_sp[m.iElectron], n_sp[m.iBackground], T_h, T_e);
  for (int sp = 0; sp < m.numSpecies; sp++) {
    const double temp = (sp == m.iElectron) ? T_e : T_h;
    h[sp] = n_sp[sp] * (m.molarCP[sp] * temp + m.formE[sp]);
  }
}

// PerfectMixture::computeSpeciesPrimitives (:882-927)
MIXBIG void mix_species_primitives(const MixParams &m, const double *U, double *X_sp, double *Y_sp, double *n_sp) {
  for (int sp = 0; sp < m.numSpecies; sp++) X_sp[sp] = Y_sp[sp] = n_sp[sp] = 0.0;
  double n_e = 0.0, n = 0.0;
  for (int sp = 0; sp < m.numActive; sp++) {
    n_sp[sp] = U[m.nvel + 2 + sp] / m.mw[sp];
    n += n_sp[sp];
    if (m.ambipolar) n_e += m.charge[sp] * n_sp[sp];
  }
  if (m.ambipolar) {
    n_sp[m.iElectron] = n_e;
    n += n_e;
  }
  double Yb = 1.;
  for (int sp = 0; sp < m.numActive; sp++) {
    Y_sp[sp] = U[m.nvel + 2 + sp] / U[0];
    Yb -= Y_sp[sp];
  }
  if (m.ambipolar) {
    Y_sp[m.iElectron] = n_e * m.mw[m.iElectron] / U[0];
    Yb -= Y_sp[m.iElectron];
  }
  Y_sp[m.iBackground] = Yb;
  n_sp[m.iBackground] = Y_sp[m.iBackground] * U[0] / m.mw[m.iBackground];
  n += n_sp[m.iBackground];
  for (int sp = 0; sp < m.numSpecies; sp++) X_sp[sp] = n_sp[sp] / n;
}

// PerfectMixture::ComputeMoleFractionGradient (:1534-1592); gradUp[eq + d*neq]; gradX[sp + d*numSpecies]
MIXBIG void mix_mole_fraction_grad(const MixParams &m, const double *n_sp, const double *gradUp, double *gradX) {
  const int ns = m.numSpecies, neq = m.neq, dim = m.dim;
  for (int d = 0; d < dim; d++)
    for (int sp = 0; sp < ns; sp++) gradX[sp + d * ns] = 0.0;
  double totalN = 0.0;
  for (int sp = 0; sp < ns; sp++) totalN += n_sp[sp];
  double neGrad[MIX_MAXDIM], nBGrad[MIX_MAXDIM], totalNGrad[MIX_MAXDIM];
  for (int d = 0; d < dim; d++) neGrad[d] = 0.0;
  if (m.ambipolar)
    for (int sp = 0; sp < m.numActive; sp++)
      for (int d = 0; d < dim; d++) neGrad[d] += gradUp[(m.nvel + 2 + sp) + d * neq] * m.charge[sp];
  for (int d = 0; d < dim; d++) nBGrad[d] = gradUp[0 + d * neq];
  for (int sp = 0; sp < m.numActive; sp++)
    for (int d = 0; d < dim; d++) nBGrad[d] -= gradUp[(m.nvel + 2 + sp) + d * neq] * m.mw[sp];
  if (m.ambipolar)
    for (int d = 0; d < dim; d++) nBGrad[d] += -m.mw[m.iElectron] * neGrad[d];
  for (int d = 0; d < dim; d++) nBGrad[d] /= m.mw[m.iBackground];
  for (int d = 0; d < dim; d++) totalNGrad[d] = 0.0;
  for (int sp = 0; sp < m.numActive; sp++)
    for (int d = 0; d < dim; d++) totalNGrad[d] += gradUp[(m.nvel + 2 + sp) + d * neq];
  if (m.ambipolar)
    for (int d = 0; d < dim; d++) totalNGrad[d] += neGrad[d];
  for (int d = 0; d < dim; d++) totalNGrad[d] += nBGrad[d];
  for (int sp = 0; sp < m.numActive; sp++)
    for (int d = 0; d < dim; d++)
      gradX[sp + d * ns] = gradUp[(m.nvel + 2 + sp) + d * neq] / totalN - n_sp[sp] / totalN / totalN * totalNGrad[d];
  if (m.ambipolar) {
    const int sp = m.iElectron;
    for (int d = 0; d < dim; d++) gradX[sp + d * ns] = neGrad[d] / totalN - n_sp[sp] / totalN / totalN * totalNGrad[d];
  }
  const int sp = m.iBackground;
  for (int d = 0; d < dim; d++) gradX[sp + d * ns] = nBGrad[d] / totalN - n_sp[sp] / totalN / totalN * totalNGrad[d];
}

// ConstantTransport: diffusion velocities V[sp + d*numSpecies] (nvel components), shared by the flux and the
// source variants (transport_properties.cpp:332-383, 385-448) with addAmbipolarEfield (:131-150), addMixtureDrift
// for a zero external field (no-op) and correctMassDiffusionFlux (:59-71).  Returns n_sp too.
MIXBIG void mix_const_diffusion(const MixParams &m, const double *U, const double *gradUp, double *V, double *n_sp,
                               double *mobility) {
  const int ns = m.numSpecies;
  double prim[MIX_MAXEQ];
  mix_prim(m, U, prim);
  const double Te = m.twoTemp ? prim[m.neq - 1] : prim[m.nvel + 1];
  const double Th = prim[m.nvel + 1];
  double X_sp[MIX_MAXSP], Y_sp[MIX_MAXSP];
  mix_species_primitives(m, U, X_sp, Y_sp, n_sp);
  for (int v = 0; v < m.nvel; v++)
    for (int sp = 0; sp < ns; sp++) V[sp + v * ns] = 0.0;
  double gradX[MIX_MAXSP * MIX_MAXDIM];
  mix_mole_fraction_grad(m, n_sp, gradUp, gradX);
  for (int sp = 0; sp < ns; sp++)
    for (int d = 0; d < m.dim; d++) V[sp + d * ns] = -m.diff[sp] * gradX[sp + d * ns] / (X_sp[sp] + MIX_XEPS);
  for (int sp = 0; sp < ns; sp++) {
    const double temp = (sp == m.trElectron) ? Te : Th;
    mobility[sp] = MIX_QE_OVER_KB * m.charge[sp] / temp * m.diff[sp];
  }
  if (m.ambipolar) {
    double mho = 0.0;
    for (int sp = 0; sp < ns; sp++) mho += mobility[sp] * n_sp[sp] * m.charge[sp];
    double ambE[MIX_MAXDIM];
    for (int v = 0; v < m.nvel; v++) ambE[v] = 0.0;
    for (int sp = 0; sp < ns; sp++)
      for (int d = 0; d < m.nvel; d++) ambE[d] -= V[sp + d * ns] * n_sp[sp] * m.charge[sp];
    for (int d = 0; d < m.nvel; d++) ambE[d] /= (mho + MIX_XEPS);
    for (int sp = 0; sp < ns; sp++)
      for (int d = 0; d < m.nvel; d++) V[sp + d * ns] += mobility[sp] * ambE[d];
  }
  // addMixtureDrift with Efield = 0 adds exactly 0
  for (int sp = 0; sp < ns; sp++) {
    if (m.charge[sp] == 0.0) continue;
    for (int d = 0; d < m.nvel; d++) V[sp + d * ns] += mobility[sp] * 0.0;
  }
  double Vc[MIX_MAXDIM];
  for (int v = 0; v < m.nvel; v++) Vc[v] = 0.0;
  for (int sp = 0; sp < ns; sp++)
    for (int d = 0; d < m.nvel; d++) Vc[d] += Y_sp[sp] * V[sp + d * ns];
  for (int sp = 0; sp < ns; sp++)
    for (int d = 0; d < m.nvel; d++) V[sp + d * ns] -= Vc[d];
}

// Fluxes::ComputeConvectiveFluxes (fluxes.cpp:135-170)
MIXBIG void mix_conv_flux(const MixParams &m, const double *s, double *f) {
  double Pe = 0.0;
  const double pres = mix_pressure(m, s, &Pe);
  const int neq = m.neq, nvel = m.nvel;
  for (int d = 0; d < m.dim; d++) {
    f[0 + d * neq] = s[d + 1];
    for (int i = 0; i < nvel; i++) f[1 + i + d * neq] = s[i + 1] * s[d + 1] / s[0];
    f[1 + d + d * neq] += pres;
  }
  const double H = (s[1 + nvel] + pres) / s[0];
  for (int d = 0; d < m.dim; d++) f[1 + nvel + d * neq] = s[d + 1] * H;
  for (int sp = 0; sp < m.numActive; sp++)
    for (int d = 0; d < m.dim; d++) f[nvel + 2 + sp + d * neq] = s[nvel + 2 + sp] * s[1 + d] / s[0];
  if (m.twoTemp) {
    const double electronEnthalpy = (s[neq - 1] + Pe) / s[0];
    for (int d = 0; d < m.dim; d++) f[neq - 1 + d * neq] = electronEnthalpy * s[1 + d];
  }
}

// Fluxes::ComputeViscousFluxes (fluxes.cpp:178-335), no SGS / sponge; radius = x[0] (axisymmetric terms only)
MIXBIG void mix_visc_flux(const MixParams &m, const double *s, const double *gr, double radius, double *f) {
  const int neq = m.neq, dim = m.dim, nvel = m.nvel, ns = m.numSpecies;
  for (int i = 0; i < neq * dim; i++) f[i] = 0.;
  if (m.eq_system == 0) return;
  double hsp[MIX_MAXSP], V[MIX_MAXSP * MIX_MAXDIM], n_sp[MIX_MAXSP], mob[MIX_MAXSP];
  mix_species_enthalpies(m, s, hsp);
  mix_const_diffusion(m, s, gr, V, n_sp, mob);
  const double visc = m.visc;
  double bulk = m.bulk;
  bulk -= 2. / 3. * visc;
  double k = m.kh;
  const double ke = m.ke;
  if (m.twoTemp) {
    for (int d = 0; d < dim; d++) {
      const double qeFlux = ke * gr[neq - 1 + d * neq];
      f[1 + nvel + d * neq] += qeFlux;
      f[neq - 1 + d * neq] += qeFlux;
      f[neq - 1 + d * neq] -= hsp[ns - 2] * V[ns - 2 + d * ns];
    }
  } else {
    k += ke;
  }
  for (int d = 0; d < dim; d++) f[0 + d * neq] = 0.;
  double gu[3][3], st[3][3], vel[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    vel[i] = i < dim ? s[1 + i] / s[0] : 0.0;
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
  const double ur = m.axisym ? s[1] / s[0] : 0.0, ut = m.axisym ? s[3] / s[0] : 0.0;
  if (m.axisym && radius > 0) divV += ur / radius;
#pragma unroll
  for (int i = 0; i < 3; i++) st[i][i] += bulk * divV;
  double tau_tr = 0, tau_tz = 0;
  if (m.axisym) {  // fluxes.cpp:285-297
    tau_tr = gr[3 + 0 * neq];
    if (radius > 0) tau_tr -= ut / radius;
    tau_tr *= visc;
    tau_tz = visc * gr[3 + 1 * neq];
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    if (j < dim) {
      double vtmp = 0.0;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        if (i < dim) f[(1 + i) + j * neq] = st[i][j];
        vtmp += st[j][i] * vel[i];
      }
      f[(1 + nvel) + j * neq] += vtmp;
      f[(1 + nvel) + j * neq] += k * gr[(1 + nvel) + j * neq];
      for (int sp = 0; sp < ns; sp++) f[(1 + nvel) + j * neq] -= hsp[sp] * V[sp + j * ns];
    }
  }
  if (m.axisym) {  // fluxes.cpp:295-296, 320-323
    f[(1 + 2) + 0 * neq] = tau_tr;
    f[(1 + 2) + 1 * neq] = tau_tz;
    f[(1 + nvel) + 0 * neq] += ut * tau_tr;
    f[(1 + nvel) + 1 * neq] += ut * tau_tz;
  }
  for (int sp = 0; sp < m.numActive; sp++)
    for (int d = 0; d < dim; d++) f[(nvel + 2 + sp) + d * neq] = -s[nvel + 2 + sp] * V[sp + d * ns];
}

// Fluxes::ComputeBdrViscousFluxes (fluxes.cpp:344-504) as the wall conditions use it: every species' normal
// diffusion flux is prescribed 0 (primFluxIdxs[0..numSpecies) = true, wallBC.cpp:66-110), and with heat_prescribed
// (adiabatic wall) the heavy and -- two-temperature -- electron heat fluxes are prescribed 0 too.  nrm = unit normal.
MIXBIG void mix_bdr_visc_flux(const MixParams &m, const double *s, const double *gr, double radius, const double *nrm,
                              bool heat_prescribed, double *nf) {
  const int neq = m.neq, dim = m.dim, nvel = m.nvel, ns = m.numSpecies;
  for (int eq = 0; eq < neq; eq++) nf[eq] = 0.;
  if (m.eq_system == 0) return;
  const double visc = m.visc;
  double bulk = m.bulk;
  bulk -= 2. / 3. * visc;
  double k = m.kh;
  const double ke = m.ke;
  // species part of normalPrimFlux is replaced by the prescribed zeros, so the species-enthalpy terms vanish
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
  const double ur = m.axisym ? s[1] / s[0] : 0.0, ut = m.axisym ? s[3] / s[0] : 0.0;
  if (m.axisym && radius > 0) divV += ur / radius;
#pragma unroll
  for (int i = 0; i < 3; i++) st[i][i] += bulk * divV;
  double pf[3] = {0, 0, 0};
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      if (i < dim && j < dim) pf[i] += st[i][j] * nn[j];
  if (m.axisym) {
    double tau_tr = gr[3 + 0 * neq];
    if (radius > 0) tau_tr -= ut / radius;
    tau_tr *= visc;
    const double tau_tz = visc * gr[3 + 1 * neq];
    pf[2] += tau_tr * nn[0];
    pf[2] += tau_tz * nn[1];
  }
  double qe = 0.0, qh = 0.0;
  if (m.twoTemp) {
    for (int d = 0; d < dim; d++) qe -= ke * gr[(neq - 1) + d * neq] * nrm[d];
    qe += 0.0;  // speciesEnthalpies[electron] * (prescribed zero flux)
  } else {
    k += ke;
  }
  for (int d = 0; d < dim; d++) qh -= k * gr[(1 + nvel) + d * neq] * nrm[d];
  if (heat_prescribed) {
    qh = 0.0;
    qe = 0.0;
  }
  (void)ns;
  // species equations: -state * (prescribed zero diffusion flux)
  for (int sp = 0; sp < m.numActive; sp++) nf[nvel + 2 + sp] = -s[nvel + 2 + sp] * 0.0;
  for (int d = 0; d < nvel; d++) nf[d + 1] = pf[d];
  for (int d = 0; d < nvel; d++) nf[nvel + 1] += pf[d] * (s[1 + d] / s[0]);
  nf[nvel + 1] -= qh;
  if (m.twoTemp) {
    nf[nvel + 1] -= qe;
    nf[neq - 1] = -qe;
  }
}

// PerfectMixture::ComputePressureFromPrimitives (equation_of_state.cpp:988-1010)
MIXFN double mix_pressure_from_prim(const MixParams &m, const double *Up) {
  const double *n_sp = Up + m.nvel + 2;
  double n_e = m.ambipolar ? mix_ambipolar_ne(m, n_sp) : n_sp[m.iElectron];
  const double rhoB = mix_background_rho(m, Up[0], n_sp, n_e);
  const double nB = rhoB / m.mw[m.iBackground];
  const double T_h = Up[m.iTh], T_e = m.twoTemp ? Up[m.iTe] : Up[m.iTh];
  return mix_pressure_base(m, n_sp, n_e, nB, T_h, T_e);
}

// ConstantTransport::GetViscosities (transport_properties.hpp:305-309)
MIXFN void mix_viscosities(const MixParams &m, const double *, const double *, double *visc) {
  visc[0] = m.visc;
  visc[1] = m.bulk;
}

// PerfectMixture::computeStagnantStateWithTemp (equation_of_state.cpp:1596-1620)
MIXBIG void mix_stagnant_state_with_temp(const MixParams &m, const double *in, double Temp, double *out) {
  double n_sp[MIX_MAXSP];
  mix_number_densities(m, in, n_sp);
  for (int eq = 0; eq < m.neq; eq++) out[eq] = in[eq];
  for (int d = 0; d < m.nvel; d++) out[1 + d] = 0.;
  const double Ch = mix_heavies_cv(m, n_sp, n_sp[m.iBackground]);
  const double Ue = n_sp[m.iElectron] * m.molarCV[m.iElectron] * Temp;
  out[m.iTh] = Ch * Temp + Ue;
  if (m.twoTemp) out[m.iTe] = Ue;
  for (int sp = 0; sp < m.numSpecies - 2; sp++) out[m.iTh] += n_sp[sp] * m.formE[sp];
}

// PerfectMixture::modifyEnergyForPressure (equation_of_state.cpp:1698-1741); in and out may alias
MIXBIG void mix_modify_energy_for_pressure(const MixParams &m, const double *in, double *out, double p, bool modifyElectronEnergy) {
  double n_sp[MIX_MAXSP];
  mix_number_densities(m, in, n_sp);
  const double inTe = m.twoTemp ? in[m.iTe] : 0.0;
  double ke = 0.0;
  for (int d = 0; d < m.nvel; d++) ke += 0.5 * in[d + 1] * in[d + 1] / in[0];
  double vk[MIX_MAXDIM];
  for (int d = 0; d < m.nvel; d++) vk[d] = 0.5 * in[d + 1] * in[d + 1] / in[0];
  for (int eq = 0; eq < m.neq; eq++) out[eq] = in[eq];
  double Th = 0., pe = 0.0;
  if (m.twoTemp && (!modifyElectronEnergy)) {
    const double Te = inTe / (n_sp[m.iElectron] + 1.0e-30) / m.molarCV[m.iElectron];
    pe = n_sp[m.iElectron] * MIX_RU * Te;
  }
  for (int sp = 0; sp < m.numSpecies; sp++) {
    if (m.twoTemp && (!modifyElectronEnergy) && (sp == m.iElectron)) continue;
    Th += n_sp[sp];
  }
  Th = (p - pe) / (Th * MIX_RU);
  const double totalHeatCapacity = mix_heavies_cv(m, n_sp, n_sp[m.iBackground]);
  double rE = totalHeatCapacity * Th;
  double electronEnergy = 0.0;
  if (m.twoTemp) {
    electronEnergy = modifyElectronEnergy ? n_sp[m.iElectron] * m.molarCV[m.iElectron] * Th : inTe;
    out[m.iTe] = electronEnergy;
  } else {
    electronEnergy = n_sp[m.iElectron] * m.molarCV[m.iElectron] * Th;
  }
  rE += electronEnergy;
  for (int d = 0; d < m.nvel; d++) rE += vk[d];
  (void)ke;
  for (int sp = 0; sp < m.numSpecies - 2; sp++) rE += n_sp[sp] * m.formE[sp];
  out[m.iTh] = rE;
}

// Chemistry::isElectronInvolvedAt (chemistry.hpp:136-138)
MIXFN bool mix_electron_involved(const MixParams &m, int r) {
  return (m.chElectron < 0) ? false : (m.reactS[m.chElectron + r * m.numSpecies] != 0);
}

// SourceTerm::updateTerms, one node (source_term.cpp:117-250): chemistry creation rates, and for two
// temperatures the electron-energy sink of electron-impact reactions, the work u.grad(p_e) and the elastic
// electron-heavy energy exchange.  Un = conserved state of the SOLUTION grid function, upn / gr from the stage
// vector (parity trap 1); the species clamp index is the reference's hard-coded 3 + 2 + sp (parity trap 2).
MIXBIG void mix_source(const MixParams &m, double *Un, double *upn, const double *gr, double *src) {
  const int neq = m.neq, nvel = m.nvel, ns = m.numSpecies;
  for (int sp = 0; sp < m.numActive; sp++) {
    const int eq = 3 + 2 + sp;
    if (eq < neq) {
      upn[eq] = fmax(upn[eq], 0.0);
      Un[eq] = fmax(Un[eq], 0.0);
    }
  }
  double V[MIX_MAXSP * MIX_MAXDIM], nsp[MIX_MAXSP], mob[MIX_MAXSP];
  mix_const_diffusion(m, Un, gr, V, nsp, mob);  // ComputeSourceMolecularTransport: n_sp, mtFreq
  for (int eq = 0; eq < neq; eq++) src[eq] = 0.0;
  const double Th = upn[1 + nvel];
  const double Te = m.twoTemp ? upn[neq - 1] : Th;
  double progress[MIX_MAXRX];
  for (int r = 0; r < m.numReactions; r++) progress[r] = 0.0;
  if (ns > 1 && m.numReactions > 0) {
    const double Thlim = fmax(Th, m.minTemp), Telim = fmax(Te, m.minTemp);
    for (int r = 0; r < m.numReactions; r++) {
      const bool el = mix_electron_involved(m, r);
      const double temp = el ? Telim : Thlim;
      double kf;
      if (m.rxModel[r] == 0) {  // Arrhenius (reaction.cpp:41-48)
        kf = m.rxA[r] * pow(temp, m.rxB[r]) * exp(-m.rxE[r] / MIX_RU / temp);
      } else {  // Hoffert-Lien (reaction.cpp:53-61)
        const double tf = m.rxE[r] / MIX_KB / temp;
        kf = m.rxA[r] * pow(temp, m.rxB[r]) * (tf + 2.0) * exp(-tf);
      }
      double kC = 0.0;
      if (m.detailed[r]) kC = m.eqA[r] * pow(temp, m.eqB[r]) * exp(-m.eqE[r] / temp);  // chemistry.cpp:204-218
      double rate = 1.;  // mass action (chemistry.cpp:238-253)
      for (int sp = 0; sp < ns; sp++) rate *= pow(nsp[sp], m.reactS[sp + r * ns]);
      if (m.detailed[r]) {
        double rateBWD = 1.;
        for (int sp = 0; sp < ns; sp++) rateBWD *= pow(nsp[sp], m.prodS[sp + r * ns]);
        rate -= rateBWD / kC;
      }
      progress[r] = kf * rate;
    }
    for (int sp = 0; sp < m.numActive; sp++) {  // creation rates (chemistry.cpp:277-300)
      double c = 0.;
      for (int r = 0; r < m.numReactions; r++) c += progress[r] * (m.prodS[sp + r * ns] - m.reactS[sp + r * ns]);
      c *= m.mw[sp];
      src[2 + nvel + sp] += c;
    }
  }
  if (m.twoTemp) {
    for (int r = 0; r < m.numReactions; r++)
      if (mix_electron_involved(m, r)) src[neq - 1] -= m.rxEnergy[r] * progress[r];
    // PerfectMixture::computeElectronPressureGrad (equation_of_state.cpp:1847-1871)
    double neGrad[MIX_MAXDIM];
    for (int d = 0; d < m.dim; d++) neGrad[d] = 0.0;
    if (m.ambipolar) {
      for (int sp = 0; sp < m.numActive; sp++)
        for (int d = 0; d < m.dim; d++) neGrad[d] += gr[(nvel + 2 + sp) + d * neq] * m.charge[sp];
    } else {
      for (int d = 0; d < m.dim; d++) neGrad[d] = gr[(nvel + ns) + d * neq];
    }
    const double ne = nsp[ns - 2];
    for (int d = 0; d < m.dim; d++) {
      const double gradPe = (neGrad[d] * Te + ne * gr[m.iTe + d * neq]) * MIX_RU;
      src[neq - 1] += gradPe * upn[d + 1];
    }
    const double me = m.mw[ns - 2];
    for (int sp = 0; sp < ns; sp++) {
      if (sp == ns - 2) continue;
      const double m_sp = m.mw[sp];
      double energy = 1.5 * MIX_RU * (Te - Th);
      energy *= 2.0 * me * m_sp / (m_sp + me) / (m_sp + me) * ne * m.mtFreq[sp];
      src[neq - 1] -= energy;
    }
  }
}

}  // namespace tpsb
