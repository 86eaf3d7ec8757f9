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

#include "physics.cuh"  // PhysParams / DryAux / dry_modify_transport: the SGS and sponge block is fluid independent

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
  // transport model: the reference's TransportModel value (dataStructures.hpp:80-88): 2 CONSTANT, 0 ARGON_MINIMAL
  int transportModel;
  // ConstantTransport (transport_properties.cpp:303-330)
  double visc, bulk, kh, ke, diff[MIX_MAXSP], mtFreq[MIX_MAXSP];
  int trElectron;
  // GasMinimalTransport, argon ternary (gas_transport.cpp:42-157): species indices, per-particle masses mw / N_A,
  // reduced masses, third-order electron conductivity switch, artificial multipliers
  int gmIon, gmElectron, gmNeutral, thirdOrderKe, multiply;
  double gmMw[MIX_MAXSP], gmMuw[MIX_MAXSP * MIX_MAXSP], fluxMult[4], mfFreqMult, diffMult, mobilMult;
  // GasMixtureTransport (transportModel 1, gas_transport.cpp:877-1650): collision type per species pair,
  // GasColl values (dataStructures.hpp:122-143) at [spI + spJ*numSpecies], spI <= spJ
  int collIdx[MIX_MAXSP * MIX_MAXSP];
  // Chemistry (chemistry.cpp:40-113)
  int numReactions, chElectron;
  double minTemp;
  int rxModel[MIX_MAXRX], detailed[MIX_MAXRX];
  double rxA[MIX_MAXRX], rxB[MIX_MAXRX], rxE[MIX_MAXRX], rxEnergy[MIX_MAXRX];
  double eqA[MIX_MAXRX], eqB[MIX_MAXRX], eqE[MIX_MAXRX];
  double reactS[MIX_MAXRX * MIX_MAXSP], prodS[MIX_MAXRX * MIX_MAXSP];  // [sp + r*numSpecies]
  // TABULATED_RXN (rxModel 2): LinearTable per reaction (table.cpp:76-97) in one device pool, reaction r at
  // tbl + tblOff[r]: x[n] | a[n-1] | b[n-1]
  const double *tbl;
  int tblOff[MIX_MAXRX + 1], tblN[MIX_MAXRX + 1], tblXlog[MIX_MAXRX + 1], tblFlog[MIX_MAXRX + 1];
  int radiation;  // NetEmission (radiation.hpp:57-70): the net-emission-coefficient table sits in slot MIX_MAXRX
  // GRIDFUNCTION_RXN (rxModel 3): externally supplied rate coefficient per node, component rxComp[r] of
  // rateField[comp][N] (reaction.cpp:86-117); NULL -> 0 like the reference
  const double *rateField;
  long long rateN;
  int rxComp[MIX_MAXRX];
  // flow/useMixingLength: MixingLengthTransport around the molecular transport (mixing_length_transport.cpp:44-57)
  int mlOn;
  double mlMax, mlPrt, mlBulk;
};


// The larger routines are deliberately NOT inlined (MIXBIG): one compiled body serves every kernel, so the body the
// point-wise parity test (tpsb_debug_point_eval vs the reference classes) certifies is the body the DG kernels run.
// nvcc 12.9 -O3 produced wrong code for the run-time-indexed small arrays of these routines in some inlining
// contexts (NaN species sources inside gen_source_kernel while the same call in gen_point_eval_kernel was exact).
#define MIXFN __host__ __device__ __forceinline__
#define MIXBIG __host__ __device__ __noinline__

// MixingLengthTransport::ComputeFluxTransportProperties (mixing_length_transport.cpp:62-121): eddy viscosity
// rho l^2 |S| with l = min(0.41 d_wall, l_max), |S| = sqrt(2 S_ij S_ij) including the axisymmetric swirl terms, added to
// the molecular viscosity; bulk += bulk_mult mu_t; kappa += mu_t (kappa / mu) Pr_ratio.  bulk is the raw buffer entry
// (before Fluxes subtracts 2/3 mu).  gr[eq + d*neq]; velocities are momentum / rho in every gas model.
MIXFN void mixlen_add(int dim, int nvel, int neq, double max_len, double prt, double bulk_mult, const double *s, const double *gr,
                      double radius, double distance, double &visc, double &bulk, double &kappa) {
  const double cp_over_Pr = kappa / visc;
  const double rho = s[0];
  double ur = 0;
  if (nvel != dim) ur = s[1] / s[0];
  double S = 0;
  for (int i = 0; i < dim; i++)
    for (int j = 0; j < dim; j++) {
      const double Sij = 0.5 * (gr[(1 + i) + j * neq] + gr[(1 + j) + i * neq]);
      S += 2 * Sij * Sij;
    }
  if (nvel != dim) {
    const double ut = s[3] / s[0];
    const double ut_r = gr[3 + 0 * neq], ut_z = gr[3 + 1 * neq];
    double Szx = 0.5 * ut_r;
    if (radius > 0) Szx -= 0.5 * ut / radius;
    const double Szy = 0.5 * ut_z;
    double Szz = 0.0;
    if (radius > 0) Szz += ur / radius;
    S += 2 * (2 * Szx * Szx + 2 * Szy * Szy + Szz * Szz);
  }
  S = sqrt(S);
  double mixing_length = 0.41 * distance;
  if (mixing_length > max_len) mixing_length = max_len;
  const double mut = rho * mixing_length * mixing_length * S;
  visc += mut;
  bulk += bulk_mult * mut;
  kappa += mut * cp_over_Pr * prt;
}

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

// ---- collision integrals (collision_integrals.cpp:53-201): charged-particle fits in units of pi lambda_D^2 over the
// Debye-length-scaled temperature, argon fits in m^2 over T [K] ----
MIXFN double coll_charged(double a, double b, double c, double d, double Tp) {
  return a * pow(log(1.0 + b * pow(Tp, c)), d) / Tp / Tp;
}
MIXFN double coll_att11(double Tp) { return coll_charged(0.2150, 5.2194, 1.0472, 1.2435, Tp); }
MIXFN double coll_att12(double Tp) { return coll_charged(0.0991, 7.4684, 1.0155, 1.1536, Tp); }
MIXFN double coll_att13(double Tp) { return coll_charged(0.0616, 7.8271, 0.9452, 1.1105, Tp); }
MIXFN double coll_att14(double Tp) { return coll_charged(0.0308, 13.9567, 0.9511, 1.1803, Tp); }
MIXFN double coll_att15(double Tp) { return coll_charged(0.0232, 13.7888, 0.9148, 1.1532, Tp); }
MIXFN double coll_att22(double Tp) { return coll_charged(0.2423, 4.6796, 1.3290, 1.1279, Tp); }
MIXFN double coll_att23(double Tp) { return coll_charged(0.1221, 8.7542, 1.3875, 1.1110, Tp); }
MIXFN double coll_att24(double Tp) { return coll_charged(0.0619, 18.2538, 1.4341, 1.1618, Tp); }
MIXFN double coll_rep11(double Tp) { return coll_charged(0.3904, 0.9100, 1.1025, 1.0544, Tp); }
MIXFN double coll_rep12(double Tp) { return coll_charged(0.1547, 1.6597, 1.1725, 0.9792, Tp); }
MIXFN double coll_rep13(double Tp) { return coll_charged(0.0814, 2.5815, 1.1948, 0.9570, Tp); }
MIXFN double coll_rep14(double Tp) { return coll_charged(0.0683, 1.9774, 1.2033, 0.8264, Tp); }
MIXFN double coll_rep15(double Tp) { return coll_charged(0.0346, 4.5177, 1.2132, 0.9294, Tp); }
MIXFN double coll_ArAr11(double T) { return 2.2910e-18 * pow(T, -0.3032); }
MIXFN double coll_rep22(double Tp) { return coll_charged(0.4128, 1.2436, 1.1830, 1.0123, Tp); }
MIXFN double coll_rep23(double Tp) { return coll_charged(0.2203, 1.8832, 1.2059, 0.9851, Tp); }
MIXFN double coll_rep24(double Tp) { return coll_charged(0.1323, 2.7248, 1.2129, 0.9847, Tp); }
MIXFN double coll_ArAr22(double T) { return 1.7e-18 * pow(T, -0.25); }
MIXFN double coll_ArAr1P11(double T) { return 4.574321e-18 * pow(T, -0.1805); }
// e-Ar (1,r): 9-term polynomial in log T, powers -1 .. 7 (collision_integrals.cpp:124-140)
__host__ __device__ __noinline__ double coll_eAr(int r, double T) {
  const double coeff[5][9] = {
      {6.36254140e-18, 1.84835040e-18, -5.87727093e-18, 3.20023027e-18, -8.50509054e-19, 1.28163820e-19, -1.11712910e-20,
       5.25649382e-22, -1.03296658e-23},
      {1.91338172e-17, 5.45418129e-18, -1.78361685e-17, 9.75657946e-18, -2.61115722e-18, 3.98310268e-19, -3.53503678e-20,
       1.70375066e-21, -3.45211955e-23},
      {3.04685398e-17, 8.39750994e-18, -2.88132528e-17, 1.60147037e-17, -4.34837891e-18, 6.73136845e-19, -6.06704580e-20,
       2.97216168e-21, -6.12760944e-23},
      {3.90777949e-17, 1.04696956e-17, -3.73774204e-17, 2.10610498e-17, -5.79029566e-18, 9.07573157e-19, -8.28466766e-20,
       4.11188110e-21, -8.59225098e-23},
      {4.41333290e-17, 1.15696010e-17, -4.25651305e-17, 2.42442440e-17, -6.73359258e-18, 1.06641697e-18, -9.83933863e-20,
       4.93775812e-21, -1.04362372e-22}};
  const double logT = log(T);
  double pw[9];
  pw[0] = 1. / logT;
  pw[1] = 1.;
  for (int k = 0; k < 7; k++) pw[k + 2] = pw[k + 1] * logT;
  double fit = 0.0;
  for (int k = 0; k < 9; k++) fit += coeff[r - 1][k] * pw[k];
  return fit;
}
// Nitrogen pairs (collision_integrals.cpp:210-625): every fit is Q(T) = pre exp(s_sum sum_k (s_coef c_k) (ln T)^k) with
// 2 to 7 terms; the scalings are applied where the reference applies them (to the coefficients for e-N / e-N2, to the
// sum for N2-N.+1) and the powers are formed with pow() term by term, so the round-off of the strongly cancelling sums
// (|terms| ~ 1e4 against a result ~ -45) follows the reference's.  Data table generated by tools/gen_nitrogen_fits.py.
struct NitFit {
  int n;
  double s_coef, s_sum;
  int pi;
  double c[7];
};
enum { NIT_NiNi11 = 0, NIT_NiNi22, NIT_NiNi1P11, NIT_N2N211, NIT_N2N222, NIT_N2N21P11, NIT_N2Ni1P11, NIT_NiN21P11, NIT_N2Ni11,
       NIT_N2Ni22, NIT_eNi11, NIT_eNi12, NIT_eNi13, NIT_eNi14, NIT_eNi15, NIT_eN211, NIT_eN212, NIT_eN213, NIT_eN214, NIT_eN215 };
__host__ __device__ __noinline__ double coll_nitrogen(int which, double T) {
  const NitFit fits[20] = {
#include "nitrogen_fits.inc"
  };
  const NitFit &f = fits[which];
  const double logT = log(T);
  double sum = f.s_coef == 1.0 ? f.c[0] : f.c[0] * f.s_coef;
  for (int k = 1; k < f.n; k++) {
    const double ck = f.s_coef == 1.0 ? f.c[k] : f.c[k] * f.s_coef;
    sum += ck * (k == 1 ? logT : pow(logT, static_cast<double>(k)));
  }
  if (f.s_sum != 1.0) sum *= f.s_sum;
  const double q = exp(sum);
  return f.pi ? 3.14159265358979323846 * q : q;
}

constexpr double MIX_PI = 3.14159265358979323846;   // equation_of_state.hpp:67
constexpr double MIX_EPS0 = 8.8541878128e-12;       // VACUUMPERMITTIVITY
constexpr double MIX_DEBYE = MIX_KB * MIX_EPS0 / MIX_QE / MIX_QE;  // gas_transport.hpp:74

// Debye-length scales shared by the argon-minimal routines (gas_transport.cpp:224-231)
struct GmDebye {
  double length, circle, nondimTe, nondimTh;
};
MIXFN GmDebye gm_debye(const MixParams &m, const double *n_sp, double Te, double Th) {
  GmDebye d;
  const double nOverT = (n_sp[m.gmElectron] + MIX_XEPS) / Te + (n_sp[m.gmIon] + MIX_XEPS) / Th;
  d.length = sqrt(MIX_DEBYE / MIX_NA / nOverT);
  d.circle = MIX_PI * d.length * d.length;
  d.nondimTe = d.length * 4.0 * MIX_PI * MIX_DEBYE * Te;
  d.nondimTh = d.length * 4.0 * MIX_PI * MIX_DEBYE * Th;
  return d;
}
// binary diffusivities D_ij = binaryDiff[i + 3 j] of the ternary argon mixture (gas_transport.cpp:275-335)
MIXFN void gm_binary_diff(const MixParams &m, double Te, double Th, double nTotal, const GmDebye &db, double *bd) {
  const double diffusivityFactor = 3. / 16. * sqrt(2.0 * MIX_PI * MIX_KB) / MIX_NA;
  const int e = m.gmElectron, n = m.gmNeutral, i = m.gmIon;
  for (int k = 0; k < 9; k++) bd[k] = 0.0;
  bd[e + n * 3] = diffusivityFactor * sqrt(Te / m.gmMuw[e + n * 3]) / nTotal / coll_eAr(1, Te);
  bd[n + e * 3] = bd[e + n * 3];
  bd[n + i * 3] = diffusivityFactor * sqrt(Th / m.gmMuw[n + i * 3]) / nTotal / coll_ArAr1P11(Th);
  bd[i + n * 3] = bd[n + i * 3];
  bd[e + i * 3] = diffusivityFactor * sqrt(Te / m.gmMuw[i + e * 3]) / nTotal / (coll_att11(db.nondimTe) * db.circle);
  bd[i + e * 3] = bd[e + i * 3];
}
// TransportProperties::CurtissHirschfelder (transport_properties.cpp:188-199) + mobilities + multipliers
MIXFN void gm_diffusivity_mobility(const MixParams &m, const double *X_sp, const double *Y_sp, const double *bd, double Te,
                                   double Th, double *diffusivity, double *mobility) {
  for (int sp = 0; sp < 3; sp++) diffusivity[sp] = 0.0;
  for (int spI = 0; spI < 3; spI++) {
    for (int spJ = 0; spJ < 3; spJ++) {
      if (spI == spJ) continue;
      diffusivity[spI] += (X_sp[spJ] + MIX_XEPS) / bd[spI + spJ * 3];
    }
    diffusivity[spI] = (1.0 - Y_sp[spI]) / diffusivity[spI];
  }
  for (int sp = 0; sp < 3; sp++) {
    const double temp = (sp == m.gmElectron) ? Te : Th;
    mobility[sp] = MIX_QE_OVER_KB * m.charge[sp] / temp * diffusivity[sp];
  }
  if (m.multiply)
    for (int sp = 0; sp < 3; sp++) {
      diffusivity[sp] *= m.diffMult;
      mobility[sp] *= m.mobilMult;
    }
}
// diffusion velocities from diffusivities / mobilities: concentration-driven part, ambipolar field
// (transport_properties.cpp:131-150), zero external field, mass-flux correction (:59-71)
MIXFN void mix_diffusion_velocity(const MixParams &m, const double *X_sp, const double *Y_sp, const double *n_sp,
                                  const double *gradUp, const double *diffusivity, const double *mobility, double *V) {
  const int ns = m.numSpecies;
  double gradX[MIX_MAXSP * MIX_MAXDIM];
  mix_mole_fraction_grad(m, n_sp, gradUp, gradX);
  for (int v = 0; v < m.nvel; v++)
    for (int sp = 0; sp < ns; sp++) V[sp + v * ns] = 0.0;
  for (int sp = 0; sp < ns; sp++)
    for (int d = 0; d < m.dim; d++) {
      const double DgradX = diffusivity[sp] * gradX[sp + d * ns];
      V[sp + d * ns] = -DgradX / (X_sp[sp] + MIX_XEPS);
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
// species viscosities -> mixture viscosity (gas_transport.cpp:233-262; also GetViscosities :775-823)
MIXFN double gm_viscosity(const MixParams &m, const double *X_sp, double Th, const GmDebye &db, double *speciesViscosity) {
  const double viscosityFactor = 5. / 16. * sqrt(MIX_PI * MIX_KB);
  speciesViscosity[m.gmIon] = viscosityFactor * sqrt(m.gmMw[m.gmIon] * Th) / (coll_rep22(db.nondimTh) * db.circle);
  speciesViscosity[m.gmNeutral] = viscosityFactor * sqrt(m.gmMw[m.gmNeutral] * Th) / coll_ArAr22(Th);
  speciesViscosity[m.gmElectron] = 0.0;
  double avg = 0.0;
  for (int sp = 0; sp < 3; sp++) avg += X_sp[sp] * speciesViscosity[sp];
  return avg;
}
// GasMinimalTransport::computeThirdOrderElectronThermalConductivity (gas_transport.cpp:400-489), argon
MIXBIG double gm_third_order_ke(const MixParams &m, const double *X_sp, const GmDebye &db, double Te) {
  const double viscosityFactor = 5. / 16. * sqrt(MIX_PI * MIX_KB), kOverEtaFactor = 15. / 4. * MIX_KB;
  double Q2[3], Q1Ion[5], Q1N[5];
  Q2[0] = db.circle * coll_rep22(db.nondimTe);
  Q2[1] = db.circle * coll_rep23(db.nondimTe);
  Q2[2] = db.circle * coll_rep24(db.nondimTe);
  Q1Ion[0] = db.circle * coll_att11(db.nondimTe);
  Q1Ion[1] = db.circle * coll_att12(db.nondimTe);
  Q1Ion[2] = db.circle * coll_att13(db.nondimTe);
  Q1Ion[3] = db.circle * coll_att14(db.nondimTe);
  Q1Ion[4] = db.circle * coll_att15(db.nondimTe);
  for (int r = 0; r < 5; r++) Q1N[r] = coll_eAr(r + 1, Te);
  auto L11ea = [](const double *Q1) { return 6.25 * Q1[0] - 15. * Q1[1] + 12. * Q1[2]; };
  auto L12ea = [](const double *Q1) { return 10.9375 * Q1[0] - 39.375 * Q1[1] + 57. * Q1[2] - 30. * Q1[3]; };
  auto L22ea = [](const double *Q1) { return 19.140625 * Q1[0] - 91.875 * Q1[1] + 199.5 * Q1[2] - 210. * Q1[3] + 90. * Q1[4]; };
  const int e = m.gmElectron, i = m.gmIon, n = m.gmNeutral;
  double L11 = sqrt(2.0) * X_sp[e] * Q2[0];
  L11 += X_sp[i] * L11ea(Q1Ion);
  L11 += X_sp[n] * L11ea(Q1N);
  double L12 = sqrt(2.0) * X_sp[e] * (1.75 * Q2[0] - 2.0 * Q2[1]);
  L12 += X_sp[i] * L12ea(Q1Ion);
  L12 += X_sp[n] * L12ea(Q1N);
  double L22 = sqrt(2.0) * X_sp[e] * (4.8125 * Q2[0] - 7.0 * Q2[1] + 5. * Q2[2]);
  L22 += X_sp[i] * L22ea(Q1Ion);
  L22 += X_sp[n] * L22ea(Q1N);
  return viscosityFactor * kOverEtaFactor * sqrt(2.0 * Te / m.gmMw[e]) * X_sp[e] / (L11 - L12 * L12 / L22);
}

// ---- GasMixtureTransport (argon mixtures of up to 7 species: Ar, Ar.+1, excited states, E) ----
// GasMinimalTransport::computeCollisionInputs (gas_transport.cpp:185-204): Debye length from ALL charged species at T_e
struct GmxColl {
  double Te, Th, debyeCircle, ndimTe, ndimTh;
};
MIXFN GmxColl gmx_collision_inputs(const MixParams &m, double Te, double Th, const double *n_sp) {
  GmxColl c;
  c.Te = Te, c.Th = Th;
  double nOverT = 0.0;
  for (int sp = 0; sp < m.numSpecies; sp++) nOverT += (n_sp[sp] + MIX_XEPS) / Te * m.charge[sp] * m.charge[sp];
  const double debyeLength = sqrt(MIX_DEBYE / MIX_NA / nOverT);
  c.debyeCircle = MIX_PI * debyeLength * debyeLength;
  c.ndimTe = debyeLength * 4.0 * MIX_PI * MIX_DEBYE * Te;
  c.ndimTh = debyeLength * 4.0 * MIX_PI * MIX_DEBYE * Th;
  return c;
}
// GasMixtureTransport::collisionIntegral (gas_transport.cpp:995-1283), charged, argon and nitrogen pairs.  Unsupported (l, r)
// return NaN (the reference asserts).
MIXBIG double gmx_collision_integral(const MixParams &m, int _spI, int _spJ, int l, int r, const GmxColl &ci) {
  const int spI = (_spI > _spJ) ? _spJ : _spI, spJ = (_spI > _spJ) ? _spI : _spJ;
  const int e = m.gmElectron;
  const int collIdx = m.collIdx[spI + spJ * m.numSpecies];
  double temp;
  if (collIdx == 0 || collIdx == 1) {
    temp = ((spI == e) || (spJ == e)) ? ci.ndimTe : ci.ndimTh;
  } else {
    temp = ((spI == e) || (spJ == e)) ? ci.Te : ci.Th;
  }
  const double nan = NAN;
  switch (collIdx) {
    case 0:  // CLMB_ATT
      if (l == 1) {
        switch (r) {
          case 1: return ci.debyeCircle * coll_att11(temp);
          case 2: return ci.debyeCircle * coll_att12(temp);
          case 3: return ci.debyeCircle * coll_att13(temp);
          case 4: return ci.debyeCircle * coll_att14(temp);
          case 5: return ci.debyeCircle * coll_att15(temp);
        }
      } else if (l == 2) {
        switch (r) {
          case 2: return ci.debyeCircle * coll_att22(temp);
          case 3: return ci.debyeCircle * coll_att23(temp);
          case 4: return ci.debyeCircle * coll_att24(temp);
        }
      }
      return nan;
    case 1:  // CLMB_REP
      if (l == 1) {
        switch (r) {
          case 1: return ci.debyeCircle * coll_rep11(temp);
          case 2: return ci.debyeCircle * coll_rep12(temp);
          case 3: return ci.debyeCircle * coll_rep13(temp);
          case 4: return ci.debyeCircle * coll_rep14(temp);
          case 5: return ci.debyeCircle * coll_rep15(temp);
        }
      } else if (l == 2) {
        switch (r) {
          case 2: return ci.debyeCircle * coll_rep22(temp);
          case 3: return ci.debyeCircle * coll_rep23(temp);
          case 4: return ci.debyeCircle * coll_rep24(temp);
        }
      }
      return nan;
    case 2:  // AR_AR1P
      return (l == 1 && r == 1) ? coll_ArAr1P11(temp) : nan;
    case 3:  // AR_E
      return (l == 1 && r >= 1 && r <= 5) ? coll_eAr(r, temp) : nan;
    case 4:  // AR_AR
      return (l == 1 && r == 1) ? coll_ArAr11(temp) : ((l == 2 && r == 2) ? coll_ArAr22(temp) : nan);
    // nitrogen ("Ni") pairs, GasColl values of src/dataStructures.hpp:133-143 (gas_transport.cpp:1160-1277)
    case 6:  // NI_NI1P
      return (l == 1 && r == 1) ? coll_nitrogen(NIT_NiNi1P11, temp) : nan;
    case 10:  // N2_NI1P
      return (l == 1 && r == 1) ? coll_nitrogen(NIT_N2Ni1P11, temp) : nan;
    case 11:  // N2_E
      return (l == 1 && r >= 1 && r <= 5) ? coll_nitrogen(NIT_eN211 + r - 1, temp) : nan;
    case 7:  // NI_E
      return (l == 1 && r >= 1 && r <= 5) ? coll_nitrogen(NIT_eNi11 + r - 1, temp) : nan;
    case 8:  // NI_NI
      return (l == 1 && r == 1) ? coll_nitrogen(NIT_NiNi11, temp) : ((l == 2 && r == 2) ? coll_nitrogen(NIT_NiNi22, temp) : nan);
    case 12:  // N2_N2
      return (l == 1 && r == 1) ? coll_nitrogen(NIT_N2N211, temp) : ((l == 2 && r == 2) ? coll_nitrogen(NIT_N2N222, temp) : nan);
    case 13:  // N2_NI
      return (l == 1 && r == 1) ? coll_nitrogen(NIT_N2Ni11, temp) : ((l == 2 && r == 2) ? coll_nitrogen(NIT_N2Ni22, temp) : nan);
  }
  return -1.0;  // NI_N21P / N2_N21P have no case in the reference either: it returns -1
}
// binary diffusivities, Curtiss-Hirschfelder, mobilities, multipliers (gas_transport.cpp:1323-1349)
MIXBIG void gmx_diffusivity_mobility(const MixParams &m, const double *X_sp, const double *Y_sp, double nTotal, const GmxColl &ci,
                                     double *diffusivity, double *mobility) {
  const int ns = m.numSpecies, e = m.gmElectron;
  const double diffusivityFactor = 3. / 16. * sqrt(2.0 * MIX_PI * MIX_KB) / MIX_NA;
  double bd[MIX_MAXSP * MIX_MAXSP];
  for (int spI = 0; spI < ns - 1; spI++)
    for (int spJ = spI + 1; spJ < ns; spJ++) {
      const double temp = ((spI == e) || (spJ == e)) ? ci.Te : ci.Th;
      bd[spI + spJ * ns] = diffusivityFactor * sqrt(temp / m.gmMuw[spI + spJ * ns]) / nTotal / gmx_collision_integral(m, spI, spJ, 1, 1, ci);
      bd[spJ + spI * ns] = bd[spI + spJ * ns];
    }
  for (int sp = 0; sp < ns; sp++) diffusivity[sp] = 0.0;
  for (int spI = 0; spI < ns; spI++) {
    for (int spJ = 0; spJ < ns; spJ++) {
      if (spI == spJ) continue;
      diffusivity[spI] += (X_sp[spJ] + MIX_XEPS) / bd[spI + spJ * ns];
    }
    diffusivity[spI] = (1.0 - Y_sp[spI]) / diffusivity[spI];
  }
  for (int sp = 0; sp < ns; sp++) {
    const double temp = (sp == e) ? ci.Te : ci.Th;
    mobility[sp] = MIX_QE_OVER_KB * m.charge[sp] / temp * diffusivity[sp];
  }
  if (m.multiply)
    for (int sp = 0; sp < ns; sp++) {
      diffusivity[sp] *= m.diffMult;
      mobility[sp] *= m.mobilMult;
    }
}
// species viscosities -> mixture viscosity (gas_transport.cpp:1300-1312, 1508-1520)
MIXFN double gmx_viscosity(const MixParams &m, const double *X_sp, const GmxColl &ci, double *speciesViscosity) {
  const double viscosityFactor = 5. / 16. * sqrt(MIX_PI * MIX_KB);
  double avg = 0.0;
  for (int sp = 0; sp < m.numSpecies; sp++) {
    speciesViscosity[sp] = (sp == m.gmElectron) ? 0.0 : viscosityFactor * sqrt(m.gmMw[sp] * ci.Th) / gmx_collision_integral(m, sp, sp, 2, 2, ci);
  }
  for (int sp = 0; sp < m.numSpecies; sp++) avg += X_sp[sp] * speciesViscosity[sp];
  return avg;
}
// GasMixtureTransport::computeThirdOrderElectronThermalConductivity (gas_transport.cpp:1388-1407)
MIXBIG double gmx_third_order_ke(const MixParams &m, const double *X_sp, const GmxColl &ci) {
  const double viscosityFactor = 5. / 16. * sqrt(MIX_PI * MIX_KB), kOverEtaFactor = 15. / 4. * MIX_KB;
  const int e = m.gmElectron;
  double Q2[3];
  for (int r = 0; r < 3; r++) Q2[r] = gmx_collision_integral(m, e, e, 2, r + 2, ci);
  double L11 = sqrt(2.0) * X_sp[e] * Q2[0];
  double L12 = sqrt(2.0) * X_sp[e] * (1.75 * Q2[0] - 2.0 * Q2[1]);
  double L22 = sqrt(2.0) * X_sp[e] * (4.8125 * Q2[0] - 7.0 * Q2[1] + 5. * Q2[2]);
  for (int sp = 0; sp < m.numSpecies; sp++) {
    if (sp == e) continue;
    double Q1[5];
    for (int r = 0; r < 5; r++) Q1[r] = gmx_collision_integral(m, sp, e, 1, r + 1, ci);
    L11 += X_sp[sp] * (6.25 * Q1[0] - 15. * Q1[1] + 12. * Q1[2]);
    L12 += X_sp[sp] * (10.9375 * Q1[0] - 39.375 * Q1[1] + 57. * Q1[2] - 30. * Q1[3]);
    L22 += X_sp[sp] * (19.140625 * Q1[0] - 91.875 * Q1[1] + 199.5 * Q1[2] - 210. * Q1[3] + 90. * Q1[4]);
  }
  return viscosityFactor * kOverEtaFactor * sqrt(2.0 * ci.Te / m.gmMw[e]) * X_sp[e] / (L11 - L12 * L12 / L22);
}

// TransportProperties::ComputeFluxTransportProperties -> {Constant, GasMinimal}Transport::ComputeFluxMolecularTransport
// (transport_properties.cpp:332-383, gas_transport.cpp:206-398): tb = {viscosity, bulk viscosity, heavy thermal
// conductivity, electron thermal conductivity}, diffusion velocities V[sp + d*numSpecies].
MIXBIG void mix_flux_transport(const MixParams &m, const double *s, const double *gr, double *tb, double *V) {
  double n_sp[MIX_MAXSP], mob[MIX_MAXSP];
  if (m.transportModel == 2) {
    tb[0] = m.visc, tb[1] = m.bulk, tb[2] = m.kh, tb[3] = m.ke;
    mix_const_diffusion(m, s, gr, V, n_sp, mob);
    return;
  }
  double prim[MIX_MAXEQ], X_sp[MIX_MAXSP], Y_sp[MIX_MAXSP];
  mix_prim(m, s, prim);
  mix_species_primitives(m, s, X_sp, Y_sp, n_sp);
  double nTotal = 0.0;
  for (int sp = 0; sp < m.numSpecies; sp++) nTotal += n_sp[sp];
  const double Te = m.twoTemp ? prim[m.neq - 1] : prim[m.nvel + 1];
  const double Th = prim[m.nvel + 1];
  if (m.transportModel == 1) {  // GasMixtureTransport::ComputeFluxMolecularTransport (gas_transport.cpp:1285-1386)
    const GmxColl ci = gmx_collision_inputs(m, Te, Th, n_sp);
    const double vF = 5. / 16. * sqrt(MIX_PI * MIX_KB), kF = 15. / 4. * MIX_KB;
    double spVisc[MIX_MAXSP], diffusivity[MIX_MAXSP];
    tb[0] = gmx_viscosity(m, X_sp, ci, spVisc);
    tb[2] = 0.0;
    for (int sp = 0; sp < m.numSpecies; sp++) tb[2] += X_sp[sp] * (sp == m.gmElectron ? 0.0 : spVisc[sp] * kF / m.gmMw[sp]);
    tb[1] = 0.0;
    if (m.thirdOrderKe) {
      tb[3] = gmx_third_order_ke(m, X_sp, ci);
    } else {
      tb[3] = vF * kF * sqrt(ci.Te / m.gmMw[m.gmElectron]) * X_sp[m.gmElectron] /
              gmx_collision_integral(m, m.gmElectron, m.gmElectron, 2, 2, ci);
    }
    gmx_diffusivity_mobility(m, X_sp, Y_sp, nTotal, ci, diffusivity, mob);
    if (m.multiply)
      for (int t = 0; t < 4; t++) tb[t] *= m.fluxMult[t];
    mix_diffusion_velocity(m, X_sp, Y_sp, n_sp, gr, diffusivity, mob, V);
    return;
  }
  const GmDebye db = gm_debye(m, n_sp, Te, Th);
  const double viscosityFactor = 5. / 16. * sqrt(MIX_PI * MIX_KB), kOverEtaFactor = 15. / 4. * MIX_KB;
  double spVisc[3], spK[3];
  tb[0] = gm_viscosity(m, X_sp, Th, db, spVisc);
  for (int sp = 0; sp < 3; sp++) spK[sp] = spVisc[sp] * kOverEtaFactor / m.gmMw[sp];
  tb[2] = 0.0;
  for (int sp = 0; sp < 3; sp++) tb[2] += X_sp[sp] * spK[sp];
  tb[1] = 0.0;
  if (m.thirdOrderKe) {
    tb[3] = gm_third_order_ke(m, X_sp, db, Te);
  } else {
    tb[3] = viscosityFactor * kOverEtaFactor * sqrt(Te / m.gmMw[m.gmElectron]) * X_sp[m.gmElectron] /
            (coll_rep22(db.nondimTe) * db.circle);
  }
  double bd[9], diffusivity[3];
  gm_binary_diff(m, Te, Th, nTotal, db, bd);
  gm_diffusivity_mobility(m, X_sp, Y_sp, bd, Te, Th, diffusivity, mob);
  if (m.multiply)
    for (int t = 0; t < 4; t++) tb[t] *= m.fluxMult[t];
  mix_diffusion_velocity(m, X_sp, Y_sp, n_sp, gr, diffusivity, mob, V);
}

// ComputeSourceMolecularTransport (transport_properties.cpp:385-448, gas_transport.cpp:592-773): number densities
// and the electron momentum-transfer collision frequencies the two-temperature source needs
MIXBIG void mix_source_transport(const MixParams &m, const double *Un, const double *upn, const double *gr, double *n_sp,
                                 double *mtFreq) {
  if (m.transportModel == 2) {
    double V[MIX_MAXSP * MIX_MAXDIM], mob[MIX_MAXSP];
    mix_const_diffusion(m, Un, gr, V, n_sp, mob);
    for (int sp = 0; sp < m.numSpecies; sp++) mtFreq[sp] = m.mtFreq[sp];
    return;
  }
  double X_sp[MIX_MAXSP], Y_sp[MIX_MAXSP];
  mix_species_primitives(m, Un, X_sp, Y_sp, n_sp);
  const double Te = m.twoTemp ? upn[m.neq - 1] : upn[m.nvel + 1];
  const double Th = upn[m.nvel + 1];
  const double mfFreqFactor = 4. / 3. * MIX_NA * sqrt(8. * MIX_KB / MIX_PI);
  if (m.transportModel == 1) {  // GasMixtureTransport::ComputeSourceMolecularTransport (gas_transport.cpp:1409-1497)
    const GmxColl ci = gmx_collision_inputs(m, Te, Th, n_sp);
    for (int sp = 0; sp < m.numSpecies; sp++) {
      mtFreq[sp] = (sp == m.gmElectron) ? 0.0
                                        : mfFreqFactor * sqrt(ci.Te / m.gmMw[m.gmElectron]) * n_sp[sp] *
                                              gmx_collision_integral(m, sp, m.gmElectron, 1, 1, ci);
      if (m.multiply) mtFreq[sp] *= m.mfFreqMult;
    }
    return;
  }
  const GmDebye db = gm_debye(m, n_sp, Te, Th);
  const double Qea = coll_eAr(1, Te), Qie = coll_att11(db.nondimTe) * db.circle;
  for (int sp = 0; sp < 3; sp++) mtFreq[sp] = 0.0;
  mtFreq[m.gmIon] = mfFreqFactor * sqrt(Te / m.gmMw[m.gmElectron]) * n_sp[m.gmIon] * Qie;
  mtFreq[m.gmNeutral] = mfFreqFactor * sqrt(Te / m.gmMw[m.gmElectron]) * n_sp[m.gmNeutral] * Qea;
  if (m.multiply)
    for (int sp = 0; sp < 3; sp++) mtFreq[sp] *= m.mfFreqMult;
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
// mod / ax != NULL: the SGS model / viscous sponge block (fluxes.cpp:224-246) with the caller's element size and point
MIXBIG void mix_visc_flux(const MixParams &m, const double *s, const double *gr, double radius, double *f, double distance = 0.0,
                          const PhysParams *mod = nullptr, const DryAux *ax = nullptr) {
  const int neq = m.neq, dim = m.dim, nvel = m.nvel, ns = m.numSpecies;
  for (int i = 0; i < neq * dim; i++) f[i] = 0.;
  if (m.eq_system == 0) return;
  double hsp[MIX_MAXSP], V[MIX_MAXSP * MIX_MAXDIM], tb[4];
  mix_species_enthalpies(m, s, hsp);
  mix_flux_transport(m, s, gr, tb, V);
  if (m.mlOn) mixlen_add(dim, nvel, neq, m.mlMax, m.mlPrt, m.mlBulk, s, gr, radius, distance, tb[0], tb[1], tb[2]);
  double visc = tb[0];
  double bulk = tb[1];
  bulk -= 2. / 3. * visc;
  double k = tb[2];
  const double ke = tb[3];
  if (mod && ax && (mod->sgs_model | mod->sponge)) {
    dry_modify_transport(*mod, s[0], gr + 1, neq, *ax, visc, bulk, k);
    if (mod->sponge) {  // the sponge also scales the active species' diffusion velocities (fluxes.cpp:241-245)
      const double wgt = dry_sponge_weight(*mod, ax->x);
      for (int sp = 0; sp < m.numActive; sp++)
        for (int d = 0; d < dim; d++) V[sp + d * ns] *= wgt;
    }
  }
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
// General form: pf[0..numSpecies) are the prescribed species diffusion velocities (NULL: zeros), hvy_prescribed /
// elec_prescribed select the prescribed heat fluxes pf[numSpecies + nvel] / pf[numSpecies + nvel + 1].
MIXBIG void mix_bdr_visc_flux_general(const MixParams &m, const double *s, const double *gr, double radius, const double *nrm,
                                      const double *pf, bool hvy_prescribed, bool elec_prescribed, double *nf,
                                      const PhysParams *mod = nullptr, const DryAux *ax = nullptr);
MIXBIG void mix_bdr_visc_flux(const MixParams &m, const double *s, const double *gr, double radius, const double *nrm,
                              bool heat_prescribed, double *nf, const PhysParams *mod = nullptr, const DryAux *ax = nullptr) {
  mix_bdr_visc_flux_general(m, s, gr, radius, nrm, nullptr, heat_prescribed, heat_prescribed, nf, mod, ax);
}
MIXBIG void mix_bdr_visc_flux_general(const MixParams &m, const double *s, const double *gr, double radius, const double *nrm,
                                      const double *pfl, bool hvy_prescribed, bool elec_prescribed, double *nf,
                                      const PhysParams *mod, const DryAux *ax) {
  const int neq = m.neq, dim = m.dim, nvel = m.nvel, ns = m.numSpecies;
  for (int eq = 0; eq < neq; eq++) nf[eq] = 0.;
  if (m.eq_system == 0) return;
  double tb[4], Vunused[MIX_MAXSP * MIX_MAXDIM];
  mix_flux_transport(m, s, gr, tb, Vunused);
  double visc = tb[0];
  double bulk = tb[1];
  bulk -= 2. / 3. * visc;
  double k = tb[2];
  const double ke = tb[3];
  if (mod && ax && (mod->sgs_model | mod->sponge)) dry_modify_transport(*mod, s[0], gr + 1, neq, *ax, visc, bulk, k);  // fluxes.cpp:386-407
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
  double vsp[MIX_MAXSP], hsp[MIX_MAXSP];  // normalPrimFlux[sp]: every wall type prescribes all of them
  for (int sp = 0; sp < ns; sp++) vsp[sp] = pfl ? pfl[sp] : 0.0;
  mix_species_enthalpies(m, s, hsp);
  double qe = 0.0, qh = 0.0;
  if (m.twoTemp) {
    for (int d = 0; d < dim; d++) qe -= ke * gr[(neq - 1) + d * neq] * nrm[d];
    qe += hsp[ns - 2] * vsp[ns - 2];
  } else {
    k += ke;
  }
  for (int d = 0; d < dim; d++) qh -= k * gr[(1 + nvel) + d * neq] * nrm[d];
  for (int sp = 0; sp < ns; sp++) {
    if (m.twoTemp && (sp == ns - 2)) continue;
    qh += hsp[sp] * vsp[sp];
  }
  if (hvy_prescribed) qh = pfl ? pfl[ns + nvel] : 0.0;
  if (elec_prescribed) qe = pfl ? pfl[ns + nvel + 1] : 0.0;
  for (int sp = 0; sp < m.numActive; sp++) nf[nvel + 2 + sp] = -s[nvel + 2 + sp] * vsp[sp];
  for (int d = 0; d < nvel; d++) nf[d + 1] = pf[d];
  for (int d = 0; d < nvel; d++) nf[nvel + 1] += pf[d] * (s[1 + d] / s[0]);
  nf[nvel + 1] -= qh;
  if (m.twoTemp) {
    nf[nvel + 1] -= qe;
    nf[neq - 1] = -qe;
  }
}

// PerfectMixture::GetConservativesFromPrimitives (equation_of_state.cpp:744-783)
MIXBIG void mix_cons(const MixParams &m, const double *primit, double *conserv) {
  const int nvel = m.nvel;
  conserv[0] = primit[0];
  for (int d = 0; d < nvel; d++) conserv[d + 1] = primit[d + 1] * primit[0];
  for (int sp = 0; sp < m.numActive; sp++) conserv[nvel + 2 + sp] = primit[nvel + 2 + sp] * m.mw[sp];
  const double *n_sp = primit + nvel + 2;
  const double n_e = m.ambipolar ? mix_ambipolar_ne(m, n_sp) : n_sp[m.iElectron];
  const double rhoB = mix_background_rho(m, primit[0], n_sp, n_e);
  const double nB = rhoB / m.mw[m.iBackground];
  if (m.twoTemp) conserv[m.iTe] = n_e * m.molarCV[m.iElectron] * primit[m.iTe];
  double totalHeatCapacity = mix_heavies_cv(m, n_sp, nB);
  if (!m.twoTemp) totalHeatCapacity += n_e * m.molarCV[m.iElectron];
  double totalEnergy = 0.0;
  for (int d = 0; d < nvel; d++) totalEnergy += primit[d + 1] * primit[d + 1];
  totalEnergy *= 0.5 * primit[0];
  totalEnergy += totalHeatCapacity * primit[m.iTh];
  if (m.twoTemp) totalEnergy += conserv[m.iTe];
  for (int sp = 0; sp < m.numSpecies - 2; sp++) totalEnergy += primit[nvel + 2 + sp] * m.formE[sp];
  conserv[m.iTh] = totalEnergy;
}

// PerfectMixture::computeSheathBdrFlux (equation_of_state.cpp:1909-1942): Bohm velocities of the positive ions, the
// electron flux that keeps the wall current-free, a fully catalytic wall for the background, and (two temperatures)
// the sheath electron heat flux.  pf[numSpecies + nvel + 2].
MIXBIG void mix_sheath_bdr_flux(const MixParams &m, const double *state, double *pf) {
  const int ns = m.numSpecies;
  double n_sp[MIX_MAXSP], T_h, T_e;
  mix_number_densities(m, state, n_sp);
  mix_temperatures(m, state, n_sp, n_sp[m.iElectron], n_sp[m.iBackground], T_h, T_e);
  for (int sp = 0; sp < ns; sp++) pf[sp] = 0.0;
  for (int sp = 0; sp < ns; sp++) {
    const double Zsp = m.charge[sp];
    if (Zsp > 0.0) {
      const double msp = m.mw[sp];
      const double VB = sqrt((T_h + Zsp * T_e) * MIX_RU / msp);
      pf[sp] = VB;
      pf[m.iElectron] += Zsp * n_sp[sp] * VB;
      pf[m.iBackground] -= msp * n_sp[sp] * VB;
    }
  }
  pf[m.iElectron] /= n_sp[m.iElectron];
  pf[m.iBackground] -= m.mw[m.iElectron] * n_sp[m.iElectron] * pf[m.iElectron];
  pf[m.iBackground] /= m.mw[m.iBackground] * n_sp[m.iBackground];
  if (m.twoTemp) {
    const double vTe = sqrt(8.0 * MIX_RU * T_e / MIX_PI / m.mw[m.iElectron]);
    const double gamma = -log(4.0 / vTe * pf[m.iElectron]);
    pf[ns + m.nvel + 1] = pf[m.iElectron] * (gamma + 2.0) * n_sp[m.iElectron] * MIX_RU * T_e;
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

// ConstantTransport::GetViscosities (transport_properties.hpp:305-309) / GasMinimalTransport::GetViscosities
// (gas_transport.cpp:775-823)
MIXBIG void mix_viscosities(const MixParams &m, const double *U, const double *Up, double *visc) {
  if (m.transportModel == 2) {
    visc[0] = m.visc;
    visc[1] = m.bulk;
    return;
  }
  double n_sp[MIX_MAXSP], X_sp[MIX_MAXSP], Y_sp[MIX_MAXSP], spVisc[MIX_MAXSP];
  mix_species_primitives(m, U, X_sp, Y_sp, n_sp);
  const double Te = m.twoTemp ? Up[m.neq - 1] : Up[m.nvel + 1];
  const double Th = Up[m.nvel + 1];
  if (m.transportModel == 1) {  // GasMixtureTransport::GetViscosities (gas_transport.cpp:1499-1546)
    const GmxColl ci = gmx_collision_inputs(m, Te, Th, n_sp);
    visc[0] = gmx_viscosity(m, X_sp, ci, spVisc);
  } else {
    const GmDebye db = gm_debye(m, n_sp, Te, Th);
    visc[0] = gm_viscosity(m, X_sp, Th, db, spVisc);
  }
  visc[1] = 0.0;
  if (m.multiply) {
    visc[0] *= m.fluxMult[0];
    visc[1] *= m.fluxMult[1];
  }
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

// TableInterpolator::findInterval + LinearTable::eval (table.cpp:52-97)
__host__ __device__ __noinline__ double mix_table_eval(const MixParams &m, int r, double xEval) {
  const int n = m.tblN[r];
  const double *x = m.tbl + m.tblOff[r], *ta = x + n, *tb = ta + n;
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
  first = max(1, min(n - 1, first));
  const int index = first - 1;
  const double xt = m.tblXlog[r] ? log(xEval) : xEval;
  double ft = ta[index] + tb[index] * xt;
  if (m.tblFlog[r]) ft = exp(ft);
  return ft;
}

// Chemistry::isElectronInvolvedAt (chemistry.hpp:136-138)
MIXFN bool mix_electron_involved(const MixParams &m, int r) {
  return (m.chElectron < 0) ? false : (m.reactS[m.chElectron + r * m.numSpecies] != 0);
}

// SourceTerm::updateTerms, one node (source_term.cpp:117-250): chemistry creation rates, and for two
// temperatures the electron-energy sink of electron-impact reactions, the work u.grad(p_e) and the elastic
// electron-heavy energy exchange.  Un = conserved state of the SOLUTION grid function, upn / gr from the stage
// vector (parity trap 1); the species clamp index is the reference's hard-coded 3 + 2 + sp (parity trap 2).
MIXBIG void mix_source(const MixParams &m, double *Un, double *upn, const double *gr, long long node, double *src) {
  const int neq = m.neq, nvel = m.nvel, ns = m.numSpecies;
  for (int sp = 0; sp < m.numActive; sp++) {
    const int eq = 3 + 2 + sp;
    if (eq < neq) {
      upn[eq] = fmax(upn[eq], 0.0);
      Un[eq] = fmax(Un[eq], 0.0);
    }
  }
  double nsp[MIX_MAXSP], mtFreq[MIX_MAXSP];
  mix_source_transport(m, Un, upn, gr, nsp, mtFreq);  // ComputeSourceMolecularTransport: n_sp, mtFreq
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
      } else if (m.rxModel[r] == 1) {  // Hoffert-Lien (reaction.cpp:53-61)
        const double tf = m.rxE[r] / MIX_KB / temp;
        kf = m.rxA[r] * pow(temp, m.rxB[r]) * (tf + 2.0) * exp(-tf);
      } else if (m.rxModel[r] == 2) {  // Tabulated (reaction.cpp:78-84)
        kf = mix_table_eval(m, r, temp);
      } else {  // GridFunctionReaction (reaction.cpp:108-117)
        kf = m.rateField ? m.rateField[node + m.rxComp[r] * m.rateN] : 0.;
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
  if (m.radiation) src[1 + nvel] += -4.0 * MIX_PI * mix_table_eval(m, MIX_MAXRX, Th);  // source_term.cpp:205-207
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
      energy *= 2.0 * me * m_sp / (m_sp + me) / (m_sp + me) * ne * mtFreq[sp];
      src[neq - 1] -= energy;
    }
  }
}

}  // namespace tpsb
