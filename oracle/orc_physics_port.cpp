// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Never linked into the product.
//
// "port" physics back end: a line-by-line CPU restatement of the reference's dry-air
// closure.  Each function cites the reference lines it follows.  It is validated against
// the reference's own object code (orc_physics_ref.cpp) by tests/test_oracle_physics.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "orc_physics.hpp"

namespace orc {

class DryAirPort : public Physics {
  OrcPhysParams p_;
  int dim_, nvel_, neq_;
  double cp_div_pr_;

 public:
  DryAirPort(const OrcPhysParams &p, int dim, int nvel, int neq) : p_(p), dim_(dim), nvel_(nvel), neq_(neq) {
    // src/transport_properties.cpp:220
    cp_div_pr_ = p.gamma * p.R / (p.Pr * (p.gamma - 1.));
  }
  const char *kind() const override { return "port"; }
  int num_active_species() const override { return 0; }
  int num_species() const override { return 1; }  // src/equation_of_state.cpp:154 (not NS_PASSIVE)
  // src/equation_of_state.cpp:365-377
  void stagnation_state(const double *U, double *out) override {
    const double p = pressure(U);
    for (int eq = 0; eq < neq_; eq++) out[eq] = U[eq];
    for (int d = 0; d < nvel_; d++) out[1 + d] = 0.;
    out[1 + nvel_] = p / (p_.gamma - 1.);
  }
  // src/equation_of_state.cpp:379-386
  void stagnant_state_with_temp(const double *U, double T, double *out) override {
    for (int eq = 0; eq < neq_; eq++) out[eq] = U[eq];
    for (int d = 0; d < nvel_; d++) out[1 + d] = 0.;
    out[1 + nvel_] = p_.R / (p_.gamma - 1.) * U[0] * T;
  }
  // src/equation_of_state.cpp:389-411
  void modify_energy_for_pressure(const double *in, double *out, double p, bool) override {
    double tmp[16];
    for (int eq = 0; eq < neq_; eq++) tmp[eq] = in[eq];
    double ke = 0.;
    for (int d = 0; d < nvel_; d++) ke += in[1 + d] * in[1 + d];
    ke *= 0.5 / in[0];
    for (int eq = 0; eq < neq_; eq++) out[eq] = tmp[eq];
    out[1 + nvel_] = p / (p_.gamma - 1.) + ke;
  }
  // src/equation_of_state.cpp:134-143
  void modify_state_from_primitive(const double *U, const double *bprim, const bool *primIdxs, double *out) override {
    double pr[16];
    prim(U, pr);
    for (int i = 0; i < neq_; i++)
      if (primIdxs[i]) pr[i] = bprim[i];
    cons(pr, out);
  }
  // src/fluxes.cpp:344-504 for dry air: one species with zero diffusion velocity
  // (src/transport_properties.cpp:236-266), no species enthalpy carried to the heat flux, single temperature.
  void bdr_visc_flux(const double *s, const double *gradUp, double *xyz, double /*delta*/, double /*dist*/,
                     const double *nrm, const double *primFlux, const bool *primFluxIdxs, double *normalFlux) override {
    for (int eq = 0; eq < neq_; eq++) normalFlux[eq] = 0.;
    if (p_.eq_system == 0) return;
    const bool axisym = (dim_ == 2 && nvel_ == 3);
    const double radius = axisym ? xyz[0] : -1;
    const int numSpecies = 1;
    double pr = pressure(s);
    double temp = pr / p_.R / s[0];
    double visc = (p_.C1 * p_.visc_mult * pow(temp, 1.5) / (temp + p_.S0));
    double bulkViscosity = p_.bulk_visc_mult * visc;
    double k = cp_div_pr_ * visc;
    bulkViscosity -= 2. / 3. * visc;
    const int primFluxSize = numSpecies + nvel_ + 1;
    double normalPrimFlux[16];
    for (int eq = 0; eq < primFluxSize; eq++) normalPrimFlux[eq] = 0.0;
    for (int i = 0; i < numSpecies; i++)
      if (primFluxIdxs[i]) normalPrimFlux[i] = primFlux[i];
    double stress[9];
    double divV = 0.;
    for (int i = 0; i < dim_; i++) {
      for (int j = 0; j < dim_; j++)
        stress[i + j * dim_] = gradUp[(1 + j) + i * neq_] + gradUp[(1 + i) + j * neq_];
      divV += gradUp[(1 + i) + i * neq_];
    }
    for (int i = 0; i < dim_; i++)
      for (int j = 0; j < dim_; j++) stress[i + j * dim_] *= visc;
    const double ur = (axisym ? s[1] / s[0] : 0), ut = (axisym ? s[3] / s[0] : 0);
    if (axisym && radius > 0) divV += ur / radius;
    for (int i = 0; i < dim_; i++) stress[i + i * dim_] += bulkViscosity * divV;
    for (int i = 0; i < dim_; i++)
      for (int j = 0; j < dim_; j++) normalPrimFlux[numSpecies + i] += stress[i + j * dim_] * nrm[j];
    if (axisym) {  // src/fluxes.cpp:452-464
      const double ut_r = gradUp[3 + 0 * neq_], ut_z = gradUp[3 + 1 * neq_];
      double tau_tr = ut_r;
      if (radius > 0) tau_tr -= ut / radius;
      tau_tr *= visc;
      const double tau_tz = visc * ut_z;
      normalPrimFlux[numSpecies + nvel_ - 1] += tau_tr * nrm[0];
      normalPrimFlux[numSpecies + nvel_ - 1] += tau_tz * nrm[1];
    }
    k += 0.0;  // ke, single temperature
    for (int d = 0; d < dim_; d++) normalPrimFlux[numSpecies + nvel_] -= k * gradUp[(1 + nvel_) + d * neq_] * nrm[d];
    // species enthalpy x diffusion flux: DryAir::computeSpeciesEnthalpies returns 0 (equation_of_state.cpp)
    for (int i = numSpecies; i < primFluxSize; i++)
      if (primFluxIdxs[i]) normalPrimFlux[i] = primFlux[i];
    double vel0[3];
    for (int d = 0; d < nvel_; d++) vel0[d] = s[1 + d] / s[0];
    for (int d = 0; d < nvel_; d++) normalFlux[d + 1] = normalPrimFlux[numSpecies + d];
    for (int d = 0; d < nvel_; d++) normalFlux[nvel_ + 1] += normalPrimFlux[numSpecies + d] * vel0[d];
    normalFlux[nvel_ + 1] -= normalPrimFlux[numSpecies + nvel_];
  }

  // src/equation_of_state.hpp:610-617 (DryAir::ComputePressure)
  double pressure(const double *s) override {
    double den_vel2 = 0;
    for (int d = 0; d < nvel_; d++) den_vel2 += s[d + 1] * s[d + 1];
    den_vel2 /= s[0];
    return (p_.gamma - 1.0) * (s[1 + nvel_] - 0.5 * den_vel2);
  }
  // src/equation_of_state.cpp:360-364 (DryAir::ComputePressureFromPrimitives)
  double pressure_from_primitives(const double *Up) override { return p_.R * Up[0] * Up[1 + nvel_]; }
  // src/transport_properties.hpp:264-269 (DryAirTransport::GetViscosities)
  void get_viscosities(const double *, const double *Up, const double *, double, double, double *visc) override {
    const double temp = Up[1 + nvel_];
    visc[0] = (p_.C1 * p_.visc_mult * pow(temp, 1.5) / (temp + p_.S0));
    visc[1] = p_.bulk_visc_mult * visc[0];
  }
  // src/equation_of_state.hpp:621-627 (DryAir::ComputeTemperature)
  double temperature(const double *s) {
    double den_vel2 = 0;
    for (int d = 0; d < nvel_; d++) den_vel2 += s[d + 1] * s[d + 1];
    den_vel2 /= s[0];
    return (p_.gamma - 1.0) / p_.R * (s[1 + nvel_] - 0.5 * den_vel2) / s[0];
  }
  // src/equation_of_state.cpp:321-335
  void prim(const double *U, double *Up) override {
    double T = temperature(U);
    for (int eq = 0; eq < neq_; eq++) Up[eq] = U[eq];
    for (int d = 0; d < nvel_; d++) Up[1 + d] /= U[0];
    Up[1 + nvel_] = T;
  }
  // src/equation_of_state.cpp:298-315
  void cons(const double *Up, double *U) override {
    for (int eq = 0; eq < neq_; eq++) U[eq] = Up[eq];
    double v2 = 0.;
    for (int d = 0; d < nvel_; d++) {
      v2 += Up[1 + d] * Up[1 + d];
      U[1 + d] *= Up[0];
    }
    U[1 + nvel_] = p_.R * Up[0] * Up[1 + nvel_] / (p_.gamma - 1.) + 0.5 * Up[0] * v2;
  }
  // src/equation_of_state.cpp:279-294
  double max_char_speed(const double *s) override {
    const double den = s[0];
    double den_vel2 = 0;
    for (int d = 0; d < nvel_; d++) den_vel2 += s[d + 1] * s[d + 1];
    den_vel2 /= den;
    const double pres = pressure(s);
    const double sound = sqrt(p_.gamma * pres / den);
    const double vel = sqrt(den_vel2 / den);
    return vel + sound;
  }
  // src/fluxes.cpp:135-170 (no species, single temperature)
  void conv_flux(const double *s, double *flux) override {
    const double pres = pressure(s);
    for (int d = 0; d < dim_; d++) {
      flux[0 + d * neq_] = s[d + 1];
      for (int i = 0; i < nvel_; i++) flux[1 + i + d * neq_] = s[i + 1] * s[d + 1] / s[0];
      flux[1 + d + d * neq_] += pres;
    }
    const double H = (s[1 + nvel_] + pres) / s[0];
    for (int d = 0; d < dim_; d++) flux[1 + nvel_ + d * neq_] = s[d + 1] * H;
  }
  // src/fluxes.cpp:178-335 with DryAirTransport::ComputeFluxMolecularTransport
  // (src/transport_properties.cpp:223-234); no SGS, no sponge.
  void visc_flux(const double *s, const double *gradUp, double *xyz, double /*delta*/, double /*dist*/,
                 double *flux) override {
    const bool axisym = (dim_ == 2 && nvel_ == 3);
    const double radius = axisym ? xyz[0] : -1;
    for (int d = 0; d < dim_; d++)
      for (int eq = 0; eq < neq_; eq++) flux[eq + d * neq_] = 0.;
    if (p_.eq_system == 0) return;  // EULER, src/fluxes.cpp:185-187

    double pr = pressure(s);
    double temp = pr / p_.R / s[0];
    double visc = (p_.C1 * p_.visc_mult * pow(temp, 1.5) / (temp + p_.S0));
    double bulkViscosity = p_.bulk_visc_mult * visc;
    double k = cp_div_pr_ * visc;
    bulkViscosity -= 2. / 3. * visc;
    double ke = 0.0;
    k += ke;  // single temperature, src/fluxes.cpp:257

    double stress[9], vel[3], vtmp[3];
    double divV = 0.;
    for (int i = 0; i < dim_; i++) {
      for (int j = 0; j < dim_; j++)
        stress[i + j * dim_] = gradUp[(1 + j) + i * neq_] + gradUp[(1 + i) + j * neq_];
      divV += gradUp[(1 + i) + i * neq_];
    }
    for (int i = 0; i < dim_; i++)
      for (int j = 0; j < dim_; j++) stress[i + j * dim_] *= visc;
    const double ur = (axisym ? s[1] / s[0] : 0), ut = (axisym ? s[3] / s[0] : 0);
    if (axisym && radius > 0) divV += ur / radius;
    for (int i = 0; i < dim_; i++) stress[i + i * dim_] += bulkViscosity * divV;
    for (int i = 0; i < dim_; i++)
      for (int j = 0; j < dim_; j++) flux[(1 + i) + j * neq_] = stress[i + j * dim_];
    double tau_tr = 0, tau_tz = 0;
    if (axisym) {  // src/fluxes.cpp:285-297
      const double ut_r = gradUp[3 + 0 * neq_], ut_z = gradUp[3 + 1 * neq_];
      tau_tr = ut_r;
      if (radius > 0) tau_tr -= ut / radius;
      tau_tr *= visc;
      tau_tz = visc * ut_z;
      flux[(1 + 2) + 0 * neq_] = tau_tr;
      flux[(1 + 2) + 1 * neq_] = tau_tz;
    }

    for (int d = 0; d < dim_; d++) vel[d] = s[1 + d] / s[0];
    for (int i = 0; i < dim_; i++) {
      vtmp[i] = 0.0;
      for (int j = 0; j < dim_; j++) vtmp[i] += stress[i + j * dim_] * vel[j];
    }
    for (int d = 0; d < dim_; d++) {
      flux[(1 + nvel_) + d * neq_] += vtmp[d];
      flux[(1 + nvel_) + d * neq_] += k * gradUp[(1 + nvel_) + d * neq_];
    }
    if (axisym) {  // src/fluxes.cpp:320-323
      flux[(1 + nvel_) + 0 * neq_] += ut * tau_tr;
      flux[(1 + nvel_) + 1 * neq_] += ut * tau_tz;
    }
  }
  // src/riemann_solver.cpp:117-206 (Eval_Roe, Roe-Lohner): two velocity components, gamma - 1 = 0.4 hard-coded (:153)
  void roe(const double *state1, const double *state2, const double *nor, double *flux) {
    const int dim = dim_, NS_eq = 2 + dim;
    double normag = 0;
    for (int i = 0; i < dim; i++) normag += nor[i] * nor[i];
    normag = sqrt(normag);
    double unitN[3], f1[48], f2[48], meanFlux[16], vel[3];
    for (int d = 0; d < dim; d++) unitN[d] = nor[d] / normag;
    conv_flux(state1, f1);
    conv_flux(state2, f2);
    for (int eq = 0; eq < NS_eq; eq++) {
      meanFlux[eq] = 0.;
      for (int d = 0; d < dim; d++) meanFlux[eq] += (f1[eq + d * neq_] + f2[eq + d * neq_]) * unitN[d];
    }
    const double r = sqrt(state1[0] * state2[0]);
    for (int i = 0; i < dim; i++) {
      vel[i] = state1[i + 1] / sqrt(state1[0]) + state2[i + 1] / sqrt(state2[0]);
      vel[i] /= sqrt(state1[0]) + sqrt(state2[0]);
    }
    double qk = 0.;
    for (int d = 0; d < dim; d++) qk += vel[d] * unitN[d];
    const double p1 = pressure(state1), p2 = pressure(state2);
    double H = (state1[1 + dim] + p1) / sqrt(state1[0]) + (state2[1 + dim] + p2) / sqrt(state2[0]);
    H /= sqrt(state1[0]) + sqrt(state2[0]);
    const double a2 = 0.4 * (H - 0.5 * (vel[0] * vel[0] + vel[1] * vel[1]));
    const double a = sqrt(a2);
    double lamb[3] = {qk, qk + a, qk - a};
    if (fabs(lamb[0]) < 1e-4) lamb[0] = 1e-4;
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
    for (int i = 0; i < 4; i++) DF1[i] *= fabs(lamb[0]);
    DF4[0] = 1.;
    DF4[1] = vel[0] + unitN[0] * a;
    DF4[2] = vel[1] + unitN[1] * a;
    DF4[3] = H + qk * a;
    for (int i = 0; i < 4; i++) DF4[i] *= fabs(lamb[1]) * (deltaP + r * a * deltaQk) * 0.5 / a2;
    DF5[0] = 1.;
    DF5[1] = vel[0] - unitN[0] * a;
    DF5[2] = vel[1] - unitN[1] * a;
    DF5[3] = H - qk * a;
    for (int i = 0; i < 4; i++) DF5[i] *= fabs(lamb[2]) * (deltaP - r * a * deltaQk) * 0.5 / a2;
    for (int i = 0; i < NS_eq; i++) flux[i] = (meanFlux[i] - (DF1[i] + DF4[i] + DF5[i])) * 0.5 * normag;
  }
  // src/riemann_solver.cpp:89-114 (Eval_LF) with ComputeFluxDotN :53-64
  void riemann(const double *s1, const double *s2, const double *nor, double *flux, bool LF) override {
    if (p_.use_roe && !LF) {
      roe(s1, s2, nor, flux);
      return;
    }
    const double maxE1 = max_char_speed(s1);
    const double maxE2 = max_char_speed(s2);
    const double maxE = fmax(maxE1, maxE2);
    double f1[16 * 3], f2[16 * 3], flux1[16], flux2[16];
    conv_flux(s1, f1);
    conv_flux(s2, f2);
    for (int eq = 0; eq < neq_; eq++) {
      flux1[eq] = 0;
      flux2[eq] = 0;
      for (int d = 0; d < dim_; d++) {
        flux1[eq] += f1[eq + d * neq_] * nor[d];
        flux2[eq] += f2[eq + d * neq_] * nor[d];
      }
    }
    double normag = 0;
    for (int i = 0; i < dim_; i++) normag += nor[i] * nor[i];
    normag = sqrt(normag);
    for (int i = 0; i < neq_; i++) flux[i] = 0.5 * (flux1[i] + flux2[i]) - 0.5 * maxE * (s2[i] - s1[i]) * normag;
  }
};

Physics *make_physics(const OrcPhysParams &p, int dim, int nvel, int neq) {
  if (p.sgs_model != 0 || p.sponge_enabled != 0 || p.use_mixing_length != 0) {
    fprintf(stderr, "oracle (port back end): SGS models / viscous sponge / mixing length are served by the reference back end only\n");
    abort();
  }
  if (p.fluid != 0) return nullptr;
  if (p.use_roe && dim != 2) return nullptr;  // Eval_Roe is written for two velocity components (riemann_solver.cpp:153-170)
  return new DryAirPort(p, dim, nvel, neq);
}

}  // namespace orc
