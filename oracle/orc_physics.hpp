// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Never linked into the product.
//
// Per-point physics interface used by the DG oracle (oracle/dg_oracle.cpp).  Two
// interchangeable back ends implement it:
//   * orc_physics_port.cpp : a plain restatement of the reference formulas
//                            (cpu_baseline.kind == "port");
//   * orc_physics_ref.cpp  : thin calls into the reference's OWN, unmodified object code
//                            (DryAir, DryAirTransport, Fluxes, RiemannSolverTPS) compiled from
//                            /root/reference/src against oracle/refstub/ -- built only into
//                            oracle/_ref/ (cpu_baseline.kind == "reference").
// Array conventions are the reference's: states are [rho, rho u(nvel), rho E, ...],
// per-point matrices are column-major f[eq + d*neq] (src/fluxes.cpp:141-145).
#pragma once

extern "C" {
// Mirrors the scalar inputs of DryAirInput (src/dataStructures.hpp:609-622) and the
// DryAirTransport constructor (src/transport_properties.cpp:208-221).
struct OrcPlasma;
struct OrcPhysParams {
  int eq_system;  // Equations enum of the reference: 0 EULER, 1 NS (src/dataStructures.hpp:65-69)
  int fluid;      // WorkingFluid: 0 DRY_AIR
  double gamma;   // specific_heat_ratio
  double R;       // gas_constant
  double visc_mult;
  double bulk_visc_mult;
  double C1, S0, Pr;  // Sutherland data (src/dataStructures.hpp:205-209)
  const OrcPlasma *plasma;  // fluid == 1 (USER_DEFINED): gas / transport / chemistry models (reference back end only)
  int use_roe;              // flow/useRoe: RiemannSolverTPS::Eval_Roe on interior faces and inviscid walls (2-D dry air)
  // Fluxes' SGS model and planar viscous sponge (src/fluxes.cpp:97-128; reference back end only)
  int sgs_model;            // flow/sgsModel: 0 none, 1 smagorinsky, 2 sigma
  double sgs_const, sgs_floor;
  int sponge_enabled;       // viscosityMultiplierFunction/isEnabled
  double sponge_normal[3], sponge_point[3], sponge_ratio, sponge_width;
  // flow/useMixingLength: MixingLengthTransport wrapped around the molecular transport (src/M2ulPhyS.cpp:265-283,
  // mixingLengthTransportData src/dataStructures.hpp:548-554; reference back end only)
  int use_mixing_length;
  double max_mixing_length, mixing_length_Prt, mixing_length_bulk_mult;
  const struct OrcLte *lte;  // fluid == 2 (LTE_FLUID): 1-D look-up tables (reference back end only)
};
// LteMixture / LteTransport with 1-D tables (flow/lte/table_dim = 1, src/M2ulPhyS.cpp:175-258): thermodynamic table
// T -> (energy, R, c) and its inverse energy -> T, transport table T -> (mu, kappa, sigma), optional net emission
// coefficient.  Same layout as tpsb_lte_tables.
struct OrcLte {
  int num_thermo;
  const double *T, *energy, *R, *c;
  int num_trans;
  const double *T_trans, *mu, *kappa, *sigma;
  int nec_table_n, nec_table_xlog, nec_table_flog;
  const double *nec_table_x, *nec_table_f;
};
// Plasma models of a user-defined fluid: PerfectMixtureInput + constantTransportData + ChemistryInput
// (src/dataStructures.hpp:537-546,623-633,690-712) flattened; same layout as tpsb_plasma_models.
struct OrcPlasma {
  int num_species, ambipolar, two_temperature;
  double mw[8], charge[8], formation_energy[8], molar_cv[8];
  int transport_model;
  double viscosity, bulk_viscosity, thermal_conductivity, electron_thermal_conductivity;
  double diffusivity[8], mt_freq[8];
  int num_reactions;
  double min_temperature;
  int model[34], detailed_balance[34];
  double rate_params[34][3];
  double reaction_energy[34];
  double equilibrium_params[34][3];
  int reactant_stoich[34][8], product_stoich[34][8];
  // transport_model == 0 (ARGON_MINIMAL): GasTransportInput scalars (src/dataStructures.hpp:644-666)
  int third_order_k_electron, multiply;
  double flux_trns_multiplier[4], mf_freq_multiplier, diff_mult, mobil_mult;
  // TABULATED_RXN tables (TableInput) and GRIDFUNCTION_RXN components
  int table_n[34], table_xlog[34], table_flog[34];
  const double *table_x[34], *table_f[34];
  int rate_component[34];
  // transport_model == 1 (ARGON_MIXTURE): GasTransportInput::collisionIndex / ionIndex / neutralIndex
  int collision_index[64];
  int ion_index, neutral_index;
  // RadiationInput: NET_EMISSION / TABULATED_NEC table (nec_table_n = 0: none)
  int nec_table_n, nec_table_xlog, nec_table_flog;
  const double *nec_table_x, *nec_table_f;
};
// One boundary condition of BCintegrator's attribute maps (src/BCintegrator.cpp:64-125).
// kind: 0 inlet, 1 outlet, 2 wall; type: the reference's InletType / OutletType / WallType value
// (src/dataStructures.hpp:168-196); data: inlet inputState (rho, u, v, w, rho Y_sp of the active species ...),
// outlet inputState (p), wall Th.
struct OrcBc {
  int attr, kind, type;
  double data[12];
};
// same layout as tpsb_forcing_desc (include/tpsb200.h): kind 0 pressure gradient, 1 heat source, 2 Joule heating, 3 sponge zone
struct OrcForcing {
  int kind;
  double pressure_grad[3];
  double hs_point1[3], hs_point2[3], hs_radius, hs_value;
  const double *joule_heating;
  int sz_type, sz_mixed_out;
  double sz_normal[3], sz_point0[3], sz_point_init[3];
  double sz_r1, sz_r2, sz_tol, sz_mult;
  double sz_target[5];
};
}

namespace orc {

struct Physics {
  virtual ~Physics() {}
  virtual const char *kind() const = 0;
  // GasMixture::GetPrimitivesFromConservatives
  virtual void prim(const double *U, double *Up) = 0;
  // GasMixture::GetConservativesFromPrimitives
  virtual void cons(const double *Up, double *U) = 0;
  // GasMixture::ComputeMaxCharSpeed
  virtual double max_char_speed(const double *U) = 0;
  // Fluxes::ComputeConvectiveFluxes
  virtual void conv_flux(const double *U, double *F) = 0;
  // Fluxes::ComputeViscousFluxes
  virtual void visc_flux(const double *U, const double *gradUp, double *xyz, double delta, double dist,
                         double *F) = 0;
  // RiemannSolverTPS::Eval(state1, state2, nor, flux, LF): Eval_Roe when useRoe && !LF, else Eval_LF
  virtual void riemann(const double *U1, const double *U2, const double *nor, double *flux, bool LF = false) = 0;
  virtual int num_active_species() const = 0;
  // SourceTerm::updateTerms for one node (src/source_term.cpp:117-250): Un = conserved state of the solution
  // grid function, upn / gradUpn = primitives and their gradients; both may be clamped in place like the
  // reference does.  Fluids without plasma sources keep the default (no forcing term registered).
  virtual bool has_source() const { return false; }
  virtual void source_term(double *Un, double *upn, const double *gradUpn, int node, double *src) {}
  // Chemistry::setRates: rate coefficients of the GRIDFUNCTION_RXN reactions, data[component][size]
  virtual void set_rates(const double *data, int size) {}
  // Whole-vector forcing term through the reference's own SourceTerm::updateTerms object code (src/source_term.cpp:62-255):
  // y += S(Usol, Up, gradUp), byNODES arrays of N nodes.  false: back end has no such object (the per-node restatement runs).
  virtual bool source_update(const double *Usol, const double *Up, const double *gradUp, long N, double *y) { return false; }
  // Wall boundary flux through the reference's own WallBC::computeBdrFlux object code (src/wallBC.cpp:268-543).
  // data = the wall's inputs as in OrcBc (VISC_ISOTH: {Th}; VISC_GNRL: {hvyThermalCond, elecThermalCond, Th, Te}).
  // false: back end has no such object (the restatement in dg_oracle.cpp runs).
  virtual bool wall_bc_flux(int wall_type, const double *data, bool use_bc_in_grad, const double *normal, const double *stateIn,
                            const double *gradState, const double *xyz, double delta, double dist, double *bdrFlux) {
    return false;
  }
  // MolecularTransport::computeMixtureAverageDiffusivity (gas_transport.hpp:132): what utils/binary_mixture_ic.cpp feeds
  // its analytic solution with; false when the back end has no collision-integral transport
  virtual bool mixture_average_diffusivity(const double *U, double *D) { return false; }
  // ---- used by AxisymmetricSource (src/forcing_terms.cpp:255-380) ----
  // GasMixture::ComputePressureFromPrimitives
  virtual double pressure_from_primitives(const double *Up) = 0;
  // TransportProperties::GetViscosities(conserved, primitive, gradUp, radius, distance, visc[2])
  virtual void get_viscosities(const double *U, const double *Up, const double *gradUp, double radius, double dist,
                               double *visc) = 0;
  // ---- used by the boundary conditions ----
  virtual int num_species() const = 0;
  // GasMixture::ComputePressure
  virtual double pressure(const double *U) = 0;
  // GasMixture::computeStagnationState / computeStagnantStateWithTemp / modifyEnergyForPressure
  virtual void stagnation_state(const double *U, double *out) = 0;
  virtual void stagnant_state_with_temp(const double *U, double T, double *out) = 0;
  virtual void modify_energy_for_pressure(const double *in, double *out, double p, bool modifyElectronEnergy) = 0;
  // GasMixture::modifyStateFromPrimitive (src/equation_of_state.cpp:118-143): primitives of U, entries flagged in
  // primIdxs replaced by prim, back to conserved
  virtual void modify_state_from_primitive(const double *U, const double *prim, const bool *primIdxs, double *out) = 0;
  // GasMixture::computeSheathBdrFlux (PerfectMixture: src/equation_of_state.cpp:1909-1942): fills primFlux
  virtual void sheath_bdr_flux(const double *wallState, double *primFlux) {}
  // Fluxes::ComputeBdrViscousFluxes with BoundaryViscousFluxData {normal (unit), primFlux, primFluxIdxs}
  virtual void bdr_visc_flux(const double *U, const double *gradUp, double *xyz, double delta, double dist,
                             const double *unit_normal, const double *primFlux, const bool *primFluxIdxs,
                             double *normalFlux) = 0;
};

// Implemented by exactly one of orc_physics_port.cpp / orc_physics_ref.cpp per library.
Physics *make_physics(const OrcPhysParams &p, int dim, int nvel, int neq);

}  // namespace orc
