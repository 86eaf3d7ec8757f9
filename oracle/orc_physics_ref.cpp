// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Never linked into the product.
//
// "reference" physics back end: forwards every per-point call to the reference's own,
// unmodified classes, whose sources are compiled in place from /root/reference/src
// (equation_of_state.cpp, transport_properties.cpp, fluxes.cpp, riemann_solver.cpp, ...)
// against the header stub in oracle/refstub/.  Built only into oracle/_ref/ by
// oracle/Makefile; the resulting shared object travels to the GPU box, the sources do not.
#include <cstring>
#include <omp.h>

#include "chemistry.hpp"
#include "equation_of_state.hpp"
#include "fluxes.hpp"
#include "gas_transport.hpp"
#include "lte_mixture.hpp"
#include "lte_transport_properties.hpp"
#include "radiation.hpp"
#include "mixing_length_transport.hpp"
#include "riemann_solver.hpp"
#include "transport_properties.hpp"

#include "source_term.hpp"
#include "wallBC.hpp"

#include "orc_physics.hpp"

// ---- glue for the two reference classes whose BASE-class constructors live in translation units that need real MFEM
// (forcing_terms.cpp, BoundaryCondition.cpp): member initialisation only, restated from src/forcing_terms.cpp:36-52 and
// src/BoundaryCondition.cpp:34-57.  Everything that computes -- SourceTerm::updateTerms, WallBC::computeBdrFlux and the
// per-type wall routines -- is the reference's own object code (source_term.cpp, wallBC.cpp compiled in place).
ForcingTerms::ForcingTerms(const int &_dim, const int &_num_equation, const int &_order, const int &_intRuleType,
                           IntegrationRules *_intRules, ParFiniteElementSpace *_vfes, ParGridFunction *U, ParGridFunction *_Up,
                           ParGridFunction *_gradUp, const precomputedIntegrationData &gpu_precomputed_data, bool axisym)
    : dim(_dim), nvel(axisym ? 3 : _dim), num_equation(_num_equation), axisymmetric_(axisym), order(_order),
      intRuleType(_intRuleType), intRules(_intRules), vfes(_vfes), U_(U), Up_(_Up), gradUp_(_gradUp),
      gpu_precomputed_data_(gpu_precomputed_data), h_num_elems_of_type(nullptr) {}
ForcingTerms::~ForcingTerms() {}
BoundaryCondition::BoundaryCondition(RiemannSolverTPS *_rsolver, GasMixture *_mixture, Equations _eqSystem,
                                     ParFiniteElementSpace *_vfes, IntegrationRules *_intRules, double &_dt, const int _dim,
                                     const int _num_equation, const int _patchNumber, const double _refLength, bool axisym)
    : rsolver(_rsolver), mixture(_mixture), eqSystem(_eqSystem), vfes(_vfes), intRules(_intRules), dt(_dt), dim_(_dim),
      nvel_(axisym ? 3 : _dim), num_equation_(_num_equation), patchNumber(_patchNumber), refLength(_refLength),
      axisymmetric_(axisym), BCinit(false) {}
BoundaryCondition::~BoundaryCondition() {}
void BoundaryCondition::computeBdrPrimitiveStateForGradient(const Vector &stateIn, Vector &stateBC) const { stateBC = stateIn; }

namespace {
// one reference WallBC object per (type, inputs, useBCinGrad), built on first use.  WallBC::computeBdrFlux writes its
// bcFlux_ member (the reference's CPU path is single-threaded per rank), so every OpenMP thread of the oracle owns a cache.
struct WallCache {
  struct Entry {
    int type;
    bool use;
    double data[4];
    WallBC *bc;
  };
  std::vector<Entry> entries;
  boundaryFaceIntegrationData bfd;
  int maxIntPoints = 64;
  double dt = 0.0;
  WallBC *get(RiemannSolverTPS *rs, GasMixture *mix, Equations eqs, Fluxes *flux, int dim, int neq, bool axisym, int type,
              const double *data, bool use) {
    for (auto &e : entries)
      if (e.type == type && e.use == use && e.data[0] == data[0] && e.data[1] == data[1] && e.data[2] == data[2] &&
          e.data[3] == data[3])
        return e.bc;
    WallData wd;
    wd.hvyThermalCond = NONE_THMCND, wd.elecThermalCond = NONE_THMCND, wd.Th = 0.0, wd.Te = 0.0;
    if (type == VISC_ISOTH) {  // M2ulPhyS::parseBCInputs: Th = Te = the wall temperature (src/M2ulPhyS.cpp, wall inputs)
      wd.hvyThermalCond = ISOTH, wd.elecThermalCond = ISOTH, wd.Th = data[0], wd.Te = data[0];
    } else if (type == VISC_GNRL) {
      wd.hvyThermalCond = static_cast<ThermalCondition>(static_cast<int>(data[0]));
      wd.elecThermalCond = static_cast<ThermalCondition>(static_cast<int>(data[1]));
      wd.Th = data[2], wd.Te = data[3];
    }
    WallBC *bc = new WallBC(rs, mix, mix, eqs, flux, nullptr, nullptr, dt, dim, neq, 1, static_cast<WallType>(type), wd, bfd,
                            maxIntPoints, axisym, use);  // wallBC.cpp:36
    entries.push_back({type, use, {data[0], data[1], data[2], data[3]}, bc});
    return bc;
  }
  bool flux(RiemannSolverTPS *rs, GasMixture *mix, Equations eqs, Fluxes *fl, int dim, int neq, bool axisym, int type,
            const double *data, bool use, const double *normal, const double *stateIn, const double *gradState,
            const double *xyz, double delta, double dist, double *bdrFlux) {
    WallBC *bc = get(rs, mix, eqs, fl, dim, neq, axisym, type, data, use);
    Vector n(dim), s(neq), t(3), f(neq);
    DenseMatrix g(neq, dim);
    for (int d = 0; d < dim; d++) n[d] = normal[d];
    for (int i = 0; i < neq; i++) s[i] = stateIn[i];
    for (int i = 0; i < neq * dim; i++) g.GetData()[i] = gradState[i];
    for (int d = 0; d < 3; d++) t[d] = d < dim ? xyz[d] : 0.0;
    f = 0.0;
    bc->computeBdrFlux(n, s, g, t, delta, 0.0, dist, f);  // wallBC.cpp:268
    for (int i = 0; i < neq; i++) bdrFlux[i] = f[i];
    return true;
  }
};
}  // namespace

// flow/useMixingLength (src/M2ulPhyS.cpp:265-283): the flux class sees the molecular transport through MixingLengthTransport
static TransportProperties *wrap_mixing_length(const OrcPhysParams &p, GasMixture *mix, MolecularTransport *molecular) {
  if (!p.use_mixing_length) return molecular;
  mixingLengthTransportData d;
  d.max_mixing_length_ = p.max_mixing_length;
  d.Prt_ = p.mixing_length_Prt;
  d.Let_ = 1.0;
  d.bulk_multiplier_ = p.mixing_length_bulk_mult;
  return new MixingLengthTransport(mix, d, molecular);
}


// Fluxes with or without its SGS model / viscous sponge: the constructor M2ulPhyS uses (fluxes.cpp:57-95) reads these from
// RunConfiguration and normalises the sponge normal; the device constructor (fluxes.cpp:97-128) takes them as given
static Fluxes *make_fluxes(const OrcPhysParams &p, GasMixture *mix, Equations eqs, TransportProperties *trans, int neq, int dim,
                           bool axisym) {
  if (p.sgs_model == 0 && p.sponge_enabled == 0) return new Fluxes(mix, eqs, trans, neq, dim, axisym);  // fluxes.cpp:34
  viscositySpongeData vsd;
  vsd.enabled = p.sponge_enabled != 0;
  double nm = 0;
  for (int d = 0; d < 3; d++) nm += p.sponge_normal[d] * p.sponge_normal[d];
  nm = std::sqrt(nm);
  for (int d = 0; d < 3; d++) {
    vsd.n[d] = vsd.enabled ? p.sponge_normal[d] / nm : 0.0;
    vsd.p[d] = p.sponge_point[d];
  }
  vsd.ratio = p.sponge_ratio;
  vsd.width = p.sponge_width;
  return new Fluxes(mix, eqs, trans, neq, dim, axisym, p.sgs_model, p.sgs_floor, p.sgs_const, vsd);
}

namespace orc {

// Single-species fluids: DryAir + DryAirTransport, or LteMixture + LteTransport over 1-D tables (fluid == 2).
class DryAirRef : public Physics {
  int dim_, nvel_, neq_;
  GasMixture *mix_;
  MolecularTransport *trans_;
  Fluxes *flux_;
  RiemannSolverTPS *rs_;
  NetEmission *rad_ = nullptr;
  bool use_roe_ = false;
  Equations eqs_;
  WallCache walls_[256];

  static TableInput table(int n, const double *x, const double *f) {
    TableInput t;
    t.Ndata = n, t.xdata = x, t.fdata = f, t.xLogScale = false, t.fLogScale = false, t.order = 1;
    return t;
  }

 public:
  DryAirRef(const OrcPhysParams &p, int dim, int nvel, int neq) : dim_(dim), nvel_(nvel), neq_(neq) {
    if (p.fluid == 2) {  // M2ulPhyS::initMixtureAndTransportModels, table_dim == 1 (M2ulPhyS.cpp:175-258)
      const OrcLte &l = *p.lte;
      // energy, R and c over T; the e -> T table is the T -> e table with its columns swapped (M2ulPhyS.cpp:193-200)
      mix_ = new LteMixture(LTE_FLUID, dim, nvel, 0.0, table(l.num_thermo, l.T, l.energy), table(l.num_thermo, l.T, l.R),
                            table(l.num_thermo, l.T, l.c), table(l.num_thermo, l.energy, l.T));
      trans_ = new LteTransport(mix_, table(l.num_trans, l.T_trans, l.mu), table(l.num_trans, l.T_trans, l.kappa),
                                table(l.num_trans, l.T_trans, l.sigma));
      if (l.nec_table_n > 0) {
        RadiationInput ri;
        ri.model = NET_EMISSION;
        ri.necModel = TABULATED_NEC;
        ri.necTableInput = table(l.nec_table_n, l.nec_table_x, l.nec_table_f);
        ri.necTableInput.xLogScale = l.nec_table_xlog != 0, ri.necTableInput.fLogScale = l.nec_table_flog != 0;
        rad_ = new NetEmission(ri);
      }
    } else {
      DryAirInput in;
      in.f = DRY_AIR;
      in.eq_sys = static_cast<Equations>(p.eq_system);
      in.specific_heat_ratio = p.gamma;
      in.gas_constant = p.R;
      mix_ = new DryAir(in, dim, nvel);                                                          // equation_of_state.cpp:150
      trans_ = new DryAirTransport(mix_, p.visc_mult, p.bulk_visc_mult, p.C1, p.S0, p.Pr);       // transport_properties.cpp:208
    }
    const bool axisym = (dim == 2 && nvel == 3);  // config.isAxisymmetric()
    flux_ = make_fluxes(p, mix_, static_cast<Equations>(p.eq_system), wrap_mixing_length(p, mix_, trans_), neq, dim, axisym);
    rs_ = new RiemannSolverTPS(neq, mix_, static_cast<Equations>(p.eq_system), flux_, p.use_roe != 0, axisym);  // riemann_solver.cpp:38
    use_roe_ = p.use_roe != 0;
    eqs_ = static_cast<Equations>(p.eq_system);
  }
  ~DryAirRef() {
    delete rad_;
    delete rs_;
    delete flux_;
    delete trans_;
    delete mix_;
  }
  const char *kind() const override { return "reference"; }
  bool wall_bc_flux(int wall_type, const double *data, bool use, const double *normal, const double *stateIn,
                    const double *gradState, const double *xyz, double delta, double dist, double *bdrFlux) override {
    return walls_[omp_get_thread_num() & 255].flux(rs_, mix_, eqs_, flux_, dim_, neq_, dim_ == 2 && nvel_ == 3, wall_type, data, use, normal, stateIn,
                       gradState, xyz, delta, dist, bdrFlux);
  }
  int num_active_species() const override { return mix_->GetNumActiveSpecies(); }
  int num_species() const override { return mix_->GetNumSpecies(); }
  double pressure(const double *U) override { return mix_->ComputePressure(U); }
  double pressure_from_primitives(const double *Up) override { return mix_->ComputePressureFromPrimitives(Up); }
  void get_viscosities(const double *U, const double *Up, const double *gradUp, double radius, double dist,
                       double *visc) override {
    trans_->GetViscosities(U, Up, gradUp, radius, dist, visc);
  }
  void stagnation_state(const double *U, double *out) override {
    Vector a(const_cast<double *>(U), neq_), b(neq_);
    mix_->computeStagnationState(a, b);  // equation_of_state.cpp:365
    for (int i = 0; i < neq_; i++) out[i] = b[i];
  }
  void stagnant_state_with_temp(const double *U, double T, double *out) override {
    Vector a(const_cast<double *>(U), neq_), b(neq_);
    mix_->computeStagnantStateWithTemp(a, T, b);  // equation_of_state.cpp:379
    for (int i = 0; i < neq_; i++) out[i] = b[i];
  }
  void modify_energy_for_pressure(const double *in, double *out, double p, bool mee) override {
    double tmp[16];
    for (int i = 0; i < neq_; i++) tmp[i] = in[i];
    mix_->modifyEnergyForPressure(tmp, out, p, mee);  // equation_of_state.cpp:402
  }
  void modify_state_from_primitive(const double *U, const double *prim, const bool *primIdxs, double *out) override {
    BoundaryPrimitiveData bc;
    for (int i = 0; i < gpudata::MAXEQUATIONS; i++) {
      bc.prim[i] = i < 16 ? prim[i] : 0.0;
      bc.primIdxs[i] = i < 16 ? primIdxs[i] : false;
    }
    mix_->modifyStateFromPrimitive(U, bc, out);  // equation_of_state.cpp:134
  }
  void bdr_visc_flux(const double *U, const double *gradUp, double *xyz, double delta, double dist, const double *nrm,
                     const double *primFlux, const bool *primFluxIdxs, double *normalFlux) override {
    BoundaryViscousFluxData bc;
    for (int d = 0; d < gpudata::MAXDIM; d++) bc.normal[d] = d < dim_ ? nrm[d] : 0.0;
    for (int i = 0; i < gpudata::MAXEQUATIONS; i++) {
      bc.primFlux[i] = i < 16 ? primFlux[i] : 0.0;
      bc.primFluxIdxs[i] = i < 16 ? primFluxIdxs[i] : false;
    }
    flux_->ComputeBdrViscousFluxes(U, gradUp, xyz, delta, dist, bc, normalFlux);  // fluxes.cpp:344
  }
  // SourceTerm::updateTerms for one species without reactions (source_term.cpp:117-250): only the radiative energy sink
  // (:205-207) reaches dU/dt; the plasma conductivity it also writes is an output field, not part of the residual
  bool has_source() const override { return rad_ != nullptr; }
  void source_term(double *Un, double *upn, const double *gradUpn, int n, double *srcTerm) override {
    for (int eq = 0; eq < neq_; eq++) srcTerm[eq] = 0.0;
    if (rad_) srcTerm[1 + nvel_] += rad_->computeEnergySink(upn[1 + nvel_]);
  }
  void prim(const double *U, double *Up) override { mix_->GetPrimitivesFromConservatives(U, Up); }
  void cons(const double *Up, double *U) override { mix_->GetConservativesFromPrimitives(Up, U); }
  double max_char_speed(const double *U) override { return mix_->ComputeMaxCharSpeed(U); }
  void conv_flux(const double *U, double *F) override { flux_->ComputeConvectiveFluxes(U, F); }
  void visc_flux(const double *U, const double *gradUp, double *xyz, double delta, double dist, double *F) override {
    flux_->ComputeViscousFluxes(U, gradUp, xyz, delta, dist, F);
  }
  void riemann(const double *U1, const double *U2, const double *nor, double *flux, bool LF) override {
    if (use_roe_ && !LF) {  // Eval_Roe exists only behind the Vector overload (riemann_solver.cpp:66-83,117-206)
      Vector a(const_cast<double *>(U1), neq_), b(const_cast<double *>(U2), neq_), n(const_cast<double *>(nor), dim_), f(neq_);
      rs_->Eval(a, b, n, f, false);
      for (int i = 0; i < neq_; i++) flux[i] = f[i];
    } else {
      rs_->Eval(U1, U2, nor, flux, true);
    }
  }
};

// PerfectMixture + ConstantTransport + Chemistry + Fluxes + RiemannSolverTPS, the reference's own classes.
class MixtureRef : public Physics {
  int dim_, nvel_, neq_;
  PerfectMixture *mix_;
  TransportProperties *trans_;
  Fluxes *flux_;
  RiemannSolverTPS *rs_;
  Chemistry *chem_ = nullptr;
  NetEmission *rad_ = nullptr;
  double rxParams_[34][3];
  Equations eqs_;
  WallCache walls_[256];
  // the reference's SourceTerm over stub grid functions of the current mesh size (built on first use)
  RunConfiguration cfg_;
  precomputedIntegrationData pre_;
  ParFiniteElementSpace *st_fes_ = nullptr;
  ParGridFunction *st_U_ = nullptr, *st_Up_ = nullptr, *st_gradUp_ = nullptr;
  SourceTerm *st_ = nullptr;
  int st_order_ = 0, st_rule_ = 0;
  long st_N_ = -1;

 public:
  MixtureRef(const OrcPhysParams &p, int dim, int nvel, int neq) : dim_(dim), nvel_(nvel), neq_(neq) {
    const OrcPlasma &pm = *p.plasma;
    PerfectMixtureInput in;
    in.f = USER_DEFINED;
    in.numSpecies = pm.num_species;
    in.isElectronIncluded = true;
    in.ambipolar = pm.ambipolar != 0;
    in.twoTemperature = pm.two_temperature != 0;
    for (int sp = 0; sp < pm.num_species; sp++) {
      in.gasParams[sp + GasParams::SPECIES_MW * pm.num_species] = pm.mw[sp];
      in.gasParams[sp + GasParams::SPECIES_CHARGES * pm.num_species] = pm.charge[sp];
      in.gasParams[sp + GasParams::FORMATION_ENERGY * pm.num_species] = pm.formation_energy[sp];
      in.gasParams[sp + GasParams::SPECIES_DEGENERACY * pm.num_species] = 1.0;
      in.molarCV[sp] = pm.molar_cv[sp];
    }
    mix_ = new PerfectMixture(in, dim, nvel);  // equation_of_state.cpp:478
    constantTransportData ct;
    ct.viscosity = pm.viscosity;
    ct.bulkViscosity = pm.bulk_viscosity;
    ct.thermalConductivity = pm.thermal_conductivity;
    ct.electronThermalConductivity = pm.electron_thermal_conductivity;
    for (int sp = 0; sp < gpudata::MAXSPECIES; sp++) {
      ct.diffusivity[sp] = sp < pm.num_species ? pm.diffusivity[sp] : 0.0;
      ct.mtFreq[sp] = sp < pm.num_species ? pm.mt_freq[sp] : 0.0;
    }
    ct.electronIndex = pm.num_species - 2;
    if (pm.transport_model == 0 || pm.transport_model == 1) {  // ARGON_MINIMAL / ARGON_MIXTURE (gas_transport.cpp:42, 877)
      GasTransportInput gi;
      memset(&gi, 0, sizeof(gi));
      gi.gas = Ar;
      gi.ionIndex = 0, gi.electronIndex = pm.num_species - 2, gi.neutralIndex = pm.num_species - 1;
      if (pm.transport_model == 1) {
        gi.ionIndex = pm.ion_index, gi.neutralIndex = pm.neutral_index;
        for (int i = 0; i < pm.num_species; i++)
          for (int j = i; j < pm.num_species; j++)
            gi.collisionIndex[i + j * pm.num_species] = static_cast<GasColl>(pm.collision_index[i + j * pm.num_species]);
      }
      gi.neutralIndex2 = -1, gi.ionIndex2 = -1;
      gi.thirdOrderkElectron = pm.third_order_k_electron != 0;
      gi.multiply = pm.multiply != 0;
      for (int t = 0; t < 4; t++) gi.fluxTrnsMultiplier[t] = pm.flux_trns_multiplier[t];
      gi.spcsTrnsMultiplier[0] = pm.mf_freq_multiplier;
      gi.diffMult = pm.diff_mult, gi.mobilMult = pm.mobil_mult;
      if (pm.transport_model == 1)
        trans_ = new GasMixtureTransport(mix_, gi);
      else
        trans_ = new GasMinimalTransport(mix_, gi);
    } else {
      trans_ = new ConstantTransport(mix_, ct);  // transport_properties.cpp:303
    }
    const Equations eqs = static_cast<Equations>(p.eq_system);
    const bool axisym = (dim == 2 && nvel == 3);  // config.isAxisymmetric()
    flux_ = make_fluxes(p, mix_, eqs, wrap_mixing_length(p, mix_, static_cast<MolecularTransport *>(trans_)), neq, dim, axisym);
    rs_ = new RiemannSolverTPS(neq, mix_, eqs, flux_, false, axisym);
    eqs_ = eqs;
    cfg_.workFluid = USER_DEFINED;
    cfg_.axisymmetric_ = axisym;
    cfg_.chemistryInput.numReactions = pm.num_reactions;
    cfg_.radiationInput.model = pm.nec_table_n > 0 ? NET_EMISSION : NONE_RAD;
    if (pm.num_reactions > 0) {
      ChemistryInput ci;
      ci.model = NUM_CHEMISTRYMODEL;
      ci.electronIndex = pm.num_species - 2;
      ci.numReactions = pm.num_reactions;
      ci.minimumTemperature = pm.min_temperature;
      for (int r = 0; r < pm.num_reactions; r++) {
        ci.reactionEnergies[r] = pm.reaction_energy[r];
        ci.detailedBalance[r] = pm.detailed_balance[r] != 0;
        ci.reactionModels[r] = static_cast<ReactionModel>(pm.model[r]);
        for (int k = 0; k < 3; k++) {
          rxParams_[r][k] = pm.rate_params[r][k];
          ci.equilibriumConstantParams[k + r * gpudata::MAXCHEMPARAMS] = pm.equilibrium_params[r][k];
        }
        ci.reactionInputs[r].modelParams = rxParams_[r];
        ci.reactionInputs[r].indexInput = pm.rate_component[r];
        if (pm.model[r] == TABULATED_RXN) {
          TableInput &ti = ci.reactionInputs[r].tableInput;
          ti.Ndata = pm.table_n[r], ti.xdata = pm.table_x[r], ti.fdata = pm.table_f[r];
          ti.xLogScale = pm.table_xlog[r] != 0, ti.fLogScale = pm.table_flog[r] != 0, ti.order = 1;
        }
        for (int sp = 0; sp < pm.num_species; sp++) {
          ci.reactantStoich[sp + r * pm.num_species] = static_cast<int16_t>(pm.reactant_stoich[r][sp]);
          ci.productStoich[sp + r * pm.num_species] = static_cast<int16_t>(pm.product_stoich[r][sp]);
        }
      }
      chem_ = new Chemistry(mix_, ci);  // chemistry.cpp:40
    }
    if (pm.nec_table_n > 0) {  // RadiationInput -> NetEmission (radiation.cpp:35)
      RadiationInput ri;
      ri.model = NET_EMISSION;
      ri.necModel = TABULATED_NEC;
      ri.necTableInput.Ndata = pm.nec_table_n, ri.necTableInput.xdata = pm.nec_table_x, ri.necTableInput.fdata = pm.nec_table_f;
      ri.necTableInput.xLogScale = pm.nec_table_xlog != 0, ri.necTableInput.fLogScale = pm.nec_table_flog != 0;
      ri.necTableInput.order = 1;
      rad_ = new NetEmission(ri);
    }
  }
  ~MixtureRef() {
    delete st_;
    delete st_U_;
    delete st_Up_;
    delete st_gradUp_;
    delete st_fes_;
    delete rad_;
    delete chem_;
    delete rs_;
    delete flux_;
    delete trans_;
    delete mix_;
  }
  const char *kind() const override { return "reference"; }
  int num_active_species() const override { return mix_->GetNumActiveSpecies(); }
  int num_species() const override { return mix_->GetNumSpecies(); }
  void prim(const double *U, double *Up) override { mix_->GetPrimitivesFromConservatives(U, Up); }
  void cons(const double *Up, double *U) override { mix_->GetConservativesFromPrimitives(Up, U); }
  double max_char_speed(const double *U) override { return mix_->ComputeMaxCharSpeed(U); }
  void conv_flux(const double *U, double *F) override { flux_->ComputeConvectiveFluxes(U, F); }
  void visc_flux(const double *U, const double *gradUp, double *xyz, double delta, double dist, double *F) override {
    flux_->ComputeViscousFluxes(U, gradUp, xyz, delta, dist, F);
  }
  void riemann(const double *U1, const double *U2, const double *nor, double *flux, bool) override {
    rs_->Eval(U1, U2, nor, flux, false);
  }
  double pressure(const double *U) override { return mix_->ComputePressure(U); }
  double pressure_from_primitives(const double *Up) override { return mix_->ComputePressureFromPrimitives(Up); }
  void get_viscosities(const double *U, const double *Up, const double *gradUp, double radius, double dist,
                       double *visc) override {
    trans_->GetViscosities(U, Up, gradUp, radius, dist, visc);
  }
  void stagnation_state(const double *U, double *out) override {
    Vector a(const_cast<double *>(U), neq_), b(neq_);
    mix_->computeStagnationState(a, b);
    for (int i = 0; i < neq_; i++) out[i] = b[i];
  }
  void stagnant_state_with_temp(const double *U, double T, double *out) override {
    Vector a(const_cast<double *>(U), neq_), b(neq_);
    mix_->computeStagnantStateWithTemp(a, T, b);
    for (int i = 0; i < neq_; i++) out[i] = b[i];
  }
  void modify_energy_for_pressure(const double *in, double *out, double p, bool mee) override {
    double tmp[16];
    for (int i = 0; i < neq_; i++) tmp[i] = in[i];
    mix_->modifyEnergyForPressure(tmp, out, p, mee);
  }
  void modify_state_from_primitive(const double *U, const double *prim, const bool *primIdxs, double *out) override {
    BoundaryPrimitiveData bc;
    for (int i = 0; i < gpudata::MAXEQUATIONS; i++) {
      bc.prim[i] = i < 16 ? prim[i] : 0.0;
      bc.primIdxs[i] = i < 16 ? primIdxs[i] : false;
    }
    mix_->modifyStateFromPrimitive(U, bc, out);  // equation_of_state.cpp:134
  }
  void sheath_bdr_flux(const double *wallState, double *primFlux) override {
    BoundaryViscousFluxData bc;
    for (int i = 0; i < gpudata::MAXEQUATIONS; i++) {
      bc.primFlux[i] = i < 16 ? primFlux[i] : 0.0;
      bc.primFluxIdxs[i] = false;
    }
    mix_->computeSheathBdrFlux(wallState, bc);  // equation_of_state.cpp:1909
    for (int i = 0; i < 16 && i < gpudata::MAXEQUATIONS; i++) primFlux[i] = bc.primFlux[i];
  }
  void bdr_visc_flux(const double *U, const double *gradUp, double *xyz, double delta, double dist, const double *nrm,
                     const double *primFlux, const bool *primFluxIdxs, double *normalFlux) override {
    BoundaryViscousFluxData bc;
    for (int d = 0; d < gpudata::MAXDIM; d++) bc.normal[d] = d < dim_ ? nrm[d] : 0.0;
    for (int i = 0; i < gpudata::MAXEQUATIONS; i++) {
      bc.primFlux[i] = i < 16 ? primFlux[i] : 0.0;
      bc.primFluxIdxs[i] = i < 16 ? primFluxIdxs[i] : false;
    }
    flux_->ComputeBdrViscousFluxes(U, gradUp, xyz, delta, dist, bc, normalFlux);
  }
  bool has_source() const override { return true; }
  // GridFunctionReaction::setGridFunction semantics (src/reaction.cpp:93-106): component c at data + c * size.
  // (Chemistry::setRates -> setData offsets by the PREVIOUS size, src/reaction.cpp:88-91; calling it twice gives the
  // intended offset.)
  void set_rates(const double *data, int size) override {
    if (chem_) {
      chem_->setRates(data, size);
      chem_->setRates(data, size);
    }
  }
  // SourceTerm::updateTerms node body (src/source_term.cpp:117-250) over the reference's transport / chemistry /
  // mixture objects; no radiation, no EM coupling output.
  bool wall_bc_flux(int wall_type, const double *data, bool use, const double *normal, const double *stateIn,
                    const double *gradState, const double *xyz, double delta, double dist, double *bdrFlux) override {
    return walls_[omp_get_thread_num() & 255].flux(rs_, mix_, eqs_, flux_, dim_, neq_, dim_ == 2 && nvel_ == 3, wall_type, data, use, normal, stateIn,
                       gradState, xyz, delta, dist, bdrFlux);
  }
  bool source_update(const double *Usol, const double *Up, const double *gradUp, long N, double *y) override {
    if (st_N_ != N) {
      delete st_;
      delete st_U_;
      delete st_Up_;
      delete st_gradUp_;
      delete st_fes_;
      st_fes_ = new ParFiniteElementSpace(static_cast<int>(N), neq_);
      st_U_ = new ParGridFunction(st_fes_, neq_);
      st_Up_ = new ParGridFunction(st_fes_, neq_);
      st_gradUp_ = new ParGridFunction(st_fes_, neq_ * dim_);
      st_ = new SourceTerm(dim_, neq_, st_order_, st_rule_, nullptr, st_fes_, st_U_, st_Up_, st_gradUp_, pre_, cfg_, mix_, mix_,
                           trans_, chem_, rad_, nullptr, nullptr);  // source_term.cpp:36
      st_N_ = N;
    }
    std::memcpy(st_U_->GetData(), Usol, sizeof(double) * N * neq_);
    std::memcpy(st_Up_->GetData(), Up, sizeof(double) * N * neq_);
    std::memcpy(st_gradUp_->GetData(), gradUp, sizeof(double) * N * neq_ * dim_);
    Vector in(y, static_cast<int>(N) * neq_);
    st_->updateTerms(in);  // source_term.cpp:62: in += S, node by node
    return true;
  }
  bool mixture_average_diffusivity(const double *U, double *D) override {
    GasMinimalTransport *g = dynamic_cast<GasMinimalTransport *>(trans_);
    if (!g) return false;
    double E[gpudata::MAXDIM] = {0, 0, 0};
    g->computeMixtureAverageDiffusivity(U, E, D, false);  // gas_transport.cpp, as utils/binary_mixture_ic.cpp:143 calls it
    return true;
  }
  void source_term(double *Un, double *upn, const double *gradUpn, int n, double *srcTerm) override {
    const int _num_equation = neq_, _nvel = nvel_, _dim = dim_;
    const int _numSpecies = mix_->GetNumSpecies(), _numActiveSpecies = mix_->GetNumActiveSpecies();
    const int _numReactions = chem_ ? chem_->getNumReactions() : 0;
    for (int sp = 0; sp < _numActiveSpecies; sp++) {
      int eq = 3 + 2 + sp;
      if (eq >= _num_equation) continue;  // the reference indexes its MAXEQUATIONS-sized scratch; harmless there
      upn[eq] = max(upn[eq], 0.0);
      Un[eq] = max(Un[eq], 0.0);
    }
    double Efield[gpudata::MAXDIM];
    for (int v = 0; v < _nvel; v++) Efield[v] = 0.0;
    double globalTransport[gpudata::MAXSPECIES];
    double speciesTransport[gpudata::MAXSPECIES * SpeciesTrns::NUM_SPECIES_COEFFS];
    double diffusionVelocity[gpudata::MAXSPECIES * gpudata::MAXDIM];
    for (int v = 0; v < _nvel; v++)
      for (int sp = 0; sp < _numSpecies; sp++) diffusionVelocity[sp + v * _numSpecies] = 0.0;
    double ns[gpudata::MAXSPECIES];
    trans_->ComputeSourceTransportProperties(Un, upn, gradUpn, Efield, 0.0, globalTransport, speciesTransport,
                                             diffusionVelocity, ns);
    for (int eq = 0; eq < _num_equation; eq++) srcTerm[eq] = 0.0;
    double Th = upn[1 + _nvel], Te = mix_->IsTwoTemperature() ? upn[_num_equation - 1] : Th;
    double progressRates[gpudata::MAXREACTIONS], creationRates[gpudata::MAXSPECIES], emissionRates[gpudata::MAXSPECIES];
    for (int r = 0; r < gpudata::MAXREACTIONS; r++) progressRates[r] = 0.0;
    if (_numSpecies > 1 && _numReactions > 0) {
      double kfwd[gpudata::MAXREACTIONS], kC[gpudata::MAXREACTIONS];
      chem_->computeForwardRateCoeffs(ns, Th, Te, n, kfwd);
      chem_->computeEquilibriumConstants(Th, Te, kC);
      for (int sp = 0; sp < _numSpecies; sp++) creationRates[sp] = 0.0;
      chem_->computeProgressRate(ns, kfwd, kC, progressRates);
      chem_->computeCreationRate(progressRates, creationRates, emissionRates);
      for (int sp = 0; sp < _numActiveSpecies; sp++) srcTerm[2 + _nvel + sp] += creationRates[sp];
    }
    if (rad_) srcTerm[1 + _nvel] += rad_->computeEnergySink(Th);  // source_term.cpp:205-207
    if (mix_->IsTwoTemperature()) {
      for (int r = 0; r < _numReactions; r++)
        if (chem_->isElectronInvolvedAt(r)) srcTerm[_num_equation - 1] -= chem_->getReactionEnergy(r) * progressRates[r];
      double gradPe[gpudata::MAXDIM];
      mix_->computeElectronPressureGrad(ns[_numSpecies - 2], Te, gradUpn, gradPe);
      for (int d = 0; d < _dim; d++) srcTerm[_num_equation - 1] += gradPe[d] * upn[d + 1];
      const double me = mix_->GetGasParams(_numSpecies - 2, GasParams::SPECIES_MW);
      const double ne = ns[_numSpecies - 2];
      for (int sp = 0; sp < _numSpecies; sp++) {
        if (sp == _numSpecies - 2) continue;
        double m_sp = mix_->GetGasParams(sp, GasParams::SPECIES_MW);
        double energy = 1.5 * UNIVERSALGASCONSTANT * (Te - Th);
        energy *= 2.0 * me * m_sp / (m_sp + me) / (m_sp + me) * ne *
                  speciesTransport[sp + SpeciesTrns::MF_FREQUENCY * _numSpecies];
        srcTerm[_num_equation - 1] -= energy;
      }
    }
  }
};

Physics *make_physics(const OrcPhysParams &p, int dim, int nvel, int neq) {
  if (p.fluid == 1 && p.plasma) return new MixtureRef(p, dim, nvel, neq);
  if (p.fluid == 2 && p.lte) return new DryAirRef(p, dim, nvel, neq);
  if (p.fluid != 0) return nullptr;
  return new DryAirRef(p, dim, nvel, neq);
}

}  // namespace orc
