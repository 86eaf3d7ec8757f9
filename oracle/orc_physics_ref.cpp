// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Never linked into the product.
//
// "reference" physics back end: forwards every per-point call to the reference's own,
// unmodified classes, whose sources are compiled in place from /root/reference/src
// (equation_of_state.cpp, transport_properties.cpp, fluxes.cpp, riemann_solver.cpp, ...)
// against the header stub in oracle/refstub/.  Built only into oracle/_ref/ by
// oracle/Makefile; the resulting shared object travels to the GPU box, the sources do not.
#include "equation_of_state.hpp"
#include "fluxes.hpp"
#include "riemann_solver.hpp"
#include "transport_properties.hpp"

#include "orc_physics.hpp"

namespace orc {

class DryAirRef : public Physics {
  int dim_, nvel_, neq_;
  DryAir *mix_;
  DryAirTransport *trans_;
  Fluxes *flux_;
  RiemannSolverTPS *rs_;

 public:
  DryAirRef(const OrcPhysParams &p, int dim, int nvel, int neq) : dim_(dim), nvel_(nvel), neq_(neq) {
    DryAirInput in;
    in.f = DRY_AIR;
    in.eq_sys = static_cast<Equations>(p.eq_system);
    in.specific_heat_ratio = p.gamma;
    in.gas_constant = p.R;
    mix_ = new DryAir(in, dim, nvel);                                                          // equation_of_state.cpp:150
    trans_ = new DryAirTransport(mix_, p.visc_mult, p.bulk_visc_mult, p.C1, p.S0, p.Pr);       // transport_properties.cpp:208
    flux_ = new Fluxes(mix_, static_cast<Equations>(p.eq_system), trans_, neq, dim, false);   // fluxes.cpp:34
    rs_ = new RiemannSolverTPS(neq, mix_, static_cast<Equations>(p.eq_system), flux_, false, false);  // riemann_solver.cpp:38
  }
  ~DryAirRef() {
    delete rs_;
    delete flux_;
    delete trans_;
    delete mix_;
  }
  const char *kind() const override { return "reference"; }
  int num_active_species() const override { return mix_->GetNumActiveSpecies(); }
  int num_species() const override { return mix_->GetNumSpecies(); }
  double pressure(const double *U) override { return mix_->ComputePressure(U); }
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
  void prim(const double *U, double *Up) override { mix_->GetPrimitivesFromConservatives(U, Up); }
  void cons(const double *Up, double *U) override { mix_->GetConservativesFromPrimitives(Up, U); }
  double max_char_speed(const double *U) override { return mix_->ComputeMaxCharSpeed(U); }
  void conv_flux(const double *U, double *F) override { flux_->ComputeConvectiveFluxes(U, F); }
  void visc_flux(const double *U, const double *gradUp, double *xyz, double delta, double dist, double *F) override {
    flux_->ComputeViscousFluxes(U, gradUp, xyz, delta, dist, F);
  }
  void riemann(const double *U1, const double *U2, const double *nor, double *flux) override {
    rs_->Eval(U1, U2, nor, flux, false);
  }
};

Physics *make_physics(const OrcPhysParams &p, int dim, int nvel, int neq) {
  if (p.fluid != 0) return nullptr;
  return new DryAirRef(p, dim, nvel, neq);
}

}  // namespace orc
