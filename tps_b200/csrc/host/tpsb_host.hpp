// Host-side C++ mirror of the reference's operator interface for the RHS path, written above the C ABI
// (include/tpsb200.h).  Class and method names, argument meaning and call order follow
// src/rhs_operator.hpp:146-184 (RHSoperator), MFEM's ODESolver (Init/Step, as used at
// src/M2ulPhyS.cpp:753,2005) and utils/compute_rhs.cpp:102, so TPS's M2ulPhyS can swap its RHSoperator for
// this one without touching the Runge-Kutta loop.  In a TPS build `Vector` is
// mfem::Vector with device memory (Read()/Write() return device pointers); here, without MFEM, a minimal
// device-backed Vector with the same accessors stands in.
//
// Error behaviour: like the reference (mfem_error / exit(ERROR), src/rhs_operator.cpp) a failed call is
// fatal for the solver -- it is reported through tpsb_last_error and raised as std::runtime_error.
#pragma once
#include <cuda_runtime_api.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/tpsb200.h"

namespace tpsb_host {

#ifndef TPSB_HAVE_MFEM
// Device-backed vector with mfem::Vector's accessor names.
class Vector {
  double *d_ = nullptr;
  int64_t n_ = 0;

 public:
  Vector() {}
  explicit Vector(int64_t n) { SetSize(n); }
  Vector(const Vector &) = delete;
  Vector &operator=(const Vector &) = delete;
  ~Vector() {
    if (d_) cudaFree(d_);
  }
  void SetSize(int64_t n) {
    if (d_) cudaFree(d_);
    d_ = nullptr;
    n_ = n;
    if (n > 0 && cudaMalloc(reinterpret_cast<void **>(&d_), sizeof(double) * n) != cudaSuccess) throw std::runtime_error("Vector: cudaMalloc failed");
  }
  int64_t Size() const { return n_; }
  void UseDevice(bool) {}
  const double *Read() const { return d_; }
  double *Write() { return d_; }
  double *ReadWrite() { return d_; }
  // host mirrors (explicit copies)
  void SetFromHost(const std::vector<double> &h) {
    if (static_cast<int64_t>(h.size()) != n_) SetSize(static_cast<int64_t>(h.size()));
    cudaMemcpy(d_, h.data(), sizeof(double) * n_, cudaMemcpyHostToDevice);
  }
  std::vector<double> HostCopy() const {
    std::vector<double> h(n_);
    cudaMemcpy(h.data(), d_, sizeof(double) * n_, cudaMemcpyDeviceToHost);
    return h;
  }
};
#else
using Vector = mfem::Vector;
#endif

// RHSoperator : public TimeDependentOperator   (src/rhs_operator.hpp:52-184)
class RHSoperator {
  tpsb_ctx *ctx_ = nullptr;
  mutable double time_ = 0.0;

  void check(int rc, const char *what) const {
    if (rc != TPSB_OK) throw std::runtime_error(std::string(what) + ": " + tpsb_last_error(ctx_));
  }

 public:
  // The reference constructor takes the FE spaces, integration rules, Fluxes, GasMixture, ... objects
  // (src/rhs_operator.cpp:38-48); everything the device path needs from them is in these POD blocks.
  RHSoperator(const tpsb_mesh_maps &maps, const tpsb_space_desc &space, const tpsb_physics &phys,
              const tpsb_bc_set *bcs = nullptr, const tpsb_halo_desc *halo = nullptr, int device = 0,
              void *cuda_stream = nullptr) {
    const int rc = tpsb_create(&maps, &space, &phys, bcs, halo, device, cuda_stream, &ctx_);
    if (rc != TPSB_OK) throw std::runtime_error(std::string("RHSoperator: ") + tpsb_last_error(nullptr));
  }
  RHSoperator(const RHSoperator &) = delete;
  virtual ~RHSoperator() { tpsb_destroy(ctx_); }

  int64_t Height() const { return tpsb_num_dofs(ctx_) * tpsb_num_equation(ctx_); }
  void SetTime(double t) const { time_ = t; }
  double GetTime() const { return time_; }

  // src/rhs_operator.hpp:157
  virtual void Mult(const Vector &x, Vector &y) const { check(tpsb_rhs_mult(ctx_, x.Read(), y.Write()), "RHSoperator::Mult"); }
  // src/rhs_operator.hpp:158-159
  void updatePrimitives(const Vector &x) const { check(tpsb_update_primitives(ctx_, x.Read()), "updatePrimitives"); }
  void updateGradients(const Vector &x, const bool &primitiveUpdated) const {
    check(tpsb_update_gradients(ctx_, x.Read(), primitiveUpdated ? 1 : 0), "updateGradients");
  }
  // device views of Up / gradUp (M2ulPhyS::getPrimitiveGF / getGradientGF)
  const double *getPrimitives() const {
    double *up = nullptr;
    check(tpsb_get_fields(ctx_, &up, nullptr), "getPrimitives");
    return up;
  }
  const double *getGradients() const {
    double *g = nullptr;
    check(tpsb_get_fields(ctx_, nullptr, &g), "getGradients");
    return g;
  }
  // max_char_speed reference member (src/rhs_operator.hpp:73), reduced on the device
  double getMaxCharSpeed() const {
    double v = 0;
    check(tpsb_get_max_char_speed(ctx_, &v), "getMaxCharSpeed");
    return v;
  }
  // the solution grid function U_ the forcing terms read (src/source_term.cpp:66) and the wall-distance grid function
  // distance_ of the mixing-length model (src/M2ulPhyS.cpp:265-283); device vectors owned by the caller
  void setSolutionView(const Vector *U) const { check(tpsb_set_solution_view(ctx_, U ? U->Read() : nullptr), "setSolutionView"); }
  void setDistance(const Vector *d) const { check(tpsb_set_distance_field(ctx_, d ? d->Read() : nullptr), "setDistance"); }
  // hmin of M2ulPhyS (src/M2ulPhyS.cpp:756-761)
  double getMinElementSize() const {
    double h = 0;
    check(tpsb_get_hmin(ctx_, &h), "getMinElementSize");
    return h;
  }
  // ForcingTerms registered by RHSoperator's constructor (src/rhs_operator.cpp:101-167): pressure gradient, sponge zones, heat
  // sources, Joule heating -- in the order of the calls
  void addForcing(const tpsb_forcing_desc &d) const { check(tpsb_add_forcing(ctx_, &d), "addForcing"); }
  void clearForcings() const { check(tpsb_clear_forcings(ctx_), "clearForcings"); }
  // BoundaryCondition::dt (a reference to M2ulPhyS::dt): the step the non-reflecting inlets / outlets advance their boundary
  // states with at every Mult (src/inletBC.cpp:703, src/outletBC.cpp:707)
  void setTimeStep(double dt) const { check(tpsb_set_time_step(ctx_, dt), "setTimeStep"); }
  // InletBC / OutletBC members meanUp and boundaryU of the non-reflecting condition on a boundary attribute
  int getBoundaryState(int attr, double *mean_up, double *boundary_u, int capacity) const {
    int n = 0;
    check(tpsb_get_bc_state(ctx_, attr, mean_up, boundary_u, capacity, &n), "getBoundaryState");
    return n;
  }
  // Averaging::addSample (src/averaging.cpp:198-420) on device vectors; the caller keeps ns_mean / ns_vari
  void addSample(const Vector &inst, int num_fields, Vector &mean, Vector *vari, int vari_start, int vari_components, int ns_mean,
                 int ns_vari, int pressure_slot) const {
    check(tpsb_averaging_add_sample(ctx_, inst.Read(), num_fields, mean.ReadWrite(), vari ? vari->ReadWrite() : nullptr, vari_start,
                                    vari_components, ns_mean, ns_vari, pressure_slot),
          "addSample");
  }
  int path() const { return tpsb_get_path(ctx_); }  // 0 general 3-D, 1 fast, 2 fused, 3 generic (include/tpsb200.h)
  tpsb_ctx *context() const { return ctx_; }
};

// mfem::ODESolver family selected at src/M2ulPhyS.cpp:721-739 (Init + Step as at :753 and :2005).
class ODESolver {
 protected:
  RHSoperator *f_ = nullptr;
  int scheme_;

 public:
  explicit ODESolver(int scheme) : scheme_(scheme) {}
  virtual ~ODESolver() {}
  virtual void Init(RHSoperator &f) { f_ = &f; }
  virtual void Step(Vector &x, double &t, double &dt) {
    const int rc = tpsb_ode_step(f_->context(), x.ReadWrite(), dt, scheme_, 1);
    if (rc != TPSB_OK) throw std::runtime_error(std::string("ODESolver::Step: ") + tpsb_last_error(f_->context()));
    t += dt;
    f_->SetTime(t);
  }
};
// M2ulPhyS::solveStep without the I/O (src/M2ulPhyS.cpp:2004-2016): Step + Check_NAN + Check_Undershoot + adaptive dt.
// cfl <= 0: constant time step.  Returns the number of NaN entries found (the reference aborts when it is non-zero).
inline int solveStep(RHSoperator &rhs, int scheme, Vector &U, double &t, double &dt, double cfl) {
  int nan = 0;
  double next = dt;
  const int rc = tpsb_solve_step(rhs.context(), U.ReadWrite(), dt, scheme, cfl, &nan, &next);
  if (rc != TPSB_OK) throw std::runtime_error(std::string("solveStep: ") + tpsb_last_error(rhs.context()));
  t += dt;
  rhs.SetTime(t);
  dt = next;
  return nan;
}
struct ForwardEulerSolver : ODESolver { ForwardEulerSolver() : ODESolver(1) {} };
struct RK2Solver : ODESolver { explicit RK2Solver(double /*a = 1.0*/ = 1.0) : ODESolver(2) {} };
struct RK3SSPSolver : ODESolver { RK3SSPSolver() : ODESolver(3) {} };
struct RK4Solver : ODESolver { RK4Solver() : ODESolver(4) {} };

}  // namespace tpsb_host
