/* TEST INFRASTRUCTURE ONLY.  Force-included (-include) ahead of every reference
 * translation unit built into oracle/_ref/.  The include guards of the three heavy
 * reference headers are pre-defined on the command line
 * (-DRUN_CONFIGURATION_HPP_ -DUTILS_HPP_ -DTPS_MFEM_WRAP_HPP_, cf.
 * src/run_configuration.hpp:33, src/utils.hpp:32, src/tps_mfem_wrap.hpp:32) so their
 * bodies are skipped; this file supplies the few names the leaf physics sources
 * still expect from them. */
#pragma once
#include "mfem.hpp"
#include "grvy.h"
#include "tps_config.h"
using namespace mfem;
#include "dataStructures.hpp"

#ifndef ERROR
#define ERROR 1
#endif

/* POD stand-in for the reference's RunConfiguration (src/run_configuration.hpp:55-420):
 * exposes only the members/accessors the leaf constructors read. */
class RunConfiguration {
 public:
  WorkingFluid workFluid = DRY_AIR;
  double const_plasma_conductivity_ = 0.0;
  DryAirInput dryAirInput;
  PerfectMixtureInput perfectMixtureInput;
  GasTransportInput gasTransportInput;
  constantTransportData constantTransport;
  ChemistryInput chemistryInput;
  RadiationInput radiationInput;
  LteMixtureInput lteMixtureInput;
  mixingLengthTransportData mix_length_trans_input_;
  SutherlandData sutherland_;
  linearlyVaryingVisc linViscData;
  double visc_mult = 1.0;
  double bulk_visc = 0.0;
  int sgsModelType = 0;
  double sgsFloor = 0.0;
  double sgs_model_const = 0.0;
  bool axisymmetric_ = false;

  WorkingFluid GetWorkingFluid() { return workFluid; }
  double GetViscMult() { return visc_mult; }
  double GetBulkViscMult() { return bulk_visc; }
  int GetSgsModelType() { return sgsModelType; }
  double GetSgsFloor() { return sgsFloor; }
  double GetSgsConstant() { return sgs_model_const; }
  linearlyVaryingVisc &GetLinearVaryingData() { return linViscData; }
  bool isAxisymmetric() const { return axisymmetric_; }
};
