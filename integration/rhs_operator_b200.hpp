// Reference-side binding of libtpsb200.so: the subclass a pecos/tps maintainer adds as src/rhs_operator_b200.hpp.
// It fills the POD blocks of include/tpsb200.h from objects TPS already owns (ParMesh, ParFiniteElementSpace,
// RunConfiguration) and forwards RHSoperator::Mult (src/rhs_operator.hpp:157) to the device library.  Everything it reads
// is named as in the reference: src/run_configuration.hpp:90-322, src/dataStructures.hpp:205-730, MFEM's ParMesh.
// tests/test_cpu_integration_binding.py compiles this file against integration/stub/ (declarations only: MFEM and the
// TPS build tree are absent in this repository's image), so it cannot rot into pseudo-code.
#ifndef RHS_OPERATOR_B200_HPP_
#define RHS_OPERATOR_B200_HPP_

#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

#include "rhs_operator.hpp"
#include "run_configuration.hpp"
#include "tpsb200.h"

class RHSoperatorB200 : public RHSoperator {
  tpsb_ctx *ctx_ = nullptr;
  const mfem::ParGridFunction *U_view_ = nullptr;      // the solution grid function the forcing terms read
  const mfem::ParGridFunction *distance_view_ = nullptr;
  mutable bool distance_sent_ = false;
  const double &dt_;                                  // M2ulPhyS::dt, held by reference like BoundaryCondition::dt

  std::vector<double> vx_;                            // [(NE + NEH)][2^dim][dim]
  std::vector<int> el1_, el2_, inf1_, inf2_, attr_;   // MFEM face tables
  std::vector<int> nbr_rank_, send_off_, send_el_, recv_off_;
  std::vector<tpsb_bc_desc> bc_;
  tpsb_plasma_models pm_;
  std::vector<std::vector<double>> tables_;           // host copies of the tabulated rates / NEC (kept until create returns)

  static void fail(const char *what, tpsb_ctx *c) { mfem::mfem_error((std::string(what) + ": " + tpsb_last_error(c)).c_str()); }

  // (a) vertices of every local and face-neighbour element, read through the element transformation so that periodic
  //     meshes are un-wrapped exactly like the L2 coordinates TPS uses (src/M2ulPhyS.cpp:472-481)
  void fill_vertices(mfem::ParMesh *mesh) {
    const int dim = mesh->Dimension(), NE = mesh->GetNE(), NEH = mesh->GetNFaceNeighborElements(), nv = 1 << dim;
    vx_.resize(size_t(NE + NEH) * nv * dim);
    for (int e = 0; e < NE + NEH; e++) {
      mfem::ElementTransformation *T =
          e < NE ? mesh->GetElementTransformation(e) : mesh->GetFaceNbrElementTransformation(e - NE);
      const mfem::DenseMatrix &pm = T->GetPointMat();  // dim x (number of nodes); linear meshes: the 2^dim vertices
      for (int v = 0; v < nv; v++)
        for (int d = 0; d < dim; d++) vx_[(size_t(e) * nv + v) * dim + d] = pm(d, v);
    }
  }

  // (b) face tables in MFEM numbering (what initIndirectionArrays reads, src/M2ulPhyS.cpp:937-1075, 1269-1420).  Shared
  //     faces: ParMesh keeps Elem2No < 0 for them; GetSharedFaceTransformations returns Elem2No = NE + neighbour index and
  //     decodes the same Elem2Inf that GetFaceInfos returns after ExchangeFaceNbrData().
  void fill_faces(mfem::ParMesh *mesh) {
    const int NF = mesh->GetNumFaces();
    el1_.resize(NF), el2_.resize(NF), inf1_.resize(NF), inf2_.resize(NF), attr_.assign(NF, 0);
    for (int f = 0; f < NF; f++) {
      mesh->GetFaceElements(f, &el1_[f], &el2_[f]);
      mesh->GetFaceInfos(f, &inf1_[f], &inf2_[f]);
      if (el2_[f] < 0) el2_[f] = -1, inf2_[f] = inf2_[f] >= 0 ? inf2_[f] : -1;
    }
    for (int s = 0; s < mesh->GetNSharedFaces(); s++) {
      const int f = mesh->GetSharedFace(s);
      mfem::FaceElementTransformations *tr = mesh->GetSharedFaceTransformations(s, true);
      int i1, i2;
      mesh->GetFaceInfos(f, &i1, &i2);
      el2_[f] = tr->Elem2No;  // = NE + face-neighbour element index
      inf2_[f] = i2;          // 64 * local face of the neighbour + orientation
    }
    for (int b = 0; b < mesh->GetNBE(); b++) attr_[mesh->GetBdrElementFaceIndex(b)] = mesh->GetBdrAttribute(b);
  }

  // (c) partition neighbours: the tables ParMesh::ExchangeFaceNbrData built (what ParFiniteElementSpace::
  //     ExchangeFaceNbrData and src/rhs_operator.cpp:716-831 exchange with MPI)
  void fill_halo(mfem::ParMesh *mesh) {
    const int nn = mesh->GetNFaceNeighbors();
    nbr_rank_.resize(nn), send_off_.resize(nn + 1), recv_off_.resize(nn + 1);
    const int *I = mesh->send_face_nbr_elements.GetI(), *J = mesh->send_face_nbr_elements.GetJ();
    for (int k = 0; k < nn; k++) nbr_rank_[k] = mesh->GetFaceNbrRank(k);
    for (int k = 0; k <= nn; k++) send_off_[k] = I[k], recv_off_[k] = mesh->face_nbr_elements_offset[k];
    send_el_.assign(J, J + I[nn]);
  }

  // tangent1 of a non-reflecting patch exactly as InletBC / OutletBC's constructors derive it (src/outletBC.cpp:88-176): the
  // unit vector from the first to the second face quadrature point of the patch's first boundary element
  static void patch_tangent(mfem::ParMesh *mesh, mfem::ParFiniteElementSpace *vfes, mfem::IntegrationRules *intRules, int attr,
                            double *t) {
    t[0] = t[1] = t[2] = 0.;
    const int dim = mesh->Dimension();
    for (int bel = 0; bel < mesh->GetNBE(); bel++) {
      if (mesh->GetBdrAttribute(bel) != attr) continue;
      mfem::FaceElementTransformations *Tr = mesh->GetBdrFaceTransformations(bel);
      const int intorder = Tr->Elem1->OrderW() + 2 * vfes->GetFE(Tr->Elem1No)->GetOrder();
      const mfem::IntegrationRule &ir = intRules->Get(Tr->GetGeometryType(), intorder);
      double x[2][3] = {{0, 0, 0}, {0, 0, 0}};
      for (int i = 0; i < 2; i++) {
        mfem::Vector xi(x[i], 3);
        Tr->Transform(ir.IntPoint(i), xi);
      }
      double m = 0.;
      for (int d = 0; d < dim; d++) m += (x[1][d] - x[0][d]) * (x[1][d] - x[0][d]);
      for (int d = 0; d < dim; d++) t[d] = (x[1][d] - x[0][d]) / std::sqrt(m);
      return;
    }
  }
  static bool is_nr(int kind, int type) {
    return (kind == TPSB_BC_INLET && (type == SUB_DENS_VEL_NR || type == SUB_VEL_CONST_ENT)) ||
           (kind == TPSB_BC_OUTLET && (type == SUB_P_NR || type == SUB_MF_NR || type == SUB_MF_NR_PW));
  }

  // (d) BCintegrator's attribute maps (src/BCintegrator.cpp:64-125): inlets, outlets, walls
  void fill_bcs(RunConfiguration &config, int num_species_active) {
    const auto *inl = config.GetInletPatchType();
    for (size_t i = 0; i < inl->size(); i++) {
      tpsb_bc_desc b{};
      b.attr = (*inl)[i].first, b.kind = TPSB_BC_INLET, b.type = int((*inl)[i].second);
      mfem::Array<double> data = config.GetInletData(int(i));  // {rho, u, v, w, rho Y_sp of the active species} src/M2ulPhyS.cpp:3609-3641
      for (int k = 0; k < data.Size() && k < TPSB_BC_NDATA; k++) b.data[k] = data[k];
      bc_.push_back(b);
    }
    const auto *out = config.GetOutletPatchType();
    for (size_t i = 0; i < out->size(); i++) {
      tpsb_bc_desc b{};
      b.attr = (*out)[i].first, b.kind = TPSB_BC_OUTLET, b.type = int((*out)[i].second);
      mfem::Array<double> data = config.GetOutletData(int(i));  // SUB_P: {p}
      for (int k = 0; k < data.Size() && k < TPSB_BC_NDATA; k++) b.data[k] = data[k];
      bc_.push_back(b);
    }
    auto *wal = config.GetWallPatchType();
    for (size_t i = 0; i < wal->size(); i++) {
      tpsb_bc_desc b{};
      b.attr = (*wal)[i].first, b.kind = TPSB_BC_WALL, b.type = int((*wal)[i].second);
      const WallData w = config.GetWallData(int(i));
      if ((*wal)[i].second == VISC_GNRL) {  // src/wallBC.cpp:112-147
        b.data[0] = double(w.hvyThermalCond), b.data[1] = double(w.elecThermalCond), b.data[2] = w.Th, b.data[3] = w.Te;
      } else {
        b.data[0] = w.Th;                   // VISC_ISOTH wall temperature; unused by INV / SLIP / VISC_ADIAB
      }
      bc_.push_back(b);
    }
    (void)num_species_active;
  }
  // non-reflecting / mass-flow conditions: flow/refLength and the patch tangent ride in data[8..11]
  void fill_nr_data(RunConfiguration &config, mfem::ParMesh *mesh, mfem::ParFiniteElementSpace *vfes,
                    mfem::IntegrationRules *intRules) {
    for (tpsb_bc_desc &b : bc_)
      if (is_nr(b.kind, b.type)) {
        b.data[8] = config.GetReferenceLength();
        patch_tangent(mesh, vfes, intRules, b.attr, &b.data[9]);
      }
  }

  // (e) plasma models of a USER_DEFINED working fluid: PerfectMixtureInput, constantTransportData / GasTransportInput,
  //     ChemistryInput and RadiationInput flattened (src/dataStructures.hpp:537-546, 624-729); species are already in
  //     mixture order there (src/M2ulPhyS.cpp:2979-3137)
  void fill_plasma(RunConfiguration &config) {
    std::memset(&pm_, 0, sizeof(pm_));
    const PerfectMixtureInput &mi = config.perfectMixtureInput;
    const int ns = mi.numSpecies;
    pm_.num_species = ns, pm_.ambipolar = mi.ambipolar, pm_.two_temperature = mi.twoTemperature;
    for (int sp = 0; sp < ns; sp++) {
      pm_.mw[sp] = mi.gasParams[sp + SPECIES_MW * ns];
      pm_.charge[sp] = mi.gasParams[sp + SPECIES_CHARGES * ns];
      pm_.formation_energy[sp] = mi.gasParams[sp + FORMATION_ENERGY * ns];
      pm_.molar_cv[sp] = mi.molarCV[sp];
    }
    const TransportModel tm = config.GetTranportModel();
    pm_.transport_model = tm == NITROGEN_MIXTURE ? int(ARGON_MIXTURE) : int(tm);  // both run GasMixtureTransport
    const constantTransportData &ct = config.constantTransport;
    pm_.viscosity = ct.viscosity, pm_.bulk_viscosity = ct.bulkViscosity;
    pm_.thermal_conductivity = ct.thermalConductivity, pm_.electron_thermal_conductivity = ct.electronThermalConductivity;
    for (int sp = 0; sp < ns; sp++) pm_.diffusivity[sp] = ct.diffusivity[sp], pm_.mt_freq[sp] = ct.mtFreq[sp];
    const GasTransportInput &gi = config.gasTransportInput;
    pm_.third_order_k_electron = gi.thirdOrderkElectron, pm_.multiply = gi.multiply;
    for (int k = 0; k < 4; k++) pm_.flux_trns_multiplier[k] = gi.fluxTrnsMultiplier[k];
    pm_.mf_freq_multiplier = gi.spcsTrnsMultiplier[0], pm_.diff_mult = gi.diffMult, pm_.mobil_mult = gi.mobilMult;
    for (int i = 0; i < ns; i++)
      for (int j = 0; j < ns; j++) pm_.collision_index[i + j * ns] = int(gi.collisionIndex[i + j * ns]);
    pm_.ion_index = gi.ionIndex, pm_.neutral_index = gi.neutralIndex;

    const ChemistryInput &ci = config.chemistryInput;
    pm_.num_reactions = ci.numReactions, pm_.min_temperature = ci.minimumTemperature;
    size_t rxn_param_idx = 0;  // rxnModelParamsHost holds one entry per Arrhenius / Hoffert-Lien reaction (src/M2ulPhyS.cpp:3470-3474)
    for (int r = 0; r < ci.numReactions; r++) {
      pm_.model[r] = int(ci.reactionModels[r]), pm_.detailed_balance[r] = ci.detailedBalance[r];
      pm_.reaction_energy[r] = ci.reactionEnergies[r];
      for (int k = 0; k < 3; k++) pm_.equilibrium_params[r][k] = ci.equilibriumConstantParams[k + r * gpudata::MAXCHEMPARAMS];
      for (int sp = 0; sp < ns; sp++) {
        pm_.reactant_stoich[r][sp] = ci.reactantStoich[sp + r * ns];
        pm_.product_stoich[r][sp] = ci.productStoich[sp + r * ns];
      }
      const ReactionInput &ri = ci.reactionInputs[r];
      if (ci.reactionModels[r] == ARRHENIUS || ci.reactionModels[r] == HOFFERTLIEN) {
        const double *p = config.rxnModelParamsHost[rxn_param_idx++].HostRead();  // host {A, b, E}; modelParams may be a device pointer
        for (int k = 0; k < 3; k++) pm_.rate_params[r][k] = p[k];
      } else if (ci.reactionModels[r] == TABULATED_RXN) {
        const TableInput &ti = ri.tableInput;  // host pointers into config.tableHost on a CPU build; copied here
        tables_.emplace_back(ti.xdata, ti.xdata + ti.Ndata), pm_.table_x[r] = tables_.back().data();
        tables_.emplace_back(ti.fdata, ti.fdata + ti.Ndata), pm_.table_f[r] = tables_.back().data();
        pm_.table_n[r] = ti.Ndata, pm_.table_xlog[r] = ti.xLogScale, pm_.table_flog[r] = ti.fLogScale;
      } else if (ci.reactionModels[r] == GRIDFUNCTION_RXN) {
        pm_.rate_component[r] = ri.indexInput;
      }
    }
    const RadiationInput &rad = config.radiationInput;
    if (rad.model == NET_EMISSION) {
      const TableInput &ti = rad.necTableInput;
      tables_.emplace_back(ti.xdata, ti.xdata + ti.Ndata), pm_.nec_table_x = tables_.back().data();
      tables_.emplace_back(ti.fdata, ti.fdata + ti.Ndata), pm_.nec_table_f = tables_.back().data();
      pm_.nec_table_n = ti.Ndata, pm_.nec_table_xlog = ti.xLogScale, pm_.nec_table_flog = ti.fLogScale;
    }
  }

  // (f) forcing terms in the order RHSoperator's constructor registers them (src/rhs_operator.cpp:101-167); SourceTerm and
  //     AxisymmetricSource are part of tpsb_rhs_mult itself
  void add_forcings(RunConfiguration &config, const mfem::ParGridFunction *joule_heating) {
    if (config.thereIsForcing()) {
      tpsb_forcing_desc d{};
      d.kind = TPSB_FORCING_PRESSURE_GRADIENT;
      for (int k = 0; k < 3; k++) d.pressure_grad[k] = config.GetImposedPressureGradient()[k];
      if (tpsb_add_forcing(ctx_, &d)) fail("pressure gradient", ctx_);
    }
    for (int sz = 0; sz < config.numSpongeRegions_; sz++) {
      const SpongeZoneData &s = config.GetSpongeZoneData(sz);
      tpsb_forcing_desc d{};
      d.kind = TPSB_FORCING_SPONGE_ZONE;
      d.sz_type = s.szType == ANNULUS, d.sz_mixed_out = s.szSolType == MIXEDOUT;
      for (int k = 0; k < 3; k++) d.sz_normal[k] = s.normal[k], d.sz_point0[k] = s.point0[k], d.sz_point_init[k] = s.pointInit[k];
      d.sz_r1 = s.r1, d.sz_r2 = s.r2, d.sz_tol = s.tol, d.sz_mult = s.multFactor;
      if (s.szSolType == USERDEF)
        for (int k = 0; k < 5; k++) d.sz_target[k] = s.targetUp[k];
      if (tpsb_add_forcing(ctx_, &d)) fail("sponge zone", ctx_);
    }
    for (int s = 0; s < config.numHeatSources; s++) {
      const heatSourceData &h = config.heatSource[s];
      if (!h.isEnabled) continue;
      tpsb_forcing_desc d{};
      d.kind = TPSB_FORCING_HEAT_SOURCE;
      for (int k = 0; k < 3; k++) d.hs_point1[k] = h.point1[k], d.hs_point2[k] = h.point2[k];
      d.hs_radius = h.radius, d.hs_value = h.value;
      if (tpsb_add_forcing(ctx_, &d)) fail("heat source", ctx_);
    }
    if (joule_heating) {
      tpsb_forcing_desc d{};
      d.kind = TPSB_FORCING_JOULE_HEATING;
      d.joule_heating = joule_heating->Read();  // device pointer, kept by reference like the reference's grid function
      if (tpsb_add_forcing(ctx_, &d)) fail("joule heating", ctx_);
    }
  }

 public:
  // lte_tables: NULL unless the working fluid is LTE_FLUID.
  // base_args: the arguments M2ulPhyS::initVariables passes to RHSoperator today (src/M2ulPhyS.cpp:692-720); the base
  // class keeps owning what the rest of TPS reads from it.  nccl_comm: created once per job with
  // tpsb_comm_get_unique_id (rank 0) + MPI_Bcast of the 128 bytes + tpsb_comm_init_rank; NULL in a serial run.
  template <class... A>
  RHSoperatorB200(mfem::ParMesh *mesh, mfem::ParFiniteElementSpace *vfes, mfem::IntegrationRules *intRules,
                  RunConfiguration &config, const double &dt, const mfem::ParGridFunction *U,
                  const mfem::ParGridFunction *distance, const mfem::ParGridFunction *joule_heating,
                  const tpsb_lte_tables *lte_tables, void *nccl_comm, void *cuda_stream, A &&...base_args)
      : RHSoperator(std::forward<A>(base_args)...), U_view_(U), distance_view_(distance), dt_(dt) {
    const int dim = mesh->Dimension(), NE = mesh->GetNE(), NEH = mesh->GetNFaceNeighborElements();
    fill_vertices(mesh);
    fill_faces(mesh);
    tpsb_mesh_maps maps{dim,         NE,          NEH,          vx_.data(),  mesh->GetNumFaces(),
                        el1_.data(), el2_.data(), inf1_.data(), inf2_.data(), attr_.data()};
    tpsb_space_desc space{config.GetSolutionOrder(), config.GetBasisType(), config.GetIntegrationRule(), vfes->GetVDim(),
                          config.isAxisymmetric() ? 3 : dim};

    tpsb_physics phys{};
    phys.eq_system = int(config.GetEquationSystem()), phys.fluid = int(config.GetWorkingFluid());
    phys.specific_heat_ratio = config.dryAirInput.specific_heat_ratio, phys.gas_constant = config.dryAirInput.gas_constant;
    phys.visc_mult = config.GetViscMult(), phys.bulk_visc_mult = config.GetBulkViscMult();
    phys.sutherland_C1 = config.sutherland_.C1, phys.sutherland_S0 = config.sutherland_.S0;
    phys.sutherland_Pr = config.sutherland_.Pr;
    phys.use_roe = config.RoeRiemannSolverTPS();
    phys.sgs_model = config.GetSgsModelType(), phys.sgs_const = config.GetSgsConstant(), phys.sgs_floor = config.GetSgsFloor();
    const linearlyVaryingVisc &lv = config.GetLinearVaryingData();  // Fluxes' planar viscous sponge, src/fluxes.cpp:57-95
    phys.sponge_enabled = lv.isEnabled;
    if (lv.isEnabled) {
      for (int k = 0; k < dim; k++) phys.sponge_normal[k] = lv.normal[k], phys.sponge_point[k] = lv.point0[k];
      phys.sponge_ratio = lv.viscRatio, phys.sponge_width = lv.width;
    }
    phys.use_mixing_length = config.use_mixing_length;  // MixingLengthTransport, src/M2ulPhyS.cpp:265-283
    phys.max_mixing_length = config.mix_length_trans_input_.max_mixing_length_;
    phys.mixing_length_Prt = config.mix_length_trans_input_.Prt_;
    phys.mixing_length_bulk_mult = config.mix_length_trans_input_.bulk_multiplier_;
    if (config.GetWorkingFluid() == USER_DEFINED) {
      fill_plasma(config);
      phys.plasma = &pm_;
    }
    // LTE_FLUID, flow/lte/table_dim = 1: the host TableInputs M2ulPhyS::initMixtureAndTransportModels reads from the thermo
    // and transport files (thermo_tables[0..2] = energy, R, c over T; trans_tables[0..2] = mu, kappa, sigma;
    // src/M2ulPhyS.cpp:178-250) handed over as plain arrays, plus radiationInput.necTableInput when NET_EMISSION is on
    if (config.GetWorkingFluid() == LTE_FLUID) phys.lte = lte_tables;

    fill_bcs(config, config.GetNumSpecies());
    fill_nr_data(config, mesh, vfes, intRules);
    tpsb_bc_set bcs{int(bc_.size()), bc_.data(), config.useBCinGrad};
    tpsb_halo_desc halo{};
    if (NEH > 0) {
      fill_halo(mesh);
      halo = tpsb_halo_desc{int(nbr_rank_.size()), nbr_rank_.data(), send_off_.data(), send_el_.data(), recv_off_.data(), nccl_comm};
    }
    if (tpsb_create(&maps, &space, &phys, bc_.empty() ? nullptr : &bcs, NEH > 0 ? &halo : nullptr, mfem::Device::GetId(),
                    cuda_stream, &ctx_))
      fail("tpsb_create", nullptr);
    tables_.clear();  // copied at create
    add_forcings(config, joule_heating);
  }
  ~RHSoperatorB200() override { tpsb_destroy(ctx_); }
  RHSoperatorB200(const RHSoperatorB200 &) = delete;
  RHSoperatorB200 &operator=(const RHSoperatorB200 &) = delete;

  tpsb_ctx *context() const { return ctx_; }

  // src/rhs_operator.hpp:157 -- what MFEM's ODESolver::Step calls (src/M2ulPhyS.cpp:753, 2005)
  void Mult(const mfem::Vector &x, mfem::Vector &y) const override {
    // forcing terms read the solution grid function, not the RK stage vector (src/source_term.cpp:66,
    // src/forcing_terms.cpp:260)
    if (U_view_) tpsb_set_solution_view(ctx_, U_view_->Read());
    if (distance_view_ && !distance_sent_) {  // once, after the distance solver ran (src/M2ulPhyS.cpp:265-283)
      if (tpsb_set_distance_field(ctx_, distance_view_->Read())) fail("distance field", ctx_);
      distance_sent_ = true;
    }
    tpsb_set_time_step(ctx_, dt_);  // the non-reflecting conditions advance their boundary states with it
    if (tpsb_rhs_mult(ctx_, x.Read(), y.Write())) fail("tpsb_rhs_mult", ctx_);
  }

  // RHSoperator::getMaxCharSpeed equivalent for M2ulPhyS::solveStep's adaptive dt (src/M2ulPhyS.cpp:2013-2016); collective
  double maxCharSpeed() const {
    double v = 0.;
    if (tpsb_get_max_char_speed(ctx_, &v)) fail("max char speed", ctx_);
    return v;
  }
  // Averaging::addSample (src/averaging.cpp:198-420) on the device; the caller keeps ns_mean / ns_vari like the reference
  void addSample(const mfem::Vector &inst, int num_fields, mfem::Vector &mean, mfem::Vector *vari, int vari_start,
                 int vari_components, int ns_mean, int ns_vari, int pressure_slot) const {
    if (tpsb_averaging_add_sample(ctx_, inst.Read(), num_fields, mean.ReadWrite(), vari ? vari->ReadWrite() : nullptr,
                                  vari_start, vari_components, ns_mean, ns_vari, pressure_slot))
      fail("averaging", ctx_);
  }
};

#endif  // RHS_OPERATOR_B200_HPP_
