/* tpsb200.h -- C ABI of libtpsb200.so: the B200-native (sm_100a, FP64) replacement for the
 * explicit DG right-hand-side path of pecos/tps (RHSoperator::Mult and everything it calls).
 *
 * The reference has no C ABI: its "plugin API" is a set of C++ classes wired with raw pointers in
 * M2ulPhyS::initVariables (src/M2ulPhyS.cpp:599-621,692-753).  Each entry point below names the
 * reference interface it replaces; INTEGRATION.md shows the C++ shim a TPS maintainer adds so that
 * M2ulPhyS, MFEM's ODESolver and utils/compute_rhs.cpp keep calling the same class methods.
 *
 * Conventions (all from the reference):
 *   - nodal vectors are Ordering::byNODES: U[n + eq*N], gradUp[n + eq*N + d*neq*N], N = vfes->GetNDofs()
 *     = num_elems * dof, node n = e*dof + local node (lexicographic, x fastest)   (src/rhs_operator.cpp:589-591)
 *   - conserved state [rho, rho u (nvel), rho E, ...], primitives [rho, u (nvel), T, ...]
 *     (src/equation_of_state.cpp:321-335)
 *   - faces follow MFEM's Mesh::GetFaceElements / GetFaceInfos numbering:
 *     ElemXInf = 64*local_face + orientation                                     (src/M2ulPhyS.cpp:937-958)
 * Plain pointers and sizes only; no C++/torch types.  Every function returns 0 on success or a
 * TPSB_E* code (message via tpsb_last_error); nothing here calls exit()/abort().
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with TPSB_ECUDA.
 */
#ifndef TPSB200_H_
#define TPSB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPSB_OK 0
#define TPSB_EINVAL 1    /* bad argument / unsupported configuration */
#define TPSB_ECUDA 2     /* CUDA runtime error (or no device)          */
#define TPSB_ENCCL 3     /* NCCL error                                  */
#define TPSB_ENOTIMPL 4  /* valid in the reference, not built yet       */

/* ---- enums: numeric values equal the reference's (src/dataStructures.hpp:65-72) ---- */
enum { TPSB_EULER = 0, TPSB_NS = 1, TPSB_NS_PASSIVE = 2 }; /* Equations     */
enum { TPSB_DRY_AIR = 0, TPSB_USER_DEFINED = 1, TPSB_LTE_FLUID = 2 }; /* WorkingFluid */

/* Mesh/connectivity tables the MFEM host code hands over (what initIndirectionArrays reads from
 * ParMesh, src/M2ulPhyS.cpp:816-1075).  All arrays are HOST memory, copied at create.          */
typedef struct {
  int dim;                      /* 3 (hexahedra) or 2 (quadrilaterals)                                      */
  int num_elems;                /* vfes->GetNE(), local elements                                           */
  int num_nbr_elems;            /* face-neighbour (halo) elements, pmesh->GetNFaceNeighborElements(); 0 serial */
  const double *elem_vertices;  /* [(num_elems+num_nbr_elems)][2^dim][dim] vertex coordinates, MFEM vertex
                                   order, taken per element from the mesh nodes (so periodic meshes are
                                   un-wrapped exactly as MFEM's L2 nodal GridFunction is)                  */
  int num_faces;                /* mesh->GetNumFaces()                                                     */
  const int *face_el1;          /* [num_faces] Elem1No                                                     */
  const int *face_el2;          /* [num_faces] Elem2No: -1 boundary; >= num_elems: num_elems + nbr index   */
  const int *face_inf1;         /* [num_faces] Elem1Inf = 64*local_face (+0)                               */
  const int *face_inf2;         /* [num_faces] Elem2Inf = 64*local_face + orientation (-1 on boundary)     */
  const int *face_attr;         /* [num_faces] boundary attribute of boundary faces (else 0); may be NULL  */
} tpsb_mesh_maps;

/* FE space description (src/M2ulPhyS.cpp:558-579). */
typedef struct {
  int order;          /* flow/order, 1..5 (1-3: sum-factorised kernels; 4, 5: generic path) */
  int basis_type;     /* flow/basisType: 0 Gauss-Legendre nodes, 1 Gauss-Lobatto      */
  int int_rule_type;  /* flow/integrationRule: 0 Gauss-Legendre, 1 Gauss-Lobatto      */
  int num_equation;   /* vfes vdim                                                    */
  int nvel;           /* dim; 3 on a 2-D mesh = axisymmetric run (x = r, y = z, third velocity u_theta;
                         config.isAxisymmetric(), src/rhs_operator.cpp:191-205, src/forcing_terms.cpp:255-380) */
} tpsb_space_desc;

/* Physics parameter block: the POD input structs of the reference flattened
 * (DryAirInput src/dataStructures.hpp:609-622; DryAirTransport ctor src/transport_properties.cpp:208;
 *  SutherlandData :205-209).                                                                     */
/* Plasma models of a user-defined working fluid (fluid = TPSB_USER_DEFINED): the reference's POD input structs
 * PerfectMixtureInput, constantTransportData and ChemistryInput (src/dataStructures.hpp:537-546,623-633,690-712)
 * flattened.  Species are in MIXTURE order: electron second to last, background last (src/M2ulPhyS.cpp:2979-3137). */
#define TPSB_MAX_SPECIES 8
#define TPSB_MAX_REACTIONS 34
typedef struct {
  /* PerfectMixture (src/equation_of_state.cpp:478-574) */
  int num_species, ambipolar, two_temperature;
  double mw[TPSB_MAX_SPECIES];               /* GasParams::SPECIES_MW [kg/mol]                 */
  double charge[TPSB_MAX_SPECIES];           /* GasParams::SPECIES_CHARGES                     */
  double formation_energy[TPSB_MAX_SPECIES]; /* GasParams::FORMATION_ENERGY [J/mol]            */
  double molar_cv[TPSB_MAX_SPECIES];         /* perfect_mixture/constant_molar_cv, units of R  */
  /* TransportModel value (src/dataStructures.hpp:80-88): 2 constant (src/transport_properties.cpp:303-448),
   * 0 argon_minimal, 1 argon_mixture (fields at the end of this struct)                                              */
  int transport_model;
  double viscosity, bulk_viscosity, thermal_conductivity, electron_thermal_conductivity;
  double diffusivity[TPSB_MAX_SPECIES], mt_freq[TPSB_MAX_SPECIES];
  /* Chemistry (src/chemistry.cpp:40-300): reaction r uses model[r] = ReactionModel (0 Arrhenius, 1 Hoffert-Lien,
   * 2 tabulated, 3 grid function; fields at the end of this struct) */
  int num_reactions;
  double min_temperature;
  int model[TPSB_MAX_REACTIONS], detailed_balance[TPSB_MAX_REACTIONS];
  double rate_params[TPSB_MAX_REACTIONS][3];         /* A, b, E                               */
  double reaction_energy[TPSB_MAX_REACTIONS];
  double equilibrium_params[TPSB_MAX_REACTIONS][3];  /* A, b, E of K_eq = A T^b exp(-E/T)     */
  int reactant_stoich[TPSB_MAX_REACTIONS][TPSB_MAX_SPECIES], product_stoich[TPSB_MAX_REACTIONS][TPSB_MAX_SPECIES];
  /* transport_model = argon_minimal (TransportModel value 0): GasMinimalTransport, Chapman-Enskog transport of the
   * ternary mixture [Ar.+1, E, Ar] from collision integrals (src/gas_transport.cpp:42-870,
   * src/collision_integrals.cpp:53-201; GasTransportInput src/dataStructures.hpp:644-666).                          */
  int third_order_k_electron;      /* argon_transport/third_order_thermal_conductivity                              */
  int multiply;                    /* artificial multipliers on (argonMinimal.multipliers.ini)                       */
  double flux_trns_multiplier[4];  /* viscosity, bulk viscosity, heavy / electron thermal conductivity               */
  double mf_freq_multiplier, diff_mult, mobil_mult;
  /* reaction model 2 (TABULATED_RXN): rate coefficient k_f(T) by linear interpolation of a table, optionally in log
   * scale of either axis (TableInput src/dataStructures.hpp:668-676, LinearTable src/table.cpp:52-97,
   * Tabulated src/reaction.cpp:62-84).  HOST arrays, copied at create; at most 1000 points (gpudata::MAXTABLE).
   * reaction model 3 (GRIDFUNCTION_RXN): k_f per node from component rate_component[r] of the field handed to
   * tpsb_set_reaction_rate_field (GridFunctionReaction src/reaction.cpp:86-117).                                      */
  int table_n[TPSB_MAX_REACTIONS], table_xlog[TPSB_MAX_REACTIONS], table_flog[TPSB_MAX_REACTIONS];
  const double *table_x[TPSB_MAX_REACTIONS], *table_f[TPSB_MAX_REACTIONS];
  int rate_component[TPSB_MAX_REACTIONS];
  /* transport_model = argon_mixture (TransportModel value 1): GasMixtureTransport (src/gas_transport.cpp:877-1650),
   * argon or nitrogen mixtures of up to 7 species.  collision_index[spI + spJ*num_species] (spI <= spJ) = GasColl value
   * (src/dataStructures.hpp:122-143: 0 CLMB_ATT, 1 CLMB_REP, 2 AR_AR1P, 3 AR_E, 4 AR_AR; nitrogen 6 NI_NI1P, 7 NI_E,
   * 8 NI_NI, 10 N2_NI1P, 11 N2_E, 12 N2_N2, 13 N2_NI with the curve fits of src/collision_integrals.cpp:210-625), as
   * M2ulPhyS::identifyCollisionType fills GasTransportInput::collisionIndex (src/M2ulPhyS.cpp:3925-3970);
   * ion_index / neutral_index: the species 'Ar.+1' and 'Ar' (GasTransportInput::ionIndex / neutralIndex).            */
  int collision_index[TPSB_MAX_SPECIES * TPSB_MAX_SPECIES];
  int ion_index, neutral_index;
  /* Radiation model NET_EMISSION with a tabulated net emission coefficient (RadiationInput
   * src/dataStructures.hpp:724-729, NetEmission src/radiation.hpp:57-70): energy sink -4 pi eps_N(T_h) on the total
   * energy equation (src/source_term.cpp:205-207).  nec_table_n = 0: no radiation.  HOST arrays, copied at create. */
  int nec_table_n, nec_table_xlog, nec_table_flog;
  const double *nec_table_x, *nec_table_f;
} tpsb_plasma_models;

/* LTE working fluid (fluid = TPSB_LTE_FLUID) with 1-D look-up tables, flow/lte/table_dim = 1: LteMixture + LteTransport
 * (src/lte_mixture.cpp, src/lte_transport_properties.cpp) as M2ulPhyS builds them from the columns of the thermo file
 * "T_energy_R_c" and the transport file (src/M2ulPhyS.cpp:175-258): one species, one temperature, num_equation = nvel + 2;
 * T from the conserved state by Newton iteration on e(T) started from the inverse table, p = rho R(T) T, speed of sound c(T),
 * viscosity mu(T), conductivity kappa(T), no bulk viscosity; optional tabulated net emission coefficient (radiative sink
 * -4 pi eps_N(T), src/source_term.cpp:205-207).  All tables are linear in both axes (LinearTable src/table.cpp:76-116), HOST
 * arrays copied at create, 2..1000 rows (gpudata::MAXTABLE); T and energy strictly increasing.  sigma (electrical
 * conductivity) only feeds the plasma-conductivity output field of the reference, not dU/dt: accepted, not used.  The
 * reference's 2-D (T, rho) tables need GSL and are CPU-only there (src/M2ulPhyS.cpp:168-170): not built.               */
typedef struct {
  int num_thermo;
  const double *T, *energy, *R, *c;
  int num_trans;
  const double *T_trans, *mu, *kappa, *sigma;
  int nec_table_n, nec_table_xlog, nec_table_flog;
  const double *nec_table_x, *nec_table_f;
} tpsb_lte_tables;

typedef struct {
  int eq_system;          /* TPSB_EULER / TPSB_NS                              */
  int fluid;              /* TPSB_DRY_AIR, TPSB_USER_DEFINED with `plasma`, or TPSB_LTE_FLUID with `lte` */
  double specific_heat_ratio;
  double gas_constant;
  double visc_mult;       /* flow/viscosityMultiplier                          */
  double bulk_visc_mult;  /* flow/bulkViscosityMultiplier                      */
  double sutherland_C1, sutherland_S0, sutherland_Pr;
  const tpsb_plasma_models *plasma; /* fluid == TPSB_USER_DEFINED: gas / transport / chemistry models; else NULL */
  int use_roe;            /* flow/useRoe: RiemannSolverTPS::Eval_Roe (src/riemann_solver.cpp:117-206) on interior
                             faces and inviscid walls; as in the reference it is written for two velocity
                             components with gamma - 1 = 0.4 hard-coded, so only 2-D dry air accepts it           */
  /* Sub-grid-scale eddy viscosity of Fluxes::ComputeViscousFluxes / ComputeBdrViscousFluxes (src/fluxes.cpp:224-231,
   * 386-392): flow/sgsModel 0 none, 1 smagorinsky (Fluxes::sgsSmag :513-541), 2 sigma (Fluxes::sgsSigma :543-650, the
   * closed-form branch a device build runs); flow/sgsModelConstant (reference defaults 0.12 / 0.135,
   * src/M2ulPhyS.cpp:2693-2699), flow/sgsFloor.  The element size is Mesh::GetElementSize(e, 1) / order, derived
   * from the vertices at create.  3-D dry air, Gauss-Legendre path.                                               */
  int sgs_model;
  double sgs_const, sgs_floor;
  /* Planar viscous sponge (viscosityMultiplierFunction/..., Fluxes::viscSpongePlanar src/fluxes.cpp:664-684): viscosity,
   * bulk viscosity and conductivity times 1 + (max(ratio,1) - 1) (tanh(dist/width - 2) + 1)/2, dist = (x - point).n
   * with n normalised as Fluxes' constructor does (src/fluxes.cpp:73-83).                                          */
  int sponge_enabled;
  double sponge_normal[3], sponge_point[3], sponge_ratio, sponge_width;
  /* flow/useMixingLength: MixingLengthTransport wrapped around the molecular transport (src/M2ulPhyS.cpp:265-283,
   * src/mixing_length_transport.cpp:62-121): mu_t = rho l^2 |S|, l = min(0.41 d_wall, max_mixing_length);
   * mixing-length/Pr_ratio scales the eddy conductivity, mixing-length/bulk-multiplier the eddy bulk viscosity.  The wall
   * distance d_wall is the nodal field handed over with tpsb_set_distance_field (zero until then).  Runs on the
   * generic path (any fluid, 2-D / axisymmetric / 3-D).                                                            */
  int use_mixing_length;
  double max_mixing_length, mixing_length_Prt, mixing_length_bulk_mult;
  const tpsb_lte_tables *lte; /* fluid == TPSB_LTE_FLUID: the look-up tables; else NULL */
} tpsb_physics;

/* Boundary conditions: BCintegrator's attribute -> {InletBC, OutletBC, WallBC} maps (src/BCintegrator.cpp:64-125).
 * kind selects the map, type is the reference's enum value (src/dataStructures.hpp:168-196):
 *   inlet  : InletType  -- built: SUB_DENS_VEL (2), data = inputState {rho, u, v, w}      (src/inletBC.cpp:729-756)
 *   outlet : OutletType -- built: SUB_P (0),        data = {p}                             (src/outletBC.cpp:731-737)
 *   wall   : WallType   -- built: INV (0), SLIP (1), VISC_ADIAB (2), VISC_ISOTH (3, data = {Th}) (src/wallBC.cpp:277-510),
 *                          VISC_GNRL (4, generic path; data = {hvyThermalCond, elecThermalCond, Th, Te} with
 *                          ThermalCondition 0 ADIAB / 1 ISOTH / 2 SHTH sheath; src/wallBC.cpp:112-147,512-543,
 *                          PerfectMixture::computeSheathBdrFlux src/equation_of_state.cpp:1909-1942)
 *   non-reflecting / mass-flow conditions (dry air, planar 2-D or 3-D; generic path; the reference refuses them for mixtures):
 *     inlet  SUB_DENS_VEL_NR (6), SUB_VEL_CONST_ENT (7): data = {rho, u, v, w}            (src/inletBC.cpp:576-727)
 *     outlet SUB_P_NR (2): data = {p}; SUB_MF_NR (3), SUB_MF_NR_PW (4): data = {mass flow} (src/outletBC.cpp:573-1027)
 *     with data[8] = flow/refLength (> 0) and data[9..11] = the patch's unit tangent `tangent1` (all zero: derived like the
 *     reference's constructors do, from the first two quadrature points of the patch's first boundary face in face order).
 *     These conditions are STATEFUL like the reference's: each keeps a conserved boundary state per face quadrature point
 *     (`boundaryU`, initialised from the interpolated primitives by the first evaluation and advanced by EVERY evaluation
 *     with the time step of tpsb_set_time_step / tpsb_ode_step) and the patch mean of the primitives, refreshed at every
 *     evaluation (InletBC / OutletBC::updateMean, src/rhs_operator.cpp:364) and all-reduced over the ranks.
 * Other types return TPSB_ENOTIMPL at create.  use_bc_in_grad = boundaryConditions/useBCinGrad
 * (src/M2ulPhyS.cpp:3480): the BR1 gradient then uses the wall state at isothermal walls
 * (src/faceGradientIntegration.cpp:96-115) and the wall Riemann state flips (src/wallBC.cpp:476-479).  */
enum { TPSB_BC_INLET = 0, TPSB_BC_OUTLET = 1, TPSB_BC_WALL = 2 };
#define TPSB_BC_NDATA 12
typedef struct {
  int attr;        /* boundary attribute (patch number) the condition applies to */
  int kind;        /* TPSB_BC_*                                                   */
  int type;        /* InletType / OutletType / WallType value                     */
  double data[TPSB_BC_NDATA]; /* inlet: inputState {rho, u, v, w, rho Y_sp (active species, mixture order) ...}
                                 (src/M2ulPhyS.cpp:3609-3641, src/inletBC.cpp:52-66); outlet {p}; wall {Th}  */
} tpsb_bc_desc;
typedef struct {
  int num_bcs;
  const tpsb_bc_desc *bcs;
  int use_bc_in_grad;
} tpsb_bc_set;

/* Partition neighbours for the face-neighbour exchange that replaces RHSoperator::initNBlockDataTransfer /
 * waitAllDataTransfer (src/rhs_operator.cpp:716-831).  NULL / num_nbr_ranks == 0 for a serial run.  */
typedef struct {
  int num_nbr_ranks;
  const int *nbr_rank;          /* [num_nbr_ranks] peer rank                                             */
  const int *send_offset;       /* [num_nbr_ranks+1] into send_elems (== send_face_nbr_elements of MFEM) */
  const int *send_elems;        /* local element ids whose dofs are sent, grouped by peer                */
  const int *recv_offset;       /* [num_nbr_ranks+1] halo-element ranges (face_nbr_elements_offset)      */
  void *nccl_comm;              /* ncclComm_t created by the host (or by tpsb_comm_init_rank)            */
} tpsb_halo_desc;

typedef struct tpsb_ctx tpsb_ctx; /* opaque, one per rank / GPU */

/* Library / build information. */
const char *tpsb_version(void);
const char *tpsb_last_error(const tpsb_ctx *ctx); /* ctx may be NULL: error of the last failed create */

/* RHSoperator::RHSoperator + Gradients::Gradients + M2ulPhyS::initIndirectionArrays
 * (src/rhs_operator.cpp:38-322, src/gradients.cpp:36-138, src/M2ulPhyS.cpp:816-1532):
 * builds every device-side table.  cuda_stream: the caller's cudaStream_t (0 = default stream);
 * all work of this context is enqueued on / ordered against it.                                   */
int tpsb_create(const tpsb_mesh_maps *maps, const tpsb_space_desc *space, const tpsb_physics *phys,
                const tpsb_bc_set *bcs /* NULL: no boundary faces */, const tpsb_halo_desc *halo, int device,
                void *cuda_stream, tpsb_ctx **out);
void tpsb_destroy(tpsb_ctx *ctx);

/* vfes->GetNDofs(), num_equation */
int64_t tpsb_num_dofs(const tpsb_ctx *ctx);
int tpsb_num_equation(const tpsb_ctx *ctx);
/* which kernel set tpsb_create selected: 0 general trilinear 3-D, 1 affine 3-D (four launches), 2 affine 3-D fused
 * (three launches, p = 3), 3 generic tensor-product path (2-D, Gauss-Lobatto, mixtures)                      */
int tpsb_get_path(const tpsb_ctx *ctx);

/* RHSoperator::Mult(const Vector &x, Vector &y) const   (src/rhs_operator.cpp:343-464)
 * d_x, d_y: DEVICE pointers, neq*N doubles each, byNODES.  Asynchronous on the context stream.    */
int tpsb_rhs_mult(tpsb_ctx *ctx, const double *d_x, double *d_y);

/* Same call on HOST buffers (pinned or pageable): copies x in, runs Mult, copies y out, synchronises.
 * This is what a host-resident MFEM Vector costs and what bench.py reports as "e2e".               */
int tpsb_rhs_mult_host(tpsb_ctx *ctx, const double *h_x, double *h_y);

/* RHSoperator::updatePrimitives (src/rhs_operator.cpp:623-651) */
int tpsb_update_primitives(tpsb_ctx *ctx, const double *d_x);
/* RHSoperator::updateGradients (src/rhs_operator.cpp:653-680) */
int tpsb_update_gradients(tpsb_ctx *ctx, const double *d_x, int primitives_updated);
/* Views of the context-owned fields: M2ulPhyS::getPrimitiveGF / getGradientGF (src/M2ulPhyS.hpp:368-470).
 * Up: neq*N, gradUp: dim*neq*N doubles, device memory, valid until destroy.                       */
int tpsb_get_fields(tpsb_ctx *ctx, double **d_Up, double **d_gradUp);
/* Note (fused path, tpsb_get_path() == 2): tpsb_rhs_mult keeps Up and gradUp on chip.  tpsb_get_fields after a
 * tpsb_rhs_mult therefore re-evaluates them from the vector of that call -- which must still be alive and unchanged
 * -- exactly as the reference's members would read; tpsb_update_primitives / tpsb_update_gradients are the explicit
 * form (RHSoperator::updatePrimitives / updateGradients, what M2ulPhyS calls before output and averaging).     */
/* The solution grid function U_ that SourceTerm / AxisymmetricSource read the conserved state from
 * (src/source_term.cpp:66,77,121): in Runge-Kutta stages it is NOT the stage vector x handed to Mult, while
 * Up / gradUp are computed from x -- the reference's behaviour, kept (parity trap 1).
 * NULL (default): use x itself.  tpsb_ode_step sets it to d_U for the duration of the step.        */
int tpsb_set_solution_view(tpsb_ctx *ctx, const double *d_U);
/* The wall-distance grid function (M2ulPhyS::distance_, src/M2ulPhyS.cpp:265-283, computed by the host's distance
 * solver or read from a file): d_distance[N] in DEVICE memory, nodal, valid until replaced.  RHSoperator::GetFlux reads
 * it at the nodes (src/rhs_operator.cpp:534-537), FaceIntegrator / BCintegrator interpolate it to the face points with
 * each side's own shape functions (src/face_integrator.cpp:304-309, src/BCintegrator.cpp:408-411).  NULL: zero.     */
int tpsb_set_distance_field(tpsb_ctx *ctx, const double *d_distance);

/* Averaging::addSample -> addSampleInternal (src/averaging.cpp:198-420): running mean and (co)variances of a nodal family.
 *   d_inst  instantaneous field, DEVICE, num_fields x N doubles byNODES (M2ulPhyS registers Up, src/M2ulPhyS.cpp:634;
 *           NULL: the context's primitive field of the last tpsb_update_primitives / tpsb_get_fields)
 *   d_mean  running mean, same shape, updated in place: mean <- (ns_mean mean + inst) / (ns_mean + 1); with
 *           pressure_slot != 0 component 1 + dim is averaged as the PRESSURE of the instantaneous primitive state
 *           (the GasMixture overload, :330-420: "the instantaneous field contains temperature, but the averaged quantity is
 *           pressure")
 *   d_vari  NULL, or vari_components (vari_components + 1) / 2 fields: variances of fields vari_start .. vari_start +
 *           vari_components - 1 about the UPDATED mean, then their covariances (i < j), each (ns_vari v + d_i d_j) / (ns_vari + 1)
 * The caller keeps the sample counters (ns_mean_, ns_vari_ of the reference class) and zeroes mean / vari when they are 0. */
int tpsb_averaging_add_sample(tpsb_ctx *ctx, const double *d_inst, int num_fields, double *d_mean, double *d_vari, int vari_start,
                              int vari_components, int ns_mean, int ns_vari, int pressure_slot);

/* ---- forcing terms (ForcingTerms subclasses, src/forcing_terms.hpp:54-330): nodal terms added to dU/dt AFTER Me^-1
 * (src/rhs_operator.cpp:451-461), in the order they are registered here -- the reference registers them in the order
 * pressure gradient, sponge zones, heat sources, (SourceTerm, AxisymmetricSource: built in), Joule heating
 * (src/rhs_operator.cpp:101-167).  Node sets are not stored: every node re-evaluates its membership from its own
 * coordinates (the trilinear map of its element's vertices), which is what the reference's constructors do once.
 *   PRESSURE_GRADIENT  ConstantPressureGradient::updateTerms, CPU branch (src/forcing_terms.cpp:115-175):
 *                      d(rho u_d)/dt -= g_d ; d(rho E)/dt -= sum_d (u_d g_d + p d u_d / d x_d)
 *   HEAT_SOURCE        HeatSource (:923-1010), type "cylinder": + value on rho E for the nodes within `radius` of the
 *                      segment point1 -> point2
 *   JOULE_HEATING      JouleHeating::updateTerms (:443-470): + max(field, 0) on rho E (and on the electron energy of a
 *                      two-temperature mixture); field = nodal DEVICE array of N doubles, kept by reference
 *   SPONGE_ZONE        SpongeZone (:472-700), dry air: - a sigma multFactor (U - U_target), sigma from the planar or annular
 *                      zone geometry, a = speed of sound of the target, U_target user defined from (rho, u, v, w, p) or
 *                      "mixed out" from the mean normal fluxes over the nodes within tol of the zone's entry plane
 *                      (computeMixedOutValues, all-reduced over the ranks; DryAir::computeConservedStateFromConvectiveFlux,
 *                      src/equation_of_state.cpp:414-442)                                                          */
enum { TPSB_FORCING_PRESSURE_GRADIENT = 0, TPSB_FORCING_HEAT_SOURCE = 1, TPSB_FORCING_JOULE_HEATING = 2, TPSB_FORCING_SPONGE_ZONE = 3 };
typedef struct {
  int kind;
  double pressure_grad[3];                                    /* flow/pressureGrad                                        */
  double hs_point1[3], hs_point2[3], hs_radius, hs_value;     /* heatSource%d/{point1, point2, radius, value}             */
  const double *joule_heating;                                /* DEVICE, N doubles                                        */
  int sz_type;                                                /* spongezone%d/type-like: 0 planar, 1 annulus              */
  int sz_mixed_out;                                           /* 0 user-defined target, 1 mixed-out target                */
  double sz_normal[3], sz_point0[3], sz_point_init[3];        /* normal (normalised here), end plane point, entry plane pt */
  double sz_r1, sz_r2, sz_tol, sz_mult;                       /* annulus radii, plane tolerance, multFactor               */
  double sz_target[5];                                        /* rho, u, v, w, p of the user-defined target               */
} tpsb_forcing_desc;
int tpsb_add_forcing(tpsb_ctx *ctx, const tpsb_forcing_desc *desc);
int tpsb_clear_forcings(tpsb_ctx *ctx);

/* Test hook, host only: the static chunk schedule tpsb_rhs_mult_host uses to overlap copy-in / kernels / copy-out on a
 * single-rank 3-D mesh.  elem_begin / face_begin (two-sided faces, in their order) / bdr_begin (boundary faces, in their
 * order; may be NULL): chunks + 1 entries; ops: (kind, chunk) pairs, kind 0 copy-in + primitives, 1 gradient, 2 face
 * fluxes of the chunk's two-sided and boundary face ranges, 3 residual + copy-out.                                   */
int tpsb_debug_host_pipe_schedule(const tpsb_mesh_maps *maps, int chunks, int *elem_begin, int *face_begin, int *bdr_begin,
                                  int *ops, int max_ops, int *num_ops);

/* BoundaryCondition::dt (src/BoundaryCondition.hpp: a reference to M2ulPhyS::dt): the time step the non-reflecting inlets /
 * outlets advance their boundary states with at every evaluation.  tpsb_ode_step / tpsb_solve_step set it themselves. */
int tpsb_set_time_step(tpsb_ctx *ctx, double dt);
/* State of the non-reflecting condition on boundary attribute attr (this rank's part): mean_up[num_equation] = meanUp,
 * boundary_u[min(points, capacity)][num_equation] = boundaryU in face order, *num_points = its boundary quadrature points.
 * Either output may be NULL.  Synchronises the stream. */
int tpsb_get_bc_state(tpsb_ctx *ctx, int attr, double *mean_up, double *boundary_u, int capacity, int *num_points);

/* Chemistry::setGridFunctionRates (src/chemistry.cpp:133-140): the externally computed rate coefficients of the
 * GRIDFUNCTION_RXN reactions, d_rates[component][N] in DEVICE memory (byNODES), valid until replaced; NULL: those
 * reactions have k_f = 0 (src/reaction.cpp:113-116).                                                  */
int tpsb_set_reaction_rate_field(tpsb_ctx *ctx, const double *d_rates, int num_components);

/* RHSoperator::computeMeanTimeDerivatives / getLocalTimeDerivatives (src/rhs_operator.cpp:833-848): mean over the
 * local nodes of |dU/dt| per equation for a DEVICE vector d_y (neq*N, e.g. the result of tpsb_rhs_mult); out[neq] on
 * the host.  The reference evaluates it every 100th call; here the caller decides.                     */
int tpsb_get_mean_time_derivatives(tpsb_ctx *ctx, const double *d_y, double *out);

/* max_char_speed of the last Mult (src/rhs_operator.cpp:549-558), reduced over this rank's nodes
 * (and over ranks when a communicator is attached); synchronises the stream.
 * COLLECTIVE when the context has a communicator: tpsb_get_max_char_speed, tpsb_get_hmin, tpsb_check_state and
 * tpsb_solve_step (which calls them) all-reduce on the context's NCCL communicator, like the MPI_Allreduce calls they
 * replace -- every rank must call them, in the same order, and not concurrently with another call on the context. */
int tpsb_get_max_char_speed(tpsb_ctx *ctx, double *out);

/* MFEM ODESolver::Step for the solvers M2ulPhyS selects (src/M2ulPhyS.cpp:721-739, :2005):
 * scheme 1 ForwardEuler, 2 RK2(1.0), 3 RK3SSP, 4 RK4.  d_U is advanced in place nsteps times with a
 * constant dt, stage vectors stay on the device.                                                   */
int tpsb_ode_step(tpsb_ctx *ctx, double *d_U, double dt, int scheme, int nsteps);
/* hmin of M2ulPhyS (src/M2ulPhyS.cpp:756-761): min over all elements (all ranks) of Mesh::GetElementSize(e, 1).   */
int tpsb_get_hmin(tpsb_ctx *ctx, double *hmin);
/* M2ulPhyS::Check_NAN (src/M2ulPhyS.cpp:2463-2519) and Check_Undershoot (:2526-2549) of a device solution vector:
 * *num_nan = NaN entries over all ranks; mixtures: negative active-species densities are set to zero in place.     */
int tpsb_check_state(tpsb_ctx *ctx, double *d_U, int *num_nan);
/* M2ulPhyS::solveStep without the I/O (src/M2ulPhyS.cpp:2004-2016): one ODESolver::Step on the device solution vector,
 * Check_NAN (:2463-2519; *num_nan = NaN entries over all ranks), Check_Undershoot for mixtures (:2526-2549: active
 * species densities clamped at 0) and, when cfl > 0, the next adaptive time step
 * *dt_next = cfl * hmin / max_char_speed / dim (:2013-2016); cfl <= 0: constant time step, *dt_next = dt.         */
int tpsb_solve_step(tpsb_ctx *ctx, double *d_U, double dt, int scheme, double cfl, int *num_nan, double *dt_next);

/* Index maps derived at create, in the layout of the reference's precomputedIntegrationData
 * (src/dataStructures.hpp:297-517) -- exported so they can be compared bit-for-bit.
 * element_to_faces: 7*num_elems ints (count + up to 6 interior face ids, src/M2ulPhyS.cpp:878-958). */
int tpsb_get_element_to_faces(const tpsb_ctx *ctx, int *out);

/* Test hook: the reference-element tables the kernels use (1-D nodes/weights, differentiation and
 * interpolation matrices, face-node maps, orientation permutations) for a given order, flattened as
 * doubles in the order documented in tps_b200/csrc/tables.hpp (struct RefTables).  Returns the number
 * of doubles written (<= cap) or a negative error.                                                */
int tpsb_get_ref_tables(int order, double *out, int cap);

/* Test hook: device views of internal buffers. which = 0: face residuals [two-sided face][neq][(p+1)^2];
 * 1: face-trace blocks of the fast path [6*num_elems + shared faces][10][(p+1)^2] (NULL on other paths). */
int tpsb_debug_buffer(tpsb_ctx *ctx, int which, double **d_ptr, int64_t *count);

/* Test hook (generic path): the context's per-point physics on n points, point-major arrays on the device:
 * which = 0 primitives (GasMixture::GetPrimitivesFromConservatives), 1 max characteristic speed, 2 convective flux
 * (Fluxes::ComputeConvectiveFluxes), 3 viscous flux (ComputeViscousFluxes; d_aux = gradUp), 4 plasma source
 * (SourceTerm::updateTerms node body; d_aux = gradUp).                                              */
int tpsb_debug_point_eval(tpsb_ctx *ctx, int which, int n, const double *d_U, const double *d_aux, double *d_out);

/* Per-kernel device timers (the role GRVY timers play in the reference, src/M2ulPhyS.cpp:2146-2155):
 * with profiling on, every launch is bracketed by CUDA events on the context stream; the accumulated
 * milliseconds and launch counts per kernel class are returned in the order
 * {prim, grad, face_flux, elem_resid, pack, axpy}.  Profiling serialises nothing but adds events.  */
#define TPSB_NUM_KERNEL_CLASSES 6
int tpsb_set_profiling(tpsb_ctx *ctx, int on);
int tpsb_get_kernel_times(tpsb_ctx *ctx, double ms[TPSB_NUM_KERNEL_CLASSES], int64_t count[TPSB_NUM_KERNEL_CLASSES]);

/* Kernel launches issued by this context since create (for bench.py's gpu_launches). */
int64_t tpsb_launch_count(const tpsb_ctx *ctx);

/* ---- meshkit: host-side stand-in for the few MFEM mesh services the path needs when MFEM is absent ----
 * Cartesian hexahedral box, elements and vertices x-fastest, hex vertex order of
 * test/meshes/periodic-cube.mesh; periodic directions identify vertices (Mesh::MakePeriodic).
 * order_mode 0: lexicographic element order; 1: blocked (8^3 tiles) locality-preserving order.
 * elem_verts: [NE][8] ints; elem_xyz: [NE][8][3] doubles (un-wrapped).                              */
int tpsb_mk_cartesian_hex(int nx, int ny, int nz, const double lo[3], const double hi[3], const int periodic[3],
                          int order_mode, int *elem_verts, double *elem_xyz);
/* MFEM face generation (GetElementToFaceTable + GenerateFaces): returns the number of faces;
 * arrays sized 6*num_elems are always sufficient.  Pass NULL outputs to only count.                */
int tpsb_mk_build_faces(int num_elems, const int *elem_verts, int *face_el1, int *face_el2, int *face_inf1,
                        int *face_inf2);

/* 2-D counterparts (quadrilaterals, Geometry::Constants<SQUARE> vertex/edge order; the meshes utils/beam_mesh.cpp
 * makes for the 2-D test cases): elem_verts [NE][4], elem_xyz [NE][4][2]; arrays of 4*num_elems ints suffice.   */
int tpsb_mk_cartesian_quad(int nx, int ny, const double lo[2], const double hi[2], const int periodic[2],
                           int *elem_verts, double *elem_xyz);
int tpsb_mk_build_faces2d(int num_elems, const int *elem_verts, int *face_el1, int *face_el2, int *face_inf1,
                          int *face_inf2);

/* Structured block partition of the same box over a procs[0] x procs[1] x procs[2] rank grid (rank index
 * x-fastest): the stand-in for Mesh::GeneratePartitioning + ParMesh's face-neighbour tables
 * (src/M2ulPhyS.cpp:332,362,421; ExchangeFaceNbrData).  Local elements come first, then the
 * face-neighbour (halo) elements grouped by owner rank and sorted by global element id -- the order in
 * which the owner lists them in its send_elems, so one contiguous message per peer suffices.
 * Call once with all array pointers NULL to obtain the sizes, then with arrays of at least:
 *   elem_verts 8*(ne+nh) ints, elem_xyz 24*(ne+nh) doubles, elem_gid (ne+nh) int64 (global
 *   lexicographic element id), face_* 6*(ne+nh) ints each, nbr_rank/num peers (<= 26),
 *   send_offset/recv_offset peers+1, send_elems num_send.                                          */
typedef struct {
  int num_elems, num_nbr_elems, num_faces, num_nbr_ranks, num_send;
} tpsb_mk_part_sizes;
int tpsb_mk_partition(const int n[3], const double lo[3], const double hi[3], const int periodic[3],
                      const int procs[3], int rank, int order_mode, tpsb_mk_part_sizes *sizes, int *elem_verts,
                      double *elem_xyz, int64_t *elem_gid, int *face_el1, int *face_el2, int *face_inf1,
                      int *face_inf2, int *nbr_rank, int *send_offset, int *send_elems, int *recv_offset);

/* General element partitions (any conforming hexahedral mesh): the stand-in for Mesh::GeneratePartitioning(nprocs, 1)
 * -- METIS 5 k-way on the element dual graph, src/M2ulPhyS.cpp:332,362 -- or recursive coordinate bisection, and for
 * ParMesh's local numbering + face-neighbour tables (src/M2ulPhyS.cpp:421): local elements in global order, then the
 * face-neighbour elements grouped by owner and sorted by global element id.  tpsb_mk_partition_general is called
 * twice like tpsb_mk_partition (all output pointers NULL: sizes only); face_gface[f] is the global (MFEM) number of
 * local face f, for carrying boundary attributes over.                                                       */
int tpsb_mk_partition_metis(int num_elems, int num_faces, const int *face_el1, const int *face_el2, int nparts,
                            int *elem_rank, int64_t *edge_cut);
int tpsb_mk_partition_rcb(int num_elems, const double *elem_xyz, int nparts, int *elem_rank);
int tpsb_mk_partition_general(int num_elems, const int *elem_verts, const double *elem_xyz, int num_faces,
                              const int *gface_el1, const int *gface_el2, const int *elem_rank, int rank,
                              tpsb_mk_part_sizes *sizes, int *l_elem_verts, double *l_elem_xyz, int64_t *elem_gid,
                              int *face_el1, int *face_el2, int *face_inf1, int *face_inf2, int *face_gface, int *nbr_rank,
                              int *send_offset, int *send_elems, int *recv_offset);

/* the same for quadrilateral (dim = 2: elem_verts [NE][4], elem_xyz [NE][4][2]) or hexahedral (dim = 3) meshes */
int tpsb_mk_partition_rcb_dim(int dim, int num_elems, const double *elem_xyz, int nparts, int *elem_rank);
int tpsb_mk_partition_general_dim(int dim, int num_elems, const int *elem_verts, const double *elem_xyz, int num_faces,
                                  const int *gface_el1, const int *gface_el2, const int *elem_rank, int rank,
                                  tpsb_mk_part_sizes *sizes, int *l_elem_verts, double *l_elem_xyz, int64_t *elem_gid,
                                  int *face_el1, int *face_el2, int *face_inf1, int *face_inf2, int *face_gface, int *nbr_rank,
                                  int *send_offset, int *send_elems, int *recv_offset);

/* ---- communicator bootstrap for the NCCL face-neighbour exchange ----
 * unique_id: 128-byte ncclUniqueId produced on rank 0 by tpsb_comm_get_unique_id and broadcast by
 * the host (MPI_Bcast in TPS, torch.distributed in bench.py).                                      */
int tpsb_comm_get_unique_id(unsigned char unique_id[128]);
int tpsb_comm_init_rank(const unsigned char unique_id[128], int nranks, int rank, int device, void **nccl_comm);
int tpsb_comm_destroy(void *nccl_comm);

#ifdef __cplusplus
}
#endif
#endif /* TPSB200_H_ */
