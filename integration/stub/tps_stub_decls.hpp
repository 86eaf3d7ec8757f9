// DECLARATIONS ONLY -- a stand-in for the parts of MFEM and of the pecos/tps headers that
// integration/rhs_operator_b200.hpp touches, so that the binding can be syntax- and type-checked in an image that has
// neither MFEM nor the TPS build tree.  Names, member types and signatures follow MFEM 4.x (mesh/pmesh.hpp,
// fem/pfespace.hpp, linalg/vector.hpp, general/array.hpp, general/table.hpp) and the reference
// (src/dataStructures.hpp:65-730, src/run_configuration.hpp:90-410, src/rhs_operator.hpp:60-160).  Nothing here is linked or run.
#ifndef TPS_STUB_DECLS_HPP_
#define TPS_STUB_DECLS_HPP_
#include <cstdint>
#include <list>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace mfem {
void mfem_error(const char *msg);
template <class T>
class Array {
 public:
  Array();
  explicit Array(int n);
  int Size() const;
  T &operator[](int i);
  const T &operator[](int i) const;
  const T *GetData() const;
};
class Vector {
 public:
  Vector();
  Vector(double *data, int size);
  int Size() const;
  double &operator[](int i);
  const double &operator[](int i) const;
  double &operator()(int i);
  const double *Read(bool on_dev = true) const;
  const double *HostRead() const;
  double *Write(bool on_dev = true);
  double *ReadWrite(bool on_dev = true);
};
class DenseMatrix {
 public:
  double &operator()(int i, int j);
  const double &operator()(int i, int j) const;
};
class Table {
 public:
  const int *GetI() const;
  const int *GetJ() const;
};
class ElementTransformation {
 public:
  const DenseMatrix &GetPointMat() const;
  int OrderW() const;
};
class IntegrationPoint {
 public:
  double x, y, z, weight;
};
class IntegrationRule {
 public:
  int GetNPoints() const;
  const IntegrationPoint &IntPoint(int i) const;
};
class IntegrationRules {
 public:
  const IntegrationRule &Get(int GeomType, int Order);
};
class FiniteElement {
 public:
  int GetOrder() const;
};
class FaceElementTransformations {
 public:
  int Elem1No, Elem2No;
  ElementTransformation *Elem1;
  int GetGeometryType() const;
  void Transform(const IntegrationPoint &ip, Vector &x);
};
class ParMesh {
 public:
  int Dimension() const;
  int GetNE() const;
  int GetNBE() const;
  int GetNumFaces() const;
  int GetNFaceNeighbors() const;
  int GetNFaceNeighborElements() const;
  int GetNSharedFaces() const;
  int GetSharedFace(int sface) const;
  int GetFaceNbrRank(int fn) const;
  ElementTransformation *GetElementTransformation(int i);
  ElementTransformation *GetFaceNbrElementTransformation(int i);
  FaceElementTransformations *GetSharedFaceTransformations(int sf, bool fill2 = true);
  FaceElementTransformations *GetBdrFaceTransformations(int BdrElemNo);
  void GetFaceElements(int Face, int *Elem1, int *Elem2) const;
  void GetFaceInfos(int Face, int *Inf1, int *Inf2) const;
  int GetBdrElementFaceIndex(int be_idx) const;
  int GetBdrAttribute(int i) const;
  Table send_face_nbr_elements;
  Array<int> face_nbr_elements_offset;
};
class ParFiniteElementSpace {
 public:
  int GetVDim() const;
  const FiniteElement *GetFE(int i) const;
};
class ParGridFunction : public Vector {};
class Device {
 public:
  static int GetId();
};
class TimeDependentOperator {
 public:
  virtual void Mult(const Vector &x, Vector &y) const = 0;
  virtual ~TimeDependentOperator();
};
}  // namespace mfem

namespace gpudata {
const int MAXSPECIES = 200, MAXEQUATIONS = 205, MAXREACTIONS = 34, MAXCHEMPARAMS = 3, MAXDIM = 3;
}

// ---- src/dataStructures.hpp ----
enum Equations { EULER, NS, NS_PASSIVE };
enum WorkingFluid { DRY_AIR, USER_DEFINED, LTE_FLUID };
enum TransportModel { ARGON_MINIMAL, ARGON_MIXTURE, CONSTANT, LTE_TRANSPORT, MIXING_LENGTH, NUM_TRANSPORTMODEL, NITROGEN_MIXTURE };
enum ReactionModel { ARRHENIUS, HOFFERTLIEN, TABULATED_RXN, GRIDFUNCTION_RXN, RADIATIVE_DECAY, NUM_REACTIONMODEL };
enum RadiationModel { NONE_RAD, NET_EMISSION, NUM_RADIATIONMODEL };
enum NetEmissionCoefficientModel { TABULATED_NEC, NUM_NECMODEL };
enum GasParams { SPECIES_MW, SPECIES_CHARGES, FORMATION_ENERGY, SPECIES_DEGENERACY, NUM_GASPARAMS };
enum FluxTrns { VISCOSITY, BULK_VISCOSITY, HEAVY_THERMAL_CONDUCTIVITY, ELECTRON_THERMAL_CONDUCTIVITY, NUM_FLUX_TRANS };
enum SpeciesTrns { MF_FREQUENCY, NUM_SPECIES_COEFFS };
enum GasColl { CLMB_ATT, CLMB_REP, AR_AR1P, AR_E, AR_AR, NONE_GASCOLL };
enum GasType { ARGON_GAS, NITROGEN_GAS };
enum InletType { UNI_DENS_VEL, INTERPOLATE, SUB_DENS_VEL, SUB_DENS_VEL_FACE_X, SUB_DENS_VEL_FACE_Y, SUB_DENS_VEL_FACE_Z, SUB_DENS_VEL_NR, SUB_VEL_CONST_ENT };
enum OutletType { SUB_P, RESIST_IN, SUB_P_NR, SUB_MF_NR, SUB_MF_NR_PW };
enum WallType { INV, SLIP, VISC_ADIAB, VISC_ISOTH, VISC_GNRL };
enum ThermalCondition { ADIAB, ISOTH, SHTH, NONE_THMCND };
enum SpongeZoneSolution { USERDEF, MIXEDOUT, NONE_SZSOL };
enum SpongeZoneType { PLANAR, ANNULUS, NONE_SZTYPE };

struct SutherlandData { double C1, S0, Pr; };
struct linearlyVaryingVisc {
  mfem::Vector normal, point0, pointInit;
  double viscRatio, width, uniformMult;
  bool isEnabled;
};
struct SpongeZoneData {
  mfem::Vector normal, point0, pointInit;
  double r1, r2;
  SpongeZoneSolution szSolType;
  SpongeZoneType szType;
  double tol;
  mfem::Vector targetUp;
  double multFactor;
};
struct heatSourceData {
  bool isEnabled;
  double value;
  std::string type;
  mfem::Vector point1, point2;
  double radius;
};
struct constantTransportData {
  double viscosity, bulkViscosity, diffusivity[gpudata::MAXSPECIES], thermalConductivity, electronThermalConductivity;
  double mtFreq[gpudata::MAXSPECIES];
  int electronIndex;
};
struct mixingLengthTransportData { double max_mixing_length_, Prt_, Let_, bulk_multiplier_; };
struct WallData {
  ThermalCondition hvyThermalCond, elecThermalCond;
  double Th, Te;
};
struct DryAirInput {
  WorkingFluid f;
  Equations eq_sys;
  double specific_heat_ratio, gas_constant;
};
struct PerfectMixtureInput {
  WorkingFluid f;
  int numSpecies;
  bool isElectronIncluded, ambipolar, twoTemperature;
  double gasParams[gpudata::MAXSPECIES * NUM_GASPARAMS];
  double molarCV[gpudata::MAXSPECIES];
};
struct GasTransportInput {
  int neutralIndex, ionIndex, neutralIndex2, ionIndex2, electronIndex;
  bool thirdOrderkElectron;
  GasColl collisionIndex[gpudata::MAXSPECIES * gpudata::MAXSPECIES];
  GasType gas;
  bool multiply;
  double fluxTrnsMultiplier[NUM_FLUX_TRANS], spcsTrnsMultiplier[NUM_SPECIES_COEFFS], diffMult, mobilMult;
  bool constantActive;
  constantTransportData constantTransport;
};
struct TableInput {
  int Ndata;
  const double *xdata, *fdata;
  bool xLogScale, fLogScale;
  int order;
};
struct ReactionInput {
  TableInput tableInput;
  const double *modelParams;
  int indexInput;
};
struct ChemistryInput {
  int electronIndex, numReactions;
  double reactionEnergies[gpudata::MAXREACTIONS];
  bool detailedBalance[gpudata::MAXREACTIONS];
  int16_t reactantStoich[gpudata::MAXSPECIES * gpudata::MAXREACTIONS], productStoich[gpudata::MAXSPECIES * gpudata::MAXREACTIONS];
  ReactionModel reactionModels[gpudata::MAXREACTIONS];
  double equilibriumConstantParams[gpudata::MAXCHEMPARAMS * gpudata::MAXREACTIONS];
  ReactionInput reactionInputs[gpudata::MAXREACTIONS];
  double minimumTemperature;
};
struct RadiationInput {
  RadiationModel model;
  NetEmissionCoefficientModel necModel;
  TableInput necTableInput;
};

// ---- src/run_configuration.hpp ----
class RunConfiguration {
 public:
  int numSpongeRegions_, numHeatSources;
  heatSourceData *heatSource;
  bool useBCinGrad, use_mixing_length;
  SutherlandData sutherland_;
  std::vector<mfem::Vector> rxnModelParamsHost;
  constantTransportData constantTransport;
  mixingLengthTransportData mix_length_trans_input_;
  DryAirInput dryAirInput;
  PerfectMixtureInput perfectMixtureInput;
  GasTransportInput gasTransportInput;
  ChemistryInput chemistryInput;
  RadiationInput radiationInput;

  int GetSgsModelType();
  int GetSolutionOrder();
  int GetIntegrationRule();
  int GetBasisType();
  bool RoeRiemannSolverTPS() const;
  WorkingFluid GetWorkingFluid();
  double GetViscMult();
  double GetBulkViscMult();
  double GetSgsFloor();
  double GetReferenceLength();
  double GetSgsConstant();
  Equations GetEquationSystem() const;
  bool isAxisymmetric() const;
  linearlyVaryingVisc &GetLinearVaryingData();
  SpongeZoneData &GetSpongeZoneData(int sz);
  bool thereIsForcing();
  double *GetImposedPressureGradient();
  const std::vector<std::pair<int, InletType>> *GetInletPatchType() const;
  mfem::Array<double> GetInletData(int i);
  const std::vector<std::pair<int, OutletType>> *GetOutletPatchType() const;
  mfem::Array<double> GetOutletData(int out);
  std::vector<std::pair<int, WallType>> *GetWallPatchType();
  WallData GetWallData(int w);
  int GetNumSpecies();
  TransportModel GetTranportModel();
};

// ---- src/rhs_operator.hpp (the real constructor takes ~25 wiring arguments; the binding forwards them untouched) ----
class RHSoperator : public mfem::TimeDependentOperator {
 public:
  explicit RHSoperator(int &iter /* ... */);
  ~RHSoperator() override;
  void Mult(const mfem::Vector &x, mfem::Vector &y) const override;
};
#endif  // TPS_STUB_DECLS_HPP_
