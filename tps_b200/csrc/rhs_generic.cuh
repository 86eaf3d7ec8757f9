// Generic tensor-product path of RHSoperator::Mult: quadrilaterals or hexahedra, Gauss-Legendre or Gauss-Lobatto
// nodes and rules (flow/basisType, flow/integrationRule), any order the tables were built for, run-time
// num_equation.  It serves every configuration the two specialised 3-D paths do not (the reference's 2-D cases:
// mms.euler_2d runs p = 2, GLL/GLL quads).  Written for generality, not speed: one CTA per element, quadrature
// done point by point from dense reference-element tables (values/derivatives of the basis at the volume and face
// quadrature points), metric terms recomputed from the element's vertices, and each face flux evaluated by both
// of its elements (no face buffers, no atomics).
//
//   gen_prim_kernel   Up = prim(U)                                              (rhs_operator.cpp:623-651)
//   gen_grad_kernel   gradUp = Me^-1 ( sum_q w|J| phi (dphi/dx Up) + face jumps ) (gradients.cpp:144-232,
//                                                                                 faceGradientIntegration.cpp:40-140)
//   gen_resid_kernel  y = Me^-1 ( sum_q w (dphi adjJ) . F(q) + face fluxes )      (rhs_operator.cpp:379-448,493-559,
//                     domain_integrator.cpp:44-99, face_integrator.cpp:194-352)
// Me is the reference's mass matrix (rule of order 2p): diagonal for GL nodes + GL rule (collocated), dense
// otherwise -- then its inverse is precomputed per element on the host (Cholesky) and applied as a dense block,
// exactly like the reference's Me_inv (rhs_operator.cpp:173-189,432-448).
#pragma once
#include "gen_physics.cuh"

namespace tpsb {

// Boundary conditions of the generic path (BCintegrator's attribute maps, BCintegrator.cpp:64-125): same kinds and
// types as the 3-D dry-air table (physics.cuh: BcDev), with the full inlet inputState (species included).
constexpr int GEN_BC_NDATA = 12;
struct GenBc {
  int kind, type;  // kind: 0 inlet, 1 outlet, 2 wall; type: InletType / OutletType / WallType (dataStructures.hpp:168-196)
  int nr;          // non-reflecting / mass-flow condition: index of its patch state (GenArgs::nr), else -1
  double d[GEN_BC_NDATA];
};
// State of one non-reflecting inlet / outlet patch (InletBC / OutletBC members boundaryU, meanUp, tangent1, area_;
// src/inletBC.cpp:38-221, src/outletBC.cpp:38-212): all device-resident, so an evaluation never synchronises the host.
struct GenNrPatch {
  const int *faces;   // the patch's boundary faces on this rank, ascending face id
  int nfaces;
  double *boundaryU;  // [nfaces * nqf][neq] conserved boundary state per face quadrature point
  double *meanUp;     // [neq] mean of the interpolated primitives over the patch's points (all ranks)
  double *sums;       // [neq + 2] reduction scratch: sum of primitives, number of points, area
  double *area;       // [1] patch area over all ranks (BoundaryCondition::aggregateArea), set by the first evaluation
  int *init;          // [1] 0 until boundaryU has been initialised from the interpolated primitives
  double tangent1[3];
  double ref_length;
};
struct GenBcTable {
  int nbc, use_bc_in_grad;
  GenBc bc[MAX_BC];
};

struct GenArgs {
  int dim, np, dof, nqv, nqf, nfe, nv, neq, nvel, eq_system;
  int NE;
  long long N;
  int me_diag;
  GenPhys phys;
  // reference-element tables
  const double *phiV;   // [nqv][dof]
  const double *dphiV;  // [nqv][dof][dim]   reference-space derivatives
  const double *wV;     // [nqv]
  const double *xiV;    // [nqv][dim]
  const double *phiF;   // [codes][nqf][dof]  code = local_face * (dim == 3 ? 8 : 2) + orientation
  const double *xiF;    // [codes][nqf][dim]  element reference point of face quadrature point q
  const double *dlocF;  // [codes][dim*(dim-1)] d(xi)/d(face coordinates), column-major
  const double *wF;     // [nqf]
  // mesh
  const double *vx;     // [NE][nv][dim]
  const int *el_face;   // [NE][nfe] face id of each local face
  const int *f_el1, *f_el2, *f_inf1, *f_inf2;
  const double *me_inv;  // [NE][dof] (me_diag) or [NE][dof][dof]
  const double *me_inv_rad;  // axisymmetric: inverse of int r phi_i phi_j, applied in Mult (rhs_operator.cpp:441-445)
  const double *xiN;     // [dof][dim] reference coordinates of the nodes
  const int *f_bc;       // [NF] boundary face -> index into bct.bc, -1: no boundary condition / interior face
  GenBcTable bct;
  // fields
  const double *U;
  double *Up, *gradUp, *y;
  unsigned long long *maxCharBits;
  const double *dist;  // nodal wall distance (tpsb_set_distance_field; mixing-length model), NULL: 0
  // face-neighbour (halo) elements NE .. NE + NEH - 1 of a partitioned mesh: element-major copies [halo element][field][dof]
  // received from their owners (RHSoperator::initNBlockDataTransfer / waitAllDataTransfer, src/rhs_operator.cpp:716-831)
  int NEH;
  const double *Uhalo, *UpHalo, *gradUpHalo, *distHalo;
  // Fluxes' SGS model / viscous sponge: delta = h_min / order of every element (local, then face-neighbour); NULL: off
  const double *elem_delta;
  // non-reflecting inlets / outlets: patch states, first boundary-state point of every face (-1: none), BoundaryCondition::dt
  const GenNrPatch *nr;
  const int *f_nr_off;
  double bc_dt;
};

// physical point of reference point xi of element vertices v (sponge distance; z = 0 on quadrilaterals)
__device__ __forceinline__ void gen_point(int dim, const double *v, const double *xi, double *X) {
  if (dim == 3) {
    hex_point(v, xi[0], xi[1], xi[2], X);
  } else {
    const double x = xi[0], y = xi[1];
    for (int i = 0; i < 2; i++) X[i] = (1 - x) * (1 - y) * v[0 + i] + x * (1 - y) * v[2 + i] + x * y * v[4 + i] + (1 - x) * y * v[6 + i];
    X[2] = 0.0;
  }
}

// field `fld` (of nfld) of element el: local elements live in the byNODES array, face-neighbour elements in the halo copy
__device__ __forceinline__ const double *gen_elem_field(const GenArgs &a, const double *local, const double *halo, int el, int fld,
                                                        int nfld) {
  return el < a.NE ? local + static_cast<long long>(el) * a.dof + static_cast<long long>(fld) * a.N
                   : halo + (static_cast<long long>(el - a.NE) * nfld + fld) * a.dof;
}

__device__ __forceinline__ int gen_code(int dim, int inf) { return (inf / 64) * (dim == 3 ? 8 : 2) + inf % 64; }

// Jacobian of the multilinear map at reference point xi: J[i + dim*j] = d x_i / d xi_j
__device__ __forceinline__ void gen_jacobian(int dim, const double *v, const double *xi, double *J) {
  if (dim == 2) {
    const double x = xi[0], y = xi[1];
    for (int i = 0; i < 2; i++) {
      const double X0 = v[0 + i], X1 = v[2 + i], X2 = v[4 + i], X3 = v[6 + i];
      J[i + 0] = (1 - y) * (X1 - X0) + y * (X2 - X3);
      J[i + 2] = (1 - x) * (X3 - X0) + x * (X2 - X1);
    }
  } else {
    hex_jacobian(v, xi[0], xi[1], xi[2], J);
  }
}
// x coordinate (the radius of an axisymmetric run) of reference point xi of a quadrilateral
__device__ __forceinline__ double gen_quad_x(const double *v, const double *xi) {
  const double x = xi[0], y = xi[1];
  return (1 - x) * (1 - y) * v[0] + x * (1 - y) * v[2] + x * y * v[4] + (1 - x) * y * v[6];
}
__device__ __forceinline__ double gen_det(int dim, const double *J) {
  return dim == 2 ? J[0] * J[3] - J[2] * J[1] : det3(J);
}
__device__ __forceinline__ void gen_adj(int dim, const double *J, double *A) {
  if (dim == 2) {
    A[0] = J[3];
    A[1] = -J[1];
    A[2] = -J[2];
    A[3] = J[0];
  } else {
    adj3(J, A);
  }
}
// CalcOrtho of the face Jacobian J * dloc
__device__ __forceinline__ void gen_face_normal(int dim, const double *J, const double *dloc, double *nor) {
  double Jf[6];
  for (int i = 0; i < dim; i++)
    for (int c = 0; c < dim - 1; c++) {
      double a = 0;
      for (int k = 0; k < dim; k++) a += J[i + dim * k] * dloc[k + dim * c];
      Jf[i + dim * c] = a;
    }
  if (dim == 3) {
    nor[0] = Jf[1] * Jf[5] - Jf[2] * Jf[4];
    nor[1] = Jf[2] * Jf[3] - Jf[0] * Jf[5];
    nor[2] = Jf[0] * Jf[4] - Jf[1] * Jf[3];
  } else {
    nor[0] = Jf[1];
    nor[1] = -Jf[0];
  }
}

// BoundaryCondition::computeBdrPrimitiveStateForGradient (BoundaryCondition.cpp:55) / WallBC override
// (wallBC.cpp:241-266): only an isothermal wall changes the state used by the BR1 jump.
__device__ __forceinline__ void gen_bc_prim_for_gradient(const GenPhys &g, const GenBc &bc, const double *primIn, double *primBC) {
  for (int eq = 0; eq < g.neq; eq++) primBC[eq] = primIn[eq];
  if (bc.kind == 2 && bc.type == 3) {
    for (int i = 0; i < g.nvel; i++) primBC[1 + i] = 0.0;
    primBC[g.nvel + 1] = bc.d[0];
  }
}

// Non-reflecting / mass-flow conditions, dry air (the reference refuses them for mixtures, inletBC.cpp:49-52,
// outletBC.cpp:50-53): InletBC::subsonicNonReflectingDensityVelocity (inletBC.cpp:576-727; SUB_DENS_VEL_NR = 6,
// SUB_VEL_CONST_ENT = 7), OutletBC::subsonicNonReflectingPressure / subsonicNonRefMassFlow / subsonicNonRefPWMassFlow
// (outletBC.cpp:573-729, 739-892, 894-1027; SUB_P_NR = 2, SUB_MF_NR = 3, SUB_MF_NR_PW = 4).  The point's boundary state
// bU is advanced by dt times the characteristic flux derivative; the face flux is Lax-Friedrichs against the state bU had
// before.  gr[eq + d*neq]: interior gradients of the primitives; nor: CalcOrtho normal (outward).
__device__ __noinline__ void gen_nr_flux(const GenPhys &g, const GenBc &bc, const GenNrPatch &pt, double dt, const double *u1,
                                         const double *gr, const double *nor, double *bU, double *fx) {
  const int neq = g.neq, dim = g.dim, nvel = g.nvel;
  const double gamma = g.dry.gamma, Rg = g.dry.R;
  const bool inlet = bc.kind == 0;
  double unitNorm[3] = {0, 0, 0}, tangent2[3] = {0, 0, 0}, meanUp[GEN_MAXEQ];
  const double *tangent1 = pt.tangent1;
  for (int eq = 0; eq < neq; eq++) meanUp[eq] = pt.meanUp[eq];
  {
    double mod = 0.;
    for (int d = 0; d < dim; d++) mod += nor[d] * nor[d];
    for (int d = 0; d < dim; d++) unitNorm[d] = nor[d] * ((inlet ? -1. : 1.) / sqrt(mod));  // inlet: into the domain
  }
  double meanVel[3] = {0, 0, 0};
  for (int d = 0; d < dim; d++) {
    meanVel[0] += unitNorm[d] * meanUp[d + 1];
    meanVel[1] += tangent1[d] * meanUp[d + 1];
  }
  if (dim == 3) {
    tangent2[0] = unitNorm[1] * tangent1[2] - unitNorm[2] * tangent1[1];
    tangent2[1] = unitNorm[2] * tangent1[0] - unitNorm[0] * tangent1[2];
    tangent2[2] = unitNorm[0] * tangent1[1] - unitNorm[1] * tangent1[0];
    for (int d = 0; d < dim; d++) meanVel[2] += tangent2[d] * meanUp[d + 1];
  }
  double normGrad[GEN_MAXEQ];
  for (int eq = 0; eq < neq; eq++) {
    normGrad[eq] = 0.;
    for (int d = 0; d < dim; d++) normGrad[eq] += unitNorm[d] * gr[eq + d * neq];
  }
  // DryAir::ComputePressureDerivative(normGrad, stateIn, false) (equation_of_state.cpp:350-359)
  const double T = dry_gen_pressure(g, u1) / (Rg * u1[0]);
  const double dpdn = Rg * (T * normGrad[0] + u1[0] * normGrad[1 + nvel]);
  const double speedSound = sqrt(gamma * Rg * meanUp[1 + nvel]);
  double meanK = 0.;
  for (int d = 0; d < nvel; d++) meanK += meanUp[1 + d] * meanUp[1 + d];
  meanK *= 0.5;
  const double sigma = speedSound / pt.ref_length;
  double L1, L2, L3, L4 = 0., L5;
  if (inlet) {
    double meanDV[3];
    for (int d = 0; d < nvel; d++) meanDV[d] = meanUp[1 + d] - bc.d[1 + d];
    L1 = 0.;
    for (int d = 0; d < dim; d++) L1 += unitNorm[d] * normGrad[1 + d];
    L1 = dpdn - meanUp[0] * speedSound * L1;
    L1 *= meanVel[0] - speedSound;
    L5 = 0.;
    for (int d = 0; d < dim; d++) L5 += meanDV[d] * unitNorm[d];
    L5 *= sigma * 2. * meanUp[0] * speedSound;
    L3 = 0.;
    for (int d = 0; d < dim; d++) L3 += meanDV[d] * tangent1[d];
    L3 *= sigma;
    if (dim == 3) {
      for (int d = 0; d < dim; d++) L4 += meanDV[d] * tangent2[d];
      L4 *= sigma;
    }
    L2 = sigma * speedSound * speedSound * (meanUp[0] - bc.d[0]) - 0.5 * L5;
    if (bc.type == 7) L2 = 0.;
  } else {
    L2 = speedSound * speedSound * normGrad[0] - dpdn;
    L2 *= meanVel[0];
    L3 = 0.;
    for (int d = 0; d < dim; d++) L3 += tangent1[d] * normGrad[1 + d];
    L3 *= meanVel[0];
    if (dim == 3) {
      for (int d = 0; d < dim; d++) L4 += tangent2[d] * normGrad[1 + d];
      L4 *= meanVel[0];
    }
    L5 = 0.;
    for (int d = 0; d < dim; d++) L5 += unitNorm[d] * normGrad[1 + d];
    L5 = dpdn + meanUp[0] * speedSound * L5;
    L5 *= meanVel[0] + speedSound;
    if (bc.type == 2) {
      const double meanP = Rg * meanUp[0] * meanUp[1 + nvel];
      L1 = sigma * (meanP - bc.d[0]);
    } else {
      double vn = meanVel[0];  // SUB_MF_NR: the patch mean; SUB_MF_NR_PW: the point's own normal velocity
      if (bc.type == 4) {
        vn = 0.;
        for (int d = 0; d < dim; d++) vn += u1[1 + d] * unitNorm[d];
        vn /= u1[0];
      }
      L1 = -sigma * (vn - bc.d[0] / meanUp[0] / pt.area[0]);
      L1 *= meanUp[0] * speedSound;
    }
  }
  const double d1 = (L2 + 0.5 * (L5 + L1)) / speedSound / speedSound;
  const double d2 = 0.5 * (L5 - L1) / meanUp[0] / speedSound;
  const double d3 = L3, d4 = L4, d5 = 0.5 * (L5 + L1);
  double dF[GEN_MAXEQ];
  for (int eq = 0; eq < neq; eq++) dF[eq] = 0.;
  dF[0] = d1;
  dF[1] = meanVel[0] * d1 + meanUp[0] * d2;
  dF[2] = meanVel[1] * d1 + meanUp[0] * d3;
  if (dim == 3) dF[3] = meanVel[2] * d1 + meanUp[0] * d4;
  dF[1 + dim] = meanUp[0] * meanVel[0] * d2;
  dF[1 + dim] += meanUp[0] * meanVel[1] * d3;
  if (dim == 3) dF[1 + dim] += meanUp[0] * meanVel[2] * d4;
  dF[1 + dim] += meanK * d1 + d5 / (gamma - 1.);
  double state2[GEN_MAXEQ], stateN[GEN_MAXEQ], newU[GEN_MAXEQ];
  for (int eq = 0; eq < neq; eq++) state2[eq] = bU[eq], stateN[eq] = bU[eq];
  for (int d = 0; d < dim; d++) stateN[1 + d] = 0.;
  for (int d = 0; d < dim; d++) {
    stateN[1] += state2[1 + d] * unitNorm[d];
    stateN[2] += state2[1 + d] * tangent1[d];
    if (dim == 3) stateN[3] += state2[1 + d] * tangent2[d];
  }
  for (int i = 0; i < neq; i++) newU[i] = stateN[i] - dt * dF[i];
  {  // back to Cartesian momentum: inverse (adjugate / determinant, mfem::CalcInverse) of the matrix with rows n, t1, t2
    double momX[3] = {0, 0, 0};
    if (dim == 2) {
      const double a = unitNorm[0], b = tangent1[0], c = unitNorm[1], d = tangent1[1];  // column-major M(0,0), M(1,0), M(0,1), M(1,1)
      const double t = 1.0 / (a * d - b * c);
      momX[0] = (d * t) * newU[1] + (-c * t) * newU[2];
      momX[1] = (-b * t) * newU[1] + (a * t) * newU[2];
    } else {
      double a[9], inv[9];
      for (int d = 0; d < 3; d++) a[0 + d * 3] = unitNorm[d], a[1 + d * 3] = tangent1[d], a[2 + d * 3] = tangent2[d];
      const double t = 1.0 / (a[0] * (a[4] * a[8] - a[5] * a[7]) - a[3] * (a[1] * a[8] - a[2] * a[7]) + a[6] * (a[1] * a[5] - a[2] * a[4]));
      inv[0] = (a[4] * a[8] - a[5] * a[7]) * t, inv[3] = (a[5] * a[6] - a[3] * a[8]) * t, inv[6] = (a[3] * a[7] - a[4] * a[6]) * t;
      inv[1] = (a[2] * a[7] - a[1] * a[8]) * t, inv[4] = (a[0] * a[8] - a[2] * a[6]) * t, inv[7] = (a[1] * a[6] - a[0] * a[7]) * t;
      inv[2] = (a[1] * a[5] - a[2] * a[4]) * t, inv[5] = (a[2] * a[3] - a[0] * a[5]) * t, inv[8] = (a[0] * a[4] - a[1] * a[3]) * t;
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) momX[i] += inv[i + j * 3] * newU[1 + j];
    }
    for (int d = 0; d < dim; d++) newU[1 + d] = momX[d];
  }
  for (int eq = 0; eq < neq; eq++) bU[eq] = newU[eq];
  gen_riemann_lf(g, u1, state2, nor, fx);
}

// BCintegrator::computeBdrFlux (BCintegrator.cpp:228-242) -> InletBC::subsonicReflectingDensityVelocity
// (inletBC.cpp:729-756), OutletBC::subsonicReflectingPressure (outletBC.cpp:731-737), WallBC::computeINVwallFlux /
// computeAdiabaticWallFlux / computeIsothermalWallFlux (wallBC.cpp:277-320, 430-510), any fluid, 2-D / 3-D /
// axisymmetric.  gr[eq + d*neq]: interior gradients of the primitives; nor: CalcOrtho normal (area weighted, outward).
__device__ __noinline__ void gen_bc_flux(const GenPhys &g, const GenBc &bc, int use_bc_in_grad, const double *u1, const double *gr,
                                         const double *nor, double radius, double *fx, double distance = 0.0,
                                         const DryAux *ax = nullptr) {
  const int neq = g.neq, dim = g.dim, nvel = g.nvel;
  double s2[GEN_MAXEQ], viscF[GEN_MAXEQ * GEN_MAXDIM], wallViscF[GEN_MAXEQ], un[3];
  double normN = 0.;
  for (int d = 0; d < dim; d++) normN += nor[d] * nor[d];
  const bool ns = g.dry.eq_system != 0 || (g.fluid && g.mix->eq_system != 0);
  if (bc.kind == 0) {
    const double pr = gen_pressure(g, u1);
    for (int eq = 0; eq < neq; eq++) s2[eq] = u1[eq];
    s2[0] = bc.d[0];
    s2[1] = bc.d[0] * bc.d[1];
    s2[2] = bc.d[0] * bc.d[2];
    if (nvel == 3) s2[3] = bc.d[0] * bc.d[3];
    const int nact = gen_num_active_species(g);
    for (int sp = 0; sp < nact; sp++) s2[nvel + 2 + sp] = bc.d[4 + sp];
    gen_modify_energy_for_pressure(g, s2, s2, pr, true);
    gen_riemann_lf(g, u1, s2, nor, fx);
    return;
  }
  if (bc.kind == 1) {
    gen_modify_energy_for_pressure(g, u1, s2, bc.d[0], false);
    gen_riemann_lf(g, u1, s2, nor, fx);
    return;
  }
  if (bc.type == 1) {  // SLIP (wallBC.cpp:326-428): mirror state, Riemann flux only (Roe when useRoe)
    double vel[3] = {0, 0, 0}, nVel[3];
    for (int d = 0; d < nvel; d++) vel[d] = u1[1 + d] / u1[0];
    slip_mirror_velocity(dim, nor, vel, nVel);
    for (int eq = 0; eq < neq; eq++) s2[eq] = u1[eq];
    s2[1] = u1[0] * nVel[0];
    s2[2] = u1[0] * nVel[1];
    if (dim == 3) s2[3] = u1[0] * nVel[2];
    gen_riemann(g, u1, s2, nor, fx);
    return;
  }
  if (bc.type == 4) {
    // VISC_GNRL: WallBC constructor (wallBC.cpp:112-147) + computeGeneralWallFlux (:512-543).
    // d = {hvyThermalCond, elecThermalCond, Th, Te}; ThermalCondition 0 ADIAB, 1 ISOTH, 2 SHTH
    const int hvy = static_cast<int>(bc.d[0]), elec = static_cast<int>(bc.d[1]);
    const bool twoT = g.fluid && g.mix->twoTemp;
    const int nsp = g.fluid ? g.mix->numSpecies : 1;
    double prim[GEN_MAXEQ], pf[GEN_MAXEQ + 4];
    gen_prim(g, u1, prim);
    for (int d = 0; d < nvel; d++) prim[1 + d] = 0.0;
    if (hvy == 1) prim[nvel + 1] = bc.d[2];
    if (elec == 1) prim[neq - 1] = bc.d[3];  // index num_equation - 1 in either temperature model (wallBC.cpp:133-134)
    gen_cons(g, prim, s2);
    gen_riemann_lf(g, u1, s2, nor, fx);
    if (!ns) return;
    for (int i = 0; i < GEN_MAXEQ + 4; i++) pf[i] = 0.0;
    if (elec == 2 && g.fluid) mix_sheath_bdr_flux(*g.mix, s2, pf);
    const double isq = 1. / sqrt(normN);
    for (int d = 0; d < dim; d++) un[d] = nor[d] * isq;
    const bool hvy_presc = hvy != 1, elec_presc = (elec == 0) || (elec == 2 && twoT);
    if (g.fluid)
      mix_bdr_visc_flux_general(*g.mix, s2, gr, radius, un, pf, hvy_presc, elec_presc, wallViscF, &g.dry, ax);
    else
      dry_gen_bdr_visc_flux(g, s2, gr, radius, un, hvy_presc, wallViscF, ax);
    (void)nsp;
    const double nm = sqrt(normN);
    for (int eq = 0; eq < neq; eq++) wallViscF[eq] *= nm;
    gen_visc_flux(g, u1, gr, radius, viscF, 0.0, ax);
    for (int eq = 1; eq < neq; eq++) {
      fx[eq] -= 0.5 * wallViscF[eq];
      for (int d = 0; d < dim; d++) fx[eq] -= 0.5 * viscF[eq + d * neq] * nor[d];
    }
    return;
  }
  if (bc.type == 0) {  // INV: mirror state
    const double norm = sqrt(normN);
    double vel[3] = {0, 0, 0};
    for (int d = 0; d < nvel; d++) vel[d] = u1[1 + d] / u1[0];
    for (int d = 0; d < dim; d++) un[d] = nor[d] / norm;
    double vn = 0;
    for (int d = 0; d < dim; d++) vn += vel[d] * un[d];
    for (int eq = 0; eq < neq; eq++) s2[eq] = u1[eq];
    s2[1] = u1[0] * (vel[0] - 2. * vn * un[0]);
    s2[2] = u1[0] * (vel[1] - 2. * vn * un[1]);
    if (dim == 3) s2[3] = u1[0] * (vel[2] - 2. * vn * un[2]);
    if (nvel == 3 && dim == 2) s2[3] = u1[0] * vel[2];
    gen_riemann(g, u1, s2, nor, fx);  // the inviscid wall does not force Lax-Friedrichs (wallBC.cpp:301)
    if (!ns) return;
    gen_visc_flux(g, s2, gr, radius, viscF, distance, ax);  // only the inviscid wall passes the wall distance on (wallBC.cpp:309-313)
    for (int eq = 0; eq < neq; eq++) {
      wallViscF[eq] = 0.;
      for (int d = 0; d < dim; d++) wallViscF[eq] += viscF[eq + d * neq] * nor[d];
    }
    gen_visc_flux(g, u1, gr, radius, viscF, distance, ax);
  } else {
    const double isq = 1. / sqrt(normN);
    for (int d = 0; d < dim; d++) un[d] = nor[d] * isq;
    if (bc.type == 2) {  // VISC_ADIAB: stagnation state
      gen_stagnation_state(g, u1, s2);
      gen_riemann_lf(g, u1, s2, nor, fx);
      if (!ns) return;
      gen_bdr_visc_flux(g, s2, gr, radius, un, true, wallViscF, ax);
    } else {  // VISC_ISOTH
      if (use_bc_in_grad) {
        for (int eq = 0; eq < neq; eq++) s2[eq] = u1[eq];
        for (int i = 0; i < nvel; i++) s2[i + 1] *= -1.0;
      } else {
        gen_stagnant_state_with_temp(g, u1, bc.d[0], s2);
      }
      gen_riemann_lf(g, u1, s2, nor, fx);
      if (!ns) return;
      gen_stagnant_state_with_temp(g, u1, bc.d[0], s2);
      gen_bdr_visc_flux(g, s2, gr, radius, un, false, wallViscF, ax);
    }
    const double nm = sqrt(normN);
    for (int eq = 0; eq < neq; eq++) wallViscF[eq] *= nm;
    gen_visc_flux(g, u1, gr, radius, viscF, 0.0, ax);
  }
  for (int eq = 1; eq < neq; eq++) {
    fx[eq] -= 0.5 * wallViscF[eq];
    for (int d = 0; d < dim; d++) fx[eq] -= 0.5 * viscF[eq + d * neq] * nor[d];
  }
}

__global__ void gen_prim_kernel(GenArgs a) {
  const long long n = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  double s[GEN_MAXEQ], up[GEN_MAXEQ];
  for (int eq = 0; eq < a.neq; eq++) s[eq] = a.U[n + eq * a.N];
  gen_prim(a.phys, s, up);
  for (int eq = 0; eq < a.neq; eq++) a.Up[n + eq * a.N] = up[eq];
}

// y_el = Me^-1 z_el for nc interleaved columns: z[k*nc + c] -> out(global, byNODES with component stride cstride)
__device__ __forceinline__ void gen_apply_minv(const GenArgs &a, const double *me_inv, int e, const double *z, int nc,
                                               double *out, long long cstride) {
  const int dof = a.dof;
  for (int t = threadIdx.x; t < dof * nc; t += blockDim.x) {
    const int j = t % dof, c = t / dof;
    double v;
    if (a.me_diag) {
      v = me_inv[static_cast<long long>(e) * dof + j] * z[j * nc + c];
    } else {
      const double *mi = me_inv + (static_cast<long long>(e) * dof + j) * dof;
      v = 0;
      for (int k = 0; k < dof; k++) v += mi[k] * z[k * nc + c];
    }
    out[static_cast<long long>(e) * dof + j + c * cstride] = v;
  }
}

// ------------------------------------------------------------------------------------------------
template <int DIM, int NVEL, int NEQ>
__global__ void __launch_bounds__(128) gen_grad_kernel(GenArgs a) {
  extern __shared__ double sm[];
  GenPhys phl;
  if (DIM > 0) {  // dry-air instantiation: sizes and the fluid switch become constants the inlined physics folds
    phl = a.phys;
    phl.dim = DIM, phl.nvel = NVEL, phl.neq = NEQ, phl.fluid = 0, phl.mix = nullptr;
  }
  const GenPhys &ph = DIM > 0 ? phl : a.phys;
  const int e = blockIdx.x, dim = DIM > 0 ? DIM : a.dim, dof = a.dof, neq = DIM > 0 ? NEQ : a.neq, nc = neq * dim;
  const int nqmax = a.nqv > a.nfe * a.nqf ? a.nqv : a.nfe * a.nqf;
  double *sUp = sm;                    // [neq][dof]
  double *sRhs = sUp + neq * dof;      // [dof][nc]
  double *sQ = sRhs + dof * nc;        // [max(nqv, nfe * nqf)][nc]
  double *sW = sQ + nqmax * nc;        // [nqv]  w |J|
  const long long N = a.N;
  const double *vx = a.vx + static_cast<long long>(e) * a.nv * dim;
  for (int t = threadIdx.x; t < neq * dof; t += blockDim.x)
    sUp[t] = a.Up[static_cast<long long>(e) * dof + (t % dof) + (t / dof) * N];
  __syncthreads();
  // volume: physical gradient of Up at the quadrature points
  for (int q = threadIdx.x; q < a.nqv; q += blockDim.x) {
    double J[9], A[9];
    gen_jacobian(dim, vx, a.xiV + q * dim, J);
    const double det = gen_det(dim, J);
    gen_adj(dim, J, A);
    sW[q] = a.wV[q] * det;
    const double *dp = a.dphiV + static_cast<long long>(q) * dof * dim;
    for (int eq = 0; eq < neq; eq++) {
      double gr[GEN_MAXDIM] = {0, 0, 0};
      for (int k = 0; k < dof; k++) {
        const double u = sUp[eq * dof + k];
        for (int r = 0; r < dim; r++) gr[r] += dp[k * dim + r] * u;
      }
      for (int d = 0; d < dim; d++) {
        double g = 0;
        for (int r = 0; r < dim; r++) g += gr[r] * A[r + dim * d];
        sQ[q * nc + eq + d * neq] = g / det;  // CalcPhysDShape = dshape * inv(J)
      }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < dof * nc; t += blockDim.x) {
    const int j = t / nc, c = t % nc;
    double v = 0;
    for (int q = 0; q < a.nqv; q++) v += a.phiV[q * dof + j] * sW[q] * sQ[q * nc + c];
    sRhs[j * nc + c] = v;
  }
  __syncthreads();
  // faces: phi_j(q) * 1/2 (Up_other - Up_own)(q) * n_out(q) w_q.  One task per (local face, face point): all faces of
  // the element are in flight at once (a p = 2 quadrilateral has 4 points per face -- looping over the faces left 4
  // threads busy), then one lift pass that adds the faces in local-face order, as the face-by-face loop did.
  double *sQF = sQ;  // [nfe * nqf][nc]  (the volume values in sQ are consumed)
  for (int it = threadIdx.x; it < a.nfe * a.nqf; it += blockDim.x) {
    const int lf = it / a.nqf, q = it - lf * a.nqf;
    const int f = a.el_face[e * a.nfe + lf];
    const int e1 = a.f_el1[f], e2 = a.f_el2[f];
    const bool bdr = e2 < 0;
    // boundary face: Up2 = Up1 (zero jump) unless useBCinGrad and a boundary condition supplies a state
    // (faceGradientIntegration.cpp:93-115)
    if (bdr && !(a.bct.use_bc_in_grad && a.f_bc[f] >= 0)) continue;
    const bool first = (e1 == e);
    const int code_own = gen_code(dim, first ? a.f_inf1[f] : a.f_inf2[f]);
    const int code_oth = bdr ? code_own : gen_code(dim, first ? a.f_inf2[f] : a.f_inf1[f]);
    const int code1 = gen_code(dim, a.f_inf1[f]);
    const int eo = bdr ? e1 : (first ? e2 : e1);
    const double *v1 = a.vx + static_cast<long long>(e1) * a.nv * dim;
    double J[9], nor[3];
    gen_jacobian(dim, v1, a.xiF + (static_cast<long long>(code1) * a.nqf + q) * dim, J);
    gen_face_normal(dim, J, a.dlocF + code1 * dim * (dim - 1), nor);
    const double sg = (first ? 1.0 : -1.0) * a.wF[q];
    const double *po = a.phiF + (static_cast<long long>(code_own) * a.nqf + q) * dof;
    const double *pn = a.phiF + (static_cast<long long>(code_oth) * a.nqf + q) * dof;
    double *dst = sQF + static_cast<long long>(it) * nc;
    if (bdr) {
      double own[GEN_MAXEQ], pbc[GEN_MAXEQ];
      for (int eq = 0; eq < neq; eq++) {
        double v = 0;
        for (int k = 0; k < dof; k++) v += po[k] * sUp[eq * dof + k];
        own[eq] = v;
      }
      gen_bc_prim_for_gradient(ph, a.bct.bc[a.f_bc[f]], own, pbc);
      for (int eq = 0; eq < neq; eq++) {
        const double jump = 0.5 * (pbc[eq] - own[eq]);
        for (int d = 0; d < dim; d++) dst[eq + d * neq] = jump * nor[d] * sg;
      }
      continue;
    }
    for (int eq = 0; eq < neq; eq++) {
      double own = 0, oth = 0;
      const double *un = gen_elem_field(a, a.Up, a.UpHalo, eo, eq, neq);
      for (int k = 0; k < dof; k++) {
        own += po[k] * sUp[eq * dof + k];
        oth += pn[k] * un[k];
      }
      const double jump = 0.5 * (oth - own);
      for (int d = 0; d < dim; d++) dst[eq + d * neq] = jump * nor[d] * sg;
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < dof * nc; t += blockDim.x) {
    const int j = t / nc, c = t % nc;
    double acc = sRhs[j * nc + c];
    for (int lf = 0; lf < a.nfe; lf++) {
      const int f = a.el_face[e * a.nfe + lf];
      if (a.f_el2[f] < 0 && !(a.bct.use_bc_in_grad && a.f_bc[f] >= 0)) continue;
      const int code_own = gen_code(dim, a.f_el1[f] == e ? a.f_inf1[f] : a.f_inf2[f]);
      const double *po = a.phiF + static_cast<long long>(code_own) * a.nqf * dof;
      const double *src = sQF + static_cast<long long>(lf) * a.nqf * nc;
      double v = 0;
      for (int q = 0; q < a.nqf; q++) v += po[q * dof + j] * src[q * nc + c];
      acc += v;
    }
    sRhs[j * nc + c] = acc;
  }
  __syncthreads();
  // gradUp[e*dof + j + eq*N + d*neq*N]: component c = eq + d*neq has stride N
  gen_apply_minv(a, a.me_inv, e, sRhs, nc, a.gradUp, N);
}

// ------------------------------------------------------------------------------------------------
// DIM / NVEL / NEQ > 0: compile-time copies of the run-time sizes (dry-air instantiations).  The per-point routines are
// inlined, so with constant sizes their equation / dimension loops unroll and the small per-thread arrays live in
// registers instead of run-time-indexed local memory; <0,0,0> is the fully run-time form (mixtures).
template <int DIM, int NVEL, int NEQ>
__global__ void __launch_bounds__(128) gen_resid_kernel(GenArgs a) {
  extern __shared__ double sm[];
  GenPhys phl;
  if (DIM > 0) {  // dry-air instantiation: sizes and the fluid switch become constants the inlined physics folds
    phl = a.phys;
    phl.dim = DIM, phl.nvel = NVEL, phl.neq = NEQ, phl.fluid = 0, phl.mix = nullptr;
  }
  const GenPhys &ph = DIM > 0 ? phl : a.phys;
  const int nvel = DIM > 0 ? NVEL : a.nvel;
  const int e = blockIdx.x, dim = DIM > 0 ? DIM : a.dim, dof = a.dof, neq = DIM > 0 ? NEQ : a.neq, nc = neq * dim;
  const int nqmax = a.nqv > a.nfe * a.nqf ? a.nqv : a.nfe * a.nqf;
  double *sU = sm;                  // [neq][dof]
  double *sG = sU + neq * dof;      // [nc][dof]   gradUp of the element
  double *sF = sG + nc * dof;       // [dof][nc]   nodal flux F_c - F_v
  double *sQ = sF + dof * nc;       // [nqmax][nc]
  double *sZ = sQ + nqmax * nc;     // [dof][neq]
  __shared__ unsigned long long sMaxBits;
  const int nact = gen_num_active_species(ph);
  const long long N = a.N;
  const double *vx = a.vx + static_cast<long long>(e) * a.nv * dim;
  if (threadIdx.x == 0) sMaxBits = 0ull;
  for (int t = threadIdx.x; t < neq * dof; t += blockDim.x) sU[t] = a.U[static_cast<long long>(e) * dof + (t % dof) + (t / dof) * N];
  for (int t = threadIdx.x; t < nc * dof; t += blockDim.x)
    sG[t] = a.gradUp[static_cast<long long>(e) * dof + (t % dof) + (t / dof) * N];
  __syncthreads();
  // nodal flux (GetFlux, rhs_operator.cpp:493-559) and max characteristic speed
  for (int k = threadIdx.x; k < dof; k += blockDim.x) {
    double s[GEN_MAXEQ], gr[GEN_MAXEQ * GEN_MAXDIM], fc[GEN_MAXEQ * GEN_MAXDIM], fv[GEN_MAXEQ * GEN_MAXDIM];
    for (int eq = 0; eq < neq; eq++) s[eq] = sU[eq * dof + k];
    // species densities are clamped >= 0 at the nodes of GetFlux (rhs_operator.cpp:512-517)
    for (int sp = 0; sp < nact; sp++) s[nvel + 2 + sp] = fmax(s[nvel + 2 + sp], 0.0);
    for (int c = 0; c < nc; c++) gr[c] = sG[c * dof + k];
    gen_conv_flux(ph, s, fc);
    if (a.eq_system != 0) {
      const double radius = ph.axisym ? gen_quad_x(vx, a.xiN + k * dim) : -1.0;  // nodal coordinate (GetFlux :526-528)
      const double dw = a.dist ? a.dist[static_cast<long long>(e) * dof + k] : 0.0;  // rhs_operator.cpp:534-537
      DryAux ax;
      if (a.elem_delta) {  // SGS model / viscous sponge: the element's delta, the node's coordinates (rhs_operator.cpp:526-533)
        ax.delta = a.elem_delta[e];
        gen_point(dim, vx, a.xiN + k * dim, ax.x);
      }
      gen_visc_flux(ph, s, gr, radius, fv, dw, a.elem_delta ? &ax : nullptr);
      for (int c = 0; c < nc; c++) fc[c] -= fv[c];
    }
    for (int c = 0; c < nc; c++) sF[k * nc + c] = fc[c];
    atomicMax(&sMaxBits, static_cast<unsigned long long>(__double_as_longlong(gen_max_char_speed(ph, s))));
  }
  __syncthreads();
  if (threadIdx.x == 0) atomicMax(a.maxCharBits, sMaxBits);
  // volume: G(q)[eq][r] = w_q sum_d adjJ(q)[r][d] F(q)[eq][d], F(q) interpolated from the nodal fluxes
  for (int q = threadIdx.x; q < a.nqv; q += blockDim.x) {
    double J[9], A[9];
    gen_jacobian(dim, vx, a.xiV + q * dim, J);
    gen_adj(dim, J, A);
    // shape *= ip.weight [* radius]  (domain_integrator.cpp:71-90)
    const double w = ph.axisym ? a.wV[q] * gen_quad_x(vx, a.xiV + q * dim) : a.wV[q];
    const double *ph = a.phiV + static_cast<long long>(q) * dof;
    for (int eq = 0; eq < neq; eq++) {
      double fq[GEN_MAXDIM] = {0, 0, 0};
      for (int k = 0; k < dof; k++)
        for (int d = 0; d < dim; d++) fq[d] += ph[k] * sF[k * nc + eq + d * neq];
      for (int r = 0; r < dim; r++) {
        double g = 0;
        for (int d = 0; d < dim; d++) g += A[r + dim * d] * fq[d];
        sQ[q * nc + eq * dim + r] = w * g;
      }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < dof * neq; t += blockDim.x) {
    const int j = t / neq, eq = t % neq;
    double v = 0;
    for (int q = 0; q < a.nqv; q++) {
      const double *dp = a.dphiV + (static_cast<long long>(q) * dof + j) * dim;
      for (int r = 0; r < dim; r++) v += dp[r] * sQ[q * nc + eq * dim + r];
    }
    sZ[j * neq + eq] = v;
  }
  __syncthreads();
  // faces: Fhat = Rusanov(u1,u2,n) - 1/2 (Fv1 + Fv2).n with n = CalcOrtho of Elem1 (face_integrator.cpp:282-351)
  // One task per (local face, face point), all faces in flight at once; the lift below adds them in local-face order.
  double *sQF = sQ;  // [nfe * nqf][neq]  (the volume values in sQ are consumed)
  {
    for (int it = threadIdx.x; it < a.nfe * a.nqf; it += blockDim.x) {
      const int lf = it / a.nqf, q = it - lf * a.nqf;
      const int f = a.el_face[e * a.nfe + lf];
      const int e1 = a.f_el1[f], e2 = a.f_el2[f];
      const bool bdr = e2 < 0;
      if (bdr && a.f_bc[f] < 0) continue;  // no boundary integrator registered for this attribute
      const bool first = (e1 == e);
      const int code_own = gen_code(dim, first ? a.f_inf1[f] : a.f_inf2[f]);
      const int code_oth = bdr ? code_own : gen_code(dim, first ? a.f_inf2[f] : a.f_inf1[f]);
      const int code1 = gen_code(dim, a.f_inf1[f]);
      const int eo = bdr ? e1 : (first ? e2 : e1);
      const double *v1 = a.vx + static_cast<long long>(e1) * a.nv * dim;
      double *dstq = sQF + static_cast<long long>(it) * neq;
      double J[9], nor[3];
      const double *xi1 = a.xiF + (static_cast<long long>(code1) * a.nqf + q) * dim;
      gen_jacobian(dim, v1, xi1, J);
      gen_face_normal(dim, J, a.dlocF + code1 * dim * (dim - 1), nor);
      const double radius = ph.axisym ? gen_quad_x(v1, xi1) : -1.0;  // Tr.Transform(ip)[0]
      const double *po = a.phiF + (static_cast<long long>(code_own) * a.nqf + q) * dof;
      const double *pn = a.phiF + (static_cast<long long>(code_oth) * a.nqf + q) * dof;
      double uo[GEN_MAXEQ], un[GEN_MAXEQ], go[GEN_MAXEQ * GEN_MAXDIM], gn[GEN_MAXEQ * GEN_MAXDIM];
      if (bdr) {  // BCintegrator::AssembleFaceVector (BCintegrator.cpp:295-441)
        for (int eq = 0; eq < neq; eq++) {
          double x = 0;
          for (int k = 0; k < dof; k++) x += po[k] * sU[eq * dof + k];
          const int sp = eq - nvel - 2;
          uo[eq] = (sp >= 0 && sp < nact) ? fmax(x, 0.0) : x;
        }
        const GenBc &bcf = a.bct.bc[a.f_bc[f]];
        for (int c = 0; c < nc; c++) {  // the Euler boundary fluxes never read the gradients (the non-reflecting ones do)
          double x = 0;
          if (a.eq_system != 0 || bcf.nr >= 0)
            for (int k = 0; k < dof; k++) x += po[k] * sG[c * dof + k];
          go[c] = x;
        }
        double fxb[GEN_MAXEQ];
        for (int eq = 0; eq < neq; eq++) fxb[eq] = 0.0;
        double dwb = 0.0;  // BCintegrator.cpp:408-411
        if (a.dist)
          for (int k = 0; k < dof; k++) dwb += po[k] * a.dist[static_cast<long long>(e1) * dof + k];
        DryAux axb;
        if (a.elem_delta) {  // BCintegrator.cpp:395-411: the face point, the boundary element's delta
          axb.delta = a.elem_delta[e1];
          gen_point(dim, v1, xi1, axb.x);
        }
        if (bcf.nr >= 0)  // every boundary point is visited by exactly one thread: its state is updated in place
          gen_nr_flux(ph, bcf, a.nr[bcf.nr], a.bc_dt, uo, go, nor,
                      a.nr[bcf.nr].boundaryU + static_cast<long long>(a.f_nr_off[f] + q) * neq, fxb);
        else
          gen_bc_flux(ph, bcf, a.bct.use_bc_in_grad, uo, go, nor, radius, fxb, dwb, a.elem_delta ? &axb : nullptr);
        const double sgb = -(ph.axisym ? a.wF[q] * radius : a.wF[q]);  // elvect -= fluxN w [r] shape1
        for (int eq = 0; eq < neq; eq++) dstq[eq] = sgb * fxb[eq];
        continue;
      }
      for (int eq = 0; eq < neq; eq++) {
        double x = 0, y = 0;
        const double *src = gen_elem_field(a, a.U, a.Uhalo, eo, eq, neq);
        for (int k = 0; k < dof; k++) {
          x += po[k] * sU[eq * dof + k];
          y += pn[k] * src[k];
        }
        uo[eq] = x;
        un[eq] = y;
      }
      for (int sp = 0; sp < nact; sp++) {  // face_integrator.cpp:297-301
        uo[nvel + 2 + sp] = fmax(uo[nvel + 2 + sp], 0.0);
        un[nvel + 2 + sp] = fmax(un[nvel + 2 + sp], 0.0);
      }
      if (a.eq_system != 0) {  // gradients enter the viscous fluxes only
        for (int c = 0; c < nc; c++) {
          double x = 0, y = 0;
          const double *src = gen_elem_field(a, a.gradUp, a.gradUpHalo, eo, c, nc);
          for (int k = 0; k < dof; k++) {
            x += po[k] * sG[c * dof + k];
            y += pn[k] * src[k];
          }
          go[c] = x;
          gn[c] = y;
        }
      }
      const double *u1 = first ? uo : un, *u2 = first ? un : uo, *g1 = first ? go : gn, *g2 = first ? gn : go;
      double fx[GEN_MAXEQ];
      gen_riemann(ph, u1, u2, nor, fx);
      if (a.eq_system != 0) {
        double f1[GEN_MAXEQ * GEN_MAXDIM], f2[GEN_MAXEQ * GEN_MAXDIM];
        double dwo = 0.0, dwn = 0.0;  // each side's own interpolation of the wall distance (face_integrator.cpp:304-309)
        if (a.dist)
          for (int k = 0; k < dof; k++) {
            dwo += po[k] * a.dist[static_cast<long long>(e) * dof + k];
            dwn += pn[k] * gen_elem_field(a, a.dist, a.distHalo, eo, 0, 1)[k];
          }
        // same physical point for both sides, each side's own delta (parity trap 5, face_integrator.cpp:251-276, 333-334)
        DryAux ax1, ax2;
        if (a.elem_delta) {
          gen_point(dim, v1, xi1, ax1.x);
          ax2.x[0] = ax1.x[0], ax2.x[1] = ax1.x[1], ax2.x[2] = ax1.x[2];
          ax1.delta = a.elem_delta[e1];
          ax2.delta = a.elem_delta[e2];
        }
        gen_visc_flux(ph, u1, g1, radius, f1, first ? dwo : dwn, a.elem_delta ? &ax1 : nullptr);
        gen_visc_flux(ph, u2, g2, radius, f2, first ? dwn : dwo, a.elem_delta ? &ax2 : nullptr);
        for (int eq = 0; eq < neq; eq++) {
          double v = 0;
          for (int d = 0; d < dim; d++) v += (-0.5 * (f1[eq + d * neq] + f2[eq + d * neq])) * nor[d];
          fx[eq] += v;
        }
      }
      // elvect1 -= phi1 Fhat w ; elvect2 += phi2 Fhat w ; axisymmetric: fluxN *= radius (face_integrator.cpp:344-350)
      const double sg = (first ? -1.0 : 1.0) * (ph.axisym ? a.wF[q] * radius : a.wF[q]);
      for (int eq = 0; eq < neq; eq++) dstq[eq] = sg * fx[eq];
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < dof * neq; t += blockDim.x) {
    const int j = t / neq, eq = t % neq;
    double acc = sZ[j * neq + eq];
    for (int lf = 0; lf < a.nfe; lf++) {
      const int f = a.el_face[e * a.nfe + lf];
      if (a.f_el2[f] < 0 && a.f_bc[f] < 0) continue;
      const int code_own = gen_code(dim, a.f_el1[f] == e ? a.f_inf1[f] : a.f_inf2[f]);
      const double *po = a.phiF + static_cast<long long>(code_own) * a.nqf * dof;
      const double *src = sQF + static_cast<long long>(lf) * a.nqf * neq;
      double v = 0;
      for (int q = 0; q < a.nqf; q++) v += po[q * dof + j] * src[q * neq + eq];
      acc += v;
    }
    sZ[j * neq + eq] = acc;
  }
  __syncthreads();
  gen_apply_minv(a, ph.axisym ? a.me_inv_rad : a.me_inv, e, sZ, neq, a.y, N);
}

// primitives of the received face-neighbour elements (element-major halo copies)
__global__ void gen_prim_halo_kernel(GenArgs a) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(a.NEH) * a.dof) return;
  const long long k = t / a.dof, ln = t % a.dof;
  double s[GEN_MAXEQ], up[GEN_MAXEQ];
  for (int eq = 0; eq < a.neq; eq++) s[eq] = a.Uhalo[(k * a.neq + eq) * a.dof + ln];
  gen_prim(a.phys, s, up);
  double *dst = const_cast<double *>(a.UpHalo);
  for (int eq = 0; eq < a.neq; eq++) dst[(k * a.neq + eq) * a.dof + ln] = up[eq];
}

// SourceTerm::updateTerms (source_term.cpp:62-255): node-wise plasma sources added to y AFTER Me^-1
// (rhs_operator.cpp:451-461).  Usol = the solution grid function U_ the reference reads the conserved state from
// (parity trap 1: in Runge-Kutta stages it differs from the stage vector Up / gradUp were computed from).
__global__ void gen_source_kernel(GenArgs a, const double *Usol) {
  const long long n = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  double Un[GEN_MAXEQ], upn[GEN_MAXEQ], gr[GEN_MAXEQ * GEN_MAXDIM], src[GEN_MAXEQ];
  for (int eq = 0; eq < a.neq; eq++) {
    Un[eq] = Usol[n + eq * a.N];
    upn[eq] = a.Up[n + eq * a.N];
    for (int d = 0; d < a.dim; d++) gr[eq + d * a.neq] = a.gradUp[n + eq * a.N + d * a.neq * a.N];
  }
  mix_source(*a.phys.mix, Un, upn, gr, n, src);
  for (int eq = 0; eq < a.neq; eq++) a.y[n + eq * a.N] += src[eq];
}

// AxisymmetricSource::updateTerms (forcing_terms.cpp:255-380): (p + rho u_t^2 - tau_tt)/r on the r-momentum and
// (-rho u_r u_t + tau_tr)/r on the theta-momentum, node-wise after Me^-1, registered after SourceTerm
// (rhs_operator.cpp:126-160).  Reads U_ (the solution grid function), Up and gradUp; the 1/r is unguarded like the
// reference's (meshes keep their nodes off the axis: Gauss-Legendre nodes are interior).
__global__ void gen_axisym_source_kernel(GenArgs a, const double *Usol) {
  const long long n = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  const int neq = a.neq;
  double U[GEN_MAXEQ], Up[GEN_MAXEQ], gr[GEN_MAXEQ * GEN_MAXDIM];
  for (int eq = 0; eq < neq; eq++) {
    U[eq] = Usol[n + eq * a.N];
    Up[eq] = a.Up[n + eq * a.N];
    for (int d = 0; d < a.dim; d++) gr[eq + d * neq] = a.gradUp[n + eq * a.N + d * neq * a.N];
  }
  const int nact = gen_num_active_species(a.phys);
  for (int sp = 0; sp < nact; sp++) {
    const int eq = 3 + 2 + sp;
    U[eq] = fmax(U[eq], 0.0);
    Up[eq] = fmax(Up[eq], 0.0);
  }
  const long long e = n / a.dof;
  const int k = static_cast<int>(n % a.dof);
  const double radius = gen_quad_x(a.vx + e * a.nv * a.dim, a.xiN + k * a.dim);
  const double rho = Up[0], ur = Up[1], ut = Up[3];
  const double pressure = gen_pressure_from_prim(a.phys, Up);
  const double rurut = rho * ur * ut, rutut = rho * ut * ut;
  double tau_tt, tau_tr;
  if (a.eq_system == 0) {
    tau_tt = tau_tr = 0.0;
  } else {
    const double ur_r = gr[1 + 0 * neq], uz_z = gr[2 + 1 * neq], ut_r = gr[3 + 0 * neq];
    double visc_vec[2];
    gen_viscosities(a.phys, U, Up, visc_vec);
    const double visc = visc_vec[0];
    double bulkVisc = visc_vec[1];
    bulkVisc -= 2. / 3. * visc;
    double divV = ur_r + uz_z;
    if (radius > 0) divV += ur / radius;
    tau_tt = (radius > 0) ? 2.0 * ur / radius * visc : 0.0;
    tau_tt += bulkVisc * divV;
    tau_tr = ut_r;
    if (radius > 0) tau_tr -= ut / radius;
    tau_tr *= visc;
  }
  a.y[n + 1 * a.N] += (pressure + rutut - tau_tt) / radius;
  a.y[n + 3 * a.N] += (-rurut + tau_tr) / radius;
}

// test hook: the per-point physics on arrays of points (point-major: U[i*neq + eq], gradUp[i*neq*dim + eq + d*neq])
__global__ void gen_point_eval_kernel(GenArgs a, int which, int n, const double *U, const double *aux, double *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int neq = a.neq, nc = a.neq * a.dim;
  double s[GEN_MAXEQ], up[GEN_MAXEQ], gr[GEN_MAXEQ * GEN_MAXDIM], f[GEN_MAXEQ * GEN_MAXDIM];
  for (int eq = 0; eq < neq; eq++) s[eq] = U[i * neq + eq];
  if (which == 0) {
    gen_prim(a.phys, s, up);
    for (int eq = 0; eq < neq; eq++) out[i * neq + eq] = up[eq];
  } else if (which == 1) {
    out[i] = gen_max_char_speed(a.phys, s);
  } else if (which == 2) {
    gen_conv_flux(a.phys, s, f);
    for (int c = 0; c < nc; c++) out[i * nc + c] = f[c];
  } else if (which == 3) {
    for (int c = 0; c < nc; c++) gr[c] = aux[i * nc + c];
    gen_visc_flux(a.phys, s, gr, 1.0, f);  // probe radius 1 (axisymmetric terms)
    for (int c = 0; c < nc; c++) out[i * nc + c] = f[c];
  } else if (which == 4 && a.phys.fluid) {
    for (int c = 0; c < nc; c++) gr[c] = aux[i * nc + c];
    gen_prim(a.phys, s, up);
    mix_source(*a.phys.mix, s, up, gr, i, f);
    for (int eq = 0; eq < neq; eq++) out[i * neq + eq] = f[eq];
  }
}

// SourceTerm::updateTerms for the LTE fluid (one species, no reactions): the net-emission radiative sink
// -4 pi eps_N(T) on the total energy (source_term.cpp:205-207, radiation.hpp:57-70), T the nodal primitive of this evaluation
__global__ void gen_lte_source_kernel(GenArgs a) {
  const long long n = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  const double Th = a.Up[n + static_cast<long long>(1 + a.nvel) * a.N];
  a.y[n + static_cast<long long>(1 + a.nvel) * a.N] += -4.0 * 3.14159265358979323846 * lte_eval(*a.phys.lte, LTE_NEC, Th);
}

// InletBC::updateMean / OutletBC::updateMean (inletBC.cpp:482-564, outletBC.cpp:470-561), called by every Mult after
// the gradients (rhs_operator.cpp:364): sum of the primitives interpolated to the patch's face quadrature points, their
// number and (for the first evaluation) the patch area; the first evaluation also initialises the boundary states with
// the conserved form of the interpolated primitives.  One CTA per patch, fixed-order reduction: run-to-run deterministic.
__global__ void __launch_bounds__(256) gen_nr_sum_kernel(GenArgs a) {
  const GenNrPatch &pt = a.nr[blockIdx.x];
  __shared__ double red[256];
  const int neq = a.neq, dim = a.dim, dof = a.dof, npts = pt.nfaces * a.nqf;
  const bool first = pt.init[0] == 0;
  double acc[GEN_MAXEQ + 2];
  for (int i = 0; i < neq + 2; i++) acc[i] = 0.0;
  for (int t = threadIdx.x; t < npts; t += blockDim.x) {
    const int f = pt.faces[t / a.nqf], q = t % a.nqf, e1 = a.f_el1[f], code1 = gen_code(dim, a.f_inf1[f]);
    const double *po = a.phiF + (static_cast<long long>(code1) * a.nqf + q) * dof;
    double up[GEN_MAXEQ], st[GEN_MAXEQ];
    for (int eq = 0; eq < neq; eq++) {
      double x = 0.;
      const double *src = a.Up + static_cast<long long>(e1) * dof + static_cast<long long>(eq) * a.N;
      for (int k = 0; k < dof; k++) x += po[k] * src[k];
      up[eq] = x;
      acc[eq] += x;
    }
    acc[neq] += 1.0;
    if (first) {
      gen_cons(a.phys, up, st);
      for (int eq = 0; eq < neq; eq++) pt.boundaryU[static_cast<long long>(t) * neq + eq] = st[eq];
      double J[9], nor[3], m = 0.;  // BoundaryCondition::aggregateArea (BoundaryCondition.cpp:59-81)
      gen_jacobian(dim, a.vx + static_cast<long long>(e1) * a.nv * dim, a.xiF + (static_cast<long long>(code1) * a.nqf + q) * dim, J);
      gen_face_normal(dim, J, a.dlocF + code1 * dim * (dim - 1), nor);
      for (int d = 0; d < dim; d++) m += nor[d] * nor[d];
      acc[neq + 1] += sqrt(m) * a.wF[q];
    }
  }
  for (int i = 0; i < neq + 2; i++) {
    red[threadIdx.x] = acc[i];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
      if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
      __syncthreads();
    }
    if (threadIdx.x == 0) pt.sums[i] = red[0];
    __syncthreads();
  }
}
// ... after the sums have been all-reduced over the ranks (MPI_Allreduce in the reference): the mean, the area, the flag
__global__ void gen_nr_finish_kernel(GenArgs a) {
  const GenNrPatch &pt = a.nr[blockIdx.x];
  const int neq = a.neq;
  if (threadIdx.x < neq) pt.meanUp[threadIdx.x] = pt.sums[threadIdx.x] * (1. / pt.sums[neq]);
  if (threadIdx.x == 0 && pt.init[0] == 0) {
    pt.area[0] = pt.sums[neq + 1];
    pt.init[0] = 1;
  }
}

inline size_t gen_grad_smem(const GenArgs &a) {
  const int nc = a.neq * a.dim, nqmax = a.nqv > a.nfe * a.nqf ? a.nqv : a.nfe * a.nqf;
  return sizeof(double) * (static_cast<size_t>(a.neq) * a.dof + static_cast<size_t>(a.dof) * nc + static_cast<size_t>(nqmax) * nc + a.nqv);
}
inline size_t gen_resid_smem(const GenArgs &a) {
  const int nc = a.neq * a.dim, nqmax = a.nqv > a.nfe * a.nqf ? a.nqv : a.nfe * a.nqf;
  return sizeof(double) * (static_cast<size_t>(a.neq) * a.dof + 2 * static_cast<size_t>(a.dof) * nc + static_cast<size_t>(nqmax) * nc +
                           static_cast<size_t>(a.dof) * a.neq);
}

}  // namespace tpsb
