// Fast path of RHSoperator::Mult for the headline configuration: 3-D hexahedra whose elements are all
// parallelepipeds (constant Jacobian), dry air, Gauss-Legendre nodes and rules.  Included once by tpsb200.cu.
//
// Three kernels per evaluation (after prim_kernel):
//
//   grad_trace_kernel  per element: BR1 gradient (Gradients::computeGradients, src/gradients.cpp:144-232 +
//                      GradFaceIntegrator, src/faceGradientIntegration.cpp:40-140) AND the face traces the flux
//                      kernel needs, written as one contiguous block per (element, local face):
//                        fields 0-4  trace of the conserved state U
//                        fields 5-9  trace of the normal-contracted viscous fields
//                                    s_i = sum_j (d_j u_i + d_i u_j) n_j, i=0..2 ; n.grad T ; div u
//                      (on a parallelepiped the face normal n is constant, the contraction is linear and
//                      commutes with the trace/interpolation, so 5 viscous fields replace 12 gradient fields)
//   face_flux_fast_kernel  one WARP per face, no CTA barriers: both sides' blocks -> (p+2)^2 quadrature points
//                      -> Rusanov + averaged viscous flux (FaceIntegrator, src/face_integrator.cpp:282-351;
//                      RiemannSolverTPS::Eval_LF, src/riemann_solver.cpp:89-114) -> projected face residual
//   elem_resid_fast_kernel nodal flux, collocated weak divergence, lift of the six face residuals, diagonal
//                      Me^-1, max characteristic speed (src/rhs_operator.cpp:379-391,432-448,493-559)
//
// Why traces are materialised: extrapolating a face trace needs ALL (p+1)^3 nodes of the element, so a
// face-centric kernel that reads U/gradUp pulls 17 x 512 B per side through L2 (8.7 KB) -- the element
// kernel has those values on chip already and emits 1.28 KB per side instead.
#pragma once
#include "rhs_kernels.cuh"

namespace tpsb {

constexpr int NTF = 10;  // trace fields per (element, face) block on the affine path
constexpr int GEO = 12;  // doubles per element in the affine metric table: adj(J) [9], det, 1/det, pad

// local faces at the - / + end of element axis r (MFEM hex face numbering, tables.hpp HEX_FACE_VERT)
__device__ constexpr int kFaceMinus[3] = {4, 1, 0};
__device__ constexpr int kFacePlus[3] = {2, 3, 5};

// ------------------------------------------------------------------------------------------------
template <int NP, int EPB, int MINB>
__global__ void __launch_bounds__(NP *NP *NP *EPB, MINB)
    grad_trace_kernel(KernelArgs a, int elem_begin, int elem_count, const int *elem_list) {
  constexpr int ND = NP * NP * NP, NF2 = NP * NP;
  constexpr int NV = 13;                               // viscous node fields: 3 axes x (s_0..2, n.gradT) + div u
  constexpr int WRK = 2 * NEQ * ND + 6 * NEQ * NF2;    // sU | sUp | sJ ; sV aliases the front
  static_assert(NV * ND <= WRK, "sV overlay");
  __shared__ __align__(16) double sWork[EPB][WRK];
  __shared__ double sGeo[EPB][GEO];
  __shared__ double sD[NP][NP], sLb[2][NP], sWn[NP];
  __shared__ int sNbr[EPB][6], sCode[EPB][6], sFp[6];
  const int le = threadIdx.x / ND, n = threadIdx.x % ND;
  const int slot = blockIdx.x * EPB + le;
  const bool active = slot < elem_count;
  const int e = active ? (elem_list ? elem_list[elem_begin + slot] : elem_begin + slot) : 0;
  const long long N = a.N;
  double *sU = &sWork[le][0], *sUp = sU + NEQ * ND, *sJ = sUp + NEQ * ND, *sV = sU;
  if (threadIdx.x < NP * NP) sD[threadIdx.x / NP][threadIdx.x % NP] = c_T.D[threadIdx.x / NP][threadIdx.x % NP];
  if (threadIdx.x < 2 * NP) sLb[threadIdx.x / NP][threadIdx.x % NP] = c_T.lb[threadIdx.x / NP][threadIdx.x % NP];
  if (threadIdx.x < NP) sWn[threadIdx.x] = c_T.wn[threadIdx.x];
  if (threadIdx.x < 6) sFp[threadIdx.x] = c_T.face_par[threadIdx.x];
  const long long o = static_cast<long long>(e) * ND + n;
  if (active) {
#pragma unroll
    for (int f = 0; f < NEQ; f++) {
      sU[f * ND + n] = a.U[o + f * N];
      sUp[f * ND + n] = a.Up[o + f * N];
    }
    for (int t = n; t < GEO; t += ND) sGeo[le][t] = a.geo[static_cast<long long>(e) * GEO + t];
    for (int t = n; t < 6; t += ND) {
      sNbr[le][t] = a.nbr_elem[e * 6 + t];
      sCode[le][t] = a.nbr_code[e * 6 + t];
    }
  }
  __syncthreads();
  const int i = n % NP, j = (n / NP) % NP, k = n / (NP * NP);
  double rg[NEQ][DIM];  // reference-space gradient: D along each axis (+ face lifts below)
  double *blk0 = a.tr + static_cast<long long>(e) * 6 * (NTF * NF2);
  if (active) {
    double dI[NP], dJ[NP], dK[NP];
#pragma unroll
    for (int m = 0; m < NP; m++) {
      dI[m] = sD[i][m];
      dJ[m] = sD[j][m];
      dK[m] = sD[k][m];
    }
#pragma unroll
    for (int f = 0; f < NEQ; f++) {
      double d0 = 0, d1 = 0, d2 = 0;
#pragma unroll
      for (int m = 0; m < NP; m++) {
        d0 += dI[m] * sUp[f * ND + m + NP * j + NP * NP * k];
        d1 += dJ[m] * sUp[f * ND + i + NP * m + NP * NP * k];
        d2 += dK[m] * sUp[f * ND + i + NP * j + NP * NP * m];
      }
      rg[f][0] = d0;
      rg[f][1] = d1;
      rg[f][2] = d2;
    }
    // traces at the face nodes of all six faces: task = (face, face node)
    for (int t = n; t < 6 * NF2; t += ND) {
      const int lf = t / NF2, ab = t % NF2, fa = ab % NP, fb = ab / NP;
      const FacePar fp = decode_face(sFp[lf]);
      const int b0 = face_node_base<NP>(fp, fa, fb), cs = axis_stride<NP>(fp.an);
      double lbo[NP];
#pragma unroll
      for (int c = 0; c < NP; c++) lbo[c] = sLb[fp.side][c];
      double pT[NEQ];
      double *blk = blk0 + lf * (NTF * NF2) + ab;
#pragma unroll
      for (int f = 0; f < NEQ; f++) {
        double u = 0, p = 0;
#pragma unroll
        for (int c = 0; c < NP; c++) {
          u += lbo[c] * sU[f * ND + b0 + c * cs];
          p += lbo[c] * sUp[f * ND + b0 + c * cs];
        }
        blk[f * NF2] = u;
        pT[f] = p;
      }
      const int nbr = sNbr[le][lf];
      if (nbr < 0) {  // boundary face: Up2 = Up1 unless useBCinGrad (faceGradientIntegration.cpp:96-115)
        if (a.bct.use_bc_in_grad && nbr <= -2) {
          double pbc[NEQ];
          dry_bc_prim_for_gradient(a.bct.bc[-2 - nbr], pT, pbc);
#pragma unroll
          for (int f = 0; f < NEQ; f++) sJ[(lf * NEQ + f) * NF2 + ab] = 0.5 * (pbc[f] - pT[f]);
        } else {
#pragma unroll
          for (int f = 0; f < NEQ; f++) sJ[(lf * NEQ + f) * NF2 + ab] = 0.0;
        }
        continue;
      }
      const int code = sCode[le][lf];
      const FacePar fq = decode_face(sFp[code & 7]);
      int a2, b2;
      apply_perm<NP>(code >> 3, fa, fb, a2, b2);
      const int q0 = face_node_base<NP>(fq, a2, b2), qs = axis_stride<NP>(fq.an);
      const bool local = nbr < a.NE;
      const double *src = local ? a.Up + static_cast<long long>(nbr) * ND + q0
                                : a.UpHalo + static_cast<long long>(nbr - a.NE) * NEQ * ND + q0;
      const long long fstride = local ? N : ND;
#pragma unroll
      for (int f = 0; f < NEQ; f++) {
        const double oth = trace_line<NP>(src + f * fstride, qs, sLb[fq.side], a.vec_ok != 0);
        sJ[(lf * NEQ + f) * NF2 + ab] = 0.5 * (oth - pT[f]);
      }
    }
  }
  __syncthreads();
  double vf[NV];
  if (active) {
    // lift of the face jumps, in reference space: both faces of axis r share the direction adj(J) row r
    // (outward normal = +-A[r,:]), so  rg[f][r] += (l_c(1) J+ - l_c(0) J-) / w_c
    const int idx[3] = {i, j, k};
#pragma unroll
    for (int r = 0; r < DIM; r++) {
      const int c = idx[r];
      const FacePar fm = decode_face(kFacePar(kFaceMinus[r])), fpl = decode_face(kFacePar(kFacePlus[r]));
      const int iam = pick3(fm.as, i, j, k), ibm = pick3(fm.at, i, j, k);
      const int abm = (fm.ss ? iam : NP - 1 - iam) + NP * (fm.st ? ibm : NP - 1 - ibm);
      const int iap = pick3(fpl.as, i, j, k), ibp = pick3(fpl.at, i, j, k);
      const int abp = (fpl.ss ? iap : NP - 1 - iap) + NP * (fpl.st ? ibp : NP - 1 - ibp);
      const double iw = 1.0 / sWn[c];
      const double cm = sLb[0][c] * iw, cp = sLb[1][c] * iw;
#pragma unroll
      for (int f = 0; f < NEQ; f++)
        rg[f][r] += cp * sJ[(kFacePlus[r] * NEQ + f) * NF2 + abp] - cm * sJ[(kFaceMinus[r] * NEQ + f) * NF2 + abm];
    }
    // physical gradient: g[f][d] = sum_r rg[f][r] * inv(J)(r,d), inv(J)(r,d) = A[r + 3 d] / det
    const double *A = sGeo[le];
    const double idet = sGeo[le][10];
    double g[NEQ][DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      const double a0 = A[0 + 3 * d] * idet, a1 = A[1 + 3 * d] * idet, a2 = A[2 + 3 * d] * idet;
#pragma unroll
      for (int f = 0; f < NEQ; f++) {
        g[f][d] = rg[f][0] * a0 + rg[f][1] * a1 + rg[f][2] * a2;
        a.gradUp[o + (f + d * NEQ) * N] = g[f][d];
      }
    }
    // normal-contracted viscous fields for the three axis directions n^r = A[r,:] (outward at the + face)
    double sym[DIM][DIM];
#pragma unroll
    for (int p = 0; p < DIM; p++)
#pragma unroll
      for (int q = 0; q < DIM; q++) sym[p][q] = g[1 + p][q] + g[1 + q][p];
#pragma unroll
    for (int r = 0; r < DIM; r++) {
      const double n0 = A[r + 0], n1 = A[r + 3], n2 = A[r + 6];
#pragma unroll
      for (int p = 0; p < DIM; p++) vf[r * 4 + p] = sym[p][0] * n0 + sym[p][1] * n1 + sym[p][2] * n2;
      vf[r * 4 + 3] = g[4][0] * n0 + g[4][1] * n1 + g[4][2] * n2;
    }
    vf[12] = g[1][0] + g[2][1] + g[3][2];
  }
  __syncthreads();  // every read of sU / sUp / sJ is done: sV may overwrite them
  if (active) {
#pragma unroll
    for (int v = 0; v < NV; v++) sV[v * ND + n] = vf[v];
  }
  __syncthreads();
  if (active) {
    for (int t = n; t < 6 * NF2; t += ND) {
      const int lf = t / NF2, ab = t % NF2, fa = ab % NP, fb = ab / NP;
      const FacePar fp = decode_face(sFp[lf]);
      const int b0 = face_node_base<NP>(fp, fa, fb), cs = axis_stride<NP>(fp.an);
      const double sg = fp.side ? 1.0 : -1.0;  // outward normal of this face = sg * A[an,:]
      double lbo[NP];
#pragma unroll
      for (int c = 0; c < NP; c++) lbo[c] = sLb[fp.side][c];
      double *blk = blk0 + lf * (NTF * NF2) + ab;
      const double *row = sV + (fp.an * 4) * ND + b0;
#pragma unroll
      for (int v = 0; v < 4; v++) {
        double s = 0;
#pragma unroll
        for (int c = 0; c < NP; c++) s += lbo[c] * row[v * ND + c * cs];
        blk[(NEQ + v) * NF2] = sg * s;
      }
      double s = 0;
#pragma unroll
      for (int c = 0; c < NP; c++) s += lbo[c] * sV[12 * ND + b0 + c * cs];
      blk[(NEQ + 4) * NF2] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// FP64 tensor-core contraction (DMMA, SASS DMMA.884): D(8x8) = A(8x4) B(4x8).  Fragments (PTX ISA, m8n8k4 .f64):
// A[row = lane/4][k = lane%4], B[k = lane%4][col = lane/4], D[row = lane/4][col = 2 (lane%4) + {0,1}].
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1)
               : "d"(a), "d"(b), "d"(0.0), "d"(0.0));
}

// grad_trace_mma_kernel (p = 3): same result as grad_trace_kernel<4,...>, but every contraction along an element
// axis -- the collocation derivative D (4 rows) AND the two end-point extrapolations lb (2 rows) -- is one DMMA
// with A = [D; lb(0); lb(1); 0; 0] and B = 8 lines of 4 nodal values.  Why: the one-thread-per-node form reads one
// shared-memory double per DFMA and ncu shows it bound by shared-memory wavefronts (l1tex data pipe 93 %, FP64
// pipe 17 %, profiles/r1h_*); a DMMA reads one double per lane for 8 FMAs and the B fragments below are
// bank-conflict free, so the shared-memory traffic of the contractions drops ~6x at the same FP64-pipe cost
// (B200: DMMA and DFMA share one pipe, 37 TF each, tools/ubench/fp64_pipes.cu).
//   node p(n) = n + 4 (n / 16): planes of 16 nodes padded to 20 doubles -> the three B-fragment patterns
//   (x: 32 consecutive nodes, y: 4x4 transposed inside a plane, z: stride one plane) hit 16 distinct banks per
//   half warp.
// Two warps per element (lane = node for the per-node phases), EPB elements per CTA.
__device__ __forceinline__ int pad_node(int n) { return n + 4 * (n >> 4); }

template <int EPB, int MINB>
__global__ void __launch_bounds__(64 * EPB, MINB)
    grad_trace_mma_kernel(KernelArgs a, int elem_begin, int elem_count, const int *elem_list) {
  constexpr int NP = 4, ND = 64, NF2 = 16, PS = 80;  // PS: padded doubles per field
  __shared__ __align__(16) double sF[EPB][2 * NEQ * PS];    // U (0-4), Up (5-9); later the viscous traces [6][5][16]
  __shared__ __align__(16) double sDr[EPB][NEQ * DIM * PS];  // reference derivatives of Up; later the 13 viscous fields
  __shared__ __align__(16) double sJ[EPB][6 * NEQ * NF2];    // own Up traces -> jumps 1/2 (Up_nbr - Up_own)
  __shared__ __align__(16) double sTU[EPB][6 * NEQ * NF2];   // traces of the conserved state
  __shared__ double sGeo[EPB][GEO];
  __shared__ double sD[NP][NP], sLb[2][NP], sWn[NP];
  __shared__ int sNbr[EPB][6], sCode[EPB][6], sFp[6];
  const int le = threadIdx.x / ND, n = threadIdx.x % ND;
  const int lane = threadIdx.x & 31, half = (threadIdx.x >> 5) & 1;
  const int slot = blockIdx.x * EPB + le;
  const bool active = slot < elem_count;
  const int e = active ? (elem_list ? elem_list[elem_begin + slot] : elem_begin + slot) : 0;
  const long long N = a.N;
  if (threadIdx.x < NP * NP) sD[threadIdx.x / NP][threadIdx.x % NP] = c_T.D[threadIdx.x / NP][threadIdx.x % NP];
  if (threadIdx.x < 2 * NP) sLb[threadIdx.x / NP][threadIdx.x % NP] = c_T.lb[threadIdx.x / NP][threadIdx.x % NP];
  if (threadIdx.x < NP) sWn[threadIdx.x] = c_T.wn[threadIdx.x];
  if (threadIdx.x < 6) sFp[threadIdx.x] = c_T.face_par[threadIdx.x];
  const long long o = static_cast<long long>(e) * ND + n;
  const int pn = pad_node(n);
  if (active) {
#pragma unroll
    for (int f = 0; f < NEQ; f++) {
      sF[le][f * PS + pn] = a.U[o + f * N];
      sF[le][(NEQ + f) * PS + pn] = a.Up[o + f * N];
    }
    for (int t = n; t < GEO; t += ND) sGeo[le][t] = a.geo[static_cast<long long>(e) * GEO + t];
    for (int t = n; t < 6; t += ND) {
      sNbr[le][t] = a.nbr_elem[e * 6 + t];
      sCode[le][t] = a.nbr_code[e * 6 + t];
    }
  }
  __syncthreads();
  // A fragment: rows 0-3 = D, row 4 = lb(xi = 0), row 5 = lb(xi = 1)
  const int fr = lane >> 2, fk = lane & 3;
  const double afrag = fr < NP ? sD[fr][fk] : (fr < NP + 2 ? sLb[fr - NP][fk] : 0.0);
  // per-axis fragment addressing for this lane.  Line L (0..15) of axis d: nodes base_d(L) + m stride_d.
  //   B fragment : m = lane%4, L = 8 half + lane/4
  //   D fragment : row = lane/4, lines L0 = 8 half + 2 (lane%4), L0 + 1
  int bOff[DIM], dOff0[DIM], dOff1[DIM], tOff0[DIM], tOff1[DIM];
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    const int str = d == 0 ? 1 : (d == 1 ? NP : NP * NP);
    auto base = [d](int L) { return d == 0 ? 4 * L : (d == 1 ? (L & 3) + 16 * (L >> 2) : L); };
    bOff[d] = pad_node(base(8 * half + fr) + fk * str);
    const int L0 = 8 * half + 2 * fk;
    dOff0[d] = pad_node(base(L0) + (fr & 3) * str);
    dOff1[d] = pad_node(base(L0 + 1) + (fr & 3) * str);
    // face node of line L on the -/+ face of axis d (rows 4 / 5)
    const FacePar fp = (fr & 1) ? decode_face(kFacePar(kFacePlus[d])) : decode_face(kFacePar(kFaceMinus[d]));
    auto abof = [&](int L) {
      const int nb = base(L), i = nb & 3, j = (nb >> 2) & 3, k = nb >> 4;
      const int ia = pick3(fp.as, i, j, k), ib = pick3(fp.at, i, j, k);
      return (fp.ss ? ia : NP - 1 - ia) + NP * (fp.st ? ib : NP - 1 - ib);
    };
    const int lf = (fr & 1) ? kFacePlus[d] : kFaceMinus[d];
    tOff0[d] = lf * (NEQ * NF2) + abof(L0);
    tOff1[d] = lf * (NEQ * NF2) + abof(L0 + 1);
  }
  if (active) {
#pragma unroll
    for (int d = 0; d < DIM; d++) {
#pragma unroll
      for (int f = 0; f < 2 * NEQ; f++) {
        double d0, d1;
        dmma884(d0, d1, afrag, sF[le][f * PS + bOff[d]]);
        if (f >= NEQ && fr < NP) {
          sDr[le][((f - NEQ) * DIM + d) * PS + dOff0[d]] = d0;
          sDr[le][((f - NEQ) * DIM + d) * PS + dOff1[d]] = d1;
        }
        if (fr == NP || fr == NP + 1) {
          double *dst = f < NEQ ? &sTU[le][f * NF2] : &sJ[le][(f - NEQ) * NF2];
          dst[tOff0[d]] = d0;
          dst[tOff1[d]] = d1;
        }
      }
    }
  }
  __syncthreads();
  if (active) {
    // jumps at the face nodes of all six faces: task = (face, face node)
    for (int t = n; t < 6 * NF2; t += ND) {
      const int lf = t / NF2, ab = t % NF2, fa = ab % NP, fb = ab / NP;
      double pT[NEQ];
#pragma unroll
      for (int f = 0; f < NEQ; f++) pT[f] = sJ[le][(lf * NEQ + f) * NF2 + ab];
      const int nbr = sNbr[le][lf];
      if (nbr < 0) {  // boundary face: Up2 = Up1 unless useBCinGrad (faceGradientIntegration.cpp:96-115)
        if (a.bct.use_bc_in_grad && nbr <= -2) {
          double pbc[NEQ];
          dry_bc_prim_for_gradient(a.bct.bc[-2 - nbr], pT, pbc);
#pragma unroll
          for (int f = 0; f < NEQ; f++) sJ[le][(lf * NEQ + f) * NF2 + ab] = 0.5 * (pbc[f] - pT[f]);
        } else {
#pragma unroll
          for (int f = 0; f < NEQ; f++) sJ[le][(lf * NEQ + f) * NF2 + ab] = 0.0;
        }
        continue;
      }
      const int code = sCode[le][lf];
      const FacePar fq = decode_face(sFp[code & 7]);
      int a2, b2;
      apply_perm<NP>(code >> 3, fa, fb, a2, b2);
      const int q0 = face_node_base<NP>(fq, a2, b2), qs = axis_stride<NP>(fq.an);
      const bool local = nbr < a.NE;
      const double *src = local ? a.Up + static_cast<long long>(nbr) * ND + q0
                                : a.UpHalo + static_cast<long long>(nbr - a.NE) * NEQ * ND + q0;
      const long long fstride = local ? N : ND;
#pragma unroll
      for (int f = 0; f < NEQ; f++) {
        const double oth = trace_line<NP>(src + f * fstride, qs, sLb[fq.side], a.vec_ok != 0);
        sJ[le][(lf * NEQ + f) * NF2 + ab] = 0.5 * (oth - pT[f]);
      }
    }
  }
  __syncthreads();
  const int i = n % NP, j = (n / NP) % NP, k = n / (NP * NP);
  if (active) {
    double rg[NEQ][DIM];
#pragma unroll
    for (int f = 0; f < NEQ; f++)
#pragma unroll
      for (int r = 0; r < DIM; r++) rg[f][r] = sDr[le][(f * DIM + r) * PS + pn];
    // lift of the face jumps in reference space (see grad_trace_kernel)
    const int idx[3] = {i, j, k};
#pragma unroll
    for (int r = 0; r < DIM; r++) {
      const int c = idx[r];
      const FacePar fm = decode_face(kFacePar(kFaceMinus[r])), fpl = decode_face(kFacePar(kFacePlus[r]));
      const int iam = pick3(fm.as, i, j, k), ibm = pick3(fm.at, i, j, k);
      const int abm = (fm.ss ? iam : NP - 1 - iam) + NP * (fm.st ? ibm : NP - 1 - ibm);
      const int iap = pick3(fpl.as, i, j, k), ibp = pick3(fpl.at, i, j, k);
      const int abp = (fpl.ss ? iap : NP - 1 - iap) + NP * (fpl.st ? ibp : NP - 1 - ibp);
      const double iw = 1.0 / sWn[c];
      const double cm = sLb[0][c] * iw, cp = sLb[1][c] * iw;
#pragma unroll
      for (int f = 0; f < NEQ; f++)
        rg[f][r] += cp * sJ[le][(kFacePlus[r] * NEQ + f) * NF2 + abp] - cm * sJ[le][(kFaceMinus[r] * NEQ + f) * NF2 + abm];
    }
    const double *A = sGeo[le];
    const double idet = sGeo[le][10];
    double g[NEQ][DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      const double a0 = A[0 + 3 * d] * idet, a1 = A[1 + 3 * d] * idet, a2 = A[2 + 3 * d] * idet;
#pragma unroll
      for (int f = 0; f < NEQ; f++) {
        g[f][d] = rg[f][0] * a0 + rg[f][1] * a1 + rg[f][2] * a2;
        a.gradUp[o + (f + d * NEQ) * N] = g[f][d];
      }
    }
    double sym[DIM][DIM];
#pragma unroll
    for (int p = 0; p < DIM; p++)
#pragma unroll
      for (int q = 0; q < DIM; q++) sym[p][q] = g[1 + p][q] + g[1 + q][p];
    // thread n only ever touches column pn of sDr, so the viscous fields can overwrite it in place
    double *sV = sDr[le];
#pragma unroll
    for (int r = 0; r < DIM; r++) {
      const double n0 = A[r + 0], n1 = A[r + 3], n2 = A[r + 6];
#pragma unroll
      for (int p = 0; p < DIM; p++) sV[(r * 4 + p) * PS + pn] = sym[p][0] * n0 + sym[p][1] * n1 + sym[p][2] * n2;
      sV[(r * 4 + 3) * PS + pn] = g[4][0] * n0 + g[4][1] * n1 + g[4][2] * n2;
    }
    sV[12 * PS + pn] = g[1][0] + g[2][1] + g[3][2];
  }
  __syncthreads();
  if (active) {
    // traces of the normal-contracted viscous fields: axis d carries fields 4 d .. 4 d + 3 and div u (12)
    const double *sV = sDr[le];
    double *sTV = sF[le];  // U / Up are dead: [6][5][16]
    const double sg = (fr & 1) ? 1.0 : -1.0;  // outward normal of the -/+ face = -/+ A[d,:]
#pragma unroll
    for (int d = 0; d < DIM; d++) {
#pragma unroll
      for (int v = 0; v < 5; v++) {
        double d0, d1;
        dmma884(d0, d1, afrag, sV[(v < 4 ? 4 * d + v : 12) * PS + bOff[d]]);
        if (fr == NP || fr == NP + 1) {
          const double s = v < 4 ? sg : 1.0;
          sTV[v * NF2 + tOff0[d]] = s * d0;
          sTV[v * NF2 + tOff1[d]] = s * d1;
        }
      }
    }
  }
  __syncthreads();
  if (active) {
    // one contiguous 7.5 KB block per element: [face][U traces 0-4 | viscous traces 5-9][face node]
    double2 *dst = reinterpret_cast<double2 *>(a.tr + static_cast<long long>(e) * 6 * (NTF * NF2));
    const double2 *su = reinterpret_cast<const double2 *>(sTU[le]), *sv = reinterpret_cast<const double2 *>(sF[le]);
    constexpr int HB = NEQ * NF2 / 2;  // double2 per (face, U | viscous) half block
#pragma unroll
    for (int t = n; t < 6 * HB; t += ND) {
      const int lf = t / HB;
      dst[t + HB * lf] = su[t];
      dst[t + HB * lf + HB] = sv[t];
    }
  }
}

// gather the trace blocks of the shared faces into the send buffer (one block per shared face, in the
// order the receiver lists its shared faces with this peer)
__global__ void pack_blocks_kernel(int nblk, int blk_doubles, const int *src_blk, const double *tr, double *dst) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(nblk) * blk_doubles) return;
  const int b = static_cast<int>(t / blk_doubles), w = static_cast<int>(t % blk_doubles);
  dst[t] = tr[static_cast<long long>(src_blk[b]) * blk_doubles + w];
}

// ------------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy helpers (sm_90+ PTX; SASS: UBLKCP / SYNCS)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// L2 prefetch of a block two faces ahead (no shared-memory buffer needed): the bulk copy issued one face ahead then hits L2
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D_%=;\n"
      "bra W_%=;\n"
      "D_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// NP contiguous doubles from shared memory through explicit ld.shared PTX (16-byte vectors when NP is even;
// p is 16-byte aligned by construction).  Inline PTX, because nvcc 12.9 mis-optimised the plain C++ loads of
// the unrolled row loop below: it fed v[0] with Y[161] in place of Y[1] (PTX: ld.shared.v2.f64 [..+6400]
// paired with the scalar load at [..+5120]) -- found by the parity test, kept out of the optimiser's reach here.
template <int NP>
__device__ __forceinline__ void load_row(const double *p, double *out) {
  const unsigned addr = smem_u32(p);
  if constexpr (NP % 2 == 0) {
#pragma unroll
    for (int q = 0; q < NP / 2; q++)
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(out[2 * q]), "=d"(out[2 * q + 1]) : "r"(addr + 16 * q) : "memory");
  } else {
#pragma unroll
    for (int q = 0; q < NP; q++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(out[q]) : "r"(addr + 8 * q) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// face_flux_fast_kernel: one WARP per face, grid-stride over faces, no CTA barriers in the loop.
//   raw[2] : the two sides' trace blocks of the current / next face, filled by TMA bulk copies
//            (cp.async.bulk + mbarrier) issued one face ahead by lane 0
//   Y      : values interpolated along a, [2*NTF][NQ][NP]
//   F, B   : weighted fluxes at the quadrature points / half-projected fluxes
// Phases: (1) interpolation along a, task (side*NTF+field, b); side 2 is read through the orientation
// permutation.  (2) one lane per quadrature point (alpha,beta): finishes the interpolation along b in
// registers (4 FMAs per field with its own row of P) and evaluates the numerical flux.  (3,4) projection
// back onto the (p+1)^2 face nodes.
template <int NP, int WPB, int MINB>
__global__ void __launch_bounds__(32 * WPB, MINB) face_flux_fast_kernel(KernelArgs a, int face_begin, int face_count) {
  constexpr int NF2 = NP * NP, NQ = NP + 1, NQ2 = NQ * NQ;
  constexpr int RAW = 2 * NTF * NF2, YS = 2 * NTF * NP * NQ;
  constexpr int FS = ((NEQ * NQ2 + 3) / 4) * 4, BS = NEQ * NQ * NP;
  constexpr int PER_WARP = 2 * RAW + YS + FS + BS;
  constexpr unsigned BLKB = NTF * NF2 * sizeof(double);
  static_assert(NQ2 <= 32, "one lane per quadrature point");
  static_assert(BLKB % 16 == 0, "bulk copy granularity");
  extern __shared__ __align__(128) double sDyn[];
  __shared__ __align__(8) unsigned long long sBar[WPB][2];
  __shared__ double sWq[NQ2], sP[NQ][NP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < NQ2) sWq[threadIdx.x] = c_T.wq[threadIdx.x % NQ] * c_T.wq[threadIdx.x / NQ];
  if (threadIdx.x < NQ * NP) sP[threadIdx.x / NP][threadIdx.x % NP] = c_T.P[threadIdx.x / NP][threadIdx.x % NP];
  double *raw = sDyn + warp * PER_WARP, *Y = raw + 2 * RAW, *F = Y + YS, *B = F + FS;
  const unsigned bar0 = smem_u32(&sBar[warp][0]), bar1 = smem_u32(&sBar[warp][1]);
  if (lane == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const double wq = lane < NQ2 ? sWq[lane] : 0.0;
  const int qa = lane % NQ, qb = lane < NQ2 ? lane / NQ : 0;
  double pb[NP];  // this lane's row of P for the interpolation along b
#pragma unroll
  for (int q = 0; q < NP; q++) pb[q] = sP[qb][q];
  const PhysParams ph = a.phys;
  const int stride = gridDim.x * WPB;
  int fi = blockIdx.x * WPB + warp;
  // descriptors are fetched two faces ahead, blocks one face ahead
  int4 fd_cur = make_int4(0, 0, 0, 0), fd_nxt = make_int4(0, 0, 0, 0);
  if (fi < face_count) fd_cur = __ldg(&a.face_desc[face_begin + fi]);
  if (fi + stride < face_count) fd_nxt = __ldg(&a.face_desc[face_begin + fi + stride]);
  if (fi < face_count && lane == 0) {
    mbar_expect_tx(bar0, 2 * BLKB);
    bulk_g2s(smem_u32(raw), a.tr + static_cast<long long>(fd_cur.x) * (NTF * NF2), BLKB, bar0);
    bulk_g2s(smem_u32(raw + NTF * NF2), a.tr + static_cast<long long>(fd_cur.y) * (NTF * NF2), BLKB, bar0);
  }
  for (int it = 0; fi < face_count; it++, fi += stride) {
    const int buf = it & 1;
    const unsigned ph_bit = (it >> 1) & 1;
    const int fc = face_begin + fi;
    double *cur = raw + buf * RAW;
    // prefetch: blocks of the next face into the other buffer, descriptor of the one after
    int4 fd_n2 = make_int4(0, 0, 0, 0);
    if (fi + 2 * stride < face_count) fd_n2 = __ldg(&a.face_desc[fc + 2 * stride]);
    if (fi + stride < face_count && lane == 0) {
      const unsigned nb = buf ? bar0 : bar1;
      double *nx = raw + (buf ^ 1) * RAW;
      mbar_expect_tx(nb, 2 * BLKB);
      bulk_g2s(smem_u32(nx), a.tr + static_cast<long long>(fd_nxt.x) * (NTF * NF2), BLKB, nb);
      bulk_g2s(smem_u32(nx + NTF * NF2), a.tr + static_cast<long long>(fd_nxt.y) * (NTF * NF2), BLKB, nb);
    }
    const double2 nrm01 = __ldg(reinterpret_cast<const double2 *>(a.face_nor) + 2 * fc);
    const double2 nrm23 = __ldg(reinterpret_cast<const double2 *>(a.face_nor) + 2 * fc + 1);
    const int pc = fd_cur.z;  // perm code: face coords -> Elem2 local face coords
    mbar_wait(buf ? bar1 : bar0, ph_bit);
    // (1) interpolation along a: Y[sf][alpha][b] = sum_a P[alpha][a] T[sf](a, b)
    for (int t = lane; t < 2 * NTF * NP; t += 32) {
      const int sf = t / NP, b = t % NP;
      int i0 = NP * b, di = 1;
      if (sf >= NTF) {  // side 2: (a,b) -> Elem2 coords; along a the image is a line of stride +-1 or +-NP
        const int fa2 = (pc & 2) ? 1 : 0, fb2 = (pc & 4) ? 1 : 0;
        if (pc & 1) {
          i0 = (fa2 ? NP - 1 - b : b) + (fb2 ? NP * (NP - 1) : 0);
          di = fb2 ? -NP : NP;
        } else {
          i0 = (fa2 ? NP - 1 : 0) + NP * (fb2 ? NP - 1 - b : b);
          di = fa2 ? -1 : 1;
        }
      }
      const double *src = cur + sf * NF2 + i0;
      double in[NP];
#pragma unroll
      for (int q = 0; q < NP; q++) in[q] = src[q * di];
#pragma unroll
      for (int al = 0; al < NQ; al++) {
        double v = 0;
#pragma unroll
        for (int q = 0; q < NP; q++) v += sP[al][q] * in[q];
        Y[sf * (NP * NQ) + al * NP + b] = v;
      }
    }
    __syncwarp();
    // (2) quadrature point (qa, qb): finish the interpolation and evaluate the flux
    if (lane < NQ2) {
      double v[2 * NTF];
#pragma unroll
      for (int sf = 0; sf < 2 * NTF; sf++) {
        double row[NP];
        load_row<NP>(Y + sf * (NP * NQ) + qa * NP, row);
        double s = 0;
#pragma unroll
        for (int q = 0; q < NP; q++) s += pb[q] * row[q];
        v[sf] = s;
      }
      const double *u1 = v, *u2 = v + NTF;
      const double n0 = nrm01.x, n1 = nrm01.y, n2 = nrm23.x, normag = nrm23.y;
      // DryAir state (equation_of_state.cpp:279-335) with 1/rho formed once
      const double ri1 = fast_rcp(u1[0]), ri2 = fast_rcp(u2[0]);
      const double vx1 = u1[1] * ri1, vy1 = u1[2] * ri1, vz1 = u1[3] * ri1;
      const double vx2 = u2[1] * ri2, vy2 = u2[2] * ri2, vz2 = u2[3] * ri2;
      const double vv1 = vx1 * vx1 + vy1 * vy1 + vz1 * vz1, vv2 = vx2 * vx2 + vy2 * vy2 + vz2 * vz2;
      const double p1 = ph.gm1 * (u1[4] - 0.5 * u1[0] * vv1), p2 = ph.gm1 * (u2[4] - 0.5 * u2[0] * vv2);
      const double pr1 = p1 * ri1, pr2 = p2 * ri2;  // p / rho = R T
      const double maxE = fmax(fast_sqrt(vv1) + fast_sqrt(ph.gamma * pr1), fast_sqrt(vv2) + fast_sqrt(ph.gamma * pr2));
      // Rusanov (riemann_solver.cpp:89-114)
      const double vn1 = vx1 * n0 + vy1 * n1 + vz1 * n2, vn2 = vx2 * n0 + vy2 * n1 + vz2 * n2;
      const double diss = 0.5 * maxE * normag;
      const double psum = 0.5 * (p1 + p2);
      double fx[NEQ];
      fx[0] = 0.5 * (u1[0] * vn1 + u2[0] * vn2) - diss * (u2[0] - u1[0]);
      fx[1] = 0.5 * (u1[1] * vn1 + u2[1] * vn2) + psum * n0 - diss * (u2[1] - u1[1]);
      fx[2] = 0.5 * (u1[2] * vn1 + u2[2] * vn2) + psum * n1 - diss * (u2[2] - u1[2]);
      fx[3] = 0.5 * (u1[3] * vn1 + u2[3] * vn2) + psum * n2 - diss * (u2[3] - u1[3]);
      fx[4] = 0.5 * ((u1[4] + p1) * vn1 + (u2[4] + p2) * vn2) - diss * (u2[4] - u1[4]);
      if (ph.eq_system != 0) {  // - 1/2 (Fv1 + Fv2).n  (face_integrator.cpp:331-341); side 2 carries n2 = -n
        // Sutherland (transport_properties.cpp:223-234): mu = C1 mult T^1.5 / (T + S0)
        const double iR = 1.0 / ph.R;
        const double T1 = pr1 * iR, T2 = pr2 * iR;
        const double mu1 = ph.C1 * ph.visc_mult * (T1 * fast_sqrt(T1)) * fast_rcp(T1 + ph.S0);
        const double mu2 = ph.C1 * ph.visc_mult * (T2 * fast_sqrt(T2)) * fast_rcp(T2 + ph.S0);
        const double bf = ph.bulk_visc_mult - 2. / 3.;
        const double bd1 = bf * mu1 * u1[NEQ + 4], bd2 = -bf * mu2 * u2[NEQ + 4];
        const double t0 = mu1 * u1[NEQ + 0] + bd1 * n0 - (mu2 * u2[NEQ + 0] + bd2 * n0);
        const double t1 = mu1 * u1[NEQ + 1] + bd1 * n1 - (mu2 * u2[NEQ + 1] + bd2 * n1);
        const double t2 = mu1 * u1[NEQ + 2] + bd1 * n2 - (mu2 * u2[NEQ + 2] + bd2 * n2);
        const double e1 = vx1 * (mu1 * u1[NEQ + 0] + bd1 * n0) + vy1 * (mu1 * u1[NEQ + 1] + bd1 * n1) +
                          vz1 * (mu1 * u1[NEQ + 2] + bd1 * n2) + ph.cp_div_pr * mu1 * u1[NEQ + 3];
        const double e2 = vx2 * (mu2 * u2[NEQ + 0] + bd2 * n0) + vy2 * (mu2 * u2[NEQ + 1] + bd2 * n1) +
                          vz2 * (mu2 * u2[NEQ + 2] + bd2 * n2) + ph.cp_div_pr * mu2 * u2[NEQ + 3];
        fx[1] -= 0.5 * t0;
        fx[2] -= 0.5 * t1;
        fx[3] -= 0.5 * t2;
        fx[4] -= 0.5 * (e1 - e2);
      }
#pragma unroll
      for (int eq = 0; eq < NEQ; eq++) F[eq * NQ2 + lane] = fx[eq] * wq;
    }
    __syncwarp();
    // (3) projection along beta: B[eq][alpha][b] = sum_beta P[beta][b] F[eq][alpha + NQ beta]
    for (int t = lane; t < NEQ * NQ; t += 32) {
      const int eq = t / NQ, al = t % NQ;
      double in[NQ];
#pragma unroll
      for (int q = 0; q < NQ; q++) in[q] = F[eq * NQ2 + al + NQ * q];
#pragma unroll
      for (int b = 0; b < NP; b++) {
        double s = 0;
#pragma unroll
        for (int q = 0; q < NQ; q++) s += sP[q][b] * in[q];
        B[t * NP + b] = s;
      }
    }
    __syncwarp();
    // (4) ... and along alpha, straight to global: R[eq][a + NP b] = sum_alpha P[alpha][a] B[eq][alpha][b]
    for (int t = lane; t < NEQ * NP; t += 32) {
      const int eq = t / NP, b = t % NP;
      double in[NQ];
#pragma unroll
      for (int q = 0; q < NQ; q++) in[q] = B[(eq * NQ + q) * NP + b];
      double *dst = a.faceRes + (static_cast<long long>(fc) * NEQ + eq) * NF2 + NP * b;
#pragma unroll
      for (int aa = 0; aa < NP; aa++) {
        double s = 0;
#pragma unroll
        for (int q = 0; q < NQ; q++) s += sP[q][aa] * in[q];
        dst[aa] = s;
      }
    }
    __syncwarp();  // all reads of cur / Y / F / B done before the next iteration overwrites them
    fd_cur = fd_nxt;
    fd_nxt = fd_n2;
  }
}

// ------------------------------------------------------------------------------------------------
// face_flux_mma_kernel (p = 3): face_flux_fast_kernel with every tensor contraction on the FP64 tensor cores.
// ncu (profiles/r1h_*): the DFMA form is bound by shared-memory wavefronts (l1tex data pipe 96 %, a third of them
// bank conflicts of the stride-4 row reads, another fifth broadcast reads of P); as DMMA fragments the same data is
// read once per 8 FMAs from conflict-free, lane-contiguous addresses.  One WARP per face as before:
//   A  Y[sf][al][b]  = sum_a  P[al][a] T[sf][a + 4 b]     10 DMMA   (A = P padded to 8 rows; side 2 permuted)
//   B  V[q][sf]      = sum_b  P[be][b] Y[sf][al][b]       13 DMMA   q = al + 5 be
//   C  one lane per quadrature point: Rusanov + averaged viscous flux (identical to face_flux_fast_kernel)
//   D  Bq[eq][al][b] = sum_be P[be][b] F[eq][al + 5 be]    4 x 2 DMMA (K = 5 padded to 8, A = P^T)
//   E  R[eq][a + 4b] = sum_al P[al][a] Bq[eq][al][b]       3 x 2 DMMA -> global
__device__ __forceinline__ void dmma884_acc(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

template <int WPB, int MINB>
__global__ void __launch_bounds__(32 * WPB, MINB) face_flux_mma_kernel(KernelArgs a, int face_begin, int face_count) {
  constexpr int NP = 4, NF2 = 16, NQ = 5, NQ2 = 25;
  constexpr int RAW = 2 * NTF * NF2;            // 320: both sides' trace blocks
  constexpr int YS = 104 * NP;                  // Y: 100 columns (sf, al) x 4, padded to 13 groups of 8
  constexpr int VS = NQ2 * 21 + 3;              // V[q][21]: q-major, odd stride
  constexpr int PER_WARP = 2 * RAW + YS + VS;   // 1584 doubles
  constexpr unsigned BLKB = NTF * NF2 * sizeof(double);
  extern __shared__ __align__(128) double sDyn[];
  __shared__ __align__(8) unsigned long long sBar[WPB][2];
  __shared__ double sWq[NQ2], sP[NQ][NP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < NQ2) sWq[threadIdx.x] = c_T.wq[threadIdx.x % NQ] * c_T.wq[threadIdx.x / NQ];
  if (threadIdx.x < NQ * NP) sP[threadIdx.x / NP][threadIdx.x % NP] = c_T.P[threadIdx.x / NP][threadIdx.x % NP];
  double *raw = sDyn + warp * PER_WARP, *Y = raw + 2 * RAW, *V = Y + YS;
  double *F = Y, *Bq = Y + 128;  // Y is dead once V is complete
  const unsigned bar0 = smem_u32(&sBar[warp][0]), bar1 = smem_u32(&sBar[warp][1]);
  if (lane == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int fr = lane >> 2, fk = lane & 3;
  // A fragments: interpolation P (5 x 4, rows 5-7 zero) and projection P^T (4 x 5 split into k = 0-3 and k = 4)
  const double aP = fr < NQ ? sP[fr][fk] : 0.0;
  const double aT0 = fr < NP ? sP[fk][fr] : 0.0;
  const double aT1 = (fr < NP && fk == 0) ? sP[4][fr] : 0.0;
  for (int t = lane; t < YS - 400; t += 32) Y[400 + t] = 0.0;  // padding columns of stage B stay finite
  const double wq = lane < NQ2 ? sWq[lane] : 0.0;
  const PhysParams ph = a.phys;
  const int stride = gridDim.x * WPB;
  int fi = blockIdx.x * WPB + warp;
  int4 fd_cur = make_int4(0, 0, 0, 0), fd_nxt = make_int4(0, 0, 0, 0);
  if (fi < face_count) fd_cur = __ldg(&a.face_desc[face_begin + fi]);
  if (fi + stride < face_count) fd_nxt = __ldg(&a.face_desc[face_begin + fi + stride]);
  if (fi < face_count && lane == 0) {
    mbar_expect_tx(bar0, 2 * BLKB);
    bulk_g2s(smem_u32(raw), a.tr + static_cast<long long>(fd_cur.x) * (NTF * NF2), BLKB, bar0);
    bulk_g2s(smem_u32(raw + NTF * NF2), a.tr + static_cast<long long>(fd_cur.y) * (NTF * NF2), BLKB, bar0);
  }
  // stage-A store offsets: lane (row al = fr, columns 2 fk, 2 fk + 1 of group g) -> Y[(2 g + fk / 2)][al][2 (fk % 2)]
  const int yOff = (fk >> 1) * (NP * NQ) + fr * NP + 2 * (fk & 1);
  for (int it = 0; fi < face_count; it++, fi += stride) {
    const int buf = it & 1;
    const unsigned ph_bit = (it >> 1) & 1;
    const int fc = face_begin + fi;
    const double *cur = raw + buf * RAW;
    int4 fd_n2 = make_int4(0, 0, 0, 0);
    if (fi + 2 * stride < face_count) fd_n2 = __ldg(&a.face_desc[fc + 2 * stride]);
    if (fi + stride < face_count && lane == 0) {
      const unsigned nb = buf ? bar0 : bar1;
      double *nx = raw + (buf ^ 1) * RAW;
      mbar_expect_tx(nb, 2 * BLKB);
      bulk_g2s(smem_u32(nx), a.tr + static_cast<long long>(fd_nxt.x) * (NTF * NF2), BLKB, nb);
      bulk_g2s(smem_u32(nx + NTF * NF2), a.tr + static_cast<long long>(fd_nxt.y) * (NTF * NF2), BLKB, nb);
    }
    if (fi + 2 * stride < face_count && lane == 1) {  // two faces ahead: into L2 only
      bulk_prefetch_l2(a.tr + static_cast<long long>(fd_n2.x) * (NTF * NF2), BLKB);
      bulk_prefetch_l2(a.tr + static_cast<long long>(fd_n2.y) * (NTF * NF2), BLKB);
    }
    const double2 nrm01 = __ldg(reinterpret_cast<const double2 *>(a.face_nor) + 2 * fc);
    const double2 nrm23 = __ldg(reinterpret_cast<const double2 *>(a.face_nor) + 2 * fc + 1);
    // side 2: face coordinates (a, b) -> index in Elem2's block (perm code as in face_flux_fast_kernel)
    int p2;
    {
      const int pc = fd_cur.z, b = fr & 3, fa2 = (pc & 2) ? 1 : 0, fb2 = (pc & 4) ? 1 : 0;
      int i0, di;
      if (pc & 1) {
        i0 = (fa2 ? NP - 1 - b : b) + (fb2 ? NP * (NP - 1) : 0);
        di = fb2 ? -NP : NP;
      } else {
        i0 = (fa2 ? NP - 1 : 0) + NP * (fb2 ? NP - 1 - b : b);
        di = fa2 ? -1 : 1;
      }
      p2 = (lane >> 4) * NF2 + i0 + fk * di;
    }
    mbar_wait(buf ? bar1 : bar0, ph_bit);
    // A: interpolation along a
#pragma unroll
    for (int g = 0; g < 10; g++) {
      const double bv = g < 5 ? cur[32 * g + lane] : cur[NTF * NF2 + 32 * (g - 5) + p2];
      double d0, d1;
      dmma884(d0, d1, aP, bv);
      if (fr < NQ) *reinterpret_cast<double2 *>(&Y[2 * g * (NP * NQ) + yOff]) = make_double2(d0, d1);
    }
    __syncwarp();
    // B: interpolation along b; all fragments are read before V (which does not alias Y) is written
#pragma unroll
    for (int g = 0; g < 13; g++) {
      double d0, d1;
      dmma884(d0, d1, aP, Y[32 * g + lane]);
      const int c0 = 8 * g + 2 * fk;  // columns (sf, al) = c / 5, c % 5; row = be
      if (fr < NQ && c0 < 100) {
        const int sf0 = (c0 * 52429) >> 18, al0 = c0 - 5 * sf0;
        V[(al0 + NQ * fr) * 21 + sf0] = d0;
        const int c1 = c0 + 1, sf1 = (c1 * 52429) >> 18, al1 = c1 - 5 * sf1;
        V[(al1 + NQ * fr) * 21 + sf1] = d1;
      }
    }
    __syncwarp();
    // C: numerical flux at quadrature point q = lane
    if (lane < NQ2) {
      double v[2 * NTF];
#pragma unroll
      for (int sf = 0; sf < 2 * NTF; sf++) v[sf] = V[lane * 21 + sf];
      const double *u1 = v, *u2 = v + NTF;
      const double n0 = nrm01.x, n1 = nrm01.y, n2 = nrm23.x, normag = nrm23.y;
      const double ri1 = fast_rcp(u1[0]), ri2 = fast_rcp(u2[0]);
      const double vx1 = u1[1] * ri1, vy1 = u1[2] * ri1, vz1 = u1[3] * ri1;
      const double vx2 = u2[1] * ri2, vy2 = u2[2] * ri2, vz2 = u2[3] * ri2;
      const double vv1 = vx1 * vx1 + vy1 * vy1 + vz1 * vz1, vv2 = vx2 * vx2 + vy2 * vy2 + vz2 * vz2;
      const double p1 = ph.gm1 * (u1[4] - 0.5 * u1[0] * vv1), p2v = ph.gm1 * (u2[4] - 0.5 * u2[0] * vv2);
      const double pr1 = p1 * ri1, pr2 = p2v * ri2;
      const double maxE = fmax(fast_sqrt(vv1) + fast_sqrt(ph.gamma * pr1), fast_sqrt(vv2) + fast_sqrt(ph.gamma * pr2));
      const double vn1 = vx1 * n0 + vy1 * n1 + vz1 * n2, vn2 = vx2 * n0 + vy2 * n1 + vz2 * n2;
      const double diss = 0.5 * maxE * normag;
      const double psum = 0.5 * (p1 + p2v);
      double fx[NEQ];
      fx[0] = 0.5 * (u1[0] * vn1 + u2[0] * vn2) - diss * (u2[0] - u1[0]);
      fx[1] = 0.5 * (u1[1] * vn1 + u2[1] * vn2) + psum * n0 - diss * (u2[1] - u1[1]);
      fx[2] = 0.5 * (u1[2] * vn1 + u2[2] * vn2) + psum * n1 - diss * (u2[2] - u1[2]);
      fx[3] = 0.5 * (u1[3] * vn1 + u2[3] * vn2) + psum * n2 - diss * (u2[3] - u1[3]);
      fx[4] = 0.5 * ((u1[4] + p1) * vn1 + (u2[4] + p2v) * vn2) - diss * (u2[4] - u1[4]);
      if (ph.eq_system != 0) {
        const double iR = 1.0 / ph.R;
        const double T1 = pr1 * iR, T2 = pr2 * iR;
        const double mu1 = ph.C1 * ph.visc_mult * (T1 * fast_sqrt(T1)) * fast_rcp(T1 + ph.S0);
        const double mu2 = ph.C1 * ph.visc_mult * (T2 * fast_sqrt(T2)) * fast_rcp(T2 + ph.S0);
        const double bf = ph.bulk_visc_mult - 2. / 3.;
        const double bd1 = bf * mu1 * u1[NEQ + 4], bd2 = -bf * mu2 * u2[NEQ + 4];
        const double t0 = mu1 * u1[NEQ + 0] + bd1 * n0 - (mu2 * u2[NEQ + 0] + bd2 * n0);
        const double t1 = mu1 * u1[NEQ + 1] + bd1 * n1 - (mu2 * u2[NEQ + 1] + bd2 * n1);
        const double t2 = mu1 * u1[NEQ + 2] + bd1 * n2 - (mu2 * u2[NEQ + 2] + bd2 * n2);
        const double e1 = vx1 * (mu1 * u1[NEQ + 0] + bd1 * n0) + vy1 * (mu1 * u1[NEQ + 1] + bd1 * n1) +
                          vz1 * (mu1 * u1[NEQ + 2] + bd1 * n2) + ph.cp_div_pr * mu1 * u1[NEQ + 3];
        const double e2 = vx2 * (mu2 * u2[NEQ + 0] + bd2 * n0) + vy2 * (mu2 * u2[NEQ + 1] + bd2 * n1) +
                          vz2 * (mu2 * u2[NEQ + 2] + bd2 * n2) + ph.cp_div_pr * mu2 * u2[NEQ + 3];
        fx[1] -= 0.5 * t0;
        fx[2] -= 0.5 * t1;
        fx[3] -= 0.5 * t2;
        fx[4] -= 0.5 * (e1 - e2);
      }
#pragma unroll
      for (int eq = 0; eq < NEQ; eq++) F[eq * NQ2 + lane] = fx[eq] * wq;
    }
    __syncwarp();
    // D: projection along beta (K = 5: k = 0-3, then k = 4 with a one-column A fragment)
#pragma unroll
    for (int g = 0; g < 4; g++) {
      const int col = min(8 * g + fr, NEQ * NQ - 1), eq = (col * 52429) >> 18, al = col - 5 * eq;
      const double *fb = F + eq * NQ2 + al;
      double d0, d1;
      dmma884(d0, d1, aT0, fb[NQ * fk]);
      dmma884_acc(d0, d1, aT1, fb[NQ * 4]);
      const int c0 = 8 * g + 2 * fk;
      if (fr < NP) {
        if (c0 < NEQ * NQ) Bq[c0 * NP + fr] = d0;
        if (c0 + 1 < NEQ * NQ) Bq[(c0 + 1) * NP + fr] = d1;
      }
    }
    __syncwarp();
    // E: projection along alpha, straight to global: R[eq][a + 4 b]
#pragma unroll
    for (int g = 0; g < 3; g++) {
      const int col = min(8 * g + fr, NEQ * NP - 1), eq = col >> 2, b = col & 3;
      const double *bb = Bq + eq * (NQ * NP) + b;
      double d0, d1;
      dmma884(d0, d1, aT0, bb[NP * fk]);
      dmma884_acc(d0, d1, aT1, bb[NP * 4]);
      const int c0 = 8 * g + 2 * fk;
      if (fr < NP && c0 < NEQ * NP) {
        double *dst = a.faceRes + static_cast<long long>(fc) * (NEQ * NF2) + (c0 >> 2) * NF2 + fr + NP * (c0 & 3);
        dst[0] = d0;
        dst[NP] = d1;  // column c0 + 1: same eq (c0 even), b + 1
      }
    }
    __syncwarp();
    fd_cur = fd_nxt;
    fd_nxt = fd_n2;
  }
}

constexpr size_t face_mma_smem_bytes(int wpb) { return static_cast<size_t>(wpb) * (2 * 320 + 104 * 4 + 25 * 21 + 3) * sizeof(double); }

template <int NP>
constexpr size_t face_fast_smem_bytes(int wpb) {
  constexpr int NF2 = NP * NP, NQ = NP + 1, NQ2 = NQ * NQ;
  return static_cast<size_t>(wpb) * (2 * (2 * NTF * NF2) + 2 * NTF * NP * NQ + ((NEQ * NQ2 + 3) / 4) * 4 + NEQ * NQ * NP) *
         sizeof(double);
}

}  // namespace tpsb
