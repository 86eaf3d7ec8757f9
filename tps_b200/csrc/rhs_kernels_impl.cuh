// Kernel bodies; included exactly once, by tpsb200.cu.  See rhs_kernels.cuh for the map
// kernel <-> reference routine.
#pragma once
#include "rhs_kernels.cuh"

namespace tpsb {

// one slot per polynomial order (NP = 2, 3, 4), immutable after the first tpsb_create on the device: contexts of
// different order can run concurrently.  Every kernel has NP in scope (template parameter or constexpr).
__constant__ RefTables c_Tab[3];
#define c_T c_Tab[NP - 2]

// ------------------------------------------------------------------------------------------------
__global__ void prim_kernel(KernelArgs a, int halo) {
  const long long n = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long cnt = halo ? a.NH : a.N;
  if (n >= cnt) return;
  double s[NEQ], up[NEQ];
  if (!halo) {
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) s[eq] = a.U[n + eq * cnt];
    dry_prim(a.phys, s, up);
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) a.Up[n + eq * cnt] = up[eq];
  } else {
    // halo buffers are element-major: [halo element][field][node]
    const long long k = n / a.ND, ln = n % a.ND;
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) s[eq] = a.Uhalo[(k * NEQ + eq) * a.ND + ln];
    dry_prim(a.phys, s, up);
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) a.UpHalo[(k * NEQ + eq) * a.ND + ln] = up[eq];
  }
}

// prim_range_kernel: updatePrimitives of the nodes [begin, begin + count) -- the chunked host-buffer pipeline
// (tpsb_rhs_mult_host) converts each element chunk as soon as its host-to-device copy has landed
__global__ void prim_range_kernel(KernelArgs a, long long begin, long long count) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const long long n = begin + t;
  double s[NEQ], up[NEQ];
#pragma unroll
  for (int eq = 0; eq < NEQ; eq++) s[eq] = a.U[n + eq * a.N];
  dry_prim(a.phys, s, up);
#pragma unroll
  for (int eq = 0; eq < NEQ; eq++) a.Up[n + eq * a.N] = up[eq];
}

// pack_kernel: gather the dofs of the elements in send_elems into an element-major send buffer
// (replaces the pack loop of RHSoperator::initNBlockDataTransfer, src/rhs_operator.cpp:786-805)
__global__ void pack_kernel(int nsend, int nd, int nfld, long long N, const int *send_elems, const double *src,
                            double *dst) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(nsend) * nfld * nd;
  if (t >= total) return;
  const int ln = t % nd;
  const int f = (t / nd) % nfld;
  const long long k = t / (static_cast<long long>(nd) * nfld);
  dst[t] = src[static_cast<long long>(send_elems[k]) * nd + ln + f * N];
}

// ------------------------------------------------------------------------------------------------
// Index helpers.  All per-face parameters come from c_T.face_par[] / perm codes staged in shared
// memory: constant-bank reads with thread-dependent indices serialise on the ADU pipe (measured: 89 %
// ADU utilisation in the first version of grad_kernel), so nothing below indexes c_T divergently.
struct FacePar {
  int as, at, an, ss, st, side;
};
// RefTables::face_par of the hexahedron (tables.hpp:186-193; MFEM's CUBE face-vertex table, independent of the order)
// as compile-time constants: with the local face index unrolled, every axis pick / flip below folds away instead of
// costing integer instructions per node (they were ~45 % of elem_resid_kernel's instruction mix, profiles/r1k_*).
// tpsb_create checks them against the table built at run time.
__host__ __device__ constexpr int kFacePar(int lf) {
  return lf == 0 ? 100 : lf == 1 ? 216 : lf == 2 ? 457 : lf == 3 ? 408 : lf == 4 ? 137 : 484;
}
__host__ __device__ __forceinline__ constexpr FacePar decode_face(int par) {
  return FacePar{par & 3, (par >> 2) & 3, (par >> 4) & 3, (par >> 6) & 1, (par >> 7) & 1, (par >> 8) & 1};
}
template <int NP>
__device__ __forceinline__ int axis_stride(int axis) {
  return axis == 0 ? 1 : (axis == 1 ? NP : NP * NP);
}
// element node index of face node (a,b) at normal index c = 0, and the stride along the normal
template <int NP>
__device__ __forceinline__ int face_node_base(const FacePar &f, int a, int b) {
  const int ia = f.ss ? a : NP - 1 - a, ib = f.st ? b : NP - 1 - b;
  return ia * axis_stride<NP>(f.as) + ib * axis_stride<NP>(f.at);
}
template <int NP>
__device__ __forceinline__ void apply_perm(int code, int a, int b, int &a2, int &b2) {
  const int x = (code & 1) ? b : a, y = (code & 1) ? a : b;
  a2 = (code & 2) ? NP - 1 - x : x;
  b2 = (code & 4) ? NP - 1 - y : y;
}
__device__ __forceinline__ int pick3(int axis, int i, int j, int k) { return axis == 0 ? i : (axis == 1 ? j : k); }

// Extrapolate one line of NP nodal values (stride cs doubles) to the face: sum_c lb[c] * r[c*cs].
// Faces whose normal is the element's x axis have the NP values contiguous; for p = 3 they are one
// 32-byte sector, fetched with a single 256-bit load (LDG.E.256 on sm_100a) instead of four strided
// 8-byte loads that each touch the same sector.  vec_ok: all field base pointers are 32-byte aligned.
template <int NP>
__device__ __forceinline__ double trace_line(const double *r, int cs, const double *lb, bool vec_ok) {
  if (NP == 4 && cs == 1 && vec_ok) {
    double x0, x1, x2, x3;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x0), "=d"(x1), "=d"(x2), "=d"(x3) : "l"(r));
    return lb[0] * x0 + lb[1] * x1 + lb[2] * x2 + lb[3] * x3;
  }
  double v = 0;
#pragma unroll
  for (int c = 0; c < NP; c++) v += lb[c] * __ldg(r + c * cs);
  return v;
}


// ------------------------------------------------------------------------------------------------
// grad_kernel: EPB elements per CTA, one thread per node.
//   volume : gradUp_j = sum_r inv(J)_rd * (D applied along r)             [= Me^-1 Ke Up, collocated]
//   faces  : + l_c(face)/(w_c |J|) * 1/2 (Up_nbr - Up_own)(a,b) * n_d(a,b)     [= Me^-1 of the face term of
//            GradFaceIntegrator (src/faceGradientIntegration.cpp:119-137).  On a trilinear hexahedron the
//            CalcOrtho normal is bilinear over the face, so phi_i * jump * n has degree 2p+1 per direction:
//            both the reference's (p+2)^2-point rule and the (p+1)^2 face nodes integrate it exactly, and the
//            term collapses onto the face nodes with the normal taken AT the face node]
// Geometry is the general trilinear map (per-node Jacobian from the 8 vertices); this is the path for
// curved/skewed meshes and boundary conditions, the all-parallelepiped case runs rhs_fast.cuh.
// Neighbour traces are extrapolated straight from global memory (L2-resident neighbour data), own
// traces from shared memory; all six faces are processed between two barriers.
template <int NP, int EPB, int MINB>
__global__ void __launch_bounds__(NP *NP *NP *EPB, MINB)
    grad_kernel(KernelArgs a, int elem_begin, int elem_count, const int *elem_list) {
  constexpr int ND = NP * NP * NP, NF2 = NP * NP;
  __shared__ double sUp[EPB][NEQ][ND];
  __shared__ double sJ[EPB][6][NEQ][NF2];
  __shared__ double sVx[EPB][24];
  __shared__ double sNor[EPB][6][NF2][3];
  __shared__ double sD[NP][NP], sLb[2][NP], sWn[NP], sXn[NP];
  __shared__ int sNbr[EPB][6], sCode[EPB][6], sFp[6], sFv[6][4];
  const int le = threadIdx.x / ND, n = threadIdx.x % ND;
  const int slot = blockIdx.x * EPB + le;
  const bool active = slot < elem_count;
  const int e = active ? (elem_list ? elem_list[elem_begin + slot] : elem_begin + slot) : 0;
  const long long N = a.N;
  if (threadIdx.x < NP * NP) sD[threadIdx.x / NP][threadIdx.x % NP] = c_T.D[threadIdx.x / NP][threadIdx.x % NP];
  if (threadIdx.x < 2 * NP) sLb[threadIdx.x / NP][threadIdx.x % NP] = c_T.lb[threadIdx.x / NP][threadIdx.x % NP];
  if (threadIdx.x < NP) sWn[threadIdx.x] = c_T.wn[threadIdx.x];
  if (threadIdx.x < NP) sXn[threadIdx.x] = c_T.xn[threadIdx.x];
  if (threadIdx.x < 6) sFp[threadIdx.x] = c_T.face_par[threadIdx.x];
  for (int t = threadIdx.x; t < 24; t += blockDim.x) sFv[t / 4][t % 4] = c_T.face_vert[t / 4][t % 4];
  if (active) {
#pragma unroll
    for (int f = 0; f < NEQ; f++) sUp[le][f][n] = a.Up[static_cast<long long>(e) * ND + n + f * N];
    for (int t = n; t < 24; t += ND) sVx[le][t] = a.vx[static_cast<long long>(e) * 24 + t];
    for (int t = n; t < 6; t += ND) {
      sNbr[le][t] = a.nbr_elem[e * 6 + t];
      sCode[le][t] = a.nbr_code[e * 6 + t];
    }
  }
  __syncthreads();
  const int i = n % NP, j = (n / NP) % NP, k = n / (NP * NP);
  double g[NEQ][DIM];
  double det = 1.0;
  if (active) {
    double J[9], A[9];
    hex_jacobian(sVx[le], sXn[i], sXn[j], sXn[k], J);
    det = det3(J);
    adj3(J, A);
    const double idet = 1.0 / det;
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
        d0 += dI[m] * sUp[le][f][m + NP * j + NP * NP * k];
        d1 += dJ[m] * sUp[le][f][i + NP * m + NP * NP * k];
        d2 += dK[m] * sUp[le][f][i + NP * j + NP * NP * m];
      }
      // inv(J)(r,d) = A[r + 3 d] / det
#pragma unroll
      for (int d = 0; d < DIM; d++) g[f][d] = (d0 * A[0 + 3 * d] + d1 * A[1 + 3 * d] + d2 * A[2 + 3 * d]) * idet;
    }
    // face jumps 1/2 (Up_nbr - Up_own) at the face nodes of all six faces: task = (face, face node)
    for (int t = n; t < 6 * NF2; t += ND) {
      const int lf = t / NF2, ab = t % NF2, fa = ab % NP, fb = ab / NP;
      const int nbr = sNbr[le][lf];
      {  // outward CalcOrtho normal at this face node (bilinear over the face of a trilinear hexahedron)
        double Xf[12], nor[3];
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
          for (int d = 0; d < 3; d++) Xf[q * 3 + d] = sVx[le][sFv[lf][q] * 3 + d];
        face_normal(Xf, sXn[fa], sXn[fb], nor);
        sNor[le][lf][ab][0] = nor[0];
        sNor[le][lf][ab][1] = nor[1];
        sNor[le][lf][ab][2] = nor[2];
      }
      const FacePar fp = decode_face(sFp[lf]);
      if (nbr < 0) {  // boundary face: Up2 = Up1 unless useBCinGrad (faceGradientIntegration.cpp:96-115)
        if (a.bct.use_bc_in_grad && nbr <= -2) {
          const int b0b = face_node_base<NP>(fp, fa, fb), csb = axis_stride<NP>(fp.an);
          double own[NEQ], pbc[NEQ];
#pragma unroll
          for (int f = 0; f < NEQ; f++) {
            double v = 0;
#pragma unroll
            for (int c = 0; c < NP; c++) v += sLb[fp.side][c] * sUp[le][f][b0b + c * csb];
            own[f] = v;
          }
          dry_bc_prim_for_gradient(a.bct.bc[-2 - nbr], own, pbc);
#pragma unroll
          for (int f = 0; f < NEQ; f++) sJ[le][lf][f][ab] = 0.5 * (pbc[f] - own[f]);
        } else {
#pragma unroll
          for (int f = 0; f < NEQ; f++) sJ[le][lf][f][ab] = 0.0;
        }
        continue;
      }
      const int code = sCode[le][lf];
      const FacePar fq = decode_face(sFp[code & 7]);
      int a2, b2;
      apply_perm<NP>(code >> 3, fa, fb, a2, b2);
      const int b0 = face_node_base<NP>(fp, fa, fb), cs = axis_stride<NP>(fp.an);
      const int q0 = face_node_base<NP>(fq, a2, b2), qs = axis_stride<NP>(fq.an);
      const bool local = nbr < a.NE;
      const double *src = local ? a.Up + static_cast<long long>(nbr) * ND + q0
                                : a.UpHalo + static_cast<long long>(nbr - a.NE) * NEQ * ND + q0;
      const long long fstride = local ? N : ND;
#pragma unroll
      for (int f = 0; f < NEQ; f++) {
        double own = 0;
#pragma unroll
        for (int c = 0; c < NP; c++) own += sLb[fp.side][c] * sUp[le][f][b0 + c * cs];
        const double oth = trace_line<NP>(src + f * fstride, qs, sLb[fq.side], a.vec_ok != 0);
        sJ[le][lf][f][ab] = 0.5 * (oth - own);
      }
    }
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int lf = 0; lf < 6; lf++) {
      const FacePar fp = decode_face(kFacePar(lf));
      const int ia = pick3(fp.as, i, j, k), ib = pick3(fp.at, i, j, k), c = pick3(fp.an, i, j, k);
      const int ab = (fp.ss ? ia : NP - 1 - ia) + NP * (fp.st ? ib : NP - 1 - ib);
      const double coef = sLb[fp.side][c] / (sWn[c] * det);
      const double n0 = sNor[le][lf][ab][0], n1 = sNor[le][lf][ab][1], n2 = sNor[le][lf][ab][2];
#pragma unroll
      for (int f = 0; f < NEQ; f++) {
        const double v = coef * sJ[le][lf][f][ab];
        g[f][0] += v * n0;
        g[f][1] += v * n1;
        g[f][2] += v * n2;
      }
    }
#pragma unroll
    for (int d = 0; d < DIM; d++)
#pragma unroll
      for (int f = 0; f < NEQ; f++) a.gradUp[static_cast<long long>(e) * ND + n + (f + d * NEQ) * N] = g[f][d];
  }
}

// ------------------------------------------------------------------------------------------------
// face_flux_kernel: FPB faces per CTA of NT threads; every phase is a flat task loop over the CTA so
// that lanes stay packed (40->34 interpolation tasks and 25 flux tasks per face do not fill a warp).
// Output faceRes[face][eq][a + NP*b] = sum_q w_q phi^face_ab(s_q) * Fhat_eq(s_q)   (face coordinates
// of Elem1), Fhat = Rusanov(U1,U2,n) - 1/2 (Fv(U1,G1) + Fv(U2,G2)).n with n = CalcOrtho (Elem1 -> Elem2)
// (src/face_integrator.cpp:282-351).  elem_resid_kernel lifts it into both elements.
//   phase 1  traces of the carried fields of both sides, extrapolated straight from global memory
//   phase 2  one task per (face, side, field): (p+1)^2 -> (p+2)^2 interpolation entirely in registers
//   phase 3  one task per (face, quadrature point): numerical flux
//   phase 4  projection back onto the (p+1)^2 face nodes
// Carried fields (dry air): U (5), grad u (9), grad T (3); grad rho never enters the dry-air viscous flux.
constexpr int NFC = NEQ + (NEQ - 1) * DIM;  // 17
__device__ __forceinline__ int carried_grad_field(int u) {  // u in [NEQ, NFC) -> index into gradUp's NEQ*DIM fields
  const int r = u - NEQ;
  return (r / (NEQ - 1)) * NEQ + (r % (NEQ - 1)) + 1;
}

// BDR = true: the faces are boundary faces (a.bdr_*): only Elem1's side exists, the numerical flux is the
// boundary-condition flux of BCintegrator::AssembleFaceVector (src/BCintegrator.cpp:295-441) and the result
// goes to faceRes slot NFint + k.
template <int NP, int FPB, int NT, bool BDR, bool MOD>
__global__ void __launch_bounds__(NT) face_flux_kernel(KernelArgs a, int face_begin, int face_count, const int *face_list) {
  constexpr int NF2 = NP * NP, NQ = NP + 1, NQ2 = NQ * NQ, ND = NP * NP * NP;
  __shared__ double sT[FPB][2][NFC][NF2 + 1];  // +1: conflict-free per-(side,field) row reads
  __shared__ double sQ[FPB][2][NFC][NQ2];
  __shared__ double sXf[FPB][12];
  __shared__ double sLb[2][NP];
  __shared__ const double *sBaseU[FPB][2], *sBaseG[FPB][2];
  __shared__ long long sStrU[FPB][2], sStrG[FPB][2];
  __shared__ int sOff[FPB][2][NF2], sCs[FPB][2], sSide[FPB][2], sFc[FPB];
  static_assert(sizeof(double) * NEQ * (NQ2 + NP * NQ) <= sizeof(double) * 2 * NFC * (NF2 + 1), "sF/sB overlay");
  const int tid = threadIdx.x;
  const long long N = a.N;
  if (tid < 2 * NP) sLb[tid / NP][tid % NP] = c_T.lb[tid / NP][tid % NP];
  // per-face setup: one thread per (face, side, face node)
  for (int t = tid; t < FPB * 2 * NF2; t += NT) {
    const int fl = t / (2 * NF2), s = (t / NF2) % 2, ab = t % NF2;
    const int slot = blockIdx.x * FPB + fl;
    if (slot >= face_count) {
      if (s == 0 && ab == 0) sFc[fl] = -1;
      continue;
    }
    const int fc = face_list ? face_list[face_begin + slot] : face_begin + slot;
    if (BDR && s == 1) continue;
    const int el = BDR ? a.bdr_el1[fc] : (s ? a.face_el2[fc] : a.face_el1[fc]);
    const int inf = BDR ? 64 * a.bdr_lf[fc] : (s ? a.face_inf2[fc] : a.face_inf1[fc]);
    const FacePar fp = decode_face(c_T.face_par[inf / 64]);
    int fa = ab % NP, fb = ab / NP;
    if (s) {
      int a2, b2;
      apply_perm<NP>(c_T.perm_code[inf % 64], fa, fb, a2, b2);
      fa = a2;
      fb = b2;
    }
    sOff[fl][s][ab] = face_node_base<NP>(fp, fa, fb);
    if (ab == 0) {
      sCs[fl][s] = axis_stride<NP>(fp.an);
      sSide[fl][s] = fp.side;
      if (el < a.NE) {
        sBaseU[fl][s] = a.U + static_cast<long long>(el) * ND;
        sBaseG[fl][s] = a.gradUp + static_cast<long long>(el) * ND;
        sStrU[fl][s] = N;
        sStrG[fl][s] = N;
      } else {
        const long long k2 = el - a.NE;
        sBaseU[fl][s] = a.Uhalo + k2 * NEQ * ND;
        sBaseG[fl][s] = a.gradUpHalo + k2 * NEQ * DIM * ND;
        sStrU[fl][s] = ND;
        sStrG[fl][s] = ND;
      }
      if (s == 0) sFc[fl] = fc;
    }
    if (s == 0 && ab < 12) sXf[fl][ab] = a.vx[static_cast<long long>(el) * 24 + c_T.face_vert[inf / 64][ab / 3] * 3 + ab % 3];
  }
  if (NF2 < 12) {  // p = 1, 2: fewer face nodes than vertex coordinates
    __syncthreads();
    for (int t = tid; t < FPB * 12; t += NT) {
      const int fl = t / 12, q = t % 12, fc = sFc[fl];
      if (fc >= 0)
        sXf[fl][q] = BDR ? a.vx[static_cast<long long>(a.bdr_el1[fc]) * 24 + c_T.face_vert[a.bdr_lf[fc]][q / 3] * 3 + q % 3]
                         : a.vx[static_cast<long long>(a.face_el1[fc]) * 24 + c_T.face_vert[a.face_inf1[fc] / 64][q / 3] * 3 + q % 3];
    }
  }
  __syncthreads();
  // 1. traces at the (a,b) face nodes, both sides, in face (= Elem1-local) coordinates:
  //    one task per (face, side, face node) sweeping all carried fields, so the index arithmetic is
  //    paid once per 17 line extrapolations (it was 57 % of the instructions when paid per field)
  for (int t = tid; t < FPB * 2 * NF2; t += NT) {
    const int ab = t % NF2, s = (t / NF2) % 2, fl = t / (2 * NF2);
    if (sFc[fl] < 0 || (BDR && s == 1)) continue;
    const int off = sOff[fl][s][ab], cs = sCs[fl][s];
    const double *lb = sLb[sSide[fl][s]];
    const bool vec = a.vec_ok != 0;
    const double *pu = sBaseU[fl][s] + off;
    const long long su = sStrU[fl][s], sg = sStrG[fl][s];
    double *dst = &sT[fl][s][0][ab];
#pragma unroll
    for (int u = 0; u < NEQ; u++) dst[u * (NF2 + 1)] = trace_line<NP>(pu + u * su, cs, lb, vec);
    const double *pg = sBaseG[fl][s] + off;
#pragma unroll
    for (int d = 0; d < DIM; d++)
#pragma unroll
      for (int i = 0; i < NEQ - 1; i++)
        dst[(NEQ + d * (NEQ - 1) + i) * (NF2 + 1)] = trace_line<NP>(pg + (d * NEQ + i + 1) * sg, cs, lb, vec);
  }
  __syncthreads();
  // 2. tensor interpolation to the quadrature points, one (face, side, field) per task, in registers:
  //    A[b][alpha] = sum_a P[alpha][a] T[a + NP b] ;  Q[alpha + NQ beta] = sum_b P[beta][b] A[b][alpha]
  for (int t = tid; t < FPB * 2 * NFC; t += NT) {
    const int u = t % NFC, s = (t / NFC) % 2, fl = t / (2 * NFC);
    if (sFc[fl] < 0 || (BDR && s == 1)) continue;
    double A[NP][NQ];
    {
      double T[NF2];
#pragma unroll
      for (int q = 0; q < NF2; q++) T[q] = sT[fl][s][u][q];
#pragma unroll
      for (int b = 0; b < NP; b++)
#pragma unroll
        for (int al = 0; al < NQ; al++) {
          double v = 0;
#pragma unroll
          for (int q = 0; q < NP; q++) v += c_T.P[al][q] * T[q + NP * b];
          A[b][al] = v;
        }
    }
#pragma unroll
    for (int be = 0; be < NQ; be++)
#pragma unroll
      for (int al = 0; al < NQ; al++) {
        double v = 0;
#pragma unroll
        for (int b = 0; b < NP; b++) v += c_T.P[be][b] * A[b][al];
        sQ[fl][s][u][al + NQ * be] = v;
      }
  }
  __syncthreads();
  // sT is dead: reuse it for the weighted fluxes and the half-projected fluxes
  constexpr int FSTR = 2 * NFC * (NF2 + 1);  // doubles per face in sT
  // 3. numerical flux at the quadrature points
  for (int t = tid; t < FPB * NQ2; t += NT) {
    const int qp = t % NQ2, fl = t / NQ2;
    if (sFc[fl] < 0) continue;
    const int al = qp % NQ, be = qp / NQ;
    double u1[NEQ], u2[NEQ], nor[3], fl1[NEQ], fl2[NEQ];
#pragma unroll
    for (int f = 0; f < NEQ; f++) {
      u1[f] = sQ[fl][0][f][qp];
      u2[f] = BDR ? 0.0 : sQ[fl][1][f][qp];
    }
    face_normal(sXf[fl], c_T.xq[al], c_T.xq[be], nor);
    // SGS model / viscous sponge: same physical point for both sides, each side's own element size (parity trap 5)
    constexpr bool mod = MOD;  // compiled out of the plain instantiation (it cost ~15 % there as a run-time branch)
    DryAux ax1, ax2;
    if constexpr (MOD) {
      const int fcm = sFc[fl];
      face_point(sXf[fl], c_T.xq[al], c_T.xq[be], ax1.x);
      ax2.x[0] = ax1.x[0], ax2.x[1] = ax1.x[1], ax2.x[2] = ax1.x[2];
      ax1.delta = a.elem_delta[BDR ? a.bdr_el1[fcm] : a.face_el1[fcm]];
      ax2.delta = BDR ? ax1.delta : a.elem_delta[a.face_el2[fcm]];
    }
    if constexpr (BDR) {
      double g[NEQ * DIM], fxb[NEQ];
#pragma unroll
      for (int d = 0; d < DIM; d++) {
        g[0 + d * NEQ] = 0.0;  // grad(rho) never enters the dry-air fluxes
#pragma unroll
        for (int i = 0; i < NEQ - 1; i++) g[1 + i + d * NEQ] = sQ[fl][0][NEQ + d * (NEQ - 1) + i][qp];
      }
      dry_bc_flux(a.phys, a.bct.bc[a.bdr_bc[sFc[fl]]], a.bct.use_bc_in_grad, u1, g, nor, fxb, mod ? &ax1 : nullptr);
      const double wb = c_T.wq[al] * c_T.wq[be];
      double *dstb = &sT[0][0][0][0] + fl * (2 * NFC * (NF2 + 1));
#pragma unroll
      for (int eq = 0; eq < NEQ; eq++) dstb[eq * NQ2 + qp] = fxb[eq] * wb;
      continue;
    }
    const DryPoint q1 = dry_point(a.phys, u1), q2 = dry_point(a.phys, u2);
    // Rusanov (riemann_solver.cpp:89-114)
    const double maxE = fmax(dry_char_speed_pt(a.phys, q1), dry_char_speed_pt(a.phys, q2));
    dry_conv_dot_n(u1, q1, nor, fl1);
    dry_conv_dot_n(u2, q2, nor, fl2);
    const double normag = sqrt(nor[0] * nor[0] + nor[1] * nor[1] + nor[2] * nor[2]);
    double fx[NEQ];
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) fx[eq] = 0.5 * (fl1[eq] + fl2[eq]) - 0.5 * maxE * (u2[eq] - u1[eq]) * normag;
    if (a.phys.eq_system != 0) {  // - 1/2 (Fv1 + Fv2).n  (face_integrator.cpp:331-341)
      double gu[9], gT[3], v1[NEQ], v2[NEQ];
#pragma unroll
      for (int d = 0; d < DIM; d++) {
#pragma unroll
        for (int i = 0; i < DIM; i++) gu[i + 3 * d] = sQ[fl][0][NEQ + d * (NEQ - 1) + i][qp];
        gT[d] = sQ[fl][0][NEQ + d * (NEQ - 1) + 3][qp];
      }
      dry_visc_dot_n(a.phys, q1, gu, gT, nor, v1, mod ? &ax1 : nullptr, u1[0]);
#pragma unroll
      for (int d = 0; d < DIM; d++) {
#pragma unroll
        for (int i = 0; i < DIM; i++) gu[i + 3 * d] = sQ[fl][1][NEQ + d * (NEQ - 1) + i][qp];
        gT[d] = sQ[fl][1][NEQ + d * (NEQ - 1) + 3][qp];
      }
      dry_visc_dot_n(a.phys, q2, gu, gT, nor, v2, mod ? &ax2 : nullptr, u2[0]);
#pragma unroll
      for (int eq = 1; eq < NEQ; eq++) fx[eq] -= 0.5 * (v1[eq] + v2[eq]);
    }
    const double w = c_T.wq[al] * c_T.wq[be];
    double *dst = &sT[0][0][0][0] + fl * FSTR;
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) dst[eq * NQ2 + qp] = fx[eq] * w;
  }
  __syncthreads();
  // 4a. project back along beta: B[b][alpha] = sum_beta P[beta][b] F[alpha + NQ beta]
  for (int t = tid; t < FPB * NEQ * NQ; t += NT) {
    const int al = t % NQ, eq = (t / NQ) % NEQ, fl = t / (NEQ * NQ);
    if (sFc[fl] < 0) continue;
    double *base = &sT[0][0][0][0] + fl * FSTR;
    double in[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) in[q] = base[eq * NQ2 + al + NQ * q];
#pragma unroll
    for (int b = 0; b < NP; b++) {
      double v = 0;
#pragma unroll
      for (int q = 0; q < NQ; q++) v += c_T.P[q][b] * in[q];
      base[NEQ * NQ2 + eq * NP * NQ + b * NQ + al] = v;
    }
  }
  __syncthreads();
  // 4b. ... and along alpha, straight to global: R[a + NP b] = sum_alpha P[alpha][a] B[b][alpha]
  for (int t = tid; t < FPB * NEQ * NP; t += NT) {
    const int b = t % NP, eq = (t / NP) % NEQ, fl = t / (NEQ * NP);
    const int fc = sFc[fl];
    if (fc < 0) continue;
    const double *base = &sT[0][0][0][0] + fl * FSTR + NEQ * NQ2 + eq * NP * NQ + b * NQ;
    double in[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) in[q] = base[q];
#pragma unroll
    for (int aa = 0; aa < NP; aa++) {
      double v = 0;
#pragma unroll
      for (int q = 0; q < NQ; q++) v += c_T.P[q][aa] * in[q];
      a.faceRes[(static_cast<long long>(BDR ? a.NFint + fc : fc) * NEQ + eq) * NF2 + aa + NP * b] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// elem_resid_kernel: EPB elements per CTA, one thread per node.
// AFF: every element is a parallelepiped (fast path): adj(J) and det come from the 12-double table a.geo
// instead of being rebuilt per node from the 8 vertices (~250 flops per node saved).
template <int NP, int EPB, int MINB, bool AFF = false, bool MOD = false, bool RK = false>
__global__ void __launch_bounds__(NP *NP *NP *EPB, MINB) elem_resid_kernel(KernelArgs a, int elem_begin, int elem_count) {
  constexpr int ND = NP * NP * NP, NF2 = NP * NP;
  __shared__ double sG[EPB][NEQ][DIM][ND];
  __shared__ double sVx[EPB][24];
  __shared__ double sD[NP][NP], sLb[2][NP], sWn[NP], sXn[NP];
  __shared__ int sFace[EPB][6], sFcode[EPB][6], sFp[6];
  __shared__ __align__(16) double sR[EPB][6][NEQ * NF2];  // face residual blocks of the six faces (cp.async)
  __shared__ unsigned long long sMaxBits;
  const int le = threadIdx.x / ND, n = threadIdx.x % ND;
  const int slot = blockIdx.x * EPB + le;
  const bool active = slot < elem_count;
  const int e = active ? elem_begin + slot : 0;
  const long long N = a.N;
  if (threadIdx.x < NP * NP) sD[threadIdx.x / NP][threadIdx.x % NP] = c_T.D[threadIdx.x / NP][threadIdx.x % NP];
  if (threadIdx.x < 2 * NP) sLb[threadIdx.x / NP][threadIdx.x % NP] = c_T.lb[threadIdx.x / NP][threadIdx.x % NP];
  if (threadIdx.x < NP) sWn[threadIdx.x] = c_T.wn[threadIdx.x];
  if (threadIdx.x < NP) sXn[threadIdx.x] = c_T.xn[threadIdx.x];
  if (threadIdx.x < 6) sFp[threadIdx.x] = c_T.face_par[threadIdx.x];
  if (threadIdx.x == 0) sMaxBits = 0ull;
  if (active) {
    if constexpr (AFF) {
      for (int t = n; t < 12; t += ND) sVx[le][t] = a.geo[static_cast<long long>(e) * 12 + t];
    } else {
      for (int t = n; t < 24; t += ND) sVx[le][t] = a.vx[static_cast<long long>(e) * 24 + t];
    }
    for (int t = n; t < 6; t += ND) {
      sFace[le][t] = a.el_face[e * 6 + t];
      sFcode[le][t] = a.el_face_code[e * 6 + t];
    }
  }
  __syncthreads();
  // asynchronous, coalesced staging of the <= 6 contiguous face-residual blocks (NEQ*NF2 doubles each):
  // issued now, consumed after the flux evaluation below
  if (active) {
    // 16-byte copies when a block is an even number of doubles (p = 1, 3), else 8-byte copies (p = 2)
    constexpr int W = (NEQ * NF2) % 2 == 0 ? 2 : 1;
    constexpr int CH = NEQ * NF2 / W;
    for (int t = n; t < 6 * CH; t += ND) {
      const int lf = t / CH, ch = t % CH, fc = sFace[le][lf];
      if (fc >= 0) {
        const double *g = a.faceRes + static_cast<long long>(fc) * NEQ * NF2 + W * ch;
        const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(&sR[le][lf][W * ch]));
        if (W == 2)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g));
        else
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(g));
      }
    }
    asm volatile("cp.async.commit_group;");
  }
  const int i = n % NP, j = (n / NP) % NP, k = n / (NP * NP);
  double det = 1.0, wnode = 1.0, mcs = 0.0;
  if (active) {
    const long long o = static_cast<long long>(e) * ND + n;
    double s[NEQ];
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) s[eq] = a.U[o + eq * N];
    const DryPoint q = dry_point(a.phys, s);
    mcs = dry_char_speed_pt(a.phys, q);  // rhs_operator.cpp:550
    double gu[9], gT[3];
    if (a.phys.eq_system != 0) {
#pragma unroll
      for (int d = 0; d < DIM; d++) {
#pragma unroll
        for (int c = 0; c < DIM; c++) gu[c + 3 * d] = a.gradUp[o + (1 + c + d * NEQ) * N];
        gT[d] = a.gradUp[o + (4 + d * NEQ) * N];
      }
    }
    double A[9];
    if constexpr (AFF) {
#pragma unroll
      for (int t = 0; t < 9; t++) A[t] = sVx[le][t];
      det = sVx[le][9];
    } else {
      double J[9];
      hex_jacobian(sVx[le], sXn[i], sXn[j], sXn[k], J);
      det = det3(J);
      adj3(J, A);
    }
    wnode = sWn[i] * sWn[j] * sWn[k];
    // SGS model / viscous sponge (general path only; the fast path never carries them)
    DryAux ax;
    constexpr bool mod = MOD && !AFF;
    if constexpr (mod) {
      ax.delta = a.elem_delta[e];
      hex_point(sVx[le], sXn[i], sXn[j], sXn[k], ax.x);
    }
    double visc = 0, bulk = 0, kth = 0;
    if constexpr (mod) {
      if (a.phys.eq_system != 0) dry_visc_coeffs(a.phys, q, gu, &ax, s[0], visc, bulk, kth);
    }
    // G[eq][r] = w_k sum_d adjJ(r,d) (F_c - F_v)[eq][d]: the flux of GetFlux (rhs_operator.cpp:532-540)
    // contracted with row r of adj(J) (DomainIntegrator, domain_integrator.cpp:71-97)
#pragma unroll
    for (int r = 0; r < DIM; r++) {
      const double ar[3] = {A[r + 0], A[r + 3], A[r + 6]};
      double fc[NEQ];
      dry_conv_dot_n(s, q, ar, fc);
      if (a.phys.eq_system != 0) {
        double fv[NEQ];
        if constexpr (mod)
          dry_visc_dot_n_c(q, visc, bulk, kth, gu, gT, ar, fv);
        else
          dry_visc_dot_n(a.phys, q, gu, gT, ar, fv);
#pragma unroll
        for (int eq = 1; eq < NEQ; eq++) fc[eq] -= fv[eq];
      }
#pragma unroll
      for (int eq = 0; eq < NEQ; eq++) sG[le][eq][r][n] = wnode * fc[eq];
    }
  }
  // max characteristic speed: block reduction, one global atomic per CTA
  if constexpr ((EPB * ND) % 32 == 0) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mcs = fmax(mcs, __shfl_xor_sync(0xffffffffu, mcs, off));
    if ((threadIdx.x & 31) == 0) atomicMax(&sMaxBits, static_cast<unsigned long long>(__double_as_longlong(mcs)));
  } else {
    atomicMax(&sMaxBits, static_cast<unsigned long long>(__double_as_longlong(mcs)));
  }
  asm volatile("cp.async.wait_all;");
  __syncthreads();
  if (threadIdx.x == 0) atomicMax(a.maxCharBits, sMaxBits);
  if (!active) return;
  double z[NEQ];
  {
    // z_j = sum_r sum_m D[m][j_r] G[eq][r][.. m ..]   (transpose of the collocation derivative)
    double tI[NP], tJ[NP], tK[NP];
#pragma unroll
    for (int m = 0; m < NP; m++) {
      tI[m] = sD[m][i];
      tJ[m] = sD[m][j];
      tK[m] = sD[m][k];
    }
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) {
      double v = 0;
#pragma unroll
      for (int m = 0; m < NP; m++) {
        v += tI[m] * sG[le][eq][0][m + NP * j + NP * NP * k];
        v += tJ[m] * sG[le][eq][1][i + NP * m + NP * NP * k];
        v += tK[m] * sG[le][eq][2][i + NP * j + NP * NP * m];
      }
      z[eq] = v;
    }
  }
  // face residuals: elvect1 -= phi1 Fhat w ; elvect2 += phi2 Fhat w   (face_integrator.cpp:348-350)
#pragma unroll
  for (int lf = 0; lf < 6; lf++) {
    const int fc = sFace[le][lf];
    if (fc < 0) continue;
    const int code = sFcode[le][lf];
    const int side = code & 1;
    const FacePar fp = decode_face(kFacePar(lf));
    const int ia = pick3(fp.as, i, j, k), ib = pick3(fp.at, i, j, k), c = pick3(fp.an, i, j, k);
    int fa = fp.ss ? ia : NP - 1 - ia, fb = fp.st ? ib : NP - 1 - ib;
    if (side) {  // own local face coordinates -> face (Elem1) coordinates
      int a2, b2;
      apply_perm<NP>(code >> 1, fa, fb, a2, b2);
      fa = a2;
      fb = b2;
    }
    const double coef = (side ? 1.0 : -1.0) * sLb[fp.side][c];
    const double *R = &sR[le][lf][fa + NP * fb];
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) z[eq] += coef * R[eq * NF2];
  }
  // y = Me^-1 z, Me = diag(w |J|)   (rhs_operator.cpp:432-448)
  const double im = 1.0 / (wnode * det);
  const long long o = static_cast<long long>(e) * ND + n;
  if constexpr (RK) {  // fused Runge-Kutta stage update (same FMAs the separate axpy kernel would issue)
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) {
      const double ki = z[eq] * im, xi = a.rk.X[o + eq * N];
      if (a.rk.Z) a.rk.Z[o + eq * N] = (a.rk.zacc ? a.rk.Z[o + eq * N] : xi) + a.rk.B * ki;
      a.rk.Y[o + eq * N] = xi + a.rk.A * ki;
    }
  } else {
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) a.y[o + eq * N] = z[eq] * im;
  }
}

__global__ void axpy2_kernel(long long n, const double *x, const double *k, double a, double *y, double b, double *z,
                             int z_accumulate) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double xi = x[i], ki = k[i];
  if (y) y[i] = xi + a * ki;
  if (z) z[i] = (z_accumulate ? z[i] : xi) + b * ki;
}

// out = a*x + b*y (RK3SSP stage combination)
__global__ void rk3_combine_kernel(long long n, const double *x, const double *y, double a, double b, double *out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = a * x[i] + b * y[i];
}

// explicit instantiations (p = 1, 2, 3)
// (kernels are instantiated implicitly by the launchers in tpsb200.cu)

}  // namespace tpsb
