// Kernel bodies; included exactly once, by tpsb200.cu.  See rhs_kernels.cuh for the map
// kernel <-> reference routine.
#pragma once
#include "rhs_kernels.cuh"

namespace tpsb {

__constant__ RefTables c_T;

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
// grad_kernel: EPB elements per CTA, one thread per node.
//   volume : gradUp_j = sum_r inv(J)_rd * (D applied along r)             [= Me^-1 Ke Up, collocated]
//   faces  : + l_c(face)/(w_c |J|) * 1/2 (Up_nbr - Up_own)(a,b) * n_d     [= Me^-1 of the face term of
//            GradFaceIntegrator (src/faceGradientIntegration.cpp:119-137); on an affine element the
//            (p+2)^2-point face rule integrates phi_i * jump exactly, so it collapses onto the
//            (p+1)^2 face nodes]
// Only parallelepiped elements take this path; create() rejects other meshes for now.
template <int NP, int EPB>
__global__ void __launch_bounds__(NP *NP *NP *EPB)
    grad_kernel(KernelArgs a, int elem_begin, int elem_count, const int *elem_list) {
  constexpr int ND = NP * NP * NP, NF2 = NP * NP;
  __shared__ double sUp[EPB][NEQ][ND];
  __shared__ double sNb[EPB][NEQ][ND];
  __shared__ double sJ[EPB][NEQ][NF2];
  __shared__ double sVx[EPB][24];
  const int le = threadIdx.x / ND, n = threadIdx.x % ND;
  const int slot = blockIdx.x * EPB + le;
  const bool active = slot < elem_count;
  const int e = active ? (elem_list ? elem_list[elem_begin + slot] : elem_begin + slot) : 0;
  const long long N = a.N;
  if (active) {
#pragma unroll
    for (int f = 0; f < NEQ; f++) sUp[le][f][n] = a.Up[static_cast<long long>(e) * ND + n + f * N];
    for (int t = n; t < 24; t += ND) sVx[le][t] = a.vx[static_cast<long long>(e) * 24 + t];
  }
  __syncthreads();
  const int i = n % NP, j = (n / NP) % NP, k = n / (NP * NP);
  double g[NEQ][DIM];
  double J[9], A[9], det = 1.0;
  if (active) {
    hex_jacobian(sVx[le], c_T.xn[i], c_T.xn[j], c_T.xn[k], J);
    det = det3(J);
    adj3(J, A);
    const double idet = 1.0 / det;
#pragma unroll
    for (int f = 0; f < NEQ; f++) {
      double d0 = 0, d1 = 0, d2 = 0;
#pragma unroll
      for (int m = 0; m < NP; m++) {
        d0 += c_T.D[i][m] * sUp[le][f][m + NP * j + NP * NP * k];
        d1 += c_T.D[j][m] * sUp[le][f][i + NP * m + NP * NP * k];
        d2 += c_T.D[k][m] * sUp[le][f][i + NP * j + NP * NP * m];
      }
      // inv(J)(r,d) = A[r + 3 d] / det
#pragma unroll
      for (int d = 0; d < DIM; d++) g[f][d] = (d0 * A[0 + 3 * d] + d1 * A[1 + 3 * d] + d2 * A[2 + 3 * d]) * idet;
    }
  }
  for (int lf = 0; lf < 6; lf++) {
    const int nbr = active ? a.nbr_elem[e * 6 + lf] : -1;
    const int code = active ? a.nbr_code[e * 6 + lf] : 0;
    __syncthreads();  // previous face done with sNb / sJ
    if (nbr >= 0) {
      if (nbr < a.NE) {
#pragma unroll
        for (int f = 0; f < NEQ; f++) sNb[le][f][n] = a.Up[static_cast<long long>(nbr) * ND + n + f * N];
      } else {
#pragma unroll
        for (int f = 0; f < NEQ; f++)
          sNb[le][f][n] = a.UpHalo[(static_cast<long long>(nbr - a.NE) * NEQ + f) * ND + n];
      }
    }
    __syncthreads();
    if (active) {
      const int side = c_T.face_side[lf], cs = c_T.face_cstride[lf];
      const int lf2 = code & 7, pidx = code >> 3;
      const int side2 = c_T.face_side[lf2], cs2 = c_T.face_cstride[lf2];
      for (int t = n; t < NEQ * NF2; t += ND) {
        const int f = t / NF2, ab = t % NF2;
        double own = 0;
        const int b0 = c_T.face_base[lf][ab];
#pragma unroll
        for (int c = 0; c < NP; c++) own += c_T.lb[side][c] * sUp[le][f][b0 + c * cs];
        double jump = 0.0;  // boundary faces: Up2 = Up1 unless useBCinGrad (faceGradientIntegration.cpp:96-115)
        if (nbr >= 0) {
          const int ab2 = (pidx < 8) ? c_T.perm[pidx][ab] : c_T.iperm[pidx - 8][ab];
          const int b2 = c_T.face_base[lf2][ab2];
          double oth = 0;
#pragma unroll
          for (int c = 0; c < NP; c++) oth += c_T.lb[side2][c] * sNb[le][f][b2 + c * cs2];
          jump = 0.5 * (oth - own);
        }
        sJ[le][f][ab] = jump;
      }
    }
    __syncthreads();
    if (active) {
      // outward area-weighted normal of own local face lf (affine: constant over the face)
      double Xf[12], nor[3];
#pragma unroll
      for (int q = 0; q < 4; q++)
#pragma unroll
        for (int d = 0; d < 3; d++) Xf[q * 3 + d] = sVx[le][c_T.face_vert[lf][q] * 3 + d];
      face_normal(Xf, 0.5, 0.5, nor);
      const int ab = c_T.node_ab[lf][n], c = c_T.node_c[lf][n];
      const double coef = c_T.lb[c_T.face_side[lf]][c] / (c_T.wn[c] * det);
#pragma unroll
      for (int f = 0; f < NEQ; f++) {
        const double v = coef * sJ[le][f][ab];
#pragma unroll
        for (int d = 0; d < DIM; d++) g[f][d] += v * nor[d];
      }
    }
  }
  if (active) {
#pragma unroll
    for (int d = 0; d < DIM; d++)
#pragma unroll
      for (int f = 0; f < NEQ; f++) a.gradUp[static_cast<long long>(e) * ND + n + (f + d * NEQ) * N] = g[f][d];
  }
}

// ------------------------------------------------------------------------------------------------
// face_flux_kernel: FPB faces per CTA, NP^3 threads per face.
// Output faceRes[face][eq][a + NP*b] = sum_q w_q phi^face_ab(s_q) * Fhat_eq(s_q)   (face coordinates
// of Elem1), Fhat = Rusanov(U1,U2,n) - 1/2 (Fv(U1,G1) + Fv(U2,G2)).n with n = CalcOrtho (Elem1 -> Elem2)
// (src/face_integrator.cpp:282-351).  elem_resid_kernel lifts it into both elements.
template <int NP, int FPB>
__global__ void __launch_bounds__(NP *NP *NP *FPB)
    face_flux_kernel(KernelArgs a, int face_begin, int face_count, const int *face_list) {
  constexpr int ND = NP * NP * NP, NF2 = NP * NP, NQ = NP + 1, NQ2 = NQ * NQ;
  // sN is reused for the quadrature-point values once the traces are extracted
  constexpr int NS = ND > NQ2 ? ND : NQ2;
  __shared__ double sN[FPB][2][NFLD][NS];
  __shared__ double sT[FPB][2][NFLD][NF2];
  __shared__ double sA[FPB][2][NFLD][NP * NQ];  // [b][alpha]
  __shared__ double sF[FPB][NEQ][NQ2];
  __shared__ double sB[FPB][NEQ][NP * NQ];  // [b][alpha]
  __shared__ double sXf[FPB][12];
  const int lfc = threadIdx.x / ND, n = threadIdx.x % ND;
  const int slot = blockIdx.x * FPB + lfc;
  const bool active = slot < face_count;
  const int fc = active ? (face_list ? face_list[face_begin + slot] : face_begin + slot) : 0;
  const long long N = a.N;
  int e1 = 0, e2 = 0, lf1 = 0, lf2 = 0, ori = 0;
  if (active) {
    e1 = a.face_el1[fc];
    e2 = a.face_el2[fc];
    lf1 = a.face_inf1[fc] / 64;
    lf2 = a.face_inf2[fc] / 64;
    ori = a.face_inf2[fc] % 64;
    const long long o1 = static_cast<long long>(e1) * ND + n;
#pragma unroll
    for (int f = 0; f < NEQ; f++) sN[lfc][0][f][n] = a.U[o1 + f * N];
#pragma unroll
    for (int f = 0; f < NEQ * DIM; f++) sN[lfc][0][NEQ + f][n] = a.gradUp[o1 + f * N];
    if (e2 < a.NE) {
      const long long o2 = static_cast<long long>(e2) * ND + n;
#pragma unroll
      for (int f = 0; f < NEQ; f++) sN[lfc][1][f][n] = a.U[o2 + f * N];
#pragma unroll
      for (int f = 0; f < NEQ * DIM; f++) sN[lfc][1][NEQ + f][n] = a.gradUp[o2 + f * N];
    } else {
      const long long k2 = e2 - a.NE;
#pragma unroll
      for (int f = 0; f < NEQ; f++) sN[lfc][1][f][n] = a.Uhalo[(k2 * NEQ + f) * ND + n];
#pragma unroll
      for (int f = 0; f < NEQ * DIM; f++) sN[lfc][1][NEQ + f][n] = a.gradUpHalo[(k2 * NEQ * DIM + f) * ND + n];
    }
    for (int t = n; t < 12; t += ND)
      sXf[lfc][t] = a.vx[static_cast<long long>(e1) * 24 + c_T.face_vert[lf1][t / 3] * 3 + t % 3];
  }
  __syncthreads();
  // 1. traces at the (a,b) face nodes, both sides, in face (= Elem1-local) coordinates
  if (active) {
    const int side1 = c_T.face_side[lf1], cs1 = c_T.face_cstride[lf1];
    const int side2 = c_T.face_side[lf2], cs2 = c_T.face_cstride[lf2];
    for (int t = n; t < 2 * NFLD * NF2; t += ND) {
      const int s = t / (NFLD * NF2), f = (t / NF2) % NFLD, ab = t % NF2;
      const int lf = s ? lf2 : lf1, side = s ? side2 : side1, cs = s ? cs2 : cs1;
      const int abl = s ? c_T.perm[ori][ab] : ab;
      const int b0 = c_T.face_base[lf][abl];
      double v = 0;
#pragma unroll
      for (int c = 0; c < NP; c++) v += c_T.lb[side][c] * sN[lfc][s][f][b0 + c * cs];
      sT[lfc][s][f][ab] = v;
    }
  }
  __syncthreads();
  // 2. interpolate along a: A[b][alpha] = sum_a P[alpha][a] T[a + NP b]
  if (active) {
    for (int t = n; t < 2 * NFLD * NP; t += ND) {
      const int s = t / (NFLD * NP), f = (t / NP) % NFLD, b = t % NP;
      double in[NP];
#pragma unroll
      for (int q = 0; q < NP; q++) in[q] = sT[lfc][s][f][q + NP * b];
#pragma unroll
      for (int al = 0; al < NQ; al++) {
        double v = 0;
#pragma unroll
        for (int q = 0; q < NP; q++) v += c_T.P[al][q] * in[q];
        sA[lfc][s][f][b * NQ + al] = v;
      }
    }
  }
  __syncthreads();
  // 3. interpolate along b: Q[alpha + NQ beta] = sum_b P[beta][b] A[b][alpha]   (into sN)
  if (active) {
    for (int t = n; t < 2 * NFLD * NQ; t += ND) {
      const int s = t / (NFLD * NQ), f = (t / NQ) % NFLD, al = t % NQ;
      double in[NP];
#pragma unroll
      for (int q = 0; q < NP; q++) in[q] = sA[lfc][s][f][q * NQ + al];
#pragma unroll
      for (int be = 0; be < NQ; be++) {
        double v = 0;
#pragma unroll
        for (int q = 0; q < NP; q++) v += c_T.P[be][q] * in[q];
        sN[lfc][s][f][al + NQ * be] = v;
      }
    }
  }
  __syncthreads();
  // 4. numerical flux at the quadrature points
  for (int qp = n; active && qp < NQ2; qp += ND) {
    const int al = qp % NQ, be = qp / NQ;
    double u1[NEQ], u2[NEQ], g1[NEQ * DIM], g2[NEQ * DIM], nor[3], fl[NEQ];
#pragma unroll
    for (int f = 0; f < NEQ; f++) {
      u1[f] = sN[lfc][0][f][qp];
      u2[f] = sN[lfc][1][f][qp];
    }
    face_normal(sXf[lfc], c_T.xq[al], c_T.xq[be], nor);
    dry_riemann_lf(a.phys, u1, u2, nor, fl);
    if (a.phys.eq_system != 0) {
#pragma unroll
      for (int f = 0; f < NEQ * DIM; f++) {
        g1[f] = sN[lfc][0][NEQ + f][qp];
        g2[f] = sN[lfc][1][NEQ + f][qp];
      }
      double v1[NEQ * DIM], v2[NEQ * DIM];
      dry_visc_flux(a.phys, u1, g1, v1);
      dry_visc_flux(a.phys, u2, g2, v2);
#pragma unroll
      for (int eq = 0; eq < NEQ; eq++) {
        double s = 0;
#pragma unroll
        for (int d = 0; d < DIM; d++) s += (-0.5 * (v1[eq + d * NEQ] + v2[eq + d * NEQ])) * nor[d];
        fl[eq] += s;
      }
    }
    const double w = c_T.wq[al] * c_T.wq[be];
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) sF[lfc][eq][qp] = fl[eq] * w;
  }
  __syncthreads();
  // 5. project back along beta: B[b][alpha] = sum_beta P[beta][b] F[alpha + NQ beta]
  if (active) {
    for (int t = n; t < NEQ * NQ; t += ND) {
      const int eq = t / NQ, al = t % NQ;
      double in[NQ];
#pragma unroll
      for (int q = 0; q < NQ; q++) in[q] = sF[lfc][eq][al + NQ * q];
#pragma unroll
      for (int b = 0; b < NP; b++) {
        double v = 0;
#pragma unroll
        for (int q = 0; q < NQ; q++) v += c_T.P[q][b] * in[q];
        sB[lfc][eq][b * NQ + al] = v;
      }
    }
  }
  __syncthreads();
  // 6. ... and along alpha, straight to global: R[a + NP b] = sum_alpha P[alpha][a] B[b][alpha]
  if (active) {
    for (int t = n; t < NEQ * NF2; t += ND) {
      const int eq = t / NF2, ab = t % NF2, aa = ab % NP, b = ab / NP;
      double v = 0;
#pragma unroll
      for (int q = 0; q < NQ; q++) v += c_T.P[q][aa] * sB[lfc][eq][b * NQ + q];
      a.faceRes[(static_cast<long long>(fc) * NEQ + eq) * NF2 + ab] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// elem_resid_kernel: EPB elements per CTA, one thread per node.
template <int NP, int EPB>
__global__ void __launch_bounds__(NP *NP *NP *EPB) elem_resid_kernel(KernelArgs a) {
  constexpr int ND = NP * NP * NP, NF2 = NP * NP;
  __shared__ double sG[EPB][NEQ][DIM][ND];
  __shared__ double sVx[EPB][24];
  __shared__ unsigned long long sMaxBits;
  const int le = threadIdx.x / ND, n = threadIdx.x % ND;
  const int e = blockIdx.x * EPB + le;
  const bool active = e < a.NE;
  const long long N = a.N;
  if (active)
    for (int t = n; t < 24; t += ND) sVx[le][t] = a.vx[static_cast<long long>(e) * 24 + t];
  __syncthreads();
  const int i = n % NP, j = (n / NP) % NP, k = n / (NP * NP);
  double det = 1.0, wnode = 1.0, mcs = 0.0;
  if (active) {
    const long long o = static_cast<long long>(e) * ND + n;
    double s[NEQ], fcv[NEQ * DIM];
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) s[eq] = a.U[o + eq * N];
    dry_conv_flux(a.phys, s, fcv);  // GetFlux, rhs_operator.cpp:532
    if (a.phys.eq_system != 0) {
      double g[NEQ * DIM], fv[NEQ * DIM];
#pragma unroll
      for (int f = 0; f < NEQ * DIM; f++) g[f] = a.gradUp[o + f * N];
      dry_visc_flux(a.phys, s, g, fv);
#pragma unroll
      for (int f = 0; f < NEQ * DIM; f++) fcv[f] -= fv[f];  // f -= fvisc, rhs_operator.cpp:540
    }
    mcs = dry_max_char_speed(a.phys, s);  // rhs_operator.cpp:550
    double J[9], A[9];
    hex_jacobian(sVx[le], c_T.xn[i], c_T.xn[j], c_T.xn[k], J);
    det = det3(J);
    adj3(J, A);
    wnode = c_T.wn[i] * c_T.wn[j] * c_T.wn[k];
    // G[eq][r] = w_k sum_d adjJ(r,d) F[eq][d]     (DomainIntegrator, domain_integrator.cpp:71-97)
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++)
#pragma unroll
      for (int r = 0; r < DIM; r++)
        sG[le][eq][r][n] =
            wnode * (A[r + 0] * fcv[eq + 0 * NEQ] + A[r + 3] * fcv[eq + 1 * NEQ] + A[r + 6] * fcv[eq + 2 * NEQ]);
  }
  // max characteristic speed: block reduction, one global atomic per CTA
  if (threadIdx.x == 0) sMaxBits = 0ull;
  __syncthreads();
  if constexpr ((EPB * ND) % 32 == 0) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mcs = fmax(mcs, __shfl_xor_sync(0xffffffffu, mcs, off));
    if ((threadIdx.x & 31) == 0) atomicMax(&sMaxBits, static_cast<unsigned long long>(__double_as_longlong(mcs)));
  } else {
    atomicMax(&sMaxBits, static_cast<unsigned long long>(__double_as_longlong(mcs)));
  }
  __syncthreads();
  if (threadIdx.x == 0) atomicMax(a.maxCharBits, sMaxBits);
  if (!active) return;
  double z[NEQ];
  // z_j = sum_r sum_m D[m][j_r] G[eq][r][.. m ..]   (transpose of the collocation derivative)
#pragma unroll
  for (int eq = 0; eq < NEQ; eq++) {
    double v = 0;
#pragma unroll
    for (int m = 0; m < NP; m++) {
      v += c_T.D[m][i] * sG[le][eq][0][m + NP * j + NP * NP * k];
      v += c_T.D[m][j] * sG[le][eq][1][i + NP * m + NP * NP * k];
      v += c_T.D[m][k] * sG[le][eq][2][i + NP * j + NP * NP * m];
    }
    z[eq] = v;
  }
  // face residuals: elvect1 -= phi1 Fhat w ; elvect2 += phi2 Fhat w   (face_integrator.cpp:348-350)
  for (int lf = 0; lf < 6; lf++) {
    const int fc = a.el_face[e * 6 + lf];
    if (fc < 0) continue;
    const int code = a.el_face_code[e * 6 + lf];
    const int side = code & 1, ori = code >> 1;
    const int abl = c_T.node_ab[lf][n], c = c_T.node_c[lf][n];
    const int ab = side ? c_T.iperm[ori][abl] : abl;
    const double coef = (side ? 1.0 : -1.0) * c_T.lb[c_T.face_side[lf]][c];
    const double *R = a.faceRes + static_cast<long long>(fc) * NEQ * NF2 + ab;
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) z[eq] += coef * R[eq * NF2];
  }
  // y = Me^-1 z, Me = diag(w |J|)   (rhs_operator.cpp:432-448)
  const double im = 1.0 / (wnode * det);
  const long long o = static_cast<long long>(e) * ND + n;
#pragma unroll
  for (int eq = 0; eq < NEQ; eq++) a.y[o + eq * N] = z[eq] * im;
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
#define TPSB_INST(NP, EPB, FPB)                                                                        \
  template __global__ void grad_kernel<NP, EPB>(KernelArgs, int, int, const int *);                    \
  template __global__ void face_flux_kernel<NP, FPB>(KernelArgs, int, int, const int *);               \
  template __global__ void elem_resid_kernel<NP, EPB>(KernelArgs);
TPSB_INST(4, 4, 1)
TPSB_INST(3, 8, 2)
TPSB_INST(2, 16, 4)

}  // namespace tpsb
