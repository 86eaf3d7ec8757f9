// Fused fast path (p = 3, all-parallelepiped hexahedra, dry air, GL/GL): RHSoperator::Mult in three launches
//
//   elem_fused_kernel   per element, everything that only needs the element and its six face neighbours:
//                       updatePrimitives (src/rhs_operator.cpp:623-651) on the fly, the BR1 gradient
//                       (Gradients::computeGradients, src/gradients.cpp:144-232 + GradFaceIntegrator,
//                       src/faceGradientIntegration.cpp:40-140), the nodal flux and its collocated weak divergence
//                       (GetFlux :493-559, Aflux :379-391, DomainIntegrator src/domain_integrator.cpp:44-99) with
//                       Me^-1 applied (:432-448) -> the VOLUME part of dU/dt, and the face-trace blocks the flux
//                       kernel consumes.  Up and gradUp never touch HBM (they are written only for elements with a
//                       boundary face -- the one-sided BC kernel reads them -- or when the caller asks for them).
//   face_flux_mma_kernel (rhs_fast.cuh, unchanged)  trace blocks -> face residuals
//   lift_kernel         dU/dt += Me^-1 (lift of the six face residuals); optional fused Runge-Kutta stage update
//
// HBM bytes per node: 40 (U) + 120 (trace blocks out) + 40 (volume part out) | 120 + 30 | 40 + 30 + 40 = 460, against
// 787 for prim -> grad_trace -> face -> resid (Up 80, gradUp 240 and the second read of U 40 are gone).
//
// Neighbour primitives for the BR1 jump are rebuilt from the neighbour's conserved lines (L2 hits with the tiled
// element order): 6 x 64 extra primitive conversions per element replace the materialised Up array.
//
// FP64 pipe budget (B200: one FP64 pipe per SM quadrant, a DFMA warp instruction holds it 2 cycles, a DMMA.884 16 --
// tools/ubench/fp64_pipes.cu): contractions whose 8-row A fragment would be mostly padding (end-point traces: 2 of 8
// rows) run as line tasks on DFMA; the derivative (+ both end points: 6 of 8 rows) and the transposed derivative of the
// weak divergence (4 of 8) stay on DMMA for their shared-memory operand reuse.
#pragma once
#include "rhs_fast.cuh"

namespace tpsb {

constexpr int FUSED_WRITE = 1;   // mode bit: write the volume part of dU/dt and the trace blocks
constexpr int FUSED_EXPORT = 2;  // mode bit: write gradUp for every element (tpsb_update_gradients / tpsb_get_fields)

// primitives [rho, u, v, w, T] of DryAir::GetPrimitivesFromConservatives (equation_of_state.cpp:321-335) with one
// reciprocal (agrees with the divide form to ~1 ulp)
__device__ __forceinline__ void dry_prim_fast(double cT, double r, double mx, double my, double mz, double E, double &u, double &v,
                                              double &w, double &T) {
  const double ri = fast_rcp(r);
  u = mx * ri;
  v = my * ri;
  w = mz * ri;
  T = cT * (E - 0.5 * (mx * u + my * v + mz * w)) * ri;
}

// max over the warp of a non-negative double, as its bit pattern (two REDUX instead of five 64-bit shuffle rounds)
__device__ __forceinline__ unsigned long long warp_max_bits(double v) {
  const unsigned hi = static_cast<unsigned>(__double2hiint(v)), lo = static_cast<unsigned>(__double2loint(v));
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  return (static_cast<unsigned long long>(mh) << 32) | ml;
}

// Persistent: gridDim.x CTAs of 64 threads (two warps = one element at a time) stride over the element range, so the
// fragment / line-task addressing and the 1-D tables are set up once per CTA, not once per element.
template <int MINB>
__global__ void __launch_bounds__(64, MINB)
    elem_fused_kernel(KernelArgs a, int elem_begin, int elem_count, const int *elem_list, int mode) {
  constexpr int NP = 4, ND = 64, NF2 = 16, PS = 80;  // PS: padded doubles per field (pad_node)
  constexpr int BLK = NTF * NF2;                      // doubles per (element, face) trace block
  // sm[0 .. 16 PS): fields 0-4 U, 5-9 Up, 10-15 the jump block sJ[6][5][16]; later G[eq][r] = flux . adj(J) row r in
  //                 field 3 eq + r, overwritten in place by its transposed derivative
  // sm[16 PS .. 31 PS): reference derivatives of Up [f][r]; later the 13 viscous fields
  __shared__ __align__(16) double sm[31 * PS];
  __shared__ double sGeo[GEO];
  __shared__ double sD[NP][NP], sLb[2][NP], sWn[NP], sLw[2][NP];
  __shared__ int sNbr[6], sCode[6], sFp[6];
  double *sF = sm, *sJ = sm + 10 * PS, *sDr = sm + 16 * PS;
  const int n = threadIdx.x;
  const int lane = n & 31, half = n >> 5;
  const long long N = a.N;
  if (n < NP * NP) sD[n / NP][n % NP] = c_T.D[n / NP][n % NP];
  if (n < 2 * NP) {
    sLb[n / NP][n % NP] = c_T.lb[n / NP][n % NP];
    sLw[n / NP][n % NP] = c_T.lb[n / NP][n % NP] / c_T.wn[n % NP];  // lift coefficient l_c(end) / w_c
  }
  if (n < NP) sWn[n] = c_T.wn[n];
  if (n < 6) sFp[n] = c_T.face_par[n];
  __syncthreads();
  const int pn = pad_node(n);
  const double cT = a.phys.gm1 / a.phys.R;
  // ---- fragment addressing (see grad_trace_mma_kernel)
  const int fr = lane >> 2, fk = lane & 3;
  const double afrag = fr < NP ? sD[fr][fk] : (fr < NP + 2 ? sLb[fr - NP][fk] : 0.0);
  // transposed derivative with the quadrature weights folded in: Dt[j][m] = D[m][j] w_m / w_j, so that
  // (1 / w_j) sum_m D[m][j] (w_m G_m) needs no separate weighting of G
  const double afragT = fr < NP ? sD[fk][fr] * sWn[fk] / sWn[fr] : 0.0;
  int bOff[DIM];
  int sOff0[DIM], sOff1[DIM];  // P1 store targets of this lane in sm: rows 0-3 -> derivative field, rows 4-5 -> own trace
  const int sFld = fr < NP ? DIM * PS : NF2;  // ... and their stride per field
  const bool sAct = fr < NP + 2;
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    const int str = d == 0 ? 1 : (d == 1 ? NP : NP * NP);
    auto base = [d](int L) { return d == 0 ? 4 * L : (d == 1 ? (L & 3) + 16 * (L >> 2) : L); };
    bOff[d] = pad_node(base(8 * half + fr) + fk * str);
    const int L0 = 8 * half + 2 * fk;
    const int dOff0 = pad_node(base(L0) + (fr & 3) * str), dOff1 = pad_node(base(L0 + 1) + (fr & 3) * str);
    const FacePar fp = (fr & 1) ? decode_face(kFacePar(kFacePlus[d])) : decode_face(kFacePar(kFaceMinus[d]));
    auto abof = [&](int L) {
      const int nb = base(L), i = nb & 3, j = (nb >> 2) & 3, k = nb >> 4;
      const int ia = pick3(fp.as, i, j, k), ib = pick3(fp.at, i, j, k);
      return (fp.ss ? ia : NP - 1 - ia) + NP * (fp.st ? ib : NP - 1 - ib);
    };
    const int lf = (fr & 1) ? kFacePlus[d] : kFaceMinus[d];
    sOff0[d] = fr < NP ? 16 * PS + d * PS + dOff0 : 10 * PS + lf * (NEQ * NF2) + abof(L0);
    sOff1[d] = fr < NP ? 16 * PS + d * PS + dOff1 : 10 * PS + lf * (NEQ * NF2) + abof(L0 + 1);
  }
  // ---- line tasks: thread t < 48 owns line L = t % 16 of axis t / 16 and extrapolates it to both ends (DFMA)
  const int lax = n >> 4, lL = n & 15;
  int lOff = 0, lStr = 0, lFm = 0, lFp = 0;  // padded base / stride of the line, trace slot on the - / + face
  // end-point coefficients l_m(0), l_m(1) are read straight from the constant bank (compile-time addresses: free operands)
#define LB0(m) c_T.lb[0][m]
#define LB1(m) c_T.lb[1][m]
  if (n < 48) {
    const int nb = lax == 0 ? 4 * lL : (lax == 1 ? (lL & 3) + 16 * (lL >> 2) : lL);
    lOff = pad_node(nb);
    lStr = lax == 0 ? 1 : (lax == 1 ? 4 : 20);
    const int i = nb & 3, j = (nb >> 2) & 3, k = nb >> 4;
    const int fm = lax == 0 ? kFaceMinus[0] : (lax == 1 ? kFaceMinus[1] : kFaceMinus[2]);
    const int fpl = lax == 0 ? kFacePlus[0] : (lax == 1 ? kFacePlus[1] : kFacePlus[2]);
    const FacePar pm = decode_face(sFp[fm]), pp = decode_face(sFp[fpl]);
    const int iam = pick3(pm.as, i, j, k), ibm = pick3(pm.at, i, j, k);
    const int iap = pick3(pp.as, i, j, k), ibp = pick3(pp.at, i, j, k);
    lFm = fm * BLK + (pm.ss ? iam : NP - 1 - iam) + NP * (pm.st ? ibm : NP - 1 - ibm);
    lFp = fpl * BLK + (pp.ss ? iap : NP - 1 - iap) + NP * (pp.st ? ibp : NP - 1 - ibp);
  }
  // ---- neighbour tasks (P2): (face, face node); faces are taken in axis pairs (4,2 | 1,3 | 0,5) so that on a
  // structured mesh a warp's 32 tasks see one neighbour axis (the x-normal pair takes the 256-bit loads)
  const int i = n & 3, j = (n >> 2) & 3, k = n >> 4;
  int liftP[DIM], liftM[DIM];  // jump slots of this node on the + / - face of each axis
  {
#pragma unroll
    for (int r = 0; r < DIM; r++) {
      const FacePar fm = decode_face(kFacePar(kFaceMinus[r])), fpl = decode_face(kFacePar(kFacePlus[r]));
      const int iam = pick3(fm.as, i, j, k), ibm = pick3(fm.at, i, j, k);
      const int iap = pick3(fpl.as, i, j, k), ibp = pick3(fpl.at, i, j, k);
      liftM[r] = kFaceMinus[r] * (NEQ * NF2) + (fm.ss ? iam : NP - 1 - iam) + NP * (fm.st ? ibm : NP - 1 - ibm);
      liftP[r] = kFacePlus[r] * (NEQ * NF2) + (fpl.ss ? iap : NP - 1 - iap) + NP * (fpl.st ? ibp : NP - 1 - ibp);
    }
  }
  unsigned long long mcs_bits = 0ull;
  const bool ns = a.phys.eq_system != 0;

  for (int slot = blockIdx.x; slot < elem_count; slot += gridDim.x) {
    const int e = elem_list ? elem_list[elem_begin + slot] : elem_begin + slot;
    const long long o = static_cast<long long>(e) * ND + n;
    if (n < 6) {
      sNbr[n] = a.nbr_elem[e * 6 + n];
      sCode[n] = a.nbr_code[e * 6 + n];
    }
    if (n >= 32 && n < 32 + GEO) sGeo[n - 32] = a.geo[static_cast<long long>(e) * GEO + (n - 32)];
    // ---- P0: own state, primitives
    double s[NEQ];
#pragma unroll
    for (int f = 0; f < NEQ; f++) s[f] = a.U[o + f * N];
    {
      double u, v, w, T;
      dry_prim_fast(cT, s[0], s[1], s[2], s[3], s[4], u, v, w, T);
#pragma unroll
      for (int f = 0; f < NEQ; f++) sF[f * PS + pn] = s[f];
      sF[5 * PS + pn] = s[0];
      sF[6 * PS + pn] = u;
      sF[7 * PS + pn] = v;
      sF[8 * PS + pn] = w;
      sF[9 * PS + pn] = T;
    }
    __syncthreads();
    // ---- P1: D and both end-point extrapolations of the primitives: one DMMA per (axis, field, 8 lines)
#pragma unroll
    for (int d = 0; d < DIM; d++) {
#pragma unroll
      for (int f = 0; f < NEQ; f++) {
        double d0, d1;
        dmma884(d0, d1, afrag, sF[(NEQ + f) * PS + bOff[d]]);
        if (sAct) {
          sm[sOff0[d] + f * sFld] = d0;
          sm[sOff1[d] + f * sFld] = d1;
        }
      }
    }
    double *blk0 = a.tr + static_cast<long long>(e) * 6 * BLK;
    if (n < 48 && (mode & FUSED_WRITE)) {  // traces of the conserved state, straight to the trace blocks
#pragma unroll
      for (int f = 0; f < NEQ; f++) {
        const double *p = &sF[f * PS + lOff];
        const double x0 = p[0], x1 = p[lStr], x2 = p[2 * lStr], x3 = p[3 * lStr];
        blk0[lFm + f * NF2] = LB0(0) * x0 + LB0(1) * x1 + LB0(2) * x2 + LB0(3) * x3;
        blk0[lFp + f * NF2] = LB1(0) * x0 + LB1(1) * x1 + LB1(2) * x2 + LB1(3) * x3;
      }
    }
    __syncthreads();  // own traces (P1) are in sJ
    // ---- P2: neighbour primitive traces at my face nodes -> jumps 1/2 (Up_nbr - Up_own) in place; boundary face:
    // Up2 = Up1 unless useBCinGrad (faceGradientIntegration.cpp:96-115)
#pragma unroll
    for (int rd = 0; rd < 2; rd++) {
      if (rd == 1 && half != 0) break;
      // warp 0: faces (4,2) then (0,5); warp 1: faces (1,3)
      const int pair = rd == 0 ? half : 2;
      const int hi = lane >> 4;
      const int lf = pair == 0 ? (hi ? 2 : 4) : (pair == 1 ? (hi ? 3 : 1) : (hi ? 5 : 0));
      const int ab = lane & 15, fa = ab & 3, fb = ab >> 2;
      double *pj = &sJ[lf * (NEQ * NF2) + ab];
      const int nbr = sNbr[lf];
      if (nbr < 0) {
        if (a.bct.use_bc_in_grad && nbr <= -2) {
          double pT[NEQ], pbc[NEQ];
#pragma unroll
          for (int f = 0; f < NEQ; f++) pT[f] = pj[f * NF2];
          dry_bc_prim_for_gradient(a.bct.bc[-2 - nbr], pT, pbc);
#pragma unroll
          for (int f = 0; f < NEQ; f++) pj[f * NF2] = 0.5 * (pbc[f] - pT[f]);
        } else {
#pragma unroll
          for (int f = 0; f < NEQ; f++) pj[f * NF2] = 0.0;
        }
        continue;
      }
      const int code = sCode[lf];
      const FacePar fq = decode_face(sFp[code & 7]);
      int a2, b2;
      apply_perm<NP>(code >> 3, fa, fb, a2, b2);
      const int q0 = face_node_base<NP>(fq, a2, b2), qs = axis_stride<NP>(fq.an);
      const bool local = nbr < a.NE;
      const double *p = (local ? a.U + static_cast<long long>(nbr) * ND : a.Uhalo + static_cast<long long>(nbr - a.NE) * (NEQ * ND)) + q0;
      const long long fs = local ? N : ND;
      double x[NEQ][NP];
      if (qs == 1 && a.vec_ok) {
#pragma unroll
        for (int f = 0; f < NEQ; f++)
          asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                       : "=d"(x[f][0]), "=d"(x[f][1]), "=d"(x[f][2]), "=d"(x[f][3])
                       : "l"(p + f * fs));
      } else {
#pragma unroll
        for (int f = 0; f < NEQ; f++) {
          const double *pf = p + f * fs;
          x[f][0] = __ldg(pf);
          x[f][1] = __ldg(pf + qs);
          x[f][2] = __ldg(pf + 2 * qs);
          x[f][3] = __ldg(pf + 3 * qs);
        }
      }
      double acc[NEQ] = {0, 0, 0, 0, 0};
#pragma unroll
      for (int m = 0; m < NP; m++) {
        const double lm = fq.side ? LB1(m) : LB0(m);
        double u, v, w, T;
        dry_prim_fast(cT, x[0][m], x[1][m], x[2][m], x[3][m], x[4][m], u, v, w, T);
        acc[0] += lm * x[0][m];
        acc[1] += lm * u;
        acc[2] += lm * v;
        acc[3] += lm * w;
        acc[4] += lm * T;
      }
#pragma unroll
      for (int f = 0; f < NEQ; f++) pj[f * NF2] = 0.5 * (acc[f] - pj[f * NF2]);
    }
    __syncthreads();
    // ---- P3: per node -- lifted physical gradient, viscous face fields, contravariant flux
    const double *A = sGeo;
    const double idet = sGeo[10];
    double G[NEQ][DIM];
    {
      double rg[NEQ][DIM];
      const int idx[3] = {i, j, k};
#pragma unroll
      for (int r = 0; r < DIM; r++) {
        const double lwM = sLw[0][idx[r]], lwP = sLw[1][idx[r]];
#pragma unroll
        for (int f = 0; f < NEQ; f++)
          rg[f][r] = sDr[(f * DIM + r) * PS + pn] + (lwP * sJ[liftP[r] + f * NF2] - lwM * sJ[liftM[r] + f * NF2]);
      }
      double g[NEQ][DIM];
      bool wr = (mode & FUSED_EXPORT) != 0;
      if (!wr) wr = (sNbr[0] < 0) | (sNbr[1] < 0) | (sNbr[2] < 0) | (sNbr[3] < 0) | (sNbr[4] < 0) | (sNbr[5] < 0);
#pragma unroll
      for (int d = 0; d < DIM; d++) {
        const double a0 = A[0 + 3 * d] * idet, a1 = A[1 + 3 * d] * idet, a2 = A[2 + 3 * d] * idet;
#pragma unroll
        for (int f = 0; f < NEQ; f++) g[f][d] = rg[f][0] * a0 + rg[f][1] * a1 + rg[f][2] * a2;
      }
      if (wr) {  // the one-sided boundary-face kernel (and the caller, on request) read gradUp from HBM
#pragma unroll
        for (int d = 0; d < DIM; d++)
#pragma unroll
          for (int f = 0; f < NEQ; f++) a.gradUp[o + (f + d * NEQ) * N] = g[f][d];
      }
      // point state (fluxes.cpp:135-170, 178-335 with DryAirTransport, transport_properties.cpp:223-234)
      const DryPoint q = dry_point(a.phys, s);
      {
        const unsigned long long b = warp_max_bits(dry_char_speed_pt(a.phys, q));  // rhs_operator.cpp:550
        mcs_bits = b > mcs_bits ? b : mcs_bits;
      }
      double visc = 0, bulk = 0, kth = 0;
      if (ns) dry_transport_pt(a.phys, q, visc, bulk, kth);
      double sym[DIM][DIM];
#pragma unroll
      for (int p = 0; p < DIM; p++)
#pragma unroll
        for (int c = 0; c < DIM; c++) sym[p][c] = g[1 + p][c] + g[1 + c][p];
      const double divu = g[1][0] + g[2][1] + g[3][2];
      double *sV = sDr;  // thread n only ever touches column pn of sDr: in place
#pragma unroll
      for (int r = 0; r < DIM; r++) {
        const double nr[3] = {A[r + 0], A[r + 3], A[r + 6]};
        double sv[DIM];
#pragma unroll
        for (int p = 0; p < DIM; p++) sv[p] = sym[p][0] * nr[0] + sym[p][1] * nr[1] + sym[p][2] * nr[2];
        const double hT = g[4][0] * nr[0] + g[4][1] * nr[1] + g[4][2] * nr[2];
#pragma unroll
        for (int p = 0; p < DIM; p++) sV[(r * 4 + p) * PS + pn] = sv[p];
        sV[(r * 4 + 3) * PS + pn] = hT;
        double fc[NEQ];
        dry_conv_dot_n(s, q, nr, fc);
        if (ns) {  // F_v . n^r from the same contracted fields
          const double bd = bulk * divu;
          const double t0 = visc * sv[0] + bd * nr[0], t1 = visc * sv[1] + bd * nr[1], t2 = visc * sv[2] + bd * nr[2];
          fc[1] -= t0;
          fc[2] -= t1;
          fc[3] -= t2;
          fc[4] -= q.vel[0] * t0 + q.vel[1] * t1 + q.vel[2] * t2 + kth * hT;
        }
#pragma unroll
        for (int eq = 0; eq < NEQ; eq++) G[eq][r] = fc[eq];
      }
      sV[12 * PS + pn] = divu;
      // fields 0-9 of sF are this thread's own column: U / Up are dead (s[] lives in registers)
#pragma unroll
      for (int eq = 0; eq < 3; eq++)
#pragma unroll
        for (int r = 0; r < DIM; r++) sF[(eq * DIM + r) * PS + pn] = G[eq][r];
      sF[9 * PS + pn] = G[3][0];
    }
    __syncthreads();  // every lift has read sJ: fields 10-14 may be overwritten; the viscous fields are complete
    sF[10 * PS + pn] = G[3][1];
    sF[11 * PS + pn] = G[3][2];
    sF[12 * PS + pn] = G[4][0];
    sF[13 * PS + pn] = G[4][1];
    sF[14 * PS + pn] = G[4][2];
    // ---- P4a: traces of the normal-contracted viscous fields (axis d carries fields 4 d .. 4 d + 3 and div u)
    if (n < 48 && (mode & FUSED_WRITE)) {
#pragma unroll
      for (int v = 0; v < 5; v++) {
        const double *p = &sDr[(v < 4 ? 4 * lax + v : 12) * PS + lOff];
        const double x0 = p[0], x1 = p[lStr], x2 = p[2 * lStr], x3 = p[3 * lStr];
        const double tm = LB0(0) * x0 + LB0(1) * x1 + LB0(2) * x2 + LB0(3) * x3;
        const double tp = LB1(0) * x0 + LB1(1) * x1 + LB1(2) * x2 + LB1(3) * x3;
        blk0[lFm + (NEQ + v) * NF2] = v < 4 ? -tm : tm;  // outward normal of the - face = -A[d,:]
        blk0[lFp + (NEQ + v) * NF2] = tp;
      }
    }
    __syncthreads();
    if (mode & FUSED_WRITE) {
      // ---- P4b: transposed derivative of G along each axis, in place (a warp reads and writes the same 8 lines)
#pragma unroll
      for (int r = 0; r < DIM; r++) {
#pragma unroll
        for (int eq = 0; eq < NEQ; eq++) {
          double d0, d1;
          double *fld = &sF[(eq * DIM + r) * PS];
          dmma884(d0, d1, afragT, fld[bOff[r]]);
          if (fr < NP) {  // rows 0-3: sOff0/1 hold 16 PS + r PS + the D-fragment node offset
            fld[sOff0[r] - (16 + r) * PS] = d0;
            fld[sOff1[r] - (16 + r) * PS] = d1;
          }
        }
      }
      __syncthreads();
      // ---- P5: volume part of dU/dt = Me^-1 (sum_r D^T_r G_r), Me = diag(w |J|)
#pragma unroll
      for (int eq = 0; eq < NEQ; eq++) {
        const double z = sF[(eq * DIM + 0) * PS + pn] + sF[(eq * DIM + 1) * PS + pn] + sF[(eq * DIM + 2) * PS + pn];
        a.y[o + eq * N] = z * idet;
      }
    }
    __syncthreads();  // the next element overwrites sm / sGeo / sNbr
  }
  if (lane == 0 && mcs_bits > __ldcg(a.maxCharBits)) atomicMax(a.maxCharBits, mcs_bits);
}

// lift_kernel: y = y_vol + Me^-1 sum_faces (+-) l_c(face) R_face   (face_integrator.cpp:348-350, rhs_operator.cpp:432-448);
// y_vol is what elem_fused_kernel left in a.y.  RK: fused Runge-Kutta stage update as in elem_resid_kernel.
// Streaming kernel (110 B per node): all loads of a thread are independent of each other (absent faces read slot 0
// with a zero coefficient), so the 35 of them are in flight together.
template <bool RK>
__global__ void __launch_bounds__(256) lift_kernel(KernelArgs a, int elem_begin, int elem_count) {
  constexpr int NP = 4, ND = 64, NF2 = 16;
  __shared__ double sLb[2][NP], sWn[NP];
  __shared__ int sFace[4][6], sFcode[4][6];
  __shared__ double sDet[4];
  if (threadIdx.x < 2 * NP) sLb[threadIdx.x / NP][threadIdx.x % NP] = c_T.lb[threadIdx.x / NP][threadIdx.x % NP];
  if (threadIdx.x < NP) sWn[threadIdx.x] = c_T.wn[threadIdx.x];
  const int le = threadIdx.x >> 6, n = threadIdx.x & 63;
  const int slot = blockIdx.x * 4 + le;
  const bool active = slot < elem_count;
  const int e = active ? elem_begin + slot : elem_begin;
  if (n < 6) {
    sFace[le][n] = __ldg(&a.el_face[e * 6 + n]);
    sFcode[le][n] = __ldg(&a.el_face_code[e * 6 + n]);
  }
  if (n == 32) sDet[le] = __ldg(&a.geo[static_cast<long long>(e) * GEO + 9]);
  const long long N = a.N;
  const long long o = static_cast<long long>(e) * ND + n;
  double yv[NEQ];
  if (active) {
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) yv[eq] = a.y[o + eq * N];
  }
  __syncthreads();
  if (!active) return;
  const int i = n & 3, j = (n >> 2) & 3, k = n >> 4;
  double R[6][NEQ], coef[6];
#pragma unroll
  for (int lf = 0; lf < 6; lf++) {
    const int fc = sFace[le][lf];
    const int code = sFcode[le][lf];
    const int side = code & 1;
    const FacePar fp = decode_face(kFacePar(lf));
    const int ia = pick3(fp.as, i, j, k), ib = pick3(fp.at, i, j, k), c = pick3(fp.an, i, j, k);
    int fa = fp.ss ? ia : NP - 1 - ia, fb = fp.st ? ib : NP - 1 - ib;
    int a2, b2;  // own local face coordinates -> face (Elem1) coordinates (identity code on side 0)
    apply_perm<NP>(code >> 1, fa, fb, a2, b2);
    coef[lf] = fc < 0 ? 0.0 : (side ? sLb[fp.side][c] : -sLb[fp.side][c]);
    const double *Rp = a.faceRes + static_cast<long long>(fc < 0 ? 0 : fc) * (NEQ * NF2) + a2 + NP * b2;
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) R[lf][eq] = __ldg(Rp + eq * NF2);
  }
  double z[NEQ] = {0, 0, 0, 0, 0};
#pragma unroll
  for (int lf = 0; lf < 6; lf++)
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) z[eq] += coef[lf] * R[lf][eq];
  const double im = 1.0 / (sWn[i] * sWn[j] * sWn[k] * sDet[le]);
  if constexpr (RK) {
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) {
      const double ki = yv[eq] + z[eq] * im, xi = a.rk.X[o + eq * N];
      if (a.rk.Z) a.rk.Z[o + eq * N] = (a.rk.zacc ? a.rk.Z[o + eq * N] : xi) + a.rk.B * ki;
      a.rk.Y[o + eq * N] = xi + a.rk.A * ki;
    }
  } else {
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) a.y[o + eq * N] = yv[eq] + z[eq] * im;
  }
}

}  // namespace tpsb
