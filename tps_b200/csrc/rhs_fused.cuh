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

// ---- shared-memory layout of elem_fused_kernel ---------------------------------------------------------------------
// Fields live in PAIRS: one 16-byte unit per node holds two fields, so the per-node and per-line accesses are 128-bit
// (ncu, profiles/r2b: a 64-bit shared access of a full warp costs 4 wavefronts like a 128-bit one -- the data pipe
// works on quarter warps -- and the first fused kernel sat at 85 % of that pipe).  Node n of a pair field sits in unit
// swz(n): an XOR swizzle of the low three bits with (j1, k0, k1) that makes EVERY access pattern of the kernel hit 8
// distinct 16-byte bank groups per quarter warp without padding: node order, the DMMA B fragments and D-fragment
// stores along x / y / z, and the line tasks (8 lines at one position) along x / y / z (checked exhaustively by
// tests/test_cpu_capi.py::test_fused_swizzle_is_conflict_free).
__host__ __device__ constexpr int fused_swz(int n) {
  return (n & 0x38) | ((n & 7) ^ ((n >> 2) & 2) ^ (((n >> 4) & 1) * 5) ^ ((n >> 4) & 2));
}
constexpr int FPU = 128;                 // doubles per pair field (64 units x 2)
constexpr int F_S0 = 0 * FPU;            // (rho, u)
constexpr int F_S1 = 1 * FPU;            // (v, w)
constexpr int F_S2 = 2 * FPU;            // (T, -)
constexpr int F_S3 = 3 * FPU;            // (rho u, rho v)
constexpr int F_S4 = 4 * FPU;            // (rho w, rho E)
constexpr int F_SJ = 5 * FPU;            // jump block [6 faces][16 face nodes][6]: (rho, u, v, w, T, -) own trace -> jump
constexpr int F_DR = F_SJ + 6 * 16 * 6;  // 8 pair fields: (d rho, d u)_r, (d v, d w)_r for r = 0,1,2; (dT_0, dT_1); (dT_2, -)
constexpr int F_TOTAL = F_DR + 8 * FPU;  // 2240 doubles = 17.9 KB
// later in the element: G = flux . adj(J) row r in pair fields 0-7 of the front region ((G0,G1)_r, (G2,G3)_r at 2 r,
// 2 r + 1; (G4_0, G4_1) at 6; (G4_2, -) at 7; pairs 5-7 overlay the dead jump block), and the viscous fields in
// place of the derivatives ((s0,s1)_r, (s2, n.gradT)_r at F_DR + 2 r, 2 r + 1; (div u, -) at F_DR + 6)

__device__ __forceinline__ double2 lds128(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void sts128(double *p, double x, double y) { *reinterpret_cast<double2 *>(p) = make_double2(x, y); }

// Persistent: gridDim.x CTAs of 64 threads (two warps = one element at a time) stride over the element range, so the
// fragment / line-task addressing and the 1-D tables are set up once per CTA, not once per element.
template <int REGS>
__global__ void __launch_bounds__(64) __maxnreg__(REGS)
    elem_fused_kernel(KernelArgs a, int elem_begin, int elem_count, const int *elem_list, int mode) {
  constexpr int NP = 4, ND = 64, NF2 = 16;
  constexpr int BLK = NTF * NF2;  // doubles per (element, face) trace block
  __shared__ __align__(16) double sm[F_TOTAL];
  __shared__ double sGeo[GEO];
  __shared__ double sD[NP][NP], sLb[2][NP], sWn[NP], sLw[2][NP];
  __shared__ int sNbr[6], sCode[6], sFp[6];
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
  const int un = 2 * fused_swz(n);  // this thread's node: offset of its unit inside a pair field
  const double cT = a.phys.gm1 / a.phys.R;
  // ---- DMMA fragment addressing.  Line L (0..15) of axis d: nodes base_d(L) + m stride_d.
  //   B fragment : m = lane%4, L = 8 half + lane/4
  //   D fragment : row = lane/4 (0-3 derivative at position row, 4 / 5 trace at the - / + end), lines L0 = 8 half + 2 (lane%4), L0 + 1
  const int fr = lane >> 2, fk = lane & 3;
  const double afrag = fr < NP ? sD[fr][fk] : (fr < NP + 2 ? sLb[fr - NP][fk] : 0.0);
  // transposed derivative with the quadrature weights folded in: Dt[j][m] = D[m][j] w_m / w_j, so that
  // (1 / w_j) sum_m D[m][j] (w_m G_m) needs no separate weighting of G
  const double afragT = fr < NP ? sD[fk][fr] * sWn[fk] / sWn[fr] : 0.0;
  int bU[DIM];             // B-fragment unit offset (doubles) inside a pair field
  int o0[DIM], o1[DIM];    // P1 store target of (rho, u): derivative pair 2 d at the D-fragment node (rows 0-3), or the jump
                           // block entry of the line's face node on the - / + face (rows 4 / 5)
  const bool isD = fr < NP, sAct = fr < NP + 2;
  const int pst = isD ? FPU : 2;  // ... offset of the (v, w) store from it
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    const int str = d == 0 ? 1 : (d == 1 ? NP : NP * NP);
    auto base = [d](int L) { return d == 0 ? 4 * L : (d == 1 ? (L & 3) + 16 * (L >> 2) : L); };
    bU[d] = 2 * fused_swz(base(8 * half + fr) + fk * str);
    const int L0 = 8 * half + 2 * fk;
    const FacePar fp = (fr & 1) ? decode_face(kFacePar(kFacePlus[d])) : decode_face(kFacePar(kFaceMinus[d]));
    auto abof = [&](int L) {
      const int nb = base(L), i = nb & 3, j = (nb >> 2) & 3, k = nb >> 4;
      const int ia = pick3(fp.as, i, j, k), ib = pick3(fp.at, i, j, k);
      return (fp.ss ? ia : NP - 1 - ia) + NP * (fp.st ? ib : NP - 1 - ib);
    };
    const int lf = (fr & 1) ? kFacePlus[d] : kFaceMinus[d];
    o0[d] = isD ? F_DR + 2 * d * FPU + 2 * fused_swz(base(L0) + (fr & 3) * str) : F_SJ + (lf * NF2 + abof(L0)) * 6;
    o1[d] = isD ? F_DR + 2 * d * FPU + 2 * fused_swz(base(L0 + 1) + (fr & 3) * str) : F_SJ + (lf * NF2 + abof(L0 + 1)) * 6;
  }
  // ---- line tasks: thread t < 48 owns line L = t % 16 of axis t / 16 and extrapolates it to both ends (DFMA)
  // end-point coefficients l_m(0), l_m(1) are read straight from the constant bank (compile-time addresses: free operands)
#define LB0(m) c_T.lb[0][m]
#define LB1(m) c_T.lb[1][m]
  const int lax = n >> 4, lL = n & 15;
  int lU[NP] = {0, 0, 0, 0};  // unit offsets (doubles) of the line's four nodes
  int lFm = 0, lFp = 0;       // trace slot on the - / + face
  if (n < 48) {
    const int nb = lax == 0 ? 4 * lL : (lax == 1 ? (lL & 3) + 16 * (lL >> 2) : lL);
    const int ls = lax == 0 ? 1 : (lax == 1 ? 4 : 16);
#pragma unroll
    for (int m = 0; m < NP; m++) lU[m] = 2 * fused_swz(nb + m * ls);
    const int i = nb & 3, j = (nb >> 2) & 3, k = nb >> 4;
    const int fm = lax == 0 ? kFaceMinus[0] : (lax == 1 ? kFaceMinus[1] : kFaceMinus[2]);
    const int fpl = lax == 0 ? kFacePlus[0] : (lax == 1 ? kFacePlus[1] : kFacePlus[2]);
    const FacePar pm = decode_face(sFp[fm]), pp = decode_face(sFp[fpl]);
    const int iam = pick3(pm.as, i, j, k), ibm = pick3(pm.at, i, j, k);
    const int iap = pick3(pp.as, i, j, k), ibp = pick3(pp.at, i, j, k);
    lFm = fm * BLK + (pm.ss ? iam : NP - 1 - iam) + NP * (pm.st ? ibm : NP - 1 - ibm);
    lFp = fpl * BLK + (pp.ss ? iap : NP - 1 - iap) + NP * (pp.st ? ibp : NP - 1 - ibp);
  }
  const int i = n & 3, j = (n >> 2) & 3, k = n >> 4;
  int liftP[DIM], liftM[DIM];  // jump-block entries of this node on the + / - face of each axis
#pragma unroll
  for (int r = 0; r < DIM; r++) {
    const FacePar fm = decode_face(kFacePar(kFaceMinus[r])), fpl = decode_face(kFacePar(kFacePlus[r]));
    const int iam = pick3(fm.as, i, j, k), ibm = pick3(fm.at, i, j, k);
    const int iap = pick3(fpl.as, i, j, k), ibp = pick3(fpl.at, i, j, k);
    liftM[r] = F_SJ + (kFaceMinus[r] * NF2 + (fm.ss ? iam : NP - 1 - iam) + NP * (fm.st ? ibm : NP - 1 - ibm)) * 6;
    liftP[r] = F_SJ + (kFacePlus[r] * NF2 + (fpl.ss ? iap : NP - 1 - iap) + NP * (fpl.st ? ibp : NP - 1 - ibp)) * 6;
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
      sts128(&sm[F_S0 + un], s[0], u);
      sts128(&sm[F_S1 + un], v, w);
      sm[F_S2 + un] = T;
      sts128(&sm[F_S3 + un], s[1], s[2]);
      sts128(&sm[F_S4 + un], s[3], s[4]);
    }
    __syncthreads();
    // ---- P1: D and both end-point extrapolations of the primitives: one DMMA per (axis, field, 8 lines)
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      double a0, a1, b0, b1;
      double2 x = lds128(&sm[F_S0 + bU[d]]);
      dmma884(a0, a1, afrag, x.x);
      dmma884(b0, b1, afrag, x.y);
      if (sAct) {
        sts128(&sm[o0[d]], a0, b0);
        sts128(&sm[o1[d]], a1, b1);
      }
      x = lds128(&sm[F_S1 + bU[d]]);
      dmma884(a0, a1, afrag, x.x);
      dmma884(b0, b1, afrag, x.y);
      if (sAct) {
        sts128(&sm[o0[d] + pst], a0, b0);
        sts128(&sm[o1[d] + pst], a1, b1);
      }
      dmma884(a0, a1, afrag, sm[F_S2 + bU[d]]);
      // dT_d: derivative pairs 6 (d = 0, 1) / 7 (d = 2); trace: slot 4 of the jump-block entry
      const int td = isD ? (6 + (d >> 1) - 2 * d) * FPU + (d & 1) : 4;
      if (sAct) {
        sm[o0[d] + td] = a0;
        sm[o1[d] + td] = a1;
      }
    }
    double *blk0 = a.tr + static_cast<long long>(e) * 6 * BLK;
    if (n < 48 && (mode & FUSED_WRITE)) {  // traces of rho u, rho v, rho w, rho E, straight to the trace blocks
      const double2 p0 = lds128(&sm[F_S3 + lU[0]]), p1 = lds128(&sm[F_S3 + lU[1]]), p2 = lds128(&sm[F_S3 + lU[2]]),
                    p3 = lds128(&sm[F_S3 + lU[3]]);
      blk0[lFm + 1 * NF2] = LB0(0) * p0.x + LB0(1) * p1.x + LB0(2) * p2.x + LB0(3) * p3.x;
      blk0[lFp + 1 * NF2] = LB1(0) * p0.x + LB1(1) * p1.x + LB1(2) * p2.x + LB1(3) * p3.x;
      blk0[lFm + 2 * NF2] = LB0(0) * p0.y + LB0(1) * p1.y + LB0(2) * p2.y + LB0(3) * p3.y;
      blk0[lFp + 2 * NF2] = LB1(0) * p0.y + LB1(1) * p1.y + LB1(2) * p2.y + LB1(3) * p3.y;
      const double2 q0 = lds128(&sm[F_S4 + lU[0]]), q1 = lds128(&sm[F_S4 + lU[1]]), q2 = lds128(&sm[F_S4 + lU[2]]),
                    q3 = lds128(&sm[F_S4 + lU[3]]);
      blk0[lFm + 3 * NF2] = LB0(0) * q0.x + LB0(1) * q1.x + LB0(2) * q2.x + LB0(3) * q3.x;
      blk0[lFp + 3 * NF2] = LB1(0) * q0.x + LB1(1) * q1.x + LB1(2) * q2.x + LB1(3) * q3.x;
      blk0[lFm + 4 * NF2] = LB0(0) * q0.y + LB0(1) * q1.y + LB0(2) * q2.y + LB0(3) * q3.y;
      blk0[lFp + 4 * NF2] = LB1(0) * q0.y + LB1(1) * q1.y + LB1(2) * q2.y + LB1(3) * q3.y;
    }
    __syncthreads();  // own traces (P1) are in the jump block
    // ---- P2: neighbour primitive traces at my face nodes -> jumps 1/2 (Up_nbr - Up_own) in place; boundary face:
    // Up2 = Up1 unless useBCinGrad (faceGradientIntegration.cpp:96-115).  Task = (face, face node); faces are taken in
    // axis pairs (4,2 | 1,3 | 0,5) so that on a structured mesh a warp's 32 tasks see one neighbour axis (the x-normal
    // pair takes the 256-bit loads).  The trace of rho (conserved = primitive) goes to the trace block from here.
#pragma unroll
    for (int rd = 0; rd < 2; rd++) {
      if (rd == 1 && half != 0) break;
      // warp 0: faces (4,2) then (0,5); warp 1: faces (1,3)
      const int pair = rd == 0 ? half : 2;
      const int hi = lane >> 4;
      const int lf = pair == 0 ? (hi ? 2 : 4) : (pair == 1 ? (hi ? 3 : 1) : (hi ? 5 : 0));
      const int ab = lane & 15, fa = ab & 3, fb = ab >> 2;
      double *pj = &sm[F_SJ + (lf * NF2 + ab) * 6];
      const double2 own01 = lds128(pj), own23 = lds128(pj + 2);
      const double own4 = pj[4];
      if (mode & FUSED_WRITE) blk0[lf * BLK + ab] = own01.x;
      const int nbr = sNbr[lf];
      if (nbr < 0) {
        if (a.bct.use_bc_in_grad && nbr <= -2) {
          const double pT[NEQ] = {own01.x, own01.y, own23.x, own23.y, own4};
          double pbc[NEQ];
          dry_bc_prim_for_gradient(a.bct.bc[-2 - nbr], pT, pbc);
          sts128(pj, 0.5 * (pbc[0] - pT[0]), 0.5 * (pbc[1] - pT[1]));
          sts128(pj + 2, 0.5 * (pbc[2] - pT[2]), 0.5 * (pbc[3] - pT[3]));
          pj[4] = 0.5 * (pbc[4] - pT[4]);
        } else {
          sts128(pj, 0.0, 0.0);
          sts128(pj + 2, 0.0, 0.0);
          pj[4] = 0.0;
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
      sts128(pj, 0.5 * (acc[0] - own01.x), 0.5 * (acc[1] - own01.y));
      sts128(pj + 2, 0.5 * (acc[2] - own23.x), 0.5 * (acc[3] - own23.y));
      pj[4] = 0.5 * (acc[4] - own4);
    }
    __syncthreads();
    // ---- P3: per node -- lifted physical gradient, viscous face fields, contravariant flux
    const double *A = sGeo;
    const double idet = sGeo[10];
    double G[NEQ][DIM];
    {
      double rg[NEQ][DIM];
      {
        const double *dr = &sm[F_DR + un];
#pragma unroll
        for (int r = 0; r < DIM; r++) {
          const double2 x0 = lds128(dr + 2 * r * FPU), x1 = lds128(dr + (2 * r + 1) * FPU);
          rg[0][r] = x0.x, rg[1][r] = x0.y, rg[2][r] = x1.x, rg[3][r] = x1.y;
        }
        const double2 t01 = lds128(dr + 6 * FPU);
        rg[4][0] = t01.x, rg[4][1] = t01.y, rg[4][2] = dr[7 * FPU];
      }
      const int idx[3] = {i, j, k};
#pragma unroll
      for (int r = 0; r < DIM; r++) {
        const double lwM = sLw[0][idx[r]], lwP = sLw[1][idx[r]];
        const double *jp = &sm[liftP[r]], *jm = &sm[liftM[r]];
        const double2 p01 = lds128(jp), p23 = lds128(jp + 2), m01 = lds128(jm), m23 = lds128(jm + 2);
        const double p4 = jp[4], m4 = jm[4];
        rg[0][r] += lwP * p01.x - lwM * m01.x;
        rg[1][r] += lwP * p01.y - lwM * m01.y;
        rg[2][r] += lwP * p23.x - lwM * m23.x;
        rg[3][r] += lwP * p23.y - lwM * m23.y;
        rg[4][r] += lwP * p4 - lwM * m4;
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
      double *sV = &sm[F_DR + un];  // this thread's own units: the viscous fields replace the derivatives in place
#pragma unroll
      for (int r = 0; r < DIM; r++) {
        const double nr[3] = {A[r + 0], A[r + 3], A[r + 6]};
        double sv[DIM];
#pragma unroll
        for (int p = 0; p < DIM; p++) sv[p] = sym[p][0] * nr[0] + sym[p][1] * nr[1] + sym[p][2] * nr[2];
        const double hT = g[4][0] * nr[0] + g[4][1] * nr[1] + g[4][2] * nr[2];
        sts128(sV + 2 * r * FPU, sv[0], sv[1]);
        sts128(sV + (2 * r + 1) * FPU, sv[2], hT);
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
      sV[6 * FPU] = divu;
      // pair fields 0-4 of the front region are this thread's own units: U / Up are dead (s[] lives in registers)
      double *sG = &sm[un];
      sts128(sG + 0 * FPU, G[0][0], G[1][0]);
      sts128(sG + 1 * FPU, G[2][0], G[3][0]);
      sts128(sG + 2 * FPU, G[0][1], G[1][1]);
      sts128(sG + 3 * FPU, G[2][1], G[3][1]);
      sts128(sG + 4 * FPU, G[0][2], G[1][2]);
    }
    __syncthreads();  // every lift has read the jump block: pair fields 5-7 may overwrite it; the viscous fields are complete
    {
      double *sG = &sm[un];
      sts128(sG + 5 * FPU, G[2][2], G[3][2]);
      sts128(sG + 6 * FPU, G[4][0], G[4][1]);
      sG[7 * FPU] = G[4][2];
    }
    // ---- P4a: traces of the normal-contracted viscous fields (axis d carries (s0,s1)_d, (s2,n.gradT)_d and div u)
    if (n < 48 && (mode & FUSED_WRITE)) {
      const double *v0 = &sm[F_DR + 2 * lax * FPU], *v1 = v0 + FPU, *v2 = &sm[F_DR + 6 * FPU];
      double tm[5] = {0, 0, 0, 0, 0}, tp[5] = {0, 0, 0, 0, 0};
#pragma unroll
      for (int m = 0; m < NP; m++) {
        const double2 x = lds128(v0 + lU[m]), y = lds128(v1 + lU[m]);
        const double z = v2[lU[m]];
        tm[0] += LB0(m) * x.x, tm[1] += LB0(m) * x.y, tm[2] += LB0(m) * y.x, tm[3] += LB0(m) * y.y, tm[4] += LB0(m) * z;
        tp[0] += LB1(m) * x.x, tp[1] += LB1(m) * x.y, tp[2] += LB1(m) * y.x, tp[3] += LB1(m) * y.y, tp[4] += LB1(m) * z;
      }
#pragma unroll
      for (int v = 0; v < 5; v++) {
        blk0[lFm + (NEQ + v) * NF2] = v < 4 ? -tm[v] : tm[v];  // outward normal of the - face = -A[d,:]
        blk0[lFp + (NEQ + v) * NF2] = tp[v];
      }
    }
    __syncthreads();
    if (mode & FUSED_WRITE) {
      // ---- P4b: transposed derivative of G along each axis, in place (a warp reads and writes the same 8 lines)
#pragma unroll
      for (int r = 0; r < DIM; r++) {
        // rows 0-3: o0 / o1 hold F_DR + 2 r FPU + the D-fragment unit
        const int u0 = o0[r] - (F_DR + 2 * r * FPU), u1 = o1[r] - (F_DR + 2 * r * FPU);
#pragma unroll
        for (int h2 = 0; h2 < 2; h2++) {
          double *fld = &sm[(2 * r + h2) * FPU];
          const double2 x = lds128(fld + bU[r]);
          double a0, a1, b0, b1;
          dmma884(a0, a1, afragT, x.x);
          dmma884(b0, b1, afragT, x.y);
          if (isD) {
            sts128(fld + u0, a0, b0);
            sts128(fld + u1, a1, b1);
          }
        }
        double *f4 = &sm[(6 + (r >> 1)) * FPU + (r & 1)];
        double a0, a1;
        dmma884(a0, a1, afragT, f4[bU[r]]);
        if (isD) {
          f4[u0] = a0;
          f4[u1] = a1;
        }
      }
      __syncthreads();
      // ---- P5: volume part of dU/dt = Me^-1 (sum_r D^T_r G_r), Me = diag(w |J|)
      const double *sG = &sm[un];
      const double2 g00 = lds128(sG), g01 = lds128(sG + 2 * FPU), g02 = lds128(sG + 4 * FPU);
      const double2 g20 = lds128(sG + FPU), g21 = lds128(sG + 3 * FPU), g22 = lds128(sG + 5 * FPU);
      const double2 g4a = lds128(sG + 6 * FPU);
      const double g4b = sG[7 * FPU];
      a.y[o + 0 * N] = (g00.x + g01.x + g02.x) * idet;
      a.y[o + 1 * N] = (g00.y + g01.y + g02.y) * idet;
      a.y[o + 2 * N] = (g20.x + g21.x + g22.x) * idet;
      a.y[o + 3 * N] = (g20.y + g21.y + g22.y) * idet;
      a.y[o + 4 * N] = (g4a.x + g4a.y + g4b) * idet;
    }
    __syncthreads();  // the next element overwrites sm / sGeo / sNbr
  }
  if (lane == 0 && mcs_bits > __ldcg(a.maxCharBits)) atomicMax(a.maxCharBits, mcs_bits);
#undef LB0
#undef LB1
}

// lift_kernel: y = y_vol + Me^-1 sum_faces (+-) l_c(face) R_face   (face_integrator.cpp:348-350, rhs_operator.cpp:432-448);
// y_vol is what elem_fused_kernel left in a.y.  RK: fused Runge-Kutta stage update as in elem_resid_kernel.
// Streaming kernel (110 B per node) whose only hazard is latency: the face ids of an element must arrive before its
// face residuals can be addressed.  Persistent CTAs of four elements; the ids of the NEXT four elements are fetched
// while the 35 independent loads per node of the current ones are in flight (absent faces read slot 0 with a zero
// coefficient, so no load hides behind a branch).
template <bool RK>
__global__ void __launch_bounds__(256) lift_kernel(KernelArgs a, int elem_begin, int elem_count) {
  constexpr int NP = 4, ND = 64, NF2 = 16;
  __shared__ double sLb[2][NP], sIw[ND];
  __shared__ int sFace[2][4][6], sFcode[2][4][6];
  __shared__ double sIdet[2][4];
  const int le = threadIdx.x >> 6, n = threadIdx.x & 63;
  const int i = n & 3, j = (n >> 2) & 3, k = n >> 4;
  if (threadIdx.x < 2 * NP) sLb[threadIdx.x / NP][threadIdx.x % NP] = c_T.lb[threadIdx.x / NP][threadIdx.x % NP];
  if (threadIdx.x < ND) sIw[n] = 1.0 / (c_T.wn[i] * c_T.wn[j] * c_T.wn[k]);
  const long long N = a.N;
  const int nquad = (elem_count + 3) / 4;
  int q = blockIdx.x;
  // ids of the first quad
  int pf = -1, pc = 0;
  double pd = 0.0;
  {
    const int slot = q * 4 + le;
    if (q < nquad && slot < elem_count) {
      const int e = elem_begin + slot;
      if (n < 6) pf = __ldg(&a.el_face[e * 6 + n]), pc = __ldg(&a.el_face_code[e * 6 + n]);
      if (n == 32) pd = __ldg(&a.geo[static_cast<long long>(e) * GEO + 10]);
    }
  }
  for (int buf = 0; q < nquad; q += gridDim.x, buf ^= 1) {
    if (n < 6) sFace[buf][le][n] = pf, sFcode[buf][le][n] = pc;
    if (n == 32) sIdet[buf][le] = pd;
    __syncthreads();  // also orders the reuse of buffer buf (two iterations ago) behind everybody's reads
    {  // prefetch the ids of the next quad of this CTA
      const int slot = (q + gridDim.x) * 4 + le;
      if (q + gridDim.x < nquad && slot < elem_count) {
        const int e = elem_begin + slot;
        if (n < 6) pf = __ldg(&a.el_face[e * 6 + n]), pc = __ldg(&a.el_face_code[e * 6 + n]);
        if (n == 32) pd = __ldg(&a.geo[static_cast<long long>(e) * GEO + 10]);
      }
    }
    const int slot = q * 4 + le;
    if (slot >= elem_count) continue;
    const int e = elem_begin + slot;
    const long long o = static_cast<long long>(e) * ND + n;
    double yv[NEQ];
#pragma unroll
    for (int eq = 0; eq < NEQ; eq++) yv[eq] = a.y[o + eq * N];
    double R[6][NEQ], coef[6];
#pragma unroll
    for (int lf = 0; lf < 6; lf++) {
      const int fc = sFace[buf][le][lf];
      const int code = sFcode[buf][le][lf];
      const int side = code & 1;
      const FacePar fp = decode_face(kFacePar(lf));
      const int ia = pick3(fp.as, i, j, k), ib = pick3(fp.at, i, j, k), c = pick3(fp.an, i, j, k);
      const int fa = fp.ss ? ia : NP - 1 - ia, fb = fp.st ? ib : NP - 1 - ib;
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
    const double im = sIw[n] * sIdet[buf][le];
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
}

}  // namespace tpsb
