// CUDA kernels of the DG right-hand side (sm_100a, FP64).  One evaluation of RHSoperator::Mult
// (src/rhs_operator.cpp:343-464) is four launches:
//
//   prim_kernel        Up = prim(U)                                   (updatePrimitives, :623-651)
//   grad_kernel        gradUp = Me^-1 (Ke Up + face jumps)            (Gradients::computeGradients,
//                                                                      src/gradients.cpp:144-232)
//   face_flux_kernel   per face: traces -> 5x5 quadrature points -> Rusanov + averaged viscous flux
//                      -> projected back onto the face nodes, computed ONCE per face
//                                                                     (FaceIntegrator, src/face_integrator.cpp:194-352)
//   elem_resid_kernel  nodal flux (GetFlux :493-559), collocated weak divergence (Aflux, :379-391),
//                      lift of the six face residuals, diagonal Me^-1 (:432-448), max char speed
//
// Sum factorisation: with Gauss-Legendre nodes AND Gauss-Legendre quadrature the reference's dense
// Me / Ke / Aflux blocks are exactly w_j|J_j| delta_ij, w_j|J_j| d_d phi_k(x_j) and
// w_k [d_xi phi_j(x_k) adj J_k]_d, i.e. 1-D differentiation along lines (tables.hpp).
#pragma once
#include <cuda_runtime.h>

#include "physics.cuh"
#include "tables.hpp"

namespace tpsb {

constexpr int NFLD = NEQ * (1 + DIM);  // U (5) + gradUp (15) fields carried to the faces

// c_T (the __constant__ copy of RefTables) is defined in rhs_kernels_impl.cuh; the library is a single
// CUDA translation unit (tpsb200.cu), so no relocatable device code is needed.

struct KernelArgs {
  // sizes
  int NE;         // local elements
  int NEH;        // halo (face-neighbour) elements
  long long N;    // NE * dof
  long long NH;   // NEH * dof
  int NFint;      // faces with two sides (interior + shared)
  int ND;         // dofs per element
  int vec_ok;     // all field base pointers 32-byte aligned: 256-bit trace loads allowed
  PhysParams phys;
  // geometry / connectivity (device)
  const double *vx;        // [(NE+NEH)][8][3]
  const int *nbr_elem;     // [NE][6] neighbour element (>= NE: halo); boundary face: -2 - (index into bct.bc)
  const int *nbr_code;     // [NE][6] (nbr local face) | (perm code << 3): own face coords -> neighbour face coords
  const int *face_el1, *face_el2, *face_inf1, *face_inf2;  // [NFint] compacted two-sided faces
  const int *el_face;      // [NE][6] compact face id or -1
  const int *el_face_code; // [NE][6] side | (perm code own -> face coords << 1)
  // fields
  const double *U;         // [NEQ][N]
  const double *Uhalo;     // [NEH][NEQ][dof]  (element-major: one contiguous block per peer)
  double *Up;              // [NEQ][N]
  double *UpHalo;          // [NEH][NEQ][dof]
  double *gradUp;          // [DIM][NEQ][N]
  const double *gradUpHalo;  // [NEH][DIM*NEQ][dof]
  double *faceRes;         // [NFint][NEQ][np*np]
  double *y;               // [NEQ][N]
  unsigned long long *maxCharBits;  // atomicMax target (bit pattern of a non-negative double)
  // fast (all-affine) path, rhs_fast.cuh
  const double *geo;       // [NE][12] adj(J) (A[r + 3 d]), det, 1/det, pad
  double *tr;              // [6 NE + shared faces][10][np*np] face-trace blocks
  const int4 *face_desc;   // [NFint] {block of side 1, block of side 2, perm code Elem2 face coords -> face coords, 0}
  const double *elem_delta;  // [NE+NEH] h_min / order (Mesh::GetElementSize(e, 1) / order): SGS models
  const double *face_nor;  // [NFint][4] CalcOrtho normal (Elem1 -> Elem2, area weighted) and its magnitude
  // boundary faces (BCintegrator): face k lifts into faceRes slot NFint + k
  int NFbdr;
  const int *bdr_el1, *bdr_lf, *bdr_bc;  // [NFbdr] element, local face, index into bct.bc
  BcTable bct;
  // Runge-Kutta stage update fused into the residual kernel (tpsb_ode_step): with k = dU/dt of this evaluation,
  // Y = X + A k and Z = (zacc ? Z : X) + B k are written instead of k itself (Y may be the vector being evaluated: a
  // CTA reads only its own nodes of U and writes them last).  X == nullptr: plain Mult, y = k.
  struct RkStage {
    const double *X;
    double *Y, *Z;
    double A, B;
    int zacc;
  } rk;
};

// ---- geometry: trilinear hexahedron from its 8 vertices (mesh nodes of order 1) ----
// J is column-major: J[i + 3*j] = d x_i / d xi_j
__device__ __forceinline__ void hex_jacobian(const double *v, double x, double y, double z, double *J) {
  const double x0 = 1.0 - x, y0 = 1.0 - y, z0 = 1.0 - z;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double X0 = v[0 * 3 + i], X1 = v[1 * 3 + i], X2 = v[2 * 3 + i], X3 = v[3 * 3 + i];
    const double X4 = v[4 * 3 + i], X5 = v[5 * 3 + i], X6 = v[6 * 3 + i], X7 = v[7 * 3 + i];
    J[i + 0] = y0 * z0 * (X1 - X0) + y * z0 * (X2 - X3) + y0 * z * (X5 - X4) + y * z * (X6 - X7);
    J[i + 3] = x0 * z0 * (X3 - X0) + x * z0 * (X2 - X1) + x0 * z * (X7 - X4) + x * z * (X6 - X5);
    J[i + 6] = x0 * y0 * (X4 - X0) + x * y0 * (X5 - X1) + x * y * (X6 - X2) + x0 * y * (X7 - X3);
  }
}
__device__ __forceinline__ double det3(const double *J) {
  return J[0] * (J[4] * J[8] - J[5] * J[7]) - J[3] * (J[1] * J[8] - J[2] * J[7]) + J[6] * (J[1] * J[5] - J[2] * J[4]);
}
// A = adj(J) = det(J) inv(J), column-major A[r + 3*d]
__device__ __forceinline__ void adj3(const double *J, double *A) {
  A[0] = J[4] * J[8] - J[7] * J[5];
  A[3] = J[6] * J[5] - J[3] * J[8];
  A[6] = J[3] * J[7] - J[6] * J[4];
  A[1] = J[7] * J[2] - J[1] * J[8];
  A[4] = J[0] * J[8] - J[6] * J[2];
  A[7] = J[6] * J[1] - J[0] * J[7];
  A[2] = J[1] * J[5] - J[4] * J[2];
  A[5] = J[3] * J[2] - J[0] * J[5];
  A[8] = J[0] * J[4] - J[3] * J[1];
}
// CalcOrtho of the face Jacobian at face point (s,t): face vertices Xf[4][3] in the face's own order
__device__ __forceinline__ void face_normal(const double *Xf, double s, double t, double *nor) {
  double ts[3], tt[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    ts[i] = (1.0 - t) * (Xf[1 * 3 + i] - Xf[0 * 3 + i]) + t * (Xf[2 * 3 + i] - Xf[3 * 3 + i]);
    tt[i] = (1.0 - s) * (Xf[3 * 3 + i] - Xf[0 * 3 + i]) + s * (Xf[2 * 3 + i] - Xf[1 * 3 + i]);
  }
  nor[0] = ts[1] * tt[2] - ts[2] * tt[1];
  nor[1] = ts[2] * tt[0] - ts[0] * tt[2];
  nor[2] = ts[0] * tt[1] - ts[1] * tt[0];
}

// physical point of a bilinear face / trilinear hexahedron
__device__ __forceinline__ void face_point(const double *Xf, double s, double t, double *x) {
#pragma unroll
  for (int i = 0; i < 3; i++)
    x[i] = (1.0 - s) * (1.0 - t) * Xf[0 * 3 + i] + s * (1.0 - t) * Xf[1 * 3 + i] + s * t * Xf[2 * 3 + i] + (1.0 - s) * t * Xf[3 * 3 + i];
}
__device__ __forceinline__ void hex_point(const double *v, double x, double y, double z, double *X) {
  const double x0 = 1.0 - x, y0 = 1.0 - y, z0 = 1.0 - z;
#pragma unroll
  for (int i = 0; i < 3; i++)
    X[i] = z0 * (y0 * (x0 * v[0 * 3 + i] + x * v[1 * 3 + i]) + y * (x * v[2 * 3 + i] + x0 * v[3 * 3 + i])) +
           z * (y0 * (x0 * v[4 * 3 + i] + x * v[5 * 3 + i]) + y * (x * v[6 * 3 + i] + x0 * v[7 * 3 + i]));
}

__device__ __forceinline__ void atomic_max_double(unsigned long long *addr, double v) {
  // v >= 0: IEEE bit patterns of non-negative doubles are ordered like unsigned integers
  atomicMax(addr, static_cast<unsigned long long>(__double_as_longlong(v)));
}

// ------------------------------------------------------------------------------------------------
// prim_kernel: one thread per node (local nodes, then halo nodes)
__global__ void prim_kernel(KernelArgs a, int halo);
__global__ void pack_kernel(int nsend, int nd, int nfld, long long N, const int *send_elems, const double *src,
                            double *dst);

// grad_kernel / face_flux_kernel / elem_resid_kernel are templates on NP = p+1; see rhs_kernels.cu
template <int NP, int EPB, int MINB>
__global__ void grad_kernel(KernelArgs a, int elem_begin, int elem_count, const int *elem_list);
template <int NP, int FPB, int NT, bool BDR, bool MOD>
__global__ void face_flux_kernel(KernelArgs a, int face_begin, int face_count, const int *face_list);
template <int NP, int EPB, int MINB, bool AFF, bool MOD, bool RK>
__global__ void elem_resid_kernel(KernelArgs a, int elem_begin, int elem_count);

// y = x + a*k ; z = x + b*k  etc. for the ODE stages
__global__ void axpy2_kernel(long long n, const double *x, const double *k, double a, double *y, double b, double *z,
                             int z_accumulate);

}  // namespace tpsb
