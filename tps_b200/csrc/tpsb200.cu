// libtpsb200: context management, launch orchestration and the C ABI declared in include/tpsb200.h.
// Single CUDA translation unit (kernels are included below); host mesh tables live in meshkit.cpp.
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/tpsb200.h"
#include "rhs_kernels_impl.cuh"
#include "rhs_fast.cuh"
#include "rhs_fused.cuh"
#include "rhs_generic.cuh"
#include "rhs_forcing.cuh"

using namespace tpsb;

static thread_local std::string g_create_error;
constexpr int MAX_DEV = 64;
// The reference-element tables of orders 1..3 sit side by side in each device's __constant__ memory (c_Tab, one slot
// per order, uploaded once per device by the first tpsb_create on it); per-device flags, because cudaFuncSetAttribute
// and __constant__ memory are per device.
static std::mutex g_dev_mutex;
static bool g_tables_uploaded[MAX_DEV] = {};

struct tpsb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;       // caller's stream: all compute is enqueued here
  cudaStream_t comm_stream = nullptr;  // NCCL face-neighbour exchange
  cudaEvent_t ev_pack = nullptr, ev_recvU = nullptr, ev_recvG = nullptr;
  int order = 0, np = 0, nd = 0;
  int NE = 0, NEH = 0, NF = 0, NFint = 0, NFlocal = 0;  // NFlocal: two-sided faces with both elements local
  long long N = 0, NH = 0;
  RefTables T;
  PhysParams phys;
  std::vector<int> element_to_faces;  // reference layout, stride 7
  // device tables
  double *d_vx = nullptr, *d_elem_delta = nullptr;
  double hmin = 0.0;            // min element size h_min over the local elements (adaptive time step)
  int *d_nan_count = nullptr;   // tpsb_solve_step: NaN counter
  int *d_nbr_elem = nullptr, *d_nbr_code = nullptr, *d_face_el1 = nullptr, *d_face_el2 = nullptr,
      *d_face_inf1 = nullptr, *d_face_inf2 = nullptr, *d_el_face = nullptr, *d_el_face_code = nullptr;
  int *d_elem_list = nullptr;  // interior elements first, then elements touching a shared face
  int n_int_elems = 0, n_pb_elems = 0;
  // device fields owned by the context
  double *d_Up = nullptr, *d_gradUp = nullptr, *d_faceRes = nullptr;
  double *d_Uhalo = nullptr, *d_UpHalo = nullptr, *d_gradUpHalo = nullptr, *d_sendU = nullptr, *d_sendG = nullptr;
  unsigned long long *d_maxBits = nullptr;
  double *d_mcs = nullptr;
  // generic tensor-product path (rhs_generic.cuh): 2-D, Gauss-Lobatto, ...
  bool generic = false;
  int dim = 3, neq = NEQ;
  const double *sol_view = nullptr;  // U_ of the reference's forcing terms (tpsb_set_solution_view)
  MixParams mix_host;                // host copy of the device MixParams (re-uploaded by tpsb_set_reaction_rate_field)
  GenArgs gen;
  std::vector<void *> gen_allocs;
  // boundary faces (BCintegrator)
  int NFbdr = 0;
  int *d_bdr_el1 = nullptr, *d_bdr_lf = nullptr, *d_bdr_bc = nullptr;
  BcTable bct;
  // fast (all-affine) path
  bool fast = false;
  bool fused = false;               // p = 3 fast path in three launches (rhs_fused.cuh): Up / gradUp are not materialised
  const double *fields_x = nullptr;  // fused path: vector of the last evaluation (tpsb_get_fields refreshes from it)
  bool fields_stale = false;
  double *d_geo = nullptr, *d_tr = nullptr, *d_face_nor = nullptr, *d_sendTr = nullptr;
  int4 *d_face_desc = nullptr;
  int *d_send_blk = nullptr;
  std::vector<int> sendblk_offset, recvblk_offset;  // per peer, in trace blocks
  int n_send_blk = 0;
  cudaEvent_t ev_recvT = nullptr;
  // halo description
  ncclComm_t comm = nullptr;
  std::vector<int> nbr_rank, send_offset, recv_offset;
  int *d_send_elems = nullptr;
  int n_send = 0;
  // ODE / host-staging work vectors (lazy)
  double *d_k = nullptr, *d_yv = nullptr, *d_z = nullptr, *d_hx = nullptr, *d_hy = nullptr;
  // chunked host-buffer pipeline of tpsb_rhs_mult_host (single rank, periodic fast path): host->device copies,
  // kernels and device->host copies of different element chunks overlap on three streams
  struct PipeOp {
    int kind, chunk;  // 0 prim (after the chunk's copy landed), 1 gradient, 2 face fluxes, 3 residual + copy out
  };
  int pipe_chunks = 0;
  std::vector<int> pipe_eb, pipe_fb, pipe_bb;  // element / two-sided face / boundary face range of each chunk
  std::vector<PipeOp> pipe_ops;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_out;
  cudaEvent_t ev_pipe0 = nullptr, ev_pipe1 = nullptr;
  // tpsb_ode_step: one Runge-Kutta step captured as a CUDA graph and replayed (keyed by solution pointer, dt, scheme)
  cudaGraphExec_t ode_exec = nullptr;
  cudaStream_t ode_stream = nullptr;
  cudaEvent_t ode_ev0 = nullptr, ode_ev1 = nullptr;
  double *ode_U = nullptr;
  double ode_dt = 0.0;
  int ode_scheme = 0;
  long long ode_launches = 0;
  KernelArgs::RkStage rk = {nullptr, nullptr, nullptr, 0.0, 0.0, 0};  // stage update fused into the residual kernel
  // forcing terms (tpsb_add_forcing), applied after Me^-1 in registration order
  std::vector<ForcingDev> forcings;
  // non-reflecting inlets / outlets (generic path): host copies of the patch descriptors, their attributes, BoundaryCondition::dt
  std::vector<GenNrPatch> nr_patches;
  std::vector<int> nr_attr;
  double bc_dt = 0.0;
  bool lte_radiation = false;  // LTE fluid with a net-emission-coefficient table
  double *d_xiN3 = nullptr;  // [dof][3] reference coordinates of the nodes (3-D dry-air paths; the generic path has its own)
  bool forcing_needs_grad = false;
  long long launches = 0;
  int tune[3] = {0, 0, 0};
  int num_sms = 148, face_ctas_per_sm = 5;
  // per-kernel device timers (tpsb_set_profiling)
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events[TPSB_NUM_KERNEL_CLASSES];
  double prof_ms[TPSB_NUM_KERNEL_CLASSES] = {0, 0, 0, 0, 0, 0};
  long long prof_count[TPSB_NUM_KERNEL_CLASSES] = {0, 0, 0, 0, 0, 0};
  std::string err;
};

static int fail(tpsb_ctx *c, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c)
    c->err = buf;
  else
    g_create_error = buf;
  return code;
}

static std::string setup_halo_desc(tpsb_ctx *c, const tpsb_halo_desc *halo, int NEH);
enum { K_PRIM = 0, K_GRAD, K_FACE, K_RESID, K_PACK, K_AXPY };
struct ProfScope {  // brackets one launch with events when profiling is on
  tpsb_ctx *c;
  int k;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  ProfScope(tpsb_ctx *c_, int k_);
  ~ProfScope();
};

#define CU(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) return fail(ctx, TPSB_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
  } while (0)
#define NC(call)                                                                                     \
  do {                                                                                               \
    ncclResult_t r_ = (call);                                                                        \
    if (r_ != ncclSuccess) return fail(ctx, TPSB_ENCCL, "%s failed: %s", #call, ncclGetErrorString(r_)); \
  } while (0)

ProfScope::ProfScope(tpsb_ctx *c_, int k_) : c(c_), k(k_) {
  c->launches++;
  if (!c->profiling) return;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, c->stream);
}
ProfScope::~ProfScope() {
  if (!c->profiling) return;
  cudaEventRecord(e1, c->stream);
  c->prof_events[k].emplace_back(e0, e1);
}

template <class T>
static cudaError_t upload(T **dst, const std::vector<T> &src) {
  *dst = nullptr;
  if (src.empty()) return cudaSuccess;
  cudaError_t e = cudaMalloc(reinterpret_cast<void **>(dst), src.size() * sizeof(T));
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}

// h_min of a trilinear hexahedron as MFEM's Mesh::GetElementSize(e, 1) defines it: the smallest singular value of the
// Jacobian at the element centre (the cube is its own "perfect" element).  One-sided Jacobi (Hestenes) SVD: rotate
// column pairs until orthogonal; the singular values are the column norms.
static double hex_min_size(const double *v) {
  double J[3][3];  // J[i][j] = d x_i / d xi_j at (1/2, 1/2, 1/2)
  static const int lo[3][4] = {{0, 3, 4, 7}, {0, 1, 4, 5}, {0, 1, 2, 3}}, hi[3][4] = {{1, 2, 5, 6}, {3, 2, 7, 6}, {4, 5, 6, 7}};
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++) {
      double s = 0;
      for (int q = 0; q < 4; q++) s += v[hi[j][q] * 3 + i] - v[lo[j][q] * 3 + i];
      J[i][j] = 0.25 * s;
    }
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        double a = 0, b = 0, g = 0;
        for (int i = 0; i < 3; i++) a += J[i][p] * J[i][p], b += J[i][q] * J[i][q], g += J[i][p] * J[i][q];
        if (std::fabs(g) <= 1e-17 * std::sqrt(a * b)) continue;
        off = std::max(off, std::fabs(g) / std::sqrt(a * b));
        const double zeta = (b - a) / (2.0 * g);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
        for (int i = 0; i < 3; i++) {
          const double xp = J[i][p], xq = J[i][q];
          J[i][p] = cs * xp - sn * xq;
          J[i][q] = sn * xp + cs * xq;
        }
      }
    if (off == 0) break;
  }
  double m = 1e300;
  for (int j = 0; j < 3; j++) m = std::min(m, std::sqrt(J[0][j] * J[0][j] + J[1][j] * J[1][j] + J[2][j] * J[2][j]));
  return m;
}

// the same for a bilinear quadrilateral: smaller singular value of the 2 x 2 centre Jacobian, closed form
static double quad_min_size(const double *v) {
  const double a = 0.5 * ((v[2] - v[0]) + (v[4] - v[6])), c = 0.5 * ((v[3] - v[1]) + (v[5] - v[7]));  // d x / d xi
  const double b = 0.5 * ((v[6] - v[0]) + (v[4] - v[2])), d = 0.5 * ((v[7] - v[1]) + (v[5] - v[3]));  // d x / d eta
  const double s1 = std::hypot(a + d, c - b), s2 = std::hypot(a - d, c + b);  // sigma_max,min = (s1 +- s2) / 2
  return 0.5 * std::fabs(s1 - s2);
}

static bool element_is_affine(const double *v) {
  double scale = 0;
  for (int i = 0; i < 3; i++) scale = std::max(scale, std::fabs(v[6 * 3 + i] - v[i]));
  const double tol = 1e-12 * std::max(scale, 1e-300);
  for (int i = 0; i < 3; i++) {
    const double X0 = v[i], e1 = v[3 + i] - X0, e2 = v[9 + i] - X0, e3 = v[12 + i] - X0;
    if (std::fabs(v[2 * 3 + i] - (X0 + e1 + e2)) > tol) return false;
    if (std::fabs(v[5 * 3 + i] - (X0 + e1 + e3)) > tol) return false;
    if (std::fabs(v[7 * 3 + i] - (X0 + e2 + e3)) > tol) return false;
    if (std::fabs(v[6 * 3 + i] - (X0 + e1 + e2 + e3)) > tol) return false;
  }
  return true;
}

// ---- generic path setup ---------------------------------------------------------------------------
namespace {

const int G_HEX_FACE_VERT[6][4] = {{3, 2, 1, 0}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
const int G_QUAD_EDGE_VERT[4][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}};
const double G_QUAD_VERT[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};

// face reference point -> element reference point [MFEM GetLocalQuadToHexTransformation / GetLocalSegToQuadTransformation]
void g_loc_map(int dim, int lf, int ori, const double *st, double *xi, double *dloc) {
  if (dim == 3) {
    const double s = st[0], t = st[1];
    const int *hv = G_HEX_FACE_VERT[lf];
    const int *qo = QUAD_ORIENT[ori];
    const double Nq[4] = {(1 - s) * (1 - t), s * (1 - t), s * t, (1 - s) * t};
    const double dNs[4] = {-(1 - t), (1 - t), t, -t}, dNt[4] = {-(1 - s), -s, s, (1 - s)};
    for (int i = 0; i < 3; i++) {
      double a = 0, b = 0, c = 0;
      for (int j = 0; j < 4; j++) {
        const double vj = HEX_VERT[hv[qo[j]]][i];
        a += Nq[j] * vj, b += dNs[j] * vj, c += dNt[j] * vj;
      }
      xi[i] = a, dloc[i] = b, dloc[i + 3] = c;
    }
  } else {
    const int *ev = G_QUAD_EDGE_VERT[lf];
    const int i0 = ori ? 1 : 0, i1 = ori ? 0 : 1;
    for (int i = 0; i < 2; i++) {
      const double v0 = G_QUAD_VERT[ev[i0]][i], v1 = G_QUAD_VERT[ev[i1]][i];
      xi[i] = (1 - st[0]) * v0 + st[0] * v1;
      dloc[i] = v1 - v0;
    }
  }
}

// tensor Lagrange basis on nodes xn (np per direction) at reference point xi: values and reference derivatives
void g_shape(int dim, int np, const double *xn, const double *xi, double *phi, double *dphi) {
  long double v[3][8], d[3][8];
  for (int a = 0; a < dim; a++) lagrange_ld(xn, np, xi[a], v[a], d[a]);
  int dof = 1;
  for (int a = 0; a < dim; a++) dof *= np;
  for (int n = 0; n < dof; n++) {
    const int idx[3] = {n % np, (n / np) % np, n / (np * np)};
    long double s = 1;
    for (int a = 0; a < dim; a++) s *= v[a][idx[a]];
    phi[n] = static_cast<double>(s);
    if (dphi)
      for (int a = 0; a < dim; a++) {
        long double q = d[a][idx[a]];
        for (int b = 0; b < dim; b++)
          if (b != a) q *= v[b][idx[b]];
        dphi[n * dim + a] = static_cast<double>(q);
      }
  }
}

template <class T>
cudaError_t g_upload(tpsb_ctx *c, const T **dst, const std::vector<T> &src) {
  T *p = nullptr;
  *dst = nullptr;
  if (src.empty()) return cudaSuccess;
  cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p), src.size() * sizeof(T));
  if (e != cudaSuccess) return e;
  c->gen_allocs.push_back(p);
  *dst = p;
  return cudaMemcpy(p, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}

// in-place inverse of an SPD matrix (row-major n x n) by Cholesky: A = L L^T, A^-1 = L^-T L^-1
bool g_spd_inverse(std::vector<double> &A, int n) {
  std::vector<long double> L(static_cast<size_t>(n) * n, 0.0L), Li(static_cast<size_t>(n) * n, 0.0L);
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      long double s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k];
      if (i == j) {
        if (s <= 0) return false;
        L[i * n + i] = sqrtl(s);
      } else {
        L[i * n + j] = s / L[j * n + j];
      }
    }
  for (int c = 0; c < n; c++)  // Li = L^-1 (lower triangular), column by column
    for (int i = c; i < n; i++) {
      long double s = (i == c) ? 1.0L : 0.0L;
      for (int k = c; k < i; k++) s -= L[i * n + k] * Li[k * n + c];
      Li[i * n + c] = s / L[i * n + i];
    }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      long double s = 0;
      for (int k = std::max(i, j); k < n; k++) s += Li[k * n + i] * Li[k * n + j];
      A[i * n + j] = static_cast<double>(s);
    }
  return true;
}

// Everything the generic kernels need; returns an error string (empty on success).
std::string create_generic(tpsb_ctx *c, const tpsb_mesh_maps *maps, const tpsb_space_desc *space, const tpsb_physics *phys,
                           const tpsb_bc_set *bcs, const tpsb_halo_desc *halo) {
  const int dim = maps->dim, p = space->order, np = p + 1, NE = c->NE, NF = maps->num_faces;
  const int nv = 1 << dim, nfe = 2 * dim, nori = dim == 3 ? 8 : 2;
  int dof = 1;
  for (int d = 0; d < dim; d++) dof *= np;
  // 1-D nodes and rules (M2ulPhyS.cpp:558-572; MFEM rule sizes: GL n = order/2+1 with order|1, GLL n = order/2+2)
  double xn[8], wtmp[8], xv[8], wv[8], xf[8], wf[8];
  if (space->basis_type == 0)
    gauss_legendre01(np, xn, wtmp);
  else
    gauss_lobatto01(np, xn, wtmp);
  const int ov = 2 * p, of = dim - 1 + 2 * p;
  const int nqv1 = space->int_rule_type == 0 ? (ov | 1) / 2 + 1 : ov / 2 + 2;
  const int nqf1 = space->int_rule_type == 0 ? (of | 1) / 2 + 1 : of / 2 + 2;
  if (nqv1 > 8 || nqf1 > 8) return "quadrature rule too large";
  if (space->int_rule_type == 0) {
    gauss_legendre01(nqv1, xv, wv);
    gauss_legendre01(nqf1, xf, wf);
  } else {
    gauss_lobatto01(nqv1, xv, wv);
    gauss_lobatto01(nqf1, xf, wf);
  }
  int nqv = 1, nqf = 1;
  for (int d = 0; d < dim; d++) nqv *= nqv1;
  for (int d = 0; d < dim - 1; d++) nqf *= nqf1;
  std::vector<double> phiV(static_cast<size_t>(nqv) * dof), dphiV(static_cast<size_t>(nqv) * dof * dim), wV(nqv), xiV(static_cast<size_t>(nqv) * dim);
  for (int q = 0; q < nqv; q++) {
    const int qi[3] = {q % nqv1, (q / nqv1) % nqv1, q / (nqv1 * nqv1)};
    double xi[3] = {0, 0, 0}, w = 1;
    for (int d = 0; d < dim; d++) xi[d] = xv[qi[d]], w *= wv[qi[d]];
    for (int d = 0; d < dim; d++) xiV[q * dim + d] = xi[d];
    wV[q] = w;
    g_shape(dim, np, xn, xi, &phiV[static_cast<size_t>(q) * dof], &dphiV[static_cast<size_t>(q) * dof * dim]);
  }
  std::vector<double> xiN(static_cast<size_t>(dof) * dim);  // reference coordinates of the nodes (x fastest)
  for (int n = 0; n < dof; n++) {
    const int ni[3] = {n % np, (n / np) % np, n / (np * np)};
    for (int d = 0; d < dim; d++) xiN[static_cast<size_t>(n) * dim + d] = xn[ni[d]];
  }
  const bool axisym = dim == 2 && space->nvel == 3;
  const int ncode = nfe * nori;
  std::vector<double> phiF(static_cast<size_t>(ncode) * nqf * dof), xiF(static_cast<size_t>(ncode) * nqf * dim), dlocF(static_cast<size_t>(ncode) * dim * (dim - 1)), wF(nqf);
  for (int q = 0; q < nqf; q++) wF[q] = dim == 3 ? wf[q % nqf1] * wf[q / nqf1] : wf[q];
  for (int lf = 0; lf < nfe; lf++)
    for (int ori = 0; ori < nori; ori++) {
      const int code = lf * nori + ori;
      for (int q = 0; q < nqf; q++) {
        const double st[2] = {xf[q % nqf1], dim == 3 ? xf[q / nqf1] : 0.0};
        double xi[3], dl[6];
        g_loc_map(dim, lf, ori, st, xi, dl);
        for (int d = 0; d < dim; d++) xiF[(static_cast<size_t>(code) * nqf + q) * dim + d] = xi[d];
        for (int k = 0; k < dim * (dim - 1); k++) dlocF[static_cast<size_t>(code) * dim * (dim - 1) + k] = dl[k];
        g_shape(dim, np, xn, xi, &phiF[(static_cast<size_t>(code) * nqf + q) * dof], nullptr);
      }
    }
  // element -> faces by local face index
  std::vector<int> el_face(static_cast<size_t>(NE) * nfe, -1);
  for (int f = 0; f < NF; f++) {
    const int e1 = maps->face_el1[f], e2 = maps->face_el2[f];
    const int lf1 = maps->face_inf1[f] / 64;
    if (e1 < 0 || e1 >= NE || lf1 < 0 || lf1 >= nfe || maps->face_inf1[f] % 64 != 0) return "invalid face tables";
    el_face[static_cast<size_t>(e1) * nfe + lf1] = f;
    if (e2 >= 0) {
      const int lf2 = maps->face_inf2[f] / 64, ori = maps->face_inf2[f] % 64;
      if (e2 >= NE + c->NEH) return "invalid face tables";
      if (lf2 < 0 || lf2 >= nfe || ori < 0 || ori >= nori) return "invalid face tables";
      if (e2 < NE) el_face[static_cast<size_t>(e2) * nfe + lf2] = f;  // shared face: Elem2 is a face-neighbour element
    }
  }
  for (int v : el_face)
    if (v < 0) return "an element face is missing from the face tables";
  // mass matrices (rhs_operator.cpp:173-189): diagonal when nodes and volume rule coincide
  const bool diag = space->basis_type == 0 && space->int_rule_type == 0;
  std::vector<double> me(static_cast<size_t>(NE) * (diag ? dof : dof * dof));
  // axisymmetric: the inverse of M_ij = int r phi_i phi_j as well (Me_inv_rad, rhs_operator.cpp:191-205)
  std::vector<double> me_rad(axisym ? me.size() : 0);
  std::vector<double> M(static_cast<size_t>(dof) * dof), MR(static_cast<size_t>(dof) * dof);
  for (int e = 0; e < NE; e++) {
    const double *v = &maps->elem_vertices[static_cast<size_t>(e) * nv * dim];
    std::fill(M.begin(), M.end(), 0.0);
    std::fill(MR.begin(), MR.end(), 0.0);
    for (int q = 0; q < nqv; q++) {
      double J[9];
      const double *xi = &xiV[static_cast<size_t>(q) * dim];
      double radius = 1.0;
      if (axisym)  // x coordinate of the quadrature point (bilinear map, MFEM quad vertex order)
        radius = (1 - xi[0]) * (1 - xi[1]) * v[0] + xi[0] * (1 - xi[1]) * v[2] + xi[0] * xi[1] * v[4] + (1 - xi[0]) * xi[1] * v[6];
      if (dim == 2) {
        for (int i = 0; i < 2; i++) {
          J[i] = (1 - xi[1]) * (v[2 + i] - v[i]) + xi[1] * (v[4 + i] - v[6 + i]);
          J[i + 2] = (1 - xi[0]) * (v[6 + i] - v[i]) + xi[0] * (v[4 + i] - v[2 + i]);
        }
      } else {
        const double x = xi[0], y = xi[1], z = xi[2], x0 = 1 - x, y0 = 1 - y, z0 = 1 - z;
        for (int i = 0; i < 3; i++) {
          const double X0 = v[i], X1 = v[3 + i], X2 = v[6 + i], X3 = v[9 + i], X4 = v[12 + i], X5 = v[15 + i], X6 = v[18 + i], X7 = v[21 + i];
          J[i] = y0 * z0 * (X1 - X0) + y * z0 * (X2 - X3) + y0 * z * (X5 - X4) + y * z * (X6 - X7);
          J[i + 3] = x0 * z0 * (X3 - X0) + x * z0 * (X2 - X1) + x0 * z * (X7 - X4) + x * z * (X6 - X5);
          J[i + 6] = x0 * y0 * (X4 - X0) + x * y0 * (X5 - X1) + x * y * (X6 - X2) + x0 * y * (X7 - X3);
        }
      }
      const double det = dim == 2 ? J[0] * J[3] - J[2] * J[1]
                                  : J[0] * (J[4] * J[8] - J[5] * J[7]) - J[3] * (J[1] * J[8] - J[2] * J[7]) + J[6] * (J[1] * J[5] - J[2] * J[4]);
      if (!(det > 0)) return "element with non-positive Jacobian";
      const double wd = wV[q] * det;
      const double *ph = &phiV[static_cast<size_t>(q) * dof];
      if (diag) {
        for (int i = 0; i < dof; i++) M[static_cast<size_t>(i) * dof + i] += wd * ph[i] * ph[i];
        if (axisym)
          for (int i = 0; i < dof; i++) MR[static_cast<size_t>(i) * dof + i] += wd * radius * ph[i] * ph[i];
      } else {
        for (int i = 0; i < dof; i++)
          for (int j = 0; j < dof; j++) M[static_cast<size_t>(i) * dof + j] += wd * ph[i] * ph[j];
        if (axisym)
          for (int i = 0; i < dof; i++)
            for (int j = 0; j < dof; j++) MR[static_cast<size_t>(i) * dof + j] += wd * radius * ph[i] * ph[j];
      }
    }
    if (diag) {
      for (int i = 0; i < dof; i++) me[static_cast<size_t>(e) * dof + i] = 1.0 / M[static_cast<size_t>(i) * dof + i];
      if (axisym)
        for (int i = 0; i < dof; i++) me_rad[static_cast<size_t>(e) * dof + i] = 1.0 / MR[static_cast<size_t>(i) * dof + i];
    } else {
      if (!g_spd_inverse(M, dof)) return "mass matrix is not positive definite";
      std::copy(M.begin(), M.end(), &me[static_cast<size_t>(e) * dof * dof]);
      if (axisym) {
        if (!g_spd_inverse(MR, dof)) return "radius-weighted mass matrix is not positive definite (element on r <= 0?)";
        std::copy(MR.begin(), MR.end(), &me_rad[static_cast<size_t>(e) * dof * dof]);
      }
    }
  }
  // boundary faces -> index into the boundary-condition table (BCintegrator's attribute maps, BCintegrator.cpp:64-125)
  std::vector<int> f_bc(NF, -1);
  GenBcTable gbt;
  memset(&gbt, 0, sizeof(gbt));
  if (bcs && bcs->num_bcs > 0) {
    gbt.nbc = bcs->num_bcs, gbt.use_bc_in_grad = bcs->use_bc_in_grad ? 1 : 0;
    for (int i = 0; i < bcs->num_bcs; i++) {
      gbt.bc[i].kind = bcs->bcs[i].kind, gbt.bc[i].type = bcs->bcs[i].type;
      for (int k = 0; k < TPSB_BC_NDATA; k++) gbt.bc[i].d[k] = bcs->bcs[i].data[k];
    }
    if (maps->face_attr)
      for (int f = 0; f < NF; f++)
        if (maps->face_el2[f] < 0)
          for (int i = 0; i < bcs->num_bcs; i++)
            if (bcs->bcs[i].attr == maps->face_attr[f]) {
              f_bc[f] = i;
              break;
            }
  }
  // non-reflecting inlets / outlets: per patch the face list, the first boundary-state point of every face, tangent1
  std::vector<int> f_nr_off(NF, -1);
  std::vector<std::vector<int>> nr_faces;
  std::vector<GenNrPatch> nr_host;
  for (int i = 0; i < gbt.nbc; i++) {
    GenBc &b = gbt.bc[i];
    b.nr = -1;
    const bool is_nr = (b.kind == 0 && (b.type == 6 || b.type == 7)) || (b.kind == 1 && b.type >= 2 && b.type <= 4);
    if (!is_nr) continue;
    b.nr = static_cast<int>(nr_host.size());
    std::vector<int> faces;
    for (int f = 0; f < NF; f++)
      if (f_bc[f] == i) {
        f_nr_off[f] = static_cast<int>(faces.size()) * nqf;
        faces.push_back(f);
      }
    GenNrPatch pt;
    memset(&pt, 0, sizeof(pt));
    pt.nfaces = static_cast<int>(faces.size());
    pt.ref_length = b.d[8];
    for (int d = 0; d < 3; d++) pt.tangent1[d] = b.d[9 + d];
    if (pt.tangent1[0] == 0.0 && pt.tangent1[1] == 0.0 && pt.tangent1[2] == 0.0 && !faces.empty()) {
      // OutletBC / InletBC constructor (src/outletBC.cpp:161-176): unit vector from the first to the second quadrature
      // point of the patch's first face
      const int f = faces[0], e1 = maps->face_el1[f], code = (maps->face_inf1[f] / 64) * nori;
      const double *v = &maps->elem_vertices[static_cast<size_t>(e1) * nv * dim];
      double X[2][3] = {{0, 0, 0}, {0, 0, 0}};
      for (int q = 0; q < 2; q++) {
        const double *xi = &xiF[(static_cast<size_t>(code) * nqf + q) * dim];
        for (int vtx = 0; vtx < nv; vtx++) {  // multilinear map, MFEM vertex order
          static const int hv[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
          double w = 1.0;
          for (int d = 0; d < dim; d++) w *= hv[vtx][d] ? xi[d] : 1.0 - xi[d];
          for (int d = 0; d < dim; d++) X[q][d] += w * v[vtx * dim + d];
        }
      }
      double m = 0.0;
      for (int d = 0; d < dim; d++) m += (X[1][d] - X[0][d]) * (X[1][d] - X[0][d]);
      for (int d = 0; d < dim; d++) pt.tangent1[d] = (X[1][d] - X[0][d]) * (1. / std::sqrt(m));
    }
    nr_faces.push_back(faces);
    nr_host.push_back(pt);
  }
  GenArgs &g = c->gen;
  memset(&g, 0, sizeof(g));
  g.dim = dim, g.np = np, g.dof = dof, g.nqv = nqv, g.nqf = nqf, g.nfe = nfe, g.nv = nv;
  g.neq = space->num_equation, g.nvel = space->nvel, g.NE = NE, g.N = static_cast<long long>(NE) * dof, g.me_diag = diag ? 1 : 0;
  g.phys.dim = dim, g.phys.nvel = space->nvel, g.phys.neq = space->num_equation, g.phys.dry = c->phys;
  g.phys.axisym = axisym ? 1 : 0;
  g.phys.use_roe = phys->use_roe ? 1 : 0;
  g.bct = gbt;
  g.eq_system = phys->eq_system;
  g.phys.fluid = 0, g.phys.mix = nullptr;
  g.phys.ml_on = phys->use_mixing_length ? 1 : 0;
  g.phys.ml_max = phys->max_mixing_length, g.phys.ml_prt = phys->mixing_length_Prt, g.phys.ml_bulk = phys->mixing_length_bulk_mult;
  std::vector<MixParams> mixv;
  if (phys->fluid == TPSB_USER_DEFINED) {
    const tpsb_plasma_models &pm = *phys->plasma;
    MixParams m;
    memset(&m, 0, sizeof(m));
    m.numSpecies = pm.num_species, m.ambipolar = pm.ambipolar ? 1 : 0, m.twoTemp = pm.two_temperature ? 1 : 0;
    m.numActive = m.ambipolar ? m.numSpecies - 2 : m.numSpecies - 1;                    // equation_of_state.hpp:131
    m.iBackground = m.numSpecies - 1, m.iElectron = m.numSpecies - 2;                   // :137-146
    m.dim = dim, m.nvel = space->nvel, m.neq = space->num_equation, m.iTh = space->nvel + 1, m.iTe = space->num_equation - 1;
    m.eq_system = phys->eq_system;
    m.axisym = axisym ? 1 : 0;
    for (int sp = 0; sp < m.numSpecies; sp++) {
      m.mw[sp] = pm.mw[sp], m.charge[sp] = pm.charge[sp], m.formE[sp] = pm.formation_energy[sp];
      m.molarCV[sp] = pm.molar_cv[sp] * MIX_RU;                                         // equation_of_state.cpp:568-571
      m.molarCP[sp] = m.molarCV[sp] + MIX_RU;
      m.diff[sp] = pm.diffusivity[sp], m.mtFreq[sp] = pm.mt_freq[sp];
    }
    m.visc = pm.viscosity, m.bulk = pm.bulk_viscosity, m.kh = pm.thermal_conductivity, m.ke = pm.electron_thermal_conductivity;
    m.trElectron = m.iElectron, m.chElectron = m.iElectron;
    m.transportModel = pm.transport_model;
    if (pm.transport_model == 0 || pm.transport_model == 1) {  // Gas{Minimal,Mixture}Transport ctor (gas_transport.cpp:42-157, 877-985)
      const int ns = m.numSpecies;
      m.gmIon = pm.transport_model == 1 ? pm.ion_index : 0, m.gmElectron = m.iElectron;
      m.gmNeutral = pm.transport_model == 1 ? pm.neutral_index : m.iBackground;
      m.thirdOrderKe = pm.third_order_k_electron ? 1 : 0, m.multiply = pm.multiply ? 1 : 0;
      for (int sp = 0; sp < ns; sp++) m.gmMw[sp] = pm.mw[sp] / MIX_NA;
      for (int i = 0; i < ns; i++)
        for (int j = i; j < ns; j++) {
          m.gmMuw[i + j * ns] = m.gmMw[i] * m.gmMw[j] / (m.gmMw[i] + m.gmMw[j]);
          if (i != j) m.gmMuw[j + i * ns] = m.gmMuw[i + j * ns];
          m.collIdx[i + j * ns] = pm.collision_index[i + j * ns];
        }
      for (int t = 0; t < 4; t++) m.fluxMult[t] = pm.flux_trns_multiplier[t];
      m.mfFreqMult = pm.mf_freq_multiplier, m.diffMult = pm.diff_mult, m.mobilMult = pm.mobil_mult;
    }
    m.numReactions = pm.num_reactions, m.minTemp = pm.min_temperature;
    for (int r = 0; r < pm.num_reactions; r++) {
      m.rxModel[r] = pm.model[r], m.detailed[r] = pm.detailed_balance[r] ? 1 : 0;
      m.rxA[r] = pm.rate_params[r][0], m.rxB[r] = pm.rate_params[r][1], m.rxE[r] = pm.rate_params[r][2];
      m.rxEnergy[r] = pm.reaction_energy[r];
      m.eqA[r] = pm.equilibrium_params[r][0], m.eqB[r] = pm.equilibrium_params[r][1], m.eqE[r] = pm.equilibrium_params[r][2];
      for (int sp = 0; sp < m.numSpecies; sp++) {
        m.reactS[sp + r * m.numSpecies] = pm.reactant_stoich[r][sp];
        m.prodS[sp + r * m.numSpecies] = pm.product_stoich[r][sp];
      }
    }
    // LinearTable constructor (table.cpp:76-85): slopes / intercepts in the (log) variables
    std::vector<double> tbl;
    m.radiation = pm.nec_table_n > 0 ? 1 : 0;
    for (int r = 0; r <= pm.num_reactions; r++) {
      const bool nec = r == pm.num_reactions;  // last pass: the net-emission table into slot MIX_MAXRX
      if (nec && !m.radiation) break;
      if (!nec) m.rxComp[r] = pm.rate_component[r];
      if (!nec && pm.model[r] != 2) continue;
      const int slot = nec ? MIX_MAXRX : r;
      const int n = nec ? pm.nec_table_n : pm.table_n[r];
      const bool xl = (nec ? pm.nec_table_xlog : pm.table_xlog[r]) != 0, fl = (nec ? pm.nec_table_flog : pm.table_flog[r]) != 0;
      m.tblOff[slot] = static_cast<int>(tbl.size()), m.tblN[slot] = n, m.tblXlog[slot] = xl, m.tblFlog[slot] = fl;
      const double *xd = nec ? pm.nec_table_x : pm.table_x[r], *fd = nec ? pm.nec_table_f : pm.table_f[r];
      tbl.insert(tbl.end(), xd, xd + n);
      std::vector<double> ta(n, 0.0), tb(n, 0.0);
      for (int k = 0; k < n - 1; k++) {
        ta[k] = fl ? log(fd[k]) : fd[k];
        const double df = fl ? (log(fd[k + 1]) - log(fd[k])) : (fd[k + 1] - fd[k]);
        tb[k] = xl ? df / (log(xd[k + 1]) - log(xd[k])) : df / (xd[k + 1] - xd[k]);
        ta[k] -= xl ? tb[k] * log(xd[k]) : tb[k] * xd[k];
      }
      tbl.insert(tbl.end(), ta.begin(), ta.end());
      tbl.insert(tbl.end(), tb.begin(), tb.end());
    }
    if (!tbl.empty()) {
      cudaError_t te = cudaSetDevice(c->device);
      if (te == cudaSuccess) te = g_upload(c, &m.tbl, tbl);
      if (te != cudaSuccess) return std::string("device setup failed: ") + cudaGetErrorString(te);
    }
    m.rateField = nullptr, m.rateN = static_cast<long long>(NE) * dof;
    m.mlOn = phys->use_mixing_length ? 1 : 0;
    m.mlMax = phys->max_mixing_length, m.mlPrt = phys->mixing_length_Prt, m.mlBulk = phys->mixing_length_bulk_mult;
    mixv.push_back(m);
    c->mix_host = m;
    g.phys.fluid = 1;
  }
  std::vector<double> vx(maps->elem_vertices, maps->elem_vertices + static_cast<size_t>(NE) * nv * dim);
  std::vector<int> fe1(maps->face_el1, maps->face_el1 + NF), fe2(maps->face_el2, maps->face_el2 + NF),
      fi1(maps->face_inf1, maps->face_inf1 + NF), fi2(maps->face_inf2, maps->face_inf2 + NF);
  cudaError_t ce = cudaSetDevice(c->device);
  if (ce == cudaSuccess) ce = g_upload(c, &g.phiV, phiV);
  if (ce == cudaSuccess) ce = g_upload(c, &g.dphiV, dphiV);
  if (ce == cudaSuccess) ce = g_upload(c, &g.wV, wV);
  if (ce == cudaSuccess) ce = g_upload(c, &g.xiV, xiV);
  if (ce == cudaSuccess) ce = g_upload(c, &g.phiF, phiF);
  if (ce == cudaSuccess) ce = g_upload(c, &g.xiF, xiF);
  if (ce == cudaSuccess) ce = g_upload(c, &g.dlocF, dlocF);
  if (ce == cudaSuccess) ce = g_upload(c, &g.wF, wF);
  if (ce == cudaSuccess) ce = g_upload(c, &g.vx, vx);
  if (ce == cudaSuccess) ce = g_upload(c, &g.el_face, el_face);
  if (ce == cudaSuccess) ce = g_upload(c, &g.f_el1, fe1);
  if (ce == cudaSuccess) ce = g_upload(c, &g.f_el2, fe2);
  if (ce == cudaSuccess) ce = g_upload(c, &g.f_inf1, fi1);
  if (ce == cudaSuccess) ce = g_upload(c, &g.f_inf2, fi2);
  if (ce == cudaSuccess) ce = g_upload(c, &g.me_inv, me);
  if (ce == cudaSuccess && axisym) ce = g_upload(c, &g.me_inv_rad, me_rad);
  if (ce == cudaSuccess) ce = g_upload(c, &g.xiN, xiN);
  if (ce == cudaSuccess) ce = g_upload(c, &g.f_bc, f_bc);
  if (!nr_host.empty()) {
    const int neq = space->num_equation;
    if (ce == cudaSuccess) ce = g_upload(c, &g.f_nr_off, f_nr_off);
    for (size_t i = 0; i < nr_host.size() && ce == cudaSuccess; i++) {
      GenNrPatch &pt = nr_host[i];
      const size_t npts = static_cast<size_t>(pt.nfaces) * nqf;
      double *block = nullptr;  // boundaryU | meanUp | sums | area, one allocation per patch
      const size_t nd = npts * neq + neq + (neq + 2) + 1;
      if (!nr_faces[i].empty()) ce = g_upload(c, &pt.faces, nr_faces[i]);
      if (ce == cudaSuccess) ce = cudaMalloc(&block, nd * sizeof(double));
      if (ce == cudaSuccess) ce = cudaMemset(block, 0, nd * sizeof(double));
      if (ce == cudaSuccess) ce = cudaMalloc(&pt.init, sizeof(int));
      if (ce == cudaSuccess) ce = cudaMemset(pt.init, 0, sizeof(int));
      if (ce != cudaSuccess) break;
      c->gen_allocs.push_back(block);
      c->gen_allocs.push_back(pt.init);
      pt.boundaryU = block, pt.meanUp = block + npts * neq, pt.sums = pt.meanUp + neq, pt.area = pt.sums + neq + 2;
    }
    if (ce == cudaSuccess) ce = g_upload(c, &g.nr, nr_host);
    c->nr_patches = nr_host;
    for (int i = 0; i < gbt.nbc; i++)
      if (gbt.bc[i].nr >= 0) c->nr_attr.push_back(bcs->bcs[i].attr);
  }
  if (ce == cudaSuccess && !mixv.empty()) ce = g_upload(c, &g.phys.mix, mixv);
  if (ce == cudaSuccess && phys->fluid == TPSB_LTE_FLUID) {
    // LinearTable's constructor (table.cpp:76-87) on the host: per table x[n] | a[n-1] | b[n-1], one device pool
    const tpsb_lte_tables &lt = *phys->lte;
    struct Src { int n; const double *x, *f; int xlog, flog; };
    const Src src[LTE_NTAB] = {{lt.num_thermo, lt.T, lt.energy, 0, 0}, {lt.num_thermo, lt.T, lt.R, 0, 0},
                               {lt.num_thermo, lt.T, lt.c, 0, 0},      {lt.num_thermo, lt.energy, lt.T, 0, 0},
                               {lt.num_trans, lt.T_trans, lt.mu, 0, 0}, {lt.num_trans, lt.T_trans, lt.kappa, 0, 0},
                               {lt.nec_table_n, lt.nec_table_x, lt.nec_table_f, lt.nec_table_xlog, lt.nec_table_flog}};
    std::vector<double> pool;
    size_t off[LTE_NTAB];
    LteParams L;
    memset(&L, 0, sizeof(L));
    for (int t = 0; t < LTE_NTAB; t++) {
      off[t] = pool.size();
      const int n = src[t].n;
      L.n[t] = n, L.xlog[t] = src[t].xlog ? 1 : 0, L.flog[t] = src[t].flog ? 1 : 0;
      if (n < 2) continue;
      pool.insert(pool.end(), src[t].x, src[t].x + n);
      std::vector<double> a(n - 1), b(n - 1);
      for (int k = 0; k < n - 1; k++) {
        const double *x = src[t].x, *f = src[t].f;
        a[k] = L.flog[t] ? std::log(f[k]) : f[k];
        const double df = L.flog[t] ? (std::log(f[k + 1]) - std::log(f[k])) : (f[k + 1] - f[k]);
        b[k] = L.xlog[t] ? df / (std::log(x[k + 1]) - std::log(x[k])) : df / (x[k + 1] - x[k]);
        a[k] -= L.xlog[t] ? b[k] * std::log(x[k]) : b[k] * x[k];
      }
      pool.insert(pool.end(), a.begin(), a.end());
      pool.insert(pool.end(), b.begin(), b.end());
    }
    const double *d_pool = nullptr;
    ce = g_upload(c, &d_pool, pool);
    for (int t = 0; t < LTE_NTAB; t++) L.x[t] = d_pool + off[t];
    c->lte_radiation = lt.nec_table_n >= 2;
    std::vector<LteParams> lv(1, L);
    if (ce == cudaSuccess) ce = g_upload(c, &g.phys.lte, lv);
  }
  const size_t nb = static_cast<size_t>(g.N) * sizeof(double);
  if (ce == cudaSuccess) ce = cudaMalloc(&c->d_Up, nb * g.neq);
  if (ce == cudaSuccess) ce = cudaMalloc(&c->d_gradUp, nb * g.neq * dim);
  if (ce == cudaSuccess) ce = cudaMalloc(&c->d_maxBits, sizeof(unsigned long long));
  if (ce == cudaSuccess) ce = cudaMalloc(&c->d_mcs, sizeof(double));
  if (ce == cudaSuccess) ce = cudaMemset(c->d_maxBits, 0, sizeof(unsigned long long));
  if (ce != cudaSuccess) return std::string("device setup failed: ") + cudaGetErrorString(ce);
  g.Up = c->d_Up, g.gradUp = c->d_gradUp, g.maxCharBits = c->d_maxBits;
  g.NEH = c->NEH;
  g.Uhalo = g.UpHalo = g.gradUpHalo = g.distHalo = nullptr;
  g.elem_delta = nullptr;
  if (c->phys.sgs_model | c->phys.sponge) {
    // delta = Mesh::GetElementSize(e, 1) / order (rhs_operator.cpp:149-156), local then face-neighbour elements
    std::vector<double> delta(static_cast<size_t>(NE + c->NEH));
    for (int e = 0; e < NE + c->NEH; e++) {
      const double *v = &maps->elem_vertices[static_cast<size_t>(e) * nv * dim];
      delta[e] = (dim == 3 ? hex_min_size(v) : quad_min_size(v)) / p;
    }
    ce = g_upload(c, &g.elem_delta, delta);
    if (ce != cudaSuccess) return std::string("device setup failed: ") + cudaGetErrorString(ce);
  }
  if (c->NEH > 0) {  // partitioned mesh: face-neighbour copies and the exchange description
    const std::string herr = setup_halo_desc(c, halo, c->NEH);
    if (!herr.empty()) return herr;
    const size_t hb = static_cast<size_t>(c->NEH) * dof * sizeof(double), sb = static_cast<size_t>(c->n_send) * dof * sizeof(double);
    ce = cudaMalloc(&c->d_Uhalo, hb * g.neq);
    if (ce == cudaSuccess) ce = cudaMalloc(&c->d_UpHalo, hb * g.neq);
    if (ce == cudaSuccess) ce = cudaMalloc(&c->d_gradUpHalo, hb * g.neq * dim);
    if (ce == cudaSuccess) ce = cudaMalloc(&c->d_sendU, sb * g.neq);
    if (ce == cudaSuccess) ce = cudaMalloc(&c->d_sendG, sb * g.neq * dim);
    if (ce != cudaSuccess) return std::string("device setup failed: ") + cudaGetErrorString(ce);
    g.Uhalo = c->d_Uhalo, g.UpHalo = c->d_UpHalo, g.gradUpHalo = c->d_gradUpHalo;
  }
  // element_to_faces (stride 1 + faces per element; 7 in 3-D as in the reference)
  std::vector<int> e2f(static_cast<size_t>(nfe + 1) * NE, 0);
  for (int f = 0; f < NF; f++) {
    if (maps->face_el2[f] < 0) continue;
    for (int e : {maps->face_el1[f], maps->face_el2[f]}) {
      if (e >= NE) continue;  // shared face: the second element lives on another rank
      const int nf = e2f[static_cast<size_t>(nfe + 1) * e];
      e2f[static_cast<size_t>(nfe + 1) * e + nf + 1] = f;
      e2f[static_cast<size_t>(nfe + 1) * e] = nf + 1;
    }
  }
  c->element_to_faces = e2f;
  return "";
}


// Static schedule of the chunked host-buffer pipeline.  Elements are cut into C contiguous chunks, the two-sided
// faces into the C ranges "Elem1 lies in chunk k" (contiguous because faces are numbered by first appearance over
// the elements).  A chunk's gradient needs the primitives of its face neighbours' chunks, a face range the trace
// blocks of both sides' chunks, a chunk's residual the face residuals of all its faces; the op list below is the
// greedy order in which those become available while the chunks arrive 0, 1, 2, ...
// bdr_el1: elements of the boundary faces (face-residual slots NFint + k), in slot order; a boundary face belongs to the
// face range of its element's chunk and needs that chunk's gradient only.
bool host_pipe_schedule(int NE, int NFint, int C, const std::vector<int> &nbr_elem, const std::vector<int> &el_face,
                        const std::vector<int> &fl_el1, const std::vector<int> &fl_el2, const std::vector<int> &bdr_el1,
                        std::vector<int> &eb, std::vector<int> &fb, std::vector<int> &bb, std::vector<tpsb_ctx::PipeOp> &ops) {
  if (C < 3 || C > 64 || C > NE) return false;
  const int NB = static_cast<int>(bdr_el1.size());
  eb.assign(C + 1, 0);
  fb.assign(C + 1, NFint);
  bb.assign(C + 1, NB);
  for (int k = 0; k <= C; k++) eb[k] = static_cast<int>(static_cast<long long>(NE) * k / C);
  auto chunk_of = [&](int e) { return static_cast<int>(std::upper_bound(eb.begin(), eb.end(), e) - eb.begin()) - 1; };
  fb[0] = 0;
  int cur = 0;
  for (int f = 0; f < NFint; f++) {
    const int cf = chunk_of(fl_el1[f]);
    if (cf < cur) return false;  // face order not monotone in Elem1: keep the unchunked path
    while (cur < cf) fb[++cur] = f;
  }
  while (cur < C) fb[++cur] = NFint;
  bb[0] = 0;
  cur = 0;
  for (int k = 0; k < NB; k++) {
    const int cb = chunk_of(bdr_el1[k]);
    if (cb < cur) return false;
    while (cur < cb) bb[++cur] = k;
  }
  while (cur < C) bb[++cur] = NB;
  using mask = unsigned long long;
  std::vector<mask> need_prim(C, 0), need_grad(C, 0), need_face(C, 0);
  for (int e = 0; e < NE; e++) {
    const int ce = chunk_of(e);
    need_prim[ce] |= mask(1) << ce;
    for (int lf = 0; lf < 6; lf++) {
      const int nb = nbr_elem[static_cast<size_t>(e) * 6 + lf], fc = el_face[static_cast<size_t>(e) * 6 + lf];
      if (nb >= 0) need_prim[ce] |= mask(1) << chunk_of(nb);
      if (fc >= 0 && fc < NFint)
        need_face[ce] |= mask(1) << (static_cast<int>(std::upper_bound(fb.begin(), fb.end(), fc) - fb.begin()) - 1);
      else if (fc >= NFint)
        need_face[ce] |= mask(1) << ce;  // its own boundary faces
    }
  }
  for (int f = 0; f < NFint; f++) {
    const int cf = chunk_of(fl_el1[f]);
    need_grad[cf] |= (mask(1) << cf) | (mask(1) << chunk_of(fl_el2[f]));
  }
  for (int k = 0; k < NB; k++) {  // a chunk's boundary faces ride in its face op and read its elements' gradients
    const int cb = chunk_of(bdr_el1[k]);
    need_grad[cb] |= mask(1) << cb;
  }
  ops.clear();
  mask primd = 0, gradd = 0, faced = 0, resd = 0;
  for (int k = 0; k < C; k++) {
    ops.push_back({0, k});
    primd |= mask(1) << k;
    for (bool progress = true; progress;) {
      progress = false;
      for (int g = 0; g < C; g++)
        if (!(gradd >> g & 1) && (need_prim[g] & ~primd) == 0) ops.push_back({1, g}), gradd |= mask(1) << g, progress = true;
      for (int f = 0; f < C; f++)
        if (!(faced >> f & 1) && (need_grad[f] & ~gradd) == 0) {
          if (fb[f + 1] > fb[f] || bb[f + 1] > bb[f]) ops.push_back({2, f});
          faced |= mask(1) << f, progress = true;
        }
      for (int r = 0; r < C; r++)
        if (!(resd >> r & 1) && (gradd >> r & 1) && (need_face[r] & ~faced) == 0)
          ops.push_back({3, r}), resd |= mask(1) << r, progress = true;
    }
  }
  return resd == (C == 64 ? ~mask(0) : (mask(1) << C) - 1);
}

void build_host_pipe(tpsb_ctx *c, const std::vector<int> &nbr_elem, const std::vector<int> &el_face,
                     const std::vector<int> &fl_el1, const std::vector<int> &fl_el2, const std::vector<int> &bdr_el1) {
  int C = std::min(64, c->NE / 2048);
  if (const char *ev = getenv("TPSB_HOST_CHUNKS")) C = atoi(ev);
  C = std::min(C, 64);
  std::vector<int> eb, fb, bb;
  std::vector<tpsb_ctx::PipeOp> ops;
  if (!host_pipe_schedule(c->NE, c->NFint, C, nbr_elem, el_face, fl_el1, fl_el2, bdr_el1, eb, fb, bb, ops)) return;
  c->pipe_chunks = C;
  c->pipe_eb = eb;
  c->pipe_fb = fb;
  c->pipe_bb = bb;
  c->pipe_ops = ops;
}

}  // namespace

// Face-neighbour exchange description shared by every path: peers, send / receive offsets, the send-element list on the
// device, the highest-priority exchange stream and its events.  Returns an error string (empty on success).
static std::string setup_halo_desc(tpsb_ctx *c, const tpsb_halo_desc *halo, int NEH) {
  if (!halo || halo->num_nbr_ranks <= 0 || !halo->nccl_comm || !halo->nbr_rank || !halo->send_offset || !halo->send_elems ||
      !halo->recv_offset)
    return "mesh has " + std::to_string(NEH) + " face-neighbour elements but no halo description";
  const int np = halo->num_nbr_ranks;
  c->comm = static_cast<ncclComm_t>(halo->nccl_comm);
  c->nbr_rank.assign(halo->nbr_rank, halo->nbr_rank + np);
  c->send_offset.assign(halo->send_offset, halo->send_offset + np + 1);
  c->recv_offset.assign(halo->recv_offset, halo->recv_offset + np + 1);
  c->n_send = c->send_offset[np];
  if (c->recv_offset[np] != NEH) return "recv_offset does not cover the " + std::to_string(NEH) + " face-neighbour elements";
  for (int k = 0; k < c->n_send; k++)
    if (halo->send_elems[k] < 0 || halo->send_elems[k] >= c->NE) return "send_elems out of range";
  std::vector<int> se(halo->send_elems, halo->send_elems + c->n_send);
  cudaError_t ce = upload(&c->d_send_elems, se);
  if (ce == cudaSuccess) {
    // highest priority: the exchange kernels must get SM slots while the interior gradient / face kernels (hundreds of
    // thousands of queued CTAs) run, or the "overlapped" exchange only starts when they drain
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    static const bool flat = getenv("TPSB_COMM_PRIO") && atoi(getenv("TPSB_COMM_PRIO")) == 0;
    ce = cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, flat ? prio_lo : prio_hi);
  }
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&c->ev_pack, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&c->ev_recvU, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&c->ev_recvG, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&c->ev_recvT, cudaEventDisableTiming);
  if (ce != cudaSuccess) return std::string("halo setup failed: ") + cudaGetErrorString(ce);
  return "";
}

extern "C" {

const char *tpsb_version(void) { return "tpsb200 0.1 (sm_100a, fp64)"; }

const char *tpsb_last_error(const tpsb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int tpsb_create(const tpsb_mesh_maps *maps, const tpsb_space_desc *space, const tpsb_physics *phys,
                const tpsb_bc_set *bcs, const tpsb_halo_desc *halo, int device, void *cuda_stream, tpsb_ctx **out) {
  tpsb_ctx *ctx = nullptr;  // errors before allocation go to the thread-local create error
  if (!maps || !space || !phys || !out) return fail(ctx, TPSB_EINVAL, "null argument");
  *out = nullptr;
  if (maps->dim != 2 && maps->dim != 3) return fail(ctx, TPSB_EINVAL, "dim must be 2 (quadrilaterals) or 3 (hexahedra); got %d", maps->dim);
  if (space->basis_type < 0 || space->basis_type > 1 || space->int_rule_type < 0 || space->int_rule_type > 1)
    return fail(ctx, TPSB_EINVAL, "basisType / integrationRule must be 0 (Gauss-Legendre) or 1 (Gauss-Lobatto)");
  // orders 1-3 have sum-factorised kernels; 4 and 5 run on the generic path (dense reference-element tables of any size:
  // the largest rule, the 3-D Gauss-Lobatto face rule at p = 5, has 8 points per direction)
  if (space->order < 1 || space->order > 5) return fail(ctx, TPSB_ENOTIMPL, "order must be 1..5");
  // config.isAxisymmetric(): a 2-D (r, z) mesh carrying three velocity components
  if (space->nvel != maps->dim && !(maps->dim == 2 && space->nvel == 3))
    return fail(ctx, TPSB_EINVAL, "nvel must equal dim (or 3 on a 2-D mesh for axisymmetric runs)");
  if (phys->fluid == TPSB_DRY_AIR) {
    if (space->num_equation != space->nvel + 2) return fail(ctx, TPSB_EINVAL, "dry air has num_equation = nvel + 2");
  } else if (phys->fluid == TPSB_USER_DEFINED) {
    const tpsb_plasma_models *pm = phys->plasma;
    if (!pm) return fail(ctx, TPSB_EINVAL, "fluid = user_defined needs tpsb_physics.plasma");
    if (pm->num_species < 3 || pm->num_species > TPSB_MAX_SPECIES) return fail(ctx, TPSB_EINVAL, "num_species must be 3..%d (electron and background included)", TPSB_MAX_SPECIES);
    if (pm->transport_model < 0 || pm->transport_model > 2)
      return fail(ctx, TPSB_ENOTIMPL, "transport_model %d not built (2 constant, 0 argon_minimal, 1 argon_mixture)", pm->transport_model);
    if (pm->transport_model == 1) {
      if (pm->num_species > 7) return fail(ctx, TPSB_EINVAL, "argon_mixture transport serves at most 7 species (gas_transport.cpp:903)");
      if (pm->ion_index < 0 || pm->ion_index >= pm->num_species || pm->neutral_index < 0 || pm->neutral_index >= pm->num_species)
        return fail(ctx, TPSB_EINVAL, "argon_mixture transport needs the indices of 'Ar.+1' and 'Ar'");
      for (int i = 0; i < pm->num_species; i++)
        for (int j = i; j < pm->num_species; j++)
          if (pm->collision_index[i + j * pm->num_species] < 0 || pm->collision_index[i + j * pm->num_species] > 14 ||
              pm->collision_index[i + j * pm->num_species] == 5)
            return fail(ctx, TPSB_EINVAL, "collision type %d of species pair (%d, %d) is not a GasColl value (src/dataStructures.hpp:122-144)",
                        pm->collision_index[i + j * pm->num_species], i, j);
    }
    if (pm->transport_model == 0 && (pm->num_species != 3 || !(pm->charge[0] > 0) || !(pm->charge[1] < 0) || pm->charge[2] != 0))
      return fail(ctx, TPSB_EINVAL, "argon_minimal transport serves the ternary mixture [Ar.+1, E, Ar] only (gas_transport.cpp:51-56)");
    if (pm->num_reactions < 0 || pm->num_reactions > TPSB_MAX_REACTIONS) return fail(ctx, TPSB_EINVAL, "too many reactions");
    for (int r = 0; r < pm->num_reactions; r++)
      if (pm->model[r] < 0 || pm->model[r] > 3)
        return fail(ctx, TPSB_ENOTIMPL, "reaction model %d not built (0 Arrhenius, 1 Hoffert-Lien, 2 tabulated, 3 grid function)", pm->model[r]);
    if (pm->nec_table_n != 0 && (pm->nec_table_n < 2 || pm->nec_table_n > 1000 || !pm->nec_table_x || !pm->nec_table_f))
      return fail(ctx, TPSB_EINVAL, "the net-emission-coefficient table needs 2..1000 points");
    for (int r = 0; r < pm->num_reactions; r++)
      if (pm->model[r] == 2 && (pm->table_n[r] < 2 || pm->table_n[r] > 1000 || !pm->table_x[r] || !pm->table_f[r]))
        return fail(ctx, TPSB_EINVAL, "reaction %d: a tabulated rate needs 2..1000 table points (gpudata::MAXTABLE)", r);
    const int nact = pm->ambipolar ? pm->num_species - 2 : pm->num_species - 1;
    if (space->num_equation != space->nvel + 2 + nact + (pm->two_temperature ? 1 : 0) || space->num_equation > GEN_MAXEQ)
      return fail(ctx, TPSB_EINVAL, "num_equation does not match the mixture (nvel + 2 + active species [+ 1 electron energy])");
  } else if (phys->fluid == TPSB_LTE_FLUID) {
    // LteMixture / LteTransport with 1-D tables (flow/lte/table_dim = 1, src/M2ulPhyS.cpp:175-258); the 2-D (GSL) tables are
    // a CPU-only feature of the reference (src/M2ulPhyS.cpp:168-170)
    const tpsb_lte_tables *lt = phys->lte;
    if (!lt) return fail(ctx, TPSB_EINVAL, "fluid = lte needs tpsb_physics.lte");
    if (lt->num_thermo < 2 || lt->num_thermo > 1000 || !lt->T || !lt->energy || !lt->R || !lt->c)
      return fail(ctx, TPSB_EINVAL, "the LTE thermodynamic table needs 2..1000 rows of T, energy, R, c (gpudata::MAXTABLE)");
    if (lt->num_trans < 2 || lt->num_trans > 1000 || !lt->T_trans || !lt->mu || !lt->kappa)
      return fail(ctx, TPSB_EINVAL, "the LTE transport table needs 2..1000 rows of T, mu, kappa");
    if (lt->nec_table_n != 0 && (lt->nec_table_n < 2 || lt->nec_table_n > 1000 || !lt->nec_table_x || !lt->nec_table_f))
      return fail(ctx, TPSB_EINVAL, "the net-emission-coefficient table needs 2..1000 points");
    for (int k = 1; k < lt->num_thermo; k++)
      if (!(lt->T[k] > lt->T[k - 1]) || !(lt->energy[k] > lt->energy[k - 1]))
        return fail(ctx, TPSB_EINVAL, "LTE tables: T and energy(T) must increase strictly (the e -> T table is the inverse)");
    if (space->num_equation != space->nvel + 2) return fail(ctx, TPSB_EINVAL, "the LTE fluid has num_equation = nvel + 2");
    if (phys->use_mixing_length || phys->sgs_model != 0 || phys->sponge_enabled)
      return fail(ctx, TPSB_ENOTIMPL, "mixing length / SGS / viscous sponge with the LTE fluid are not built");
    for (int i = 0; bcs && bcs->bcs && i < bcs->num_bcs; i++)
      if (bcs->bcs[i].kind == TPSB_BC_WALL && bcs->bcs[i].type == 4)
        return fail(ctx, TPSB_ENOTIMPL, "the general wall (VISC_GNRL) with the LTE fluid is not built");
  } else {
    return fail(ctx, TPSB_ENOTIMPL, "working fluid %d not built", phys->fluid);
  }
  // 3-D Gauss-Legendre dry air runs the specialised kernels; everything else the generic tensor-product path
  bool want_generic = maps->dim != 3 || space->basis_type != 0 || space->int_rule_type != 0 || phys->fluid != TPSB_DRY_AIR ||
                      space->order > 3;
  if (const char *pth = getenv("TPSB_PATH")) want_generic = want_generic || strcmp(pth, "generic") == 0;
  if (phys->use_roe && !(maps->dim == 2 && space->nvel == 2 && phys->fluid == TPSB_DRY_AIR))
    return fail(ctx, TPSB_ENOTIMPL, "useRoe: Eval_Roe of the reference is written for 2-D dry air only (riemann_solver.cpp:117-206)");
  if (phys->use_mixing_length) want_generic = true;  // the mixing-length model lives on the generic path
  auto bc_is_nr = [](const tpsb_bc_desc &b) {  // non-reflecting / mass-flow inlets and outlets: generic path
    return (b.kind == TPSB_BC_INLET && (b.type == 6 || b.type == 7)) || (b.kind == TPSB_BC_OUTLET && b.type >= 2 && b.type <= 4);
  };
  if (bcs && bcs->bcs)
    for (int i = 0; i < bcs->num_bcs && i < MAX_BC; i++)
      if (bc_is_nr(bcs->bcs[i])) want_generic = true;
  for (int i = 0; bcs && bcs->bcs && i < bcs->num_bcs; i++)  // ... and so does the general wall (WallType VISC_GNRL)
    if (bcs->bcs[i].kind == TPSB_BC_WALL && bcs->bcs[i].type == 4) want_generic = true;
  const bool visc_mod = phys->sgs_model != 0 || phys->sponge_enabled != 0;
  if (phys->sgs_model < 0 || phys->sgs_model > 2) return fail(ctx, TPSB_EINVAL, "sgs_model %d: 0 none, 1 smagorinsky, 2 sigma", phys->sgs_model);
  if (phys->sgs_model != 0 && maps->dim != 3)
    return fail(ctx, TPSB_ENOTIMPL, "the SGS models read a 3 x 3 velocity gradient (fluxes.cpp:513-650): 3-D runs only");
  if (phys->sponge_enabled) {
    const double *n = phys->sponge_normal;
    if (!(n[0] * n[0] + n[1] * n[1] + n[2] * n[2] > 0) || !(phys->sponge_width > 0))
      return fail(ctx, TPSB_EINVAL, "viscous sponge needs a non-zero normal and a positive width");
  }
  if (phys->eq_system != TPSB_EULER && phys->eq_system != TPSB_NS)
    return fail(ctx, TPSB_ENOTIMPL, "equation system %d not built", phys->eq_system);
  if (maps->num_elems <= 0 || maps->num_faces <= 0 || !maps->elem_vertices || !maps->face_el1 || !maps->face_el2 ||
      !maps->face_inf1 || !maps->face_inf2)
    return fail(ctx, TPSB_EINVAL, "incomplete mesh maps");
  BcTable bct;
  memset(&bct, 0, sizeof(bct));
  if (bcs) {
    if (bcs->num_bcs < 0 || bcs->num_bcs > MAX_BC || (bcs->num_bcs > 0 && !bcs->bcs))
      return fail(ctx, TPSB_EINVAL, "at most %d boundary conditions are supported", MAX_BC);
    bct.nbc = bcs->num_bcs;
    bct.use_bc_in_grad = bcs->use_bc_in_grad ? 1 : 0;
    for (int i = 0; i < bcs->num_bcs; i++) {
      const tpsb_bc_desc &b = bcs->bcs[i];
      const bool ok = (b.kind == TPSB_BC_INLET && b.type == 2) || (b.kind == TPSB_BC_OUTLET && b.type == 0) ||
                      (b.kind == TPSB_BC_WALL && b.type >= 0 && b.type <= 3) ||
                      (b.kind == TPSB_BC_WALL && b.type == 4 && want_generic) ||  // VISC_GNRL: generic path
                      bc_is_nr(b);
      if (bc_is_nr(b)) {
        // the reference refuses them for mixtures (src/inletBC.cpp:49-52, src/outletBC.cpp:50-53) and indexes the energy
        // with 1 + dim, which is wrong for the three-velocity axisymmetric state
        if (phys->fluid != TPSB_DRY_AIR || space->nvel != maps->dim)
          return fail(ctx, TPSB_ENOTIMPL, "non-reflecting boundary conditions: dry air, planar 2-D or 3-D only (attribute %d)", b.attr);
        if (!(b.data[8] > 0.0)) return fail(ctx, TPSB_EINVAL, "non-reflecting boundary condition on attribute %d needs data[8] = refLength > 0", b.attr);
      }
      if (!ok)
        return fail(ctx, TPSB_ENOTIMPL, "boundary condition kind %d type %d (attribute %d) not built yet", b.kind, b.type, b.attr);
      bct.bc[i].kind = b.kind;
      bct.bc[i].type = b.type;
      for (int k = 0; k < 4; k++) bct.bc[i].d[k] = b.data[k];
      if (b.kind == TPSB_BC_INLET && phys->fluid == TPSB_USER_DEFINED && !want_generic)
        return fail(ctx, TPSB_EINVAL, "mixture inlets run on the generic path");
    }
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(ctx, TPSB_ECUDA, "no CUDA device: libtpsb200 has no CPU fallback");
  if (device < 0 || device >= ndev || device >= MAX_DEV) return fail(ctx, TPSB_EINVAL, "device %d out of range", device);

  tpsb_ctx *c = new tpsb_ctx;
  ctx = c;
  c->device = device;
  c->stream = static_cast<cudaStream_t>(cuda_stream);
  cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device);
  if (const char *fc = getenv("TPSB_FACE_CTAS")) c->face_ctas_per_sm = std::max(1, atoi(fc));
  if (const char *tn = getenv("TPSB_TUNE")) sscanf(tn, "%d,%d,%d", &c->tune[0], &c->tune[1], &c->tune[2]);
  c->order = space->order;
  c->np = space->order + 1;
  c->nd = c->np * c->np * c->np;
  c->NE = maps->num_elems;
  c->NEH = maps->num_nbr_elems;
  c->NF = maps->num_faces;
  c->N = static_cast<long long>(c->NE) * c->nd;
  c->NH = static_cast<long long>(c->NEH) * c->nd;
  if (!build_ref_tables(std::min(c->order, 3), c->T)) {  // orders 4, 5: generic path only, these tables are not used
    delete c;
    return fail(nullptr, TPSB_EINVAL, "reference tables");
  }
  for (int lf = 0; lf < 6; lf++)
    if (c->T.face_par[lf] != kFacePar(lf)) {
      delete c;
      return fail(nullptr, TPSB_EINVAL, "compile-time face parameters disagree with the reference-element tables");
    }
  c->bct = bct;
  c->phys.eq_system = phys->eq_system;
  c->phys.gamma = phys->specific_heat_ratio;
  c->phys.R = phys->gas_constant;
  c->phys.gm1 = phys->specific_heat_ratio - 1.0;
  c->phys.visc_mult = phys->visc_mult;
  c->phys.bulk_visc_mult = phys->bulk_visc_mult;
  c->phys.C1 = phys->sutherland_C1;
  c->phys.S0 = phys->sutherland_S0;
  c->phys.Pr = phys->sutherland_Pr;
  c->phys.sgs_model = phys->sgs_model;
  c->phys.sgs_const = phys->sgs_const;
  c->phys.sgs_floor = phys->sgs_floor;
  c->phys.sponge = phys->sponge_enabled ? 1 : 0;
  c->phys.sp_ratio = 1.0, c->phys.sp_width = 1.0;
  for (int d = 0; d < 3; d++) c->phys.sp_n[d] = c->phys.sp_p[d] = 0.0;
  if (phys->sponge_enabled) {  // Fluxes' constructor normalises the normal (fluxes.cpp:73-83)
    const double *n = phys->sponge_normal;
    const double nm = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    for (int d = 0; d < 3; d++) c->phys.sp_n[d] = n[d] / nm, c->phys.sp_p[d] = phys->sponge_point[d];
    c->phys.sp_ratio = phys->sponge_ratio;
    c->phys.sp_width = phys->sponge_width;
  }
  c->phys.cp_div_pr = phys->specific_heat_ratio * phys->gas_constant /
                      (phys->sutherland_Pr * (phys->specific_heat_ratio - 1.));

  // hmin = min over the local elements of Mesh::GetElementSize(e, 1) (src/M2ulPhyS.cpp:756-761): the adaptive time step
  c->hmin = 1.0e18;
  for (int e = 0; e < c->NE; e++)
    c->hmin = std::min(c->hmin, maps->dim == 3 ? hex_min_size(&maps->elem_vertices[static_cast<size_t>(e) * 24])
                                               : quad_min_size(&maps->elem_vertices[static_cast<size_t>(e) * 8]));

  if (want_generic) {
    c->generic = true;
    c->dim = maps->dim;
    c->neq = space->num_equation;
    c->nd = 1;
    for (int d = 0; d < maps->dim; d++) c->nd *= c->np;
    c->N = static_cast<long long>(c->NE) * c->nd;
    const std::string err = create_generic(c, maps, space, phys, bcs, halo);
    if (!err.empty()) {
      tpsb_destroy(c);
      return fail(nullptr, TPSB_EINVAL, "%s", err.c_str());
    }
    *out = c;
    return TPSB_OK;
  }
  const int NE = c->NE, NEH = c->NEH;
  bool all_affine = true;
  for (int e = 0; e < NE && all_affine; e++) all_affine = element_is_affine(&maps->elem_vertices[static_cast<size_t>(e) * 24]);

  // ---- derive the index maps (M2ulPhyS::initIndirectionArrays, src/M2ulPhyS.cpp:816-1075) ----
  std::vector<int> e2f(static_cast<size_t>(7) * NE, 0);
  std::vector<int> fl_el1, fl_el2, fl_inf1, fl_inf2, sh_el1, sh_el2, sh_inf1, sh_inf2;
  std::vector<int> b_el1, b_lf, b_bc;
  for (int f = 0; f < c->NF; f++) {
    const int e1 = maps->face_el1[f], e2 = maps->face_el2[f];
    if (e1 < 0 || e1 >= NE || e2 >= NE + NEH) {
      delete c;
      return fail(nullptr, TPSB_EINVAL, "face %d has invalid elements (%d,%d)", f, e1, e2);
    }
    if (e2 < 0) {  // boundary face: BCintegrator (src/BCintegrator.cpp:127-197 builds the same per-patch lists)
      const int attr = maps->face_attr ? maps->face_attr[f] : 0;
      int ibc = -1;
      for (int i = 0; bcs && i < bcs->num_bcs; i++)
        if (bcs->bcs[i].attr == attr) ibc = i;
      const int lf = maps->face_inf1[f] / 64;
      if (ibc < 0 || lf < 0 || lf > 5) {
        delete c;
        return fail(nullptr, TPSB_EINVAL, "boundary face %d (attribute %d) has no boundary condition", f, attr);
      }
      b_el1.push_back(e1), b_lf.push_back(lf), b_bc.push_back(ibc);
      continue;
    }
    // element_to_faces: interior (incl. shared) faces in ascending face order
    int nf = e2f[7 * e1];
    if (nf >= 6) {
      delete c;
      return fail(nullptr, TPSB_EINVAL, "element %d has more than 6 faces", e1);
    }
    e2f[7 * e1 + nf + 1] = f;
    e2f[7 * e1] = nf + 1;
    if (e2 < NE) {
      nf = e2f[7 * e2];
      if (nf >= 6) {
        delete c;
        return fail(nullptr, TPSB_EINVAL, "element %d has more than 6 faces", e2);
      }
      e2f[7 * e2 + nf + 1] = f;
      e2f[7 * e2] = nf + 1;
      fl_el1.push_back(e1), fl_el2.push_back(e2), fl_inf1.push_back(maps->face_inf1[f]), fl_inf2.push_back(maps->face_inf2[f]);
    } else {
      sh_el1.push_back(e1), sh_el2.push_back(e2), sh_inf1.push_back(maps->face_inf1[f]), sh_inf2.push_back(maps->face_inf2[f]);
    }
  }
  c->element_to_faces = e2f;
  c->NFlocal = static_cast<int>(fl_el1.size());
  fl_el1.insert(fl_el1.end(), sh_el1.begin(), sh_el1.end());
  fl_el2.insert(fl_el2.end(), sh_el2.begin(), sh_el2.end());
  fl_inf1.insert(fl_inf1.end(), sh_inf1.begin(), sh_inf1.end());
  fl_inf2.insert(fl_inf2.end(), sh_inf2.begin(), sh_inf2.end());
  c->NFint = static_cast<int>(fl_el1.size());

  std::vector<int> nbr_elem(static_cast<size_t>(6) * NE, -1), nbr_code(static_cast<size_t>(6) * NE, 0);
  std::vector<int> el_face(static_cast<size_t>(6) * NE, -1), el_face_code(static_cast<size_t>(6) * NE, 0);
  std::vector<char> touches_shared(NE, 0);
  for (int fc = 0; fc < c->NFint; fc++) {
    const int e1 = fl_el1[fc], e2 = fl_el2[fc];
    const int lf1 = fl_inf1[fc] / 64, lf2 = fl_inf2[fc] / 64, ori = fl_inf2[fc] % 64;
    if (lf1 < 0 || lf1 > 5 || lf2 < 0 || lf2 > 5 || ori < 0 || ori > 7 || fl_inf1[fc] % 64 != 0) {
      delete c;
      return fail(nullptr, TPSB_EINVAL, "face info codes out of range on two-sided face %d", fc);
    }
    nbr_elem[e1 * 6 + lf1] = e2;
    nbr_code[e1 * 6 + lf1] = lf2 | (c->T.perm_code[ori] << 3);  // face coords == own coords: perm[ori]
    el_face[e1 * 6 + lf1] = fc;
    el_face_code[e1 * 6 + lf1] = 0;
    if (e2 < NE) {
      nbr_elem[e2 * 6 + lf2] = e1;
      nbr_code[e2 * 6 + lf2] = lf1 | (c->T.iperm_code[ori] << 3);  // own coords -> face coords: iperm[ori]
      el_face[e2 * 6 + lf2] = fc;
      el_face_code[e2 * 6 + lf2] = 1 | (c->T.iperm_code[ori] << 1);
    } else {
      touches_shared[e1] = 1;
    }
  }
  c->NFbdr = static_cast<int>(b_el1.size());
  for (int k = 0; k < c->NFbdr; k++) {
    nbr_elem[b_el1[k] * 6 + b_lf[k]] = -2 - b_bc[k];
    el_face[b_el1[k] * 6 + b_lf[k]] = c->NFint + k;  // lifted like Elem1's side: elvect -= phi Fhat w
    el_face_code[b_el1[k] * 6 + b_lf[k]] = 0;
  }
  std::vector<int> elem_list;
  elem_list.reserve(NE);
  for (int e = 0; e < NE; e++)
    if (!touches_shared[e]) elem_list.push_back(e);
  c->n_int_elems = static_cast<int>(elem_list.size());
  for (int e = 0; e < NE; e++)
    if (touches_shared[e]) elem_list.push_back(e);
  c->n_pb_elems = NE - c->n_int_elems;

  // ---- fast path tables (rhs_fast.cuh): affine metric per element, per-face block ids / normals ----
  // all-parallelepiped meshes run the fast path; anything else (trilinear/skewed elements) the general one
  c->fast = all_affine && !visc_mod;  // the SGS models need the full velocity gradient at the face points
  if (const char *pth = getenv("TPSB_PATH")) c->fast = c->fast && strcmp(pth, "legacy") != 0 && strcmp(pth, "general") != 0;
  c->fused = c->fast && c->np == 4;
  if (const char *pth = getenv("TPSB_PATH")) c->fused = c->fused && strcmp(pth, "unfused") != 0;
  std::vector<double> geo, face_nor;
  std::vector<int4> face_desc;
  std::vector<int> send_blk;
  const int nshared = c->NFint - c->NFlocal;
  if (c->fast) {
    geo.assign(static_cast<size_t>(NE) * GEO, 0.0);
    for (int e = 0; e < NE; e++) {
      const double *v = &maps->elem_vertices[static_cast<size_t>(e) * 24];
      double J[9], A[9];
      for (int i = 0; i < 3; i++) {
        J[i + 0] = v[1 * 3 + i] - v[i];
        J[i + 3] = v[3 * 3 + i] - v[i];
        J[i + 6] = v[4 * 3 + i] - v[i];
      }
      const double det = J[0] * (J[4] * J[8] - J[5] * J[7]) - J[3] * (J[1] * J[8] - J[2] * J[7]) +
                         J[6] * (J[1] * J[5] - J[2] * J[4]);
      A[0] = J[4] * J[8] - J[7] * J[5];
      A[3] = J[6] * J[5] - J[3] * J[8];
      A[6] = J[3] * J[7] - J[6] * J[4];
      A[1] = J[7] * J[2] - J[1] * J[8];
      A[4] = J[0] * J[8] - J[6] * J[2];
      A[7] = J[6] * J[1] - J[0] * J[7];
      A[2] = J[1] * J[5] - J[4] * J[2];
      A[5] = J[3] * J[2] - J[0] * J[5];
      A[8] = J[0] * J[4] - J[3] * J[1];
      double *g = &geo[static_cast<size_t>(e) * GEO];
      for (int q = 0; q < 9; q++) g[q] = A[q];
      g[9] = det;
      g[10] = 1.0 / det;
    }
    // shared faces: position of each one in the peer-grouped receive order, key (halo element, its local face)
    std::vector<int> shared_pos(nshared, 0);
    {
      std::vector<std::pair<long long, int>> key(nshared);
      for (int s = 0; s < nshared; s++) {
        const int fc = c->NFlocal + s;
        key[s] = {static_cast<long long>(fl_el2[fc] - NE) * 8 + fl_inf2[fc] / 64, s};
      }
      std::sort(key.begin(), key.end());
      for (int pos = 0; pos < nshared; pos++) shared_pos[key[pos].second] = pos;
    }
    face_desc.resize(c->NFint);
    face_nor.assign(static_cast<size_t>(c->NFint) * 4, 0.0);
    for (int fc = 0; fc < c->NFint; fc++) {
      const int e1 = fl_el1[fc], e2 = fl_el2[fc], lf1 = fl_inf1[fc] / 64, lf2 = fl_inf2[fc] / 64, ori = fl_inf2[fc] % 64;
      int4 d;
      d.x = e1 * 6 + lf1;
      d.y = e2 < NE ? e2 * 6 + lf2 : 6 * NE + shared_pos[fc - c->NFlocal];
      d.z = c->T.perm_code[ori];
      d.w = 0;
      face_desc[fc] = d;
      // CalcOrtho of Elem1's face Jacobian (constant on a parallelogram face): (X1 - X0) x (X3 - X0)
      const double *v = &maps->elem_vertices[static_cast<size_t>(e1) * 24];
      const int *fv = c->T.face_vert[lf1];
      double ts[3], tt[3];
      for (int i = 0; i < 3; i++) {
        ts[i] = v[fv[1] * 3 + i] - v[fv[0] * 3 + i];
        tt[i] = v[fv[3] * 3 + i] - v[fv[0] * 3 + i];
      }
      double *nr = &face_nor[static_cast<size_t>(fc) * 4];
      nr[0] = ts[1] * tt[2] - ts[2] * tt[1];
      nr[1] = ts[2] * tt[0] - ts[0] * tt[2];
      nr[2] = ts[0] * tt[1] - ts[1] * tt[0];
      nr[3] = std::sqrt(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]);
    }
  }

  if (NEH == 0) build_host_pipe(c, nbr_elem, el_face, fl_el1, fl_el2, b_el1);

  // ---- device allocations ----
  cudaError_t ce = cudaSetDevice(device);
  std::vector<double> vx(maps->elem_vertices, maps->elem_vertices + static_cast<size_t>(NE + NEH) * 24);
  if (ce == cudaSuccess) ce = upload(&c->d_vx, vx);
  if (ce == cudaSuccess && visc_mod) {
    // delta = Mesh::GetElementSize(e, 1) / order: smallest singular value of the Jacobian at the element centre
    std::vector<double> delta(static_cast<size_t>(NE + NEH));
    for (int e = 0; e < NE + NEH; e++) delta[e] = hex_min_size(&maps->elem_vertices[static_cast<size_t>(e) * 24]) / c->order;
    ce = upload(&c->d_elem_delta, delta);
  }
  if (ce == cudaSuccess) ce = upload(&c->d_nbr_elem, nbr_elem);
  if (ce == cudaSuccess) ce = upload(&c->d_nbr_code, nbr_code);
  if (ce == cudaSuccess) ce = upload(&c->d_face_el1, fl_el1);
  if (ce == cudaSuccess) ce = upload(&c->d_face_el2, fl_el2);
  if (ce == cudaSuccess) ce = upload(&c->d_face_inf1, fl_inf1);
  if (ce == cudaSuccess) ce = upload(&c->d_face_inf2, fl_inf2);
  if (ce == cudaSuccess) ce = upload(&c->d_el_face, el_face);
  if (ce == cudaSuccess) ce = upload(&c->d_el_face_code, el_face_code);
  if (ce == cudaSuccess) ce = upload(&c->d_elem_list, elem_list);
  const size_t nb = static_cast<size_t>(c->N) * sizeof(double);
  if (ce == cudaSuccess) ce = cudaMalloc(&c->d_Up, nb * NEQ);
  if (ce == cudaSuccess) ce = cudaMalloc(&c->d_gradUp, nb * NEQ * DIM);
  if (ce == cudaSuccess)
    ce = cudaMalloc(&c->d_faceRes, static_cast<size_t>(c->NFint + c->NFbdr) * NEQ * c->np * c->np * sizeof(double));
  if (ce == cudaSuccess) ce = upload(&c->d_bdr_el1, b_el1);
  if (ce == cudaSuccess) ce = upload(&c->d_bdr_lf, b_lf);
  if (ce == cudaSuccess) ce = upload(&c->d_bdr_bc, b_bc);
  if (ce == cudaSuccess) ce = cudaMalloc(&c->d_maxBits, sizeof(unsigned long long));
  if (ce == cudaSuccess) ce = cudaMalloc(&c->d_mcs, sizeof(double));
  if (ce == cudaSuccess) ce = cudaMemset(c->d_maxBits, 0, sizeof(unsigned long long));
  if (ce == cudaSuccess && device < MAX_DEV) {
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    if (!g_tables_uploaded[device]) {
      for (int p = 1; p <= 3 && ce == cudaSuccess; p++) {
        RefTables Tp;
        build_ref_tables(p, Tp);
        ce = cudaMemcpyToSymbol(c_Tab, &Tp, sizeof(RefTables), static_cast<size_t>(p - 1) * sizeof(RefTables));
      }
      g_tables_uploaded[device] = ce == cudaSuccess;
    }
  }
  if (c->fast) {
    if (ce == cudaSuccess) ce = upload(&c->d_geo, geo);
    if (ce == cudaSuccess) ce = upload(&c->d_face_desc, face_desc);
    if (ce == cudaSuccess) ce = upload(&c->d_face_nor, face_nor);
    const size_t blk = static_cast<size_t>(NTF) * c->np * c->np * sizeof(double);
    if (ce == cudaSuccess) ce = cudaMalloc(&c->d_tr, (static_cast<size_t>(6) * NE + nshared) * blk);
  }

  // ---- halo ----
  if (ce == cudaSuccess && NEH > 0) {
    const std::string herr = setup_halo_desc(c, halo, NEH);
    if (!herr.empty()) {
      tpsb_destroy(c);
      return fail(nullptr, TPSB_EINVAL, "%s", herr.c_str());
    }
    const int np = halo->num_nbr_ranks;
    std::vector<int> se(halo->send_elems, halo->send_elems + c->n_send);
    const size_t hb = static_cast<size_t>(c->NH) * sizeof(double), sb = static_cast<size_t>(c->n_send) * c->nd * sizeof(double);
    if (ce == cudaSuccess) ce = cudaMalloc(&c->d_Uhalo, hb * NEQ);
    if (ce == cudaSuccess) ce = cudaMalloc(&c->d_UpHalo, hb * NEQ);
    if (ce == cudaSuccess) ce = cudaMalloc(&c->d_gradUpHalo, hb * NEQ * DIM);
    if (ce == cudaSuccess) ce = cudaMalloc(&c->d_sendU, sb * NEQ);
    if (ce == cudaSuccess) ce = cudaMalloc(&c->d_sendG, sb * NEQ * DIM);
    if (c->fast) {
      // trace blocks this rank sends to peer p: for each send element (in send order = the receiver's halo
      // order) its local faces, ascending, whose neighbour is a halo element owned by p.  The receiver sorts
      // its shared faces by (halo element, that element's local face): the same order.
      auto peer_of = [&](int h) {
        for (int q = 0; q < np; q++)
          if (h >= c->recv_offset[q] && h < c->recv_offset[q + 1]) return q;
        return -1;
      };
      c->sendblk_offset.assign(np + 1, 0);
      c->recvblk_offset.assign(np + 1, 0);
      for (int q = 0; q < np; q++) {
        for (int k = c->send_offset[q]; k < c->send_offset[q + 1]; k++) {
          const int es = se[k];
          for (int lf = 0; lf < 6; lf++) {
            const int nb = nbr_elem[es * 6 + lf];
            if (nb >= NE && peer_of(nb - NE) == q) send_blk.push_back(es * 6 + lf);
          }
        }
        c->sendblk_offset[q + 1] = static_cast<int>(send_blk.size());
      }
      for (int s2 = 0; s2 < nshared; s2++) c->recvblk_offset[peer_of(fl_el2[c->NFlocal + s2] - NE) + 1]++;
      for (int q = 0; q < np; q++) c->recvblk_offset[q + 1] += c->recvblk_offset[q];
      c->n_send_blk = static_cast<int>(send_blk.size());
      if (c->n_send_blk != nshared) {
        tpsb_destroy(c);
        return fail(nullptr, TPSB_EINVAL, "halo description inconsistent: %d trace blocks to send, %d shared faces",
                    c->n_send_blk, nshared);
      }
      if (ce == cudaSuccess) ce = upload(&c->d_send_blk, send_blk);
      if (ce == cudaSuccess)
        ce = cudaMalloc(&c->d_sendTr, static_cast<size_t>(c->n_send_blk) * NTF * c->np * c->np * sizeof(double));
    }
  }
  if (ce != cudaSuccess) {
    const std::string msg = cudaGetErrorString(ce);
    tpsb_destroy(c);
    return fail(nullptr, TPSB_ECUDA, "device setup failed: %s", msg.c_str());
  }
  *out = c;
  return TPSB_OK;
}

void tpsb_destroy(tpsb_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  void *ptrs[] = {c->d_vx,       c->d_nbr_elem,  c->d_nbr_code,     c->d_face_el1, c->d_face_el2, c->d_face_inf1,
                  c->d_face_inf2, c->d_el_face,  c->d_el_face_code, c->d_elem_list, c->d_Up,       c->d_gradUp,
                  c->d_faceRes,  c->d_Uhalo,     c->d_UpHalo,       c->d_gradUpHalo, c->d_sendU,   c->d_sendG,
                  c->d_maxBits,  c->d_mcs,       c->d_send_elems,   c->d_k,        c->d_yv,       c->d_z,
                  c->d_hx,       c->d_hy,        c->d_geo,          c->d_tr,       c->d_face_nor, c->d_sendTr,
                  c->d_face_desc, c->d_send_blk, c->d_bdr_el1,      c->d_bdr_lf,   c->d_bdr_bc,   c->d_elem_delta,
                  c->d_nan_count, c->d_xiN3};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  for (auto &f : c->forcings)
    if (f.mix) cudaFree(f.mix);
  for (void *p : c->gen_allocs) cudaFree(p);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  if (c->ev_pack) cudaEventDestroy(c->ev_pack);
  if (c->ev_recvU) cudaEventDestroy(c->ev_recvU);
  if (c->ev_recvG) cudaEventDestroy(c->ev_recvG);
  if (c->ev_recvT) cudaEventDestroy(c->ev_recvT);
  if (c->ode_exec) cudaGraphExecDestroy(c->ode_exec);
  if (c->ode_stream) cudaStreamDestroy(c->ode_stream);
  if (c->ode_ev0) cudaEventDestroy(c->ode_ev0);
  if (c->ode_ev1) cudaEventDestroy(c->ode_ev1);
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  for (cudaEvent_t e : c->ev_in) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_out) cudaEventDestroy(e);
  if (c->ev_pipe0) cudaEventDestroy(c->ev_pipe0);
  if (c->ev_pipe1) cudaEventDestroy(c->ev_pipe1);
  delete c;
}

// Test hook (host only, no device needed): the chunk schedule of the host-buffer pipeline for a single-rank mesh whose
// faces are all two-sided.  ops = (kind, chunk) pairs: 0 copy-in + primitives, 1 gradient, 2 face range, 3 residual + copy-out.
int tpsb_debug_host_pipe_schedule(const tpsb_mesh_maps *maps, int chunks, int *elem_begin, int *face_begin, int *bdr_begin,
                                  int *ops, int max_ops, int *num_ops) {
  if (!maps || maps->dim != 3 || !elem_begin || !face_begin || !ops || !num_ops) return TPSB_EINVAL;
  const int NE = maps->num_elems, NF = maps->num_faces;
  std::vector<int> nbr(static_cast<size_t>(NE) * 6, -1), elf(static_cast<size_t>(NE) * 6, -1), e1, e2, b1, blf;
  for (int f = 0; f < NF; f++) {  // two-sided faces keep their order; boundary faces (el2 < 0) follow as slots NFint + k
    const int a1 = maps->face_el1[f], a2 = maps->face_el2[f];
    if (a1 < 0 || a2 >= NE) return TPSB_EINVAL;
    const int lf1 = maps->face_inf1[f] / 64;
    if (a2 < 0) {
      b1.push_back(a1), blf.push_back(lf1);
      continue;
    }
    const int lf2 = maps->face_inf2[f] / 64, fc = static_cast<int>(e1.size());
    nbr[static_cast<size_t>(a1) * 6 + lf1] = a2, nbr[static_cast<size_t>(a2) * 6 + lf2] = a1;
    elf[static_cast<size_t>(a1) * 6 + lf1] = fc, elf[static_cast<size_t>(a2) * 6 + lf2] = fc;
    e1.push_back(a1), e2.push_back(a2);
  }
  const int NFint = static_cast<int>(e1.size());
  for (size_t k = 0; k < b1.size(); k++) elf[static_cast<size_t>(b1[k]) * 6 + blf[k]] = NFint + static_cast<int>(k);
  std::vector<int> eb, fb, bb;
  std::vector<tpsb_ctx::PipeOp> po;
  if (!host_pipe_schedule(NE, NFint, chunks, nbr, elf, e1, e2, b1, eb, fb, bb, po)) return TPSB_ENOTIMPL;
  if (static_cast<int>(po.size()) > max_ops) return TPSB_EINVAL;
  std::copy(eb.begin(), eb.end(), elem_begin);
  std::copy(fb.begin(), fb.end(), face_begin);
  if (bdr_begin) std::copy(bb.begin(), bb.end(), bdr_begin);
  for (size_t k = 0; k < po.size(); k++) ops[2 * k] = po[k].kind, ops[2 * k + 1] = po[k].chunk;
  *num_ops = static_cast<int>(po.size());
  return TPSB_OK;
}

int64_t tpsb_num_dofs(const tpsb_ctx *c) { return c ? c->N : 0; }
int tpsb_num_equation(const tpsb_ctx *c) { return c ? c->neq : 0; }
int tpsb_get_path(const tpsb_ctx *c) { return !c ? -1 : c->generic ? 3 : c->fused ? 2 : c->fast ? 1 : 0; }
int64_t tpsb_launch_count(const tpsb_ctx *c) { return c ? c->launches : 0; }

int tpsb_get_element_to_faces(const tpsb_ctx *c, int *out) {
  if (!c || !out) return TPSB_EINVAL;
  std::copy(c->element_to_faces.begin(), c->element_to_faces.end(), out);
  return TPSB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
static KernelArgs make_args(tpsb_ctx *c, const double *d_x, double *d_y) {
  KernelArgs a;
  a.NE = c->NE;
  a.NEH = c->NEH;
  a.N = c->N;
  a.NH = c->NH;
  a.NFint = c->NFint;
  a.ND = c->nd;
  a.phys = c->phys;
  a.vx = c->d_vx;
  a.nbr_elem = c->d_nbr_elem;
  a.nbr_code = c->d_nbr_code;
  a.face_el1 = c->d_face_el1;
  a.face_el2 = c->d_face_el2;
  a.face_inf1 = c->d_face_inf1;
  a.face_inf2 = c->d_face_inf2;
  a.el_face = c->d_el_face;
  a.el_face_code = c->d_el_face_code;
  a.U = d_x;
  a.Uhalo = c->d_Uhalo;
  a.Up = c->d_Up;
  a.UpHalo = c->d_UpHalo;
  a.gradUp = c->d_gradUp;
  a.gradUpHalo = c->d_gradUpHalo;
  a.faceRes = c->d_faceRes;
  a.y = d_y;
  a.maxCharBits = c->d_maxBits;
  a.NFbdr = c->NFbdr;
  a.bdr_el1 = c->d_bdr_el1;
  a.bdr_lf = c->d_bdr_lf;
  a.bdr_bc = c->d_bdr_bc;
  a.bct = c->bct;
  a.geo = c->d_geo;
  a.elem_delta = c->d_elem_delta;
  a.rk = c->rk;
  a.tr = c->d_tr;
  a.face_desc = c->d_face_desc;
  a.face_nor = c->d_face_nor;
  auto al32 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 31u) == 0; };
  a.vec_ok = (al32(d_x) && al32(c->d_Up) && al32(c->d_gradUp) && al32(c->d_Uhalo) && al32(c->d_UpHalo) &&
              al32(c->d_gradUpHalo) && (c->N % 4 == 0))
                 ? 1
                 : 0;
  return a;
}

template <int NP, int EPB, int MINB = 1>
static void launch_grad(tpsb_ctx *c, const KernelArgs &a, int begin, int count, const int *list) {
  if (count <= 0) return;
  ProfScope ps(c, K_GRAD);
  grad_kernel<NP, EPB, MINB><<<(count + EPB - 1) / EPB, NP * NP * NP * EPB, 0, c->stream>>>(a, begin, count, list);
}
template <int NP, int FPB, int NTF>
static void launch_face(tpsb_ctx *c, const KernelArgs &a, int begin, int count) {
  if (count <= 0) return;
  ProfScope ps(c, K_FACE);
  if (c->phys.sgs_model | c->phys.sponge)
    face_flux_kernel<NP, FPB, NTF, false, true><<<(count + FPB - 1) / FPB, NTF, 0, c->stream>>>(a, begin, count, nullptr);
  else
    face_flux_kernel<NP, FPB, NTF, false, false><<<(count + FPB - 1) / FPB, NTF, 0, c->stream>>>(a, begin, count, nullptr);
}
// boundary faces: BCintegrator (src/BCintegrator.cpp:295-441), same kernel in one-sided mode
static void bdr_faces(tpsb_ctx *c, const KernelArgs &a, int begin = 0, int count = -1) {
  if (count < 0) count = c->NFbdr;
  if (count <= 0) return;
  ProfScope ps(c, K_FACE);
  const bool mod = (c->phys.sgs_model | c->phys.sponge) != 0;
  if (c->np == 4) {
    if (mod) face_flux_kernel<4, 4, 160, true, true><<<(count + 3) / 4, 160, 0, c->stream>>>(a, begin, count, nullptr);
    else face_flux_kernel<4, 4, 160, true, false><<<(count + 3) / 4, 160, 0, c->stream>>>(a, begin, count, nullptr);
  } else if (c->np == 3) {
    if (mod) face_flux_kernel<3, 4, 128, true, true><<<(count + 3) / 4, 128, 0, c->stream>>>(a, begin, count, nullptr);
    else face_flux_kernel<3, 4, 128, true, false><<<(count + 3) / 4, 128, 0, c->stream>>>(a, begin, count, nullptr);
  } else {
    if (mod) face_flux_kernel<2, 8, 128, true, true><<<(count + 7) / 8, 128, 0, c->stream>>>(a, begin, count, nullptr);
    else face_flux_kernel<2, 8, 128, true, false><<<(count + 7) / 8, 128, 0, c->stream>>>(a, begin, count, nullptr);
  }
}
template <int NP, int EPB, int MINB = 1>
static void launch_resid(tpsb_ctx *c, const KernelArgs &a, int begin, int count) {
  if (count <= 0) return;
  ProfScope ps(c, K_RESID);
  const dim3 grid((count + EPB - 1) / EPB), block(NP * NP * NP * EPB);
  const bool mod = (c->phys.sgs_model | c->phys.sponge) != 0, rk = a.rk.X != nullptr;
#define RESID_LAUNCH(AFF, MOD, RK) elem_resid_kernel<NP, EPB, MINB, AFF, MOD, RK><<<grid, block, 0, c->stream>>>(a, begin, count)
  if (c->fast) {
    if (rk) RESID_LAUNCH(true, false, true);
    else RESID_LAUNCH(true, false, false);
  } else if (mod) {
    if (rk) RESID_LAUNCH(false, true, true);
    else RESID_LAUNCH(false, true, false);
  } else {
    if (rk) RESID_LAUNCH(false, false, true);
    else RESID_LAUNCH(false, false, false);
  }
#undef RESID_LAUNCH
}

// Launch-shape selection.  p = 3 is the tuned case; tune[] (TPSB_TUNE="g,f,r", development knob) picks
// among pre-instantiated CTA shapes.
static void grad(tpsb_ctx *c, const KernelArgs &a, int begin, int count, const int *list) {
  if (c->np == 4) {
    switch (c->tune[0]) {
      case 1: launch_grad<4, 1, 16>(c, a, begin, count, list); break;
      case 2: launch_grad<4, 2, 6>(c, a, begin, count, list); break;
      case 3: launch_grad<4, 2, 8>(c, a, begin, count, list); break;
      case 4: launch_grad<4, 4, 3>(c, a, begin, count, list); break;
      case 5: launch_grad<4, 1, 8>(c, a, begin, count, list); break;
      default: launch_grad<4, 1, 12>(c, a, begin, count, list); break;
    }
  } else if (c->np == 3) {
    launch_grad<3, 8, 3>(c, a, begin, count, list);
  } else {
    launch_grad<2, 16, 4>(c, a, begin, count, list);
  }
}
static void face(tpsb_ctx *c, const KernelArgs &a, int begin, int count) {
  if (c->np == 4) {
    switch (c->tune[1]) {
      case 1: launch_face<4, 4, 128>(c, a, begin, count); break;
      case 2: launch_face<4, 2, 96>(c, a, begin, count); break;
      case 3: launch_face<4, 4, 192>(c, a, begin, count); break;
      case 4: launch_face<4, 3, 128>(c, a, begin, count); break;
      case 5: launch_face<4, 4, 256>(c, a, begin, count); break;
      default: launch_face<4, 4, 160>(c, a, begin, count); break;
    }
  } else if (c->np == 3) {
    launch_face<3, 4, 128>(c, a, begin, count);
  } else {
    launch_face<2, 8, 128>(c, a, begin, count);
  }
}
static void resid(tpsb_ctx *c, const KernelArgs &a, int begin = 0, int count = -1) {
  if (count < 0) count = c->NE;
  if (c->np == 4) {
    switch (c->tune[2]) {
      case 1: launch_resid<4, 1, 16>(c, a, begin, count); break;
      case 2: launch_resid<4, 2, 5>(c, a, begin, count); break;
      case 3: launch_resid<4, 1, 12>(c, a, begin, count); break;
      case 4: launch_resid<4, 4, 2>(c, a, begin, count); break;
      case 5: launch_resid<4, 2, 6>(c, a, begin, count); break;
      case 6: launch_resid<4, 1, 10>(c, a, begin, count); break;
      default:
        // the trilinear instantiation rebuilds the metric per node: 102 registers (10 CTAs / SM) beat 80 with spills
        if (c->fast) launch_resid<4, 1, 12>(c, a, begin, count);
        else launch_resid<4, 1, 10>(c, a, begin, count);
        break;
    }
  } else if (c->np == 3) {
    launch_resid<3, 8, 2>(c, a, begin, count);
  } else {
    launch_resid<2, 16, 4>(c, a, begin, count);
  }
}
#define DISPATCH(c, CALL) CALL

static void launch_prim(tpsb_ctx *c, const KernelArgs &a, int halo) {
  const long long cnt = halo ? c->NH : c->N;
  if (cnt <= 0) return;
  ProfScope ps(c, K_PRIM);
  prim_kernel<<<static_cast<unsigned>((cnt + 255) / 256), 256, 0, c->stream>>>(a, halo);
}

// One grouped ncclSend/ncclRecv round on the communication stream: replaces
// RHSoperator::initNBlockDataTransfer (MPI_Isend/Irecv per neighbour, src/rhs_operator.cpp:775-822).
static int exchange(tpsb_ctx *ctx, const double *src, int nfld, double *sendbuf, double *recvbuf, cudaEvent_t done) {
  tpsb_ctx *c = ctx;
  const long long total = static_cast<long long>(c->n_send) * nfld * c->nd;
  {
    ProfScope ps(c, K_PACK);
    pack_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, c->stream>>>(c->n_send, c->nd, nfld, c->N,
                                                                                  c->d_send_elems, src, sendbuf);
  }
  CU(cudaEventRecord(c->ev_pack, c->stream));
  CU(cudaStreamWaitEvent(c->comm_stream, c->ev_pack, 0));
  const size_t per = static_cast<size_t>(nfld) * c->nd;
  NC(ncclGroupStart());
  for (size_t p = 0; p < c->nbr_rank.size(); p++) {
    NC(ncclSend(sendbuf + c->send_offset[p] * per, (c->send_offset[p + 1] - c->send_offset[p]) * per, ncclDouble,
                c->nbr_rank[p], c->comm, c->comm_stream));
    NC(ncclRecv(recvbuf + c->recv_offset[p] * per, (c->recv_offset[p + 1] - c->recv_offset[p]) * per, ncclDouble,
                c->nbr_rank[p], c->comm, c->comm_stream));
  }
  NC(ncclGroupEnd());
  CU(cudaEventRecord(done, c->comm_stream));
  return TPSB_OK;
}

// gradient pass = updatePrimitives + Gradients::computeGradients, with the exchange overlapped as in
// the reference's GPU branch (src/rhs_operator.cpp:349-361)
static int run_gradients(tpsb_ctx *ctx, const KernelArgs &a, bool prims_done) {
  tpsb_ctx *c = ctx;
  if (!prims_done) launch_prim(c, a, 0);
  if (c->NEH > 0) {
    int rc = exchange(c, a.U, NEQ, c->d_sendU, c->d_Uhalo, c->ev_recvU);
    if (rc) return rc;
    DISPATCH(c, grad(c, a, 0, c->n_int_elems, c->d_elem_list));
    CU(cudaStreamWaitEvent(c->stream, c->ev_recvU, 0));
    launch_prim(c, a, 1);
    DISPATCH(c, grad(c, a, c->n_int_elems, c->n_pb_elems, c->d_elem_list));
  } else {
    DISPATCH(c, grad(c, a, 0, c->NE, nullptr));
  }
  CU(cudaGetLastError());
  return TPSB_OK;
}

// ---- fast path (rhs_fast.cuh) ----
template <int NP, int EPB, int MINB>
static void launch_grad_trace(tpsb_ctx *c, const KernelArgs &a, int begin, int count, const int *list) {
  if (count <= 0) return;
  ProfScope ps(c, K_GRAD);
  grad_trace_kernel<NP, EPB, MINB><<<(count + EPB - 1) / EPB, NP * NP * NP * EPB, 0, c->stream>>>(a, begin, count, list);
}
template <int NP, int WPB, int MINB>
static void launch_face_fast(tpsb_ctx *c, const KernelArgs &a, int begin, int count) {
  if (count <= 0) return;
  ProfScope ps(c, K_FACE);
  int grid = (count + WPB - 1) / WPB;
  const int cap = c->num_sms * c->face_ctas_per_sm;
  if (grid > cap) grid = cap;
  const size_t smem = face_fast_smem_bytes<NP>(WPB);
  static bool attr_set[MAX_DEV] = {};  // per instantiation and device (cudaFuncSetAttribute acts on the current device)
  if (!attr_set[c->device]) {
    cudaFuncSetAttribute(face_flux_fast_kernel<NP, WPB, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    attr_set[c->device] = true;
  }
  face_flux_fast_kernel<NP, WPB, MINB><<<grid, 32 * WPB, smem, c->stream>>>(a, begin, count);
}
template <int EPB, int MINB>
static void launch_grad_trace_mma(tpsb_ctx *c, const KernelArgs &a, int begin, int count, const int *list) {
  if (count <= 0) return;
  ProfScope ps(c, K_GRAD);
  grad_trace_mma_kernel<EPB, MINB><<<(count + EPB - 1) / EPB, 64 * EPB, 0, c->stream>>>(a, begin, count, list);
}
static void grad_trace(tpsb_ctx *c, const KernelArgs &a, int begin, int count, const int *list) {
  if (c->np == 4) {
    switch (c->tune[0]) {
      case 6: launch_grad_trace<4, 1, 10>(c, a, begin, count, list); break;  // DFMA form (before the DMMA kernel)
      case 7: launch_grad_trace_mma<2, 4>(c, a, begin, count, list); break;
      case 8: launch_grad_trace_mma<1, 6>(c, a, begin, count, list); break;
      case 1: launch_grad_trace<4, 1, 12>(c, a, begin, count, list); break;
      case 2: launch_grad_trace<4, 2, 5>(c, a, begin, count, list); break;
      case 3: launch_grad_trace<4, 2, 6>(c, a, begin, count, list); break;
      case 4: launch_grad_trace<4, 4, 3>(c, a, begin, count, list); break;
      case 5: launch_grad_trace<4, 1, 8>(c, a, begin, count, list); break;
      default: launch_grad_trace_mma<1, 9>(c, a, begin, count, list); break;
    }
  } else if (c->np == 3) {
    launch_grad_trace<3, 4, 4>(c, a, begin, count, list);
  } else {
    launch_grad_trace<2, 16, 4>(c, a, begin, count, list);
  }
}
static int comm_grain();
static int face_grain();
template <int WPB, int MINB>
static void launch_face_mma(tpsb_ctx *c, const KernelArgs &a, int begin, int count) {
  if (count <= 0) return;
  ProfScope ps(c, K_FACE);
  int grid = (count + WPB - 1) / WPB;
  const int cap = c->num_sms * MINB;
  if (grid > cap) {
    const int g = c->NEH > 0 ? comm_grain() : face_grain();  // partitioned runs: warps retire after g faces so the exchange kernels get SM slots
    grid = g > 0 ? std::max(cap, (count + WPB * g - 1) / (WPB * g)) : cap;
  }
  const size_t smem = face_mma_smem_bytes(WPB);
  static bool attr_set[MAX_DEV] = {};  // per instantiation and device
  if (!attr_set[c->device]) {
    cudaFuncSetAttribute(face_flux_mma_kernel<WPB, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    attr_set[c->device] = true;
  }
  face_flux_mma_kernel<WPB, MINB><<<grid, 32 * WPB, smem, c->stream>>>(a, begin, count);
}
static void face_fast(tpsb_ctx *c, const KernelArgs &a, int begin, int count) {
  if (c->np == 4) {
    switch (c->tune[1]) {
      case 6: launch_face_fast<4, 4, 5>(c, a, begin, count); break;  // DFMA form (before the DMMA kernel)
      case 7: launch_face_mma<4, 3>(c, a, begin, count); break;
      case 8: launch_face_mma<2, 8>(c, a, begin, count); break;
      case 9: launch_face_mma<8, 2>(c, a, begin, count); break;
      case 1: launch_face_fast<4, 4, 4>(c, a, begin, count); break;
      case 2: launch_face_fast<4, 2, 10>(c, a, begin, count); break;
      case 3: launch_face_fast<4, 8, 2>(c, a, begin, count); break;
      case 4: launch_face_fast<4, 6, 3>(c, a, begin, count); break;
      default: launch_face_mma<4, 4>(c, a, begin, count); break;
    }
  } else if (c->np == 3) {
    launch_face_fast<3, 8, 2>(c, a, begin, count);
  } else {
    launch_face_fast<2, 8, 2>(c, a, begin, count);
  }
}

// grouped ncclSend/ncclRecv of the shared faces' trace blocks (replaces the gradUp half of
// RHSoperator::initNBlockDataTransfer: 1.28 KB per shared face instead of a 7.7 KB element)
static int exchange_traces(tpsb_ctx *ctx) {
  tpsb_ctx *c = ctx;
  const int bd = NTF * c->np * c->np;
  const long long total = static_cast<long long>(c->n_send_blk) * bd;
  {
    ProfScope ps(c, K_PACK);
    pack_blocks_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, c->stream>>>(c->n_send_blk, bd, c->d_send_blk,
                                                                                         c->d_tr, c->d_sendTr);
  }
  CU(cudaEventRecord(c->ev_pack, c->stream));
  CU(cudaStreamWaitEvent(c->comm_stream, c->ev_pack, 0));
  double *recv = c->d_tr + static_cast<size_t>(6) * c->NE * bd;
  NC(ncclGroupStart());
  for (size_t p = 0; p < c->nbr_rank.size(); p++) {
    NC(ncclSend(c->d_sendTr + static_cast<size_t>(c->sendblk_offset[p]) * bd,
                static_cast<size_t>(c->sendblk_offset[p + 1] - c->sendblk_offset[p]) * bd, ncclDouble, c->nbr_rank[p], c->comm,
                c->comm_stream));
    NC(ncclRecv(recv + static_cast<size_t>(c->recvblk_offset[p]) * bd,
                static_cast<size_t>(c->recvblk_offset[p + 1] - c->recvblk_offset[p]) * bd, ncclDouble, c->nbr_rank[p], c->comm,
                c->comm_stream));
  }
  NC(ncclGroupEnd());
  CU(cudaEventRecord(c->ev_recvT, c->comm_stream));
  return TPSB_OK;
}

static int run_gradients_fast(tpsb_ctx *ctx, const KernelArgs &a, bool prims_done) {
  tpsb_ctx *c = ctx;
  if (!prims_done) launch_prim(c, a, 0);
  if (c->NEH > 0) {
    int rc = exchange(c, a.U, NEQ, c->d_sendU, c->d_Uhalo, c->ev_recvU);
    if (rc) return rc;
    grad_trace(c, a, 0, c->n_int_elems, c->d_elem_list);
    CU(cudaStreamWaitEvent(c->stream, c->ev_recvU, 0));
    launch_prim(c, a, 1);
    grad_trace(c, a, c->n_int_elems, c->n_pb_elems, c->d_elem_list);
  } else {
    grad_trace(c, a, 0, c->NE, nullptr);
  }
  CU(cudaGetLastError());
  return TPSB_OK;
}

static void resid(tpsb_ctx *c, const KernelArgs &a, int begin, int count);

// Partitioned runs: a persistent grid that fills every SM keeps the NCCL send / receive kernels of the overlapped exchange
// out until it drains -- stream priority does not preempt resident CTAs -- so the "overlapped" exchange ran after the
// interior kernel (0.8 ms of idle stream per evaluation at 8 ranks).  With a halo the element and face kernels therefore use
// CTAs that retire after `grain` elements (faces per warp), which frees SM slots every few tens of microseconds.
// TPSB_COMM_GRAIN=0 restores the persistent grids.
static int comm_grain() {
  static const int g = getenv("TPSB_COMM_GRAIN") ? std::max(0, atoi(getenv("TPSB_COMM_GRAIN"))) : 8;
  return g;
}
// Single-rank runs: CTAs of the element kernel that retire after 32 elements are also 3-5 % faster than one persistent CTA
// per resident slot (96^3: 5.38 against 5.65 ms; grains 4 / 8 / 16 / 32 / 64 / 128 / 256 measured: 5.63 / 5.52 / 5.48 / 5.38 /
// 5.46 / 5.57 / 5.72 ms) -- the resident CTAs drift out of phase instead of hitting the DMMA, shared-memory and load phases
// together, and the tail is one short CTA instead of the slowest of 1184 long ones.  0 = persistent.
static int elem_grain() {
  static const int g = getenv("TPSB_ELEM_GRAIN") ? std::max(0, atoi(getenv("TPSB_ELEM_GRAIN"))) : 32;
  return g;
}
// the same for the face kernel (faces per warp): 0 / 4 / 8 / 16 / 32 / 64 -> 3.904 / 4.009 / 3.913 / 3.862 / 3.853 / 3.879 ms
static int face_grain() {
  static const int g = getenv("TPSB_FACE_GRAIN") ? std::max(0, atoi(getenv("TPSB_FACE_GRAIN"))) : 32;
  return g;
}
// ---- fused fast path (rhs_fused.cuh) ----
static void elem_fused(tpsb_ctx *c, const KernelArgs &a, int begin, int count, const int *list, int mode) {
  if (count <= 0) return;
  ProfScope ps(c, K_GRAD);
// REGS: register cap of the instantiation; the persistent grid is exactly what is resident (occupancy query per device)
#define FUSED_LAUNCH(REGS)                                                                                             \
  do {                                                                                                                 \
    static int per_sm[MAX_DEV] = {};                                                                                   \
    if (!per_sm[c->device]) {                                                                                          \
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[c->device], elem_fused_kernel<REGS>, 64, 0);               \
      if (per_sm[c->device] < 1) per_sm[c->device] = 1;                                                                \
    }                                                                                                                  \
    int grid = std::min(count, c->num_sms * per_sm[c->device]);                                                       \
    const int grain = c->NEH > 0 ? comm_grain() : elem_grain();                                                       \
    if (grain > 0) grid = std::min(count, std::max(grid, (count + grain - 1) / grain));                               \
    elem_fused_kernel<REGS><<<grid, 64, 0, c->stream>>>(a, begin, count, list, mode);                                  \
  } while (0)
  switch (c->tune[0]) {
    case 1: FUSED_LAUNCH(96); break;
    case 2: FUSED_LAUNCH(104); break;
    case 3: FUSED_LAUNCH(112); break;
    case 4: FUSED_LAUNCH(144); break;
    case 5: FUSED_LAUNCH(168); break;
    default: FUSED_LAUNCH(128); break;
  }
#undef FUSED_LAUNCH
}
static void lift(tpsb_ctx *c, const KernelArgs &a, int begin = 0, int count = -1) {
  if (count < 0) count = c->NE;
  if (count <= 0) return;
  ProfScope ps(c, K_RESID);
  static const int per_sm = getenv("TPSB_LIFT_CTAS") ? std::max(1, atoi(getenv("TPSB_LIFT_CTAS"))) : 2;
  const int grid = std::min((count + 3) / 4, c->num_sms * per_sm);
  if (a.rk.X) lift_kernel<true><<<grid, 256, 0, c->stream>>>(a, begin, count);
  else lift_kernel<false><<<grid, 256, 0, c->stream>>>(a, begin, count);
}
// element pass with the exchange of the partition-boundary elements' state overlapped (src/rhs_operator.cpp:349-361)
static int run_elem_fused(tpsb_ctx *ctx, const KernelArgs &a, int mode) {
  tpsb_ctx *c = ctx;
  if (c->NEH > 0) {
    int rc = exchange(c, a.U, NEQ, c->d_sendU, c->d_Uhalo, c->ev_recvU);
    if (rc) return rc;
    elem_fused(c, a, 0, c->n_int_elems, c->d_elem_list, mode);
    CU(cudaStreamWaitEvent(c->stream, c->ev_recvU, 0));
    elem_fused(c, a, c->n_int_elems, c->n_pb_elems, c->d_elem_list, mode);
  } else {
    elem_fused(c, a, 0, c->NE, nullptr, mode);
  }
  CU(cudaGetLastError());
  return TPSB_OK;
}
static int run_mult_fused(tpsb_ctx *ctx, const double *d_x, double *d_y) {
  tpsb_ctx *c = ctx;
  CU(cudaSetDevice(c->device));
  KernelArgs a = make_args(c, d_x, d_y);
  CU(cudaMemsetAsync(c->d_maxBits, 0, sizeof(unsigned long long), c->stream));
  int rc = run_elem_fused(c, a, c->forcing_needs_grad ? (FUSED_WRITE | FUSED_EXPORT) : FUSED_WRITE);
  if (rc) return rc;
  c->fields_x = d_x;
  c->fields_stale = true;
  if (c->NEH > 0) {
    rc = exchange_traces(c);
    if (rc) return rc;
    face_fast(c, a, 0, c->NFlocal);
    CU(cudaStreamWaitEvent(c->stream, c->ev_recvT, 0));
    face_fast(c, a, c->NFlocal, c->NFint - c->NFlocal);
  } else {
    face_fast(c, a, 0, c->NFint);
  }
  bdr_faces(c, a);
  lift(c, a);
  CU(cudaGetLastError());
  return TPSB_OK;
}
// primitives and gradients on request (RHSoperator::updatePrimitives / updateGradients, getGradientGF)
static int run_gradients_fused(tpsb_ctx *ctx, const KernelArgs &a, bool prims_done) {
  tpsb_ctx *c = ctx;
  if (!prims_done) launch_prim(c, a, 0);
  const int rc = run_elem_fused(c, a, FUSED_EXPORT);
  c->fields_stale = false;
  return rc;
}

static int run_mult_fast(tpsb_ctx *ctx, const double *d_x, double *d_y) {
  tpsb_ctx *c = ctx;
  if (c->fused) return run_mult_fused(ctx, d_x, d_y);
  CU(cudaSetDevice(c->device));
  KernelArgs a = make_args(c, d_x, d_y);
  CU(cudaMemsetAsync(c->d_maxBits, 0, sizeof(unsigned long long), c->stream));
  int rc = run_gradients_fast(c, a, false);
  if (rc) return rc;
  if (c->NEH > 0) {
    rc = exchange_traces(c);
    if (rc) return rc;
    face_fast(c, a, 0, c->NFlocal);
    CU(cudaStreamWaitEvent(c->stream, c->ev_recvT, 0));
    face_fast(c, a, c->NFlocal, c->NFint - c->NFlocal);
  } else {
    face_fast(c, a, 0, c->NFint);
  }
  bdr_faces(c, a);
  resid(c, a);
  CU(cudaGetLastError());
  return TPSB_OK;
}

// ---- generic path (rhs_generic.cuh) ----
// One CTA per element; the per-point loops have max(dof, nqv) independent items (9 / 16 on a p = 2 quadrilateral), and
// the kernels hold 96-128 registers per thread: sizing the CTA to the element instead of a fixed 128 threads keeps
// 2-4x more elements resident per SM (C1, 25 600 quads: 4.6 -> see DESIGN.md ms per evaluation).
// dry-air size combinations with compile-time instantiations of the generic kernels: 1 planar 2-D, 2 axisymmetric, 3 3-D
static int gen_specialisation(const GenArgs &g) {
  static const bool off = getenv("TPSB_GEN_SPEC") && atoi(getenv("TPSB_GEN_SPEC")) == 0;
  if (off || g.phys.fluid) return 0;
  if (g.dim == 2 && g.nvel == 2 && g.neq == 4) return 1;
  if (g.dim == 2 && g.nvel == 3 && g.neq == 5) return 2;
  if (g.dim == 3 && g.nvel == 3 && g.neq == 5) return 3;
  return 0;
}
static int gen_block_threads(const GenArgs &g) {
  static const int forced = getenv("TPSB_GEN_THREADS") ? atoi(getenv("TPSB_GEN_THREADS")) : 0;
  if (forced >= 32 && forced <= 128) return (forced / 32) * 32;
  const int items = std::max(std::max(g.dof, g.nqv), g.nfe * g.nqf);
  return std::min(128, std::max(32, ((items + 31) / 32) * 32));
}
static int exchange(tpsb_ctx *ctx, const double *src, int nfld, double *sendbuf, double *recvbuf, cudaEvent_t done);
static int run_gradients_generic(tpsb_ctx *ctx, const double *d_x, bool prims_done) {
  tpsb_ctx *c = ctx;
  GenArgs g = c->gen;
  g.U = d_x;
  if (c->NEH > 0) {  // face-neighbour state first (RHSoperator::Mult, src/rhs_operator.cpp:349-361)
    int rc = exchange(c, d_x, g.neq, c->d_sendU, c->d_Uhalo, c->ev_recvU);
    if (rc) return rc;
  }
  if (!prims_done) {
    ProfScope ps(c, K_PRIM);
    gen_prim_kernel<<<static_cast<unsigned>((g.N + 255) / 256), 256, 0, c->stream>>>(g);
  }
  if (c->NEH > 0) {
    CU(cudaStreamWaitEvent(c->stream, c->ev_recvU, 0));
    ProfScope ps(c, K_PRIM);
    const long long nh = static_cast<long long>(c->NEH) * g.dof;
    gen_prim_halo_kernel<<<static_cast<unsigned>((nh + 255) / 256), 256, 0, c->stream>>>(g);
  }
  {
    ProfScope ps(c, K_GRAD);
    const size_t smem = gen_grad_smem(g);
    const int spec = gen_specialisation(g);
#define GEN_LAUNCH_GRAD(D, V, Q)                                                                                          \
  do {                                                                                                                    \
    if (smem > 48 * 1024)                                                                                                 \
      CU(cudaFuncSetAttribute(gen_grad_kernel<D, V, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
    gen_grad_kernel<D, V, Q><<<g.NE, gen_block_threads(g), smem, c->stream>>>(g);                                       \
  } while (0)
    if (spec == 1) GEN_LAUNCH_GRAD(2, 2, 4);
    else if (spec == 2) GEN_LAUNCH_GRAD(2, 3, 5);
    else if (spec == 3) GEN_LAUNCH_GRAD(3, 3, 5);
    else GEN_LAUNCH_GRAD(0, 0, 0);
#undef GEN_LAUNCH_GRAD
  }
  CU(cudaGetLastError());
  return TPSB_OK;
}
static int run_mult_generic(tpsb_ctx *ctx, const double *d_x, double *d_y) {
  tpsb_ctx *c = ctx;
  CU(cudaSetDevice(c->device));
  CU(cudaMemsetAsync(c->d_maxBits, 0, sizeof(unsigned long long), c->stream));
  int rc = run_gradients_generic(c, d_x, false);
  if (rc) return rc;
  GenArgs g = c->gen;
  g.U = d_x;
  g.y = d_y;
  g.bc_dt = c->bc_dt;
  if (!c->nr_patches.empty()) {  // bcIntegrator->updateBCMean(Up) (rhs_operator.cpp:364): patch means on the device
    ProfScope ps(c, K_FACE);
    const int np = static_cast<int>(c->nr_patches.size());
    gen_nr_sum_kernel<<<np, 256, 0, c->stream>>>(g);
    if (c->comm)  // the MPI_Allreduce of updateMean; every rank holds every patch (possibly with no faces of it)
      for (int i = 0; i < np; i++)
        NC(ncclAllReduce(c->nr_patches[i].sums, c->nr_patches[i].sums, g.neq + 2, ncclDouble, ncclSum, c->comm, c->stream));
    gen_nr_finish_kernel<<<np, 32, 0, c->stream>>>(g);
  }
  if (c->NEH > 0 && g.eq_system != 0) {  // face-neighbour gradients for the viscous face fluxes (rhs_operator.cpp:363-375)
    rc = exchange(c, c->d_gradUp, g.neq * g.dim, c->d_sendG, c->d_gradUpHalo, c->ev_recvG);
    if (rc) return rc;
    CU(cudaStreamWaitEvent(c->stream, c->ev_recvG, 0));
  }
  {
    ProfScope ps(c, K_RESID);
    const size_t smem = gen_resid_smem(g);
    const int spec = gen_specialisation(g);
#define GEN_LAUNCH_RESID(D, V, Q)                                                                                          \
  do {                                                                                                                     \
    if (smem > 48 * 1024)                                                                                                  \
      CU(cudaFuncSetAttribute(gen_resid_kernel<D, V, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
    gen_resid_kernel<D, V, Q><<<g.NE, gen_block_threads(g), smem, c->stream>>>(g);                                       \
  } while (0)
    if (spec == 1) GEN_LAUNCH_RESID(2, 2, 4);
    else if (spec == 2) GEN_LAUNCH_RESID(2, 3, 5);
    else if (spec == 3) GEN_LAUNCH_RESID(3, 3, 5);
    else GEN_LAUNCH_RESID(0, 0, 0);
#undef GEN_LAUNCH_RESID
  }
  if (g.phys.lte && c->lte_radiation) {  // SourceTerm of the LTE fluid: the radiative sink only (source_term.cpp:205-207)
    ProfScope ps(c, K_RESID);
    gen_lte_source_kernel<<<static_cast<unsigned>((g.N + 127) / 128), 128, 0, c->stream>>>(g);
  }
  if (g.phys.fluid) {  // forcing terms are added after Me^-1 (rhs_operator.cpp:451-461)
    ProfScope ps(c, K_RESID);
    gen_source_kernel<<<static_cast<unsigned>((g.N + 127) / 128), 128, 0, c->stream>>>(g, c->sol_view ? c->sol_view : d_x);
  }
  if (g.phys.axisym) {  // AxisymmetricSource, registered after SourceTerm (rhs_operator.cpp:159-160)
    ProfScope ps(c, K_RESID);
    gen_axisym_source_kernel<<<static_cast<unsigned>((g.N + 127) / 128), 128, 0, c->stream>>>(g, c->sol_view ? c->sol_view : d_x);
  }
  CU(cudaGetLastError());
  return TPSB_OK;
}

// ForcingTerms::updateTerms of every registered term (src/rhs_operator.cpp:451-461), after Me^-1
static int apply_forcings(tpsb_ctx *ctx, const double *d_x, double *d_y) {
  tpsb_ctx *c = ctx;
  if (c->forcings.empty()) return TPSB_OK;
  ForcingArgs a;
  memset(&a, 0, sizeof(a));
  a.NE = c->NE, a.N = c->N, a.dof = c->nd, a.x = d_x, a.y = d_y, a.gradUp = c->d_gradUp;
  if (c->generic) {
    a.dim = c->gen.dim, a.nvel = c->gen.nvel, a.neq = c->gen.neq, a.nv = c->gen.nv, a.phys = c->gen.phys;
    a.vx = c->gen.vx, a.xiN = c->gen.xiN;
  } else {
    a.dim = 3, a.nvel = 3, a.neq = NEQ, a.nv = 8, a.vx = c->d_vx, a.xiN = c->d_xiN3;
    a.phys.dim = 3, a.phys.nvel = 3, a.phys.neq = NEQ, a.phys.fluid = 0, a.phys.dry = c->phys;
  }
  a.nf = static_cast<int>(c->forcings.size());
  for (int i = 0; i < a.nf; i++) a.f[i] = c->forcings[i];
  const unsigned nb = static_cast<unsigned>((c->N + 127) / 128);
  for (int i = 0; i < a.nf; i++)
    if (a.f[i].kind == TPSB_FORCING_SPONGE_ZONE && a.f[i].sz_mixed) {  // SpongeZone::computeMixedOutValues
      CU(cudaMemsetAsync(a.f[i].mix, 0, sizeof(double) * (a.neq + 1), c->stream));
      forcing_mixed_out_sum_kernel<<<nb, 128, 0, c->stream>>>(a, i);
      c->launches++;
      if (c->comm) NC(ncclAllReduce(a.f[i].mix, a.f[i].mix, a.neq + 1, ncclDouble, ncclSum, c->comm, c->stream));
      forcing_mixed_out_target_kernel<<<1, 1, 0, c->stream>>>(a, i);
      c->launches++;
    }
  {
    ProfScope ps(c, K_RESID);
    forcing_kernel<<<nb, 128, 0, c->stream>>>(a);
  }
  CU(cudaGetLastError());
  return TPSB_OK;
}

static int run_mult_paths(tpsb_ctx *ctx, const double *d_x, double *d_y);
static int run_mult(tpsb_ctx *ctx, const double *d_x, double *d_y) {
  const int rc = run_mult_paths(ctx, d_x, d_y);
  if (rc || ctx->forcings.empty()) return rc;
  return apply_forcings(ctx, d_x, d_y);
}
static int run_mult_paths(tpsb_ctx *ctx, const double *d_x, double *d_y) {
  tpsb_ctx *c = ctx;
  if (c->generic) return run_mult_generic(ctx, d_x, d_y);
  if (c->fast) return run_mult_fast(ctx, d_x, d_y);
  CU(cudaSetDevice(c->device));
  KernelArgs a = make_args(c, d_x, d_y);
  CU(cudaMemsetAsync(c->d_maxBits, 0, sizeof(unsigned long long), c->stream));
  int rc = run_gradients(c, a, false);
  if (rc) return rc;
  if (c->NEH > 0) {
    rc = exchange(c, c->d_gradUp, NEQ * DIM, c->d_sendG, c->d_gradUpHalo, c->ev_recvG);
    if (rc) return rc;
    DISPATCH(c, face(c, a, 0, c->NFlocal));
    CU(cudaStreamWaitEvent(c->stream, c->ev_recvG, 0));
    DISPATCH(c, face(c, a, c->NFlocal, c->NFint - c->NFlocal));
  } else {
    DISPATCH(c, face(c, a, 0, c->NFint));
  }
  bdr_faces(c, a);
  DISPATCH(c, resid(c, a));
  CU(cudaGetLastError());
  return TPSB_OK;
}

static __global__ void bits_to_double_kernel(const unsigned long long *bits, double *out) {
  *out = __longlong_as_double(static_cast<long long>(*bits));
}

static void ode_graph_invalidate(tpsb_ctx *ctx) {
  if (ctx->ode_exec) cudaGraphExecDestroy(ctx->ode_exec);
  ctx->ode_exec = nullptr;
}

static int ensure_work(tpsb_ctx *ctx, double **p) {
  if (*p) return TPSB_OK;
  CU(cudaMalloc(p, static_cast<size_t>(ctx->N) * ctx->neq * sizeof(double)));
  return TPSB_OK;
}

// tpsb_rhs_mult_host on a single rank (3-D dry-air paths, with or without boundary faces): the evaluation is PCIe-bound (2 x 40 B per node against
// ~0.23 ns of kernel time per node), so the three legs are overlapped chunk by chunk -- copies in on s_in, kernels on
// the context stream in the order of ctx->pipe_ops, copies out on s_out (PCIe is full duplex).  Same kernels, same
// per-element arithmetic: the result is bit-identical to tpsb_rhs_mult.
static int run_mult_host_pipelined(tpsb_ctx *ctx, const double *h_x, double *h_y) {
  tpsb_ctx *c = ctx;
  const int C = c->pipe_chunks;
  if (!c->s_in) {
    CU(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    c->ev_in.assign(C, nullptr);
    c->ev_out.assign(C, nullptr);
    for (int k = 0; k < C; k++) {
      CU(cudaEventCreateWithFlags(&c->ev_in[k], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&c->ev_out[k], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&c->ev_pipe0, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_pipe1, cudaEventDisableTiming));
  }
  KernelArgs a = make_args(c, c->d_hx, c->d_hy);
  CU(cudaMemsetAsync(c->d_maxBits, 0, sizeof(unsigned long long), c->stream));
  static const bool dbg = getenv("TPSB_PIPE_DEBUG") != nullptr;  // development: where the three legs end
  static cudaEvent_t dbg_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  static cudaEvent_t dbg_r[64] = {}, dbg_o[64] = {}, dbg_i[64] = {};
  if (dbg) {
    for (auto &e : dbg_ev)
      if (!e) cudaEventCreate(&e);
    CU(cudaEventRecord(dbg_ev[0], c->stream));
  }
  CU(cudaEventRecord(c->ev_pipe0, c->stream));  // earlier work on the caller's stream may still use d_hx / d_hy
  CU(cudaStreamWaitEvent(c->s_in, c->ev_pipe0, 0));
  CU(cudaStreamWaitEvent(c->s_out, c->ev_pipe0, 0));
  const size_t pitch = static_cast<size_t>(c->N) * sizeof(double);  // byNODES: one row per equation
  static const bool row_copies = !(getenv("TPSB_HOST_COPY2D") && atoi(getenv("TPSB_HOST_COPY2D")) != 0);
  for (int k = 0; k < C; k++) {
    const size_t off = static_cast<size_t>(c->pipe_eb[k]) * c->nd;
    const size_t width = static_cast<size_t>(c->pipe_eb[k + 1] - c->pipe_eb[k]) * c->nd * sizeof(double);
    if (row_copies) {  // one contiguous copy per equation row (TPSB_HOST_COPY2D=1 selects the pitched form)
      for (int eq = 0; eq < NEQ; eq++)
        CU(cudaMemcpyAsync(c->d_hx + off + static_cast<size_t>(eq) * c->N, h_x + off + static_cast<size_t>(eq) * c->N, width,
                           cudaMemcpyHostToDevice, c->s_in));
    } else {
      CU(cudaMemcpy2DAsync(c->d_hx + off, pitch, h_x + off, pitch, width, NEQ, cudaMemcpyHostToDevice, c->s_in));
    }
    CU(cudaEventRecord(c->ev_in[k], c->s_in));
    if (dbg && k < 64) {
      if (!dbg_i[k]) cudaEventCreate(&dbg_i[k]);
      CU(cudaEventRecord(dbg_i[k], c->s_in));
    }
  }
  for (const tpsb_ctx::PipeOp &op : c->pipe_ops) {
    const int k = op.chunk, e0 = c->pipe_eb[k], ne = c->pipe_eb[k + 1] - e0;
    switch (op.kind) {
      case 0: {
        CU(cudaStreamWaitEvent(c->stream, c->ev_in[k], 0));
        const long long cnt = static_cast<long long>(ne) * c->nd;
        if (c->fused) break;  // primitives are rebuilt inside the element kernel
        ProfScope ps(c, K_PRIM);
        prim_range_kernel<<<static_cast<unsigned>((cnt + 255) / 256), 256, 0, c->stream>>>(a, static_cast<long long>(e0) * c->nd, cnt);
        break;
      }
      case 1:
        if (c->fused) elem_fused(c, a, e0, ne, nullptr, FUSED_WRITE);
        else if (c->fast) grad_trace(c, a, e0, ne, nullptr);
        else grad(c, a, e0, ne, nullptr);
        break;
      case 2:
        if (c->fast) face_fast(c, a, c->pipe_fb[k], c->pipe_fb[k + 1] - c->pipe_fb[k]);
        else face(c, a, c->pipe_fb[k], c->pipe_fb[k + 1] - c->pipe_fb[k]);
        bdr_faces(c, a, c->pipe_bb[k], c->pipe_bb[k + 1] - c->pipe_bb[k]);  // BCintegrator faces of this chunk's elements
        break;
      default: {
        if (c->fused) lift(c, a, e0, ne);
        else resid(c, a, e0, ne);
        CU(cudaEventRecord(c->ev_out[k], c->stream));
        if (dbg && k < 64) {
          if (!dbg_r[k]) cudaEventCreate(&dbg_r[k]), cudaEventCreate(&dbg_o[k]);
          CU(cudaEventRecord(dbg_r[k], c->stream));
        }
        CU(cudaStreamWaitEvent(c->s_out, c->ev_out[k], 0));
        const size_t off = static_cast<size_t>(e0) * c->nd;
        const size_t wout = static_cast<size_t>(ne) * c->nd * sizeof(double);
        if (row_copies) {
          for (int eq = 0; eq < NEQ; eq++)
            CU(cudaMemcpyAsync(h_y + off + static_cast<size_t>(eq) * c->N, c->d_hy + off + static_cast<size_t>(eq) * c->N, wout,
                               cudaMemcpyDeviceToHost, c->s_out));
        } else {
          CU(cudaMemcpy2DAsync(h_y + off, pitch, c->d_hy + off, pitch, wout, NEQ, cudaMemcpyDeviceToHost, c->s_out));
        }
        if (dbg && k < 64) CU(cudaEventRecord(dbg_o[k], c->s_out));
        break;
      }
    }
  }
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->ev_pipe1, c->s_out));
  CU(cudaStreamWaitEvent(c->stream, c->ev_pipe1, 0));
  if (dbg) {
    CU(cudaEventRecord(dbg_ev[1], c->s_in));
    CU(cudaEventRecord(dbg_ev[2], c->s_out));
    CU(cudaEventRecord(dbg_ev[3], c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  if (dbg) {
    float t_in = 0, t_out = 0, t_all = 0;
    cudaEventSynchronize(dbg_ev[1]);
    cudaEventElapsedTime(&t_in, dbg_ev[0], dbg_ev[1]);
    cudaEventElapsedTime(&t_out, dbg_ev[0], dbg_ev[2]);
    cudaEventElapsedTime(&t_all, dbg_ev[0], dbg_ev[3]);
    fprintf(stderr, "[tpsb pipe] copy-in done %.2f ms, copy-out done %.2f ms, all %.2f ms after the start\n", t_in, t_out, t_all);
    for (int k = 0; k < C && k < 64; k++) {
      float tr = 0, to = 0;
      cudaEventElapsedTime(&tr, dbg_ev[0], dbg_r[k]);
      cudaEventElapsedTime(&to, dbg_ev[0], dbg_o[k]);
      float ti = 0;
      cudaEventElapsedTime(&ti, dbg_ev[0], dbg_i[k]);
      fprintf(stderr, "[tpsb pipe]   chunk %2d: copied in %.2f ms, residual done %.2f ms, copied out %.2f ms\n", k, ti, tr, to);
    }
  }
  return TPSB_OK;
}

extern "C" {

int tpsb_rhs_mult(tpsb_ctx *ctx, const double *d_x, double *d_y) {
  if (!ctx || !d_x || !d_y) return TPSB_EINVAL;
  return run_mult(ctx, d_x, d_y);
}

int tpsb_rhs_mult_host(tpsb_ctx *ctx, const double *h_x, double *h_y) {
  if (!ctx || !h_x || !h_y) return TPSB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  int rc = ensure_work(ctx, &ctx->d_hx);
  if (!rc) rc = ensure_work(ctx, &ctx->d_hy);
  if (rc) return rc;
  if (ctx->pipe_chunks > 0 && ctx->forcings.empty()) return run_mult_host_pipelined(ctx, h_x, h_y);
  const size_t nb = static_cast<size_t>(ctx->N) * ctx->neq * sizeof(double);
  CU(cudaMemcpyAsync(ctx->d_hx, h_x, nb, cudaMemcpyHostToDevice, ctx->stream));
  rc = run_mult(ctx, ctx->d_hx, ctx->d_hy);
  if (rc) return rc;
  CU(cudaMemcpyAsync(h_y, ctx->d_hy, nb, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return TPSB_OK;
}

int tpsb_update_primitives(tpsb_ctx *ctx, const double *d_x) {
  if (!ctx || !d_x) return TPSB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  if (ctx->generic) {
    GenArgs g = ctx->gen;
    g.U = d_x;
    ProfScope ps(ctx, K_PRIM);
    gen_prim_kernel<<<static_cast<unsigned>((g.N + 255) / 256), 256, 0, ctx->stream>>>(g);
    CU(cudaGetLastError());
    return TPSB_OK;
  }
  KernelArgs a = make_args(ctx, d_x, nullptr);
  launch_prim(ctx, a, 0);
  CU(cudaGetLastError());
  return TPSB_OK;
}

int tpsb_update_gradients(tpsb_ctx *ctx, const double *d_x, int primitives_updated) {
  if (!ctx || !d_x) return TPSB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  if (ctx->generic) return run_gradients_generic(ctx, d_x, primitives_updated != 0);
  KernelArgs a = make_args(ctx, d_x, nullptr);
  if (ctx->fused) return run_gradients_fused(ctx, a, primitives_updated != 0);
  if (ctx->fast) return run_gradients_fast(ctx, a, primitives_updated != 0);
  return run_gradients(ctx, a, primitives_updated != 0);
}

int tpsb_get_fields(tpsb_ctx *ctx, double **d_Up, double **d_gradUp) {
  if (!ctx) return TPSB_EINVAL;
  if (ctx->fused && ctx->fields_stale && ctx->fields_x) {
    // the fused evaluation keeps Up / gradUp on chip: rebuild them from the vector of the last tpsb_rhs_mult
    // (which must still be alive; call tpsb_update_gradients(x) instead to be explicit)
    CU(cudaSetDevice(ctx->device));
    KernelArgs a = make_args(ctx, ctx->fields_x, nullptr);
    const int rc = run_gradients_fused(ctx, a, false);
    if (rc) return rc;
  }
  if (d_Up) *d_Up = ctx->d_Up;
  if (d_gradUp) *d_gradUp = ctx->d_gradUp;
  return TPSB_OK;
}

int tpsb_debug_buffer(tpsb_ctx *ctx, int which, double **d_ptr, int64_t *count) {
  if (!ctx || !d_ptr || !count) return TPSB_EINVAL;
  const int64_t nf2 = static_cast<int64_t>(ctx->np) * ctx->np;
  if (which == 0) {
    *d_ptr = ctx->d_faceRes;
    *count = static_cast<int64_t>(ctx->NFint) * NEQ * nf2;
  } else if (which == 1) {
    *d_ptr = ctx->d_tr;
    *count = ctx->d_tr ? (static_cast<int64_t>(6) * ctx->NE + (ctx->NFint - ctx->NFlocal)) * NTF * nf2 : 0;
  } else {
    return TPSB_EINVAL;
  }
  return TPSB_OK;
}

int tpsb_debug_point_eval(tpsb_ctx *ctx, int which, int n, const double *d_U, const double *d_aux, double *d_out) {
  if (!ctx || !d_U || !d_out || n < 0 || which < 0 || which > 4) return TPSB_EINVAL;
  if (!ctx->generic) return fail(ctx, TPSB_ENOTIMPL, "point evaluation is a test hook of the generic path");
  CU(cudaSetDevice(ctx->device));
  gen_point_eval_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->gen, which, n, d_U, d_aux, d_out);
  ctx->launches++;
  CU(cudaGetLastError());
  return TPSB_OK;
}

int tpsb_set_solution_view(tpsb_ctx *ctx, const double *d_U) {
  if (!ctx) return TPSB_EINVAL;
  ctx->sol_view = d_U;
  return TPSB_OK;
}

int tpsb_set_distance_field(tpsb_ctx *ctx, const double *d_distance) {
  if (!ctx) return TPSB_EINVAL;
  if (!ctx->generic) return fail(ctx, TPSB_EINVAL, "the wall distance is read by the mixing-length model (generic path) only");
  ctx->gen.dist = d_distance;
  ode_graph_invalidate(ctx);  // the pointer is baked into the captured launches
  if (ctx->NEH > 0 && d_distance) {  // the neighbours' side of a shared face interpolates ITS distance (face_integrator.cpp:304-309)
    CU(cudaSetDevice(ctx->device));
    double *dh = nullptr, *sb = nullptr;
    CU(cudaMalloc(&dh, static_cast<size_t>(ctx->NEH) * ctx->nd * sizeof(double)));
    CU(cudaMalloc(&sb, static_cast<size_t>(std::max(ctx->n_send, 1)) * ctx->nd * sizeof(double)));
    ctx->gen_allocs.push_back(dh);
    const int rc = exchange(ctx, d_distance, 1, sb, dh, ctx->ev_recvU);
    if (rc) return rc;
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_recvU, 0));
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(sb);
    ctx->gen.distHalo = dh;
  }
  return TPSB_OK;
}

int tpsb_averaging_add_sample(tpsb_ctx *ctx, const double *d_inst, int num_fields, double *d_mean, double *d_vari, int vari_start,
                              int vari_components, int ns_mean, int ns_vari, int pressure_slot) {
  if (!ctx || !d_mean || ns_mean < 0 || ns_vari < 0) return TPSB_EINVAL;
  tpsb_ctx *c = ctx;
  if (num_fields < 1 || num_fields > GEN_MAXEQ + 4) return fail(ctx, TPSB_EINVAL, "1..%d fields per family", GEN_MAXEQ + 4);
  if (d_vari && (vari_start < 0 || vari_components < 1 || vari_start + vari_components > num_fields))
    return fail(ctx, TPSB_EINVAL, "variance components out of range");
  CU(cudaSetDevice(c->device));
  if (!d_inst) {  // the primitive state, as M2ulPhyS registers it (src/M2ulPhyS.cpp:634)
    double *up = nullptr;
    const int rc = tpsb_get_fields(ctx, &up, nullptr);
    if (rc) return rc;
    d_inst = up;
    if (num_fields != c->neq) return fail(ctx, TPSB_EINVAL, "the primitive family has %d fields", c->neq);
  }
  GenPhys ph;
  memset(&ph, 0, sizeof(ph));
  if (c->generic) {
    ph = c->gen.phys;
  } else {
    ph.dim = 3, ph.nvel = 3, ph.neq = NEQ, ph.fluid = 0, ph.dry = c->phys;
  }
  if (pressure_slot && num_fields != c->neq) return fail(ctx, TPSB_EINVAL, "the pressure slot needs the full primitive state");
  averaging_kernel<<<static_cast<unsigned>((c->N + 127) / 128), 128, 0, c->stream>>>(
      c->N, num_fields, c->generic ? c->gen.dim : 3, d_inst, d_mean, d_vari, vari_start, vari_components,
      static_cast<double>(ns_mean), static_cast<double>(ns_vari), pressure_slot, ph);
  c->launches++;
  CU(cudaGetLastError());
  return TPSB_OK;
}

int tpsb_clear_forcings(tpsb_ctx *ctx) {
  if (!ctx) return TPSB_EINVAL;
  for (auto &f : ctx->forcings)
    if (f.mix) cudaFree(f.mix);
  ctx->forcings.clear();
  ctx->forcing_needs_grad = false;
  ode_graph_invalidate(ctx);
  return TPSB_OK;
}

int tpsb_add_forcing(tpsb_ctx *ctx, const tpsb_forcing_desc *d) {
  if (!ctx || !d) return TPSB_EINVAL;
  tpsb_ctx *c = ctx;
  if (c->forcings.size() >= MAX_FORCING) return fail(ctx, TPSB_EINVAL, "at most %d forcing terms", MAX_FORCING);
  CU(cudaSetDevice(c->device));
  const int dim = c->generic ? c->gen.dim : 3, nvel = c->generic ? c->gen.nvel : 3, neq = c->neq;
  const bool mixture = c->generic && c->gen.phys.fluid != 0;
  if (c->generic && c->gen.phys.lte && d->kind != TPSB_FORCING_HEAT_SOURCE && d->kind != TPSB_FORCING_JOULE_HEATING)
    return fail(ctx, TPSB_ENOTIMPL, "forcing kind %d with the LTE fluid is not built (its dry-air formulas read gamma and R)", d->kind);
  ForcingDev f;
  memset(&f, 0, sizeof(f));
  f.kind = d->kind;
  switch (d->kind) {
    case TPSB_FORCING_PRESSURE_GRADIENT:
      for (int k = 0; k < 3; k++) f.g[k] = d->pressure_grad[k];
      if ((c->generic ? c->gen.eq_system : c->phys.eq_system) == 0)
        return fail(ctx, TPSB_EINVAL, "the pressure-gradient forcing reads gradUp: Navier-Stokes runs only");
      c->forcing_needs_grad = true;
      break;
    case TPSB_FORCING_HEAT_SOURCE: {
      double len = 0;
      for (int k = 0; k < dim; k++) len += (d->hs_point2[k] - d->hs_point1[k]) * (d->hs_point2[k] - d->hs_point1[k]);
      len = std::sqrt(len);
      if (!(len > 0) || !(d->hs_radius > 0)) return fail(ctx, TPSB_EINVAL, "heat source: degenerate cylinder");
      for (int k = 0; k < dim; k++) f.p1[k] = d->hs_point1[k], f.axis[k] = (d->hs_point2[k] - d->hs_point1[k]) / len;
      f.len = len, f.radius = d->hs_radius, f.value = d->hs_value;
      break;
    }
    case TPSB_FORCING_JOULE_HEATING:
      if (!d->joule_heating) return fail(ctx, TPSB_EINVAL, "Joule heating needs its nodal field");
      if (nvel != 3) return fail(ctx, TPSB_EINVAL, "JouleHeating asserts nvel == 3 (src/forcing_terms.cpp:444)");
      f.field = d->joule_heating;
      break;
    case TPSB_FORCING_SPONGE_ZONE: {
      if (mixture) return fail(ctx, TPSB_ENOTIMPL, "sponge zones are built for dry air");
      double nm = 0;
      for (int k = 0; k < dim; k++) nm += d->sz_normal[k] * d->sz_normal[k];
      nm = std::sqrt(nm);
      if (!(nm > 0)) return fail(ctx, TPSB_EINVAL, "sponge zone: zero normal");
      for (int k = 0; k < dim; k++) f.n[k] = d->sz_normal[k] / nm, f.p0[k] = d->sz_point0[k], f.pi[k] = d->sz_point_init[k];
      f.sz_type = d->sz_type, f.sz_mixed = d->sz_mixed_out ? 1 : 0;
      if (f.sz_type == 1 && dim != 3) return fail(ctx, TPSB_EINVAL, "annular sponge zones are three-dimensional");
      f.r1 = d->sz_r1, f.r2 = d->sz_r2, f.tol = d->sz_tol, f.mult = d->sz_mult;
      if (!f.sz_mixed) {  // SpongeZone::SpongeZone, USERDEF (src/forcing_terms.cpp:486-520) with DryAir::modifyEnergyForPressure
        const PhysParams &ph = c->generic ? c->gen.phys.dry : c->phys;
        const double rho = d->sz_target[0], p = d->sz_target[4];
        double ke = 0;
        f.targetU[0] = rho;
        for (int k = 0; k < nvel; k++) f.targetU[1 + k] = rho * d->sz_target[1 + k], ke += f.targetU[1 + k] * f.targetU[1 + k];
        ke *= 0.5 / rho;
        f.targetU[1 + nvel] = p / ph.gm1 + ke;
        f.sound = std::sqrt(ph.gamma * p / rho);  // sqrt(gamma R T) of the target
      } else {
        CU(cudaMalloc(&f.mix, sizeof(double) * (2 * neq + 2)));
      }
      break;
    }
    default:
      return fail(ctx, TPSB_ENOTIMPL, "forcing kind %d not built", d->kind);
  }
  if (!c->generic && !c->d_xiN3) {  // reference coordinates of the GL nodes, x fastest
    std::vector<double> xi(static_cast<size_t>(c->nd) * 3);
    for (int n = 0; n < c->nd; n++) {
      xi[3 * n + 0] = c->T.xn[n % c->np];
      xi[3 * n + 1] = c->T.xn[(n / c->np) % c->np];
      xi[3 * n + 2] = c->T.xn[n / (c->np * c->np)];
    }
    CU(upload(&c->d_xiN3, xi));
  }
  c->forcings.push_back(f);
  ode_graph_invalidate(ctx);
  return TPSB_OK;
}

int tpsb_set_time_step(tpsb_ctx *ctx, double dt) {
  if (!ctx) return TPSB_EINVAL;
  if (!(dt >= 0.0) || !std::isfinite(dt)) return fail(ctx, TPSB_EINVAL, "time step must be finite and >= 0");
  if (ctx->bc_dt != dt) ode_graph_invalidate(ctx);
  ctx->bc_dt = dt;
  return TPSB_OK;
}

int tpsb_get_bc_state(tpsb_ctx *ctx, int attr, double *mean_up, double *boundary_u, int capacity, int *num_points) {
  if (!ctx || !num_points) return TPSB_EINVAL;
  tpsb_ctx *c = ctx;
  for (size_t i = 0; i < c->nr_patches.size(); i++)
    if (c->nr_attr[i] == attr) {
      const GenNrPatch &pt = c->nr_patches[i];
      const int npts = pt.nfaces * c->gen.nqf, neq = c->gen.neq;
      *num_points = npts;
      CU(cudaSetDevice(c->device));
      CU(cudaStreamSynchronize(c->stream));
      if (mean_up) CU(cudaMemcpy(mean_up, pt.meanUp, sizeof(double) * neq, cudaMemcpyDeviceToHost));
      if (boundary_u && capacity > 0 && npts > 0)
        CU(cudaMemcpy(boundary_u, pt.boundaryU, sizeof(double) * neq * std::min(npts, capacity), cudaMemcpyDeviceToHost));
      return TPSB_OK;
    }
  return fail(ctx, TPSB_EINVAL, "no non-reflecting boundary condition on attribute %d", attr);
}

int tpsb_set_reaction_rate_field(tpsb_ctx *ctx, const double *d_rates, int num_components) {
  if (!ctx) return TPSB_EINVAL;
  if (!ctx->generic || !ctx->gen.phys.fluid) return fail(ctx, TPSB_EINVAL, "reaction rate fields belong to a plasma mixture");
  for (int r = 0; r < ctx->mix_host.numReactions; r++)
    if (ctx->mix_host.rxModel[r] == 3 && d_rates && (ctx->mix_host.rxComp[r] < 0 || ctx->mix_host.rxComp[r] >= num_components))
      return fail(ctx, TPSB_EINVAL, "reaction %d reads component %d of a %d-component rate field", r, ctx->mix_host.rxComp[r], num_components);
  CU(cudaSetDevice(ctx->device));
  ctx->mix_host.rateField = d_rates;
  CU(cudaMemcpyAsync(const_cast<MixParams *>(ctx->gen.phys.mix), &ctx->mix_host, sizeof(MixParams), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return TPSB_OK;
}

static __global__ void mean_abs_kernel(long long N, int neq, const double *y, double *out) {
  __shared__ double sm[256];
  const int eq = blockIdx.y;
  double s = 0;
  for (long long n = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; n < N; n += static_cast<long long>(gridDim.x) * blockDim.x)
    s += fabs(y[n + eq * N]);
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(&out[eq], sm[0] / static_cast<double>(N));
}

int tpsb_get_mean_time_derivatives(tpsb_ctx *ctx, const double *d_y, double *out) {
  if (!ctx || !d_y || !out) return TPSB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  const int neq = tpsb_num_equation(ctx);
  const long long N = tpsb_num_dofs(ctx);
  double *d_out = nullptr;
  CU(cudaMalloc(&d_out, neq * sizeof(double)));
  CU(cudaMemsetAsync(d_out, 0, neq * sizeof(double), ctx->stream));
  const int nb = static_cast<int>(std::min<long long>((N + 255) / 256, 1024));
  mean_abs_kernel<<<dim3(nb, neq), 256, 0, ctx->stream>>>(N, neq, d_y, d_out);
  cudaError_t e = cudaMemcpyAsync(out, d_out, neq * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_out);
  CU(e);
  return TPSB_OK;
}

int tpsb_get_max_char_speed(tpsb_ctx *ctx, double *out) {
  if (!ctx || !out) return TPSB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  bits_to_double_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_maxBits, ctx->d_mcs);
  ctx->launches++;
  if (ctx->comm)  // MPI_Allreduce(MAX) of src/rhs_operator.cpp:557-558
    NC(ncclAllReduce(ctx->d_mcs, ctx->d_mcs, 1, ncclDouble, ncclMax, ctx->comm, ctx->stream));
  CU(cudaMemcpyAsync(out, ctx->d_mcs, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return TPSB_OK;
}

// MFEM ODESolver::Step restated (third party): ForwardEuler, RK2(a=1) (Heun),
// RK3SSP, RK4 -- the solvers M2ulPhyS::initVariables can select (src/M2ulPhyS.cpp:721-739).
// One step, enqueued on ctx->stream.
static int ode_one_step(tpsb_ctx *ctx, double *x, double dt, int scheme) {
  int rc;
  const long long n = ctx->N * ctx->neq;
  const unsigned nb = static_cast<unsigned>((n + 255) / 256);
  double *k = ctx->d_k, *y = ctx->d_yv, *z = ctx->d_z;
  cudaStream_t st = ctx->stream;
  // One stage: k = f(in), then Y = X + A k and (Z given) Z = (ACC ? Z : X) + B k.  On the 3-D dry-air paths the update
  // is the epilogue of the residual kernel (no k round trip, no axpy sweep: 240 -> 120 B per node and stage less);
  // the generic path runs its forcing-term kernels after the residual, so it keeps the separate sweep.
  static const bool no_fuse = getenv("TPSB_ODE_FUSE") && atoi(getenv("TPSB_ODE_FUSE")) == 0;
  const bool fuse = !ctx->generic && !no_fuse && ctx->forcings.empty();  // forcing terms act on k before the update
  auto stage = [&](const double *in, const double *X, double A, double *Y, double B, double *Z, int ACC) -> int {
    if (fuse) {
      ctx->rk = {X, Y, Z, A, B, ACC};
      const int r = run_mult(ctx, in, k);
      ctx->rk = {nullptr, nullptr, nullptr, 0.0, 0.0, 0};
      return r;
    }
    const int r = run_mult(ctx, in, k);
    if (r) return r;
    axpy2_kernel<<<nb, 256, 0, st>>>(n, X, k, A, Y, B, Z, ACC);
    ctx->launches++;
    return TPSB_OK;
  };
  if (scheme == 1) {  // x += dt f(x)
    if ((rc = stage(x, x, dt, x, 0.0, nullptr, 0))) return rc;
  } else if (scheme == 2) {  // RK2Solver(a = 1): y = x + dt k1; x += dt/2 (k1 + k2)
    if ((rc = stage(x, x, dt, y, 0.5 * dt, z, 0))) return rc;
    if ((rc = stage(y, z, 0.5 * dt, x, 0.0, nullptr, 0))) return rc;
  } else if (scheme == 3) {  // RK3SSPSolver
    if ((rc = stage(x, x, dt, y, 0.0, nullptr, 0))) return rc;  // y = x + dt k
    // y = 3/4 x + 1/4 (y + dt k)
    if ((rc = stage(y, y, dt, y, 0.0, nullptr, 0))) return rc;
    {
      ProfScope ps(ctx, K_AXPY);
      rk3_combine_kernel<<<nb, 256, 0, st>>>(n, x, y, 0.75, 0.25, y);
    }
    // x = 1/3 x + 2/3 (y + dt k)
    if ((rc = stage(y, y, dt, y, 0.0, nullptr, 0))) return rc;
    {
      ProfScope ps(ctx, K_AXPY);
      rk3_combine_kernel<<<nb, 256, 0, st>>>(n, x, y, 1.0 / 3.0, 2.0 / 3.0, x);
    }
  } else {  // RK4Solver
    if ((rc = stage(x, x, dt / 2, y, dt / 6, z, 0))) return rc;
    if ((rc = stage(y, x, dt / 2, y, dt / 3, z, 1))) return rc;
    if ((rc = stage(y, x, dt, y, dt / 3, z, 1))) return rc;
    if ((rc = stage(y, z, dt / 6, x, 0.0, nullptr, 0))) return rc;
  }
  return TPSB_OK;
}

// The first step runs eagerly on the caller's stream (lazy allocations, function attributes, constant tables); the
// remaining ones replay one captured step: a Runge-Kutta step is 20-30 short launches, and on the small 2-D / plasma
// configurations their launch latency, not the kernels, sets the step time.  Single rank only
// (the halo exchange keeps its own stream and events); TPSB_ODE_GRAPH=0 switches the replay off.
int tpsb_ode_step(tpsb_ctx *ctx, double *d_U, double dt, int scheme, int nsteps) {
  if (!ctx || !d_U || nsteps < 0) return TPSB_EINVAL;
  if (scheme < 1 || scheme > 4) return fail(ctx, TPSB_ENOTIMPL, "ODE scheme %d not built", scheme);
  CU(cudaSetDevice(ctx->device));
  int rc = ensure_work(ctx, &ctx->d_k);
  if (!rc) rc = ensure_work(ctx, &ctx->d_yv);
  if (!rc) rc = ensure_work(ctx, &ctx->d_z);
  if (rc) return rc;
  const double *saved_view = ctx->sol_view;
  ctx->sol_view = d_U;  // the forcing terms read the solution vector, not the stage vector (parity trap 1)
  ctx->bc_dt = dt;      // BoundaryCondition::dt is a reference to M2ulPhyS::dt, the step being taken
  const char *env = getenv("TPSB_ODE_GRAPH");
  const bool use_graph = !ctx->comm && !ctx->profiling && nsteps >= 3 && !(env && atoi(env) == 0);
  int done = 0;
  for (; done < nsteps && (!use_graph || done < 1); done++)
    if ((rc = ode_one_step(ctx, d_U, dt, scheme))) break;
  if (!rc && use_graph && done < nsteps) {
    cudaStream_t caller = ctx->stream;
    if (!ctx->ode_stream) {
      CU(cudaStreamCreateWithFlags(&ctx->ode_stream, cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&ctx->ode_ev0, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&ctx->ode_ev1, cudaEventDisableTiming));
    }
    if (ctx->ode_exec && (ctx->ode_U != d_U || ctx->ode_dt != dt || ctx->ode_scheme != scheme)) ode_graph_invalidate(ctx);
    if (!ctx->ode_exec) {
      const long long l0 = ctx->launches;
      cudaGraph_t graph = nullptr;
      CU(cudaStreamBeginCapture(ctx->ode_stream, cudaStreamCaptureModeThreadLocal));
      ctx->stream = ctx->ode_stream;
      rc = ode_one_step(ctx, d_U, dt, scheme);
      ctx->stream = caller;
      const cudaError_t ce = cudaStreamEndCapture(ctx->ode_stream, &graph);
      ctx->ode_launches = ctx->launches - l0;
      ctx->launches = l0;  // nothing ran yet
      if (rc || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        ctx->sol_view = saved_view;
        return rc ? rc : fail(ctx, TPSB_ECUDA, "capturing the Runge-Kutta step failed: %s", cudaGetErrorString(ce));
      }
      const cudaError_t ie = cudaGraphInstantiate(&ctx->ode_exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ie != cudaSuccess) {
        ctx->ode_exec = nullptr;
        ctx->sol_view = saved_view;
        return fail(ctx, TPSB_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
      }
      ctx->ode_U = d_U, ctx->ode_dt = dt, ctx->ode_scheme = scheme;
    }
    CU(cudaEventRecord(ctx->ode_ev0, caller));
    CU(cudaStreamWaitEvent(ctx->ode_stream, ctx->ode_ev0, 0));
    for (; done < nsteps; done++) {
      CU(cudaGraphLaunch(ctx->ode_exec, ctx->ode_stream));
      ctx->launches += ctx->ode_launches;
    }
    CU(cudaEventRecord(ctx->ode_ev1, ctx->ode_stream));
    CU(cudaStreamWaitEvent(caller, ctx->ode_ev1, 0));
  }
  ctx->sol_view = saved_view;
  if (rc) return rc;
  CU(cudaGetLastError());
  return TPSB_OK;
}

// M2ulPhyS::Check_NAN (src/M2ulPhyS.cpp:2463-2519: count the NaNs of the solution) and Check_Undershoot (:2526-2549:
// active species densities clamped at zero, mixtures only) in one sweep over the solution vector
static __global__ void check_state_kernel(long long N, int neq, int clamp_begin, int clamp_end, double *U, int *nan_count) {
  const long long n = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  int bad = 0;
  if (n < N) {
    for (int eq = 0; eq < neq; eq++) {
      const double v = U[n + eq * N];
      if (v != v) bad++;
      if (eq >= clamp_begin && eq < clamp_end && v < 0.0) U[n + eq * N] = 0.0;  // max(v, 0): a NaN stays a NaN
    }
  }
  bad = __reduce_add_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(nan_count, bad);
}

int tpsb_get_hmin(tpsb_ctx *ctx, double *hmin) {
  if (!ctx || !hmin) return TPSB_EINVAL;
  *hmin = ctx->hmin;
  if (ctx->comm) {  // MPI_Allreduce(MIN) of src/M2ulPhyS.cpp:761
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(ctx->d_mcs, hmin, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NC(ncclAllReduce(ctx->d_mcs, ctx->d_mcs, 1, ncclDouble, ncclMin, ctx->comm, ctx->stream));
    CU(cudaMemcpyAsync(hmin, ctx->d_mcs, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return TPSB_OK;
}

// tpsb_check_state: Check_NAN + Check_Undershoot of one solution vector.
// M2ulPhyS::solveStep (src/M2ulPhyS.cpp:2004-2016) without the I/O: one ODESolver::Step, Check_NAN, Check_Undershoot
// for mixtures, and -- when cfl > 0 -- the next adaptive time step CFL hmin / max_char_speed / dim with the
// characteristic speed of the step's last stage, reduced over the ranks.  cfl <= 0: constant time step, *dt_next = dt.
int tpsb_check_state(tpsb_ctx *ctx, double *d_U, int *num_nan) {
  if (!ctx || !d_U) return TPSB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  if (!ctx->d_nan_count) CU(cudaMalloc(&ctx->d_nan_count, sizeof(int)));
  CU(cudaMemsetAsync(ctx->d_nan_count, 0, sizeof(int), ctx->stream));
  int cb = 0, ce = 0;
  if (ctx->generic && ctx->gen.phys.fluid) {
    cb = ctx->gen.nvel + 2;
    ce = cb + (ctx->mix_host.ambipolar ? ctx->mix_host.numSpecies - 2 : ctx->mix_host.numSpecies - 1);
  }
  check_state_kernel<<<static_cast<unsigned>((ctx->N + 255) / 256), 256, 0, ctx->stream>>>(ctx->N, ctx->neq, cb, ce, d_U, ctx->d_nan_count);
  ctx->launches++;
  if (ctx->comm) NC(ncclAllReduce(ctx->d_nan_count, ctx->d_nan_count, 1, ncclInt, ncclSum, ctx->comm, ctx->stream));
  int count = 0;
  CU(cudaMemcpyAsync(&count, ctx->d_nan_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (num_nan) *num_nan = count;
  return TPSB_OK;
}

int tpsb_solve_step(tpsb_ctx *ctx, double *d_U, double dt, int scheme, double cfl, int *num_nan, double *dt_next) {
  if (!ctx || !d_U) return TPSB_EINVAL;
  int rc = tpsb_ode_step(ctx, d_U, dt, scheme, 1);
  if (rc) return rc;
  if ((rc = tpsb_check_state(ctx, d_U, num_nan))) return rc;
  if (dt_next) {
    *dt_next = dt;
    if (cfl > 0.0) {
      double mcs = 0.0, hmin = 0.0;
      if ((rc = tpsb_get_max_char_speed(ctx, &mcs))) return rc;
      if ((rc = tpsb_get_hmin(ctx, &hmin))) return rc;
      // a quiescent (mcs == 0) or NaN state has no CFL limit to offer: keep the current step and say so
      if (!(mcs > 0.0) || !std::isfinite(mcs)) return fail(ctx, TPSB_EINVAL, "adaptive time step: max characteristic speed is %g", mcs);
      *dt_next = cfl * hmin / mcs / static_cast<double>(ctx->dim);
    }
  }
  return TPSB_OK;
}

int tpsb_set_profiling(tpsb_ctx *ctx, int on) {
  if (!ctx) return TPSB_EINVAL;
  ctx->profiling = on != 0;
  return TPSB_OK;
}

int tpsb_get_kernel_times(tpsb_ctx *ctx, double ms[TPSB_NUM_KERNEL_CLASSES], int64_t count[TPSB_NUM_KERNEL_CLASSES]) {
  if (!ctx || !ms || !count) return TPSB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < TPSB_NUM_KERNEL_CLASSES; k++) {
    for (auto &ev : ctx->prof_events[k]) {
      float t = 0;
      CU(cudaEventElapsedTime(&t, ev.first, ev.second));
      ctx->prof_ms[k] += t;
      ctx->prof_count[k]++;
      cudaEventDestroy(ev.first);
      cudaEventDestroy(ev.second);
    }
    ctx->prof_events[k].clear();
    ms[k] = ctx->prof_ms[k];
    count[k] = ctx->prof_count[k];
    ctx->prof_ms[k] = 0;
    ctx->prof_count[k] = 0;
  }
  return TPSB_OK;
}

int tpsb_comm_get_unique_id(unsigned char unique_id[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return TPSB_ENCCL;
  memcpy(unique_id, &id, sizeof(id));
  return TPSB_OK;
}

int tpsb_comm_init_rank(const unsigned char unique_id[128], int nranks, int rank, int device, void **nccl_comm) {
  if (!unique_id || !nccl_comm) return TPSB_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return TPSB_ECUDA;
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclComm_t comm;
  if (ncclCommInitRank(&comm, nranks, id, rank) != ncclSuccess) return TPSB_ENCCL;
  *nccl_comm = comm;
  return TPSB_OK;
}

int tpsb_comm_destroy(void *nccl_comm) {
  if (!nccl_comm) return TPSB_OK;
  return ncclCommDestroy(static_cast<ncclComm_t>(nccl_comm)) == ncclSuccess ? TPSB_OK : TPSB_ENCCL;
}

}  // extern "C"
