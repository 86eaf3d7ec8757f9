/* TEST INFRASTRUCTURE ONLY -- not part of the product, never shipped or linked into it.
 *
 * A toy stand-in for the slice of MFEM's public interface that the reference's
 * per-point physics translation units (src/equation_of_state.cpp, src/fluxes.cpp,
 * src/riemann_solver.cpp, src/transport_properties.cpp, src/gas_transport.cpp,
 * src/collision_integrals.cpp, src/chemistry.cpp, src/reaction.cpp, src/table.cpp,
 * src/radiation.cpp, src/mixing_length_transport.cpp) touch.  With it those files
 * compile IN PLACE from /root/reference, unmodified, into oracle/_ref/ and serve as
 * the bit-for-bit CPU oracle of the physics layer.
 * MFEM itself is absent from this environment; none of its code is reproduced here --
 * only the handful of container signatures the reference calls.
 */
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#define MFEM_HOST_DEVICE
#define MFEM_ASSERT(cond, msg) assert(cond)
#define MFEM_VERIFY(cond, msg) assert(cond)
#define MFEM_ABORT(msg) abort()
#define MFEM_FORALL(i, N, ...)            \
  for (int i = 0; i < (N); i++) {         \
    __VA_ARGS__                           \
  }

typedef int MPI_Request;
typedef int MPI_Status;
typedef int MPI_Comm;
#define MPI_COMM_WORLD 0
static inline int MPI_Barrier(MPI_Comm) { return 0; }

namespace mfem {

using std::max;
using std::min;

inline void mfem_error(const char *msg = NULL) {
  if (msg) fprintf(stderr, "mfem_error: %s\n", msg);
  abort();
}

struct Mpi {
  static bool Root() { return true; }
};

template <class T>
class Array {
  std::vector<T> d_;

 public:
  Array() {}
  explicit Array(int n) : d_(n) {}
  void SetSize(int n) { d_.resize(n); }
  void SetSize(int n, const T &v) { d_.assign(n, v); }
  int Size() const { return static_cast<int>(d_.size()); }
  T &operator[](int i) { return d_[i]; }
  const T &operator[](int i) const { return d_[i]; }
  Array &operator=(const T &v) {
    std::fill(d_.begin(), d_.end(), v);
    return *this;
  }
  T *GetData() { return d_.data(); }
  const T *GetData() const { return d_.data(); }
  const T *Read() const { return d_.data(); }
  T *Write() { return d_.data(); }
  T *ReadWrite() { return d_.data(); }
  const T *HostRead() const { return d_.data(); }
  T *HostWrite() { return d_.data(); }
  T *HostReadWrite() { return d_.data(); }
  void Append(const T &v) { d_.push_back(v); }
  void DeleteAll() { d_.clear(); }
};

class Vector {
  double *p_ = NULL;
  int n_ = 0;
  bool own_ = false;

 public:
  Vector() {}
  explicit Vector(int n) { SetSize(n); }
  Vector(double *d, int n) : p_(d), n_(n), own_(false) {}
  Vector(const Vector &o) {
    SetSize(o.n_);
    if (n_) memcpy(p_, o.p_, sizeof(double) * n_);
  }
  ~Vector() {
    if (own_) delete[] p_;
  }
  Vector &operator=(const Vector &o) {
    if (this == &o) return *this;
    if (n_ != o.n_) SetSize(o.n_);
    if (n_) memcpy(p_, o.p_, sizeof(double) * n_);
    return *this;
  }
  Vector &operator=(double v) {
    for (int i = 0; i < n_; i++) p_[i] = v;
    return *this;
  }
  Vector &operator*=(double v) {
    for (int i = 0; i < n_; i++) p_[i] *= v;
    return *this;
  }
  Vector &operator/=(double v) {
    for (int i = 0; i < n_; i++) p_[i] /= v;
    return *this;
  }
  Vector &operator+=(const Vector &o) {
    for (int i = 0; i < n_; i++) p_[i] += o.p_[i];
    return *this;
  }
  Vector &operator-=(const Vector &o) {
    for (int i = 0; i < n_; i++) p_[i] -= o.p_[i];
    return *this;
  }
  double operator*(const Vector &o) const {
    double s = 0;
    for (int i = 0; i < n_; i++) s += p_[i] * o.p_[i];
    return s;
  }
  void SetSize(int n) {
    if (own_ && n <= n_ && p_) {
      n_ = n;
      return;
    }
    if (own_) delete[] p_;
    p_ = n > 0 ? new double[n]() : NULL;
    n_ = n;
    own_ = true;
  }
  void UseDevice(bool) {}
  void SetDataAndSize(double *d, int n) {
    if (own_) delete[] p_;
    p_ = d;
    n_ = n;
    own_ = false;
  }
  void NewDataAndSize(double *d, int n) { SetDataAndSize(d, n); }
  int Size() const { return n_; }
  double &operator[](int i) { return p_[i]; }
  const double &operator[](int i) const { return p_[i]; }
  double &operator()(int i) { return p_[i]; }
  const double &operator()(int i) const { return p_[i]; }
  double &Elem(int i) { return p_[i]; }
  double *GetData() const { return p_; }
  const double *Read() const { return p_; }
  double *Write() { return p_; }
  double *ReadWrite() { return p_; }
  const double *HostRead() const { return p_; }
  double *HostWrite() { return p_; }
  double *HostReadWrite() { return p_; }
  Vector &Set(double a, const Vector &x) {
    for (int i = 0; i < n_; i++) p_[i] = a * x.p_[i];
    return *this;
  }
  Vector &Add(double a, const Vector &x) {
    for (int i = 0; i < n_; i++) p_[i] += a * x.p_[i];
    return *this;
  }
  double Norml2() const { return std::sqrt((*this) * (*this)); }
  double Sum() const {
    double s = 0;
    for (int i = 0; i < n_; i++) s += p_[i];
    return s;
  }
  double Max() const {
    double m = p_[0];
    for (int i = 1; i < n_; i++) m = std::max(m, p_[i]);
    return m;
  }
  void Print(std::ostream &os = std::cout, int = 8) const {
    for (int i = 0; i < n_; i++) os << p_[i] << " ";
    os << "\n";
  }
};

/* column-major dense matrix */
class DenseMatrix {
  double *p_ = NULL;
  int h_ = 0, w_ = 0;
  bool own_ = false;

 public:
  DenseMatrix() {}
  explicit DenseMatrix(int s) { SetSize(s, s); }
  DenseMatrix(int h, int w) { SetSize(h, w); }
  DenseMatrix(double *d, int h, int w) : p_(d), h_(h), w_(w), own_(false) {}
  DenseMatrix(const DenseMatrix &o) {
    SetSize(o.h_, o.w_);
    if (h_ * w_) memcpy(p_, o.p_, sizeof(double) * h_ * w_);
  }
  ~DenseMatrix() {
    if (own_) delete[] p_;
  }
  DenseMatrix &operator=(const DenseMatrix &o) {
    if (this == &o) return *this;
    if (h_ != o.h_ || w_ != o.w_) SetSize(o.h_, o.w_);
    if (h_ * w_) memcpy(p_, o.p_, sizeof(double) * h_ * w_);
    return *this;
  }
  DenseMatrix &operator=(double v) {
    for (int i = 0; i < h_ * w_; i++) p_[i] = v;
    return *this;
  }
  DenseMatrix &operator*=(double v) {
    for (int i = 0; i < h_ * w_; i++) p_[i] *= v;
    return *this;
  }
  DenseMatrix &operator+=(const DenseMatrix &o) {
    for (int i = 0; i < h_ * w_; i++) p_[i] += o.p_[i];
    return *this;
  }
  DenseMatrix &operator-=(const DenseMatrix &o) {
    for (int i = 0; i < h_ * w_; i++) p_[i] -= o.p_[i];
    return *this;
  }
  void SetSize(int s) { SetSize(s, s); }
  void SetSize(int h, int w) {
    if (own_) delete[] p_;
    p_ = (h * w > 0) ? new double[h * w]() : NULL;
    h_ = h;
    w_ = w;
    own_ = true;
  }
  void UseExternalData(double *d, int h, int w) {
    if (own_) delete[] p_;
    p_ = d;
    h_ = h;
    w_ = w;
    own_ = false;
  }
  int Height() const { return h_; }
  int Width() const { return w_; }
  int NumRows() const { return h_; }
  int NumCols() const { return w_; }
  double &operator()(int i, int j) { return p_[i + j * h_]; }
  const double &operator()(int i, int j) const { return p_[i + j * h_]; }
  double *GetData() const { return p_; }
  double *Data() const { return p_; }
  const double *Read() const { return p_; }
  double *Write() { return p_; }
  double *ReadWrite() { return p_; }
  const double *HostRead() const { return p_; }
  double *HostWrite() { return p_; }
  void Mult(const Vector &x, Vector &y) const {
    for (int i = 0; i < h_; i++) {
      double s = 0;
      for (int j = 0; j < w_; j++) s += (*this)(i, j) * x[j];
      y[i] = s;
    }
  }
  void Mult(const double *x, double *y) const {
    for (int i = 0; i < h_; i++) {
      double s = 0;
      for (int j = 0; j < w_; j++) s += (*this)(i, j) * x[j];
      y[i] = s;
    }
  }
  void MultTranspose(const Vector &x, Vector &y) const {
    for (int j = 0; j < w_; j++) {
      double s = 0;
      for (int i = 0; i < h_; i++) s += (*this)(i, j) * x[i];
      y[j] = s;
    }
  }
  void AddMult(const Vector &x, Vector &y) const {
    for (int i = 0; i < h_; i++) {
      double s = 0;
      for (int j = 0; j < w_; j++) s += (*this)(i, j) * x[j];
      y[i] += s;
    }
  }
  void GetColumn(int c, Vector &col) const {
    col.SetSize(h_);
    for (int i = 0; i < h_; i++) col[i] = (*this)(i, c);
  }
  void GetRow(int r, Vector &row) const {
    row.SetSize(w_);
    for (int j = 0; j < w_; j++) row[j] = (*this)(r, j);
  }
  void SetCol(int c, const Vector &col) {
    for (int i = 0; i < h_; i++) (*this)(i, c) = col[i];
  }
  void SetRow(int r, const Vector &row) {
    for (int j = 0; j < w_; j++) (*this)(r, j) = row[j];
  }
  void Print(std::ostream &os = std::cout, int = 4) const {
    for (int i = 0; i < h_; i++) {
      for (int j = 0; j < w_; j++) os << (*this)(i, j) << " ";
      os << "\n";
    }
  }
};

// [MFEM CalcInverse, linalg/densemat.cpp]: adjugate / determinant for the 1 x 1, 2 x 2 and 3 x 3 matrices the
// reference inverts per point (WallBC::computeSlipWallFlux, src/wallBC.cpp:414)
inline void CalcInverse(const DenseMatrix &a, DenseMatrix &inva) {
  const int n = a.Height();
  if (n == 1) {
    inva(0, 0) = 1.0 / a(0, 0);
  } else if (n == 2) {
    const double t = 1.0 / (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0));
    inva(0, 0) = a(1, 1) * t;
    inva(0, 1) = -a(0, 1) * t;
    inva(1, 0) = -a(1, 0) * t;
    inva(1, 1) = a(0, 0) * t;
  } else {
    const double det = a(0, 0) * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) - a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) +
                       a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
    const double t = 1.0 / det;
    inva(0, 0) = (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) * t;
    inva(0, 1) = (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) * t;
    inva(0, 2) = (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) * t;
    inva(1, 0) = (a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2)) * t;
    inva(1, 1) = (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) * t;
    inva(1, 2) = (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) * t;
    inva(2, 0) = (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0)) * t;
    inva(2, 1) = (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) * t;
    inva(2, 2) = (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) * t;
  }
}

class DenseTensor {
  std::vector<double> d_;
  int ni_ = 0, nj_ = 0, nk_ = 0;

 public:
  DenseTensor() {}
  DenseTensor(int i, int j, int k) { SetSize(i, j, k); }
  void SetSize(int i, int j, int k) {
    ni_ = i;
    nj_ = j;
    nk_ = k;
    d_.assign(static_cast<size_t>(i) * j * k, 0.0);
  }
  int SizeI() const { return ni_; }
  int SizeJ() const { return nj_; }
  int SizeK() const { return nk_; }
  double &operator()(int i, int j, int k) { return d_[i + ni_ * (j + static_cast<size_t>(nj_) * k)]; }
  const double &operator()(int i, int j, int k) const { return d_[i + ni_ * (j + static_cast<size_t>(nj_) * k)]; }
};

/* y_mat += v * w^T  (v: height, w: width) */
inline void AddMultVWt(const Vector &v, const Vector &w, DenseMatrix &VWt) {
  for (int i = 0; i < v.Size(); i++)
    for (int j = 0; j < w.Size(); j++) VWt(i, j) += v[i] * w[j];
}
inline void AddMult_a_VWt(double a, const Vector &v, const Vector &w, DenseMatrix &VWt) {
  for (int i = 0; i < v.Size(); i++)
    for (int j = 0; j < w.Size(); j++) VWt(i, j) += a * v[i] * w[j];
}
inline void Mult(const DenseMatrix &b, const DenseMatrix &c, DenseMatrix &a) {
  for (int i = 0; i < b.Height(); i++)
    for (int j = 0; j < c.Width(); j++) {
      double s = 0;
      for (int k = 0; k < b.Width(); k++) s += b(i, k) * c(k, j);
      a(i, j) = s;
    }
}

struct Ordering {
  enum Type { byNODES, byVDIM };
};

class FiniteElementSpace {
  int ndofs_ = 0, vdim_ = 1;

 public:
  explicit FiniteElementSpace(int n = 0, int vdim = 1) : ndofs_(n), vdim_(vdim) {}
  int GetNDofs() const { return ndofs_; }
  int GetVDim() const { return vdim_; }
  Ordering::Type GetOrdering() const { return Ordering::byNODES; }
};
typedef FiniteElementSpace ParFiniteElementSpace;

class GridFunction : public Vector {
  FiniteElementSpace *fes_ = NULL;

 public:
  GridFunction() {}
  explicit GridFunction(FiniteElementSpace *f) : Vector(f->GetNDofs()), fes_(f) {}
  GridFunction(FiniteElementSpace *f, int vdim) : Vector(f->GetNDofs() * vdim), fes_(f) {}
  FiniteElementSpace *FESpace() const { return fes_; }
  FiniteElementSpace *ParFESpace() const { return fes_; }
};
typedef GridFunction ParGridFunction;

class IntegrationRules {};
class ParMesh {};
class Mesh {};

}  // namespace mfem
