// PVFMM stand-in for the oracle build (TEST INFRASTRUCTURE ONLY -- never linked
// into the product library).
//
// The reference (arashb/tbslas @ 0e66711) is header-only C++ templates layered on
// PVFMM (github dmalhotra/pvfmm), which is NOT vendored, NOT version-pinned
// (reference Makefile:1-13 only includes $(PVFMM_DIR)/MakeVariables) and not
// installed here.  This header restates, from PVFMM's published sources, exactly
// the primitives the semi-Lagrangian hot path touches so that the reference's
// own headers (src/tree/tree_functor.h, src/semilag/*.h, src/utils/cubic.h,
// src/tree/tree_set_functor.h, src/tree/tree_extrap_functor.h, src/utils/cheb.h)
// compile VERBATIM from /root/reference and run single-rank.  Everything that is
// arithmetic on the path and lives in the reference itself (vec_eval, the coordinate
// rescale, RK2, fast_interp, InterpCubic1D, the functors) is therefore the
// reference's own code; only the items below are restatements:
//   * MortonId (integer anchor at depth MAX_DEPTH=15, z-major comparator)
//   * cheb_poly (three-term recurrence, all-zero outside [-1,1]) and cheb_eval
//   * containers Vector/Matrix (+ Jacobi-SVD pinv), intrinsics wrappers, the sort,
//     Profile (no-op), par::Scatter* (np == 1: no-ops).
#ifndef TBSLAS_ORACLE_SHIM_PVFMM_SHIM_HPP_
#define TBSLAS_ORACLE_SHIM_PVFMM_SHIM_HPP_

#include <mpi.h>
#include <omp.h>
#include <stdint.h>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <parallel/algorithm>
#include <vector>

#if defined(__AVX__)
#include <immintrin.h>
#endif

#ifndef MAX_DEPTH
#define MAX_DEPTH 15
#endif
#define COORD_DIM 3
#define ASSERT_WITH_MSG(cond, msg) assert((cond) && (msg))

namespace pvfmm {

enum BoundaryType { FreeSpace, Periodic };

// ---------------------------------------------------------------- containers
template <class T>
class Vector {
 public:
  Vector() {}
  explicit Vector(size_t n) : v_(n) {}
  size_t Dim() const { return v_.size(); }
  void Resize(size_t n) { v_.resize(n); }
  void ReInit(size_t n) { v_.assign(n, T()); }
  void SetZero() { std::fill(v_.begin(), v_.end(), T()); }
  T &operator[](size_t i) { return v_.data()[i]; }
  const T &operator[](size_t i) const { return v_.data()[i]; }

 private:
  std::vector<T> v_;
};

template <class T>
class Matrix {
 public:
  Matrix() : r_(0), c_(0), p_(NULL), own_(true) {}
  Matrix(size_t r, size_t c, T *ptr = NULL, bool own = true)
      : r_(r), c_(c), p_(NULL), own_(own) {
    if (own_) {
      p_ = (r * c) ? new T[r * c]() : NULL;
      if (ptr) std::memcpy(p_, ptr, r * c * sizeof(T));
    } else {
      p_ = ptr;
    }
  }
  Matrix(const Matrix &m) : r_(m.r_), c_(m.c_), p_(NULL), own_(true) {
    p_ = (r_ * c_) ? new T[r_ * c_] : NULL;
    if (p_) std::memcpy(p_, m.p_, r_ * c_ * sizeof(T));
  }
  ~Matrix() {
    if (own_) delete[] p_;
  }
  Matrix &operator=(const Matrix &m) {
    if (this == &m) return *this;
    if (own_) {
      if (r_ * c_ != m.r_ * m.c_) {
        delete[] p_;
        p_ = (m.r_ * m.c_) ? new T[m.r_ * m.c_] : NULL;
      }
      r_ = m.r_;
      c_ = m.c_;
    } else {
      assert(r_ * c_ == m.r_ * m.c_);
      r_ = m.r_;
      c_ = m.c_;
    }
    if (p_) std::memcpy(p_, m.p_, r_ * c_ * sizeof(T));
    return *this;
  }
  size_t Dim(size_t i) const { return i == 0 ? r_ : c_; }
  void Resize(size_t r, size_t c) { ReInit(r, c); }
  void ReInit(size_t r, size_t c, T *ptr = NULL, bool own = true) {
    if (own_) delete[] p_;
    r_ = r;
    c_ = c;
    own_ = own;
    if (own_) {
      p_ = (r * c) ? new T[r * c]() : NULL;
      if (ptr) std::memcpy(p_, ptr, r * c * sizeof(T));
    } else {
      p_ = ptr;
    }
  }
  void SetZero() {
    if (p_) std::memset(p_, 0, r_ * c_ * sizeof(T));
  }
  T *operator[](size_t i) { return p_ + i * c_; }
  const T *operator[](size_t i) const { return p_ + i * c_; }

  Matrix Transpose() const {
    Matrix t(c_, r_);
    for (size_t i = 0; i < r_; i++)
      for (size_t j = 0; j < c_; j++) t[j][i] = (*this)[i][j];
    return t;
  }
  // C = A * B (row major, plain triple loop; off the measured path).
  static void GEMM(Matrix &C, const Matrix &A, const Matrix &B) {
    assert(A.c_ == B.r_ && C.r_ == A.r_ && C.c_ == B.c_);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)A.r_; i++) {
      T *c = C[i];
      for (size_t j = 0; j < B.c_; j++) c[j] = 0;
      for (size_t k = 0; k < A.c_; k++) {
        const T a = A[i][k];
        const T *b = B[k];
        for (size_t j = 0; j < B.c_; j++) c[j] += a * b[j];
      }
    }
  }
  // Moore-Penrose pseudo-inverse by one-sided Jacobi SVD (PVFMM calls LAPACK
  // dgesvd and drops singular values below eps*max; same truncation rule).
  Matrix pinv(T eps = -1) const {
    const bool tall = r_ >= c_;
    Matrix A = tall ? *this : Transpose();  // m x n, m >= n
    const size_t m = A.r_, n = A.c_;
    Matrix V(n, n);
    for (size_t i = 0; i < n; i++) V[i][i] = 1;
    for (int sweep = 0; sweep < 60; sweep++) {
      double off = 0;
      for (size_t p = 0; p + 1 < n; p++)
        for (size_t q = p + 1; q < n; q++) {
          double a = 0, b = 0, g = 0;
          for (size_t i = 0; i < m; i++) {
            a += (double)A[i][p] * A[i][p];
            b += (double)A[i][q] * A[i][q];
            g += (double)A[i][p] * A[i][q];
          }
          if (g == 0 || std::fabs(g) <= 1e-17 * std::sqrt(a * b)) continue;
          off = std::max(off, std::fabs(g) / std::sqrt(a * b));
          const double zeta = (b - a) / (2 * g);
          const double t = (zeta >= 0 ? 1.0 : -1.0) /
                           (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
          const double cs = 1 / std::sqrt(1 + t * t), sn = cs * t;
          for (size_t i = 0; i < m; i++) {
            const T x = A[i][p], y = A[i][q];
            A[i][p] = cs * x - sn * y;
            A[i][q] = sn * x + cs * y;
          }
          for (size_t i = 0; i < n; i++) {
            const T x = V[i][p], y = V[i][q];
            V[i][p] = cs * x - sn * y;
            V[i][q] = sn * x + cs * y;
          }
        }
      if (off < 1e-15) break;
    }
    std::vector<double> s(n);
    double smax = 0;
    for (size_t j = 0; j < n; j++) {
      double a = 0;
      for (size_t i = 0; i < m; i++) a += (double)A[i][j] * A[i][j];
      s[j] = std::sqrt(a);
      smax = std::max(smax, s[j]);
    }
    if (eps < 0) {
      eps = 1;
      while (eps + (T)1 > (T)1) eps *= 0.5;
      eps = std::sqrt(eps);
    }
    // A = U S V^T with U = A_rot / s  =>  pinv(A) = V S^-1 U^T   (n x m)
    Matrix P(n, m);
    for (size_t j = 0; j < n; j++) {
      if (s[j] <= eps * smax) continue;
      const double inv = 1.0 / (s[j] * s[j]);
      for (size_t a = 0; a < n; a++) {
        const double vaj = V[a][j] * inv;
        for (size_t i = 0; i < m; i++) P[a][i] += vaj * A[i][j];
      }
    }
    return tall ? P : P.Transpose();
  }

 private:
  size_t r_, c_;
  T *p_;
  bool own_;
};

namespace mem {
template <class T>
struct TypeTraits {
  static uintptr_t ID() {
    static char tag;
    return (uintptr_t)&tag;
  }
};
}  // namespace mem

// ------------------------------------------------------------------ MortonId
// Integer anchor (x,y,z) at depth MAX_DEPTH plus the node depth.  Ordering is the
// Morton (Z-) curve with z the most significant axis, ties (same anchor) broken
// by depth -- i.e. an ancestor sorts before its descendants.
typedef uint32_t UINT_T;

class MortonId {
 public:
  MortonId() : x(0), y(0), z(0), depth(0) {}
  template <class T>
  MortonId(T x_f, T y_f, T z_f, uint8_t depth_ = MAX_DEPTH) : depth(depth_) {
    const UINT_T max_int = ((UINT_T)1) << (MAX_DEPTH);
    x = (UINT_T)std::floor(x_f * max_int);
    y = (UINT_T)std::floor(y_f * max_int);
    z = (UINT_T)std::floor(z_f * max_int);
  }
  template <class T>
  explicit MortonId(T *coord, uint8_t depth_ = MAX_DEPTH) : depth(depth_) {
    const UINT_T max_int = ((UINT_T)1) << (MAX_DEPTH);
    x = (UINT_T)std::floor(coord[0] * max_int);
    y = (UINT_T)std::floor(coord[1] * max_int);
    z = (UINT_T)std::floor(coord[2] * max_int);
  }
  unsigned int GetDepth() const { return depth; }
  int operator<(const MortonId &m) const {
    if (x == m.x && y == m.y && z == m.z) return depth < m.depth;
    const UINT_T x_ = (x ^ m.x), y_ = (y ^ m.y), z_ = (z ^ m.z);
    if ((z_ > x_ || ((z_ ^ x_) < x_ && (z_ ^ x_) < z_)) &&
        (z_ > y_ || ((z_ ^ y_) < y_ && (z_ ^ y_) < z_)))
      return z < m.z;
    if (y_ > x_ || ((y_ ^ x_) < x_ && (y_ ^ x_) < y_)) return y < m.y;
    return x < m.x;
  }
  int operator>(const MortonId &m) const { return m < *this; }
  int operator==(const MortonId &m) const {
    return x == m.x && y == m.y && z == m.z && depth == m.depth;
  }
  int operator!=(const MortonId &m) const { return !(*this == m); }
  int operator<=(const MortonId &m) const { return !(m < *this); }
  int operator>=(const MortonId &m) const { return !(*this < m); }

  UINT_T x, y, z;
  uint8_t depth;
};

// --------------------------------------------------------------- cheb basis
// T_0..T_d at n points, degree-major out[i*n+j]; every basis value is 0 when the
// point lies outside [-1,1] (this is what makes out-of-leaf points evaluate to 0).
template <class T>
inline void cheb_poly(int d, const T *in, int n, T *out) {
  if (d == 0) {
    for (int i = 0; i < n; i++) out[i] = (std::fabs(in[i]) <= 1 ? 1.0 : 0);
  } else if (d == 1) {
    for (int i = 0; i < n; i++) {
      out[i] = (std::fabs(in[i]) <= 1 ? 1.0 : 0);
      out[i + n] = (std::fabs(in[i]) <= 1 ? in[i] : 0);
    }
  } else {
    for (int j = 0; j < n; j++) {
      const T x = (std::fabs(in[j]) <= 1 ? in[j] : 0);
      T y0 = (std::fabs(in[j]) <= 1 ? 1.0 : 0);
      out[j] = y0;
      out[j + n] = x;
      T y1 = x;
      T *y2 = &out[2 * n + j];
      for (int i = 2; i <= d; i++) {
        *y2 = 2 * x * y1 - y0;
        y0 = y1;
        y1 = *y2;
        y2 = &y2[n];
      }
    }
  }
}

// All tensor basis values T_i(z) T_j(y) T_k(x), i+j+k <= deg, at ONE point, in the
// coefficient storage order (i outermost, k innermost).
template <class T>
inline void cheb_eval(int cheb_deg, T *coord, T *coeff0, T *buff) {
  const int d = cheb_deg + 1;
  T *p = buff;  // p[i*3+axis]
  cheb_poly(cheb_deg, coord, 3, p);
  int indx = 0;
  for (int i = 0; i < d; i++)
    for (int j = 0; i + j < d; j++)
      for (int k = 0; i + j + k < d; k++)
        coeff0[indx++] = p[i * 3 + 2] * (p[j * 3 + 1] * p[k * 3 + 0]);
}

template <class T>
inline std::vector<T> cheb_nodes(int deg, int dim) {
  const unsigned int d = deg + 1;
  std::vector<T> x(d);
  for (unsigned int i = 0; i < d; i++)
    x[i] = -std::cos((i + (T)0.5) * (T)M_PI / d) * 0.5 + 0.5;
  if (dim == 1) return x;
  unsigned int n1 = 1;
  for (int i = 0; i < dim; i++) n1 *= d;
  std::vector<T> y(n1 * dim);
  for (int i = 0; i < dim; i++) {
    unsigned int n2 = 1;
    for (int k = 0; k < i; k++) n2 *= d;
    for (unsigned int j = 0; j < n1; j++) y[j * dim + i] = x[(j / n2) % d];
  }
  return y;
}

template <class T>
inline T cos(T x) {
  return std::cos(x);
}
template <class T>
inline T const_pi() {
  return (T)3.14159265358979323846264338327950288L;
}
template <class T>
inline T pow(T b, int e) {
  T r = 1;
  for (int i = 0; i < e; i++) r *= b;
  return r;
}

// --------------------------------------------------------------- intrinsics
template <class T>
inline T zero_intrin() {
  return (T)0;
}
template <class T, class Real>
inline T set_intrin(const Real &a) {
  return a;
}
template <class T, class Real>
inline T load_intrin(Real const *a) {
  return a[0];
}
template <class T, class Real>
inline void store_intrin(Real *a, const T &b) {
  a[0] = b;
}
template <class T>
inline T mul_intrin(const T &a, const T &b) {
  return a * b;
}
template <class T>
inline T add_intrin(const T &a, const T &b) {
  return a + b;
}
#if defined(__AVX__)
template <>
inline __m256d zero_intrin() {
  return _mm256_setzero_pd();
}
template <>
inline __m256d set_intrin(const double &a) {
  return _mm256_set1_pd(a);
}
template <>
inline __m256d load_intrin(double const *a) {
  return _mm256_loadu_pd(a);
}
template <>
inline void store_intrin(double *a, const __m256d &b) {
  _mm256_storeu_pd(a, b);
}
template <>
inline __m256d mul_intrin(const __m256d &a, const __m256d &b) {
  return _mm256_mul_pd(a, b);
}
template <>
inline __m256d add_intrin(const __m256d &a, const __m256d &b) {
  return _mm256_add_pd(a, b);
}
template <>
inline __m256 zero_intrin() {
  return _mm256_setzero_ps();
}
template <>
inline __m256 set_intrin(const float &a) {
  return _mm256_set1_ps(a);
}
template <>
inline __m256 load_intrin(float const *a) {
  return _mm256_loadu_ps(a);
}
template <>
inline void store_intrin(float *a, const __m256 &b) {
  _mm256_storeu_ps(a, b);
}
template <>
inline __m256 mul_intrin(const __m256 &a, const __m256 &b) {
  return _mm256_mul_ps(a, b);
}
template <>
inline __m256 add_intrin(const __m256 &a, const __m256 &b) {
  return _mm256_add_ps(a, b);
}
#endif

// ------------------------------------------------------------------ Profile
class Profile {
 public:
  static void Tic(const char *, const MPI_Comm * = NULL, bool = false, int = 0) {}
  static void Toc() {}
  static long long Add_FLOP(long long f) {
    long long &c = flop_();
    c += f;
    return c;
  }
  static bool Enable(bool) { return false; }
  static void print(const MPI_Comm * = NULL) {}
  static long long &flop_() {
    static long long c = 0;
    return c;
  }
};

// ---------------------------------------------------------------- sort / par
namespace omp_par {
template <class It>
inline void merge_sort(It a, It b) {
  // PVFMM: OpenMP merge sort (not stable).  Any sort gives the same leaf
  // assignment; equal keys only permute points inside one leaf.
  __gnu_parallel::stable_sort(a, b);
}
}  // namespace omp_par

namespace par {
template <typename T, typename D>
struct SortPair {
  T key;
  D data;
  int operator<(const SortPair<T, D> &p) const { return key < p.key; }
};
template <typename T>
struct Mpi_datatype {
  static MPI_Datatype value() { return (MPI_Datatype)sizeof(T); }
};
// np == 1: every key already lives on its owner; the scatter is the identity on
// an (empty) outsider set.
template <typename T>
inline int SortScatterIndex(const Vector<T> &key, Vector<size_t> &scatter_index,
                            const MPI_Comm &, const T * = NULL) {
  scatter_index.Resize(key.Dim());
  for (size_t i = 0; i < key.Dim(); i++) scatter_index[i] = i;
  return 0;
}
template <typename T>
inline int ScatterForward(Vector<T> &, const Vector<size_t> &, const MPI_Comm &) {
  return 0;
}
template <typename T>
inline int ScatterReverse(Vector<T> &, const Vector<size_t> &, const MPI_Comm &,
                          size_t = 0) {
  return 0;
}
// Declared only so that tree_utils.h / tree_semilag.h parse (non-dependent names);
// they belong to tree construction, merging and curl, which are outside the hot path and
// are never called through this stand-in.
template <class V, class W>
inline int HyperQuickSort(const V &, W &, const MPI_Comm &) {
  abort();
}
template <class T, class V>
inline int partitionW(V &, unsigned int *, const MPI_Comm &) {
  abort();
}
}  // namespace par
template <class T, class Y>
inline T cheb_approx(T *, int, int, T *, void * = NULL) {
  abort();
}
template <class T>
inline void cheb_curl(T *, int, T *, void * = NULL) {
  abort();
}

// -------------------------------------------------------------- tree / nodes
template <class Real>
class Cheb_Node {
 public:
  typedef Real Real_t;
  Cheb_Node() : depth_(0), deg_(0), dof_(1), leaf_(true), ghost_(false) {
    coord_[0] = coord_[1] = coord_[2] = 0;
  }
  bool IsLeaf() const { return leaf_; }
  bool IsGhost() const { return ghost_; }
  int DataDOF() const { return dof_; }
  int ChebDeg() const { return deg_; }
  Vector<Real> &ChebData() { return data_; }
  Real *Coord() { return coord_; }
  size_t Depth() const { return depth_; }
  MortonId GetMortonId() { return MortonId(coord_, (uint8_t)depth_); }

  Real coord_[3];
  size_t depth_;
  int deg_, dof_;
  bool leaf_, ghost_;
  Vector<Real> data_;
};

// Flat leaf list standing in for pvfmm::MPI_Tree: leaves are stored in Morton
// order, which is what PVFMM's preorder/postorder traversals visit leaves in.
template <class Node>
class MPI_Tree {
 public:
  typedef Node Node_t;
  typedef typename Node::Real_t Real_t;
  MPI_Tree() : comm_(MPI_COMM_WORLD) {}
  ~MPI_Tree() {
    for (size_t i = 0; i < nodes_.size(); i++) delete nodes_[i];
  }
  std::vector<Node_t *> &GetNodeList() { return nodes_; }
  const MPI_Comm *Comm() const { return &comm_; }
  int Dim() const { return COORD_DIM; }
  Node_t *RootNode() { return nodes_.empty() ? NULL : nodes_[0]; }
  Node_t *PostorderFirst() { return nodes_.empty() ? NULL : nodes_[0]; }
  Node_t *PostorderNxt(Node_t *n) {
    // linear position lookup is fine at oracle sizes; cache the last hit
    if (last_ < nodes_.size() && nodes_[last_] == n) {
      last_++;
    } else {
      last_ = (std::find(nodes_.begin(), nodes_.end(), n) - nodes_.begin()) + 1;
    }
    return last_ < nodes_.size() ? nodes_[last_] : NULL;
  }
  Node_t *PreorderFirst() { return PostorderFirst(); }
  Node_t *PreorderNxt(Node_t *n) { return PostorderNxt(n); }

 private:
  std::vector<Node_t *> nodes_;
  MPI_Comm comm_;
  size_t last_ = 0;
};

}  // namespace pvfmm

#endif  // TBSLAS_ORACLE_SHIM_PVFMM_SHIM_HPP_
