// oracle/ref_build/include/RcppArmadillo.h
//
// TEST INFRASTRUCTURE ONLY (see oracle/ref_build/README.md).  A minimal stand-in for the part of
// Armadillo that GPvecchia's hot-path sources use, so that the UNMODIFIED files
//   /root/reference/src/{U_NZentries,Matern,Esqe,dist,ic0}.cpp
// compile in an image that has neither R nor RcppArmadillo nor BH.  Nothing here is taken from
// Armadillo's sources (they are not in the image); it re-creates the documented public behaviour of
// the few classes / functions those five files call:
//   Mat / Col / Row (column-major, uninitialised sized constructor), uword = 64-bit unsigned
//   (RcppArmadillo's ARMA_64BIT_WORD default), span, zeros, ones, find, elem, rows, row, submat, t,
//   element-wise + - %, diagmat, chol(X,"upper"), solve(A,b).
// chol  -> LAPACK dpotrf('U') then the strict lower triangle set to zero; failure throws
//          std::runtime_error("chol(): decomposition failed"), which is what U_NZentries.cpp:64 catches.
// solve -> Armadillo's default solve() detects an upper-triangular square A and calls LAPACK dtrtrs
//          (for a 1x1 system it goes through the general LU path: x = b / a, the same bits).  Its
//          rcond < eps fall-back to an SVD least-squares solution is not re-created: a factor that
//          dpotrf accepted in fp64 has cond(R) <= ~1e8.
// LAPACK is bound at run time (gpv_ref_bind_lapack in ref_entry.cpp) from the OpenBLAS inside scipy;
// unbound, the published unblocked dpotf2 / back-substitution algorithms are used.
#ifndef GPV_REF_STUB_RCPPARMADILLO_H
#define GPV_REF_STUB_RCPPARMADILLO_H

#include <cmath>
#include <cstddef>
#include <cstring>
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <utility>

namespace arma {

typedef unsigned long long uword;
typedef long long sword;

struct span {
  uword a, b;
  span(uword a_, uword b_) : a(a_), b(b_) {}
};

namespace lapack_stub {
typedef void (*dpotrf_fn)(const char*, const int*, double*, const int*, int*);
typedef void (*dtrtrs_fn)(const char*, const char*, const char*, const int*, const int*, const double*,
                          const int*, double*, const int*, int*);
inline dpotrf_fn g_dpotrf = nullptr;
inline dtrtrs_fn g_dtrtrs = nullptr;
inline int g_force_textbook = 0;
}  // namespace lapack_stub

template <class T> class Col;
template <class T> class Row;
template <class T> class Mat;

// Lentries(k, span(a, b)) = M.t()
template <class T>
class subview_row_span {
 public:
  Mat<T>& m; uword r, a, b;
  subview_row_span(Mat<T>& m_, uword r_, uword a_, uword b_) : m(m_), r(r_), a(a_), b(b_) {}
  void operator=(const Mat<T>& x) {
    if (x.n_elem != b - a + 1) throw std::logic_error("copy into submatrix: incompatible matrix dimensions");
    for (uword j = a; j <= b; ++j) m.at(r, j) = x.mem[j - a];
  }
};

template <class T>
class Mat {
 public:
  uword n_rows = 0, n_cols = 0, n_elem = 0;
  T* mem = nullptr;

 protected:
  static const uword prealloc = 16;          // Armadillo keeps <= 16 elements inside the object
  T local[prealloc];
  void init(uword r, uword c) {
    n_rows = r; n_cols = c; n_elem = r * c;
    mem = (n_elem <= prealloc) ? local : new T[n_elem];
  }
  void release() { if (mem && mem != local) delete[] mem; mem = nullptr; }

 public:
  Mat() {}
  Mat(uword r, uword c) { init(r, c); }
  Mat(const Mat& o) { init(o.n_rows, o.n_cols); if (n_elem) std::memcpy(mem, o.mem, sizeof(T) * n_elem); }
  Mat(Mat&& o) noexcept { steal(o); }
  Mat(const T* src, uword r, uword c) { init(r, c); if (n_elem) std::memcpy(mem, src, sizeof(T) * n_elem); }
  ~Mat() { release(); }
  Mat& operator=(const Mat& o) {
    if (this == &o) return *this;
    if (n_elem != o.n_elem) { release(); init(o.n_rows, o.n_cols); }
    else { n_rows = o.n_rows; n_cols = o.n_cols; }
    if (n_elem) std::memcpy(mem, o.mem, sizeof(T) * n_elem);
    return *this;
  }
  Mat& operator=(Mat&& o) noexcept { if (this != &o) { release(); steal(o); } return *this; }

 protected:
  void steal(Mat& o) {
    n_rows = o.n_rows; n_cols = o.n_cols; n_elem = o.n_elem;
    if (o.mem == o.local) { mem = local; if (n_elem) std::memcpy(local, o.local, sizeof(T) * n_elem); }
    else { mem = o.mem; }
    o.mem = nullptr; o.n_rows = o.n_cols = o.n_elem = 0;
  }

 public:
  T* memptr() { return mem; }
  const T* memptr() const { return mem; }
  uword size() const { return n_elem; }
  void fill(T v) { for (uword i = 0; i < n_elem; ++i) mem[i] = v; }

  T& at(uword i, uword j) { return mem[i + j * n_rows]; }
  const T& at(uword i, uword j) const { return mem[i + j * n_rows]; }
  T& operator()(uword i, uword j) { return mem[i + j * n_rows]; }
  const T& operator()(uword i, uword j) const { return mem[i + j * n_rows]; }
  T& operator()(uword i) { return mem[i]; }
  const T& operator()(uword i) const { return mem[i]; }
  T& operator[](uword i) { return mem[i]; }
  const T& operator[](uword i) const { return mem[i]; }
  subview_row_span<T> operator()(uword r, const span& s) { return subview_row_span<T>(*this, r, s.a, s.b); }

  Row<T> row(uword k) const;
  Mat rows(const Col<uword>& idx) const;
  Mat submat(const Col<uword>& ri, const Col<uword>& ci) const;
  Col<T> elem(const Col<uword>& idx) const;
  Mat t() const {
    Mat out(n_cols, n_rows);
    for (uword i = 0; i < n_rows; ++i)
      for (uword j = 0; j < n_cols; ++j) out.at(j, i) = at(i, j);
    return out;
  }
};

template <class T>
class Col : public Mat<T> {
 public:
  Col() { this->n_cols = 1; }
  explicit Col(uword n) : Mat<T>(n, 1) {}
  Col(std::initializer_list<T> l) : Mat<T>(l.size(), 1) { uword i = 0; for (const T& v : l) this->mem[i++] = v; }
  Col(const T* src, uword n) : Mat<T>(src, n, 1) {}
  Col(const Col& o) : Mat<T>(o) {}
  Col(Col&& o) noexcept : Mat<T>(std::move(o)) {}
  Col& operator=(const Col& o) { Mat<T>::operator=(o); return *this; }
  Col& operator=(Col&& o) noexcept { Mat<T>::operator=(std::move(o)); return *this; }
  using Mat<T>::operator();
  Col operator()(const span& s) const {
    Col out(s.b - s.a + 1);
    for (uword i = s.a; i <= s.b; ++i) out.mem[i - s.a] = this->mem[i];
    return out;
  }
  Row<T> t() const;
};

template <class T>
class Row : public Mat<T> {
 public:
  Row() { this->n_rows = 1; }
  explicit Row(uword n) : Mat<T>(1, n) {}
  Row(const Row& o) : Mat<T>(o) {}
  Row(Row&& o) noexcept : Mat<T>(std::move(o)) {}
  Row& operator=(const Row& o) { Mat<T>::operator=(o); return *this; }
  Row& operator=(Row&& o) noexcept { Mat<T>::operator=(std::move(o)); return *this; }
  Col<T> t() const { Col<T> out(this->n_elem); for (uword i = 0; i < this->n_elem; ++i) out.mem[i] = this->mem[i]; return out; }
};

template <class T>
Row<T> Col<T>::t() const { Row<T> out(this->n_elem); for (uword i = 0; i < this->n_elem; ++i) out.mem[i] = this->mem[i]; return out; }

template <class T>
Row<T> Mat<T>::row(uword k) const {
  Row<T> out(n_cols);
  for (uword j = 0; j < n_cols; ++j) out.mem[j] = at(k, j);
  return out;
}
template <class T>
Mat<T> Mat<T>::rows(const Col<uword>& idx) const {
  Mat out(idx.n_elem, n_cols);
  for (uword j = 0; j < n_cols; ++j)
    for (uword i = 0; i < idx.n_elem; ++i) out.at(i, j) = at(idx.mem[i], j);
  return out;
}
template <class T>
Mat<T> Mat<T>::submat(const Col<uword>& ri, const Col<uword>& ci) const {
  Mat out(ri.n_elem, ci.n_elem);
  for (uword j = 0; j < ci.n_elem; ++j)
    for (uword i = 0; i < ri.n_elem; ++i) out.at(i, j) = at(ri.mem[i], ci.mem[j]);
  return out;
}
template <class T>
Col<T> Mat<T>::elem(const Col<uword>& idx) const {
  Col<T> out(idx.n_elem);
  for (uword i = 0; i < idx.n_elem; ++i) out.mem[i] = mem[idx.mem[i]];
  return out;
}

typedef Mat<double> mat;
typedef Col<double> vec;
typedef Col<double> colvec;
typedef Row<double> rowvec;
typedef Mat<uword> umat;
typedef Col<uword> uvec;

inline mat zeros(uword r, uword c) { mat out(r, c); out.fill(0.0); return out; }
inline vec zeros(uword n) { vec out(n); out.fill(0.0); return out; }
inline vec ones(uword n) { vec out(n); out.fill(1.0); return out; }

template <class T>
Col<uword> find(const Col<T>& x) {
  uword cnt = 0;
  for (uword i = 0; i < x.n_elem; ++i) cnt += (x.mem[i] != T(0));
  Col<uword> out(cnt);
  cnt = 0;
  for (uword i = 0; i < x.n_elem; ++i) if (x.mem[i] != T(0)) out.mem[cnt++] = i;
  return out;
}

template <class T, class S>
Col<T> operator-(const Col<T>& a, S k) { Col<T> o(a.n_elem); for (uword i = 0; i < a.n_elem; ++i) o.mem[i] = a.mem[i] - T(k); return o; }
template <class T>
Col<T> operator-(const Col<T>& a, const Col<T>& b) {
  if (a.n_elem != b.n_elem) throw std::logic_error("subtraction: incompatible matrix dimensions");
  Col<T> o(a.n_elem); for (uword i = 0; i < a.n_elem; ++i) o.mem[i] = a.mem[i] - b.mem[i]; return o;
}
template <class T>
Col<T> operator%(const Col<T>& a, const Col<T>& b) {
  if (a.n_elem != b.n_elem) throw std::logic_error("element-wise multiplication: incompatible matrix dimensions");
  Col<T> o(a.n_elem); for (uword i = 0; i < a.n_elem; ++i) o.mem[i] = a.mem[i] * b.mem[i]; return o;
}
template <class T>
Mat<T> operator+(const Mat<T>& a, const Mat<T>& b) {
  if (a.n_rows != b.n_rows || a.n_cols != b.n_cols) throw std::logic_error("addition: incompatible matrix dimensions");
  Mat<T> o(a.n_rows, a.n_cols); for (uword i = 0; i < a.n_elem; ++i) o.mem[i] = a.mem[i] + b.mem[i]; return o;
}
template <class T>
Mat<T> diagmat(const Col<T>& v) {
  Mat<T> o(v.n_elem, v.n_elem); o.fill(T(0));
  for (uword i = 0; i < v.n_elem; ++i) o.at(i, i) = v.mem[i];
  return o;
}

// chol(X, "upper")
inline mat chol(const mat& X, const char* layout = "upper") {
  if (X.n_rows != X.n_cols) throw std::logic_error("chol(): given matrix must be square sized");
  if (layout[0] != 'u') throw std::logic_error("chol(): stub implements layout \"upper\" only");
  mat out(X);
  const int n = (int)out.n_rows;
  int info = 0;
  if (n > 0) {
    if (lapack_stub::g_dpotrf && !lapack_stub::g_force_textbook) {
      lapack_stub::g_dpotrf("U", &n, out.memptr(), &n, &info);
    } else {                                   // published unblocked algorithm (dpotf2, upper)
      double* a = out.memptr();
      for (int j = 0; j < n && info == 0; ++j) {
        double ajj = a[j + (size_t)j * n];
        for (int i = 0; i < j; ++i) ajj -= a[i + (size_t)j * n] * a[i + (size_t)j * n];
        if (!(ajj > 0)) { info = j + 1; break; }
        ajj = std::sqrt(ajj);
        a[j + (size_t)j * n] = ajj;
        for (int c = j + 1; c < n; ++c) {
          double v = a[j + (size_t)c * n];
          for (int i = 0; i < j; ++i) v -= a[i + (size_t)j * n] * a[i + (size_t)c * n];
          a[j + (size_t)c * n] = v / ajj;
        }
      }
    }
  }
  if (info != 0) throw std::runtime_error("chol(): decomposition failed");
  for (int j = 0; j < n; ++j)
    for (int i = j + 1; i < n; ++i) out.at(i, j) = 0.0;
  return out;
}

// solve(A, b) for the one shape the reference uses: square upper-triangular A, one right-hand side
inline vec solve(const mat& A, const vec& b) {
  if (A.n_rows != A.n_cols || A.n_rows != b.n_elem) throw std::logic_error("solve(): number of rows in given matrices must be the same");
  const int n = (int)A.n_rows;
  for (int j = 0; j < n; ++j)
    for (int i = j + 1; i < n; ++i)
      if (A.at(i, j) != 0.0) throw std::logic_error("solve(): stub handles upper-triangular systems only");
  vec x(b);
  int info = 0;
  if (n == 1) { x.mem[0] = b.mem[0] / A.mem[0]; if (A.mem[0] == 0.0) info = 1; }
  else if (n > 1 && lapack_stub::g_dtrtrs && !lapack_stub::g_force_textbook) {
    const int one = 1;
    lapack_stub::g_dtrtrs("U", "N", "N", &n, &one, A.memptr(), &n, x.memptr(), &n, &info);
  } else {
    for (int i = n - 1; i >= 0; --i) {
      double v = x.mem[i];
      for (int j = i + 1; j < n; ++j) v -= A.at(i, j) * x.mem[j];
      if (A.at(i, i) == 0.0) { info = i + 1; break; }
      x.mem[i] = v / A.at(i, i);
    }
  }
  if (info != 0) throw std::runtime_error("solve(): solution not found");
  return x;
}

}  // namespace arma

#include <Rcpp.h>
#endif
