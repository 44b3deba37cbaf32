// oracle/ref_build/include/Rcpp.h -- TEST INFRASTRUCTURE ONLY (oracle/ref_build/README.md).
// Stand-in for the handful of Rcpp names GPvecchia's hot-path sources use: List / List::create / _[""],
// NumericVector (a shared handle: copies alias the same storage, as an R vector passed by value to a
// C++ function does -- ic0() relies on that, src/ic0.cpp:69-70,89), Rcerr / Rcout (counted, then dropped).
#ifndef GPV_REF_STUB_RCPP_H
#define GPV_REF_STUB_RCPP_H

#include <atomic>
#include <map>
#include <memory>
#include <ostream>
#include <string>
#include <RcppArmadillo.h>

namespace Rcpp {

class CountingStream {
 public:
  std::atomic<long> lines{0};
  template <class T> CountingStream& operator<<(const T&) { return *this; }
  CountingStream& operator<<(std::ostream& (*)(std::ostream&)) { lines.fetch_add(1, std::memory_order_relaxed); return *this; }
};
inline CountingStream Rcerr;
inline CountingStream Rcout;

template <class T> struct named_object { std::string name; const T& object; };
class Named {
 public:
  std::string name;
  explicit Named(const std::string& n) : name(n) {}
  template <class T> named_object<T> operator=(const T& o) const { return named_object<T>{name, o}; }
};
namespace internal {
struct NamedPlaceHolder { Named operator[](const std::string& n) const { return Named(n); } };
}  // namespace internal
static internal::NamedPlaceHolder _;

class List {
 public:
  std::map<std::string, arma::mat> items;
  template <class... Args>
  static List create(const Args&... args) { List l; (l.items.emplace(args.name, arma::mat(args.object)), ...); return l; }
  const arma::mat& operator[](const std::string& n) const { return items.at(n); }
};

class NumericVector {
  std::shared_ptr<double[]> own;
  double* p = nullptr;
  long n = 0;

 public:
  NumericVector() {}
  explicit NumericVector(long n_) : own(new double[n_ > 0 ? n_ : 1]()), p(own.get()), n(n_) {}
  NumericVector(double* ext, long n_) : p(ext), n(n_) {}      // view of caller memory (an R vector)
  long size() const { return n; }
  double& operator[](long i) { return p[i]; }
  const double& operator[](long i) const { return p[i]; }
  double* begin() { return p; }
};

}  // namespace Rcpp
#endif
