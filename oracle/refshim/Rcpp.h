// oracle/refshim/Rcpp.h — TEST INFRASTRUCTURE, not product code.
//
// A minimal stand-in for the slice of the Rcpp / R API that the reference's four
// translation units touch (SURVEY.md Appendix D lists the surface: R::runif, R::rnorm,
// Rcpp::stop, Rcpp::Rcout, Rcpp::List{size,operator[],create}, Rcpp::Named,
// Rcpp::as<NumericVector | std::vector<double>>).  With this header on the include path
// the reference sources under /root/reference/Multiview compile UNMODIFIED into
// oracle/_ref/libmvref.so (see oracle/Makefile).  R itself is not installed in this image.
//
// The random numbers behind R::runif / R::rnorm come from ref_shim.cpp: either a
// call-ordered Philox4x32-10 stream (seedable) or a scripted queue of uniforms, so that
// tests can drive the reference and the restated oracle with the same uniforms.
#ifndef MV_ORACLE_RCPP_STANDIN_H
#define MV_ORACLE_RCPP_STANDIN_H

#include <cmath>
#include <cstddef>
#include <ostream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace R {
double runif(double a, double b);
double rnorm(double mean, double sd);
}  // namespace R

namespace Rcpp {

class NumericVector : public std::vector<double> {
 public:
  using std::vector<double>::vector;
  NumericVector() = default;
  NumericVector(const std::vector<double>& v) : std::vector<double>(v) {}
};

// One element of a List: the stand-in only ever carries numeric vectors inward.
struct ListElem {
  std::vector<double> values;
};

// Named("x") = value  — outward values are dropped; the shim reads the reference's
// saved_* globals directly instead of unpacking a returned list.
struct NamedPlaceholder {
  template <class T>
  NamedPlaceholder& operator=(const T&) { return *this; }
};
inline NamedPlaceholder Named(const char*) { return NamedPlaceholder(); }
inline NamedPlaceholder Named(const std::string&) { return NamedPlaceholder(); }

class List {
 public:
  List() = default;
  int size() const { return static_cast<int>(elems_.size()); }
  const ListElem& operator[](int i) const { return elems_.at(static_cast<size_t>(i)); }
  void push_back(const std::vector<double>& v) { elems_.push_back(ListElem{v}); }
  template <class... Args>
  static List create(const Args&...) { return List(); }

 private:
  std::vector<ListElem> elems_;
};

template <class T>
T as(const ListElem& e);
template <>
inline NumericVector as<NumericVector>(const ListElem& e) { return NumericVector(e.values); }
template <>
inline std::vector<double> as<std::vector<double>>(const ListElem& e) { return e.values; }

[[noreturn]] inline void stop(const std::string& msg) { throw std::runtime_error(msg); }

extern std::ostream Rcout;  // a sink (defined in ref_shim.cpp); progress lines are discarded

}  // namespace Rcpp

#endif  // MV_ORACLE_RCPP_STANDIN_H
