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

class IntegerVector : public std::vector<int> {
 public:
  using std::vector<int>::vector;
  IntegerVector() = default;
  IntegerVector(const std::vector<int>& v) : std::vector<int>(v) {}
};

// One element of a List.  The reference only ever passes numeric vectors; the host layer of this repo also accepts what
// data pre-process.R builds — numeric matrices (column-major, `nrow` x `ncol`) and Matrix::dgCMatrix objects (an S4 object
// with slots i, p, x, Dim: compressed COLUMNS) — so the stand-in can carry those too.
struct ListElem {
  std::vector<double> values;          // vector, or matrix in column-major order, or the x slot of a dgCMatrix
  int nrow = 0, ncol = 0;              // > 0: a matrix
  bool is_dgc = false;                 // a dgCMatrix: values = x, slot_i, slot_p, nrow / ncol = Dim
  std::vector<int> slot_i, slot_p;
};
}  // namespace Rcpp
typedef const Rcpp::ListElem* SEXP;    // (R's SEXP is an opaque pointer too)
inline bool Rf_isMatrix(SEXP e) { return !e->is_dgc && e->nrow > 0; }
inline bool Rf_isS4(SEXP e) { return e->is_dgc; }
namespace Rcpp {

class NumericMatrix {                  // Rcpp::NumericMatrix: column-major view
 public:
  explicit NumericMatrix(SEXP e) : e_(e) {}
  int nrow() const { return e_->nrow; }
  int ncol() const { return e_->ncol; }
  double operator()(int i, int j) const { return e_->values[(size_t)j * e_->nrow + i]; }
 private:
  SEXP e_;
};

class S4 {                             // Rcpp::S4: is(class), slot(name)
 public:
  struct SlotProxy {
    SEXP e; std::string name;
    operator IntegerVector() const {
      if (name == "i") return IntegerVector(e->slot_i);
      if (name == "p") return IntegerVector(e->slot_p);
      if (name == "Dim") return IntegerVector(std::vector<int>{e->nrow, e->ncol});
      throw std::runtime_error("no integer slot " + name);
    }
    operator NumericVector() const;
  };
  explicit S4(SEXP e) : e_(e) {}
  bool is(const std::string& cls) const { return e_->is_dgc && cls == "dgCMatrix"; }
  SlotProxy slot(const std::string& name) const { return SlotProxy{e_, name}; }
 private:
  SEXP e_;
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
  struct ElemRef {                     // what operator[] yields: usable as the element and convertible to SEXP, like Rcpp's proxy
    const ListElem* e;
    operator SEXP() const { return e; }
    operator const ListElem&() const { return *e; }
  };
  ElemRef operator[](int i) const { return ElemRef{&elems_.at(static_cast<size_t>(i))}; }
  void push_back(const std::vector<double>& v) { ListElem e; e.values = v; elems_.push_back(e); }
  void push_back_matrix(const std::vector<double>& colmajor, int nrow, int ncol) {          // stand-in only
    ListElem e; e.values = colmajor; e.nrow = nrow; e.ncol = ncol; elems_.push_back(e);
  }
  void push_back_dgc(const std::vector<int>& i, const std::vector<int>& p, const std::vector<double>& x, int nrow, int ncol) {
    ListElem e; e.values = x; e.slot_i = i; e.slot_p = p; e.nrow = nrow; e.ncol = ncol; e.is_dgc = true; elems_.push_back(e);
  }
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
template <class T>
T as(SEXP e) { return as<T>(*e); }
template <class T>
T as(const List::ElemRef& r) { return as<T>(*r.e); }
inline S4::SlotProxy::operator NumericVector() const {
  if (name == "x") return NumericVector(e->values);
  throw std::runtime_error("no numeric slot " + name);
}

[[noreturn]] inline void stop(const std::string& msg) { throw std::runtime_error(msg); }

extern std::ostream Rcout;  // a sink (defined in ref_shim.cpp); progress lines are discarded

}  // namespace Rcpp

#endif  // MV_ORACLE_RCPP_STANDIN_H
