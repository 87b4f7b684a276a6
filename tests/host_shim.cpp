// tests/host_shim.cpp — TEST scaffolding for the host layer (multiview-clustering_b200/host): R is not
// installed here, so the sources are compiled against the stand-in Rcpp.h of oracle/refshim and driven
// through these C hooks instead of through R.  Nothing here is product code.
#include <Rcpp.h>

#include <cstring>
#include <sstream>

#include "multiview_gibbs.h"
#include "multiview_hyper.h"
#include "multiview_rng.h"
#include "multiview_state.h"
#include "multiview_utils_decl.h"

namespace R {
double runif(double, double) { return 0.5; }   // unused by the host layer (the device draws from Philox)
double rnorm(double, double) { return 0.0; }
}  // namespace R
namespace Rcpp {
static std::ostringstream g_sink;
std::ostream Rcout(g_sink.rdbuf());
}  // namespace Rcpp

static std::string g_err;

extern "C" {
const char* host_last_error() { return g_err.c_str(); }

// run_gibbs_cpp on d scalar views of n customers (y is [d][n]); returns the number of saved states
int host_run(int n_, int d_, const double* y_, int M, int burn_in, int thin, int cap, unsigned long long seed) {
  try {
    mvhost::table_capacity = cap;
    mvhost::seed = seed;
    Rcpp::List data;
    for (int v = 0; v < d_; ++v) data.push_back(std::vector<double>(y_ + (size_t)v * n_, y_ + (size_t)(v + 1) * n_));
    run_gibbs_cpp(data, M, burn_in, thin);
    return (int)saved_table_of.size();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
// run_gibbs_cpp on a list of [matrix (n x D, column-major as in R), dgCMatrix (n x W, compressed columns), vector]
int host_run_mixed(int n_, int D, const double* colmajor, int W, int nnz, const int* ci, const int* cp, const double* cx,
                   const double* vec, int M, int burn_in, int thin, int cap, unsigned long long seed) {
  try {
    mvhost::table_capacity = cap;
    mvhost::seed = seed;
    mvhost::view_dim.clear();
    Rcpp::List data;
    data.push_back_matrix(std::vector<double>(colmajor, colmajor + (size_t)n_ * D), n_, D);
    data.push_back_dgc(std::vector<int>(ci, ci + nnz), std::vector<int>(cp, cp + W + 1), std::vector<double>(cx, cx + nnz), n_, W);
    data.push_back(std::vector<double>(vec, vec + n_));
    run_gibbs_cpp(data, M, burn_in, thin);
    return (int)saved_table_of.size();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
// the same through MVG_ENGINE_SEQ (mvhost::sequential)
int host_run_seq(int n_, int d_, const double* y_, int M, int burn_in, int thin, unsigned long long seed) {
  mvhost::sequential = true;
  const int rc = host_run(n_, d_, y_, M, burn_in, thin, 64, seed);
  mvhost::sequential = false;
  return rc;
}
int host_saved_T(int s) { return (int)saved_dish_of[(size_t)s][0].size(); }
void host_saved_table_of(int s, int* out) { std::memcpy(out, saved_table_of[(size_t)s].data(), sizeof(int) * saved_table_of[(size_t)s].size()); }
void host_saved_dish_of(int s, int v, int* out) { std::memcpy(out, saved_dish_of[(size_t)s][(size_t)v].data(), sizeof(int) * saved_dish_of[(size_t)s][(size_t)v].size()); }
void host_saved_hypers(int s, int d_, double* out) {   // alpha_v[d], sigma_v[d], tau_v[d], alpha_g, sigma_g
  for (int v = 0; v < d_; ++v) {
    out[v] = saved_alpha_v[(size_t)v][(size_t)s];
    out[d_ + v] = saved_sigma_v[(size_t)v][(size_t)s];
    out[2 * d_ + v] = saved_tau_v[(size_t)v][(size_t)s];
  }
  out[3 * d_] = saved_alpha_global[(size_t)s];
  out[3 * d_ + 1] = saved_sigma_global[(size_t)s];
}
// inspectors on the final mirrored state
double host_log_EPPF(int v, double a, double s) { return log_EPPF(v, a, s); }
double host_log_posterior_given_tau(int v, double tau) { return log_posterior_given_tau(v, tau); }
double host_compute_log_likelihood() { return compute_log_likelihood(); }
int host_final_T() { return T; }
int host_final_K(int v) { return views[(size_t)v].K; }
int host_saved_loglik(double* out) { for (size_t s = 0; s < saved_loglik.size(); ++s) out[s] = saved_loglik[s]; return (int)saved_loglik.size(); }
int host_final_table_of(int* out) { for (int i = 0; i < n; ++i) out[i] = table_of[(size_t)i]; return T; }
void host_final_dish_of(int v, int* out) { for (int t = 0; t < T; ++t) out[t] = dish_of[(size_t)v][(size_t)t]; }
void host_final_hypers(double* out) {
  for (int v = 0; v < d; ++v) { out[v] = views[(size_t)v].alpha_v; out[d + v] = views[(size_t)v].sigma_v; out[2 * d + v] = views[(size_t)v].tau_v; }
  out[3 * d] = alpha_global; out[3 * d + 1] = sigma_global;
}
// compute_table_probs_with_cache / compute_f_vk / compute_f_vk_new on the mirrored state (multiview_utils.h)
int host_table_probs(int i, double* pe, double* pn) {
  try {
    std::vector<double> p;
    std::vector<std::unordered_map<int, double>> cache;
    mvu::table_probs(i, p, *pn, cache);
    for (size_t t = 0; t < p.size(); ++t) pe[t] = p[t];
    return (int)p.size();
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
double host_f_vk(int v, int k, int i) { return mvu::f_vk(v, k, i); }
double host_f_vk_new(int v, int i) { return mvu::f_vk_new(v, i); }
int host_mutator_raises(int which) {               // 1 if the call raised (the documented behaviour)
  try {
    switch (which) {
      case 0: mvu::remove(0); break;
      case 1: mvu::add_existing(0, 0); break;
      case 2: mvu::new_table(); break;
      default: mvu::assign_dishes(0, 0); break;
    }
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
  return 0;
}
double host_uniform01(unsigned seed, int skip) { set_rng_seed(seed); double u = 0; for (int i = 0; i <= skip; ++i) u = uniform01(); return u; }
}
