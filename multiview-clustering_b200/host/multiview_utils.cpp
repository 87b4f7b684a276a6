// multiview_utils.cpp — see multiview_utils.h.  Reference: /root/reference/Multiview/multiview_utils.cpp.
#include "multiview_utils.h"

#include <cmath>
#include <stdexcept>
#include <string>

#include "../../include/mvg.h"
#include "multiview_gibbs.h"     // compute_log_likelihood

#ifdef MVHOST_WITH_RCPP
void mvhost_rcpp_stop(const std::string& msg);   // multiview_gibbs.cpp: Rcpp::stop, i.e. an R error
#endif

namespace {
int dim_of(int v) {
  if ((size_t)v < mvhost::csr_views.size() && mvhost::csr_views[(size_t)v].vocab > 0) return 0;   // a count view has no Gaussian coordinates
  return mvhost::view_dim.empty() ? 1 : mvhost::view_dim[(size_t)v];
}

[[noreturn]] void not_on_device(const char* what) {
  const std::string msg = std::string(what) + ": not supported on the device chain (the sweep moves every customer on the "
                          "GPU: mvg_sweep / gibbs_sampler; the mirrored state is read-only)";
#ifdef MVHOST_WITH_RCPP
  mvhost_rcpp_stop(msg);
#endif
  throw std::runtime_error(msg);
}

// log N(y_i; S1/(tau+n), tau (tau+n+1)/(tau+n)) summed over the D coordinates: the closed form of
// exp(logp(S + {i}) - logp(S)) of multiview_utils.cpp:307-338.
double log_f_dish(int v, int k, int i) {
  const ViewState& V = views[(size_t)v];
  const int D = dim_of(v);
  const double tau = V.tau_v, nk = (double)V.n_vk[(size_t)k];
  const double var = tau * (tau + nk + 1.0) / (tau + nk);
  double dist = 0.0;
  for (int j = 0; j < D; ++j) {
    const double diff = y[(size_t)v][(size_t)i * D + j] - V.sum_y[(size_t)k * D + j] / (tau + nk);
    dist += diff * diff;
  }
  return -0.5 * (double)D * std::log(2.0 * M_PI * var) - 0.5 * dist / var;
}
}  // namespace

// multiview_utils.cpp:16-35 computes per-view variances nobody reads; kept as the no-op it effectively is.
void ensure_global_variances_calculated() {}

double compute_f_vk(int v, int k, int i) {
  if (v < 0 || v >= d || k < 0 || k >= views[(size_t)v].K || i < 0 || i >= n) not_on_device("compute_f_vk: index out of range");
  return std::exp(log_f_dish(v, k, i));
}

double compute_f_vk_new(int v, int i) {              // :340-350 — NOT the n = 0 case of compute_f_vk (that is N(0, tau + 1))
  const int D = dim_of(v);
  const double tau = views[(size_t)v].tau_v;
  double q = 0.0;
  for (int j = 0; j < D; ++j) q += y[(size_t)v][(size_t)i * D + j] * y[(size_t)v][(size_t)i * D + j];
  return std::exp(-0.5 * (double)D * std::log(2.0 * M_PI * tau) - 0.5 * q / tau);
}

// :71-136 with compute_marginal_likelihood_new_table (:40-69) inlined
void compute_table_probs_with_cache(int i, std::vector<double>& prob_existing, double& prob_new,
                                    std::vector<std::unordered_map<int, double>>& cache_fvk) {
  cache_fvk.assign((size_t)d, {});                   // :77-79: the memo is per customer
  prob_existing.assign((size_t)T, 0.0);
  for (int t = 0; t < T; ++t) {
    double logp = 0.0;
    for (int v = 0; v < d; ++v) {
      const int k = dish_of[(size_t)v][(size_t)t];
      auto it = cache_fvk[(size_t)v].find(k);
      double f;
      if (it != cache_fvk[(size_t)v].end()) f = it->second;
      else { f = compute_f_vk(v, k, i); cache_fvk[(size_t)v][k] = f; }
      logp += std::log(f);
    }
    const double w = (double)n_t[(size_t)t] - sigma_global;         // :110-115
    prob_existing[(size_t)t] = (w > 0.0) ? w * std::exp(logp) : 0.0;
  }
  double logm = 0.0;
  for (int v = 0; v < d; ++v) {                                      // :40-69
    const ViewState& V = views[(size_t)v];
    double num = 0.0, den = V.alpha_v;
    int K_act = 0;
    for (int k = 0; k < V.K; ++k) {
      const int l = V.l_vk[(size_t)k];
      den += (double)l;
      if (l <= 0) continue;
      K_act++;
      const double w = (double)l - V.sigma_v;
      if (w > 0.0) num += w * compute_f_vk(v, k, i);
    }
    const double f_new = compute_f_vk_new(v, i);
    const double wn = V.alpha_v + (double)K_act * V.sigma_v;
    if (wn > 0.0) num += wn * f_new;
    logm += std::log(den > 0.0 ? num / den : f_new);
  }
  int T_ne = 0;
  for (int t = 0; t < T; ++t) T_ne += n_t[(size_t)t] > 0;
  const double wn = alpha_global + sigma_global * (double)T_ne;      // :124-135
  prob_new = (wn > 0.0) ? wn * std::exp(logm) : 0.0;
}

void remove_customer(int) { not_on_device("remove_customer"); }
void add_customer_to_existing_table(int, int) { not_on_device("add_customer_to_existing_table"); }
int create_empty_table() { not_on_device("create_empty_table"); }
void add_customer_to_new_table(int, int) { not_on_device("add_customer_to_new_table"); }
int sample_dish_for_new_table(int, int) { not_on_device("sample_dish_for_new_table"); }
void assign_dishes_new_table(int, int) { not_on_device("assign_dishes_new_table"); }

// :291-303 on the mirrored state; saved_loglik (declared by the reference, multiview_state.h:38, never written there)
// receives the collapsed log marginal likelihood of the kept state.
void save_state() {
  saved_table_of.push_back(table_of);
  saved_dish_of.push_back(dish_of);
  saved_loglik.push_back(compute_log_likelihood());
  if ((int)saved_alpha_v.size() < d) { saved_alpha_v.resize((size_t)d); saved_sigma_v.resize((size_t)d); saved_tau_v.resize((size_t)d); }
  for (int v = 0; v < d; ++v) {
    saved_alpha_v[(size_t)v].push_back(views[(size_t)v].alpha_v);
    saved_sigma_v[(size_t)v].push_back(views[(size_t)v].sigma_v);
    saved_tau_v[(size_t)v].push_back(views[(size_t)v].tau_v);
  }
  saved_alpha_global.push_back(alpha_global);
  saved_sigma_global.push_back(sigma_global);
}

// :305-306 — the reference wraps R::runif / R::rnorm; here the call-ordered host Philox stream (the same numbers
// multiview_rng.h hands out; that header and this file define the same name, so — as in the reference — a translation
// unit includes one of the two).
namespace {
struct HostStream { unsigned long long seed = 1999ull, calls = 0; };
HostStream& host_stream() { static thread_local HostStream s; return s; }
}  // namespace
double uniform01() {
  HostStream& g = host_stream();
  return mvg_philox_uniform_f64(mvhost::seed, 0u, 6u, 0u, 0u, g.calls++);
}
double rnorm_scalar(double mean, double sd) {
  HostStream& g = host_stream();
  return mean + sd * mvg_philox_normal(mvhost::seed, 0u, 6u, 1u, 0u, g.calls++);
}
