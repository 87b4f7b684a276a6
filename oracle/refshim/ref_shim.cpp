// oracle/refshim/ref_shim.cpp — TEST INFRASTRUCTURE, not product code.
//
// extern "C" access to the UNMODIFIED reference sampler, for ctypes.  This file is linked
// with /root/reference/Multiview/multiview_{gibbs,utils,hyper,state}.cpp (compiled where
// they lie, see oracle/Makefile) into oracle/_ref/libmvref.so.  Nothing here restates the
// reference's arithmetic: every number returned is produced by the reference's own
// functions operating on the reference's own globals (multiview_state.h:21-45).
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load the result.
#include <Rcpp.h>

#include <cstdint>
#include <cstring>
#include <deque>
#include <sstream>

#include "multiview_gibbs.h"
#include "multiview_hyper.h"
#include "multiview_state.h"
#include "multiview_utils.h"

extern "C" {
#include "../mv_philox_ref.h"
}

// Has external linkage in multiview_utils.cpp:40 but is not declared in multiview_utils.h.
double compute_marginal_likelihood_new_table(int v, int i);

// ---------------------------------------------------------------------------------------
// RNG behind R::runif / R::rnorm
// ---------------------------------------------------------------------------------------
namespace {
uint64_t g_seed = 1999;          // echoes set.seed(1999), New_Simulation.R:12
uint64_t g_calls = 0;            // call-ordered Philox counter (domain 6)
std::deque<double> g_scripted;   // when non-empty, uniforms are popped from here first
uint64_t g_n_unif = 0, g_n_norm = 0;
std::string g_last_error;

class NullBuf : public std::streambuf {
  int overflow(int c) override { return c; }
};
NullBuf g_nullbuf;
}  // namespace

namespace Rcpp {
std::ostream Rcout(&g_nullbuf);
}

namespace R {
double runif(double a, double b) {
  ++g_n_unif;
  double u;
  if (!g_scripted.empty()) {
    u = g_scripted.front();
    g_scripted.pop_front();
  } else {
    u = mvo_uniform53(g_seed, 0, MVO_DOM_CALLSEQ, 0, 0, g_calls++);
  }
  return a + (b - a) * u;
}
double rnorm(double mean, double sd) {
  ++g_n_norm;
  double z = mvo_normal(g_seed, 0, MVO_DOM_CALLSEQ, 1, 0, g_calls++);
  return mean + sd * z;
}
}  // namespace R

// ---------------------------------------------------------------------------------------
// C surface
// ---------------------------------------------------------------------------------------
extern "C" {

const char* ref_last_error() { return g_last_error.c_str(); }

void ref_set_seed(uint64_t seed) {
  g_seed = seed;
  g_calls = 0;
  g_scripted.clear();
  g_n_unif = g_n_norm = 0;
}
void ref_push_uniforms(const double* u, int count) {
  for (int j = 0; j < count; ++j) g_scripted.push_back(u[j]);
}
int ref_scripted_left() { return static_cast<int>(g_scripted.size()); }
uint64_t ref_uniform_calls() { return g_n_unif; }
uint64_t ref_normal_calls() { return g_n_norm; }

// Load a complete sampler state into the reference's globals.  table_in[i] may be -1
// (customer unseated).  Membership lists and sufficient statistics are built the way the
// reference's own init loop builds them (multiview_gibbs.cpp:24-33, :64-73).
// yflat is [d][n]; dish_in is [d][T_in]; K_in[v] is the number of dish slots of view v.
int ref_load_state(int n_in, int d_in, const double* yflat, const int* table_in, int T_in,
                   const int* dish_in, const int* K_in, const double* alpha_v_in,
                   const double* sigma_v_in, const double* tau_v_in, double alpha_g,
                   double sigma_g) {
  n = n_in;
  d = d_in;
  y.assign(d, std::vector<double>());
  for (int v = 0; v < d; ++v) y[v].assign(yflat + (size_t)v * n, yflat + (size_t)(v + 1) * n);
  T = T_in;
  table_of.assign(table_in, table_in + n);
  n_t.assign(T, 0);
  customers_at_table.assign(T, std::vector<int>());
  for (int i = 0; i < n; ++i) {
    int t = table_of[i];
    if (t < 0) continue;
    if (t >= T) { g_last_error = "ref_load_state: table index out of range"; return 1; }
    customers_at_table[t].push_back(i);
    n_t[t]++;
  }
  dish_of.assign(d, std::vector<int>());
  views.assign(d, ViewState());
  for (int v = 0; v < d; ++v) {
    dish_of[v].assign(dish_in + (size_t)v * T, dish_in + (size_t)(v + 1) * T);
    ViewState& V = views[v];
    V.K = K_in[v];
    V.n_vk.assign(V.K, 0);
    V.l_vk.assign(V.K, 0);
    V.sum_y.assign(V.K, 0.0);
    V.sum_y2.assign(V.K, 0.0);
    V.customers_at_dish.assign(V.K, std::vector<int>());
    for (int t = 0; t < T; ++t) {
      int k = dish_of[v][t];
      if (k < 0 || k >= V.K) { g_last_error = "ref_load_state: dish index out of range"; return 2; }
      V.l_vk[k]++;
    }
    for (int i = 0; i < n; ++i) {
      int t = table_of[i];
      if (t < 0) continue;
      int k = dish_of[v][t];
      double val = y[v][i];
      V.n_vk[k]++;
      V.sum_y[k] += val;
      V.sum_y2[k] += val * val;
      V.customers_at_dish[k].push_back(i);
    }
    V.alpha_v = alpha_v_in[v];
    V.sigma_v = sigma_v_in[v];
    V.tau_v = tau_v_in[v];
  }
  alpha_global = alpha_g;
  sigma_global = sigma_g;
  return 0;
}

int ref_get_dims(int* n_out, int* d_out, int* T_out) {
  *n_out = n; *d_out = d; *T_out = T;
  return 0;
}
int ref_get_K(int v) { return views[v].K; }
void ref_get_tables(int* table_out, int* n_t_out) {
  std::memcpy(table_out, table_of.data(), sizeof(int) * (size_t)n);
  if (T > 0) std::memcpy(n_t_out, n_t.data(), sizeof(int) * (size_t)T);
}
void ref_get_dish_of(int v, int* out) {
  if (T > 0) std::memcpy(out, dish_of[v].data(), sizeof(int) * (size_t)T);
}
void ref_get_view_stats(int v, int* n_vk_out, int* l_vk_out, double* sum_y_out, double* sum_y2_out) {
  const ViewState& V = views[v];
  for (int k = 0; k < V.K; ++k) {
    n_vk_out[k] = V.n_vk[k];
    l_vk_out[k] = V.l_vk[k];
    sum_y_out[k] = V.sum_y[k];
    sum_y2_out[k] = V.sum_y2[k];
  }
}
void ref_get_hypers(double* alpha_v_out, double* sigma_v_out, double* tau_v_out, double* global2) {
  for (int v = 0; v < d; ++v) {
    alpha_v_out[v] = views[v].alpha_v;
    sigma_v_out[v] = views[v].sigma_v;
    tau_v_out[v] = views[v].tau_v;
  }
  global2[0] = alpha_global;
  global2[1] = sigma_global;
}

// --- likelihood kernels (multiview_utils.cpp:307-350, :40-69) ---
double ref_compute_f_vk(int v, int k, int i) { return compute_f_vk(v, k, i); }
double ref_compute_f_vk_new(int v, int i) { return compute_f_vk_new(v, i); }
double ref_marginal_new_table(int v, int i) { return compute_marginal_likelihood_new_table(v, i); }

// --- table weights (multiview_utils.cpp:71-136); customer i must be unseated ---
int ref_table_probs(int i, double* prob_existing_out, double* prob_new_out) {
  try {
    std::vector<double> pe(T, 0.0);
    double pn = 0.0;
    std::vector<std::unordered_map<int, double>> cache(d);
    compute_table_probs_with_cache(i, pe, pn, cache);
    for (int t = 0; t < T; ++t) prob_existing_out[t] = pe[t];
    *prob_new_out = pn;
    return 0;
  } catch (const std::exception& e) { g_last_error = e.what(); return 1; }
}

int ref_remove_customer(int i) {
  try { remove_customer(i); return 0; }
  catch (const std::exception& e) { g_last_error = e.what(); return 1; }
}
void ref_add_customer_to_existing_table(int i, int t) { add_customer_to_existing_table(i, t); }
int ref_seat_at_new_table(int i) {  // multiview_gibbs.cpp:193-196
  int t_new = create_empty_table();
  add_customer_to_new_table(i, t_new);
  assign_dishes_new_table(i, t_new);
  return t_new;
}
int ref_sample_dish_for_new_table(int v, int i) { return sample_dish_for_new_table(v, i); }

// --- hyperparameter step (multiview_hyper.cpp) ---
double ref_log_EPPF(int v, double a, double s) { return log_EPPF(v, a, s); }
double ref_log_prior_alpha(double a) { return log_prior_alpha(a); }
double ref_log_prior_sigma(double s) { return log_prior_sigma(s); }
double ref_log_posterior_given_tau(int v, double tau) { return log_posterior_given_tau(v, tau); }
void ref_update_hyperparameters() { update_hyperparameters(); }

// --- whole chain through the reference's own entry point (multiview_gibbs.cpp:105-131) ---
// Returns the number of saved states; fetch them with ref_saved_*.  M = 0 runs only the
// reference's random initialisation (initialize_state_from_data is file-static).
int ref_run_gibbs(int n_in, int d_in, const double* yflat, int M, int burn_in, int thin) {
  try {
    Rcpp::List views_in;
    for (int v = 0; v < d_in; ++v)
      views_in.push_back(std::vector<double>(yflat + (size_t)v * n_in, yflat + (size_t)(v + 1) * n_in));
    run_gibbs_cpp(views_in, M, burn_in, thin);
    return static_cast<int>(saved_table_of.size());
  } catch (const std::exception& e) { g_last_error = e.what(); return -1; }
}
// Continue the current chain for M more sweeps without re-initialising.
int ref_gibbs_sampler(int M, int burn_in, int thin) {
  try { gibbs_sampler(M, burn_in, thin); return static_cast<int>(saved_table_of.size()); }
  catch (const std::exception& e) { g_last_error = e.what(); return -1; }
}
int ref_saved_count() { return static_cast<int>(saved_table_of.size()); }
int ref_saved_T(int s) { return static_cast<int>(saved_dish_of[s].empty() ? 0 : saved_dish_of[s][0].size()); }
void ref_saved_table_of(int s, int* out) {
  std::memcpy(out, saved_table_of[s].data(), sizeof(int) * saved_table_of[s].size());
}
void ref_saved_dish_of(int s, int v, int* out) {
  std::memcpy(out, saved_dish_of[s][v].data(), sizeof(int) * saved_dish_of[s][v].size());
}
void ref_saved_hypers(int s, double* alpha_v_out, double* sigma_v_out, double* tau_v_out, double* global2) {
  for (int v = 0; v < d; ++v) {
    alpha_v_out[v] = saved_alpha_v[v][s];
    sigma_v_out[v] = saved_sigma_v[v][s];
    tau_v_out[v] = saved_tau_v[v][s];
  }
  global2[0] = saved_alpha_global[s];
  global2[1] = saved_sigma_global[s];
}

}  // extern "C"
