// multiview_hyper.cpp — see multiview_hyper.h.  Reference: /root/reference/Multiview/multiview_hyper.cpp.
#include "multiview_hyper.h"

#include <limits>

#include "../../include/mvg.h"
#include "multiview_utils.h"   // rnorm_scalar, uniform01 (as multiview_hyper.cpp:10 of the reference)

namespace {
constexpr double kEps = 1e-6;                        // multiview_hyper.cpp:13
constexpr double kNegInf = -std::numeric_limits<double>::infinity();
int dim_of(int v) {
  if ((size_t)v < mvhost::csr_views.size() && mvhost::csr_views[(size_t)v].vocab > 0) return 0;   // a count view has no Gaussian coordinates
  return mvhost::view_dim.empty() ? 1 : mvhost::view_dim[(size_t)v];
}
}  // namespace

// multiview_hyper.cpp:137-163: the literals every chain starts from; on the device they are set by
// mvg_init_state_reference together with tau_v = 0.0025 Var(y_v) (multiview_gibbs.cpp:75-98).
void initialize_hyperparameters() {
  views.resize((size_t)d);
  for (ViewState& V : views) { V.alpha_v = 1.0; V.sigma_v = 0.5; V.tau_v = 1.0; }
  alpha_global = 1.0;
  sigma_global = 0.6;
}

// :166-174 — log-normal random walk with step 0.3 (host stream; the device draws its own)
double propose_tau(double tau_old) {
  if (tau_old <= 0.0) tau_old = kEps;
  return std::exp(std::log(tau_old) + rnorm_scalar(0.0, 0.3));
}

// :176-209 — Gaussian part from the within-dish sums of squares, InvGamma(2, 1) prior
double log_posterior_given_tau(int v, double tau) {
  if (tau <= 0.0) return kNegInf;
  const ViewState& V = views[(size_t)v];
  const int D = dim_of(v);
  const double lg = std::log(2.0 * M_PI * tau);
  double loglik = 0.0;
  for (int k = 0; k < V.K; ++k) {
    const int nk = V.n_vk[(size_t)k];
    if (nk == 0) continue;
    double s1sq = 0.0;
    for (int j = 0; j < D; ++j) s1sq += V.sum_y[(size_t)k * D + j] * V.sum_y[(size_t)k * D + j];
    double sse = V.sum_y2[(size_t)k] - s1sq / (double)nk;
    if (sse < 0.0) sse = 0.0;
    loglik += -0.5 * (double)nk * (double)D * lg - 0.5 * (sse / tau);
  }
  const double a_tau = 2.0, b_tau = 1.0;             // :133-134
  return loglik + (a_tau * std::log(b_tau) - std::lgamma(a_tau) - (a_tau + 1.0) * std::log(tau) - b_tau / tau);
}

// :211-231 on the device, then the mirror is refreshed
void update_tau_v_MH() {
  if (!mvhost::chain()) return;
  if (mvg_hyper_step_parts(mvhost::chain(), MVG_HYPER_TAU) != MVG_OK) mvhost::fail("update_tau_v_MH");
  mvhost::pull_state();
}

// :233-292 on the device (tau_v, then alpha_v/sigma_v per view, then the franchise pair)
void update_hyperparameters() {
  if (views.empty()) initialize_hyperparameters();   // :234-235
  if (!mvhost::chain()) return;
  if (mvg_hyper_step(mvhost::chain()) != MVG_OK) mvhost::fail("update_hyperparameters");
  mvhost::pull_state();
}

// :295-342 — Pitman-Yor EPPF of the tables over the dishes of view v
double log_EPPF(int v, double alpha, double sigma) {
  if (!(sigma > kEps && sigma < 1.0 - kEps) || alpha <= -sigma) return kNegInf;
  const ViewState& V = views[(size_t)v];
  long total = 0;
  int live = 0;
  for (int k = 0; k < V.K; ++k) { total += V.l_vk[(size_t)k]; live += V.l_vk[(size_t)k] > 0; }
  if (total <= 0) return 0.0;
  double lp = 0.0;
  for (int j = 0; j < live; ++j) {
    const double term = alpha + (double)j * sigma;
    if (term <= 0.0) return kNegInf;
    lp += std::log(term);
  }
  for (long i = 1; i < total; ++i) lp -= std::log(alpha + (double)i);
  for (int k = 0; k < V.K; ++k)
    for (int m = 1; m < V.l_vk[(size_t)k]; ++m) lp += std::log((double)m - sigma);
  return lp;
}

double log_prior_alpha(double alpha) {               // :344-351, Gamma(4, 3)
  return (alpha <= 0.0) ? kNegInf : (4.0 - 1.0) * std::log(alpha) - 3.0 * alpha;
}
double log_prior_sigma(double sigma) {               // :353-360, Beta(1, 5)
  return (sigma <= 0.0 || sigma >= 1.0) ? kNegInf : (1.0 - 1.0) * std::log(sigma) + (5.0 - 1.0) * std::log(1.0 - sigma);
}
