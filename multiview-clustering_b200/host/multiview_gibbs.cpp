// multiview_gibbs.cpp — Rcpp entry, initialisation and chain loop of the B200 sampler.
// Reference: /root/reference/Multiview/multiview_gibbs.cpp (run_gibbs_cpp :105-131,
// initialize_state_from_data :12-103, gibbs_sampler :134-212).  Built with sourceCpp / R CMD SHLIB and
// linked against the prebuilt libmvg_b200.so; every sweep runs on the GPU.
#include "multiview_gibbs.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/mvg.h"
#include "multiview_hyper.h"
#include "multiview_state.h"
#include "multiview_utils.h"

void mvhost_rcpp_stop(const std::string& msg) { Rcpp::stop(msg); }

// multiview_gibbs.cpp:12-103: T = 4 random tables, 2 random dishes per view, statistics rebuild,
// alpha_v = 1, sigma_v = .5, tau_v = 0.0025 Var(y_v), alpha_global = 1, sigma_global = .6 — on the
// device, from the Philox initialisation domains.
static void initialize_state_from_data() {
  mvhost::open_chain();
  if (mvg_init_state_reference(mvhost::chain()) != MVG_OK) mvhost::fail("initialize_state_from_data");
  mvhost::pull_state();
  saved_table_of.clear(); saved_dish_of.clear(); saved_loglik.clear();
  saved_alpha_v.assign((size_t)d, {}); saved_sigma_v.assign((size_t)d, {}); saved_tau_v.assign((size_t)d, {});
  saved_alpha_global.clear(); saved_sigma_global.clear();
}

// multiview_gibbs.cpp:134-212.  One mvg_sweep = the loop over all customers (:157-200) followed by
// update_hyperparameters (:202); launches are asynchronous, the device is only read on kept sweeps.
void gibbs_sampler(int M, int burn_in, int thin) {
  if (!mvhost::chain()) Rcpp::stop("gibbs_sampler: state not initialised");
  if (thin <= 0) Rcpp::stop("gibbs_sampler: thin must be positive");
  for (int iter = 0; iter < M; ++iter) {
    if ((iter + 1) % 100 == 0) Rcpp::Rcout << "Iteration " << (iter + 1) << " / " << M << "\n";   // :152-155
    if (mvg_sweep(mvhost::chain(), 1, 1) != MVG_OK) mvhost::fail("gibbs_sampler");
    if (iter >= burn_in && ((iter - burn_in) % thin == 0)) {                                       // :205
      mvhost::pull_state();
      save_state();                                  // multiview_utils.cpp:291-303 (+ saved_loglik)
    }
  }
  mvhost::pull_state();
}

// Declared by the reference (multiview_gibbs.h:13) but never defined there.  Here: the collapsed log
// marginal likelihood of the data given the current partition, sum over views and dishes of
// logp(n, S1, S2) as multiview_utils.cpp:307-338 writes it (prior mean 0, prior variance 1).
double compute_log_likelihood() {
  double total = 0.0;
  for (int v = 0; v < d; ++v) {
    const ViewState& V = views[(size_t)v];
    if ((size_t)v < mvhost::csr_views.size() && mvhost::csr_views[(size_t)v].vocab > 0) continue;   // Gaussian views only (as mvg_log_likelihood)
    const int D = mvhost::view_dim.empty() ? 1 : mvhost::view_dim[(size_t)v];
    const double tau = V.tau_v;
    for (int k = 0; k < V.K; ++k) {
      const double nk = (double)V.n_vk[(size_t)k];
      if (nk <= 0.0) continue;
      double s1sq = 0.0;
      for (int j = 0; j < D; ++j) s1sq += V.sum_y[(size_t)k * D + j] * V.sum_y[(size_t)k * D + j];
      total += (double)D * (-0.5 * nk * std::log(2.0 * M_PI * tau) - 0.5 * std::log(tau * (tau + nk)))
               - 0.5 * V.sum_y2[(size_t)k] / tau + 0.5 * s1sq / (tau * (tau + nk));
    }
  }
  return total;
}

// [[Rcpp::export]]
Rcpp::List run_gibbs_cpp(const Rcpp::List& data_views,
                         int M, int burn_in, int thin) {
  // data_views: per view a numeric vector (the reference's scalar views, multiview_gibbs.cpp:109-115), a numeric matrix
  // (customers x features) or a Matrix::dgCMatrix of counts (customers x vocabulary: the X_body / X_title / X_topics of
  // dataset/reuters/data pre-process.R:104-108, as sparse matrices).
  d = data_views.size();
  y.clear();
  y.resize((size_t)d);
  const std::vector<int> flat_dims = mvhost::view_dim;         // caller-declared dims of row-major FLATTENED matrices, if any
  mvhost::csr_views.assign((size_t)d, mvhost::CsrView{});
  std::vector<int> dims((size_t)d, 1);
  bool any_matrix = false;
  n = -1;
  for (int v = 0; v < d; ++v) {
    SEXP el = data_views[v];
    int rows = 0;
    if (Rf_isS4(el) && Rcpp::S4(el).is("dgCMatrix")) {
      // compressed sparse COLUMNS -> the compressed sparse ROWS the device takes (counting sort by row)
      Rcpp::S4 m(el);
      const Rcpp::IntegerVector mi = m.slot("i"), mp = m.slot("p"), dim = m.slot("Dim");
      const Rcpp::NumericVector mx = m.slot("x");
      rows = dim[0];
      mvhost::CsrView& cv = mvhost::csr_views[(size_t)v];
      cv.vocab = dim[1];
      cv.rowptr.assign((size_t)rows + 1, 0);
      for (size_t k = 0; k < (size_t)mi.size(); ++k) cv.rowptr[(size_t)mi[k] + 1] += 1;
      for (int r = 0; r < rows; ++r) cv.rowptr[(size_t)r + 1] += cv.rowptr[(size_t)r];
      cv.col.assign((size_t)mi.size(), 0);
      cv.val.assign((size_t)mi.size(), 0.f);
      std::vector<int> fill(cv.rowptr.begin(), cv.rowptr.end() - 1);
      for (int c = 0; c < dim[1]; ++c)
        for (int k = mp[c]; k < mp[c + 1]; ++k) {
          const int pos = fill[(size_t)mi[k]]++;
          cv.col[(size_t)pos] = c;
          cv.val[(size_t)pos] = (float)mx[k];
        }
      dims[(size_t)v] = 0;
      any_matrix = true;
    } else if (Rf_isMatrix(el)) {
      const Rcpp::NumericMatrix M(el);                         // column-major in R; the device takes row-major
      rows = M.nrow();
      dims[(size_t)v] = M.ncol();
      y[(size_t)v].resize((size_t)rows * M.ncol());
      for (int i = 0; i < rows; ++i)
        for (int j = 0; j < M.ncol(); ++j) y[(size_t)v][(size_t)i * M.ncol() + j] = M(i, j);
      any_matrix = true;
    } else {
      y[(size_t)v] = Rcpp::as<std::vector<double>>(el);
      const int D = flat_dims.empty() ? 1 : flat_dims[(size_t)v];
      rows = (int)y[(size_t)v].size() / D;
      dims[(size_t)v] = D;
      if (D != 1) any_matrix = true;
    }
    if (n < 0) n = rows;
    else if (rows != n) Rcpp::stop("run_gibbs_cpp: the views have different numbers of customers");
  }
  if (any_matrix) mvhost::view_dim = dims;

  if (mvhost::sequential) {
    // MVG_ENGINE_SEQ: the reference's sequential sampler rule for rule on the device (csrc/mv_seq_core.h); scalar views
    if (any_matrix) Rcpp::stop("run_gibbs_cpp: the sequential engine takes scalar views");
    if (thin <= 0) Rcpp::stop("run_gibbs_cpp: thin must be positive");
    const int S = (M > burn_in) ? (M - burn_in + thin - 1) / thin : 0, t_cap = 1024;
    std::vector<double> flat((size_t)d * n);
    for (int v = 0; v < d; ++v) std::copy(y[(size_t)v].begin(), y[(size_t)v].end(), flat.begin() + (size_t)v * n);
    std::vector<int32_t> tab((size_t)std::max(S, 1) * n), Ts((size_t)std::max(S, 1)), dish((size_t)std::max(S, 1) * d * t_cap);
    std::vector<double> hyp((size_t)std::max(S, 1) * (3 * d + 2));
    int32_t ns = 0;
    if (mvg_seq_run(0, n, d, flat.data(), M, burn_in, thin, mvhost::seed, t_cap, 2 * M + 64, S, tab.data(), Ts.data(), dish.data(),
                    hyp.data(), &ns, nullptr) != MVG_OK)
      Rcpp::stop(std::string("run_gibbs_cpp (sequential engine): ") + mvg_seq_last_error());
    saved_table_of.clear(); saved_dish_of.clear(); saved_loglik.clear();
    saved_alpha_v.assign((size_t)d, {}); saved_sigma_v.assign((size_t)d, {}); saved_tau_v.assign((size_t)d, {});
    saved_alpha_global.clear(); saved_sigma_global.clear();
    for (int s = 0; s < ns; ++s) {
      saved_table_of.emplace_back(tab.begin() + (size_t)s * n, tab.begin() + (size_t)(s + 1) * n);
      std::vector<std::vector<int>> dd((size_t)d);
      for (int v = 0; v < d; ++v) {
        const int32_t* row = dish.data() + ((size_t)s * d + v) * t_cap;
        dd[(size_t)v].assign(row, row + Ts[(size_t)s]);
      }
      saved_dish_of.push_back(dd);
      const double* h = hyp.data() + (size_t)s * (3 * d + 2);
      for (int v = 0; v < d; ++v) {
        saved_alpha_v[(size_t)v].push_back(h[v]);
        saved_sigma_v[(size_t)v].push_back(h[d + v]);
        saved_tau_v[(size_t)v].push_back(h[2 * d + v]);
      }
      saved_alpha_global.push_back(h[3 * d]);
      saved_sigma_global.push_back(h[3 * d + 1]);
    }
  } else {
    initialize_state_from_data();
    gibbs_sampler(M, burn_in, thin);
    mvhost::close_chain();
  }

  mvhost::view_dim = flat_dims;                                // (the caller's setting, not what this call derived)
  return Rcpp::List::create(
      Rcpp::Named("table_of") = saved_table_of,
      Rcpp::Named("dish_of") = saved_dish_of,
      Rcpp::Named("loglik") = saved_loglik,
      Rcpp::Named("alpha_v") = saved_alpha_v,
      Rcpp::Named("sigma_v") = saved_sigma_v,
      Rcpp::Named("tau_v") = saved_tau_v,
      Rcpp::Named("alpha_global") = saved_alpha_global,
      Rcpp::Named("sigma_global") = saved_sigma_global);
}
