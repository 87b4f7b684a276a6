// multiview_state.cpp — definitions of the mirrored chain state and the device chain behind it
// (reference: /root/reference/Multiview/multiview_state.cpp:4-27 holds the same globals).
#include "multiview_state.h"

#include <algorithm>
#include <stdexcept>
#include <string>

#include "../../include/mvg.h"

int n = 0, d = 0;
std::vector<std::vector<double>> y;
double alpha_global = 1.0;
double sigma_global = 0.5;
int T = 0;
std::vector<int> table_of;
std::vector<int> n_t;
std::vector<std::vector<int>> customers_at_table;
std::vector<std::vector<int>> dish_of;
std::vector<ViewState> views;
std::vector<std::vector<int>> saved_table_of;
std::vector<std::vector<std::vector<int>>> saved_dish_of;
std::vector<double> saved_loglik;
std::vector<std::vector<double>> saved_alpha_v;
std::vector<std::vector<double>> saved_sigma_v;
std::vector<std::vector<double>> saved_tau_v;
std::vector<double> saved_alpha_global;
std::vector<double> saved_sigma_global;

#ifdef MVHOST_WITH_RCPP
void mvhost_rcpp_stop(const std::string& msg);   // multiview_gibbs.cpp: Rcpp::stop, i.e. an R error
#endif

namespace mvhost {

int table_capacity = 64;
unsigned long long seed = 1999ull;
int engine = MVG_ENGINE_AUTO;
std::vector<int> view_dim;
std::vector<CsrView> csr_views;
bool sequential = false;
static mvg_handle* g_chain = nullptr;

mvg_handle* chain() { return g_chain; }

void fail(const char* where) {
  std::string msg = std::string(where) + ": " + mvg_last_error(g_chain);
#ifdef MVHOST_WITH_RCPP
  mvhost_rcpp_stop(msg);
#endif
  throw std::runtime_error(msg);
}

static bool is_counts(int v) { return (size_t)v < csr_views.size() && csr_views[(size_t)v].vocab > 0; }
static int dim_of(int v) { return is_counts(v) ? 0 : (view_dim.empty() ? 1 : view_dim[(size_t)v]); }

void close_chain() {
  if (g_chain) mvg_destroy(g_chain);
  g_chain = nullptr;
}

void open_chain() {
  close_chain();
  mvg_config cfg{};
  cfg.abi_version = MVG_ABI_VERSION;
  cfg.device = 0;
  cfg.n_rows = n;
  cfg.n_rows_global = n;
  cfg.row_offset = 0;
  cfg.n_views = d;
  cfg.cap = table_capacity;
  cfg.seed = seed;
  cfg.chain = 0;
  cfg.engine = engine;
  cfg.rank = 0;
  cfg.world = 1;
  if (mvg_create(&cfg, &g_chain) != MVG_OK) fail("mvg_create");
  for (int v = 0; v < d; ++v) {
    if (is_counts(v)) {
      const CsrView& cv = csr_views[(size_t)v];
      if (mvg_upload_view_csr(g_chain, v, cv.rowptr.data(), cv.col.data(), cv.val.data(), (int64_t)cv.col.size(), cv.vocab) != MVG_OK)
        fail("mvg_upload_view_csr");
    } else if (mvg_upload_view_f64(g_chain, v, y[(size_t)v].data(), dim_of(v)) != MVG_OK) {
      fail("mvg_upload_view_f64");
    }
  }
}

void pull_state() {
  if (!g_chain) throw std::runtime_error("pull_state: no device chain");
  const int cap = table_capacity;
  int dsum = 0;
  for (int v = 0; v < d; ++v) dsum += dim_of(v);
  std::vector<int32_t> tab((size_t)n), nt((size_t)cap), dish((size_t)d * cap), nvk((size_t)d * cap), lvk((size_t)d * cap);
  std::vector<double> s1((size_t)cap * dsum), s2((size_t)d * cap), av((size_t)d), sv((size_t)d), tv((size_t)d);
  double ag[2] = {0, 0};
  mvg_state_host st{};
  st.table_of = tab.data(); st.n_t = nt.data(); st.dish_of = dish.data(); st.n_vk = nvk.data(); st.l_vk = lvk.data();
  st.sum_y = s1.data(); st.sum_y2 = s2.data(); st.alpha_v = av.data(); st.sigma_v = sv.data(); st.tau_v = tv.data();
  st.alpha_sigma_global = ag;
  if (mvg_get_state(g_chain, &st) != MVG_OK) fail("mvg_get_state");

  // compact the table slots
  std::vector<int> tmap((size_t)cap, -1);
  T = 0;
  for (int t = 0; t < cap; ++t) if (nt[(size_t)t] > 0) tmap[(size_t)t] = T++;
  table_of.assign((size_t)n, 0);
  n_t.assign((size_t)T, 0);
  customers_at_table.assign((size_t)T, {});
  for (int i = 0; i < n; ++i) {
    const int t = tmap[(size_t)tab[(size_t)i]];
    table_of[(size_t)i] = t;
    n_t[(size_t)t] += 1;
    customers_at_table[(size_t)t].push_back(i);
  }
  dish_of.assign((size_t)d, std::vector<int>((size_t)T, 0));
  views.resize((size_t)d);
  int doff = 0;
  for (int v = 0; v < d; ++v) {
    const int D = dim_of(v);
    std::vector<int> kmap((size_t)cap, -1);
    int K = 0;
    for (int k = 0; k < cap; ++k) if (lvk[(size_t)v * cap + k] > 0) kmap[(size_t)k] = K++;
    ViewState& V = views[(size_t)v];
    V.K = K;
    V.n_vk.assign((size_t)K, 0); V.l_vk.assign((size_t)K, 0);
    V.sum_y.assign((size_t)K * D, 0.0); V.sum_y2.assign((size_t)K, 0.0);
    V.customers_at_dish.assign((size_t)K, {});
    for (int k = 0; k < cap; ++k) {
      const int kk = kmap[(size_t)k];
      if (kk < 0) continue;
      V.n_vk[(size_t)kk] = nvk[(size_t)v * cap + k];
      V.l_vk[(size_t)kk] = lvk[(size_t)v * cap + k];
      V.sum_y2[(size_t)kk] = s2[(size_t)v * cap + k];
      for (int j = 0; j < D; ++j) V.sum_y[(size_t)kk * D + j] = s1[(size_t)cap * doff + (size_t)k * D + j];
    }
    for (int t = 0; t < cap; ++t)
      if (tmap[(size_t)t] >= 0) dish_of[(size_t)v][(size_t)tmap[(size_t)t]] = kmap[(size_t)dish[(size_t)v * cap + t]];
    for (int i = 0; i < n; ++i) V.customers_at_dish[(size_t)dish_of[(size_t)v][(size_t)table_of[(size_t)i]]].push_back(i);
    V.alpha_v = av[(size_t)v]; V.sigma_v = sv[(size_t)v]; V.tau_v = tv[(size_t)v];
    doff += D;
  }
  alpha_global = ag[0];
  sigma_global = ag[1];
}

}  // namespace mvhost
