// tests/seq_host_check.cpp — TEST scaffolding: the sequential engine's algorithm (csrc/mv_seq_core.h) compiled for the
// HOST, so that its logic can be compared with the compiled reference on a box without a GPU.  The product runs the same
// source on the device (csrc/mv_seq.cu); nothing here is product code.
#define MV_SEQ_FN inline
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mv_seq_core.h"

extern "C" int seq_host_run(int n, int d, const double* y, int M, int burn_in, int thin, unsigned long long seed, int t_cap,
                            int k_cap, int n_saved_max, int* saved_table_of, int* saved_T, int* saved_dish_of,
                            double* saved_hypers, unsigned long long* calls) {
  using mv::seq::State;
  State s{};
  s.n = n; s.d = d; s.t_cap = t_cap; s.k_cap = k_cap; s.seed = seed; s.calls = 0; s.err = 0; s.y = y;
  std::vector<int> table_of(n), n_t(t_cap), dish_of((size_t)d * t_cap), K(d), n_vk((size_t)d * k_cap), l_vk((size_t)d * k_cap), cand(k_cap);
  std::vector<double> sy((size_t)d * k_cap), sy2((size_t)d * k_cap), av(d), sv(d), tv(d), prob(t_cap), wts(k_cap + 1);
  s.table_of = table_of.data(); s.n_t = n_t.data(); s.dish_of = dish_of.data(); s.K = K.data(); s.n_vk = n_vk.data();
  s.l_vk = l_vk.data(); s.sum_y = sy.data(); s.sum_y2 = sy2.data(); s.alpha_v = av.data(); s.sigma_v = sv.data();
  s.tau_v = tv.data(); s.prob = prob.data(); s.wts = wts.data(); s.cand = cand.data();
  mv::seq::start(s);
  int saved = 0;
  for (int iter = 0; iter < M; ++iter) {
    mv::seq::sweep(s);
    if (s.err) return -s.err;
    if (iter >= burn_in && ((iter - burn_in) % thin == 0)) {
      if (saved >= n_saved_max) return -100;
      std::memcpy(saved_table_of + (size_t)saved * n, s.table_of, sizeof(int) * n);
      saved_T[saved] = s.T;
      std::memcpy(saved_dish_of + (size_t)saved * d * t_cap, s.dish_of, sizeof(int) * (size_t)d * t_cap);
      double* o = saved_hypers + (size_t)saved * (3 * d + 2);
      for (int v = 0; v < d; ++v) { o[v] = s.alpha_v[v]; o[d + v] = s.sigma_v[v]; o[2 * d + v] = s.tau_v[v]; }
      o[3 * d] = s.alpha_g; o[3 * d + 1] = s.sigma_g;
      ++saved;
    }
  }
  if (calls) *calls = s.calls;
  return saved;
}
