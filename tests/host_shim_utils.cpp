// tests/host_shim_utils.cpp — TEST scaffolding, see multiview_utils_decl.h.
#include <Rcpp.h>

#include "multiview_utils.h"
#include "multiview_utils_decl.h"

namespace mvu {
void table_probs(int i, std::vector<double>& pe, double& pn, std::vector<std::unordered_map<int, double>>& cache) {
  compute_table_probs_with_cache(i, pe, pn, cache);
}
double f_vk(int v, int k, int i) { return compute_f_vk(v, k, i); }
double f_vk_new(int v, int i) { return compute_f_vk_new(v, i); }
void remove(int i) { remove_customer(i); }
void add_existing(int i, int t) { add_customer_to_existing_table(i, t); }
int new_table() { return create_empty_table(); }
void assign_dishes(int i, int t) { assign_dishes_new_table(i, t); }
}  // namespace mvu
