// tests/multiview_utils_decl.h — TEST scaffolding: multiview_rng.h and multiview_utils.h both declare uniform01 (as in
// the reference, where the two headers cannot share a translation unit), so host_shim.cpp reaches the utilities through
// these forwarders, defined in host_shim_utils.cpp which includes multiview_utils.h alone.
#pragma once
#include <unordered_map>
#include <vector>
namespace mvu {
void table_probs(int i, std::vector<double>& pe, double& pn, std::vector<std::unordered_map<int, double>>& cache);
double f_vk(int v, int k, int i);
double f_vk_new(int v, int i);
void remove(int i);
void add_existing(int i, int t);
int new_table();
void assign_dishes(int i, int t);
}  // namespace mvu
