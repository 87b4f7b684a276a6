// multiview_utils.h — the sampling utilities of the reference's interface, over the device chain.
//
// Drop-in for /root/reference/Multiview/multiview_utils.h:11-37: the thirteen declarations the reference's other
// translation units include (multiview_gibbs.cpp:5-8, multiview_hyper.cpp:10), with the same names and argument lists.
// On the B200 build the sweep itself — remove / score / draw / insert for every customer — runs on the device behind
// mvg_sweep (include/mvg.h), so these functions divide into
//
//   inspectors   compute_f_vk, compute_f_vk_new, compute_table_probs_with_cache: the reference's arithmetic
//                (multiview_utils.cpp:40-136, :307-350) evaluated on the MIRRORED state of multiview_state.h as it stands
//                (the reference calls them after remove_customer(i); the device applies that removal as a leave-one-out
//                view inside the kernel).  D > 1 views use the per-coordinate product that reduces to the reference at D = 1.
//   bookkeeping  save_state (appends the mirror to the saved_* traces, filling saved_loglik, which the reference declares
//                and never writes: multiview_state.h:38), ensure_global_variances_calculated (a no-op, as in the reference
//                its result is unused: multiview_utils.cpp:16-35), uniform01 / rnorm_scalar (the host Philox stream).
//   mutators     remove_customer, add_customer_to_existing_table, create_empty_table, add_customer_to_new_table,
//                sample_dish_for_new_table, assign_dishes_new_table: single-customer edits of a chain that lives in HBM
//                are not offered; they raise Rcpp::stop (the reference's error convention, multiview_utils.cpp:141,145).
#pragma once
#ifndef MULTIVIEW_UTILS_H
#define MULTIVIEW_UTILS_H

#include <unordered_map>
#include <vector>

#include "multiview_state.h"

void ensure_global_variances_calculated();

double compute_f_vk(int v, int k, int i);          // posterior-predictive density of y[v][i] under dish k (:307-338)
double compute_f_vk_new(int v, int i);             // density under a new dish, N(y; 0, tau_v) (:340-350)

void compute_table_probs_with_cache(               // unnormalised table weights and the new-table weight (:71-136)
    int i,
    std::vector<double> &prob_existing,
    double &prob_new,
    std::vector<std::unordered_map<int, double>> &cache_fvk
);

void remove_customer(int i);
void add_customer_to_existing_table(int i, int t);

int create_empty_table();
void add_customer_to_new_table(int i, int t_new);

int sample_dish_for_new_table(int v, int i);
void assign_dishes_new_table(int i, int t_new);

void save_state();
double uniform01();
double rnorm_scalar(double mean, double sd);

#endif  // MULTIVIEW_UTILS_H
