// multiview_gibbs.h — chain entry points of the B200 sampler.
//
// Drop-in for /root/reference/Multiview/multiview_gibbs.h:8-13: the four declarations of the reference, unchanged in
// name and signature.  run_gibbs_cpp keeps its Rcpp signature and the eight names of its result list
// (multiview_gibbs.cpp:105-131), so New_Simulation.R:128-133 calls it as before; the sweep itself runs on the GPU
// behind the C ABI of include/mvg.h.
#pragma once
#ifndef MULTIVIEW_GIBBS_H
#define MULTIVIEW_GIBBS_H

#include <Rcpp.h>

// M sweeps of the device chain; after sweep `iter` with iter >= burn_in and (iter - burn_in) % thin == 0 the state is
// pulled from the GPU and appended to the saved_* traces (save rule of multiview_gibbs.cpp:205).  One sweep = the loop
// over all customers (:157-200) followed by the hyperparameter step (:202).
void gibbs_sampler(int M, int burn_in, int thin);

// The R entry point ([[Rcpp::export]] in the .cpp): data_views is a list of numeric vectors (one per view, the
// reference's scalar views) or of row-major flattened matrices when mvhost::view_dim is set.  Returns the list
// table_of, dish_of, loglik, alpha_v, sigma_v, tau_v, alpha_global, sigma_global.
Rcpp::List run_gibbs_cpp(const Rcpp::List& data_views,
                         int M, int burn_in, int thin);

// Joint log marginal likelihood of the data given the current partition.  The reference declares this function and
// never defines it (multiview_gibbs.h:13); here it is the sum over live dishes of log p(y_S) as
// multiview_utils.cpp:316-320 writes it.
double compute_log_likelihood();

// Reference initialisation (multiview_gibbs.cpp:12-103): T = 4 random tables, two random dishes per view, statistics,
// tau_v = 0.0025 Var(y_v) — performed on the device from the Philox initialisation domains.
static void initialize_state_from_data();

using namespace Rcpp;   // the reference's header exports the namespace to its includers; kept for them

#endif  // MULTIVIEW_GIBBS_H
