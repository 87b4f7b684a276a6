// multiview_gibbs.h — chain entry points of the B200 sampler with the reference's declarations
// (/root/reference/Multiview/multiview_gibbs.h:8-13).  run_gibbs_cpp keeps its Rcpp signature and the
// eight names of its result list (multiview_gibbs.cpp:105-131), so New_Simulation.R:128-133 calls it
// unchanged; the sweep itself runs on the GPU behind the C ABI of include/mvg.h.
#ifndef MULTIVIEW_GIBBS_H
#define MULTIVIEW_GIBBS_H

#include <Rcpp.h>
using namespace Rcpp;

Rcpp::List run_gibbs_cpp(const Rcpp::List& data_views,
                         int M, int burn_in, int thin);

void gibbs_sampler(int M, int burn_in, int thin);
static void initialize_state_from_data();
double compute_log_likelihood();

#endif
