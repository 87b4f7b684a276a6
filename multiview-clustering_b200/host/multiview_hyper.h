// multiview_hyper.h — hyperparameter step of the B200 sampler, with the declarations of the reference's
// header (/root/reference/Multiview/multiview_hyper.h:9-22).
//
// update_hyperparameters() and update_tau_v_MH() run ON THE DEVICE (one small kernel, Philox numbers
// addressed by the position the reference's sequential code would draw them at); there is no host
// implementation of the Metropolis-Hastings updates.  The remaining functions are pure inspectors of the
// mirrored state (multiview_state.h), written from the formulas of multiview_hyper.cpp so that code
// calling them keeps working: log EPPF, the priors, the tau log-posterior, a tau proposal.
#ifndef MULTIVIEW_HYPER_H
#define MULTIVIEW_HYPER_H

#include <cmath>
#include <vector>

#include "multiview_state.h"

extern double alpha_global;
extern double sigma_global;

void initialize_hyperparameters();

double propose_tau(double tau_old);
double log_posterior_given_tau(int v, double tau_candidate);
void update_tau_v_MH();

void update_hyperparameters();

double log_EPPF(int v, double alpha, double sigma);
double log_prior_alpha(double alpha);
double log_prior_sigma(double sigma);

#endif
