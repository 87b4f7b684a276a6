// multiview_hyper.h — hyperparameter step of the B200 sampler.
//
// Drop-in for /root/reference/Multiview/multiview_hyper.h:9-22: every function the reference declares there is
// declared here with the same name, argument list and meaning, so translation units written against the
// reference keep compiling.  What differs is where the work happens:
//
//   * The Metropolis-Hastings updates themselves — update_hyperparameters() and update_tau_v_MH() — run ON THE DEVICE
//     inside k_finalize (csrc/mv_state_kernels.cu), with Philox numbers addressed by the position the reference's
//     sequential code would draw them at.  The host functions below only trigger them through the C ABI
//     (mvg_hyper_step / mvg_hyper_step_parts) and refresh the mirrored state; there is no host implementation.
//   * The remaining functions are pure inspectors of the mirrored state (multiview_state.h), written from the
//     formulas of multiview_hyper.cpp: log EPPF, the two priors, the tau log-posterior, a tau proposal.
#pragma once
#ifndef MULTIVIEW_HYPER_H
#define MULTIVIEW_HYPER_H

#include "multiview_state.h"   // alpha_global, sigma_global, views[v].{alpha_v, sigma_v, tau_v}

#include <cmath>
#include <vector>

// ---- device-side steps (thin triggers) ------------------------------------------------------------------------

// One full hyperparameter step on the device: tau_v for every view, then (alpha_v, sigma_v) per view, then the
// franchise pair (alpha_global, sigma_global) — the order of multiview_hyper.cpp:233-292.  Called once per sweep by
// the reference (multiview_gibbs.cpp:202); here mvg_sweep already includes it, this entry exists for callers that
// drive the pieces themselves.
void update_hyperparameters();

// Only the kernel-variance part of the step, update_tau_v_MH of multiview_hyper.cpp:211-231.
void update_tau_v_MH();

// The reference's literals (alpha = 1, sigma = .5, alpha_global = 1, sigma_global = .6; multiview_gibbs.cpp:75-98)
// written into the mirrored state; the device applies the same ones in mvg_init_state_reference.
void initialize_hyperparameters();

// ---- inspectors of the mirrored state ---------------------------------------------------------------------------

// log EPPF of the tables-per-dish partition of view v under (alpha, sigma): multiview_hyper.cpp:295-342.
double log_EPPF(int v, double alpha, double sigma);

// log of the Gamma(4, 3) prior on a concentration and of the Beta(1, 5) prior on a discount (:344-360).
double log_prior_alpha(double alpha);
double log_prior_sigma(double sigma);

// Unnormalised log posterior of the kernel variance of view v at tau_candidate (:176-209), from the mirrored
// per-dish statistics n_vk, sum_y, sum_y2.
double log_posterior_given_tau(int v, double tau_candidate);

// A log-normal random-walk proposal around tau_old with the reference's step 0.3 (:166-174), on the host stream of
// multiview_rng.h.
double propose_tau(double tau_old);

// franchise-level hyperparameters (defined with the rest of the mirrored state)
extern double alpha_global;
extern double sigma_global;

#endif  // MULTIVIEW_HYPER_H
