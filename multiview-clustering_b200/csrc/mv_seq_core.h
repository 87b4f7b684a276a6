// mv_seq_core.h — the SEQUENTIAL sampler of the reference, restated for one thread of control.
//
// MVG_ENGINE_SEQ is the reference-exact mode (SURVEY.md §7.3-1 "SEQ"): customers are re-seated one after the other,
// each against the state left by its predecessors, with unbounded table and dish slots, swap-with-last deletion of an
// emptied table and dish slots that are never recycled — every rule of
//   /root/reference/Multiview/multiview_gibbs.cpp:12-103 (start), :157-200 (sweep),
//   /root/reference/Multiview/multiview_utils.cpp:40-136 (weights), :138-289 (remove / add / births), :307-350 (densities),
//   /root/reference/Multiview/multiview_hyper.cpp:166-292 (hyper step), :53-83, :295-360 (EPPF, priors)
// in FP64, in the reference's operation order, on the call-ordered Philox stream the compiled reference is driven with
// (oracle/refshim: R::runif / R::rnorm over domain kDomCallSeq).  Given the same data and seed the chain visits the same
// integer states as the unmodified reference (tests/test_seq_engine.py: table_of and dish_of of every kept sweep
// identical, hyperparameters to 1e-9): the doubles differ only by the last bits of exp / log / cos between libm and the
// CUDA math library, which a draw or an acceptance test sees with probability ~1e-15 each.
//
// It exists for the reference's own configuration (config 1: N ~ 500, scalar views) and as the statistical anchor of
// the data-parallel engines; it is one thread per chain and not a performance path (≈ 200 sweeps/s at N = 500).
//
// Pure C++ (no containers, no CUDA intrinsics) so that the same source is compiled for the device by mv_seq.cu and —
// for logic checks on a host without a GPU, TEST ONLY — by tests/seq_host_check.cpp.  Compile without FMA contraction.
#pragma once
#include <math.h>
#include <stdint.h>

#include "mv_philox.h"

#ifndef MV_SEQ_FN
#define MV_SEQ_FN __host__ __device__ inline
#endif

namespace mv {
namespace seq {

constexpr double kEpsS = 1e-6;               // multiview_hyper.cpp:13
constexpr double kPiS = 3.14159265358979323846;

struct State {
  int n, d;                 // customers, views
  int T, t_cap, k_cap;      // live tables (dense 0..T-1), capacities of the table and dish-slot arrays
  const double* y;          // [d][n]
  int* table_of;            // [n]
  int* n_t;                 // [t_cap]
  int* dish_of;             // [d][t_cap]
  int* K;                   // [d] dish slots in use (never shrinks)
  int* n_vk;                // [d][k_cap]
  int* l_vk;                // [d][k_cap]
  double* sum_y;            // [d][k_cap]
  double* sum_y2;           // [d][k_cap]
  double* alpha_v;          // [d]
  double* sigma_v;          // [d]
  double* tau_v;            // [d]
  double alpha_g, sigma_g;
  uint64_t seed, calls;     // call-ordered stream position
  int err;                  // 1: table capacity, 2: dish capacity, 4: customer without a table
  double* prob;             // scratch [t_cap]
  double* wts;              // scratch [k_cap + 1]
  int* cand;                // scratch [k_cap]
};

// ---- the stream: R::runif / R::rnorm of the reference, in call order -----------------------------------------------
MV_SEQ_FN double s_runif(State& s, double a, double b) {
  const U4 r = stream_block(s.seed, 0u, kDomCallSeq, 0u, 0u, s.calls++);
  return a + (b - a) * uniform_f64_from(r.x, r.y);
}
MV_SEQ_FN double s_rnorm(State& s, double mean, double sd) {
  const U4 r = stream_block(s.seed, 0u, kDomCallSeq, 1u, 0u, s.calls++);
  const double u1 = uniform_f64_from(r.x, r.y), u2 = uniform_f64_from(r.z, r.w);
  return mean + sd * (sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2));
}

// ---- densities (multiview_utils.cpp:307-350), in the reference's operation order -----------------------------------
MV_SEQ_FN double f_dish(const State& s, int v, int k, int i) {
  const double yi = s.y[(size_t)v * s.n + i], tau = s.tau_v[v];
  const size_t o = (size_t)v * s.k_cap + k;
  const int m = s.n_vk[o];
  const double S1 = s.sum_y[o], S2 = s.sum_y2[o];
  const double q_old = -0.5 * S2 / tau;
  const double r_old = 0.5 * (S1 * S1) / (tau * (tau + m));
  const double l_old = -0.5 * m * log(2.0 * kPiS * tau) - 0.5 * log(tau * (tau + m));
  const int m1 = m + 1;
  const double S1n = S1 + yi, S2n = S2 + yi * yi;
  const double q_new = -0.5 * S2n / tau;
  const double r_new = 0.5 * (S1n * S1n) / (tau * (tau + m1));
  const double l_new = -0.5 * m1 * log(2.0 * kPiS * tau) - 0.5 * log(tau * (tau + m1));
  return exp((l_new + q_new + r_new) - (l_old + q_old + r_old));
}
MV_SEQ_FN double f_fresh(const State& s, int v, int i) {
  const double yi = s.y[(size_t)v * s.n + i], tau = s.tau_v[v];
  const double a = -0.5 * log(2.0 * kPiS * tau);
  const double b = -0.5 * (yi * yi) / tau;
  return exp(a + b);
}

// ---- a customer leaves its table (multiview_utils.cpp:138-192) -------------------------------------------------------
MV_SEQ_FN void leave(State& s, int i) {
  const int t = s.table_of[i];
  if (t < 0 || t >= s.T) { s.err |= 4; return; }
  s.n_t[t] -= 1;
  for (int v = 0; v < s.d; ++v) {
    const size_t o = (size_t)v * s.k_cap + s.dish_of[(size_t)v * s.t_cap + t];
    const double yi = s.y[(size_t)v * s.n + i];
    s.n_vk[o] -= 1;
    s.sum_y[o] -= yi;
    s.sum_y2[o] -= yi * yi;
  }
  s.table_of[i] = -1;
  if (s.n_t[t] != 0) return;
  // the table is empty: its dishes lose a table, the LAST table takes its index (:168-191)
  for (int v = 0; v < s.d; ++v) {
    const int k = s.dish_of[(size_t)v * s.t_cap + t];
    const size_t o = (size_t)v * s.k_cap + k;
    if (k >= 0 && s.l_vk[o] > 0) s.l_vk[o] -= 1;
  }
  const int last = s.T - 1;
  if (t != last) {
    s.n_t[t] = s.n_t[last];
    for (int v = 0; v < s.d; ++v) s.dish_of[(size_t)v * s.t_cap + t] = s.dish_of[(size_t)v * s.t_cap + last];
    for (int j = 0; j < s.n; ++j) if (s.table_of[j] == last) s.table_of[j] = t;
  }
  s.T -= 1;
}

MV_SEQ_FN void join(State& s, int i, int t) {                      // :194-207
  s.table_of[i] = t;
  s.n_t[t] += 1;
  for (int v = 0; v < s.d; ++v) {
    const size_t o = (size_t)v * s.k_cap + s.dish_of[(size_t)v * s.t_cap + t];
    const double yi = s.y[(size_t)v * s.n + i];
    s.n_vk[o] += 1;
    s.sum_y[o] += yi;
    s.sum_y2[o] += yi * yi;
  }
}

MV_SEQ_FN int fresh_dish(State& s, int v) {                        // a new dish slot is always appended (:250-258, :268-275)
  if (s.K[v] >= s.k_cap) { s.err |= 2; return s.k_cap - 1; }
  const int k = s.K[v]++;
  const size_t o = (size_t)v * s.k_cap + k;
  s.n_vk[o] = 0; s.l_vk[o] = 0; s.sum_y[o] = 0.0; s.sum_y2[o] = 0.0;
  return k;
}

// dish of a new table in view v (:224-276)
MV_SEQ_FN int pick_dish(State& s, int v, int i) {
  int nc = 0;
  for (int k = 0; k < s.K[v]; ++k) {
    const size_t o = (size_t)v * s.k_cap + k;
    if (s.l_vk[o] > 0) {
      double w = (s.l_vk[o] - s.sigma_v[v]) * f_dish(s, v, k, i);
      if (w < 0) w = 0;
      s.wts[nc] = w;
      s.cand[nc] = k;
      ++nc;
    }
  }
  double wn = (s.alpha_v[v] + s.sigma_v[v] * nc) * f_fresh(s, v, i);
  if (wn < 0) wn = 0;
  s.wts[nc] = wn;
  double total = 0;
  for (int j = 0; j <= nc; ++j) total += s.wts[j];
  if (total <= 0) return fresh_dish(s, v);
  const double u = s_runif(s, 0.0, total);
  double cum = 0;
  for (int j = 0; j < nc; ++j) {
    cum += s.wts[j];
    if (u < cum) return s.cand[j];
  }
  return fresh_dish(s, v);
}

// marginal density of customer i at a NEW table in view v (:40-69)
MV_SEQ_FN double fresh_table_density(const State& s, int v, int i) {
  double tables = 0.0;
  for (int k = 0; k < s.K[v]; ++k) tables += s.l_vk[(size_t)v * s.k_cap + k];
  const double den = s.alpha_v[v] + tables;
  if (den <= 0.0) return f_fresh(s, v, i);
  double acc = 0.0;
  int live = 0;
  for (int k = 0; k < s.K[v]; ++k) {
    const size_t o = (size_t)v * s.k_cap + k;
    if (s.l_vk[o] > 0) {
      ++live;
      double w = (s.l_vk[o] - s.sigma_v[v]);
      if (w < 0.0) w = 0.0;
      acc += w * f_dish(s, v, k, i);
    }
  }
  double wn = (s.alpha_v[v] + live * s.sigma_v[v]);
  if (wn < 0.0) wn = 0.0;
  acc += wn * f_fresh(s, v, i);
  return acc / den;
}

// one customer: leave, weigh every table and a new one, draw, sit down (multiview_gibbs.cpp:157-200)
MV_SEQ_FN void reseat(State& s, int i) {
  leave(s, i);
  if (s.err) return;
  const int T = s.T;
  for (int t = 0; t < T; ++t) {                                     // multiview_utils.cpp:83-115
    if (s.n_t[t] == 0) { s.prob[t] = 0.0; continue; }
    double lp = 0.0;
    for (int v = 0; v < s.d; ++v) lp += log(f_dish(s, v, s.dish_of[(size_t)v * s.t_cap + t], i));
    const double mass = s.n_t[t] - s.sigma_g;
    s.prob[t] = (mass <= 0.0) ? 0.0 : mass * exp(lp);
  }
  double lnew = 0.0;
  for (int v = 0; v < s.d; ++v) lnew += log(fresh_table_density(s, v, i));
  int live = 0;
  for (int t = 0; t < T; ++t) if (s.n_t[t] > 0) ++live;
  const double mass_new = s.alpha_g + s.sigma_g * live;
  double p_new = (mass_new <= 0.0) ? 0.0 : mass_new * exp(lnew);
  double total = p_new;
  for (int t = 0; t < T; ++t) total += s.prob[t];
  if (total <= 0.0) {                                               // nothing has weight: table 0, no draw (:172-176)
    if (T <= 0) { s.err |= 4; return; }
    join(s, i, 0);
    return;
  }
  for (int t = 0; t < T; ++t) s.prob[t] /= total;
  p_new /= total;
  const double u = s_runif(s, 0.0, 1.0);
  double cum = 0.0;
  int pick = -1;
  for (int t = 0; t < T; ++t) {
    cum += s.prob[t];
    if (u < cum) { pick = t; break; }
  }
  if (pick >= 0) { join(s, i, pick); return; }
  // a new table at index T, one customer, a dish per view in view order (:193-196, :278-289)
  if (s.T >= s.t_cap) { s.err |= 1; join(s, i, 0); return; }
  const int t_new = s.T++;
  s.table_of[i] = t_new;
  s.n_t[t_new] = 1;
  for (int v = 0; v < s.d; ++v) s.dish_of[(size_t)v * s.t_cap + t_new] = -1;
  for (int v = 0; v < s.d; ++v) {
    const int k = pick_dish(s, v, i);
    s.dish_of[(size_t)v * s.t_cap + t_new] = k;
    const size_t o = (size_t)v * s.k_cap + k;
    const double yi = s.y[(size_t)v * s.n + i];
    s.l_vk[o] += 1;
    s.n_vk[o] += 1;
    s.sum_y[o] += yi;
    s.sum_y2[o] += yi * yi;
  }
}

// ---- hyperparameters (multiview_hyper.cpp) ---------------------------------------------------------------------------
MV_SEQ_FN double prior_alpha(double a) { return (a <= 0.0) ? -INFINITY : (4.0 - 1.0) * log(a) - 3.0 * a; }          // :344-351
MV_SEQ_FN double prior_sigma(double x) {                                                                               // :353-360
  return (x <= 0.0 || x >= 1.0) ? -INFINITY : (1.0 - 1.0) * log(x) + (5.0 - 1.0) * log(1.0 - x);
}
// log EPPF of a partition given by `sizes` (entries <= 0 skipped when `skip_empty`), `items` items in total (:53-83, :295-342)
MV_SEQ_FN double eppf(const int* sizes, int count, bool view_level, int n_customers, double alpha, double sigma) {
  if (!(sigma > kEpsS && sigma < 1.0 - kEpsS)) return -INFINITY;
  if (alpha <= -sigma) return -INFINITY;
  int clusters = 0, items = 0;
  if (view_level) {                                                 // tables per dish; dead slots do not count
    for (int k = 0; k < count; ++k) if (sizes[k] > 0) { ++clusters; items += sizes[k]; }
    if (items == 0) return 0.0;
  } else {                                                          // customers per table: all T tables, n customers
    if (count <= 0) return 0.0;
    clusters = count;
    items = n_customers;
  }
  double lp = 0.0;
  for (int j = 0; j < clusters; ++j) {
    const double term = alpha + j * sigma;
    if (term <= 0.0) return -INFINITY;
    lp += log(term);
  }
  for (int q = 1; q < items; ++q) {
    const double term = alpha + q;
    if (term <= 0.0) return -INFINITY;
    lp -= log(term);
  }
  for (int k = 0; k < count; ++k) {
    if (view_level && sizes[k] <= 0) continue;
    for (int m = 1; m < sizes[k]; ++m) {
      const double term = (double)m - sigma;
      if (term <= 0.0) return -INFINITY;
      lp += log(term);
    }
  }
  return lp;
}
MV_SEQ_FN double tau_target(const State& s, int v, double tau) {                                                      // :176-209
  if (tau <= 0.0) return -INFINITY;
  double ll = 0.0;
  for (int k = 0; k < s.K[v]; ++k) {
    const size_t o = (size_t)v * s.k_cap + k;
    const int m = s.n_vk[o];
    if (m == 0) continue;
    double sse = s.sum_y2[o] - (s.sum_y[o] * s.sum_y[o]) / (double)m;
    if (sse < 0.0) sse = 0.0;
    ll += -0.5 * m * log(2.0 * kPiS * tau) - 0.5 * (sse / tau);
  }
  const double a_tau = 2.0, b_tau = 1.0;
  return ll + (a_tau * log(b_tau) - lgamma(a_tau) - (a_tau + 1.0) * log(tau) - b_tau / tau);
}
MV_SEQ_FN double walk_alpha(State& s, double a_old) {                                                                 // :100-108
  double la = log(a_old > kEpsS ? a_old : kEpsS);
  la += s_rnorm(s, 0.0, 0.1);
  const double c = exp(la);
  return (c > kEpsS) ? c : kEpsS;
}
MV_SEQ_FN double walk_sigma(State& s, double x_old) {                                                                 // :110-128
  double p = x_old + s_rnorm(s, 0.0, 0.05);
  while (p <= kEpsS || p >= 1.0 - kEpsS) {
    if (p <= kEpsS) p = 2.0 * kEpsS - p;
    if (p >= 1.0 - kEpsS) p = 2.0 * (1.0 - kEpsS) - p;
  }
  return p < kEpsS ? kEpsS : (p > 1.0 - kEpsS ? 1.0 - kEpsS : p);
}
MV_SEQ_FN double sigma_target_view(const State& s, int v, double x) {
  if (x <= kEpsS || x >= 1.0 - kEpsS) return -INFINITY;
  return eppf(s.l_vk + (size_t)v * s.k_cap, s.K[v], true, 0, s.alpha_v[v], x) + prior_sigma(x);
}
MV_SEQ_FN double sigma_target_global(const State& s, double x) {
  if (x <= kEpsS || x >= 1.0 - kEpsS) return -INFINITY;
  return eppf(s.n_t, s.T, false, s.n, s.alpha_g, x) + prior_sigma(x);
}

MV_SEQ_FN void hyper_step(State& s) {                                                                                  // :211-292
  for (int v = 0; v < s.d; ++v) {                                   // tau_v
    double t_old = s.tau_v[v];
    if (t_old <= 0.0) t_old = kEpsS;
    const double lo = tau_target(s, v, t_old);
    const double t_new = exp(log(t_old) + s_rnorm(s, 0.0, 0.3));
    if (t_new <= 0.0) continue;
    const double ln = tau_target(s, v, t_new);
    const double acc = (ln - lo) + (log(t_new) - log(t_old));
    if (log(s_runif(s, 0.0, 1.0)) < acc) s.tau_v[v] = t_new;
  }
  for (int v = 0; v < s.d; ++v) {                                   // alpha_v, sigma_v
    const int* l = s.l_vk + (size_t)v * s.k_cap;
    double a_old = s.alpha_v[v];
    if (a_old <= 0.0) a_old = kEpsS;
    const double a_new = walk_alpha(s, a_old);
    const double lo = (a_old <= 0.0) ? -INFINITY : eppf(l, s.K[v], true, 0, a_old, s.sigma_v[v]) + prior_alpha(a_old);
    const double ln = (a_new <= 0.0) ? -INFINITY : eppf(l, s.K[v], true, 0, a_new, s.sigma_v[v]) + prior_alpha(a_new);
    const double acc = (ln - lo) + (log(a_new) - log(a_old));
    if (log(s_runif(s, 0.0, 1.0)) < acc) s.alpha_v[v] = a_new;
    const double x_old = s.sigma_v[v];
    const double x_new = walk_sigma(s, x_old);
    const double lu = log(s_runif(s, 0.0, 1.0));
    if (lu < sigma_target_view(s, v, x_new) - sigma_target_view(s, v, x_old)) s.sigma_v[v] = x_new;
  }
  {                                                                 // alpha_global, sigma_global
    double a_old = s.alpha_g;
    if (a_old <= 0.0) a_old = kEpsS;
    const double a_new = walk_alpha(s, a_old);
    const double lo = (a_old <= 0.0) ? -INFINITY : eppf(s.n_t, s.T, false, s.n, a_old, s.sigma_g) + prior_alpha(a_old);
    const double ln = (a_new <= 0.0) ? -INFINITY : eppf(s.n_t, s.T, false, s.n, a_new, s.sigma_g) + prior_alpha(a_new);
    const double acc = (ln - lo) + (log(a_new) - log(a_old));
    if (log(s_runif(s, 0.0, 1.0)) < acc) s.alpha_g = a_new;
    const double x_old = s.sigma_g;
    const double x_new = walk_sigma(s, x_old);
    const double lu = log(s_runif(s, 0.0, 1.0));
    if (lu < sigma_target_global(s, x_new) - sigma_target_global(s, x_old)) s.sigma_g = x_new;
  }
}

// ---- the reference's start (multiview_gibbs.cpp:12-103) ---------------------------------------------------------------
MV_SEQ_FN void start(State& s) {
  const int T0 = 4, K0 = 2;
  s.T = T0;
  for (int t = 0; t < T0; ++t) s.n_t[t] = 0;
  for (int i = 0; i < s.n; ++i) {
    int t = (int)floor(s_runif(s, 0.0, (double)T0));
    if (t < 0) t = 0;
    if (t >= T0) t = T0 - 1;
    s.table_of[i] = t;
    s.n_t[t] += 1;
  }
  for (int v = 0; v < s.d; ++v) {
    s.K[v] = K0;
    for (int k = 0; k < K0; ++k) {
      const size_t o = (size_t)v * s.k_cap + k;
      s.n_vk[o] = 0; s.l_vk[o] = 0; s.sum_y[o] = 0.0; s.sum_y2[o] = 0.0;
    }
    for (int t = 0; t < T0; ++t) {
      int k = (int)floor(s_runif(s, 0.0, (double)K0));
      if (k < 0) k = 0;
      if (k >= K0) k = K0 - 1;
      s.dish_of[(size_t)v * s.t_cap + t] = k;
      s.l_vk[(size_t)v * s.k_cap + k] += 1;
    }
    const double* yv = s.y + (size_t)v * s.n;
    for (int i = 0; i < s.n; ++i) {
      const size_t o = (size_t)v * s.k_cap + s.dish_of[(size_t)v * s.t_cap + s.table_of[i]];
      s.n_vk[o] += 1;
      s.sum_y[o] += yv[i];
      s.sum_y2[o] += yv[i] * yv[i];
    }
    s.alpha_v[v] = 1.0;
    s.sigma_v[v] = 0.5;
    double s1 = 0.0;
    for (int i = 0; i < s.n; ++i) s1 += yv[i];
    const double mean = s1 / (s.n > 1 ? s.n : 1);
    double var = 0.0;
    if (s.n > 1) {
      for (int i = 0; i < s.n; ++i) { const double df = yv[i] - mean; var += df * df; }
      var /= (s.n - 1);
    } else {
      var = 1.0;
    }
    if (var <= 0.0) var = 1.0;
    s.tau_v[v] = var * 0.25 * 0.01;
  }
  s.alpha_g = 1;
  s.sigma_g = 0.6;
}

MV_SEQ_FN void sweep(State& s) {                                    // multiview_gibbs.cpp:157-202
  for (int i = 0; i < s.n && !s.err; ++i) reseat(s, i);
  if (!s.err) hyper_step(s);
}

}  // namespace seq
}  // namespace mv
