/* oracle/mv_oracle.c — TEST INFRASTRUCTURE (CPU oracle), not product code.
 * See mv_oracle.h for the parity status of each part.  Plain C11, FP64 unless a function name
 * ends in _f32 (those restate the device's FP32 epilogue operation by operation and must be
 * compiled with -ffp-contract=off, which oracle/Makefile does).
 */
#include "mv_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mv_philox_ref.h"

#ifdef _OPENMP
#include <omp.h>
#endif

#define K_EPS 1e-6 /* multiview_hyper.cpp:13 */

/* ======================================================================================
 * Philox
 * ==================================================================================== */
void mvo_philox_raw(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  mvo_philox4x32_10(ctr, key, out);
}
void mvo_philox_block(uint64_t seed, uint32_t chain, uint32_t domain, uint32_t slot, uint32_t sweep,
                      uint64_t index, uint32_t out[4]) {
  mvo_stream_block(seed, chain, domain, slot, sweep, index, out);
}
float mvo_uf(uint64_t seed, uint32_t chain, uint32_t domain, uint32_t slot, uint32_t sweep, uint64_t index) {
  return mvo_uniform_f32(seed, chain, domain, slot, sweep, index);
}
double mvo_u53(uint64_t seed, uint32_t chain, uint32_t domain, uint32_t slot, uint32_t sweep, uint64_t index) {
  return mvo_uniform53(seed, chain, domain, slot, sweep, index);
}
double mvo_z(uint64_t seed, uint32_t chain, uint32_t domain, uint32_t slot, uint32_t sweep, uint64_t index) {
  return mvo_normal(seed, chain, domain, slot, sweep, index);
}

/* ======================================================================================
 * State
 * ==================================================================================== */
static int is_counts(const mvo_state* s, int v) { return s->csr != NULL && s->csr[v].rowptr != NULL; }

int mvo_rebuild_stats(mvo_state* s) {
  const int cap = s->cap, V = s->V;
  memset(s->n_t, 0, sizeof(int32_t) * (size_t)cap);
  for (int i = 0; i < s->n; ++i) {
    int t = s->table_of[i];
    if (t < 0 || t >= cap) return 1;
    s->n_t[t]++;
  }
  for (int v = 0; v < V; ++v) {
    const int D = s->D[v];
    int32_t* n_vk = s->n_vk + (size_t)v * cap;
    int32_t* l_vk = s->l_vk + (size_t)v * cap;
    int32_t* dish = s->dish_of + (size_t)v * cap;
    double* S1 = s->S1[v];
    double* S2 = s->S2 + (size_t)v * cap;
    memset(n_vk, 0, sizeof(int32_t) * (size_t)cap);
    memset(l_vk, 0, sizeof(int32_t) * (size_t)cap);
    memset(S1, 0, sizeof(double) * (size_t)cap * D);
    memset(S2, 0, sizeof(double) * (size_t)cap);
    for (int t = 0; t < cap; ++t) {
      if (s->n_t[t] == 0) { dish[t] = -1; continue; }
      if (dish[t] < 0 || dish[t] >= cap) return 2;
      l_vk[dish[t]]++;                                   /* multiview_gibbs.cpp:60 */
    }
    if (is_counts(s, v)) {                                 /* count view: word counts per dish */
      const mvo_csr* cs = &s->csr[v];
      memset(cs->cd, 0, sizeof(int64_t) * (size_t)cap * cs->vocab);
      memset(cs->ctot, 0, sizeof(int64_t) * (size_t)cap);
      for (int i = 0; i < s->n; ++i) {
        int k = dish[s->table_of[i]];
        double tot = 0.0;
        for (int j = cs->rowptr[i]; j < cs->rowptr[i + 1]; ++j) {
          cs->cd[(size_t)k * cs->vocab + cs->col[j]] += (int64_t)cs->val[j];
          tot += (double)cs->val[j];
        }
        cs->ctot[k] += (int64_t)tot;
        n_vk[k]++;
        S2[k] += tot;                                      /* the device keeps the token totals in this slot */
      }
      continue;
    }
    for (int i = 0; i < s->n; ++i) {                      /* multiview_gibbs.cpp:64-73 */
      int k = dish[s->table_of[i]];
      const float* xi = s->x[v] + (size_t)i * D;
      double q = 0.0;
      for (int dd = 0; dd < D; ++dd) {
        double val = (double)xi[dd];
        S1[(size_t)k * D + dd] += val;
        q += val * val;
      }
      n_vk[k]++;
      S2[k] += q;
    }
  }
  return 0;
}

int mvo_init_reference(mvo_state* s) {
  const int T0 = 4, K0 = 2;                               /* multiview_gibbs.cpp:14-15 */
  if (s->cap < T0) return 1;
  for (int i = 0; i < s->n; ++i) {                         /* :25-33 */
    double u = mvo_uniform53(s->seed, s->chain, MVO_DOM_INIT_TABLE, 0, 0, (uint64_t)(s->row_offset + i));
    int t = (int)floor(u * (double)T0);
    if (t < 0) t = 0;
    if (t >= T0) t = T0 - 1;
    s->table_of[i] = t;
  }
  for (int v = 0; v < s->V; ++v) {                         /* :55-62 */
    int32_t* dish = s->dish_of + (size_t)v * s->cap;
    for (int t = 0; t < s->cap; ++t) dish[t] = -1;
    for (int t = 0; t < T0; ++t) {
      double u = mvo_uniform53(s->seed, s->chain, MVO_DOM_INIT_DISH, (uint32_t)v, 0, (uint64_t)t);
      int k = (int)floor(u * (double)K0);
      if (k < 0) k = 0;
      if (k >= K0) k = K0 - 1;
      dish[t] = k;
    }
    s->alpha_v[v] = 1.0;                                   /* :75-76 */
    s->sigma_v[v] = 0.5;
    /* :78-94, pooled over the D coordinates (equal to the reference at D = 1) */
    const int D = s->D[v];
    if (is_counts(s, v)) { s->tau_v[v] = 1.0; continue; }   /* no kernel variance in a count view */
    double var_sum = 0.0;
    for (int dd = 0; dd < D; ++dd) {
      double s1 = 0.0;
      for (int i = 0; i < s->n; ++i) s1 += (double)s->x[v][(size_t)i * D + dd];
      double mean = s1 / (double)(s->n > 1 ? s->n : 1);
      double var = 0.0;
      if (s->n > 1) {
        for (int i = 0; i < s->n; ++i) {
          double diff = (double)s->x[v][(size_t)i * D + dd] - mean;
          var += diff * diff;
        }
        var /= (double)(s->n - 1);
      } else {
        var = 1.0;
      }
      var_sum += var;
    }
    double var = var_sum / (double)D;
    if (var <= 0.0) var = 1.0;
    s->tau_v[v] = var * 0.25 * 0.01;
  }
  s->alpha_g = 1.0;                                        /* :97-98 */
  s->sigma_g = 0.6;
  /* a table slot that received no row is free: its dish entries are dropped by the rebuild */
  return mvo_rebuild_stats(s);
}

/* ======================================================================================
 * FP64 per-row arithmetic
 * ==================================================================================== */
double mvo_log_f_vk(const mvo_state* s, int v, int k, const float* x, int loo) {
  const int D = s->D[v];
  const double tau = s->tau_v[v];
  const double* S1 = s->S1[v] + (size_t)k * D;
  double n = (double)s->n_vk[(size_t)v * s->cap + k] - (loo ? 1.0 : 0.0);
  /* N(x; S1/(tau+n), tau(tau+n+1)/(tau+n)) per coordinate — closed form of multiview_utils.cpp:307-338 */
  double var = tau * (tau + n + 1.0) / (tau + n);
  double dist = 0.0;
  for (int dd = 0; dd < D; ++dd) {
    double xv = (double)x[dd];
    double s1 = S1[dd] - (loo ? xv : 0.0);
    double diff = xv - s1 / (tau + n);
    dist += diff * diff;
  }
  return -0.5 * (double)D * log(2.0 * M_PI * var) - 0.5 * dist / var;
}

double mvo_log_f_new(const mvo_state* s, int v, const float* x) {
  const int D = s->D[v];
  const double tau = s->tau_v[v];
  double q = 0.0;
  for (int dd = 0; dd < D; ++dd) q += (double)x[dd] * (double)x[dd];
  return -0.5 * (double)D * log(2.0 * M_PI * tau) - 0.5 * q / tau;  /* :346-349 */
}

/* Count views: plug-in multinomial, leave-one-out for the row's own dish. */
static double counts_log_f_vk(const mvo_state* s, int v, int k, int i, int loo) {
  const mvo_csr* cs = &s->csr[v];
  double tot = 0.0;
  for (int j = cs->rowptr[i]; j < cs->rowptr[i + 1]; ++j) tot += (double)cs->val[j];
  const double den = (double)cs->vocab * s->count_beta + (double)cs->ctot[k] - (loo ? tot : 0.0);
  double lf = 0.0;
  for (int j = cs->rowptr[i]; j < cs->rowptr[i + 1]; ++j) {
    const double c = (double)cs->cd[(size_t)k * cs->vocab + cs->col[j]] - (loo ? (double)cs->val[j] : 0.0);
    lf += (double)cs->val[j] * log((s->count_beta + c) / den);
  }
  return lf;
}
static double counts_log_f_new(const mvo_state* s, int v, int i) {
  const mvo_csr* cs = &s->csr[v];
  double tot = 0.0;
  for (int j = cs->rowptr[i]; j < cs->rowptr[i + 1]; ++j) tot += (double)cs->val[j];
  return -tot * log((double)cs->vocab);
}
/* log f of ROW i under dish k / a new dish, whatever the kind of view v */
static double row_log_f_vk(const mvo_state* s, int v, int k, int i, int loo) {
  if (is_counts(s, v)) return counts_log_f_vk(s, v, k, i, loo);
  return mvo_log_f_vk(s, v, k, s->x[v] + (size_t)i * s->D[v], loo);
}
static double row_log_f_new(const mvo_state* s, int v, int i) {
  if (is_counts(s, v)) return counts_log_f_new(s, v, i);
  return mvo_log_f_new(s, v, s->x[v] + (size_t)i * s->D[v]);
}

static double lse2(double a, double b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  double m = a > b ? a : b;
  return m + log(exp(a - m) + exp(b - m));
}

/* log marginal likelihood of a new table in view v for row x whose current table is t0
 * (multiview_utils.cpp:40-69), with the row's own contribution removed. lf[k] is filled for live k. */
static double log_marginal_new_table(const mvo_state* s, int v, int i, int t0, double* lf,
                                     double lf_new) {
  const int cap = s->cap;
  const int32_t* l_vk = s->l_vk + (size_t)v * cap;
  const int k0 = s->dish_of[(size_t)v * cap + t0];
  const int single = (s->n_t[t0] == 1);
  const double alpha = s->alpha_v[v], sigma = s->sigma_v[v];
  double total_tables = 0.0;
  int K_active = 0;
  double acc = -INFINITY;
  for (int k = 0; k < cap; ++k) {
    int l = l_vk[k] - ((single && k == k0) ? 1 : 0);
    lf[k] = -INFINITY;
    if (l <= 0) continue;
    total_tables += (double)l;
    K_active++;
    lf[k] = row_log_f_vk(s, v, k, i, k == k0);
    double w = (double)l - sigma;                         /* :56-57 */
    if (w > 0.0) acc = lse2(acc, log(w) + lf[k]);
  }
  double denominator = alpha + total_tables;               /* :46-47 */
  if (denominator <= 0.0) return lf_new;
  double w_new = alpha + (double)K_active * sigma;         /* :63-64 */
  if (w_new > 0.0) acc = lse2(acc, log(w_new) + lf_new);
  return acc - log(denominator);
}

int mvo_row_logweights(const mvo_state* s, int i, double* lw, double* Lvt) {
  const int cap = s->cap, V = s->V;
  const int t0 = s->table_of[i];
  const int single = (s->n_t[t0] == 1);
  double* lf = (double*)malloc(sizeof(double) * (size_t)cap);
  int T_nonempty = 0, F = 0;
  for (int t = 0; t < cap; ++t) {
    if (s->n_t[t] > 0) T_nonempty++; else F++;
    lw[t] = 0.0;
  }
  double log_new = 0.0;
  for (int v = 0; v < V; ++v) {
    double lf_new = row_log_f_new(s, v, i);
    double lm = log_marginal_new_table(s, v, i, t0, lf, lf_new);
    log_new += lm;                                         /* multiview_utils.cpp:119-122 */
    for (int t = 0; t < cap; ++t) {
      int k = s->dish_of[(size_t)v * cap + t];
      double val = (k >= 0) ? lf[k] : -INFINITY;
      lw[t] += val;                                        /* :91-108 */
      if (Lvt) Lvt[(size_t)v * (cap + 1) + t] = val;
    }
    if (Lvt) Lvt[(size_t)v * (cap + 1) + cap] = lf_new;
  }
  for (int t = 0; t < cap; ++t) {
    int nt = s->n_t[t] - (t == t0 ? 1 : 0);
    double mass = (double)nt - s->sigma_g;                 /* :110-115 */
    if (nt <= 0 || mass <= 0.0) lw[t] = -INFINITY;
    else lw[t] += log(mass);
  }
  double mass_new = s->alpha_g + s->sigma_g * (double)(T_nonempty - single);  /* :124-135 */
  if (mass_new <= 0.0 || F == 0) lw[cap] = -INFINITY;      /* F == 0: no free slot (capacity rule) */
  else lw[cap] = log(mass_new) + log_new;
  free(lf);
  return 0;
}

int mvo_draw_from_logweights(const mvo_state* s, int i, const double* lw, double u) {
  const int cap = s->cap;
  double M = -INFINITY;
  for (int t = 0; t <= cap; ++t) if (lw[t] > M) M = lw[t];
  if (M == -INFINITY) return s->table_of[i];               /* degenerate: stay (cf. :172-176) */
  double sum_p = exp(lw[cap] - M);                          /* multiview_gibbs.cpp:169-170 */
  for (int t = 0; t < cap; ++t) sum_p += exp(lw[t] - M);
  double cum = 0.0;
  for (int t = 0; t < cap; ++t) {                           /* :181-191 */
    cum += exp(lw[t] - M) / sum_p;
    if (u < cum) return t;
  }
  if (lw[cap] == -INFINITY) {                               /* rounding fall-through with no new-table mass */
    for (int t = cap - 1; t >= 0; --t) if (lw[t] > -INFINITY) return t;
    return s->table_of[i];
  }
  return MVO_NEW;
}

int mvo_draw_rows(const mvo_state* s, int32_t* choice, int threads) {
  const int cap = s->cap;
  if (threads <= 0) threads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
  {
    double* lw = (double*)malloc(sizeof(double) * (size_t)(cap + 1));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int i = 0; i < s->n; ++i) {
      mvo_row_logweights(s, i, lw, NULL);
      double u = (double)mvo_uniform_f32(s->seed, s->chain, MVO_DOM_TABLE, 0, s->sweep,
                                       (uint64_t)(s->row_offset + i));
      choice[i] = mvo_draw_from_logweights(s, i, lw, u);
    }
    free(lw);
  }
  return 0;
}

/* sample_dish_for_new_table (multiview_utils.cpp:224-276) for birth row b; l_live carries the
 * increments of earlier births of this sweep. Returns the dish slot; *is_new = 1 for a new dish. */
static int sample_dish_birth(const mvo_state* s, int v, int b, const int32_t* l_live, double* w_out) {
  const int cap = s->cap;
  const int t0 = s->table_of[b];
  const int k0 = s->dish_of[(size_t)v * cap + t0];
  const int single = (s->n_t[t0] == 1);
  const double alpha = s->alpha_v[v], sigma = s->sigma_v[v];
  double* lwt = (double*)malloc(sizeof(double) * (size_t)(cap + 1));
  int K_active = 0;
  double M = -INFINITY;
  for (int k = 0; k < cap; ++k) {
    int l = l_live[k] - ((single && k == k0) ? 1 : 0);
    lwt[k] = -INFINITY;
    if (l <= 0) continue;
    K_active++;
    double w = (double)l - sigma;                          /* :232-233 */
    if (w <= 0.0) continue;
    /* a dish opened earlier in this sweep has no rows in the sweep-start statistics: n = 0 */
    lwt[k] = log(w) + row_log_f_vk(s, v, k, b, (k == k0) && s->n_vk[(size_t)v * cap + k] > 0);
    if (lwt[k] > M) M = lwt[k];
  }
  double w_new = alpha + sigma * (double)K_active;          /* :241-243 */
  lwt[cap] = (w_new > 0.0) ? log(w_new) + row_log_f_new(s, v, b) : -INFINITY;
  if (lwt[cap] > M) M = lwt[cap];
  int pick = -1;
  if (M > -INFINITY) {
    double total = 0.0;
    for (int k = 0; k <= cap; ++k) {
      double w = (lwt[k] == -INFINITY) ? 0.0 : exp(lwt[k] - M);
      if (w_out) w_out[k] = w;
      total += w;                                          /* :247-248, candidates ascending then new */
    }
    double u = mvo_uniform53(s->seed, s->chain, MVO_DOM_DISH, (uint32_t)v, s->sweep,
                             (uint64_t)(s->row_offset + b)) * total;   /* :261 */
    double cum = 0.0;
    for (int k = 0; k < cap; ++k) {
      if (lwt[k] == -INFINITY) continue;
      cum += exp(lwt[k] - M);
      if (u < cum) { pick = k; break; }
    }
  } else if (w_out) {
    for (int k = 0; k <= cap; ++k) w_out[k] = 0.0;
  }
  free(lwt);
  if (pick >= 0) return pick;
  for (int k = 0; k < cap; ++k) if (l_live[k] == 0) return k;   /* new dish: lowest free slot (:268-275) */
  return -1;
}

int mvo_reseat(mvo_state* s, const int32_t* choice, int32_t* n_seated, int32_t* birth_rows, double* birth_w) {
  const int cap = s->cap, V = s->V;
  int32_t* free_slots = (int32_t*)malloc(sizeof(int32_t) * (size_t)cap);
  int32_t* l_live = (int32_t*)malloc(sizeof(int32_t) * (size_t)V * cap);
  int32_t* new_table = (int32_t*)malloc(sizeof(int32_t) * (size_t)s->n);
  int32_t* new_dish = (int32_t*)malloc(sizeof(int32_t) * (size_t)V * cap);
  int F = 0, seated = 0;
  for (int t = 0; t < cap; ++t) if (s->n_t[t] == 0) free_slots[F++] = t;
  memcpy(l_live, s->l_vk, sizeof(int32_t) * (size_t)V * cap);
  memcpy(new_dish, s->dish_of, sizeof(int32_t) * (size_t)V * cap);
  for (int i = 0; i < s->n; ++i) {
    if (choice[i] != MVO_NEW) { new_table[i] = choice[i]; continue; }
    if (seated >= F) { new_table[i] = s->table_of[i]; continue; }   /* overflow: stay */
    int tn = free_slots[seated];
    new_table[i] = tn;
    for (int v = 0; v < V; ++v) {                                    /* multiview_utils.cpp:278-289 */
      double* w_out = birth_w ? birth_w + ((size_t)seated * V + v) * (cap + 1) : NULL;
      int k = sample_dish_birth(s, v, i, l_live + (size_t)v * cap, w_out);
      if (k < 0) { free(free_slots); free(l_live); free(new_table); free(new_dish); return 3; }
      new_dish[(size_t)v * cap + tn] = k;
      l_live[(size_t)v * cap + k]++;
    }
    if (birth_rows) birth_rows[seated] = i;
    seated++;
  }
  if (n_seated) *n_seated = seated;
  memcpy(s->table_of, new_table, sizeof(int32_t) * (size_t)s->n);
  memcpy(s->dish_of, new_dish, sizeof(int32_t) * (size_t)V * cap);
  free(free_slots); free(l_live); free(new_table); free(new_dish);
  return mvo_rebuild_stats(s);   /* deaths: slots left without rows lose their dishes here */
}

/* ======================================================================================
 * Hyperparameter step (multiview_hyper.cpp)
 * ==================================================================================== */
static double log_prior_alpha(double alpha) {              /* :344-351 */
  if (alpha <= 0.0) return -INFINITY;
  return (4.0 - 1.0) * log(alpha) - 3.0 * alpha;
}
static double log_prior_sigma(double sigma) {              /* :353-360 */
  if (sigma <= 0.0 || sigma >= 1.0) return -INFINITY;
  return (1.0 - 1.0) * log(sigma) + (5.0 - 1.0) * log(1.0 - sigma);
}

static double eppf_core(const int32_t* counts, int n_counts, long total_items, int n_clusters_for_first_sum,
                        double alpha, double sigma, int use_lgamma) {
  double logp = 0.0;
  for (int j = 0; j < n_clusters_for_first_sum; ++j) {      /* :60-64 / :319-323 */
    double term = alpha + (double)j * sigma;
    if (term <= 0.0) return -INFINITY;
    logp += log(term);
  }
  if (use_lgamma) {
    if (total_items > 1) {
      if (alpha + 1.0 <= 0.0) return -INFINITY;
      logp -= lgamma(alpha + (double)total_items) - lgamma(alpha + 1.0);
    }
    for (int c = 0; c < n_counts; ++c)
      if (counts[c] > 1) logp += lgamma((double)counts[c] - sigma) - lgamma(1.0 - sigma);
  } else {
    for (long i = 1; i < total_items; ++i) {                /* :68-72 / :326-330 */
      double term = alpha + (double)i;
      if (term <= 0.0) return -INFINITY;
      logp -= log(term);
    }
    for (int c = 0; c < n_counts; ++c)                      /* :74-80 / :333-339 */
      for (int m = 1; m < counts[c]; ++m) {
        double term = (double)m - sigma;
        if (term <= 0.0) return -INFINITY;
        logp += log(term);
      }
  }
  return logp;
}

double mvo_log_EPPF_view(const mvo_state* s, int v, double alpha, double sigma, int use_lgamma) {
  if (v < 0 || v >= s->V) return -INFINITY;                 /* :296-299 */
  if (!(sigma > K_EPS && sigma < 1.0 - K_EPS)) return -INFINITY;
  if (alpha <= -sigma) return -INFINITY;
  const int cap = s->cap;
  int32_t* sizes = (int32_t*)malloc(sizeof(int32_t) * (size_t)cap);
  int K_active = 0;
  long total_tables = 0;
  for (int k = 0; k < cap; ++k) {                           /* :306-312 */
    int c = s->l_vk[(size_t)v * cap + k];
    if (c > 0) { sizes[K_active++] = c; total_tables += c; }
  }
  double r = 0.0;
  if (total_tables > 0) r = eppf_core(sizes, K_active, total_tables, K_active, alpha, sigma, use_lgamma);
  free(sizes);
  return r;
}

double mvo_log_EPPF_global(const mvo_state* s, double alpha, double sigma, int use_lgamma) {
  if (!(sigma > K_EPS && sigma < 1.0 - K_EPS)) return -INFINITY;   /* :54-56 */
  if (alpha <= -sigma) return -INFINITY;
  const int cap = s->cap;
  int32_t* sizes = (int32_t*)malloc(sizeof(int32_t) * (size_t)cap);
  int T_live = 0;
  for (int t = 0; t < cap; ++t) if (s->n_t[t] > 0) sizes[T_live++] = s->n_t[t];
  double r = 0.0;
  /* the reference's T counts only live tables (empty ones are swap-deleted, multiview_utils.cpp:168-191) */
  if (T_live > 0) r = eppf_core(sizes, T_live, (long)s->n_global, T_live, alpha, sigma, use_lgamma);
  free(sizes);
  return r;
}

double mvo_log_posterior_tau(const mvo_state* s, int v, double tau) {   /* :176-209 */
  if (tau <= 0.0) return -INFINITY;
  const int cap = s->cap, D = s->D[v];
  double loglik = 0.0;
  for (int k = 0; k < cap; ++k) {
    int n_k = s->n_vk[(size_t)v * cap + k];
    if (n_k == 0) continue;
    const double* S1 = s->S1[v] + (size_t)k * D;
    double s1sq = 0.0;
    for (int dd = 0; dd < D; ++dd) s1sq += S1[dd] * S1[dd];
    double sse = s->S2[(size_t)v * cap + k] - s1sq / (double)n_k;     /* :191 */
    if (sse < 0.0) sse = 0.0;
    loglik += -0.5 * (double)n_k * (double)D * log(2.0 * M_PI * tau) - 0.5 * (sse / tau);
  }
  const double a_tau = 2.0, b_tau = 1.0;                    /* :133-134 */
  double logprior = a_tau * log(b_tau) - lgamma(a_tau) - (a_tau + 1.0) * log(tau) - b_tau / tau;
  return loglik + logprior;
}

static double reflect_unit(double value) {                  /* :110-122 */
  double prop = value;
  while (prop <= K_EPS || prop >= 1.0 - K_EPS) {
    if (prop <= K_EPS) prop = 2.0 * K_EPS - prop;
    if (prop >= 1.0 - K_EPS) prop = 2.0 * (1.0 - K_EPS) - prop;
  }
  if (prop < K_EPS) prop = K_EPS;
  if (prop > 1.0 - K_EPS) prop = 1.0 - K_EPS;
  return prop;
}

int mvo_hyper_step(mvo_state* s, const double* z_in, const double* u_in, int use_lgamma) {
  const int V = s->V;
  int hz = 0, hu = 0;
#define NEXT_Z() (z_in ? z_in[hz++] : mvo_normal(s->seed, s->chain, MVO_DOM_HYPER_NORMAL, 0, s->sweep, (uint64_t)(hz++)))
#define NEXT_U() (u_in ? u_in[hu++] : mvo_uniform53(s->seed, s->chain, MVO_DOM_HYPER_UNIF, 0, s->sweep, (uint64_t)(hu++)))
  for (int v = 0; v < V; ++v) {                             /* update_tau_v_MH, :211-231 */
    if (is_counts(s, v)) { (void)NEXT_Z(); (void)NEXT_U(); continue; }   /* no tau; stream positions stay fixed */
    double tau_old = s->tau_v[v];
    if (tau_old <= 0.0) tau_old = K_EPS;
    double log_old = mvo_log_posterior_tau(s, v, tau_old);
    double tau_prop = exp(log(tau_old) + 0.0 + 0.3 * NEXT_Z());        /* propose_tau :166-174 */
    /* :221 `continue` on tau_prop <= 0 is unreachable (exp > 0); indices stay fixed per step */
    double log_new = mvo_log_posterior_tau(s, v, tau_prop);
    double log_acc = (log_new - log_old) + (log(tau_prop) - log(tau_old));
    if (log(NEXT_U()) < log_acc) s->tau_v[v] = tau_prop;
  }
  for (int v = 0; v < V; ++v) {                             /* :239-266 */
    double alpha_old = s->alpha_v[v];
    if (alpha_old <= 0.0) alpha_old = K_EPS;
    double cand = exp(log(alpha_old > K_EPS ? alpha_old : K_EPS) + 0.0 + 0.1 * NEXT_Z());   /* :100-108 */
    double alpha_prop = cand > K_EPS ? cand : K_EPS;
    double sigma = s->sigma_v[v];
    double lo = mvo_log_EPPF_view(s, v, alpha_old, sigma, use_lgamma) + log_prior_alpha(alpha_old);
    double ln = mvo_log_EPPF_view(s, v, alpha_prop, sigma, use_lgamma) + log_prior_alpha(alpha_prop);
    double log_acc = (ln - lo) + (log(alpha_prop) - log(alpha_old));
    if (log(NEXT_U()) < log_acc) s->alpha_v[v] = alpha_prop;

    double sigma_old = s->sigma_v[v];
    double sigma_prop = reflect_unit(sigma_old + 0.0 + 0.05 * NEXT_Z());                    /* :124-128 */
    double alpha = s->alpha_v[v];
    double lpn = (sigma_prop <= K_EPS || sigma_prop >= 1.0 - K_EPS) ? -INFINITY
                 : mvo_log_EPPF_view(s, v, alpha, sigma_prop, use_lgamma) + log_prior_sigma(sigma_prop);
    double lpo = (sigma_old <= K_EPS || sigma_old >= 1.0 - K_EPS) ? -INFINITY
                 : mvo_log_EPPF_view(s, v, alpha, sigma_old, use_lgamma) + log_prior_sigma(sigma_old);
    if (log(NEXT_U()) < lpn - lpo) s->sigma_v[v] = sigma_prop;
  }
  {                                                         /* :268-291 */
    double ag_old = s->alpha_g;
    if (ag_old <= 0.0) ag_old = K_EPS;
    double cand = exp(log(ag_old > K_EPS ? ag_old : K_EPS) + 0.0 + 0.1 * NEXT_Z());
    double ag_prop = cand > K_EPS ? cand : K_EPS;
    double lo = mvo_log_EPPF_global(s, ag_old, s->sigma_g, use_lgamma) + log_prior_alpha(ag_old);
    double ln = mvo_log_EPPF_global(s, ag_prop, s->sigma_g, use_lgamma) + log_prior_alpha(ag_prop);
    double log_acc = (ln - lo) + (log(ag_prop) - log(ag_old));
    if (log(NEXT_U()) < log_acc) s->alpha_g = ag_prop;

    double sg_old = s->sigma_g;
    double sg_prop = reflect_unit(sg_old + 0.0 + 0.05 * NEXT_Z());
    double lpn = (sg_prop <= K_EPS || sg_prop >= 1.0 - K_EPS) ? -INFINITY
                 : mvo_log_EPPF_global(s, s->alpha_g, sg_prop, use_lgamma) + log_prior_sigma(sg_prop);
    double lpo = (sg_old <= K_EPS || sg_old >= 1.0 - K_EPS) ? -INFINITY
                 : mvo_log_EPPF_global(s, s->alpha_g, sg_old, use_lgamma) + log_prior_sigma(sg_old);
    if (log(NEXT_U()) < lpn - lpo) s->sigma_g = sg_prop;
  }
#undef NEXT_Z
#undef NEXT_U
  return 0;
}

int mvo_sweep(mvo_state* s, int n_sweeps, int threads, int do_hyper) {
  int32_t* choice = (int32_t*)malloc(sizeof(int32_t) * (size_t)s->n);
  int rc = 0;
  for (int it = 0; it < n_sweeps && rc == 0; ++it) {
    rc = mvo_draw_rows(s, choice, threads);
    if (rc == 0) rc = mvo_reseat(s, choice, NULL, NULL, NULL);
    if (rc == 0 && do_hyper) rc = mvo_hyper_step(s, NULL, NULL, 1);
    s->sweep++;
  }
  free(choice);
  return rc;
}

/* ======================================================================================
 * FP32 mirror of the device epilogue (log2 domain).  Every operation below is a single
 * correctly-rounded FP32 operation, in the order the CUDA kernel performs it.
 * ==================================================================================== */
static inline float f32_from_bits(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static inline uint32_t bits_from_f32(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }

float mvo_exp2m(float d) {
  /* 2^d for d <= 0: round-to-nearest split d = n + f, |f| <= 1/2, degree-5 polynomial, exponent add */
  d = fmaxf(d, -125.0f);
  float r = d + 12582912.0f;            /* 1.5 * 2^23: the low mantissa bits of r now hold n */
  float n = r - 12582912.0f;
  float f = d - n;
  float p = 0x1.5bba14p-10f;
  p = fmaf(p, f, 0x1.3cea88p-7f);
  p = fmaf(p, f, 0x1.c6b752p-5f);
  p = fmaf(p, f, 0x1.ebf9bcp-3f);
  p = fmaf(p, f, 0x1.62e42ap-1f);
  p = fmaf(p, f, 1.0f);
  return f32_from_bits(bits_from_f32(p) + (bits_from_f32(r) << 23));
}

float mvo_log2m(float s) {
  /* log2(s) for normal s > 0: s = 2^e * m, m in [sqrt(1/2), sqrt(2)); atanh series in t = (m-1)/(m+1) */
  uint32_t b = bits_from_f32(s);
  int32_t e = (int32_t)(b >> 23) - 127;
  float m = f32_from_bits((b & 0x007FFFFFu) | 0x3F800000u);
  if (m > 1.41421354f) { m = m * 0.5f; e += 1; }
  /* the quotient as the device forms it: linear seed, three Newton steps, one residual correction */
  float num = m - 1.0f, den = m + 1.0f;
  float y = fmaf(-0.24264069f, den, 0.99258476f);
  y = fmaf(y, fmaf(-den, y, 1.0f), y);
  y = fmaf(y, fmaf(-den, y, 1.0f), y);
  y = fmaf(y, fmaf(-den, y, 1.0f), y);
  float t = num * y;
  t = fmaf(fmaf(-t, den, num), y, t);
  float t2 = t * t;
  float q = 0x1.c71c72p-4f;             /* 1/9 */
  q = fmaf(q, t2, 0x1.24924ap-3f);      /* 1/7 */
  q = fmaf(q, t2, 0x1.99999ap-3f);      /* 1/5 */
  q = fmaf(q, t2, 0x1.555556p-2f);      /* 1/3 */
  q = fmaf(q, t2, 1.0f);
  float r = (t * q) * 0x1.715476p+1f;   /* 2/ln 2 */
  return (float)e + r;
}

void mvo_stageA_f32(const float* x, int D, const float* m, int cap, float* acc, float* xx) {
  float q = 0.0f;
  for (int dd = 0; dd < D; ++dd) q = fmaf(x[dd], x[dd], q);
  *xx = q;
  for (int t = 0; t < cap; ++t) {
    float a = 0.0f;
    const float* mt = m + (size_t)t * D;
    for (int dd = 0; dd < D; ++dd) a = fmaf(x[dd], mt[dd], a);
    acc[t] = a;
  }
}

/* Restates HalfEpilogue / merge_view / RowEpilogue::finish of multiview-clustering_b200/csrc/mv_device.cuh
 * operation by operation.  The cap tables are two halves of cap/2; per view each half streams a
 * log-sum-exp over its dishes in chunks of 16 tables (2-way chunk max, rescale, 4-way partial sums),
 * the halves are merged (half A first) together with the new-dish term.  The draw: global max, per-half
 * totals (4-way partials + tree), total = (HA + HB) + new, per-half prefix counts (half B starts at HA).
 * margin_out, if not NULL, receives the distance of u*total to the nearest CDF edge relative to total
 * (how close the draw was to flipping) — used to grade engines whose weights are tolerance-level. */
int mvo_stageB_f32_ex(const mvo_params_f32* p, const float* acc, const float* xx, int t0, float uf,
                      float* lw_out, float* margin_out) {
  return mvo_stageB_f32_mixed(p, NULL, acc, xx, NULL, t0, uf, lw_out, margin_out);
}

void mvo_stageA_counts_f32(const int32_t* col, const float* val, int nnz, const float* l2t, const int32_t* cdt,
                           int cap, int t0, float beta, float wbeta_plus_ctot, float* acc, float* acc_loo, float* rowtot) {
  float tot = 0.0f;
  for (int j = 0; j < nnz; ++j) tot = tot + val[j];
  *rowtot = tot;
  for (int t = 0; t < cap; ++t) {
    float a = 0.0f;
    for (int j = 0; j < nnz; ++j) a = fmaf(val[j], l2t[(size_t)col[j] * cap + t], a);
    acc[t] = a;
  }
  const float lden = mvo_log2m(wbeta_plus_ctot - tot);
  float a = 0.0f;
  for (int j = 0; j < nnz; ++j) {
    const float c = (float)cdt[(size_t)col[j] * cap + t0];
    a = fmaf(val[j], mvo_log2m((beta + c) - val[j]) - lden, a);
  }
  *acc_loo = a;
}

int mvo_stageB_f32_mixed(const mvo_params_f32* p, const int32_t* kind, const float* acc, const float* xx,
                         const float* acc_loo, int t0, float uf, float* lw_out, float* margin_out) {
  const int V = p->V, cap = p->cap, half = cap / 2;
  float* lw = (float*)malloc(sizeof(float) * (size_t)cap);
  float* term = (float*)malloc(sizeof(float) * (size_t)cap);
  const int single = p->single[t0];
  float lnew = single ? p->LMN[1] : p->LMN[0];
  for (int t = 0; t < cap; ++t) lw[t] = (t == t0) ? p->LM1[t0] : p->LM[t];
  for (int v = 0; v < V; ++v) {
    const size_t o = (size_t)v * cap;
    const int k0 = p->dish[o + t0];
    const int counts = (kind != NULL && kind[v] != 0);
    const float A1r = counts ? 0.0f : p->A1[o + t0], C1r = counts ? acc_loo[v] : p->C1[o + t0];
    const float nxx = counts ? -0.0f : -xx[v];
    float hmx[2], hs[2];
    for (int h = 0; h < 2; ++h) {
      float mx = MVO_MASKED, s = 0.0f;
      for (int base = h * half; base < (h + 1) * half; base += 16) {   /* kEpiChunk = 16 tables at a time */
        float cp[2] = {MVO_MASKED, MVO_MASKED};
        for (int j = 0; j < 16; ++j) {
          const int t = base + j;
          float e = fmaf(2.0f, acc[o + t], nxx);
          int same = (p->dish[o + t] == k0);
          float L = same ? fmaf(A1r, e, C1r) : fmaf(p->A[o + t], e, p->C[o + t]);
          lw[t] = lw[t] + L;
          float w = (same && single) ? p->W1[o + t] : p->W[o + t];
          term[t] = L + w;
          cp[j & 1] = fmaxf(cp[j & 1], term[t]);
        }
        float mn = fmaxf(mx, fmaxf(cp[0], cp[1]));
        s = s * mvo_exp2m(mx - mn);
        mx = mn;
        float sp[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int j = 0; j < 16; ++j) sp[j & 3] = sp[j & 3] + mvo_exp2m(term[base + j] - mn);
        s = s + ((sp[0] + sp[1]) + (sp[2] + sp[3]));
      }
      hmx[h] = mx; hs[h] = s;
    }
    /* merge_view */
    float mn = fmaxf(hmx[0], hmx[1]);
    float s = (hs[0] * mvo_exp2m(hmx[0] - mn)) + (hs[1] * mvo_exp2m(hmx[1] - mn));
    float Lnew = fmaf(-p->AN[v], xx[v], p->CN[v]);
    float termnew = Lnew + ((single && p->lone[o + t0]) ? p->WN[2 * v + 1] : p->WN[2 * v]);
    float m2 = fmaxf(mn, termnew);
    s = s * mvo_exp2m(mn - m2);
    s = s + mvo_exp2m(termnew - m2);
    float logmarg = (m2 + mvo_log2m(s)) - (single ? p->LD[2 * v + 1] : p->LD[2 * v]);
    lnew = lnew + logmarg;
  }
  float M = lnew;
  for (int t = 0; t < cap; ++t) M = fmaxf(M, lw[t]);
  if (lw_out) { memcpy(lw_out, lw, sizeof(float) * (size_t)cap); lw_out[cap] = lnew; }
  if (margin_out) *margin_out = 1.0f;
  int choice = MVO_NEW;
  if (!(M > -1.0e29f)) {
    choice = t0;
  } else {
    float H[2];
    for (int h = 0; h < 2; ++h) {
      float qp[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      for (int j = 0; j < half; ++j) {
        const int t = h * half + j;
        term[t] = mvo_exp2m(lw[t] - M);
        qp[j & 3] = qp[j & 3] + term[t];
      }
      H[h] = (qp[0] + qp[1]) + (qp[2] + qp[3]);
    }
    float total = (H[0] + H[1]) + mvo_exp2m(lnew - M);
    float target = uf * total;
    float margin = 1.0f;
    int cnt = 0;
    for (int h = 0; h < 2; ++h) {
      float cum = h ? H[0] : 0.0f;
      for (int j = 0; j < half; ++j) {
        cum = cum + term[h * half + j];
        cnt += (target < cum) ? 0 : 1;
        float dist = fabsf(target - cum) / total;
        if (dist < margin) margin = dist;
      }
    }
    if (margin_out) *margin_out = margin;
    choice = (cnt < cap) ? cnt : MVO_NEW;
    if (cnt >= cap && !(lnew > -1.0e29f)) {
      choice = t0;
      for (int t = 0; t < cap; ++t) if (term[t] > 1.0e-30f) choice = t;
    }
  }
  free(lw); free(term);
  return choice;
}

/* The tensor-core engine's epilogue (multiview-clustering_b200/csrc/mv_draw_tc.cu), restated operation by operation.
 * acc[v][t] = x_v . b_vt with the PRE-SCALED means b = 2 A m (mvo_scaled_means), so that
 *     lw[t] = base[t] + sum_v A_vt (-|x_v|^2) + sum_v acc[v][t]        base[t] = LM[t] + sum_v C_vt  (view order)
 * is the log2 weight of table t with the customer still counted.  Removing it (multiview_utils.cpp:138-192) is a
 * scalar correction per view from the dot product at its own table t0,
 *     plain_v = acc[v][t0] + (A_vt0 (-|x_v|^2) + C_vt0),   loo_v = R_vt0 acc[v][t0] + (A1_vt0 (-|x_v|^2) + C1_vt0),
 *     delta_v = loo_v - plain_v,   R = A1 / A (FP32 division),
 * added (in view order, then LM1 - LM for t0 itself) to every table serving the same dish as t0 in view v.
 * lnew_dev: the log2 weight of a new table as the DEVICE evaluated it (a float statistic checked against the FP64
 * restatement by tolerance; the draw is bit-exact given it).  Then the common draw: global max, exact exp2, half
 * totals, inverse-CDF scan. */
static int draw_from_lw_f32(const float* lw, int cap, float lnew, int t0, float uf, float* margin_out) {
  const int half = cap / 2;
  float* term = (float*)malloc(sizeof(float) * (size_t)cap);
  float M = lnew;
  for (int t = 0; t < cap; ++t) M = fmaxf(M, lw[t]);
  if (margin_out) *margin_out = 1.0f;
  int choice = MVO_NEW;
  if (!(M > -1.0e29f)) {
    choice = t0;
  } else {
    float H[2];
    for (int h = 0; h < 2; ++h) {
      float qp[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      for (int j = 0; j < half; ++j) {
        const int t = h * half + j;
        term[t] = mvo_exp2m(lw[t] - M);
        qp[j & 3] = qp[j & 3] + term[t];
      }
      H[h] = (qp[0] + qp[1]) + (qp[2] + qp[3]);
    }
    float total = (H[0] + H[1]) + mvo_exp2m(lnew - M);
    float target = uf * total;
    float margin = 1.0f;
    int cnt = 0;
    for (int h = 0; h < 2; ++h) {
      float cum = h ? H[0] : 0.0f;
      for (int j = 0; j < half; ++j) {
        cum = cum + term[h * half + j];
        cnt += (target < cum) ? 0 : 1;
        float dist = fabsf(target - cum) / total;
        if (dist < margin) margin = dist;
      }
    }
    if (margin_out) *margin_out = margin;
    choice = (cnt < cap) ? cnt : MVO_NEW;
    if (cnt >= cap && !(lnew > -1.0e29f)) {
      choice = t0;
      for (int t = 0; t < cap; ++t) if (term[t] > 1.0e-30f) choice = t;
    }
  }
  free(term);
  return choice;
}

int mvo_stageB_tc(const mvo_params_f32* p, const float* acc, const float* xx, int t0, float uf, float lnew_dev,
                  float* lw_out, float* margin_out) {
  const int V = p->V, cap = p->cap;
  float* lw = (float*)malloc(sizeof(float) * (size_t)cap);
  float delta[64];
  if (V > 64) { free(lw); return -2; }
  for (int t = 0; t < cap; ++t) {
    float b = p->LM[t];
    for (int v = 0; v < V; ++v) b = b + p->C[(size_t)v * cap + t];
    for (int v = 0; v < V; ++v) b = fmaf(p->A[(size_t)v * cap + t], -xx[v], b);
    for (int v = 0; v < V; ++v) b = b + acc[(size_t)v * cap + t];
    lw[t] = b;
  }
  for (int v = 0; v < V; ++v) {
    const size_t o = (size_t)v * cap;
    const float nxx = -xx[v], a = acc[o + t0];
    const float A = p->A[o + t0];
    const float R = (A != 0.0f) ? p->A1[o + t0] / A : 0.0f;
    const float plain = a + fmaf(A, nxx, p->C[o + t0]);
    const float loo = fmaf(R, a, fmaf(p->A1[o + t0], nxx, p->C1[o + t0]));
    delta[v] = loo + (-plain);
  }
  for (int t = 0; t < cap; ++t) {
    float c = 0.0f;
    for (int v = 0; v < V; ++v) {
      const size_t o = (size_t)v * cap;
      if (p->dish[o + t] >= 0 && p->dish[o + t] == p->dish[o + t0]) c = c + delta[v];
    }
    if (t == t0) c = c + (p->LM1[t0] + (-p->LM[t0]));
    lw[t] = lw[t] + c;
  }
  if (lw_out) { memcpy(lw_out, lw, sizeof(float) * (size_t)cap); lw_out[cap] = lnew_dev; }
  const int choice = draw_from_lw_f32(lw, cap, lnew_dev, t0, uf, margin_out);
  free(lw);
  return choice;
}

/* b = (float)(2 A m): the B operand of the tensor-core engine for table t of one view, one rounding of the exact
 * product (k_finalize computes the same expression). */
void mvo_scaled_means(const float* A, const float* m, int cap, int D, float* b) {
  for (int t = 0; t < cap; ++t)
    for (int dd = 0; dd < D; ++dd) b[(size_t)t * D + dd] = (float)(2.0 * (double)A[t] * (double)m[(size_t)t * D + dd]);
}

int mvo_stageB_f32(const mvo_params_f32* p, const float* acc, const float* xx, int t0, float uf,
                   float* lw_out) {
  return mvo_stageB_f32_ex(p, acc, xx, t0, uf, lw_out, NULL);
}

int mvo_make_params(const mvo_state* s, int32_t* dish, float* A, float* C, float* A1, float* C1,
                    float* W, float* W1, int32_t* lone, float* AN, float* CN, float* WN, float* LD,
                    float* LM, float* LM1, int32_t* single, float* LMN, float* const* m) {
  const int V = s->V, cap = s->cap;
  const double LOG2E = 1.4426950408889634074;
  int T_ne = 0, F = 0;
  for (int t = 0; t < cap; ++t) { if (s->n_t[t] > 0) T_ne++; else F++; }
  for (int v = 0; v < V; ++v) {
    const int D = s->D[v];
    const double tau = s->tau_v[v], alpha = s->alpha_v[v], sigma = s->sigma_v[v];
    const size_t o = (size_t)v * cap;
    int K_act = 0;
    long sum_l = 0;
    for (int k = 0; k < cap; ++k) if (s->l_vk[o + k] > 0) { K_act++; sum_l += s->l_vk[o + k]; }
    for (int t = 0; t < cap; ++t) {
      const int k = (s->n_t[t] > 0) ? s->dish_of[o + t] : -1;
      dish[o + t] = k;
      float* mt = m[v] + (size_t)t * D;
      if (k < 0) {
        A[o + t] = 0.f; C[o + t] = MVO_MASKED; A1[o + t] = 0.f; C1[o + t] = MVO_MASKED;
        W[o + t] = MVO_MASKED; W1[o + t] = MVO_MASKED; lone[o + t] = 0;
        for (int dd = 0; dd < D; ++dd) mt[dd] = 0.f;
        continue;
      }
      if (is_counts(s, v)) {          /* acc is log2 f itself: log2 f = 0 + 0.5 * (2 acc - 0) */
        A[o + t] = 0.5f; C[o + t] = 0.f; A1[o + t] = 0.f; C1[o + t] = 0.f;
        int rep = 1;
        for (int t2 = 0; t2 < t; ++t2) if (s->n_t[t2] > 0 && s->dish_of[o + t2] == k) { rep = 0; break; }
        double w = (double)s->l_vk[o + k] - sigma, w1 = w - 1.0;
        W[o + t] = (rep && w > 0.0) ? (float)log2(w) : MVO_MASKED;
        W1[o + t] = (rep && w1 > 0.0) ? (float)log2(w1) : MVO_MASKED;
        lone[o + t] = (s->l_vk[o + k] == 1);
        continue;
      }
      const double n = (double)s->n_vk[o + k];
      double mm = 0.0;
      for (int dd = 0; dd < D; ++dd) {
        mt[dd] = (float)(s->S1[v][(size_t)k * D + dd] / (tau + n));
        mm += (double)mt[dd] * (double)mt[dd];
      }
      double a = (tau + n) / (2.0 * tau * (tau + n + 1.0));
      double c = -0.5 * log(2.0 * M_PI * tau * (tau + n + 1.0) / (tau + n));
      A[o + t] = (float)(LOG2E * a);
      C[o + t] = (float)(LOG2E * ((double)D * c - a * mm));
      if (n >= 2.0) {
        double a1 = (tau + n) / (2.0 * tau * (tau + n - 1.0));
        double c1 = -0.5 * log(2.0 * M_PI * tau * (tau + n) / (tau + n - 1.0));
        A1[o + t] = (float)(LOG2E * a1);
        C1[o + t] = (float)(LOG2E * ((double)D * c1 - a1 * mm));
      } else {
        A1[o + t] = 0.f; C1[o + t] = MVO_MASKED;
      }
      int rep = 1;
      for (int t2 = 0; t2 < t; ++t2) if (s->n_t[t2] > 0 && s->dish_of[o + t2] == k) { rep = 0; break; }
      double w = (double)s->l_vk[o + k] - sigma, w1 = w - 1.0;
      W[o + t] = (rep && w > 0.0) ? (float)log2(w) : MVO_MASKED;
      W1[o + t] = (rep && w1 > 0.0) ? (float)log2(w1) : MVO_MASKED;
      lone[o + t] = (s->l_vk[o + k] == 1);
    }
    AN[v] = (float)(LOG2E / (2.0 * tau));
    CN[v] = (float)(LOG2E * (-0.5 * (double)D * log(2.0 * M_PI * tau)));
    if (is_counts(s, v)) { AN[v] = (float)log2((double)s->csr[v].vocab); CN[v] = 0.f; }   /* log2 f_new = -|x| log2 W */
    double wn0 = alpha + (double)K_act * sigma, wn1 = alpha + (double)(K_act - 1) * sigma;
    WN[2 * v] = wn0 > 0.0 ? (float)log2(wn0) : MVO_MASKED;
    WN[2 * v + 1] = wn1 > 0.0 ? (float)log2(wn1) : MVO_MASKED;
    double d0 = alpha + (double)sum_l, d1 = alpha + (double)(sum_l - 1);
    LD[2 * v] = d0 > 0.0 ? (float)log2(d0) : 0.f;
    LD[2 * v + 1] = d1 > 0.0 ? (float)log2(d1) : 0.f;
  }
  for (int t = 0; t < cap; ++t) {
    double mass = (double)s->n_t[t] - s->sigma_g, mass1 = mass - 1.0;
    LM[t] = (s->n_t[t] > 0 && mass > 0.0) ? (float)log2(mass) : MVO_MASKED;
    LM1[t] = (s->n_t[t] > 1 && mass1 > 0.0) ? (float)log2(mass1) : MVO_MASKED;
    single[t] = (s->n_t[t] == 1);
  }
  double mn0 = s->alpha_g + s->sigma_g * (double)T_ne, mn1 = s->alpha_g + s->sigma_g * (double)(T_ne - 1);
  LMN[0] = (F > 0 && mn0 > 0.0) ? (float)log2(mn0) : MVO_MASKED;
  LMN[1] = (F > 0 && mn1 > 0.0) ? (float)log2(mn1) : MVO_MASKED;
  return 0;
}
