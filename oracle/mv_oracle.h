/* oracle/mv_oracle.h — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * PARITY STATUS
 *   D = 1 likelihood, table weights, dish sampling, hyper log-posteriors: PINNED against the
 *     compiled, unmodified reference (oracle/_ref/libmvref.so) and against the known answers
 *     of SURVEY.md Appendix B (tests/test_oracle_vs_reference.py, tests/golden/).
 *   D > 1 isotropic generalisation and count views (A.3): the reference has no
 *     such code — "parity unpinned"; they reduce to the pinned D = 1 path (tested).
 *   Synchronous sweep: the reference sweep is sequential (multiview_gibbs.cpp:157-200).  The
 *     data-parallel sweep is a different Markov kernel built from the reference's per-customer
 *     arithmetic; per-row it equals "remove_customer(i) on the sweep-start state, then
 *     compute_table_probs_with_cache(i)" of the reference (pinned), the composition is ours.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.
 *
 * Layout conventions (shared with the CUDA library):
 *   cap        capacity of table slots and of dish slots per view (T_cap = K_cap)
 *   table_of   [n]       table slot of every row, always in [0,cap)
 *   n_t        [cap]     rows per table slot (0 = free slot)
 *   dish_of    [V*cap]   dish slot eaten by table t in view v, -1 for a free table slot
 *   n_vk,l_vk  [V*cap]   rows / tables per dish slot
 *   S1         per view [cap*D]  sum of rows per dish slot  (ViewState::sum_y,  multiview_state.h:11)
 *   S2         [V*cap]   sum of squared norms per dish slot (ViewState::sum_y2, multiview_state.h:12)
 */
#ifndef MV_ORACLE_H
#define MV_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVO_NEW (-1)           /* choice value: "open a new table" */
#define MVO_MASKED (-1.0e30f)  /* FP32 sentinel for a zero-weight option in the log2 domain */

/* A sparse COUNT view in CSR form (no reference counterpart, parity unpinned).  Model: the rows of a
 * dish are draws from one multinomial over the vocabulary whose probabilities are the plug-in estimate
 * theta_kw = (beta + c_kw) / (W beta + C_k) from the dish's word counts c_kw (C_k their total), so that
 *   log f_vk(x) = sum_w x_w log theta_kw            (the multinomial coefficient cancels across k),
 * with the row's own counts removed first for its own dish, and log f_new(x) = -|x| log W (an empty dish). */
typedef struct mvo_csr {
  const int32_t* rowptr;   /* [n+1]; NULL: view v is dense */
  const int32_t* col;      /* [nnz] in [0, vocab) */
  const float* val;        /* [nnz] non-negative integer counts */
  int32_t vocab;
  int32_t pad;
  int64_t* cd;             /* [cap*vocab] word counts per dish slot, rebuilt by mvo_rebuild_stats */
  int64_t* ctot;           /* [cap] their totals */
} mvo_csr;

typedef struct mvo_state {
  int32_t n;            /* rows held here (all of them: the oracle is single-process) */
  int32_t V;            /* views */
  int32_t cap;          /* table/dish slot capacity */
  int32_t reserved0;
  int64_t row_offset;   /* global index of row 0, for Philox addressing */
  int64_t n_global;     /* total customers, enters log_global_EPPF (multiview_hyper.cpp:66) */
  const int32_t* D;     /* [V] feature dimension of each dense view */
  const float* const* x;/* [V] row-major [n][D_v], the FP32 data the GPU sees */
  int32_t* table_of;
  int32_t* n_t;
  int32_t* dish_of;
  int32_t* n_vk;
  int32_t* l_vk;
  double* const* S1;    /* [V] -> [cap*D_v] */
  double* S2;           /* [V*cap] */
  double* alpha_v;      /* [V] */
  double* sigma_v;      /* [V] */
  double* tau_v;        /* [V] */
  double alpha_g;
  double sigma_g;
  uint64_t seed;
  uint32_t chain;
  uint32_t sweep;       /* index of the NEXT sweep to run */
  const mvo_csr* csr;   /* [V] or NULL (all views dense); a count view has D[v] = 0 */
  double count_beta;    /* symmetric Dirichlet pseudo-count of the count views */
} mvo_state;

/* --- Philox (host restatement) ------------------------------------------------------- */
void mvo_philox_block(uint64_t seed, uint32_t chain, uint32_t domain, uint32_t slot, uint32_t sweep,
                      uint64_t index, uint32_t out[4]);
void mvo_philox_raw(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
float mvo_uf(uint64_t seed, uint32_t chain, uint32_t domain, uint32_t slot, uint32_t sweep, uint64_t index);
double mvo_u53(uint64_t seed, uint32_t chain, uint32_t domain, uint32_t slot, uint32_t sweep, uint64_t index);
double mvo_z(uint64_t seed, uint32_t chain, uint32_t domain, uint32_t slot, uint32_t sweep, uint64_t index);

/* --- state ---------------------------------------------------------------------------- */
/* Rebuild n_t, n_vk, l_vk, S1, S2 from table_of + dish_of (the loop of multiview_gibbs.cpp:64-73). */
int mvo_rebuild_stats(mvo_state* s);
/* The reference's random initialisation (multiview_gibbs.cpp:12-103) on Philox domains 4/5. */
int mvo_init_reference(mvo_state* s);

/* --- FP64 restatement of the per-row arithmetic ---------------------------------------- */
/* log posterior-predictive density of row x (length D_v) under dish k of view v
 * (multiview_utils.cpp:307-338 generalised per SURVEY.md A.2); loo != 0 removes x from the dish first. */
double mvo_log_f_vk(const mvo_state* s, int v, int k, const float* x, int loo);
double mvo_log_f_new(const mvo_state* s, int v, const float* x);   /* multiview_utils.cpp:340-350 */
/* Natural-log weights of the cap tables and (index cap) of a new table for row i, with row i
 * removed from the sweep-start state (multiview_utils.cpp:71-136, :40-69).  lw has cap+1 entries.
 * Lvt, if not NULL, receives V*(cap+1) log f values: [v][t] by table, [v][cap] = log f_new. */
int mvo_row_logweights(const mvo_state* s, int i, double* lw, double* Lvt);
/* The draw of multiview_gibbs.cpp:169-199 from log-weights (max-subtracted) and uniform u. */
int mvo_draw_from_logweights(const mvo_state* s, int i, const double* lw, double u);
/* All rows: choice[i] in [0,cap) or MVO_NEW.  threads <= 0 uses one thread. */
int mvo_draw_rows(const mvo_state* s, int32_t* choice, int threads);
/* Births in row order / overflow / deaths; rewrites table_of, dish_of, then rebuilds stats.
 * birth_rows/birth_w, if not NULL, receive the seated birth rows and the V*(cap+1) dish
 * weights each one saw (max-normalised), for comparison with the device's export. */
int mvo_reseat(mvo_state* s, const int32_t* choice, int32_t* n_seated, int32_t* birth_rows, double* birth_w);
/* update_hyperparameters() (multiview_hyper.cpp:233-292).  If z/u are not NULL they replace the
 * Philox normals/uniforms (3V+2 of each, in call order).  use_lgamma != 0 evaluates the EPPF sums in
 * closed form (what the GPU does) instead of the reference's loops. */
int mvo_hyper_step(mvo_state* s, const double* z, const double* u, int use_lgamma);
double mvo_log_EPPF_view(const mvo_state* s, int v, double alpha, double sigma, int use_lgamma);
double mvo_log_EPPF_global(const mvo_state* s, double alpha, double sigma, int use_lgamma);
double mvo_log_posterior_tau(const mvo_state* s, int v, double tau);
/* One or more full synchronous sweeps: draws, reseat, stats rebuild, hyper step. */
int mvo_sweep(mvo_state* s, int n_sweeps, int threads, int do_hyper);

/* --- FP32 mirror of the device's per-row epilogue --------------------------------------- */
typedef struct mvo_params_f32 {   /* the per-sweep parameter block, table-major */
  int32_t V, cap;
  const int32_t* dish;   /* [V*cap] */
  const float* A;        /* [V*cap] */
  const float* C;        /* [V*cap] */
  const float* A1;       /* [V*cap] leave-one-out variants */
  const float* C1;       /* [V*cap] */
  const float* W;        /* [V*cap] log2 dish weight (representative table only) */
  const float* W1;       /* [V*cap] same with l-1 */
  const int32_t* lone;   /* [V*cap] 1 if the table's dish is served by exactly one table */
  const float* AN;       /* [V] */
  const float* CN;       /* [V] */
  const float* WN;       /* [V*2] */
  const float* LD;       /* [V*2] */
  const float* LM;       /* [cap] */
  const float* LM1;      /* [cap] */
  const int32_t* single; /* [cap] n_t == 1 */
  const float* LMN;      /* [2] */
} mvo_params_f32;

/* Fill an FP32 parameter block (and the per-table means m[v] = [cap*D_v]) from the FP64 state;
 * every array must be preallocated by the caller. */
int mvo_make_params(const mvo_state* s, int32_t* dish, float* A, float* C, float* A1, float* C1,
                    float* W, float* W1, int32_t* lone, float* AN, float* CN, float* WN, float* LD,
                    float* LM, float* LM1, int32_t* single, float* LMN, float* const* m);
/* Stage A on CUDA cores, restated: acc[t] = sum_d x[d]*m[t][d] as one ascending fmaf chain,
 * xx = sum_d x[d]^2 likewise.  acc has cap entries. */
void mvo_stageA_f32(const float* x, int D, const float* m, int cap, float* acc, float* xx);
/* Stage B: from per-view dot products acc[v][t] (V*cap) and squared norms xx[v] of ONE row, its
 * current table t0 and uniform uf, reproduce the device's choice bit for bit. */
int mvo_stageB_f32(const mvo_params_f32* p, const float* acc, const float* xx, int t0, float uf,
                   float* lw_out /* cap+1 or NULL */);
/* Same, also reporting how close u*total came to a CDF edge (relative to total). */
/* The tensor-core engine's draw stage (pre-scaled means, leave-one-out as a scalar correction); see mv_oracle.c. */
int mvo_stageB_tc(const mvo_params_f32* p, const float* acc, const float* xx, int t0, float uf, float lnew_dev,
                  float* lw_out, float* margin_out);
void mvo_scaled_means(const float* A, const float* m, int cap, int D, float* b);
int mvo_stageB_f32_ex(const mvo_params_f32* p, const float* acc, const float* xx, int t0, float uf,
                      float* lw_out, float* margin_out);
/* Same for a mix of dense and count views: kind[v] != 0 marks a count view, for which acc[v][t] is already
 * log2 f under table t's dish, acc_loo[v] the leave-one-out value for the row's own dish and xx[v] the row's
 * total count (it only enters the new-dish term).  kind == NULL: all dense. */
int mvo_stageB_f32_mixed(const mvo_params_f32* p, const int32_t* kind, const float* acc, const float* xx,
                         const float* acc_loo, int t0, float uf, float* lw_out, float* margin_out);
/* Stage A of a count view on CUDA cores, restated: acc[t] = sum_j val_j * l2t[col_j*cap + t] as one ascending
 * fmaf chain over the row's nonzeros; *acc_loo = sum_j val_j * (log2m(beta + cdt[col_j*cap + t0] - val_j) -
 * log2m(W beta + ctot_t0 - rowtot)) likewise; *rowtot = sum_j val_j.  l2t / cdt are the device's tables:
 * [vocab][cap] log2 theta and dish word counts as seen from each table slot. */
void mvo_stageA_counts_f32(const int32_t* col, const float* val, int nnz, const float* l2t, const int32_t* cdt,
                           int cap, int t0, float beta, float wbeta_plus_ctot, float* acc, float* acc_loo, float* rowtot);
float mvo_exp2m(float d);
float mvo_log2m(float s);

#ifdef __cplusplus
}
#endif
#endif /* MV_ORACLE_H */
