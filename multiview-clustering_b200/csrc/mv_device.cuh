// mv_device.cuh — shared device-side definitions of the allocation sweep:
//   * the per-sweep FP32 parameter block (table-major) the likelihood+draw kernels read,
//   * the bit-reproducible exp2 / log2 used by the draw,
//   * RowEpilogue<CAP>: leave-one-out table weights, log-sum-exp marginal of a new table and the
//     inverse-CDF draw for ONE customer held by ONE thread.
//
// Reference arithmetic being replaced (paths under /root/reference/Multiview):
//   compute_f_vk / compute_f_vk_new               multiview_utils.cpp:307-350
//   compute_marginal_likelihood_new_table          multiview_utils.cpp:40-69
//   compute_table_probs_with_cache                 multiview_utils.cpp:71-136
//   normalise + inverse-CDF draw                   multiview_gibbs.cpp:169-199
//   remove_customer (as a leave-one-out view)      multiview_utils.cpp:138-192
//
// Every FP32 operation of the epilogue is an explicit round-to-nearest intrinsic in a fixed
// order, so that oracle/mv_oracle.c:mvo_stageB_f32 (an independent plain-C restatement) produces
// the same integer draw bit for bit from the same dot products.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mv_philox.h"

namespace mv {

constexpr int kMaxViews = 16;
constexpr float kMasked = -1.0e30f;   // log2-domain sentinel of a zero-weight option
constexpr int kNewTable = -1;

// ---- per-sweep parameter block (built by the finalize kernel) ----------------
struct __align__(16) TableParam {   // one per (view, table slot); 32 bytes
  float A, C;        // log2 f = C + A*e,  e = 2 x.m - |x|^2      (dish statistics as they are)
  float A1, C1;      // same with the customer itself removed from the dish (n-1, S1-x)
  float W, W1;       // log2 (l_vk - sigma_v)+ carried by the lowest table of each dish; W1: l_vk-1
  int32_t dish;      // dish slot of this table in this view, -1 = free table slot
  int32_t lone;      // 1 if that dish is served by exactly one table
};
struct __align__(16) ViewParam {    // one per view; 32 bytes
  float AN, CN;      // log2 f_new = CN - AN |x|^2
  float WN0, WN1;    // log2 (alpha_v + K_act sigma_v)+ ; WN1 with K_act-1
  float LD0, LD1;    // log2 (alpha_v + sum_k l_vk) ; LD1 with the sum reduced by one
  float pad0, pad1;
};
struct __align__(16) TableMass {    // one per table slot; 16 bytes
  float LM, LM1;     // log2 (n_t - sigma_g)+ ; LM1 with n_t-1
  int32_t single;    // n_t == 1
  int32_t pad;
};
struct __align__(16) GlobalParam {
  float LMN0, LMN1;  // log2 (alpha_g + sigma_g T_nonempty)+ ; LMN1 with T_nonempty-1; masked if no free slot
  int32_t nfree;     // free table slots at sweep start
  uint32_t sweep;    // index of the sweep these parameters are for
};

// ---- bit-reproducible transcendental pieces ---------------------------------------------
// 2^d for d <= 0 (clamped at -125): n = rint(d) via the 1.5*2^23 trick, 2^f by a degree-5
// polynomial on [-1/2,1/2] (rel. err 1.9e-7), exponent added in the integer domain.
__device__ __forceinline__ float exp2m(float d) {
  d = fmaxf(d, -125.0f);
  const float r = __fadd_rn(d, 12582912.0f);
  const float n = __fadd_rn(r, -12582912.0f);
  const float f = __fadd_rn(d, -n);
  float p = 0x1.5bba14p-10f;
  p = __fmaf_rn(p, f, 0x1.3cea88p-7f);
  p = __fmaf_rn(p, f, 0x1.c6b752p-5f);
  p = __fmaf_rn(p, f, 0x1.ebf9bcp-3f);
  p = __fmaf_rn(p, f, 0x1.62e42ap-1f);
  p = __fmaf_rn(p, f, 1.0f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(r) << 23));
}

// log2(s) for a normal s > 0: atanh series in t = (m-1)/(m+1), m in [sqrt(1/2), sqrt(2)).
__device__ __forceinline__ float log2m(float s) {
  const uint32_t b = __float_as_uint(s);
  int32_t e = (int32_t)(b >> 23) - 127;
  float m = __uint_as_float((b & 0x007FFFFFu) | 0x3F800000u);
  if (m > 1.41421354f) { m = __fmul_rn(m, 0.5f); e += 1; }
  const float t = __fdiv_rn(__fadd_rn(m, -1.0f), __fadd_rn(m, 1.0f));
  const float t2 = __fmul_rn(t, t);
  float q = 0x1.c71c72p-4f;
  q = __fmaf_rn(q, t2, 0x1.24924ap-3f);
  q = __fmaf_rn(q, t2, 0x1.99999ap-3f);
  q = __fmaf_rn(q, t2, 0x1.555556p-2f);
  q = __fmaf_rn(q, t2, 1.0f);
  const float r = __fmul_rn(__fmul_rn(t, q), 0x1.715476p+1f);
  return __fadd_rn((float)e, r);
}

// ---- one customer's epilogue ---------------------------------------------------------------
template <int CAP>
struct RowEpilogue {
  float lw[CAP];   // running log2 weight of each table
  float lnew;      // running log2 weight of a new table
  int t0;          // current table of the customer
  int single;      // the customer sits alone at t0

  __device__ __forceinline__ void begin(const TableMass* __restrict__ tm, const GlobalParam& g, int t0_) {
    t0 = t0_;
    single = tm[t0_].single;
    lnew = single ? g.LMN1 : g.LMN0;
  }

  // acc[t] = x . m_{v,t} (consumed and overwritten), xx = |x|^2; tp = this view's TableParam[CAP].
  __device__ __forceinline__ void view(const TableParam* __restrict__ tp, const ViewParam& vp,
                                       float (&acc)[CAP], float xx, bool first) {
    const int k0 = tp[t0].dish;
    const int lone0 = tp[t0].lone;
    const float nxx = -xx;
#pragma unroll
    for (int t = 0; t < CAP; ++t) {
      const TableParam q = tp[t];
      const float e = __fmaf_rn(2.0f, acc[t], nxx);
      const bool same = (q.dish == k0);
      const float L = same ? __fmaf_rn(q.A1, e, q.C1) : __fmaf_rn(q.A, e, q.C);
      lw[t] = first ? L : __fadd_rn(lw[t], L);
      const float w = (same && single) ? q.W1 : q.W;
      acc[t] = __fadd_rn(L, w);
    }
    const float Lnew = __fmaf_rn(-vp.AN, xx, vp.CN);
    const float termnew = __fadd_rn(Lnew, (single && lone0) ? vp.WN1 : vp.WN0);
    float mx = termnew;
#pragma unroll
    for (int t = 0; t < CAP; ++t) mx = fmaxf(mx, acc[t]);
    float s = exp2m(__fadd_rn(termnew, -mx));
#pragma unroll
    for (int t = 0; t < CAP; ++t) s = __fadd_rn(s, exp2m(__fadd_rn(acc[t], -mx)));
    const float logmarg = __fadd_rn(__fadd_rn(mx, log2m(s)), -(single ? vp.LD1 : vp.LD0));
    lnew = __fadd_rn(lnew, logmarg);
  }

  // uf in (0,1). Returns the table slot or kNewTable.  lw[] is left holding the probabilities.
  __device__ __forceinline__ int finish(const TableMass* __restrict__ tm, float uf) {
#pragma unroll
    for (int t = 0; t < CAP; ++t) lw[t] = __fadd_rn(lw[t], (t == t0) ? tm[t].LM1 : tm[t].LM);
    float M = lnew;
#pragma unroll
    for (int t = 0; t < CAP; ++t) M = fmaxf(M, lw[t]);
    if (!(M > -1.0e29f)) return t0;   // nothing has weight: stay (cf. multiview_gibbs.cpp:172-176)
    const bool new_masked = !(lnew > -1.0e29f);
    int last_live = t0;
    float total = exp2m(__fadd_rn(lnew, -M));
#pragma unroll
    for (int t = 0; t < CAP; ++t) {
      if (lw[t] > -1.0e29f) last_live = t;
      lw[t] = exp2m(__fadd_rn(lw[t], -M));
      total = __fadd_rn(total, lw[t]);
    }
    const float target = __fmul_rn(uf, total);
    float cum = 0.0f;
    int choice = kNewTable;
    bool found = false;
#pragma unroll
    for (int t = 0; t < CAP; ++t) {
      cum = __fadd_rn(cum, lw[t]);
      if (!found && target < cum) { choice = t; found = true; }
    }
    if (!found && new_masked) choice = last_live;   // rounding fall-through with no new-table mass
    return choice;
  }
};

}  // namespace mv
