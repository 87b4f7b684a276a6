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
struct __align__(16) TableParam {   // one per (view, table slot); 32 bytes, hot half first
  float A, C;        // log2 f = C + A*e,  e = 2 x.m - |x|^2      (dish statistics as they are)
  float W;           // log2 (l_vk - sigma_v)+ carried by the lowest table of each dish, else masked
  int32_t dish;      // dish slot of this table in this view, -1 = free table slot
  float A1, C1;      // A, C with the customer itself removed from the dish (n-1, S1-x)
  float W1;          // W with l_vk-1
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

// log2(s) for a normal s > 0: atanh series in t = (m-1)/(m+1), m in [sqrt(1/2), sqrt(2)).  The
// quotient is a fixed sequence of fused multiply-adds (linear seed on [1.707, 2.414], three Newton
// steps, one residual correction): no division subroutine, and the CPU mirror repeats it exactly.
__device__ __forceinline__ float log2m(float s) {
  const uint32_t b = __float_as_uint(s);
  int32_t e = (int32_t)(b >> 23) - 127;
  float m = __uint_as_float((b & 0x007FFFFFu) | 0x3F800000u);
  if (m > 1.41421354f) { m = __fmul_rn(m, 0.5f); e += 1; }
  const float num = __fadd_rn(m, -1.0f), den = __fadd_rn(m, 1.0f);
  float y = __fmaf_rn(-0.24264069f, den, 0.99258476f);
  y = __fmaf_rn(y, __fmaf_rn(-den, y, 1.0f), y);
  y = __fmaf_rn(y, __fmaf_rn(-den, y, 1.0f), y);
  y = __fmaf_rn(y, __fmaf_rn(-den, y, 1.0f), y);
  float t = __fmul_rn(num, y);
  t = __fmaf_rn(__fmaf_rn(-t, den, num), y, t);
  const float t2 = __fmul_rn(t, t);
  float q = 0x1.c71c72p-4f;
  q = __fmaf_rn(q, t2, 0x1.24924ap-3f);
  q = __fmaf_rn(q, t2, 0x1.99999ap-3f);
  q = __fmaf_rn(q, t2, 0x1.555556p-2f);
  q = __fmaf_rn(q, t2, 1.0f);
  const float r = __fmul_rn(__fmul_rn(t, q), 0x1.715476p+1f);
  return __fadd_rn((float)e, r);
}

// 2^d for the weights.  EXACT: the polynomial above (the CPU mirror reproduces it bit for bit).
// FAST: one MUFU.EX2 (ex2.approx.ftz, <= 2 ulp) — the weights are then tolerance-level and only
// the scan that turns them into a table index is mirrored exactly.
template <bool FAST>
__device__ __forceinline__ float exp2w(float d) {
  if (FAST) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return r;
  } else {
    return exp2m(d);
  }
}

// ---- one customer's epilogue ---------------------------------------------------------------
// Order of operations (restated by oracle/mv_oracle.c:mvo_stageB_f32):
//   begin       lw[t] = log2 mass of table t (customer removed), lnew = log2 mass of a new table
//   view_begin / view_chunk x CAP/16 / view_end, once per view:
//               L[t] = log2 f under table t's dish (leave-one-out for the customer's own dish),
//               lw[t] += L[t]; streaming log-sum-exp over the dishes in chunks of 16 tables
//               (running max mx, sum s rescaled when mx moves), then the new-dish term
//               -> marginal of a new table -> lnew
//   finish      max-normalised weights, total, inverse-CDF count
// Chunks of 16 keep the live state at lw[CAP] + 16 terms, so the tensor-core kernel can feed a chunk
// straight from one tcgen05.ld.x16 and stay inside its register budget (measured: no spills at 168).
constexpr int kEpiChunk = 16;

template <int CAP, bool FAST = false>
struct RowEpilogue {
  static_assert(CAP % kEpiChunk == 0, "table capacity must be a multiple of the epilogue chunk");
  float lw[CAP];   // running log2 weight of each table
  float lnew;      // running log2 weight of a new table
  int t0;          // current table of the customer
  int single;      // the customer sits alone at t0
  // per-view state
  float mx, s, nxx, A1r, C1r;
  int k0, lone0, any_single;

  __device__ __forceinline__ void begin(const TableMass* __restrict__ tm, const GlobalParam& g, int t0_) {
    t0 = t0_;
    const TableMass own = tm[t0_];
    single = own.single;
    any_single = __any_sync(0xffffffffu, single);   // customers alone at their table are rare: warp-uniform slow path
    lnew = single ? g.LMN1 : g.LMN0;
#pragma unroll
    for (int t = 0; t < CAP; ++t) lw[t] = (t == t0_) ? own.LM1 : tm[t].LM;
  }

  __device__ __forceinline__ void view_begin(const TableParam* __restrict__ tp, float xx) {
    const TableParam own = tp[t0];
    k0 = own.dish; lone0 = own.lone; A1r = own.A1; C1r = own.C1;
    nxx = -xx;
    mx = kMasked;
    s = 0.0f;
  }

  // acc[j] = x . m_{v, base + j} for the kEpiChunk tables of chunk `base / kEpiChunk` (consumed).
  template <int BASE, bool SINGLE>
  __device__ __forceinline__ void chunk_impl(const TableParam* __restrict__ tp, float (&acc)[kEpiChunk]) {
    float c0 = kMasked, c1 = kMasked;
#pragma unroll
    for (int j = 0; j < kEpiChunk; ++j) {
      const int t = BASE + j;
      const float4 q = *reinterpret_cast<const float4*>(&tp[t]);          // A, C, W, dish
      const float e = __fmaf_rn(2.0f, acc[j], nxx);
      const bool same = (__float_as_int(q.w) == k0);
      const float L = same ? __fmaf_rn(A1r, e, C1r) : __fmaf_rn(q.x, e, q.y);
      lw[t] = __fadd_rn(lw[t], L);
      float w = q.z;
      if (SINGLE) w = (same && single) ? tp[t].W1 : q.z;
      const float term = __fadd_rn(L, w);
      acc[j] = term;
      if (j & 1) c1 = fmaxf(c1, term); else c0 = fmaxf(c0, term);
      if ((j & 7) == 7) asm volatile("" ::: "memory");   // keep at most 8 parameter loads in flight (registers)
    }
    const float mn = fmaxf(mx, fmaxf(c0, c1));
    s = __fmul_rn(s, exp2w<FAST>(__fadd_rn(mx, -mn)));
    mx = mn;
    float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;
#pragma unroll
    for (int j = 0; j < kEpiChunk; j += 4) {
      p0 = __fadd_rn(p0, exp2w<FAST>(__fadd_rn(acc[j], -mn)));
      p1 = __fadd_rn(p1, exp2w<FAST>(__fadd_rn(acc[j + 1], -mn)));
      p2 = __fadd_rn(p2, exp2w<FAST>(__fadd_rn(acc[j + 2], -mn)));
      p3 = __fadd_rn(p3, exp2w<FAST>(__fadd_rn(acc[j + 3], -mn)));
    }
    s = __fadd_rn(s, __fadd_rn(__fadd_rn(p0, p1), __fadd_rn(p2, p3)));
  }

  template <int BASE>
  __device__ __forceinline__ void view_chunk(const TableParam* __restrict__ tp, float (&acc)[kEpiChunk]) {
    if (any_single) chunk_impl<BASE, true>(tp, acc);
    else chunk_impl<BASE, false>(tp, acc);
  }

  __device__ __forceinline__ void view_end(const ViewParam& vp, float xx) {
    const float Lnew = __fmaf_rn(-vp.AN, xx, vp.CN);
    const float termnew = __fadd_rn(Lnew, (single && lone0) ? vp.WN1 : vp.WN0);
    const float mn = fmaxf(mx, termnew);
    s = __fmul_rn(s, exp2w<FAST>(__fadd_rn(mx, -mn)));
    s = __fadd_rn(s, exp2w<FAST>(__fadd_rn(termnew, -mn)));
    const float logmarg = __fadd_rn(__fadd_rn(mn, log2m(s)), -(single ? vp.LD1 : vp.LD0));
    lnew = __fadd_rn(lnew, logmarg);
  }

  // Whole view from an array of CAP dot products (CUDA-core engine).
  template <int BASE>
  __device__ __forceinline__ void view_chunks_from(const TableParam* __restrict__ tp, float (&acc)[CAP]) {
    if constexpr (BASE < CAP) {
      float ch[kEpiChunk];
#pragma unroll
      for (int j = 0; j < kEpiChunk; ++j) ch[j] = acc[BASE + j];
      view_chunk<BASE>(tp, ch);
      view_chunks_from<BASE + kEpiChunk>(tp, acc);
    }
  }
  __device__ __forceinline__ void view(const TableParam* __restrict__ tp, const ViewParam& vp,
                                       float (&acc)[CAP], float xx) {
    view_begin(tp, xx);
    view_chunks_from<0>(tp, acc);
    view_end(vp, xx);
  }

  // uf in (0,1). Returns the table slot or kNewTable.  lw[] is left holding the weights.
  __device__ __forceinline__ int finish(float uf) {
    float M0 = lnew, M1 = kMasked;
#pragma unroll
    for (int t = 0; t < CAP; t += 2) { M0 = fmaxf(M0, lw[t]); M1 = fmaxf(M1, lw[t + 1]); }
    const float M = fmaxf(M0, M1);
    if (!(M > -1.0e29f)) return t0;   // nothing has weight: stay (cf. multiview_gibbs.cpp:172-176)
    float q0 = 0.0f, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f;
#pragma unroll
    for (int t = 0; t < CAP; t += 4) {
      lw[t] = exp2w<FAST>(__fadd_rn(lw[t], -M));         q0 = __fadd_rn(q0, lw[t]);
      lw[t + 1] = exp2w<FAST>(__fadd_rn(lw[t + 1], -M)); q1 = __fadd_rn(q1, lw[t + 1]);
      lw[t + 2] = exp2w<FAST>(__fadd_rn(lw[t + 2], -M)); q2 = __fadd_rn(q2, lw[t + 2]);
      lw[t + 3] = exp2w<FAST>(__fadd_rn(lw[t + 3], -M)); q3 = __fadd_rn(q3, lw[t + 3]);
    }
    const float total = __fadd_rn(__fadd_rn(__fadd_rn(q0, q1), __fadd_rn(q2, q3)), exp2w<FAST>(__fadd_rn(lnew, -M)));
    const float target = __fmul_rn(uf, total);
    // first t with target < cum_t  ==  number of t with cum_t <= target (cum is non-decreasing)
    float cum = 0.0f;
    int cnt = 0;
#pragma unroll
    for (int t = 0; t < CAP; ++t) {
      cum = __fadd_rn(cum, lw[t]);
      cnt += (target < cum) ? 0 : 1;
    }
    int choice = (cnt < CAP) ? cnt : kNewTable;
    if (cnt >= CAP && !(lnew > -1.0e29f)) {   // rounding fall-through with no new-table mass: last live table
      choice = t0;
#pragma unroll
      for (int t = 0; t < CAP; ++t) if (lw[t] > 1.0e-30f) choice = t;
    }
    return choice;
  }
};

}  // namespace mv
