/* oracle/mv_philox_ref.h — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Independent plain-C restatement of the counter-based random stream the B200 sampler
 * uses (the product's copy lives in multiview-clustering_b200/csrc/mv_philox.h; the two
 * are written separately and tests/test_philox.py checks them against each other and
 * against the Random123 known-answer vectors for Philox4x32-10).
 *
 * The reference has no pinned RNG: its live stream is R's (R::runif / R::rnorm,
 * /root/reference/Multiview/multiview_utils.cpp:305-306, :261; multiview_gibbs.cpp:26,56)
 * and multiview_rng.h:9-24 is an unused std::mt19937 alternative.  SURVEY.md Appendix C
 * therefore defines the stream; this file is that definition:
 *
 *   key     = ( lo32(seed),  hi32(seed) + chain )                       (mod 2^32)
 *   counter = ( lo32(index), hi32(index), sweep, (domain << 24) | slot )
 *
 *   domain 0  table draw of row `index` in sweep `sweep`   (multiview_gibbs.cpp:181)
 *   domain 1  dish draw, view `slot`, for a table born at row `index`   (multiview_utils.cpp:261)
 *   domain 2  hyper-step normal number `index`              (multiview_hyper.cpp:104,126,170)
 *   domain 3  hyper-step uniform number `index`             (multiview_hyper.cpp:228,253,260,279,286)
 *   domain 4  init: table of row `index`                    (multiview_gibbs.cpp:26)
 *   domain 5  init: dish of table `index` in view `slot`    (multiview_gibbs.cpp:56)
 *   domain 6  call-ordered stream (index = call number) backing R::runif/R::rnorm when
 *             the compiled reference is driven from ref_shim.cpp
 *
 *   uf  = ((x0 >> 9) + 0.5) * 2^-23          float  in (0,1); k + 0.5 is exact in FP32 for k < 2^23
 *   u53 = ((x0 >> 5) * 2^26 + (x1 >> 6) + 0.5) * 2^-53     double in (0,1) (rounded once)
 *   z   = sqrt(-2 ln u53(x0,x1)) * cos(2 pi u53(x2,x3))    standard normal (Box-Muller)
 */
#ifndef MV_PHILOX_REF_H
#define MV_PHILOX_REF_H

#include <math.h>
#include <stdint.h>

enum {
  MVO_DOM_TABLE = 0,
  MVO_DOM_DISH = 1,
  MVO_DOM_HYPER_NORMAL = 2,
  MVO_DOM_HYPER_UNIF = 3,
  MVO_DOM_INIT_TABLE = 4,
  MVO_DOM_INIT_DISH = 5,
  MVO_DOM_CALLSEQ = 6
};

static inline void mvo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * (uint64_t)c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * (uint64_t)c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline void mvo_stream_block(uint64_t seed, uint32_t chain, uint32_t domain, uint32_t slot,
                                    uint32_t sweep, uint64_t index, uint32_t out[4]) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32) + chain};
  uint32_t ctr[4] = {(uint32_t)index, (uint32_t)(index >> 32), sweep, (domain << 24) | (slot & 0xFFFFFFu)};
  mvo_philox4x32_10(ctr, key, out);
}

static inline float mvo_uf_from(uint32_t x0) {
  return ((float)(x0 >> 9) + 0.5f) * 1.1920928955078125e-07f; /* 2^-23; every step is exact */
}

static inline double mvo_u53_from(uint32_t a, uint32_t b) {
  uint64_t k = ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6);
  return ((double)k + 0.5) * 1.1102230246251565e-16; /* 2^-53 */
}

static inline float mvo_uniform_f32(uint64_t seed, uint32_t chain, uint32_t domain, uint32_t slot,
                                  uint32_t sweep, uint64_t index) {
  uint32_t x[4];
  mvo_stream_block(seed, chain, domain, slot, sweep, index, x);
  return mvo_uf_from(x[0]);
}

static inline double mvo_uniform53(uint64_t seed, uint32_t chain, uint32_t domain, uint32_t slot,
                                   uint32_t sweep, uint64_t index) {
  uint32_t x[4];
  mvo_stream_block(seed, chain, domain, slot, sweep, index, x);
  return mvo_u53_from(x[0], x[1]);
}

static inline double mvo_normal(uint64_t seed, uint32_t chain, uint32_t domain, uint32_t slot,
                                uint32_t sweep, uint64_t index) {
  uint32_t x[4];
  mvo_stream_block(seed, chain, domain, slot, sweep, index, x);
  double u1 = mvo_u53_from(x[0], x[1]);
  double u2 = mvo_u53_from(x[2], x[3]);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

#endif /* MV_PHILOX_REF_H */
