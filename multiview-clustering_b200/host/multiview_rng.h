// multiview_rng.h — host random stream with the surface of the reference's (unused) header
// /root/reference/Multiview/multiview_rng.h:9-24: global_rng(), uniform01(), rnorm(mean, sd) and the
// set_rng_seed() it left commented out (:26-30).  The generator is the same counter-based Philox4x32-10
// the GPU kernels use (exported by the C ABI as mvg_philox_*), taken in call order on its own domain, so
// a host test can draw exactly the numbers a device thread would draw for a given (sweep, row, slot).
#pragma once

#include <cstdint>
#include <limits>
#include <random>

#include "../../include/mvg.h"

struct mv_philox_engine {            // satisfies UniformRandomBitGenerator
  using result_type = uint32_t;
  uint64_t seed = 1999;
  uint64_t calls = 0;
  static constexpr result_type min() { return 0; }
  static constexpr result_type max() { return std::numeric_limits<uint32_t>::max(); }
  result_type operator()() {
    uint32_t ctr[4] = {(uint32_t)calls, (uint32_t)(calls >> 32), 0u, 6u << 24}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, out[4];
    ++calls;
    mvg_philox4x32_10(ctr, key, out);
    return out[0];
  }
};

inline mv_philox_engine& philox_rng() {
  static thread_local mv_philox_engine rng;
  return rng;
}
// The reference's signature (multiview_rng.h:9): a std::mt19937 for callers that want a standard engine.  It is seeded
// from the Philox stream (set_rng_seed re-seeds both); uniform01() / rnorm() below do not go through it.
inline std::mt19937& global_rng() {
  static thread_local std::mt19937 rng(1999u);
  return rng;
}
inline void set_rng_seed(unsigned int seed) {
  philox_rng().seed = seed;
  philox_rng().calls = 0;
  global_rng().seed(seed);
}
// Uniform(0, 1), 53 bits, never 0 or 1 (domain 6 = call-ordered host stream)
inline double uniform01() {
  mv_philox_engine& g = philox_rng();
  return mvg_philox_uniform_f64(g.seed, 0u, 6u, 0u, 0u, g.calls++);
}
// Normal(mean, sd^2) by Box-Muller on one Philox block
inline double rnorm(double mean, double sd) {
  mv_philox_engine& g = philox_rng();
  return mean + sd * mvg_philox_normal(g.seed, 0u, 6u, 1u, 0u, g.calls++);
}
