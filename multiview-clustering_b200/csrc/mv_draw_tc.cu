// mv_draw_tc.cu — likelihood + draw on the 5th-generation tensor cores (engines MVG_ENGINE_TCGEN05*).
//
// Shape: cap = 64 table slots, every view dense with dim 64, one to three views (BASELINE config C3 has three).
// One persistent CTA per SM walks row tiles of 128 customers.  For every (tile, view):
//
//   TMA producer (1 thread)    cp.async.bulk.tensor: the tile's [128 x 64] FP32 features arrive in
//                              shared memory as two K-halves of [128 x 32] in the 128B-swizzled
//                              K-major layout UMMA reads directly (4-deep ring plus an L2 prefetch of the tile-views behind it).
//   converters (2 x 128 thr.)  one warpgroup per K-half; thread r owns customer r: the remainder
//                              x_lo = x - trunc_tf32(x) (exact in FP32; two instructions per element), written
//                              with tcgen05.st into TMEM lane r (the remainder tile never touches shared
//                              memory).  |x|^2 is not recomputed here: it was stored when the view was uploaded.
//   MMA issuer (1 thread)      tcgen05.mma kind::tf32, M=128 N=64 K=8, three passes accumulated in
//                              one TMEM tile:  x.m_hi + x.m_lo (A = the raw tile in shared memory; the
//                              hardware reads the top 19 bits of each FP32 word, so it serves as
//                              x_hi) and x_lo.m_hi (A = the remainder tile in TMEM).  The split
//                              restores ~2^-19 relative accuracy (north_star: FP32 tolerance).
//   epilogue (4 x 128 threads) two PAIRS of warpgroups take alternate row tiles; inside a pair each
//                              warpgroup owns one half of the 64 tables.  tcgen05.ld: thread r of either
//                              warpgroup reads TMEM lane r = customer r, 16 dot products at a time, and
//                              feeds HalfEpilogue (mv_device.cuh): leave-one-out weights and a streaming
//                              log-sum-exp over its half.  The two threads of a customer trade nine
//                              scalars through shared memory (three named-barrier rendezvous per tile)
//                              to merge marginals, totals and the inverse-CDF counts.  Twice the warps
//                              at half the registers each: the epilogue is latency-bound, not issue-bound.
//
// TMEM (512 columns): [0,384) six accumulator tiles, [384,512) two remainder tiles.
// The [N x 64] log-likelihood matrices never exist in memory: HBM traffic is the features once
// (N*V*256 B) plus 8 B per customer (table in, choice out).
//
// Replaces, per customer: remove_customer + compute_table_probs_with_cache + the draw of
// /root/reference/Multiview/multiview_utils.cpp:71-192, :307-350 and multiview_gibbs.cpp:157-199.
#include <cuda.h>

#include "mv_ctx.h"

namespace mv {

namespace {

constexpr int kTileRows = 128;
constexpr int kHalfCols = 32;                       // floats per 128-byte swizzled row
constexpr int kHalfBytes = kTileRows * 128;         // 16 KB: one K-half of an A tile
constexpr int kBHalfBytes = 64 * 128;               // 8 KB: one K-half of a B matrix (64 tables)
constexpr int kRawStages = 4;                       // K-halves of raw features in flight (64 KB) + L2 prefetch two tile-views ahead
constexpr int kLoStages = 2;                        // TMEM remainder tiles (64 columns each)
constexpr int kDStages = 6;                         // TMEM accumulator tiles (64 columns each): three per epilogue pair
constexpr int kTmemCols = 512;
constexpr int kLoCol0 = kDStages * 64;              // first TMEM column of the remainder tiles
constexpr int kMaxTcViews = 3;
constexpr int kThreads = 896;                       // WG0: control, WG1+WG2: converters (one per K-half), WG3..WG6: epilogue
constexpr int kMmaWarps = 2;                        // MMA-issuing warps of WG0; tile-view i belongs to warp 1 + i % 2
constexpr int kExFields = 9;                        // scalars two epilogue threads of one customer trade per tile
constexpr int kEpiGroups = 2;

struct __align__(64) TcMaps {
  CUtensorMap x[kMaxTcViews];
  CUtensorMap mean_hi;
  CUtensorMap mean_lo;
};

// ---- shared memory carve-up (dynamic, 1024-byte aligned for SWIZZLE_128B) --------------------
struct SmemLayout {
  static constexpr int b_off = 0;                                            // [V][hi,lo][2 halves][8 KB]
  static constexpr int raw_off = b_off + kMaxTcViews * 4 * kBHalfBytes;      // 96 KB
  static constexpr int tp_off = raw_off + kRawStages * kHalfBytes;           // +112 KB
  static constexpr int tm_off = tp_off + kMaxTcViews * 64 * (int)sizeof(TableParam);   // per view: PairHot[32] then TableCold[64]
  static constexpr int vp_off = tm_off + 64 * (int)sizeof(TableMass);
  static constexpr int lm_off = vp_off + kMaxTcViews * (int)sizeof(ViewParam);              // float[64]: LM of every table
  static constexpr int same_off = lm_off + 64 * (int)sizeof(float);                         // u64[V][64]: same-dish table masks
  static constexpr int ex_off = same_off + kMaxTcViews * 64 * 8;
  static constexpr int bar_off = ex_off + 4 * kExFields * kTileRows * (int)sizeof(float);   // ex: [pair][half][field][row]
  static constexpr int n_bars = 2 * kRawStages + 2 * kLoStages + 2 * kDStages + 1;
  static constexpr int misc_off = bar_off + n_bars * 8;
  static constexpr int total = misc_off + 64;
};
static_assert(SmemLayout::total <= 227 * 1024, "shared memory budget");

// ---- PTX helpers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"   // %2: suspend-time hint (ns): a waiting warp
      "@p bra DONE;\n\t"                                                 //     sleeps in hardware instead of polling
      "bra WAIT_LOOP;\n\t"                                               //     away the epilogue's issue slots
      "DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity), "r"(2000u)
      : "memory");
}
// mbar_wait that also adds the cycles spent waiting to *acc when profiling is on (debug_export & 2).
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, bool prof, long long& acc) {
  if (prof) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
  } else {
    mbar_wait(bar, parity);
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// Same with an L2 cache policy: the feature stream (768 MB per sweep at C3) is marked evict-first so that it
// does not push the resident working set (parameters, packets, the other kernels' code) out of the 126 MB L2.
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// cute::UMMA::InstrDescriptor for kind::tf32: D = F32 (1 at [4,6)), A = B = TF32 (2 at [7,10) and
// [10,13)), both K-major, N >> 3 at [17,23), M >> 4 at [24,29).
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

// The shared-memory operand descriptor is 64 bits; everything that varies per MMA (the start address,
// bits [0,14), in 16-byte units) lives in the low word, so the issuer only ever adds to `lo`.
template <bool ACC>
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(kIdesc), "n"(ACC ? 1 : 0)
      : "memory");
}
// Same with the A operand in tensor memory: lane = row, one 32-bit column per K element.
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t desc_hi) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 db;\n\t"
      "setp.ne.b32 p, 1, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(desc_hi), "r"(kIdesc)
      : "memory");
}
__device__ __forceinline__ void tmem_st_16(uint32_t taddr, const uint32_t (&u)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]),
        "r"(u[8]), "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, SWIZZLE_128B operand descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits
// [0,14), leading byte offset [16,30) (unused for swizzled K-major: 1), stride byte offset [32,46) =
// 1024 B between 8-row groups, version 1 at [46,48), layout type SWIZZLE_128B = 2 at [61,64).
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);

// One lane of a converged warp (cute::elect_one_sync): lets ptxas keep single-lane tcgen05 issue
// free of the lane-serialising loop it otherwise wraps around uniform-datapath instructions.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n" : "=r"(pred));
  return pred != 0;
}

template <int REGS> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }

struct Ring {   // stage index + mbarrier phase parity of one pipeline role
  int stage = 0;
  uint32_t phase = 0;
  int n;
  __device__ explicit Ring(int n_) : n(n_) {}
  __device__ __forceinline__ void next() { if (++stage == n) { stage = 0; phase ^= 1u; } }
};

}  // namespace

// 16 consecutive accumulator columns of this thread's TMEM lane (issue only; pair with tmem_ld_wait).
__device__ __forceinline__ void tmem_ld_16(uint32_t taddr, uint32_t (&u)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One 16-table chunk of one view (its dot products already sit in `cur`): export them when asked, then
// run the half-epilogue on them.
template <int BASE, bool WITH_NEW, bool FAST>
__device__ __forceinline__ void epi_chunk(HalfEpilogue<32, FAST>& epi, const PairHot* hoth, const TableCold* coldh,
                                          uint32_t (&cur)[16], const Ctx& c, int row, int v, int tbase, bool live) {
  float ch[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) ch[t] = __uint_as_float(cur[t]);
  if ((c.debug_export & 1) && live) {
    float* da = c.dbg_acc + ((size_t)row * c.V + v) * 64 + tbase + BASE;
#pragma unroll
    for (int t = 0; t < 16; ++t) da[t] = ch[t];
  }
  epi.template view_chunk<BASE, WITH_NEW, true>(hoth, coldh, ch);
}

template <bool FAST>
__global__ void __launch_bounds__(kThreads, 1) k_draw_tc(const Ctx c, const __grid_constant__ TcMaps maps) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = c.V;

  unsigned char* s_tp = smem + SmemLayout::tp_off;   // per view 2 KB: PairHot[32] (hot, pair-interleaved) then TableCold[64]
  TableMass* s_tm = reinterpret_cast<TableMass*>(smem + SmemLayout::tm_off);
  ViewParam* s_vp = reinterpret_cast<ViewParam*>(smem + SmemLayout::vp_off);
  float* s_lm = reinterpret_cast<float*>(smem + SmemLayout::lm_off);
  uint32_t* s_misc = reinterpret_cast<uint32_t*>(smem + SmemLayout::misc_off);   // [0] TMEM base, [1] sweep, [2..3] GlobalParam floats

  // barrier addresses
  const uint32_t bar0 = sbase + SmemLayout::bar_off;
  auto raw_full = [&](int s) { return bar0 + 8u * s; };
  auto raw_empty = [&](int s) { return bar0 + 8u * (kRawStages + s); };
  auto lo_full = [&](int s) { return bar0 + 8u * (2 * kRawStages + s); };
  auto lo_empty = [&](int s) { return bar0 + 8u * (2 * kRawStages + kLoStages + s); };
  auto d_full = [&](int s) { return bar0 + 8u * (2 * kRawStages + 2 * kLoStages + s); };
  auto d_empty = [&](int s) { return bar0 + 8u * (2 * kRawStages + 2 * kLoStages + kDStages + s); };
  const uint32_t b_full = bar0 + 8u * (2 * kRawStages + 2 * kLoStages + 2 * kDStages);

  // ---- one-time setup --------------------------------------------------------------------------
  for (int v = 0; v < V; ++v)
    stage_view_params(c.tparam + v * 64, 64, reinterpret_cast<PairHot*>(s_tp + v * 2048),
                      reinterpret_cast<TableCold*>(s_tp + v * 2048 + 1024), tid, kThreads);
  for (int i = tid; i < 64; i += kThreads) { s_tm[i] = c.tmass[i]; s_lm[i] = c.tmass[i].LM; }
  unsigned long long* s_same = reinterpret_cast<unsigned long long*>(smem + SmemLayout::same_off);
  for (int i = tid; i < V * 64; i += kThreads) s_same[i] = c.tsame[i];
  if (tid < V) s_vp[tid] = c.vparam[tid];
  if (tid == 0) {
    const GlobalParam g = *c.gparam;
    s_misc[1] = g.sweep;
    s_misc[2] = __float_as_uint(g.LMN0);
    s_misc[3] = __float_as_uint(g.LMN1);
    for (int s = 0; s < kRawStages; ++s) { mbar_init(raw_full(s), 1); mbar_init(raw_empty(s), 129); }   // one converter warpgroup + the MMA commit
    for (int s = 0; s < kLoStages; ++s) { mbar_init(lo_full(s), 256); mbar_init(lo_empty(s), 1); }      // both converter warpgroups
    for (int s = 0; s < kDStages; ++s) { mbar_init(d_full(s), 1); mbar_init(d_empty(s), 256); }   // both halves of a pair
    mbar_init(b_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {   // TMEM allocation: one warp, address published through shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_misc[0])), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_misc[0];
  const uint32_t sweep = s_misc[1];
  GlobalParam gp;
  gp.LMN0 = __uint_as_float(s_misc[2]);
  gp.LMN1 = __uint_as_float(s_misc[3]);

  const int n_tiles = (c.n_rows + kTileRows - 1) / kTileRows;
  const bool prof = (c.debug_export & 2) != 0 && c.dbg_prof != nullptr;
  long long* prof_out = prof ? c.dbg_prof + (size_t)blockIdx.x * 16 : nullptr;
  long long w0 = 0, w1 = 0, w2 = 0;
  const long long t_start = prof ? clock64() : 0;

  if (warp < 4) {
    // =========================== WG0: control ==================================================
    reg_dec<24>();
    if (warp == 0 && lane == 0) {
      // ---- TMA producer ----
      mbar_expect_tx(b_full, (uint32_t)(V * 4 * kBHalfBytes));
      for (int v = 0; v < V; ++v)
        for (int part = 0; part < 2; ++part)         // 0: m_hi, 1: m_lo
          for (int h = 0; h < 2; ++h)
            tma_load_2d(sbase + SmemLayout::b_off + ((v * 2 + part) * 2 + h) * kBHalfBytes,
                        part ? &maps.mean_lo : &maps.mean_hi, b_full, h * kHalfCols, v * 64);
      Ring r(kRawStages);
      const uint64_t stream_policy = l2_evict_first_policy();
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int v = 0; v < V; ++v)
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            mbar_wait_t(raw_empty(r.stage), r.phase ^ 1u, prof, w0);
            mbar_expect_tx(raw_full(r.stage), kHalfBytes);
            tma_load_2d_hint(sbase + SmemLayout::raw_off + r.stage * kHalfBytes, &maps.x[v], raw_full(r.stage),
                             h * kHalfCols, tile * kTileRows, stream_policy);
            {                                           // pull the same half of the CTA's next tile into L2 meanwhile
              const int tn = tile + gridDim.x;
              if (tn < n_tiles)
                asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.L2::cache_hint [%0, {%1, %2}], %3;"
                             ::"l"(reinterpret_cast<uint64_t>(&maps.x[v])), "r"(h * kHalfCols), "r"(tn * kTileRows),
                               "l"(stream_policy) : "memory");
            }
            r.next();
          }
      if (prof) { prof_out[0] = w0; prof_out[1] = clock64() - t_start; }
    } else if (warp >= 1 && warp <= kMmaWarps) {
      // ---- MMA issuers: warps 1 and 2 take alternate tile-views.  One tcgen05.mma of this shape keeps the
      //      tensor pipe busy for 32 cycles but costs its issuing warp ~100 cycles of uniform-datapath
      //      bookkeeping, so a single issuer cannot feed the pipe.  An mbarrier phase is one parity bit: a
      //      waiter that is a whole use ahead would sail through, so every barrier must be waited on in
      //      use order by ONE thread of control.  The stage counts are chosen for that: tile-view i goes to
      //      warp i % 2, and raw stages (4 = 2 per tile-view), remainder stages (2) and accumulator stages
      //      (6) all give consecutive uses of a stage the same parity of i.  The whole warp walks the loop
      //      (descriptor arithmetic stays warp-uniform); one elected lane issues. ----
      mbar_wait(b_full, 0);
      tc_fence_after();
      const int me = warp - 1;
      const uint32_t a_me = desc_lo(sbase + SmemLayout::raw_off + me * 2 * kHalfBytes);   // this warp's two raw stages
      const uint32_t b_lo0 = desc_lo(sbase + SmemLayout::b_off);
      const uint32_t lo_tmem = tmem_base + (uint32_t)(kLoCol0 + me * 64);                 // this warp's remainder stage
      const int n_tv = ((n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * V;
      int v = me % V, ds = me;                       // view and accumulator stage of tile-view i
      const int vstep = kMmaWarps % V;
      uint32_t use = 0;                              // how many tile-views this warp has issued
      uint32_t dph = 0;                              // accumulator-stage phase: flips every third tile-view of this warp
      int dcnt = 0;
      for (int i = me; i < n_tv; i += kMmaWarps) {
        const uint32_t ph = use & 1u;                // raw and remainder stages: one use per tile-view of this warp
        mbar_wait_t(d_empty(ds), dph ^ 1u, prof, w0);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(ds * 64);
        const uint32_t b_v = b_lo0 + (uint32_t)(v * 4 * (kBHalfBytes >> 4));   // [hi|lo][h] blocks of this view
        // passes 1+2 on the raw halves: x_hi . m_hi  and  x_hi . m_lo   (4 x K=8 per 128-byte row, +32 B each)
        mbar_wait_t(raw_full(2 * me), ph, prof, w1);
        tc_fence_after();
        if (elect_one()) {
          umma_tf32<false>(d_tmem, a_me, b_v, kDescHi);
#pragma unroll
          for (int k = 1; k < 4; ++k) umma_tf32<true>(d_tmem, a_me + 2 * k, b_v + 2 * k, kDescHi);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32<true>(d_tmem, a_me + 2 * k, b_v + 2 * (kBHalfBytes >> 4) + 2 * k, kDescHi);
          umma_commit(raw_empty(2 * me));             // this raw half is free once these MMAs retire
        }
        __syncwarp();
        mbar_wait_t(raw_full(2 * me + 1), ph, prof, w1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a = a_me + (uint32_t)(kHalfBytes >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32<true>(d_tmem, a + 2 * k, b_v + (kBHalfBytes >> 4) + 2 * k, kDescHi);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32<true>(d_tmem, a + 2 * k, b_v + 3 * (kBHalfBytes >> 4) + 2 * k, kDescHi);
          umma_commit(raw_empty(2 * me + 1));
        }
        __syncwarp();
        // pass 3 from tensor memory: x_lo . m_hi
        mbar_wait_t(lo_full(me), ph, prof, w2);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32_ts(d_tmem, lo_tmem + (uint32_t)(k * 8), b_v + 2 * k, kDescHi);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32_ts(d_tmem, lo_tmem + (uint32_t)(32 + k * 8), b_v + (kBHalfBytes >> 4) + 2 * k, kDescHi);
          umma_commit(lo_empty(me));
          umma_commit(d_full(ds));
        }
        __syncwarp();
        ++use;
        v += vstep; if (v >= V) v -= V;              // view of tile-view i + 2 (vstep = 2 mod V, no division on the issue path)
        ds += kMmaWarps; if (ds >= kDStages) ds -= kDStages;
        if (++dcnt == kDStages / kMmaWarps) { dcnt = 0; dph ^= 1u; }
      }
      if (prof && lane == 0 && me == 0) { prof_out[2] = w0; prof_out[3] = w1; prof_out[4] = w2; prof_out[5] = clock64() - t_start; }
    }
  } else if (warp < 12) {
    // =========================== WG1, WG2: converters, one K-half each ===========================
    reg_dec<48>();
    const int h = (warp >> 2) - 1;                     // K-half of this warpgroup
    const int r = tid & 127;                           // row of the tile = 128-byte line of the half = TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t line = (uint32_t)r * 128u;
    int i = 0;                                         // tile-view counter of this CTA
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
      for (int v = 0; v < V; ++v, ++i) {
        const int rs = (2 * i + h) % kRawStages;
        const uint32_t rph = (uint32_t)((2 * i + h) / kRawStages) & 1u;
        const int ls = i % kLoStages;
        const uint32_t lph = (uint32_t)(i / kLoStages) & 1u;
        mbar_wait_t(raw_full(rs), rph, prof, w0);
        mbar_wait_t(lo_empty(ls), lph ^ 1u, prof, w1);
        tc_fence_after();
        const unsigned char* src = smem + SmemLayout::raw_off + rs * kHalfBytes + line;
        const uint32_t lo_addr = tmem_base + lane_base + (uint32_t)(kLoCol0 + ls * 64 + h * 32);
#pragma unroll
        for (int half16 = 0; half16 < 2; ++half16) {
          uint32_t lo[16];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {             // logical 16-byte chunk cidx sits at physical chunk cidx ^ (r & 7)
            const int cidx = half16 * 4 + cc;
            const float4 x4 = *reinterpret_cast<const float4*>(src + ((cidx ^ (r & 7)) << 4));
            const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // the tensor core reads trunc_tf32(x); the remainder x - trunc_tf32(x) is exact in FP32 and the
              // hardware again keeps its top 19 bits: x_hi + x_lo carries >= 21 significant bits of x
              lo[cc * 4 + e] = __float_as_uint(__fadd_rn(xs[e], -__uint_as_float(__float_as_uint(xs[e]) & 0xFFFFE000u)));
            }
          }
          tmem_st_16(lo_addr + (uint32_t)(half16 * 16), lo);
        }
        mbar_arrive(raw_empty(rs));                    // this thread no longer reads the raw half
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(lo_full(ls));
      }
    if (prof && r == 0) { prof_out[6 + 3 * h] = w0; prof_out[7 + 3 * h] = w1; prof_out[8 + 3 * h] = clock64() - t_start; }
  } else {
    // =========================== WG3..WG6: epilogue (thread r <-> TMEM lane r <-> customer r) ====
    // pool: 896 threads x 72 registers = 128 x (24 + 2*48 + 4*96)
    reg_inc<96>();
    const int wg = (warp >> 2) - 3;
    const int pair = wg >> 1;                          // this pair takes tiles j = pair, pair+2, ...
    const int hf = wg & 1;                             // tables [32 hf, 32 hf + 32)
    const int r = tid & 127;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    float* ex_own = reinterpret_cast<float*>(smem + SmemLayout::ex_off) + ((pair * 2 + hf) * kExFields) * kTileRows + r;
    float* ex_oth = reinterpret_cast<float*>(smem + SmemLayout::ex_off) + ((pair * 2 + (hf ^ 1)) * kExFields) * kTileRows + r;
    auto rendezvous = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(1 + pair) : "memory"); };
    // no free table slot => a new table has no weight and the per-view marginals are skipped (CTA-uniform)
    const bool with_new = (gp.LMN0 > -1.0e29f) || (gp.LMN1 > -1.0e29f);
    int j = pair;
    const int tile0 = blockIdx.x + pair * gridDim.x;
    int t0_next = (tile0 < n_tiles) ? c.table_cur[min(tile0 * kTileRows + r, c.n_rows - 1)] : 0;
    // |x|^2 of the customer in every view: computed once when the view was uploaded (c.xx), fetched a tile ahead
    float xx_next[kMaxTcViews];
#pragma unroll
    for (int v = 0; v < kMaxTcViews; ++v)
      xx_next[v] = (tile0 < n_tiles) ? c.xx[(size_t)v * c.xx_stride + min(tile0 * kTileRows + r, c.n_rows - 1)] : 0.0f;   // (xx has >= 3 rows)
    for (int tile = tile0; tile < n_tiles; tile += kEpiGroups * gridDim.x, j += kEpiGroups) {
      const int row = tile * kTileRows + r;
      const bool live = row < c.n_rows;
      const int rowc = live ? row : (c.n_rows - 1);
      HalfEpilogue<32, FAST> epi;
      epi.begin(s_tm, s_lm, t0_next, 32 * hf);
      float xxv[kMaxTcViews];
#pragma unroll
      for (int v = 0; v < kMaxTcViews; ++v) xxv[v] = xx_next[v];
      {                                               // the next tile's tables and norms travel while this one is computed
        const int tn = tile + kEpiGroups * gridDim.x;
        if (tn < n_tiles) {
          const int rn = min(tn * kTileRows + r, c.n_rows - 1);
          t0_next = c.table_cur[rn];
#pragma unroll
          for (int v = 0; v < kMaxTcViews; ++v) xx_next[v] = c.xx[(size_t)v * c.xx_stride + rn];
        }
      }
      float lnew = epi.single ? gp.LMN1 : gp.LMN0;
      for (int v = 0; v < V; ++v) {
        const int idx = j * V + v;
        const int stage = idx % kDStages;             // V = 3: stages 0-2 serve this CTA's even tiles, 3-5 the odd ones
        mbar_wait_t(d_full(stage), (uint32_t)(idx / kDStages) & 1u, prof, w0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_base + (uint32_t)(stage * 64 + 32 * hf);
        const float xx = (v == 0) ? xxv[0] : ((v == 1) ? xxv[1] : xxv[2]);
        const PairHot* hot = reinterpret_cast<const PairHot*>(s_tp + v * 2048);
        const TableCold* cold = reinterpret_cast<const TableCold*>(s_tp + v * 2048 + 1024);
        epi.view_begin(hot, cold, xx);
        epi.samemask = (uint32_t)(s_same[v * 64 + epi.t0] >> (32 * hf));
        if ((c.debug_export & 1) && live && hf == 0) c.dbg_xx[(size_t)row * V + v] = xx;
        uint32_t ua[16], ub[16];
        tmem_ld_16(taddr, ua);
        tmem_ld_wait();
        tmem_ld_16(taddr + 16, ub);
        if (with_new) epi_chunk<0, true>(epi, hot + 16 * hf, cold + 32 * hf, ua, c, row, v, 32 * hf, live);
        else epi_chunk<0, false>(epi, hot + 16 * hf, cold + 32 * hf, ua, c, row, v, 32 * hf, live);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(d_empty(stage));                  // this half's columns are in registers
        if (with_new) {
          epi_chunk<16, true>(epi, hot + 16 * hf, cold + 32 * hf, ub, c, row, v, 32 * hf, live);
          ex_own[(2 * v) * kTileRows] = epi.mx;
          ex_own[(2 * v + 1) * kTileRows] = epi.s;
        } else {
          epi_chunk<16, false>(epi, hot + 16 * hf, cold + 32 * hf, ub, c, row, v, 32 * hf, live);
        }
      }
      float uf = 0.0f;
      if (hf == 0) {
        const U4 rnd = stream_block(c.seed, c.chain, kDomTable, 0, sweep, (uint64_t)(c.row_offset + rowc));
        uf = uniform_f32_from(rnd.x);
        ex_own[8 * kTileRows] = uf;
      }
      ex_own[6 * kTileRows] = epi.halfmax();
      rendezvous();                                   // #1: streaming sums, half maxima, the uniform
      for (int v = 0; with_new && v < V; ++v) {
        const float mo = ex_oth[(2 * v) * kTileRows], so = ex_oth[(2 * v + 1) * kTileRows];
        const float mi = ex_own[(2 * v) * kTileRows], si = ex_own[(2 * v + 1) * kTileRows];
        const TableCold* cold = reinterpret_cast<const TableCold*>(s_tp + v * 2048 + 1024);
        const float xx = (v == 0) ? xxv[0] : ((v == 1) ? xxv[1] : xxv[2]);
        lnew = __fadd_rn(lnew, merge_view<FAST>(hf ? mo : mi, hf ? so : si, hf ? mi : mo, hf ? si : so, s_vp[v], xx,
                                                epi.single, cold[epi.t0].lone));
      }
      if (hf == 1) uf = ex_oth[8 * kTileRows];
      const float M = hf ? fmaxf(fmaxf(lnew, ex_oth[6 * kTileRows]), ex_own[6 * kTileRows])
                         : fmaxf(fmaxf(lnew, ex_own[6 * kTileRows]), ex_oth[6 * kTileRows]);
      int choice = epi.t0;                            // nothing has weight: stay (cf. multiview_gibbs.cpp:172-176)
      const bool any_weight = (M > -1.0e29f);         // identical in both threads of the customer
      float Hown = 0.0f;
      if (any_weight) Hown = epi.weights(M);
      ex_own[7 * kTileRows] = Hown;
      rendezvous();                                   // #2: half totals
      const float Hoth = ex_oth[7 * kTileRows];
      const float HA = hf ? Hoth : Hown, HB = hf ? Hown : Hoth;
      const float total = __fadd_rn(__fadd_rn(HA, HB), exp2w<FAST>(__fadd_rn(lnew, -M)));
      const float target = __fmul_rn(uf, total);
      int cnt = 0;
      if (any_weight) cnt = epi.scan(target, hf ? HA : 0.0f);
      if (hf == 1) {
        int enc = cnt;
        if (cnt == 32) enc |= (epi.last_live() + 1) << 8;   // only needed when the draw ran past every table
        ex_own[8 * kTileRows] = __int_as_float(enc);
      }
      rendezvous();                                   // #3: the upper half's count
      if (hf == 0) {
        if (any_weight) {
          const int enc = __float_as_int(ex_oth[8 * kTileRows]);
          const int total_cnt = cnt + (enc & 0xFF);
          choice = (total_cnt < 64) ? total_cnt : kNewTable;
          if (total_cnt >= 64 && !(lnew > -1.0e29f)) {   // rounding fall-through with no new-table mass: last live table
            const int lb = (enc >> 8) - 1, la = epi.last_live();
            choice = (lb >= 0) ? (32 + lb) : ((la >= 0) ? la : epi.t0);
          }
        }
        if (live) {
          c.choice[row] = choice;
          if (c.debug_export & 1) c.dbg_choice[row] = choice;
        }
        const unsigned births = __ballot_sync(0xffffffffu, live && choice == kNewTable);
        if (lane == 0 && (row >> 5) < c.n_chunks) c.birthmask[row >> 5] = births;
      }
    }
    if (prof && r == 0 && hf == 0) { prof_out[12 + 2 * pair] = w0 + w1; prof_out[13 + 2 * pair] = clock64() - t_start; }
  }

  // ---- teardown ------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

// =================================================================================================
// host side
// =================================================================================================
bool draw_tc_supported(const Ctx& c) {
  if (c.cap != 64 || c.V < 1 || c.V > kMaxTcViews) return false;   // shared memory and TMEM are laid out for up to three views
  for (int v = 0; v < c.V; ++v)
    if (c.D[v] != 64 || (reinterpret_cast<uintptr_t>(c.x[v]) & 15) != 0) return false;
  return true;
}

size_t draw_tc_maps_bytes() { return sizeof(TcMaps); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static cudaError_t encode_2d(EncodeTiledFn fn, CUtensorMap* map, const void* base, uint64_t rows, uint32_t box_rows) {
  cuuint64_t dims[2] = {64, rows};
  cuuint64_t strides[1] = {64 * sizeof(float)};
  cuuint32_t box[2] = {kHalfCols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t draw_tc_make_maps(const Ctx& c, void* maps_out) {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  if (e != cudaSuccess) return e;
  if (!fnp || q != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(fnp);
  TcMaps* m = static_cast<TcMaps*>(maps_out);
  for (int v = 0; v < c.V; ++v)
    if ((e = encode_2d(fn, &m->x[v], c.x[v], (uint64_t)c.n_rows, kTileRows)) != cudaSuccess) return e;
  for (int v = c.V; v < kMaxTcViews; ++v) m->x[v] = m->x[0];
  if ((e = encode_2d(fn, &m->mean_hi, c.mean_hi, (uint64_t)c.V * 64, 64)) != cudaSuccess) return e;
  if ((e = encode_2d(fn, &m->mean_lo, c.mean_lo, (uint64_t)c.V * 64, 64)) != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t launch_draw_tc(const Ctx& c, const void* maps, bool fast, cudaStream_t s) {
  if (c.n_rows <= 0) return cudaSuccess;
  if (!draw_tc_supported(c) || !maps) return cudaErrorInvalidValue;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n_tiles = (c.n_rows + kTileRows - 1) / kTileRows;
  const int grid = n_tiles < sms ? n_tiles : sms;
  auto kern = fast ? k_draw_tc<true> : k_draw_tc<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout::total);
  if (e != cudaSuccess) return e;
  kern<<<grid, kThreads, SmemLayout::total, s>>>(c, *static_cast<const TcMaps*>(maps));
  return cudaGetLastError();
}

}  // namespace mv
