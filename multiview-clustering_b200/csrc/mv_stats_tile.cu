// mv_stats_tile.cu — sufficient statistics for the C3 shape (cap = 64 table slots, every view dense with
// dim 64, at most 3 views): a bulk-copy (TMA) fed streaming kernel with the sums held in REGISTERS.
//
// One persistent CTA per SM walks a contiguous range of 64-row tiles.
//
//   warp 16 (producer, one lane)   per tile: waits for a free stage of the 4-deep ring and issues cp.async.bulk
//                                  copies: per view the 64 rows x 256 B of features (contiguous in the
//                                  row-major view: no tensor map needed) and their precomputed squared norms,
//                                  plus the sweep's raw draws of the tile's rows.
//   warp 17 (resolver)             turns the raw draws into the table each row now sits at (births: seated
//                                  candidate / overflow stays put, exactly as k_stats does), writes table_cur
//                                  and leaves the resolved tables in the stage.  No global load in the common
//                                  case, so it never stalls the ring.
//   warps 0..15 (accumulators)     warp w owns tables 4w..4w+3 for the whole kernel: lane l keeps, per table
//                                  and view, the sums of columns 2l, 2l+1 in two registers (24 + 4 registers
//                                  per lane), lane v < V the sum of squared norms of view v.  Per tile a warp
//                                  finds the rows of its tables with two ballots per table and adds them in
//                                  ascending row order straight from shared memory: no read-modify-write on
//                                  memory at all, every row is added by exactly one warp.
//
// Summation tree (fixed for a given launch shape, so the float sums are reproducible): rows ascending inside
// a CTA per table, then k_reduce adds the CTAs in ascending order in FP64.  Counts are integers (popc of the
// ballots) and exact.
//
// DELTA mode (incremental statistics, what the reference's remove_customer / add_customer do:
// multiview_utils.cpp:151-163, 199-206): only the 64-row tiles that contain a row whose draw differs from its table
// (movedmask, written by the draw kernel) are fetched; a moved row is added to its new table and subtracted from its
// old one, so the per-CTA partials hold the CHANGE of the statistics and k_reduce_x adds it to the shard's running
// FP64 sums.  With nothing moved the kernel reads the 4 bytes of mask per 32 rows and nothing else; with everything
// moved it costs what the full rebuild costs.  Same row -> CTA -> warp mapping, same fixed order.
// ROW mode of DELTA: when the moved rows are a small part of the tiles they sit in (one moved row in fifty still touches
// 72 % of the 64-row tiles), a CTA compacts the moved ROWS of its range instead and fetches only those — one 256-byte
// bulk copy per (row, view), 64 rows to a stage — so that the statistics kernel reads what actually changed: the sweep
// then streams X once (the draw kernel) plus the moved rows, the reference's remove/add bookkeeping at its own cost.
//
// HBM traffic: the features once (N*V*256 B), 4 B of squared norm per (row, view), 4 B read + 4 B written of
// assignment per row.  Replaces the rebuild loop of /root/reference/Multiview/multiview_gibbs.cpp:64-73 (and
// the incremental updates of multiview_utils.cpp:151-163, 199-206) for this shape.
#include "mv_ctx.h"

namespace mv {

namespace {

constexpr int kTile = 64;                 // rows per stage
constexpr int kStages = 4;
constexpr int kAccWarps = 16;             // x 4 tables = cap 64
constexpr int kTabPerWarp = 4;
constexpr int kThreadsST = (kAccWarps + 2) * 32;
constexpr int kMaxV = 3;
constexpr int kMaxAct = 4096;             // DELTA: entries of the CTA's work list: active tiles, or moved rows in row mode (16 KB)
constexpr int kRowModeTiles = 2048;       // tile mode: tiles per CTA the list may hold

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
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// With an L2 evict-first policy: the feature stream must not push the resident working set out of L2.
__device__ __forceinline__ void bulk_load_stream(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
               : "memory");
}

template <int V>
struct Stage {
  float x[V][kTile][64];      // V x 16 KB, each filled by one bulk copy
  float xx[kMaxV + 1][kTile]; // squared norms of the rows (row V.. unused), one bulk copy per view
  int32_t raw[kTile];         // the sweep's raw draws (bulk copy)
  int32_t tab[kTile];         // resolved table of each row; < 0: nothing to add
  int32_t old[kTile];         // DELTA: the table each row sat at before the sweep (bulk copy)
  int32_t tabold[kTile];      // DELTA: the table a moved row leaves; < 0: nothing to subtract
};

template <int V, bool DELTA>
__global__ void __launch_bounds__(kThreadsST, 1) k_stats_tile(const Ctx c) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Stage<V>* st = reinterpret_cast<Stage<V>*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + sizeof(Stage<V>) * kStages);   // full, ready, empty [kStages] each
  int32_t* act = reinterpret_cast<int32_t*>(bars + 3 * kStages);                          // DELTA: [kMaxAct] active tiles, ascending
  __shared__ int s_nact, s_rowmode, s_nmoved;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  pdl_trigger();
  pdl_wait();

  const int n_tiles = (c.n_rows + kTile - 1) / kTile;
  const int per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int t_lo = min((int)blockIdx.x * per_cta, n_tiles);
  const int t_hi = min(t_lo + per_cta, n_tiles);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&bars[s]), 1);                      // full: the producer's expect_tx arrive + the bytes
      mbar_init(smem_u32(&bars[kStages + s]), 32);           // ready: the 32 resolver lanes
      mbar_init(smem_u32(&bars[2 * kStages + s]), kAccWarps);// empty: one lane of every accumulator warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (DELTA && wid == 0) {
    // this CTA's range: how many tiles hold a moved row, how many rows moved (a 64-row tile = two mask words)
    int n_act = 0, n_mv = 0;
    for (int base = t_lo; base < t_hi; base += 32) {
      const int tile = base + lane;
      unsigned w0 = 0u, w1 = 0u;
      if (tile < t_hi) {
        const int ch = 2 * tile;
        w0 = c.movedmask[ch];
        w1 = (ch + 1 < c.n_chunks) ? c.movedmask[ch + 1] : 0u;
      }
      n_act += __popc(__ballot_sync(0xffffffffu, (w0 | w1) != 0u));
      n_mv += __popc(w0) + __popc(w1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_mv += __shfl_xor_sync(0xffffffffu, n_mv, o);
    // row mode when the moved rows are at most a quarter of the rows of the tiles they sit in
    const bool rowmode = n_mv > 0 && n_mv <= kMaxAct && n_mv * 4 <= n_act * kTile;
    int n = 0;
    for (int base = t_lo; base < t_hi; base += 32) {
      const int tile = base + lane;
      unsigned w0 = 0u, w1 = 0u;
      if (tile < t_hi) {
        const int ch = 2 * tile;
        w0 = c.movedmask[ch];
        w1 = (ch + 1 < c.n_chunks) ? c.movedmask[ch + 1] : 0u;
      }
      if (rowmode) {                               // the moved rows, ascending
        const int mine = __popc(w0) + __popc(w1);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += y;
        }
        int at = n + incl - mine;
        for (unsigned m = w0; m; m &= m - 1u) act[at++] = tile * kTile + __ffs(m) - 1;
        for (unsigned m = w1; m; m &= m - 1u) act[at++] = tile * kTile + 32 + __ffs(m) - 1;
        n += __shfl_sync(0xffffffffu, incl, 31);
      } else {                                     // the tiles with a moved row, ascending
        const bool any = (w0 | w1) != 0u;
        const unsigned m = __ballot_sync(0xffffffffu, any);
        if (any) act[n + __popc(m & ((1u << lane) - 1u))] = tile;
        n += __popc(m);
      }
    }
    if (lane == 0) { s_rowmode = rowmode ? 1 : 0; s_nmoved = rowmode ? n : 0; s_nact = rowmode ? (n + kTile - 1) / kTile : n; }
  }
  __syncthreads();
  const bool rowmode = DELTA && s_rowmode != 0;
  const int n_moved = DELTA ? s_nmoved : 0;
  const int n_iter = DELTA ? s_nact : (t_hi - t_lo);
  if (DELTA) {
    // nothing moved in this CTA's rows: its partials would be all zero; k_reduce_x skips them (adding zeros changes nothing)
    if (tid == 0) c.cta_active[blockIdx.x] = (n_iter > 0) ? 1 : 0;
    if (n_iter == 0) return;
  }

  if (wid == kAccWarps) {
    // ---------------- producer: one lane keeps the ring full ----------------
    if (rowmode) {
      // row mode: the whole warp issues — lane l copies rows l and l + 32 of the stage's (up to) 64 moved rows
      uint64_t policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      for (int k = 0; k < n_iter; ++k) {
        const int s = k % kStages;
        const uint32_t full = smem_u32(&bars[s]);
        const int valid = min(kTile, n_moved - k * kTile);
        if (lane == 0) {
          if (k >= kStages) mbar_wait(smem_u32(&bars[2 * kStages + s]), ((k / kStages) - 1) & 1);
          mbar_expect_tx(full, (uint32_t)(V * valid * 256));
        }
        __syncwarp();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int r = hf * 32 + lane;
          if (r < valid) {
            const int row = act[k * kTile + r];
#pragma unroll
            for (int v = 0; v < V; ++v)
              bulk_load_stream(smem_u32(&st[s].x[v][r][0]), c.x[v] + (size_t)row * 64, 256u, full, policy);
          }
        }
      }
    } else if (lane == 0) {
      uint64_t policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      for (int k = 0; k < n_iter; ++k) {
        const int tile = DELTA ? act[k] : (t_lo + k);
        const int s = k % kStages;
        const uint32_t full = smem_u32(&bars[s]);
        if (k >= kStages) mbar_wait(smem_u32(&bars[2 * kStages + s]), ((k / kStages) - 1) & 1);
        const int row0 = tile * kTile;
        const int valid = min(kTile, c.n_rows - row0);
        const uint32_t small = (uint32_t)((valid + 3) & ~3) * 4u;       // whole 16-byte groups (arrays are padded)
        mbar_expect_tx(full, (uint32_t)(V * valid * 256) + (uint32_t)(V + 1 + (DELTA ? 1 : 0)) * small);
        bulk_load(smem_u32(&st[s].raw[0]), c.choice + row0, small, full);
        if (DELTA) bulk_load(smem_u32(&st[s].old[0]), c.table_cur + row0, small, full);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          bulk_load(smem_u32(&st[s].xx[v][0]), c.xx + (size_t)v * c.xx_stride + row0, small, full);
          bulk_load_stream(smem_u32(&st[s].x[v][0][0]), c.x[v] + (size_t)row0 * 64, (uint32_t)(valid * 256), full, policy);
        }
      }
    }
  } else if (wid == kAccWarps + 1) {
    // ---------------- resolver: raw draws -> the table each row now sits at ----------------
    const int nfree = c.gparam->nfree;
    for (int k = 0; k < n_iter; ++k) {
      const int tile = (DELTA && !rowmode) ? act[k] : (t_lo + k);
      const int s = k % kStages;
      mbar_wait(smem_u32(&bars[s]), (k / kStages) & 1);
      const int row0 = tile * kTile;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int r = hf * 32 + lane;
        int row = row0 + r;
        bool have = row < c.n_rows;
        if (rowmode) {                                     // the stage's rows come from the list; draws, tables, norms from global memory
          have = k * kTile + r < n_moved;
          row = have ? act[k * kTile + r] : 0;
        }
        int t = -3, told = -3;
        if (have) {
          t = rowmode ? c.choice[row] : st[s].raw[r];
          const int cur = DELTA ? (rowmode ? c.table_cur[row] : st[s].old[r]) : 0;
          if (rowmode) {
#pragma unroll
            for (int v = 0; v < V; ++v) st[s].xx[v][r] = c.xx[(size_t)v * c.xx_stride + row];
          }
          if (t == kNewTable) {                            // a birth: seated (candidate, -2) or overflow (stays put)
            const int ch = row >> 5;                       // (tile mode: row0 is a multiple of 64, a chunk is one half of the tile)
            const unsigned m = c.birthmask[ch];
            const int rank = c.chunk_prefix[ch] + __popc(m & ((1u << (row & 31)) - 1u));
            t = (rank < nfree) ? -2 : (DELTA ? cur : c.table_cur[row]);
            c.choice[row] = t;
          }
          if (DELTA) {
            if (t != cur) { told = cur; if (t >= 0) c.table_cur[row] = t; }   // a move: leaves `cur` (a seated birth joins its new table in k_finalize)
            else t = -3;                                                       // stays: nothing to add, nothing to subtract
          } else if (t >= 0) {
            c.table_cur[row] = t;
          }
        }
        st[s].tab[r] = t;
        if (DELTA) st[s].tabold[r] = told;
      }
      mbar_arrive(smem_u32(&bars[kStages + s]));
    }
  } else {
    // ---------------- accumulators ----------------
    float2 acc[kTabPerWarp][V];
    float s2[kTabPerWarp];
    int cnt[kTabPerWarp];
#pragma unroll
    for (int j = 0; j < kTabPerWarp; ++j) {
      s2[j] = 0.0f;
      cnt[j] = 0;
#pragma unroll
      for (int v = 0; v < V; ++v) acc[j][v] = make_float2(0.0f, 0.0f);
    }
    const int tbase = wid * kTabPerWarp;
    for (int k = 0; k < n_iter; ++k) {
      const int s = k % kStages;
      mbar_wait(smem_u32(&bars[s]), (k / kStages) & 1);              // features and squared norms have landed
      mbar_wait(smem_u32(&bars[kStages + s]), (k / kStages) & 1);    // and the rows' tables are resolved
      const Stage<V>& S = st[s];
      const int ta = S.tab[lane], tb = S.tab[32 + lane];
      const int oa = DELTA ? S.tabold[lane] : -3, ob = DELTA ? S.tabold[32 + lane] : -3;
#pragma unroll
      for (int j = 0; j < kTabPerWarp; ++j) {
        unsigned ma = __ballot_sync(0xffffffffu, ta == tbase + j);
        unsigned mb = __ballot_sync(0xffffffffu, tb == tbase + j);
        cnt[j] += __popc(ma) + __popc(mb);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          unsigned m = hf ? mb : ma;
          while (m) {
            const int r = hf * 32 + __ffs(m) - 1;
            m &= m - 1;
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const float2 xv = *reinterpret_cast<const float2*>(&S.x[v][r][2 * lane]);
              acc[j][v].x = __fadd_rn(acc[j][v].x, xv.x);
              acc[j][v].y = __fadd_rn(acc[j][v].y, xv.y);
            }
            s2[j] = __fadd_rn(s2[j], S.xx[lane & 3][r]);      // lane v < V keeps view v's sum; the others are ignored
          }
        }
        if (DELTA) {                                          // the rows that left this table, after the arrivals
          unsigned na = __ballot_sync(0xffffffffu, oa == tbase + j);
          unsigned nb = __ballot_sync(0xffffffffu, ob == tbase + j);
          cnt[j] -= __popc(na) + __popc(nb);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            unsigned m = hf ? nb : na;
            while (m) {
              const int r = hf * 32 + __ffs(m) - 1;
              m &= m - 1;
#pragma unroll
              for (int v = 0; v < V; ++v) {
                const float2 xv = *reinterpret_cast<const float2*>(&S.x[v][r][2 * lane]);
                acc[j][v].x = __fadd_rn(acc[j][v].x, -xv.x);
                acc[j][v].y = __fadd_rn(acc[j][v].y, -xv.y);
              }
              s2[j] = __fadd_rn(s2[j], -S.xx[lane & 3][r]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars[2 * kStages + s]));
    }
    // this CTA's partials, in the layout k_reduce expects
    const size_t part_stride = (size_t)64 * c.Dsum + (size_t)V * 64;
    float* part = c.partial_f + (size_t)blockIdx.x * part_stride;
    float* part_s2 = part + (size_t)64 * c.Dsum;
#pragma unroll
    for (int j = 0; j < kTabPerWarp; ++j) {
      const int t = tbase + j;
#pragma unroll
      for (int v = 0; v < V; ++v)
        *reinterpret_cast<float2*>(part + (size_t)64 * (64 * v) + (size_t)t * 64 + 2 * lane) = acc[j][v];
      if (lane < V) part_s2[lane * 64 + t] = s2[j];
      if (lane == 0) c.partial_n[(size_t)blockIdx.x * 64 + t] = cnt[j];
    }
  }
}

template <int V, bool DELTA>
cudaError_t launch_v(const Ctx& c, cudaStream_t s) {
  const int smem = (int)sizeof(Stage<V>) * kStages + 3 * kStages * 8 + (DELTA ? kMaxAct * 4 : 0);
  static_assert(sizeof(Stage<V>) * kStages + 3 * kStages * 8 + kMaxAct * 4 <= 227 * 1024 - 64, "shared memory budget");
  cudaError_t e = cudaFuncSetAttribute(k_stats_tile<V, DELTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  return launch_chain(k_stats_tile<V, DELTA>, dim3(c.stat_ctas), dim3(kThreadsST), (size_t)smem, s, c.pdl != 0, c);
}

}  // namespace

bool stats_tile_supported(const Ctx& c) {
  if (c.cap != 64 || c.V < 1 || c.V > kMaxV) return false;
  for (int v = 0; v < c.V; ++v)
    if (c.D[v] != 64 || (reinterpret_cast<uintptr_t>(c.x[v]) & 15) != 0) return false;
  return true;
}

bool stats_delta_supported(const Ctx& c) {
  if (!stats_tile_supported(c) || c.stat_ctas <= 0) return false;
  const int n_tiles = (c.n_rows + kTile - 1) / kTile;
  return (n_tiles + c.stat_ctas - 1) / c.stat_ctas <= kRowModeTiles;
}

cudaError_t launch_stats_tile(const Ctx& c, bool delta, cudaStream_t s) {
  if (delta) {
    switch (c.V) {
      case 1: return launch_v<1, true>(c, s);
      case 2: return launch_v<2, true>(c, s);
      default: return launch_v<3, true>(c, s);
    }
  }
  switch (c.V) {
    case 1: return launch_v<1, false>(c, s);
    case 2: return launch_v<2, false>(c, s);
    default: return launch_v<3, false>(c, s);
  }
}

}  // namespace mv
