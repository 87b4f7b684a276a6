// mv_counts.cu — sufficient statistics and per-sweep tables of the sparse COUNT views (CSR).
//
// The reference has no count likelihood: this is new capability, parity pinned only to the
// FP64 restatement in oracle/mv_oracle.c (counts_log_f_vk).  Model: the rows of a dish are draws of one
// multinomial whose probabilities are the plug-in estimate theta_kw = (beta + c_kw) / (W beta + C_k) from the
// dish's word counts; log f_vk(x) = sum_w x_w log theta_kw, the row's own counts removed for its own dish.
//
//   k_rowtotals       |x| of every row, once per upload (kept in the squared-norm slot of the view)
//   k_counts_scatter  word counts per TABLE slot: cnt_t[w][t] += x_w over the rows seated at t.  Integer
//                     atomics: the sums are exact and independent of the order (north_star: integer counts
//                     bit-exact)
//   k_counts_tables   per (word, table slot): the counts of the DISH the slot serves (sum over the slots of that
//                     dish, ascending) and log2 theta, the table the likelihood kernel reads feature-major
#include "mv_ctx.h"

namespace mv {

__global__ void k_rowtotals(const int32_t* __restrict__ rowptr, const float* __restrict__ val, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float tot = 0.0f;
  for (int j = rowptr[i]; j < rowptr[i + 1]; ++j) tot = __fadd_rn(tot, val[j]);   // ascending: the mirror's order
  out[i] = tot;
}

__global__ void k_counts_scatter(const Ctx c, const int v) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int32_t* __restrict__ rp = c.rowptr[v];
  const int32_t* __restrict__ col = c.col[v];
  const float* __restrict__ val = c.val[v];
  int32_t* cnt = c.cnt_t[v];
  for (int i = warp; i < c.n_rows; i += nwarps) {
    const int t = c.table_cur[i];
    for (int j = rp[i] + lane; j < rp[i + 1]; j += 32) atomicAdd(&cnt[(size_t)col[j] * c.cap + t], (int32_t)val[j]);
  }
}

__global__ void k_counts_tables(const Ctx c, const int v) {
  __shared__ int32_t s_dish[64];
  __shared__ double s_den[64];
  const int cap = c.cap;
  if (threadIdx.x < cap) {
    const int k = c.dish_of[v * cap + threadIdx.x];
    s_dish[threadIdx.x] = k;
    // S2k of a count view holds the token total of the dish (an exact integer)
    s_den[threadIdx.x] = (k >= 0) ? (double)c.vocab[v] * (double)c.count_beta + c.S2k[v * cap + k] : 1.0;
  }
  __syncthreads();
  const size_t total = (size_t)c.vocab[v] * cap;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(e % cap);
    const size_t w = e / cap;
    const int k = s_dish[t];
    int32_t cd = 0;
    float l2 = 0.0f;
    if (k >= 0) {
      const int32_t* row = c.cnt_t[v] + w * cap;
      for (int t2 = 0; t2 < cap; ++t2) if (s_dish[t2] == k) cd += row[t2];
      l2 = (float)log2(((double)c.count_beta + (double)cd) / s_den[t]);
    }
    c.cnt_d[v][e] = cd;
    c.l2t[v][e] = l2;
  }
}

cudaError_t launch_rowtotals(const int32_t* rowptr, const float* val, float* out, int n, cudaStream_t s) {
  k_rowtotals<<<(n + 255) / 256, 256, 0, s>>>(rowptr, val, out, n);
  return cudaGetLastError();
}

cudaError_t launch_counts_rebuild(const Ctx& c, cudaStream_t s) {
  for (int v = 0; v < c.V; ++v) {
    if (!c.kind[v]) continue;
    const size_t cells = (size_t)c.vocab[v] * c.cap;
    cudaError_t e = cudaMemsetAsync(c.cnt_t[v], 0, sizeof(int32_t) * cells, s);
    if (e != cudaSuccess) return e;
    int blocks = (c.n_rows + 7) / 8;                 // 8 warps per block, one row per warp per step
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_counts_scatter<<<blocks, 256, 0, s>>>(c, v);
    int tb = (int)((cells + 255) / 256);
    if (tb > 148 * 16) tb = 148 * 16;
    k_counts_tables<<<tb, 256, 0, s>>>(c, v);
  }
  return cudaGetLastError();
}

}  // namespace mv
