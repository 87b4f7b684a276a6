// mv_seq.cu — MVG_ENGINE_SEQ: the reference's sequential sampler on the device, one thread per chain
// (mv_seq_core.h has the algorithm and the reference citations).  Compiled WITHOUT FMA contraction (--fmad=false): the
// compiled reference (g++ -O2 on x86-64) does not fuse a*b+c either, and the chain must follow it state for state.
//
// C ABI: mvg_seq_run = run_gibbs_cpp(data_views, M, burn_in, thin) of /root/reference/Multiview/multiview_gibbs.cpp:105-131
// for scalar views: reference start, M sweeps, the state kept after sweep `iter` when iter >= burn_in and
// (iter - burn_in) % thin == 0 (:205).  The data are uploaded once; the chain state lives in device memory; kept states
// are copied out as they occur.
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/mvg.h"
#define MV_SEQ_FN __device__ inline
#include "mv_seq_core.h"

namespace {

__global__ void k_seq_start(mv::seq::State* st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) mv::seq::start(*st);
}
__global__ void k_seq_sweeps(mv::seq::State* st, int n_sweeps) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  mv::seq::State s = *st;                       // scalars in registers; the arrays stay where they are
  for (int it = 0; it < n_sweeps && !s.err; ++it) mv::seq::sweep(s);
  *st = s;
}

struct DevBuf {
  std::vector<void*> ptrs;
  ~DevBuf() { for (void* p : ptrs) cudaFree(p); }
  template <class T> T* get(size_t count) {
    void* q = nullptr;
    if (cudaMalloc(&q, sizeof(T) * (count ? count : 1)) != cudaSuccess) return nullptr;
    cudaMemset(q, 0, sizeof(T) * (count ? count : 1));
    ptrs.push_back(q);
    return static_cast<T*>(q);
  }
};

thread_local std::string g_seq_error;

int seq_fail(int code, const std::string& msg) { g_seq_error = msg; return code; }

}  // namespace

extern "C" const char* mvg_seq_last_error(void) { return g_seq_error.c_str(); }

extern "C" int mvg_seq_run(int32_t device, int32_t n, int32_t d, const double* y, int32_t M, int32_t burn_in, int32_t thin,
                           uint64_t seed, int32_t t_cap, int32_t k_cap, int32_t n_saved_max, int32_t* saved_table_of,
                           int32_t* saved_T, int32_t* saved_dish_of, double* saved_hypers, int32_t* n_saved,
                           uint64_t* stream_calls) {
  using mv::seq::State;
  if (!y || n <= 0 || d <= 0 || d > 64 || M < 0 || thin <= 0 || t_cap < 8 || k_cap < 8)
    return seq_fail(MVG_EINVAL, "mvg_seq_run: bad argument (n, d > 0; thin >= 1; t_cap, k_cap >= 8)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return seq_fail(MVG_ECUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)");
  if (device < 0 || device >= ndev) return seq_fail(MVG_EINVAL, "device ordinal out of range");
  if ((e = cudaSetDevice(device)) != cudaSuccess) return seq_fail(MVG_ECUDA, cudaGetErrorString(e));
  DevBuf B;
  State h{};
  h.n = n; h.d = d; h.T = 0; h.t_cap = t_cap; h.k_cap = k_cap; h.seed = seed; h.calls = 0; h.err = 0;
  double* dy = B.get<double>((size_t)d * n);
  h.table_of = B.get<int>(n); h.n_t = B.get<int>(t_cap); h.dish_of = B.get<int>((size_t)d * t_cap); h.K = B.get<int>(d);
  h.n_vk = B.get<int>((size_t)d * k_cap); h.l_vk = B.get<int>((size_t)d * k_cap);
  h.sum_y = B.get<double>((size_t)d * k_cap); h.sum_y2 = B.get<double>((size_t)d * k_cap);
  h.alpha_v = B.get<double>(d); h.sigma_v = B.get<double>(d); h.tau_v = B.get<double>(d);
  h.prob = B.get<double>(t_cap); h.wts = B.get<double>((size_t)k_cap + 1); h.cand = B.get<int>(k_cap);
  State* ds = B.get<State>(1);
  if (!dy || !h.table_of || !h.n_t || !h.dish_of || !h.K || !h.n_vk || !h.l_vk || !h.sum_y || !h.sum_y2 || !h.alpha_v ||
      !h.sigma_v || !h.tau_v || !h.prob || !h.wts || !h.cand || !ds)
    return seq_fail(MVG_ENOMEM, "mvg_seq_run: cudaMalloc failed");
  h.y = dy;
#define SEQ_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return seq_fail(MVG_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } while (0)
  SEQ_CUDA(cudaMemcpy(dy, y, sizeof(double) * (size_t)d * n, cudaMemcpyHostToDevice));
  SEQ_CUDA(cudaMemcpy(ds, &h, sizeof(State), cudaMemcpyHostToDevice));
  k_seq_start<<<1, 32>>>(ds);
  SEQ_CUDA(cudaGetLastError());
  int saved = 0;
  auto run = [&](int k) -> int {                      // k sweeps in launches of at most 64 (a launch stays well under a second)
    while (k > 0) {
      const int b = k < 64 ? k : 64;
      k_seq_sweeps<<<1, 32>>>(ds, b);
      cudaError_t e2 = cudaGetLastError();
      if (e2 != cudaSuccess) return seq_fail(MVG_ECUDA, std::string("k_seq_sweeps: ") + cudaGetErrorString(e2));
      k -= b;
    }
    return MVG_OK;
  };
  State cur{};
  int done = 0;
  for (int iter = 0; iter < M; ++iter) {
    const bool keep = iter >= burn_in && ((iter - burn_in) % thin == 0);
    if (!keep && iter + 1 < M) continue;              // sweeps are issued up to the next kept one (or the end)
    int rc = run(iter + 1 - done);
    if (rc != MVG_OK) return rc;
    done = iter + 1;
    SEQ_CUDA(cudaMemcpy(&cur, ds, sizeof(State), cudaMemcpyDeviceToHost));
    if (cur.err) return seq_fail(MVG_EINVAL, "mvg_seq_run: capacity exceeded or inconsistent state, flags=" + std::to_string(cur.err) +
                                                 " (1: t_cap tables, 2: k_cap dish slots, 4: customer without a table)");
    if (!keep) break;
    if (saved >= n_saved_max) return seq_fail(MVG_EINVAL, "mvg_seq_run: trace buffers too small");
    if (saved_table_of) SEQ_CUDA(cudaMemcpy(saved_table_of + (size_t)saved * n, h.table_of, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
    if (saved_T) saved_T[saved] = cur.T;
    if (saved_dish_of) SEQ_CUDA(cudaMemcpy(saved_dish_of + (size_t)saved * d * t_cap, h.dish_of, sizeof(int) * (size_t)d * t_cap, cudaMemcpyDeviceToHost));
    if (saved_hypers) {
      double* o = saved_hypers + (size_t)saved * (3 * d + 2);
      SEQ_CUDA(cudaMemcpy(o, h.alpha_v, sizeof(double) * d, cudaMemcpyDeviceToHost));
      SEQ_CUDA(cudaMemcpy(o + d, h.sigma_v, sizeof(double) * d, cudaMemcpyDeviceToHost));
      SEQ_CUDA(cudaMemcpy(o + 2 * d, h.tau_v, sizeof(double) * d, cudaMemcpyDeviceToHost));
      o[3 * d] = cur.alpha_g;
      o[3 * d + 1] = cur.sigma_g;
    }
    ++saved;
  }
  if (done < M) { int rc = run(M - done); if (rc != MVG_OK) return rc; }
  SEQ_CUDA(cudaDeviceSynchronize());
  SEQ_CUDA(cudaMemcpy(&cur, ds, sizeof(State), cudaMemcpyDeviceToHost));
  if (cur.err) return seq_fail(MVG_EINVAL, "mvg_seq_run: capacity exceeded, flags=" + std::to_string(cur.err));
  if (n_saved) *n_saved = saved;
  if (stream_calls) *stream_calls = cur.calls;
#undef SEQ_CUDA
  return MVG_OK;
}
