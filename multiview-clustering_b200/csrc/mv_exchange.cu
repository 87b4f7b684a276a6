// mv_exchange.cu — the once-per-sweep exchange of the shards' packets over NVLink peer memory.
//
// Every rank owns a receive buffer  xrecv = [2 parities][world][pkt.bytes] + counters [2][world]  that its peers have
// mapped (CUDA IPC).  After k_reduce has finished this rank's packet, ONE kernel (k_exchange) does both directions:
//
//   push blocks (g, 0..7)   store the packet straight into rank g's receive slot (P2P stores through NVLink /
//                           NVSwitch) and count their arrival in g's buffer with a system-scope release;
//   wait blocks (g, 8..11)  wait (acquire) until all of rank g's push blocks have arrived in the LOCAL buffer and
//                           copy that slot into the working packet array k_finalize reads.
//
// Two parities: a rank can run at most one exchange ahead of a peer that is still reading the previous packets (it
// cannot finish exchange s+1 without that peer's packet s+1).  The sequence number is kept by the host (every rank
// issues the same number of exchanges).  The waits are bounded (2^24 polls, a few seconds): a peer that never arrives
// raises status bit 8 instead of hanging the GPU.  ncclAllGather remains the transport when no peer buffers are attached.
#include "mv_ctx.h"

namespace mv {

constexpr int kPushBlocks = 8;     // blocks per destination rank: the stores of one slot are spread over 8 SMs
constexpr int kWaitBlocks = 4;     // blocks per source rank copying the received slot into place

// Arrivals are COUNTED: every push block adds 1 to counter[parity][source] in the destination's buffer after its part
// of the slot is globally visible; the k-th exchange on a parity is complete at kPushBlocks * k (counters only grow).
__device__ __forceinline__ uint32_t arrivals_expected(uint32_t seq) { return (uint32_t)kPushBlocks * ((seq + 1u) >> 1); }

// ONE launch: blocks (g, 0 .. kPushBlocks-1) push to rank g, blocks (g, kPushBlocks ..) wait for rank g and copy its
// slot into place.  All blocks are resident at once (<= 16 x 12), so pushes and waits overlap.
__global__ void __launch_bounds__(256) k_exchange(const Ctx c, const XchgPeers peers, unsigned char* __restrict__ recv_local,
                                                  const uint32_t seq) {
  const int g = blockIdx.x;
  const int parity = seq & 1u;
  const size_t bytes = (size_t)c.pkt.bytes;                    // multiple of 16
  const size_t n16 = bytes / 16;
  if (blockIdx.y < kPushBlocks) {
    // ---- push this rank's packet into rank g's receive slot ----
    if (g == c.rank) return;
    const size_t per = (n16 + kPushBlocks - 1) / kPushBlocks;
    const size_t lo = (size_t)blockIdx.y * per, hi = (lo + per < n16) ? lo + per : n16;
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(c.packet + (size_t)c.rank * bytes);
    unsigned char* base = peers.recv[g];
    uint4* dst = reinterpret_cast<uint4*>(base + ((size_t)parity * c.world + c.rank) * bytes);
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = src[i];
    __syncthreads();                                           // the block's stores happen-before thread 0's release
    if (threadIdx.x == 0) {
      uint32_t* counter = reinterpret_cast<uint32_t*>(base + (size_t)2 * c.world * bytes) + parity * c.world + c.rank;
      asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    }
  } else {
    // ---- wait for rank g's packet and copy it into the working array ----
    if (g == c.rank) return;                                   // own slot is already in place
    __shared__ int ok;
    if (threadIdx.x == 0) {
      const uint32_t* counter = reinterpret_cast<const uint32_t*>(recv_local + (size_t)2 * c.world * bytes) + parity * c.world + g;
      const uint32_t want = arrivals_expected(seq);
      uint32_t v = 0;
      long long spins = 0;
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      } while ((int32_t)(v - want) < 0 && ++spins < (1ll << 24));   // seconds at most: a missing peer must not hang the GPU
      ok = ((int32_t)(v - want) >= 0);
      if (!ok) atomicOr(c.status, 8);
    }
    __syncthreads();
    if (!ok) return;
    const int wb = blockIdx.y - kPushBlocks;
    const size_t per = (n16 + kWaitBlocks - 1) / kWaitBlocks;
    const size_t lo = (size_t)wb * per, hi = (lo + per < n16) ? lo + per : n16;
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(recv_local + ((size_t)parity * c.world + g) * bytes);
    uint4* dst = reinterpret_cast<uint4*>(c.packet + (size_t)g * bytes);
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = src[i];
  }
}

cudaError_t launch_exchange_p2p(const Ctx& c, const XchgPeers& peers, unsigned char* recv_local, uint32_t seq, cudaStream_t s) {
  k_exchange<<<dim3(c.world, kPushBlocks + kWaitBlocks), 256, 0, s>>>(c, peers, recv_local, seq);
  return cudaGetLastError();
}

}  // namespace mv
