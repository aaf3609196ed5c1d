// TMA bulk-copy row pipeline shared by the streaming kernels.
//
// Why: a warp-per-row kernel that loads straight into registers has (warps resident) x (one row) bytes in
// flight; the gate kernels need 70-140 registers, so only 8-24 warps fit per SM and they sat at 2.6-3.2 TB/s
// on long-scoreboard stalls (profiles/r01_*).  Here one elected producer thread streams R-row tiles of every
// input matrix into a shared-memory ring with cp.async.bulk (completion by mbarrier expect_tx); the bytes in
// flight are (stages x tile bytes) per SM, independent of the consumers' register use
// (scripts/microbench_rows.cu: 6.5 TB/s for 2 reads + 1 write per row with R=8, 8 stages).
#pragma once
#include "common.cuh"

namespace mrg {
namespace pipe {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t it = 0; it < (1u << 26); ++it)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

constexpr int kMaxStreams = 6;
constexpr int kConsumerWarps = 16;            // the consumers are issue/latency bound: 8 warps left the SM 58 % idle
constexpr int kTileRows = kConsumerWarps;     // rows per tile = one row per consumer warp
constexpr int kPipeThreads = (kConsumerWarps + 1) * 32;

// Streams of [rows, D] fp32 matrices, all tiled identically.
struct Streams {
  const float* base[kMaxStreams];
  int n;
};

// Shared-memory ring: stage s holds n_streams tiles of kTileRows x D floats.
struct Ring {
  float* tiles;        // [stages][n_streams][kTileRows * D]
  uint64_t* full;      // [stages]
  uint64_t* empty;     // [stages]
  int stages, n_streams, D;
  __device__ __forceinline__ const float* tile(int stage, int stream) const {
    return tiles + ((size_t)stage * n_streams + stream) * kTileRows * D;
  }
  __device__ __forceinline__ void init() {
    if (threadIdx.x == 0) {
      for (int s = 0; s < stages; ++s) {
        mbar_init(full + s, 1);
        mbar_init(empty + s, kConsumerWarps);
      }
      fence_barrier_init();
    }
    __syncthreads();
  }
  // producer: one thread walks this CTA's tiles (tile = blockIdx.x + it * gridDim.x)
  __device__ __forceinline__ void produce(const Streams& st, int64_t rows) {
    const int64_t ntiles = (rows + kTileRows - 1) / kTileRows;
    int it = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int s = it % stages;
      if (it >= stages) mbar_wait(empty + s, ((it / stages) - 1) & 1);
      const int64_t r0 = t * kTileRows;
      const int64_t nr = rows - r0 < kTileRows ? rows - r0 : kTileRows;
      const uint32_t bytes = (uint32_t)(nr * D * 4);
      mbar_expect_tx(full + s, bytes * st.n);
      for (int k = 0; k < st.n; ++k)
        bulk_g2s(const_cast<float*>(tile(s, k)), st.base[k] + (size_t)r0 * D, bytes, full + s);
    }
  }
};

inline size_t ring_bytes(int stages, int n_streams, int D) {
  return (size_t)stages * n_streams * kTileRows * D * 4 + 2 * (size_t)stages * 8;
}

}  // namespace pipe
}  // namespace mrg
