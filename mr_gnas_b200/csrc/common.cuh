// Shared device/host helpers for libmrgnas (sm_100a).  Warp-per-row kernels: each lane owns
// NV float4 column groups (col4 = lane + 32*v), so a D<=512 fp32 row is moved with 128-bit
// loads/stores and per-column statistics stay in registers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/mrgnas.h"

namespace mrg {

constexpr int kThreads = 256;            // 8 warps per CTA
constexpr int kWarpsPerBlock = kThreads / 32;
constexpr int kNumSMs = 148;             // B200
constexpr int kStatsBlocksPerSM = 6;
constexpr int kMaxParts = kNumSMs * kStatsBlocksPerSM;  // 592 persistent CTAs
constexpr int kFoldRows = 32;            // fold fp32 partial sums into (shared-memory) doubles every 32 rows

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define MRG_CHECK_ARG(cond, msg)                 \
  do {                                           \
    if (!(cond)) {                               \
      mrg::set_error("invalid argument: %s", msg); \
      return MRG_ERR_INVALID;                    \
    }                                            \
  } while (0)

#define MRG_LAUNCH_CHECK(what)                                   \
  do {                                                           \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return mrg::cuda_fail(e__, what);    \
  } while (0)

inline bool valid_D(int D) { return D > 0 && D % 4 == 0 && D <= 512; }
inline int nv_for(int D) { return D <= 128 ? 1 : (D <= 256 ? 2 : 4); }
constexpr int kMinRowsPerWarp = 8;       // small (node-level) launches: fewer CTAs -> fewer partials for the finalize kernels
inline int stats_grid(int64_t rows) {
  int64_t need = (rows + kWarpsPerBlock * kMinRowsPerWarp - 1) / (kWarpsPerBlock * kMinRowsPerWarp);
  if (need < 1) need = 1;
  return (int)(need < kMaxParts ? need : kMaxParts);
}

// Persistent row kernels walk rows with stride gridDim*warps and assume every CTA is resident: a grid larger
// than (resident CTAs per SM) x SMs runs in waves and the last, partial wave costs a full pass (measured:
// 592 CTAs at 3 resident/SM = 444 + 148 -> 1.5x the time).  Cap the grid at what is co-resident.
template <typename K>
inline int resident_grid(K kernel, size_t smem, int want, int threads = kThreads) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess || occ < 1) {
    cudaGetLastError();
    occ = 1;
  }
  const int g = occ * kNumSMs;
  return want < g ? want : g;
}
// per-CTA partial buffers are sized for stats_grid(rows) parts: zero the slots a capped grid does not write
inline void zero_unwritten_parts(double* parts, int grid, int nparts, size_t doubles_per_part, cudaStream_t st) {
  if (parts && grid < nparts)
    cudaMemsetAsync(parts + (size_t)grid * doubles_per_part, 0, (size_t)(nparts - grid) * doubles_per_part * sizeof(double), st);
}

// dispatch on NV (float4 groups per lane)
#define MRG_DISPATCH_NV(D, ...)                      \
  do {                                               \
    int nv__ = mrg::nv_for(D);                       \
    if (nv__ == 1) { constexpr int NV = 1; __VA_ARGS__; } \
    else if (nv__ == 2) { constexpr int NV = 2; __VA_ARGS__; } \
    else { constexpr int NV = 4; __VA_ARGS__; }      \
  } while (0)

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming (read-once) load: bypass L1 allocation
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w));
}

// One row of a streamed [rows, D] matrix held in registers.  Persistent row kernels keep the NEXT row's loads in
// flight while the current row is processed (ncu: with one row per warp in flight the gate kernels sat at
// 2.6 TB/s on long-scoreboard stalls; bytes in flight per SM were ~19 KB against the ~40 KB HBM3e needs).
template <int NV>
struct RowBuf {
  float4 v[NV];
  __device__ __forceinline__ void load(const float* __restrict__ base, int64_t row, int D, int D4, int lane) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c4 = lane + 32 * k;
      if (c4 < D4) v[k] = ld_stream4(base + (size_t)row * D + 4 * c4);
    }
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// correctly rounded reciprocal instead of the IEEE division subroutine (one call per row in the gate kernels)
__device__ __forceinline__ float sigmoidf_(float t) { return __frcp_rn(1.0f + expf(-t)); }

// one element of sigmoid + BCELoss with torch's log clamp at -100 (operations_lp.py:126, mr_lp_train.py:116)
__device__ __forceinline__ float bce_term(float logit, float yv, float* p_out) {
  const float p = sigmoidf_(logit);
  *p_out = p;
  const float lp = fmaxf(logf(p), -100.f);
  const float l1p = fmaxf(logf(1.f - p), -100.f);
  return -(yv * lp + (1.f - yv) * l1p);
}

// Per-lane view of an mrg_act: column affine held in registers.
template <int NV>
struct ActRegs {
  float4 sc[NV], sh[NV];
  bool affine, relu;
  __device__ __forceinline__ void init(const mrg_act& a, int lane, int D4) {
    affine = a.scale != nullptr;
    relu = a.relu != 0;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      int c4 = lane + 32 * v;
      if (affine && c4 < D4) {
        sc[v] = ldg4(a.scale + 4 * c4);
        sh[v] = ldg4(a.shift + 4 * c4);
      } else {
        sc[v] = make_float4(1.f, 1.f, 1.f, 1.f);
        sh[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  __device__ __forceinline__ float4 apply(float4 x, int v) const {
    if (affine) {
      x.x = fmaf(sc[v].x, x.x, sh[v].x);
      x.y = fmaf(sc[v].y, x.y, sh[v].y);
      x.z = fmaf(sc[v].z, x.z, sh[v].z);
      x.w = fmaf(sc[v].w, x.w, sh[v].w);
    }
    if (relu) {
      x.x = x.x > 0.f ? x.x : 0.f;
      x.y = x.y > 0.f ? x.y : 0.f;
      x.z = x.z > 0.f ? x.z : 0.f;
      x.w = x.w > 0.f ? x.w : 0.f;
    }
    return x;
  }
};

// Column (sum, sum of squares) accumulation: fp32 per lane, folded every kFoldRows rows into
// this warp's DOUBLE accumulators in shared memory (lane-owned columns: no conflicts, no atomics,
// and no double registers in the streaming loop), reduced over the CTA's warps in a fixed order
// and written as this CTA's partial.  smem need: stats_smem_doubles(D) doubles.
__host__ __device__ inline int stats_smem_doubles(int D) { return kWarpsPerBlock * 2 * D; }

template <int NV>
struct ColStats {
  float4 fs[NV], fq[NV];
  int pending;
  double* ws;  // this warp's [2][D] doubles
  __device__ __forceinline__ void init(double* smem, int D, int D4) {
    pending = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ws = smem + (size_t)warp * 2 * D;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      fs[v] = fq[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int c4 = lane + 32 * v;
      if (c4 < D4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) ws[4 * c4 + k] = ws[D + 4 * c4 + k] = 0.0;
      }
    }
  }
  __device__ __forceinline__ void add(float4 a, float4 b, int v) {  // sum += a ; sq += b
    fs[v].x += a.x; fs[v].y += a.y; fs[v].z += a.z; fs[v].w += a.w;
    fq[v].x += b.x; fq[v].y += b.y; fq[v].z += b.z; fq[v].w += b.w;
  }
  __device__ __forceinline__ void add_sq(float4 a, int v) {
    add(a, make_float4(a.x * a.x, a.y * a.y, a.z * a.z, a.w * a.w), v);
  }
  __device__ __forceinline__ void fold(int D, int D4) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c4 = lane + 32 * v;
      if (c4 < D4) {
        double* s = ws + 4 * c4;
        double* q = ws + D + 4 * c4;
        s[0] += fs[v].x; s[1] += fs[v].y; s[2] += fs[v].z; s[3] += fs[v].w;
        q[0] += fq[v].x; q[1] += fq[v].y; q[2] += fq[v].z; q[3] += fq[v].w;
      }
      fs[v] = fq[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    pending = 0;
  }
  __device__ __forceinline__ void row_done(int D, int D4) {
    if (++pending == kFoldRows) fold(D, D4);
  }
  // part: this CTA's [2][D] doubles; smem: the base passed to init()
  __device__ __forceinline__ void write_block(double* part, double* smem, int D, int D4) {
    fold(D, D4);
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < kWarpsPerBlock; ++w) t += smem[(size_t)w * 2 * D + c];
      part[c] = t;
    }
  }
};

// Per-column constants (gate vectors, BN affine of an mrg_act) staged once per CTA in shared memory and
// read back with conflict-free 128-bit loads, instead of pinning 8*NV registers each for the whole kernel.
struct ActSmem {
  const float* sc;  // smem, or nullptr if identity
  const float* sh;
  bool relu;
  __device__ __forceinline__ void init(const mrg_act& a, float* smem_sc, float* smem_sh, int D) {
    relu = a.relu != 0;
    if (a.scale) {
      for (int c = threadIdx.x; c < D; c += blockDim.x) {
        smem_sc[c] = a.scale[c];
        smem_sh[c] = a.shift[c];
      }
      sc = smem_sc;
      sh = smem_sh;
    } else {
      sc = sh = nullptr;
    }
  }
  __device__ __forceinline__ float4 apply(float4 x, int c4) const {
    if (sc) {
      const float4 a = *reinterpret_cast<const float4*>(sc + 4 * c4);
      const float4 b = *reinterpret_cast<const float4*>(sh + 4 * c4);
      x.x = fmaf(a.x, x.x, b.x); x.y = fmaf(a.y, x.y, b.y); x.z = fmaf(a.z, x.z, b.z); x.w = fmaf(a.w, x.w, b.w);
    }
    if (relu) {
      x.x = x.x > 0.f ? x.x : 0.f; x.y = x.y > 0.f ? x.y : 0.f;
      x.z = x.z > 0.f ? x.z : 0.f; x.w = x.w > 0.f ? x.w : 0.f;
    }
    return x;
  }
};

}  // namespace mrg
