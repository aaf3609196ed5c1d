// Backward of a_max (operations_lp.py:230-235) exploiting the sparsity of max-routing.
//
// The reference back-propagates through update_all(max) + ReLU + Linear with two DENSE
// [E,D]x[D,D] GEMMs (dX = dM W, dW = dM^T X) although dM has exactly one non-zero per
// (destination, feature): N*D of E*D entries (1/avg-degree dense).  Here both products are
// evaluated on the non-zeros only -- 2*N*D*D FMAs instead of 4*E*D*D -- and dM is never formed:
//
//   dX[e,:]  = sum_{f : arg[n,f]==e} g[n,f] * W[f,:]          n = dst[e]       (amax_bwd_dx)
//   dW[f,:]  = sum_{n : arg[n,f]>=0} g[n,f] * x(arg[n,f],:)                     (amax_bwd_dw)
//   db[f]    = sum_{n : arg[n,f]>=0} g[n,f]
//
// arg uses the encoded form of the forward kernels: >=0 edge id with a positive message,
// <0 no gradient (isolated destination, or ReLU gated the maximum off).
// Both kernels are deterministic: dX rows have one writer; dW is accumulated per CTA in shared
// memory by a fixed warp<->row ownership and the per-CTA partials are folded in CTA order.
#include "common.cuh"

namespace mrg {

constexpr int kBwdThreads = 1024;  // one CTA per SM (W / dW slice resident in smem) -> fill it with 32 warps
constexpr int kBwdWarps = kBwdThreads / 32;

// ---------------------------------------------------------------------------------------
// dX: one warp per <=32-edge chunk of a destination's CSR row.  Lane l keeps g[n,f], arg[n,f]
// for f = l + 32 j in registers; for every edge a ballot per 32-feature strip finds the features
// routed to it and their W rows are accumulated from shared memory (W resident when it fits).
// ---------------------------------------------------------------------------------------
template <int NJ, bool W_SMEM>
__global__ void __launch_bounds__(kBwdThreads, 1) amax_bwd_dx_kernel(
    const float* __restrict__ g, const int32_t* __restrict__ arg, const float* __restrict__ W,
    const int32_t* __restrict__ ptr, const int32_t* __restrict__ eid, const int32_t* __restrict__ chunk_first,
    const int32_t* __restrict__ chunk_seg, int64_t nseg, int D, float* __restrict__ dX) {
  extern __shared__ float smem_w[];  // [D][D] when W_SMEM
  const int lane = threadIdx.x & 31;
  if (W_SMEM) {
    for (int i = threadIdx.x * 4; i < D * D; i += blockDim.x * 4)
      *reinterpret_cast<float4*>(smem_w + i) = ldg4(W + i);
    __syncthreads();
  }
  const float* Wp = W_SMEM ? smem_w : W;
  const int64_t warp0 = (int64_t)blockIdx.x * kBwdWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kBwdWarps;
  const int64_t nchunks = chunk_first[nseg];
  for (int64_t ch = warp0; ch < nchunks; ch += nwarps) {
    const int64_t n = chunk_seg[ch];
    const int32_t p0 = ptr[n], p1 = ptr[n + 1];
    const int32_t lo = p0 + (int32_t)(ch - chunk_first[n]) * MRG_CHUNK_ROWS;
    const int32_t hi = min(lo + MRG_CHUNK_ROWS, p1);
    float gr[NJ];
    int32_t ar[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int f = lane + 32 * j;
      gr[j] = f < D ? __ldg(g + (size_t)n * D + f) : 0.f;
      ar[j] = f < D ? __ldg(arg + (size_t)n * D + f) : -1;
    }
    for (int32_t pz = lo; pz < hi; ++pz) {
      const int32_t e = __ldg(eid + pz);
      float acc[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[j] = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        unsigned m = __ballot_sync(0xffffffffu, ar[j] == e);
        while (m) {
          const int src_lane = __ffs(m) - 1;
          m &= m - 1;
          const float gv = __shfl_sync(0xffffffffu, gr[j], src_lane);
          const float* wrow = Wp + (size_t)(src_lane + 32 * j) * D;
#pragma unroll
          for (int t = 0; t < NJ; ++t) {
            const int k = lane + 32 * t;
            if (k < D) acc[t] = fmaf(gv, W_SMEM ? wrow[k] : __ldg(wrow + k), acc[t]);
          }
        }
      }
      float* out = dX + (size_t)e * D;
#pragma unroll
      for (int t = 0; t < NJ; ++t) {
        const int k = lane + 32 * t;
        if (k < D) out[k] = acc[t];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// dW / db: every CTA accumulates a private [D][KW] slice in shared memory (warp w owns rows
// f = w, w+8, ...), walking destinations n = cta, cta+stride, ...; x rows are gathered through
// the lazy BatchNorm+ReLU view.  Partials: part[cta][D][KW] (+ db in part_b[cta][D]).
// ---------------------------------------------------------------------------------------
template <int NT>   // NT = float4 groups per lane covering the CTA's KW-wide column slice
__global__ void __launch_bounds__(kBwdThreads, 1) amax_bwd_dw_kernel(const float* __restrict__ g,
                                                                     const int32_t* __restrict__ arg, mrg_act x,
                                                                     int64_t N, int D, int KW, int kslices,
                                                                     float* __restrict__ part,
                                                                     float* __restrict__ part_b) {
  extern __shared__ float smem[];  // dW_s [D][KW] | db_s [D]
  float* dW_s = smem;
  float* db_s = dW_s + (size_t)D * KW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slice = blockIdx.x % kslices;
  const int k0 = slice * KW;
  const int kw4 = min(KW, D - k0) >> 2;   // float4 groups in this slice (KW % 4 == 0)
  for (int i = threadIdx.x; i < D * KW; i += blockDim.x) dW_s[i] = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) db_s[i] = 0.f;
  __syncthreads();
  const bool affine = x.scale != nullptr, relu = x.relu != 0;
  float4 sc[NT], sh[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int c4 = lane + 32 * t;
    sc[t] = (affine && c4 < kw4) ? ldg4(x.scale + k0 + 4 * c4) : make_float4(1.f, 1.f, 1.f, 1.f);
    sh[t] = (affine && c4 < kw4) ? ldg4(x.shift + k0 + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // Warp w owns the dW rows f = w + 32 j: no other warp ever touches them, so there is no block-level
  // synchronisation in the node loop and every warp keeps a batch of gathered x rows in flight.
  constexpr int FB = 4;                       // features (gathered rows) in flight per batch
  const int nf = (D - warp + kBwdWarps - 1) / kBwdWarps;   // features owned by this warp (<= 32)
  const int64_t stride = gridDim.x / kslices;
  for (int64_t n = blockIdx.x / kslices; n < N; n += stride) {
    // lane j holds (arg, g) of the warp's j-th feature
    const int fl = warp + kBwdWarps * lane;
    const int32_t a_l = (lane < nf) ? __ldg(arg + (size_t)n * D + fl) : -1;
    const float g_l = (lane < nf) ? __ldg(g + (size_t)n * D + fl) : 0.f;
    for (int j0 = 0; j0 < nf; j0 += FB) {
      float4 xv[FB][NT];
      int32_t ev[FB];
#pragma unroll
      for (int u = 0; u < FB; ++u) {
        ev[u] = (j0 + u < nf) ? __shfl_sync(0xffffffffu, a_l, (j0 + u) & 31) : -1;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const int c4 = lane + 32 * t;
          xv[u][t] = (ev[u] >= 0 && c4 < kw4) ? ldg4(x.data + (size_t)ev[u] * D + k0 + 4 * c4)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < FB; ++u) {
        const float gv = __shfl_sync(0xffffffffu, g_l, (j0 + u) & 31);
        if (ev[u] >= 0) {   // warp-uniform
          const int f = warp + kBwdWarps * (j0 + u);
          if (lane == 0 && slice == 0) db_s[f] += gv;
          float* wr = dW_s + (size_t)f * KW;
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            const int c4 = lane + 32 * t;
            if (c4 < kw4) {
              float4 v = xv[u][t];
              if (affine) {
                v.x = fmaf(sc[t].x, v.x, sh[t].x); v.y = fmaf(sc[t].y, v.y, sh[t].y);
                v.z = fmaf(sc[t].z, v.z, sh[t].z); v.w = fmaf(sc[t].w, v.w, sh[t].w);
              }
              if (relu) {
                v.x = v.x > 0.f ? v.x : 0.f; v.y = v.y > 0.f ? v.y : 0.f;
                v.z = v.z > 0.f ? v.z : 0.f; v.w = v.w > 0.f ? v.w : 0.f;
              }
              float4 a = *reinterpret_cast<float4*>(wr + 4 * c4);
              a.x = fmaf(gv, v.x, a.x); a.y = fmaf(gv, v.y, a.y); a.z = fmaf(gv, v.z, a.z); a.w = fmaf(gv, v.w, a.w);
              *reinterpret_cast<float4*>(wr + 4 * c4) = a;
            }
          }
        }
      }
    }
  }
  __syncthreads();
  float* p = part + (size_t)blockIdx.x * D * KW;
  for (int i = threadIdx.x; i < D * KW; i += blockDim.x) p[i] = dW_s[i];
  if (slice == 0)
    for (int i = threadIdx.x; i < D; i += blockDim.x) part_b[(size_t)(blockIdx.x / kslices) * D + i] = db_s[i];
}

// fold the per-CTA partials in CTA order: dW[f][k0+k] = sum_c part[c*kslices+slice][f][k]
__global__ void amax_bwd_dw_fold_kernel(const float* __restrict__ part, const float* __restrict__ part_b, int nctas,
                                        int kslices, int D, int KW, float* __restrict__ dW, float* __restrict__ db) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int groups = nctas / kslices;
  if (idx < D * D) {
    const int f = idx / D, k = idx % D;
    const int slice = k / KW, kk = k % KW;
    float t = 0.f;
    for (int c = 0; c < groups; ++c) t += part[((size_t)(c * kslices + slice) * D + f) * KW + kk];
    dW[idx] = t;
  } else if (idx < D * D + D) {
    const int f = idx - D * D;
    float t = 0.f;
    for (int c = 0; c < groups; ++c) t += part_b[(size_t)c * D + f];
    if (db) db[f] = t;
  }
}

}  // namespace mrg

using namespace mrg;

static inline int dw_kw(int D, int ks) { return ((D + ks - 1) / ks + 3) / 4 * 4; }
static inline int dw_kslices(int D) {
  int ks = 1;
  while ((size_t)D * dw_kw(D, ks) * 4 > 190 * 1024) ++ks;
  return ks;
}
static inline int dw_grid(int D) {
  const int ks = dw_kslices(D);
  return kNumSMs / ks * ks;
}

extern "C" size_t mrg_amax_bwd_workspace_bytes(int32_t D) {
  const int ks = dw_kslices(D), KW = dw_kw(D, ks);
  return (size_t)dw_grid(D) * D * KW * sizeof(float) + (size_t)dw_grid(D) * D * sizeof(float) + 256;
}

extern "C" int mrg_amax_bwd(const float* g, const int32_t* arg, mrg_act x, const float* W, const int32_t* csr_ptr,
                            const int32_t* csr_eid, const int32_t* chunk_first, const int32_t* chunk_seg, int64_t N,
                            int64_t E, int64_t max_chunks, int32_t D, float* dX, float* dW, float* db, void* workspace,
                            size_t workspace_bytes, void* stream) {
  MRG_CHECK_ARG(g && arg && x.data && W && csr_ptr && chunk_first && chunk_seg && workspace, "amax_bwd: null pointer");
  MRG_CHECK_ARG(E == 0 || csr_eid, "amax_bwd: null csr_eid");
  MRG_CHECK_ARG(valid_D(D) && D <= 256, "amax_bwd: D must be a multiple of 4 and <= 256");
  if (workspace_bytes < mrg_amax_bwd_workspace_bytes(D)) {
    set_error("amax_bwd: workspace too small");
    return MRG_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if (dX && E > 0) {
    const bool w_smem = (size_t)D * D * 4 <= 200 * 1024;
    const size_t smem = w_smem ? (size_t)D * D * 4 : 0;
    const int grid = kNumSMs;  // persistent: one 32-warp CTA per SM walks the chunk list
    (void)max_chunks;
#define LDX(NJ, WS)                                                                                              \
  do {                                                                                                           \
    if (smem > 48 * 1024) {                                                                                      \
      e = cudaFuncSetAttribute(amax_bwd_dx_kernel<NJ, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return cuda_fail(e, "amax_bwd dx smem attr");                                        \
    }                                                                                                            \
    amax_bwd_dx_kernel<NJ, WS><<<grid, kBwdThreads, smem, st>>>(g, arg, W, csr_ptr, csr_eid, chunk_first, chunk_seg, N, D, dX); \
  } while (0)
    const int nj = (D + 31) / 32;
    if (nj <= 2) { if (w_smem) LDX(2, true); else LDX(2, false); }
    else if (nj <= 4) { if (w_smem) LDX(4, true); else LDX(4, false); }
    else if (nj <= 7) { if (w_smem) LDX(7, true); else LDX(7, false); }
    else { if (w_smem) LDX(8, true); else LDX(8, false); }
#undef LDX
  }
  if (dW) {
    const int ks = dw_kslices(D), KW = dw_kw(D, ks), grid = dw_grid(D);
    float* part = (float*)workspace;
    float* part_b = part + (size_t)grid * D * KW;
    const size_t smem = ((size_t)D * KW + (size_t)D) * 4;
    const int nt = (KW / 4 + 31) / 32;
#define LDW(NT)                                                                                               \
  do {                                                                                                        \
    if (smem > 48 * 1024) {                                                                                   \
      e = cudaFuncSetAttribute(amax_bwd_dw_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return cuda_fail(e, "amax_bwd dw smem attr");                                     \
    }                                                                                                         \
    amax_bwd_dw_kernel<NT><<<grid, kBwdThreads, smem, st>>>(g, arg, x, N, D, KW, ks, part, part_b);               \
  } while (0)
    if (nt <= 1) LDW(1);
    else LDW(2);
#undef LDW
    const int n = D * D + D;
    amax_bwd_dw_fold_kernel<<<(n + 255) / 256, 256, 0, st>>>(part, part_b, grid, ks, D, KW, dW, db);
  }
  MRG_LAUNCH_CHECK("amax_bwd");
  return MRG_OK;
}
