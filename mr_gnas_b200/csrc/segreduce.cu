// K5/K11: deterministic segmented reductions of gathered fp32 rows (sum / mean / max+arg),
// used forward over the dst-CSR (DGL update_all replacement) and backward over the src-CSC /
// relation segments (index_add_ replacement).  Segments are cut into <=32-row chunks so hub
// nodes and frequent relations spread over many warps; multi-chunk segments are combined in
// ascending chunk order by a second kernel, so results are bit-reproducible.
#include <cub/cub.cuh>

#include "common.cuh"

namespace mrg {

constexpr int CH = MRG_CHUNK_ROWS;

__global__ void chunk_count_kernel(const int32_t* __restrict__ ptr, int64_t nseg, int32_t* __restrict__ cnt) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s > nseg) return;
  cnt[s] = s < nseg ? (ptr[s + 1] - ptr[s] + CH - 1) / CH : 0;
}
__global__ void chunk_fill_kernel(const int32_t* __restrict__ chunk_first, int64_t nseg,
                                  int32_t* __restrict__ chunk_seg) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  for (int32_t k = chunk_first[s]; k < chunk_first[s + 1]; ++k) chunk_seg[k] = (int32_t)s;
}

template <int NV>
struct RedRegs {
  float4 val[NV];
  int4 arg[NV];
};

// Encoded arg: >=0 edge/row id; -1 empty segment; (-2 - id) max attained but value <= 0 under
// ReLU (gradient is gated off; id still recoverable for the bit-exact argmax check).
template <int NV, int KIND>
__device__ __forceinline__ void finalize_seg(RedRegs<NV>& acc, int64_t seg, int len, int D, int D4, int lane,
                                             float alpha, bool gate_relu, const ActRegs<NV>& ares, const mrg_act& res,
                                             int accumulate, float* __restrict__ out, int32_t* __restrict__ arg) {
  float scale = alpha;
  if (KIND == MRG_RED_MEAN) scale = alpha / (float)(len > 0 ? len : 1);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    int c4 = lane + 32 * v;
    if (c4 < D4) {
      const size_t off = (size_t)seg * D + 4 * c4;
      float4 o = acc.val[v];
      if (KIND == MRG_RED_MAX) {
        int4 a = acc.arg[v];
        if (len == 0) {
          o = make_float4(0.f, 0.f, 0.f, 0.f);
          a = make_int4(-1, -1, -1, -1);
        } else if (gate_relu) {
          if (!(o.x > 0.f)) a.x = -2 - a.x;
          if (!(o.y > 0.f)) a.y = -2 - a.y;
          if (!(o.z > 0.f)) a.z = -2 - a.z;
          if (!(o.w > 0.f)) a.w = -2 - a.w;
        }
        if (arg) *reinterpret_cast<int4*>(arg + off) = a;
      } else {
        o = make_float4(scale * o.x, scale * o.y, scale * o.z, scale * o.w);
      }
      if (res.data) {
        float4 rv = ares.apply(ldg4(res.data + off), v);
        o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
      }
      if (accumulate) {
        float4 p = *reinterpret_cast<const float4*>(out + off);
        o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
      }
      st4(out + off, o);
    }
  }
}

template <int NV, int KIND, bool MUL>
__global__ void __launch_bounds__(kThreads) seg_reduce_chunk_kernel(
    mrg_act m, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
    const int32_t* __restrict__ chunk_first, const int32_t* __restrict__ chunk_seg, int64_t nseg, int D,
    const float* __restrict__ mul, const int32_t* __restrict__ mul_idx, float alpha, mrg_act res, int accumulate,
    float* __restrict__ out, int32_t* __restrict__ arg, float* __restrict__ pval, int32_t* __restrict__ parg) {
  const int lane = threadIdx.x & 31;
  const int D4 = D >> 2;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  const int64_t nchunks = chunk_first[nseg];
  ActRegs<NV> am, ares;
  am.init(m, lane, D4);
  ares.init(res, lane, D4);
  for (int64_t ch = warp0; ch < nchunks; ch += nwarps) {
    const int64_t seg = chunk_seg[ch];
    const int32_t cf = chunk_first[seg];
    const int nch = chunk_first[seg + 1] - cf;
    const int32_t p0 = ptr[seg], p1 = ptr[seg + 1];
    const int32_t lo = p0 + (int32_t)(ch - cf) * CH;
    const int32_t hi = min(lo + CH, p1);
    RedRegs<NV> acc;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      acc.val[v] = KIND == MRG_RED_MAX ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
      acc.arg[v] = make_int4(-1, -1, -1, -1);
    }
    for (int32_t j0 = lo; j0 < hi; j0 += 4) {
      int32_t rid[4];
      float4 a[4][NV], b[4][NV];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int32_t j = j0 + u;
        rid[u] = j < hi ? (idx ? __ldg(idx + j) : j) : -1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (rid[u] >= 0) {
          const float* rp = m.data + (size_t)rid[u] * D;
          const float* mp = nullptr;
          if (MUL) mp = mul + (size_t)(mul_idx ? __ldg(mul_idx + rid[u]) : rid[u]) * D;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            int c4 = lane + 32 * v;
            if (c4 < D4) {
              a[u][v] = ld_stream4(rp + 4 * c4);
              if (MUL) b[u][v] = ldg4(mp + 4 * c4);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (rid[u] >= 0) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            int c4 = lane + 32 * v;
            if (c4 < D4) {
              float4 x = am.apply(a[u][v], v);
              if (MUL) { x.x *= b[u][v].x; x.y *= b[u][v].y; x.z *= b[u][v].z; x.w *= b[u][v].w; }
              if (KIND == MRG_RED_MAX) {
                if (x.x > acc.val[v].x) { acc.val[v].x = x.x; acc.arg[v].x = rid[u]; }
                if (x.y > acc.val[v].y) { acc.val[v].y = x.y; acc.arg[v].y = rid[u]; }
                if (x.z > acc.val[v].z) { acc.val[v].z = x.z; acc.arg[v].z = rid[u]; }
                if (x.w > acc.val[v].w) { acc.val[v].w = x.w; acc.arg[v].w = rid[u]; }
              } else {
                acc.val[v].x += x.x; acc.val[v].y += x.y; acc.val[v].z += x.z; acc.val[v].w += x.w;
              }
            }
          }
        }
      }
    }
    if (nch == 1) {
      finalize_seg<NV, KIND>(acc, seg, p1 - p0, D, D4, lane, alpha, m.relu != 0, ares, res, accumulate, out, arg);
    } else {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        int c4 = lane + 32 * v;
        if (c4 < D4) {
          st4(pval + (size_t)ch * D + 4 * c4, acc.val[v]);
          if (KIND == MRG_RED_MAX) *reinterpret_cast<int4*>(parg + (size_t)ch * D + 4 * c4) = acc.arg[v];
        }
      }
    }
  }
}

// Segments with 0 or >1 chunks: one CTA per segment.  Warp w folds chunks c0+w, c0+w+8, ... in
// ascending order, then warp 0 merges the 8 warp partials in warp order (fixed order => bit
// reproducible; MAX ties resolve to the lowest row id explicitly).
template <int NV, int KIND>
__global__ void __launch_bounds__(kThreads) seg_reduce_combine_kernel(
    mrg_act m, const int32_t* __restrict__ ptr, const int32_t* __restrict__ chunk_first, int64_t nseg, int D,
    float alpha, mrg_act res, int accumulate, float* __restrict__ out, int32_t* __restrict__ arg,
    const float* __restrict__ pval, const int32_t* __restrict__ parg) {
  extern __shared__ float smem_f[];  // [kWarpsPerBlock][D] values, then [kWarpsPerBlock][D] args (MAX)
  int32_t* smem_a = reinterpret_cast<int32_t*>(smem_f + kWarpsPerBlock * D);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D4 = D >> 2;
  ActRegs<NV> ares;
  ares.init(res, lane, D4);
  for (int64_t seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
    const int32_t c0 = chunk_first[seg], c1 = chunk_first[seg + 1];
    if (c1 - c0 == 1) continue;  // finalized by the chunk kernel (uniform over the CTA)
    RedRegs<NV> acc;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      acc.val[v] = KIND == MRG_RED_MAX ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
      acc.arg[v] = make_int4(-1, -1, -1, -1);
    }
    for (int32_t ch = c0 + warp; ch < c1; ch += kWarpsPerBlock) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        int c4 = lane + 32 * v;
        if (c4 < D4) {
          float4 x = ldg4(pval + (size_t)ch * D + 4 * c4);
          if (KIND == MRG_RED_MAX) {
            int4 a = *reinterpret_cast<const int4*>(parg + (size_t)ch * D + 4 * c4);
            if (x.x > acc.val[v].x) { acc.val[v].x = x.x; acc.arg[v].x = a.x; }
            if (x.y > acc.val[v].y) { acc.val[v].y = x.y; acc.arg[v].y = a.y; }
            if (x.z > acc.val[v].z) { acc.val[v].z = x.z; acc.arg[v].z = a.z; }
            if (x.w > acc.val[v].w) { acc.val[v].w = x.w; acc.arg[v].w = a.w; }
          } else {
            acc.val[v].x += x.x; acc.val[v].y += x.y; acc.val[v].z += x.z; acc.val[v].w += x.w;
          }
        }
      }
    }
    const int nw = min(c1 - c0, kWarpsPerBlock);  // warps that hold a partial
    __syncthreads();
    if (warp > 0 && warp < nw) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        int c4 = lane + 32 * v;
        if (c4 < D4) {
          *reinterpret_cast<float4*>(smem_f + warp * D + 4 * c4) = acc.val[v];
          if (KIND == MRG_RED_MAX) *reinterpret_cast<int4*>(smem_a + warp * D + 4 * c4) = acc.arg[v];
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      for (int w = 1; w < nw; ++w) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          int c4 = lane + 32 * v;
          if (c4 < D4) {
            float4 x = *reinterpret_cast<const float4*>(smem_f + w * D + 4 * c4);
            if (KIND == MRG_RED_MAX) {
              int4 a = *reinterpret_cast<const int4*>(smem_a + w * D + 4 * c4);
#define MRG_MAXTIE(F)                                                                              \
  if (x.F > acc.val[v].F || (x.F == acc.val[v].F && a.F >= 0 && (acc.arg[v].F < 0 || a.F < acc.arg[v].F))) { \
    acc.val[v].F = x.F;                                                                            \
    acc.arg[v].F = a.F;                                                                            \
  }
              MRG_MAXTIE(x) MRG_MAXTIE(y) MRG_MAXTIE(z) MRG_MAXTIE(w)
#undef MRG_MAXTIE
            } else {
              acc.val[v].x += x.x; acc.val[v].y += x.y; acc.val[v].z += x.z; acc.val[v].w += x.w;
            }
          }
        }
      }
      finalize_seg<NV, KIND>(acc, seg, ptr[seg + 1] - ptr[seg], D, D4, lane, alpha, m.relu != 0, ares, res, accumulate,
                             out, arg);
    }
  }
}

// backward w.r.t. message rows, in edge-id order (coalesced writes)
template <int NV, int KIND>
__global__ void __launch_bounds__(kThreads) seg_reduce_bwd_kernel(const float* __restrict__ g,
                                                                  const int32_t* __restrict__ arg, mrg_act m,
                                                                  const int32_t* __restrict__ dst,
                                                                  const int32_t* __restrict__ ptr, int64_t E,
                                                                  int64_t n_self, int D, float* dm, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int D4 = D >> 2;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  const int64_t rows = E + n_self;
  ActRegs<NV> am;
  const bool gate = m.data != nullptr && m.relu != 0 && KIND != MRG_RED_MAX;
  if (gate) am.init(m, lane, D4);
  for (int64_t row = warp0; row < rows; row += nwarps) {
    const bool self = row >= E;
    const int64_t n = self ? row - E : (int64_t)__ldg(dst + row);
    float scale = 1.f;
    if (!self && KIND == MRG_RED_MEAN) {
      int len = __ldg(ptr + n + 1) - __ldg(ptr + n);
      scale = 1.f / (float)(len > 0 ? len : 1);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      int c4 = lane + 32 * v;
      if (c4 < D4) {
        const size_t off = (size_t)row * D + 4 * c4;
        float4 o = ldg4(g + (size_t)n * D + 4 * c4);
        if (!self) {
          if (KIND == MRG_RED_MAX) {
            int4 a = __ldg(reinterpret_cast<const int4*>(arg + (size_t)n * D + 4 * c4));
            const int e = (int)row;
            o.x = a.x == e ? o.x : 0.f;
            o.y = a.y == e ? o.y : 0.f;
            o.z = a.z == e ? o.z : 0.f;
            o.w = a.w == e ? o.w : 0.f;
          } else {
            o = make_float4(scale * o.x, scale * o.y, scale * o.z, scale * o.w);
            if (gate) {
              float4 s = am.apply(ld_stream4(m.data + off), v);
              o.x = s.x > 0.f ? o.x : 0.f;
              o.y = s.y > 0.f ? o.y : 0.f;
              o.z = s.z > 0.f ? o.z : 0.f;
              o.w = s.w > 0.f ? o.w : 0.f;
            }
          }
        }
        if (accumulate) {
          float4 p = *reinterpret_cast<const float4*>(dm + off);
          o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
        }
        st4(dm + off, o);
      }
    }
  }
}

}  // namespace mrg

using namespace mrg;

extern "C" int64_t mrg_chunk_capacity(int64_t total_rows, int64_t nseg) { return total_rows / CH + nseg + 1; }

extern "C" size_t mrg_chunk_workspace_bytes(int64_t nseg) {
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum((void*)nullptr, tmp, (int32_t*)nullptr, (int32_t*)nullptr, (int)(nseg + 1));
  return ((size_t)(nseg + 1) * sizeof(int32_t) + 255) / 256 * 256 + tmp + 256;
}

extern "C" int mrg_chunk_build(const int32_t* ptr, int64_t nseg, int32_t* chunk_first, int32_t* chunk_seg,
                               void* workspace, size_t workspace_bytes, void* stream) {
  MRG_CHECK_ARG(ptr && chunk_first && chunk_seg && workspace, "chunk_build: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t cnt_bytes = ((size_t)(nseg + 1) * sizeof(int32_t) + 255) / 256 * 256;
  if (workspace_bytes < mrg_chunk_workspace_bytes(nseg)) {
    set_error("chunk_build: workspace too small");
    return MRG_ERR_WORKSPACE;
  }
  int32_t* cnt = (int32_t*)workspace;
  void* tmp = (char*)workspace + cnt_bytes;
  size_t tmp_bytes = workspace_bytes - cnt_bytes;
  const int blocks = (int)((nseg + 1 + 255) / 256);
  chunk_count_kernel<<<blocks, 256, 0, st>>>(ptr, nseg, cnt);
  cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, chunk_first, (int)(nseg + 1), st);
  if (e != cudaSuccess) return cuda_fail(e, "chunk_build scan");
  if (nseg > 0) chunk_fill_kernel<<<(int)((nseg + 255) / 256), 256, 0, st>>>(chunk_first, nseg, chunk_seg);
  MRG_LAUNCH_CHECK("chunk_build");
  return MRG_OK;
}

extern "C" size_t mrg_seg_reduce_workspace_bytes(int64_t max_chunks, int32_t D, int32_t kind) {
  size_t per = (size_t)max_chunks * D * sizeof(float);
  return (kind == MRG_RED_MAX ? 2 * per : per) + 256;
}

extern "C" int mrg_seg_reduce_fwd(int32_t kind, mrg_act m, const int32_t* ptr, const int32_t* idx,
                                  const int32_t* chunk_first, const int32_t* chunk_seg, int64_t nseg,
                                  int64_t max_chunks, int32_t D, const float* mul, const int32_t* mul_idx, float alpha,
                                  mrg_act residual, int32_t accumulate, float* out, int32_t* arg, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  MRG_CHECK_ARG(m.data && ptr && chunk_first && chunk_seg && out && workspace, "seg_reduce_fwd: null pointer");
  MRG_CHECK_ARG(valid_D(D), "seg_reduce_fwd: D");
  MRG_CHECK_ARG(kind == MRG_RED_SUM || kind == MRG_RED_MEAN || kind == MRG_RED_MAX, "seg_reduce_fwd: kind");
  MRG_CHECK_ARG(!(kind == MRG_RED_MAX && mul), "seg_reduce_fwd: max with multiplier unsupported");
  if (nseg <= 0) return MRG_OK;
  if (workspace_bytes < mrg_seg_reduce_workspace_bytes(max_chunks, D, kind)) {
    set_error("seg_reduce_fwd: workspace too small");
    return MRG_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* pval = (float*)workspace;
  int32_t* parg = (int32_t*)(pval + (size_t)max_chunks * D);
  const int want_c = stats_grid(max_chunks);
  const int grid_s = (int)(nseg < 4 * kMaxParts ? nseg : 4 * kMaxParts);
  const size_t comb_smem = (size_t)kWarpsPerBlock * D * 8;
#define L(KIND)                                                                                                   \
  MRG_DISPATCH_NV(D, {                                                                                            \
    if (mul) seg_reduce_chunk_kernel<NV, KIND, true><<<resident_grid(seg_reduce_chunk_kernel<NV, KIND, true>, 0, want_c), kThreads, 0, st>>>(m, ptr, idx, chunk_first, chunk_seg, nseg, D, mul, mul_idx, alpha, residual, accumulate, out, arg, pval, parg); \
    else seg_reduce_chunk_kernel<NV, KIND, false><<<resident_grid(seg_reduce_chunk_kernel<NV, KIND, false>, 0, want_c), kThreads, 0, st>>>(m, ptr, idx, chunk_first, chunk_seg, nseg, D, mul, mul_idx, alpha, residual, accumulate, out, arg, pval, parg); \
    seg_reduce_combine_kernel<NV, KIND><<<grid_s, kThreads, comb_smem, st>>>(m, ptr, chunk_first, nseg, D, alpha, residual, accumulate, out, arg, pval, parg); \
  })
  if (kind == MRG_RED_SUM) L(MRG_RED_SUM);
  else if (kind == MRG_RED_MEAN) L(MRG_RED_MEAN);
  else L(MRG_RED_MAX);
#undef L
  MRG_LAUNCH_CHECK("seg_reduce_fwd");
  return MRG_OK;
}

extern "C" int mrg_seg_reduce_bwd(int32_t kind, const float* g, const int32_t* arg, const float* out, mrg_act m,
                                  const int32_t* dst, const int32_t* ptr, int64_t E, int64_t n_self, int32_t D,
                                  float* dm, int32_t accumulate, void* stream) {
  (void)out;
  MRG_CHECK_ARG(g && dm && (dst || E == 0), "seg_reduce_bwd: null pointer");
  MRG_CHECK_ARG(valid_D(D), "seg_reduce_bwd: D");
  MRG_CHECK_ARG(kind != MRG_RED_MAX || arg, "seg_reduce_bwd: max needs arg");
  MRG_CHECK_ARG(kind != MRG_RED_MEAN || ptr, "seg_reduce_bwd: mean needs ptr");
  MRG_CHECK_ARG(kind == MRG_RED_SUM || kind == MRG_RED_MEAN || kind == MRG_RED_MAX, "seg_reduce_bwd: kind");
  if (E + n_self <= 0) return MRG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int want = stats_grid(E + n_self);
#define grid resident_grid(seg_reduce_bwd_kernel<NV, MRG_RED_SUM>, 0, want)
  if (kind == MRG_RED_SUM) MRG_DISPATCH_NV(D, seg_reduce_bwd_kernel<NV, MRG_RED_SUM><<<grid, kThreads, 0, st>>>(g, arg, m, dst, ptr, E, n_self, D, dm, accumulate));
  else if (kind == MRG_RED_MEAN) MRG_DISPATCH_NV(D, seg_reduce_bwd_kernel<NV, MRG_RED_MEAN><<<grid, kThreads, 0, st>>>(g, arg, m, dst, ptr, E, n_self, D, dm, accumulate));
  else MRG_DISPATCH_NV(D, seg_reduce_bwd_kernel<NV, MRG_RED_MAX><<<grid, kThreads, 0, st>>>(g, arg, m, dst, ptr, E, n_self, D, dm, accumulate));
#undef grid
  MRG_LAUNCH_CHECK("seg_reduce_bwd");
  return MRG_OK;
}
