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
__global__ void __launch_bounds__(kBwdThreads, 1) amax_bwd_dx_kernel(  // NT4 = ceil(NJ / 4) float4 groups per lane
    const float* __restrict__ g, const int32_t* __restrict__ arg, const float* __restrict__ W,
    const int32_t* __restrict__ ptr, const int32_t* __restrict__ eid, const int32_t* __restrict__ chunk_first,
    const int32_t* __restrict__ chunk_seg, int64_t nseg, int D, float* __restrict__ dX) {
  extern __shared__ float smem_w[];  // [D][D] when W_SMEM
  constexpr int NT4 = (NJ + 3) / 4;
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
    const int D4 = D >> 2;
    int32_t e_next = __ldg(eid + lo);
    for (int32_t pz = lo; pz < hi; ++pz) {
      const int32_t e = e_next;
      if (pz + 1 < hi) e_next = __ldg(eid + pz + 1);
      // lane owns the float4 column groups c4 = lane + 32 t: 128-bit shared-memory reads of W and 128-bit stores
      float4 acc[NT4];
#pragma unroll
      for (int t = 0; t < NT4; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        unsigned m = __ballot_sync(0xffffffffu, ar[j] == e);
        while (m) {
          const int src_lane = __ffs(m) - 1;
          m &= m - 1;
          const float gv = __shfl_sync(0xffffffffu, gr[j], src_lane);
          const float* wrow = Wp + (size_t)(src_lane + 32 * j) * D;
#pragma unroll
          for (int t = 0; t < NT4; ++t) {
            const int c4 = lane + 32 * t;
            if (c4 < D4) {
              const float4 w = W_SMEM ? *reinterpret_cast<const float4*>(wrow + 4 * c4) : ldg4(wrow + 4 * c4);
              acc[t].x = fmaf(gv, w.x, acc[t].x); acc[t].y = fmaf(gv, w.y, acc[t].y);
              acc[t].z = fmaf(gv, w.z, acc[t].z); acc[t].w = fmaf(gv, w.w, acc[t].w);
            }
          }
        }
      }
      float* out = dX + (size_t)e * D;
#pragma unroll
      for (int t = 0; t < NT4; ++t) {
        const int c4 = lane + 32 * t;
        if (c4 < D4) st_stream4(out + 4 * c4, acc[t]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// dW / db.  Every (destination, feature) pair routes g[n,f] to ONE edge row, so dW^T is a sum of
// N*D rank-1 contributions g[n,f] * x(arg[n,f],:).  Gathering an 800-byte x row per pair would move
// N*D*b bytes (5x the edge tensor at C1); instead each CTA owns a contiguous range of dst-CSR
// positions, streams the x rows of that range ONCE into shared memory (cp.async, 3-stage ring of
// kDwCap-row windows, lazy BatchNorm+ReLU applied in place when a window lands) and resolves every
// pair against the staged window by a binary search of its edge id in the window's (ascending) edge
// list.  Warp w owns the dW rows f in [w*NF, (w+1)*NF) with the accumulators in REGISTERS
// (acc[u][t]: feature w*NF+u, column k0 + lane + 32 t): no shared-memory read-modify-write, no atomics,
// a fixed accumulation order -> deterministic.  Destinations that straddle windows / CTAs are simply
// seen by both; each pair is counted where its edge position falls.
// Partials: part[cta][D][KW] (+ db in part_b[group][D]), folded in CTA order.
// ---------------------------------------------------------------------------------------
constexpr int kDwThreads = 512;
constexpr int kDwWarps = kDwThreads / 32;
constexpr int kDwCap = 64;      // edge rows per staged window
constexpr int kDwStages = 3;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// pos[n,f] = dst-CSR position of edge arg[n,f] (edge ids ascend inside a destination), -1 when no gradient flows
__global__ void amax_pos_kernel(const int32_t* __restrict__ arg, const int32_t* __restrict__ ptr,
                                const int32_t* __restrict__ eid, int64_t N, int D, int32_t* __restrict__ pos) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * D) return;
  const int64_t n = i / D;
  const int32_t a = __ldg(arg + i);
  int32_t r = -1;
  if (a >= 0) {
    int32_t l = __ldg(ptr + n), h = __ldg(ptr + n + 1);
    const int32_t hi = h;
    while (l < h) {
      const int32_t mid = (l + h) >> 1;
      if (__ldg(eid + mid) < a) l = mid + 1; else h = mid;
    }
    if (l < hi && __ldg(eid + l) == a) r = l;
  }
  pos[i] = r;
}

template <int NF, int NT>
__global__ void __launch_bounds__(kDwThreads, 1) amax_bwd_dw_kernel(
    const float* __restrict__ g, const int32_t* __restrict__ pos, mrg_act x, const int32_t* __restrict__ ptr,
    const int32_t* __restrict__ eid, int64_t N, int64_t E, int D, int KW, int kslices, float* __restrict__ part,
    float* __restrict__ part_b) {
  extern __shared__ float smem[];
  // layout: xs[kDwStages][kDwCap][KW] | pad[32] | sc[KW] | sh[KW]
  float* xs = smem;
  float* sc_s = xs + (size_t)kDwStages * kDwCap * KW + 32;
  float* sh_s = sc_s + KW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slice = blockIdx.x % kslices, group = blockIdx.x / kslices, ngroups = gridDim.x / kslices;
  const int k0 = slice * KW;
  const int kw = min(KW, D - k0);   // valid columns of this slice (multiple of 4)
  const int kw4 = kw >> 2;
  const bool affine = x.scale != nullptr, relu = x.relu != 0;
  if (affine)
    for (int c = threadIdx.x; c < kw; c += blockDim.x) {
      sc_s[c] = x.scale[k0 + c];
      sh_s[c] = x.shift[k0 + c];
    }
  const int64_t P0 = E * group / ngroups, P1 = E * (group + 1) / ngroups;
  const int nb = (int)((P1 - P0 + kDwCap - 1) / kDwCap);

  float acc[NF][NT];
#pragma unroll
  for (int u = 0; u < NF; ++u)
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[u][t] = 0.f;
  float db_acc = 0.f;
  const int f_l = warp * NF + lane;                 // the feature whose (pos, g) this lane fetches
  const bool f_ok = lane < NF && f_l < D;
  const float* xs_lane = xs + lane;

  auto issue = [&](int b) {   // warp w copies rows w, w+16, ... of window b
    const int64_t w_lo = P0 + (int64_t)b * kDwCap;
    const int nrows = (int)min((int64_t)kDwCap, P1 - w_lo);
    float* st = xs + (size_t)(b % kDwStages) * kDwCap * KW;
    for (int r = warp; r < nrows; r += kDwWarps) {
      const int32_t e = __ldg(eid + w_lo + r);
      const float* src = x.data + (size_t)e * D + k0;
      for (int c4 = lane; c4 < kw4; c4 += 32) cp_async16(st + (size_t)r * KW + 4 * c4, src + 4 * c4);
    }
  };

  if (nb > 0) issue(0);
  cp_async_commit();
  if (nb > 1) issue(1);
  cp_async_commit();

  // first destination whose CSR row reaches past P0: largest n with ptr[n] <= P0 (empty rows skipped below)
  int64_t n_cur = 0;
  {
    int64_t lo = 0, hi = N;   // invariant: ptr[lo] <= P0, answer in [lo, hi)
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ldg(ptr + mid) <= P0) lo = mid; else hi = mid;
    }
    n_cur = lo;
  }
  // (pos, g) of destination n_pf, fetched one destination ahead of use
  int64_t n_pf = n_cur;
  int32_t p_pf = f_ok ? __ldg(pos + (size_t)n_pf * D + f_l) : -1;
  float g_pf = f_ok ? __ldg(g + (size_t)n_pf * D + f_l) : 0.f;

  for (int b = 0; b < nb; ++b) {
    cp_async_wait<1>();
    const int64_t w_lo = P0 + (int64_t)b * kDwCap;
    const int64_t w_hi = min(w_lo + kDwCap, P1);
    const int nrows = (int)(w_hi - w_lo);
    const int stage_off = (b % kDwStages) * kDwCap * KW;
    if (affine || relu) {     // each thread post-processes exactly the 16-byte pieces it copied
      float* st = xs + stage_off;
      for (int r = warp; r < nrows; r += kDwWarps)
        for (int c4 = lane; c4 < kw4; c4 += 32) {
          float4 v = *reinterpret_cast<float4*>(st + (size_t)r * KW + 4 * c4);
          if (affine) {
            const float4 a = *reinterpret_cast<const float4*>(sc_s + 4 * c4);
            const float4 s = *reinterpret_cast<const float4*>(sh_s + 4 * c4);
            v.x = fmaf(a.x, v.x, s.x); v.y = fmaf(a.y, v.y, s.y); v.z = fmaf(a.z, v.z, s.z); v.w = fmaf(a.w, v.w, s.w);
          }
          if (relu) {
            v.x = v.x > 0.f ? v.x : 0.f; v.y = v.y > 0.f ? v.y : 0.f;
            v.z = v.z > 0.f ? v.z : 0.f; v.w = v.w > 0.f ? v.w : 0.f;
          }
          *reinterpret_cast<float4*>(st + (size_t)r * KW + 4 * c4) = v;
        }
    }
    __syncthreads();          // window b complete for everyone; everyone finished computing window b-1
    if (b + 2 < nb) issue(b + 2);
    cp_async_commit();
    const float* st_lane = xs_lane + stage_off;

    int64_t n = n_cur;
    while (n < N) {
      const int32_t p0 = __ldg(ptr + n), p1 = __ldg(ptr + n + 1);
      if (p0 >= w_hi) break;
      if (p1 > w_lo && p1 > p0) {
        int32_t pp;
        float gv;
        if (n == n_pf) {
          pp = p_pf; gv = g_pf;
        } else {
          pp = f_ok ? __ldg(pos + (size_t)n * D + f_l) : -1;
          gv = f_ok ? __ldg(g + (size_t)n * D + f_l) : 0.f;
        }
        if (n + 1 < N && n_pf != n + 1) {   // next destination's routing, in flight while this one is processed
          n_pf = n + 1;
          p_pf = f_ok ? __ldg(pos + (size_t)(n + 1) * D + f_l) : -1;
          g_pf = f_ok ? __ldg(g + (size_t)(n + 1) * D + f_l) : 0.f;
        }
        const int r = (pp >= w_lo && pp < w_hi) ? (int)(pp - w_lo) : -1;
        if (r >= 0) db_acc += gv;
        const unsigned mask = __ballot_sync(0xffffffffu, r >= 0);
        if (mask) {
#pragma unroll
          for (int u = 0; u < NF; ++u) {
            if (mask & (1u << u)) {      // warp-uniform
              const int ru = __shfl_sync(0xffffffffu, r, u);
              const float gu = __shfl_sync(0xffffffffu, gv, u);
              const float* row = st_lane + ru * KW;
#pragma unroll
              for (int t = 0; t < NT; ++t) acc[u][t] = fmaf(gu, row[32 * t], acc[u][t]);
            }
          }
        }
      }
      if (p1 > w_hi) break;     // this destination continues in the next window
      ++n;
    }
    n_cur = n;
  }
  cp_async_wait<0>();
  float* p = part + (size_t)blockIdx.x * D * KW;
#pragma unroll
  for (int u = 0; u < NF; ++u) {
    const int f = warp * NF + u;
    if (f < D) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int k = lane + 32 * t;
        if (k < KW) p[(size_t)f * KW + k] = k < kw ? acc[u][t] : 0.f;
      }
    }
  }
  if (slice == 0 && f_ok) part_b[(size_t)group * D + f_l] = db_acc;
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

static inline int dw_kslices(int D) { return D > 208 ? 2 : 1; }
static inline int dw_kw(int D, int ks) { return ((D + ks - 1) / ks + 3) / 4 * 4; }
static inline int dw_grid(int D) {
  const int ks = dw_kslices(D);
  return kNumSMs / ks * ks;
}
static inline size_t dw_smem(int KW) { return ((size_t)kDwStages * kDwCap * KW + 32) * 4 + 2 * (size_t)KW * 4; }
static inline size_t dw_part_bytes(int D) {
  const int ks = dw_kslices(D), KW = dw_kw(D, ks);
  return ((size_t)dw_grid(D) * D * KW * sizeof(float) + (size_t)dw_grid(D) * D * sizeof(float) + 255) / 256 * 256;
}

extern "C" size_t mrg_amax_bwd_workspace_bytes(int64_t N, int32_t D) {
  return dw_part_bytes(D) + (size_t)N * D * sizeof(int32_t) + 256;   // per-CTA dW partials | pos[N,D]
}

extern "C" int mrg_amax_bwd(const float* g, const int32_t* arg, mrg_act x, const float* W, const int32_t* csr_ptr,
                            const int32_t* csr_eid, const int32_t* chunk_first, const int32_t* chunk_seg, int64_t N,
                            int64_t E, int64_t max_chunks, int32_t D, float* dX, float* dW, float* db, void* workspace,
                            size_t workspace_bytes, void* stream) {
  MRG_CHECK_ARG(g && arg && x.data && W && csr_ptr && chunk_first && chunk_seg && workspace, "amax_bwd: null pointer");
  MRG_CHECK_ARG(E == 0 || csr_eid, "amax_bwd: null csr_eid");
  MRG_CHECK_ARG(valid_D(D) && D <= 256, "amax_bwd: D must be a multiple of 4 and <= 256");
  if (workspace_bytes < mrg_amax_bwd_workspace_bytes(N, D)) {
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
    const size_t smem = dw_smem(KW);
    const int nf = (D + kDwWarps - 1) / kDwWarps, nt = (KW + 31) / 32;
    int32_t* pos = (int32_t*)((char*)workspace + dw_part_bytes(D));
    if (N > 0) amax_pos_kernel<<<(unsigned)((N * D + 255) / 256), 256, 0, st>>>(arg, csr_ptr, csr_eid, N, D, pos);
#define LDW(NF, NT)                                                                                               \
  do {                                                                                                            \
    e = cudaFuncSetAttribute(amax_bwd_dw_kernel<NF, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return cuda_fail(e, "amax_bwd dw smem attr");                                           \
    amax_bwd_dw_kernel<NF, NT><<<grid, kDwThreads, smem, st>>>(g, pos, x, csr_ptr, csr_eid, N, E, D, KW, ks, part, \
                                                               part_b);                                           \
  } while (0)
    if (nf <= 4 && nt <= 2) LDW(4, 2);
    else if (nf <= 8 && nt <= 4) LDW(8, 4);
    else if (nf <= 13 && nt <= 7) LDW(13, 7);
    else LDW(16, 4);
#undef LDW
    const int n = D * D + D;
    amax_bwd_dw_fold_kernel<<<(n + 255) / 256, 256, 0, st>>>(part, part_b, grid, ks, D, KW, dW, db);
  }
  MRG_LAUNCH_CHECK("amax_bwd");
  return MRG_OK;
}
