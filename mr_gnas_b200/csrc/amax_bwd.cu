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
//
// Round 2: the routing is turned ONCE per call (amax_route_kernel) into two bit tables over the dst-CSR
// positions p of the routed edges, so that neither product searches for its pairs any more:
//   rmask[p][j]  (32-bit words, j = f / 32)      bit f % 32 set  <=>  feature f of dst(p) is routed to position p
//   fmask[p / 64][f] (64-bit words)              bit p % 64 set  <=>  the same, indexed by (64-row window, feature)
// dX walks rmask row by row (ascending f => fixed summation order); dW streams 64-row windows of x through
// shared memory and each warp reads the 64-bit row sets of the features it owns.  Both are deterministic:
// bitwise OR is order independent, dX rows have one writer, dW accumulators have one owner warp and the per-CTA
// partials are folded in CTA order.
#include "common.cuh"
#include "pipe.cuh"

namespace mrg {

constexpr int kBwdThreads = 1024;  // one CTA per SM (W resident in smem) -> fill it with 32 warps
constexpr int kBwdWarps = kBwdThreads / 32;
constexpr int kRW = 8;             // rmask words per position (D <= 256)
constexpr int kWin = 64;           // rows per fmask window == rows per staged dW window

// ---------------------------------------------------------------------------------------
// routing tables.  One thread per (destination, feature): dst-CSR position of edge arg[n,f] by binary search
// (edge ids ascend inside a destination), then two atomic ORs.
// ---------------------------------------------------------------------------------------
__global__ void amax_route_kernel(const int32_t* __restrict__ arg, const int32_t* __restrict__ ptr,
                                  const int32_t* __restrict__ eid, int64_t N, int D, uint32_t* __restrict__ rmask,
                                  unsigned long long* __restrict__ fmask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * D) return;
  const int64_t n = i / D;
  const int f = (int)(i - n * D);
  const int32_t a = __ldg(arg + i);
  if (a < 0) return;
  int32_t l = __ldg(ptr + n), h = __ldg(ptr + n + 1);
  const int32_t hi = h;
  while (l < h) {
    const int32_t mid = (l + h) >> 1;
    if (__ldg(eid + mid) < a) l = mid + 1; else h = mid;
  }
  if (l >= hi || __ldg(eid + l) != a) return;
  atomicOr(rmask + (size_t)l * kRW + (f >> 5), 1u << (f & 31));
  atomicOr(fmask + (size_t)(l / kWin) * D + f, 1ull << (l % kWin));
}

// ---------------------------------------------------------------------------------------
// dX: one warp per <=32-edge chunk of a destination's CSR row.  Lane l keeps g[n, l + 32 j] in registers and loads
// word l of the position's rmask row; the set bits are walked in ascending feature order and the W rows are
// accumulated from shared memory (W resident when it fits).  Rows without a routed feature are written as zeros.
// ---------------------------------------------------------------------------------------
template <int NJ, bool W_SMEM>
__global__ void __launch_bounds__(kBwdThreads, 1) amax_bwd_dx_kernel(
    const float* __restrict__ g, const uint32_t* __restrict__ rmask, const float* __restrict__ W,
    const int32_t* __restrict__ ptr, const int32_t* __restrict__ eid, const int32_t* __restrict__ chunk_first,
    const int32_t* __restrict__ chunk_seg, int64_t nseg, int D, float* __restrict__ dX,
    unsigned int* __restrict__ next_chunk) {
  extern __shared__ float smem_w[];  // [D][D] when W_SMEM
  constexpr int NT4 = (NJ + 3) / 4;   // float4 column groups per lane
  const int lane = threadIdx.x & 31;
  if (W_SMEM) {
    for (int i = threadIdx.x * 4; i < D * D; i += blockDim.x * 4)
      *reinterpret_cast<float4*>(smem_w + i) = ldg4(W + i);
    __syncthreads();
  }
  const float* Wp = W_SMEM ? smem_w : W;
  const int D4 = D >> 2;
  // Chunks cost between ~30 (a hub's rows, hardly any routed feature) and ~3,500 (the single edge of a
  // degree-1 destination carries all D features) instructions: a static round-robin left the SM sub-partitions
  // between 24 % and 76 % busy (ncu, round 2).  Warps now draw chunks from a global counter; every dX row still
  // has exactly one writer and a fixed summation order, so the result does not depend on who draws what.
  const int64_t nchunks = chunk_first[nseg];
  for (;;) {
    unsigned int ch_u = 0;
    if (lane == 0) ch_u = atomicAdd(next_chunk, 1u);
    const int64_t ch = __shfl_sync(0xffffffffu, ch_u, 0);
    if (ch >= nchunks) break;
    const int64_t n = chunk_seg[ch];
    const int32_t p0 = ptr[n], p1 = ptr[n + 1];
    const int32_t lo = p0 + (int32_t)(ch - chunk_first[n]) * MRG_CHUNK_ROWS;
    const int32_t hi = min(lo + MRG_CHUNK_ROWS, p1);
    float gr[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int f = lane + 32 * j;
      gr[j] = f < D ? __ldg(g + (size_t)n * D + f) : 0.f;
    }
    // lane l owns position lo + l of the chunk: its edge id, and (lanes < kRW of the broadcast below) mask words
    const int32_t my_e = lo + lane < hi ? __ldg(eid + lo + lane) : 0;
    uint32_t w_next = lane < kRW ? __ldg(rmask + (size_t)lo * kRW + lane) : 0u;
    for (int32_t pz = lo; pz < hi; ++pz) {
      const uint32_t w_cur = w_next;
      if (pz + 1 < hi && lane < kRW) w_next = __ldg(rmask + (size_t)(pz + 1) * kRW + lane);
      const int32_t e = __shfl_sync(0xffffffffu, my_e, pz - lo);
      float4 acc[NT4];
#pragma unroll
      for (int t = 0; t < NT4; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (__ballot_sync(0xffffffffu, w_cur != 0u)) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          uint32_t m = __shfl_sync(0xffffffffu, w_cur, j);
          while (m) {
            const int src_lane = __ffs(m) - 1;
            m &= m - 1;
            const float gv = __shfl_sync(0xffffffffu, gr[j], src_lane);
            const float* wrow = Wp + (src_lane + 32 * j) * D + 4 * lane;
#pragma unroll
            for (int t = 0; t < NT4; ++t) {
              if (lane + 32 * t < D4) {
                const float4 w = W_SMEM ? *reinterpret_cast<const float4*>(wrow + 128 * t) : ldg4(wrow + 128 * t);
                acc[t].x = fmaf(gv, w.x, acc[t].x); acc[t].y = fmaf(gv, w.y, acc[t].y);
                acc[t].z = fmaf(gv, w.z, acc[t].z); acc[t].w = fmaf(gv, w.w, acc[t].w);
              }
            }
          }
        }
      }
      float* out = dX + (size_t)e * D + 4 * lane;
#pragma unroll
      for (int t = 0; t < NT4; ++t)
        if (lane + 32 * t < D4) st_stream4(out + 128 * t, acc[t]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// dW / db.  Every (destination, feature) pair routes g[n,f] to ONE edge row, so dW^T is a sum of
// N*D rank-1 contributions g[n,f] * x(arg[n,f],:).
//
// Warp w of a CTA owns the dW rows f in [w*NF, (w+1)*NF) with the accumulators in REGISTERS (acc[u][t]: feature
// w*NF+u, column k0 + lane + 32 t); the CTA owns a contiguous range of 64-row windows of dst-CSR positions.  For
// every window lane u reads the 64-bit row set of its feature (fmask) and the warp walks the set bits: the x row
// of a pair is read STRAIGHT from global memory through L1 (7 coalesced 128-byte requests at D = 200), activated
// in registers (lazy BatchNorm + ReLU) and accumulated.  The 16 warps of a CTA sweep the same window range, so a
// row is fetched from HBM once and the ~5 further features routed to it hit L1/L2; the rows of the NEXT window are
// prefetched to L2 while the current one is processed.
//
// Round-2 history (profiles/r02_ncu_top_kernels.md): the shared-memory staged versions -- per-thread cp.async,
// then cp.async.bulk + mbarrier rings, with the routing prefetched a window ahead -- all sat at ~7 us per window
// whatever their instruction count: one CTA-wide barrier per window made every window cost what its slowest warp
// cost (31 % of the stall samples on that barrier, 0.44 instructions per cycle and scheduler).  Without a staging
// buffer there is nothing to hand over and no barrier: warps run independently.
// No atomics, one owner per accumulator, fixed order (windows ascending, rows ascending) -> deterministic.
// Partials: part[cta][D][KW] (+ db in part_b[group][D]), folded in CTA order.
// ---------------------------------------------------------------------------------------
constexpr int kDwThreads = 512;
constexpr int kDwWarps = kDwThreads / 32;

template <int NF, int NT>
__global__ void __launch_bounds__(kDwThreads, 1) amax_bwd_dw_kernel(
    const float* __restrict__ g, const unsigned long long* __restrict__ fmask, mrg_act x,
    const int32_t* __restrict__ eid, const int32_t* __restrict__ csr_dst, int64_t E, int D, int KW, int kslices,
    float* __restrict__ part, float* __restrict__ part_b) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slice = blockIdx.x % kslices, group = blockIdx.x / kslices, ngroups = gridDim.x / kslices;
  const int k0 = slice * KW;
  const int kw = min(KW, D - k0);   // valid columns of this slice
  const bool affine = x.scale != nullptr, relu = x.relu != 0;
  const int64_t nwin = (E + kWin - 1) / kWin;
  const int64_t W0 = nwin * group / ngroups, W1 = nwin * (group + 1) / ngroups;

  float acc[NF][NT], sc[NT], sh[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int k = lane + 32 * t;
    sc[t] = (affine && k < kw) ? __ldg(x.scale + k0 + k) : 1.f;
    sh[t] = (affine && k < kw) ? __ldg(x.shift + k0 + k) : 0.f;
#pragma unroll
    for (int u = 0; u < NF; ++u) acc[u][t] = 0.f;
  }
  float db_acc = 0.f;
  const int f_l = warp * NF + lane;                 // the feature whose row set this lane holds
  const bool f_ok = lane < NF && f_l < D;
  const float* xcol = x.data + k0 + lane;
  const float* gcol = g + warp * NF;

  // per-window state, loaded one window ahead: row set of my feature, edge id + destination of rows lane, lane + 32
  unsigned long long m_n = 0ull;
  int32_t e0_n = 0, e1_n = 0, d0_n = 0, d1_n = 0;
  auto load_window = [&](int64_t w) {
    m_n = 0ull;
    if (w < W1) {
      const int64_t p0 = w * kWin + lane, p1 = p0 + 32;
      if (f_ok) m_n = __ldg(fmask + (size_t)w * D + f_l);
      e0_n = p0 < E ? __ldg(eid + p0) : 0;
      e1_n = p1 < E ? __ldg(eid + p1) : 0;
      d0_n = __ldg(csr_dst + p0);            // csr_dst is padded by one window
      d1_n = __ldg(csr_dst + p1);
    }
  };
  load_window(W0);
  for (int64_t w = W0; w < W1; ++w) {
    const unsigned long long m = m_n;
    const int32_t e0 = e0_n, e1 = e1_n, d0 = d0_n, d1 = d1_n;
    load_window(w + 1);
    // rows of the next window towards L2 / L1 while this one is processed: warp w takes rows w, w+16, w+32, w+48
    if (w + 1 < W1) {
#pragma unroll
      for (int i = 0; i < kWin / kDwWarps; ++i) {
        const int r = warp + kDwWarps * i;
        const int32_t e = __shfl_sync(0xffffffffu, r < 32 ? e0_n : e1_n, r & 31);
        if (lane < NT) asm volatile("prefetch.global.L2 [%0];" ::"l"(xcol - lane + (size_t)e * D + 32 * lane));
      }
    }
#pragma unroll
    for (int u = 0; u < NF; ++u) {
      unsigned long long mu = __shfl_sync(0xffffffffu, m, u);
      while (mu) {
        const int r = __ffsll((long long)mu) - 1;
        mu &= mu - 1;
        const int32_t e = __shfl_sync(0xffffffffu, r < 32 ? e0 : e1, r & 31);
        const int32_t n = __shfl_sync(0xffffffffu, r < 32 ? d0 : d1, r & 31);
        const float gv = __ldg(gcol + (size_t)n * D + u);
        const float* row = xcol + (size_t)e * D;
        float v[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) v[t] = (lane + 32 * t < kw) ? __ldg(row + 32 * t) : 0.f;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          float a = fmaf(sc[t], v[t], sh[t]);
          if (relu) a = fmaxf(a, 0.f);
          acc[u][t] = fmaf(gv, a, acc[u][t]);
        }
        if (lane == u) db_acc += gv;
      }
    }
  }
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
static inline size_t align256(size_t b) { return (b + 255) / 256 * 256; }
static inline size_t dw_part_bytes(int D) {
  const int ks = dw_kslices(D), KW = dw_kw(D, ks);
  return align256((size_t)dw_grid(D) * D * KW * sizeof(float) + (size_t)dw_grid(D) * D * sizeof(float));
}
static inline size_t rmask_bytes(int64_t E) { return align256((size_t)(E > 0 ? E : 1) * kRW * sizeof(uint32_t)); }
static inline size_t fmask_bytes(int64_t E, int D) {
  return align256((size_t)((E + kWin - 1) / kWin + 1) * D * sizeof(unsigned long long));
}

extern "C" size_t mrg_amax_bwd_workspace_bytes(int64_t N, int64_t E, int32_t D) {
  (void)N;
  return dw_part_bytes(D) + rmask_bytes(E) + fmask_bytes(E, D) + 512;   // dW partials | rmask | fmask | counter
}

extern "C" int mrg_amax_bwd(const float* g, const int32_t* arg, mrg_act x, const float* W, const int32_t* csr_ptr,
                            const int32_t* csr_eid, const int32_t* csr_dst, const int32_t* chunk_first,
                            const int32_t* chunk_seg, int64_t N, int64_t E, int64_t max_chunks, int32_t D, float* dX,
                            float* dW, float* db, void* workspace, size_t workspace_bytes, void* stream) {
  MRG_CHECK_ARG(g && arg && x.data && W && csr_ptr && chunk_first && chunk_seg && workspace, "amax_bwd: null pointer");
  MRG_CHECK_ARG(E == 0 || (csr_eid && csr_dst), "amax_bwd: null csr_eid / csr_dst");
  MRG_CHECK_ARG(valid_D(D) && D <= 256, "amax_bwd: D must be a multiple of 4 and <= 256");
  if (workspace_bytes < mrg_amax_bwd_workspace_bytes(N, E, D)) {
    set_error("amax_bwd: workspace too small");
    return MRG_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  (void)max_chunks;
  uint32_t* rmask = (uint32_t*)((char*)workspace + dw_part_bytes(D));
  unsigned long long* fmask = (unsigned long long*)((char*)rmask + rmask_bytes(E));
  unsigned int* counter = (unsigned int*)((char*)fmask + fmask_bytes(E, D));     // dX chunk counter (zeroed with the tables)
  e = cudaMemsetAsync(rmask, 0, rmask_bytes(E) + fmask_bytes(E, D) + 256, st);
  if (e != cudaSuccess) return cuda_fail(e, "amax_bwd memset");
  if (N > 0 && E > 0)
    amax_route_kernel<<<(unsigned)((N * D + 255) / 256), 256, 0, st>>>(arg, csr_ptr, csr_eid, N, D, rmask, fmask);
  if (dX && E > 0) {
    const bool w_smem = (size_t)D * D * 4 <= 200 * 1024;
    const size_t smem = w_smem ? (size_t)D * D * 4 : 0;
    const int grid = kNumSMs;  // persistent: one 32-warp CTA per SM walks the chunk list
#define LDX(NJ, WS)                                                                                              \
  do {                                                                                                           \
    if (smem > 48 * 1024) {                                                                                      \
      e = cudaFuncSetAttribute(amax_bwd_dx_kernel<NJ, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return cuda_fail(e, "amax_bwd dx smem attr");                                        \
    }                                                                                                            \
    amax_bwd_dx_kernel<NJ, WS><<<grid, kBwdThreads, smem, st>>>(g, rmask, W, csr_ptr, csr_eid, chunk_first, chunk_seg, N, D, dX, counter); \
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
    const int nf = (D + kDwWarps - 1) / kDwWarps, nt = (KW + 31) / 32;
#define LDW(NF, NT)                                                                                               \
  do {                                                                                                            \
    e = cudaFuncSetAttribute(amax_bwd_dw_kernel<NF, NT>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);      \
    if (e != cudaSuccess) return cuda_fail(e, "amax_bwd dw carveout attr");                                       \
    amax_bwd_dw_kernel<NF, NT><<<grid, kDwThreads, 0, st>>>(g, fmask, x, csr_eid, csr_dst, E, D, KW, ks, part,     \
                                                            part_b);                                              \
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
