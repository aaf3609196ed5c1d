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
// N*D rank-1 contributions g[n,f] * x(arg[n,f],:).  Each CTA owns a contiguous range of 64-row windows of
// dst-CSR positions and streams the x rows of that range ONCE through a 3-stage shared-memory ring:
//   * warp 0 issues one cp.async.bulk per row (800 B at D=200; edge ids prefetched one window ahead, so the issue
//     never waits on a load) plus the window's fmask row and destination ids; completion by mbarrier expect_tx.
//     Round-2 profile of the per-thread cp.async version: 15 % of the instructions and 29 % of the stall samples sat
//     in the copy loop (address arithmetic, waiting for the edge ids);
//   * lazy BatchNorm + ReLU is applied in place by all threads, one float4 column group per thread (scale / shift
//     in registers), then one __syncthreads;
//   * warp w owns the dW rows f in [w*NF, (w+1)*NF) with the accumulators in REGISTERS (acc[u][t]: feature w*NF+u,
//     column k0 + lane + 32 t): lane u reads the 64-bit row set of its feature, the warp pops rows (ascending) into
//     a small shared list, fetches their g values in one parallel round and accumulates feature by feature.
// No shared-memory read-modify-write, no atomics, fixed order -> deterministic.
// Partials: part[cta][D][KW] (+ db in part_b[group][D]), folded in CTA order.
// ---------------------------------------------------------------------------------------
// Measured (round 2): the time of this kernel is (windows per CTA) x (a fixed ~7 us per window), whatever the
// instruction count -- 32 warps with two column slices (twice the windows per CTA) took 650 us against 413 us.
// The fixed cost was the exposed L2 latency of the g gather (one dependent round per pop batch, inside the
// per-window barrier).  The first batch of window b+1 is therefore popped and its g loads issued BEFORE window b
// is accumulated; the ring is 4 deep so that waiting for window b+1 early still leaves two windows in flight.
constexpr int kDwThreads = 512;
constexpr int kDwWarps = kDwThreads / 32;
constexpr int kDwStages = 4;
constexpr int kPop = 3;         // rows popped per feature and batch

template <int NF, int NT>
__global__ void __launch_bounds__(kDwThreads, 1) amax_bwd_dw_kernel(
    const float* __restrict__ g, const unsigned long long* __restrict__ fmask, mrg_act x,
    const int32_t* __restrict__ eid, const int32_t* __restrict__ csr_dst, int64_t E, int D, int KW, int kslices,
    float* __restrict__ part, float* __restrict__ part_b, long long* __restrict__ prof) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // optional phase timing (mrg_debug_set_dw_prof): cycles of warp 1 lane 0 per phase, summed over CTAs
  // (compiled in only with -DMRG_DW_PROF: the counters cost registers)
#ifdef MRG_DW_PROF
  long long t_aff = 0, t_bar = 0, t_pub = 0, t_wait = 0, t_acc = 0, t_mark = 0;
  const bool do_prof = prof != nullptr && threadIdx.x == 32;
#define DW_MARK(var) do { if (do_prof) { const long long now_ = clock64(); var += now_ - t_mark; t_mark = now_; } } while (0)
#else
  (void)prof;
#define DW_MARK(var) do { } while (0)
#endif
  // layout: xs[kDwStages][kWin][KW] f32 | fm[kDwStages][D] u64 | wd[kDwStages][kWin] i32 | full[kDwStages] u64 |
  //         lr[warps][NF][kPop] i32 | lg[warps][NF][kPop] f32 | sc[KW] | sh[KW]
  float* xs = reinterpret_cast<float*>(smem_raw);
  unsigned long long* fm = reinterpret_cast<unsigned long long*>(xs + (size_t)kDwStages * kWin * KW);
  int32_t* wd = reinterpret_cast<int32_t*>(fm + (size_t)kDwStages * D);
  uint64_t* full = reinterpret_cast<uint64_t*>(wd + kDwStages * kWin);
  int32_t* lr = reinterpret_cast<int32_t*>(full + kDwStages);
  float* lg = reinterpret_cast<float*>(lr + kDwWarps * NF * kPop);
  float* sc_s = lg + kDwWarps * NF * kPop + ((kDwWarps * NF * kPop) & 1 ? 1 : 0) + 2;   // keep 16-byte alignment
  sc_s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sc_s) + 15) & ~(uintptr_t)15);
  float* sh_s = sc_s + KW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slice = blockIdx.x % kslices, group = blockIdx.x / kslices, ngroups = gridDim.x / kslices;
  const int k0 = slice * KW;
  const int kw = min(KW, D - k0);   // valid columns of this slice (multiple of 4)
  const int kw4 = kw >> 2;
  const bool affine = x.scale != nullptr, relu = x.relu != 0;
  const int64_t nwin = (E + kWin - 1) / kWin;
  const int64_t W0 = nwin * group / ngroups, W1 = nwin * (group + 1) / ngroups;
  const int nb = (int)(W1 - W0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kDwStages; ++s) pipe::mbar_init(full + s, 1);
    pipe::fence_barrier_init();
  }
  // in-place activation: thread t owns float4 column group t % 64 (when < kw4) of rows t / 64 + 8 i
  const int my_c4 = threadIdx.x & 63, my_r0 = threadIdx.x >> 6;
  if (affine)
    for (int c = threadIdx.x; c < kw; c += blockDim.x) {
      sc_s[c] = x.scale[k0 + c];
      sh_s[c] = x.shift[k0 + c];
    }

  float acc[NF][NT];
#pragma unroll
  for (int u = 0; u < NF; ++u)
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[u][t] = 0.f;
  float db_acc = 0.f;
  const int f_l = warp * NF + lane;                 // the feature whose row set this lane pops
  const bool f_ok = lane < NF && f_l < D;
  int32_t* lr_w = lr + warp * NF * kPop;
  float* lg_w = lg + warp * NF * kPop;
  __syncthreads();

  // ---- copy issue, spread over all warps (a single issuing warp arrived late at every window barrier and the
  // other 15 waited for it: 31 % of the stall samples): warp w copies rows w, w+16, w+32, w+48 of a window, one
  // cp.async.bulk per lane 0..3; their edge ids are prefetched one window ahead
  int32_t e_pf = 0;
  const int my_row = warp + kDwWarps * lane;          // lanes < kWin / kDwWarps
  auto prefetch_eids = [&](int b) {
    if (b < nb && lane < kWin / kDwWarps) {
      const int64_t p = (W0 + b) * kWin + my_row;
      e_pf = p < E ? __ldg(eid + p) : 0;
    }
  };
  auto issue = [&](int b) {     // every warp; the stage must have been released by a __syncthreads
    const int64_t w_lo = (W0 + b) * kWin;
    const int nrows = (int)min((int64_t)kWin, E - w_lo);
    const int s = b % kDwStages;
    float* st = xs + (size_t)s * kWin * KW;
    const uint32_t row_bytes = (uint32_t)kw * 4u;
    if (threadIdx.x == 0) {
      // (a row copy of another warp may complete before this expect_tx: the transaction count of an mbarrier is
      //  signed, the phase cannot complete before this arrival)
      pipe::mbar_expect_tx(full + s, (uint32_t)nrows * row_bytes + (uint32_t)D * 8u + (uint32_t)kWin * 4u);
      pipe::bulk_g2s(fm + (size_t)s * D, fmask + (size_t)(W0 + b) * D, (uint32_t)D * 8u, full + s);
      pipe::bulk_g2s(wd + s * kWin, csr_dst + w_lo, (uint32_t)kWin * 4u, full + s);   // csr_dst is padded by kWin entries
    }
    if (lane < kWin / kDwWarps && my_row < nrows)
      pipe::bulk_g2s(st + (size_t)my_row * KW, x.data + (size_t)e_pf * D + k0, row_bytes, full + s);
  };
  prefetch_eids(0);
  if (nb > 0) issue(0);
  prefetch_eids(1);
  if (nb > 1) issue(1);
  prefetch_eids(2);
  if (nb > 2) issue(2);
  prefetch_eids(3);

  // pop up to kPop rows (ascending) of this lane's feature from row set `m` of the window in stage `s` and issue
  // the loads of their g values; nothing is consumed here, so the L2 latency overlaps whatever follows
  unsigned long long mA = 0ull, mB = 0ull;
  int cntA = 0, cntB = 0;
  float gA[kPop], gB[kPop];
  int rA[kPop], rB[kPop];
  auto pop = [&](unsigned long long& m, int s, int& cnt, int* rq, float* gq) {
    const int32_t* wd_s = wd + s * kWin;
    cnt = 0;
#pragma unroll
    for (int q = 0; q < kPop; ++q) {
      rq[q] = 0;
      gq[q] = 0.f;
      if (m) {
        rq[q] = __ffsll((long long)m) - 1;
        m &= m - 1;
        gq[q] = __ldg(g + (size_t)wd_s[rq[q]] * D + f_l);
        ++cnt;
      }
    }
  };
  auto publish = [&](const int* rq, const float* gq) {     // lanes < NF: my batch -> the warp's shared list
    if (lane < NF) {
#pragma unroll
      for (int q = 0; q < kPop; ++q) {
        lr_w[lane * kPop + q] = rq[q] * KW;
        lg_w[lane * kPop + q] = gq[q];
        db_acc += gq[q];
      }
    }
    __syncwarp();
  };
  if (nb > 0) {
    pipe::mbar_wait(full + 0, 0u);
    mA = f_ok ? fm[f_l] : 0ull;
    pop(mA, 0, cntA, rA, gA);
  }

#ifdef MRG_DW_PROF
  if (do_prof) t_mark = clock64();
#endif
  for (int b = 0; b < nb; ++b) {
    const int64_t w_lo = (W0 + b) * kWin;
    const int nrows = (int)min((int64_t)kWin, E - w_lo);
    const int s = b % kDwStages;          // landed: waited for in the previous iteration / the prologue
    if ((affine || relu) && my_c4 < kw4) {
      float* col = xs + (size_t)s * kWin * KW + 4 * my_c4;
      float4 my_sc = make_float4(1.f, 1.f, 1.f, 1.f), my_sh = make_float4(0.f, 0.f, 0.f, 0.f);
      if (affine) {
        my_sc = *reinterpret_cast<const float4*>(sc_s + 4 * my_c4);
        my_sh = *reinterpret_cast<const float4*>(sh_s + 4 * my_c4);
      }
      for (int r = my_r0; r < nrows; r += kDwThreads / 64) {
        float4 v = *reinterpret_cast<float4*>(col + (size_t)r * KW);
        if (affine) {
          v.x = fmaf(my_sc.x, v.x, my_sh.x); v.y = fmaf(my_sc.y, v.y, my_sh.y);
          v.z = fmaf(my_sc.z, v.z, my_sh.z); v.w = fmaf(my_sc.w, v.w, my_sh.w);
        }
        if (relu) {
          v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        *reinterpret_cast<float4*>(col + (size_t)r * KW) = v;
      }
    }
    // generic-proxy writes of this window (and reads of window b-1) are ordered before the bulk copies that will
    // overwrite the ring: proxy fence by every thread, then the CTA barrier, then warp 0 issues window b+3
    DW_MARK(t_aff);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    DW_MARK(t_bar);
    if (b + 3 < nb) {
      issue(b + 3);              // into the stage of window b-1, which everyone has left
      prefetch_eids(b + 4);
    }
    const float* st_lane = xs + (size_t)s * kWin * KW + lane;
    publish(rA, gA);                                   // first batch of window b (its g loads were issued a window ago)
    DW_MARK(t_pub);
    if (b + 1 < nb) {                                  // first batch of window b+1: loads in flight during the FMAs
      const int s1 = (b + 1) % kDwStages;
      pipe::mbar_wait(full + s1, (uint32_t)((b + 1) / kDwStages) & 1u);
      DW_MARK(t_wait);
      mB = f_ok ? fm[(size_t)s1 * D + f_l] : 0ull;
      pop(mB, s1, cntB, rB, gB);
    }
    int cnt = cntA;
    for (;;) {
#pragma unroll
      for (int u = 0; u < NF; ++u) {
        const int cu = __shfl_sync(0xffffffffu, cnt, u);
        for (int q = 0; q < cu; ++q) {
          const float gu = lg_w[u * kPop + q];
          const float* row = st_lane + lr_w[u * kPop + q];
#pragma unroll
          for (int t = 0; t < NT; ++t) acc[u][t] = fmaf(gu, row[32 * t], acc[u][t]);
        }
      }
      __syncwarp();
      if (!__ballot_sync(0xffffffffu, mA != 0ull)) break;
      pop(mA, s, cnt, rA, gA);                         // features with more than kPop rows in this window (rare)
      publish(rA, gA);
    }
    mA = mB;
    cntA = cntB;
#pragma unroll
    for (int q = 0; q < kPop; ++q) {
      rA[q] = rB[q];
      gA[q] = gB[q];
    }
    DW_MARK(t_acc);
  }
#ifdef MRG_DW_PROF
  if (do_prof) {
    atomicAdd((unsigned long long*)prof + 0, (unsigned long long)t_aff);
    atomicAdd((unsigned long long*)prof + 1, (unsigned long long)t_bar);
    atomicAdd((unsigned long long*)prof + 2, (unsigned long long)t_pub);
    atomicAdd((unsigned long long*)prof + 3, (unsigned long long)t_wait);
    atomicAdd((unsigned long long*)prof + 4, (unsigned long long)t_acc);
    atomicAdd((unsigned long long*)prof + 5, (unsigned long long)nb);
  }
#endif
#undef DW_MARK
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
static inline size_t dw_smem(int D, int KW, int NF) {
  return (size_t)kDwStages * kWin * KW * 4 + (size_t)kDwStages * D * 8 + (size_t)kDwStages * kWin * 4 +
         (size_t)kDwStages * 8 + (size_t)kDwWarps * NF * kPop * 8 + 2 * (size_t)KW * 4 + 256;
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

static long long* g_dw_prof = nullptr;
/* debugging aid (not part of the data path): device buffer of 8 int64 that the dW kernel adds its per-phase cycle
 * counts to (in-place activation, barrier, list publish, wait for the next window, accumulate, windows) */
extern "C" int mrg_debug_set_dw_prof(long long* dev_buf) {
  g_dw_prof = dev_buf;
  return MRG_OK;
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
    const size_t smem = dw_smem(D, KW, NF);                                                                       \
    e = cudaFuncSetAttribute(amax_bwd_dw_kernel<NF, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return cuda_fail(e, "amax_bwd dw smem attr");                                           \
    amax_bwd_dw_kernel<NF, NT><<<grid, kDwThreads, smem, st>>>(g, fmask, x, csr_eid, csr_dst, E, D, KW, ks, part,  \
                                                               part_b, g_dw_prof);                                \
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
