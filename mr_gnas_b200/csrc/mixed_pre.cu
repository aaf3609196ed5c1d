// K9b  DARTS MixedOp over the PRE candidate list with ONE shared read of its inputs (cell_lp.py:25-33 over
// PRE_OPS = pre_mult / pre_sub / pre_add, operations_lp.py:71-98):
//
//   out = sum_k w_k * relu( BN_k( comp_k(a, b) ) ),      comp in { a - b, a * b, a + b }
//
// The per-candidate form (round 1/2: K compositions written to HBM, K column-statistics passes, one K-way
// mixed-sum pass; backward K x (BN reduce + BN apply + composition backward) + the gradient adds of autograd)
// moves (2 + 1) K + K + 1 = 13 row passes forward and ~34 backward for K = 3.  Every candidate is an elementwise
// function of the SAME two rows, so here no candidate output ever exists:
//   forward   pass 1  read a, b            -> column sums / sums of squares of all K compositions
//             (K x mrg_bn_finalize: batch statistics -> BN affine per candidate)
//             pass 2  read a, b, write out -> recompute, normalise, ReLU, alpha-weighted sum
//   backward  pass 1  read dout, a, b      -> BN-backward column sums of all K candidates
//             (K x mrg_bn_bwd_finalize + mrg_mixed_bwd_scale: coefficients, dgamma, dbeta, dalpha)
//             pass 2  read dout, a, b, write da, db -> gradients of both inputs, summed over the candidates
// 5 row passes forward, 8 backward.  Statistics use the same fp32-per-lane / double-per-warp / fixed-order fold
// as every other producer (ColStats), so the result does not depend on the launch.
#include "common.cuh"

namespace mrg {

template <int COMP>
__device__ __forceinline__ float4 comp4(float4 a, float4 b) {
  if (COMP == MRG_COMP_SUB) return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
  if (COMP == MRG_COMP_MULT) return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 comp4(int comp, float4 a, float4 b) {
  return comp == MRG_COMP_SUB ? comp4<MRG_COMP_SUB>(a, b) : (comp == MRG_COMP_MULT ? comp4<MRG_COMP_MULT>(a, b) : comp4<MRG_COMP_ADD>(a, b));
}

struct MixedPreArgs {
  const float* a;          // [rows, D]  first operand rows  (the gathered entity rows)
  const float* b;          // [rows, D]  second operand rows (the gathered relation rows)
  int64_t rows;
  int D, K;
  int comp[MRG_MAX_MIXED];
  const float* scale;      // [K, D] BN affine a_k = gamma * invstd
  const float* shift;      // [K, D]           b_k = beta - a_k * mean
  const float* w;          // [K] softmax(alpha) row
  const float* coef;       // [K, 3, D] backward coefficients (c0, c1, c2), already scaled by w_k
};

// forward pass 1 / backward pass 1.  BWD = false: sums of v_k and v_k^2.  BWD = true: sums of dz_k and dz_k * v_k with
// dz_k = dout * [a_k v_k + b_k > 0].  stats layout [K][nparts][2][D] doubles (one mrg_bn_finalize per candidate).
template <int NV, int KC, bool BWD>
__global__ void __launch_bounds__(kThreads) mixed_pre_stats_kernel(const MixedPreArgs p, const float* __restrict__ dout,
                                                                   double* __restrict__ stats, int nparts) {
  extern __shared__ double smem_d[];
  const int lane = threadIdx.x & 31;
  const int D = p.D, D4 = D >> 2;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  ColStats<NV> cs[KC];
#pragma unroll
  for (int k = 0; k < KC; ++k) cs[k].init(smem_d + (size_t)k * stats_smem_doubles(D), D, D4);
  RowBuf<NV> na, nb, ng;
  if (warp0 < p.rows) {
    na.load(p.a, warp0, D, D4, lane);
    nb.load(p.b, warp0, D, D4, lane);
    if (BWD) ng.load(dout, warp0, D, D4, lane);
  }
  for (int64_t row = warp0; row < p.rows; row += nwarps) {
    const RowBuf<NV> ca = na, cb = nb, cg = ng;
    if (row + nwarps < p.rows) {
      na.load(p.a, row + nwarps, D, D4, lane);
      nb.load(p.b, row + nwarps, D, D4, lane);
      if (BWD) ng.load(dout, row + nwarps, D, D4, lane);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c4 = lane + 32 * v;
      if (c4 < D4) {
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          const float4 x = comp4(p.comp[k], ca.v[v], cb.v[v]);
          if (!BWD) {
            cs[k].add_sq(x, v);
          } else {
            const float4 sc = ldg4(p.scale + (size_t)k * D + 4 * c4), sh = ldg4(p.shift + (size_t)k * D + 4 * c4);
            float4 g = cg.v[v];
            g.x = fmaf(sc.x, x.x, sh.x) > 0.f ? g.x : 0.f;
            g.y = fmaf(sc.y, x.y, sh.y) > 0.f ? g.y : 0.f;
            g.z = fmaf(sc.z, x.z, sh.z) > 0.f ? g.z : 0.f;
            g.w = fmaf(sc.w, x.w, sh.w) > 0.f ? g.w : 0.f;
            cs[k].add(g, make_float4(g.x * x.x, g.y * x.y, g.z * x.z, g.w * x.w), v);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KC; ++k) cs[k].row_done(D, D4);
  }
#pragma unroll
  for (int k = 0; k < KC; ++k)
    cs[k].write_block(stats + ((size_t)k * nparts + blockIdx.x) * 2 * D, smem_d + (size_t)k * stats_smem_doubles(D), D, D4);
}

// forward pass 2: out = sum_k w_k relu(a_k comp_k(a, b) + b_k), Python's sum() association order
template <int NV>
__global__ void __launch_bounds__(kThreads) mixed_pre_fwd_kernel(const MixedPreArgs p, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int D = p.D, D4 = D >> 2;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  float wk[MRG_MAX_MIXED];
  for (int k = 0; k < p.K; ++k) wk[k] = __ldg(p.w + k);
  RowBuf<NV> na, nb;
  if (warp0 < p.rows) {
    na.load(p.a, warp0, D, D4, lane);
    nb.load(p.b, warp0, D, D4, lane);
  }
  for (int64_t row = warp0; row < p.rows; row += nwarps) {
    const RowBuf<NV> ca = na, cb = nb;
    if (row + nwarps < p.rows) {
      na.load(p.a, row + nwarps, D, D4, lane);
      nb.load(p.b, row + nwarps, D, D4, lane);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c4 = lane + 32 * v;
      if (c4 < D4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < p.K; ++k) {
          float4 x = comp4(p.comp[k], ca.v[v], cb.v[v]);
          const float4 sc = ldg4(p.scale + (size_t)k * D + 4 * c4), sh = ldg4(p.shift + (size_t)k * D + 4 * c4);
          x.x = fmaxf(fmaf(sc.x, x.x, sh.x), 0.f);
          x.y = fmaxf(fmaf(sc.y, x.y, sh.y), 0.f);
          x.z = fmaxf(fmaf(sc.z, x.z, sh.z), 0.f);
          x.w = fmaxf(fmaf(sc.w, x.w, sh.w), 0.f);
          acc.x += wk[k] * x.x; acc.y += wk[k] * x.y; acc.z += wk[k] * x.z; acc.w += wk[k] * x.w;
        }
        st_stream4(out + (size_t)row * D + 4 * c4, acc);
      }
    }
  }
}

// backward pass 2: dv_k = c2_k dout [s_k > 0] + c1_k v_k + c0_k  (the BatchNorm + ReLU + alpha backward of candidate k,
// coefficients from mrg_bn_bwd_finalize / mrg_mixed_bwd_scale), chained through the composition:
//   sub: da += dv, db -= dv      mult: da += dv * b, db += dv * a      add: da += dv, db += dv
template <int NV>
__global__ void __launch_bounds__(kThreads) mixed_pre_bwd_kernel(const MixedPreArgs p, const float* __restrict__ dout,
                                                                 float* __restrict__ da, float* __restrict__ db) {
  const int lane = threadIdx.x & 31;
  const int D = p.D, D4 = D >> 2;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  RowBuf<NV> na, nb, ng;
  if (warp0 < p.rows) {
    na.load(p.a, warp0, D, D4, lane);
    nb.load(p.b, warp0, D, D4, lane);
    ng.load(dout, warp0, D, D4, lane);
  }
  for (int64_t row = warp0; row < p.rows; row += nwarps) {
    const RowBuf<NV> ca = na, cb = nb, cg = ng;
    if (row + nwarps < p.rows) {
      na.load(p.a, row + nwarps, D, D4, lane);
      nb.load(p.b, row + nwarps, D, D4, lane);
      ng.load(dout, row + nwarps, D, D4, lane);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c4 = lane + 32 * v;
      if (c4 < D4) {
        const float4 a = ca.v[v], b = cb.v[v], g = cg.v[v];
        float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga;
        for (int k = 0; k < p.K; ++k) {
          const float4 x = comp4(p.comp[k], a, b);
          const float4 sc = ldg4(p.scale + (size_t)k * D + 4 * c4), sh = ldg4(p.shift + (size_t)k * D + 4 * c4);
          const float* cf = p.coef + (size_t)k * 3 * D + 4 * c4;
          const float4 c0 = ldg4(cf), c1 = ldg4(cf + D), c2 = ldg4(cf + 2 * D);
          float4 dv;
          dv.x = fmaf(c2.x, fmaf(sc.x, x.x, sh.x) > 0.f ? g.x : 0.f, fmaf(c1.x, x.x, c0.x));
          dv.y = fmaf(c2.y, fmaf(sc.y, x.y, sh.y) > 0.f ? g.y : 0.f, fmaf(c1.y, x.y, c0.y));
          dv.z = fmaf(c2.z, fmaf(sc.z, x.z, sh.z) > 0.f ? g.z : 0.f, fmaf(c1.z, x.z, c0.z));
          dv.w = fmaf(c2.w, fmaf(sc.w, x.w, sh.w) > 0.f ? g.w : 0.f, fmaf(c1.w, x.w, c0.w));
          if (p.comp[k] == MRG_COMP_MULT) {
            ga.x = fmaf(dv.x, b.x, ga.x); ga.y = fmaf(dv.y, b.y, ga.y); ga.z = fmaf(dv.z, b.z, ga.z); ga.w = fmaf(dv.w, b.w, ga.w);
            gb.x = fmaf(dv.x, a.x, gb.x); gb.y = fmaf(dv.y, a.y, gb.y); gb.z = fmaf(dv.z, a.z, gb.z); gb.w = fmaf(dv.w, a.w, gb.w);
          } else {
            const float sgn = p.comp[k] == MRG_COMP_SUB ? -1.f : 1.f;
            ga.x += dv.x; ga.y += dv.y; ga.z += dv.z; ga.w += dv.w;
            gb.x = fmaf(sgn, dv.x, gb.x); gb.y = fmaf(sgn, dv.y, gb.y); gb.z = fmaf(sgn, dv.z, gb.z); gb.w = fmaf(sgn, dv.w, gb.w);
          }
        }
        const size_t off = (size_t)row * D + 4 * c4;
        if (da) st_stream4(da + off, ga);
        if (db) st_stream4(db + off, gb);
      }
    }
  }
}

}  // namespace mrg

using namespace mrg;

static int fill_args(MixedPreArgs& p, const float* a, const float* b, int64_t rows, int32_t D, const int32_t* comps,
                     int32_t K, const float* scale, const float* shift, const float* w, const float* coef) {
  MRG_CHECK_ARG(a && b && comps, "mixed_pre: null pointer");
  MRG_CHECK_ARG(valid_D(D), "mixed_pre: D must be a multiple of 4 and <= 512");
  MRG_CHECK_ARG(K >= 1 && K <= MRG_MAX_MIXED, "mixed_pre: 1 <= K <= MRG_MAX_MIXED");
  p.a = a;
  p.b = b;
  p.rows = rows;
  p.D = D;
  p.K = K;
  for (int k = 0; k < K; ++k) {
    MRG_CHECK_ARG(comps[k] == MRG_COMP_SUB || comps[k] == MRG_COMP_MULT || comps[k] == MRG_COMP_ADD, "mixed_pre: bad composition");
    p.comp[k] = comps[k];
  }
  p.scale = scale;
  p.shift = shift;
  p.w = w;
  p.coef = coef;
  return MRG_OK;
}

template <bool BWD>
static int launch_stats(const MixedPreArgs& p, const float* dout, double* stats, cudaStream_t st) {
  if (p.rows <= 0) return MRG_OK;
  const size_t smem = (size_t)p.K * stats_smem_doubles(p.D) * sizeof(double);
  const int nparts = stats_grid(p.rows);      // partial blocks per candidate, as every other statistics producer
#define MRG_MP_STATS(KC)                                                                                      \
  MRG_DISPATCH_NV(p.D, {                                                                                      \
    auto kern = mixed_pre_stats_kernel<NV, KC, BWD>;                                                          \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    const int grid = resident_grid(kern, smem, nparts);                                                       \
    for (int k = 0; k < KC; ++k) zero_unwritten_parts(stats + (size_t)k * nparts * 2 * p.D, grid, nparts, 2 * (size_t)p.D, st); \
    kern<<<grid, kThreads, smem, st>>>(p, dout, stats, nparts);                                               \
  })
  switch (p.K) {
    case 1: MRG_MP_STATS(1); break;
    case 2: MRG_MP_STATS(2); break;
    case 3: MRG_MP_STATS(3); break;
    default: set_error("mixed_pre: statistics pass supports K <= 3 candidates"); return MRG_ERR_INVALID;
  }
#undef MRG_MP_STATS
  MRG_LAUNCH_CHECK("mixed_pre_stats");
  return MRG_OK;
}

extern "C" int mrg_mixed_pre_stats(const float* a, const float* b, int64_t rows, int32_t D, const int32_t* comps, int32_t K,
                                   double* stats, void* stream) {
  MixedPreArgs p;
  int rc = fill_args(p, a, b, rows, D, comps, K, nullptr, nullptr, nullptr, nullptr);
  if (rc != MRG_OK) return rc;
  MRG_CHECK_ARG(stats, "mixed_pre_stats: null stats");
  return launch_stats<false>(p, nullptr, stats, (cudaStream_t)stream);
}

extern "C" int mrg_mixed_pre_fwd(const float* a, const float* b, int64_t rows, int32_t D, const int32_t* comps, int32_t K,
                                 const float* scale, const float* shift, const float* w, float* out, void* stream) {
  MixedPreArgs p;
  int rc = fill_args(p, a, b, rows, D, comps, K, scale, shift, w, nullptr);
  if (rc != MRG_OK) return rc;
  MRG_CHECK_ARG(scale && shift && w && out, "mixed_pre_fwd: null pointer");
  if (rows <= 0) return MRG_OK;
  MRG_DISPATCH_NV(D, {
    const int grid = resident_grid(mixed_pre_fwd_kernel<NV>, 0, (int)((rows + kWarpsPerBlock - 1) / kWarpsPerBlock < kMaxParts
                                                                          ? (rows + kWarpsPerBlock - 1) / kWarpsPerBlock
                                                                          : kMaxParts));
    mixed_pre_fwd_kernel<NV><<<grid, kThreads, 0, (cudaStream_t)stream>>>(p, out);
  });
  MRG_LAUNCH_CHECK("mixed_pre_fwd");
  return MRG_OK;
}

extern "C" int mrg_mixed_pre_bwd_stats(const float* dout, const float* a, const float* b, int64_t rows, int32_t D,
                                       const int32_t* comps, int32_t K, const float* scale, const float* shift,
                                       double* bwd_stats, void* stream) {
  MixedPreArgs p;
  int rc = fill_args(p, a, b, rows, D, comps, K, scale, shift, nullptr, nullptr);
  if (rc != MRG_OK) return rc;
  MRG_CHECK_ARG(dout && scale && shift && bwd_stats, "mixed_pre_bwd_stats: null pointer");
  return launch_stats<true>(p, dout, bwd_stats, (cudaStream_t)stream);
}

extern "C" int mrg_mixed_pre_bwd(const float* dout, const float* a, const float* b, int64_t rows, int32_t D,
                                 const int32_t* comps, int32_t K, const float* scale, const float* shift,
                                 const float* coef, float* da, float* db, void* stream) {
  MixedPreArgs p;
  int rc = fill_args(p, a, b, rows, D, comps, K, scale, shift, nullptr, coef);
  if (rc != MRG_OK) return rc;
  MRG_CHECK_ARG(dout && scale && shift && coef, "mixed_pre_bwd: null pointer");
  if (rows <= 0) return MRG_OK;
  MRG_DISPATCH_NV(D, {
    const int grid = resident_grid(mixed_pre_bwd_kernel<NV>, 0, (int)((rows + kWarpsPerBlock - 1) / kWarpsPerBlock < kMaxParts
                                                                          ? (rows + kWarpsPerBlock - 1) / kWarpsPerBlock
                                                                          : kMaxParts));
    mixed_pre_bwd_kernel<NV><<<grid, kThreads, 0, (cudaStream_t)stream>>>(p, dout, da, db);
  });
  MRG_LAUNCH_CHECK("mixed_pre_bwd");
  return MRG_OK;
}
