// K6 (fused form): backward of the collapsed sparse gate (f_sparse_op_comp / f_sparse_op(_last),
// operations_lp.py:304-354,405-416) fused with the BatchNorm1d+ReLU backward on BOTH sides of it:
//
//   * the incoming gradient is read through the lazy BN backward of the gate's own output state k,
//       dz = ds_k * [a_k*y_k + b_k > 0],   dy = c2*dz + c0 + c1*y_k          (mrg_grad; coef from mrg_bn_bwd_finalize)
//     so the standalone bn_bwd_apply pass (read ds, read y, write dy) disappears;
//   * when this gate is the LAST consumer of its input state x, the kernel holds the final ds_x row in registers
//     next to the raw y_x it just loaded, so it also emits the BN-backward column sums of that state
//     (sum dz_x, sum dz_x*y_x): the standalone bn_bwd_reduce pass (read ds, read y) disappears as well.
//
// Data movement per row: dy, y_k, x, (xin), (old dx), (old dxin) in, dx (dxin) out -- streamed through a TMA
// bulk-copy ring (pipe.cuh), one row per consumer warp per tile; per-column constants live in shared memory.
// Deterministic: every consumer warp keeps fp32 column partials in registers for its ~rows/(148*16) rows; at the end
// the 16 warps' partials are summed in double in warp order, per-CTA partials go to fixed slots and the finalize
// kernels fold them in slot order.
#include "pipe.cuh"

namespace mrg {

struct GateBwdArgs {
  pipe::Streams st;
  int s_dy, s_yk, s_x, s_in, s_dxo, s_dio;  // stream slots (-1: absent)
  mrg_act yk, x, xin;
  const float* coef;
  const float* gate;
  const float* v1;
  const float* v2;
  const float* row_scale;
  float base_scale;
  float* dx;
  float* dxin;
  double* dparam;   // [grid][2D+1]
  double* xstats;   // [grid][2D] or null
  int64_t rows;
  int D, stages;
};

enum { K_KSC = 0, K_KSH, K_C0, K_C1, K_C2, K_XSC, K_XSH, K_ISC, K_ISH, K_V1, K_V2, K_NCONST };

template <int NV, bool HAS_IN, bool SAME>
__global__ void __launch_bounds__(pipe::kPipeThreads, 1) gate_bwd_pipe_kernel(const GateBwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int D = a.D, D4 = D >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool lazy = a.coef != nullptr, xst = a.xstats != nullptr;
  // xin read through the same view as x (the fused cell's "x is x_in" gates): one activation serves both
  const bool same_act = SAME && a.xin.scale == a.x.scale && a.xin.shift == a.x.shift && a.xin.relu == a.x.relu;
  float* cst = reinterpret_cast<float*>(smem_raw);                   // [K_NCONST][D]
  size_t off = (size_t)K_NCONST * D * 4;
  off = (off + 127) / 128 * 128;
  pipe::Ring ring;
  ring.tiles = reinterpret_cast<float*>(smem_raw + off);
  ring.stages = a.stages;
  ring.n_streams = a.st.n;
  ring.D = D;
  ring.full = reinterpret_cast<uint64_t*>(ring.tiles + (size_t)a.stages * a.st.n * pipe::kTileRows * D);
  ring.empty = ring.full + a.stages;

  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    cst[K_KSC * D + c] = (lazy && a.yk.scale) ? a.yk.scale[c] : 1.f;
    cst[K_KSH * D + c] = (lazy && a.yk.scale) ? a.yk.shift[c] : 0.f;
    cst[K_C0 * D + c] = lazy ? a.coef[c] : 0.f;
    cst[K_C1 * D + c] = lazy ? a.coef[D + c] : 0.f;
    cst[K_C2 * D + c] = lazy ? a.coef[2 * D + c] : 1.f;
    cst[K_XSC * D + c] = a.x.scale ? a.x.scale[c] : 1.f;
    cst[K_XSH * D + c] = a.x.scale ? a.x.shift[c] : 0.f;
    cst[K_ISC * D + c] = (HAS_IN && a.xin.scale) ? a.xin.scale[c] : 1.f;
    cst[K_ISH * D + c] = (HAS_IN && a.xin.scale) ? a.xin.shift[c] : 0.f;
    cst[K_V1 * D + c] = a.v1[c];
    cst[K_V2 * D + c] = HAS_IN ? a.v2[c] : 0.f;
  }
  ring.init();   // includes __syncthreads: constants visible

  // per-warp fp32 column partials: dv1 | dv2 (dparam) and sum dz | sum dz*y (statistics of x's state)
  float4 pa1[NV], pa2[NV], pb1[NV], pb2[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) pa1[v] = pa2[v] = pb1[v] = pb2[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  float dc_f = 0.f;

  if (warp == pipe::kConsumerWarps) {
    if (lane == 0) ring.produce(a.st, a.rows);
  } else {
    const bool k_relu = lazy && a.yk.relu != 0, x_relu = a.x.relu != 0, i_relu = HAS_IN && a.xin.relu != 0;
    const int64_t ntiles = (a.rows + pipe::kTileRows - 1) / pipe::kTileRows;
    int it = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int s = it % a.stages;
      const int64_t row = t * pipe::kTileRows + warp;
      float g = 0.f, rs = 1.f;
      if (row < a.rows) {   // issue the two per-row scalars before blocking on the tile
        g = __ldg(a.gate + row);
        if (a.row_scale) rs = __ldg(a.row_scale + row);
      }
      pipe::mbar_wait(ring.full + s, (it / a.stages) & 1);
      if (row < a.rows) {
        const float* t_dy = ring.tile(s, a.s_dy) + warp * D;
        const float* t_yk = lazy ? ring.tile(s, a.s_yk) + warp * D : nullptr;
        const float* t_x = ring.tile(s, a.s_x) + warp * D;
        const float* t_in = (HAS_IN && !SAME) ? ring.tile(s, a.s_in) + warp * D : nullptr;
        float4 dv[NV], xv[NV], iv[NV], xr[NV];
        float dot = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c4 = lane + 32 * v;
          dv[v] = xv[v] = iv[v] = xr[v] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c4 < D4) {
#define CST(K) (*reinterpret_cast<const float4*>(cst + (K) * D + 4 * c4))
            float4 d = *reinterpret_cast<const float4*>(t_dy + 4 * c4);
            if (lazy) {
              const float4 y = *reinterpret_cast<const float4*>(t_yk + 4 * c4);
              if (k_relu) {
                const float4 ks = CST(K_KSC), kh = CST(K_KSH);
                d.x = fmaf(ks.x, y.x, kh.x) > 0.f ? d.x : 0.f;
                d.y = fmaf(ks.y, y.y, kh.y) > 0.f ? d.y : 0.f;
                d.z = fmaf(ks.z, y.z, kh.z) > 0.f ? d.z : 0.f;
                d.w = fmaf(ks.w, y.w, kh.w) > 0.f ? d.w : 0.f;
              }
              const float4 c0 = CST(K_C0), c1 = CST(K_C1), c2 = CST(K_C2);
              d.x = fmaf(c2.x, d.x, fmaf(c1.x, y.x, c0.x));
              d.y = fmaf(c2.y, d.y, fmaf(c1.y, y.y, c0.y));
              d.z = fmaf(c2.z, d.z, fmaf(c1.z, y.z, c0.z));
              d.w = fmaf(c2.w, d.w, fmaf(c1.w, y.w, c0.w));
            }
            const float4 raw = *reinterpret_cast<const float4*>(t_x + 4 * c4);
            const float4 xs = CST(K_XSC), xh = CST(K_XSH);
            float4 x = make_float4(fmaf(xs.x, raw.x, xh.x), fmaf(xs.y, raw.y, xh.y), fmaf(xs.z, raw.z, xh.z),
                                   fmaf(xs.w, raw.w, xh.w));
            if (x_relu) {
              x.x = x.x > 0.f ? x.x : 0.f; x.y = x.y > 0.f ? x.y : 0.f;
              x.z = x.z > 0.f ? x.z : 0.f; x.w = x.w > 0.f ? x.w : 0.f;
            }
            if (HAS_IN) {
              if (same_act) {
                iv[v] = x;
              } else {
                const float4 rin = SAME ? raw : *reinterpret_cast<const float4*>(t_in + 4 * c4);
                const float4 is = CST(K_ISC), ih = CST(K_ISH);
                float4 q = make_float4(fmaf(is.x, rin.x, ih.x), fmaf(is.y, rin.y, ih.y), fmaf(is.z, rin.z, ih.z),
                                       fmaf(is.w, rin.w, ih.w));
                if (i_relu) {
                  q.x = q.x > 0.f ? q.x : 0.f; q.y = q.y > 0.f ? q.y : 0.f;
                  q.z = q.z > 0.f ? q.z : 0.f; q.w = q.w > 0.f ? q.w : 0.f;
                }
                iv[v] = q;
              }
            }
            dv[v] = d;
            xv[v] = x;
            xr[v] = raw;
            dot += d.x * x.x + d.y * x.y + d.z * x.z + d.w * x.w;
          }
        }
        dot = warp_sum(dot);
        const float sc = a.base_scale * rs;
        const float dt = sc * g * (1.f - g) * dot;
        const float sg = sc * g;
        if (lane == 0) dc_f += dt;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c4 = lane + 32 * v;
          if (c4 < D4) {
            const size_t goff = (size_t)row * D + 4 * c4;
            const float4 w1 = CST(K_V1), w2 = CST(K_V2);
            float4 o = make_float4(fmaf(sg, dv[v].x, dt * w1.x), fmaf(sg, dv[v].y, dt * w1.y),
                                   fmaf(sg, dv[v].z, dt * w1.z), fmaf(sg, dv[v].w, dt * w1.w));
            float4 oi = make_float4(dt * w2.x, dt * w2.y, dt * w2.z, dt * w2.w);
            if (HAS_IN && SAME) { o.x += oi.x; o.y += oi.y; o.z += oi.z; o.w += oi.w; }
            if (a.s_dxo >= 0) {
              const float4 p = *reinterpret_cast<const float4*>(ring.tile(s, a.s_dxo) + warp * D + 4 * c4);
              o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
            }
            if (a.dx) st_stream4(a.dx + goff, o);
            if (HAS_IN && !SAME && a.dxin) {
              if (a.s_dio >= 0) {
                const float4 p = *reinterpret_cast<const float4*>(ring.tile(s, a.s_dio) + warp * D + 4 * c4);
                oi.x += p.x; oi.y += p.y; oi.z += p.z; oi.w += p.w;
              }
              st_stream4(a.dxin + goff, oi);
            }
            pa1[v].x = fmaf(dt, xv[v].x, pa1[v].x); pa1[v].y = fmaf(dt, xv[v].y, pa1[v].y);
            pa1[v].z = fmaf(dt, xv[v].z, pa1[v].z); pa1[v].w = fmaf(dt, xv[v].w, pa1[v].w);
            if (HAS_IN) {
              pa2[v].x = fmaf(dt, iv[v].x, pa2[v].x); pa2[v].y = fmaf(dt, iv[v].y, pa2[v].y);
              pa2[v].z = fmaf(dt, iv[v].z, pa2[v].z); pa2[v].w = fmaf(dt, iv[v].w, pa2[v].w);
            }
            if (xst) {
              float4 z = o;
              if (x_relu) {
                z.x = xv[v].x > 0.f ? z.x : 0.f; z.y = xv[v].y > 0.f ? z.y : 0.f;
                z.z = xv[v].z > 0.f ? z.z : 0.f; z.w = xv[v].w > 0.f ? z.w : 0.f;
              }
              pb1[v].x += z.x; pb1[v].y += z.y; pb1[v].z += z.z; pb1[v].w += z.w;
              pb2[v].x = fmaf(z.x, xr[v].x, pb2[v].x); pb2[v].y = fmaf(z.y, xr[v].y, pb2[v].y);
              pb2[v].z = fmaf(z.z, xr[v].z, pb2[v].z); pb2[v].w = fmaf(z.w, xr[v].w, pb2[v].w);
            }
#undef CST
          }
        }
      }
      __syncwarp();
      if (lane == 0) pipe::mbar_arrive(ring.empty + s);
    }
  }
  // ---- fold the 16 warps' partials in warp order (the ring memory is free: every tile was consumed)
  __syncthreads();
  float* red = ring.tiles;                        // [kConsumerWarps][4*D + 4] floats (rows stay 16-byte aligned)
  const int stride = 4 * D + 4;
  if (warp < pipe::kConsumerWarps) {
    float* r = red + (size_t)warp * stride;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c4 = lane + 32 * v;
      if (c4 < D4) {
        *reinterpret_cast<float4*>(r + 4 * c4) = pa1[v];
        *reinterpret_cast<float4*>(r + D + 4 * c4) = pa2[v];
        *reinterpret_cast<float4*>(r + 2 * D + 4 * c4) = pb1[v];
        *reinterpret_cast<float4*>(r + 3 * D + 4 * c4) = pb2[v];
      }
    }
    if (lane == 0) r[4 * D] = dc_f;
  }
  __syncthreads();
  double* part = a.dparam + (size_t)blockIdx.x * (2 * D + 1);
  for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) {
    double t = 0.0, u = 0.0;
#pragma unroll 4
    for (int w = 0; w < pipe::kConsumerWarps; ++w) {
      t += (double)red[(size_t)w * stride + c];
      u += (double)red[(size_t)w * stride + 2 * D + c];
    }
    part[c] = t;
    if (xst) a.xstats[(size_t)blockIdx.x * 2 * D + c] = u;
  }
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < pipe::kConsumerWarps; ++w) t += (double)red[(size_t)w * stride + 4 * D];
    part[2 * D] = t;
  }
}

}  // namespace mrg

using namespace mrg;

extern "C" int mrg_sparse_gate_bwd_fused_supported(int32_t D) { return valid_D(D) && D <= 256; }

extern "C" int mrg_sparse_gate_bwd_fused(mrg_grad dy, mrg_act x, mrg_act xin, const float* gate, int64_t rows,
                                         int32_t D, const float* v1, const float* v2, const float* row_scale,
                                         float base_scale, float* dx, float* dxin, int32_t accumulate, double* dparam,
                                         double* x_bwd_stats, void* stream) {
  MRG_CHECK_ARG(dy.ds && x.data && gate && v1 && dparam, "sparse_gate_bwd_fused: null pointer");
  MRG_CHECK_ARG(mrg_sparse_gate_bwd_fused_supported(D), "sparse_gate_bwd_fused: D must be a multiple of 4 and <= 256");
  MRG_CHECK_ARG(!dy.coef || dy.y.data, "sparse_gate_bwd_fused: lazy BN gradient needs y");
  MRG_CHECK_ARG(!xin.data || v2, "sparse_gate_bwd_fused: xin needs v2");
  MRG_CHECK_ARG(!(accumulate & 1) || dx, "sparse_gate_bwd_fused: accumulate without dx");
  cudaStream_t st = (cudaStream_t)stream;
  const int want = stats_grid(rows);
  const bool has_in = xin.data != nullptr;
  const bool same = has_in && xin.data == x.data && dxin == dx;
  GateBwdArgs a;
  a.st.n = 0;
  auto add = [&](const float* p) { a.st.base[a.st.n] = p; return a.st.n++; };
  a.s_dy = add(dy.ds);
  a.s_yk = dy.coef ? add(dy.y.data) : -1;
  a.s_x = add(x.data);
  a.s_in = (has_in && !same) ? add(xin.data) : -1;
  a.s_dxo = (accumulate & 1) ? add(dx) : -1;
  a.s_dio = (has_in && !same && dxin && (accumulate & 2)) ? add(dxin) : -1;
  a.yk = dy.y; a.x = x; a.xin = xin; a.coef = dy.coef; a.gate = gate; a.v1 = v1; a.v2 = v2;
  a.row_scale = row_scale; a.base_scale = base_scale; a.dx = dx; a.dxin = dxin; a.dparam = dparam;
  a.xstats = x_bwd_stats; a.rows = rows; a.D = D;
  const size_t fixed = (size_t)K_NCONST * D * 4 + 256;
  const size_t budget = 220 * 1024;
  int stages = 8;
  while (stages > 2 && fixed + pipe::ring_bytes(stages, a.st.n, D) > budget) --stages;
  // the final fold reuses the ring memory: kConsumerWarps * (4D+4) floats <= 2 streams x 2 stages x 16 rows x D floats
  if (pipe::ring_bytes(stages, a.st.n, D) < (size_t)pipe::kConsumerWarps * (4 * D + 4) * 4) {
    set_error("sparse_gate_bwd_fused: ring smaller than the fold scratch");
    return MRG_ERR_INVALID;
  }
  a.stages = stages;
  const size_t smem = fixed + pipe::ring_bytes(stages, a.st.n, D);
  if (smem > budget) {
    set_error("sparse_gate_bwd_fused: shared memory budget exceeded");
    return MRG_ERR_INVALID;
  }
  const int64_t ntiles = (rows + pipe::kTileRows - 1) / pipe::kTileRows;
  int grid = (int)(ntiles < kNumSMs ? (ntiles > 0 ? ntiles : 1) : kNumSMs);
  if (grid > want) grid = want;
  cudaError_t e;
#define L(HI, SM_)                                                                                                  \
  do {                                                                                                              \
    e = cudaFuncSetAttribute(gate_bwd_pipe_kernel<NV, HI, SM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return cuda_fail(e, "sparse_gate_bwd_fused smem attr");                                   \
    gate_bwd_pipe_kernel<NV, HI, SM_><<<grid, pipe::kPipeThreads, smem, st>>>(a);                                   \
  } while (0)
  if (D <= 128) { constexpr int NV = 1; if (!has_in) L(false, false); else if (same) L(true, true); else L(true, false); }
  else { constexpr int NV = 2; if (!has_in) L(false, false); else if (same) L(true, true); else L(true, false); }
#undef L
  zero_unwritten_parts(dparam, grid, want, 2 * (size_t)D + 1, st);
  if (x_bwd_stats) zero_unwritten_parts(x_bwd_stats, grid, want, 2 * (size_t)D, st);
  MRG_LAUNCH_CHECK("sparse_gate_bwd_fused");
  return MRG_OK;
}
