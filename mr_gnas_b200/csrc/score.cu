// 1-N scoring with the TransE scorer (reference: models/operations_lp.py:101-112, sf_TransE_op):
//   x[b, n] = gamma - || (sub_emb[b] + rel_emb[b]) - all_ent[n] ||_1 ,   score = sigmoid(x)
// The reference broadcasts obj_emb.unsqueeze(1) - all_ent to a [B, N, D] tensor (2.98 GB at C1) before the
// L1 norm; here the distance is a register-tiled "GEMM with |a - b| instead of a * b" and its backward two
// sign-weighted reductions, nothing of size B*N*D ever exists.
//
//   forward : x[b,n]    = gamma - sum_k |q[b,k] - e[n,k]|
//   backward: dq[b,k]   = - sum_n dx[b,n] * sgn(q[b,k] - e[n,k])
//             de[n,k]   = + sum_b dx[b,n] * sgn(q[b,k] - e[n,k])          (sgn(0) = 0, as torch.sign)
// Deterministic: fixed tile ownership, split reductions folded in split order.
#include "common.cuh"

namespace mrg {

constexpr int kST = 64;        // tile edge (rows of either operand)
constexpr int kSK = 32;        // reduction chunk staged in shared memory
constexpr int kSThreads = 256; // 16 x 16 threads, 4 x 4 outputs each

// ---- forward: grid (ceil(N/64), ceil(B/64)) ----------------------------------------------------------------
__global__ void __launch_bounds__(kSThreads) transe_fwd_kernel(const float* __restrict__ q, const float* __restrict__ e,
                                                              int B, int64_t N, int D, float gamma,
                                                              float* __restrict__ x) {
  __shared__ __align__(16) float qs[kSK][kST + 4];   // [k][b]
  __shared__ __align__(16) float es[kSK][kST + 4];   // [k][n]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t n0 = (int64_t)blockIdx.x * kST;
  const int b0 = blockIdx.y * kST;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < D; k0 += kSK) {
    // stage: thread t loads row (t / 4) [+ 64 rows per operand], 8 consecutive k -- two float4, transposed into [k][row]
    {
      const int r = threadIdx.x >> 2, kq = (threadIdx.x & 3) * 8;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = k0 + kq + 4 * h;
        float4 vq = make_float4(0.f, 0.f, 0.f, 0.f), ve = vq;
        if (k < D) {      // D % 4 == 0
          if (b0 + r < B) vq = ldg4(q + (size_t)(b0 + r) * D + k);
          if (n0 + r < N) ve = ldg4(e + (size_t)(n0 + r) * D + k);
        }
        qs[kq + 4 * h + 0][r] = vq.x; qs[kq + 4 * h + 1][r] = vq.y; qs[kq + 4 * h + 2][r] = vq.z; qs[kq + 4 * h + 3][r] = vq.w;
        es[kq + 4 * h + 0][r] = ve.x; es[kq + 4 * h + 1][r] = ve.y; es[kq + 4 * h + 2][r] = ve.z; es[kq + 4 * h + 3][r] = ve.w;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kSK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&qs[k][4 * ty]);
      const float4 c = *reinterpret_cast<const float4*>(&es[k][4 * tx]);
      const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += fabsf(av[i] - cv[j]);   // zero-padded k contributes |0 - 0| = 0
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + 4 * ty + i;
    if (b >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + 4 * tx + j;
      if (n < N) x[(size_t)b * N + n] = gamma - acc[i][j];
    }
  }
}

// ---- backward: out[r, k] = sign * sum_{c in split} dxv(r, c) * sgn(A[r,k] - Bm[c,k]) ------------------------
// DXT = false: dxv(r, c) = dx[r * ld + c]   (r = query, c = entity  -> dq, sign = -1)
// DXT = true : dxv(r, c) = dx[c * ld + r]   (r = entity, c = query  -> de, sign = -1 with A = e, Bm = q)
// grid (ceil(R/64), ceil(D/64), splits); partial results to part[split][R][D] (or straight to out when splits == 1)
template <bool DXT>
__global__ void __launch_bounds__(kSThreads) transe_bwd_kernel(const float* __restrict__ dx, int64_t ld,
                                                              const float* __restrict__ A, const float* __restrict__ Bm,
                                                              int64_t R, int64_t C, int D, float sign,
                                                              float* __restrict__ part) {
  __shared__ __align__(16) float ds[kSK][kST + 4];   // [c][r]
  __shared__ __align__(16) float bs[kSK][kST + 4];   // [c][k]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t r0 = (int64_t)blockIdx.x * kST;
  const int kk0 = blockIdx.y * kST;
  const int splits = gridDim.z;
  const int64_t cper = ((C + splits - 1) / splits + kSK - 1) / kSK * kSK;
  const int64_t c_lo = (int64_t)blockIdx.z * cper, c_hi = min(C, c_lo + cper);
  float av[4][4], acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t r = r0 + 4 * ty + i;
      const int k = kk0 + 4 * tx + j;
      av[i][j] = (r < R && k < D) ? __ldg(A + (size_t)r * D + k) : 0.f;
      acc[i][j] = 0.f;
    }
  for (int64_t c0 = c_lo; c0 < c_hi; c0 += kSK) {
    // stage dx tile [32 c][64 r] and Bm tile [32 c][64 k]
    for (int t = threadIdx.x; t < kSK * kST; t += kSThreads) {
      int cc, rr;
      if (DXT) { cc = t / kST; rr = t % kST; }     // dx[c * ld + r]: r contiguous
      else     { rr = t / kSK; cc = t % kSK; }     // dx[r * ld + c]: c contiguous
      const int64_t c = c0 + cc, r = r0 + rr;
      float v = 0.f;
      if (c < c_hi && r < R) v = __ldg(dx + (DXT ? (size_t)c * ld + r : (size_t)r * ld + c));
      ds[cc][rr] = v;
    }
    for (int t = threadIdx.x; t < kSK * kST; t += kSThreads) {
      const int cc = t / kST, k = t % kST;
      const int64_t c = c0 + cc;
      bs[cc][k] = (c < c_hi && kk0 + k < D) ? __ldg(Bm + (size_t)c * D + kk0 + k) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int cc = 0; cc < kSK; ++cc) {
      const float4 d4 = *reinterpret_cast<const float4*>(&ds[cc][4 * ty]);
      const float4 b4 = *reinterpret_cast<const float4*>(&bs[cc][4 * tx]);
      const float dv[4] = {d4.x, d4.y, d4.z, d4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float diff = av[i][j] - bv[j];
          const float s = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);   // padded c: dv == 0 -> contributes 0
          acc[i][j] = fmaf(dv[i], s, acc[i][j]);
        }
    }
    __syncthreads();
  }
  float* out = part + (size_t)blockIdx.z * R * D;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + 4 * ty + i;
    if (r >= R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kk0 + 4 * tx + j;
      if (k < D) out[(size_t)r * D + k] = sign * acc[i][j];
    }
  }
}

__global__ void fold_splits_kernel(const float* __restrict__ part, int splits, int64_t n, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float t = 0.f;
    for (int s = 0; s < splits; ++s) t += part[(size_t)s * n + i];
    out[i] = t;
  }
}

}  // namespace mrg

using namespace mrg;

static inline int transe_dq_splits(int64_t B, int64_t N, int D) {
  const int64_t tiles = ((B + kST - 1) / kST) * ((D + kST - 1) / kST);
  int64_t s = (4 * (int64_t)kNumSMs + tiles - 1) / tiles;          // ~4 CTAs per SM
  const int64_t smax = (N + 4 * kSK - 1) / (4 * kSK);              // at least 128 entities per split
  if (s > smax) s = smax;
  return (int)(s < 1 ? 1 : s);
}

extern "C" size_t mrg_transe_bwd_workspace_bytes(int64_t B, int64_t N, int32_t D) {
  return (size_t)transe_dq_splits(B, N, D) * B * D * sizeof(float) + 256;
}

extern "C" int mrg_transe_fwd(const float* query, const float* ent, int64_t B, int64_t N, int32_t D, float gamma,
                              float* logit, void* stream) {
  MRG_CHECK_ARG(query && ent && logit, "transe_fwd: null pointer");
  MRG_CHECK_ARG(B > 0 && N > 0 && D > 0 && D % 4 == 0 && B < (1 << 22), "transe_fwd: sizes (D % 4 == 0)");
  dim3 grid((unsigned)((N + kST - 1) / kST), (unsigned)((B + kST - 1) / kST));
  transe_fwd_kernel<<<grid, kSThreads, 0, (cudaStream_t)stream>>>(query, ent, (int)B, N, D, gamma, logit);
  MRG_LAUNCH_CHECK("transe_fwd");
  return MRG_OK;
}

extern "C" int mrg_transe_bwd(const float* dlogit, const float* query, const float* ent, int64_t B, int64_t N,
                              int32_t D, float* dquery, float* dent, void* workspace, size_t workspace_bytes,
                              void* stream) {
  MRG_CHECK_ARG(dlogit && query && ent, "transe_bwd: null pointer");
  MRG_CHECK_ARG(B > 0 && N > 0 && D > 0 && D % 4 == 0, "transe_bwd: sizes (D % 4 == 0)");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned ktiles = (unsigned)((D + kST - 1) / kST);
  if (dent) {   // every entity row sums over all B queries inside one CTA: no split
    dim3 grid((unsigned)((N + kST - 1) / kST), ktiles, 1);
    transe_bwd_kernel<true><<<grid, kSThreads, 0, st>>>(dlogit, N, ent, query, N, B, D, -1.f, dent);
  }
  if (dquery) {
    const int splits = transe_dq_splits(B, N, D);
    if (splits > 1 && (!workspace || workspace_bytes < mrg_transe_bwd_workspace_bytes(B, N, D))) {
      set_error("transe_bwd: workspace too small");
      return MRG_ERR_WORKSPACE;
    }
    dim3 grid((unsigned)((B + kST - 1) / kST), ktiles, (unsigned)splits);
    float* part = splits > 1 ? (float*)workspace : dquery;
    transe_bwd_kernel<false><<<grid, kSThreads, 0, st>>>(dlogit, N, query, ent, B, N, D, -1.f, part);
    if (splits > 1) {
      const int64_t n = B * D;
      fold_splits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, splits, n, dquery);
    }
  }
  MRG_LAUNCH_CHECK("transe_bwd");
  return MRG_OK;
}
