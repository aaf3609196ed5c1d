// K5 (tensor-core form): fused edge-tile GEMM + bias + ReLU + destination max aggregation
//   out[n,f] = max_{e in CSR row n} relu( sum_k W[f,k] * X[eid(e),k] + b[f] )  (+ residual[n,f])
// for a_max_op (operations_lp.py:230-235; operations.py:112-121) without ever materialising
// the [E,D] message matrix.
//
// tcgen05 mapping (sm_100a): the GEMM is computed TRANSPOSED, C^T[f, e] = W[f,:] . X[e,:],
// so the UMMA "M" axis (TMEM lanes) is the output feature and the "N" axis (TMEM columns)
// is the edge position inside a 128-edge tile of the dst-sorted CSR.  Every epilogue thread
// then owns ONE feature and walks its TMEM row across the tile's edges, so the segmented
// max over destinations is a register-only scan and the writes out[n, f0..f0+31] of a warp are
// coalesced.  Segments that straddle tiles are merged with a 64-bit atomicMax on the packed
// key (float bits of the non-negative value, ~edge id): order independent, hence deterministic,
// and the lowest edge id wins ties (the stated tie-break).
//
// fp32 accuracy on the TF32 pipe: both operands are split x = hi + lo (hi = rna_tf32(x),
// lo = rna_tf32(x - hi)) and three MMAs accumulate hi*hi + hi*lo + lo*hi in fp32 TMEM
// ("3xTF32"), giving ~2^-21 relative operand error, i.e. SGEMM-class results.
//
// Pipeline per CTA (persistent, one CTA per SM, 26 warps; tiles are walked in PAIRS -- TMEM slots 0/1 -- with
// their K chunks interleaved so each W chunk streamed from L2 feeds two tiles):
//   warps 0-15  X producers: 4 threads per tile row gather X by CSR edge id with a 4-chunk-deep register
//                            prefetch ring -> (BN affine + ReLU) -> hi/lo split -> 128B-swizzled K-major smem
//   warp  25    W loader   : one thread, cp.async.bulk (UBLKCP) of the pre-swizzled hi/lo K-chunk image,
//                            completion by mbarrier expect_tx
//   warp  24    MMA issuer : one elected thread, tcgen05.mma kind::tf32, M=128 per half, N=128, K=8
//   warps 16-23 epilogue   : one 4-warp set per TMEM slot: tcgen05.ld -> +bias, ReLU -> segmented max scan
//                            -> atomicMax
#include "common.cuh"

namespace mrg {
namespace tc {

constexpr int TILE_E = 128;   // edges per tile  (UMMA N)
constexpr int KCH = 32;       // fp32 elements per K chunk = one 128-byte swizzled row
constexpr int STAGES = 2;
constexpr int TPR = 4;         // producer threads per tile row
constexpr int NU = 8 / TPR;    // 16-byte units of a 128-byte K chunk handled by one producer thread
constexpr int PROD_THREADS = 128 * TPR;
constexpr int EPI_THREADS = 128;
constexpr int THREADS = PROD_THREADS + 2 * EPI_THREADS + 64;   // 16 producer + 8 epilogue + MMA + W loader warps
constexpr int PD = 4;         // X prefetch depth (items in flight per producer thread)
constexpr int EPI_WARP0 = PROD_THREADS / 32;   // multiple of 4 -> TMEM lane quadrants 0..3 for both warp sets
constexpr int MMA_WARP = EPI_WARP0 + 8;
constexpr int WLD_WARP = MMA_WARP + 1;
constexpr uint32_t TILE_BYTES = 128 * 128;  // 128 rows x 128 B

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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// K-major, 128B-swizzled smem matrix descriptor: start>>4 | LBO(=1)<<16 | SBO(=1024B>>4)<<32 |
// version 1 << 46 | SWIZZLE_128B (2) << 61     (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::tf32, fp32 accumulate, A and B K-major, M=128, N=128   (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// kind::f16 with bf16 operands, fp32 accumulate, A and B K-major
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Segmented running max of 32 consecutive TMEM columns (one feature per thread) with the lowest-column tie-break.
// `fl` (warp-uniform) flags the columns where a new destination starts.  Groups of 8 columns without a flag -- most
// of them: a destination has tens of in-edges -- take a branch-free path: a pairwise (value, column) tournament
// whose strict `later > earlier` comparisons keep the earliest column on ties, merged into the running maximum by
// one more strict comparison; 4 instructions per column with log-depth dependences instead of the ~11 of the
// flag-testing path (BSSY/BSYNC pairs, moves) that the round-2 SASS showed for every column.
template <class Close>
__device__ __forceinline__ void seg_scan32(const uint32_t (&v)[32], uint32_t fl, int cb, float bias, float& best,
                                           int& bcol, Close&& close_seg) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (((fl >> (8 * g)) & 0xffu) == 0u) {
      float m[8];
      int ix[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        m[j] = __uint_as_float(v[8 * g + j]) + bias;
        ix[j] = cb + 8 * g + j;
      }
#pragma unroll
      for (int w = 1; w < 8; w <<= 1) {
#pragma unroll
        for (int j = 0; j < 8; j += 2 * w) {
          const bool gt = m[j + w] > m[j];
          m[j] = gt ? m[j + w] : m[j];
          ix[j] = gt ? ix[j + w] : ix[j];
        }
      }
      const bool gt = m[0] > best;
      best = gt ? m[0] : best;
      bcol = gt ? ix[0] : bcol;
    } else {
#pragma unroll
      for (int j = 8 * g; j < 8 * g + 8; ++j) {
        if (fl & (1u << j)) {   // warp-uniform
          close_seg(cb + j);
          best = 0.f;
          bcol = cb + j;
        }
        const float val = __uint_as_float(v[j]) + bias;
        const bool gt = val > best;
        best = gt ? val : best;
        bcol = gt ? cb + j : bcol;
      }
    }
  }
}

// W [Dout, Din] fp32 -> hi/lo TF32 pair in the exact shared-memory image of the pipeline: per K chunk
// [hi half0..MH-1 | lo half0..MH-1], each half a 128-row x 128-byte tile, K-major, 128B-swizzled
// (16-byte unit j of row r stored at unit j ^ (r & 7)), zero padded.  One bulk copy per chunk then lands it.
__global__ void tf32_split_kernel(const float* __restrict__ W, int Dout, int Din, int MH, int nchunks,
                                  float* __restrict__ img) {
  const int64_t per_chunk = (int64_t)2 * MH * 128 * KCH;  // floats
  const int64_t n = (int64_t)nchunks * MH * 128 * KCH;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i % KCH);
    const int row = (int)((i / KCH) % (MH * 128));
    const int c = (int)(i / ((int64_t)KCH * MH * 128));
    const int k = c * KCH + kk;
    const float x = (row < Dout && k < Din) ? W[(size_t)row * Din + k] : 0.f;
    const float hi = tf32_rna(x);
    const int h = row >> 7, rr = row & 127, j = kk >> 2, e = kk & 3;
    const int64_t off = (int64_t)(rr >> 3) * 256 + (rr & 7) * 32 + ((j ^ (rr & 7)) << 2) + e;  // floats within a tile
    img[c * per_chunk + (int64_t)h * 4096 + off] = hi;
    img[c * per_chunk + (int64_t)(MH + h) * 4096 + off] = tf32_rna(x - hi);
  }
}

// W [Dout, Din] fp32 -> bf16 shared-memory image: per K chunk of 64 elements MH tiles of 128 rows x 128 bytes,
// K-major, 128B-swizzled (16-byte unit j of row r at unit j ^ (r & 7)), zero padded.
__global__ void bf16_image_kernel(const float* __restrict__ W, int Dout, int Din, int MH, int nchunks,
                                  unsigned short* __restrict__ img) {
  const int64_t per_chunk = (int64_t)MH * 128 * 64;  // bf16 elements
  const int64_t n = (int64_t)nchunks * per_chunk;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i % 64);
    const int row = (int)((i / 64) % (MH * 128));
    const int c = (int)(i / per_chunk);
    const int k = c * 64 + kk;
    const float x = (row < Dout && k < Din) ? W[(size_t)row * Din + k] : 0.f;
    unsigned short b;
    asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(b) : "f"(x));
    const int h = row >> 7, rr = row & 127, j = kk >> 3, e = kk & 7;
    const int64_t off = (int64_t)(rr >> 3) * 512 + (rr & 7) * 64 + ((j ^ (rr & 7)) << 3) + e;  // bf16 within a tile
    img[c * per_chunk + (int64_t)h * 8192 + off] = b;
  }
}

#ifdef MRG_TC_PROF
__device__ unsigned long long* g_tc_prof = nullptr;      // debugging aid: MMA-warp wait breakdown (scripts/prof_amax_tc.py)
#endif

struct AmaxParams {
  mrg_act x;                 // [rows, D] message source rows, read through act
  const int32_t* csr_eid;    // [E] row id per CSR position
  const int32_t* dst;        // [E] destination per edge id
  const float* wimg;         // [nchunks][2*MH tiles of 128x128B] pre-swizzled hi/lo image of W
  const float* bias;         // [D] or null
  unsigned long long* packed;  // [N, D] zero-initialised
  int64_t E;
  int D, MH, Kp, nchunks, num_tiles;
  int Dout;                  // rows of W actually present (amax: D; DistMult: query rows of this launch)
  // EPI_DISTMULT: logits [Dout, ldl] row-major (row b, column = tile position), labels likewise
  const float* label;
  float* logit;
  double* loss_partial;      // [gridDim.x]
  int64_t ldl;
  // EPI_STORE: store[(tile position) * lds + f] = acc + bias[f]   (plain Linear: out = x W^T + b)
  float* store;
  int64_t lds;
  int l2_prefetch;           // CTA-pair kernel: prefetch the next tile's gathered rows to L2
};

enum { EPI_AMAX = 0, EPI_DISTMULT = 1, EPI_STORE = 2 };

// One work ITEM = (tile, K chunk).  A CTA walks its tiles in PAIRS (TMEM slots 0/1) with the chunks of the
// two tiles interleaved, so every W chunk fetched from L2 feeds two tiles (W re-streaming is the dominant
// L2->SM traffic of this kernel):  items of pair p:  (A,c0) (B,c0) (A,c1) (B,c1) ...
struct Item {
  int pair, c, slot;
};
__device__ __forceinline__ Item item_of(int q, int nchunks, int full_pairs) {
  Item it;
  const int per_pair = 2 * nchunks;
  if (q < full_pairs * per_pair) {
    it.pair = q / per_pair;
    const int rem = q - it.pair * per_pair;
    it.c = rem >> 1;
    it.slot = rem & 1;
  } else {  // trailing single tile
    it.pair = full_pairs;
    it.c = q - full_pairs * per_pair;
    it.slot = 0;
  }
  return it;
}

// dynamic smem (1024-B aligned): W ring [WS][Whi MH tiles | Wlo MH tiles], X ring [XS][Xhi | Xlo], small arrays
// PREC: 0 = 3xTF32 (fp32-class accuracy: hi/lo operand pairs, three MMAs per K step);
//       2 = BF16 (operands rounded to bf16, ONE kind::f16 MMA per K step, K chunk = 64 elements per 128-byte row):
//           a third of the tensor work, half the operand bytes per K and 3x the X stages -- the reduced-precision
//           variant of the north_star, with its own stated tolerance (tests/test_gpu_ops_lp.py::test_amax_bf16_variant)
template <int MH, int EPI, int PREC = 0>
__global__ void __launch_bounds__(THREADS, 1) amax_tc_kernel(const AmaxParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr bool BF = PREC == 2;
  constexpr int KC = BF ? 64 : KCH;                       // operand elements per K chunk (one 128-byte row)
  constexpr int STAGES = BF ? 3 : tc::STAGES;             // W ring depth (shadows the namespace constant)
  constexpr uint32_t WSTAGE = (BF ? 1 : 2) * MH * TILE_BYTES;
  constexpr uint32_t XSTAGE = (BF ? 1 : 2) * TILE_BYTES;
  constexpr int XS = BF ? 6 : (MH == 1 ? 4 : 2);
  // align with pointer arithmetic on the __shared__ symbol (an integer round-trip would demote every access
  // below to generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_w = smem;
  uint8_t* smem_x = smem + STAGES * WSTAGE;
  uint8_t* tail = smem_x + XS * XSTAGE;
  uint64_t* xfull_bar = (uint64_t*)tail;              // [XS]      X item written (256 producer arrivals)
  uint64_t* xempty_bar = xfull_bar + XS;              // [XS]      MMAs that read the X stage retired
  uint64_t* wfull_bar = xempty_bar + XS;              // [STAGES]  W chunk landed (expect_tx)
  uint64_t* wempty_bar = wfull_bar + STAGES;          // [STAGES]  MMAs of both slots on this chunk retired
  uint64_t* tfull_bar = wempty_bar + STAGES;          // [2]       accumulator of slot complete
  uint64_t* tempty_bar = tfull_bar + 2;               // [2]       epilogue drained the slot
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);  // [1] (+pad)
  float* s_scale = (float*)(tmem_slot + 4);           // [256]
  float* s_shift = s_scale + 256;                     // [256]
  float* s_bias = s_shift + 256;                      // [256]
  int32_t* s_dst = (int32_t*)(s_bias + 256);          // [2 slots][128]
  int32_t* s_eid = s_dst + 2 * TILE_E;                // [2 slots][128]
  uint32_t* s_flag = (uint32_t*)(s_eid + 2 * TILE_E); // [2 slots][4] segment-start bitmask of the tile's columns
  double* s_loss = (double*)(s_flag + 8);             // [8] per-epilogue-warp BCE partial (EPI_DISTMULT)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = p.D;
  constexpr uint32_t TMEM_COLS = 2 * MH * TILE_E;  // 256 or 512

  for (int c = threadIdx.x; c < 256; c += THREADS) {
    s_scale[c] = (p.x.scale && c < D) ? p.x.scale[c] : 1.f;
    s_shift[c] = (p.x.scale && c < D) ? p.x.shift[c] : 0.f;
    s_bias[c] = (p.bias && c < p.Dout) ? p.bias[c] : 0.f;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < XS; ++s) {
      mbar_init(&xfull_bar[s], PROD_THREADS);
      mbar_init(&xempty_bar[s], 1);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&wfull_bar[s], 1);
      mbar_init(&wempty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], EPI_THREADS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int my_tiles = p.num_tiles > (int)blockIdx.x ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int full_pairs = my_tiles >> 1;
  const int npairs = (my_tiles + 1) >> 1;
  const int total = my_tiles * p.nchunks;   // items this CTA streams, in order
  auto tile_of = [&](int pair, int slot) { return (int64_t)blockIdx.x + (int64_t)(2 * pair + slot) * gridDim.x; };

  if (warp < EPI_WARP0) {
    // ================================ X PRODUCERS ================================
    constexpr int RPW = 32 / TPR;              // tile rows per producer warp
    const int r = warp * RPW + (lane % RPW);   // tile row
    const int half = lane / RPW;               // which 16-byte units of a 128-byte chunk: TPR*j + half
    const bool affine = p.x.scale != nullptr, relu = p.x.relu != 0;
    const uint32_t roff = (uint32_t)(r >> 3) * 1024 + (uint32_t)(r & 7) * 128;
    // The producers share the SM's issue slots with the epilogue and sit on the critical path (profiles/r01: ~540
    // warp instructions per item and warp before this rewrite): item coordinates are tracked incrementally (no
    // divisions), per-thread shared-memory offsets and per-tile row pointers are hoisted, the BN affine is read
    // with 128-bit shared loads.
    struct Cur { int pair, c, slot; };
    auto advance = [&](Cur& k) {
      if (k.pair < full_pairs && k.slot == 0) { k.slot = 1; return; }
      k.slot = 0;
      if (++k.c == p.nchunks) { k.c = 0; ++k.pair; }
    };
    uint32_t uoff[NU];
    int ucol[NU];
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      const int u = TPR * j + half;
      uoff[j] = roff + (uint32_t)((u ^ (r & 7)) << 4);
      ucol[j] = 4 * u;
    }
    const float* xrow0 = nullptr;   // load cursor's row per slot (null: past the last edge)
    const float* xrow1 = nullptr;
    bool cvalid0 = false, cvalid1 = false;   // consume cursor's row validity per slot
    constexpr int NB = BF ? 2 * NU : NU;     // float4 pieces per item and thread (a bf16 unit packs two float4)
    constexpr int PDP = BF ? 2 : PD;         // items in flight per producer thread
    float4 buf[PDP][NB];
    Cur lc = {0, 0, 0}, cc = {0, 0, 0};
    int lq = 0;
    if (BF) {
#pragma unroll
      for (int j = 0; j < NU; ++j) ucol[j] *= 2;          // a 16-byte bf16 unit covers 8 columns
    }
    // Gathered rows (a_max: x rows by CSR edge id) are prefetched to L2 one tile ahead: the loads of an item
    // then see L2 latency instead of HBM latency (round-2 ncu: 45 % of the stall samples were producers waiting
    // for rows, with only 2-4 items x 64-128 B per thread in flight).  Edge ids are pipelined two tiles ahead so
    // that neither the prefetch nor the row pointer waits on a load: e_cur (this tile), e_n1 (next), e_n2 (in flight).
    int32_t e_cur[2] = {-1, -1}, e_n1[2] = {-1, -1}, e_n2[2] = {-1, -1};
    bool primed[2] = {false, false};
    auto eid_at = [&](int pair, int slot) -> int32_t {
      const int64_t pos = tile_of(pair, slot) * TILE_E + r;
      return pos < p.E ? __ldg(p.csr_eid + pos) : -1;
    };
    auto load = [&](float4(&b)[NB]) {
      if (lc.c == 0) {
        const int sl = lc.slot;
        if (BF && p.csr_eid) {      // measured: 344 -> 322 us for the bf16 variant, 442 -> 459 us for 3xTF32 (not latency bound)
          if (!primed[sl]) {
            e_n1[sl] = eid_at(lc.pair, sl);
            e_n2[sl] = eid_at(lc.pair + 1, sl);
            primed[sl] = true;
          }
          e_cur[sl] = e_n1[sl];
          e_n1[sl] = e_n2[sl];
          e_n2[sl] = eid_at(lc.pair + 2, sl);
          if (e_n1[sl] >= 0) {      // rows of the NEXT tile of this slot -> L2 (4 threads per row, 2 lines each)
            const char* nrow = reinterpret_cast<const char*>(p.x.data + (size_t)e_n1[sl] * D);
            const int lines = (D * 4 + 127) / 128 + 1;          // rows are not 128-byte aligned
#pragma unroll
            for (int l = 0; l < 2; ++l)
              if (2 * half + l < lines) asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow + (2 * half + l) * 128));
          }
          const float* v = e_cur[sl] >= 0 ? p.x.data + (size_t)e_cur[sl] * D : nullptr;
          if (sl) xrow1 = v;
          else xrow0 = v;
        } else {
          const int64_t pos = tile_of(lc.pair, sl) * TILE_E + r;
          const float* v = pos < p.E ? p.x.data + (size_t)(p.csr_eid ? __ldg(p.csr_eid + pos) : (int32_t)pos) * D
                                     : nullptr;
          if (sl) xrow1 = v;
          else xrow0 = v;
        }
      }
      const float* xr = lc.slot ? xrow1 : xrow0;
      const int c0 = lc.c * KC;
      if (BF) {
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          const int col = c0 + ucol[j];
          b[2 * j] = (xr && col < D) ? ld_stream4(xr + col) : make_float4(0.f, 0.f, 0.f, 0.f);
          b[2 * j + 1] = (xr && col + 4 < D) ? ld_stream4(xr + col + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
#pragma unroll
        for (int j = 0; j < NU; ++j)
          b[j] = (xr && c0 + ucol[j] < D) ? ld_stream4(xr + c0 + ucol[j]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      advance(lc);
    };
    auto activate = [&](float4 v, int col, bool valid) {     // lazy BN affine + ReLU of 4 columns
      if (valid && col < D) {
        if (affine) {
          const float4 sc = *reinterpret_cast<const float4*>(s_scale + col);
          const float4 sh = *reinterpret_cast<const float4*>(s_shift + col);
          v.x = fmaf(sc.x, v.x, sh.x); v.y = fmaf(sc.y, v.y, sh.y);
          v.z = fmaf(sc.z, v.z, sh.z); v.w = fmaf(sc.w, v.w, sh.w);
        }
        if (relu) {
          v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
      }
      return v;
    };
    auto consume = [&](float4(&b)[NB], int q) {
      if (cc.c == 0) {
        const bool ok = tile_of(cc.pair, cc.slot) * TILE_E + r < p.E;
        if (cc.slot) cvalid1 = ok;
        else cvalid0 = ok;
      }
      const bool valid = cc.slot ? cvalid1 : cvalid0;
      const int c0 = cc.c * KC;
      const int s = q % XS;
      const uint32_t ph = (q / XS) & 1;
      mbar_wait(&xempty_bar[s], ph ^ 1);
      uint8_t* xhi = smem_x + (size_t)s * XSTAGE;
      uint8_t* xlo = xhi + TILE_BYTES;
      if (BF) {
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          const int col = c0 + ucol[j];
          const float4 a = activate(b[2 * j], col, valid), c4 = activate(b[2 * j + 1], col + 4, valid);
          uint4 pk;      // 8 bf16: cvt.rn.bf16x2 packs (hi operand -> upper half)
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.x) : "f"(a.y), "f"(a.x));
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.y) : "f"(a.w), "f"(a.z));
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.z) : "f"(c4.y), "f"(c4.x));
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.w) : "f"(c4.w), "f"(c4.z));
          *reinterpret_cast<uint4*>(xhi + uoff[j]) = pk;
        }
      } else {
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        const int col = c0 + ucol[j];
        float4 v = b[j];
        if (valid && col < D) {
          if (affine) {
            const float4 sc = *reinterpret_cast<const float4*>(s_scale + col);
            const float4 sh = *reinterpret_cast<const float4*>(s_shift + col);
            v.x = fmaf(sc.x, v.x, sh.x);
            v.y = fmaf(sc.y, v.y, sh.y);
            v.z = fmaf(sc.z, v.z, sh.z);
            v.w = fmaf(sc.w, v.w, sh.w);
          }
          if (relu) {
            v.x = v.x > 0.f ? v.x : 0.f;
            v.y = v.y > 0.f ? v.y : 0.f;
            v.z = v.z > 0.f ? v.z : 0.f;
            v.w = v.w > 0.f ? v.w : 0.f;
          }
        }
        const float4 hi = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
        const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);  // exact; MMA truncates to tf32
        *reinterpret_cast<float4*>(xhi + uoff[j]) = hi;
        *reinterpret_cast<float4*>(xlo + uoff[j]) = lo;
      }
      }
      fence_proxy_async();
      mbar_arrive(&xfull_bar[s]);
      advance(cc);
    };
#pragma unroll
    for (int u = 0; u < PDP; ++u) {
      if (lq < total) load(buf[u]);
      ++lq;
    }
    for (int q = 0; q < total; q += PDP) {
#pragma unroll
      for (int u = 0; u < PDP; ++u) {
        if (q + u < total) {
          consume(buf[u], q + u);
          if (lq < total) load(buf[u]);
          ++lq;
        }
      }
    }
  } else if (warp == WLD_WARP) {
    // ================================ W LOADER ================================
    if (lane == 0) {
      const int wtotal = npairs * p.nchunks;
      for (int wq = 0; wq < wtotal; ++wq) {
        const int s = wq % STAGES;
        const uint32_t ph = (wq / STAGES) & 1;
        const int c = wq % p.nchunks;
        mbar_wait(&wempty_bar[s], ph ^ 1);
        mbar_expect_tx(&wfull_bar[s], WSTAGE);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wimg) + (size_t)c * WSTAGE;
        const uint32_t dst = smem_u32(smem_w + (size_t)s * WSTAGE);
        for (uint32_t o = 0; o < WSTAGE; o += TILE_BYTES) bulk_g2s(dst + o, src + o, TILE_BYTES, &wfull_bar[s]);
      }
    }
  } else if (warp == MMA_WARP) {
    // ================================ MMA ISSUER ================================
    const uint32_t idesc = BF ? umma_idesc_bf16(128, TILE_E) : umma_idesc(128, TILE_E);
    int q = 0, wq = 0;
#ifdef MRG_TC_PROF
    long long t_te = 0, t_w = 0, t_x = 0, t_iss = 0, t_m = clock64();
#define TC_MARK(v) do { const long long n_ = clock64(); v += n_ - t_m; t_m = n_; } while (0)
#else
#define TC_MARK(v) do { } while (0)
#endif
    for (int pr = 0; pr < npairs; ++pr) {
      const int ntp = min(2, my_tiles - 2 * pr);
      TC_MARK(t_iss);
      for (int ts = 0; ts < ntp; ++ts) mbar_wait(&tempty_bar[ts], (pr & 1) ^ 1);
      TC_MARK(t_te);
      tc_fence_after();
      for (int c = 0; c < p.nchunks; ++c, ++wq) {
        const int ws = wq % STAGES;
        TC_MARK(t_iss);
        mbar_wait(&wfull_bar[ws], (wq / STAGES) & 1);
        TC_MARK(t_w);
        const uint32_t wbase = smem_u32(smem_w + (size_t)ws * WSTAGE);
        const int ksteps = BF ? (min(KC, D - c * KC) + 15) / 16 : min(KCH, D - c * KCH) / 8;   // zero padded
        for (int ts = 0; ts < ntp; ++ts, ++q) {
          const int xs = q % XS;
          TC_MARK(t_iss);
          mbar_wait(&xfull_bar[xs], (q / XS) & 1);
          TC_MARK(t_x);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t xhi = smem_u32(smem_x + (size_t)xs * XSTAGE), xlo = xhi + TILE_BYTES;
            for (int h = 0; h < MH; ++h) {
              const uint32_t whi = wbase + h * TILE_BYTES, wlo = wbase + (MH + h) * TILE_BYTES;
              const uint32_t d_tmem = tmem_base + ts * (MH * TILE_E) + h * TILE_E;
              for (int k = 0; k < ksteps; ++k) {
                const uint32_t ko = k * 32;  // 8 tf32 / 16 bf16 = 32 bytes inside the swizzled 128B row
                const uint32_t acc = (c > 0 || k > 0) ? 1u : 0u;
                if (BF) {
                  umma_bf16(d_tmem, umma_desc(wbase + h * TILE_BYTES + ko), umma_desc(xhi + ko), idesc, acc);
                } else {
                  umma_tf32(d_tmem, umma_desc(wlo + ko), umma_desc(xhi + ko), idesc, acc);  // small terms first
                  umma_tf32(d_tmem, umma_desc(whi + ko), umma_desc(xlo + ko), idesc, 1u);
                  umma_tf32(d_tmem, umma_desc(whi + ko), umma_desc(xhi + ko), idesc, 1u);
                }
              }
            }
            umma_commit(&xempty_bar[xs]);                          // X stage free when these MMAs retire
            if (c == p.nchunks - 1) umma_commit(&tfull_bar[ts]);   // accumulator of this slot complete
          }
          __syncwarp();
        }
        if (lane == 0) umma_commit(&wempty_bar[ws]);               // W stage free (both slots done with it)
        __syncwarp();
      }
    }
#ifdef MRG_TC_PROF
    TC_MARK(t_iss);
    if (lane == 0 && EPI == EPI_AMAX && g_tc_prof) {
      atomicAdd(g_tc_prof + 0, (unsigned long long)t_te);
      atomicAdd(g_tc_prof + 1, (unsigned long long)t_w);
      atomicAdd(g_tc_prof + 2, (unsigned long long)t_x);
      atomicAdd(g_tc_prof + 3, (unsigned long long)t_iss);
      atomicAdd(g_tc_prof + 4, (unsigned long long)(my_tiles * p.nchunks));
    }
#endif
#undef TC_MARK
  } else {
    // ================================ EPILOGUE (2 warp sets, one per TMEM slot) ================================
    const int ts = (warp - EPI_WARP0) >> 2;           // slot served by this warp set
    const int et = (threadIdx.x - PROD_THREADS) & 127;  // 0..127 == TMEM lane
    const int quad = warp & 3;                        // TMEM lane quadrant of this warp
    int32_t* sd = s_dst + ts * TILE_E;
    int32_t* se = s_eid + ts * TILE_E;
    if (EPI == EPI_STORE) {
      // ---- plain GEMM epilogue: TMEM lane = output feature f, TMEM column = input row of the tile; the 32 lanes of a
      // warp write 32 consecutive features of one output row (coalesced).
      for (int pr = 0; pr < npairs; ++pr) {
        if (2 * pr + ts >= my_tiles) break;
        const int64_t pos0 = tile_of(pr, ts) * TILE_E;
        const int cnt = (int)min((int64_t)TILE_E, p.E - pos0);
        mbar_wait(&tfull_bar[ts], pr & 1);
        tc_fence_after();
        for (int h = 0; h < MH; ++h) {
          const int f = h * 128 + et;
          const bool fvalid = f < p.Dout;
          const float bias = s_bias[f & 255];
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + ts * (MH * TILE_E) + h * TILE_E;
          float* o = p.store + (size_t)pos0 * p.lds + f;
#pragma unroll 1
          for (int w = 0; w < 4; ++w) {
            const int cb = 32 * w;
            if (cb >= cnt) break;   // warp-uniform
            uint32_t v[32];
            tmem_ld32(taddr + cb, v);
            if (fvalid) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (cb + j < cnt) o[(size_t)(cb + j) * p.lds] = __uint_as_float(v[j]) + bias;
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[ts]);
      }
    } else if (EPI == EPI_DISTMULT) {
      // ---- DistMult 1-N scores + BCE: TMEM lane = query row b, TMEM column = entity of the tile.
      // logit[b, n] is stored for the backward; the loss terms are summed per thread (fp32 per 32 columns,
      // double across), per warp, per CTA -> loss_partial[cta]; nothing else leaves the SM.
      double acc = 0.0;
      for (int pr = 0; pr < npairs; ++pr) {
        if (2 * pr + ts >= my_tiles) break;
        const int64_t pos0 = tile_of(pr, ts) * TILE_E;
        const int cnt = (int)min((int64_t)TILE_E, p.E - pos0);
        mbar_wait(&tfull_bar[ts], pr & 1);
        tc_fence_after();
        // TMEM lane = query row, so a thread holds 32 consecutive entities of ITS row: read / written straight from
        // there, a warp touches 32 different rows of label and logit per instruction (4 useful bytes per sector; the
        // round-2 profile had this kernel latency bound at 134 us).  Each 32 x 16 block therefore goes through a small
        // shared-memory tile and leaves transposed: 2 rows x 16 consecutive entities per instruction.
        float* tb = reinterpret_cast<float*>(tail + 8192) + (size_t)(warp - EPI_WARP0) * 32 * 17;
        const int cc = lane & 15, r2 = lane >> 4;
        for (int h = 0; h < MH; ++h) {
          const int brow0 = h * 128 + quad * 32;      // first query row of this warp's TMEM lanes
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + ts * (MH * TILE_E) + h * TILE_E;
#pragma unroll 1
          for (int w = 0; w < 4; ++w) {
            const int cb = 32 * w;
            if (cb >= cnt) break;   // warp-uniform
            uint32_t v[32];
            tmem_ld32(taddr + cb, v);
            float facc = 0.f;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
              for (int j = 0; j < 16; ++j) tb[lane * 17 + j] = __uint_as_float(v[16 * hh + j]);
              __syncwarp();
              const int col = cb + 16 * hh + cc;
#pragma unroll 4
              for (int it = 0; it < 16; ++it) {
                const int rr = 2 * it + r2, b = brow0 + rr;
                if (b < p.Dout && col < cnt) {
                  const float z = tb[rr * 17 + cc];
                  const size_t off = (size_t)b * p.ldl + pos0 + col;
                  float pr_;
                  facc += bce_term(z, __ldg(p.label + off), &pr_);
                  p.logit[off] = z;
                }
              }
              __syncwarp();
            }
            acc += (double)facc;
          }
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[ts]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) s_loss[warp - EPI_WARP0] = acc;
    } else {
#ifdef MRG_TC_PROF
    long long e_pre = 0, e_wait = 0, e_scan = 0, e_ld = 0, e_m = clock64();
#define EP_MARK(v) do { const long long n_ = clock64(); v += n_ - e_m; e_m = n_; } while (0)
#else
#define EP_MARK(v) do { } while (0)
#endif
    for (int pr = 0; pr < npairs; ++pr) {
      if (2 * pr + ts >= my_tiles) break;
      const int64_t pos0 = tile_of(pr, ts) * TILE_E;
      const int cnt = (int)min((int64_t)TILE_E, p.E - pos0);
      EP_MARK(e_scan);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + ts) : "memory");   // previous tile's readers are done with sd/se
      {
        int32_t e = -1, d = -1;
        if (et < cnt) {
          e = __ldg(p.csr_eid + pos0 + et);
          d = __ldg(p.dst + e);
        }
        se[et] = e;
        sd[et] = d;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + ts) : "memory");
      {
        // bit j of word w: a new destination starts at column 32 w + j (column 0 never flags)
        const bool starts = et > 0 && et < cnt && sd[et] != sd[et - 1];
        const uint32_t word = __ballot_sync(0xffffffffu, starts);
        if (lane == 0) s_flag[ts * 4 + quad] = word;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + ts) : "memory");
      EP_MARK(e_pre);
      mbar_wait(&tfull_bar[ts], pr & 1);
      EP_MARK(e_wait);
      tc_fence_after();
      for (int h = 0; h < MH; ++h) {
        const int f = h * 128 + et;
        const bool fvalid = f < p.Dout;
        const float bias = s_bias[f & 255];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + ts * (MH * TILE_E) + h * TILE_E;
        // Register-only segmented max.  relu(v) >= 0, so starting every segment at (best = 0, bcol = first column)
        // and updating on a strict v > best IS max(relu(v)) with the lowest-column tie-break -- 4 instructions per
        // column (FADD, FSETP, FSEL, SEL); memory is touched only when a segment closes (a few times per tile).
        float best = 0.f;
        int bcol = 0;
        auto close_seg = [&](int col) {   // columns [.., col) of the running destination are done
          if (fvalid) {
            const unsigned long long key = ((unsigned long long)__float_as_uint(best) << 32) |
                                           (unsigned long long)(0xFFFFFFFFu - (uint32_t)se[bcol]);
            atomicMax(p.packed + (size_t)sd[col - 1] * D + f, key);
          }
        };
#pragma unroll 1
        for (int w = 0; w < 4; ++w) {
          const int cb = 32 * w;
          if (cb >= cnt) break;   // warp-uniform
          uint32_t v[32];
#ifdef MRG_TC_PROF
          const long long t_ld0 = clock64();
#endif
          tmem_ld32(taddr + cb, v);
#ifdef MRG_TC_PROF
          e_ld += clock64() - t_ld0;
#endif
          const uint32_t fl = s_flag[ts * 4 + w];
          if (cb + 32 <= cnt) {
            seg_scan32(v, fl, cb, bias, best, bcol, close_seg);
          } else {   // ragged last tile
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (cb + j < cnt) {
                if (fl & (1u << j)) {
                  close_seg(cb + j);
                  best = 0.f;
                  bcol = cb + j;
                }
                const float val = __uint_as_float(v[j]) + bias;
                const bool gt = val > best;
                best = gt ? val : best;
                bcol = gt ? cb + j : bcol;
              }
            }
          }
        }
        if (cnt > 0) close_seg(cnt);
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[ts]);
    }
#ifdef MRG_TC_PROF
    EP_MARK(e_scan);
    if (et == 0 && ts == 0 && g_tc_prof) {
      atomicAdd(g_tc_prof + 5, (unsigned long long)e_pre);
      atomicAdd(g_tc_prof + 6, (unsigned long long)e_wait);
      atomicAdd(g_tc_prof + 7, (unsigned long long)e_scan);
      atomicAdd(g_tc_prof + 8, (unsigned long long)e_ld);
    }
#endif
#undef EP_MARK
    }
  }

  tc_fence_before();
  __syncthreads();
  if (EPI == EPI_DISTMULT && threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_loss[w];
    p.loss_partial[blockIdx.x] = t;
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// packed key -> out (+ residual) and the encoded argmax (see segreduce.cu for the code)
__global__ void amax_finalize_kernel(const unsigned long long* __restrict__ packed, int64_t n, int D, mrg_act res,
                                     float* __restrict__ out, int32_t* __restrict__ arg) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = packed[i];
    float val = 0.f;
    int32_t a = -1;
    if (key != 0ull) {
      val = __uint_as_float((uint32_t)(key >> 32));
      const int32_t e = (int32_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
      a = val > 0.f ? e : -2 - e;
    }
    if (res.data) {
      float rv = res.data[i];
      const int c = (int)(i % D);
      if (res.scale) rv = fmaf(res.scale[c], rv, res.shift[c]);
      if (res.relu) rv = rv > 0.f ? rv : 0.f;
      val += rv;
    }
    out[i] = val;
    if (arg) arg[i] = a;
  }
}

// ------------------------------------------------------------------------------------------
// CTA-PAIR form of the fused a_max forward (tcgen05 cta_group::2), 128 < D <= 256, 3xTF32.
//
// What the single-CTA kernel above taught (scripts/prof_amax_tc.py, profiles/r02_ncu_amax_tc.md): per
// (128 edges x 32 K) item the MMA warp spends ~2,050 cycles issuing/back-pressured -- 24 MMAs that each read an
// 8 KB operand pair from shared memory: 192 KB + 32 KB of producer stores + 32 KB of W landing = 256 KB at
// 128 B/clk, i.e. the kernel is SHARED-MEMORY-BANDWIDTH bound, not tensor bound -- and ~1,150 cycles waiting
// for the epilogue, because both TMEM slots of a tile pair finish together (the pair exists so that one W
// chunk streamed from L2 feeds 256 edges).
//
// A CTA pair removes both: UMMA M = 256 puts feature half r on CTA r (each CTA loads and reads only ITS half of
// every W chunk), UMMA N = 256 edges per tile of which CTA r gathers / converts rows [128 r, 128 r + 128) -- the
// B operand is shared across the pair, so each SM's shared memory feeds 8 KB per MMA of twice the size
// (64 B/clk) -- and a tile is 256 TMEM columns per CTA, so consecutive tiles alternate between two slots and the
// epilogue of tile i overlaps the MMAs of tile i + 1.  W is still re-streamed once per 256 edges.
//
// Cross-CTA protocol (leader = cluster rank 0 issues every MMA):
//   xfull[s]   leader   32 arrivals: one per producer warp of BOTH CTAs (rank 1 arrives remotely through mapa)
//   wland[s]   local    the CTA's own W half landed (expect_tx); rank 1's idle MMA warp relays it to
//   wfull[s]   leader   2 arrivals: own loader (arrive.expect_tx) + the relay
//   xempty[s], wempty[s], tfull[slot]   in BOTH CTAs, arrived by tcgen05.commit ... multicast::cluster
//   tempty[slot] leader 8 arrivals: the 4 epilogue warps of the slot in both CTAs
// ------------------------------------------------------------------------------------------
constexpr int TILE2 = 256;
constexpr int PD2 = 2;          // items in flight per producer thread (2, 3, 4 measured alike: 285-292 us; 6 spills)
__host__ __device__ constexpr uint32_t tc2_stage_bytes(int prec) { return (prec == 2 ? 1 : 2) * TILE_BYTES; }   // [hi | lo] or bf16
__host__ __device__ constexpr int tc2_xstages(int prec) { return prec == 2 ? 6 : 3; }
__host__ __device__ constexpr int tc2_wstages(int prec) { return prec == 2 ? 4 : 3; }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster.  Default (.release.cta)
// semantics on purpose: .release.cluster compiles to MEMBAR.ALL.GPU + ERRBAR per arrival (measured: 2,600 cycles
// per producer item).  What the arrival publishes is shared memory that the writer already made visible to the
// async proxy (MEMBAR.CTA + FENCE.VIEW.ASYNC) and that only the tensor core reads, or TMEM reads that
// tcgen05.fence::before_thread_sync ordered.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once the MMAs issued so far retire) on the barrier at this offset in both CTAs of the pair
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}

// PREC 0: 3xTF32 (hi | lo tiles per stage, three MMAs per K step of 8); PREC 2: bf16 operands (one 16 KB tile per
// stage holding 64 K elements per 128-byte row, one kind::f16 MMA per K step of 16, twice the stages)
template <int PREC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) amax_tc2_kernel(const AmaxParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr bool BF = PREC == 2;
  constexpr int KC = BF ? 64 : KCH;
  constexpr uint32_t STAGE2 = tc2_stage_bytes(PREC);
  constexpr int XS2 = tc2_xstages(PREC), WS2 = tc2_wstages(PREC);
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_w = smem;
  uint8_t* smem_x = smem + WS2 * STAGE2;
  uint8_t* tail = smem_x + XS2 * STAGE2;
  uint64_t* xfull_bar = (uint64_t*)tail;          // [XS2]
  uint64_t* xempty_bar = xfull_bar + XS2;         // [XS2]
  uint64_t* wfull_bar = xempty_bar + XS2;         // [WS2]
  uint64_t* wland_bar = wfull_bar + WS2;          // [WS2]
  uint64_t* wempty_bar = wland_bar + WS2;         // [WS2]
  uint64_t* tfull_bar = wempty_bar + WS2;         // [2]
  uint64_t* tempty_bar = tfull_bar + 2;           // [2]
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);
  float* s_scale = (float*)(tail + 512);          // [256]  (16-byte aligned: read with 128-bit loads)
  float* s_shift = s_scale + 256;                 // [256]
  float* s_bias = s_shift + 256;                  // [256]
  int32_t* s_dst = (int32_t*)(s_bias + 256);      // [2 slots][256]
  int32_t* s_eid = s_dst + 2 * TILE2;             // [2 slots][256]
  uint32_t* s_flag = (uint32_t*)(s_eid + 2 * TILE2);  // [2 slots][8]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = p.D;
  const uint32_t rank = cluster_ctarank();
  const int ncl = (int)gridDim.x >> 1, cl = (int)blockIdx.x >> 1;

  for (int c = threadIdx.x; c < 256; c += THREADS) {
    s_scale[c] = (p.x.scale && c < D) ? p.x.scale[c] : 1.f;
    s_shift[c] = (p.x.scale && c < D) ? p.x.shift[c] : 0.f;
    s_bias[c] = (p.bias && c < D) ? p.bias[c] : 0.f;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < XS2; ++s) {
      mbar_init(&xfull_bar[s], 2 * (PROD_THREADS / 32));   // one arrival per producer warp of both CTAs
      mbar_init(&xempty_bar[s], 1);
    }
    for (int s = 0; s < WS2; ++s) {
      mbar_init(&wfull_bar[s], 2);
      mbar_init(&wland_bar[s], 1);
      mbar_init(&wempty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();       // barriers of the peer are initialised, both TMEM halves allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int my_tiles = p.num_tiles > cl ? (p.num_tiles - 1 - cl) / ncl + 1 : 0;   // 256-edge tiles of this pair
  const int nch = p.nchunks;
  const int total = my_tiles * nch;
  auto tile_pos0 = [&](int i) { return ((int64_t)cl + (int64_t)i * ncl) * TILE2; };

  if (warp < EPI_WARP0) {
    // ================================ X PRODUCERS (both CTAs, 128 rows each) ================================
    constexpr int RPW = 32 / TPR;
    const int r = warp * RPW + (lane % RPW);   // row of this CTA's half tile
    const int half = lane / RPW;
    const bool affine = p.x.scale != nullptr, relu = p.x.relu != 0;
    const uint32_t roff = (uint32_t)(r >> 3) * 1024 + (uint32_t)(r & 7) * 128;
    uint32_t uoff[NU];
    int ucol[NU];
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      const int u = TPR * j + half;
      uoff[j] = roff + (uint32_t)((u ^ (r & 7)) << 4);
      ucol[j] = 4 * u;
    }
    constexpr int NB = BF ? 2 * NU : NU;     // float4 pieces per item and thread (a bf16 unit packs two float4)
    if (BF) {
#pragma unroll
      for (int j = 0; j < NU; ++j) ucol[j] *= 2;          // a 16-byte bf16 unit covers 8 columns
    }
    float4 buf[PD2][NB];
    int l_i = 0, l_c = 0, c_i = 0, c_c = 0, lq = 0;
    const float* xrow = nullptr;
    bool cvalid = false;
    // Row ids run one tile ahead of the load cursor: the row pointer never waits on the id, and one item into a tile
    // (the id has arrived by then) the 4 threads that share a row prefetch the NEXT tile's whole row to L2, two
    // 128-byte lines each.  An item reads only 128 B of each row and comes back for the next piece ~2,000 cycles
    // later; without the prefetch every piece is a separate DRAM visit whose latency the producers eat (round-2 ncu:
    // 30 % of all stall samples on the first use of the gathered registers, independent of the ring depth).
    auto eid_of = [&](int i) -> int32_t {
      if (i >= my_tiles) return -1;
      const int64_t pos = tile_pos0(i) + 128 * (int)rank + r;
      if (pos >= p.E) return -1;
      return p.csr_eid ? __ldg(p.csr_eid + pos) : (int32_t)pos;
    };
    int32_t e_nxt = eid_of(0);
    const int pf_lines = (D * 4 + 127) / 128 + 1;     // rows are not 128-byte aligned
    auto load = [&](float4(&b)[NB]) {
      if (l_c == 0) {
        xrow = e_nxt >= 0 ? p.x.data + (size_t)e_nxt * D : nullptr;
        e_nxt = eid_of(l_i + 1);
      } else if (l_c == 1 && p.l2_prefetch && e_nxt >= 0) {
        const char* nrow = reinterpret_cast<const char*>(p.x.data + (size_t)e_nxt * D);
#pragma unroll
        for (int l = 0; l < 2; ++l)
          if (2 * half + l < pf_lines) asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow + (2 * half + l) * 128));
      }
      const int c0 = l_c * KC;
      if (BF) {
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          const int col = c0 + ucol[j];
          b[2 * j] = (xrow && col < D) ? ld_stream4(xrow + col) : make_float4(0.f, 0.f, 0.f, 0.f);
          b[2 * j + 1] = (xrow && col + 4 < D) ? ld_stream4(xrow + col + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
#pragma unroll
        for (int j = 0; j < NU; ++j)
          b[j] = (xrow && c0 + ucol[j] < D) ? ld_stream4(xrow + c0 + ucol[j]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (++l_c == nch) { l_c = 0; ++l_i; }
    };
    auto activate = [&](float4 v, int col) {     // lazy BN affine + ReLU of 4 columns of a valid row
      if (cvalid && col < D) {
        if (affine) {
          const float4 sc = *reinterpret_cast<const float4*>(s_scale + col);
          const float4 sh = *reinterpret_cast<const float4*>(s_shift + col);
          v.x = fmaf(sc.x, v.x, sh.x);
          v.y = fmaf(sc.y, v.y, sh.y);
          v.z = fmaf(sc.z, v.z, sh.z);
          v.w = fmaf(sc.w, v.w, sh.w);
        }
        if (relu) {
          v.x = v.x > 0.f ? v.x : 0.f;
          v.y = v.y > 0.f ? v.y : 0.f;
          v.z = v.z > 0.f ? v.z : 0.f;
          v.w = v.w > 0.f ? v.w : 0.f;
        }
      }
      return v;
    };
    auto consume = [&](float4(&b)[NB], int q) {
      if (c_c == 0) cvalid = tile_pos0(c_i) + 128 * (int)rank + r < p.E;
      const int c0 = c_c * KC;
      const int s = q % XS2;
      const uint32_t ph = (q / XS2) & 1;
      mbar_wait(&xempty_bar[s], ph ^ 1);
      uint8_t* xhi = smem_x + (size_t)s * STAGE2;
      uint8_t* xlo = xhi + TILE_BYTES;
      if (BF) {
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          const int col = c0 + ucol[j];
          const float4 a = activate(b[2 * j], col), c4 = activate(b[2 * j + 1], col + 4);
          uint4 pk;      // 8 bf16: cvt.rn.bf16x2 packs (first operand -> upper half)
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.x) : "f"(a.y), "f"(a.x));
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.y) : "f"(a.w), "f"(a.z));
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.z) : "f"(c4.y), "f"(c4.x));
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.w) : "f"(c4.w), "f"(c4.z));
          *reinterpret_cast<uint4*>(xhi + uoff[j]) = pk;
        }
      } else {
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          const float4 v = activate(b[j], c0 + ucol[j]);
          const float4 hi = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
          const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);   // exact; the MMA truncates to tf32
          *reinterpret_cast<float4*>(xhi + uoff[j]) = hi;
          *reinterpret_cast<float4*>(xlo + uoff[j]) = lo;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&xfull_bar[s], 0);   // one arrival per warp, on the leader's barrier
      if (++c_c == nch) { c_c = 0; ++c_i; }
    };
#pragma unroll
    for (int u = 0; u < PD2; ++u) {
      if (lq < total) load(buf[u]);
      ++lq;
    }
    for (int q = 0; q < total; q += PD2) {
#pragma unroll
      for (int u = 0; u < PD2; ++u) {
        if (q + u < total) {
          consume(buf[u], q + u);
          if (lq < total) load(buf[u]);
          ++lq;
        }
      }
    }
  } else if (warp == WLD_WARP) {
    // ================================ W LOADER (own feature half: hi tile, lo tile) ================================
    if (lane == 0) {
      for (int q = 0; q < total; ++q) {
        const int s = q % WS2;
        const uint32_t ph = (q / WS2) & 1;
        const int c = q % nch;
        mbar_wait(&wempty_bar[s], ph ^ 1);
        uint64_t* bar = rank == 0 ? &wfull_bar[s] : &wland_bar[s];
        mbar_expect_tx(bar, STAGE2);
        // image of chunk c, 16 KB tiles: [hi half 0 | hi half 1 | lo half 0 | lo half 1]  (bf16: [half 0 | half 1])
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wimg) + (size_t)c * (BF ? 2 : 4) * TILE_BYTES;
        const uint32_t dst = smem_u32(smem_w + (size_t)s * STAGE2);
        bulk_g2s(dst, src + (size_t)rank * TILE_BYTES, TILE_BYTES, bar);
        if (!BF) bulk_g2s(dst + TILE_BYTES, src + (size_t)(2 + rank) * TILE_BYTES, TILE_BYTES, bar);
      }
    }
  } else if (warp == MMA_WARP) {
    if (rank != 0) {
      // ---- relay: this CTA's W half has landed -> tell the leader
      if (lane == 0) {
        for (int q = 0; q < total; ++q) {
          const int s = q % WS2;
          mbar_wait(&wland_bar[s], (q / WS2) & 1);
          mbar_arrive_cluster(&wfull_bar[s], 0);
        }
      }
    } else {
      // ================================ MMA ISSUER (leader CTA) ================================
      const uint32_t idesc = BF ? umma_idesc_bf16(256, TILE2) : umma_idesc(256, TILE2);
      int q = 0;
#ifdef MRG_TC_PROF
      long long t_te = 0, t_w = 0, t_x = 0, t_iss = 0, t_m = clock64();
#define TC_MARK(v) do { const long long n_ = clock64(); v += n_ - t_m; t_m = n_; } while (0)
#else
#define TC_MARK(v) do { } while (0)
#endif
      for (int i = 0; i < my_tiles; ++i) {
        const int slot = i & 1;
        TC_MARK(t_iss);
        mbar_wait(&tempty_bar[slot], ((i >> 1) & 1) ^ 1);
        TC_MARK(t_te);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + slot * TILE2;
        for (int c = 0; c < nch; ++c, ++q) {
          const int ws = q % WS2, xs = q % XS2;
          TC_MARK(t_iss);
          mbar_wait(&wfull_bar[ws], (q / WS2) & 1);
          TC_MARK(t_w);
          mbar_wait(&xfull_bar[xs], (q / XS2) & 1);
          TC_MARK(t_x);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t whi = smem_u32(smem_w + (size_t)ws * STAGE2), wlo = whi + TILE_BYTES;
            const uint32_t xhi = smem_u32(smem_x + (size_t)xs * STAGE2), xlo = xhi + TILE_BYTES;
            const int ksteps = BF ? (min(KC, D - c * KC) + 15) / 16 : min(KC, D - c * KC) / 8;   // zero padded
            for (int k = 0; k < ksteps; ++k) {
              const uint32_t ko = k * 32;    // 8 tf32 / 16 bf16 = 32 bytes inside the swizzled 128-byte row
              const uint32_t acc = (c > 0 || k > 0) ? 1u : 0u;
              if (BF) {
                umma2_bf16(d_tmem, umma_desc(whi + ko), umma_desc(xhi + ko), idesc, acc);
              } else {
                umma2_tf32(d_tmem, umma_desc(wlo + ko), umma_desc(xhi + ko), idesc, acc);   // small terms first
                umma2_tf32(d_tmem, umma_desc(whi + ko), umma_desc(xlo + ko), idesc, 1u);
                umma2_tf32(d_tmem, umma_desc(whi + ko), umma_desc(xhi + ko), idesc, 1u);
              }
            }
            umma2_commit(&xempty_bar[xs]);
            umma2_commit(&wempty_bar[ws]);
            if (c == nch - 1) umma2_commit(&tfull_bar[slot]);
          }
          __syncwarp();
        }
      }
#ifdef MRG_TC_PROF
      TC_MARK(t_iss);
      if (lane == 0 && g_tc_prof) {
        atomicAdd(g_tc_prof + 0, (unsigned long long)t_te);
        atomicAdd(g_tc_prof + 1, (unsigned long long)t_w);
        atomicAdd(g_tc_prof + 2, (unsigned long long)t_x);
        atomicAdd(g_tc_prof + 3, (unsigned long long)t_iss);
        atomicAdd(g_tc_prof + 4, (unsigned long long)total);
      }
#endif
#undef TC_MARK
    }
  } else {
    // ================================ EPILOGUE (4 warps per TMEM slot, thread = feature 128 rank + lane id) ==========
    const int ts = (warp - EPI_WARP0) >> 2;
    const int et = (threadIdx.x - PROD_THREADS) & 127;
    const int quad = warp & 3;
    int32_t* sd = s_dst + ts * TILE2;
    int32_t* se = s_eid + ts * TILE2;
    const int f = 128 * (int)rank + et;
    const bool fvalid = f < D;
    const float bias = s_bias[f];
#ifdef MRG_TC_PROF
    long long e_pre = 0, e_wait = 0, e_scan = 0, e_m = clock64();
#define EP_MARK(v) do { const long long n_ = clock64(); v += n_ - e_m; e_m = n_; } while (0)
#else
#define EP_MARK(v) do { } while (0)
#endif
    for (int i = ts; i < my_tiles; i += 2) {
      const int64_t pos0 = tile_pos0(i);
      const int cnt = (int)min((int64_t)TILE2, p.E - pos0);
      EP_MARK(e_scan);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + ts) : "memory");   // previous tile's readers are done with sd/se
#pragma unroll
      for (int hcol = 0; hcol < 2; ++hcol) {
        const int col = et + 128 * hcol;
        int32_t e = -1, d = -1;
        if (col < cnt) {
          e = __ldg(p.csr_eid + pos0 + col);
          d = __ldg(p.dst + e);
        }
        se[col] = e;
        sd[col] = d;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + ts) : "memory");
#pragma unroll
      for (int hcol = 0; hcol < 2; ++hcol) {
        const int col = et + 128 * hcol;
        const bool starts = col > 0 && col < cnt && sd[col] != sd[col - 1];
        const uint32_t word = __ballot_sync(0xffffffffu, starts);
        if (lane == 0) s_flag[ts * 8 + 4 * hcol + quad] = word;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + ts) : "memory");
      EP_MARK(e_pre);
      mbar_wait(&tfull_bar[ts], ((i >> 1) & 1));
      EP_MARK(e_wait);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + ts * TILE2;
      float best = 0.f;
      int bcol = 0;
      auto close_seg = [&](int col) {
        if (fvalid) {
          const unsigned long long key = ((unsigned long long)__float_as_uint(best) << 32) |
                                         (unsigned long long)(0xFFFFFFFFu - (uint32_t)se[bcol]);
          atomicMax(p.packed + (size_t)sd[col - 1] * D + f, key);
        }
      };
#pragma unroll 1
      for (int w = 0; w < 8; ++w) {
        const int cb = 32 * w;
        if (cb >= cnt) break;
        uint32_t v[32];
        tmem_ld32(taddr + cb, v);
        const uint32_t fl = s_flag[ts * 8 + w];
        if (cb + 32 <= cnt) {
          seg_scan32(v, fl, cb, bias, best, bcol, close_seg);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (cb + j < cnt) {
              if (fl & (1u << j)) {
                close_seg(cb + j);
                best = 0.f;
                bcol = cb + j;
              }
              const float val = __uint_as_float(v[j]) + bias;
              const bool gt = val > best;
              best = gt ? val : best;
              bcol = gt ? cb + j : bcol;
            }
          }
        }
      }
      if (cnt > 0) close_seg(cnt);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty_bar[ts], 0);
    }
#ifdef MRG_TC_PROF
    EP_MARK(e_scan);
    if (et == 0 && ts == 0 && rank == 0 && g_tc_prof) {
      atomicAdd(g_tc_prof + 5, (unsigned long long)e_pre);
      atomicAdd(g_tc_prof + 6, (unsigned long long)e_wait);
      atomicAdd(g_tc_prof + 7, (unsigned long long)e_scan);
    }
#endif
#undef EP_MARK
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();       // no remote arrival or peer shared-memory read may outlive either CTA
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// Reduction GEMM on tcgen05 (3xTF32):  C[F1, F2] = sum_n A[n, F1] * B[n, F2]
// -- the weight gradient dW = dY^T X of every nn.Linear on the path (model_lp.py:70-71,124; operations_lp.py
// Linear candidates) and both backward GEMMs of sf_DisMult_op (operations_lp.py:115-127): dq = dl . ent (A given
// K-major: dl is [B, N]) and dent = dl^T . q.  Round 2 had these on cuBLAS fp32 SIMT kernels (0.5 ms of the step).
//
// Both operands have the reduction index as their SLOW dimension, the opposite of what a K-major UMMA operand
// wants, so the 16 producer warps transpose on the way into shared memory: lane = reduction row n, a thread loads
// 8 consecutive features of its row (two 128-bit loads) and scatters them with 32-bit stores into 8 rows of the
// 128B-swizzled K-major tile.  For a fixed tile row the 32 lanes hit word ((n >> 2) ^ (row & 7)) * 4 + (n & 3):
// 32 distinct banks, no conflicts.  CTA tile = 256 x 128 of C (two M = 128 accumulators, 256 TMEM columns), the
// reduction is split over `slices` CTAs per tile so that the grid fills the GPU; partials go to a workspace
// TRANSPOSED ([slice][F2][F1]: TMEM lane = C row, so a warp writes 32 consecutive C rows of one column) and a second
// kernel folds the slices in a fixed order while transposing back -- deterministic.
// ------------------------------------------------------------------------------------------
constexpr int RM = 256, RN = 128, RK = 32, RSTAGES = 2;
constexpr uint32_t RA_TILE = 2 * TILE_BYTES;                 // 256 rows x 128 B
constexpr uint32_t RB_TILE = TILE_BYTES;                     // 128 rows x 128 B
constexpr uint32_t RSTAGE = 2 * RA_TILE + 2 * RB_TILE;       // A hi | A lo | B hi | B lo = 96 KB
constexpr int RWARPS = 16;
constexpr int RTHREADS = RWARPS * 32 + 32;                   // producer / epilogue warps + the MMA warp

struct RedParams {
  const float* A;       // [rows, F1] (lda), or with a_kmajor: At [F1, rows] (lda)
  const float* B;       // [rows, F2] (ldb)
  int64_t lda, ldb, rows;
  int a_kmajor, a_vec, b_vec;   // *_vec: 128-bit loads allowed (16-byte aligned rows, feature count % 8 == 0)
  int F1, F2, m_tiles, n_tiles, slices, cps, nchunks;
  int ones_col;         // 1: B gets a virtual column F2 of ones (column sums of A = the bias gradient) -> F2 + 1 columns
  float* part;          // [slices][n_tiles * RN][m_tiles * RM]
  float* C;             // direct epilogue (slices == 1): C [F1, F2] (ldc) written by the main kernel, no fold pass
  float* colsum;
  const float* bias;    // [F2] added to every row of C (nn.Linear forward), or null
  int64_t ldc;
  int direct;
};

// `ones_at` >= 0: the virtual feature of that index reads as 1 for valid rows (bias-gradient column)
__device__ __forceinline__ void red_load8(const float* row, int f0, int F, bool rvalid, bool vec, float (&v)[8],
                                          int ones_at = -1) {
  if (rvalid && vec && f0 + 8 <= F) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(row + f0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(row + f0 + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (rvalid && f0 + e < F) ? __ldg(row + f0 + e) : ((rvalid && f0 + e == ones_at) ? 1.f : 0.f);
  }
}
// 8 consecutive tile rows (R0 .. R0 + 7, R0 % 8 == 0), column = lane, hi and lo tiles
__device__ __forceinline__ void red_scatter8(uint8_t* hi_tile, uint32_t lo_off, int R0, int lane, const float (&v)[8]) {
  uint8_t* base = hi_tile + (uint32_t)(R0 >> 7) * TILE_BYTES + (uint32_t)((R0 & 127) >> 3) * 1024 + (uint32_t)(lane & 3) * 4;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    uint8_t* a = base + e * 128 + ((((uint32_t)lane >> 2) ^ (uint32_t)e) << 4);
    // lo = the exact fp32 difference; the MMA truncates it to tf32.  (Rounding it here instead changed nothing: 5.1e-6 vs
    // 5.3e-6 of max|C| at 14,541 rows -- the error of this kernel is the TMEM accumulator's, which does not round to nearest
    // and grows with the number of MMAs accumulated per CTA.)
    const float h = tf32_rna(v[e]);
    *reinterpret_cast<float*>(a) = h;
    *reinterpret_cast<float*>(a + lo_off) = v[e] - h;
  }
}

__global__ void __launch_bounds__(RTHREADS, 1) gemm_red_kernel(const RedParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* tail = smem + RSTAGES * RSTAGE;
  uint64_t* full_bar = (uint64_t*)tail;            // [RSTAGES] 16 producer-warp arrivals
  uint64_t* empty_bar = full_bar + RSTAGES;        // [RSTAGES] MMAs that read the stage retired
  uint64_t* tfull_bar = empty_bar + RSTAGES;       // accumulators complete
  uint32_t* tmem_slot = (uint32_t*)(tfull_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int mt = blockIdx.x % p.m_tiles, nt = (blockIdx.x / p.m_tiles) % p.n_tiles, sl = blockIdx.x / (p.m_tiles * p.n_tiles);
  const int m0 = mt * RM, k0 = nt * RN;
  const int c_begin = sl * p.cps, c_end = min(p.nchunks, c_begin + p.cps);
  const int nch = c_end - c_begin;                 // >= 1 by construction of `slices`

  if (threadIdx.x == 0) {
    for (int s = 0; s < RSTAGES; ++s) {
      mbar_init(&full_bar[s], RWARPS);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == RWARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < RWARPS) {
    // ================================ PRODUCERS (transpose into K-major tiles) ================================
    // The loads of chunk ci + 1 are issued before chunk ci is scattered (two register sets, loop unrolled by two): the
    // first version loaded, waited and stored per chunk and ran at 5,600 cycles per chunk against 2,040 of MMA work.
    struct Regs {
      float a0[8], a1[8], b[8];
      float ka[RM / RWARPS];
    };
    auto load = [&](Regs& r, int ci) {
      const int64_t n = (int64_t)(c_begin + ci) * RK + lane;     // reduction row of this lane
      const bool nvalid = n < p.rows;
      if (p.a_kmajor) {          // At[F1][rows]: a warp copies 16 tile rows, lane = reduction index (coalesced)
#pragma unroll
        for (int i = 0; i < RM / RWARPS; ++i) {
          const int f = m0 + warp * (RM / RWARPS) + i;
          r.ka[i] = (nvalid && f < p.F1) ? __ldg(p.A + (size_t)f * p.lda + n) : 0.f;
        }
      } else {
        const float* arow = p.A + (size_t)(nvalid ? n : 0) * p.lda;
        red_load8(arow, m0 + 8 * warp, p.F1, nvalid, p.a_vec != 0, r.a0);
        red_load8(arow, m0 + 8 * (warp + RWARPS), p.F1, nvalid, p.a_vec != 0, r.a1);
      }
      red_load8(p.B + (size_t)(nvalid ? n : 0) * p.ldb, k0 + 8 * warp, p.F2, nvalid, p.b_vec != 0, r.b, p.ones_col ? p.F2 : -1);
    };
    auto store = [&](const Regs& r, int ci) {
      const int s = ci % RSTAGES;
      mbar_wait(&empty_bar[s], ((ci / RSTAGES) & 1) ^ 1);
      uint8_t* st = smem + (size_t)s * RSTAGE;
      if (p.a_kmajor) {
#pragma unroll
        for (int i = 0; i < RM / RWARPS; ++i) {
          const int R = warp * (RM / RWARPS) + i;
          uint8_t* a = st + (uint32_t)(R >> 7) * TILE_BYTES + (uint32_t)((R & 127) >> 3) * 1024 + (uint32_t)(R & 7) * 128 +
                       ((((uint32_t)lane >> 2) ^ (uint32_t)(R & 7)) << 4) + (uint32_t)(lane & 3) * 4;
          const float h = tf32_rna(r.ka[i]);
          *reinterpret_cast<float*>(a) = h;
          *reinterpret_cast<float*>(a + RA_TILE) = r.ka[i] - h;
        }
      } else {
        red_scatter8(st, RA_TILE, 8 * warp, lane, r.a0);
        red_scatter8(st, RA_TILE, 8 * (warp + RWARPS), lane, r.a1);
      }
      red_scatter8(st + 2 * RA_TILE, RB_TILE, 8 * warp, lane, r.b);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    };
    Regs r0, r1;
    load(r0, 0);
    for (int ci = 0; ci < nch; ci += 2) {
      if (ci + 1 < nch) load(r1, ci + 1);
      store(r0, ci);
      if (ci + 1 < nch) {
        if (ci + 2 < nch) load(r0, ci + 2);
        store(r1, ci + 1);
      }
    }
    // ================================ EPILOGUE (same warps): TMEM -> transposed partial ================================
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const int quad = warp & 3, h = (warp >> 2) & 1, chalf = warp >> 3;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + h * RN + chalf * 64;
    const int64_t M2 = (int64_t)p.m_tiles * RM, N2 = (int64_t)p.n_tiles * RN;
    if (p.direct) {
      // one slice: no partials.  TMEM lane = C row, so the warp's 32 x 64 block is transposed through shared memory
      // (the operand stages are free: every MMA has retired) and leaves as 256-byte row segments.
      float* tb = reinterpret_cast<float*>(smem) + (size_t)warp * 32 * 65;
#pragma unroll 1
      for (int cb = 0; cb < 64; cb += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + cb, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) tb[lane * 65 + cb + j] = __uint_as_float(v[j]);
      }
      __syncwarp();
      const int f0 = m0 + h * 128 + quad * 32, kc0 = k0 + chalf * 64;
#pragma unroll 1
      for (int rr = 0; rr < 32; ++rr) {
        const int f = f0 + rr;
        if (f >= p.F1) break;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int k = kc0 + lane + 32 * t;
          const float val = tb[rr * 65 + lane + 32 * t];
          if (k < p.F2) p.C[(int64_t)f * p.ldc + k] = p.bias ? val + __ldg(p.bias + k) : val;
          else if (k == p.F2 && p.colsum) p.colsum[f] = val;
        }
      }
    } else {
      float* o = p.part + ((int64_t)sl * N2 + k0 + chalf * 64) * M2 + m0 + h * 128 + quad * 32 + lane;
#pragma unroll 1
      for (int cb = 0; cb < 64; cb += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + cb, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) o[(int64_t)(cb + j) * M2] = __uint_as_float(v[j]);
      }
    }
  } else {
    // ================================ MMA ISSUER ================================
    const uint32_t idesc = umma_idesc(128, RN);
    for (int ci = 0; ci < nch; ++ci) {
      const int s = ci % RSTAGES;
      mbar_wait(&full_bar[s], (ci / RSTAGES) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t st = smem_u32(smem + (size_t)s * RSTAGE);
        const uint32_t bhi = st + 2 * RA_TILE, blo = bhi + RB_TILE;
        for (int h = 0; h < 2; ++h) {
          const uint32_t ahi = st + h * TILE_BYTES, alo = ahi + RA_TILE;
          const uint32_t d_tmem = tmem_base + h * RN;
          for (int k = 0; k < RK / 8; ++k) {
            const uint32_t ko = k * 32;
            const uint32_t acc = (ci > 0 || k > 0) ? 1u : 0u;
            umma_tf32(d_tmem, umma_desc(alo + ko), umma_desc(bhi + ko), idesc, acc);   // small terms first
            umma_tf32(d_tmem, umma_desc(ahi + ko), umma_desc(blo + ko), idesc, 1u);
            umma_tf32(d_tmem, umma_desc(ahi + ko), umma_desc(bhi + ko), idesc, 1u);
          }
        }
        umma_commit(&empty_bar[s]);
        if (ci == nch - 1) umma_commit(tfull_bar);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == RWARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// C[f, k] = sum_s part[s][k][f]  (fixed slice order; 32 x 32 tiles transposed through shared memory)
// colsum (optional): column F2 of the partials = sum_n A[n, f] (the virtual ones column of B)
__global__ void __launch_bounds__(256) gemm_red_fold_kernel(const float* __restrict__ part, int slices, int64_t N2, int64_t M2,
                                                            int F1, int F2, float* __restrict__ C, int64_t ldc,
                                                            float* __restrict__ colsum, const float* __restrict__ bias) {
  if (colsum && blockIdx.y == 0 && threadIdx.x < 32) {
    const int f = blockIdx.x * 32 + threadIdx.x;
    if (f < F1) {
      float acc = 0.f;
      for (int s0 = 0; s0 < slices; s0 += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = s0 + u < slices ? __ldg(part + ((int64_t)(s0 + u) * N2 + F2) * M2 + f) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
      }
      colsum[f] = acc;
    }
  }
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ft = blockIdx.x * 32, kt = blockIdx.y * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int kk = ty + 8 * i;
    float acc = 0.f;
    if (kt + kk < F2 && ft + tx < F1) {
      const float* src = part + ((int64_t)kt + kk) * M2 + ft + tx;
      for (int s0 = 0; s0 < slices; s0 += 8) {      // 8 independent loads in flight, added in the fixed slice order
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = s0 + u < slices ? __ldg(src + (int64_t)(s0 + u) * N2 * M2) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
      }
    }
    tile[kk][tx] = acc;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ff = ty + 8 * i;
    if (ft + ff < F1 && kt + tx < F2) C[(int64_t)(ft + ff) * ldc + kt + tx] = tile[tx][ff] + (bias ? __ldg(bias + kt + tx) : 0.f);
  }
}

}  // namespace tc
}  // namespace mrg

using namespace mrg;

extern "C" size_t mrg_amax_tc_workspace_bytes(int64_t N, int32_t D) {
  const int MH = D <= 128 ? 1 : 2;
  const int Kp = (D + tc::KCH - 1) / tc::KCH * tc::KCH;
  return (size_t)N * D * sizeof(unsigned long long) + (size_t)2 * MH * 128 * Kp * sizeof(float) + 1024;
}

extern "C" int mrg_amax_tc_supported(int32_t D) { return (D % 8 == 0 && D >= 8 && D <= 256) ? 1 : 0; }

static int amax_tc_fwd_impl(int prec, mrg_act x, const float* W, const float* bias, const int32_t* csr_eid,
                            const int32_t* dst, int64_t E, int64_t N, int32_t D, mrg_act residual, float* out,
                            int32_t* arg, void* workspace, size_t workspace_bytes, void* stream) {
  MRG_CHECK_ARG(x.data && W && out && workspace, "amax_tc_fwd: null pointer");
  MRG_CHECK_ARG(E == 0 || (csr_eid && dst), "amax_tc_fwd: null graph arrays");
  MRG_CHECK_ARG(mrg_amax_tc_supported(D), "amax_tc_fwd: D must be a multiple of 8 and <= 256");
  if (workspace_bytes < mrg_amax_tc_workspace_bytes(N, D)) {
    set_error("amax_tc_fwd: workspace too small");
    return MRG_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int MH = D <= 128 ? 1 : 2;
  const int kch = prec == 2 ? 64 : tc::KCH;
  const int Kp = (D + kch - 1) / kch * kch;
  unsigned long long* packed = (unsigned long long*)workspace;
  const size_t packed_bytes = ((size_t)N * D * sizeof(unsigned long long) + 255) / 256 * 256;
  float* wsplit = (float*)((char*)workspace + packed_bytes);
  cudaError_t e = cudaMemsetAsync(packed, 0, (size_t)N * D * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return cuda_fail(e, "amax_tc_fwd memset");
  if (prec == 2) tc::bf16_image_kernel<<<64, 256, 0, st>>>(W, D, D, MH, Kp / kch, (unsigned short*)wsplit);
  else tc::tf32_split_kernel<<<64, 256, 0, st>>>(W, D, D, MH, Kp / kch, wsplit);
  tc::AmaxParams p;
  p.x = x;
  p.csr_eid = csr_eid;
  p.dst = dst;
  p.wimg = wsplit;
  p.bias = bias;
  p.packed = packed;
  p.E = E;
  p.D = D;
  p.MH = MH;
  p.Kp = Kp;
  p.nchunks = Kp / kch;
  p.num_tiles = (int)((E + tc::TILE_E - 1) / tc::TILE_E);
  p.Dout = D;
  p.label = nullptr;
  p.logit = nullptr;
  p.loss_partial = nullptr;
  p.ldl = 0;
  p.store = nullptr;
  p.lds = 0;
  p.l2_prefetch = 0;
  // CTA-pair kernel (cta_group::2) for the 3xTF32 form at 128 < D <= 256; MRG_AMAX_PAIR=0 selects the single-CTA kernel
  static const bool use_pair = [] {
    const char* v = getenv("MRG_AMAX_PAIR");
    return !(v && v[0] == '0');
  }();
  if (MH == 2 && use_pair && E > 0) {
    p.num_tiles = (int)((E + tc::TILE2 - 1) / tc::TILE2);
    const size_t smem = (size_t)(tc::tc2_xstages(prec) + tc::tc2_wstages(prec)) * tc::tc2_stage_bytes(prec) + 1024 /*align*/ +
                        8192 /*tail*/;
    int grid = 2 * p.num_tiles < kNumSMs ? 2 * p.num_tiles : kNumSMs;
    grid &= ~1;
    static const int pf = [] {
      const char* v = getenv("MRG_TC2_PREFETCH");
      return v ? atoi(v) : 1;
    }();
    p.l2_prefetch = pf;
    if (prec == 2) {
      e = cudaFuncSetAttribute(tc::amax_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "amax_tc_fwd smem attr");
      tc::amax_tc2_kernel<2><<<grid, tc::THREADS, smem, st>>>(p);
    } else {
      e = cudaFuncSetAttribute(tc::amax_tc2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "amax_tc_fwd smem attr");
      tc::amax_tc2_kernel<0><<<grid, tc::THREADS, smem, st>>>(p);
    }
  } else if (p.num_tiles > 0) {
    const size_t smem = prec == 2
        ? (size_t)3 * MH * tc::TILE_BYTES + (size_t)6 * tc::TILE_BYTES + 1024 + 8192
        : (size_t)tc::STAGES * 2 * MH * tc::TILE_BYTES + (size_t)(MH == 1 ? 4 : 2) * 2 * tc::TILE_BYTES +
              1024 /*align*/ + 8192 /*tail*/;
    const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
    if (prec == 2 && MH == 1) {
      e = cudaFuncSetAttribute(tc::amax_tc_kernel<1, tc::EPI_AMAX, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "amax_tc_fwd smem attr");
      tc::amax_tc_kernel<1, tc::EPI_AMAX, 2><<<grid, tc::THREADS, smem, st>>>(p);
    } else if (prec == 2) {
      e = cudaFuncSetAttribute(tc::amax_tc_kernel<2, tc::EPI_AMAX, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "amax_tc_fwd smem attr");
      tc::amax_tc_kernel<2, tc::EPI_AMAX, 2><<<grid, tc::THREADS, smem, st>>>(p);
    } else if (MH == 1) {
      e = cudaFuncSetAttribute(tc::amax_tc_kernel<1, tc::EPI_AMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "amax_tc_fwd smem attr");
      tc::amax_tc_kernel<1, tc::EPI_AMAX><<<grid, tc::THREADS, smem, st>>>(p);
    } else {
      e = cudaFuncSetAttribute(tc::amax_tc_kernel<2, tc::EPI_AMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "amax_tc_fwd smem attr");
      tc::amax_tc_kernel<2, tc::EPI_AMAX><<<grid, tc::THREADS, smem, st>>>(p);
    }
  }
  const int64_t n = N * D;
  tc::amax_finalize_kernel<<<(int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096), 256, 0, st>>>(packed, n, D, residual,
                                                                                                  out, arg);
  MRG_LAUNCH_CHECK("amax_tc_fwd");
  return MRG_OK;
}

#ifdef MRG_TC_PROF
extern "C" int mrg_debug_set_tc_prof(unsigned long long* dev_buf) {
  cudaMemcpyToSymbol(tc::g_tc_prof, &dev_buf, sizeof(dev_buf));
  return MRG_OK;
}
#endif

extern "C" int mrg_amax_tc_fwd(mrg_act x, const float* W, const float* bias, const int32_t* csr_eid,
                               const int32_t* dst, int64_t E, int64_t N, int32_t D, mrg_act residual, float* out,
                               int32_t* arg, void* workspace, size_t workspace_bytes, void* stream) {
  return amax_tc_fwd_impl(0, x, W, bias, csr_eid, dst, E, N, D, residual, out, arg, workspace, workspace_bytes, stream);
}

/* bf16 variant: operands (activated x rows, W) rounded to bf16, fp32 accumulation in TMEM, one MMA per K step */
extern "C" int mrg_amax_tc_fwd_bf16(mrg_act x, const float* W, const float* bias, const int32_t* csr_eid,
                                    const int32_t* dst, int64_t E, int64_t N, int32_t D, mrg_act residual, float* out,
                                    int32_t* arg, void* workspace, size_t workspace_bytes, void* stream) {
  return amax_tc_fwd_impl(2, x, W, bias, csr_eid, dst, E, N, D, residual, out, arg, workspace, workspace_bytes, stream);
}

// ------------------------------------------------------------------------------------------
// K7+K8 fused: DistMult 1-N scoring + sigmoid + BCE on the same tcgen05 main loop.
//   logit[b, n] = sum_k query[b,k] * ent[n,k]   (3xTF32, fp32-class accuracy),  loss = mean BCE(sigmoid(logit), label)
// "W" of the main loop is the [B<=256, D] query block (pre-split hi/lo image), the tile rows are entities.
// ------------------------------------------------------------------------------------------
__global__ void distmult_loss_finalize_kernel(const double* __restrict__ partial, int nparts, double n, float* loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int p = 0; p < nparts; ++p) t += partial[p];
    loss[0] = (float)(t / n);
  }
}

extern "C" int mrg_distmult_bce_supported(int32_t D) { return mrg_amax_tc_supported(D); }
extern "C" int32_t mrg_distmult_bce_nparts(int64_t B) { return (int32_t)((B + 255) / 256) * kNumSMs; }
extern "C" size_t mrg_distmult_bce_workspace_bytes(int32_t D) {
  const int Kp = (D + tc::KCH - 1) / tc::KCH * tc::KCH;
  return (size_t)2 * 2 * 128 * Kp * sizeof(float) + 1024;
}

extern "C" int mrg_distmult_bce_fwd(const float* query, const float* ent, const float* label, int64_t B, int64_t N,
                                    int32_t D, float* logit, double* partial, float* loss, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  MRG_CHECK_ARG(query && ent && label && logit && partial && loss && workspace, "distmult_bce_fwd: null pointer");
  MRG_CHECK_ARG(B > 0 && N > 0, "distmult_bce_fwd: sizes");
  MRG_CHECK_ARG(mrg_distmult_bce_supported(D), "distmult_bce_fwd: D must be a multiple of 8 and <= 256");
  if (workspace_bytes < mrg_distmult_bce_workspace_bytes(D)) {
    set_error("distmult_bce_fwd: workspace too small");
    return MRG_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int Kp = (D + tc::KCH - 1) / tc::KCH * tc::KCH;
  const int nparts = mrg_distmult_bce_nparts(B);
  cudaError_t e = cudaMemsetAsync(partial, 0, (size_t)nparts * sizeof(double), st);
  if (e != cudaSuccess) return cuda_fail(e, "distmult_bce_fwd memset");
  int chunk = 0;
  for (int64_t b0 = 0; b0 < B; b0 += 256, ++chunk) {
    const int rows = (int)(B - b0 < 256 ? B - b0 : 256);
    const int MH = rows <= 128 ? 1 : 2;
    tc::tf32_split_kernel<<<64, 256, 0, st>>>(query + (size_t)b0 * D, rows, D, MH, Kp / tc::KCH, (float*)workspace);
    tc::AmaxParams p;
    p.x.data = ent; p.x.scale = nullptr; p.x.shift = nullptr; p.x.relu = 0;
    p.csr_eid = nullptr; p.dst = nullptr;
    p.wimg = (const float*)workspace;
    p.bias = nullptr; p.packed = nullptr;
    p.E = N; p.D = D; p.MH = MH; p.Kp = Kp; p.nchunks = Kp / tc::KCH;
    p.num_tiles = (int)((N + tc::TILE_E - 1) / tc::TILE_E);
    p.Dout = rows;
    p.label = label + (size_t)b0 * N;
    p.logit = logit + (size_t)b0 * N;
    p.loss_partial = partial + (size_t)chunk * kNumSMs;
    p.ldl = N;
    p.store = nullptr;
    p.lds = 0;
    const size_t smem = (size_t)tc::STAGES * 2 * MH * tc::TILE_BYTES + (size_t)(MH == 1 ? 4 : 2) * 2 * tc::TILE_BYTES +
                        1024 /*align*/ + 8192 /*tail*/ + 8 * 32 * 17 * sizeof(float) /*epilogue transpose tiles*/;
    const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
    if (MH == 1) {
      e = cudaFuncSetAttribute(tc::amax_tc_kernel<1, tc::EPI_DISTMULT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "distmult_bce_fwd smem attr");
      tc::amax_tc_kernel<1, tc::EPI_DISTMULT><<<grid, tc::THREADS, smem, st>>>(p);
    } else {
      e = cudaFuncSetAttribute(tc::amax_tc_kernel<2, tc::EPI_DISTMULT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "distmult_bce_fwd smem attr");
      tc::amax_tc_kernel<2, tc::EPI_DISTMULT><<<grid, tc::THREADS, smem, st>>>(p);
    }
  }
  distmult_loss_finalize_kernel<<<1, 32, 0, st>>>(partial, nparts, (double)B * (double)N, loss);
  MRG_LAUNCH_CHECK("distmult_bce_fwd");
  return MRG_OK;
}

// ------------------------------------------------------------------------------------------
// Node-level Linear on the same main loop: out[r, f] = sum_k x[r,k] * W[f,k] + bias[f]  (3xTF32, fp32-class
// accuracy).  Replaces the cuBLAS fp32 SIMT GEMMs of nn.Linear on the path (model_lp.py:70-71 `concat`,
// :124 `linear_e`) and, called with W^T, their input gradients.  256 output features per launch.
// ------------------------------------------------------------------------------------------
extern "C" int mrg_linear_tc_supported(int32_t K) { return (K > 0 && K % 8 == 0) ? 1 : 0; }
extern "C" size_t mrg_linear_tc_workspace_bytes(int32_t K) {
  const int Kp = (K + tc::KCH - 1) / tc::KCH * tc::KCH;
  return (size_t)2 * 2 * 128 * Kp * sizeof(float) + 1024;
}

extern "C" int mrg_linear_tc_fwd(const float* x, const float* W, const float* bias, int64_t rows, int32_t K, int32_t F,
                                 float* out, int64_t ldo, void* workspace, size_t workspace_bytes, void* stream) {
  MRG_CHECK_ARG(x && W && out && workspace, "linear_tc_fwd: null pointer");
  MRG_CHECK_ARG(rows >= 0 && F > 0 && ldo >= F, "linear_tc_fwd: sizes");
  MRG_CHECK_ARG(mrg_linear_tc_supported(K), "linear_tc_fwd: K must be a positive multiple of 8");
  if (workspace_bytes < mrg_linear_tc_workspace_bytes(K)) {
    set_error("linear_tc_fwd: workspace too small");
    return MRG_ERR_WORKSPACE;
  }
  if (rows == 0) return MRG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int Kp = (K + tc::KCH - 1) / tc::KCH * tc::KCH;
  cudaError_t e;
  for (int f0 = 0; f0 < F; f0 += 256) {
    const int nf = F - f0 < 256 ? F - f0 : 256;
    const int MH = nf <= 128 ? 1 : 2;
    tc::tf32_split_kernel<<<64, 256, 0, st>>>(W + (size_t)f0 * K, nf, K, MH, Kp / tc::KCH, (float*)workspace);
    tc::AmaxParams p;
    p.x.data = x; p.x.scale = nullptr; p.x.shift = nullptr; p.x.relu = 0;
    p.csr_eid = nullptr; p.dst = nullptr;
    p.wimg = (const float*)workspace;
    p.bias = bias ? bias + f0 : nullptr;
    p.packed = nullptr;
    p.E = rows; p.D = K; p.MH = MH; p.Kp = Kp; p.nchunks = Kp / tc::KCH;
    p.num_tiles = (int)((rows + tc::TILE_E - 1) / tc::TILE_E);
    p.Dout = nf;
    p.label = nullptr; p.logit = nullptr; p.loss_partial = nullptr; p.ldl = 0;
    p.store = out + f0;
    p.lds = ldo;
    const size_t smem = (size_t)tc::STAGES * 2 * MH * tc::TILE_BYTES + (size_t)(MH == 1 ? 4 : 2) * 2 * tc::TILE_BYTES +
                        1024 /*align*/ + 8192 /*tail*/;
    const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
    if (MH == 1) {
      e = cudaFuncSetAttribute(tc::amax_tc_kernel<1, tc::EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "linear_tc_fwd smem attr");
      tc::amax_tc_kernel<1, tc::EPI_STORE><<<grid, tc::THREADS, smem, st>>>(p);
    } else {
      e = cudaFuncSetAttribute(tc::amax_tc_kernel<2, tc::EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "linear_tc_fwd smem attr");
      tc::amax_tc_kernel<2, tc::EPI_STORE><<<grid, tc::THREADS, smem, st>>>(p);
    }
  }
  MRG_LAUNCH_CHECK("linear_tc_fwd");
  return MRG_OK;
}

// ------------------------------------------------------------------------------------------
// Reduction GEMM  C[F1, F2] = sum_n A[n, F1] * B[n, F2]   (tc::gemm_red_kernel)
// ------------------------------------------------------------------------------------------
static void gemm_red_plan(int64_t rows, int32_t F1, int32_t F2, int* m_tiles, int* n_tiles, int* slices, int* cps,
                          int* nchunks) {
  *m_tiles = (F1 + tc::RM - 1) / tc::RM;
  *n_tiles = (F2 + 1 + tc::RN - 1) / tc::RN;      // room for the virtual ones column (bias gradient)
  *nchunks = (int)((rows + tc::RK - 1) / tc::RK);
  int want = kNumSMs / (*m_tiles * *n_tiles);
  if (want < 1) want = 1;
  if (want > *nchunks) want = *nchunks;
  *cps = (*nchunks + want - 1) / want;
  *slices = (*nchunks + *cps - 1) / *cps;
}

extern "C" size_t mrg_gemm_red_workspace_bytes(int64_t rows, int32_t F1, int32_t F2) {
  if (rows <= 0 || F1 <= 0 || F2 <= 0) return 0;
  int mt, nt, sl, cps, nch;
  gemm_red_plan(rows, F1, F2, &mt, &nt, &sl, &cps, &nch);
  if (sl == 1) return 16;      // one slice: the main kernel writes C itself (transposed through shared memory), no partials
  return (size_t)sl * nt * tc::RN * mt * tc::RM * sizeof(float);
}

extern "C" int mrg_gemm_red(const float* A, int64_t lda, int32_t a_kmajor, const float* B, int64_t ldb, int64_t rows,
                            int32_t F1, int32_t F2, float* C, int64_t ldc, float* colsum, const float* bias,
                            void* workspace, size_t workspace_bytes, void* stream) {
  MRG_CHECK_ARG(F1 > 0 && F2 > 0 && rows >= 0 && C, "gemm_red: bad shape");
  MRG_CHECK_ARG(ldc >= F2, "gemm_red: ldc < F2");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {
    cudaError_t e0 = cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)F2 * sizeof(float), F1, st);
    if (e0 != cudaSuccess) return cuda_fail(e0, "gemm_red memset");
    if (colsum) e0 = cudaMemsetAsync(colsum, 0, (size_t)F1 * sizeof(float), st);
    if (e0 != cudaSuccess) return cuda_fail(e0, "gemm_red memset");
    return MRG_OK;
  }
  MRG_CHECK_ARG(A && B && workspace, "gemm_red: null pointer");
  MRG_CHECK_ARG(a_kmajor ? lda >= rows : lda >= F1, "gemm_red: lda too small");
  MRG_CHECK_ARG(ldb >= F2, "gemm_red: ldb < F2");
  if (workspace_bytes < mrg_gemm_red_workspace_bytes(rows, F1, F2)) {
    set_error("gemm_red: workspace too small");
    return MRG_ERR_WORKSPACE;
  }
  tc::RedParams p;
  p.A = A;
  p.B = B;
  p.lda = lda;
  p.ldb = ldb;
  p.rows = rows;
  p.a_kmajor = a_kmajor ? 1 : 0;
  p.a_vec = (!a_kmajor && lda % 4 == 0 && ((uintptr_t)A & 15) == 0) ? 1 : 0;
  p.b_vec = (ldb % 4 == 0 && ((uintptr_t)B & 15) == 0) ? 1 : 0;
  p.F1 = F1;
  p.F2 = F2;
  p.ones_col = colsum ? 1 : 0;
  gemm_red_plan(rows, F1, F2, &p.m_tiles, &p.n_tiles, &p.slices, &p.cps, &p.nchunks);
  p.part = (float*)workspace;
  p.C = C;
  p.colsum = colsum;
  p.bias = bias;
  p.ldc = ldc;
  p.direct = p.slices == 1 ? 1 : 0;
  const size_t smem = (size_t)tc::RSTAGES * tc::RSTAGE + 1024 /*align*/ + 256 /*tail*/;
  cudaError_t e = cudaFuncSetAttribute(tc::gemm_red_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "gemm_red smem attr");
  tc::gemm_red_kernel<<<p.m_tiles * p.n_tiles * p.slices, tc::RTHREADS, smem, st>>>(p);
  if (!p.direct) {
    dim3 fg((F1 + 31) / 32, (F2 + 31) / 32);
    tc::gemm_red_fold_kernel<<<fg, 256, 0, st>>>(p.part, p.slices, (int64_t)p.n_tiles * tc::RN, (int64_t)p.m_tiles * tc::RM,
                                                 F1, F2, C, ldc, colsum, bias);
  }
  MRG_LAUNCH_CHECK("gemm_red");
  return MRG_OK;
}
