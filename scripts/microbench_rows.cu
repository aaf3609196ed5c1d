// Micro-benchmark that decided the structure of the streaming row kernels:
// how close to the measured HBM peak does a "2 reads + 1 write per row, one warp-wide dot product per row"
// kernel get on B200 as a function of how the loads are put in flight?
//   v0  flat float4 elementwise (no row structure): upper bound for 2R+1W
//   v1  warp per row, loads straight into registers (the round-1 row kernels), CTAs/SM = 2..6
//   v2  warp per row, two rows in flight per warp
//   v3  TMA bulk-copy ring: one elected thread streams R-row tiles of both inputs into shared memory
//       (cp.async.bulk + mbarrier expect_tx, S stages), 8 consumer warps read rows from smem
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o microbench_rows microbench_rows.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void v0_flat(const float* a, const float* b, float* c, int64_t n4) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    float4 x = ld_stream4(a + 4 * i), y = ld_stream4(b + 4 * i);
    st_stream4(c + 4 * i, make_float4(x.x * 0.5f + y.x, x.y * 0.5f + y.y, x.z * 0.5f + y.z, x.w * 0.5f + y.w));
  }
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) v1_rows(const float* a, const float* b, float* c, int64_t rows, int D) {
  const int lane = threadIdx.x & 31, D4 = D >> 2;
  const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * 8;
  for (int64_t r = w0; r < rows; r += nw) {
    float4 x[2], y[2];
    float dot = 0.f;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      int c4 = lane + 32 * v;
      x[v] = y[v] = make_float4(0, 0, 0, 0);
      if (c4 < D4) {
        x[v] = ld_stream4(a + r * D + 4 * c4);
        y[v] = ld_stream4(b + r * D + 4 * c4);
        dot += x[v].x * y[v].x + x[v].y * y[v].y + x[v].z * y[v].z + x[v].w * y[v].w;
      }
    }
    dot = warp_sum(dot) * 1e-3f;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      int c4 = lane + 32 * v;
      if (c4 < D4)
        st_stream4(c + r * D + 4 * c4, make_float4(dot * x[v].x + y[v].x, dot * x[v].y + y[v].y, dot * x[v].z + y[v].z, dot * x[v].w + y[v].w));
    }
  }
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) v2_rows2(const float* a, const float* b, float* c, int64_t rows, int D) {
  const int lane = threadIdx.x & 31, D4 = D >> 2;
  const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * 8;
  for (int64_t r0 = 2 * w0; r0 < rows; r0 += 2 * nw) {
    float4 x[2][2], y[2][2];
    float dot[2] = {0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        int c4 = lane + 32 * v;
        x[u][v] = y[u][v] = make_float4(0, 0, 0, 0);
        if (c4 < D4 && r0 + u < rows) {
          x[u][v] = ld_stream4(a + (r0 + u) * D + 4 * c4);
          y[u][v] = ld_stream4(b + (r0 + u) * D + 4 * c4);
        }
      }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int v = 0; v < 2; ++v)
        dot[u] += x[u][v].x * y[u][v].x + x[u][v].y * y[u][v].y + x[u][v].z * y[u][v].z + x[u][v].w * y[u][v].w;
      dot[u] = warp_sum(dot[u]) * 1e-3f;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        int c4 = lane + 32 * v;
        if (c4 < D4 && r0 + u < rows)
          st_stream4(c + (r0 + u) * D + 4 * c4, make_float4(dot[u] * x[u][v].x + y[u][v].x, dot[u] * x[u][v].y + y[u][v].y,
                                                            dot[u] * x[u][v].z + y[u][v].z, dot[u] * x[u][v].w + y[u][v].w));
      }
  }
}

// ---- v3: TMA bulk ring
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
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t it = 0; it < (1u << 26) && !ok; ++it)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  if (!ok) __trap();
}

template <int R, int S>
__global__ void __launch_bounds__(288, 1) v3_tma(const float* a, const float* b, float* c, int64_t rows, int D) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t tile_bytes = R * D * 4;
  float* ta = reinterpret_cast<float*>(smem);
  float* tb = ta + (size_t)S * R * D;
  uint64_t* full = reinterpret_cast<uint64_t*>(tb + (size_t)S * R * D);
  uint64_t* empty = full + S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, D4 = D >> 2;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t ntiles = (rows + R - 1) / R;
  if (warp == 8) {
    if (lane == 0) {
      int it = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int s = it % S;
        if (it >= S) mbar_wait(empty + s, ((it / S) - 1) & 1);
        const int64_t r0 = t * R;
        const uint32_t bytes = (uint32_t)(min((int64_t)R, rows - r0) * D * 4);
        mbar_expect_tx(full + s, 2 * bytes);
        bulk_g2s(smem_u32(ta + (size_t)s * R * D), a + r0 * D, bytes, full + s);
        bulk_g2s(smem_u32(tb + (size_t)s * R * D), b + r0 * D, bytes, full + s);
      }
    }
    return;
  }
  int it = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int s = it % S;
    mbar_wait(full + s, (it / S) & 1);
    const int64_t r0 = t * R;
    const float* sa = ta + (size_t)s * R * D;
    const float* sb = tb + (size_t)s * R * D;
#pragma unroll
    for (int rr = warp; rr < R; rr += 8) {
      if (r0 + rr < rows) {
        float4 x[2], y[2];
        float dot = 0.f;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          int c4 = lane + 32 * v;
          x[v] = y[v] = make_float4(0, 0, 0, 0);
          if (c4 < D4) {
            x[v] = *reinterpret_cast<const float4*>(sa + rr * D + 4 * c4);
            y[v] = *reinterpret_cast<const float4*>(sb + rr * D + 4 * c4);
            dot += x[v].x * y[v].x + x[v].y * y[v].y + x[v].z * y[v].z + x[v].w * y[v].w;
          }
        }
        dot = warp_sum(dot) * 1e-3f;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          int c4 = lane + 32 * v;
          if (c4 < D4)
            st_stream4(c + (r0 + rr) * D + 4 * c4, make_float4(dot * x[v].x + y[v].x, dot * x[v].y + y[v].y, dot * x[v].z + y[v].z, dot * x[v].w + y[v].w));
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);
  }
  (void)tile_bytes;
}

template <typename F>
float time_it(F f, int iters = 10) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(e1);
  CK(cudaEventSynchronize(e1));
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / iters;
}

int main(int argc, char** argv) {
  const int64_t rows = argc > 1 ? atoll(argv[1]) : 558771;
  const int D = argc > 2 ? atoi(argv[2]) : 200;
  const size_t n = (size_t)rows * D;
  float *a, *b, *c;
  CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&c, n * 4));
  CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4));
  const double gb = 3.0 * n * 4 / 1e9;
  auto report = [&](const char* name, float ms) { printf("%-28s %8.3f ms  %8.1f GB/s\n", name, ms, gb / (ms * 1e-3)); };
  report("memcpy d2d (1R+1W) x1.5", time_it([&] { cudaMemcpyAsync(c, a, n * 4, cudaMemcpyDeviceToDevice); }) * 1.5f);
  report("v0 flat", time_it([&] { v0_flat<<<(unsigned)((n / 4 + 255) / 256), 256>>>(a, b, c, n / 4); }));
  report("v1 rows 2 CTA/SM", time_it([&] { v1_rows<2><<<148 * 2, 256>>>(a, b, c, rows, D); }));
  report("v1 rows 4 CTA/SM", time_it([&] { v1_rows<4><<<148 * 4, 256>>>(a, b, c, rows, D); }));
  report("v1 rows 6 CTA/SM", time_it([&] { v1_rows<6><<<148 * 6, 256>>>(a, b, c, rows, D); }));
  report("v1 rows 8 CTA/SM", time_it([&] { v1_rows<8><<<148 * 8, 256>>>(a, b, c, rows, D); }));
  report("v1 rows nonpersist", time_it([&] { v1_rows<6><<<(unsigned)((rows + 7) / 8), 256>>>(a, b, c, rows, D); }));
  report("v2 rows2 2 CTA/SM", time_it([&] { v2_rows2<2><<<148 * 2, 256>>>(a, b, c, rows, D); }));
  report("v2 rows2 4 CTA/SM", time_it([&] { v2_rows2<4><<<148 * 4, 256>>>(a, b, c, rows, D); }));
#define V3(R, S)                                                                                                  \
  {                                                                                                               \
    size_t sm = (size_t)2 * S * R * D * 4 + 2 * S * 8 + 128;                                                      \
    CK(cudaFuncSetAttribute(v3_tma<R, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));                 \
    char nm[64];                                                                                                  \
    snprintf(nm, 64, "v3 tma R=%d S=%d (%zuKB)", R, S, sm / 1024);                                                \
    report(nm, time_it([&] { v3_tma<R, S><<<148, 288, sm>>>(a, b, c, rows, D); }));                              \
  }
  V3(8, 4) V3(8, 8) V3(16, 4) V3(16, 6) V3(32, 3)
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
