// K0: integer graph build on the device -- COO (edge-id order) -> in-degrees, degree norms,
// dst-CSR, src-CSC and relation segments over the M = E + N edge-expanded rows.  Stable LSD
// radix sorts (cub) keep ids ascending inside every segment, which is what makes the max
// aggregator's tie-break ("lowest edge id") and every segmented sum order deterministic.
#include <cub/cub.cuh>

#include "common.cuh"

namespace mrg {

// self_const: key of every self-loop row when use_const, else the offset added to the self row's local index
__global__ void iota_keys_kernel(const int32_t* __restrict__ key_edges, int64_t E, int64_t M, int32_t self_const,
                                 int use_const, int32_t* __restrict__ keys, int32_t* __restrict__ vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  keys[i] = i < E ? key_edges[i] : (use_const ? self_const : self_const + (int32_t)(i - E));
  vals[i] = (int32_t)i;
}

// ptr[s] = lower_bound(sorted_keys, s) for s in [0, nseg]
__global__ void lower_bound_kernel(const int32_t* __restrict__ sorted, int64_t n, int64_t nseg,
                                   int32_t* __restrict__ ptr) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s > nseg) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)sorted[mid] < s) lo = mid + 1;
    else hi = mid;
  }
  ptr[s] = (int32_t)lo;
}

__global__ void degree_norm_kernel(const int32_t* __restrict__ csr_ptr, int64_t N, int32_t* __restrict__ in_deg,
                                   float* __restrict__ n_norm) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int32_t d = csr_ptr[n + 1] - csr_ptr[n];
  if (in_deg) in_deg[n] = d;
  // float32(in_deg) ** -0.5 with inf -> 0 (mr_lp_train.py:82-84); correctly rounded via double
  if (n_norm) n_norm[n] = d > 0 ? (float)(1.0 / sqrt((double)d)) : 0.f;
}

__global__ void edge_norm_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                 const float* __restrict__ n_norm, int64_t E, float* __restrict__ edge_norm) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  edge_norm[e] = n_norm[dst[e]] * n_norm[src[e]];
}

static inline int bits_for(int64_t n) {
  int b = 1;
  while (((int64_t)1 << b) < n && b < 31) ++b;
  return b;
}

static size_t sort_temp_bytes(int64_t n) {
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs((void*)nullptr, tmp, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  return tmp;
}

static inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace mrg

using namespace mrg;

extern "C" size_t mrg_graph_workspace_bytes(int64_t E, int64_t N, int64_t n_rel_rows) {
  (void)n_rel_rows;
  const int64_t M = E + N;
  return 3 * align256((size_t)M * sizeof(int32_t)) + align256(sort_temp_bytes(M)) + 256;
}

// General form: n_src source nodes (gather tables have n_src rows), n_dst destination nodes owned by this graph
// (dst ids local, 0..n_dst), self-loop row E+j has source self_off+j.  The full graph is n_src == n_dst, self_off 0.
static int graph_build_impl(const int32_t* src, const int32_t* dst, const int32_t* etype, int64_t E, int64_t n_src,
                            int64_t n_dst, int64_t self_off, int64_t n_rel_rows, int32_t* in_deg, float* n_norm,
                            float* edge_norm, int32_t* csr_ptr, int32_t* csr_eid, int32_t* csc_ptr, int32_t* csc_row,
                            int32_t* rel_ptr, int32_t* rel_row, void* workspace, size_t workspace_bytes,
                            void* stream) {
  MRG_CHECK_ARG(E >= 0 && n_src > 0 && n_dst > 0 && n_rel_rows > 0, "graph_build: sizes");
  MRG_CHECK_ARG(self_off >= 0 && self_off + n_dst <= n_src, "graph_build: owned destination range outside the sources");
  MRG_CHECK_ARG(E == 0 || (src && dst), "graph_build: null src/dst");
  MRG_CHECK_ARG(csr_ptr && csr_eid && workspace, "graph_build: null output");
  MRG_CHECK_ARG(E + n_src < ((int64_t)1 << 31), "graph_build: E + N must fit int32");
  MRG_CHECK_ARG(!edge_norm || n_norm, "graph_build: edge_norm needs n_norm");
  MRG_CHECK_ARG(!(n_norm || edge_norm) || n_src == n_dst, "graph_build: degree norms need the full graph");
  if (workspace_bytes < mrg_graph_workspace_bytes(E, n_dst, n_rel_rows)) {
    set_error("graph_build: workspace too small");
    return MRG_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t M = E + n_dst;
  const size_t arr = align256((size_t)M * sizeof(int32_t));
  char* w = (char*)workspace;
  int32_t* keys_in = (int32_t*)w;
  int32_t* vals_in = (int32_t*)(w + arr);
  int32_t* keys_out = (int32_t*)(w + 2 * arr);
  void* tmp = w + 3 * arr;
  size_t tmp_bytes = workspace_bytes - 3 * arr;
  const int T = 256;
  cudaError_t e;

  // dst-CSR over the E edges
  if (E > 0) {
    iota_keys_kernel<<<(int)((E + T - 1) / T), T, 0, st>>>(dst, E, E, 0, 0, keys_in, vals_in);
    e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, csr_eid, (int)E, 0, bits_for(n_dst), st);
    if (e != cudaSuccess) return cuda_fail(e, "graph_build sort(dst)");
  }
  lower_bound_kernel<<<(int)((n_dst + 1 + T - 1) / T), T, 0, st>>>(keys_out, E, n_dst, csr_ptr);
  if (in_deg || n_norm) degree_norm_kernel<<<(int)((n_dst + T - 1) / T), T, 0, st>>>(csr_ptr, n_dst, in_deg, n_norm);
  if (edge_norm && E > 0) edge_norm_kernel<<<(int)((E + T - 1) / T), T, 0, st>>>(src, dst, n_norm, E, edge_norm);

  // src-CSC over the M rows (self row E+j has source self_off+j)
  if (csc_ptr && csc_row) {
    iota_keys_kernel<<<(int)((M + T - 1) / T), T, 0, st>>>(src, E, M, (int32_t)self_off, 0, keys_in, vals_in);
    e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, csc_row, (int)M, 0, bits_for(n_src), st);
    if (e != cudaSuccess) return cuda_fail(e, "graph_build sort(src)");
    lower_bound_kernel<<<(int)((n_src + 1 + T - 1) / T), T, 0, st>>>(keys_out, M, n_src, csc_ptr);
  }
  // relation segments over the M rows (self rows carry relation n_rel_rows-1)
  if (rel_ptr && rel_row) {
    MRG_CHECK_ARG(E == 0 || etype, "graph_build: null etype");
    iota_keys_kernel<<<(int)((M + T - 1) / T), T, 0, st>>>(etype, E, M, (int32_t)(n_rel_rows - 1), 1, keys_in, vals_in);
    e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, rel_row, (int)M, 0,
                                        bits_for(n_rel_rows), st);
    if (e != cudaSuccess) return cuda_fail(e, "graph_build sort(etype)");
    lower_bound_kernel<<<(int)((n_rel_rows + 1 + T - 1) / T), T, 0, st>>>(keys_out, M, n_rel_rows, rel_ptr);
  }
  MRG_LAUNCH_CHECK("graph_build");
  return MRG_OK;
}

extern "C" int mrg_graph_build(const int32_t* src, const int32_t* dst, const int32_t* etype, int64_t E, int64_t N,
                               int64_t n_rel_rows, int32_t* in_deg, float* n_norm, float* edge_norm,
                               int32_t* csr_ptr, int32_t* csr_eid, int32_t* csc_ptr, int32_t* csc_row,
                               int32_t* rel_ptr, int32_t* rel_row, void* workspace, size_t workspace_bytes,
                               void* stream) {
  return graph_build_impl(src, dst, etype, E, N, N, 0, n_rel_rows, in_deg, n_norm, edge_norm, csr_ptr, csr_eid, csc_ptr,
                          csc_row, rel_ptr, rel_row, workspace, workspace_bytes, stream);
}

extern "C" int mrg_graph_build_part(const int32_t* src, const int32_t* dst_local, const int32_t* etype, int64_t E,
                                    int64_t n_src, int64_t n_dst, int64_t node_lo, int64_t n_rel_rows,
                                    int32_t* csr_ptr, int32_t* csr_eid, int32_t* csc_ptr, int32_t* csc_row,
                                    int32_t* rel_ptr, int32_t* rel_row, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  return graph_build_impl(src, dst_local, etype, E, n_src, n_dst, node_lo, n_rel_rows, nullptr, nullptr, nullptr,
                          csr_ptr, csr_eid, csc_ptr, csc_row, rel_ptr, rel_row, workspace, workspace_bytes, stream);
}

extern "C" int mrg_edge_norm(const int32_t* src, const int32_t* dst, const float* n_norm, int64_t E, float* edge_norm,
                             void* stream) {
  MRG_CHECK_ARG(E == 0 || (src && dst && n_norm && edge_norm), "edge_norm: null pointer");
  if (E > 0) edge_norm_kernel<<<(int)((E + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, n_norm, E, edge_norm);
  MRG_LAUNCH_CHECK("edge_norm");
  return MRG_OK;
}
