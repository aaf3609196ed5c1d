/*
 * mrgnas.h -- C ABI of libmrgnas.so: hand-written sm_100a kernels for the MR-GNAS
 * multi-relational message-passing hot path (BASELINE.json:north_star).
 *
 * The upstream project (Amanda-Zheng/MR-GNAS) is pure Python: it has no FFI of its own.
 * On this path it reaches native code through two third-party libraries, DGL
 * (update_all / apply_edges -> libdgl gspmm/gsddmm) and torch ATen (index_select,
 * index_add_, batch_norm, sigmoid, binary_cross_entropy).  Each entry point below states
 * the reference call site(s) (file:line, relative to the upstream repo root) whose native
 * work it replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - all feature matrices are fp32, row-major, leading dimension == D, D % 4 == 0,
 *     D <= 512, base pointers 16-byte aligned; indices are int32;
 *   - the library never allocates, frees or retains memory: outputs, saved tensors and
 *     workspaces are owned by the caller (sizes from the *_bytes / *_count helpers);
 *   - calls only enqueue work on `stream` (a cudaStream_t passed as void*), never
 *     synchronise, and are re-entrant across streams/devices (one process per GPU);
 *   - return value 0 = ok, otherwise an mrg_status; mrg_last_error() gives the message of
 *     the last failure on the calling thread; no C++ exception crosses the boundary;
 *   - results are deterministic run to run (no floating-point atomics anywhere).
 */
#ifndef MRGNAS_H_
#define MRGNAS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRG_ABI_VERSION 1

typedef enum {
  MRG_OK = 0,
  MRG_ERR_INVALID = 1,  /* bad argument (null pointer, D % 4 != 0, ...) */
  MRG_ERR_CUDA = 2,     /* a CUDA runtime call / launch failed            */
  MRG_ERR_WORKSPACE = 3 /* workspace too small                            */
} mrg_status;

/* composition of source-entity and relation features (operations_lp.py:71-98) */
typedef enum { MRG_COMP_SUB = 0, MRG_COMP_MULT = 1, MRG_COMP_ADD = 2 } mrg_comp;
/* destination reductions = DGL update_all(copy_e, max|sum|mean) (operations_lp.py:233,248,262)
 * and the NC std reducer (operations.py:167-190) */
typedef enum { MRG_RED_SUM = 0, MRG_RED_MEAN = 1, MRG_RED_MAX = 2 } mrg_reduce;

/* A feature matrix read through an optional per-column affine + ReLU:
 *   value(i,c) = relu?( scale[c] * data[i,c] + shift[c] )       (scale == NULL: identity affine)
 * This is how BatchNorm1d(+ReLU) that follows every op in the reference
 * (model_lp.py:31-33, cell_lp.py:21,31-32) is folded into the consumer's loads. */
typedef struct {
  const float* data;
  const float* scale;
  const float* shift;
  int32_t relu;
} mrg_act;

/* up to MRG_MAX_MIXED candidate outputs of one DARTS MixedOp */
#define MRG_MAX_MIXED 8
typedef struct {
  mrg_act acts[MRG_MAX_MIXED];
  int32_t n;
} mrg_act_list;

int mrg_abi_version(void);
const char* mrg_last_error(void);

/* ------------------------------------------------------------------------------------
 * K0  graph build (integer, bit-exact).  Replaces dgl.DGLGraph()/add_edges/in_degrees/
 * apply_edges in build_graph (train/mr_lp_train.py:77-89), comp_deg_norm
 * (utils/utils_rgcn.py:120-127) and the stable COO->CSR DGL does inside update_all.
 *
 * Inputs: E directed edges (src,dst,etype) in EDGE-ID order, N nodes, n_rel_rows relation
 * rows (2R+1; id n_rel_rows-1 is the self-loop relation).  "Rows" are the M = E + N
 * edge-expanded rows of the reference (model_lp.py:126-131): row i<E is edge i, row E+n is
 * node n's self loop with src n and relation n_rel_rows-1.
 * Outputs (caller-allocated):
 *   in_deg[N], n_norm[N] = in_deg^-1/2 (inf->0), edge_norm[E] = n_norm[dst]*n_norm[src]
 *     (norm pointers may be NULL to skip);
 *   dst-CSR  csr_ptr[N+1], csr_eid[E]   : edge ids ascending inside each destination;
 *   src-CSC  csc_ptr[N+1], csc_row[M]   : row ids ascending inside each source;
 *   relation rel_ptr[n_rel_rows+1], rel_row[M] : row ids ascending inside each relation.
 * ---------------------------------------------------------------------------------- */
size_t mrg_graph_workspace_bytes(int64_t E, int64_t N, int64_t n_rel_rows);
int mrg_graph_build(const int32_t* src, const int32_t* dst, const int32_t* etype, int64_t E, int64_t N,
                    int64_t n_rel_rows, int32_t* in_deg, float* n_norm, float* edge_norm, int32_t* csr_ptr,
                    int32_t* csr_eid, int32_t* csc_ptr, int32_t* csc_row, int32_t* rel_ptr, int32_t* rel_row,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Destination-partitioned form (the reference is single-device): this process owns the
 * destinations [node_lo, node_lo + n_dst) of an n_src-node graph and the E edges that point at them
 * (dst_local = dst - node_lo), in global edge-id order.  Rows are the M = E + n_dst local edge-expanded rows;
 * self row E+j has source node_lo+j.  csr_ptr[n_dst+1]; csc_ptr[n_src+1] (gather tables keep n_src rows);
 * workspace: mrg_graph_workspace_bytes(E, n_dst, n_rel_rows).  Degree norms come from the GLOBAL in-degrees
 * (caller) through mrg_edge_norm. */
int mrg_graph_build_part(const int32_t* src, const int32_t* dst_local, const int32_t* etype, int64_t E, int64_t n_src,
                         int64_t n_dst, int64_t node_lo, int64_t n_rel_rows, int32_t* csr_ptr, int32_t* csr_eid,
                         int32_t* csc_ptr, int32_t* csc_row, int32_t* rel_ptr, int32_t* rel_row, void* workspace,
                         size_t workspace_bytes, void* stream);

/* edge_norm[e] = n_norm[dst[e]] * n_norm[src[e]] (one IEEE fp32 multiply; the apply_edges UDF of
 * train/mr_lp_train.py:86 and search/mr_lp_search.py:30-36) for a caller-supplied node norm. */
int mrg_edge_norm(const int32_t* src, const int32_t* dst, const float* n_norm, int64_t E, float* edge_norm,
                  void* stream);

/* Chunk table over a segment list (ptr[nseg+1]): splits every segment into pieces of at
 * most MRG_CHUNK_ROWS rows so hub nodes / frequent relations are reduced by many warps.
 * chunk_first[nseg+1] (exclusive scan of chunks per segment) and chunk_seg[max_chunks].
 * max_chunks = mrg_chunk_capacity(total_rows, nseg). */
#define MRG_CHUNK_ROWS 32
int64_t mrg_chunk_capacity(int64_t total_rows, int64_t nseg);
size_t mrg_chunk_workspace_bytes(int64_t nseg);
int mrg_chunk_build(const int32_t* ptr, int64_t nseg, int32_t* chunk_first, int32_t* chunk_seg, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * column statistics for training-mode BatchNorm1d over edge rows
 * (model_lp.py:23,32,50,72; cell_lp.py:21; torch native_batch_norm).
 * Producers write per-block partial (sum, sum of squares) in double to `stats`
 * ([mrg_stats_nparts(rows)][2][D] doubles); mrg_bn_finalize folds them in a fixed order.
 * ---------------------------------------------------------------------------------- */
int32_t mrg_stats_nparts(int64_t rows);
int32_t mrg_stats_max_parts(void);
int mrg_colstats(mrg_act x, int64_t rows, int32_t D, double* stats, void* stream);
/* mean/invstd (biased var, eps), a = gamma*invstd, b = beta - a*mean; updates
 * running_mean/var (unbiased var, momentum) when non-NULL.  nparts = total partial blocks. */
int mrg_bn_finalize(const double* stats, int32_t nparts, int64_t rows, int32_t D, const float* gamma,
                    const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                    float* mean, float* invstd, float* a, float* b, void* stream);
/* out = relu?(a*x+b): materialises an activation (BN-apply + ReLU in one pass). */
int mrg_affine_act(mrg_act x, int64_t rows, int32_t D, float* out, void* stream);
/* backward of s = relu?(a*y+b) with a,b from batch statistics of y:
 *   reduce: partial sums of dz and dz*y, dz = ds * [s>0]          -> bwd_stats
 *   finalize: dgamma, dbeta and coef[3][D] with dy = coef2*dz + coef0 + coef1*y
 *   apply:  dy (may alias ds).  accumulate!=0 adds into dy instead of overwriting. */
int mrg_bn_bwd_reduce(const float* ds, mrg_act y, int64_t rows, int32_t D, double* bwd_stats, void* stream);
int mrg_bn_bwd_finalize(const double* bwd_stats, int32_t nparts, int64_t rows, int32_t D, const float* gamma,
                        const float* mean, const float* invstd, float* dgamma, float* dbeta, float* coef,
                        void* stream);
int mrg_bn_bwd_apply(const float* ds, mrg_act y, const float* coef, int64_t rows, int32_t D, float* dy,
                     int32_t accumulate, void* stream);

/* ------------------------------------------------------------------------------------
 * K1  edge-expanded gather + composition.  Replaces all_ent_emb[src_id_final],
 * rel_embed[edge_type_final] (model_lp.py:126-131; ATen index_select) fused with
 * pre_sub/pre_mult/pre_add (operations_lp.py:71-98):
 *   y[i,:] = h[h_idx[i],:]  (-|*|+)  r[r_idx[i],:]      (idx NULL: identity rows)
 * plus the BN column statistics of y (stats may be NULL).
 * Backward (replaces ATen index_add_ atomics with deterministic segmented sums) is
 * expressed with mrg_seg_reduce over the src-CSC and the relation segments.
 * ---------------------------------------------------------------------------------- */
int mrg_compose_fwd(const float* h, const int32_t* h_idx, const float* r, const int32_t* r_idx, int64_t rows,
                    int32_t D, int32_t comp, float* y, double* stats, void* stream);
/* elementwise backward for the un-gathered op form: dx = dy (*r), dr = (+|-)dy (*x) */
int mrg_compose_bwd_rows(const float* dy, const float* x, const float* r, int64_t rows, int32_t D, int32_t comp,
                         float* dx, float* dr, void* stream);

/* ------------------------------------------------------------------------------------
 * K3/K6  collapsed sparse gate.  Replaces f_sparse_op_comp (operations_lp.py:304-343),
 * f_sparse_op (345-354; operations.py:205-216) and f_sparse_op_last (405-416) on a
 * contiguous row range that shares one (W,a) pair:
 *   t_i = x_i . v1 + xin_i . v2 + c ;  gate_i = sigmoid(t_i)
 *   y[i,:] = base_scale * (row_scale ? row_scale[i] : 1) * gate_i * x[i,:]
 * where v = a.weight @ W.weight (split into the x / xin halves) and c = a.weight . W.bias
 * (no non-linearity sits between W and a in the reference).  xin.data == NULL drops the
 * second term; xin.data == x.data reads the row once.  `gate` [rows] is saved for backward.
 * Backward returns dx/dxin w.r.t. the activated inputs (accumulate bit 0: dx +=, bit 1: dxin +=) and per-block
 * partials of dv1, dv2, dc in `dparam` ([nparts][2*D+1] doubles... see mrg_gate_dparam_count).
 * ---------------------------------------------------------------------------------- */
int mrg_sparse_gate_fwd(mrg_act x, mrg_act xin, int64_t rows, int32_t D, const float* v1, const float* v2,
                        const float* c, const float* row_scale, float base_scale, float* y, float* gate,
                        double* stats, void* stream);
/* The collapse itself and its backward for up to three segments (in / out / self) in one launch each:
 *   v[s,:] = a[s] @ W[s]   (a [D], W [D, K] = nn.Linear(K, D).weight),   c[s] = a[s] . b[s]   (b NULL: 0)
 * written as v1 = v[:, :D1] ([nseg, D1]) and v2 = v[:, D1:] ([nseg, K-D1]; NULL when K == D1);
 *   dW[s] = a[s]^T dv[s],   da[s] = W[s] dv[s] + dc[s] b[s],   db[s] = dc[s] a[s].
 * Replaces the a.weight @ W.weight matmuls that autograd would otherwise run per segment, gate and step. */
typedef struct {
  const float* W[3];
  const float* b[3];
  const float* a[3];
} mrg_gate_params;
typedef struct {
  float* dW[3];
  float* db[3];
  float* da[3];
} mrg_gate_grads;
int mrg_gate_collapse_fwd(mrg_gate_params p, int32_t nseg, int32_t D, int32_t K, int32_t D1, float* v1, float* v2,
                          float* c, void* stream);
int mrg_gate_collapse_bwd(mrg_gate_params p, int32_t nseg, int32_t D, int32_t K, int32_t D1, const float* dv1,
                          const float* dv2, const float* dc, mrg_gate_grads g, void* stream);
int64_t mrg_gate_dparam_count(int64_t rows, int32_t D);
int mrg_sparse_gate_bwd(const float* dy, mrg_act x, mrg_act xin, const float* gate, int64_t rows, int32_t D,
                        const float* v1, const float* v2, const float* row_scale, float base_scale, float* dx,
                        float* dxin, int32_t accumulate, double* dparam, void* stream);
/* folds the partials: dv1[D], dv2[D] (NULL if no xin), dc[1] */
int mrg_sparse_gate_bwd_finalize(const double* dparam, int64_t rows, int32_t D, float* dv1, float* dv2, float* dc,
                                 void* stream);

/* A gradient read through the lazy BatchNorm1d(+ReLU) backward of the state it belongs to
 * (model_lp.py:31-33 backward; torch native_batch_norm_backward + threshold_backward):
 *   dz(i,c)    = ds[i,c] * [ y.relu ? y.scale[c]*y.data[i,c] + y.shift[c] > 0 : 1 ]
 *   value(i,c) = coef ? coef[2D+c]*dz + coef[c] + coef[D+c]*y.data[i,c] : ds[i,c]
 * coef is the [3][D] output of mrg_bn_bwd_finalize; with coef == NULL it is a plain gradient matrix. */
typedef struct {
  const float* ds;
  mrg_act y;
  const float* coef;
} mrg_grad;

/* mrg_sparse_gate_bwd with the BatchNorm backward of the neighbouring states folded in (TMA-staged row
 * pipeline, D <= 256): the incoming gradient is an mrg_grad view, so no separate mrg_bn_bwd_apply pass runs, and, when
 * x_bwd_stats != NULL, the kernel also writes the per-CTA partial column sums
 *   sum_i dz_x(i,c), sum_i dz_x(i,c) * x.data[i,c],   dz_x = (final dx row) * [x.relu ? x(i,c) > 0 : 1]
 * in the layout mrg_bn_bwd_finalize consumes ([mrg_stats_nparts(rows)][2][D] doubles) -- valid as the BN
 * backward statistics of x's own state when this call contributes the LAST term of dx (no separate
 * mrg_bn_bwd_reduce pass).  Everything else as mrg_sparse_gate_bwd. */
int mrg_sparse_gate_bwd_fused_supported(int32_t D);
int mrg_sparse_gate_bwd_fused(mrg_grad dy, mrg_act x, mrg_act xin, const float* gate, int64_t rows, int32_t D,
                              const float* v1, const float* v2, const float* row_scale, float base_scale, float* dx,
                              float* dxin, int32_t accumulate, double* dparam, double* x_bwd_stats, void* stream);

/* dense gate epilogue: y = scale_i * sigmoid?(z) * x  (f_dense_op_comp / f_comp_op /
 * f_dense_op(_last), operations_lp.py:266-288,356-401) after the edge-tile GEMM z. */
int mrg_dense_gate_fwd(const float* z, mrg_act x, int64_t rows, int32_t D, int32_t use_sigmoid,
                       const float* row_scale, float base_scale, float* y, double* stats, void* stream);
int mrg_dense_gate_bwd(const float* dy, const float* z, mrg_act x, int64_t rows, int32_t D, int32_t use_sigmoid,
                       const float* row_scale, float base_scale, float* dz, float* dx, int32_t accumulate,
                       void* stream);

/* ------------------------------------------------------------------------------------
 * K9  DARTS MixedOp (cell_lp.py:25-33; cell.py:23-31):
 *   out = sum_k w[k] * relu?(scale_k * y_k + shift_k)     (left-to-right, as Python's sum())
 * i.e. every candidate's BatchNorm-apply + ReLU and the softmax(alpha)-weighted sum in one
 * pass over the rows.  The backward reuses mrg_bn_bwd_reduce/apply per candidate (ds_k = w_k*dout).
 * A candidate whose output is identically zero (f_zero_op, operations_lp.py:214-220) is passed with data == NULL and
 * its BatchNorm affine in scale / shift: it is never materialised (mrg_bn_bwd_reduce accepts the same view).
 * ---------------------------------------------------------------------------------- */
int mrg_mixed_sum_fwd(mrg_act_list ys, const float* w, int64_t rows, int32_t D, float* out, void* stream);
/* MixedOp backward, candidate k, after mrg_bn_bwd_finalize of its BatchNorm: dw[k] = sum(dout * relu(a y + b)) from
 * the finalized statistics, then coef / dgamma / dbeta are scaled by w[k] in place (training = 0: coef = [0,0,a w_k]).
 * One launch instead of ~12 elementwise / reduction launches per candidate. */
int mrg_mixed_bwd_scale(float* coef, float* dgamma, float* dbeta, const float* a, const float* b, const float* mean,
                        const float* invstd, const float* w, int32_t k, float* dw, int32_t D, int32_t training,
                        void* stream);

/* ------------------------------------------------------------------------------------
 * K9b  MixedOp over the PRE candidate list with ONE shared read of its inputs (cell_lp.py:25-33 over
 * PRE_OPS = pre_mult / pre_sub / pre_add, operations_lp.py:71-98):
 *   out = sum_k w_k relu(BN_k(comp_k(a, b))),  comps[k] in MRG_COMP_{SUB, MULT, ADD}, K <= 3.
 * No candidate output is ever written: statistics of all K compositions in one pass over (a, b), then (after one
 * mrg_bn_finalize per candidate: scale / shift [K, D]) the mixed sum in a second pass; the backward mirrors it
 * (column sums of all K BatchNorm backwards in one pass over (dout, a, b); after mrg_bn_bwd_finalize +
 * mrg_mixed_bwd_scale per candidate: coef [K, 3, D]; then da, db in one pass).  stats / bwd_stats:
 * [K][mrg_stats_nparts(rows)][2][D] doubles.  da or db may be NULL.
 * ---------------------------------------------------------------------------------- */
int mrg_mixed_pre_stats(const float* a, const float* b, int64_t rows, int32_t D, const int32_t* comps, int32_t K,
                        double* stats, void* stream);
int mrg_mixed_pre_fwd(const float* a, const float* b, int64_t rows, int32_t D, const int32_t* comps, int32_t K,
                      const float* scale, const float* shift, const float* w, float* out, void* stream);
int mrg_mixed_pre_bwd_stats(const float* dout, const float* a, const float* b, int64_t rows, int32_t D,
                            const int32_t* comps, int32_t K, const float* scale, const float* shift, double* bwd_stats,
                            void* stream);
int mrg_mixed_pre_bwd(const float* dout, const float* a, const float* b, int64_t rows, int32_t D, const int32_t* comps,
                      int32_t K, const float* scale, const float* shift, const float* coef, float* da, float* db,
                      void* stream);

/* ------------------------------------------------------------------------------------
 * K5/K11  segmented reduction of gathered rows.  One kernel family serves
 *   (i)  DGL update_all(copy_e, max|sum|mean) over the dst-CSR (operations_lp.py:233,248,
 *        262; compgcn.py:87) and the NC UDF reducers incl. std (operations.py:105-190);
 *   (ii) the backward of the edge-expanded gathers (ATen index_add_ in the reference):
 *        dh over the src-CSC, d rel_embed over the relation segments.
 *   red[s,:] = REDUCE_{j in [ptr[s],ptr[s+1])}  value(idx[j])  (.*  mul[mul_idx[idx[j]],:])
 *   out[s,:] = alpha * red (/ max(len,1) for MEAN) + (residual ? residual(s,:) : 0)
 * MAX: empty segment -> 0; arg[s,c] = lowest idx attaining the max (-1 if empty) = the
 * stated tie-break; values are compared after the optional ReLU of `m`.
 * workspace: mrg_seg_reduce_workspace_bytes(n_chunks, D, kind).
 * ---------------------------------------------------------------------------------- */
size_t mrg_seg_reduce_workspace_bytes(int64_t max_chunks, int32_t D, int32_t kind);
int mrg_seg_reduce_fwd(int32_t kind, mrg_act m, const int32_t* ptr, const int32_t* idx, const int32_t* chunk_first,
                       const int32_t* chunk_seg, int64_t nseg, int64_t max_chunks, int32_t D, const float* mul,
                       const int32_t* mul_idx, float alpha, mrg_act residual, int32_t accumulate, float* out,
                       int32_t* arg, void* workspace, size_t workspace_bytes, void* stream);
/* backward of (i) w.r.t. the (pre-ReLU) message rows, in edge-id order:
 *   SUM : dm[e,:] = g[dst[e],:]              MEAN: g[dst[e],:] / max(deg,1)
 *   MAX : dm[e,c] = g[n,c] if arg[n,c]==e    (SUM/MEAN gated by [m>0] when m.relu; MAX by the arg code)
 * rows >= E (self-loop rows, when n_self>0) receive g[row-E,:] (the residual path). */
int mrg_seg_reduce_bwd(int32_t kind, const float* g, const int32_t* arg, const float* out, mrg_act m,
                       const int32_t* dst, const int32_t* ptr, int64_t E, int64_t n_self, int32_t D, float* dm,
                       int32_t accumulate, void* stream);

/* ------------------------------------------------------------------------------------
 * K5 (tensor-core form)  fused a_max: edge-tile GEMM + bias + ReLU + destination max in one
 * tcgen05 kernel (3xTF32 split for fp32-class accuracy), never materialising the [E,D]
 * messages.  Replaces F.relu(self.linear(src_emb[:E])) + update_all(copy_e, max) + the self-loop
 * residual of a_max_op (operations_lp.py:230-235; NC: operations.py:112-121, residual = null).
 *   out[n,f] = max_{e: dst[e]=n} relu(W[f,:] . x(e,:) + bias[f]) + residual(n,f)
 * x rows are addressed by edge id through the dst-CSR (csr_eid); `arg` gets the same encoded
 * argmax as mrg_seg_reduce_fwd(MAX).  Requires D % 8 == 0, D <= 256 (mrg_amax_tc_supported).
 * ---------------------------------------------------------------------------------- */
int mrg_amax_tc_supported(int32_t D);
size_t mrg_amax_tc_workspace_bytes(int64_t N, int32_t D);
int mrg_amax_tc_fwd(mrg_act x, const float* W, const float* bias, const int32_t* csr_eid, const int32_t* dst,
                    int64_t E, int64_t N, int32_t D, mrg_act residual, float* out, int32_t* arg, void* workspace,
                    size_t workspace_bytes, void* stream);
/* Reduced-precision variant of the same call (north_star "bf16 variants"): the activated message rows and W are
 * rounded to bf16 on the way into shared memory, ONE tcgen05 kind::f16 MMA per K step accumulates in fp32 TMEM
 * (a third of the tensor work of the 3xTF32 path, three times the X stages).  Same outputs, same argmax encoding;
 * stated tolerance: |out - out_fp32| <= 2e-2 * max|out| (tests/test_gpu_ops_lp.py::test_amax_bf16_variant). */
int mrg_amax_tc_fwd_bf16(mrg_act x, const float* W, const float* bias, const int32_t* csr_eid, const int32_t* dst,
                         int64_t E, int64_t N, int32_t D, mrg_act residual, float* out, int32_t* arg, void* workspace,
                         size_t workspace_bytes, void* stream);

/* K5 backward (sparse): gradients of a_max w.r.t. the E message-source rows (dX, edge-id order,
 * every row written once), the Linear weight (dW [D,D]) and bias (db [D]) from g = dL/d(out),
 * the encoded argmax and the (lazily activated) input rows.  Replaces the reference's two dense
 * [E,D]x[D,D] backward GEMMs with 2*N*D*D FMAs on the routed entries only.  dX / dW may be NULL.
 * csr_dst[p] = destination of the edge at dst-CSR position p (= dst[csr_eid[p]], graph-static), allocated with 64
 * extra entries (whole 64-row windows are bulk-copied).  The call first
 * turns `arg` into two bit tables over CSR positions (workspace), which both products then walk. */
size_t mrg_amax_bwd_workspace_bytes(int64_t N, int64_t E, int32_t D);
/* debugging aid: 8 x int64 device buffer receiving the dW kernel's per-phase cycle counts (NULL switches it off) */
int mrg_debug_set_dw_prof(long long* dev_buf);
int mrg_amax_bwd(const float* g, const int32_t* arg, mrg_act x, const float* W, const int32_t* csr_ptr,
                 const int32_t* csr_eid, const int32_t* csr_dst, const int32_t* chunk_first,
                 const int32_t* chunk_seg, int64_t N, int64_t E, int64_t max_chunks, int32_t D, float* dX, float* dW,
                 float* db, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * TransE 1-N scoring.  Replaces sf_TransE_op.forward (operations_lp.py:101-112):
 *   logit[b,n] = gamma - || query[b,:] - ent[n,:] ||_1 ,  query = sub_emb + rel_emb;  score = sigmoid(logit)
 * without the reference's [B, N, D] broadcast (2.98 GB at C1).  Backward: dquery[b,k] = -sum_n dlogit[b,n] *
 * sgn(query[b,k] - ent[n,k]), dent[n,k] = +sum_b (same), sgn(0) = 0; either output may be NULL.  D % 4 == 0.
 * ---------------------------------------------------------------------------------- */
int mrg_transe_fwd(const float* query, const float* ent, int64_t B, int64_t N, int32_t D, float gamma, float* logit,
                   void* stream);
size_t mrg_transe_bwd_workspace_bytes(int64_t B, int64_t N, int32_t D);
int mrg_transe_bwd(const float* dlogit, const float* query, const float* ent, int64_t B, int64_t N, int32_t D,
                   float* dquery, float* dent, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * K8  DistMult 1-N scoring epilogue + BCE.  Replaces torch.sigmoid + nn.BCELoss
 * (operations_lp.py:121-127; train/mr_lp_train.py:116,235): loss = mean over n elements of
 * -(y*max(log p,-100) + (1-y)*max(log(1-p),-100)), p = sigmoid(logit).
 * fwd writes per-block partial sums (double) then the mean to loss[0]; pred (nullable)
 * receives p.  bwd: dlogit = gscale * (p - y) * p(1-p) / max(p(1-p), 1e-12) / n.
 * ---------------------------------------------------------------------------------- */
int32_t mrg_bce_nparts(int64_t n);
int mrg_sigmoid_bce_fwd(const float* logit, const float* label, int64_t n, float* pred, double* partial,
                        float* loss, void* stream);
int mrg_sigmoid_bce_bwd(const float* logit, const float* label, int64_t n, const float* gscale, float* dlogit,
                        void* stream);

/* ------------------------------------------------------------------------------------
 * Label pipeline.  Replaces TrainDataset.get_label + label smoothing
 * (utils/data_set.py:17-33) and the per-step host->device copy of the dense [B, N] label matrix
 * (train/mr_lp_train.py:227): out[b, n - col_lo] = pos if n is listed in idx[ptr[b] .. ptr[b+1]) else neg, for
 * n in [col_lo, col_hi).  The caller passes neg = fl32(1/N), pos = fl32(fl32(1 - lbl_smooth) + fl32(1/N)) (or 0 / 1
 * without smoothing), which makes the rows bit-identical to the reference's.  out: [B, col_hi - col_lo] fp32.
 * ---------------------------------------------------------------------------------- */
int mrg_labels_from_csr(const int32_t* ptr, const int32_t* idx, int64_t B, int64_t col_lo, int64_t col_hi, float neg,
                        float pos, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Evaluation: filtered rank of the target object of every query.  Replaces the
 * torch.where + double argsort over [B, N] of predict() (train/mr_lp_train.py:289-302):
 *   rank[b] = 1 + #{ n : v(b,n) > v(b,obj[b])  or  (v(b,n) == v(b,obj[b]) and n < obj[b]) },
 *   v(b,n) = -1e7 if label[b,n] != 0 and n != obj[b] else pred[b,n]        (stable descending order)
 * pred, label: [B, N] fp32 row-major; obj: [B] int64 (the reference's triplets[:, 2]); rank: [B] int32.
 * ---------------------------------------------------------------------------------- */
int mrg_filtered_rank(const float* pred, const float* label, const int64_t* obj, int64_t B, int64_t N, int32_t* rank,
                      void* stream);

/* ------------------------------------------------------------------------------------
 * K7+K8 fused  DistMult 1-N scoring + sigmoid + BCE in one tcgen05 kernel (3xTF32 operands, fp32 TMEM
 * accumulation): replaces torch.mm(sub_emb * rel_emb, all_ent.T) (operations_lp.py:121-125) + torch.sigmoid
 * (:126) + nn.BCELoss (train/mr_lp_train.py:116,235) for the training loss (model_lp.py:148-150).
 *   logit[b,n] = sum_k query[b,k] * ent[n,k]        query = sub_emb * rel_emb  [B, D],  ent [N, D]
 *   loss       = mean_{b,n} -( y*max(log p,-100) + (1-y)*max(log(1-p),-100) ),  p = sigmoid(logit), y = label[b,n]
 * The [B,N] logits are written once (the backward, mrg_sigmoid_bce_bwd, needs them); probabilities are never
 * materialised.  partial: mrg_distmult_bce_nparts(B) doubles.  Requires D % 8 == 0, D <= 256; any B (256 query
 * rows per launch), any N (one destination-partition rank passes its own entity rows and label columns).
 * ---------------------------------------------------------------------------------- */
int mrg_distmult_bce_supported(int32_t D);
int32_t mrg_distmult_bce_nparts(int64_t B);
size_t mrg_distmult_bce_workspace_bytes(int32_t D);
int mrg_distmult_bce_fwd(const float* query, const float* ent, const float* label, int64_t B, int64_t N, int32_t D,
                         float* logit, double* partial, float* loss, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ------------------------------------------------------------------------------------
 * Node-level Linear on the tcgen05 main loop (3xTF32, fp32-class accuracy):
 *   out[r, f] = sum_k x[r,k] * W[f,k] + bias[f]      x [rows, K], W [F, K] (= nn.Linear.weight), out [rows, ldo]
 * Replaces the fp32 SIMT library GEMMs of nn.Linear on the path -- the cell's `concat` Linear (model_lp.py:70-71)
 * and `linear_e` (model_lp.py:124) -- and, called with W^T, their input gradients.  K % 8 == 0; any F
 * (256 output features per launch); bias may be NULL.
 * ---------------------------------------------------------------------------------- */
int mrg_linear_tc_supported(int32_t K);
size_t mrg_linear_tc_workspace_bytes(int32_t K);
int mrg_linear_tc_fwd(const float* x, const float* W, const float* bias, int64_t rows, int32_t K, int32_t F, float* out,
                      int64_t ldo, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Reduction GEMM on the tcgen05 main loop (3xTF32, fp32-class accuracy), the reduction running over the ROWS:
 *   C[f1, f2] = sum_n A[n, f1] * B[n, f2]          A [rows, F1] (lda), B [rows, F2] (ldb), C [F1, F2] (ldc)
 *   a_kmajor != 0: A is given transposed, At [F1, rows] (lda) -- the reduction index already contiguous.
 * Replaces the fp32 SIMT library GEMMs behind autograd on the path: the weight gradient dW = dY^T X of every
 * nn.Linear (model_lp.py:70-71,124; operations_lp.py:230-250,266-288,356-390; model.py:22-28) and the two backward
 * GEMMs of sf_DisMult_op's `torch.mm(obj_emb, all_ent.transpose(1, 0))` (operations_lp.py:115-127): dq = dl . ent
 * (a_kmajor, A = dl [B, N]) and dent = dl^T . q.  Split over the rows into per-CTA partials folded in a fixed order
 * (deterministic).  Any F1, F2 >= 1; rows >= 0 (rows == 0 zeroes C).  colsum (may be NULL): [F1] = sum_n A[n, f1], the
 * bias gradient `dY.sum(0)` of the same Linear, computed as one extra (virtual, all-ones) column of B.  bias (may be
 * NULL): [F2], added to every row of C -- with a_kmajor this is nn.Linear's forward x W^T + b (A = x [rows_out, K] as the
 * K-major operand, B = W^T) and, with B = W, its input gradient dY W.
 * ---------------------------------------------------------------------------------- */
size_t mrg_gemm_red_workspace_bytes(int64_t rows, int32_t F1, int32_t F2);
int mrg_gemm_red(const float* A, int64_t lda, int32_t a_kmajor, const float* B, int64_t ldb, int64_t rows, int32_t F1,
                 int32_t F2, float* C, int64_t ldc, float* colsum, const float* bias, void* workspace, size_t workspace_bytes,
                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MRGNAS_H_ */
