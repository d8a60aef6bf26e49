/*
 * gglb200.h -- C ABI of libgglb200.so: the B200 (sm_100a) message-passing aggregation backend that
 * replaces GammaGL's `gammagl.mpops` torch_ext operators.
 *
 * Reference interface replaced (all under /root/reference/gammagl/mpops/torch_ext unless stated):
 *   c_segment_sum / c_segment_mean / c_segment_max   src/operators.cpp:14-27  (SegmentSum/Mean/Max, src/segment_*.cpp)
 *   c_spmm_sum / c_spmm_mean / c_spmm_max            src/operators.cpp:29-43  (SpMMSum/Mean/Max,  src/gspmm.cpp:26-202)
 *   c_bspmm_sum                                      src/operators.cpp:45-49  (BSpMMSum,          src/gspmm.cpp:204-260)
 *   ind2ptr / ptr2ind (COO <-> CSR)                  gammagl/ops/sparse/cpu/convert.cpp:58-128
 *   segment_softmax                                  gammagl/utils/softmax.py:29-35
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch types.  Every data pointer is a DEVICE pointer
 *     owned by the caller (inputs, outputs and workspace); the library never allocates, never
 *     synchronises with the host and never reads device data on the host.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  All work is enqueued on it.
 *   - Return value: 0 = GGL_OK, negative = GGL_E* below, positive = a cudaError_t.  A human readable
 *     message for the last failure on the calling thread is available from gglb200_last_error().
 *   - Re-entrant from any thread; no global mutable state besides the thread-local error string.
 *   - Indices inside the library are int32 (CSR) -- E and N must be < 2^31.  The int64 COO form the
 *     reference API delivers (edge_index[2,E], segment_ids[E]) is consumed by gglb200_csr_build.
 *   - Edge order: every CSR built here is a STABLE sort of the COO input, so edges inside a row keep
 *     ascending original edge id.  Kernels reduce a row sequentially in that order, which is the
 *     order of the reference's serial CPU loops (cpu/spmm_sum_cpu.cpp:27-39) -- fp32 sums are
 *     therefore bit-identical to the reference CPU operators for rows that are not split (see
 *     GGLB200_LONG_ROW) and max/arg-max tie-breaks (first edge wins) are identical always.
 */
#ifndef GGLB200_H_
#define GGLB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GGLB200_ABI_VERSION 2

#if defined(__GNUC__)
#define GGL_API __attribute__((visibility("default")))
#else
#define GGL_API
#endif

/* status codes (negative); positive values are cudaError_t */
#define GGL_OK              0
#define GGL_EINVAL         -1   /* bad argument (null pointer, negative size, unknown enum)            */
#define GGL_EUNSUPPORTED   -2   /* dtype / reduce / shape combination not implemented                   */
#define GGL_EWORKSPACE     -3   /* workspace too small                                                  */
#define GGL_ERANGE         -4   /* E or N does not fit int32                                            */

/* reducers: aggr strings of MessagePassing.aggregate (gammagl/layers/conv/message_passing.py:84-92) */
#define GGL_SUM   0
#define GGL_MEAN  1
#define GGL_MAX   2

/* element types of x / out (AT_DISPATCH_ALL_TYPES_AND2(Half,BFloat16), cpu/segment_sum_cpu.cpp:32-34) */
#define GGL_F32   0
#define GGL_BF16  1
#define GGL_F16   2
#define GGL_F64   3
#define GGL_I8    4
#define GGL_I16   5
#define GGL_I32   6
#define GGL_I64   7
#define GGL_U8    8

/* Rows with more than GGLB200_LONG_ROW edges are reduced in chunks of GGLB200_LONG_CHUNK edges whose
 * partial results are combined in ascending chunk order by a second kernel (deterministic; fp32 sums
 * of such rows differ from the serial order by rounding only). */
#define GGLB200_LONG_ROW    1024
#define GGLB200_LONG_CHUNK   512
/* The row kernel works on "edge tiles": tile t owns the rows whose first edge lies in
 * [t*TILE_EDGES, (t+1)*TILE_EDGES) -- equal work per warp regardless of the degree distribution. */
#define GGLB200_TILE_EDGES   256

GGL_API int         gglb200_abi_version(void);
GGL_API const char* gglb200_last_error(void);
/* number of kernels this library has launched from the calling process (bench.py's gpu_launches) */
GGL_API int64_t     gglb200_launch_count(void);

/* Device-wide tuning the gather kernels rely on: sets cudaLimitMaxL2FetchGranularity of the CURRENT device
 * to `bytes` (32, 64 or 128; the driver default is 64).  Random 400-byte row gathers fetch ~25 % fewer DRAM
 * sectors at 32 (measured: 16.6 -> 12.3 sectors per edge at K=100).  Returns the previous value through
 * *previous when non-NULL.  This is the only entry point that changes device state. */
GGL_API int gglb200_set_l2_fetch_granularity(int bytes, int* previous);

/* ---------------------------------------------------------------------------------------------
 * COO -> CSR ingest.  Replaces SparseStorage's host argsort (gammagl/sparse/storage.py:124-140) and
 * ind2ptr (gammagl/ops/sparse/cpu/convert.cpp:58-90).
 *
 *   key   [E] int64  the index rows are formed over (edge_index[1]=dst for the forward CSR,
 *                    edge_index[0]=src for the transposed CSR, segment_ids for the segment ops)
 *   other [E] int64  the index gathered into `col` (src resp. dst); may be NULL when col is NULL
 *   rowptr[N+1] int32 out     rowptr[r]..rowptr[r+1] = edges of row r
 *   col   [E] int32 out       other[perm[i]]           (may be NULL)
 *   perm  [E] int32 out       original edge id of the i-th sorted edge (stable: ascending inside a row)
 *   plan  [gglb200_csr_plan_ints(E)] int32 out   long-row work list consumed by the reduce kernels
 *   n_other  number of rows `col` may address (x.shape[0]); < 0 = unchecked (col carries an encoding, e.g. peer ids)
 *   Range errors are never silent: keys outside [0,N) are COUNTED in plan[6], `other` ids outside [0,n_other) in
 *   plan[7] (and folded to a valid id so that no kernel dereferences them).  The caller reads plan[0..7] after the
 *   build -- the one host sync it needs anyway for plan[0] -- and raises when either count is non-zero, like the
 *   reference's CPU operators do (TORCH_CHECK_INDEX idx < N, cpu/segment_max_cpu.cpp:50).
 * ------------------------------------------------------------------------------------------- */
GGL_API size_t gglb200_csr_build_workspace_bytes(int64_t E, int64_t N);
GGL_API size_t gglb200_csr_plan_ints(int64_t E);
GGL_API int    gglb200_csr_build(const int64_t* key, const int64_t* other, int64_t E, int64_t N, int64_t n_other,
                         int32_t* rowptr, int32_t* col, int32_t* perm, int32_t* plan,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ind2ptr / ptr2ind of gammagl/ops/sparse/sparse.py:38-75 (cpu/convert.cpp:58-128, cuda/convert.cu:41-127) on int64:
 * ptr[M+1] from a SORTED row index ind[E] (empty rows allowed), and its inverse ind[E] from ptr[M+1]. */
GGL_API int gglb200_ind2ptr(const int64_t* ind, int64_t E, int64_t M, int64_t* ptr, void* stream);
GGL_API int gglb200_ptr2ind(const int64_t* ptr, int64_t M, int64_t E, int64_t* ind, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Two-key edge sort + duplicate detection (replaces the host argsort of `major * n + minor` in
 * gammagl/utils/sort_edge_index.py:31-45, gammagl/utils/coalesce.py:41-78 and
 * gammagl/sparse/storage.py:124-140,200-217).
 *   major/minor   int64 [E], ids in [0, n_major) / [0, n_minor)
 *   sorted_major / sorted_minor  int64 [E] out (either may be NULL): the pairs in ascending (major, minor) order;
 *                 equal pairs keep their input order (stable)
 *   perm          int32 [E] out: sorted position -> input position
 *   group         int64 [E] out or NULL: rank of the distinct pair held by each sorted position (0-based,
 *                 non-decreasing) -- the segment ids coalesce() sums duplicate attributes over (coalesce.py:73-76)
 *   n_unique      device int64 scalar out or NULL: number of distinct pairs
 *   n_bad         device int64 scalar out or NULL: number of ids outside [0, n_major) / [0, n_minor) -- a caller error
 *                 the host side raises on (such ids are folded into the key width, never returned silently)
 * gglb200_unique_edges keeps the first pair of every group: uniq[group[p]] = sorted[p]  (coalesce.py:69).
 * ------------------------------------------------------------------------------------------- */
GGL_API size_t gglb200_sort_edges_workspace_bytes(int64_t E, int64_t n_major, int64_t n_minor);
GGL_API int gglb200_sort_edges(const int64_t* major, const int64_t* minor, int64_t E, int64_t n_major, int64_t n_minor,
                       int64_t* sorted_major, int64_t* sorted_minor, int32_t* perm, int64_t* group,
                       int64_t* n_unique, int64_t* n_bad, void* workspace, size_t workspace_bytes, void* stream);
GGL_API int gglb200_unique_edges(const int64_t* sorted_major, const int64_t* sorted_minor, const int64_t* group, int64_t E,
                         int64_t* uniq_major, int64_t* uniq_minor, void* stream);

/* out[i] = in[perm[i]] for i < n; `width` consecutive elements of `elem_bytes` (4 or 8 ... any
 * multiple of 4) per index.  Used to bring edge weights into CSR order once per (graph, weight). */
GGL_API int gglb200_permute_rows(const void* in, const int32_t* perm, void* out, int64_t n, int64_t row_bytes,
                         void* stream);

/* ---------------------------------------------------------------------------------------------
 * The aggregation kernel family (one entry point):
 *
 *   out[r, k] = REDUCE_{i in [rowptr[r], rowptr[r+1])}  w(i, k) * x[col[i], k]          r < n_rows
 *
 *   reduce   GGL_SUM | GGL_MEAN | GGL_MAX
 *            MEAN divides the sum by the (unweighted) edge count of the row when count > 1
 *              (cpu/spmm_mean_cpu.cpp:50-58, cpu/segment_mean_cpu.cpp:67-76);
 *            MAX starts from lowest() and takes `cur < v` in ascending edge order
 *              (cpu/spmm_max_cpu.cpp:41-49, cpu/segment_max_cpu.cpp:54-62); empty rows stay lowest().
 *   w        NULL (unweighted: segment ops, GIN, SAGE-mean), or float32
 *            [E]    when heads == 1   (gspmm,  cpu/spmm_sum_cpu.cpp:31-38)
 *            [E,H]  when heads == H   (bspmm,  cpu/bspmm_sum_cpu.cpp:43-55; K = H*C, head of k = k / C)
 *            indexed by i (already in CSR order) when wperm == NULL, else by wperm[i].
 *   x        [n_src, K] dtype, row major (every col[i] < n_src);  out [n_rows, K] dtype (written fully, no
 *            pre-zeroing needed).  n_src bounds the 16-byte-aligned superset copies of unaligned rows.
 *   arg      NULL, or int64 [n_rows, K] (MAX only): col[i*] of the winning edge; `arg_empty` for
 *            rows without edges (segment_max: N, cpu/segment_max_cpu.cpp:27; spmm_max: 0,
 *            cpu/spmm_max_cpu.cpp:23-24).  For the segment ops col := perm, so this is the edge id.
 *   out_scale NULL, or fp32 [n_rows] (GGL_SUM, floating dtypes): out[r,:] = out_scale[r] * sum(...).  With
 *            gglb200_scale_rows on the input this evaluates D_dst^a A D_src^b x -- GCNConv's normalised propagate
 *            (layers/conv/gcn_conv.py:88-104) -- WITHOUT a per-edge weight stream (no w[perm[i]] gather).
 *   plan     the array gglb200_csr_build produced for this rowptr (required)
 *   workspace partials of split rows.  gglb200_spmm_workspace_bytes(E, K, reduce) is always sufficient; a
 *            caller that has read plan[0] (= number of chunks, once after gglb200_csr_build) may pass the
 *            smaller gglb200_spmm_workspace_bytes_for_chunks(plan[0], K, reduce).  NULL is allowed ONLY when
 *            no row has more than GGLB200_LONG_ROW edges (plan[1] == 0): rows are then never split.  A
 *            workspace smaller than plan[0] chunks is a contract violation (kernels stay memory-safe but
 *            rows whose chunks do not fit are left unwritten).
 *   dtype    GGL_F32 / GGL_BF16 / GGL_F16 accumulate in fp32; GGL_F64 in fp64; integer types in int64
 *            truncated to the type (w must be NULL for non-floating dtypes).
 * ------------------------------------------------------------------------------------------- */
GGL_API size_t gglb200_spmm_workspace_bytes(int64_t E, int64_t K, int reduce);
GGL_API size_t gglb200_spmm_workspace_bytes_for_chunks(int64_t n_chunks, int64_t K, int reduce);
GGL_API int    gglb200_spmm_csr(int reduce,
                        const int32_t* rowptr, const int32_t* col, const int32_t* plan,
                        const float* w, const int32_t* wperm, int32_t heads,
                        const void* x, void* out, int64_t* arg, int64_t arg_empty, const float* out_scale,
                        int64_t n_rows, int64_t n_src, int64_t E, int64_t K, int dtype,
                        void* workspace, size_t workspace_bytes, void* stream);

/* MAX with the winning EDGE recorded: the fused form of `message` + `unsorted_segment_max` (message_passing.py:56-59,
 * 88-90; src/segment_max.cpp:37-61) that SAGEConv(aggr='pool') and every aggr='max' layer run -- without the [E,K]
 * message tensor.  out[r,k] = max_i w[perm[i]] * x[col[i],k] exactly like gglb200_spmm_csr(GGL_MAX), but
 * arg_eid[r,k] = perm[i*] = the ORIGINAL edge id of the first maximal edge (ascending edge id, strict <; arg_empty for rows
 * without edges).  Backward on the transposed CSR (rows = source): gx[s,k] = sum over out-edges e = (s -> d) with
 * arg_eid[d,k] == e of w[e] * g[d,k] -- one edge per (d,k), so parallel edges are NOT double counted (the gradient of
 * the reference's layers on multigraphs). */
GGL_API int gglb200_spmm_csr_max_edge(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const int32_t* plan,
                              const float* w, const void* x, void* out, int64_t* arg_eid, int64_t arg_empty,
                              int64_t n_rows, int64_t n_src, int64_t E, int64_t K, int dtype, void* workspace,
                              size_t workspace_bytes, void* stream);
GGL_API int gglb200_spmm_max_edge_bwd(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* perm_t, const float* w,
                              const float* g, const int64_t* arg_eid, float* gx, int64_t n_rows, int64_t K, void* stream);

/* Multi-GPU variant (1-D node partition): the same kernel with x spread over the GPUs of one node.
 * x_table is a DEVICE array of P base pointers (x shard of every rank, mapped into this process: CUDA IPC /
 * symmetric memory; all shards [rows_r, K] with the same K and dtype) and col[i] = owner << owner_shift | local_row.
 * Remote rows are fetched by the same bulk async copies straight from the owner's HBM over NVLink, so the
 * transfer overlaps the reduction tile by tile -- no pack / all-to-all / halo buffer.  Requires K*sizeof(dtype) % 16
 * == 0, GGL_SUM or GGL_MEAN.  The caller orders the kernel after every rank has finished writing its shard. */
/* gglb200_spmm_csr(GGL_SUM, ...) with the layer epilogue fused into the row write:
 *   out[r, k] = act( out_scale[r] * sum(...) + bias[k] ),  act = max(., 0) when relu != 0.
 * out_scale [n_rows] and bias [K] are fp32 and may be NULL; each step is rounded separately, so the result is
 * bit-identical to running `out + bias` (gcn_conv.py:104-106) and the model's ReLU as separate passes. */
GGL_API int gglb200_spmm_csr_epilogue(const int32_t* rowptr, const int32_t* col, const int32_t* plan, const float* w,
                              const int32_t* wperm, int32_t heads, const void* x, void* out, const float* out_scale,
                              const float* bias, int relu, int64_t n_rows, int64_t n_src, int64_t E, int64_t K,
                              int dtype, void* workspace, size_t workspace_bytes, void* stream);
GGL_API int gglb200_spmm_csr_peer(int reduce, const int32_t* rowptr, const int32_t* col, const int32_t* plan,
                          const float* w, const int32_t* wperm, int32_t heads, const void* const* x_table,
                          int32_t owner_shift, void* out, const float* out_scale, int64_t n_rows, int64_t E,
                          int64_t K, int dtype, void* workspace, size_t workspace_bytes, void* stream);

/* Halo-cache fill of the 1-D node partition: dst[i, :] = x_table[ids[i] >> owner_shift][(ids[i] & mask), :] for i < n
 * (row_bytes per row, a multiple of 16).  Rows that several local edges reference are pulled over NVLink ONCE per pass
 * into a cache region behind the rank's own shard (remote reads: no pack on the owner, no all-to-all); `col` of those
 * edges then addresses owner = self, row = B + slot, and gglb200_spmm_csr_peer finds them in local HBM. */
GGL_API int gglb200_gather_rows_peer(const void* const* x_table, int32_t owner_shift, const int32_t* ids, void* dst,
                             int64_t n, int64_t row_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Neighbour sampling for the mini-batch path (ops/sparse/cpu/sample.cpp:10-135 `sample_adj`, the per-hop rule of
 * ops/sparse/cpu/neighbor_sample.cpp:22-100; ops/sparse/cuda/neighbor_sample.cu): for every seed idx[i] pick `fanout` of its
 * neighbours in the int64 CSR (rowptr, col) -- all of them when fanout < 0 or (no replacement and fanout >= degree),
 * uniformly with replacement, or without replacement by Floyd's algorithm (the reference's rule).
 *   count   out_ptr[n_seeds+1] = exclusive scan of the edges per seed (out_ptr[n_seeds] = total); *n_bad = seeds outside [0,n_nodes)
 *   pick    e_id[total] = picked CSR positions grouped by seed, nbr[total] = col[e_id]; deterministic in (seed, inputs)
 * ------------------------------------------------------------------------------------------- */
GGL_API size_t gglb200_sample_workspace_bytes(int64_t n_seeds);
GGL_API int gglb200_sample_count(const int64_t* rowptr, const int64_t* idx, int64_t n_seeds, int64_t n_nodes, int64_t fanout,
                         int replace, int64_t* out_ptr, int64_t* n_bad, void* workspace, size_t workspace_bytes, void* stream);
GGL_API int gglb200_sample_pick(const int64_t* rowptr, const int64_t* col, const int64_t* idx, int64_t n_seeds, int64_t n_nodes,
                        int64_t fanout, int replace, uint64_t seed, const int64_t* out_ptr, int64_t* e_id, int64_t* nbr,
                        void* stream);

/* ---------------------------------------------------------------------------------------------
 * The dense feature x weight GEMM next to the aggregation, on the tcgen05 tensor cores with a TMEM accumulator:
 *   out[M,N] = act(A[M,K] @ W[K,N] + bias[N]),  fp32 in / out, act = ReLU when relu != 0, bias may be NULL.
 * The Linear of GCNConv / SAGEConv / GINConv (layers/conv/gcn_conv.py:79,104-106, sage_conv.py:80, gin_conv.py:57) applied to
 * aggregated rows (K = F_in <= 256, a multiple of 4; N <= 256, any width).  fp32-accurate on TF32 tensor cores by
 * operand splitting (3xTF32: a_hi.w_hi + a_hi.w_lo + a_lo.w_hi), ~1e-6 relative.  Other shapes: GGL_EUNSUPPORTED (the
 * host side then uses the library GEMM).
 * ------------------------------------------------------------------------------------------- */
GGL_API size_t gglb200_linear_workspace_bytes(int64_t K, int64_t N);   /* W split into hi / lo tensor-core tiles, rebuilt per call */
GGL_API int gglb200_linear_tf32x3(const float* A, const float* W, const float* bias, int relu, float* out, int64_t M, int64_t K,
                          int64_t N, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Sort-free unsorted segment reduce (the role of mpops/torch_ext/cuda/segment_{sum,mean,max}_cuda.cu -- scalar atomicAdd /
 * CAS loops there): ONE streaming pass over x[E,K] in its original order, reductions resolved by the L2.
 *   SUM   vector reductions (red.global.add.v4.f32); fp32 sums differ from the serial loop by rounding only (exact for
 *         integer-valued data, e.g. degree() = unsorted_segment_sum(ones), utils/degree.py:35-40)
 *   MEAN  SUM, cnt[N] (int32 out, required), rows with cnt > 1 divided (cpu/segment_mean_cpu.cpp:67-76)
 *   MAX   bit-exact and deterministic incl. arg[N,K] (first maximal edge, cpu/segment_max_cpu.cpp:54-62); cnt required;
 *         workspace = gglb200_segment_reduce_workspace_bytes(GGL_MAX, N, K)
 * ids outside [0,N) are skipped and counted in *n_bad (device int64, may be NULL).  float32 only.  Use the CSR path
 * (gglb200_csr_build + gglb200_spmm_csr(col = perm)) when the same ids are reduced again and again.
 * ------------------------------------------------------------------------------------------- */
GGL_API size_t gglb200_segment_reduce_workspace_bytes(int reduce, int64_t N, int64_t K);
GGL_API int gglb200_segment_reduce(int reduce, const void* x, const int64_t* ids, void* out, int64_t* arg, int64_t arg_empty,
                           int32_t* cnt, int64_t* n_bad, int64_t E, int64_t K, int64_t N, int dtype, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backward helpers of the segment ops (src/segment_sum.cpp:43-54, src/segment_mean.cpp:44-63,
 * src/segment_max.cpp:48-61).
 *   gather:  gx[e, k] = g[ids[e], k]            (SUM)
 *            gx[e, k] = g[ids[e], k] / cnt[ids[e]]   (MEAN; cnt = rowptr differences; float dtypes)
 *   max:     gx[E,K] zero-filled by the callee, then gx[arg[r,k], k] = g[r,k] for every non-empty row r
 *            (empty rows carry the sentinel and are skipped -- the reference leaks them into row N).
 * ------------------------------------------------------------------------------------------- */
GGL_API int gglb200_segment_gather_bwd(int reduce, const void* g, const int64_t* ids, const int32_t* rowptr,
                               void* gx, int64_t E, int64_t K, int dtype, void* stream);
GGL_API int gglb200_segment_max_bwd(const void* g, const int64_t* arg, const int32_t* rowptr,
                            void* gx, int64_t E, int64_t n_rows, int64_t K, int dtype, void* stream);
/* The same SUM / MEAN backward in scatter form through the segment CSR of the forward pass
 * (rowptr, perm, plan from gglb200_csr_build(key = ids, other = NULL)):  gx[perm[p], k] = g[row(p), k] (/ cnt[row(p)]).
 * Reads every gradient row once instead of once per message; identical results to the gather form. */
GGL_API int gglb200_segment_expand_bwd(int reduce, const void* g, const int32_t* rowptr, const int32_t* perm,
                               const int32_t* plan, void* gx, int64_t E, int64_t n_rows, int64_t K, int dtype,
                               void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backward of spmm_mean / spmm_max on the TRANSPOSED CSR (rows = source nodes).
 *   mean:  call gglb200_scale_rows_by_count first (g / indeg, cpu/spmm_mean_cpu.cpp:96-98), then
 *          gglb200_spmm_csr(SUM) on the transposed CSR.
 *   max:   gx[s,k] = sum over edges (s -> d) with argsrc[d,k] == s of w * g[d,k]
 *          (cpu/spmm_max_cpu.cpp:84-96; parallel edges count twice, as there).
 * ------------------------------------------------------------------------------------------- */
/* Folded GCN normalisation helpers: out[r,:] = scale[r] * x[r,:] (floating dtypes), and
 * out[r] = deg(r)^power with deg = rowptr[r+1]-rowptr[r], 0 for isolated rows (utils/degree.py + tlx.pow(deg, -0.5)). */
GGL_API int gglb200_scale_rows(const void* x, const float* scale, void* out, int64_t n_rows, int64_t K, int dtype, void* stream);
GGL_API int gglb200_degree_power(const int32_t* rowptr, float* out, int64_t n_rows, float power, void* stream);
GGL_API int gglb200_scale_rows_by_count(const float* g, const int32_t* rowptr, float* out, int64_t n_rows,
                                int64_t K, void* stream);
GGL_API int gglb200_spmm_max_bwd(const int32_t* rowptr_t, const int32_t* col_t, const float* w,
                         const int32_t* wperm, const float* g, const int64_t* argsrc, float* gx,
                         int64_t n_rows, int64_t K, void* stream);

/* ---------------------------------------------------------------------------------------------
 * bspmm backward wrt the per-head edge weight (SDDMM), cpu/bspmm_sum_cpu.cpp:95-107:
 *   gw[e, h] = sum_c x[src_e, h, c] * g[dst_e, h, c]
 * evaluated over the forward CSR (rows = dst) in pure edge tiles (plan = the CSR's plan array): x rows are
 * bulk-copied through the shared-memory ring, the g row of the current destination stays in registers, results
 * are written to gw[perm[i], h] (every edge exactly once; gw needs no initialisation).
 * ------------------------------------------------------------------------------------------- */
GGL_API int gglb200_bspmm_bwd_w(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const int32_t* plan,
                        const float* x, const float* g, float* gw,
                        int64_t n_rows, int64_t E, int32_t H, int32_t C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Edge softmax over incoming edges (gammagl/utils/softmax.py:29-35), the reference composition on the segment
 * kernels (rowptr/perm/plan = CSR over `ids`, built with other = NULL):
 *   m = segment_max(logit);  p = exp(logit - m[ids]);  s = segment_sum(p);  alpha = p / (s[ids] + 1e-16)
 * bwd: glogit = alpha * (galpha - segment_sum(alpha * galpha)[ids])  (analytic gradient; the max is a constant
 * shift).  logit / alpha / galpha / glogit are [E, H] float32 in ORIGINAL edge order; ids [E] int64.
 * workspace: gglb200_edge_softmax_workspace_bytes(n_rows, H, n_chunks) with n_chunks = plan[0] (0 if unknown and
 * the graph has no row above GGLB200_LONG_ROW edges; E / GGLB200_LONG_CHUNK + E / GGLB200_LONG_ROW + 2 is always safe).
 * ------------------------------------------------------------------------------------------- */
/* Sort-free form: pass rowptr = perm = plan = NULL (workspace = gglb200_edge_softmax_workspace_bytes(n_rows, H, 0)).  Every pass
 * streams [E,H] in its original order -- no CSR, no sort, no permuted gather: exact segment maxima by red.global.max.u32 on
 * order-preserving keys, denominators / Jacobian row terms by vector reductions (arrival-order sums), then the same element-wise
 * passes.  The two building blocks are exported as well (stat_m / stat_s / dot [n_rows,H] fp32, written fully). */
GGL_API int gglb200_edge_softmax_stats_oneshot(const int64_t* ids, const float* logit, float* stat_m, float* stat_s, float* p_or_null,
                                       int64_t n_rows, int64_t E, int32_t H, void* stream);
GGL_API int gglb200_edge_softmax_dot_oneshot(const int64_t* ids, const float* alpha, const float* galpha, float* dot, int64_t n_rows,
                                     int64_t E, int32_t H, void* stream);
GGL_API size_t gglb200_edge_softmax_workspace_bytes(int64_t n_rows, int32_t H, int64_t n_chunks);
GGL_API int gglb200_edge_softmax_fwd(const int32_t* rowptr, const int32_t* perm, const int32_t* plan, const int64_t* ids,
                             const float* logit, float* alpha, int64_t n_rows, int64_t E, int32_t H,
                             void* workspace, size_t workspace_bytes, void* stream);
GGL_API int gglb200_edge_softmax_bwd(const int32_t* rowptr, const int32_t* perm, const int32_t* plan, const int64_t* ids,
                             const float* alpha, const float* galpha, float* glogit, int64_t n_rows, int64_t E,
                             int32_t H, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused GAT aggregation (layers/conv/gat_conv.py:98-122 + utils/softmax.py:29-35) without any [E,H] tensor:
 *   pre_e = s_src[j,h] + s_dst[i,h];  alpha = softmax_i(leaky_relu(pre));  out[i,h,:] = sum_e alpha_e x[j,h,:]
 * fwd     forward CSR (rows = i, col = j).  xcat [N, xcat_stride] = [x (H*C) | s_src (H) | pad]; online softmax per
 *         (row, head); writes out [N,H*C] and the softmax statistics stat_m / stat_s [N,H].
 * bwd_src transposed CSR (rows = j, col = i).  gcat [N, H*C + 4H] = [gout | {s_dst, m, 1/s, D} per head] with
 *         D[i,h] = <out[i,h,:], gout[i,h,:]>; writes grad x [N,H*C] and grad s_src [N,H].
 * bwd_dst forward CSR again; writes grad s_dst [N,H].
 * Requires C % 4 == 0, C/4 a power of two, H*C <= 128, fp32, 16-byte aligned rows.  workspace:
 * gglb200_gat_workspace_bytes(n_chunks of that CSR, H, C) (NULL only if the CSR has no row above GGLB200_LONG_ROW).
 * ------------------------------------------------------------------------------------------- */
GGL_API size_t gglb200_gat_workspace_bytes(int64_t n_chunks, int32_t H, int32_t C);
GGL_API int gglb200_gat_fwd(const int32_t* rowptr, const int32_t* col, const int32_t* plan, const float* xcat,
                    int64_t xcat_stride, const float* s_dst, float slope, float* out, float* stat_m, float* stat_s,
                    int64_t n_rows, int64_t E, int32_t H, int32_t C, void* workspace, size_t workspace_bytes, void* stream);
GGL_API int gglb200_gat_bwd_src(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* plan_t, const float* gcat,
                        int64_t gcat_stride, const float* xcat, int64_t xcat_stride, float slope, float* gx,
                        float* gs_src, int64_t n_rows, int64_t E, int32_t H, int32_t C, void* workspace,
                        size_t workspace_bytes, void* stream);
/* bwd     the whole backward in ONE pass over the transposed CSR: gglb200_gat_bwd_src plus grad s_dst accumulated by L2 reductions
 *         (red.global.add.f32 per head and edge; gs_dst [N,H] zero-filled by the callee, summed in arrival order).  gcat rows
 *         are [gout | {s_dst, m, 1/s, D} x H | int32 row id, 0, 0, 0] (stride H*C + 4H + 4). */
GGL_API int gglb200_gat_bwd(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* plan_t, const float* gcat,
                    int64_t gcat_stride, const float* xcat, int64_t xcat_stride, float slope, float* gx, float* gs_src,
                    float* gs_dst, int64_t n_rows, int64_t E, int32_t H, int32_t C, void* workspace, size_t workspace_bytes,
                    void* stream);
GGL_API int gglb200_gat_bwd_dst(const int32_t* rowptr, const int32_t* col, const int32_t* plan, const float* xcat,
                        int64_t xcat_stride, const float* gcat, int64_t gcat_stride, float slope, float* gs_dst,
                        int64_t n_rows, int64_t E, int32_t H, int32_t C, void* workspace, size_t workspace_bytes,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GGLB200_H_ */
