/* rpg.h -- C ABI of the B200-native RelPose-GNN message-passing library (librpg_b200.so).
 *
 * The reference (nianticlabs/relpose-gnn) is pure Python and has no FFI; the boundary it
 * offers for this path is the Python class API of
 *     python/niantic/modules/my_gnn_layer.py:277-311   simpleConvEdge_upt(in, edge, out).forward(x, edge_index, edge_attr)
 *     python/niantic/modules/posenet.py:1053-1091      the GNN part of PoseNetX_R2.forward
 *     python/niantic/modules/posenet.py:1021-1031      PoseNetX_R2.compute_RP
 *     python/niantic/modules/criterion.py:42-60        PoseNetCriterion.forward
 *     python/niantic/training/train.py:236-248         edge-dropout mask
 * Each entry point below names the reference code it replaces.  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *  - all memory is owned by the caller (PyTorch); the library keeps no pointer after the
 *    kernels are enqueued on `stream`; no host synchronisation inside any call;
 *  - return value: 0 = ok, < 0 = argument error (RPG_E_*), > 0 = cudaError_t / CUresult;
 *    rpg_last_error_string() describes the last failure on the calling thread;
 *  - graphs are batched PyG-style: G graphs x N nodes, node row g*N+n, edge row g*Ep+k where
 *    the per-graph edge template (src[k], dst[k]), k in [0,Ep), is shared by the whole batch
 *    (the FC enumeration of dataset_7Scenes_multi.py:377-385,418-422, optionally thinned by
 *    the batch-shared edge-dropout mask of train.py:238-242);
 *  - "bf16 mode": activations/weights bf16 (uint16_t bit patterns), fp32 accumulation in TMEM.
 */
#ifndef RPG_H_
#define RPG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* rpg_stream_t; /* cudaStream_t */
typedef uint16_t rpg_bf16;

enum {
  RPG_OK = 0,
  RPG_E_ARG = -1,       /* bad shape / null pointer / unsupported size */
  RPG_E_GRAPH = -2,     /* edge_index is not a batched uniform template */
  RPG_E_UNSUPPORTED = -3,
  RPG_E_DRIVER = -4     /* could not resolve cuTensorMapEncodeTiled */
};

const char* rpg_last_error_string(void);
int rpg_version(void);
/* Number of kernels this library has launched in this process (bench.py reports the per-step delta). */
int64_t rpg_launch_count(void);
/* Per-launch CUDA-event timing of the tcgen05 GEMM kernel between begin/end (bench.py roofline leg):
 * summed durations (ms), launch counts and executed FLOPs for the NT and TN modes.  end() synchronises. */
int rpg_profile_begin(void);
int rpg_profile_end(double* nt_ms, double* tn_ms, int* nt_launches, int* tn_launches,
                    double* nt_flops, double* tn_flops);
/* Per-launch records of the same profiling window, for EVERY kernel class of the path (rpg_profile_records drains what
 * rpg_profile_begin .. the last launch collected and ends the window; it synchronises on the recorded events).
 * cls: RPG_PROF_*; ms: CUDA-event duration on the launch stream; flops / bytes: executed FLOPs and ALGORITHMIC bytes
 * of the launch (DESIGN.md section 5 gives the per-unit figures); M, N, K: GEMM shape (0 otherwise).
 * While a window is open every kernel is launched WITHOUT programmatic dependent launch, so that consecutive kernels do
 * not overlap and an event pair brackets exactly one kernel.                                           */
enum { RPG_PROF_GEMM_NT = 0, RPG_PROF_GEMM_TN = 1, RPG_PROF_SEGMENT_SUM = 2, RPG_PROF_ATTENTION_FWD = 3,
       RPG_PROF_ATTENTION_BWD = 4, RPG_PROF_EDGE_INIT = 5, RPG_PROF_REDUCE = 6, RPG_PROF_OTHER = 7, RPG_PROF_CLASSES = 8 };
typedef struct {
  int32_t cls, M, N, K;
  float ms;
  double flops, bytes;
} rpg_prof_rec_t;
int rpg_profile_records(rpg_prof_rec_t* out, int max_records, int* n_records);
/* Device properties the host side sizes grids with (SM count etc.); also proves the .so loads. */
int rpg_device_sm_count(int device, int* sm_count);

/* ------------------------------------------------------------------------------------------
 * Graph structure
 * ---------------------------------------------------------------------------------------- */

/* Writes n int32 words from HOST memory (pageable is fine) to device memory on `stream`; the words travel as kernel
 * parameters, so the call neither synchronises nor uses a copy engine.  For the per-step graph template tables. */
int rpg_upload_words(int32_t* dst, const int32_t* src_host, int64_t n, rpg_stream_t stream);

/* Checks that edge_index [2, Et] (int64, device) is G copies of one per-graph template with node
 * offset g*N, i.e. what PyG batching of the reference datasets produces (train.py:24,132), and
 * extracts the template of graph 0 into tmpl_src/tmpl_dst [Ep] (int32, device).
 * *bad_count (device int32, zeroed by the callee) receives the number of violating columns; the
 * caller reads it back and raises ValueError (SURVEY.md 8b error convention).                    */
int rpg_validate_edge_index(const int64_t* edge_index, int64_t Et, int G, int N, int Ep,
                            int32_t* tmpl_src, int32_t* tmpl_dst, int32_t* bad_count, rpg_stream_t stream);

/* Per-graph template tables living in device memory (built on the host by the Python mirror from the
 * validated template; a few hundred bytes).                                                        */
typedef struct {
  int G, N, Ep;               /* graphs, nodes per graph, edges per graph                         */
  const int32_t* src;         /* [Ep] source node (local index) of template edge k                */
  const int32_t* dst;         /* [Ep] destination node                                            */
  const int32_t* in_ptr;      /* [N+1] CSR over destination: in-edges of node n                   */
  const int32_t* in_idx;      /* [Ep]                                                             */
  const int32_t* out_ptr;     /* [N+1] CSR over source: out-edges of node n                       */
  const int32_t* out_idx;     /* [Ep]                                                             */
  const float* inv_deg;       /* [N] 1 / max(1, in-degree)   (PyG mean aggregation [3p])          */
  const float* deg;           /* [N] in-degree as float                                           */
  const int32_t* min_ptr;     /* [N+1] CSR over min(src,dst): edges whose lower endpoint is n     */
  const int32_t* min_idx;     /* [Ep]                                                             */
  const int32_t* max_ptr;     /* [N+1] CSR over max(src,dst)                                      */
  const int32_t* max_idx;     /* [Ep]                                                             */
  /* optional one-hot selection patterns for the K-panel gathers (NULL => gathers run in the epilogue):
   * [sel_patterns * 128, 64] bf16 each, see rpg_gemm_t.gsel                                          */
  const rpg_bf16* sel_src;
  const rpg_bf16* sel_dst;
  int sel_patterns, sel_div;
  const float* has_in;        /* [N] 1 if the node has incoming edges, else 0 (a mean over nothing is 0 [3p])   */
  /* per-graph edge sets (rpg_per_graph_tables: G = 1 above): edges / nodes of ONE member graph, 0 otherwise.  With
   * them set, sel_src / sel_dst hold one pattern tile per 128-row block (rpg_selection_patterns_rows, sel_div = 0).   */
  int pg_Ep, pg_N;
} rpg_graph_t;

/* Tables for a batch whose graphs have different edge sets of equal size Ep (dynamic kNN rewiring, posenet.py:1043-1050):
 * graph g owns edge columns [g*Ep, (g+1)*Ep) of edge_index [2, G*Ep] and nodes [g*N, (g+1)*N).  Built on the device, no
 * host round trip.  `tables`: rpg_per_graph_tables_words(G, N, Ep) int32 words laid out as
 *   src | dst | in_ptr | in_idx | out_ptr | out_idx | min_ptr | min_idx | max_ptr | max_idx | inv_deg | deg | has_in
 * (edge tables G*Ep words, ptr tables G*N + 1, node tables G*N; each rounded up to a multiple of 4 words), to be used as
 * an rpg_graph_t with G = 1, N = G*N, Ep = G*Ep (global rows).  bad[0] (device int32) counts edges that leave their
 * graph's node range. */
int64_t rpg_per_graph_tables_words(int G, int N, int Ep);
int rpg_per_graph_tables(const int64_t* edge_index, int G, int N, int Ep, int32_t* tables, int32_t* bad, rpg_stream_t stream);

/* Template tables of a batch of G IDENTICAL graphs built on the device in one launch from the template endpoints
 * tmpl_src / tmpl_dst (device int32 [Ep], as rpg_validate_edge_index extracts them): no host round trip between
 * receiving a fresh edge_index and launching the layer.  `tables`: rpg_template_tables_words(N, Ep) int32 words,
 *   src | dst | in_ptr | in_idx | out_ptr | out_idx | min_ptr | min_idx | max_ptr | max_idx | inv_deg | deg | has_in
 * (edge tables Ep words, ptr tables N + 1, node tables N; each rounded up to a multiple of 4 words; CSR order = edge
 * order within a node, i.e. a stable counting sort -- identical to the host-built tables).  N <= 1024, Ep <= 65536. */
int64_t rpg_template_tables_words(int N, int Ep);
int rpg_template_tables(const int32_t* tmpl_src, const int32_t* tmpl_dst, int N, int Ep, int32_t* tables, rpg_stream_t stream);

/* The batched edge_index of the template (what PyG's Batch hands the model, train.py:24,132), on the device:
 * edge_index [2, G*Ep] int64 with column g*Ep + k = (g*N + src[k], g*N + dst[k]). */
int rpg_build_edge_index(const rpg_graph_t* graph, int64_t* edge_index, rpg_stream_t stream);

/* Builds the one-hot selection tiles of rpg_graph_t.sel_src / sel_dst on the device from a template endpoint table
 * (src or dst): sel [patterns * 128, 64] bf16, div = gcd(128, Ep), patterns = Ep / div.  The caller checks first
 * that no 128-row block references more than 64 node rows.                                             */
int rpg_selection_patterns(const int32_t* endpoint, int Ep, int N, int div, int patterns, rpg_bf16* sel,
                           rpg_stream_t stream);
/* The same for per-graph edge sets (no period): one tile per 128-row block b of the edge rows, column =
 * endpoint[row] - ((b * 128) / pg_Ep) * pg_N with `endpoint` the GLOBAL node row of every edge ([Et]); sel
 * [ceil(Et / 128) * 128, 64].  bad[0] (device int32, pre-zeroed by the caller) counts columns outside [0, 64). */
int rpg_selection_patterns_rows(const int32_t* endpoint, int64_t Et, int pg_Ep, int pg_N, rpg_bf16* sel, int32_t* bad,
                                rpg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Generic tcgen05 GEMM with fused epilogue (the workhorse; exposed for unit tests)
 *   C[M,N] = epilogue( sum_s A_s[M,K_s] * B[N, K_0+K_1+..]^T )            (mode NT, K-major)
 *   C[M,N] = sum over rows r of A[r, M]^T B[r, N]   split over `splits`    (mode TN, MN-major)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int mode;                   /* 0 = NT (A [M,K] row-major, B [N,K] row-major), 1 = TN (A [R,M], B [R,N]) */
  int M, N;                   /* output shape                                                       */
  int n_seg;                  /* NT: number of A segments (1..6) concatenated along K                */
  const rpg_bf16* A[6];       /* NT: segment s is [M, K[s]] with row pitch lda[s] (elements)          */
  int K[6];
  int lda[6];
  const rpg_bf16* B;          /* NT: [N, sum K] pitch ldb;   TN: [R, N] pitch ldb                     */
  int ldb;
  int R;                      /* TN: number of contracted rows                                       */
  int splits;                 /* TN: split-R factor; partial s is written at out_f32 + s*split_stride */
  int64_t split_stride;
  int block_n;                /* UMMA N tile: multiple of 16 in [16,256]; 0 = library default        */
  /* ---- epilogue (NT mode), applied in this order ---- */
  const float* bias;          /* [N] or NULL                                                         */
  const rpg_bf16* gadd[2];    /* + gadd[i][ node_row(row, gmap[i]) , col ]  (row pitch gadd_ld[i])    */
  const int32_t* gmap[2];     /* template table (rpg_graph_t.src or .dst), node_row = (row/Ep)*N + gmap[row%Ep] */
  int gadd_ld[2];
  int Ep, Nn;
  const rpg_bf16* resid;      /* + resid[row, col] (pitch resid_ld) or NULL                           */
  int resid_ld;
  const float* row_scale;     /* * row_scale[row % row_scale_mod] or NULL (e.g. 1/deg)               */
  int row_scale_mod;
  const rpg_bf16* mask;       /* * (mask[row, col] > 0)  (ReLU backward) or NULL                      */
  int mask_ld;
  int relu;                   /* max(.,0) on the stored value                                        */
  rpg_bf16* out;              /* bf16 output (pitch ldo) or NULL                                      */
  rpg_bf16* out_relu;         /* second bf16 output holding max(value,0) or NULL                      */
  int ldo;
  float* out_f32;             /* fp32 output (pitch ldo_f32) or NULL                                  */
  int ldo_f32;
  /* ReLU patterns as bit matrices [M, ld bytes], bit (col & 7) of byte (col >> 3): 16x less traffic than a bf16 mask */
  const uint8_t* mask_bits;   /* * pattern bit (applied where `mask` would be) or NULL                */
  int mask_bits_ld;
  uint8_t* out_bits;          /* receives (stored value > 0) or NULL; needs N % 64 == 0               */
  int out_bits_ld;
  /* fp32 mode ("split bf16"): an fp32 value travels as a (hi, lo) bf16 pair, v = hi + lo, and a Linear is
   * evaluated as [A_hi | A_lo | A_hi] [W_hi | W_hi | W_lo]^T (K segments) with fp32 accumulation: ~2^-16 relative. */
  const float* gadd_f32[2];   /* gathered fp32 node rows (uses gmap[i]) or NULL                        */
  int gadd_f32_ld[2];
  const rpg_bf16* resid_lo;   /* low plane of resid (pitch resid_ld) or NULL                          */
  rpg_bf16* out_lo;           /* low plane of out: bf16(v - float(bf16(v))) (pitch ldo) or NULL        */
  rpg_bf16* out_relu_lo;      /* low plane of out_relu or NULL                                        */
  /* Gathered node adds as one-hot K panels (preferred over gadd): for each of n_gseg (0..4) operands one extra
   * 64-wide k-block is multiplied, A = the 128 x 64 one-hot selection tile of the row block -- which for a batch of
   * identical graph templates depends only on (row0 mod Ep), so gsel holds gsel_patterns tiles, pattern index
   * ((row0 % Ep) / gsel_div), or row block index row0 / 128 when gsel_div = 0 -- and B = rows [(row0 / Ep) * Nn, +64) of gsrc [gsrc_rows, >= N] (pitch gsrc_ld),
   * loaded MN-major.  Exact: 1.0 * bf16 accumulates in fp32.  The epilogue stays the plain one.        */
  int n_gseg;
  const rpg_bf16* gsel[4];
  int gsel_patterns, gsel_div;
  const rpg_bf16* gsrc[4];
  int gsrc_ld[4];
  int gsrc_rows;
  /* TN mode: per-split column sums of A, i.e. sum_r A[r, m] (the bias gradient when A is dY), written to
   * a_colsum [splits, M]; accumulated from the shared-memory operand tiles by the warps that are idle during the
   * main loop, so the gradient tensor is not read a second time.  NULL = off.                          */
  float* a_colsum;
  /* NT mode, plain epilogue: counter-based feature dropout fused into the `out_relu` output (posenet.py:1073-1075 applied
   * by the kernel that produces the last layer's outputs): out_relu = keep ? max(v, 0) / (1 - drop_p) : 0 and out_bits =
   * pattern of THAT tensor; `out` stays undropped.  The keep decision is the one of rpg_dropout_mask(drop_seed, drop_p).
   * drop_p = 0: off.                                                                                      */
  uint64_t drop_seed;
  float drop_p;
} rpg_gemm_t;

int rpg_gemm(const rpg_gemm_t* g, rpg_stream_t stream);
/* Tuning knob: CTAs per thread-block cluster of the GEMM kernel (1, or 2 = weight tiles fetched once per pair
 * and TMA-multicast into both shared memories; the default).                                        */
int rpg_set_gemm_cluster(int ctas_per_cluster);

/* Weight gradient dW[M,N] += A[R,M]^T B[R,N] (fp32, pitch ldo): TN GEMM split over R into `ws`
 * (rpg_layer_bwd_ws_floats elements) followed by the deterministic reduction below.                 */
int rpg_wgrad(const rpg_bf16* A, int lda, int M, const rpg_bf16* B, int ldb, int N, int64_t R,
              float* ws, float* out, int ldo, rpg_stream_t stream);
/* Same, and bias[m] += sum_r A[r, m] from the same pass (bias may be NULL).                            */
int rpg_wgrad_bias(const rpg_bf16* A, int lda, int M, const rpg_bf16* B, int ldb, int N, int64_t R,
                   float* ws, float* out, int ldo, float* bias, rpg_stream_t stream);
/* Several weight gradients that share B and whose A operands are adjacent column blocks of ONE tensor (the two halves
 * of proj_edge.weight, posenet.py:1014-1017: dW[:, blk] += A[:, blk*M:(blk+1)*M]^T B): one product of M * nblocks
 * rows, folded block by block into outs[blk] (pitch ldo).  bias (optional) += column sums of block 0.  nblocks <= 4. */
int rpg_wgrad_blocks(const rpg_bf16* A, int lda, int M, int nblocks, const rpg_bf16* B, int ldb, int N, int64_t R,
                     float* ws, float* const* outs, int ldo, float* bias, rpg_stream_t stream);
/* sizeof / offsetof probes so a foreign-language mirror of the structs can verify its layout.      */
void rpg_struct_sizes(int32_t* out16);   /* 13 values used, 16 slots */

/* out[r, c] (+)= sum_s partial[s, r, c]  -- deterministic second stage of the TN split.            */
int rpg_reduce_splits(const float* partial, int splits, int64_t split_stride, int rows, int cols,
                      float* out, int ldo, int accumulate, rpg_stream_t stream);

/* Several folds in ONE launch (a layer backward queues all of its weight-gradient folds): every descriptor is an
 * accumulating rpg_reduce_splits; outputs of different descriptors must not overlap. */
#define RPG_REDUCE_BATCH_MAX 32
typedef struct rpg_reduce_desc {
    const float* part;      /* [splits][rows*cols] partials, split stride `stride` floats */
    float* out;             /* [rows, ldo] accumulated into */
    int64_t stride;
    int32_t splits, rows, cols, ldo;
} rpg_reduce_desc_t;
typedef struct rpg_reduce_batch {
    rpg_reduce_desc_t d[RPG_REDUCE_BATCH_MAX];
    int32_t n;
} rpg_reduce_batch_t;
int rpg_reduce_splits_batch(const rpg_reduce_batch_t* batch, rpg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Weights: fp32 master parameters (reference state_dict layout) -> packed bf16 operands
 * ---------------------------------------------------------------------------------------- */

/* dst[r, c] = bf16(src[r0 + r, c0 + c]) for a rows x cols window of a row-major fp32 matrix;
 * transpose != 0 writes dst[c, r] instead.  Used to cut the concatenated-input weights of
 * my_gnn_layer.py:232,280,284 into per-source blocks and to build the dgrad (transposed) copies. */
int rpg_pack_weight(const float* src, int ld_src, int r0, int c0, int rows, int cols,
                    rpg_bf16* dst, int ld_dst, int transpose, rpg_stream_t stream);

/* Many rpg_pack_weight windows in ONE launch (a training step re-packs every operand after optimizer.step()):
 * descriptor i converts the rows x cols window at (r0, c0) of the fp32 matrix src (pitch ld_src) into dst (pitch ld_dst),
 * transposed if `transpose`; dst is bf16, or fp32 when dst_f32 != 0 (plain strided copy: head matrices, bias vectors),
 * or the bf16 LOW plane bf16(w - float(bf16(w))) when lo_plane != 0 (fp32 mode).                          */
#define RPG_PACK_BATCH_MAX 64
typedef struct {
  const float* src;
  void* dst;
  int32_t ld_src, r0, c0, rows, cols, ld_dst;
  int16_t transpose, dst_f32, lo_plane, pad_;
} rpg_pack_desc_t;
typedef struct {
  rpg_pack_desc_t d[RPG_PACK_BATCH_MAX];
  int32_t n;
} rpg_pack_batch_t;
int rpg_pack_weights_batch(const rpg_pack_batch_t* batch, rpg_stream_t stream);

/* Small fp32 products with the master weights (weight composition and its backward), several in ONE launch:
 *   C[m, n] (+)= sum_k opA(m, k) opB(k, n)     opA = A[m, k] or A[k, m] (transA), opB = B[k, n] or B[n, k] (transB)
 * plus an optional rank-1 term u[m] v[n]; optional bf16 copies of the result, plain (Cb) and transposed (CbT).     */
#define RPG_SGEMM_BATCH_MAX 8
typedef struct {
  const float* A; const float* B; float* C;
  const float* u; const float* v;
  rpg_bf16* Cb; rpg_bf16* CbT;
  int32_t M, N, K, lda, ldb, ldc, ldcb, ldcbT;
  int32_t transA, transB, accumulate, pad_;
} rpg_sgemm_desc_t;
typedef struct {
  rpg_sgemm_desc_t d[RPG_SGEMM_BATCH_MAX];
  int32_t n;
} rpg_sgemm_batch_t;
int rpg_sgemm_batch(const rpg_sgemm_batch_t* batch, rpg_stream_t stream);

/* Adam step (torch.optim.Adam semantics as used by train.py:211: L2 weight decay added to the gradient, bias
 * correction, no amsgrad) over flat fp32 buffers of n elements, one launch:
 *   g = grad * grad_scale + wd * p;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;
 *   p -= lr / (1 - b1^t) * m / (sqrt(v / (1 - b2^t)) + eps)
 * grad_scale folds the 1/world of a summed data-parallel gradient into the step.  step = t >= 1.          */
int rpg_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, float grad_scale, int64_t step, rpg_stream_t stream);

int rpg_cast_f32_to_bf16(const float* src, rpg_bf16* dst, int64_t n, rpg_stream_t stream);
int rpg_cast_bf16_to_f32(const rpg_bf16* src, float* dst, int64_t n, rpg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Bandwidth-bound kernels of the path
 * ---------------------------------------------------------------------------------------- */

/* AttentionBlock core, att.py:25-30: y[e,i] = sum_j softmax_j(phi[e,i]*theta[e,j]) * g[e,j].
 * gtp [Et, 3c] fp32 holds (g | theta | phi) rows; y [Et, ldy] bf16 (only the first c columns written). */
/* Two implementations with the same result to fp32 rounding: the SERIES form (default when aux == NULL and c % 16 == 0;
 * csrc/rpg_attention.cu: the rank-1 logits make exp(phi_i theta_j) separable, O(c K) multiply-adds per row with the order
 * K chosen from the row's range and an exact exp2 path for rows beyond the series bound) and the exp2 form (aux != NULL,
 * or RPG_ATT_SERIES=0).  rpg_attention_series_enabled() tells the caller whether `aux` is still worth allocating. */
int rpg_attention_series_enabled(void);
int rpg_attention_fwd(const float* gtp, int64_t Et, int c, rpg_bf16* y, int ldy, rpg_bf16* y_lo /* NULL in bf16 mode */,
                      float* aux /* optional [Et, 4c] fp32 row statistics for the backward, or NULL */,
                      rpg_stream_t stream);
/* The series form on bf16 projections [Et, 3c] (dense rows; the bf16 mode of the layer: half the HBM traffic of the
 * fp32 tensor and twice the resident warps).  Same results up to the bf16 rounding of (g | theta | phi).          */
int rpg_attention_fwd_bf16(const rpg_bf16* gtp16, int64_t Et, int c, rpg_bf16* y, int ldy, rpg_stream_t stream);
int rpg_attention_bwd_bf16(const rpg_bf16* gtp16, const float* dyn, int ld_dyn, const rpg_graph_t* graph, int64_t Et, int c,
                           rpg_bf16* dgtp, int ld_dgtp, rpg_stream_t stream);
/* Backward of the above with dy[e,:] = dyn[node(dst(e)), :] gathered through the template:
 * dgtp [Et, ld_dgtp] bf16 = (dg | dtheta | dphi) in the first 3c columns.                          */
int rpg_attention_bwd(const float* gtp, const float* dyn, int ld_dyn, const rpg_graph_t* graph,
                      int64_t Et, int c, rpg_bf16* dgtp, int ld_dgtp,
                      const float* aux /* the forward's statistics: c^2 instead of 2 c^2 exps per row; or NULL */,
                      rpg_stream_t stream);

/* PyG mean aggregation over destination [3p], reached from my_gnn_layer.py:301:
 * a[g*N+n, :] = inv_deg[n] * sum_{k in in-edges(n)} z[g*Ep+k, :]   (fixed order, no atomics).       */
int rpg_aggregate_mean(const rpg_bf16* z, int ldz, const rpg_graph_t* graph, int D,
                       rpg_bf16* a, int lda, rpg_stream_t stream);
/* Segment sums edge rows -> node rows, by source (by_src=1) or destination (0), fp32 accumulate:
 * out[g*N+n, :] = sum_{k in out/in-edges(n)} v[g*Ep+k, :]                                          */
int rpg_edge_to_node_sum(const rpg_bf16* v, int ldv, const rpg_graph_t* graph, int D, int by_src,
                         rpg_bf16* out, int ldo, rpg_stream_t stream);
/* General form over any CSR of the template, with an optional ReLU mask (mask[row,c] > 0) on the edge
 * rows and an optional per-node scale: out[g*N+n,:] = scale[n] * sum_{k in csr(n)} (v * [mask>0])[g*Ep+k,:] */
int rpg_segment_sum(const rpg_bf16* v, int ldv, const rpg_bf16* mask, int ldm, const int32_t* csr_ptr,
                    const int32_t* csr_idx, const float* scale, const rpg_graph_t* graph, int D,
                    rpg_bf16* out, int ldo, rpg_stream_t stream);

/* Two segment sums of the same edge tensor in one pass over HBM (the backward's scatters come in pairs: by source and by
 * destination, by lower and by upper endpoint): which_* selects the CSR, 0 = in (destination), 1 = out (source),
 * 2 = min endpoint, 3 = max endpoint.  out_x[g*N+n, :] = sum_{k in csr_x(n)} v[g*Ep+k, :].                     */
int rpg_segment_sum2(const rpg_bf16* v, int ldv, const rpg_graph_t* graph, int which_a, int which_b, int D, rpg_bf16* out_a,
                     int ldo_a, rpg_bf16* out_b, int ldo_b, rpg_stream_t stream);

/* Edge-feature initialiser, posenet.py:1014-1017 + :1053-1055 in factorised form:
 * e0[g*Ep+k, :] = relu(pmin[node(min(s,t)), :] + pmax[node(max(s,t)), :] + bias), where
 * pminmax [Nt, 2D] = x * [W_min; W_max]^T (proj_edge.weight[:, 0:D] and [:, D:2D]).
 * Backward = rpg_segment_sum over min_ptr/max_ptr with mask = e0, then node-level GEMMs.             */
int rpg_edge_init_fwd(const rpg_bf16* pminmax, int ldp, const float* bias, const rpg_graph_t* graph, int D,
                      rpg_bf16* e0, int lde, uint8_t* e0_bits, rpg_stream_t stream);

/* Feature dropout + pose heads, posenet.py:1073-1086: pose[r, 0:3] = fc_xyz(drop(f[r])), [3:6] = fc_wpqr(..).
 * Dropout follows F.dropout (kept entries scaled by 1/(1-p)).  The keep decision is either an explicit
 * uint8 [rows, D] mask `keep` (parity tests: ATen's Philox stream cannot be reproduced) or, when keep is
 * NULL and p_drop > 0, a counter-based hash of (seed, row, col / 4) evaluated in-kernel, 8 bits per element,
 * i.e. the drop probability is quantised to multiples of 1/256 (no mask traffic);
 * rpg_dropout_mask materialises that same decision so an oracle can consume it.                      */
int rpg_dropout_mask(uint64_t seed, float p_drop, int64_t rows, int D, uint8_t* keep, rpg_stream_t stream);
int rpg_head_fwd(const rpg_bf16* feat, const rpg_bf16* feat_lo /* fp32 mode: low plane, else NULL */, int ldf,
                 int64_t rows, int D, const uint8_t* keep, uint64_t seed, float p_drop, const float* w6,
                 const float* b6, float* pose, rpg_stream_t stream);
/* dfeat = (dpose * W6) * keep * scale * (feat > 0 if mask_relu); weight/bias gradients of the translation head
 * (rows 0..2 of W6: dw_t [3, D], db_t [3]) and of the rotation head (rows 3..5: dw_q, db_q), (+)= if accumulate;
 * ws: rpg_head_bwd_ws_floats(rows, D) floats of scratch for the deterministic two-stage reduction.   */
int64_t rpg_head_bwd_ws_floats(int64_t rows, int D);
int rpg_head_bwd(const float* dpose, const rpg_bf16* feat, int ldf, int64_t rows, int D, const uint8_t* keep,
                 uint64_t seed, float p_drop, const float* w6, int mask_relu, rpg_bf16* dfeat, int lddf,
                 float* dw_t, float* dw_q, float* db_t, float* db_q, int accumulate, float* ws,
                 rpg_stream_t stream);

/* Pose heads on tensor cores, for features that the producing GEMM already dropped and rescaled (rpg_gemm_t.drop_*):
 * forward is a plain rpg_gemm (pose = feat_d W6^T + b6 with W6 padded to 8 rows); this is its backward:
 *   dfeat = scale * (dpose W6) * [feat_d > 0]  (bits = pattern of feat_d = keep & relu; dfeat may be NULL)
 *   dw_t / dw_q (+)= (dpose^T feat_d) rows 0..2 / 3..5 ;  db_t / db_q (+)= column sums of dpose.
 * dpose fp32 [rows, 6] travels as a bf16 K panel [hi | lo] (dp16: scratch [rows, 64]); w6T_ext bf16 [D, 64] holds
 * W6[j, :] in columns j and 8 + j (j < 6), zeros elsewhere; ws: rpg_head_bwd_tc_ws_floats(D) floats. */
int64_t rpg_head_bwd_tc_ws_floats(int D);
int rpg_pack_dpose(const float* dpose, int64_t rows, rpg_bf16* dp16, float scale, float* scale_out, rpg_stream_t stream);
int rpg_head_bwd_tc(const float* dpose, const rpg_bf16* feat_d, int ldf, const uint8_t* bits, int64_t rows, int D,
                    float scale, const rpg_bf16* w6T_ext, rpg_bf16* dp16, rpg_bf16* dfeat, int lddf, float* dw_t,
                    float* dw_q, float* db_t, float* db_q, float* ws, rpg_stream_t stream);

/* compute_RP (posenet.py:1021-1031) + the L1 sums of PoseNetCriterion (criterion.py:51-52) fused:
 * target[e] = poses[src(e)] - poses[dst(e)];  sums[0] = sum|pred_t - targ_t|, sums[1] = sum|pred_q - targ_q|
 * (fp32, deterministic two-stage reduction); dpred[e,j] = sign(pred - target) * grad_scale[j<3 ? 0 : 1]
 * (grad_scale: device [2], e.g. exp(-sax)/(3 Et), exp(-saq)/(3 Et); NULL = 1).  target/dpred may be NULL. */
int64_t rpg_pose_loss_ws_floats(int64_t Et);
int rpg_pose_loss(const float* pred, const float* poses, const rpg_graph_t* graph, int64_t Et,
                  const float* grad_scale, float* target, float* sums, float* dpred, float* ws,
                  rpg_stream_t stream);
/* The same with the learned loss weights of PoseNetCriterion folded in (criterion.py:55-57; sax, saq: device
 * scalars): out7 = [sum_t, sum_q, loss, t_loss, q_loss, dloss/dsax, dloss/dsaq] with t_loss = sum_t / (3 Et),
 * loss = exp(-sax) t_loss + sax + exp(-saq) q_loss + saq; dpred = dloss/dpred.  No host round trip. */
int rpg_pose_criterion(const float* pred, int ld_pred, const float* poses, const rpg_graph_t* graph, int64_t Et,
                       const float* sax, const float* saq, float* target, float* out7, float* dpred, float* ws,
                       rpg_stream_t stream);   /* ld_pred: row pitch of pred in floats (>= 6; the tensor-core heads emit 8) */

/* Dynamic kNN rewiring: torch_cluster.knn_graph(x, k, batch, loop=False) [3p, torch-cluster 1.5.9] as called at
 * posenet.py:1043-1050 for G graphs of N nodes each (2 <= N <= 64, k < N): edge_index [2, G*N*k] int64 with
 * column (g*N + i)*k + r = (r-th nearest other node of i in graph g  ->  i); squared Euclidean distances in fp32 on
 * x [G*N, D] (pitch ldx), ties to the lower node index. */
int rpg_knn_graph(const float* x, int ldx, int G, int N, int D, int k, int64_t* edge_index, rpg_stream_t stream);
/* The same for a batch of graphs of DIFFERENT sizes (a general PyG `batch` vector): graph g owns the node rows
 * [node_ptr[g], node_ptr[g+1]) and writes min(k, n_g - 1) edges per node (fewer candidates than k: all of them, as
 * torch_cluster does) starting at column edge_ptr[g]; node_ptr / edge_ptr: device int64 [G+1]; max_nodes = the largest
 * n_g (<= 64); n_edges = edge_ptr[G] = the number of columns of edge_index [2, n_edges]. */
int rpg_knn_graph_ragged(const float* x, int ldx, int G, const int64_t* node_ptr, const int64_t* edge_ptr, int max_nodes, int D,
                         int k, int64_t n_edges, int64_t* edge_index, rpg_stream_t stream);

/* pose_utils.qexp (pose_utils.py:340-348): q[i] = [cos|v|, sinc(|v|/pi) v] for n log-quaternions v [n, 3] -> q [n, 4]. */
int rpg_qexp(const float* v, int64_t n, float* q, rpg_stream_t stream);
/* Evaluation composition (test.py:227-243) for G graphs sharing the template: with k = ref_k the template edge
 * (src(k) -> node 0) chosen by the caller (the reference takes the ref_node-th edge whose destination is node 0),
 *   out = poses[g*N + src(k)] - pred_edges[g*Ep + k];  out_pred[g] = [out_t * pose_s + pose_m | qexp(out_q)]   [G, 7]
 *   out_targ[g] = the same map applied to the ground truth of node 0 (may be NULL).
 * pose_m / pose_s: HOST float[3] (NULL = 0 / 1), the dataset's translation normalisation. */
int rpg_eval_compose(const float* pred_edges, const float* poses, const rpg_graph_t* graph, int ref_k,
                     const float* pose_m, const float* pose_s, float* out_pred, float* out_targ, rpg_stream_t stream);

/* Train-time edge dropout applied to a per-edge tensor (train.py:238-247: the tiled mask indexes data.edge_attr, and
 * -- documented intent -- data.edge_index): keeps the rows of the surviving template slots of every graph,
 *   out[g * Ep_kept + j, :] = in[g * Ep_full + kept_idx[j], :]     rows of row_bytes bytes (multiple of 4).
 * kept_idx: device int32 [Ep_kept], ascending slot indices of the kept directed edges. */
int rpg_edge_mask_apply(const void* in, int64_t G, int Ep_full, int Ep_kept, const int32_t* kept_idx, int row_bytes,
                        void* out, rpg_stream_t stream);

/* Per-edge gather of node rows through the template (simpleConv, my_gnn_layer.py:394-412: its first Linear acts on
 * cat[x_i, x_j] only, so it factors into two per-node products and this gather):
 *   out[e] = act(pa[node_a(e)] + pb[node_b(e)] + bias) * bit(e),   which_x: 0 = source, 1 = destination of edge e;
 * pb, bias, mask_bits ([Et, D/8] bit pattern), out_bits (pattern of the result) may be NULL; relu: 0 / 1. */
int rpg_edge_gather(const rpg_bf16* pa, int lda, int which_a, const rpg_bf16* pb, int ldb, int which_b, const float* bias,
                    const rpg_graph_t* graph, int D, int relu, const uint8_t* mask_bits, rpg_bf16* out, int ldo,
                    uint8_t* out_bits, rpg_stream_t stream);

/* out[r, :] = v[r, :] * scale[r % mod]   (bf16 rows; the 1/deg of a mean's backward). */
int rpg_scale_rows(const rpg_bf16* v, int ldv, int64_t rows, int D, const float* scale, int mod, rpg_bf16* out, int ldo,
                   rpg_stream_t stream);

/* Evaluation error metrics (test.py:202-203, 262-265) over n pose pairs [n, 7] = (t, unit quaternion):
 * t_err[i] = ||t_pred - t_gt||, q_err[i] = 2 acos(min(1, |<q_pred, q_gt>|)) in degrees (pose_utils.py:420-431). */
int rpg_pose_errors(const float* pred7, const float* targ7, int64_t n, float* t_err, float* q_err, rpg_stream_t stream);

/* Column sums (bias gradients): out[c] (+)= sum_r w[r % mod] * v[r, c]; deterministic; row_w may be NULL. */
int rpg_colsum_bf16(const rpg_bf16* v, int ldv, int64_t rows, int cols, const float* row_w, int row_w_mod,
                    float* out, int accumulate, float* scratch, rpg_stream_t stream);
int64_t rpg_colsum_scratch_floats(int64_t rows, int cols);

/* ------------------------------------------------------------------------------------------
 * The layer: simpleConvEdge_upt.forward (my_gnn_layer.py:293-311) and its backward
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int D;                          /* in = edge = out channels (every reference call site)  */
  /* packed bf16 operands, forward */
  const rpg_bf16* Wn;             /* [3D, D]  rows: edge_mlp.0[:,0:D] | edge_mlp.0[:,D:2D] | mlp.0[:,0:D] */
  const rpg_bf16* W1e_e;          /* [D, D]   edge_mlp.0[:, 2D:3D]                                       */
  const rpg_bf16* W2e;            /* [D, D]   edge_mlp.2                                                 */
  const rpg_bf16* W1m_e;          /* [D, D]   mlp.0[:, D:2D]                                             */
  const rpg_bf16* W2m;            /* [D, D]   mlp.2                                                      */
  const rpg_bf16* Wgtp;           /* [3c, D]  att.g | att.theta | att.phi                                */
  const rpg_bf16* WW;             /* [D, c]   att.W                                                      */
  const rpg_bf16* WWI;            /* [D, pad64(c) + D] = [att.W | I]: z = [y | m] [W | I]^T carries the residual
                                     of att.py:33 through the TMA/MMA pipeline instead of the epilogue         */
  const rpg_bf16* W1u;            /* [D, 2D]  mlp_updating.0                                             */
  const rpg_bf16* W2u;            /* [D, D]   mlp_updating.2                                             */
  /* transposed copies, backward (dgrad) */
  const rpg_bf16* WnT;            /* [D, 3D] */
  const rpg_bf16* W1e_eT, *W2eT, *W1m_eT, *W2mT, *W2uT;   /* [D, D]                                      */
  const rpg_bf16* WgtpT;          /* [D, 3c] */
  const rpg_bf16* WWT;            /* [c, D]  */
  const rpg_bf16* W1uT;           /* [2D, D] */
  /* fp32 biases */
  const float* b1e, *b2e, *b1m, *b2m, *bgtp, *bW, *b1u, *b2u;
  /* The message m = h2 W2m^T + b2m (my_gnn_layer.py:282) only ever enters LINEAR maps -- the attention projections
   * (att.py:20-24) and, through z = W(y) + m, the mean over incoming edges -- so it is never materialised:
   *   (g | theta | phi) = h2 Wgc^T + bgc      with Wgc = Wgtp W2m, bgc = Wgtp b2m + bgtp   (composed in fp32, rpg_compose)
   *   mean(z) = [mean(y) | mean(h2)] [WW | W2m]^T + (bW + b2m)
   * and the backward follows the same factorisation (dh2 = (dgtp Wgc + (dan W2m)[dst]) * [h2 > 0]; the weight gradients of
   * mlp.2 / att.{g,theta,phi} from T = dgtp^T h2 and small fp32 products with the master weights).                */
  const rpg_bf16* Wgc;            /* [3c, D]            composed attention projection                            */
  const rpg_bf16* WgcT;           /* [D, pad64(3c)]     its transpose (dgrad)                                    */
  const rpg_bf16* WWM;            /* [D, pad64(c) + D]  [att.W | mlp.2]                                          */
  const float* bgc;               /* [3c]                                                                        */
  const float* bWm;               /* [D]                bW + b2m                                                 */
  const float* Wgtp_f32;          /* [3c, D] fp32       att.g | att.theta | att.phi master weights, concatenated */
  const float* W2m_f32;           /* [D, D]  fp32       mlp.2.weight (master)                                    */
  /* 0 = simpleConvEdge_upt (above).
   * 1 = simpleConvEdge (my_gnn_layer.py:242-274): message = att(mlp(cat[x_i, x_j, e'])) and NO update MLP, the layer
   *     output is the mean itself (acts.a).  Then Wn is [4D, D] with rows edge_mlp.0[:,0:D] | edge_mlp.0[:,D:2D] |
   *     mlp.0[:,D:2D] (x_j = source) | mlp.0[:,0:D] (x_i = destination), WnT is [D, 4D], W1m_e = mlp.0[:,2D:3D],
   *     acts.P / grads.dP are [Nt, 4D], g_mlp0_w is [D, 3D]; W1u, W2u, h3, out, dh3, dxu and the g_upd* are unused. */
  int variant;
} rpg_layer_weights_t;

typedef struct {                  /* activations of one layer call; all bf16 unless noted               */
  const rpg_bf16* x;              /* [Nt, D] in                                                         */
  const rpg_bf16* e;              /* [Et, D] in                                                         */
  rpg_bf16* P;                    /* [Nt, 3D] scratch: x*Wn^T                                           */
  rpg_bf16* h1;                   /* [Et, D]                                                            */
  rpg_bf16* e_new;                /* [Et, D] out (pre-ReLU)                                             */
  rpg_bf16* e_new_relu;           /* [Et, D] optional: relu(e_new) for the caller (posenet.py:1065)     */
  rpg_bf16* h2;                   /* [Et, D]                                                            */
  rpg_bf16* m;                    /* [Et, D]                                                            */
  float* gtp;                     /* [Et, 3c] fp32                                                      */
  rpg_bf16* y;                    /* [Et, max(c,64)]                                                    */
  rpg_bf16* z;                    /* unused (kept for layout stability): z = W(y) + m is never materialised   */
  rpg_bf16* a;                    /* [Nt, D]  mean over in-edges of z = mean(y) WW^T + bW + mean(m)           */
  rpg_bf16* h3;                   /* [Nt, D]                                                            */
  rpg_bf16* out;                  /* [Nt, D] out (pre-ReLU)                                             */
  rpg_bf16* out_relu;             /* [Nt, D] optional relu(out) (posenet.py:1064)                       */
  /* ReLU bit patterns [rows, D/8 bytes] written by the forward and consumed by the backward epilogues    */
  float* att_aux;                 /* [Et, 4c] fp32 optional: attention row statistics kept for the backward  */
  uint8_t* h1_bits;               /* [Et, D/8] */
  uint8_t* h2_bits;               /* [Et, D/8] */
  uint8_t* h3_bits;               /* [Nt, D/8] */
  uint8_t* e_new_bits;            /* [Et, D/8] optional: (e_new > 0) for the next round's mask_de        */
  uint8_t* out_bits;              /* [Nt, D/8] optional: (out > 0) for the next round's mask_dx          */
  const uint8_t* x_bits;          /* [Nt, D/8] optional pattern of the input x (used when mask_dx)       */
  const uint8_t* e_bits;          /* [Et, D/8] optional pattern of the input e (used when mask_de)       */
  rpg_bf16* ybar;                 /* [Nt, max(c,64)] scratch: mean over in-edges of y                    */
  rpg_bf16* mbar;                 /* [Nt, D]         scratch: mean over in-edges of m                    */
  /* feature dropout of the stack's last round fused into the producing GEMMs (rpg_gemm_t.drop_*): out_relu and
   * e_new_relu are then the DROPPED, rescaled features the pose heads consume, out_bits / e_new_bits their patterns. */
  uint64_t drop_seed_x, drop_seed_e;
  float drop_p;                   /* 0 = off */
  rpg_bf16* gtp16;                /* [Et, 3c] bf16 optional: the attention projections in bf16 instead of `gtp` (series
                                     attention only; then `gtp` may be NULL)                                       */
} rpg_layer_acts_t;

int rpg_layer_fwd(const rpg_layer_weights_t* w, const rpg_graph_t* graph, const rpg_layer_acts_t* t,
                  rpg_stream_t stream);

typedef struct {
  const rpg_bf16* d_out;          /* [Nt, D] grad wrt out (pre-ReLU) or NULL (= zeros)                  */
  const rpg_bf16* d_e_new;        /* [Et, D] grad wrt e_new (pre-ReLU) or NULL                          */
  int mask_dx;                    /* multiply dx by (x > 0): the caller's x is a ReLU output            */
  int mask_de;                    /* multiply de by (e > 0)                                             */
  rpg_bf16* dx;                   /* [Nt, D] out                                                        */
  rpg_bf16* de;                   /* [Et, D] out                                                        */
  /* scratch */
  rpg_bf16* dh3;                  /* [Nt, D]    */
  rpg_bf16* dxu;                  /* [Nt, D]    */
  rpg_bf16* dan;                  /* [Nt, D]    */
  float* dyn;                     /* [Nt, c] fp32 */
  rpg_bf16* dgtp;                 /* [Et, 3c]   */
  rpg_bf16* dm;                   /* [Et, D]    */
  rpg_bf16* dh2;                  /* [Et, D]    */
  rpg_bf16* de_tot;               /* [Et, D]    */
  rpg_bf16* dh1;                  /* [Et, D]    */
  rpg_bf16* dP;                   /* [Nt, 3D]   */
  rpg_bf16* ysum;                 /* [Nt, max(c,64)] */
  float* split_ws;                /* fp32 split-R workspace, rpg_layer_bwd_ws_floats() elements         */
  float* colsum_ws;               /* rpg_colsum_scratch_floats(Et, D) floats                            */
  float* gtp_bias_tmp;            /* [pad64(3c)] fp32: column sums of dgtp                              */
  rpg_bf16* Q;                    /* [Nt, D]    dan W2m (node level)                                    */
  rpg_bf16* h2sum;                /* [Nt, D]    deg * mean(h2) = sum over in-edges of h2                */
  float* T_tmp;                   /* [3c, D] fp32: dgtp^T h2                                            */
  /* fp32 weight gradients in the reference's state_dict layout, accumulated (+=) */
  float* g_mlp0_w;  float* g_mlp0_b;      /* [D, 2D], [D]   mlp.0            */
  float* g_mlp2_w;  float* g_mlp2_b;      /* [D, D]         mlp.2            */
  float* g_upd0_w;  float* g_upd0_b;      /* [D, 2D]        mlp_updating.0   */
  float* g_upd2_w;  float* g_upd2_b;      /* [D, D]         mlp_updating.2   */
  float* g_edge0_w; float* g_edge0_b;     /* [D, 3D]        edge_model.edge_mlp.0 */
  float* g_edge2_w; float* g_edge2_b;     /* [D, D]         edge_model.edge_mlp.2 */
  float* g_att_g_w; float* g_att_g_b;     /* [c, D]         att.g / theta / phi */
  float* g_att_theta_w; float* g_att_theta_b;
  float* g_att_phi_w; float* g_att_phi_b;
  float* g_att_W_w; float* g_att_W_b;     /* [D, c], [D]    att.W            */
  /* 0: dxu and dan are separate [Nt, D] tensors and dan = da / deg (mean backward applied by the producing GEMM).
   * 2D (simpleConvEdge_upt only): dxu and dan are the column halves of ONE [Nt, 2D] tensor (dan == dxu + D, row pitch
   * 2D) written by a single GEMM [dx_u | da] = dh3 W1u; dan then holds da UNSCALED and the 1/deg is applied where da
   * is consumed (row scale of the dyn and Q GEMMs; the weight gradients contract da with the forward's means ybar,
   * mbar instead of the in-edge sums; the bias sums skip nodes without in-edges).  h2sum / ysum are not used. */
  int dxa_ld;
} rpg_layer_grads_t;

/* ------------------------------------------------------------------------------------------
 * fp32 mode ("split bf16", BASELINE config B; forward / inference in this round)
 * An fp32 value is carried as two bf16 planes (hi = bf16(v), lo = bf16(v - hi)); a Linear becomes
 * [A_hi | A_lo | A_hi] [W_hi | W_hi | W_lo]^T on the same tcgen05 kernel (fp32 accumulation), which keeps
 * ~16 mantissa bits per operand (single-pass TF32 keeps 10 and cannot meet the 1e-4 tolerance).
 * ---------------------------------------------------------------------------------------- */
int rpg_cast_f32_to_split(const float* src, rpg_bf16* hi, rpg_bf16* lo, int64_t n, rpg_stream_t stream);
int rpg_split_to_f32(const rpg_bf16* hi, const rpg_bf16* lo, float* dst, int64_t n, rpg_stream_t stream);
/* low plane of a weight window (the high plane is rpg_pack_weight's output)                          */
int rpg_pack_weight_lo(const float* src, int ld_src, int r0, int c0, int rows, int cols, rpg_bf16* dst,
                       int ld_dst, rpg_stream_t stream);
int rpg_edge_init_fwd_f32(const float* pminmax, int ldp, const float* bias, const rpg_graph_t* graph, int D,
                          rpg_bf16* e_hi, rpg_bf16* e_lo, int lde, rpg_stream_t stream);
/* the same with the ReLU bit pattern of e0 [Et, D/8] for the backward (e_bits may be NULL) */
int rpg_edge_init_fwd_split(const float* pminmax, int ldp, const float* bias, const rpg_graph_t* graph, int D,
                            rpg_bf16* e_hi, rpg_bf16* e_lo, int lde, uint8_t* e_bits, rpg_stream_t stream);
/* rpg_head_bwd for (hi, lo) features; dfeat as (hi, lo) planes */
int rpg_head_bwd_split(const float* dpose, const rpg_bf16* feat_hi, const rpg_bf16* feat_lo, int ldf, int64_t rows, int D,
                       const uint8_t* keep, uint64_t seed, float p_drop, const float* w6, int mask_relu, rpg_bf16* dfeat_hi,
                       rpg_bf16* dfeat_lo, int lddf, float* dw_t, float* dw_q, float* db_t, float* db_q, int accumulate,
                       float* ws, rpg_stream_t stream);
int rpg_aggregate_mean_split(const rpg_bf16* z_hi, const rpg_bf16* z_lo, int ldz, const rpg_graph_t* graph, int D,
                             rpg_bf16* a_hi, rpg_bf16* a_lo, int lda, rpg_stream_t stream);

typedef struct {
  int D;
  /* every operand is [N, 3K] = [W_hi | W_hi | W_lo] along K (W1u3: [D, 6D] = x part then a part)      */
  const rpg_bf16 *Wn3, *W1e_e3, *W2e3, *W1m_e3, *W2m3, *Wgtp3, *WW3, *W1u3, *W2u3;
  const float *b1e, *b2e, *b1m, *b2m, *bgtp, *bW, *b1u, *b2u;
  /* composed operands (the message m is never materialised, see rpg_layer_weights_t): Wgc3 [3c, 3D] of Wgtp W2m,
   * WWM3 [D, 3 pad64(c) + 3D] = [WW3 | W2m3], bgc [3c] = Wgtp b2m + bgtp, bWm [D] = bW + b2m                     */
  const rpg_bf16 *Wgc3, *WWM3;
  const float *bgc, *bWm;
  /* backward (dgrad) operands, [N, 3K] = [W^T_hi | W^T_hi | W^T_lo]: WnT3 [D, 9D], W1uT3 [2D, 3D] (x rows, then a rows),
   * WgcT3 [D, 3 pad64(3c)], WWT3 [c, 3D], the others [D, 3D]; NULL for inference                                  */
  const rpg_bf16 *WnT3, *W1e_eT3, *W2eT3, *W1m_eT3, *W2mT3, *WgcT3, *WWT3, *W1uT3, *W2uT3;
  const rpg_bf16 *WnT3_sd;                  /* [D, 6D]: WnT3 restricted to the two edge-MLP blocks (rounds without d_out) */
  const float *Wgtp_f32, *W2m_f32;          /* master weights for the small fp32 products of the backward          */
} rpg_layer_weights_split_t;

typedef struct {
  const rpg_bf16 *x_hi, *x_lo, *e_hi, *e_lo;        /* inputs  [Nt, D], [Et, D]                          */
  float* P;                                        /* [Nt, 3D] fp32 node projections (epilogue-gather path) */
  rpg_bf16 *h1_hi, *h1_lo, *e_new_hi, *e_new_lo, *e_new_relu_hi, *e_new_relu_lo, *h2_hi, *h2_lo, *m_hi, *m_lo;
  float* gtp;                                      /* [Et, 3c] fp32                                     */
  rpg_bf16 *y_hi, *y_lo;                           /* [Et, pad64(c)]                                    */
  rpg_bf16 *z_hi, *z_lo, *a_hi, *a_lo, *h3_hi, *h3_lo, *out_hi, *out_lo, *out_relu_hi, *out_relu_lo;
  rpg_bf16* ybar_hi, *ybar_lo;    /* [Nt, max(c,64)] scratch: mean over in-edges of y                  */
  rpg_bf16* mbar_hi, *mbar_lo;    /* [Nt, D]         scratch: mean over in-edges of m   (z is never materialised) */
  rpg_bf16* P_hi, *P_lo;          /* [Nt, 3D] node projections as (hi, lo) planes: with selection patterns in the graph
                                     the gathered terms are one-hot K panels over both planes (P may then be NULL)  */
  /* ReLU bit patterns for the backward (optional, as in rpg_layer_acts_t)                                          */
  uint8_t *h1_bits, *h2_bits, *h3_bits, *e_new_bits, *out_bits;
  const uint8_t *x_bits, *e_bits;
} rpg_layer_acts_split_t;

int rpg_layer_fwd_split(const rpg_layer_weights_split_t* w, const rpg_graph_t* graph,
                        const rpg_layer_acts_split_t* t, rpg_stream_t stream);

/* Backward of rpg_layer_fwd_split: the sequence of rpg_layer_bwd with every value a (hi, lo) pair, every dgrad GEMM the
 * 3-segment form [dY_hi | dY_lo | dY_hi] [W^T_hi | W^T_hi | W^T_lo]^T and every weight gradient three TN launches
 * (hi^T hi + lo^T hi + hi^T lo, fp32 partials folded in a fixed order).  split_ws: 3 * rpg_layer_bwd_ws_floats().      */
typedef struct {
  const rpg_bf16 *d_out_hi, *d_out_lo;        /* [Nt, D] grad wrt out (pre-ReLU) or NULL                            */
  const rpg_bf16 *d_e_new_hi, *d_e_new_lo;    /* [Et, D] grad wrt e_new (pre-ReLU) or NULL                          */
  int mask_dx, mask_de;
  rpg_bf16 *dx_hi, *dx_lo, *de_hi, *de_lo;    /* outputs                                                            */
  /* scratch (shapes as rpg_layer_grads_t) */
  rpg_bf16 *dh3_hi, *dh3_lo, *dxu_hi, *dxu_lo, *dan_hi, *dan_lo;
  float* dyn;
  rpg_bf16 *dgtp_hi, *dgtp_lo, *Q_hi, *Q_lo, *dh2_hi, *dh2_lo, *de_tot_hi, *de_tot_lo, *dh1_hi, *dh1_lo, *dP_hi, *dP_lo;
  rpg_bf16 *ysum_hi, *ysum_lo, *h2sum_hi, *h2sum_lo;
  float *split_ws, *colsum_ws, *gtp_bias_tmp, *T_tmp;
  float* Q_f32;                               /* [Nt, D] fp32: dan W2m for templates without selection patterns     */
  /* fp32 weight gradients in the reference's state_dict layout, accumulated (+=) */
  float* g_mlp0_w;  float* g_mlp0_b;  float* g_mlp2_w;  float* g_mlp2_b;
  float* g_upd0_w;  float* g_upd0_b;  float* g_upd2_w;  float* g_upd2_b;
  float* g_edge0_w; float* g_edge0_b; float* g_edge2_w; float* g_edge2_b;
  float* g_att_g_w; float* g_att_g_b; float* g_att_theta_w; float* g_att_theta_b; float* g_att_phi_w; float* g_att_phi_b;
  float* g_att_W_w; float* g_att_W_b;
} rpg_layer_grads_split_t;
int rpg_layer_bwd_split(const rpg_layer_weights_split_t* w, const rpg_graph_t* graph, const rpg_layer_acts_split_t* t,
                        const rpg_layer_grads_split_t* g, rpg_stream_t stream);
/* split-plane forms of the bandwidth kernels the fp32-mode training step needs */
int rpg_segment_sum_split(const rpg_bf16* v_hi, const rpg_bf16* v_lo, int ldv, const int32_t* csr_ptr, const int32_t* csr_idx,
                          const float* scale, const rpg_graph_t* graph, int D, rpg_bf16* out_hi, rpg_bf16* out_lo, int ldo,
                          rpg_stream_t stream);
int rpg_attention_bwd_split(const float* gtp, const float* dyn, int ld_dyn, const rpg_graph_t* graph, int64_t Et, int c,
                            rpg_bf16* dgtp_hi, rpg_bf16* dgtp_lo, int ld_dgtp, rpg_stream_t stream);
/* dW[M, N] += (A_hi + A_lo)^T (B_hi + B_lo) to first order (the lo^T lo term, 2^-18 relative, is dropped), and
 * bias[m] += column sums of A_hi + A_lo (bias may be NULL); ws: 3 * rpg_layer_bwd_ws_floats() floats.               */
int rpg_wgrad_split(const rpg_bf16* A_hi, const rpg_bf16* A_lo, int lda, int M, const rpg_bf16* B_hi, const rpg_bf16* B_lo,
                    int ldb, int N, int64_t R, float* ws, float* out, int ldo, float* bias, rpg_stream_t stream);

int64_t rpg_layer_bwd_ws_floats(int D, int64_t Et, int64_t Nt);
int rpg_layer_bwd(const rpg_layer_weights_t* w, const rpg_graph_t* graph, const rpg_layer_acts_t* t,
                  const rpg_layer_grads_t* g, rpg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RPG_H_ */
