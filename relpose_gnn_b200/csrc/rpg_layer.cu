// Host-side sequencing of one simpleConvEdge_upt call (my_gnn_layer.py:293-311) and its backward over the
// tcgen05 GEMM and the bandwidth kernels, plus the small C-ABI utilities.  No host synchronisation: every
// kernel is enqueued on the caller's stream.
//
// Concatenated inputs are never materialised.  [x_src | x_dst | e] W^T is evaluated as
//   P_s[src] + P_d[dst] + e W_e^T   with   [P_s | P_d | P_m] = x [W_s; W_d; W_m]^T   computed once per NODE,
// which is algebraically identical to the reference formulation and removes the gathers and 40 % of the FLOPs.
#include <cuda_runtime.h>

#include <atomic>
#include <cstddef>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "../../include/rpg.h"
#include "rpg_internal.h"

namespace rpg {

static thread_local char g_err[512] = "";

int set_error(int code, const char* msg) {
    snprintf(g_err, sizeof g_err, "%s", msg ? msg : "");
    return code;
}

static std::atomic<long long> g_launches{0};

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

static inline int pad64(int v) { return (v + 63) / 64 * 64; }

// Convenience builder for the NT GEMM  out = epi(A[M,K] * B[N,K]^T)
static rpg_gemm_t nt(int M, int N, const rpg_bf16* A, int K, int lda, const rpg_bf16* B, int ldb) {
    rpg_gemm_t g;
    memset(&g, 0, sizeof g);
    g.mode = 0; g.M = M; g.N = N; g.n_seg = 1;
    g.A[0] = A; g.K[0] = K; g.lda[0] = lda; g.B = B; g.ldb = ldb;
    return g;
}

// Partial products of dW[M,N] = A[R,M]^T B[R,N] through the split-R TN kernel; returns the split count (> 0)
// or an error (< 0 / cudaError as negative is impossible, so errors are reported through *rc).
static int wgrad_partials(const rpg_bf16* A, int lda, int M, const rpg_bf16* B, int ldb, int N, long long R, float* ws,
                          int sm_count, cudaStream_t s, int* splits_out, bool with_colsum = false) {
    rpg_gemm_t g;
    memset(&g, 0, sizeof g);
    g.mode = 1; g.M = M; g.N = N; g.A[0] = A; g.lda[0] = lda; g.B = B; g.ldb = ldb; g.R = (int)R;
    const int block_n = N >= 256 ? 256 : pad64(N);
    g.block_n = block_n;
    const int tiles = ((M + 127) / 128) * ((N + block_n - 1) / block_n);
    const long long kb = (R + 63) / 64;
    long long splits = sm_count / tiles;                              // one wave of work items: fp32 partials cost HBM traffic
    if (splits > kb) splits = kb;
    if (splits < 1) splits = 1;
    const long long per = (kb + splits - 1) / splits;
    splits = (kb + per - 1) / per;                                    // no empty splits
    g.splits = (int)splits;
    g.split_stride = (long long)M * N;
    g.out_f32 = ws; g.ldo_f32 = N;
    if (with_colsum) g.a_colsum = ws + (size_t)g.splits * M * N;       // [splits, M] right after the dW partials
    *splits_out = g.splits;
    return gemm_launch(&g, s);
}

}  // namespace rpg
bool rpg::pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("RPG_PDL");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}
namespace rpg {

static int sm_count_cached() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

// Weight gradients of one layer backward: every TN launch writes its split partials into its own slice of the workspace
// and queues the fold; all folds run as ONE launch at the end (fixed order per element => deterministic).
// RPG_TN_GROUP_MAX=k caps the problems per grouped launch (A/B comparisons; default: TN_GROUP_MAX)
static int tn_group_limit() {
    static int v = 0;
    if (!v) {
        const char* e = getenv("RPG_TN_GROUP_MAX");
        v = e ? atoi(e) : TN_GROUP_MAX;
        if (v < 1 || v > TN_GROUP_MAX) v = TN_GROUP_MAX;
    }
    return v;
}

struct WgradQueue {
    float* ws;
    size_t off;
    int sms;
    cudaStream_t s;
    rpg_reduce_batch_t batch;
    TnDesc pend[TN_GROUP_MAX];       // weight-gradient products queued for ONE grouped launch (rpg_gemm_tn.cu)
    int npend;
    WgradQueue(float* ws_, int sms_, cudaStream_t s_) : ws(ws_), off(0), sms(sms_), s(s_), npend(0) { batch.n = 0; }
    int add(const float* part, int splits, long long stride, int rows, int cols, float* out, int ldo) {
        if (batch.n == RPG_REDUCE_BATCH_MAX) {
            int rc = flush();
            if (rc) return rc;
        }
        rpg_reduce_desc_t& d = batch.d[batch.n++];
        d.part = part; d.out = out; d.stride = stride; d.splits = splits; d.rows = rows; d.cols = cols; d.ldo = ldo;
        return 0;
    }
    // partials of A^T B at ws + off: queued for the grouped pair kernel when the shapes qualify, launched at once otherwise
    int product(const rpg_bf16* A, int lda, int M, const rpg_bf16* B, int ldb, int N, long long R, bool with_colsum,
                float** part, int* splits) {
        float* p = ws + off;
        *part = p;
        if (tn_group_supported(A, lda, M, B, ldb, N, R)) {
            if (npend >= tn_group_limit()) {
                int rc = launch();
                if (rc) return rc;
            }
            *splits = tn_group_splits(M, N, R, sms);
            TnDesc& d = pend[npend++];
            d.A = A; d.lda = lda; d.M = M; d.B = B; d.ldb = ldb; d.N = N; d.R = R; d.part = p; d.splits = *splits;
            d.colsum = with_colsum ? p + (size_t)*splits * M * N : nullptr;
        } else {
            int rc = wgrad_partials(A, lda, M, B, ldb, N, R, p, sms, s, splits, with_colsum);
            if (rc) return rc;
        }
        off += ((size_t)*splits * M * N + (with_colsum ? (size_t)*splits * M : 0) + 63) & ~(size_t)63;
        return 0;
    }
    // dW (+)= A^T B over R rows, optionally bias (+)= colsum(A); partials only, fold queued
    int wgrad(const rpg_bf16* A, int lda, int M, const rpg_bf16* B, int ldb, int N, long long R, float* out, int ldo,
              float* bias = nullptr) {
        int splits = 0;
        float* p = nullptr;
        int rc = product(A, lda, M, B, ldb, N, R, bias != nullptr, &p, &splits);
        if (rc) return rc;
        if ((rc = add(p, splits, (long long)M * N, M, N, out, ldo))) return rc;
        if (bias) rc = add(p + (size_t)splits * M * N, splits, M, 1, M, bias, M);
        return rc;
    }
    // partials of a stacked product whose row blocks go to different parameters: the caller queues the folds
    int partials(const rpg_bf16* A, int lda, int M, const rpg_bf16* B, int ldb, int N, long long R, bool with_colsum,
                 float** part, int* splits) {
        return product(A, lda, M, B, ldb, N, R, with_colsum, part, splits);
    }
    // launches the queued products (their partials exist on the stream afterwards)
    int launch() {
        if (!npend) return 0;
        int rc = tn_group_launch(pend, npend, s);
        npend = 0;
        return rc;
    }
    int flush() {
        int rc = launch();
        if (rc) return rc;
        if (!batch.n) return 0;
        rc = rpg_reduce_splits_batch(&batch, (rpg_stream_t)s);
        batch.n = 0;
        return rc;
    }
};

}  // namespace rpg

using namespace rpg;

extern "C" {

const char* rpg_last_error_string(void) { return g_err; }
int rpg_version(void) { return 100; }

int rpg_device_sm_count(int device, int* sm_count) {
    if (!sm_count) return set_error(RPG_E_ARG, "sm_count: null");
    cudaError_t e = cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
    return 0;
}

int rpg_gemm(const rpg_gemm_t* g, rpg_stream_t stream) { return gemm_launch(g, as_stream(stream)); }

int rpg_set_gemm_cluster(int cl) { return set_gemm_cluster(cl); }

int rpg_attention_series_enabled(void) { return attention_series_enabled() ? 1 : 0; }

int64_t rpg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int rpg_profile_begin(void) { return profile_begin(); }
int rpg_profile_end(double* nt_ms, double* tn_ms, int* nt_launches, int* tn_launches, double* nt_flops, double* tn_flops) {
    return profile_end(nt_ms, tn_ms, nt_launches, tn_launches, nt_flops, tn_flops);
}

int rpg_wgrad(const rpg_bf16* A, int lda, int M, const rpg_bf16* B, int ldb, int N, int64_t R, float* ws, float* out, int ldo,
              rpg_stream_t stream) {
    if (!A || !B || !ws || !out) return set_error(RPG_E_ARG, "wgrad: null pointer");
    WgradQueue q(ws, sm_count_cached(), as_stream(stream));          // a group of one: the CTA-pair kernel when the shape qualifies
    int rc = q.wgrad(A, lda, M, B, ldb, N, R, out, ldo);
    return rc ? rc : q.flush();
}

int rpg_wgrad_bias(const rpg_bf16* A, int lda, int M, const rpg_bf16* B, int ldb, int N, int64_t R, float* ws, float* out,
                   int ldo, float* bias, rpg_stream_t stream) {
    if (!A || !B || !ws || !out) return set_error(RPG_E_ARG, "wgrad_bias: null pointer");
    WgradQueue q(ws, sm_count_cached(), as_stream(stream));
    int rc = q.wgrad(A, lda, M, B, ldb, N, R, out, ldo, bias);
    return rc ? rc : q.flush();
}

int rpg_wgrad_blocks(const rpg_bf16* A, int lda, int M, int nblocks, const rpg_bf16* B, int ldb, int N, int64_t R, float* ws,
                     float* const* outs, int ldo, float* bias, rpg_stream_t stream) {
    if (!A || !B || !ws || !outs || nblocks < 1 || nblocks > 4 || M <= 0) return set_error(RPG_E_ARG, "wgrad_blocks: bad arguments");
    for (int i = 0; i < nblocks; ++i)
        if (!outs[i]) return set_error(RPG_E_ARG, "wgrad_blocks: null output");
    WgradQueue q(ws, sm_count_cached(), as_stream(stream));
    const int Mp = M * nblocks;
    float* part = nullptr;
    int splits = 0;
    int rc = q.partials(A, lda, Mp, B, ldb, N, R, bias != nullptr, &part, &splits);
    for (int i = 0; i < nblocks && !rc; ++i)
        rc = q.add(part + (size_t)i * M * N, splits, (long long)Mp * N, M, N, outs[i], ldo);
    if (!rc && bias) rc = q.add(part + (size_t)splits * Mp * N, splits, Mp, 1, M, bias, M);
    return rc ? rc : q.flush();
}

void rpg_struct_sizes(int32_t* out) {
    out[0] = (int32_t)sizeof(rpg_graph_t);
    out[1] = (int32_t)sizeof(rpg_gemm_t);
    out[2] = (int32_t)sizeof(rpg_layer_weights_t);
    out[3] = (int32_t)sizeof(rpg_layer_acts_t);
    out[4] = (int32_t)sizeof(rpg_layer_grads_t);
    out[5] = (int32_t)offsetof(rpg_gemm_t, out_f32);
    out[6] = (int32_t)offsetof(rpg_layer_grads_t, g_mlp0_w);
    out[7] = (int32_t)offsetof(rpg_layer_weights_t, b1e);
    out[8] = (int32_t)sizeof(rpg_layer_weights_split_t);
    out[9] = (int32_t)sizeof(rpg_layer_acts_split_t);
    out[10] = (int32_t)sizeof(rpg_pack_desc_t);
    out[11] = (int32_t)sizeof(rpg_pack_batch_t);
    out[12] = (int32_t)sizeof(rpg_prof_rec_t);
    out[13] = (int32_t)sizeof(rpg_sgemm_batch_t);
    out[14] = (int32_t)sizeof(rpg_layer_grads_split_t);
    out[15] = 0;
}

#define RPG_TRY(expr)            \
    do {                         \
        int rc__ = (expr);       \
        if (rc__) return rc__;   \
    } while (0)

int rpg_layer_fwd(const rpg_layer_weights_t* w, const rpg_graph_t* gr, const rpg_layer_acts_t* t, rpg_stream_t stream) {
    if (!w || !gr || !t) return set_error(RPG_E_ARG, "layer_fwd: null argument");
    const int D = w->D, c = D / 8, c3 = 3 * c;
    if (D % 128) return set_error(RPG_E_UNSUPPORTED, "layer_fwd: channel count must be a multiple of 128");
    const long long Nt = (long long)gr->G * gr->N, Et = (long long)gr->G * gr->Ep;
    if (Nt > 0x7fffffff || Et > 0x7fffffff) return set_error(RPG_E_UNSUPPORTED, "layer_fwd: more than 2^31 rows");
    cudaStream_t s = as_stream(stream);
    const int cp = pad64(c);
    rpg_gemm_t g;

    // (1) per-node projections  P = x [W1e_src; W1e_dst; W1m_src (; W1m_dst)]^T          [Nt, 3D] ([Nt, 4D] for variant 1)
    const int np = w->variant == 1 ? 4 : 3, ldP = np * D;
    g = nt((int)Nt, ldP, t->x, D, D, w->Wn, D);
    g.out = t->P; g.ldo = ldP;
    RPG_TRY(gemm_launch(&g, s));

    // (2) edge MLP layer 1 (my_gnn_layer.py:232,237-238): h1 = relu(e W1e_e^T + P_s[src] + P_d[dst] + b)
    g = nt((int)Et, D, t->e, D, D, w->W1e_e, D);
    g.bias = w->b1e;
    if (gr->sel_src && gr->sel_dst) {            // gathers as one-hot K panels (plain epilogue)
        g.n_gseg = 2; g.gsel_patterns = gr->sel_patterns; g.gsel_div = gr->sel_div; g.gsrc_rows = (int)Nt;
        g.gsel[0] = gr->sel_src; g.gsrc[0] = t->P;     g.gsrc_ld[0] = ldP;
        g.gsel[1] = gr->sel_dst; g.gsrc[1] = t->P + D; g.gsrc_ld[1] = ldP;
    } else {
        g.gadd[0] = t->P;     g.gmap[0] = gr->src; g.gadd_ld[0] = ldP;
        g.gadd[1] = t->P + D; g.gmap[1] = gr->dst; g.gadd_ld[1] = ldP;
    }
    g.Ep = gr->Ep; g.Nn = gr->N; if (g.n_gseg && gr->pg_Ep) { g.Ep = gr->pg_Ep; g.Nn = gr->pg_N; } g.relu = 1;
    g.out = t->h1; g.ldo = D; g.out_bits = t->h1_bits; g.out_bits_ld = D / 8;
    RPG_TRY(gemm_launch(&g, s));

    // (3) edge MLP layer 2 (my_gnn_layer.py:234): e' = h1 W2e^T + b  (+ relu'd copy for the caller, posenet.py:1065)
    g = nt((int)Et, D, t->h1, D, D, w->W2e, D);
    g.bias = w->b2e; g.out = t->e_new; g.out_relu = t->e_new_relu; g.ldo = D;
    g.out_bits = t->e_new_bits; g.out_bits_ld = D / 8;
    if (t->drop_p > 0.f && t->e_new_relu) { g.drop_p = t->drop_p; g.drop_seed = t->drop_seed_e; }
    RPG_TRY(gemm_launch(&g, s));

    // (4) message MLP layer 1 (my_gnn_layer.py:280,305): h2 = relu(e' W1m_e^T + P_m[src] + b)   (x_j = source)
    g = nt((int)Et, D, t->e_new, D, D, w->W1m_e, D);
    g.bias = w->b1m;
    //     variant 1 (my_gnn_layer.py:269-270): + P_mi[dst]   (x_i = destination)
    if (gr->sel_src && (np == 3 || gr->sel_dst)) {
        g.n_gseg = np - 2; g.gsel_patterns = gr->sel_patterns; g.gsel_div = gr->sel_div; g.gsrc_rows = (int)Nt;
        g.gsel[0] = gr->sel_src; g.gsrc[0] = t->P + 2 * D; g.gsrc_ld[0] = ldP;
        if (np == 4) { g.gsel[1] = gr->sel_dst; g.gsrc[1] = t->P + 3 * D; g.gsrc_ld[1] = ldP; }
    } else {
        g.gadd[0] = t->P + 2 * D; g.gmap[0] = gr->src; g.gadd_ld[0] = ldP;
        if (np == 4) { g.gadd[1] = t->P + 3 * D; g.gmap[1] = gr->dst; g.gadd_ld[1] = ldP; }
    }
    g.Ep = gr->Ep; g.Nn = gr->N; if (g.n_gseg && gr->pg_Ep) { g.Ep = gr->pg_Ep; g.Nn = gr->pg_N; } g.relu = 1;
    g.out = t->h2; g.ldo = D; g.out_bits = t->h2_bits; g.out_bits_ld = D / 8;
    RPG_TRY(gemm_launch(&g, s));

    // (5) message MLP layer 2 (my_gnn_layer.py:282): m = h2 W2m^T + b2m is NEVER materialised -- it only enters linear
    //     maps (the attention projections and, through z = W(y) + m, the mean over incoming edges), so both are taken
    //     straight from h2 with composed weights (rpg.h: rpg_layer_weights_t.Wgc / WWM).
    // (6) attention projections (att.py:20-24): (g | theta | phi) = m Wgtp^T + b = h2 (Wgtp W2m)^T + (Wgtp b2m + bgtp)
    if (!w->Wgc || !w->bgc || !w->WWM || !w->bWm) return set_error(RPG_E_ARG, "layer_fwd: composed operands (Wgc / WWM) missing");
    const bool gtp_bf16 = t->gtp16 != nullptr && attention_series_enabled() && c % 16 == 0 && c <= 256;
    if (!gtp_bf16 && !t->gtp) return set_error(RPG_E_ARG, "layer_fwd: gtp (fp32) or gtp16 needed");
    g = nt((int)Et, c3, t->h2, D, D, w->Wgc, D);
    g.bias = w->bgc;
    if (gtp_bf16) { g.out = t->gtp16; g.ldo = c3; }
    else { g.out_f32 = t->gtp; g.ldo_f32 = c3; }
    RPG_TRY(gemm_launch(&g, s));

    // (7) rank-1 softmax attention (att.py:25-30)
    if (gtp_bf16) RPG_TRY(rpg_attention_fwd_bf16(t->gtp16, Et, c, t->y, cp, stream));
    else RPG_TRY(rpg_attention_fwd(t->gtp, Et, c, t->y, cp, nullptr, attention_series_enabled() ? nullptr : t->att_aux, stream));

    // (8)+(9) z = y WW^T + bW + m (att.py:32-33) is only ever averaged over the incoming edges (PyG aggregate [3p],
    //     my_gnn_layer.py:301), and the mean is linear: a = mean(y) WW^T + bW + mean(h2) W2m^T + b2m.  So neither the
    //     edge-level GEMMs nor the [Et, D] tensors m, z exist: two segment means (y is only c wide) and ONE node-level
    //     GEMM [ybar | h2bar] [WW | W2m]^T + (bW + b2m), zeroed for nodes without incoming edges (their mean is 0).
    if (!t->ybar || !t->mbar || !gr->has_in) return set_error(RPG_E_ARG, "layer_fwd: ybar / mbar / has_in missing");
    RPG_TRY(rpg_aggregate_mean(t->y, cp, gr, cp, t->ybar, cp, stream));
    RPG_TRY(rpg_aggregate_mean(t->h2, D, gr, D, t->mbar, D, stream));      // mbar holds mean(h2)
    g = nt((int)Nt, D, t->ybar, cp, cp, w->WWM, cp + D);
    g.n_seg = 2; g.A[1] = t->mbar; g.K[1] = D; g.lda[1] = D;
    g.bias = w->bWm; g.row_scale = gr->has_in; g.row_scale_mod = gr->N; g.out = t->a; g.ldo = D;
    RPG_TRY(gemm_launch(&g, s));
    if (w->variant == 1) return 0;               // simpleConvEdge: the mean is the layer output (no update step)

    // (10) update MLP (my_gnn_layer.py:284-286,309-311): out = relu([x | a] W1u^T + b) W2u^T + b
    g = nt((int)Nt, D, t->x, D, D, w->W1u, 2 * D);
    g.n_seg = 2; g.A[1] = t->a; g.K[1] = D; g.lda[1] = D;
    g.bias = w->b1u; g.relu = 1; g.out = t->h3; g.ldo = D; g.out_bits = t->h3_bits; g.out_bits_ld = D / 8;
    RPG_TRY(gemm_launch(&g, s));
    g = nt((int)Nt, D, t->h3, D, D, w->W2u, D);
    g.bias = w->b2u; g.out = t->out; g.out_relu = t->out_relu; g.ldo = D;
    g.out_bits = t->out_bits; g.out_bits_ld = D / 8;
    if (t->drop_p > 0.f && t->out_relu) { g.drop_p = t->drop_p; g.drop_seed = t->drop_seed_x; }
    RPG_TRY(gemm_launch(&g, s));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 mode (forward only in this round): every activation is a (hi, lo) bf16 pair and every Linear is evaluated as
//   [A_hi | A_lo | A_hi] [W_hi | W_hi | W_lo]^T   (the lo*lo term, 2^-18 relative, is dropped)
// on the same tcgen05 kernel with fp32 accumulation; node projections stay fp32.  Same sequence as rpg_layer_fwd.
static void set3(rpg_gemm_t& g, int first, const rpg_bf16* hi, const rpg_bf16* lo, int K, int ld) {
    g.A[first] = hi; g.A[first + 1] = lo; g.A[first + 2] = hi;
    for (int i = 0; i < 3; ++i) { g.K[first + i] = K; g.lda[first + i] = ld; }
}

int rpg_layer_fwd_split(const rpg_layer_weights_split_t* w, const rpg_graph_t* gr, const rpg_layer_acts_split_t* t,
                        rpg_stream_t stream) {
    if (!w || !gr || !t) return set_error(RPG_E_ARG, "layer_fwd_split: null argument");
    const int D = w->D, c = D / 8, c3 = 3 * c, cp = pad64(c);
    if (D % 128) return set_error(RPG_E_UNSUPPORTED, "layer_fwd_split: channel count must be a multiple of 128");
    const long long Nt = (long long)gr->G * gr->N, Et = (long long)gr->G * gr->Ep;
    if (Nt > 0x7fffffff || Et > 0x7fffffff) return set_error(RPG_E_UNSUPPORTED, "layer_fwd_split: more than 2^31 rows");
    cudaStream_t s = as_stream(stream);
    rpg_gemm_t g;
    auto base = [&](int M, int N, const rpg_bf16* B, int ldb) {
        memset(&g, 0, sizeof g);
        g.mode = 0; g.M = M; g.N = N; g.B = B; g.ldb = ldb;
    };
    // (1) P = x Wn^T  [Nt, 3D]: (hi, lo) planes when the gathers below can be one-hot K panels (a one-hot row times a
    //     bf16 plane is exact, so two panels per gather add P_hi + P_lo = P to 2^-17), else fp32 for the epilogue gathers
    const bool panels = gr->sel_src && gr->sel_dst && t->P_hi && t->P_lo;
    if (!panels && !t->P) return set_error(RPG_E_ARG, "layer_fwd_split: P (fp32) or P_hi / P_lo with selection patterns needed");
    const int pEp = gr->pg_Ep ? gr->pg_Ep : gr->Ep, pNn = gr->pg_Ep ? gr->pg_N : gr->N;
    base((int)Nt, 3 * D, w->Wn3, 3 * D); g.n_seg = 3; set3(g, 0, t->x_hi, t->x_lo, D, D);
    if (panels) { g.out = t->P_hi; g.out_lo = t->P_lo; g.ldo = 3 * D; }
    else { g.out_f32 = t->P; g.ldo_f32 = 3 * D; }
    RPG_TRY(gemm_launch(&g, s));
    auto panel = [&](int i, const rpg_bf16* sel, const rpg_bf16* plane, int col0) {
        g.gsel[i] = sel; g.gsrc[i] = plane + col0; g.gsrc_ld[i] = 3 * D;
    };
    // (2) h1 = relu(e W1e_e^T + P_s[src] + P_d[dst] + b)
    base((int)Et, D, w->W1e_e3, 3 * D); g.n_seg = 3; set3(g, 0, t->e_hi, t->e_lo, D, D);
    g.bias = w->b1e; g.relu = 1;
    if (panels) {
        g.n_gseg = 4; g.gsel_patterns = gr->sel_patterns; g.gsel_div = gr->sel_div; g.gsrc_rows = (int)Nt; g.Ep = pEp; g.Nn = pNn;
        panel(0, gr->sel_src, t->P_hi, 0); panel(1, gr->sel_src, t->P_lo, 0);
        panel(2, gr->sel_dst, t->P_hi, D); panel(3, gr->sel_dst, t->P_lo, D);
    } else {
        g.Ep = gr->Ep; g.Nn = gr->N;
        g.gadd_f32[0] = t->P;     g.gmap[0] = gr->src; g.gadd_f32_ld[0] = 3 * D;
        g.gadd_f32[1] = t->P + D; g.gmap[1] = gr->dst; g.gadd_f32_ld[1] = 3 * D;
    }
    g.out = t->h1_hi; g.out_lo = t->h1_lo; g.ldo = D; g.out_bits = t->h1_bits; g.out_bits_ld = D / 8;
    RPG_TRY(gemm_launch(&g, s));
    // (3) e' = h1 W2e^T + b  (+ relu'd copy)
    base((int)Et, D, w->W2e3, 3 * D); g.n_seg = 3; set3(g, 0, t->h1_hi, t->h1_lo, D, D);
    g.bias = w->b2e; g.out = t->e_new_hi; g.out_lo = t->e_new_lo; g.ldo = D;
    g.out_relu = t->e_new_relu_hi; g.out_relu_lo = t->e_new_relu_lo; g.out_bits = t->e_new_bits; g.out_bits_ld = D / 8;
    RPG_TRY(gemm_launch(&g, s));
    // (4) h2 = relu(e' W1m_e^T + P_m[src] + b)
    base((int)Et, D, w->W1m_e3, 3 * D); g.n_seg = 3; set3(g, 0, t->e_new_hi, t->e_new_lo, D, D);
    g.bias = w->b1m; g.relu = 1;
    if (panels) {
        g.n_gseg = 2; g.gsel_patterns = gr->sel_patterns; g.gsel_div = gr->sel_div; g.gsrc_rows = (int)Nt; g.Ep = pEp; g.Nn = pNn;
        panel(0, gr->sel_src, t->P_hi, 2 * D); panel(1, gr->sel_src, t->P_lo, 2 * D);
    } else {
        g.Ep = gr->Ep; g.Nn = gr->N;
        g.gadd_f32[0] = t->P + 2 * D; g.gmap[0] = gr->src; g.gadd_f32_ld[0] = 3 * D;
    }
    g.out = t->h2_hi; g.out_lo = t->h2_lo; g.ldo = D; g.out_bits = t->h2_bits; g.out_bits_ld = D / 8;
    RPG_TRY(gemm_launch(&g, s));
    // (5)+(6) the message m = h2 W2m^T + b2m is never materialised (see rpg_layer_fwd): (g | theta | phi) = h2 Wgc^T + bgc
    if (!w->Wgc3 || !w->WWM3 || !w->bgc || !w->bWm) return set_error(RPG_E_ARG, "layer_fwd_split: composed operands missing");
    base((int)Et, c3, w->Wgc3, 3 * D); g.n_seg = 3; set3(g, 0, t->h2_hi, t->h2_lo, D, D);
    g.bias = w->bgc; g.out_f32 = t->gtp; g.ldo_f32 = c3;
    RPG_TRY(gemm_launch(&g, s));
    // (7) attention -> y (hi, lo)
    RPG_TRY(rpg_attention_fwd(t->gtp, Et, c, t->y_hi, cp, t->y_lo, nullptr, stream));
    // (8)+(9) a = mean(y) WW^T + mean(h2) W2m^T + (bW + b2m), zero for nodes without incoming edges
    if (!t->ybar_hi || !t->ybar_lo || !t->mbar_hi || !t->mbar_lo || !gr->has_in)
        return set_error(RPG_E_ARG, "layer_fwd_split: ybar / mbar planes or has_in missing");
    RPG_TRY(rpg_aggregate_mean_split(t->y_hi, t->y_lo, cp, gr, cp, t->ybar_hi, t->ybar_lo, cp, stream));
    RPG_TRY(rpg_aggregate_mean_split(t->h2_hi, t->h2_lo, D, gr, D, t->mbar_hi, t->mbar_lo, D, stream));   // mean(h2)
    base((int)Nt, D, w->WWM3, 3 * cp + 3 * D); g.n_seg = 6;
    set3(g, 0, t->ybar_hi, t->ybar_lo, cp, cp); set3(g, 3, t->mbar_hi, t->mbar_lo, D, D);
    g.bias = w->bWm; g.row_scale = gr->has_in; g.row_scale_mod = gr->N;
    g.out = t->a_hi; g.out_lo = t->a_lo; g.ldo = D;
    RPG_TRY(gemm_launch(&g, s));
    // (10) out = relu([x | a] W1u^T + b) W2u^T + b
    base((int)Nt, D, w->W1u3, 6 * D); g.n_seg = 6; set3(g, 0, t->x_hi, t->x_lo, D, D); set3(g, 3, t->a_hi, t->a_lo, D, D);
    g.bias = w->b1u; g.relu = 1; g.out = t->h3_hi; g.out_lo = t->h3_lo; g.ldo = D; g.out_bits = t->h3_bits; g.out_bits_ld = D / 8;
    RPG_TRY(gemm_launch(&g, s));
    base((int)Nt, D, w->W2u3, 3 * D); g.n_seg = 3; set3(g, 0, t->h3_hi, t->h3_lo, D, D);
    g.bias = w->b2u; g.out = t->out_hi; g.out_lo = t->out_lo; g.ldo = D;
    g.out_relu = t->out_relu_hi; g.out_relu_lo = t->out_relu_lo; g.out_bits = t->out_bits; g.out_bits_ld = D / 8;
    RPG_TRY(gemm_launch(&g, s));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// fp32-mode backward.  Same sequence as rpg_layer_bwd below; every value is a (hi, lo) bf16 pair.
namespace {
struct Pl { const rpg_bf16* hi; const rpg_bf16* lo; };
struct PlW { rpg_bf16* hi; rpg_bf16* lo; };

// out = epi([A_hi | A_lo | A_hi] B3^T) with B3 = [W_hi | W_hi | W_lo] ([N, 3K], pitch ldb)
rpg_gemm_t nt3(int M, int N, Pl A, int K, int lda, const rpg_bf16* B3, int ldb) {
    rpg_gemm_t g;
    memset(&g, 0, sizeof g);
    g.mode = 0; g.M = M; g.N = N; g.n_seg = 3; g.B = B3; g.ldb = ldb;
    g.A[0] = A.hi; g.A[1] = A.lo; g.A[2] = A.hi;
    for (int i = 0; i < 3; ++i) { g.K[i] = K; g.lda[i] = lda; }
    return g;
}

// Three TN launches per weight gradient (hi^T hi, lo^T hi, hi^T lo), each queue with its own workspace third and its own
// fold launch (descriptors of one launch must not share outputs).
struct WgradQueue3 {
    WgradQueue q0, q1, q2;
    WgradQueue3(float* ws, size_t third, int sms, cudaStream_t s) : q0(ws, sms, s), q1(ws + third, sms, s), q2(ws + 2 * third, sms, s) {}
    int wgrad(Pl A, int lda, int M, Pl B, int ldb, int N, long long R, float* out, int ldo, float* bias = nullptr) {
        int rc = q0.wgrad(A.hi, lda, M, B.hi, ldb, N, R, out, ldo, bias);
        if (!rc) rc = q1.wgrad(A.lo, lda, M, B.hi, ldb, N, R, out, ldo, bias);
        if (!rc) rc = q2.wgrad(A.hi, lda, M, B.lo, ldb, N, R, out, ldo, nullptr);
        return rc;
    }
    int partials(Pl A, int lda, int M, Pl B, int ldb, int N, long long R, bool with_colsum, float* part[3], int splits[3]) {
        int rc = q0.partials(A.hi, lda, M, B.hi, ldb, N, R, with_colsum, &part[0], &splits[0]);
        if (!rc) rc = q1.partials(A.lo, lda, M, B.hi, ldb, N, R, with_colsum, &part[1], &splits[1]);
        if (!rc) rc = q2.partials(A.hi, lda, M, B.lo, ldb, N, R, false, &part[2], &splits[2]);
        return rc;
    }
    // folds of a stacked product: the same (row block -> parameter) mapping for the three launches
    int add(float* const part[3], const int splits[3], size_t block_off, long long stride, int rows, int cols, float* out, int ldo) {
        int rc = q0.add(part[0] + block_off, splits[0], stride, rows, cols, out, ldo);
        if (!rc) rc = q1.add(part[1] + block_off, splits[1], stride, rows, cols, out, ldo);
        if (!rc) rc = q2.add(part[2] + block_off, splits[2], stride, rows, cols, out, ldo);
        return rc;
    }
    int launch() {
        int rc = q0.launch();
        if (!rc) rc = q1.launch();
        if (!rc) rc = q2.launch();
        return rc;
    }
    int flush() {
        int rc = launch();
        if (!rc) rc = q0.flush();
        if (!rc) rc = q1.flush();
        if (!rc) rc = q2.flush();
        return rc;
    }
};
}  // namespace

int rpg_wgrad_split(const rpg_bf16* A_hi, const rpg_bf16* A_lo, int lda, int M, const rpg_bf16* B_hi, const rpg_bf16* B_lo,
                    int ldb, int N, int64_t R, float* ws, float* out, int ldo, float* bias, rpg_stream_t stream) {
    if (!A_hi || !A_lo || !B_hi || !B_lo || !ws || !out) return set_error(RPG_E_ARG, "wgrad_split: null pointer");
    WgradQueue3 q(ws, (size_t)rpg_layer_bwd_ws_floats(M > N ? M : N, 0, 0), sm_count_cached(), as_stream(stream));
    RPG_TRY(q.wgrad({A_hi, A_lo}, lda, M, {B_hi, B_lo}, ldb, N, R, out, ldo, bias));
    return q.flush();
}

int rpg_layer_bwd_split(const rpg_layer_weights_split_t* w, const rpg_graph_t* gr, const rpg_layer_acts_split_t* t,
                        const rpg_layer_grads_split_t* b, rpg_stream_t stream) {
    if (!w || !gr || !t || !b) return set_error(RPG_E_ARG, "layer_bwd_split: null argument");
    if (!w->WnT3 || !w->W2eT3 || !w->WgcT3) return set_error(RPG_E_ARG, "layer_bwd_split: backward operands missing");
    const int D = w->D, c = D / 8, c3 = 3 * c, cp = pad64(c), c3p = pad64(c3);
    if (D % 128) return set_error(RPG_E_UNSUPPORTED, "layer_bwd_split: channel count must be a multiple of 128");
    const long long Nt = (long long)gr->G * gr->N, Et = (long long)gr->G * gr->Ep;
    cudaStream_t s = as_stream(stream);
    const int sms = sm_count_cached();
    rpg_gemm_t g;
    const bool have_out = b->d_out_hi != nullptr;
    const int ldP = 3 * D;
    const bool panels = gr->sel_src && gr->sel_dst;
    const int pEp = gr->pg_Ep ? gr->pg_Ep : gr->Ep, pNn = gr->pg_Ep ? gr->pg_N : gr->N;
    if (!panels && have_out && !b->Q_f32) return set_error(RPG_E_ARG, "layer_bwd_split: Q_f32 needed for templates without selection patterns");
    auto outp = [&](PlW o) { g.out = o.hi; g.out_lo = o.lo; g.ldo = D; };

    if (have_out) {
        // dh3 = (d_out W2u) * [h3 > 0]
        g = nt3((int)Nt, D, {b->d_out_hi, b->d_out_lo}, D, D, w->W2uT3, 3 * D);
        g.mask_bits = t->h3_bits; g.mask_bits_ld = D / 8; outp({b->dh3_hi, b->dh3_lo});
        RPG_TRY(gemm_launch(&g, s));
        // [dx_u | da] = dh3 W1u ; da / deg -> dan
        g = nt3((int)Nt, D, {b->dh3_hi, b->dh3_lo}, D, D, w->W1uT3, 3 * D);
        outp({b->dxu_hi, b->dxu_lo});
        RPG_TRY(gemm_launch(&g, s));
        g = nt3((int)Nt, D, {b->dh3_hi, b->dh3_lo}, D, D, w->W1uT3 + (size_t)D * 3 * D, 3 * D);
        g.row_scale = gr->inv_deg; g.row_scale_mod = gr->N; outp({b->dan_hi, b->dan_lo});
        RPG_TRY(gemm_launch(&g, s));
        // dyn = dan WW   fp32 [Nt, c]
        g = nt3((int)Nt, c, {b->dan_hi, b->dan_lo}, D, D, w->WWT3, 3 * D);
        g.out_f32 = b->dyn; g.ldo_f32 = c;
        RPG_TRY(gemm_launch(&g, s));
        // attention backward -> dgtp (hi, lo) [Et, pad64(3c)]
        RPG_TRY(rpg_attention_bwd_split(t->gtp, b->dyn, c, gr, Et, c, b->dgtp_hi, b->dgtp_lo, c3p, stream));
        // Q = dan W2m ; dh2 = (dgtp Wgc + Q[dst]) * [h2 > 0]
        g = nt3((int)Nt, D, {b->dan_hi, b->dan_lo}, D, D, w->W2mT3, 3 * D);
        if (panels) outp({b->Q_hi, b->Q_lo});
        else { g.out_f32 = b->Q_f32; g.ldo_f32 = D; }          // epilogue gather of fp32 rows (tiny / irregular templates)
        RPG_TRY(gemm_launch(&g, s));
        g = nt3((int)Et, D, {b->dgtp_hi, b->dgtp_lo}, c3p, c3p, w->WgcT3, 3 * c3p);
        if (panels) {
            g.n_gseg = 2; g.gsel_patterns = gr->sel_patterns; g.gsel_div = gr->sel_div; g.gsrc_rows = (int)Nt; g.Ep = pEp; g.Nn = pNn;
            g.gsel[0] = gr->sel_dst; g.gsrc[0] = b->Q_hi; g.gsrc_ld[0] = D;
            g.gsel[1] = gr->sel_dst; g.gsrc[1] = b->Q_lo; g.gsrc_ld[1] = D;
        } else {
            g.Ep = gr->Ep; g.Nn = gr->N;
            g.gadd_f32[0] = b->Q_f32; g.gmap[0] = gr->dst; g.gadd_f32_ld[0] = D;
        }
        g.mask_bits = t->h2_bits; g.mask_bits_ld = D / 8; outp({b->dh2_hi, b->dh2_lo});
        RPG_TRY(gemm_launch(&g, s));
        // de'_tot = dh2 W1m_e + d_e_new
        g = nt3((int)Et, D, {b->dh2_hi, b->dh2_lo}, D, D, w->W1m_eT3, 3 * D);
        if (b->d_e_new_hi) { g.resid = b->d_e_new_hi; g.resid_lo = b->d_e_new_lo; g.resid_ld = D; }
        outp({b->de_tot_hi, b->de_tot_lo});
        RPG_TRY(gemm_launch(&g, s));
    }
    const Pl de_tot = have_out ? Pl{b->de_tot_hi, b->de_tot_lo} : Pl{b->d_e_new_hi, b->d_e_new_lo};
    if (!de_tot.hi || !de_tot.lo) return set_error(RPG_E_ARG, "layer_bwd_split: neither d_out nor d_e_new given");

    // dh1 = (de'_tot W2e) * [h1 > 0] ; de = dh1 W1e_e (* [e > 0])
    g = nt3((int)Et, D, de_tot, D, D, w->W2eT3, 3 * D);
    g.mask_bits = t->h1_bits; g.mask_bits_ld = D / 8; outp({b->dh1_hi, b->dh1_lo});
    RPG_TRY(gemm_launch(&g, s));
    g = nt3((int)Et, D, {b->dh1_hi, b->dh1_lo}, D, D, w->W1e_eT3, 3 * D);
    if (b->mask_de) { g.mask_bits = t->e_bits; g.mask_bits_ld = D / 8; }
    outp({b->de_hi, b->de_lo});
    RPG_TRY(gemm_launch(&g, s));

    // node side: dP = [sum_src dh1 | sum_dst dh1 | sum_src dh2], dx = dP Wn + dx_u
    RPG_TRY(rpg_segment_sum_split(b->dh1_hi, b->dh1_lo, D, gr->out_ptr, gr->out_idx, nullptr, gr, D, b->dP_hi, b->dP_lo, ldP, stream));
    RPG_TRY(rpg_segment_sum_split(b->dh1_hi, b->dh1_lo, D, gr->in_ptr, gr->in_idx, nullptr, gr, D, b->dP_hi + D, b->dP_lo + D, ldP, stream));
    if (have_out) {
        RPG_TRY(rpg_segment_sum_split(b->dh2_hi, b->dh2_lo, D, gr->out_ptr, gr->out_idx, nullptr, gr, D, b->dP_hi + 2 * D, b->dP_lo + 2 * D, ldP, stream));
        g = nt3((int)Nt, D, {b->dP_hi, b->dP_lo}, ldP, ldP, w->WnT3, 3 * ldP);
        g.resid = b->dxu_hi; g.resid_lo = b->dxu_lo; g.resid_ld = D;
    } else {
        // only the two edge-MLP blocks carry a gradient: K = 2D with the operand cut to those columns
        if (!w->WnT3_sd) return set_error(RPG_E_ARG, "layer_bwd_split: WnT3_sd missing");
        g = nt3((int)Nt, D, {b->dP_hi, b->dP_lo}, 2 * D, ldP, w->WnT3_sd, 6 * D);
    }
    if (b->mask_dx) { g.mask_bits = t->x_bits; g.mask_bits_ld = D / 8; }
    outp({b->dx_hi, b->dx_lo});
    RPG_TRY(gemm_launch(&g, s));

    // ---- weight gradients
    const size_t third = (size_t)rpg_layer_bwd_ws_floats(D, 0, 0);
    WgradQueue3 q(b->split_ws, third, sms, s);
    float* part[3];
    int splits[3];
    RPG_TRY(q.wgrad(de_tot, D, D, {t->h1_hi, t->h1_lo}, D, D, Et, b->g_edge2_w, D, b->g_edge2_b));
    RPG_TRY(q.wgrad({b->dh1_hi, b->dh1_lo}, D, D, {t->e_hi, t->e_lo}, D, D, Et, b->g_edge0_w + 2 * D, 3 * D, b->g_edge0_b));
    {
        const int Mp = have_out ? ldP : 2 * D;
        RPG_TRY(q.partials({b->dP_hi, b->dP_lo}, ldP, Mp, {t->x_hi, t->x_lo}, D, D, Nt, false, part, splits));
        const long long stride = (long long)Mp * D;
        RPG_TRY(q.add(part, splits, 0, stride, D, D, b->g_edge0_w, 3 * D));
        RPG_TRY(q.add(part, splits, (size_t)D * D, stride, D, D, b->g_edge0_w + D, 3 * D));
        if (have_out) RPG_TRY(q.add(part, splits, 2 * (size_t)D * D, stride, D, D, b->g_mlp0_w, 2 * D));
    }
    if (!have_out) return q.flush();
    RPG_TRY(q.wgrad({b->dh2_hi, b->dh2_lo}, D, D, {t->e_new_hi, t->e_new_lo}, D, D, Et, b->g_mlp0_w + D, 2 * D, b->g_mlp0_b));
    {
        // T = dgtp^T h2, csg = colsum(dgtp): three partial sets folded NOW into zeroed scratch (three small launches)
        RPG_TRY(q.partials({b->dgtp_hi, b->dgtp_lo}, c3p, c3, {t->h2_hi, t->h2_lo}, D, D, Et, true, part, splits));
        RPG_TRY(q.launch());
        cudaMemsetAsync(b->T_tmp, 0, (size_t)c3 * D * sizeof(float), s);
        cudaMemsetAsync(b->gtp_bias_tmp, 0, (size_t)c3 * sizeof(float), s);
        for (int i = 0; i < 3; ++i) {
            rpg_reduce_batch_t fold;
            fold.n = i < 2 ? 2 : 1;
            fold.d[0] = {part[i], b->T_tmp, (long long)c3 * D, splits[i], c3, D, D};
            if (i < 2) fold.d[1] = {part[i] + (size_t)splits[i] * c3 * D, b->gtp_bias_tmp, c3, splits[i], 1, c3, c3};
            RPG_TRY(rpg_reduce_splits_batch(&fold, stream));
        }
        rpg_sgemm_batch_t sg;
        memset(&sg, 0, sizeof sg);
        float* dW[3] = {b->g_att_g_w, b->g_att_theta_w, b->g_att_phi_w};
        float* dB[3] = {b->g_att_g_b, b->g_att_theta_b, b->g_att_phi_b};
        for (int i = 0; i < 3; ++i) {
            rpg_sgemm_desc_t& d = sg.d[sg.n++];
            d.A = b->T_tmp + (size_t)i * c * D; d.lda = D; d.B = w->W2m_f32; d.ldb = D; d.transB = 1;
            d.C = dW[i]; d.ldc = D; d.M = c; d.N = D; d.K = D; d.accumulate = 1;
            d.u = b->gtp_bias_tmp + i * c; d.v = w->b2m;
        }
        {
            rpg_sgemm_desc_t& d = sg.d[sg.n++];
            d.A = w->Wgtp_f32; d.lda = D; d.transA = 1; d.B = b->T_tmp; d.ldb = D;
            d.C = b->g_mlp2_w; d.ldc = D; d.M = D; d.N = D; d.K = c3; d.accumulate = 1;
        }
        {
            rpg_sgemm_desc_t& d = sg.d[sg.n++];
            d.A = w->Wgtp_f32; d.lda = D; d.transA = 1; d.B = b->gtp_bias_tmp; d.ldb = 1;
            d.C = b->g_mlp2_b; d.ldc = 1; d.M = D; d.N = 1; d.K = c3; d.accumulate = 1;
        }
        RPG_TRY(rpg_sgemm_batch(&sg, stream));
        // att.{g,theta,phi}.bias += csg  (fp32 copy-add: a 1 x c fold of one "split")
        rpg_reduce_batch_t bb;
        bb.n = 3;
        for (int i = 0; i < 3; ++i) bb.d[i] = {b->gtp_bias_tmp + i * c, dB[i], c, 1, 1, c, c};
        RPG_TRY(rpg_reduce_splits_batch(&bb, stream));
    }
    // node-level parts
    RPG_TRY(rpg_segment_sum_split(t->h2_hi, t->h2_lo, D, gr->in_ptr, gr->in_idx, nullptr, gr, D, b->h2sum_hi, b->h2sum_lo, D, stream));
    RPG_TRY(q.wgrad({b->dan_hi, b->dan_lo}, D, D, {b->h2sum_hi, b->h2sum_lo}, D, D, Nt, b->g_mlp2_w, D));
    RPG_TRY(rpg_segment_sum_split(t->y_hi, t->y_lo, cp, gr->in_ptr, gr->in_idx, nullptr, gr, cp, b->ysum_hi, b->ysum_lo, cp, stream));
    RPG_TRY(q.wgrad({b->dan_hi, b->dan_lo}, D, D, {b->ysum_hi, b->ysum_lo}, cp, c, Nt, b->g_att_W_w, c));
    for (int pass = 0; pass < 2; ++pass) {                  // sum_n deg(n) dan[n] -> att.W.bias and mlp.2.bias (hi + lo planes)
        float* dst = pass ? b->g_mlp2_b : b->g_att_W_b;
        RPG_TRY(rpg_colsum_bf16(b->dan_hi, D, Nt, D, gr->deg, gr->N, dst, 1, b->colsum_ws, stream));
        RPG_TRY(rpg_colsum_bf16(b->dan_lo, D, Nt, D, gr->deg, gr->N, dst, 1, b->colsum_ws, stream));
    }
    RPG_TRY(q.wgrad({b->d_out_hi, b->d_out_lo}, D, D, {t->h3_hi, t->h3_lo}, D, D, Nt, b->g_upd2_w, D, b->g_upd2_b));
    RPG_TRY(q.wgrad({b->dh3_hi, b->dh3_lo}, D, D, {t->x_hi, t->x_lo}, D, D, Nt, b->g_upd0_w, 2 * D, b->g_upd0_b));
    RPG_TRY(q.wgrad({b->dh3_hi, b->dh3_lo}, D, D, {t->a_hi, t->a_lo}, D, D, Nt, b->g_upd0_w + D, 2 * D));
    RPG_TRY(q.flush());
    return 0;
}

int64_t rpg_head_bwd_tc_ws_floats(int D) {
    int sms = sm_count_cached();
    if (sms <= 0 || sms > 160) sms = 160;
    return 64 + (int64_t)sms * 16 * D + (int64_t)sms * 16 + 64;
}

int rpg_head_bwd_tc(const float* dpose, const rpg_bf16* feat_d, int ldf, const uint8_t* bits, int64_t rows, int D,
                    float scale, const rpg_bf16* w6T_ext, rpg_bf16* dp16, rpg_bf16* dfeat, int lddf, float* dw_t,
                    float* dw_q, float* db_t, float* db_q, float* ws, rpg_stream_t stream) {
    if (!dpose || !feat_d || !w6T_ext || !dp16 || !dw_t || !dw_q || !db_t || !db_q || !ws || rows <= 0 || D % 64 || ldf % 8)
        return set_error(RPG_E_ARG, "head_bwd_tc: bad arguments");
    if (rows > 0x7fffffff) return set_error(RPG_E_UNSUPPORTED, "head_bwd_tc: more than 2^31 rows");
    cudaStream_t s = as_stream(stream);
    const int sms = sm_count_cached();
    // K panel [hi | lo] of dpose; ws[0] = scale (1-entry row-scale table)
    RPG_TRY(rpg_pack_dpose(dpose, rows, dp16, scale, ws, stream));
    if (dfeat) {
        // dfeat = scale * (dpose W6) * [feat_d > 0]      (K = 64: slots j and 8 + j both multiply W6[j, :])
        if (!bits) return set_error(RPG_E_ARG, "head_bwd_tc: dfeat needs the bit pattern of the dropped features");
        rpg_gemm_t g = nt((int)rows, D, dp16, 64, 64, w6T_ext, 64);
        g.row_scale = ws; g.row_scale_mod = 1;
        g.mask_bits = bits; g.mask_bits_ld = D / 8;
        g.out = dfeat; g.ldo = lddf;
        RPG_TRY(gemm_launch(&g, s));
    }
    // dW6 = dpose^T feat_d: rows j (hi) and 8 + j (lo) of a [16, D] TN product; db6 = column sums of the panel
    float* part = ws + 64;
    int splits = 0;
    RPG_TRY(wgrad_partials(dp16, 64, 16, feat_d, ldf, D, rows, part, sms, s, &splits, /*with_colsum=*/true));
    const float* cs = part + (size_t)splits * 16 * D;
    const long long stride = 16LL * D;
    for (int half = 0; half < 2; ++half) {           // hi rows, then lo rows: two folds into the same outputs
        rpg_reduce_batch_t b;
        b.n = 4;
        const int r0 = half * 8;
        b.d[0] = {part + (size_t)(r0 + 0) * D, dw_t, stride, splits, 3, D, D};
        b.d[1] = {part + (size_t)(r0 + 3) * D, dw_q, stride, splits, 3, D, D};
        b.d[2] = {cs + r0 + 0, db_t, 16, splits, 1, 3, 3};
        b.d[3] = {cs + r0 + 3, db_q, 16, splits, 1, 3, 3};
        RPG_TRY(rpg_reduce_splits_batch(&b, stream));
    }
    return 0;
}

int64_t rpg_layer_bwd_ws_floats(int D, int64_t Et, int64_t Nt) {
    (void)Et; (void)Nt;
    // 11 TN launches per layer backward, each with its own slice: splits * tiles <= sm_count, so the dW partials of one
    // launch fit sm_count * 128 * 256 floats (+ M * N when a single split is forced, M * N <= 3 D^2), plus [splits, M]
    // column sums (M <= 3 D) and the 64-float rounding.  Sized for up to 160 SMs.
    int sms = sm_count_cached();
    if (sms <= 0 || sms > 160) sms = 160;
    return 11LL * ((long long)sms * 128 * 256 + 4LL * D * D + (long long)sms * 4 * D + 64);
}

int rpg_layer_bwd(const rpg_layer_weights_t* w, const rpg_graph_t* gr, const rpg_layer_acts_t* t,
                  const rpg_layer_grads_t* b, rpg_stream_t stream) {
    if (!w || !gr || !t || !b) return set_error(RPG_E_ARG, "layer_bwd: null argument");
    const int D = w->D, c = D / 8, c3 = 3 * c, cp = pad64(c), c3p = pad64(c3);
    if (D % 128) return set_error(RPG_E_UNSUPPORTED, "layer_bwd: channel count must be a multiple of 128");
    const long long Nt = (long long)gr->G * gr->N, Et = (long long)gr->G * gr->Ep;
    cudaStream_t s = as_stream(stream);
    const int sms = sm_count_cached();
    rpg_gemm_t g;
    const bool have_out = b->d_out != nullptr;
    const bool v1 = w->variant == 1;
    const int np = v1 ? 4 : 3, ldP = np * D;

    if (have_out && v1) {
        // simpleConvEdge: out IS the mean, so dan = d_out / deg directly
        RPG_TRY(rpg_scale_rows(b->d_out, D, Nt, D, gr->inv_deg, gr->N, b->dan, D, stream));
    }
    // ---- update MLP backward (only when a gradient reaches `out`)
    // merged update-MLP dgrad (rpg_layer_grads_t.dxa_ld): da stays unscaled, its consumers apply 1 / deg
    const bool merged = have_out && !v1 && b->dxa_ld != 0;
    if (merged && (b->dxa_ld != 2 * D || b->dan != b->dxu + D || !gr->has_in))
        return set_error(RPG_E_ARG, "layer_bwd: dxa_ld must be 2D with dan == dxu + D (and the graph needs has_in)");
    const int ld_dan = merged ? 2 * D : D;
    if (have_out && !v1) {
        // dh3 = (d_out W2u) * [h3 > 0]
        g = nt((int)Nt, D, b->d_out, D, D, w->W2uT, D);
        if (t->h3_bits) { g.mask_bits = t->h3_bits; g.mask_bits_ld = D / 8; } else { g.mask = t->h3; g.mask_ld = D; }
        g.out = b->dh3; g.ldo = D;
        RPG_TRY(gemm_launch(&g, s));
        if (merged) {
            // [dx_u | da] = dh3 W1u in one launch (N = 2D)
            g = nt((int)Nt, 2 * D, b->dh3, D, D, w->W1uT, D);
            g.out = b->dxu; g.ldo = 2 * D;
            RPG_TRY(gemm_launch(&g, s));
        } else {
            // [dx_u | da] = dh3 W1u ; da is scaled by 1/deg (mean backward) -> dan
            g = nt((int)Nt, D, b->dh3, D, D, w->W1uT, D);
            g.out = b->dxu; g.ldo = D;
            RPG_TRY(gemm_launch(&g, s));
            g = nt((int)Nt, D, b->dh3, D, D, w->W1uT + (size_t)D * D, D);
            g.row_scale = gr->inv_deg; g.row_scale_mod = gr->N; g.out = b->dan; g.ldo = D;
            RPG_TRY(gemm_launch(&g, s));
        }
    }
    if (have_out) {
        // dy per destination node: dyn = dan WW   fp32 [Nt, c]
        g = nt((int)Nt, c, b->dan, D, ld_dan, w->WWT, D);
        if (merged) { g.row_scale = gr->inv_deg; g.row_scale_mod = gr->N; }
        g.out_f32 = b->dyn; g.ldo_f32 = c;
        RPG_TRY(gemm_launch(&g, s));
        // attention backward -> dgtp [Et, 3c]
        if (t->gtp16 != nullptr && attention_series_enabled() && c % 16 == 0 && c <= 256)
            RPG_TRY(rpg_attention_bwd_bf16(t->gtp16, b->dyn, c, gr, Et, c, b->dgtp, c3p, stream));
        else
            RPG_TRY(rpg_attention_bwd(t->gtp, b->dyn, c, gr, Et, c, b->dgtp, c3p, attention_series_enabled() ? nullptr : t->att_aux, stream));
        // dh2 = (dm W2m) * [h2 > 0] with dm = dgtp Wgtp + dan[dst] (never materialised):
        //     dh2 = (dgtp (Wgtp W2m) + Q[dst]) * [h2 > 0],  Q = dan W2m at node level
        if (!b->Q || !w->WgcT) return set_error(RPG_E_ARG, "layer_bwd: Q / WgcT missing");
        g = nt((int)Nt, D, b->dan, D, ld_dan, w->W2mT, D);
        if (merged) { g.row_scale = gr->inv_deg; g.row_scale_mod = gr->N; }
        g.out = b->Q; g.ldo = D;
        RPG_TRY(gemm_launch(&g, s));
        g = nt((int)Et, D, b->dgtp, c3p, c3p, w->WgcT, c3p);
        g.Ep = gr->Ep; g.Nn = gr->N;
        if (gr->sel_dst) {
            g.n_gseg = 1; g.gsel_patterns = gr->sel_patterns; g.gsel_div = gr->sel_div; g.gsrc_rows = (int)Nt;
            g.gsel[0] = gr->sel_dst; g.gsrc[0] = b->Q; g.gsrc_ld[0] = D;
            if (gr->pg_Ep) { g.Ep = gr->pg_Ep; g.Nn = gr->pg_N; }       // per-graph edge sets: window of a member graph
        } else {
            g.gadd[0] = b->Q; g.gmap[0] = gr->dst; g.gadd_ld[0] = D;
        }
        if (t->h2_bits) { g.mask_bits = t->h2_bits; g.mask_bits_ld = D / 8; } else { g.mask = t->h2; g.mask_ld = D; }
        g.out = b->dh2; g.ldo = D;
        RPG_TRY(gemm_launch(&g, s));
        // de'_tot = dh2 W1m_e + d_e_new
        g = nt((int)Et, D, b->dh2, D, D, w->W1m_eT, D);
        g.resid = b->d_e_new; g.resid_ld = D; g.out = b->de_tot; g.ldo = D;
        RPG_TRY(gemm_launch(&g, s));
    }
    const rpg_bf16* de_tot = have_out ? b->de_tot : b->d_e_new;
    if (!de_tot) return set_error(RPG_E_ARG, "layer_bwd: neither d_out nor d_e_new given");

    // ---- edge MLP backward
    // dh1 = (de'_tot W2e) * [h1 > 0]
    g = nt((int)Et, D, de_tot, D, D, w->W2eT, D);
    if (t->h1_bits) { g.mask_bits = t->h1_bits; g.mask_bits_ld = D / 8; } else { g.mask = t->h1; g.mask_ld = D; }
    g.out = b->dh1; g.ldo = D;
    RPG_TRY(gemm_launch(&g, s));
    // de = dh1 W1e_e  (optionally * [e > 0] for a ReLU'd input)
    g = nt((int)Et, D, b->dh1, D, D, w->W1e_eT, D);
    if (b->mask_de) {
        if (t->e_bits) { g.mask_bits = t->e_bits; g.mask_bits_ld = D / 8; } else { g.mask = t->e; g.mask_ld = D; }
    }
    g.out = b->de; g.ldo = D;
    RPG_TRY(gemm_launch(&g, s));

    // ---- node side: dP = [sum_src dh1 | sum_dst dh1 | sum_src dh2], dx = dP Wn + dx_u
    RPG_TRY(rpg_segment_sum2(b->dh1, D, gr, /*by source*/ 1, /*by destination*/ 0, D, b->dP, ldP, b->dP + D, ldP, stream));
    if (have_out) {
        RPG_TRY(rpg_edge_to_node_sum(b->dh2, D, gr, D, 1, b->dP + 2 * D, ldP, stream));
        if (v1) RPG_TRY(rpg_edge_to_node_sum(b->dh2, D, gr, D, 0, b->dP + 3 * D, ldP, stream));
        g = nt((int)Nt, D, b->dP, ldP, ldP, w->WnT, ldP);
        if (!v1) { g.resid = b->dxu; g.resid_ld = merged ? 2 * D : D; }
    } else {
        g = nt((int)Nt, D, b->dP, 2 * D, ldP, w->WnT, ldP);
    }
    if (b->mask_dx) {
        if (t->x_bits) { g.mask_bits = t->x_bits; g.mask_bits_ld = D / 8; } else { g.mask = t->x; g.mask_ld = D; }
    }
    g.out = b->dx; g.ldo = D;
    RPG_TRY(gemm_launch(&g, s));

    // ---- weight gradients (fp32, accumulated into the reference's state_dict layout)
    WgradQueue q(b->split_ws, sms, s);
    float* part = nullptr;
    int splits = 0;
    // edge_model.edge_mlp.2: dW = de'_tot^T h1 ; db = colsum(de'_tot)
    RPG_TRY(q.wgrad(de_tot, D, D, t->h1, D, D, Et, b->g_edge2_w, D, b->g_edge2_b));
    // edge_model.edge_mlp.0, edge columns [2D,3D): dW = dh1^T e ; db = colsum(dh1)
    RPG_TRY(q.wgrad(b->dh1, D, D, t->e, D, D, Et, b->g_edge0_w + 2 * D, 3 * D, b->g_edge0_b));
    // node-side blocks in one launch: dP^T x = [edge_mlp.0[:,0:D]; edge_mlp.0[:,D:2D]; mlp.0[:,0:D]]
    {
        const int Mp = have_out ? ldP : 2 * D;
        RPG_TRY(q.partials(b->dP, ldP, Mp, t->x, D, D, Nt, false, &part, &splits));
        const long long stride = (long long)Mp * D;
        RPG_TRY(q.add(part, splits, stride, D, D, b->g_edge0_w, 3 * D));
        RPG_TRY(q.add(part + (size_t)D * D, splits, stride, D, D, b->g_edge0_w + D, 3 * D));
        if (have_out && !v1) RPG_TRY(q.add(part + 2 * (size_t)D * D, splits, stride, D, D, b->g_mlp0_w, 2 * D));
        if (have_out && v1) {                     // mlp.0 = [x_i (dst) | x_j (src) | e']: block 2 is x_j, block 3 is x_i
            RPG_TRY(q.add(part + 2 * (size_t)D * D, splits, stride, D, D, b->g_mlp0_w + D, 3 * D));
            RPG_TRY(q.add(part + 3 * (size_t)D * D, splits, stride, D, D, b->g_mlp0_w, 3 * D));
        }
    }
    if (have_out) {
        // mlp.0 edge columns [D,2D): dW = dh2^T e'
        RPG_TRY(q.wgrad(b->dh2, D, D, t->e_new, D, D, Et, b->g_mlp0_w + (v1 ? 2 * D : D), v1 ? 3 * D : 2 * D, b->g_mlp0_b));
        // mlp.2 and att.g / theta / phi through the factorisation dm = dgtp Wgtp + dan[dst], m = h2 W2m^T + b2m:
        //   T = dgtp^T h2 [3c, D], csg = colsum(dgtp)                      (one TN launch over the edge rows)
        //   d att.{g,theta,phi}.weight = T W2m^T + csg b2m^T ;  bias = csg
        //   d mlp.2.weight = Wgtp^T T + dan^T h2sum ;  bias = Wgtp^T csg + sum_n deg(n) dan[n]
        // with h2sum[n] = sum over in-edges of h2 = deg(n) * mean(h2)[n] (the forward's node-level mean).
        if (!b->T_tmp || (!merged && !b->h2sum) || !w->Wgtp_f32 || !w->W2m_f32) return set_error(RPG_E_ARG, "layer_bwd: T_tmp / h2sum / master weights missing");
        RPG_TRY(q.partials(b->dgtp, c3p, c3, t->h2, D, D, Et, true, &part, &splits));
        // node-level parts: dan^T h2sum -> mlp.2.weight, dan^T ysum -> att.W.weight, sum_n deg(n) dan[n] -> both biases
        if (merged) {
            // (da / deg)^T (deg * mean) = da^T mean: the forward's means as they are (0 for nodes without in-edges),
            // and sum_n deg(n) dan[n] = sum over the nodes WITH in-edges of da[n]
            RPG_TRY(q.wgrad(b->dan, ld_dan, D, t->mbar, D, D, Nt, b->g_mlp2_w, D));
            RPG_TRY(q.wgrad(b->dan, ld_dan, D, t->ybar, cp, c, Nt, b->g_att_W_w, c));
            RPG_TRY(colsum_bf16_2(b->dan, ld_dan, Nt, D, gr->has_in, gr->N, b->g_att_W_b, b->g_mlp2_b, 1, b->colsum_ws, s));
        } else {
            RPG_TRY(rpg_scale_rows(t->mbar, D, Nt, D, gr->deg, gr->N, b->h2sum, D, stream));
            RPG_TRY(q.wgrad(b->dan, D, D, b->h2sum, D, D, Nt, b->g_mlp2_w, D));
            RPG_TRY(rpg_scale_rows(t->ybar, cp, Nt, cp, gr->deg, gr->N, b->ysum, cp, stream));
            RPG_TRY(q.wgrad(b->dan, D, D, b->ysum, cp, c, Nt, b->g_att_W_w, c));
            RPG_TRY(colsum_bf16_2(b->dan, D, Nt, D, gr->deg, gr->N, b->g_att_W_b, b->g_mlp2_b, 1, b->colsum_ws, s));
        }
        if (!v1) {
            // mlp_updating.2: dW = d_out^T h3 ; mlp_updating.0: dW = dh3^T [x | a]
            RPG_TRY(q.wgrad(b->d_out, D, D, t->h3, D, D, Nt, b->g_upd2_w, D, b->g_upd2_b));
            RPG_TRY(q.wgrad(b->dh3, D, D, t->x, D, D, Nt, b->g_upd0_w, 2 * D, b->g_upd0_b));
            RPG_TRY(q.wgrad(b->dh3, D, D, t->a, D, D, Nt, b->g_upd0_w + D, 2 * D));
        }
    }
    if (have_out) {
            // (after every product of this layer has been queued: ONE grouped launch carries all of them)
            const float* cs = part + (size_t)splits * c3 * D;            // [splits, 3c] column sums of dgtp
            RPG_TRY(q.add(cs, splits, c3, 1, c, b->g_att_g_b, c));
            RPG_TRY(q.add(cs + c, splits, c3, 1, c, b->g_att_theta_b, c));
            RPG_TRY(q.add(cs + 2 * c, splits, c3, 1, c, b->g_att_phi_b, c));
            // T and csg are needed NOW by the small fp32 products: the queued products run, then their own fold launch
            // into zeroed scratch
            RPG_TRY(q.launch());
            cudaMemsetAsync(b->T_tmp, 0, (size_t)c3 * D * sizeof(float), s);
            cudaMemsetAsync(b->gtp_bias_tmp, 0, (size_t)c3 * sizeof(float), s);
            rpg_reduce_batch_t fold;
            fold.n = 2;
            fold.d[0] = {part, b->T_tmp, (long long)c3 * D, splits, c3, D, D};
            fold.d[1] = {cs, b->gtp_bias_tmp, c3, splits, 1, c3, c3};
            RPG_TRY(rpg_reduce_splits_batch(&fold, stream));
            rpg_sgemm_batch_t sg;
            memset(&sg, 0, sizeof sg);
            float* dW[3] = {b->g_att_g_w, b->g_att_theta_w, b->g_att_phi_w};
            for (int i = 0; i < 3; ++i) {                                   // [c, D] += T_i W2m^T + csg_i b2m^T
                rpg_sgemm_desc_t& d = sg.d[sg.n++];
                d.A = b->T_tmp + (size_t)i * c * D; d.lda = D; d.B = w->W2m_f32; d.ldb = D; d.transB = 1;
                d.C = dW[i]; d.ldc = D; d.M = c; d.N = D; d.K = D; d.accumulate = 1;
                d.u = b->gtp_bias_tmp + i * c; d.v = w->b2m;
            }
            {                                                               // mlp.2.weight [D, D] += Wgtp^T T
                rpg_sgemm_desc_t& d = sg.d[sg.n++];
                d.A = w->Wgtp_f32; d.lda = D; d.transA = 1; d.B = b->T_tmp; d.ldb = D;
                d.C = b->g_mlp2_w; d.ldc = D; d.M = D; d.N = D; d.K = c3; d.accumulate = 1;
            }
            {                                                               // mlp.2.bias [D] += Wgtp^T csg
                rpg_sgemm_desc_t& d = sg.d[sg.n++];
                d.A = w->Wgtp_f32; d.lda = D; d.transA = 1; d.B = b->gtp_bias_tmp; d.ldb = 1;
                d.C = b->g_mlp2_b; d.ldc = 1; d.M = D; d.N = 1; d.K = c3; d.accumulate = 1;
            }
            RPG_TRY(rpg_sgemm_batch(&sg, stream));
        }
    RPG_TRY(q.flush());
    return 0;
}

}  // extern "C"
