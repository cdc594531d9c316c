// tcgen05 / TMEM / TMA GEMM with fused epilogues -- the tensor-core workhorse of the path.
//
// One persistent, warp-specialised kernel (1 CTA per SM, 320 threads):
//   warp 0      TMA producer   : cp.async.bulk.tensor tiles (128B swizzle) into a 4-stage smem ring
//   warp 1      MMA issuer     : one thread issues tcgen05.mma (M=128, N=block_n, K=16), fp32 accumulators in TMEM,
//                                two accumulator stages of 256 columns so the epilogue of tile i overlaps tile i+1
//   warps 2..9  epilogue       : tcgen05.ld -> registers -> fused bias / gathered adds / residual / ReLU(-backward)
//                                -> bf16 / fp32 stores (each thread owns one output row: 64 B contiguous per chunk)
// MODE 0 (NT): C[M,N] = sum_s A_s[M,K_s] * B[N,K]^T, both operands K-major            (all forward + dgrad GEMMs)
// MODE 1 (TN): C[M,N] = sum_r A[r,M]^T * B[r,N], both operands MN-major, split over r  (weight gradients)
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "../../include/rpg.h"
#include "rpg_internal.h"
#include "rpg_ptx.cuh"

namespace rpg {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int MAX_BLOCK_N = 256;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;        // 16 KB
constexpr int B_STAGE_BYTES = MAX_BLOCK_N * BLOCK_K * 2;    // 32 KB
constexpr int ACC_STAGES = 2;
constexpr int ACC_STRIDE = 256;      // TMEM columns per accumulator stage
constexpr int TMEM_COLS = 512;
constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = 32 * (2 + EPI_WARPS);
constexpr int SMEM_RING_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
constexpr int SMEM_BAR_BYTES = 1024;                        // barriers + TMEM slot (keeps the staging 1024-aligned)
constexpr int EPI_STAGE_BYTES = 32 * 128;                   // per epilogue warp: 32 rows x 128 B, TMA 128B-swizzle layout
constexpr int SMEM_BYTES = SMEM_RING_BYTES + SMEM_BAR_BYTES + EPI_WARPS * EPI_STAGE_BYTES + 1024;   // + align slack
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct GemmKParams {
    int mode, M, N, block_n;
    int num_kb[6];                 // NT: 64-wide k-blocks per A segment
    int total_kb;                  // NT: sum; TN: ceil(R / 64)
    int splits, kb_per_split;      // TN
    long long split_stride;
    const float* bias;
    const __nv_bfloat16* gadd[2];
    const int* gmap[2];
    int gadd_ld[2];
    int Ep, Nn;
    const __nv_bfloat16* resid;
    int resid_ld;
    const float* row_scale;
    int row_scale_mod;
    const __nv_bfloat16* mask;
    int mask_ld;
    int relu;
    __nv_bfloat16* out;
    __nv_bfloat16* out_relu;
    int ldo;
    float* out_f32;
    int ldo_f32;
    const uint8_t* mask_bits;      // ReLU pattern as 1 bit per element: [M, mask_bits_ld bytes]
    int mask_bits_ld;
    uint8_t* out_bits;             // pattern (value > 0) of the stored result
    int out_bits_ld;
    // fp32 ("split bf16") mode: values are (hi, lo) bf16 pairs, v = hi + lo
    const float* gadd_f32[2];      // gathered fp32 node rows (same gmap as gadd)
    int gadd_f32_ld[2];
    const __nv_bfloat16* resid_lo; // low plane of resid
    __nv_bfloat16* out_lo;         // low plane of out:      bf16(v - float(bf16(v)))
    __nv_bfloat16* out_relu_lo;    // low plane of out_relu
    // gathered node adds as one-hot K panels (one extra 64-wide k-block each): A = selection pattern of this row block,
    // B = the node rows of the gathered matrix (MN-major)
    int n_gseg;                    // 0..2
    int gsel_div;                  // pattern index of row block = ((m_blk * 128) % Ep) / gsel_div
    // TN mode: column sums of A (= bias gradient when A is dY) accumulated by the otherwise idle epilogue warps
    float* a_colsum;               // [splits, M] partial sums, or NULL
    int resid_tma;                 // pair kernels: the residual tile arrives by TMA in a per-warp operand buffer
    unsigned long long drop_seed;  // fused feature dropout of the out_relu output (plain epilogue)
    uint32_t drop_thresh;          // keep iff hash byte >= thresh; 0 = off
    float drop_scale;              // 1 / (1 - p)
    int l2_prefetch;               // producer pulls the NEXT work item's streamed tiles into L2 while it feeds this one
};

struct GemmTmaps {
    CUtensorMap a[6];              // A segments (NT) / a[0] = A (TN)
    CUtensorMap b;
    CUtensorMap out, out_relu, out_f32, out_lo, out_relu_lo;
    CUtensorMap ga[4];             // one-hot selection patterns [npat * 128, 64] (K-major A tiles)
    CUtensorMap gb[4];             // gathered matrix [rows, N] (MN-major B tiles: 64 rows x 64 columns per box)
    CUtensorMap resid;             // residual [M, N] in 32-row x 64-column boxes (same geometry as the output maps)
};

__device__ __forceinline__ void add_bf16x8(float* f, const uint4& u) {
    f[0] += bf16_lo(u.x); f[1] += bf16_hi(u.x); f[2] += bf16_lo(u.y); f[3] += bf16_hi(u.y);
    f[4] += bf16_lo(u.z); f[5] += bf16_hi(u.z); f[6] += bf16_lo(u.w); f[7] += bf16_hi(u.w);
}
__device__ __forceinline__ void mask_bf16x8(float* f, const uint4& u) {
    // bf16 > 0  <=>  sign bit clear and magnitude non-zero
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t lo = w[i] & 0xFFFFu, hi = w[i] >> 16;
        if (!(lo != 0 && lo < 0x8000u)) f[2 * i] = 0.f;
        if (!(hi != 0 && hi < 0x8000u)) f[2 * i + 1] = 0.f;
    }
}

// Development build only (-DRPG_GEMM_TRACE, tools/gemm_trace.py): cycles each single-thread role of a CTA spends waiting,
// accumulated per CTA.  slots: 0 producer waits for a free ring stage, 1 producer total, 2 MMA issuer waits for operands,
// 3 MMA issuer waits for a drained accumulator, 4 MMA issuer total, 5 epilogue warp 0 waits for the accumulator,
// 6 epilogue warp 0 waits for its staging tile (previous TMA store still reading), 7 epilogue warp 0 total.
#ifdef RPG_GEMM_TRACE
__device__ unsigned long long g_gemm_trace[512][8];
#define TR_DECL(a) unsigned long long a = 0ull
#define TR_T0(v) const long long v = clock64()
#define TR_ACC(a, v) a += (unsigned long long)(clock64() - (v))
#define TR_FLUSH(slot, a) g_gemm_trace[blockIdx.x & 511][slot] += (a)
#else
#define TR_DECL(a)
#define TR_T0(v)
#define TR_ACC(a, v)
#define TR_FLUSH(slot, a)
#endif

// CL = CTAs per cluster.  With CL = 2 the two CTAs of a cluster work on adjacent 128-row blocks of the same column
// block and each fetches HALF of every weight (B) tile, multicast into both shared memories: the L2 -> SM traffic for
// B, which dominates at the small N, K of this path (every tile re-reads the whole weight panel), is halved.
// OPS = the epilogue reads tensor operands (gathered node rows / residual / bf16 mask); compiled separately so that the
// plain epilogue keeps its smaller register footprint and schedule.
// EPI = 2 additionally handles the fp32 ("split bf16") mode: fp32 gathered rows, (hi, lo) residuals and outputs.
// PAIR = the two CTAs of a cluster form ONE cta_group::2 MMA (M = 256): each CTA stages its own 128 rows of A and only
// HALF of the B tile (no copy of the peer's half), the leader issues the MMAs for both, and each CTA drains its own
// half of the accumulator.  Shared-memory fill + operand-read traffic per CTA drops by a third (48 -> 32 KB per
// k-block), which is what bounds the single-CTA form at these shapes, and the ring deepens from 4 to 6 stages.
// (NT mode without one-hot panels: the panels' node windows differ between the two row blocks.)
// WS = weight-stationary pair kernel (plain epilogue, K <= 512, no one-hot panels): the cluster's half-tiles of B for ALL
// k-blocks stay resident in shared memory (8 x 16 KB per CTA) and only A streams through a 4-stage ring.  The operand
// fill from L2 is what bounds the streaming form at these shapes (32 KB per k-block and CTA, ~8.4 TB/s chip-wide at
// the measured item time -- the L2 -> SM limit); with B resident it halves.  The work items of a cluster all share one
// column block (the host sizes the grid so that the cluster count is a multiple of the number of column blocks, and the
// item stride then keeps item % num_n_blocks constant).
template <int MODE, int CL, int EPI, bool PAIR = false, bool WS = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmTmaps tm, const GemmKParams p) {
    static_assert(!PAIR || (CL == 2 && MODE == 0), "CTA pairs: NT mode, clusters of 2");
    static_assert(!WS || (PAIR && EPI == 0), "weight-stationary: plain pair kernels");
    constexpr bool OPS = EPI == 1 || EPI == 2;
    constexpr bool RTMA = EPI == 3;        // the only tensor operand is a residual, fetched by TMA (pair kernels)
    static_assert(!RTMA || PAIR, "TMA-staged residuals are a pair-kernel feature");
    // EPI = 3 gives one ring stage (32 KB) to per-warp operand buffers: the residual tile of a chunk is fetched by TMA
    // (coalesced, asynchronous, one chunk ahead) instead of row-per-thread loads, and the operand-register arrays of
    // the general EPI = 1 epilogue (which spill) are not compiled in.
    constexpr bool OPBUF = RTMA;
    // STG_BUFS = 2 would give one ring stage of the plain pair kernel to a SECOND staging tile per epilogue warp, so
    // that a chunk's TMA store need not finish reading shared memory before the next chunk is written.  Measured on
    // B200: no gain (plain 155648 x 512 x 512 in the training step 87 -> 91 us with the 5-stage ring it costs) -- the
    // wait is not what paces an item -- so the single tile and the 6-stage ring stay.
    constexpr int STG_BUFS = 1;
    constexpr int NS = WS ? 4 : ((OPBUF || STG_BUFS == 2) ? 5 : (PAIR ? 6 : STAGES));  // ring depth
    constexpr int BST = PAIR ? B_STAGE_BYTES / 2 : B_STAGE_BYTES;          // B bytes per stage in this CTA
    constexpr int WS_KB = 8;                                               // resident k-blocks (K <= 512)
    static_assert(WS ? (NS * A_STAGE_BYTES + WS_KB * BST == SMEM_RING_BYTES)
                     : (NS * (A_STAGE_BYTES + BST) + ((OPBUF || STG_BUFS == 2) ? EPI_WARPS * EPI_STAGE_BYTES : 0) == SMEM_RING_BYTES),
                  "ring carve");
    const CUtensorMap& tmA0 = tm.a[0];
    const CUtensorMap& tmB = tm.b;
    const CUtensorMap& tmOut = tm.out;
    const CUtensorMap& tmOutRelu = tm.out_relu;
    const CUtensorMap& tmOutF32 = tm.out_f32;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + NS * A_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_RING_BYTES);
    uint64_t* full_bar = bars;                    // [NS]
    uint64_t* empty_bar = bars + NS;              // [NS]
    uint64_t* acc_full = bars + 2 * NS;           // [ACC_STAGES]
    uint64_t* acc_empty = acc_full + ACC_STAGES;  // [ACC_STAGES]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);
    uint64_t* op_bar = bars + 32;                 // [EPI_WARPS] operand-buffer barriers (OPBUF kernels)
    uint64_t* b_full = bars + 24;                 // WS: the resident B tiles have landed (leader's barrier)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0);
        tma_prefetch_desc(&tmB);
        // a stage is free when every CTA's MMAs have drained it (+ the local column-sum warps in TN mode)
        // (pair mode: one multicast commit per stage; the leader's accumulator barrier hears both CTAs' epilogue warps)
        const uint32_t empty_count = PAIR ? 1 : CL + ((MODE == 1 && p.a_colsum) ? EPI_WARPS : 0);
        for (int i = 0; i < NS; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], empty_count); }
        for (int i = 0; i < ACC_STAGES; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], PAIR ? 2 * EPI_WARPS : EPI_WARPS);
        }
        if (OPBUF) for (int i = 0; i < EPI_WARPS; ++i) mbar_init(&op_bar[i], 1);
        if (WS) mbar_init(b_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        if (PAIR) { tmem_alloc_pair(tmem_slot, TMEM_COLS); tmem_relinquish_pair(); }
        else { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
    }
    if (warp == 2 && lane == 0) {          // the store-side descriptors too: the first TMA store of a tile must not miss
        if (p.out) tma_prefetch_desc(&tmOut);
        if (p.out_relu) tma_prefetch_desc(&tmOutRelu);
        if (p.out_f32) tma_prefetch_desc(&tmOutF32);
        if (RTMA) tma_prefetch_desc(&tm.resid);
        if (MODE == 0 && p.n_gseg) { tma_prefetch_desc(&tm.ga[0]); tma_prefetch_desc(&tm.gb[0]); }
    }
    tc_fence_before();
    if (CL == 1) __syncthreads(); else cluster_sync_all();   // peer barriers must be initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // The set-up above overlapped the previous kernel's tail (programmatic dependent launch); its results are visible
    // only after the wait, and nothing before this line touched global memory.
    pdl_prologue();

    const int cta_rank = CL == 1 ? 0 : (int)cluster_ctarank();
    const int worker = blockIdx.x / CL, num_workers = gridDim.x / CL;     // a worker = one cluster
    const int num_m_blocks = (p.M + BLOCK_M - 1) / BLOCK_M;
    const int num_m_units = (num_m_blocks + CL - 1) / CL;                  // CL adjacent row blocks per work item
    const int num_n_blocks = (p.N + p.block_n - 1) / p.block_n;
    const int num_tiles = num_m_units * num_n_blocks;
    const int num_items = MODE == 0 ? num_tiles : num_tiles * p.splits;
    const uint32_t stage_tx_bytes = A_STAGE_BYTES + p.block_n * BLOCK_K * 2;
    constexpr uint16_t kAllCtas = (1u << CL) - 1;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------ TMA producer
            int stage = 0;
            uint32_t phase = 0;
            TR_T0(tr_prod);
            TR_DECL(tr_w0);
            TR_DECL(tr_tot);
            if (WS && worker < num_items) {
                // every item of this cluster has the same column block: its B half-tiles for all k-blocks, once
                const int n_blk = worker % num_n_blocks;
                const int half_rows = p.block_n / 2;
                const uint32_t lead_bfull = mapa_u32(smem_u32(b_full), 0);
                if (cta_rank == 0) mbar_arrive_expect_tx(b_full, 2u * (uint32_t)p.total_kb * (uint32_t)half_rows * BLOCK_K * 2);
                for (int kb = 0; kb < p.total_kb; ++kb)
                    tma_load_2d_pair(&tmB, lead_bfull, smem_b + kb * BST, kb * BLOCK_K, n_blk * p.block_n + cta_rank * half_rows);
            }
            for (int item = worker; item < num_items; item += num_workers) {
                const int tile = MODE == 0 ? item : item / p.splits;
                const int m_blk = (tile / num_n_blocks) * CL + cta_rank, n_blk = tile % num_n_blocks;
                if (MODE == 0) {
                    // L2 prefetch: the ring (64-96 KB of A per CTA) is shallower than the HBM latency under load (the MMA
                    // issuer waited for operands 12-24 % of its time, tools/gemm_trace.py); the next item's A tiles are
                    // pulled into L2 one item ahead, so its ring loads see the L2 latency instead.
                    const int item_nx = item + num_workers;
                    const bool pf = p.l2_prefetch && item_nx < num_items;
                    const int m_blk_nx = ((item_nx / num_n_blocks) * CL + cta_rank) * BLOCK_M;
                    int kb_global = 0;
                    for (int s = 0; s < 6; ++s) {
                        const CUtensorMap* tmA = &tm.a[s];
                        for (int kb = 0; kb < p.num_kb[s]; ++kb, ++kb_global) {
                            if (pf && m_blk_nx < p.M) tma_prefetch_2d(tmA, kb * BLOCK_K, m_blk_nx);
                            { TR_T0(t); mbar_wait(&empty_bar[stage], phase ^ 1); TR_ACC(tr_w0, t); }
                            if (WS) {
                                const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
                                if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * A_STAGE_BYTES);
                                tma_load_2d_pair(tmA, lead_full, smem_a + stage * A_STAGE_BYTES, kb * BLOCK_K, m_blk * BLOCK_M);
                                if (++stage == NS) { stage = 0; phase ^= 1; }
                                continue;
                            }
                            if (PAIR) {
                                // both CTAs' tiles complete on the LEADER's full barrier: it expects the bytes of both
                                const int half_rows = p.block_n / 2;
                                const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
                                if (cta_rank == 0)
                                    mbar_arrive_expect_tx(&full_bar[stage], 2 * (A_STAGE_BYTES + half_rows * BLOCK_K * 2));
                                tma_load_2d_pair(tmA, lead_full, smem_a + stage * A_STAGE_BYTES, kb * BLOCK_K, m_blk * BLOCK_M);
                                tma_load_2d_pair(&tmB, lead_full, smem_b + stage * BST, kb_global * BLOCK_K,
                                                 n_blk * p.block_n + cta_rank * half_rows);
                                if (++stage == NS) { stage = 0; phase ^= 1; }
                                continue;
                            }
                            mbar_arrive_expect_tx(&full_bar[stage], stage_tx_bytes);
                            tma_load_2d(tmA, &full_bar[stage], smem_a + stage * A_STAGE_BYTES, kb * BLOCK_K,
                                        m_blk * BLOCK_M);
                            if (CL == 1) {
                                tma_load_2d(&tmB, &full_bar[stage], smem_b + stage * BST, kb_global * BLOCK_K,
                                            n_blk * p.block_n);
                            } else {       // my half of the weight rows, delivered to both CTAs
                                const int half_rows = p.block_n / CL;
                                tma_load_2d_mc(&tmB, &full_bar[stage],
                                               smem_b + stage * BST + cta_rank * half_rows * (BLOCK_K * 2),
                                               kb_global * BLOCK_K, n_blk * p.block_n + cta_rank * half_rows, kAllCtas);
                            }
                            if (++stage == NS) { stage = 0; phase ^= 1; }
                        }
                    }
                    // one-hot K panels: selection pattern (A) + the node rows this row block can reference (B, MN-major).
                    // The two CTAs of a cluster own different row blocks => different node windows: no multicast here.
                    for (int gs = 0; gs < p.n_gseg; ++gs) {
                        const int r0 = m_blk * BLOCK_M;
                        const int pat = p.gsel_div ? (r0 % p.Ep) / p.gsel_div : m_blk;   // periodic template | per-block tiles
                        const int win = (r0 / p.Ep) * p.Nn;
                        { TR_T0(t); mbar_wait(&empty_bar[stage], phase ^ 1); TR_ACC(tr_w0, t); }
                        mbar_arrive_expect_tx(&full_bar[stage], stage_tx_bytes);
                        tma_load_2d(&tm.ga[gs], &full_bar[stage], smem_a + stage * A_STAGE_BYTES, 0, pat * BLOCK_M);
                        for (int j = 0; j < p.block_n / 64; ++j)
                            tma_load_2d(&tm.gb[gs], &full_bar[stage], smem_b + stage * BST + j * 8192,
                                        n_blk * p.block_n + j * 64, win);
                        if (++stage == NS) { stage = 0; phase ^= 1; }
                    }
                } else {
                    const int split = item % p.splits;
                    const int kb0 = split * p.kb_per_split;
                    const int kb1 = min(kb0 + p.kb_per_split, p.total_kb);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        { TR_T0(t); mbar_wait(&empty_bar[stage], phase ^ 1); TR_ACC(tr_w0, t); }
                        mbar_arrive_expect_tx(&full_bar[stage], stage_tx_bytes);
                        // MN-major tiles: one 64(MN) x 64(K rows) box per 128-byte slab of the MN extent
                        for (int j = 0; j < BLOCK_M / 64; ++j)
                            tma_load_2d(&tmA0, &full_bar[stage], smem_a + stage * A_STAGE_BYTES + j * 8192,
                                        m_blk * BLOCK_M + j * 64, kb * BLOCK_K);
                        for (int j = 0; j < p.block_n / 64; ++j) {
                            if (CL == 1)
                                tma_load_2d(&tmB, &full_bar[stage], smem_b + stage * BST + j * 8192,
                                            n_blk * p.block_n + j * 64, kb * BLOCK_K);
                            else if (j % CL == cta_rank)
                                tma_load_2d_mc(&tmB, &full_bar[stage], smem_b + stage * BST + j * 8192,
                                               n_blk * p.block_n + j * 64, kb * BLOCK_K, kAllCtas);
                        }
                        if (++stage == NS) { stage = 0; phase ^= 1; }
                    }
                }
            }
            TR_ACC(tr_tot, tr_prod);
            TR_FLUSH(0, tr_w0);
            TR_FLUSH(1, tr_tot);
        }
    } else if (warp == 1) {
        if (lane == 0 && !(PAIR && cta_rank != 0)) {
            // ------------------------------------------------------------ MMA issuer (single thread; pair: leader only)
            const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * BLOCK_M : BLOCK_M, p.block_n, MODE, MODE);
            // K-major SW128: 8-row groups 1024 B apart (SBO), LBO unused.
            // MN-major SW128: 64-element MN slabs 8192 B apart (LBO), 8-row K groups 1024 B apart (SBO).
            const uint32_t lbo = MODE == 0 ? 0u : 8192u, sbo = 1024u;
            const uint32_t k_step = MODE == 0 ? (UMMA_K * 2) >> 4 : (UMMA_K * 128) >> 4;   // desc address units (16 B)
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            TR_T0(tr_mma);
            TR_DECL(tr_w2);
            TR_DECL(tr_w3);
            TR_DECL(tr_tot);
            if (WS && worker < num_items) { mbar_wait(b_full, 0); tc_fence_after(); }
            for (int item = worker; item < num_items; item += num_workers, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                int n_kb;
                if (MODE == 0) {
                    n_kb = p.total_kb + p.n_gseg;
                } else {
                    const int kb0 = (item % p.splits) * p.kb_per_split;
                    n_kb = min(kb0 + p.kb_per_split, p.total_kb) - kb0;
                }
                { TR_T0(t); mbar_wait(&acc_empty[acc], acc_phase ^ 1); TR_ACC(tr_w3, t); }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
                for (int kb = 0; kb < n_kb; ++kb) {
                    { TR_T0(t); mbar_wait(&full_bar[stage], phase); TR_ACC(tr_w2, t); }
                    tc_fence_after();
                    const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + stage * A_STAGE_BYTES), lbo, sbo);
                    if (MODE == 0 && kb >= n_kb - p.n_gseg) {
                        // one-hot panel: A K-major (selection), B MN-major (node rows): only the B side changes layout
                        const uint64_t bg_desc = make_smem_desc_sw128(smem_u32(smem_b + stage * BST), 8192u, 1024u);
                        const uint32_t idesc_g = make_idesc_bf16(BLOCK_M, p.block_n, 0, 1);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            umma_bf16(d_tmem, a_desc + k * k_step, bg_desc + k * ((UMMA_K * 128) >> 4), idesc_g, (kb | k) != 0);
                    } else {
                        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + (WS ? kb : stage) * BST), lbo, sbo);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            if (PAIR) umma_bf16_pair(d_tmem, a_desc + k * k_step, b_desc + k * k_step, idesc, (kb | k) != 0);
                            else umma_bf16(d_tmem, a_desc + k * k_step, b_desc + k * k_step, idesc, (kb | k) != 0);
                        }
                    }
                    if (PAIR) umma_commit_pair(&empty_bar[stage], kAllCtas);   // both CTAs' slots, both producers
                    else if (CL == 1) umma_commit(&empty_bar[stage]);   // frees the smem slot when these MMAs retire
                    else umma_commit_mc(&empty_bar[stage], kAllCtas);   // ... in every CTA that multicasts into it
                    if (kb == n_kb - 1) {
                        if (PAIR) umma_commit_pair(&acc_full[acc], kAllCtas);   // both epilogues
                        else umma_commit(&acc_full[acc]);
                    }
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
                if (n_kb <= 0) umma_commit(&acc_full[acc]);   // empty split: nothing to accumulate (epilogue writes zeros)
            }
            TR_ACC(tr_tot, tr_mma);
            TR_FLUSH(2, tr_w2);
            TR_FLUSH(3, tr_w3);
            TR_FLUSH(4, tr_tot);
        }
    } else {
        // ---------------------------------------------------------------- epilogue warps
        // Each thread owns one accumulator row (TMEM lane).  Row-per-thread global accesses would touch 32 different
        // lines per instruction (measured: 37 % of all stall samples behind those stores), so everything goes through a
        // per-warp staging tile [32 rows][128 B] in the TMA 128-byte-swizzle layout (16-byte chunk c of row r lives at
        // chunk c ^ (r & 7): conflict-free for the row view and for the coalesced view):
        //   results  : registers -> staging -> ONE TMA store per 32 x 64 tile (asynchronous, clipped at the tensor edge)
        //   operands : ReLU patterns arrive as 8 bytes of bits per row and chunk; tensor operands (gathered node rows,
        //              residuals) are read row-per-thread, issued ahead of the TMEM read
        const int ew = warp - 2;
        const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
        const int half = ew >> 2;                  // the two warps of a quadrant alternate 64-column chunks
        const int n_chunks = (p.block_n + 63) / 64;
        uint8_t* const stg0 = smem + SMEM_RING_BYTES + SMEM_BAR_BYTES + ew * EPI_STAGE_BYTES;
        uint8_t* const stg1 = smem + NS * (A_STAGE_BYTES + BST) + ew * EPI_STAGE_BYTES;   // only with STG_BUFS == 2
        uint8_t* stg = stg0;
        uint8_t* my_row = stg + lane * 128;
        int sbuf = 0;
        TR_DECL(tr_w5);
        TR_DECL(tr_w6);
        TR_DECL(tr_etot);
        // a staging tile that no earlier TMA store of this warp still reads (stores alternate between the tiles)
        auto acquire_stage = [&]() {
            if (lane == 0) {
                TR_T0(t);
                if (STG_BUFS == 2) bulk_wait_read_1(); else bulk_wait_read_all();
                TR_ACC(tr_w6, t);
            }
            __syncwarp();
            if (STG_BUFS == 2) {
                sbuf ^= 1;
                stg = sbuf ? stg1 : stg0;
                my_row = stg + lane * 128;
            }
        };
        // operand buffer of this warp (OPBUF kernels): the ring stage that was given up, [32 rows][128 B] swizzled
        uint8_t* opbuf = smem + NS * (A_STAGE_BYTES + BST) + ew * EPI_STAGE_BYTES;
        constexpr bool rtma = RTMA;
        uint32_t op_phase = 0;
        const int my_sw = lane & 7;
        // The residual tiles are prefetched one chunk ahead, across work items: next valid (item, chunk) of this warp in
        // processing order, with the coordinates of its 32 x 64 box.
        auto next_resid = [&](int it_, int c_, int& r0_, int& n0_) -> bool {
            for (;;) {
                if (c_ >= n_chunks) { it_ += num_workers; c_ = half; }
                if (it_ >= num_items) return false;
                const int mb = (it_ / num_n_blocks) * CL + cta_rank, nb = it_ % num_n_blocks;
                r0_ = mb * BLOCK_M + quad * 32;
                n0_ = nb * p.block_n + c_ * 64;
                if (r0_ < p.M && n0_ < p.N) return true;
                c_ += 2;
            }
        };
        if (rtma && lane == 0) {
            int pr0, pn0;
            if (next_resid(worker, half, pr0, pn0)) {
                mbar_arrive_expect_tx(&op_bar[ew], EPI_STAGE_BYTES);
                tma_load_2d(&tm.resid, &op_bar[ew], opbuf, pn0, pr0);
            }
        }
        int it = 0;
        TR_T0(tr_epi);
        int cs_stage = 0;                          // TN column sums: this role walks the smem ring like the MMA warp
        uint32_t cs_phase = 0;
        for (int item = worker; item < num_items; item += num_workers, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int tile = MODE == 0 ? item : item / p.splits;
            const int m_blk = (tile / num_n_blocks) * CL + cta_rank, n_blk = tile % num_n_blocks;
            if (MODE == 1 && p.a_colsum) {
                // While the tensor core works through this item's k-blocks the epilogue warps have nothing to do: they
                // sum the A tile (dY^T, 64 rows x 128 columns, MN-major 128B-swizzled) over its rows.  Thread = (16-byte
                // column chunk, row group); 4 LDS.128 + 32 FADD per stage.  Fixed order => deterministic.
                const int kb0 = (item % p.splits) * p.kb_per_split;
                const int n_kb = min(kb0 + p.kb_per_split, p.total_kb) - kb0;
                const int et = threadIdx.x - 64;                 // 0..255 over the epilogue warps
                const int c16 = et & 15, rg = et >> 4;
                float cs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&full_bar[cs_stage], cs_phase);
                    if (n_blk == 0) {
                        const uint8_t* a_tile = smem_a + cs_stage * A_STAGE_BYTES + (c16 >> 3) * 8192;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int r = rg + 16 * j;
                            const uint4 u = *reinterpret_cast<const uint4*>(a_tile + r * 128 + (((c16 & 7) ^ (r & 7)) << 4));
                            cs[0] += bf16_lo(u.x); cs[1] += bf16_hi(u.x); cs[2] += bf16_lo(u.y); cs[3] += bf16_hi(u.y);
                            cs[4] += bf16_lo(u.z); cs[5] += bf16_hi(u.z); cs[6] += bf16_lo(u.w); cs[7] += bf16_hi(u.w);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[cs_stage]);
                    if (++cs_stage == NS) { cs_stage = 0; cs_phase ^= 1; }
                }
                if (n_blk == 0) {
                    // fold the 16 row groups through the (idle) staging area, then one thread per column writes
                    float* red = reinterpret_cast<float*>(smem + SMEM_RING_BYTES + SMEM_BAR_BYTES);   // [16][128]
                    if (lane == 0) bulk_wait_read_all();                         // earlier TMA stores have left the staging area
                    __syncwarp();
                    asm volatile("bar.sync 1, 256;" ::: "memory");               // previous item's readers are done
#pragma unroll
                    for (int q = 0; q < 8; ++q) red[rg * 128 + c16 * 8 + q] = cs[q];
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (et < 128) {
                        float t = 0.f;
#pragma unroll
                        for (int g2 = 0; g2 < 16; ++g2) t += red[g2 * 128 + et];
                        const int col = m_blk * BLOCK_M + et;
                        if (col < p.M) p.a_colsum[(size_t)(item % p.splits) * p.M + col] = t;
                    }
                }
            }
            const int row0 = m_blk * BLOCK_M + quad * 32;
            const int row = row0 + lane;
            const bool row_ok = row < p.M;
            bool zero_acc = false;
            if (MODE == 1) {
                const int kb0 = (item % p.splits) * p.kb_per_split;
                zero_acc = kb0 >= p.total_kb;
            }
            int gnode0 = -1, gnode1 = -1;
            float rscale = 1.f;
            if (OPS && MODE == 0 && row_ok) {
                if (p.gadd[0] || p.gadd[1] || (EPI == 2 && (p.gadd_f32[0] || p.gadd_f32[1]))) {
                    const int gidx = row / p.Ep, k = row - gidx * p.Ep;
                    if (p.gmap[0]) gnode0 = gidx * p.Nn + __ldg(p.gmap[0] + k);
                    if (p.gmap[1]) gnode1 = gidx * p.Nn + __ldg(p.gmap[1] + k);
                }
                if (p.row_scale) rscale = __ldg(p.row_scale + (row % p.row_scale_mod));
            }
            if (!OPS && MODE == 0 && row_ok && p.row_scale) rscale = __ldg(p.row_scale + (row % p.row_scale_mod));
            // ---- everything that does not need the accumulator is requested BEFORE waiting for it: the ReLU bit
            // patterns of both chunks (8 bytes each) and the first two operand tiles of the first chunk.
            unsigned long long mbits0 = ~0ull, mbits1 = ~0ull;
            if (MODE == 0 && p.mask_bits && row_ok) {
                const int n0a = n_blk * p.block_n + half * 64, n0b = n0a + 128;
                if (half < n_chunks && n0a < p.N)
                    mbits0 = __ldg(reinterpret_cast<const unsigned long long*>(p.mask_bits + (size_t)row * p.mask_bits_ld + (n0a >> 3)));
                if (half + 2 < n_chunks && n0b < p.N)
                    mbits1 = __ldg(reinterpret_cast<const unsigned long long*>(p.mask_bits + (size_t)row * p.mask_bits_ld + (n0b >> 3)));
            }
            { TR_T0(t); mbar_wait(&acc_full[acc], acc_phase); TR_ACC(tr_w5, t); }
            tc_fence_after();
            bool released = false;

            for (int c = half, ci = 0; c < n_chunks; c += 2, ++ci) {
                const int n0 = n_blk * p.block_n + c * 64;
                const int ncols = min(64, min(p.N - n0, p.block_n - c * 64));   // multiple of 8 (host-checked)
                if (ncols <= 0 || row0 >= p.M) break;
                const int nq = ncols >> 3;                                       // valid 16-byte bf16 chunks per row

                // ---- epilogue operands: each thread fetches its own row (64 bf16 = 8 x 16 B) directly, issued before the
                // TMEM read so the latencies overlap.  (Staging them through shared memory for coalescing was measured
                // slower: the extra warp-synchronous round trip outweighs the saved L1 tag cycles.)
                uint4 opr0[8], opr1[8];
                const __nv_bfloat16* src0 = nullptr;
                const __nv_bfloat16* src1 = nullptr;
                int kind0 = 0, kind1 = 0;                                        // 1 = add, 2 = bf16 ReLU mask
                if (OPS && MODE == 0 && row_ok) {
                    const __nv_bfloat16* cand[4] = {
                        (gnode0 >= 0 && p.gadd[0]) ? p.gadd[0] + (size_t)gnode0 * p.gadd_ld[0] : nullptr,
                        (gnode1 >= 0 && p.gadd[1]) ? p.gadd[1] + (size_t)gnode1 * p.gadd_ld[1] : nullptr,
                        (p.resid && !rtma) ? p.resid + (size_t)row * p.resid_ld : nullptr,
                        p.mask ? p.mask + (size_t)row * p.mask_ld : nullptr};
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        if (!cand[o]) continue;
                        if (!src0) { src0 = cand[o] + n0; kind0 = o == 3 ? 2 : 1; }
                        else if (!src1) { src1 = cand[o] + n0; kind1 = o == 3 ? 2 : 1; }
                    }
                    if (src0) {
#pragma unroll
                        for (int qq = 0; qq < 8; ++qq)
                            opr0[qq] = qq < nq ? __ldg(reinterpret_cast<const uint4*>(src0) + qq) : make_uint4(0, 0, 0, 0);
                    }
                    if (src1) {
#pragma unroll
                        for (int qq = 0; qq < 8; ++qq)
                            opr1[qq] = qq < nq ? __ldg(reinterpret_cast<const uint4*>(src1) + qq) : make_uint4(0, 0, 0, 0);
                    }
                }

                // ---- accumulator -> registers
                float f[64];
                const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * ACC_STRIDE + c * 64;
                tmem_ld_32x32(taddr, reinterpret_cast<uint32_t(&)[32]>(f[0]));
                tmem_ld_32x32(taddr + 32, reinterpret_cast<uint32_t(&)[32]>(f[32]));
                tmem_ld_wait();
                if (c + 2 >= n_chunks) {          // last TMEM read of this tile by this warp: hand the stage back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (PAIR && cta_rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[acc]), 0));
                        else mbar_arrive(&acc_empty[acc]);
                    }
                    released = true;
                }
                if (zero_acc) {
#pragma unroll
                    for (int j = 0; j < 64; ++j) f[j] = 0.f;
                }
                uint4 rsd[8];
                if (rtma) {
                    mbar_wait(&op_bar[ew], op_phase);
                    op_phase ^= 1;
#pragma unroll
                    for (int qq = 0; qq < 8; ++qq)
                        rsd[qq] = *reinterpret_cast<const uint4*>(opbuf + lane * 128 + ((qq ^ my_sw) << 4));
                    __syncwarp();
                    if (lane == 0) {                         // the next chunk's tile (possibly of the next work item)
                        int pr0, pn0;                        // overlaps this chunk's math and store
                        if (next_resid(item, c + 2, pr0, pn0)) {
                            fence_proxy_async_smem();
                            mbar_arrive_expect_tx(&op_bar[ew], EPI_STAGE_BYTES);
                            tma_load_2d(&tm.resid, &op_bar[ew], opbuf, pn0, pr0);
                        }
                    }
                }
                acquire_stage();

                if (MODE == 0) {
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 64; j += 4) {
                            if (j < ncols) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                                f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
                            }
                        }
                    }
                    // order: gathered adds, residual, row scale, ReLU mask (at most two tensor operands per GEMM here;
                    // a third/fourth one is fetched late)
                    int used = 0;
                    bool scaled = p.row_scale == nullptr;
                    auto apply = [&](const uint4 (&o)[8], int kind) {
                        if (kind == 2 && !scaled) {
#pragma unroll
                            for (int j = 0; j < 64; ++j) f[j] *= rscale;
                            scaled = true;
                        }
#pragma unroll
                        for (int qq = 0; qq < 8; ++qq) {
                            if (kind == 2) mask_bf16x8(f + 8 * qq, o[qq]); else add_bf16x8(f + 8 * qq, o[qq]);
                        }
                    };
                    if (OPS && src0) { apply(opr0, kind0); ++used; }
                    if (OPS && src1) { apply(opr1, kind1); ++used; }
                    if (rtma) {                                      // host guarantees: no other tensor operand in this launch
#pragma unroll
                        for (int qq = 0; qq < 8; ++qq) add_bf16x8(f + 8 * qq, rsd[qq]);
                    }
                    if (OPS && row_ok) {                             // rare: more than two tensor operands
                        const __nv_bfloat16* cand[4] = {
                            (gnode0 >= 0 && p.gadd[0]) ? p.gadd[0] + (size_t)gnode0 * p.gadd_ld[0] : nullptr,
                            (gnode1 >= 0 && p.gadd[1]) ? p.gadd[1] + (size_t)gnode1 * p.gadd_ld[1] : nullptr,
                            (p.resid && !rtma) ? p.resid + (size_t)row * p.resid_ld : nullptr,
                            p.mask ? p.mask + (size_t)row * p.mask_ld : nullptr};
                        int seen = 0;
                        for (int o = 0; o < 4; ++o) {
                            if (!cand[o]) continue;
                            if (seen++ < 2) continue;
                            uint4 late[8];
#pragma unroll
                            for (int qq = 0; qq < 8; ++qq)
                                late[qq] = qq < nq ? __ldg(reinterpret_cast<const uint4*>(cand[o] + n0) + qq) : make_uint4(0, 0, 0, 0);
                            apply(late, o == 3 ? 2 : 1);
                        }
                    }
                    if (EPI == 2 && row_ok) {            // fp32 mode: fp32 gathered node rows, low plane of the residual
#pragma unroll
                        for (int o = 0; o < 2; ++o) {
                            const int gn = o == 0 ? gnode0 : gnode1;
                            if (p.gadd_f32[o] && gn >= 0) {
                                const float4* src = reinterpret_cast<const float4*>(p.gadd_f32[o] + (size_t)gn * p.gadd_f32_ld[o] + n0);
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    if (4 * j < ncols) {
                                        const float4 v = __ldg(src + j);
                                        f[4 * j] += v.x; f[4 * j + 1] += v.y; f[4 * j + 2] += v.z; f[4 * j + 3] += v.w;
                                    }
                                }
                            }
                        }
                        if (p.resid_lo) {
                            const uint4* src = reinterpret_cast<const uint4*>(p.resid_lo + (size_t)row * p.resid_ld + n0);
#pragma unroll
                            for (int qq = 0; qq < 8; ++qq) if (qq < nq) add_bf16x8(f + 8 * qq, __ldg(src + qq));
                        }
                    }
                    if (!scaled) {
#pragma unroll
                        for (int j = 0; j < 64; ++j) f[j] *= rscale;
                    }
                    if (p.mask_bits) {
                        const unsigned long long mb = ci == 0 ? mbits0 : mbits1;
#pragma unroll
                        for (int j = 0; j < 64; ++j) if (!((mb >> j) & 1ull)) f[j] = 0.f;
                    }
                    if (p.relu) {
#pragma unroll
                        for (int j = 0; j < 64; ++j) f[j] = fmaxf(f[j], 0.f);
                    }
                    // fused feature dropout (last layer of the stack): keep pattern of this row's 64 columns, the same
                    // counter-based decision the head kernels and rpg_dropout_mask make
                    unsigned long long keepm = ~0ull;
                    if (EPI == 0 && p.drop_thresh) {
                        keepm = 0ull;
#pragma unroll
                        for (int q4 = 0; q4 < 16; ++q4) {
                            const uint32_t h = keep_hash4(p.drop_seed, row, (n0 >> 2) + q4);
#pragma unroll
                            for (int b4 = 0; b4 < 4; ++b4)
                                keepm |= (unsigned long long)(((h >> (8 * b4)) & 0xFFu) >= p.drop_thresh) << (4 * q4 + b4);
                        }
                    }
                    if (p.out_bits && row_ok) {
                        unsigned long long ob = 0ull;
#pragma unroll
                        for (int j = 0; j < 64; ++j) ob |= (unsigned long long)(f[j] > 0.f) << j;
                        ob &= keepm;
                        *reinterpret_cast<unsigned long long*>(p.out_bits + (size_t)row * p.out_bits_ld + (n0 >> 3)) = ob;
                    }
                    if (p.out) {
#pragma unroll
                        for (int qq = 0; qq < 8; ++qq) {
                            uint4 u;
                            u.x = pack_bf16x2(f[8 * qq], f[8 * qq + 1]); u.y = pack_bf16x2(f[8 * qq + 2], f[8 * qq + 3]);
                            u.z = pack_bf16x2(f[8 * qq + 4], f[8 * qq + 5]); u.w = pack_bf16x2(f[8 * qq + 6], f[8 * qq + 7]);
                            *reinterpret_cast<uint4*>(my_row + ((qq ^ my_sw) << 4)) = u;
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) { tma_store_2d(&tmOut, stg, n0, row0); bulk_commit(); }
                    }
                    if (p.out_relu) {
                        if (p.out) acquire_stage();
                        if (EPI == 0 && p.drop_thresh) {
#pragma unroll
                            for (int j = 0; j < 64; ++j) f[j] = ((keepm >> j) & 1ull) ? fmaxf(f[j], 0.f) * p.drop_scale : 0.f;
                        }
#pragma unroll
                        for (int qq = 0; qq < 8; ++qq) {
                            uint4 u;
                            u.x = pack_bf16x2(fmaxf(f[8 * qq], 0.f), fmaxf(f[8 * qq + 1], 0.f));
                            u.y = pack_bf16x2(fmaxf(f[8 * qq + 2], 0.f), fmaxf(f[8 * qq + 3], 0.f));
                            u.z = pack_bf16x2(fmaxf(f[8 * qq + 4], 0.f), fmaxf(f[8 * qq + 5], 0.f));
                            u.w = pack_bf16x2(fmaxf(f[8 * qq + 6], 0.f), fmaxf(f[8 * qq + 7], 0.f));
                            *reinterpret_cast<uint4*>(my_row + ((qq ^ my_sw) << 4)) = u;
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) { tma_store_2d(&tmOutRelu, stg, n0, row0); bulk_commit(); }
                    }
                    if (EPI == 2) {                      // low planes: lo = bf16(v - float(bf16(v)))
#pragma unroll
                        for (int pass = 0; pass < 2; ++pass) {
                            if (pass == 0 ? p.out_lo == nullptr : p.out_relu_lo == nullptr) continue;
                            acquire_stage();
#pragma unroll
                            for (int qq = 0; qq < 8; ++qq) {
                                float r[8];
#pragma unroll
                                for (int t = 0; t < 8; ++t) {
                                    const float v = pass == 0 ? f[8 * qq + t] : fmaxf(f[8 * qq + t], 0.f);
                                    r[t] = v - __bfloat162float(__float2bfloat16_rn(v));
                                }
                                uint4 u;
                                u.x = pack_bf16x2(r[0], r[1]); u.y = pack_bf16x2(r[2], r[3]);
                                u.z = pack_bf16x2(r[4], r[5]); u.w = pack_bf16x2(r[6], r[7]);
                                *reinterpret_cast<uint4*>(my_row + ((qq ^ my_sw) << 4)) = u;
                            }
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) { tma_store_2d(pass == 0 ? &tm.out_lo : &tm.out_relu_lo, stg, n0, row0); bulk_commit(); }
                        }
                    }
                }
                if (p.out_f32) {
                    const int split = MODE == 1 ? item % p.splits : 0;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {               // 32 fp32 columns = 128 B per row per pass
                        if (h * 32 >= ncols) break;
                        if (h == 1 || (MODE == 0 && (p.out || p.out_relu))) acquire_stage();
#pragma unroll
                        for (int qq = 0; qq < 8; ++qq)
                            *reinterpret_cast<float4*>(my_row + ((qq ^ my_sw) << 4)) =
                                make_float4(f[h * 32 + 4 * qq], f[h * 32 + 4 * qq + 1], f[h * 32 + 4 * qq + 2], f[h * 32 + 4 * qq + 3]);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) { tma_store_3d(&tmOutF32, stg, n0 + h * 32, row0, split); bulk_commit(); }
                    }
                }
            }
            if (!released) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                        if (PAIR && cta_rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[acc]), 0));
                        else mbar_arrive(&acc_empty[acc]);
                    }
            }
        }
        if (lane == 0) bulk_wait_all();            // all results are in global memory before the CTA retires
        TR_ACC(tr_etot, tr_epi);
        if (ew == 0 && lane == 0) { TR_FLUSH(5, tr_w5); TR_FLUSH(6, tr_w6); TR_FLUSH(7, tr_etot); }
    }

    __syncwarp();
    tc_fence_before();
    if (CL == 1) __syncthreads(); else cluster_sync_all();   // no CTA may exit while its peer can still signal it
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
        else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// -------------------------------------------------------------------------------------- host side

// RPG_GEMM_WS=0 keeps the plain pair GEMMs on the streaming-B form (A/B comparisons).
static bool gemm_ws_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("RPG_GEMM_WS");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}
// RPG_GEMM_L2PF=0 turns the producer's L2 prefetch of the next work item off (A/B comparisons).
static bool gemm_l2_prefetch_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("RPG_GEMM_L2PF");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}
// RPG_GEMM_PAIR=0 keeps every GEMM on the single-CTA MMA form (A/B comparisons).
static bool gemm_pair_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("RPG_GEMM_PAIR");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

// (Round 2 measured a per-thread memo table for the descriptors -- cuTensorMapEncodeTiled is a pure function of its
// arguments and a step asks for the same ~400 every time: no difference in the host time of a step, 1.38-1.58 ms either
// way; the driver encodes one in ~0.2 us.  Not kept.)
// rank 2 or 3; dims / strides (bytes, rank - 1 of them) / box in the driver's order (innermost first); bf16 or fp32
// elements; 128-byte swizzle, 256-byte L2 promotion, no OOB fill value (zeros).
int tmap_encode(::CUtensorMap_st* tm, int f32, int rank, const void* base, const uint64_t* dims, const uint64_t* strides,
                const uint32_t* box) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_error(RPG_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
    if (rank < 2 || rank > 3) return set_error(RPG_E_ARG, "tmap_encode: rank must be 2 or 3");
    struct { uint64_t dims[3]; uint64_t strides[2]; uint32_t box[3]; } key;
    memset(&key, 0, sizeof key);
    for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) key.strides[i] = strides[i];
    cuuint64_t d[3] = {key.dims[0], key.dims[1], key.dims[2]};
    cuuint64_t st[2] = {key.strides[0], key.strides[1]};
    cuuint32_t bx[3] = {key.box[0], key.box[1], key.box[2]};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank,
                    const_cast<void*>(base), d, st, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[200];
        snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed (%d): rank=%d dims=%llu,%llu,%llu pitch=%llu box=%u,%u,%u %s", (int)r, rank,
                 (unsigned long long)key.dims[0], (unsigned long long)key.dims[1], (unsigned long long)key.dims[2],
                 (unsigned long long)key.strides[0], key.box[0], key.box[1], key.box[2], f32 ? "f32" : "bf16");
        return set_error((int)r, msg);
    }
    return 0;
}

// 2-D bf16 tensor map: inner extent `inner` (contiguous), outer extent `outer`, row pitch `ld` elements.
static int make_tmap(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                     uint32_t box_inner, uint32_t box_outer) {
    const uint64_t dims[2] = {inner, outer}, strides[1] = {ld * 2};
    const uint32_t box[2] = {box_inner, box_outer};
    return tmap_encode(tm, 0, 2, base, dims, strides, box);
}

// 3-D fp32 tensor map [planes][rows][cols] for the epilogue's fp32 results (split-R partials are the planes).
static int make_tmap_f32_3d(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows, uint64_t planes, uint64_t ld,
                            uint64_t plane_stride) {
    const uint64_t dims[3] = {cols, rows, planes}, strides[2] = {ld * 4, (planes > 1 ? plane_stride : ld * rows) * 4};
    const uint32_t box[3] = {32, 32, 1};
    return tmap_encode(tm, 1, 3, base, dims, strides, box);
}

// ---- per-launch event timing (enabled only between rpg_profile_begin and rpg_profile_end / rpg_profile_records)
struct ProfRec { cudaEvent_t e0, e1; int cls; double flops, bytes; int M, N, K, flags; };
static std::mutex g_prof_mu;
static std::atomic<bool> g_prof_on{false};
static std::vector<ProfRec> g_prof;

bool prof_active() { return g_prof_on.load(std::memory_order_relaxed); }

int profile_begin() {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_prof.clear();
    g_prof_on = true;
    return 0;
}

int prof_open(int cls, double flops, double bytes, int M, int N, int K, cudaStream_t s) {
    if (!prof_active()) return -1;
    ProfRec rec;
    cudaEventCreate(&rec.e0); cudaEventCreate(&rec.e1);
    rec.cls = cls; rec.flops = flops; rec.bytes = bytes; rec.M = M; rec.N = N; rec.K = K; rec.flags = 0;
    cudaEventRecord(rec.e0, s);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(rec);
    return (int)g_prof.size() - 1;
}

void prof_close(int slot, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (slot >= 0 && slot < (int)g_prof.size()) cudaEventRecord(g_prof[slot].e1, s);
}

// Drains the window: every record with its measured duration.
int profile_records(rpg_prof_rec_t* out, int max_records, int* n_records) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = false;
    int n = 0;
    for (auto& r : g_prof) {
        cudaEventSynchronize(r.e1);
        float t = 0.f;
        cudaEventElapsedTime(&t, r.e0, r.e1);
        if (out && n < max_records) {
            out[n].cls = r.cls; out[n].M = r.M; out[n].N = r.N; out[n].K = r.K;
            out[n].ms = t; out[n].flops = r.flops; out[n].bytes = r.bytes;
        }
        ++n;
        if (getenv("RPG_PROFILE_VERBOSE"))
            fprintf(stderr, "[rpg prof] cls=%d M=%d N=%d K=%d  %.1f us  %.0f TFLOP/s  %.0f GB/s\n", r.cls, r.M, r.N, r.K, t * 1e3,
                    r.flops / (t * 1e-3) / 1e12, r.bytes / (t * 1e-3) / 1e9);
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    g_prof.clear();
    if (n_records) *n_records = n;
    return 0;
}

int profile_end(double* nt_ms, double* tn_ms, int* nt_launches, int* tn_launches, double* nt_flops, double* tn_flops) {
    std::vector<rpg_prof_rec_t> recs;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        recs.resize(g_prof.size());
    }
    int n = 0;
    profile_records(recs.data(), (int)recs.size(), &n);
    double ms[2] = {0, 0}, fl[2] = {0, 0};
    int cnt[2] = {0, 0};
    for (int i = 0; i < n && i < (int)recs.size(); ++i) {
        if (recs[i].cls > RPG_PROF_GEMM_TN) continue;
        ms[recs[i].cls] += recs[i].ms; fl[recs[i].cls] += recs[i].flops; cnt[recs[i].cls]++;
    }
    if (nt_ms) *nt_ms = ms[0];
    if (tn_ms) *tn_ms = ms[1];
    if (nt_launches) *nt_launches = cnt[0];
    if (tn_launches) *tn_launches = cnt[1];
    if (nt_flops) *nt_flops = fl[0];
    if (tn_flops) *tn_flops = fl[1];
    return 0;
}

static thread_local int g_sm_count = 0;
static std::mutex g_attr_mu;
static bool g_attr_done[64] = {false};
static int g_sm_counts[64] = {0};
static int g_cluster = 2;       // CTAs per cluster (1 or 2); see rpg_set_gemm_cluster

int set_gemm_cluster(int cl) {
    if (cl != 1 && cl != 2) return set_error(RPG_E_ARG, "gemm cluster size must be 1 or 2");
    g_cluster = cl;
    return 0;
}

static int aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static void set_smem_attrs() {
    cudaFuncSetAttribute(gemm_tc_kernel<0, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(gemm_tc_kernel<0, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(gemm_tc_kernel<0, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(gemm_tc_kernel<1, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(gemm_tc_kernel<0, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(gemm_tc_kernel<0, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(gemm_tc_kernel<0, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(gemm_tc_kernel<1, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(gemm_tc_kernel<0, 2, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(gemm_tc_kernel<0, 2, 0, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(gemm_tc_kernel<0, 2, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(gemm_tc_kernel<0, 2, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(gemm_tc_kernel<0, 2, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
}

int gemm_launch(const rpg_gemm_t* g, cudaStream_t stream) {
    if (!g) return set_error(RPG_E_ARG, "rpg_gemm: null descriptor");
    if (g->M <= 0 || g->N <= 0) return set_error(RPG_E_ARG, "rpg_gemm: empty output");
    if (g->N % 8) return set_error(RPG_E_ARG, "rpg_gemm: N must be a multiple of 8");
    int block_n = g->block_n;
    if (block_n == 0) {
        block_n = g->N >= 256 ? 256 : ((g->N + 15) / 16) * 16;
        if (g->mode == 1) block_n = g->N >= 256 ? 256 : ((g->N + 63) / 64) * 64;
    }
    if (block_n < 16 || block_n > 256 || block_n % 16) return set_error(RPG_E_ARG, "rpg_gemm: bad block_n");
    if (g->mode == 1 && block_n % 64) return set_error(RPG_E_ARG, "rpg_gemm: TN mode needs block_n % 64 == 0");
    if (!g->B || !aligned16(g->B) || g->ldb % 8) return set_error(RPG_E_ARG, "rpg_gemm: B pointer/pitch alignment");

    const int cl = g_cluster;
    GemmKParams p;
    memset(&p, 0, sizeof p);
    p.mode = g->mode; p.M = g->M; p.N = g->N; p.block_n = block_n;
    static_assert(sizeof(GemmTmaps) + sizeof(GemmKParams) < 4000, "kernel parameter space");
    GemmTmaps tmaps;
    memset(&tmaps, 0, sizeof tmaps);
    CUtensorMap* tmA = tmaps.a;
    CUtensorMap& tmB = tmaps.b;
    int rc;
    if (g->mode == 0) {
        if (g->n_seg < 1 || g->n_seg > 6) return set_error(RPG_E_ARG, "rpg_gemm: n_seg must be 1..6");
        int ktot = 0;
        for (int s = 0; s < g->n_seg; ++s) {
            if (!g->A[s] || !aligned16(g->A[s]) || g->lda[s] % 8 || g->K[s] <= 0 || g->K[s] % BLOCK_K)
                return set_error(RPG_E_ARG, "rpg_gemm: A segment pointer/pitch/K (K must be a multiple of 64)");
            p.num_kb[s] = g->K[s] / BLOCK_K;
            ktot += g->K[s];
            if ((rc = make_tmap(&tmA[s], g->A[s], g->K[s], g->M, g->lda[s], BLOCK_K, BLOCK_M))) return rc;
        }
        for (int s = g->n_seg; s < 6; ++s) tmA[s] = tmA[0];
        p.total_kb = ktot / BLOCK_K;
        if ((rc = make_tmap(&tmB, g->B, ktot, g->N, g->ldb, BLOCK_K, block_n / cl))) return rc;
        p.splits = 1;
        if ((g->gadd[0] && !g->gmap[0]) || (g->gadd[1] && !g->gmap[1]) || ((g->gadd[0] || g->gadd[1]) && (g->Ep <= 0 || g->Nn <= 0)))
            return set_error(RPG_E_ARG, "rpg_gemm: gathered add needs gmap, Ep, Nn");
        if (!g->out && !g->out_relu && !g->out_f32 && !g->out_bits) return set_error(RPG_E_ARG, "rpg_gemm: no output");
    } else if (g->mode == 1) {
        if (!g->A[0] || !aligned16(g->A[0]) || g->lda[0] % 8 || g->R <= 0)
            return set_error(RPG_E_ARG, "rpg_gemm: TN operand A");
        if (!g->out_f32 || g->splits < 1) return set_error(RPG_E_ARG, "rpg_gemm: TN mode writes fp32 partials");
        p.total_kb = (g->R + BLOCK_K - 1) / BLOCK_K;
        p.splits = g->splits;
        p.kb_per_split = (p.total_kb + g->splits - 1) / g->splits;
        p.split_stride = g->split_stride;
        if ((rc = make_tmap(&tmA[0], g->A[0], g->M, g->R, g->lda[0], 64, BLOCK_K))) return rc;
        for (int s = 1; s < 6; ++s) tmA[s] = tmA[0];
        if ((rc = make_tmap(&tmB, g->B, g->N, g->R, g->ldb, 64, BLOCK_K))) return rc;
    } else {
        return set_error(RPG_E_ARG, "rpg_gemm: mode must be 0 (NT) or 1 (TN)");
    }
    p.bias = g->bias;
    for (int i = 0; i < 2; ++i) {
        p.gadd[i] = reinterpret_cast<const __nv_bfloat16*>(g->gadd[i]);
        p.gmap[i] = g->gmap[i];
        p.gadd_ld[i] = g->gadd_ld[i];
    }
    p.Ep = g->Ep; p.Nn = g->Nn;
    p.resid = reinterpret_cast<const __nv_bfloat16*>(g->resid); p.resid_ld = g->resid_ld;
    p.row_scale = g->row_scale; p.row_scale_mod = g->row_scale_mod > 0 ? g->row_scale_mod : 1;
    p.mask = reinterpret_cast<const __nv_bfloat16*>(g->mask); p.mask_ld = g->mask_ld;
    p.relu = g->relu;
    p.out = reinterpret_cast<__nv_bfloat16*>(g->out);
    p.out_relu = reinterpret_cast<__nv_bfloat16*>(g->out_relu);
    p.ldo = g->ldo;
    p.out_f32 = g->out_f32; p.ldo_f32 = g->ldo_f32;
    p.a_colsum = g->mode == 1 ? g->a_colsum : nullptr;
    p.l2_prefetch = gemm_l2_prefetch_enabled() ? 1 : 0;
    if (g->drop_p > 0.f) {
        if (g->mode != 0 || !g->out_relu || g->out_f32 || g->gadd[0] || g->gadd[1] || g->resid || g->mask || g->drop_p >= 1.f)
            return set_error(RPG_E_ARG, "rpg_gemm: fused dropout needs NT mode, an out_relu output, the plain epilogue, 0 < p < 1");
        p.drop_seed = g->drop_seed;
        p.drop_thresh = (uint32_t)(g->drop_p * 256.0f + 0.5f);
        if (p.drop_thresh > 255u) p.drop_thresh = 255u;
        p.drop_scale = 256.f / (256.f - (float)p.drop_thresh);      // the rate actually applied (p quantised to 1/256): E[out] = in
    }
    p.mask_bits = g->mask_bits; p.mask_bits_ld = g->mask_bits_ld;
    p.out_bits = g->out_bits; p.out_bits_ld = g->out_bits_ld;
    if ((p.mask_bits || p.out_bits) && (g->mode != 0 || p.N % 64 || block_n % 64))
        return set_error(RPG_E_ARG, "rpg_gemm: bit patterns need NT mode and N, block_n multiples of 64");
    if ((p.mask_bits && (p.mask_bits_ld % 8 || (reinterpret_cast<uintptr_t>(p.mask_bits) & 7))) ||
        (p.out_bits && (p.out_bits_ld % 8 || (reinterpret_cast<uintptr_t>(p.out_bits) & 7))))
        return set_error(RPG_E_ARG, "rpg_gemm: bit-pattern pointers and pitches must be 8-byte aligned");
    if ((p.out || p.out_relu) && p.ldo % 8) return set_error(RPG_E_ARG, "rpg_gemm: ldo must be a multiple of 8");
    if (p.out_f32 && p.ldo_f32 % 4) return set_error(RPG_E_ARG, "rpg_gemm: ldo_f32 must be a multiple of 4");
    if ((p.out && !aligned16(p.out)) || (p.out_relu && !aligned16(p.out_relu)) || (p.out_f32 && !aligned16(p.out_f32)))
        return set_error(RPG_E_ARG, "rpg_gemm: output pointers must be 16-byte aligned");
    // result tiles leave through TMA stores: 32 rows x 64 bf16 (or 32 fp32) columns per box, clipped at the tensor edge
    if (g->n_gseg) {
        if (g->mode != 0 || g->n_gseg < 0 || g->n_gseg > 4 || block_n % 64 || g->Ep <= 0 || g->Nn <= 0 || g->gsel_div < 0 ||
            g->gsel_patterns <= 0 || g->gsrc_rows <= 0)
            return set_error(RPG_E_ARG, "rpg_gemm: one-hot gather panels need NT mode, block_n % 64 == 0, Ep, Nn, gsel_div, patterns");
        for (int i = 0; i < g->n_gseg; ++i) {
            if (!g->gsel[i] || !g->gsrc[i] || !aligned16(g->gsel[i]) || !aligned16(g->gsrc[i]) || g->gsrc_ld[i] % 8)
                return set_error(RPG_E_ARG, "rpg_gemm: one-hot gather panel pointers / pitch");
            if ((rc = make_tmap(&tmaps.ga[i], g->gsel[i], BLOCK_K, (uint64_t)g->gsel_patterns * BLOCK_M, BLOCK_K, BLOCK_K, BLOCK_M))) return rc;
            if ((rc = make_tmap(&tmaps.gb[i], g->gsrc[i], g->N, g->gsrc_rows, g->gsrc_ld[i], 64, BLOCK_K))) return rc;
        }
        p.n_gseg = g->n_gseg; p.gsel_div = g->gsel_div; p.Ep = g->Ep; p.Nn = g->Nn;
    }
    for (int i = 0; i < 2; ++i) { p.gadd_f32[i] = g->gadd_f32[i]; p.gadd_f32_ld[i] = g->gadd_f32_ld[i]; }
    p.resid_lo = reinterpret_cast<const __nv_bfloat16*>(g->resid_lo);
    p.out_lo = reinterpret_cast<__nv_bfloat16*>(g->out_lo);
    p.out_relu_lo = reinterpret_cast<__nv_bfloat16*>(g->out_relu_lo);
    const bool split_mode = p.gadd_f32[0] || p.gadd_f32[1] || p.resid_lo || p.out_lo || p.out_relu_lo;
    if (split_mode && g->mode != 0) return set_error(RPG_E_ARG, "rpg_gemm: fp32 (split) operands need NT mode");
    if ((p.out_lo && !p.out) || (p.out_relu_lo && !p.out_relu) || (p.resid_lo && !p.resid))
        return set_error(RPG_E_ARG, "rpg_gemm: a low plane needs its high plane");
    if ((p.gadd_f32[0] && (!g->gmap[0] || g->gadd_f32_ld[0] % 4)) || (p.gadd_f32[1] && (!g->gmap[1] || g->gadd_f32_ld[1] % 4)) ||
        ((p.gadd_f32[0] || p.gadd_f32[1]) && (g->Ep <= 0 || g->Nn <= 0)))
        return set_error(RPG_E_ARG, "rpg_gemm: fp32 gathered add needs gmap, Ep, Nn and a pitch that is a multiple of 4");
    if (p.out && (rc = make_tmap(&tmaps.out, p.out, p.N, p.M, p.ldo, 64, 32))) return rc;
    if (p.out_relu && (rc = make_tmap(&tmaps.out_relu, p.out_relu, p.N, p.M, p.ldo, 64, 32))) return rc;
    if (p.out_lo && (rc = make_tmap(&tmaps.out_lo, p.out_lo, p.N, p.M, p.ldo, 64, 32))) return rc;
    if (p.out_relu_lo && (rc = make_tmap(&tmaps.out_relu_lo, p.out_relu_lo, p.N, p.M, p.ldo, 64, 32))) return rc;
    if (p.out_f32 && (rc = make_tmap_f32_3d(&tmaps.out_f32, p.out_f32, p.N, p.M, p.splits, p.ldo_f32, p.split_stride))) return rc;

    // the shared-memory opt-in is a per-device function attribute: set it once on every device the library is used on
    {
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(g_attr_mu);
        if (dev >= 0 && dev < 64 && !g_attr_done[dev]) {
            cudaDeviceGetAttribute(&g_sm_counts[dev], cudaDevAttrMultiProcessorCount, dev);
            set_smem_attrs();
            g_attr_done[dev] = true;
        }
        g_sm_count = g_sm_counts[dev < 64 && dev >= 0 ? dev : 0];
    }
    const int num_m_blocks = (p.M + BLOCK_M - 1) / BLOCK_M;
    const int num_n_blocks = (p.N + block_n - 1) / block_n;
    const int items = ((num_m_blocks + cl - 1) / cl) * num_n_blocks * p.splits;      // one item per cluster
    const int max_workers = g_sm_count / cl;
    int grid = (items < max_workers ? items : max_workers) * cl;

    const bool prof = prof_active();
    int prof_slot = -1;
    if (prof) {
        const double kdim = g->mode == 0 ? (double)(p.total_kb + p.n_gseg) * BLOCK_K : (double)g->R;
        // algorithmic bytes of the launch: operands read once, results written once
        double bytes = 0.0;
        if (g->mode == 0) {
            bytes = (double)p.M * p.total_kb * BLOCK_K * 2 + (double)p.N * kdim * 2;
            if (p.out) bytes += (double)p.M * p.N * 2;
            if (p.out_relu) bytes += (double)p.M * p.N * 2;
            if (p.out_f32) bytes += (double)p.M * p.N * 4;
            if (p.resid) bytes += (double)p.M * p.N * 2;
            if (p.mask_bits) bytes += (double)p.M * p.N / 8;
            if (p.out_bits) bytes += (double)p.M * p.N / 8;
        } else {
            bytes = (double)g->R * (p.M + p.N) * 2 + (double)p.splits * p.M * p.N * 4;
        }
        prof_slot = prof_open(g->mode == 0 ? RPG_PROF_GEMM_NT : RPG_PROF_GEMM_TN, 2.0 * p.M * p.N * kdim, bytes, p.M, p.N,
                              (int)kdim, stream);
    }
    const bool ops = g->mode == 0 && (p.gadd[0] || p.gadd[1] || p.resid || p.mask);
    // weight-stationary form: plain pair kernel, K <= 512, every cluster bound to one column block (cluster count a
    // multiple of the column-block count), enough items per cluster to amortise the resident B load
    bool ws = false;
    if (cl == 2 && g->mode == 0 && !split_mode && !ops && gemm_pair_enabled() && gemm_ws_enabled() && p.total_kb <= 8 &&
        g->n_gseg == 0 && num_n_blocks <= max_workers) {
        const int workers = (max_workers / num_n_blocks) * num_n_blocks;
        if (workers > 0 && items >= 3 * workers) { ws = true; grid = workers * cl; }
    }
    void (*kern)(GemmTmaps, GemmKParams);
    if (cl == 1)
        kern = g->mode == 1 ? gemm_tc_kernel<1, 1, 0>
                            : (split_mode ? gemm_tc_kernel<0, 1, 2> : (ops ? gemm_tc_kernel<0, 1, 1> : gemm_tc_kernel<0, 1, 0>));
    else if (g->mode == 0 && split_mode && g->n_gseg == 0 && gemm_pair_enabled())
        kern = gemm_tc_kernel<0, 2, 2, true>;           // fp32 mode on CTA pairs
    else if (g->mode == 0 && !split_mode && g->n_gseg == 0 && gemm_pair_enabled()) {
        kern = ops ? gemm_tc_kernel<0, 2, 1, true> : (ws ? gemm_tc_kernel<0, 2, 0, true, true> : gemm_tc_kernel<0, 2, 0, true>);
        if (ops && p.resid && !p.mask && !p.gadd[0] && !p.gadd[1] && aligned16(p.resid) && p.resid_ld % 8 == 0) {
            int rc2 = make_tmap(&tmaps.resid, p.resid, p.N, p.M, p.resid_ld, 64, 32);
            if (rc2) return rc2;
            p.resid_tma = 1;
            kern = gemm_tc_kernel<0, 2, 3, true>;
        }
    }
    else
        kern = g->mode == 1 ? gemm_tc_kernel<1, 2, 0>
                            : (split_mode ? gemm_tc_kernel<0, 2, 2> : (ops ? gemm_tc_kernel<0, 2, 1> : gemm_tc_kernel<0, 2, 0>));
    {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(GEMM_THREADS);
        cfg.dynamicSmemBytes = SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = (pdl_enabled() && !prof) ? 2 : 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmaps, p);
        (void)e;
    }
    if (prof_slot >= 0) prof_close(prof_slot, stream);
    return check_launch("gemm_tc_kernel");
}

}  // namespace rpg

#ifdef RPG_GEMM_TRACE
// Development build only: copies (and optionally clears) the per-CTA wait-cycle counters of gemm_tc_kernel.
extern "C" int rpg_debug_gemm_trace(unsigned long long* out_host, int n_blocks, int reset) {
    if (n_blocks < 0 || n_blocks > 512) return -1;
    cudaDeviceSynchronize();
    if (out_host && n_blocks &&
        cudaMemcpyFromSymbol(out_host, rpg::g_gemm_trace, sizeof(unsigned long long) * 8 * n_blocks) != cudaSuccess) return -2;
    if (reset) {
        void* sym = nullptr;
        if (cudaGetSymbolAddress(&sym, rpg::g_gemm_trace) != cudaSuccess || cudaMemset(sym, 0, sizeof(unsigned long long) * 8 * 512) != cudaSuccess)
            return -3;
    }
    return 0;
}
#endif
