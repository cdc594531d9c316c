// Grouped weight-gradient GEMMs on CTA pairs: C_p[M_p, N_p] = A_p^T B_p over R_p rows for up to TN_GROUP_MAX independent
// problems in ONE persistent launch (the 11 weight gradients of a layer backward, my_gnn_layer.py:232-239,280-282,309-311
// and att.py:20-24 seen from autograd).
//
// Why a second TN kernel next to gemm_tc_kernel<1, ...> (rpg_gemm.cu):
//  * every weight gradient of this path is a single wave of work items (split-R partials sized to the SM count), so a
//    launch per problem pays prologue + ring fill + accumulator drain + teardown once per problem with nothing to
//    amortise them over, and consecutive GEMM launches cannot overlap (each CTA owns all of TMEM and ~225 KB of shared
//    memory).  Here the ring and the double-buffered accumulator run straight through the problem boundaries.
//  * both operands are MN-major activations streamed from HBM/L2; with one CTA per tile the shared-memory port carries
//    48 KB of TMA fill + 48 KB of MMA operand reads per 512-cycle k-block (190 B/clk against 128 B/clk).  As a
//    cta_group::2 pair each CTA stages its own 128 columns of A and HALF of the B tile: 32 + 32 KB per k-block.
//
// Roles per CTA (320 threads): warp 0 lane 0 = TMA producer, warp 1 lane 0 of the LEADER = MMA issuer for the pair,
// warps 2..9 = epilogue (column sums of A out of the ring while the MMAs run, then TMEM -> fp32 partials via TMA store).
#include <cuda.h>

#include <mutex>

#include "rpg_internal.h"
#include "rpg_ptx.cuh"

namespace rpg {

namespace {

constexpr int T_BLOCK_M = 128;                 // output rows (columns of A) per CTA; 256 per pair
constexpr int T_BLOCK_K = 64;                  // contraction rows per k-block
constexpr int T_UMMA_K = 16;
constexpr int T_NS = 6;                        // ring depth
constexpr int T_A_BYTES = T_BLOCK_M * T_BLOCK_K * 2;      // 16 KB: two 64-column boxes
constexpr int T_B_BYTES = 128 * T_BLOCK_K * 2;            // 16 KB: this CTA's half of a 256-column B tile
constexpr int T_EPI_WARPS = 8;                 // 4 accumulator-drain warps + 4 column-sum warps
constexpr int T_DRAIN_WARPS = 4;
constexpr int T_CS_WARPS = T_EPI_WARPS - T_DRAIN_WARPS;
constexpr int T_THREADS = 32 * (2 + T_EPI_WARPS);
constexpr int T_RING_BYTES = T_NS * (T_A_BYTES + T_B_BYTES);
constexpr int T_BAR_BYTES = 1024;
constexpr int T_STG_BYTES = 32 * 128;          // per epilogue warp: 32 rows x 32 fp32
constexpr int T_SMEM_BYTES = T_RING_BYTES + T_BAR_BYTES + (T_DRAIN_WARPS + 1) * T_STG_BYTES + 1024;   // + column-sum scratch
static_assert(T_SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct TnProblem {
    CUtensorMap a;        // A [R, M] bf16, boxes 64 (M) x 64 (R)
    CUtensorMap b;        // B [R, N] bf16, boxes 64 (N) x 64 (R)
    CUtensorMap o;        // partials fp32 [splits][M][N], boxes 32 x 32 x 1
    float* colsum;        // [splits, M] partial column sums of A, or null
    int M, N, block_n, total_kb, splits, kb_per_split, num_n_blocks, pad;
};

struct TnBatch {
    int n;
    int item0[TN_GROUP_MAX + 1];       // first global work item of every problem (pair-items)
    TnProblem pr[TN_GROUP_MAX];
};
static_assert(sizeof(TnBatch) < 16 * 1024, "kernel parameter space");

__global__ void __launch_bounds__(T_THREADS, 1) tn_pair_group_kernel(const __grid_constant__ TnBatch batch) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + T_NS * T_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T_RING_BYTES);
    uint64_t* full_bar = bars;                     // [NS]  leader: both CTAs' tiles of a stage have landed
    uint64_t* empty_bar = bars + T_NS;             // [NS]  per CTA: the local column-sum warps are done with the stage
    uint64_t* done_bar = bars + 2 * T_NS;          // [NS]  per CTA: the pair's MMAs on this stage have retired (multicast commit)
    uint64_t* acc_full = bars + 3 * T_NS;          // [2]
    uint64_t* acc_empty = acc_full + 2;            // [2]   leader: both CTAs' epilogue warps have drained the stage
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    uint8_t* stg_base = smem + T_RING_BYTES + T_BAR_BYTES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&batch.pr[0].a);
        tma_prefetch_desc(&batch.pr[0].b);
        for (int i = 0; i < T_NS; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], T_CS_WARPS);
            mbar_init(&done_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 2 * T_DRAIN_WARPS); }
        fence_mbar_init();
    }
    if (warp == 1) { tmem_alloc_pair(tmem_slot, 512); tmem_relinquish_pair(); }
    if (warp == 2 && lane == 0) tma_prefetch_desc(&batch.pr[0].o);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_prologue();

    const int cta_rank = (int)cluster_ctarank();
    const int worker = blockIdx.x >> 1, num_workers = gridDim.x >> 1;
    const int total_items = batch.item0[batch.n];

    // (problem, tile, split) of a global work item; `pi` only ever moves forward for a worker
    struct Item { int m0, n_blk, split, kb0, n_kb; };
    auto resolve = [&](int item, int& pi) -> Item {
        while (item >= batch.item0[pi + 1]) ++pi;
        const TnProblem& pr = batch.pr[pi];
        const int li = item - batch.item0[pi];
        const int tile = li / pr.splits;
        Item it;
        it.split = li - tile * pr.splits;
        it.n_blk = tile % pr.num_n_blocks;
        it.m0 = ((tile / pr.num_n_blocks) * 2 + cta_rank) * T_BLOCK_M;
        it.kb0 = it.split * pr.kb_per_split;
        const int kb1 = min(it.kb0 + pr.kb_per_split, pr.total_kb);
        it.n_kb = kb1 > it.kb0 ? kb1 - it.kb0 : 0;
        return it;
    };

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------ TMA producer (each CTA: its A columns, its B half)
            int stage = 0, pi = 0;
            uint32_t phase = 0;
            const uint32_t lead_full0 = mapa_u32(smem_u32(&full_bar[0]), 0);
            for (int item = worker; item < total_items; item += num_workers) {
                const Item it = resolve(item, pi);
                const TnProblem& pr = batch.pr[pi];
                const int half_n = pr.block_n >> 1;
                const int nb = half_n >> 6;                                     // 64-column boxes of my B half
                const int n0 = it.n_blk * pr.block_n + cta_rank * half_n;
                const uint32_t tx = 2u * (uint32_t)(T_A_BYTES + nb * 8192);     // both CTAs' bytes land on the leader's barrier
                for (int kb = it.kb0; kb < it.kb0 + it.n_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    const uint32_t lead_full = lead_full0 + stage * 8;
                    if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx);
                    uint8_t* sa = smem_a + stage * T_A_BYTES;
                    uint8_t* sb = smem_b + stage * T_B_BYTES;
                    tma_load_2d_pair(&pr.a, lead_full, sa, it.m0, kb * T_BLOCK_K);
                    tma_load_2d_pair(&pr.a, lead_full, sa + 8192, it.m0 + 64, kb * T_BLOCK_K);
                    for (int j = 0; j < nb; ++j) tma_load_2d_pair(&pr.b, lead_full, sb + j * 8192, n0 + j * 64, kb * T_BLOCK_K);
                    if (++stage == T_NS) { stage = 0; phase ^= 1; }
                }
                if (pi + 1 < batch.n && item + num_workers >= batch.item0[pi + 1]) {   // next item is in the next problem
                    tma_prefetch_desc(&batch.pr[pi + 1].a);
                    tma_prefetch_desc(&batch.pr[pi + 1].b);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && cta_rank == 0) {
            // ------------------------------------------------------------ MMA issuer (leader, for both CTAs)
            int stage = 0, pi = 0, n_it = 0;
            uint32_t phase = 0;
            constexpr uint32_t k_step = (T_UMMA_K * 128) >> 4;                  // 16 contraction rows of 128 B, in 16-byte units
            for (int item = worker; item < total_items; item += num_workers, ++n_it) {
                const Item it = resolve(item, pi);
                const TnProblem& pr = batch.pr[pi];
                const int acc = n_it & 1;
                const uint32_t acc_phase = (n_it >> 1) & 1;
                const uint32_t idesc = make_idesc_bf16(2 * T_BLOCK_M, pr.block_n, 1, 1);
                mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int kb = 0; kb < it.n_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    // MN-major SW128: 64-element MN slabs 8192 B apart (LBO), 8-row K groups 1024 B apart (SBO)
                    const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + stage * T_A_BYTES), 8192u, 1024u);
                    const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + stage * T_B_BYTES), 8192u, 1024u);
#pragma unroll
                    for (int k = 0; k < T_BLOCK_K / T_UMMA_K; ++k)
                        umma_bf16_pair(d_tmem, a_desc + k * k_step, b_desc + k * k_step, idesc, (kb | k) != 0);
                    umma_commit_pair(&done_bar[stage], 3);       // both CTAs: the stage has been consumed
                    if (kb == it.n_kb - 1) umma_commit_pair(&acc_full[acc], 3);
                    if (++stage == T_NS) { stage = 0; phase ^= 1; }
                }
                if (it.n_kb == 0) umma_commit_pair(&acc_full[acc], 3);      // empty split: the epilogue writes zeros
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue warps (both CTAs)
        // Two independent halves, so that neither can hold the other (or the ring) up:
        //   warps 2..5 (one per TMEM lane quadrant) drain finished accumulators to the fp32 partials;
        //   warps 6..9 walk the ring with the tensor core and sum the A tiles over their rows (bias gradients).
        const int ew = warp - 2;
        if (ew < T_DRAIN_WARPS) {
            const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
            uint8_t* const stg = stg_base + ew * T_STG_BYTES;
            uint8_t* const my_row = stg + lane * 128;
            const int my_sw = lane & 7;
            int pi = 0, n_it = 0;
            const uint32_t lead_acc_empty0 = mapa_u32(smem_u32(&acc_empty[0]), 0);
            for (int item = worker; item < total_items; item += num_workers, ++n_it) {
                const Item it = resolve(item, pi);
                const TnProblem& pr = batch.pr[pi];
                const int acc = n_it & 1;
                const uint32_t acc_phase = (n_it >> 1) & 1;
                const int row0 = it.m0 + quad * 32;
                const int n_chunks = pr.block_n >> 6;
                mbar_wait(&acc_full[acc], acc_phase);
                tc_fence_after();
                for (int c = 0; c < n_chunks; ++c) {
                    const int n0 = it.n_blk * pr.block_n + c * 64;
                    float f[64];
                    const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * 256 + c * 64;
                    tmem_ld_32x32(taddr, reinterpret_cast<uint32_t(&)[32]>(f[0]));
                    tmem_ld_32x32(taddr + 32, reinterpret_cast<uint32_t(&)[32]>(f[32]));
                    tmem_ld_wait();
                    if (c == n_chunks - 1) {              // last TMEM read of this item by this warp: hand the stage back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (cta_rank != 0) mbar_arrive_cluster(lead_acc_empty0 + acc * 8);
                            else mbar_arrive(&acc_empty[acc]);
                        }
                    }
                    if (it.n_kb == 0) {
#pragma unroll
                        for (int j = 0; j < 64; ++j) f[j] = 0.f;
                    }
                    if (row0 >= pr.M || n0 >= pr.N) continue;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {               // 32 fp32 columns = 128 B per row per pass
                        if (n0 + h * 32 >= pr.N) break;
                        if (lane == 0) bulk_wait_read_all();    // the previous store of this warp has left the staging tile
                        __syncwarp();
#pragma unroll
                        for (int qq = 0; qq < 8; ++qq)
                            *reinterpret_cast<float4*>(my_row + ((qq ^ my_sw) << 4)) =
                                make_float4(f[h * 32 + 4 * qq], f[h * 32 + 4 * qq + 1], f[h * 32 + 4 * qq + 2], f[h * 32 + 4 * qq + 3]);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) { tma_store_3d(&pr.o, stg, n0 + h * 32, row0, it.split); bulk_commit(); }
                    }
                }
                if (lane == 0 && pi + 1 < batch.n && item + num_workers >= batch.item0[pi + 1]) tma_prefetch_desc(&batch.pr[pi + 1].o);
            }
            if (lane == 0) bulk_wait_all();            // all results are in global memory before the CTA retires
        } else {
            const int et = threadIdx.x - 32 * (2 + T_DRAIN_WARPS);    // 0..127 over the column-sum warps
            const int c16 = et & 15, rg = et >> 4;                      // 16-byte column chunk, row group (8 groups of 8 rows)
            float* const red = reinterpret_cast<float*>(stg_base + T_DRAIN_WARPS * T_STG_BYTES);   // [8][128]
            int cs_stage = 0, pi = 0;
            uint32_t cs_phase = 0;
            // A stage is recycled by these warps: they wait until the pair's MMAs on it have retired (the leader's commit
            // is multicast to both CTAs -- no software hand-off between the CTAs), add their rows of the A tile to the
            // column sums, and only then hand the stage back to the local producer.
            for (int item = worker; item < total_items; item += num_workers) {
                const Item it = resolve(item, pi);
                const TnProblem& pr = batch.pr[pi];
                const bool do_cs = pr.colsum != nullptr && it.n_blk == 0;
                float cs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int kb = 0; kb < it.n_kb; ++kb) {
                    mbar_wait(&done_bar[cs_stage], cs_phase);
                    if (do_cs) {
                        const uint8_t* a_tile = smem_a + cs_stage * T_A_BYTES + (c16 >> 3) * 8192;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int r = rg + 8 * j;
                            const uint4 u = *reinterpret_cast<const uint4*>(a_tile + r * 128 + (((c16 & 7) ^ (r & 7)) << 4));
                            cs[0] += bf16_lo(u.x); cs[1] += bf16_hi(u.x); cs[2] += bf16_lo(u.y); cs[3] += bf16_hi(u.y);
                            cs[4] += bf16_lo(u.z); cs[5] += bf16_hi(u.z); cs[6] += bf16_lo(u.w); cs[7] += bf16_hi(u.w);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[cs_stage]);
                    if (++cs_stage == T_NS) { cs_stage = 0; cs_phase ^= 1; }
                }
                if (do_cs) {
                    // fold the 8 row groups through the scratch tile, then one thread per column writes.  Fixed order.
                    asm volatile("bar.sync 1, 128;" ::: "memory");               // the previous item's readers are done
#pragma unroll
                    for (int q = 0; q < 8; ++q) red[rg * 128 + c16 * 8 + q] = cs[q];
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    float t = 0.f;
#pragma unroll
                    for (int g2 = 0; g2 < 8; ++g2) t += red[g2 * 128 + et];
                    const int col = it.m0 + et;
                    if (col < pr.M) pr.colsum[(size_t)it.split * pr.M + col] = t;
                }
            }
        }
    }

    __syncwarp();
    tc_fence_before();
    cluster_sync_all();                            // no CTA may exit while its peer can still signal it
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

bool tn_group_env() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("RPG_TN_GROUP");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

std::mutex g_mu;
bool g_attr_done[64] = {false};
int g_sms[64] = {0};

int device_sms() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_attr_done[dev]) {
        cudaDeviceGetAttribute(&g_sms[dev], cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(tn_pair_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM_BYTES);
        g_attr_done[dev] = true;
    }
    return g_sms[dev];
}

}  // namespace

bool tn_group_supported(const rpg_bf16* A, int lda, int M, const rpg_bf16* B, int ldb, int N, long long R) {
    if (!tn_group_env()) return false;
    if (!A || !B || M <= 0 || N < 64 || N % 64 || M % 8 || R <= 0 || R > 0x7fffffffLL) return false;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || lda % 8 || ldb % 8) return false;
    return true;
}

static int tn_block_n(int N) { return N >= 256 ? 256 : 128; }

int tn_group_splits(int M, int N, long long R, int sm_count) {
    const int block_n = tn_block_n(N);
    const int tiles = ((M + 2 * T_BLOCK_M - 1) / (2 * T_BLOCK_M)) * ((N + block_n - 1) / block_n);
    const long long kb = (R + T_BLOCK_K - 1) / T_BLOCK_K;
    long long splits = (sm_count / 2) / tiles;                        // one wave of pair-items per problem
    if (splits > kb) splits = kb;
    if (splits < 1) splits = 1;
    const long long per = (kb + splits - 1) / splits;
    return (int)((kb + per - 1) / per);                               // no empty splits
}

int tn_group_launch(const TnDesc* d, int n, cudaStream_t stream) {
    if (!d || n < 1 || n > TN_GROUP_MAX) return set_error(RPG_E_ARG, "tn_group: 1..TN_GROUP_MAX problems");
    const int sms = device_sms();
    TnBatch batch;
    memset(&batch, 0, sizeof batch);
    batch.n = n;
    double flops = 0.0, bytes = 0.0;
    int rc;
    for (int i = 0; i < n; ++i) {
        const TnDesc& t = d[i];
        if (!tn_group_supported(t.A, t.lda, t.M, t.B, t.ldb, t.N, t.R) || !t.part || t.splits < 1 ||
            (reinterpret_cast<uintptr_t>(t.part) & 15))
            return set_error(RPG_E_ARG, "tn_group: unsupported problem (N % 64, M % 8, 16-byte aligned operands)");
        TnProblem& p = batch.pr[i];
        p.M = t.M; p.N = t.N; p.block_n = tn_block_n(t.N);
        p.total_kb = (int)((t.R + T_BLOCK_K - 1) / T_BLOCK_K);
        p.splits = t.splits;
        p.kb_per_split = (p.total_kb + t.splits - 1) / t.splits;
        p.num_n_blocks = (t.N + p.block_n - 1) / p.block_n;
        p.colsum = t.colsum;
        {
            const uint64_t dims[2] = {(uint64_t)t.M, (uint64_t)t.R}, strides[1] = {(uint64_t)t.lda * 2};
            const uint32_t box[2] = {64, T_BLOCK_K};
            if ((rc = tmap_encode(&p.a, 0, 2, t.A, dims, strides, box))) return rc;
        }
        {
            const uint64_t dims[2] = {(uint64_t)t.N, (uint64_t)t.R}, strides[1] = {(uint64_t)t.ldb * 2};
            const uint32_t box[2] = {64, T_BLOCK_K};
            if ((rc = tmap_encode(&p.b, 0, 2, t.B, dims, strides, box))) return rc;
        }
        {
            const uint64_t dims[3] = {(uint64_t)t.N, (uint64_t)t.M, (uint64_t)t.splits};
            const uint64_t strides[2] = {(uint64_t)t.N * 4, (uint64_t)t.M * t.N * 4};
            const uint32_t box[3] = {32, 32, 1};
            if ((rc = tmap_encode(&p.o, 1, 3, t.part, dims, strides, box))) return rc;
        }
        const int m_units = (t.M + 2 * T_BLOCK_M - 1) / (2 * T_BLOCK_M);
        batch.item0[i + 1] = batch.item0[i] + m_units * p.num_n_blocks * p.splits;
        flops += 2.0 * t.M * t.N * (double)t.R;
        bytes += (double)t.R * (t.M + t.N) * 2 + (double)t.splits * t.M * t.N * 4;
    }
    const int total = batch.item0[n];
    const int max_workers = sms / 2;
    const int grid = 2 * (total < max_workers ? total : max_workers);
    const bool prof = prof_active();
    int slot = -1;
    if (prof) slot = prof_open(RPG_PROF_GEMM_TN, flops, bytes, 0, 0, n, stream);
    {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(T_THREADS);
        cfg.dynamicSmemBytes = T_SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = (pdl_enabled() && !prof) ? 2 : 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, tn_pair_group_kernel, batch);
        (void)e;
    }
    if (slot >= 0) prof_close(slot, stream);
    return check_launch("tn_pair_group_kernel");
}

}  // namespace rpg
