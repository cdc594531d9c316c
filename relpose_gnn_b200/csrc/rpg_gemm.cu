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
constexpr int SMEM_BYTES = SMEM_RING_BYTES + 256 + 1024;    // + barriers + alignment slack

struct GemmKParams {
    int mode, M, N, block_n;
    int num_kb[3];                 // NT: 64-wide k-blocks per A segment
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

template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB,
               const GemmKParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_RING_BYTES);
    uint64_t* full_bar = bars;                    // [STAGES]
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]
    uint64_t* acc_full = bars + 2 * STAGES;       // [ACC_STAGES]
    uint64_t* acc_empty = acc_full + ACC_STAGES;  // [ACC_STAGES]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], EPI_WARPS); }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_m_blocks = (p.M + BLOCK_M - 1) / BLOCK_M;
    const int num_n_blocks = (p.N + p.block_n - 1) / p.block_n;
    const int num_tiles = num_m_blocks * num_n_blocks;
    const int num_items = MODE == 0 ? num_tiles : num_tiles * p.splits;
    const uint32_t stage_tx_bytes = A_STAGE_BYTES + p.block_n * BLOCK_K * 2;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------ TMA producer
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int tile = MODE == 0 ? item : item / p.splits;
                const int m_blk = tile / num_n_blocks, n_blk = tile % num_n_blocks;
                if (MODE == 0) {
                    int kb_global = 0;
                    for (int s = 0; s < 3; ++s) {
                        const CUtensorMap* tmA = s == 0 ? &tmA0 : (s == 1 ? &tmA1 : &tmA2);
                        for (int kb = 0; kb < p.num_kb[s]; ++kb, ++kb_global) {
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            mbar_arrive_expect_tx(&full_bar[stage], stage_tx_bytes);
                            tma_load_2d(tmA, &full_bar[stage], smem_a + stage * A_STAGE_BYTES, kb * BLOCK_K,
                                        m_blk * BLOCK_M);
                            tma_load_2d(&tmB, &full_bar[stage], smem_b + stage * B_STAGE_BYTES, kb_global * BLOCK_K,
                                        n_blk * p.block_n);
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                } else {
                    const int split = item % p.splits;
                    const int kb0 = split * p.kb_per_split;
                    const int kb1 = min(kb0 + p.kb_per_split, p.total_kb);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&full_bar[stage], stage_tx_bytes);
                        // MN-major tiles: one 64(MN) x 64(K rows) box per 128-byte slab of the MN extent
                        for (int j = 0; j < BLOCK_M / 64; ++j)
                            tma_load_2d(&tmA0, &full_bar[stage], smem_a + stage * A_STAGE_BYTES + j * 8192,
                                        m_blk * BLOCK_M + j * 64, kb * BLOCK_K);
                        for (int j = 0; j < p.block_n / 64; ++j)
                            tma_load_2d(&tmB, &full_bar[stage], smem_b + stage * B_STAGE_BYTES + j * 8192,
                                        n_blk * p.block_n + j * 64, kb * BLOCK_K);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------------------ MMA issuer (single thread)
            const uint32_t idesc = make_idesc_bf16(BLOCK_M, p.block_n, MODE, MODE);
            // K-major SW128: 8-row groups 1024 B apart (SBO), LBO unused.
            // MN-major SW128: 64-element MN slabs 8192 B apart (LBO), 8-row K groups 1024 B apart (SBO).
            const uint32_t lbo = MODE == 0 ? 0u : 8192u, sbo = 1024u;
            const uint32_t k_step = MODE == 0 ? (UMMA_K * 2) >> 4 : (UMMA_K * 128) >> 4;   // desc address units (16 B)
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                int n_kb;
                if (MODE == 0) {
                    n_kb = p.total_kb;
                } else {
                    const int kb0 = (item % p.splits) * p.kb_per_split;
                    n_kb = min(kb0 + p.kb_per_split, p.total_kb) - kb0;
                }
                mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + stage * A_STAGE_BYTES), lbo, sbo);
                    const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + stage * B_STAGE_BYTES), lbo, sbo);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        umma_bf16(d_tmem, a_desc + k * k_step, b_desc + k * k_step, idesc, (kb | k) != 0);
                    umma_commit(&empty_bar[stage]);           // frees the smem slot when these MMAs retire
                    if (kb == n_kb - 1) umma_commit(&acc_full[acc]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (n_kb <= 0) umma_commit(&acc_full[acc]);   // empty split: nothing to accumulate (epilogue writes zeros)
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue warps
        const int ew = warp - 2;
        const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
        const int half = ew >> 2;                  // column interleave between the two warps of a quadrant
        const int n_chunks = (p.block_n + 31) / 32;
        int it = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int tile = MODE == 0 ? item : item / p.splits;
            const int m_blk = tile / num_n_blocks, n_blk = tile % num_n_blocks;
            const int row = m_blk * BLOCK_M + quad * 32 + lane;
            const bool row_ok = row < p.M;
            bool zero_acc = false;
            if (MODE == 1) {
                const int kb0 = (item % p.splits) * p.kb_per_split;
                zero_acc = kb0 >= p.total_kb;
            }
            mbar_wait(&acc_full[acc], acc_phase);
            tc_fence_after();

            const __nv_bfloat16* g0 = nullptr;
            const __nv_bfloat16* g1 = nullptr;
            float rscale = 1.f;
            if (MODE == 0 && row_ok) {
                if (p.gadd[0] || p.gadd[1]) {
                    const int gidx = row / p.Ep, k = row - gidx * p.Ep;
                    if (p.gadd[0]) g0 = p.gadd[0] + (size_t)(gidx * p.Nn + __ldg(p.gmap[0] + k)) * p.gadd_ld[0];
                    if (p.gadd[1]) g1 = p.gadd[1] + (size_t)(gidx * p.Nn + __ldg(p.gmap[1] + k)) * p.gadd_ld[1];
                }
                if (p.row_scale) rscale = __ldg(p.row_scale + (row % p.row_scale_mod));
            }

            for (int c = half; c < n_chunks; c += 2) {
                const int n0 = n_blk * p.block_n + c * 32;
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + acc * ACC_STRIDE + c * 32, v);
                tmem_ld_wait();
                if (!row_ok || n0 >= p.N) continue;
                const int nvalid = min(32, min(p.N - n0, p.block_n - c * 32));   // multiple of 8
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = zero_acc ? 0.f : __uint_as_float(v[j]);

                if (MODE == 0) {
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (j < nvalid) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                                f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
                            }
                        }
                    }
                    if (g0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8)
                            if (j < nvalid) add_bf16x8(f + j, __ldg(reinterpret_cast<const uint4*>(g0 + n0 + j)));
                    }
                    if (g1) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8)
                            if (j < nvalid) add_bf16x8(f + j, __ldg(reinterpret_cast<const uint4*>(g1 + n0 + j)));
                    }
                    if (p.resid) {
                        const __nv_bfloat16* r = p.resid + (size_t)row * p.resid_ld + n0;
#pragma unroll
                        for (int j = 0; j < 32; j += 8)
                            if (j < nvalid) add_bf16x8(f + j, __ldg(reinterpret_cast<const uint4*>(r + j)));
                    }
                    if (p.row_scale) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] *= rscale;
                    }
                    if (p.mask) {
                        const __nv_bfloat16* mk = p.mask + (size_t)row * p.mask_ld + n0;
#pragma unroll
                        for (int j = 0; j < 32; j += 8)
                            if (j < nvalid) mask_bf16x8(f + j, __ldg(reinterpret_cast<const uint4*>(mk + j)));
                    }
                    if (p.relu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                    }
                    if (p.out) {
                        __nv_bfloat16* o = p.out + (size_t)row * p.ldo + n0;
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            if (j < nvalid) {
                                uint4 u;
                                u.x = pack_bf16x2(f[j], f[j + 1]); u.y = pack_bf16x2(f[j + 2], f[j + 3]);
                                u.z = pack_bf16x2(f[j + 4], f[j + 5]); u.w = pack_bf16x2(f[j + 6], f[j + 7]);
                                *reinterpret_cast<uint4*>(o + j) = u;
                            }
                        }
                    }
                    if (p.out_relu) {
                        __nv_bfloat16* o = p.out_relu + (size_t)row * p.ldo + n0;
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            if (j < nvalid) {
                                uint4 u;
                                u.x = pack_bf16x2(fmaxf(f[j], 0.f), fmaxf(f[j + 1], 0.f));
                                u.y = pack_bf16x2(fmaxf(f[j + 2], 0.f), fmaxf(f[j + 3], 0.f));
                                u.z = pack_bf16x2(fmaxf(f[j + 4], 0.f), fmaxf(f[j + 5], 0.f));
                                u.w = pack_bf16x2(fmaxf(f[j + 6], 0.f), fmaxf(f[j + 7], 0.f));
                                *reinterpret_cast<uint4*>(o + j) = u;
                            }
                        }
                    }
                }
                if (p.out_f32) {
                    float* o = p.out_f32 + (MODE == 1 ? (size_t)(item % p.splits) * p.split_stride : 0) +
                               (size_t)row * p.ldo_f32 + n0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (j < nvalid) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                }
            }
            // all tcgen05.ld of this accumulator stage have completed (wait::ld above): hand it back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
        }
    }

    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// -------------------------------------------------------------------------------------- host side

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

// 2-D bf16 tensor map: inner extent `inner` (contiguous), outer extent `outer`, row pitch `ld` elements.
static int make_tmap(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                     uint32_t box_inner, uint32_t box_outer) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_error(RPG_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[160];
        snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu ld=%llu box=%ux%u", (int)r,
                 (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer);
        return set_error((int)r, msg);
    }
    return 0;
}

// ---- per-launch event timing (enabled only between rpg_profile_begin / rpg_profile_end)
struct ProfRec { cudaEvent_t e0, e1; int mode; double flops; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;

int profile_begin() {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.clear();
    g_prof_on = true;
    return 0;
}

int profile_end(double* nt_ms, double* tn_ms, int* nt_launches, int* tn_launches, double* nt_flops, double* tn_flops) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = false;
    double ms[2] = {0, 0}, fl[2] = {0, 0};
    int n[2] = {0, 0};
    for (auto& r : g_prof) {
        cudaEventSynchronize(r.e1);
        float t = 0.f;
        cudaEventElapsedTime(&t, r.e0, r.e1);
        ms[r.mode] += t; fl[r.mode] += r.flops; n[r.mode]++;
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    g_prof.clear();
    if (nt_ms) *nt_ms = ms[0];
    if (tn_ms) *tn_ms = ms[1];
    if (nt_launches) *nt_launches = n[0];
    if (tn_launches) *tn_launches = n[1];
    if (nt_flops) *nt_flops = fl[0];
    if (tn_flops) *tn_flops = fl[1];
    return 0;
}

static int g_sm_count = 0;
static std::once_flag g_attr_once;

static int aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

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

    GemmKParams p;
    memset(&p, 0, sizeof p);
    p.mode = g->mode; p.M = g->M; p.N = g->N; p.block_n = block_n;
    CUtensorMap tmA[3], tmB;
    memset(tmA, 0, sizeof tmA);
    int rc;
    if (g->mode == 0) {
        if (g->n_seg < 1 || g->n_seg > 3) return set_error(RPG_E_ARG, "rpg_gemm: n_seg must be 1..3");
        int ktot = 0;
        for (int s = 0; s < g->n_seg; ++s) {
            if (!g->A[s] || !aligned16(g->A[s]) || g->lda[s] % 8 || g->K[s] <= 0 || g->K[s] % BLOCK_K)
                return set_error(RPG_E_ARG, "rpg_gemm: A segment pointer/pitch/K (K must be a multiple of 64)");
            p.num_kb[s] = g->K[s] / BLOCK_K;
            ktot += g->K[s];
            if ((rc = make_tmap(&tmA[s], g->A[s], g->K[s], g->M, g->lda[s], BLOCK_K, BLOCK_M))) return rc;
        }
        for (int s = g->n_seg; s < 3; ++s) tmA[s] = tmA[0];
        p.total_kb = ktot / BLOCK_K;
        if ((rc = make_tmap(&tmB, g->B, ktot, g->N, g->ldb, BLOCK_K, block_n))) return rc;
        p.splits = 1;
        if ((g->gadd[0] && !g->gmap[0]) || (g->gadd[1] && !g->gmap[1]) || ((g->gadd[0] || g->gadd[1]) && (g->Ep <= 0 || g->Nn <= 0)))
            return set_error(RPG_E_ARG, "rpg_gemm: gathered add needs gmap, Ep, Nn");
        if (!g->out && !g->out_relu && !g->out_f32) return set_error(RPG_E_ARG, "rpg_gemm: no output");
    } else if (g->mode == 1) {
        if (!g->A[0] || !aligned16(g->A[0]) || g->lda[0] % 8 || g->R <= 0)
            return set_error(RPG_E_ARG, "rpg_gemm: TN operand A");
        if (!g->out_f32 || g->splits < 1) return set_error(RPG_E_ARG, "rpg_gemm: TN mode writes fp32 partials");
        p.total_kb = (g->R + BLOCK_K - 1) / BLOCK_K;
        p.splits = g->splits;
        p.kb_per_split = (p.total_kb + g->splits - 1) / g->splits;
        p.split_stride = g->split_stride;
        if ((rc = make_tmap(&tmA[0], g->A[0], g->M, g->R, g->lda[0], 64, BLOCK_K))) return rc;
        tmA[1] = tmA[2] = tmA[0];
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
    if ((p.out || p.out_relu) && p.ldo % 8) return set_error(RPG_E_ARG, "rpg_gemm: ldo must be a multiple of 8");
    if (p.out_f32 && p.ldo_f32 % 4) return set_error(RPG_E_ARG, "rpg_gemm: ldo_f32 must be a multiple of 4");

    std::call_once(g_attr_once, [] {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(gemm_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    });
    const int num_m_blocks = (p.M + BLOCK_M - 1) / BLOCK_M;
    const int num_n_blocks = (p.N + block_n - 1) / block_n;
    const int items = num_m_blocks * num_n_blocks * p.splits;
    const int grid = items < g_sm_count ? items : g_sm_count;
    ProfRec rec;
    bool prof = false;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        prof = g_prof_on;
    }
    if (prof) {
        cudaEventCreate(&rec.e0); cudaEventCreate(&rec.e1);
        rec.mode = g->mode;
        rec.flops = 2.0 * p.M * p.N * (g->mode == 0 ? (double)p.total_kb * BLOCK_K : (double)g->R);
        cudaEventRecord(rec.e0, stream);
    }
    if (g->mode == 0)
        gemm_tc_kernel<0><<<grid, GEMM_THREADS, SMEM_BYTES, stream>>>(tmA[0], tmA[1], tmA[2], tmB, p);
    else
        gemm_tc_kernel<1><<<grid, GEMM_THREADS, SMEM_BYTES, stream>>>(tmA[0], tmA[1], tmA[2], tmB, p);
    if (prof) {
        cudaEventRecord(rec.e1, stream);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof.push_back(rec);
    }
    return check_launch("gemm_tc_kernel");
}

}  // namespace rpg
