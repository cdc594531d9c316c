// Host-facing utilities of the path that are not tensor-core or bandwidth kernels of the layer itself:
//  * template tables built on the device (a fresh edge_index needs no host round trip),
//  * batched weight packing (one launch re-packs every operand after an optimizer step),
//  * the Adam step over the flat parameter / gradient buckets (train.py:211,274).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/rpg.h"
#include "rpg_internal.h"
#include "rpg_ptx.cuh"

namespace rpg {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------
// Template tables: four CSRs (by destination / source / lower / upper endpoint) + degrees of ONE graph template.
// One block; thread n owns node n (strided): count pass, block-wide exclusive scan, stable fill pass.
// ------------------------------------------------------------------------------------------------
constexpr int TT_THREADS = 128;          // one warp per CSR (destination / source / lower / upper endpoint)
constexpr int TT_MAX_N = 1024;

__device__ __forceinline__ int tt_key(int t, int s, int d) { return t == 0 ? d : t == 1 ? s : t == 2 ? min(s, d) : max(s, d); }

// Warp t builds CSR t: for every node n in turn the lanes scan the template 32 edges at a time, a ballot marks the edges
// keyed by n and each marked lane writes its edge at (running offset + number of marked lanes below it) -- a stable
// counting sort (edge order within a node), identical to the host tables.  ~N * Ep / 32 ballots: microseconds for the
// templates of the path (N <= 17, Ep <= 272).
__global__ void __launch_bounds__(TT_THREADS)
template_tables_kernel(const int* __restrict__ tsrc, const int* __restrict__ tdst, int N, int Ep, int* __restrict__ src,
                       int* __restrict__ dst, int* __restrict__ in_ptr, int* __restrict__ in_idx, int* __restrict__ out_ptr,
                       int* __restrict__ out_idx, int* __restrict__ min_ptr, int* __restrict__ min_idx,
                       int* __restrict__ max_ptr, int* __restrict__ max_idx, float* __restrict__ inv_deg,
                       float* __restrict__ deg, float* __restrict__ has_in) {
    pdl_prologue();
    for (int k = threadIdx.x; k < Ep; k += TT_THREADS) { src[k] = tsrc[k]; dst[k] = tdst[k]; }
    const int t = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* ptr = t == 0 ? in_ptr : t == 1 ? out_ptr : t == 2 ? min_ptr : max_ptr;
    int* idx = t == 0 ? in_idx : t == 1 ? out_idx : t == 2 ? min_idx : max_idx;
    int base = 0;
    for (int n = 0; n < N; ++n) {
        const int start = base;
        for (int k0 = 0; k0 < Ep; k0 += 32) {
            const int k = k0 + lane;
            const bool hit = k < Ep && tt_key(t, __ldg(tsrc + k), __ldg(tdst + k)) == n;
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (hit) idx[base + __popc(m & ((1u << lane) - 1u))] = k;
            base += __popc(m);
        }
        if (lane == 0) {
            ptr[n] = start;
            if (t == 0) {
                const int c = base - start;
                deg[n] = (float)c;
                inv_deg[n] = 1.f / (float)max(c, 1);
                has_in[n] = c > 0 ? 1.f : 0.f;
            }
        }
    }
    if (lane == 0) ptr[N] = base;
}

// ------------------------------------------------------------------------------------------------
// Batched weight packing: blockIdx.y = descriptor, blockIdx.x strides over its 32 x 32 tiles.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_weights_batch_kernel(const __grid_constant__ rpg_pack_batch_t batch) {
    pdl_prologue();
    __shared__ float tile[32][33];
    const rpg_pack_desc_t& d = batch.d[blockIdx.y];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8
    const int tiles_x = (d.cols + 31) / 32, tiles_y = (d.rows + 31) / 32;
    for (int tl = blockIdx.x; tl < tiles_x * tiles_y; tl += gridDim.x) {
        const int bx = (tl % tiles_x) * 32, by = (tl / tiles_x) * 32;
        for (int j = ty; j < 32; j += 8) {
            const int r = by + j, c = bx + tx;
            float v = (r < d.rows && c < d.cols) ? d.src[(size_t)(d.r0 + r) * d.ld_src + d.c0 + c] : 0.f;
            if (d.lo_plane) v = v - __bfloat162float(__float2bfloat16_rn(v));
            tile[j][tx] = v;
        }
        __syncthreads();
        for (int j = ty; j < 32; j += 8) {
            int r, c;
            float v;
            if (!d.transpose) { r = by + j; c = bx + tx; v = tile[j][tx]; }
            else { c = bx + j; r = by + tx; v = tile[tx][j]; }
            if (r < d.rows && c < d.cols) {
                const size_t o = d.transpose ? (size_t)c * d.ld_dst + r : (size_t)r * d.ld_dst + c;
                if (d.dst_f32) reinterpret_cast<float*>(d.dst)[o] = v;
                else reinterpret_cast<bf16*>(d.dst)[o] = __float2bfloat16_rn(v);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam: L2 weight decay in the gradient, bias correction, no amsgrad), 4 elements per thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 long long n, float lr_c, float b1, float b2, float eps, float wd, float gscale, float inv_sqrt_bc2) {
    pdl_prologue();
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n) {
            float4 pp = *reinterpret_cast<float4*>(p + i);
            const float4 gg = *reinterpret_cast<const float4*>(g + i);
            float4 mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
            float* P = &pp.x; const float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float gr = fmaf(wd, P[q], G[q] * gscale);
                M[q] = fmaf(b1, M[q], (1.f - b1) * gr);
                V[q] = fmaf(b2, V[q], (1.f - b2) * gr * gr);
                P[q] -= lr_c * M[q] / (sqrtf(V[q]) * inv_sqrt_bc2 + eps);
            }
            *reinterpret_cast<float4*>(p + i) = pp;
            *reinterpret_cast<float4*>(m + i) = mm;
            *reinterpret_cast<float4*>(v + i) = vv;
        } else {
            for (long long j = i; j < n; ++j) {
                const float gr = fmaf(wd, p[j], g[j] * gscale);
                m[j] = fmaf(b1, m[j], (1.f - b1) * gr);
                v[j] = fmaf(b2, v[j], (1.f - b2) * gr * gr);
                p[j] -= lr_c * m[j] / (sqrtf(v[j]) * inv_sqrt_bc2 + eps);
            }
        }
    }
}

}  // namespace rpg

using namespace rpg;

extern "C" {

static inline long long al4(long long v) { return (v + 3) / 4 * 4; }

int64_t rpg_template_tables_words(int N, int Ep) { return 6 * al4(Ep) + 4 * al4(N + 1) + 3 * al4(N); }

int rpg_template_tables(const int32_t* tsrc, const int32_t* tdst, int N, int Ep, int32_t* tables, rpg_stream_t stream) {
    if (!tsrc || !tdst || !tables || N <= 0 || Ep <= 0) return set_error(RPG_E_ARG, "template_tables: bad arguments");
    if (N > TT_MAX_N || Ep > 65536 || (long long)N * Ep > (1LL << 22))
        return set_error(RPG_E_UNSUPPORTED, "template_tables: N <= 1024, Ep <= 65536, N * Ep <= 2^22 (larger templates: host tables)");
    int32_t* p = tables;
    int32_t* src = p; p += al4(Ep);
    int32_t* dst = p; p += al4(Ep);
    int32_t* in_ptr = p; p += al4(N + 1);
    int32_t* in_idx = p; p += al4(Ep);
    int32_t* out_ptr = p; p += al4(N + 1);
    int32_t* out_idx = p; p += al4(Ep);
    int32_t* min_ptr = p; p += al4(N + 1);
    int32_t* min_idx = p; p += al4(Ep);
    int32_t* max_ptr = p; p += al4(N + 1);
    int32_t* max_idx = p; p += al4(Ep);
    float* inv_deg = reinterpret_cast<float*>(p); p += al4(N);
    float* deg = reinterpret_cast<float*>(p); p += al4(N);
    float* has_in = reinterpret_cast<float*>(p);
    launch_pdl(template_tables_kernel, dim3(1), dim3(TT_THREADS), 0, as_stream(stream), tsrc, tdst, N, Ep, src, dst, in_ptr,
               in_idx, out_ptr, out_idx, min_ptr, min_idx, max_ptr, max_idx, inv_deg, deg, has_in);
    return check_launch("template_tables_kernel");
}

int rpg_pack_weights_batch(const rpg_pack_batch_t* batch, rpg_stream_t stream) {
    if (!batch || batch->n < 1 || batch->n > RPG_PACK_BATCH_MAX) return set_error(RPG_E_ARG, "pack_weights_batch: bad arguments");
    int most = 1;
    for (int i = 0; i < batch->n; ++i) {
        const rpg_pack_desc_t& d = batch->d[i];
        if (!d.src || !d.dst || d.rows <= 0 || d.cols <= 0) return set_error(RPG_E_ARG, "pack_weights_batch: bad descriptor");
        const int tiles = ((d.cols + 31) / 32) * ((d.rows + 31) / 32);
        if (tiles > most) most = tiles;
    }
    static_assert(sizeof(rpg_pack_batch_t) <= 4000, "kernel parameter space");
    launch_pdl(pack_weights_batch_kernel, dim3((unsigned)(most > 256 ? 256 : most), (unsigned)batch->n), dim3(256), 0,
               as_stream(stream), *batch);
    return check_launch("pack_weights_batch_kernel");
}

int rpg_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, float grad_scale, int64_t step, rpg_stream_t stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq || n <= 0 || step < 1) return set_error(RPG_E_ARG, "adam_step: bad arguments");
    if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
         reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
        return set_error(RPG_E_ARG, "adam_step: buffers must be 16-byte aligned");
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    long long blocks = (n / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_pdl(adam_step_kernel, dim3((unsigned)blocks), dim3(256), 0, as_stream(stream), param, grad, exp_avg, exp_avg_sq,
               (long long)n, (float)(lr / bc1), beta1, beta2, eps, weight_decay, grad_scale, (float)(1.0 / sqrt(bc2)));
    return check_launch("adam_step_kernel");
}

int rpg_profile_records(rpg_prof_rec_t* out, int max_records, int* n_records) { return profile_records(out, max_records, n_records); }

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Small fp32 products (weight composition Wgc = Wgtp W2m and the matching backward), ~50 MFLOP each.
// ------------------------------------------------------------------------------------------------
namespace rpg {
// BM x BN output tile per block (64 x 64, or 32 x 32 for launches that would leave most SMs idle), 32-deep k chunks,
// (BM/16) x (BN/16) outputs per thread; the next chunk's global loads are issued before the current chunk's FMAs
// (register prefetch), and the thread -> element mapping of the loads follows the operand's storage order so that they
// coalesce for every transpose flag.  Fixed summation order: deterministic.
constexpr int SG_BK = 32;
template <int BM, int BN>
__global__ void __launch_bounds__(256)
sgemm_batch_kernel(const __grid_constant__ rpg_sgemm_batch_t batch) {
    constexpr int TM = BM / 16, TN = BN / 16;                        // outputs per thread
    constexpr int LA = BM * SG_BK / 256, LB = BN * SG_BK / 256;      // operand elements per thread and chunk
    pdl_prologue();
    const rpg_sgemm_desc_t& d = batch.d[blockIdx.z];
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (m0 >= d.M || n0 >= d.N) return;
    __shared__ __align__(16) float As[SG_BK][BM + 4], Bs[SG_BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;                         // outputs: rows m0 + TM ty + i, columns n0 + TN tx + j
    // A tile element (m, k): stored A[m, k] (k contiguous) or, transA, A[k, m] (m contiguous)
    int am[LA], ak[LA], bk[LB], bn[LB];
#pragma unroll
    for (int i = 0; i < LA; ++i) {
        const int e = tid + 256 * i;
        if (d.transA) { am[i] = e % BM; ak[i] = e / BM; } else { ak[i] = e % SG_BK; am[i] = e / SG_BK; }
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
        const int e = tid + 256 * i;
        if (d.transB) { bk[i] = e % SG_BK; bn[i] = e / SG_BK; } else { bn[i] = e % BN; bk[i] = e / BN; }
    }
    auto ldA = [&](int k0, int i) -> float {
        const int m = m0 + am[i], k = k0 + ak[i];
        if (m >= d.M || k >= d.K) return 0.f;
        return d.transA ? d.A[(size_t)k * d.lda + m] : d.A[(size_t)m * d.lda + k];
    };
    auto ldB = [&](int k0, int i) -> float {
        const int n = n0 + bn[i], k = k0 + bk[i];
        if (n >= d.N || k >= d.K) return 0.f;
        return d.transB ? d.B[(size_t)n * d.ldb + k] : d.B[(size_t)k * d.ldb + n];
    };
    float ra[LA], rb[LB], acc[TM][TN];
#pragma unroll
    for (int i = 0; i < LA; ++i) ra[i] = ldA(0, i);
#pragma unroll
    for (int i = 0; i < LB; ++i) rb[i] = ldB(0, i);
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < d.K; k0 += SG_BK) {
#pragma unroll
        for (int i = 0; i < LA; ++i) As[ak[i]][am[i]] = ra[i];
#pragma unroll
        for (int i = 0; i < LB; ++i) Bs[bk[i]][bn[i]] = rb[i];
        __syncthreads();
        if (k0 + SG_BK < d.K) {
#pragma unroll
            for (int i = 0; i < LA; ++i) ra[i] = ldA(k0 + SG_BK, i);
#pragma unroll
            for (int i = 0; i < LB; ++i) rb[i] = ldB(k0 + SG_BK, i);
        }
#pragma unroll
        for (int k = 0; k < SG_BK; ++k) {
            float av[TM], bv[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) av[i] = As[k][TM * ty + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) bv[j] = Bs[k][TN * tx + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + TM * ty + i;
        if (m >= d.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + TN * tx + j;
            if (n >= d.N) continue;
            float v = acc[i][j];
            if (d.u && d.v) v = fmaf(d.u[m], d.v[n], v);
            if (d.C) {
                float* o = d.C + (size_t)m * d.ldc + n;
                v = d.accumulate ? *o + v : v;
                *o = v;
            }
            if (d.Cb) reinterpret_cast<bf16*>(d.Cb)[(size_t)m * d.ldcb + n] = __float2bfloat16_rn(v);
            if (d.CbT) reinterpret_cast<bf16*>(d.CbT)[(size_t)n * d.ldcbT + m] = __float2bfloat16_rn(v);
        }
    }
}
}  // namespace rpg

extern "C" int rpg_sgemm_batch(const rpg_sgemm_batch_t* batch, rpg_stream_t stream) {
    if (!batch || batch->n < 1 || batch->n > RPG_SGEMM_BATCH_MAX) return set_error(RPG_E_ARG, "sgemm_batch: bad arguments");
    int mx = 1, nx = 1;
    for (int i = 0; i < batch->n; ++i) {
        const rpg_sgemm_desc_t& d = batch->d[i];
        if (!d.A || !d.B || (!d.C && !d.Cb && !d.CbT) || d.M <= 0 || d.N <= 0 || d.K <= 0)
            return set_error(RPG_E_ARG, "sgemm_batch: bad descriptor");
        mx = d.M > mx ? d.M : mx;
        nx = d.N > nx ? d.N : nx;
    }
    // these products are small: with 64 x 64 tiles most SMs would idle behind a few long k loops, so launches of fewer than
    // two waves of big tiles take 32 x 32 tiles (4x the blocks, a quarter of the work per k chunk each)
    long long blocks64 = 0;
    for (int i = 0; i < batch->n; ++i) blocks64 += (long long)((batch->d[i].M + 63) / 64) * ((batch->d[i].N + 63) / 64);
    if (blocks64 < 2 * 148)
        launch_pdl(sgemm_batch_kernel<32, 32>, dim3((nx + 31) / 32, (mx + 31) / 32, batch->n), dim3(256), 0, as_stream(stream), *batch);
    else
        launch_pdl(sgemm_batch_kernel<64, 64>, dim3((nx + 63) / 64, (mx + 63) / 64, batch->n), dim3(256), 0, as_stream(stream), *batch);
    return check_launch("sgemm_batch_kernel");
}
