// Bandwidth-bound kernels of the RelPose-GNN path: graph validation, weight packing, channel attention
// (MUFU-bound), deterministic segment reductions over the per-graph edge template, edge-feature initialiser,
// dropout + pose heads, pose loss, column sums.  All accesses are 16-byte vectorised and coalesced along the
// feature dimension; no atomics on floating-point data (fixed summation order => bitwise reproducible).
#include <algorithm>
#include <mutex>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/rpg.h"
#include "rpg_internal.h"
#include "rpg_ptx.cuh"

namespace rpg {

typedef __nv_bfloat16 bf16;

// Raises a kernel's dynamic shared-memory limit once per size class.  Called from the forward thread and from autograd's
// backward thread: the cached limit is guarded by a mutex.
static std::mutex g_smem_mu;
struct SmemLimit { size_t bytes[64]; SmemLimit() { for (auto& b : bytes) b = 48 * 1024; } };
template <typename K>
static void ensure_dyn_smem(K kern, size_t bytes, SmemLimit* configured) {
    int dev = 0;
    cudaGetDevice(&dev);                          // the opt-in is a per-device attribute of the function
    if (dev < 0 || dev >= 64) dev = 0;
    std::lock_guard<std::mutex> lk(g_smem_mu);
    if (bytes > configured->bytes[dev]) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        configured->bytes[dev] = bytes;
    }
}

static inline int grid_for(long long work, int block, int cap = 148 * 16) {
    long long g = (work + block - 1) / block;
    if (g < 1) g = 1;
    return (int)(g > cap ? cap : g);
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    return u;
}

// ------------------------------------------------------------------------------------------------
// edge_index validation (replaces PyG's generic gather/scatter indexing with an implicit template)
// ------------------------------------------------------------------------------------------------
__global__ void validate_edge_index_kernel(const long long* __restrict__ ei, long long Et, int G, int N, int Ep,
                                           int* __restrict__ tsrc, int* __restrict__ tdst, int* __restrict__ bad) {
    pdl_prologue();
    int local_bad = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < Et; i += (long long)gridDim.x * blockDim.x) {
        const long long g = i / Ep;
        const int k = (int)(i - g * Ep);
        const long long s = ei[i], d = ei[Et + i];
        const long long ts = ei[k], td = ei[Et + k];
        if (ts < 0 || ts >= N || td < 0 || td >= N || s != ts + g * N || d != td + g * N) ++local_bad;
        if (g == 0) { tsrc[k] = (int)ts; tdst[k] = (int)td; }
    }
    if (local_bad) atomicAdd(bad, local_bad);   // integer atomic: result is order-independent
}

// One-hot selection tiles for the K-panel gathers (rpg_gemm_t.gsel): tile p, row i selects node column
// ((p*g+i) / Ep) * N + endpoint[(p*g+i) % Ep].  One thread per (tile row, 8 columns).
__global__ void selection_patterns_kernel(const int* __restrict__ endpoint, int Ep, int N, int g, int npat,
                                          bf16* __restrict__ sel) {
    pdl_prologue();
    const int total = npat * 128 * 8;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int row = t >> 3, c8 = (t & 7) << 3;
        const int p = row >> 7, i = row & 127;
        const int le = p * g + i;
        const int col = (le / Ep) * N + __ldg(endpoint + le % Ep);
        uint4 u = make_uint4(0, 0, 0, 0);
        if (col >= c8 && col < c8 + 8) {
            uint32_t* w = reinterpret_cast<uint32_t*>(&u);
            w[(col - c8) >> 1] = ((col - c8) & 1) ? 0x3F800000u : 0x00003F80u;     // bf16 1.0 = 0x3F80
        }
        *reinterpret_cast<uint4*>(sel + (size_t)row * 64 + c8) = u;
    }
}

// ------------------------------------------------------------------------------------------------
// weights / casts
// ------------------------------------------------------------------------------------------------
// Per-row-block variant for per-graph edge sets: no period, one tile per 128 edge rows.
__global__ void selection_patterns_rows_kernel(const int* __restrict__ endpoint, long long Et, int pg_Ep, int pg_N,
                                               bf16* __restrict__ sel, int* __restrict__ bad) {
    pdl_prologue();
    const long long rows = (Et + 127) / 128 * 128;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < rows * 8; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t >> 3;
        const int cg = (int)(t & 7);
        int col = -1;
        if (r < Et) {
            const long long win = ((r / 128 * 128) / pg_Ep) * pg_N;
            col = (int)(__ldg(endpoint + r) - win);
            if ((col < 0 || col >= 64) && cg == 0) atomicAdd(bad, 1);
        }
        float f[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) f[q] = (cg * 8 + q == col) ? 1.f : 0.f;
        *reinterpret_cast<uint4*>(sel + r * 64 + cg * 8) = pack8(f);
    }
}

__global__ void pack_weight_kernel(const float* __restrict__ src, int ld_src, int r0, int c0, int rows, int cols,
                                   bf16* __restrict__ dst, int ld_dst, int transpose) {
    pdl_prologue();
    __shared__ float tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;   // bx: column block, by: row block (source window coords)
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = by + j, c = bx + threadIdx.x;
        tile[j][threadIdx.x] = (r < rows && c < cols) ? src[(size_t)(r0 + r) * ld_src + c0 + c] : 0.f;
    }
    __syncthreads();
    if (!transpose) {
        for (int j = threadIdx.y; j < 32; j += blockDim.y) {
            const int r = by + j, c = bx + threadIdx.x;
            if (r < rows && c < cols) dst[(size_t)r * ld_dst + c] = __float2bfloat16_rn(tile[j][threadIdx.x]);
        }
    } else {
        for (int j = threadIdx.y; j < 32; j += blockDim.y) {
            const int c = bx + j, r = by + threadIdx.x;     // dst[c, r]
            if (r < rows && c < cols) dst[(size_t)c * ld_dst + r] = __float2bfloat16_rn(tile[threadIdx.x][j]);
        }
    }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
    pdl_prologue();
    const long long stride = (long long)gridDim.x * blockDim.x * 8;
    for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
        if (i + 8 <= n) {
            const float4 a = *reinterpret_cast<const float4*>(src + i);
            const float4 b = *reinterpret_cast<const float4*>(src + i + 4);
            const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            *reinterpret_cast<uint4*>(dst + i) = pack8(f);
        } else {
            for (long long j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
        }
    }
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, long long n) {
    pdl_prologue();
    const long long stride = (long long)gridDim.x * blockDim.x * 8;
    for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
        if (i + 8 <= n) {
            float f[8];
            unpack8(*reinterpret_cast<const uint4*>(src + i), f);
            *reinterpret_cast<float4*>(dst + i) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(dst + i + 4) = make_float4(f[4], f[5], f[6], f[7]);
        } else {
            for (long long j = i; j < n; ++j) dst[j] = __bfloat162float(src[j]);
        }
    }
}

// fp32 mode: v -> (hi, lo) = (bf16(v), bf16(v - hi)) and back
__global__ void cast_f32_split_kernel(const float* __restrict__ src, bf16* __restrict__ hi, bf16* __restrict__ lo, long long n) {
    pdl_prologue();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = src[i];
        const bf16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}
__global__ void split_to_f32_kernel(const bf16* __restrict__ hi, const bf16* __restrict__ lo, float* __restrict__ dst, long long n) {
    pdl_prologue();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = __bfloat162float(hi[i]) + __bfloat162float(lo[i]);
}
// window of an fp32 weight -> low plane bf16(w - float(bf16(w)))   (the high plane is pack_weight_kernel's output)
__global__ void pack_weight_lo_kernel(const float* __restrict__ src, int ld_src, int r0, int c0, int rows, int cols,
                                      bf16* __restrict__ dst, int ld_dst) {
    pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c < cols && r < rows) {
        const float v = src[(size_t)(r0 + r) * ld_src + c0 + c];
        dst[(size_t)r * ld_dst + c] = __float2bfloat16_rn(v - __bfloat162float(__float2bfloat16_rn(v)));
    }
}

// ------------------------------------------------------------------------------------------------
// Channel attention (att.py:25-30).  Logits are rank-1 (phi_i * theta_j), so the row maximum is
// phi_i * max_j(theta) or phi_i * min_j(theta): no c x c tensor is ever stored.  One warp per edge row.
// ------------------------------------------------------------------------------------------------
constexpr int ATT_WARPS = 8;
constexpr float LOG2E = 1.4426950408889634f;

// The attention kernels are bound by the MUFU.EX2 pipe (c^2 exps per edge row, 16/clk/SM).  Everything around the exps
// is packed fp32x2 math (FFMA2/FADD2, sm_100+): a lane owns two rows (forward) or two columns (backward sweep 2) of the
// c x c score matrix as one 64-bit register pair and the broadcast operand is a scalar, so a pair of exps costs
// 5-7 issue slots and 3-5 FMA-pipe instructions against 16 MUFU cycles.  (Scalar FFMA issues every other cycle on this
// part, which made the scalar form FMA-pipe bound.)  Tried and measured neutral or worse: moving a quarter of the exps
// to an FMA-pipe polynomial (scalar form: neutral; packed Cody-Waite form with FFMA2: forward 205 -> 228 us, the clamp /
// exponent-splice ALU work and issue slots cost more than the MUFU cycles they free) and the packed ex2.approx.f16x2 /
// bf16x2 forms (ptxas lowers them to two scalar MUFU.EX2).
// AUX: also store the per-row statistics [1/den | y | sum_j S_ij theta_j | sum_j S_ij g_j theta_j] (fp32 [Et, 4c]) that
// let the backward skip its first sweep.
__device__ __forceinline__ float2 ex2_pair(float2 a) { return make_float2(exp2f(a.x), exp2f(a.y)); }
__device__ __forceinline__ float2 bcast2(float v) { return make_float2(v, v); }

template <int C_MAX, bool AUX>
__global__ void __launch_bounds__(ATT_WARPS * 32)
attention_fwd_kernel(const float* __restrict__ gtp, long long Et, int c, bf16* __restrict__ y, int ldy,
                     bf16* __restrict__ y_lo, float* __restrict__ aux) {
    pdl_prologue();
    __shared__ __align__(16) float s_g[ATT_WARPS][C_MAX];
    __shared__ __align__(16) float s_t[ATT_WARPS][C_MAX];
    __shared__ __align__(16) float s_gt[AUX ? ATT_WARPS : 1][AUX ? C_MAX : 4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sg = s_g[warp];
    float* st = s_t[warp];
    float* sgt = s_gt[AUX ? warp : 0];
    for (long long row = blockIdx.x * (long long)ATT_WARPS + warp; row < Et; row += (long long)gridDim.x * ATT_WARPS) {
        const float* r = gtp + row * 3 * c;
        float tmax = -INFINITY, tmin = INFINITY;
        for (int j = lane; j < c; j += 32) {
            const float g = r[j], t = r[c + j];
            sg[j] = g; st[j] = t;
            if (AUX) sgt[j] = g * t;
            tmax = fmaxf(tmax, t); tmin = fminf(tmin, t);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
            tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
        }
        __syncwarp();
        for (int i0 = 2 * lane; i0 < c; i0 += 64) {
            const float2 ph = *reinterpret_cast<const float2*>(r + 2 * c + i0);
            const float2 P = make_float2(ph.x * LOG2E, ph.y * LOG2E);
            const float2 nM = make_float2(-(P.x >= 0.f ? P.x * tmax : P.x * tmin), -(P.y >= 0.f ? P.y * tmax : P.y * tmin));
            float2 num = make_float2(0.f, 0.f), den = num, at = num, agt = num;
#pragma unroll 4
            for (int j = 0; j < c; j += 4) {
                const float4 t4 = *reinterpret_cast<const float4*>(st + j);
                const float4 g4 = *reinterpret_cast<const float4*>(sg + j);
                float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (AUX) x4 = *reinterpret_cast<const float4*>(sgt + j);
                const float tt[4] = {t4.x, t4.y, t4.z, t4.w};
                const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
                const float xx[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 e = ex2_pair(__ffma2_rn(P, bcast2(tt[q]), nM));
                    num = __ffma2_rn(e, bcast2(gg[q]), num);
                    den = __fadd2_rn(den, e);
                    if (AUX) {
                        at = __ffma2_rn(e, bcast2(tt[q]), at);
                        agt = __ffma2_rn(e, bcast2(xx[q]), agt);
                    }
                }
            }
            const float y0 = num.x / den.x, y1 = num.y / den.y;
            if (AUX) {
                const float inv0 = 1.f / den.x, inv1 = 1.f / den.y;
                float* a = aux + row * 4 * c;
                *reinterpret_cast<float2*>(a + i0) = make_float2(inv0, inv1);
                *reinterpret_cast<float2*>(a + c + i0) = make_float2(y0, y1);
                *reinterpret_cast<float2*>(a + 2 * c + i0) = make_float2(at.x * inv0, at.y * inv1);
                *reinterpret_cast<float2*>(a + 3 * c + i0) = make_float2(agt.x * inv0, agt.y * inv1);
            }
            *reinterpret_cast<uint32_t*>(y + row * ldy + i0) = pack_bf16x2(y0, y1);
            if (y_lo)
                *reinterpret_cast<uint32_t*>(y_lo + row * ldy + i0) =
                    pack_bf16x2(y0 - __bfloat162float(__float2bfloat16_rn(y0)), y1 - __bfloat162float(__float2bfloat16_rn(y1)));
        }
        __syncwarp();
    }
}

// Backward, nothing of size c x c is stored:
//  sweep 1 (lane owns i): den_i, y_i and  dphi_i = w_i (sum_j p_ij g_j theta_j - y_i sum_j p_ij theta_j),  w_i = dy_i / den_i
//          -- read from the forward's saved statistics when `aux` is given, recomputed (c^2 more exps) otherwise
//  sweep 2 (lane owns j): p_ij recomputed;  dg_j = sum_i w_i p_ij,  dtheta_j = g_j sum_i w_i phi_i p_ij - sum_i w_i y_i phi_i p_ij
// with p_ij = exp2(phi_i log2e theta_j - m_i), m_i the rank-1 row maximum.  One warp per edge row.
__global__ void __launch_bounds__(ATT_WARPS * 32)
attention_bwd_kernel(const float* __restrict__ gtp, const float* __restrict__ dyn, int ld_dyn,
                     const int* __restrict__ tdst, int Ep, int Nn, long long Et, int c,
                     bf16* __restrict__ dgtp, int ld_dgtp, const float* __restrict__ aux) {
    pdl_prologue();
    extern __shared__ __align__(16) float att_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* base = att_smem + (size_t)warp * 8 * c;
    float* sg = base;             // g_j
    float* st = sg + c;           // theta_j
    float* sgt = st + c;          // g_j * theta_j
    float* sp = sgt + c;          // phi_i * log2e
    float* sm = sp + c;           // -(row maximum) (log2 domain)
    float* sw = sm + c;           // w_i
    float* swp = sw + c;          // w_i * phi_i
    float* swyp = swp + c;        // w_i * y_i * phi_i
    for (long long row = blockIdx.x * (long long)ATT_WARPS + warp; row < Et; row += (long long)gridDim.x * ATT_WARPS) {
        const float* r = gtp + row * 3 * c;
        const long long gi = row / Ep;
        const int k = (int)(row - gi * Ep);
        const float* dy = dyn + (gi * Nn + __ldg(tdst + k)) * ld_dyn;
        float tmax = -INFINITY, tmin = INFINITY;
        for (int j = lane; j < c; j += 32) {
            const float g = r[j], t = r[c + j];
            sg[j] = g; st[j] = t; sgt[j] = g * t;
            tmax = fmaxf(tmax, t); tmin = fminf(tmin, t);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
            tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
        }
        __syncwarp();
        // ---- sweep 1 from the forward's saved statistics (no exps): [1/den | y | sum S theta | sum S g theta]
        if (aux) {
            const float* a = aux + row * 4 * c;
            for (int i = lane; i < c; i += 32) {
                const float phi = r[2 * c + i];
                const float pl = phi * LOG2E;
                const float w = dy[i] * a[i], yi = a[c + i];
                sp[i] = pl; sm[i] = -(pl >= 0.f ? pl * tmax : pl * tmin);
                sw[i] = w; swp[i] = w * phi; swyp[i] = w * yi * phi;
                dgtp[row * ld_dgtp + 2 * c + i] = __float2bfloat16_rn(dy[i] * (a[3 * c + i] - yi * a[2 * c + i]));
            }
        }
        // ---- sweep 1 (recompute): a lane owns rows i0 = ib + lane and i1 = i0 + 32 as one fp32x2 pair
        for (int ib = 0; ib < (aux ? 0 : c); ib += 64) {
            const int i0 = ib + lane, i1 = i0 + 32;
            const bool v0 = i0 < c, v1 = i1 < c;
            const float phi0 = v0 ? r[2 * c + i0] : 0.f, phi1 = v1 ? r[2 * c + i1] : 0.f;
            const float2 P = make_float2(phi0 * LOG2E, phi1 * LOG2E);
            const float2 nM = make_float2(-(P.x >= 0.f ? P.x * tmax : P.x * tmin), -(P.y >= 0.f ? P.y * tmax : P.y * tmin));
            float2 den = make_float2(0.f, 0.f), num = den, at = den, agt = den;
#pragma unroll 2
            for (int j = 0; j < c; j += 4) {
                const float4 t4 = *reinterpret_cast<const float4*>(st + j);
                const float4 g4 = *reinterpret_cast<const float4*>(sg + j);
                const float4 x4 = *reinterpret_cast<const float4*>(sgt + j);
                const float tt[4] = {t4.x, t4.y, t4.z, t4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w}, xx[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float2 e = ex2_pair(__ffma2_rn(P, bcast2(tt[u]), nM));
                    den = __fadd2_rn(den, e);
                    num = __ffma2_rn(e, bcast2(gg[u]), num);
                    at = __ffma2_rn(e, bcast2(tt[u]), at);
                    agt = __ffma2_rn(e, bcast2(xx[u]), agt);
                }
            }
            if (v0) {
                const float inv = 1.f / den.x, y = num.x * inv, w = dy[i0] * inv;
                sp[i0] = P.x; sm[i0] = nM.x; sw[i0] = w; swp[i0] = w * phi0; swyp[i0] = w * y * phi0;
                dgtp[row * ld_dgtp + 2 * c + i0] = __float2bfloat16_rn(w * (agt.x - y * at.x));
            }
            if (v1) {
                const float inv = 1.f / den.y, y = num.y * inv, w = dy[i1] * inv;
                sp[i1] = P.y; sm[i1] = nM.y; sw[i1] = w; swp[i1] = w * phi1; swyp[i1] = w * y * phi1;
                dgtp[row * ld_dgtp + 2 * c + i1] = __float2bfloat16_rn(w * (agt.y - y * at.y));
            }
        }
        __syncwarp();
        // ---- sweep 2: a lane owns columns j0 = jb + lane and j1 = j0 + 32 as one fp32x2 pair
        for (int jb = 0; jb < c; jb += 64) {
            const int j0 = jb + lane, j1 = j0 + 32;
            const bool v0 = j0 < c, v1 = j1 < c;
            const float2 TH = make_float2(v0 ? st[j0] : 0.f, v1 ? st[j1] : 0.f);
            float2 dg = make_float2(0.f, 0.f), sa = dg, sb = dg;
#pragma unroll 2
            for (int i = 0; i < c; i += 4) {
                const float4 p4 = *reinterpret_cast<const float4*>(sp + i);
                const float4 m4 = *reinterpret_cast<const float4*>(sm + i);
                const float4 w4 = *reinterpret_cast<const float4*>(sw + i);
                const float4 q4 = *reinterpret_cast<const float4*>(swp + i);
                const float4 z4 = *reinterpret_cast<const float4*>(swyp + i);
                const float pp[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w};
                const float ww[4] = {w4.x, w4.y, w4.z, w4.w}, qq[4] = {q4.x, q4.y, q4.z, q4.w}, zz[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float2 e = ex2_pair(__ffma2_rn(TH, bcast2(pp[u]), bcast2(mm[u])));
                    dg = __ffma2_rn(e, bcast2(ww[u]), dg);
                    sa = __ffma2_rn(e, bcast2(qq[u]), sa);
                    sb = __ffma2_rn(e, bcast2(zz[u]), sb);
                }
            }
            if (v0) {
                dgtp[row * ld_dgtp + j0] = __float2bfloat16_rn(dg.x);
                dgtp[row * ld_dgtp + c + j0] = __float2bfloat16_rn(sg[j0] * sa.x - sb.x);
            }
            if (v1) {
                dgtp[row * ld_dgtp + j1] = __float2bfloat16_rn(dg.y);
                dgtp[row * ld_dgtp + c + j1] = __float2bfloat16_rn(sg[j1] * sa.y - sb.y);
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Segment reductions over the per-graph template (mean aggregation, dgrad scatters): gather formulation,
// fixed order, one thread per 8 feature columns.  (Round 2 measured the alternative -- whole graph blocks bulk-copied
// into shared memory, several in flight per SM, sums gathered from there: same DRAM bytes, same 38-40 us / 56-58 us
// per launch under ncu on the 155 648 x 512 tensor, so the extra kernel was dropped)
//   out[g*N+n, c] = scale[n] * sum_{k in csr(n)} v[g*Ep+k, c] * (mask ? mask[g*Ep+k, c] > 0 : 1)
// ------------------------------------------------------------------------------------------------
// PLAIN: no ReLU mask and no fp32-mode lo plane (every launch of the bf16 path) -- a third of the registers, so more
// rows are in flight per SM.
template <typename I, bool PLAIN>   // I: index type of the flattened (row, column-group) space: 32-bit whenever it fits
__global__ void segment_sum_kernel(const bf16* __restrict__ v, int ldv, const bf16* __restrict__ mask_, int ldm,
                                   const int* __restrict__ ptr, const int* __restrict__ idx,
                                   const float* __restrict__ scale, long long Nt, int N, int Ep, int D,
                                   bf16* __restrict__ out, int ldo, const bf16* __restrict__ v_lo_,
                                   bf16* __restrict__ out_lo_) {
    pdl_prologue();
    const bf16* __restrict__ mask = PLAIN ? nullptr : mask_;
    const bf16* __restrict__ v_lo = PLAIN ? nullptr : v_lo_;
    bf16* __restrict__ out_lo = PLAIN ? nullptr : out_lo_;
    const int tpr = D >> 3;                                   // threads per row
    const I total = (I)Nt * (I)tpr;
    for (I t = (I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x; t < total; t += (I)gridDim.x * (I)blockDim.x) {
        const I row_i = t / (I)tpr;
        const int c = (int)(t - row_i * (I)tpr) << 3;
        const I g_i = row_i / (I)N;
        const int n = (int)(row_i - g_i * (I)N);
        const long long row = (long long)row_i, g = (long long)g_i;
        const int k0 = __ldg(ptr + n), k1 = __ldg(ptr + n + 1);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int kb = k0; kb < k1; kb += 4) {                 // 4 independent row loads in flight per thread
            uint4 u[4], um[4], ul[4];
            bool ok[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ok[j] = kb + j < k1;
                const long long er = g * Ep + (ok[j] ? __ldg(idx + kb + j) : 0);
                u[j] = ok[j] ? __ldg(reinterpret_cast<const uint4*>(v + er * ldv + c)) : make_uint4(0, 0, 0, 0);
                if (mask) um[j] = ok[j] ? __ldg(reinterpret_cast<const uint4*>(mask + er * ldm + c)) : make_uint4(0, 0, 0, 0);
                if (v_lo) ul[j] = ok[j] ? __ldg(reinterpret_cast<const uint4*>(v_lo + er * ldv + c)) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {                     // fixed summation order: deterministic
                float f[8];
                unpack8(u[j], f);
                if (mask) {
                    float mf[8];
                    unpack8(um[j], mf);
#pragma unroll
                    for (int q = 0; q < 8; ++q) if (!(mf[q] > 0.f)) f[q] = 0.f;
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] += f[q];
                if (v_lo) {                                   // fp32 mode: value = hi + lo
                    unpack8(ul[j], f);
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc[q] += f[q];
                }
            }
        }
        if (scale) {
            const float s = __ldg(scale + n);
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] *= s;
        }
        *reinterpret_cast<uint4*>(out + row * ldo + c) = pack8(acc);
        if (out_lo) {
            float r[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) r[q] = acc[q] - __bfloat162float(__float2bfloat16_rn(acc[q]));
            *reinterpret_cast<uint4*>(out_lo + row * ldo + c) = pack8(r);
        }
    }
}

// Two segment sums of the SAME edge tensor in one launch (by source and by destination of dh1; by lower and by upper
// endpoint of the edge-initialiser gradient): a thread owns 8 columns of one node row of ONE of the two CSRs; the two
// CSR passes of a node row sit next to each other in the flattened index space, so a graph's edge rows are read by
// both passes within a few hundred cycles: the second read is an L2 hit and DRAM sees the tensor once.  Fixed order.
template <typename I>
__global__ void segment_sum2_kernel(const bf16* __restrict__ v, int ldv, const int* __restrict__ ptr_a,
                                    const int* __restrict__ idx_a, const int* __restrict__ ptr_b,
                                    const int* __restrict__ idx_b, long long Nt, int N, int Ep, int D,
                                    bf16* __restrict__ out_a, int ldo_a, bf16* __restrict__ out_b, int ldo_b) {
    pdl_prologue();
    const int tpr = D >> 3;
    const I total = (I)Nt * (I)tpr * (I)2;
    for (I t = (I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x; t < total; t += (I)gridDim.x * (I)blockDim.x) {
        const I unit = t / (I)tpr;                             // (node row, pass)
        const int c = (int)(t - unit * (I)tpr) << 3;
        const I row_i = unit >> 1;
        const int pass = (int)(unit & 1);
        const I g_i = row_i / (I)N;
        const int n = (int)(row_i - g_i * (I)N);
        const long long row = (long long)row_i, g = (long long)g_i;
        const int* __restrict__ ptr = pass ? ptr_b : ptr_a;
        const int* __restrict__ idx = pass ? idx_b : idx_a;
        const int k0 = __ldg(ptr + n), k1 = __ldg(ptr + n + 1);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int kb = k0; kb < k1; kb += 4) {
            uint4 u[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool ok = kb + j < k1;
                const long long er = g * Ep + (ok ? __ldg(idx + kb + j) : 0);
                u[j] = ok ? __ldg(reinterpret_cast<const uint4*>(v + er * ldv + c)) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float f[8];
                unpack8(u[j], f);
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] += f[q];
            }
        }
        bf16* o = pass ? out_b + row * ldo_b : out_a + row * ldo_a;
        *reinterpret_cast<uint4*>(o + c) = pack8(acc);
    }
}

__global__ void scale_rows_kernel(const bf16* __restrict__ v, int ldv, long long rows, int D, const float* __restrict__ scale,
                                  int mod, bf16* __restrict__ out, int ldo) {
    pdl_prologue();
    const int tpr = D >> 3;
    const long long total = rows * tpr;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long row = t / tpr;
        const int c = (int)(t - row * tpr) << 3;
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(v + row * ldv + c)), f);
        const float s = __ldg(scale + (int)(row % mod));
#pragma unroll
        for (int q = 0; q < 8; ++q) f[q] *= s;
        *reinterpret_cast<uint4*>(out + row * ldo + c) = pack8(f);
    }
}

// e0 = relu(pmin[node(min(s,t))] + pmax[node(max(s,t))] + bias)      (posenet.py:1014-1017, 1053-1055)
template <typename I>
__global__ void edge_init_fwd_kernel(const bf16* __restrict__ pmm, int ldp, const float* __restrict__ bias,
                                     const int* __restrict__ tsrc, const int* __restrict__ tdst, long long Et, int N,
                                     int Ep, int D, bf16* __restrict__ e0, int lde, uint8_t* __restrict__ bits) {
    pdl_prologue();
    const int tpr = D >> 3;
    const I total = (I)Et * (I)tpr;
    for (I t = (I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x; t < total; t += (I)gridDim.x * (I)blockDim.x) {
        const I row_i = t / (I)tpr;
        const int c = (int)(t - row_i * (I)tpr) << 3;
        const I g_i = row_i / (I)Ep;
        const int k = (int)(row_i - g_i * (I)Ep);
        const long long row = (long long)row_i, g = (long long)g_i;
        const int s = __ldg(tsrc + k), d = __ldg(tdst + k);
        const long long nlo = g * N + min(s, d), nhi = g * N + max(s, d);
        float a[8], b[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(pmm + nlo * ldp + c)), a);
        unpack8(__ldg(reinterpret_cast<const uint4*>(pmm + nhi * ldp + D + c)), b);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c + 4));
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] = fmaxf(a[q] + b[q] + bb[q], 0.f);
        *reinterpret_cast<uint4*>(e0 + row * lde + c) = pack8(a);
        if (bits) {
            uint32_t m = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) m |= (uint32_t)(a[q] > 0.f) << q;
            bits[row * (D >> 3) + (c >> 3)] = (uint8_t)m;
        }
    }
}

// Edge-dropout mask applied to a per-edge tensor (train.py:242-247 masks data.edge_attr with the tiled mask): rows of
// the kept template slots, graph by graph:  out[g * Ep_kept + j, :] = in[g * Ep_full + kept_idx[j], :]   (any 4-byte type)
__global__ void edge_rows_select_kernel(const uint32_t* __restrict__ in, long long G, int Ep_full, int Ep_kept,
                                        const int* __restrict__ kept_idx, int words, uint32_t* __restrict__ out) {
    pdl_prologue();
    const long long total = G * Ep_kept * (long long)words;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long row = t / words;
        const int w = (int)(t - row * words);
        const long long g = row / Ep_kept;
        const int j = (int)(row - g * Ep_kept);
        out[t] = in[(g * Ep_full + __ldg(kept_idx + j)) * words + w];
    }
}

// General per-edge gather of node rows through the template (sibling layers without an edge-feature GEMM):
//   out[e] = act( pa[node_a(e)] + pb[node_b(e)] + bias ) * bit(e)      node_x = source (0) or destination (1)
// pb, bias, mask_bits optional; act = ReLU or identity; optional pattern output.
__global__ void edge_gather_kernel(const bf16* __restrict__ pa, int lda, int which_a, const bf16* __restrict__ pb, int ldb,
                                   int which_b, const float* __restrict__ bias, const int* __restrict__ tsrc,
                                   const int* __restrict__ tdst, long long Et, int N, int Ep, int D, int relu,
                                   const uint8_t* __restrict__ mask_bits, bf16* __restrict__ out, int ldo,
                                   uint8_t* __restrict__ out_bits) {
    pdl_prologue();
    const int tpr = D >> 3;
    const long long total = Et * tpr;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long row = t / tpr;
        const int c = (int)(t - row * tpr) << 3;
        const long long g = row / Ep;
        const int k = (int)(row - g * Ep);
        const int s = __ldg(tsrc + k), d = __ldg(tdst + k);
        float a[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(pa + (g * N + (which_a ? d : s)) * lda + c)), a);
        if (pb) {
            float b[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(pb + (g * N + (which_b ? d : s)) * ldb + c)), b);
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] += b[q];
        }
        if (bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c + 4));
            a[0] += b0.x; a[1] += b0.y; a[2] += b0.z; a[3] += b0.w; a[4] += b1.x; a[5] += b1.y; a[6] += b1.z; a[7] += b1.w;
        }
        if (relu) {
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] = fmaxf(a[q], 0.f);
        }
        if (mask_bits) {
            const uint32_t m = mask_bits[row * (D >> 3) + (c >> 3)];
#pragma unroll
            for (int q = 0; q < 8; ++q) if (!((m >> q) & 1u)) a[q] = 0.f;
        }
        *reinterpret_cast<uint4*>(out + row * ldo + c) = pack8(a);
        if (out_bits) {
            uint32_t m = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) m |= (uint32_t)(a[q] > 0.f) << q;
            out_bits[row * (D >> 3) + (c >> 3)] = (uint8_t)m;
        }
    }
}

// fp32 mode: pmm fp32 [Nt, 2D] -> e0 (hi, lo)
__global__ void edge_init_fwd_f32_kernel(const float* __restrict__ pmm, int ldp, const float* __restrict__ bias,
                                         const int* __restrict__ tsrc, const int* __restrict__ tdst, long long Et, int N,
                                         int Ep, int D, bf16* __restrict__ e_hi, bf16* __restrict__ e_lo, int lde,
                                         uint8_t* __restrict__ bits) {
    pdl_prologue();
    const int tpr = D >> 2;
    const long long total = Et * tpr;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long row = t / tpr;
        const int c = (int)(t - row * tpr) << 2;
        const long long g = row / Ep;
        const int k = (int)(row - g * Ep);
        const int s = __ldg(tsrc + k), d = __ldg(tdst + k);
        const long long nlo = g * N + min(s, d), nhi = g * N + max(s, d);
        const float4 a = __ldg(reinterpret_cast<const float4*>(pmm + nlo * ldp + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(pmm + nhi * ldp + D + c));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c));
        const float v[4] = {fmaxf(a.x + b.x + bb.x, 0.f), fmaxf(a.y + b.y + bb.y, 0.f), fmaxf(a.z + b.z + bb.z, 0.f),
                            fmaxf(a.w + b.w + bb.w, 0.f)};
        float r[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) r[q] = v[q] - __bfloat162float(__float2bfloat16_rn(v[q]));
        uint2 h, l;
        h.x = pack_bf16x2(v[0], v[1]); h.y = pack_bf16x2(v[2], v[3]);
        l.x = pack_bf16x2(r[0], r[1]); l.y = pack_bf16x2(r[2], r[3]);
        *reinterpret_cast<uint2*>(e_hi + row * lde + c) = h;
        *reinterpret_cast<uint2*>(e_lo + row * lde + c) = l;
        if (bits) {                                        // 4 pattern bits per thread; the two threads of a byte merge by shuffle
            unsigned m = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) m |= (unsigned)(v[q] > 0.f) << q;
            const unsigned other = __shfl_xor_sync(__activemask(), m, 1);
            if (!(t & 1)) bits[row * (D >> 3) + (c >> 3)] = (uint8_t)(m | (other << 4));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Dropout keep decision: explicit uint8 mask (parity tests) or a counter-based hash of (seed,row,col)
// ------------------------------------------------------------------------------------------------
__global__ void dropout_mask_kernel(unsigned long long seed, uint32_t thresh, long long rows, int D,
                                    uint8_t* __restrict__ keep) {
    pdl_prologue();
    const long long total = rows * D;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / D;
        const int c = (int)(i - r * D);
        keep[i] = keep_from_hash(keep_hash4(seed, r, c >> 2), c, thresh) ? 1 : 0;
    }
}

// pose[r, j] = sum_c drop(feat[r, c]) * w6[j, c] + b6[j]   (posenet.py:1073-1086).  One warp per row; lane owns
// columns [lane*8, +8) of every 256-column pass.
//  * PASSES 1, 2 (D <= 512): the lane's slice of the six weight rows lives in registers as fp32x2 pairs, the products
//    are FFMA2 (two columns per instruction), a warp walks two rows per iteration so four 16-byte loads are in flight
//    per lane, and the six lane-partials are reduced with an 8-shuffle transposing butterfly.
//  * PASSES 0 (wider layers): weights from shared memory, scalar math.
constexpr int HEAD_WARPS = 8;

// Transposing butterfly: v[0..5] per lane -> every lane of quad q holds the full sum of value j(q):
// q = lane >> 2:  0 -> 0, 1 -> 1, 2 -> 2, 4 -> 3, 5 -> 4, 6 -> 5  (quads 3 and 7 hold padding).
__device__ __forceinline__ float head_reduce6(const float* v, int lane) {
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
    float r[4];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float mine = h16 ? v[k + 3] : v[k], theirs = h16 ? v[k] : v[k + 3];
        r[k] = mine + __shfl_xor_sync(0xffffffffu, theirs, 16);
    }
    r[3] = 0.f;
    float s[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float mine = h8 ? r[k + 2] : r[k], theirs = h8 ? r[k] : r[k + 2];
        s[k] = mine + __shfl_xor_sync(0xffffffffu, theirs, 8);
    }
    float t = (h4 ? s[1] : s[0]) + __shfl_xor_sync(0xffffffffu, h4 ? s[0] : s[1], 4);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    return t;
}

// PLAIN: no fp32-mode lo plane and no explicit keep bytes (the production path) -- fewer live registers.
template <int PASSES, bool PLAIN>
__global__ void __launch_bounds__(HEAD_WARPS * 32, PASSES > 0 ? 2 : 1)
head_fwd_kernel(const bf16* __restrict__ feat, int ldf, long long rows, int D, const uint8_t* __restrict__ keep_,
                unsigned long long seed, uint32_t thresh, int use_seed, float scale, const float* __restrict__ w6,
                const float* __restrict__ b6, float* __restrict__ pose, const bf16* __restrict__ feat_lo_) {
    pdl_prologue();
    const uint8_t* __restrict__ keep = PLAIN ? nullptr : keep_;
    const bf16* __restrict__ feat_lo = PLAIN ? nullptr : feat_lo_;
    extern __shared__ __align__(16) float s_w[];   // [6][D], only used when PASSES == 0
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if constexpr (PASSES > 0) {
        constexpr int NP = PASSES > 0 ? PASSES : 1;
        float2 wr[NP][6][4];
#pragma unroll
        for (int ps = 0; ps < NP; ++ps) {
            const int c = ps * 256 + lane * 8;
#pragma unroll
            for (int j = 0; j < 6; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    wr[ps][j][q] = c < D ? __ldg(reinterpret_cast<const float2*>(w6 + j * D + c + 2 * q)) : make_float2(0.f, 0.f);
        }
        const int quad = lane >> 2;
        const int jout = (quad >> 2) * 3 + (quad & 3);
        const bool writer = (lane & 3) == 0 && (quad & 3) != 3;
        const float bj = writer ? b6[jout] : 0.f;
        const long long stride = (long long)gridDim.x * HEAD_WARPS * 2;
        for (long long row0 = (blockIdx.x * (long long)HEAD_WARPS + warp) * 2; row0 < rows; row0 += stride) {
            uint4 u[2][NP], ul[2][NP];
            uint2 k8[2][NP];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const long long row = row0 + rr < rows ? row0 + rr : row0;
#pragma unroll
                for (int ps = 0; ps < NP; ++ps) {
                    const int c = ps * 256 + lane * 8;
                    const bool ok = c < D;
                    u[rr][ps] = ok ? __ldg(reinterpret_cast<const uint4*>(feat + row * ldf + c)) : make_uint4(0, 0, 0, 0);
                    if (feat_lo) ul[rr][ps] = ok ? __ldg(reinterpret_cast<const uint4*>(feat_lo + row * ldf + c)) : make_uint4(0, 0, 0, 0);
                    if (keep) k8[rr][ps] = ok ? __ldg(reinterpret_cast<const uint2*>(keep + row * D + c)) : make_uint2(0, 0);
                }
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const long long row = row0 + rr < rows ? row0 + rr : row0;
                float2 acc[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
                for (int ps = 0; ps < NP; ++ps) {
                    const int c = ps * 256 + lane * 8;
                    float f[8];
                    unpack8(u[rr][ps], f);
                    if (feat_lo) {                                // fp32 mode: value = hi + lo
                        float fl[8];
                        unpack8(ul[rr][ps], fl);
#pragma unroll
                        for (int q = 0; q < 8; ++q) f[q] += fl[q];
                    }
                    if (keep) {
                        const uint32_t kw[2] = {k8[rr][ps].x, k8[rr][ps].y};
#pragma unroll
                        for (int q = 0; q < 8; ++q) if (!((kw[q >> 2] >> ((q & 3) * 8)) & 0xFF)) f[q] = 0.f;
                    } else if (use_seed) {
                        const uint32_t h0 = keep_hash4(seed, row, c >> 2), h1 = keep_hash4(seed, row, (c >> 2) + 1);
#pragma unroll
                        for (int q = 0; q < 8; ++q) if (!keep_from_hash(q < 4 ? h0 : h1, q, thresh)) f[q] = 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 6; ++j)
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            acc[j] = __ffma2_rn(make_float2(f[2 * q], f[2 * q + 1]), wr[ps][j][q], acc[j]);
                }
                float v[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) v[j] = acc[j].x + acc[j].y;
                const float t = head_reduce6(v, lane);
                if (writer && row0 + rr < rows) pose[row * 6 + jout] = fmaf(t, scale, bj);
            }
        }
    } else {
    for (int i = threadIdx.x; i < 6 * D; i += blockDim.x) s_w[i] = w6[i];
    __syncthreads();
    const int n_pass = (D + 255) / 256;
    for (long long row = blockIdx.x * (long long)HEAD_WARPS + warp; row < rows; row += (long long)gridDim.x * HEAD_WARPS) {
        float acc[6] = {0, 0, 0, 0, 0, 0};
        for (int ps = 0; ps < n_pass; ++ps) {
            const int c = ps * 256 + lane * 8;
            if (c >= D) continue;
            float f[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(feat + row * ldf + c)), f);
            if (feat_lo) {
                float fl[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(feat_lo + row * ldf + c)), fl);
#pragma unroll
                for (int q = 0; q < 8; ++q) f[q] += fl[q];
            }
            if (keep) {
                const uint2 k8 = __ldg(reinterpret_cast<const uint2*>(keep + row * D + c));
                const uint32_t kw[2] = {k8.x, k8.y};
#pragma unroll
                for (int q = 0; q < 8; ++q) if (!((kw[q >> 2] >> ((q & 3) * 8)) & 0xFF)) f[q] = 0.f;
            } else if (use_seed) {
                const uint32_t h0 = keep_hash4(seed, row, c >> 2), h1 = keep_hash4(seed, row, (c >> 2) + 1);
#pragma unroll
                for (int q = 0; q < 8; ++q) if (!keep_from_hash(q < 4 ? h0 : h1, q, thresh)) f[q] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const float4 w0 = *reinterpret_cast<const float4*>(s_w + j * D + c);
                const float4 w1 = *reinterpret_cast<const float4*>(s_w + j * D + c + 4);
                acc[j] += f[0] * w0.x + f[1] * w0.y + f[2] * w0.z + f[3] * w0.w + f[4] * w1.x + f[5] * w1.y +
                          f[6] * w1.z + f[7] * w1.w;
            }
        }
        const float t = head_reduce6(acc, lane);
        const int quad = lane >> 2;
        if ((lane & 3) == 0 && (quad & 3) != 3) {
            const int jout = (quad >> 2) * 3 + (quad & 3);
            pose[row * 6 + jout] = fmaf(t, scale, b6[jout]);
        }
    }
    }
}

// Backward of the head.  A block is (row phase) x (column thread); a thread owns 8 columns and walks the rows of its
// phase inside the block's slab, accumulating dW6[j, c] = sum_r dpose[r, j] * drop(feat)[r, c] in registers as fp32x2
// pairs (FFMA2).  Phases are folded through shared memory in a fixed order, slabs through head_bwd_reduce_kernel:
// deterministic.
constexpr int HEADB_THREADS = 256;
constexpr int HEADB_ROWS_PER_BLOCK = 256;
__global__ void __launch_bounds__(HEADB_THREADS, 2)
head_bwd_kernel(const float* __restrict__ dpose, const bf16* __restrict__ feat, int ldf, long long rows,
                int D, const uint8_t* __restrict__ keep, unsigned long long seed, uint32_t thresh,
                int use_seed, float scale, const float* __restrict__ w6, int mask_relu,
                bf16* __restrict__ dfeat, int lddf, float* __restrict__ dw6_part,
                float* __restrict__ db6_part, const bf16* __restrict__ feat_lo, bf16* __restrict__ dfeat_lo) {
    pdl_prologue();
    extern __shared__ __align__(16) float hb_smem[];   // [phases-1][48][ct] + [phases][6]
    const int ct = D >> 3;                              // column threads
    const int phases = HEADB_THREADS / ct;
    const int tx = threadIdx.x % ct, ty = threadIdx.x / ct;
    const int c = tx * 8;
    float2 gw[6][4];
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) gw[j][q] = make_float2(0.f, 0.f);
    float gb[6] = {0, 0, 0, 0, 0, 0};
    if (ty < phases) {
        float2 w[6][4];
#pragma unroll
        for (int j = 0; j < 6; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) w[j][q] = __ldg(reinterpret_cast<const float2*>(w6 + j * D + c + 2 * q));
        const long long r0 = (long long)blockIdx.x * HEADB_ROWS_PER_BLOCK;
        const long long r1 = min(r0 + HEADB_ROWS_PER_BLOCK, rows);
#pragma unroll 2
        for (long long row = r0 + ty; row < r1; row += phases) {
            float dp[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) dp[j] = __ldg(dpose + row * 6 + j);
            float f[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(feat + row * ldf + c)), f);
            if (feat_lo) {                                    // fp32 mode: value = hi + lo
                float fl[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(feat_lo + row * ldf + c)), fl);
#pragma unroll
                for (int q = 0; q < 8; ++q) f[q] += fl[q];
            }
            float km[8];
            if (keep) {
                const uint2 k8 = __ldg(reinterpret_cast<const uint2*>(keep + row * D + c));
                const uint32_t kw[2] = {k8.x, k8.y};
#pragma unroll
                for (int q = 0; q < 8; ++q) km[q] = ((kw[q >> 2] >> ((q & 3) * 8)) & 0xFF) ? scale : 0.f;
            } else if (use_seed) {
                const uint32_t h0 = keep_hash4(seed, row, c >> 2), h1 = keep_hash4(seed, row, (c >> 2) + 1);
#pragma unroll
                for (int q = 0; q < 8; ++q) km[q] = keep_from_hash(q < 4 ? h0 : h1, q, thresh) ? scale : 0.f;
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) km[q] = 1.f;
            }
            float d[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float2 sacc = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 6; ++j) sacc = __ffma2_rn(w[j][q], make_float2(dp[j], dp[j]), sacc);
                const float2 k2 = make_float2(km[2 * q], km[2 * q + 1]);
                const float2 dd = __fmul2_rn(sacc, k2);
                d[2 * q] = (mask_relu && !(f[2 * q] > 0.f)) ? 0.f : dd.x;
                d[2 * q + 1] = (mask_relu && !(f[2 * q + 1] > 0.f)) ? 0.f : dd.y;
                const float2 fd = __fmul2_rn(make_float2(f[2 * q], f[2 * q + 1]), k2);
#pragma unroll
                for (int j = 0; j < 6; ++j) gw[j][q] = __ffma2_rn(fd, make_float2(dp[j], dp[j]), gw[j][q]);
            }
            if (dfeat) *reinterpret_cast<uint4*>(dfeat + row * lddf + c) = pack8(d);
            if (dfeat_lo) {
                float dl[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) dl[q] = d[q] - __bfloat162float(__float2bfloat16_rn(d[q]));
                *reinterpret_cast<uint4*>(dfeat_lo + row * lddf + c) = pack8(dl);
            }
            if (tx == 0) {
#pragma unroll
                for (int j = 0; j < 6; ++j) gb[j] += dp[j];
            }
        }
    }
    // fold the phases (fixed order)
    float* red = hb_smem;
    float* redb = hb_smem + (size_t)(phases - 1) * 48 * ct;
    if (ty >= 1 && ty < phases) {
        float* o = red + (size_t)(ty - 1) * 48 * ct;
#pragma unroll
        for (int j = 0; j < 6; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                o[(j * 8 + 2 * q) * ct + tx] = gw[j][q].x;
                o[(j * 8 + 2 * q + 1) * ct + tx] = gw[j][q].y;
            }
    }
    if (tx == 0 && ty < phases) {
#pragma unroll
        for (int j = 0; j < 6; ++j) redb[ty * 6 + j] = gb[j];
    }
    __syncthreads();
    if (ty == 0) {
        for (int ph = 1; ph < phases; ++ph) {
            const float* o = red + (size_t)(ph - 1) * 48 * ct;
#pragma unroll
            for (int j = 0; j < 6; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    gw[j][q].x += o[(j * 8 + 2 * q) * ct + tx];
                    gw[j][q].y += o[(j * 8 + 2 * q + 1) * ct + tx];
                }
        }
        float* o = dw6_part + (size_t)blockIdx.x * 6 * D;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            *reinterpret_cast<float4*>(o + j * D + c) = make_float4(gw[j][0].x, gw[j][0].y, gw[j][1].x, gw[j][1].y);
            *reinterpret_cast<float4*>(o + j * D + c + 4) = make_float4(gw[j][2].x, gw[j][2].y, gw[j][3].x, gw[j][3].y);
        }
        if (tx < 6) {
            float sb = 0.f;
            for (int ph = 0; ph < phases; ++ph) sb += redb[ph * 6 + tx];
            db6_part[blockIdx.x * 6 + tx] = sb;
        }
    }
}

// Folds the slab partials of head_bwd_kernel: columns [0, 6D) are dW6 (rows 0..2 -> dw_t, 3..5 -> dw_q), the last block
// handles the 6 bias sums.  32 columns x 8 part-phases per block, fixed order.
__global__ void __launch_bounds__(256)
head_bwd_reduce_kernel(const float* __restrict__ dw_part, const float* __restrict__ db_part, int nparts, int D,
                       float* __restrict__ dw_t, float* __restrict__ dw_q, float* __restrict__ db_t,
                       float* __restrict__ db_q, int accumulate) {
    pdl_prologue();
    __shared__ float red[8][32];
    const int lane = threadIdx.x & 31, ph = threadIdx.x >> 5;
    const int n = 6 * D;
    const bool bias_block = blockIdx.x == gridDim.x - 1;
    const int col = bias_block ? lane : blockIdx.x * 32 + lane;
    const int ncol = bias_block ? 6 : n;
    const float* part = bias_block ? db_part : dw_part;
    const long long stride = bias_block ? 6 : n;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (col < ncol) {
        int b = ph;
        for (; b + 24 < nparts; b += 32) {
            s0 += part[(size_t)b * stride + col];
            s1 += part[(size_t)(b + 8) * stride + col];
            s2 += part[(size_t)(b + 16) * stride + col];
            s3 += part[(size_t)(b + 24) * stride + col];
        }
        for (; b < nparts; b += 8) s0 += part[(size_t)b * stride + col];
    }
    red[ph][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (ph == 0 && col < ncol) {
        float s = red[0][lane];
#pragma unroll
        for (int k = 1; k < 8; ++k) s += red[k][lane];
        float* o;
        if (bias_block) o = col < 3 ? db_t + col : db_q + (col - 3);
        else o = col < 3 * D ? dw_t + col : dw_q + (col - 3 * D);
        *o = accumulate ? *o + s : s;
    }
}

// out[i] (+)= sum_b part[b, i]
__global__ void reduce_partials_kernel(const float* __restrict__ part, int nparts, long long stride, long long n,
                                       float* __restrict__ out, int accumulate) {
    pdl_prologue();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int b = 0; b < nparts; ++b) s += part[(size_t)b * stride + i];
        out[i] = accumulate ? out[i] + s : s;
    }
}
// 2-D variant for strided outputs: out[r*ldo + c] (+)= sum_s part[s*stride + r*cols + c]
__global__ void reduce_splits_kernel(const float* __restrict__ part, int splits, long long stride, int rows, int cols,
                                     float* __restrict__ out, int ldo, int accumulate) {
    pdl_prologue();
    const long long n = (long long)rows * cols;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
        float s = 0.f;
        for (int b = 0; b < splits; ++b) s += part[(size_t)b * stride + i];
        float* o = out + (size_t)r * ldo + c;
        *o = accumulate ? *o + s : s;
    }
}

// ------------------------------------------------------------------------------------------------
// Evaluation composition (test.py:227-243) and qexp (pose_utils.py:340-348): one thread per graph / per vector.
// Double arithmetic: G threads of work, and the reference's own numbers are numpy on the host.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void qexp_d(const double* v, float* q) {
    const double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const double sc = n > 0.0 ? sin(n) / n : 1.0;          // np.sinc(n / pi)
    q[0] = (float)cos(n); q[1] = (float)(sc * v[0]); q[2] = (float)(sc * v[1]); q[3] = (float)(sc * v[2]);
}
__global__ void qexp_kernel(const float* __restrict__ v, long long n, float* __restrict__ q) {
    pdl_prologue();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double d[3] = {v[i * 3], v[i * 3 + 1], v[i * 3 + 2]};
        qexp_d(d, q + i * 4);
    }
}
struct EvalNorm { float m[3], s[3]; };
__global__ void eval_compose_kernel(const float* __restrict__ pred_edges, const float* __restrict__ poses,
                                    const int* __restrict__ tsrc, int G, int N, int Ep, int ref_k, EvalNorm nm,
                                    float* __restrict__ out_pred, float* __restrict__ out_targ) {
    pdl_prologue();
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < G; g += gridDim.x * blockDim.x) {
        const float* ap = poses + ((long long)g * N + __ldg(tsrc + ref_k)) * 6;      // absolute pose of the reference image
        const float* rp = pred_edges + ((long long)g * Ep + ref_k) * 6;              // predicted RP = p[src] - p[query]
        const float* tq = poses + (long long)g * N * 6;                              // ground truth of the query (node 0)
        double o[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) o[j] = (double)ap[j] - (double)rp[j];
        float* op = out_pred + (long long)g * 7;
#pragma unroll
        for (int j = 0; j < 3; ++j) op[j] = (float)(o[j] * nm.s[j] + nm.m[j]);
        qexp_d(o + 3, op + 3);
        if (out_targ) {
            float* ot = out_targ + (long long)g * 7;
            const double t[3] = {tq[3], tq[4], tq[5]};
#pragma unroll
            for (int j = 0; j < 3; ++j) ot[j] = (float)((double)tq[j] * nm.s[j] + nm.m[j]);
            qexp_d(t, ot + 3);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Dynamic kNN rewiring (posenet.py:1043-1050 -> torch_cluster.knn_graph [3p, 1.5.9], loop=False, source_to_target):
// per graph, every node i gets edges (j -> i) from its k nearest other nodes j in embedding space (squared Euclidean
// distance, fp32), grouped by centre, nearest first; ties go to the lower index.  One block per graph: warps compute
// the N(N-1)/2 distances, then one thread per centre selects k times.
// ------------------------------------------------------------------------------------------------
constexpr int KNN_MAX_N = 64;
// One CTA per graph.  Uniform batches: N nodes and N*k edges per graph.  Ragged batches (node_ptr / edge_ptr given):
// graph g owns nodes [node_ptr[g], node_ptr[g+1]) and edges from edge_ptr[g], min(k, n_g - 1) per node.
__global__ void __launch_bounds__(256)
knn_graph_kernel(const float* __restrict__ x, int ldx, int N_uniform, int D, int k_max, long long* __restrict__ ei, long long Et,
                 const long long* __restrict__ node_ptr, const long long* __restrict__ edge_ptr) {
    pdl_prologue();
    __shared__ float dist[KNN_MAX_N][KNN_MAX_N + 1];
    const int g = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const long long node0 = node_ptr ? node_ptr[g] : (long long)g * N_uniform;
    const int N = node_ptr ? (int)(node_ptr[g + 1] - node0) : N_uniform;
    if (N > KNN_MAX_N) __trap();                                  // the host checks the largest graph; never reached
    const int k = k_max < N - 1 ? k_max : N - 1;
    const long long edge0 = edge_ptr ? edge_ptr[g] : (long long)g * N_uniform * k_max;
    const float* xg = x + (size_t)node0 * ldx;
    const int npairs = N * (N - 1) / 2;
    for (int p = warp; p < npairs; p += nwarps) {
        // unrank p -> (i, j), i < j
        int i = 0, rem = p;
        while (rem >= N - 1 - i) { rem -= N - 1 - i; ++i; }
        const int j = i + 1 + rem;
        const float* a = xg + (size_t)i * ldx;
        const float* b = xg + (size_t)j * ldx;
        float acc = 0.f;
        for (int c = lane * 4; c < D; c += 128) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(a + c));
            const float4 v = __ldg(reinterpret_cast<const float4*>(b + c));
            const float d0 = u.x - v.x, d1 = u.y - v.y, d2 = u.z - v.z, d3 = u.w - v.w;
            acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) { dist[i][j] = acc; dist[j][i] = acc; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        unsigned long long taken = 1ull << i;                 // self excluded (loop=False)
        for (int r = 0; r < k; ++r) {
            int best = -1;
            float bd = INFINITY;
            for (int j = 0; j < N; ++j) {                     // strict <: equal distances keep the lower node index
                if ((taken >> j) & 1ull) continue;
                const float d = dist[i][j];
                if (best < 0 || d < bd) { best = j; bd = d; }
            }
            taken |= 1ull << best;
            const long long e = edge0 + (long long)i * k + r;
            ei[e] = node0 + best;                             // row 0: source = neighbour
            ei[Et + e] = node0 + i;                           // row 1: destination = centre
        }
    }
}

// dpose fp32 [rows, 6] -> bf16 [rows, 64] K-panel for the tensor-core head backward: columns j = hi, 8 + j = lo
// (dpose = hi + lo to ~2^-17), everything else zero.  Also drops `scale` into scale_out[0] (a 1-entry row-scale table).
__global__ void pack_dpose_kernel(const float* __restrict__ dpose, long long rows, bf16* __restrict__ dp16, float scale,
                                  float* __restrict__ scale_out) {
    pdl_prologue();
    if (blockIdx.x == 0 && threadIdx.x == 0 && scale_out) scale_out[0] = scale;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
        float hi[8] = {0, 0, 0, 0, 0, 0, 0, 0}, lo[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const float v = dpose[r * 6 + j];
            hi[j] = __bfloat162float(__float2bfloat16_rn(v));
            lo[j] = v - hi[j];
        }
        uint4* o = reinterpret_cast<uint4*>(dp16 + r * 64);
        o[0] = pack8(hi);
        o[1] = pack8(lo);
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int q = 2; q < 8; ++q) o[q] = z;
    }
}

// Evaluation error metrics (test.py:202-203, 262-265): translation error ||t_pred - t_gt|| and the quaternion angular
// error 2 acos(min(1, |<q1, q2>|)) in degrees (pose_utils.py:420-431), one thread per pose pair, double arithmetic.
__global__ void pose_errors_kernel(const float* __restrict__ pred7, const float* __restrict__ targ7, long long n,
                                   float* __restrict__ t_err, float* __restrict__ q_err) {
    pdl_prologue();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float* p = pred7 + i * 7;
        const float* t = targ7 + i * 7;
        double st = 0.0, d = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) { const double v = (double)p[j] - (double)t[j]; st += v * v; }
#pragma unroll
        for (int j = 3; j < 7; ++j) d += (double)p[j] * (double)t[j];
        d = fmin(1.0, fabs(d));
        t_err[i] = (float)sqrt(st);
        q_err[i] = (float)(2.0 * acos(d) * 180.0 / 3.14159265358979323846);
    }
}

// PyG-style batched edge_index of G template copies (what torch_geometric's Batch produces, train.py:24,132):
// ei[0, g*Ep + k] = g*N + src[k], ei[1, g*Ep + k] = g*N + dst[k].
__global__ void build_edge_index_kernel(const int* __restrict__ tsrc, const int* __restrict__ tdst, long long G, int N, int Ep,
                                        long long* __restrict__ ei) {
    pdl_prologue();
    const long long Et = G * Ep;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < Et; e += (long long)gridDim.x * blockDim.x) {
        const long long g = e / Ep;
        const int k = (int)(e - g * Ep);
        ei[e] = g * N + __ldg(tsrc + k);
        ei[Et + e] = g * N + __ldg(tdst + k);
    }
}

// Tables of a batch whose graphs have DIFFERENT edge sets of equal size (kNN rewiring: every graph its own 4-NN graph):
// graph g owns edge rows [g*Ep, (g+1)*Ep) and node rows [g*N, (g+1)*N).  The compute kernels see it as ONE graph of
// G*N nodes and G*Ep edges (rpg_graph_t with G = 1), so the tables hold global rows; they are built here, one thread
// per graph (counting sorts over <= a few hundred edges), without any host round trip.  bad[0] counts edges that leave
// their graph's node range.
__global__ void per_graph_tables_kernel(const long long* __restrict__ ei, int G, int N, int Ep, int* __restrict__ src,
                                        int* __restrict__ dst, int* __restrict__ in_ptr, int* __restrict__ in_idx,
                                        int* __restrict__ out_ptr, int* __restrict__ out_idx, int* __restrict__ min_ptr,
                                        int* __restrict__ min_idx, int* __restrict__ max_ptr, int* __restrict__ max_idx,
                                        float* __restrict__ inv_deg, float* __restrict__ deg, float* __restrict__ has_in,
                                        int* __restrict__ bad) {
    pdl_prologue();
    const long long Et = (long long)G * Ep;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < G; g += gridDim.x * blockDim.x) {
        const long long e0 = (long long)g * Ep;
        const int n0 = g * N;
        int nbad = 0;
        for (int k = 0; k < Ep; ++k) {
            const long long s = ei[e0 + k], d = ei[Et + e0 + k];
            if (s < n0 || s >= n0 + N || d < n0 || d >= n0 + N) ++nbad;
            src[e0 + k] = (int)s;
            dst[e0 + k] = (int)d;
        }
        if (nbad) { atomicAdd(bad, nbad); continue; }
        // four CSRs keyed by destination / source / lower / upper endpoint (stable: edge order within a node)
        for (int t = 0; t < 4; ++t) {
            int* ptr = t == 0 ? in_ptr : t == 1 ? out_ptr : t == 2 ? min_ptr : max_ptr;
            int* idx = t == 0 ? in_idx : t == 1 ? out_idx : t == 2 ? min_idx : max_idx;
            int fill = (int)e0;
            for (int n = 0; n < N; ++n) {
                ptr[n0 + n] = fill;
                for (int k = 0; k < Ep; ++k) {
                    const int s = src[e0 + k], d = dst[e0 + k];
                    const int key = t == 0 ? d : t == 1 ? s : t == 2 ? min(s, d) : max(s, d);
                    if (key == n0 + n) idx[fill++] = (int)(e0 + k);
                }
                if (t == 0) {
                    const int dg = fill - ptr[n0 + n];
                    deg[n0 + n] = (float)dg;
                    inv_deg[n0 + n] = 1.f / (float)max(dg, 1);
                    has_in[n0 + n] = dg > 0 ? 1.f : 0.f;
                }
            }
            if (g == G - 1) ptr[(long long)G * N] = (int)Et;
        }
    }
}

// Small host tables (graph templates) travel as KERNEL PARAMETERS: no staging buffer, no copy engine -- an upload can
// never queue behind a large host->device copy of another stream.
constexpr int UPLOAD_WORDS = 2032;
struct UploadWords { int32_t w[UPLOAD_WORDS]; };
__global__ void __launch_bounds__(256)
upload_words_kernel(const __grid_constant__ UploadWords p, int32_t* __restrict__ dst, int n) {
    pdl_prologue();
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) dst[i] = p.w[i];
}

// Batched form: blockIdx.y picks the descriptor; always accumulating.  A block covers 32 consecutive elements with 8
// split phases (phase q sums splits q, q + 8, ...; the phases are folded through shared memory in a fixed order), so
// small outputs with many splits do not serialise on one thread.  Deterministic.
// (small outputs; large ones use reduce_splits_batch_kernel below: one thread per element, splits in sequence)
__global__ void __launch_bounds__(256)
reduce_splits_batch_phased_kernel(const __grid_constant__ rpg_reduce_batch_t batch) {
    pdl_prologue();
    __shared__ float red[8][32];
    const rpg_reduce_desc_t& d = batch.d[blockIdx.y];
    const long long n = (long long)d.rows * d.cols;
    const int lane = threadIdx.x & 31, ph = threadIdx.x >> 5;
    for (long long base = blockIdx.x * 32LL; base < n; base += (long long)gridDim.x * 32) {
        const long long i = base + lane;
        float s0 = 0.f, s1 = 0.f;
        if (i < n) {
            int b = ph;
            for (; b + 8 < d.splits; b += 16) {
                s0 += d.part[(size_t)b * d.stride + i];
                s1 += d.part[(size_t)(b + 8) * d.stride + i];
            }
            if (b < d.splits) s0 += d.part[(size_t)b * d.stride + i];
        }
        red[ph][lane] = s0 + s1;
        __syncthreads();
        if (ph == 0 && i < n) {
            float s = red[0][lane];
#pragma unroll
            for (int k = 1; k < 8; ++k) s += red[k][lane];
            const int r = (int)(i / d.cols), c = (int)(i - (long long)r * d.cols);
            float* o = d.out + (size_t)r * d.ldo + c;
            *o += s;
        }
        __syncthreads();
    }
}

// One thread per 4 consecutive elements (float4) when the shapes allow, 8 independent loads in flight per thread:
// the fold is a pure HBM stream over the split partials (~200 MB per layer backward).
__global__ void __launch_bounds__(256)
reduce_splits_batch_kernel(const __grid_constant__ rpg_reduce_batch_t batch) {
    pdl_prologue();
    const rpg_reduce_desc_t& d = batch.d[blockIdx.y];
    const long long n = (long long)d.rows * d.cols;
    const bool vec = (d.cols % 4 == 0) && (d.ldo % 4 == 0) && (d.stride % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(d.part) | reinterpret_cast<uintptr_t>(d.out)) % 16 == 0);
    if (vec) {
        const long long n4 = n >> 2;
        for (long long i4 = blockIdx.x * 256LL + threadIdx.x; i4 < n4; i4 += (long long)gridDim.x * 256) {
            const long long i = i4 << 2;
            const int r = (int)(i / d.cols), c = (int)(i - (long long)r * d.cols);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            int b = 0;
            for (; b + 8 <= d.splits; b += 8) {
                float4 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __ldg(reinterpret_cast<const float4*>(d.part + (size_t)(b + j) * d.stride + i));
#pragma unroll
                for (int j = 0; j < 8; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
            }
            for (; b < d.splits; ++b) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(d.part + (size_t)b * d.stride + i));
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            float4* o = reinterpret_cast<float4*>(d.out + (size_t)r * d.ldo + c);
            float4 cur = *o;
            cur.x += acc.x; cur.y += acc.y; cur.z += acc.z; cur.w += acc.w;
            *o = cur;
        }
        return;
    }
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const int r = (int)(i / d.cols), c = (int)(i - (long long)r * d.cols);
        float s0 = 0.f;
        for (int b = 0; b < d.splits; ++b) s0 += d.part[(size_t)b * d.stride + i];
        float* o = d.out + (size_t)r * d.ldo + c;
        *o += s0;
    }
}

// ------------------------------------------------------------------------------------------------
// compute_RP (posenet.py:1021-1031) + L1 sums of PoseNetCriterion (criterion.py:51-52) + sign gradient
// ------------------------------------------------------------------------------------------------
constexpr int LOSS_THREADS = 256;
__global__ void __launch_bounds__(LOSS_THREADS)
pose_loss_kernel(const float* __restrict__ pred, const float* __restrict__ poses, const int* __restrict__ tsrc,
                 const int* __restrict__ tdst, long long Et, int N, int Ep, const float* __restrict__ grad_scale,
                 const float* __restrict__ sax, const float* __restrict__ saq,
                 float* __restrict__ target, float* __restrict__ dpred, float* __restrict__ partial, int ld_pred) {
    pdl_prologue();
    float st = 0.f, sq = 0.f;
    float gs_t = grad_scale ? grad_scale[0] : 1.f, gs_q = grad_scale ? grad_scale[1] : 1.f;
    if (sax) {   // criterion.py:55-57: d loss / d pred = exp(-s) * sign / (3 Et)
        gs_t = (float)(exp(-(double)sax[0]) / (3.0 * (double)Et));
        gs_q = (float)(exp(-(double)saq[0]) / (3.0 * (double)Et));
    }
    for (long long e = blockIdx.x * (long long)LOSS_THREADS + threadIdx.x; e < Et; e += (long long)gridDim.x * LOSS_THREADS) {
        const long long g = e / Ep;
        const int k = (int)(e - g * Ep);
        const float* ps = poses + (g * N + __ldg(tsrc + k)) * 6;
        const float* pd = poses + (g * N + __ldg(tdst + k)) * 6;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const float t = ps[j] - pd[j];
            const float diff = pred[e * ld_pred + j] - t;
            if (target) target[e * 6 + j] = t;
            if (j < 3) st += fabsf(diff); else sq += fabsf(diff);
            if (dpred) dpred[e * 6 + j] = (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)) * (j < 3 ? gs_t : gs_q);
        }
    }
    __shared__ float s_t[LOSS_THREADS / 32], s_q[LOSS_THREADS / 32];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        st += __shfl_xor_sync(0xffffffffu, st, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if ((threadIdx.x & 31) == 0) { s_t[threadIdx.x >> 5] = st; s_q[threadIdx.x >> 5] = sq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < LOSS_THREADS / 32; ++i) { a += s_t[i]; b += s_q[i]; }
        partial[blockIdx.x * 2] = a;
        partial[blockIdx.x * 2 + 1] = b;
    }
}
// sums[0..1] = L1 sums; with sax/saq also the learned-weight combination of criterion.py:55-57:
// sums[2..6] = loss, t_loss, q_loss, d loss/d sax, d loss/d saq.
__global__ void pose_loss_final_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ sums,
                                       const float* __restrict__ sax, const float* __restrict__ saq, long long Et) {
    pdl_prologue();
    float a = 0.f, b = 0.f;                                  // one warp, fixed order
    for (int i = threadIdx.x; i < nblocks; i += 32) { a += partial[2 * i]; b += partial[2 * i + 1]; }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        sums[0] = a; sums[1] = b;
        if (sax) {
            const float n = 3.f * (float)Et;
            const float tl = a / n, ql = b / n;
            const float ex = (float)exp(-(double)sax[0]), eq = (float)exp(-(double)saq[0]);
            sums[2] = ex * tl + sax[0] + eq * ql + saq[0];
            sums[3] = tl; sums[4] = ql;
            sums[5] = 1.f - ex * tl; sums[6] = 1.f - eq * ql;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Column sums of a bf16 matrix (bias gradients): stage 1 per row-slab, stage 2 over slabs.
// ------------------------------------------------------------------------------------------------
constexpr int COLSUM_ROWS = 64;      // short slabs: enough blocks to fill the machine at node-level row counts
constexpr int COLSUM_THREADS = 256;
// A block owns COLSUM_ROWS rows and all columns: thread = (column group of 8, row phase); phases are folded in smem.
__global__ void __launch_bounds__(COLSUM_THREADS)
colsum_stage1_kernel(const bf16* __restrict__ v, int ldv, long long rows, int cols,
                     const float* __restrict__ row_w, int row_w_mod, float* __restrict__ part) {
    pdl_prologue();
    __shared__ float red[8][COLSUM_THREADS];
    const int cg = cols >> 3;                                   // <= COLSUM_THREADS (host-checked)
    const int phases = COLSUM_THREADS / cg;
    const int my_cg = threadIdx.x % cg, my_ph = threadIdx.x / cg;
    const long long r0 = (long long)blockIdx.x * COLSUM_ROWS, r1 = min(r0 + COLSUM_ROWS, rows);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (my_ph < phases) {
#pragma unroll 4
        for (long long r = r0 + my_ph; r < r1; r += phases) {
            float f[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(v + r * ldv + my_cg * 8)), f);
            const float wgt = row_w ? __ldg(row_w + (int)((unsigned long long)r % (unsigned)row_w_mod)) : 1.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = fmaf(wgt, f[q], acc[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) red[q][threadIdx.x] = acc[q];
    __syncthreads();
    if (my_ph == 0) {
        for (int ph = 1; ph < phases; ++ph) {
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] += red[q][ph * cg + my_cg];
        }
        float* o = part + (size_t)blockIdx.x * cols + my_cg * 8;
        *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}
// out[i] (+)= sum_b part[b, i] with the parts spread over 8 phases per column (fixed order => deterministic)
__global__ void __launch_bounds__(256)
reduce_partials_wide_kernel(const float* __restrict__ part, int nparts, long long stride, int n,
                            float* __restrict__ out, int accumulate, float* __restrict__ out2) {
    pdl_prologue();
    __shared__ float red[8][32];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31), ph = threadIdx.x >> 5;
    float s = 0.f;
    if (col < n)
        for (int b = ph; b < nparts; b += 8) s += part[(size_t)b * stride + col];
    red[ph][threadIdx.x & 31] = s;
    __syncthreads();
    if (ph == 0 && col < n) {
#pragma unroll
        for (int k = 1; k < 8; ++k) s += red[k][threadIdx.x & 31];
        out[col] = accumulate ? out[col] + s : s;
        if (out2) out2[col] = accumulate ? out2[col] + s : s;      // the same sums into a second parameter gradient
    }
}

// column sums into one or two outputs (two parameter gradients that receive the same sums: one pass over the tensor)
int colsum_bf16_2(const rpg_bf16* v, int ldv, int64_t rows, int cols, const float* row_w, int row_w_mod, float* out,
                       float* out2, int accumulate, float* scratch, cudaStream_t s) {
    if (!v || !out || !scratch || rows <= 0 || cols % 8 || ldv % 8 || cols / 8 > COLSUM_THREADS)
        return set_error(RPG_E_ARG, "colsum: bad arguments (cols must be a multiple of 8, at most 2048)");
    const int slabs = (int)((rows + COLSUM_ROWS - 1) / COLSUM_ROWS);
    launch_pdl(colsum_stage1_kernel, dim3(slabs), dim3(COLSUM_THREADS), 0, s, reinterpret_cast<const bf16*>(v), ldv, rows, cols, row_w,
                                                          row_w_mod > 0 ? row_w_mod : 1, scratch);
    int rc = check_launch("colsum_stage1_kernel");
    if (rc) return rc;
    launch_pdl(reduce_partials_wide_kernel, dim3((cols + 31) / 32), dim3(256), 0, s, scratch, slabs, cols, cols, out, accumulate, out2);
    return check_launch("reduce_partials_wide_kernel");
}


}  // namespace rpg

// =================================================================================================
// C ABI
// =================================================================================================
using namespace rpg;

extern "C" {

int rpg_validate_edge_index(const int64_t* edge_index, int64_t Et, int G, int N, int Ep, int32_t* tmpl_src,
                            int32_t* tmpl_dst, int32_t* bad_count, rpg_stream_t stream) {
    if (!edge_index || !tmpl_src || !tmpl_dst || !bad_count) return set_error(RPG_E_ARG, "validate: null pointer");
    if (G <= 0 || N <= 0 || Ep <= 0 || Et != (int64_t)G * Ep) return set_error(RPG_E_GRAPH, "validate: Et != G * Ep");
    cudaStream_t s = as_stream(stream);
    cudaMemsetAsync(bad_count, 0, sizeof(int32_t), s);
    launch_pdl(validate_edge_index_kernel, dim3(grid_for(Et, 256)), dim3(256), 0, s, reinterpret_cast<const long long*>(edge_index), Et, G, N,
                                                                  Ep, tmpl_src, tmpl_dst, bad_count);
    return check_launch("validate_edge_index_kernel");
}

int rpg_selection_patterns(const int32_t* endpoint, int Ep, int N, int div, int patterns, rpg_bf16* sel, rpg_stream_t stream) {
    if (!endpoint || !sel || Ep <= 0 || N <= 0 || div <= 0 || patterns <= 0) return set_error(RPG_E_ARG, "selection_patterns: bad arguments");
    launch_pdl(selection_patterns_kernel, dim3(grid_for((long long)patterns * 128 * 8, 256)), dim3(256), 0, as_stream(stream), 
        endpoint, Ep, N, div, patterns, reinterpret_cast<bf16*>(sel));
    return check_launch("selection_patterns_kernel");
}

int rpg_pack_weight(const float* src, int ld_src, int r0, int c0, int rows, int cols, rpg_bf16* dst, int ld_dst,
                    int transpose, rpg_stream_t stream) {
    if (!src || !dst || rows <= 0 || cols <= 0) return set_error(RPG_E_ARG, "pack_weight: bad arguments");
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    launch_pdl(pack_weight_kernel, dim3(grid), dim3(block), 0, as_stream(stream), src, ld_src, r0, c0, rows, cols,
                                                              reinterpret_cast<bf16*>(dst), ld_dst, transpose);
    return check_launch("pack_weight_kernel");
}

int rpg_cast_f32_to_bf16(const float* src, rpg_bf16* dst, int64_t n, rpg_stream_t stream) {
    if (!src || !dst || n < 0) return set_error(RPG_E_ARG, "cast: bad arguments");
    if (n == 0) return 0;
    launch_pdl(cast_f32_bf16_kernel, dim3(grid_for((n + 7) / 8, 256)), dim3(256), 0, as_stream(stream), src, reinterpret_cast<bf16*>(dst), n);
    return check_launch("cast_f32_bf16_kernel");
}
int rpg_cast_bf16_to_f32(const rpg_bf16* src, float* dst, int64_t n, rpg_stream_t stream) {
    if (!src || !dst || n < 0) return set_error(RPG_E_ARG, "cast: bad arguments");
    if (n == 0) return 0;
    launch_pdl(cast_bf16_f32_kernel, dim3(grid_for((n + 7) / 8, 256)), dim3(256), 0, as_stream(stream), reinterpret_cast<const bf16*>(src), dst, n);
    return check_launch("cast_bf16_f32_kernel");
}

int rpg_cast_f32_to_split(const float* src, rpg_bf16* hi, rpg_bf16* lo, int64_t n, rpg_stream_t stream) {
    if (!src || !hi || !lo || n <= 0) return set_error(RPG_E_ARG, "cast_f32_to_split: bad arguments");
    launch_pdl(cast_f32_split_kernel, dim3(grid_for(n, 256)), dim3(256), 0, as_stream(stream), src, reinterpret_cast<bf16*>(hi), reinterpret_cast<bf16*>(lo), n);
    return check_launch("cast_f32_split_kernel");
}
int rpg_split_to_f32(const rpg_bf16* hi, const rpg_bf16* lo, float* dst, int64_t n, rpg_stream_t stream) {
    if (!dst || !hi || !lo || n <= 0) return set_error(RPG_E_ARG, "split_to_f32: bad arguments");
    launch_pdl(split_to_f32_kernel, dim3(grid_for(n, 256)), dim3(256), 0, as_stream(stream), reinterpret_cast<const bf16*>(hi), reinterpret_cast<const bf16*>(lo), dst, n);
    return check_launch("split_to_f32_kernel");
}
int rpg_pack_weight_lo(const float* src, int ld_src, int r0, int c0, int rows, int cols, rpg_bf16* dst, int ld_dst,
                       rpg_stream_t stream) {
    if (!src || !dst || rows <= 0 || cols <= 0) return set_error(RPG_E_ARG, "pack_weight_lo: bad arguments");
    dim3 grid((cols + 127) / 128, rows);
    launch_pdl(pack_weight_lo_kernel, dim3(grid), dim3(128), 0, as_stream(stream), src, ld_src, r0, c0, rows, cols, reinterpret_cast<bf16*>(dst), ld_dst);
    return check_launch("pack_weight_lo_kernel");
}
static int edge_init_f32_impl(const float* pminmax, int ldp, const float* bias, const rpg_graph_t* graph, int D, rpg_bf16* e_hi,
                              rpg_bf16* e_lo, int lde, uint8_t* bits, rpg_stream_t stream) {
    if (!pminmax || !bias || !graph || !e_hi || !e_lo || D % 8 || ldp % 4 || lde % 4) return set_error(RPG_E_ARG, "edge_init_fwd_f32: bad arguments");
    const long long Et = (long long)graph->G * graph->Ep;
    // total threads is even (D / 4 is even) and a block is a multiple of 32: shuffle partners always exist
    launch_pdl(edge_init_fwd_f32_kernel, dim3(grid_for(Et * (D / 4), 256)), dim3(256), 0, as_stream(stream),
        pminmax, ldp, bias, graph->src, graph->dst, Et, graph->N, graph->Ep, D, reinterpret_cast<bf16*>(e_hi),
        reinterpret_cast<bf16*>(e_lo), lde, bits);
    return check_launch("edge_init_fwd_f32_kernel");
}
int rpg_edge_init_fwd_f32(const float* pminmax, int ldp, const float* bias, const rpg_graph_t* graph, int D, rpg_bf16* e_hi,
                          rpg_bf16* e_lo, int lde, rpg_stream_t stream) {
    return edge_init_f32_impl(pminmax, ldp, bias, graph, D, e_hi, e_lo, lde, nullptr, stream);
}
int rpg_edge_init_fwd_split(const float* pminmax, int ldp, const float* bias, const rpg_graph_t* graph, int D, rpg_bf16* e_hi,
                            rpg_bf16* e_lo, int lde, uint8_t* e_bits, rpg_stream_t stream) {
    return edge_init_f32_impl(pminmax, ldp, bias, graph, D, e_hi, e_lo, lde, e_bits, stream);
}

int rpg_attention_fwd(const float* gtp, int64_t Et, int c, rpg_bf16* y, int ldy, rpg_bf16* y_lo, float* aux,
                      rpg_stream_t stream) {
    if (!gtp || !y || Et <= 0) return set_error(RPG_E_ARG, "attention_fwd: bad arguments");
    if (!aux && attention_series_enabled() && c % 16 == 0 && c >= 16 && c <= 256 && ldy % 8 == 0)
        return attention_series_fwd(gtp, 0, Et, c, y, ldy, y_lo, as_stream(stream));
    if (c % 4 || c < 4 || c > 256 || ldy % 2) return set_error(RPG_E_UNSUPPORTED, "attention_fwd: c must be a multiple of 4 in [4,256]");
    const int grid = grid_for(Et, ATT_WARPS, 148 * 8);
    ProfScope prof(RPG_PROF_ATTENTION_FWD, (double)Et * c * (12.0 + 2.0 + (y_lo ? 2.0 : 0.0) + (aux ? 16.0 : 0.0)), as_stream(stream),
                   (double)Et * c * c);          // "flops" field = exp evaluations (the kernel is MUFU-bound)
    if (aux)
        launch_pdl((attention_fwd_kernel<256, true>), dim3(grid), dim3(ATT_WARPS * 32), 0, as_stream(stream), gtp, Et, c, reinterpret_cast<bf16*>(y), ldy,
                                                                                      reinterpret_cast<bf16*>(y_lo), aux);
    else
        launch_pdl((attention_fwd_kernel<256, false>), dim3(grid), dim3(ATT_WARPS * 32), 0, as_stream(stream), gtp, Et, c, reinterpret_cast<bf16*>(y), ldy,
                                                                                       reinterpret_cast<bf16*>(y_lo), nullptr);
    return check_launch("attention_fwd_kernel");
}

int rpg_attention_bwd(const float* gtp, const float* dyn, int ld_dyn, const rpg_graph_t* graph, int64_t Et, int c,
                      rpg_bf16* dgtp, int ld_dgtp, const float* aux, rpg_stream_t stream) {
    if (!gtp || !dyn || !graph || !dgtp || Et <= 0) return set_error(RPG_E_ARG, "attention_bwd: bad arguments");
    if (!aux && attention_series_enabled() && c % 16 == 0 && c >= 16 && c <= 256 && ld_dgtp % 8 == 0 && ld_dyn % 4 == 0)
        return attention_series_bwd(gtp, 0, dyn, ld_dyn, graph, Et, c, dgtp, ld_dgtp, nullptr, as_stream(stream));
    if (c % 4 || c < 4 || c > 512) return set_error(RPG_E_UNSUPPORTED, "attention_bwd: c must be a multiple of 4 in [4,512]");
    const size_t smem = (size_t)ATT_WARPS * 8 * c * sizeof(float);
    static SmemLimit configured;
    ensure_dyn_smem(attention_bwd_kernel, smem, &configured);
    const int grid = grid_for(Et, ATT_WARPS, 148 * 8);
    ProfScope prof(RPG_PROF_ATTENTION_BWD, (double)Et * c * (12.0 + 6.0 + (aux ? 16.0 : 0.0)) + (double)graph->G * graph->N * c * 4.0,
                   as_stream(stream), (double)Et * c * c * (aux ? 1.0 : 2.0));
    launch_pdl(attention_bwd_kernel, dim3(grid), dim3(ATT_WARPS * 32), smem, as_stream(stream), gtp, dyn, ld_dyn, graph->dst, graph->Ep, graph->N,
                                                                           Et, c, reinterpret_cast<bf16*>(dgtp), ld_dgtp, aux);
    return check_launch("attention_bwd_kernel");
}

static int launch_segment(const rpg_bf16* v, int ldv, const rpg_bf16* mask, int ldm, const int32_t* ptr, const int32_t* idx,
                          const float* scale, const rpg_graph_t* g, int D, rpg_bf16* out, int ldo, cudaStream_t s,
                          const rpg_bf16* v_lo = nullptr, rpg_bf16* out_lo = nullptr) {
    if (!v || !g || !out || !ptr || !idx || D % 8 || ldv % 8 || ldo % 8) return set_error(RPG_E_ARG, "segment_sum: bad arguments");
    const long long Nt = (long long)g->G * g->N;
    const bool small = Nt * (D / 8) + 148LL * 16 * 256 < (1LL << 31);
    const bool plain = !mask && !v_lo && !out_lo;
    // algorithmic bytes: every edge row of the CSR read once, every node row written once
    ProfScope prof(RPG_PROF_SEGMENT_SUM, ((double)g->G * g->Ep * (mask ? 2 : 1) * (v_lo ? 2 : 1) + (double)Nt * (out_lo ? 2 : 1)) * D * 2.0, s);
    auto kern = small ? (plain ? segment_sum_kernel<unsigned, true> : segment_sum_kernel<unsigned, false>)
                      : (plain ? segment_sum_kernel<long long, true> : segment_sum_kernel<long long, false>);
    launch_pdl(kern, dim3(grid_for(Nt * (D / 8), 256)), dim3(256), 0, s, reinterpret_cast<const bf16*>(v), ldv,
               reinterpret_cast<const bf16*>(mask), ldm, ptr, idx, scale, Nt, g->N, g->Ep, D, reinterpret_cast<bf16*>(out), ldo,
               reinterpret_cast<const bf16*>(v_lo), reinterpret_cast<bf16*>(out_lo));
    return check_launch("segment_sum_kernel");
}

int rpg_aggregate_mean(const rpg_bf16* z, int ldz, const rpg_graph_t* graph, int D, rpg_bf16* a, int lda,
                       rpg_stream_t stream) {
    if (!graph) return set_error(RPG_E_ARG, "aggregate_mean: null graph");
    return launch_segment(z, ldz, nullptr, 0, graph->in_ptr, graph->in_idx, graph->inv_deg, graph, D, a, lda, as_stream(stream));
}

int rpg_aggregate_mean_split(const rpg_bf16* z_hi, const rpg_bf16* z_lo, int ldz, const rpg_graph_t* graph, int D,
                             rpg_bf16* a_hi, rpg_bf16* a_lo, int lda, rpg_stream_t stream) {
    if (!graph || !z_lo || !a_lo) return set_error(RPG_E_ARG, "aggregate_mean_split: null argument");
    return launch_segment(z_hi, ldz, nullptr, 0, graph->in_ptr, graph->in_idx, graph->inv_deg, graph, D, a_hi, lda,
                          as_stream(stream), z_lo, a_lo);
}

int rpg_attention_fwd_bf16(const rpg_bf16* gtp16, int64_t Et, int c, rpg_bf16* y, int ldy, rpg_stream_t stream) {
    if (!gtp16 || !y || Et <= 0) return set_error(RPG_E_ARG, "attention_fwd_bf16: bad arguments");
    return attention_series_fwd(gtp16, 1, Et, c, y, ldy, nullptr, as_stream(stream));
}
int rpg_attention_bwd_bf16(const rpg_bf16* gtp16, const float* dyn, int ld_dyn, const rpg_graph_t* graph, int64_t Et, int c,
                           rpg_bf16* dgtp, int ld_dgtp, rpg_stream_t stream) {
    if (!gtp16 || !dyn || !graph || !dgtp || Et <= 0) return set_error(RPG_E_ARG, "attention_bwd_bf16: bad arguments");
    return attention_series_bwd(gtp16, 1, dyn, ld_dyn, graph, Et, c, dgtp, ld_dgtp, nullptr, as_stream(stream));
}

int rpg_segment_sum_split(const rpg_bf16* v_hi, const rpg_bf16* v_lo, int ldv, const int32_t* csr_ptr, const int32_t* csr_idx,
                          const float* scale, const rpg_graph_t* graph, int D, rpg_bf16* out_hi, rpg_bf16* out_lo, int ldo,
                          rpg_stream_t stream) {
    if (!graph || !v_lo || !out_lo) return set_error(RPG_E_ARG, "segment_sum_split: null argument");
    return launch_segment(v_hi, ldv, nullptr, 0, csr_ptr, csr_idx, scale, graph, D, out_hi, ldo, as_stream(stream), v_lo, out_lo);
}

int rpg_attention_bwd_split(const float* gtp, const float* dyn, int ld_dyn, const rpg_graph_t* graph, int64_t Et, int c,
                            rpg_bf16* dgtp_hi, rpg_bf16* dgtp_lo, int ld_dgtp, rpg_stream_t stream) {
    if (!gtp || !dyn || !graph || !dgtp_hi || !dgtp_lo || Et <= 0) return set_error(RPG_E_ARG, "attention_bwd_split: bad arguments");
    return attention_series_bwd(gtp, 0, dyn, ld_dyn, graph, Et, c, dgtp_hi, ld_dgtp, dgtp_lo, as_stream(stream));
}

int rpg_edge_to_node_sum(const rpg_bf16* v, int ldv, const rpg_graph_t* graph, int D, int by_src, rpg_bf16* out, int ldo,
                         rpg_stream_t stream) {
    if (!graph) return set_error(RPG_E_ARG, "edge_to_node_sum: null graph");
    return launch_segment(v, ldv, nullptr, 0, by_src ? graph->out_ptr : graph->in_ptr, by_src ? graph->out_idx : graph->in_idx,
                          nullptr, graph, D, out, ldo, as_stream(stream));
}

int rpg_segment_sum(const rpg_bf16* v, int ldv, const rpg_bf16* mask, int ldm, const int32_t* csr_ptr, const int32_t* csr_idx,
                    const float* scale, const rpg_graph_t* graph, int D, rpg_bf16* out, int ldo, rpg_stream_t stream) {
    return launch_segment(v, ldv, mask, ldm, csr_ptr, csr_idx, scale, graph, D, out, ldo, as_stream(stream));
}

int rpg_segment_sum2(const rpg_bf16* v, int ldv, const rpg_graph_t* g, int which_a, int which_b, int D, rpg_bf16* out_a,
                     int ldo_a, rpg_bf16* out_b, int ldo_b, rpg_stream_t stream) {
    if (!v || !g || !out_a || !out_b || D % 8 || ldv % 8 || ldo_a % 8 || ldo_b % 8 || (which_a & ~3) || (which_b & ~3))
        return set_error(RPG_E_ARG, "segment_sum2: bad arguments");
    const int32_t* ptrs[4] = {g->in_ptr, g->out_ptr, g->min_ptr, g->max_ptr};
    const int32_t* idxs[4] = {g->in_idx, g->out_idx, g->min_idx, g->max_idx};
    const long long Nt = (long long)g->G * g->N;
    cudaStream_t s = as_stream(stream);
    // algorithmic bytes: the edge tensor once, two node tensors
    ProfScope prof(RPG_PROF_SEGMENT_SUM, ((double)g->G * g->Ep + 2.0 * (double)Nt) * D * 2.0, s);
    const bool small = 2 * Nt * (D / 8) + 148LL * 16 * 256 < (1LL << 31);
    auto kern = small ? segment_sum2_kernel<unsigned> : segment_sum2_kernel<long long>;
    launch_pdl(kern, dim3(grid_for(2 * Nt * (D / 8), 256)), dim3(256), 0, s, reinterpret_cast<const bf16*>(v), ldv, ptrs[which_a],
               idxs[which_a], ptrs[which_b], idxs[which_b], Nt, g->N, g->Ep, D, reinterpret_cast<bf16*>(out_a), ldo_a,
               reinterpret_cast<bf16*>(out_b), ldo_b);
    return check_launch("segment_sum2_kernel");
}

int rpg_edge_init_fwd(const rpg_bf16* pminmax, int ldp, const float* bias, const rpg_graph_t* graph, int D, rpg_bf16* e0,
                      int lde, uint8_t* e0_bits, rpg_stream_t stream) {
    if (!pminmax || !bias || !graph || !e0 || D % 8 || ldp % 8 || lde % 8) return set_error(RPG_E_ARG, "edge_init_fwd: bad arguments");
    const long long Et = (long long)graph->G * graph->Ep;
    const bool small = Et * (D / 8) + 148LL * 16 * 256 < (1LL << 31);
    ProfScope prof(RPG_PROF_EDGE_INIT, (double)Et * D * 2.0 + (e0_bits ? (double)Et * D / 8 : 0.0) + (double)graph->G * graph->N * 2 * D * 2.0,
                   as_stream(stream));
    auto kern = small ? edge_init_fwd_kernel<unsigned> : edge_init_fwd_kernel<long long>;
    launch_pdl(kern, dim3(grid_for(Et * (D / 8), 256)), dim3(256), 0, as_stream(stream),
        reinterpret_cast<const bf16*>(pminmax), ldp, bias, graph->src, graph->dst, Et, graph->N, graph->Ep, D,
        reinterpret_cast<bf16*>(e0), lde, e0_bits);
    return check_launch("edge_init_fwd_kernel");
}

int rpg_dropout_mask(uint64_t seed, float p_drop, int64_t rows, int D, uint8_t* keep, rpg_stream_t stream) {
    if (!keep || rows <= 0 || D <= 0 || p_drop < 0.f || p_drop >= 1.f) return set_error(RPG_E_ARG, "dropout_mask: bad arguments");
    const uint32_t thresh = (uint32_t)(p_drop * 256.0f + 0.5f);
    launch_pdl(dropout_mask_kernel, dim3(grid_for(rows * D, 256)), dim3(256), 0, as_stream(stream), seed, thresh, rows, D, keep);
    return check_launch("dropout_mask_kernel");
}

int rpg_head_fwd(const rpg_bf16* feat, const rpg_bf16* feat_lo, int ldf, int64_t rows, int D, const uint8_t* keep,
                 uint64_t seed, float p_drop, const float* w6, const float* b6, float* pose, rpg_stream_t stream) {
    if (!feat || !w6 || !b6 || !pose || rows <= 0 || D % 8 || ldf % 8) return set_error(RPG_E_ARG, "head_fwd: bad arguments");
    const int use_seed = (!keep && p_drop > 0.f) ? 1 : 0;
    const uint32_t thresh = (uint32_t)(p_drop * 256.0f + 0.5f);
    // explicit masks: F.dropout's 1 / (1 - p); seeded: the rate actually applied (p quantised to 1/256), so E[out] = in
    const float scale = keep ? 1.f / (1.f - p_drop) : (use_seed ? 256.f / (256.f - (float)thresh) : 1.f);
    const int grid = grid_for((rows + 1) / 2, HEAD_WARPS, 148 * 8);
    const bf16* f = reinterpret_cast<const bf16*>(feat);
    const bf16* fl = reinterpret_cast<const bf16*>(feat_lo);
    cudaStream_t st = as_stream(stream);
    const bool plain = !keep && !fl;
    if (D <= 256) {
        if (plain) launch_pdl((head_fwd_kernel<1, true>), dim3(grid), dim3(HEAD_WARPS * 32), 0, st, f, ldf, rows, D, keep, seed, thresh, use_seed, scale, w6, b6, pose, fl);
        else launch_pdl((head_fwd_kernel<1, false>), dim3(grid), dim3(HEAD_WARPS * 32), 0, st, f, ldf, rows, D, keep, seed, thresh, use_seed, scale, w6, b6, pose, fl);
    } else if (D <= 512) {
        if (plain) launch_pdl((head_fwd_kernel<2, true>), dim3(grid), dim3(HEAD_WARPS * 32), 0, st, f, ldf, rows, D, keep, seed, thresh, use_seed, scale, w6, b6, pose, fl);
        else launch_pdl((head_fwd_kernel<2, false>), dim3(grid), dim3(HEAD_WARPS * 32), 0, st, f, ldf, rows, D, keep, seed, thresh, use_seed, scale, w6, b6, pose, fl);
    } else {
        const size_t smem = (size_t)6 * D * sizeof(float);
        if (D > 2048) return set_error(RPG_E_UNSUPPORTED, "head_fwd: D > 2048");
        static SmemLimit configured;
        ensure_dyn_smem(head_fwd_kernel<0, false>, smem, &configured);
        launch_pdl((head_fwd_kernel<0, false>), dim3(grid), dim3(HEAD_WARPS * 32), smem, st, f, ldf, rows, D, keep, seed, thresh, use_seed, scale, w6, b6, pose, fl);
    }
    return check_launch("head_fwd_kernel");
}

int64_t rpg_head_bwd_ws_floats(int64_t rows, int D) {
    const int64_t blocks = (rows + HEADB_ROWS_PER_BLOCK - 1) / HEADB_ROWS_PER_BLOCK;
    return blocks * (6 * (int64_t)D + 6);
}

static int head_bwd_impl(const float* dpose, const rpg_bf16* feat, const rpg_bf16* feat_lo, int ldf, int64_t rows, int D,
                         const uint8_t* keep, uint64_t seed, float p_drop, const float* w6, int mask_relu, rpg_bf16* dfeat,
                         rpg_bf16* dfeat_lo, int lddf, float* dw_t, float* dw_q, float* db_t, float* db_q, int accumulate,
                         float* ws, rpg_stream_t stream);

int rpg_head_bwd(const float* dpose, const rpg_bf16* feat, int ldf, int64_t rows, int D, const uint8_t* keep, uint64_t seed,
                 float p_drop, const float* w6, int mask_relu, rpg_bf16* dfeat, int lddf, float* dw_t, float* dw_q,
                 float* db_t, float* db_q, int accumulate, float* ws, rpg_stream_t stream) {
    return head_bwd_impl(dpose, feat, nullptr, ldf, rows, D, keep, seed, p_drop, w6, mask_relu, dfeat, nullptr, lddf, dw_t, dw_q,
                         db_t, db_q, accumulate, ws, stream);
}

int rpg_head_bwd_split(const float* dpose, const rpg_bf16* feat_hi, const rpg_bf16* feat_lo, int ldf, int64_t rows, int D,
                       const uint8_t* keep, uint64_t seed, float p_drop, const float* w6, int mask_relu, rpg_bf16* dfeat_hi,
                       rpg_bf16* dfeat_lo, int lddf, float* dw_t, float* dw_q, float* db_t, float* db_q, int accumulate,
                       float* ws, rpg_stream_t stream) {
    if (!feat_lo || (dfeat_hi && !dfeat_lo)) return set_error(RPG_E_ARG, "head_bwd_split: low planes missing");
    return head_bwd_impl(dpose, feat_hi, feat_lo, ldf, rows, D, keep, seed, p_drop, w6, mask_relu, dfeat_hi, dfeat_lo, lddf,
                         dw_t, dw_q, db_t, db_q, accumulate, ws, stream);
}

static int head_bwd_impl(const float* dpose, const rpg_bf16* feat, const rpg_bf16* feat_lo, int ldf, int64_t rows, int D,
                         const uint8_t* keep, uint64_t seed, float p_drop, const float* w6, int mask_relu, rpg_bf16* dfeat,
                         rpg_bf16* dfeat_lo, int lddf, float* dw_t, float* dw_q, float* db_t, float* db_q, int accumulate,
                         float* ws, rpg_stream_t stream) {
    if (!dpose || !feat || !w6 || !dw_t || !dw_q || !db_t || !db_q || !ws || rows <= 0 || D % 8 || ldf % 8 || D > 8 * 1024)
        return set_error(RPG_E_ARG, "head_bwd: bad arguments");
    const int blocks = (int)((rows + HEADB_ROWS_PER_BLOCK - 1) / HEADB_ROWS_PER_BLOCK);
    float* dw_part = ws;
    float* db_part = ws + (size_t)blocks * 6 * D;
    const int use_seed = (!keep && p_drop > 0.f) ? 1 : 0;
    const uint32_t thresh = (uint32_t)(p_drop * 256.0f + 0.5f);
    // explicit masks: F.dropout's 1 / (1 - p); seeded: the rate actually applied (p quantised to 1/256), so E[out] = in
    const float scale = keep ? 1.f / (1.f - p_drop) : (use_seed ? 256.f / (256.f - (float)thresh) : 1.f);
    const int ct = D / 8;
    if (ct > HEADB_THREADS) return set_error(RPG_E_UNSUPPORTED, "head_bwd: D > 2048");
    const int phases = HEADB_THREADS / ct;
    const size_t smem = ((size_t)(phases - 1) * 48 * ct + (size_t)phases * 6) * sizeof(float);
    static SmemLimit configured;
    ensure_dyn_smem(head_bwd_kernel, smem, &configured);
    cudaStream_t s = as_stream(stream);
    launch_pdl(head_bwd_kernel, dim3(blocks), dim3(HEADB_THREADS), smem, s, dpose, reinterpret_cast<const bf16*>(feat), ldf, rows, D, keep, seed,
                                                        thresh, use_seed, scale, w6, mask_relu,
                                                        reinterpret_cast<bf16*>(dfeat), lddf, dw_part, db_part,
                                                        reinterpret_cast<const bf16*>(feat_lo), reinterpret_cast<bf16*>(dfeat_lo));
    int rc = check_launch("head_bwd_kernel");
    if (rc) return rc;
    // rows 0..2 -> translation head (fc_xyz*), rows 3..5 -> rotation head (fc_wpqr*)
    launch_pdl(head_bwd_reduce_kernel, dim3((6 * D + 31) / 32 + 1), dim3(256), 0, s, dw_part, db_part, blocks, D, dw_t, dw_q, db_t, db_q, accumulate);
    return check_launch("head_bwd_reduce_kernel");
}

int64_t rpg_pose_loss_ws_floats(int64_t Et) { return 2 * (int64_t)grid_for(Et, LOSS_THREADS, 296); }

static int pose_loss_launch(const float* pred, const float* poses, const rpg_graph_t* graph, int64_t Et, const float* grad_scale,
                            const float* sax, const float* saq,
                            float* target, float* sums, float* dpred, float* ws, rpg_stream_t stream, int ld_pred = 6) {
    if (!pred || !poses || !graph || !sums || !ws || Et <= 0 || ld_pred < 6) return set_error(RPG_E_ARG, "pose_loss: bad arguments");
    const int blocks = grid_for(Et, LOSS_THREADS, 296);
    cudaStream_t s = as_stream(stream);
    launch_pdl(pose_loss_kernel, dim3(blocks), dim3(LOSS_THREADS), 0, s, pred, poses, graph->src, graph->dst, Et, graph->N, graph->Ep, grad_scale,
                                                     sax, saq, target, dpred, ws, ld_pred);
    int rc = check_launch("pose_loss_kernel");
    if (rc) return rc;
    launch_pdl(pose_loss_final_kernel, dim3(1), dim3(32), 0, s, ws, blocks, sums, sax, saq, Et);
    return check_launch("pose_loss_final_kernel");
}

int rpg_pose_loss(const float* pred, const float* poses, const rpg_graph_t* graph, int64_t Et, const float* grad_scale,
                  float* target, float* sums, float* dpred, float* ws, rpg_stream_t stream) {
    return pose_loss_launch(pred, poses, graph, Et, grad_scale, nullptr, nullptr, target, sums, dpred, ws, stream);
}

int rpg_pose_criterion(const float* pred, int ld_pred, const float* poses, const rpg_graph_t* graph, int64_t Et, const float* sax,
                       const float* saq, float* target, float* out7, float* dpred, float* ws, rpg_stream_t stream) {
    if (!sax || !saq) return set_error(RPG_E_ARG, "pose_criterion: sax/saq must be device pointers");
    return pose_loss_launch(pred, poses, graph, Et, nullptr, sax, saq, target, out7, dpred, ws, stream, ld_pred);
}

int64_t rpg_colsum_scratch_floats(int64_t rows, int cols) {
    return ((rows + COLSUM_ROWS - 1) / COLSUM_ROWS) * (int64_t)cols;
}

int rpg_colsum_bf16(const rpg_bf16* v, int ldv, int64_t rows, int cols, const float* row_w, int row_w_mod, float* out,
                    int accumulate, float* scratch, rpg_stream_t stream) {
    return colsum_bf16_2(v, ldv, rows, cols, row_w, row_w_mod, out, nullptr, accumulate, scratch, as_stream(stream));
}

int rpg_reduce_splits(const float* partial, int splits, int64_t split_stride, int rows, int cols, float* out, int ldo,
                      int accumulate, rpg_stream_t stream) {
    if (!partial || !out || splits < 1 || rows <= 0 || cols <= 0) return set_error(RPG_E_ARG, "reduce_splits: bad arguments");
    launch_pdl(reduce_splits_kernel, dim3(grid_for((long long)rows * cols, 256)), dim3(256), 0, as_stream(stream), partial, splits, split_stride, rows,
                                                                                              cols, out, ldo, accumulate);
    return check_launch("reduce_splits_kernel");
}

int rpg_edge_gather(const rpg_bf16* pa, int lda, int which_a, const rpg_bf16* pb, int ldb, int which_b, const float* bias,
                    const rpg_graph_t* graph, int D, int relu, const uint8_t* mask_bits, rpg_bf16* out, int ldo,
                    uint8_t* out_bits, rpg_stream_t stream) {
    if (!pa || !graph || !out || D % 8 || lda % 8 || ldo % 8 || (pb && ldb % 8) || (which_a & ~1) || (which_b & ~1))
        return set_error(RPG_E_ARG, "edge_gather: bad arguments");
    const long long Et = (long long)graph->G * graph->Ep;
    launch_pdl(edge_gather_kernel, dim3(grid_for(Et * (D / 8), 256)), dim3(256), 0, as_stream(stream),
               reinterpret_cast<const bf16*>(pa), lda, which_a, reinterpret_cast<const bf16*>(pb), ldb, which_b, bias,
               graph->src, graph->dst, Et, graph->N, graph->Ep, D, relu, mask_bits, reinterpret_cast<bf16*>(out), ldo,
               out_bits);
    return check_launch("edge_gather_kernel");
}

int rpg_edge_mask_apply(const void* in, int64_t G, int Ep_full, int Ep_kept, const int32_t* kept_idx, int row_bytes, void* out,
                        rpg_stream_t stream) {
    if (!in || !out || !kept_idx || G <= 0 || Ep_full <= 0 || Ep_kept <= 0 || Ep_kept > Ep_full || row_bytes <= 0 || row_bytes % 4)
        return set_error(RPG_E_ARG, "edge_mask_apply: bad arguments (rows must be a multiple of 4 bytes)");
    const int words = row_bytes / 4;
    launch_pdl(edge_rows_select_kernel, dim3(grid_for(G * Ep_kept * (long long)words, 256)), dim3(256), 0, as_stream(stream),
               reinterpret_cast<const uint32_t*>(in), (long long)G, Ep_full, Ep_kept, kept_idx, words,
               reinterpret_cast<uint32_t*>(out));
    return check_launch("edge_rows_select_kernel");
}

int rpg_scale_rows(const rpg_bf16* v, int ldv, int64_t rows, int D, const float* scale, int mod, rpg_bf16* out, int ldo,
                   rpg_stream_t stream) {
    if (!v || !scale || !out || rows <= 0 || D % 8 || ldv % 8 || ldo % 8 || mod <= 0)
        return set_error(RPG_E_ARG, "scale_rows: bad arguments");
    launch_pdl(scale_rows_kernel, dim3(grid_for(rows * (D / 8), 256)), dim3(256), 0, as_stream(stream),
               reinterpret_cast<const bf16*>(v), ldv, rows, D, scale, mod, reinterpret_cast<bf16*>(out), ldo);
    return check_launch("scale_rows_kernel");
}

int rpg_pack_dpose(const float* dpose, int64_t rows, rpg_bf16* dp16, float scale, float* scale_out, rpg_stream_t stream) {
    if (!dpose || !dp16 || rows <= 0) return set_error(RPG_E_ARG, "pack_dpose: bad arguments");
    launch_pdl(pack_dpose_kernel, dim3(grid_for(rows, 256)), dim3(256), 0, as_stream(stream), dpose, rows,
               reinterpret_cast<bf16*>(dp16), scale, scale_out);
    return check_launch("pack_dpose_kernel");
}

int rpg_per_graph_tables(const int64_t* edge_index, int G, int N, int Ep, int32_t* tables, int32_t* bad, rpg_stream_t stream) {
    if (!edge_index || !tables || !bad || G <= 0 || N <= 0 || Ep <= 0 || (long long)G * Ep > 0x7fffffffLL || (long long)G * N > 0x7ffffff0LL)
        return set_error(RPG_E_ARG, "per_graph_tables: bad arguments");
    const long long Et = (long long)G * Ep, Nt = (long long)G * N;
    // layout of `tables` (int32 words, every table 4-word aligned): see rpg_per_graph_tables_words()
    auto al = [](long long v) { return (v + 3) / 4 * 4; };
    int32_t* p = tables;
    int32_t* src = p; p += al(Et);
    int32_t* dst = p; p += al(Et);
    int32_t* in_ptr = p; p += al(Nt + 1);
    int32_t* in_idx = p; p += al(Et);
    int32_t* out_ptr = p; p += al(Nt + 1);
    int32_t* out_idx = p; p += al(Et);
    int32_t* min_ptr = p; p += al(Nt + 1);
    int32_t* min_idx = p; p += al(Et);
    int32_t* max_ptr = p; p += al(Nt + 1);
    int32_t* max_idx = p; p += al(Et);
    float* inv_deg = reinterpret_cast<float*>(p); p += al(Nt);
    float* deg = reinterpret_cast<float*>(p); p += al(Nt);
    float* has_in = reinterpret_cast<float*>(p);
    cudaMemsetAsync(bad, 0, sizeof(int32_t), as_stream(stream));
    launch_pdl(per_graph_tables_kernel, dim3(grid_for(G, 64)), dim3(64), 0, as_stream(stream),
               reinterpret_cast<const long long*>(edge_index), G, N, Ep, src, dst, in_ptr, in_idx, out_ptr, out_idx, min_ptr,
               min_idx, max_ptr, max_idx, inv_deg, deg, has_in, bad);
    return check_launch("per_graph_tables_kernel");
}

int64_t rpg_per_graph_tables_words(int G, int N, int Ep) {
    const long long Et = (long long)G * Ep, Nt = (long long)G * N;
    auto al = [](long long v) { return (v + 3) / 4 * 4; };
    return 6 * al(Et) + 4 * al(Nt + 1) + 3 * al(Nt);
}

int rpg_selection_patterns_rows(const int32_t* endpoint, int64_t Et, int pg_Ep, int pg_N, rpg_bf16* sel, int32_t* bad,
                                rpg_stream_t stream) {
    if (!endpoint || !sel || !bad || Et <= 0 || pg_Ep <= 0 || pg_N <= 0) return set_error(RPG_E_ARG, "selection_patterns_rows: bad arguments");
    const long long rows = (Et + 127) / 128 * 128;
    launch_pdl(selection_patterns_rows_kernel, dim3(grid_for(rows * 8, 256)), dim3(256), 0, as_stream(stream), endpoint,
               (long long)Et, pg_Ep, pg_N, reinterpret_cast<bf16*>(sel), bad);
    return check_launch("selection_patterns_rows_kernel");
}

int rpg_build_edge_index(const rpg_graph_t* graph, int64_t* edge_index, rpg_stream_t stream) {
    if (!graph || !edge_index || graph->G <= 0 || graph->Ep <= 0) return set_error(RPG_E_ARG, "build_edge_index: bad arguments");
    const long long Et = (long long)graph->G * graph->Ep;
    launch_pdl(build_edge_index_kernel, dim3(grid_for(Et, 256)), dim3(256), 0, as_stream(stream), graph->src, graph->dst,
               (long long)graph->G, graph->N, graph->Ep, reinterpret_cast<long long*>(edge_index));
    return check_launch("build_edge_index_kernel");
}

int rpg_knn_graph(const float* x, int ldx, int G, int N, int D, int k, int64_t* edge_index, rpg_stream_t stream) {
    if (!x || !edge_index || G <= 0 || N < 2 || N > KNN_MAX_N || k < 1 || k >= N || D % 4 || ldx % 4)
        return set_error(RPG_E_ARG, "knn_graph: need 2 <= N <= 64, 1 <= k < N, D and pitch multiples of 4");
    const long long Et = (long long)G * N * k;
    launch_pdl(knn_graph_kernel, dim3(G), dim3(256), 0, as_stream(stream), x, ldx, N, D, k,
               reinterpret_cast<long long*>(edge_index), Et, (const long long*)nullptr, (const long long*)nullptr);
    return check_launch("knn_graph_kernel");
}

int rpg_knn_graph_ragged(const float* x, int ldx, int G, const int64_t* node_ptr, const int64_t* edge_ptr, int max_nodes, int D,
                         int k, int64_t n_edges, int64_t* edge_index, rpg_stream_t stream) {
    if (!x || !edge_index || !node_ptr || !edge_ptr || G <= 0 || max_nodes < 1 || max_nodes > KNN_MAX_N || k < 1 || n_edges < 0 ||
        D % 4 || ldx % 4)
        return set_error(RPG_E_ARG, "knn_graph_ragged: need graphs of <= 64 nodes, k >= 1, D and pitch multiples of 4");
    if (n_edges == 0) return 0;
    launch_pdl(knn_graph_kernel, dim3(G), dim3(256), 0, as_stream(stream), x, ldx, 0, D, k,
               reinterpret_cast<long long*>(edge_index), (long long)n_edges, reinterpret_cast<const long long*>(node_ptr),
               reinterpret_cast<const long long*>(edge_ptr));
    return check_launch("knn_graph_kernel");
}

int rpg_pose_errors(const float* pred7, const float* targ7, int64_t n, float* t_err, float* q_err, rpg_stream_t stream) {
    if (!pred7 || !targ7 || !t_err || !q_err || n <= 0) return set_error(RPG_E_ARG, "pose_errors: bad arguments");
    launch_pdl(pose_errors_kernel, dim3(grid_for(n, 128)), dim3(128), 0, as_stream(stream), pred7, targ7, n, t_err, q_err);
    return check_launch("pose_errors_kernel");
}

int rpg_qexp(const float* v, int64_t n, float* q, rpg_stream_t stream) {
    if (!v || !q || n <= 0) return set_error(RPG_E_ARG, "qexp: bad arguments");
    launch_pdl(qexp_kernel, dim3(grid_for(n, 128)), dim3(128), 0, as_stream(stream), v, n, q);
    return check_launch("qexp_kernel");
}

int rpg_eval_compose(const float* pred_edges, const float* poses, const rpg_graph_t* graph, int ref_k, const float* pose_m,
                     const float* pose_s, float* out_pred, float* out_targ, rpg_stream_t stream) {
    if (!pred_edges || !poses || !graph || !out_pred || ref_k < 0 || ref_k >= graph->Ep)
        return set_error(RPG_E_ARG, "eval_compose: bad arguments");
    EvalNorm nm;
    for (int j = 0; j < 3; ++j) { nm.m[j] = pose_m ? pose_m[j] : 0.f; nm.s[j] = pose_s ? pose_s[j] : 1.f; }
    launch_pdl(eval_compose_kernel, dim3(grid_for(graph->G, 128)), dim3(128), 0, as_stream(stream), pred_edges, poses,
               graph->src, graph->G, graph->N, graph->Ep, ref_k, nm, out_pred, out_targ);
    return check_launch("eval_compose_kernel");
}

int rpg_upload_words(int32_t* dst, const int32_t* src_host, int64_t n, rpg_stream_t stream) {
    if (!dst || !src_host || n <= 0) return set_error(RPG_E_ARG, "upload_words: bad arguments");
    static_assert(sizeof(UploadWords) <= 8192, "parameter block");
    for (int64_t off = 0; off < n; off += UPLOAD_WORDS) {
        UploadWords p;
        const int cnt = (int)std::min<int64_t>(UPLOAD_WORDS, n - off);
        memcpy(p.w, src_host + off, (size_t)cnt * sizeof(int32_t));
        launch_pdl(upload_words_kernel, dim3((cnt + 255) / 256), dim3(256), 0, as_stream(stream), p, dst + off, cnt);
        int rc = check_launch("upload_words_kernel");
        if (rc) return rc;
    }
    return 0;
}

int rpg_reduce_splits_batch(const rpg_reduce_batch_t* batch, rpg_stream_t stream) {
    if (!batch || batch->n < 1 || batch->n > RPG_REDUCE_BATCH_MAX) return set_error(RPG_E_ARG, "reduce_splits_batch: bad arguments");
    long long most = 1;
    for (int i = 0; i < batch->n; ++i) {
        const rpg_reduce_desc_t& d = batch->d[i];
        if (!d.part || !d.out || d.splits < 1 || d.rows <= 0 || d.cols <= 0)
            return set_error(RPG_E_ARG, "reduce_splits_batch: bad descriptor");
        most = std::max(most, (long long)d.rows * d.cols);
    }
    if (most <= 16384) {      // few elements, many splits (head gradients): spread the splits over 8 phases per element
        dim3 grid((unsigned)grid_for(most, 32, 8192), (unsigned)batch->n);
        launch_pdl(reduce_splits_batch_phased_kernel, dim3(grid), dim3(256), 0, as_stream(stream), *batch);
        return check_launch("reduce_splits_batch_phased_kernel");
    }
    dim3 grid((unsigned)grid_for(most, 256, 1024), (unsigned)batch->n);
    launch_pdl(reduce_splits_batch_kernel, dim3(grid), dim3(256), 0, as_stream(stream), *batch);
    return check_launch("reduce_splits_batch_kernel");
}

}  // extern "C"
