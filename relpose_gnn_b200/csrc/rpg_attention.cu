// Channel attention (att.py:25-30) without transcendental work: the logits of one edge row are RANK ONE,
//   f_ij = phi_i * theta_j,   y_i = sum_j softmax_j(f_ij) g_j,
// so with theta centred (beta_j = theta_j - b0, b0 = mid-range of theta; the factor exp(phi_i b0) cancels in the softmax)
//   exp(phi_i beta_j) = sum_k phi_i^k beta_j^k / k!
// is SEPARABLE in (i, j):  D_i = sum_j exp(.) = sum_k phi_i^k U_k,  N_i = sum_j exp(.) g_j = sum_k phi_i^k T_k  with the 2 (K+1)
// moments U_k = sum_j beta_j^k / k!, T_k = sum_j beta_j^k g_j / k!.  One row costs O(c K) fused multiply-adds instead of c^2
// exponentials (c = 64: 4096 MUFU.EX2 forward and again backward, which made the former kernels MUFU-bound at 200 us per
// launch); the backward is separable in the same way (moments over i).  The order K follows from the row's own range
// R = max_i |phi_i| * (max_j theta_j - min_j theta_j) / 2: the truncation error relative to the softmax denominator is
// <= R^(K+1) / (K+1)!; K is taken one order higher than a 2e-8 bound needs (fp32 mode) or a 2e-6 bound (bf16 mode, where
// the result is rounded to 8 bits right after), the extra order covering the derivative polynomials of the backward.
// Rows beyond the K = 20 bound (R > 3.4 / 4.3) take the EXACT path -- warp-per-row exp2 with the rank-1 row maximum --
// inside the same kernel, so the result is the reference's softmax to fp32 accuracy for every input.  Measured on B200
// (294 912 rows, c = 64, K = 10): forward 115 us, backward 194 us against 323 / 701 us for the exp2 kernels; rows on the
// exact path cost about twice the exp2 kernels (8 warps per SM hide the MUFU latency less well).
//
// Mapping: one THREAD per edge row (the moments are private sums: no shuffles), rows staged through shared memory with
// 16-byte cp.async (coalesced; row pitch 3c + 4 floats keeps the per-thread float4 reads bank-conflict free), results
// written back through the same tile (coalesced).  Arithmetic is packed fp32x2 (FFMA2): the pairs (U_k, T_k), (D, N),
// (A_k, B_k) share one instruction.  Bound: HBM (3c * 4 B in, c * 2 B out per row) once K is small.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "../../include/rpg.h"
#include "rpg_internal.h"
#include "rpg_ptx.cuh"

namespace rpg {

typedef __nv_bfloat16 bf16;

constexpr float LOG2E_F = 1.4426950408889634f;

__device__ __forceinline__ float2 b2(float v) { return make_float2(v, v); }

// 1 / k!
__device__ __forceinline__ float inv_fact(int k) {
    constexpr float t[24] = {1.f, 1.f, 0.5f, 1.f / 6, 1.f / 24, 1.f / 120, 1.f / 720, 1.f / 5040, 1.f / 40320, 1.f / 362880,
                             1.f / 3628800, 1.f / 39916800, 1.f / 479001600, 1.f / 6227020800.f, 1.f / 87178291200.f,
                             1.f / 1307674368000.f, 1.f / 20922789888000.f, 1.f / 355687428096000.f,
                             1.f / 6402373705728000.f, 1.f / 121645100408832000.f, 1.f / 2432902008176640000.f,
                             1.f / 51090942171709440000.f, 1.f / 1124000727777607680000.f, 1.f / 25852016738884976640000.f};
    return t[k];
}

// Series class of a row from its range R (see the header): 0..4 -> K = 4 / 6 / 10 / 14 / 20, 5 -> exact path.
// strict (fp32 mode, results kept as (hi, lo) planes): R^K / K! <= 2e-8 -- the order K - 1 bound, one order of margin for
// the derivative polynomials of the backward; measured against an fp64 softmax at the class limits: truncation < 1e-9,
// fp32 evaluation 1e-7 (tests/test_gpu_attention.py).  Otherwise (bf16 mode: the result is rounded to 8 bits right
// after): R^K / K! <= 2e-6.
constexpr int EXACT_CLASS = 5;
__device__ __forceinline__ int series_class(float R, bool strict) {
    if (strict) return R <= 0.026f ? 0 : (R <= 0.15f ? 1 : (R <= 0.75f ? 2 : (R <= 1.7f ? 3 : (R <= 3.4f ? 4 : 5))));
    return R <= 0.083f ? 0 : (R <= 0.33f ? 1 : (R <= 1.2f ? 2 : (R <= 2.4f ? 3 : (R <= 4.3f ? 4 : 5))));
}

// ---- staged rows: the projections (g | theta | phi) arrive as fp32 (fp32 mode) or bf16 (bf16 mode: half the HBM
// traffic and half the shared memory, i.e. twice the resident warps); a thread reads and writes its own row in 16-byte
// pieces (VEC = 4 floats / 8 bf16) -- with a row pitch of 3c elements + 16 bytes every quarter-warp hits 32 distinct banks.
template <typename T> struct RowT;
template <> struct RowT<float> {
    static constexpr int VEC = 4;
    static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
        const float4 a = *reinterpret_cast<const float4*>(p);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
    static __device__ __forceinline__ void st(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
    static __device__ __forceinline__ float ld1(const float* p) { return *p; }
};
template <> struct RowT<bf16> {
    static constexpr int VEC = 8;
    static __device__ __forceinline__ void ld(const bf16* p, float (&v)[8]) {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
        v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
    }
    static __device__ __forceinline__ void st(bf16* p, const float (&v)[8]) {
        uint4 u;
        u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(p) = u;
    }
    static __device__ __forceinline__ float ld1(const bf16* p) { return __bfloat162float(*p); }
};

// (U_k, T_k) = (sum_j beta_j^k, sum_j beta_j^k g_j) / k!  for k = 0..KM over the staged row r = (g | theta | phi).
// Two columns j share every instruction: P = (beta_a^k, beta_b^k), U2_k += P, T2_k += P * (g_a, g_b), P *= (beta_a, beta_b).
template <int KM, typename T>
__device__ __forceinline__ void series_moments(const T* r, int c, float b0, float2 (&UT)[KM + 1]) {
    constexpr int V = RowT<T>::VEC;
    float2 U2[KM + 1], T2[KM + 1];
#pragma unroll
    for (int k = 0; k <= KM; ++k) { U2[k] = make_float2(0.f, 0.f); T2[k] = make_float2(0.f, 0.f); }
    const float2 nb0 = b2(-b0);
    for (int j = 0; j < c; j += V) {
        float tv[V], gv[V];
        RowT<T>::ld(r + c + j, tv);
        RowT<T>::ld(r + j, gv);
#pragma unroll
        for (int h = 0; h < V / 2; ++h) {
            const float2 beta = __fadd2_rn(make_float2(tv[2 * h], tv[2 * h + 1]), nb0);
            const float2 gg = make_float2(gv[2 * h], gv[2 * h + 1]);
            T2[0] = __fadd2_rn(T2[0], gg);                   // k = 0: P = 1
            float2 P = beta;
#pragma unroll
            for (int k = 1; k <= KM; ++k) {
                U2[k] = __fadd2_rn(U2[k], P);
                T2[k] = __ffma2_rn(P, gg, T2[k]);
                if (k < KM) P = __fmul2_rn(P, beta);
            }
        }
    }
    UT[0] = make_float2((float)c, T2[0].x + T2[0].y);
#pragma unroll
    for (int k = 1; k <= KM; ++k) UT[k] = make_float2((U2[k].x + U2[k].y) * inv_fact(k), (T2[k].x + T2[k].y) * inv_fact(k));
}

// forward of one row: y_i = N_i / D_i written over g (r[i]); the row's g values are consumed before
template <int KM, typename T>
__device__ __forceinline__ void series_fwd_row(T* r, int c, float b0) {
    constexpr int V = RowT<T>::VEC;
    float2 UT[KM + 1];
    series_moments<KM, T>(r, c, b0, UT);
    for (int i = 0; i < c; i += V) {
        float pv[V], yv[V];
        RowT<T>::ld(r + 2 * c + i, pv);
#pragma unroll
        for (int q = 0; q < V; ++q) {
            float2 DN = UT[KM];
#pragma unroll
            for (int k = KM - 1; k >= 0; --k) DN = __ffma2_rn(DN, b2(pv[q]), UT[k]);
            yv[q] = __fdividef(DN.y, DN.x);
        }
        RowT<T>::st(r + i, yv);
    }
}

// backward of one row, in place: r = (g | theta | phi) -> (dg | dtheta | dphi); dy: the gradient w.r.t. y of this row
template <int KM, typename T>
__device__ __forceinline__ void series_bwd_row(T* r, int c, float b0, const float* __restrict__ dy) {
    constexpr int V = RowT<T>::VEC;
    float2 UT[KM + 1];
    series_moments<KM, T>(r, c, b0, UT);
    // (A_k, B_k) = (sum_i w_i phi_i^k, sum_i w_i y_i phi_i^k), k = 0..KM+1, two rows i per instruction
    float2 A2[KM + 2], B2[KM + 2];
#pragma unroll
    for (int k = 0; k <= KM + 1; ++k) { A2[k] = make_float2(0.f, 0.f); B2[k] = make_float2(0.f, 0.f); }
    for (int i = 0; i < c; i += V) {
        float pv[V], dphi[V], yv[V], wv[V];
        RowT<T>::ld(r + 2 * c + i, pv);
#pragma unroll
        for (int q4 = 0; q4 < V; q4 += 4) {
            const float4 d4 = __ldg(reinterpret_cast<const float4*>(dy + i + q4));
            const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
                const int q = q4 + qq;
                const float2 ph = b2(pv[q]);
                float2 DN = UT[KM], dDN = make_float2(0.f, 0.f);     // polynomial and its derivative (Horner)
#pragma unroll
                for (int k = KM - 1; k >= 0; --k) {
                    dDN = __ffma2_rn(dDN, ph, DN);
                    DN = __ffma2_rn(DN, ph, UT[k]);
                }
                const float invD = __fdividef(1.f, DN.x);
                yv[q] = DN.y * invD;
                wv[q] = dd[qq] * invD;
                dphi[q] = wv[q] * (dDN.y - yv[q] * dDN.x);           // sum_j dl_ij theta_j (the b0 part sums to zero)
            }
        }
        RowT<T>::st(r + 2 * c + i, dphi);
#pragma unroll
        for (int h = 0; h < V / 2; ++h) {
            const float2 ph = make_float2(pv[2 * h], pv[2 * h + 1]), y2 = make_float2(yv[2 * h], yv[2 * h + 1]);
            float2 Q = make_float2(wv[2 * h], wv[2 * h + 1]);
#pragma unroll
            for (int k = 0; k <= KM + 1; ++k) {
                A2[k] = __fadd2_rn(A2[k], Q);
                B2[k] = __ffma2_rn(Q, y2, B2[k]);
                if (k <= KM) Q = __fmul2_rn(Q, ph);
            }
        }
    }
    // dg_j = sum_k A_k beta^k / k!;  dtheta_j = g_j sum_k A_{k+1} beta^k / k! - sum_k B_{k+1} beta^k / k!
    float C[KM + 1], SA[KM + 1], SB[KM + 1];
#pragma unroll
    for (int k = 0; k <= KM; ++k) {
        C[k] = (A2[k].x + A2[k].y) * inv_fact(k);
        SA[k] = (A2[k + 1].x + A2[k + 1].y) * inv_fact(k);
        SB[k] = (B2[k + 1].x + B2[k + 1].y) * inv_fact(k);
    }
    const float2 nb0 = b2(-b0);
    for (int j = 0; j < c; j += V) {
        float tv[V], gv[V], dgo[V], dto[V];
        RowT<T>::ld(r + c + j, tv);
        RowT<T>::ld(r + j, gv);
#pragma unroll
        for (int h = 0; h < V / 2; ++h) {
            const float2 beta = __fadd2_rn(make_float2(tv[2 * h], tv[2 * h + 1]), nb0);
            float2 a0 = b2(C[KM]), sa = b2(SA[KM]), sb = b2(SB[KM]);
#pragma unroll
            for (int k = KM - 1; k >= 0; --k) {
                a0 = __ffma2_rn(a0, beta, b2(C[k]));
                sa = __ffma2_rn(sa, beta, b2(SA[k]));
                sb = __ffma2_rn(sb, beta, b2(SB[k]));
            }
            dgo[2 * h] = a0.x; dgo[2 * h + 1] = a0.y;
            dto[2 * h] = gv[2 * h] * sa.x - sb.x; dto[2 * h + 1] = gv[2 * h + 1] * sa.y - sb.y;
        }
        RowT<T>::st(r + j, dgo);
        RowT<T>::st(r + c + j, dto);
    }
}

// ---- exact path (rows whose range exceeds the series bound): one WARP per row, exp2 with the rank-1 row maximum
__device__ __forceinline__ float2 ex2_2(float2 a) { return make_float2(exp2f(a.x), exp2f(a.y)); }

template <typename T>
__device__ void exact_fwd_row(const T* r, int c, float tmax, float tmin, bf16* __restrict__ y, bf16* __restrict__ y_lo,
                              int lane) {
    for (int i0 = 2 * lane; i0 < c; i0 += 64) {
        const float2 P = make_float2(RowT<T>::ld1(r + 2 * c + i0) * LOG2E_F, RowT<T>::ld1(r + 2 * c + i0 + 1) * LOG2E_F);
        const float2 nM = make_float2(-(P.x >= 0.f ? P.x * tmax : P.x * tmin), -(P.y >= 0.f ? P.y * tmax : P.y * tmin));
        float2 num = make_float2(0.f, 0.f), den = num;
        for (int j = 0; j < c; ++j) {
            const float2 e = ex2_2(__ffma2_rn(P, b2(RowT<T>::ld1(r + c + j)), nM));
            num = __ffma2_rn(e, b2(RowT<T>::ld1(r + j)), num);
            den = __fadd2_rn(den, e);
        }
        const float y0 = num.x / den.x, y1 = num.y / den.y;
        *reinterpret_cast<uint32_t*>(y + i0) = pack_bf16x2(y0, y1);
        if (y_lo)
            *reinterpret_cast<uint32_t*>(y_lo + i0) =
                pack_bf16x2(y0 - __bfloat162float(__float2bfloat16_rn(y0)), y1 - __bfloat162float(__float2bfloat16_rn(y1)));
    }
}

// scratch: 5c floats per warp (P_i, -M_i, w_i, w_i phi_i, w_i y_i phi_i)
template <typename T>
__device__ void exact_bwd_row(const T* r, int c, float tmax, float tmin, const float* __restrict__ dy,
                              float* scratch, bf16* __restrict__ dgtp, int lane) {
    float* sp = scratch;
    float* sm = sp + c;
    float* sw = sm + c;
    float* swp = sw + c;
    float* swyp = swp + c;
    for (int i = lane; i < c; i += 32) {
        const float phi = RowT<T>::ld1(r + 2 * c + i);
        const float P = phi * LOG2E_F;
        const float nM = -(P >= 0.f ? P * tmax : P * tmin);
        float den = 0.f, num = 0.f, at = 0.f, agt = 0.f;
        for (int j = 0; j < c; ++j) {
            const float th = RowT<T>::ld1(r + c + j), gj = RowT<T>::ld1(r + j);
            const float e = exp2f(fmaf(P, th, nM));
            den += e; num = fmaf(e, gj, num); at = fmaf(e, th, at); agt = fmaf(e, gj * th, agt);
        }
        const float inv = 1.f / den, yv = num * inv, w = __ldg(dy + i) * inv;
        sp[i] = P; sm[i] = nM; sw[i] = w; swp[i] = w * phi; swyp[i] = w * yv * phi;
        dgtp[2 * c + i] = __float2bfloat16_rn(w * (agt - yv * at));
    }
    __syncwarp();
    for (int j = lane; j < c; j += 32) {
        const float th = RowT<T>::ld1(r + c + j);
        float dg = 0.f, sa = 0.f, sb = 0.f;
        for (int i = 0; i < c; ++i) {
            const float e = exp2f(fmaf(th, sp[i], sm[i]));
            dg = fmaf(e, sw[i], dg); sa = fmaf(e, swp[i], sa); sb = fmaf(e, swyp[i], sb);
        }
        dgtp[j] = __float2bfloat16_rn(dg);
        dgtp[c + j] = __float2bfloat16_rn(RowT<T>::ld1(r + j) * sa - sb);
    }
    __syncwarp();
}

// Row tiles are staged with 16-byte cp.async (consecutive lanes -> consecutive 16 bytes of the dense [Et, 3c] tensor);
// rows beyond Et are zero-filled.  (A double-buffered variant with one warp per scheduler was measured slower: the
// kernel is issue-bound, so the second tile costs more in lost warps than the overlap buys.)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
template <typename T>
__device__ __forceinline__ void stage_rows(const T* __restrict__ gtp, long long row0, long long Et, int c, T* tile,
                                           int pitch, int lane) {
    constexpr int V = RowT<T>::VEC;
    const int per_row = 3 * c / V;                       // 16-byte chunks per row
    const T* src = gtp + row0 * 3 * c;
    int rr = 0, q = lane;
    while (q >= per_row) { q -= per_row; ++rr; }
    while (rr < 32) {
        T* dst = tile + rr * pitch + V * q;
        if (row0 + rr < Et) cp_async16(dst, src + (size_t)rr * 3 * c + V * q);
        else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
        q += 32;
        while (q >= per_row) { q -= per_row; ++rr; }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
}

// per-thread row statistics: theta extremes and the series class
template <typename T>
__device__ __forceinline__ int row_stats(const T* r, int c, float& tmax, float& tmin, float& b0, bool strict) {
    constexpr int V = RowT<T>::VEC;
    tmax = -INFINITY; tmin = INFINITY;
    float amax = 0.f;
    for (int j = 0; j < c; j += V) {
        float tv[V], pv[V];
        RowT<T>::ld(r + c + j, tv);
        RowT<T>::ld(r + 2 * c + j, pv);
#pragma unroll
        for (int q = 0; q < V; ++q) {
            tmax = fmaxf(tmax, tv[q]); tmin = fminf(tmin, tv[q]); amax = fmaxf(amax, fabsf(pv[q]));
        }
    }
    b0 = 0.5f * (tmax + tmin);
    const float R = amax * 0.5f * (tmax - tmin);
    return (R == R) ? series_class(R, strict) : EXACT_CLASS;       // NaN / inf inputs propagate through the exact path
}

// rows of the tile (first `cols` elements of each) -> bf16 (and optionally the low plane) rows in global memory
template <typename T>
__device__ __forceinline__ void store_rows_bf16(const T* tile, int pitch, int cols, long long row0, long long Et,
                                                unsigned skip_rows, bf16* __restrict__ out, int ld, bf16* __restrict__ out_lo,
                                                int lane) {
    const int per_row = cols / 8;
    int rr = 0, q = lane;
    while (q >= per_row) { q -= per_row; ++rr; }
    while (rr < 32) {
        if (row0 + rr < Et && !((skip_rows >> rr) & 1u)) {
            const T* src = tile + rr * pitch + 8 * q;
            float f[8];
            if (RowT<T>::VEC == 8) {
                *reinterpret_cast<uint4*>(out + (row0 + rr) * ld + 8 * q) = *reinterpret_cast<const uint4*>(src);
            } else {
                float a[4], b[4];
                RowT<float>::ld(reinterpret_cast<const float*>(src), a);
                RowT<float>::ld(reinterpret_cast<const float*>(src) + 4, b);
#pragma unroll
                for (int t = 0; t < 4; ++t) { f[t] = a[t]; f[4 + t] = b[t]; }
                uint4 u;
                u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
                *reinterpret_cast<uint4*>(out + (row0 + rr) * ld + 8 * q) = u;
                if (out_lo) {
                    float l[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) l[t] = f[t] - __bfloat162float(__float2bfloat16_rn(f[t]));
                    uint4 v;
                    v.x = pack_bf16x2(l[0], l[1]); v.y = pack_bf16x2(l[2], l[3]); v.z = pack_bf16x2(l[4], l[5]); v.w = pack_bf16x2(l[6], l[7]);
                    *reinterpret_cast<uint4*>(out_lo + (row0 + rr) * ld + 8 * q) = v;
                }
            }
        }
        q += 32;
        while (q >= per_row) { q -= per_row; ++rr; }
    }
}

// bf16 rows: four CTAs per SM fit in shared memory, so the register budget is capped at 128 -- the rarely taken high-order
// branches (K >= 14) spill, the common ones do not.
template <typename T>
__global__ void __launch_bounds__(128, sizeof(T) == 2 ? 4 : 2)
attention_series_fwd_kernel(const T* __restrict__ gtp, long long Et, int c, bf16* __restrict__ y, int ldy,
                            bf16* __restrict__ y_lo, int force_exact) {
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char ats_smem_raw[];
    T* ats_smem = reinterpret_cast<T*>(ats_smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int pitch = 3 * c + RowT<T>::VEC;
    T* tile = ats_smem + (size_t)warp * 32 * pitch;
    T* r = tile + lane * pitch;
    const bool strict = y_lo != nullptr;
    const long long nblk = (Et + 31) / 32;
    for (long long blk = (long long)blockIdx.x * nwarps + warp; blk < nblk; blk += (long long)gridDim.x * nwarps) {
        const long long row0 = blk * 32;
        stage_rows<T>(gtp, row0, Et, c, tile, pitch, lane);
        float tmax, tmin, b0;
        int cls = row_stats<T>(r, c, tmax, tmin, b0, strict);
        if (force_exact) cls = EXACT_CLASS;
        if (row0 + lane >= Et) cls = 0;
        const int top = __reduce_max_sync(0xffffffffu, cls < EXACT_CLASS ? cls : 0);   // one order for the warp's series rows
        if (cls < EXACT_CLASS) {
            switch (top) {
                case 0: series_fwd_row<4, T>(r, c, b0); break;
                case 1: series_fwd_row<6, T>(r, c, b0); break;
                case 2: series_fwd_row<10, T>(r, c, b0); break;
                case 3: series_fwd_row<14, T>(r, c, b0); break;
                default: series_fwd_row<20, T>(r, c, b0); break;
            }
        }
        __syncwarp();
        unsigned exact = __ballot_sync(0xffffffffu, cls == EXACT_CLASS);
        const unsigned exact_rows = exact;
        while (exact) {                                   // whole warp on one row at a time
            const int rr = __ffs(exact) - 1;
            exact &= exact - 1;
            const float xmax = __shfl_sync(0xffffffffu, tmax, rr), xmin = __shfl_sync(0xffffffffu, tmin, rr);
            exact_fwd_row<T>(tile + rr * pitch, c, xmax, xmin, y + (row0 + rr) * ldy, y_lo ? y_lo + (row0 + rr) * ldy : nullptr, lane);
        }
        // coalesced write-back of the series rows: y_i sits in the first c elements of every staged row
        store_rows_bf16<T>(tile, pitch, c, row0, Et, exact_rows, y, ldy, y_lo, lane);
        __syncwarp();
    }
}

template <typename T>
__global__ void __launch_bounds__(128, sizeof(T) == 2 ? 4 : 2)
attention_series_bwd_kernel(const T* __restrict__ gtp, const float* __restrict__ dyn, int ld_dyn,
                            const int* __restrict__ tdst, int Ep, int Nn, long long Et, int c, bf16* __restrict__ dgtp,
                            int ld_dgtp, bf16* __restrict__ dgtp_lo, int force_exact) {
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char ats_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int pitch = 3 * c + RowT<T>::VEC;
    const size_t warp_bytes = (size_t)32 * pitch * sizeof(T) + (size_t)5 * c * sizeof(float);
    T* tile = reinterpret_cast<T*>(ats_smem_raw + (size_t)warp * warp_bytes);
    float* scratch = reinterpret_cast<float*>(ats_smem_raw + (size_t)warp * warp_bytes + (size_t)32 * pitch * sizeof(T));
    T* r = tile + lane * pitch;
    const bool strict = dgtp_lo != nullptr;
    const long long nblk = (Et + 31) / 32;
    for (long long blk = (long long)blockIdx.x * nwarps + warp; blk < nblk; blk += (long long)gridDim.x * nwarps) {
        const long long row0 = blk * 32;
        stage_rows<T>(gtp, row0, Et, c, tile, pitch, lane);
        const bool row_ok = row0 + lane < Et;
        const long long row = row_ok ? row0 + lane : Et - 1;
        const long long gi = row / Ep;
        const float* dy = dyn + (gi * Nn + __ldg(tdst + (int)(row - gi * Ep))) * ld_dyn;
        float tmax, tmin, b0;
        int cls = row_stats<T>(r, c, tmax, tmin, b0, strict);
        if (force_exact) cls = EXACT_CLASS;
        if (!row_ok) cls = 0;
        const int top = __reduce_max_sync(0xffffffffu, cls < EXACT_CLASS ? cls : 0);
        if (cls < EXACT_CLASS) {
            switch (top) {
                case 0: series_bwd_row<4, T>(r, c, b0, dy); break;
                case 1: series_bwd_row<6, T>(r, c, b0, dy); break;
                case 2: series_bwd_row<10, T>(r, c, b0, dy); break;
                case 3: series_bwd_row<14, T>(r, c, b0, dy); break;
                default: series_bwd_row<20, T>(r, c, b0, dy); break;
            }
        }
        __syncwarp();
        unsigned exact = __ballot_sync(0xffffffffu, cls == EXACT_CLASS);
        const unsigned exact_rows = exact;
        while (exact) {
            const int rr = __ffs(exact) - 1;
            exact &= exact - 1;
            const float xmax = __shfl_sync(0xffffffffu, tmax, rr), xmin = __shfl_sync(0xffffffffu, tmin, rr);
            const unsigned long long dyp = __shfl_sync(0xffffffffu, (unsigned long long)reinterpret_cast<uintptr_t>(dy), rr);
            exact_bwd_row<T>(tile + rr * pitch, c, xmax, xmin, reinterpret_cast<const float*>(dyp), scratch,
                             dgtp + (row0 + rr) * ld_dgtp, lane);
            if (dgtp_lo) {                                // fp32 mode: the exact path's bf16 rounding has no low plane
                for (int q = lane; q < 3 * c; q += 32) dgtp_lo[(row0 + rr) * ld_dgtp + q] = __float2bfloat16_rn(0.f);
            }
        }
        // coalesced write-back: (dg | dtheta | dphi) in the staged rows -> bf16 [Et, ld_dgtp]
        store_rows_bf16<T>(tile, pitch, 3 * c, row0, Et, exact_rows, dgtp, ld_dgtp, dgtp_lo, lane);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
static std::mutex g_ats_mu;
static size_t g_ats_limit[4][64];

template <typename K>
static void ats_smem_limit(K kern, int which, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    std::lock_guard<std::mutex> lk(g_ats_mu);
    if (bytes > g_ats_limit[which][dev]) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        g_ats_limit[which][dev] = bytes;
    }
}

// RPG_ATT_SERIES=0: the former exp2 kernels (A/B comparisons); RPG_ATT_EXACT=1: series kernels, every row on the exact path
bool attention_series_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("RPG_ATT_SERIES");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}
static int force_exact() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("RPG_ATT_EXACT");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v;
}

// warps per CTA and grid: as many CTAs per SM as shared memory allows (the kernels are issue-bound: warps matter)
static void ats_geometry(size_t per_warp, long long nblk, int* warps, size_t* smem, long long* grid) {
    int w = 4;
    while (w > 1 && (size_t)w * per_warp > 110 * 1024) w >>= 1;
    *warps = w;
    *smem = (size_t)w * per_warp;
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const long long per_sm = std::max<long long>(1, (long long)(227 * 1024) / (long long)(*smem + 1024));
    long long g = (nblk + w - 1) / w;
    const long long cap = (long long)sms * per_sm * 4;
    *grid = g < cap ? g : cap;
}

int attention_series_fwd(const void* gtp, int gtp_bf16, long long Et, int c, rpg_bf16* y, int ldy, rpg_bf16* y_lo, cudaStream_t s) {
    if (c % 16 || c < 16 || c > 256 || ldy % 8) return set_error(RPG_E_UNSUPPORTED, "attention_fwd (series): c must be a multiple of 16 in [16,256], ldy of 8");
    if (gtp_bf16 && y_lo) return set_error(RPG_E_ARG, "attention_fwd (series): the fp32 mode takes fp32 projections");
    const size_t esz = gtp_bf16 ? 2 : 4;
    const size_t per_warp = (size_t)32 * (3 * c * esz + 16);
    int warps; size_t smem; long long grid;
    ats_geometry(per_warp, (Et + 31) / 32, &warps, &smem, &grid);
    if (smem > 227 * 1024) return set_error(RPG_E_UNSUPPORTED, "attention_fwd (series): row too wide for shared memory");
    ProfScope prof(RPG_PROF_ATTENTION_FWD, (double)Et * c * (3.0 * esz + 2.0 + (y_lo ? 2.0 : 0.0)), s, 0.0);
    if (gtp_bf16) {
        ats_smem_limit(attention_series_fwd_kernel<bf16>, 0, smem);
        launch_pdl(attention_series_fwd_kernel<bf16>, dim3((unsigned)grid), dim3(warps * 32), smem, s, reinterpret_cast<const bf16*>(gtp),
                   Et, c, reinterpret_cast<bf16*>(y), ldy, reinterpret_cast<bf16*>(y_lo), force_exact());
    } else {
        ats_smem_limit(attention_series_fwd_kernel<float>, 1, smem);
        launch_pdl(attention_series_fwd_kernel<float>, dim3((unsigned)grid), dim3(warps * 32), smem, s, reinterpret_cast<const float*>(gtp),
                   Et, c, reinterpret_cast<bf16*>(y), ldy, reinterpret_cast<bf16*>(y_lo), force_exact());
    }
    return check_launch("attention_series_fwd_kernel");
}

int attention_series_bwd(const void* gtp, int gtp_bf16, const float* dyn, int ld_dyn, const rpg_graph_t* graph, long long Et,
                         int c, rpg_bf16* dgtp, int ld_dgtp, rpg_bf16* dgtp_lo, cudaStream_t s) {
    if (c % 16 || c < 16 || c > 256 || ld_dgtp % 8 || ld_dyn % 4)
        return set_error(RPG_E_UNSUPPORTED, "attention_bwd (series): c must be a multiple of 16 in [16,256]");
    if (gtp_bf16 && dgtp_lo) return set_error(RPG_E_ARG, "attention_bwd (series): the fp32 mode takes fp32 projections");
    const size_t esz = gtp_bf16 ? 2 : 4;
    const size_t per_warp = (size_t)32 * (3 * c * esz + 16) + (size_t)5 * c * sizeof(float);
    int warps; size_t smem; long long grid;
    ats_geometry(per_warp, (Et + 31) / 32, &warps, &smem, &grid);
    if (smem > 227 * 1024) return set_error(RPG_E_UNSUPPORTED, "attention_bwd (series): row too wide for shared memory");
    ProfScope prof(RPG_PROF_ATTENTION_BWD, (double)Et * c * (3.0 * esz + 6.0 + (dgtp_lo ? 6.0 : 0.0)) + (double)graph->G * graph->N * c * 4.0, s, 0.0);
    if (gtp_bf16) {
        ats_smem_limit(attention_series_bwd_kernel<bf16>, 2, smem);
        launch_pdl(attention_series_bwd_kernel<bf16>, dim3((unsigned)grid), dim3(warps * 32), smem, s, reinterpret_cast<const bf16*>(gtp),
                   dyn, ld_dyn, graph->dst, graph->Ep, graph->N, Et, c, reinterpret_cast<bf16*>(dgtp), ld_dgtp,
                   reinterpret_cast<bf16*>(dgtp_lo), force_exact());
    } else {
        ats_smem_limit(attention_series_bwd_kernel<float>, 3, smem);
        launch_pdl(attention_series_bwd_kernel<float>, dim3((unsigned)grid), dim3(warps * 32), smem, s, reinterpret_cast<const float*>(gtp),
                   dyn, ld_dyn, graph->dst, graph->Ep, graph->N, Et, c, reinterpret_cast<bf16*>(dgtp), ld_dgtp,
                   reinterpret_cast<bf16*>(dgtp_lo), force_exact());
    }
    return check_launch("attention_series_bwd_kernel");
}

}  // namespace rpg
