// Channel attention (att.py:25-30) without transcendental work: the logits of one edge row are RANK ONE,
//   f_ij = phi_i * theta_j,   y_i = sum_j softmax_j(f_ij) g_j,
// so with theta centred (beta_j = theta_j - b0, b0 = mid-range of theta; the factor exp(phi_i b0) cancels in the softmax)
//   exp(phi_i beta_j) = sum_k phi_i^k beta_j^k / k!
// is SEPARABLE in (i, j):  D_i = sum_j exp(.) = sum_k phi_i^k U_k,  N_i = sum_j exp(.) g_j = sum_k phi_i^k T_k  with the 2 (K+1)
// moments U_k = sum_j beta_j^k / k!, T_k = sum_j beta_j^k g_j / k!.  One row costs O(c K) fused multiply-adds instead of c^2
// exponentials (c = 64: 4096 MUFU.EX2 forward and again backward, which made the former kernels MUFU-bound at 200 us per
// launch); the backward is separable in the same way (moments over i).  The order K follows from the row's own range
// R = max_i |phi_i| * (max_j theta_j - min_j theta_j) / 2: the truncation error relative to the softmax denominator is
// <= R^(K+1) / (K+1)!, kept below 1e-7 (fp32 rounding level) with two extra orders for the derivative polynomials.  Rows
// with R > 3.4 (K would exceed 20) take the EXACT path -- the former warp-per-row exp2 code -- inside the same kernel, so
// the result is the reference's softmax to fp32 accuracy for every input.
//
// Mapping: one THREAD per edge row (the moments are private sums: no shuffles), rows staged through shared memory with
// 16-byte cp.async (coalesced; row pitch 3c + 4 floats keeps the per-thread float4 reads bank-conflict free), results
// written back through the same tile (coalesced).  Arithmetic is packed fp32x2 (FFMA2): the pairs (U_k, T_k), (D, N),
// (A_k, B_k) share one instruction.  Bound: HBM (3c * 4 B in, c * 2 B out per row) once K is small.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

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

// Series class of a row from its range R (see the header): 0..3 -> K = 6 / 10 / 14 / 20, 4 -> exact path.
// Bounds: R^K / K! <= 2e-8 (the order K - 1 bound: margin for the derivative polynomials of the backward); measured
// against an fp64 softmax at the class limits: truncation < 1e-9, fp32 evaluation 1e-7 (tests/test_gpu_attention.py).
__device__ __forceinline__ int series_class(float R) {
    return R <= 0.15f ? 0 : (R <= 0.75f ? 1 : (R <= 1.7f ? 2 : (R <= 3.4f ? 3 : 4)));
}

// (U_k, T_k) = (sum_j beta_j^k, sum_j beta_j^k g_j) / k!  for k = 0..KM over the staged row r = (g | theta | phi)
template <int KM>
__device__ __forceinline__ void series_moments(const float* r, int c, float b0, float2 (&UT)[KM + 1]) {
#pragma unroll
    for (int k = 0; k <= KM; ++k) UT[k] = make_float2(0.f, 0.f);
    for (int j = 0; j < c; j += 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(r + c + j);
        const float4 g4 = *reinterpret_cast<const float4*>(r + j);
        const float tt[4] = {t4.x, t4.y, t4.z, t4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float beta = tt[q] - b0;
            const float2 og = make_float2(1.f, gg[q]);
            float p = 1.f;
#pragma unroll
            for (int k = 0; k <= KM; ++k) {
                UT[k] = __ffma2_rn(b2(p), og, UT[k]);
                p *= beta;
            }
        }
    }
#pragma unroll
    for (int k = 2; k <= KM; ++k) UT[k] = __fmul2_rn(UT[k], b2(inv_fact(k)));
}

// forward of one row: y_i = N_i / D_i written over g (r[i]); the row's g values are consumed before
template <int KM>
__device__ __forceinline__ void series_fwd_row(float* r, int c, float b0) {
    float2 UT[KM + 1];
    series_moments<KM>(r, c, b0, UT);
    for (int i = 0; i < c; i += 4) {
        const float4 p4 = *reinterpret_cast<const float4*>(r + 2 * c + i);
        const float pp[4] = {p4.x, p4.y, p4.z, p4.w};
        float yy[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float2 DN = UT[KM];
#pragma unroll
            for (int k = KM - 1; k >= 0; --k) DN = __ffma2_rn(DN, b2(pp[q]), UT[k]);
            yy[q] = __fdividef(DN.y, DN.x);
        }
        *reinterpret_cast<float4*>(r + i) = make_float4(yy[0], yy[1], yy[2], yy[3]);
    }
}

// backward of one row, in place: r = (g | theta | phi) -> (dg | dtheta | dphi); dy: the gradient w.r.t. y of this row
template <int KM>
__device__ __forceinline__ void series_bwd_row(float* r, int c, float b0, const float* __restrict__ dy) {
    float2 UT[KM + 1];
    series_moments<KM>(r, c, b0, UT);
    float2 AB[KM + 2];                                   // (sum_i w_i phi_i^k, sum_i w_i y_i phi_i^k), k = 0..KM+1
#pragma unroll
    for (int k = 0; k <= KM + 1; ++k) AB[k] = make_float2(0.f, 0.f);
    for (int i = 0; i < c; i += 4) {
        const float4 p4 = *reinterpret_cast<const float4*>(r + 2 * c + i);
        const float4 d4 = __ldg(reinterpret_cast<const float4*>(dy + i));
        const float pp[4] = {p4.x, p4.y, p4.z, p4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
        float dphi[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 ph = b2(pp[q]);
            float2 DN = UT[KM], dDN = make_float2(0.f, 0.f);     // polynomial and its derivative (Horner)
#pragma unroll
            for (int k = KM - 1; k >= 0; --k) {
                dDN = __ffma2_rn(dDN, ph, DN);
                DN = __ffma2_rn(DN, ph, UT[k]);
            }
            const float invD = __fdividef(1.f, DN.x);
            const float y = DN.y * invD, w = dd[q] * invD;
            dphi[q] = w * (dDN.y - y * dDN.x);               // sum_j dl_ij theta_j (the b0 part sums to zero)
            const float2 oy = make_float2(1.f, y);
            float qv = w;
#pragma unroll
            for (int k = 0; k <= KM + 1; ++k) {
                AB[k] = __ffma2_rn(b2(qv), oy, AB[k]);
                qv *= pp[q];
            }
        }
        *reinterpret_cast<float4*>(r + 2 * c + i) = make_float4(dphi[0], dphi[1], dphi[2], dphi[3]);
    }
    // dg_j = sum_k A_k beta^k / k!;  dtheta_j = g_j sum_k A_{k+1} beta^k / k! - sum_k B_{k+1} beta^k / k!
    float C[KM + 1];
#pragma unroll
    for (int k = 0; k <= KM; ++k) {
        C[k] = AB[k].x * inv_fact(k);
        AB[k] = __fmul2_rn(AB[k + 1], b2(inv_fact(k)));      // AB[k] now holds (A_{k+1}, B_{k+1}) / k!
    }
    for (int j = 0; j < c; j += 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(r + c + j);
        const float4 g4 = *reinterpret_cast<const float4*>(r + j);
        const float tt[4] = {t4.x, t4.y, t4.z, t4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
        float dg[4], dt[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float beta = tt[q] - b0;
            float a0 = C[KM];
            float2 s = AB[KM];
#pragma unroll
            for (int k = KM - 1; k >= 0; --k) {
                a0 = fmaf(a0, beta, C[k]);
                s = __ffma2_rn(s, b2(beta), AB[k]);
            }
            dg[q] = a0;
            dt[q] = gg[q] * s.x - s.y;
        }
        *reinterpret_cast<float4*>(r + j) = make_float4(dg[0], dg[1], dg[2], dg[3]);
        *reinterpret_cast<float4*>(r + c + j) = make_float4(dt[0], dt[1], dt[2], dt[3]);
    }
}

// ---- exact path (rows whose range exceeds the series bound): one WARP per row, exp2 with the rank-1 row maximum
__device__ __forceinline__ float2 ex2_2(float2 a) { return make_float2(exp2f(a.x), exp2f(a.y)); }

__device__ void exact_fwd_row(const float* r, int c, float tmax, float tmin, bf16* __restrict__ y, bf16* __restrict__ y_lo,
                              int lane) {
    const float* sg = r;
    const float* st = r + c;
    for (int i0 = 2 * lane; i0 < c; i0 += 64) {
        const float2 ph = *reinterpret_cast<const float2*>(r + 2 * c + i0);
        const float2 P = make_float2(ph.x * LOG2E_F, ph.y * LOG2E_F);
        const float2 nM = make_float2(-(P.x >= 0.f ? P.x * tmax : P.x * tmin), -(P.y >= 0.f ? P.y * tmax : P.y * tmin));
        float2 num = make_float2(0.f, 0.f), den = num;
        for (int j = 0; j < c; ++j) {
            const float2 e = ex2_2(__ffma2_rn(P, b2(st[j]), nM));
            num = __ffma2_rn(e, b2(sg[j]), num);
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
__device__ void exact_bwd_row(const float* r, int c, float tmax, float tmin, const float* __restrict__ dy,
                              float* scratch, bf16* __restrict__ dgtp, int lane) {
    const float* sg = r;
    const float* st = r + c;
    float* sp = scratch;
    float* sm = sp + c;
    float* sw = sm + c;
    float* swp = sw + c;
    float* swyp = swp + c;
    for (int i = lane; i < c; i += 32) {
        const float phi = r[2 * c + i];
        const float P = phi * LOG2E_F;
        const float nM = -(P >= 0.f ? P * tmax : P * tmin);
        float den = 0.f, num = 0.f, at = 0.f, agt = 0.f;
        for (int j = 0; j < c; ++j) {
            const float e = exp2f(fmaf(P, st[j], nM));
            den += e; num = fmaf(e, sg[j], num); at = fmaf(e, st[j], at); agt = fmaf(e, sg[j] * st[j], agt);
        }
        const float inv = 1.f / den, yv = num * inv, w = __ldg(dy + i) * inv;
        sp[i] = P; sm[i] = nM; sw[i] = w; swp[i] = w * phi; swyp[i] = w * yv * phi;
        dgtp[2 * c + i] = __float2bfloat16_rn(w * (agt - yv * at));
    }
    __syncwarp();
    for (int j = lane; j < c; j += 32) {
        const float th = st[j];
        float dg = 0.f, sa = 0.f, sb = 0.f;
        for (int i = 0; i < c; ++i) {
            const float e = exp2f(fmaf(th, sp[i], sm[i]));
            dg = fmaf(e, sw[i], dg); sa = fmaf(e, swp[i], sa); sb = fmaf(e, swyp[i], sb);
        }
        dgtp[j] = __float2bfloat16_rn(dg);
        dgtp[c + j] = __float2bfloat16_rn(sg[j] * sa - sb);
    }
    __syncwarp();
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Stages 32 rows of gtp [Et, 3c] (row0 ..) into the warp's tile (pitch `pitch` floats); rows beyond Et are zero-filled.
__device__ __forceinline__ void stage_rows(const float* __restrict__ gtp, long long row0, long long Et, int c, float* tile,
                                           int pitch, int lane) {
    const int per_row = 3 * c / 4;                       // 16-byte chunks per row
    for (int idx = lane; idx < 32 * per_row; idx += 32) {
        const int rr = idx / per_row, q = idx - rr * per_row;
        float* dst = tile + rr * pitch + 4 * q;
        if (row0 + rr < Et) cp_async16(dst, gtp + (row0 + rr) * 3 * c + 4 * q);
        else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_wait_all();
    __syncwarp();
}

// per-thread row statistics: theta extremes and the series class
__device__ __forceinline__ int row_stats(const float* r, int c, float& tmax, float& tmin, float& b0) {
    tmax = -INFINITY; tmin = INFINITY;
    float amax = 0.f;
    for (int j = 0; j < c; j += 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(r + c + j);
        const float4 p4 = *reinterpret_cast<const float4*>(r + 2 * c + j);
        tmax = fmaxf(fmaxf(tmax, fmaxf(t4.x, t4.y)), fmaxf(t4.z, t4.w));
        tmin = fminf(fminf(tmin, fminf(t4.x, t4.y)), fminf(t4.z, t4.w));
        amax = fmaxf(fmaxf(amax, fmaxf(fabsf(p4.x), fabsf(p4.y))), fmaxf(fabsf(p4.z), fabsf(p4.w)));
    }
    b0 = 0.5f * (tmax + tmin);
    const float R = amax * 0.5f * (tmax - tmin);
    return (R == R) ? series_class(R) : 4;               // NaN / inf inputs propagate through the exact path
}

__global__ void __launch_bounds__(128)
attention_series_fwd_kernel(const float* __restrict__ gtp, long long Et, int c, bf16* __restrict__ y, int ldy,
                            bf16* __restrict__ y_lo, int force_exact) {
    pdl_prologue();
    extern __shared__ __align__(16) float ats_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int pitch = 3 * c + 4;
    float* tile = ats_smem + (size_t)warp * 32 * pitch;
    float* r = tile + lane * pitch;
    const long long nblk = (Et + 31) / 32;
    for (long long blk = (long long)blockIdx.x * nwarps + warp; blk < nblk; blk += (long long)gridDim.x * nwarps) {
        const long long row0 = blk * 32;
        stage_rows(gtp, row0, Et, c, tile, pitch, lane);
        float tmax, tmin, b0;
        int cls = row_stats(r, c, tmax, tmin, b0);
        if (force_exact) cls = 4;
        const bool row_ok = row0 + lane < Et;
        if (!row_ok) cls = 0;
        const int top = __reduce_max_sync(0xffffffffu, cls < 4 ? cls : 0);   // one order for all series rows of the warp
        if (cls < 4) {
            switch (top) {
                case 0: series_fwd_row<6>(r, c, b0); break;
                case 1: series_fwd_row<10>(r, c, b0); break;
                case 2: series_fwd_row<14>(r, c, b0); break;
                default: series_fwd_row<20>(r, c, b0); break;
            }
        }
        __syncwarp();
        unsigned exact = __ballot_sync(0xffffffffu, cls == 4);
        const unsigned exact_rows = exact;
        while (exact) {                                   // whole warp on one row at a time
            const int rr = __ffs(exact) - 1;
            exact &= exact - 1;
            const float xmax = __shfl_sync(0xffffffffu, tmax, rr), xmin = __shfl_sync(0xffffffffu, tmin, rr);
            exact_fwd_row(tile + rr * pitch, c, xmax, xmin, y + (row0 + rr) * ldy, y_lo ? y_lo + (row0 + rr) * ldy : nullptr, lane);
        }
        // coalesced write-back of the series rows: y_i sits in the first c floats of every staged row
        const int per_row = c / 8;
        for (int idx = lane; idx < 32 * per_row; idx += 32) {
            const int rr = idx / per_row, q = idx - rr * per_row;
            if (row0 + rr >= Et || ((exact_rows >> rr) & 1u)) continue;
            const float* src = tile + rr * pitch + 8 * q;
            const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
            uint4 u;
            u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w); u.z = pack_bf16x2(b.x, b.y); u.w = pack_bf16x2(b.z, b.w);
            *reinterpret_cast<uint4*>(y + (row0 + rr) * ldy + 8 * q) = u;
            if (y_lo) {
                const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                float l[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) l[t] = f[t] - __bfloat162float(__float2bfloat16_rn(f[t]));
                uint4 v;
                v.x = pack_bf16x2(l[0], l[1]); v.y = pack_bf16x2(l[2], l[3]); v.z = pack_bf16x2(l[4], l[5]); v.w = pack_bf16x2(l[6], l[7]);
                *reinterpret_cast<uint4*>(y_lo + (row0 + rr) * ldy + 8 * q) = v;
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(128)
attention_series_bwd_kernel(const float* __restrict__ gtp, const float* __restrict__ dyn, int ld_dyn,
                            const int* __restrict__ tdst, int Ep, int Nn, long long Et, int c, bf16* __restrict__ dgtp,
                            int ld_dgtp, bf16* __restrict__ dgtp_lo, int force_exact) {
    pdl_prologue();
    extern __shared__ __align__(16) float ats_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int pitch = 3 * c + 4;
    float* tile = ats_smem + (size_t)warp * (32 * pitch + 5 * c);
    float* scratch = tile + 32 * pitch;
    float* r = tile + lane * pitch;
    const long long nblk = (Et + 31) / 32;
    for (long long blk = (long long)blockIdx.x * nwarps + warp; blk < nblk; blk += (long long)gridDim.x * nwarps) {
        const long long row0 = blk * 32;
        stage_rows(gtp, row0, Et, c, tile, pitch, lane);
        const bool row_ok = row0 + lane < Et;
        const long long row = row_ok ? row0 + lane : Et - 1;
        const long long gi = row / Ep;
        const float* dy = dyn + (gi * Nn + __ldg(tdst + (int)(row - gi * Ep))) * ld_dyn;
        float tmax, tmin, b0;
        int cls = row_stats(r, c, tmax, tmin, b0);
        if (force_exact) cls = 4;
        if (!row_ok) cls = 0;
        const int top = __reduce_max_sync(0xffffffffu, cls < 4 ? cls : 0);
        if (cls < 4) {
            switch (top) {
                case 0: series_bwd_row<6>(r, c, b0, dy); break;
                case 1: series_bwd_row<10>(r, c, b0, dy); break;
                case 2: series_bwd_row<14>(r, c, b0, dy); break;
                default: series_bwd_row<20>(r, c, b0, dy); break;
            }
        }
        __syncwarp();
        unsigned exact = __ballot_sync(0xffffffffu, cls == 4);
        const unsigned exact_rows = exact;
        while (exact) {
            const int rr = __ffs(exact) - 1;
            exact &= exact - 1;
            const float xmax = __shfl_sync(0xffffffffu, tmax, rr), xmin = __shfl_sync(0xffffffffu, tmin, rr);
            const unsigned long long dyp = __shfl_sync(0xffffffffu, (unsigned long long)reinterpret_cast<uintptr_t>(dy), rr);
            exact_bwd_row(tile + rr * pitch, c, xmax, xmin, reinterpret_cast<const float*>(dyp), scratch,
                          dgtp + (row0 + rr) * ld_dgtp, lane);
            if (dgtp_lo) {                                // fp32 mode: the exact path's bf16 rounding has no low plane
                for (int q = lane; q < 3 * c; q += 32) dgtp_lo[(row0 + rr) * ld_dgtp + q] = __float2bfloat16_rn(0.f);
            }
        }
        // coalesced write-back: (dg | dtheta | dphi) fp32 in the staged rows -> bf16 [Et, ld_dgtp]
        const int per_row = 3 * c / 8;
        for (int idx = lane; idx < 32 * per_row; idx += 32) {
            const int rr = idx / per_row, q = idx - rr * per_row;
            if (row0 + rr >= Et || ((exact_rows >> rr) & 1u)) continue;
            const float* src = tile + rr * pitch + 8 * q;
            const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
            uint4 u;
            u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w); u.z = pack_bf16x2(b.x, b.y); u.w = pack_bf16x2(b.z, b.w);
            *reinterpret_cast<uint4*>(dgtp + (row0 + rr) * ld_dgtp + 8 * q) = u;
            if (dgtp_lo) {
                const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                float l[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) l[t] = f[t] - __bfloat162float(__float2bfloat16_rn(f[t]));
                uint4 v;
                v.x = pack_bf16x2(l[0], l[1]); v.y = pack_bf16x2(l[2], l[3]); v.z = pack_bf16x2(l[4], l[5]); v.w = pack_bf16x2(l[6], l[7]);
                *reinterpret_cast<uint4*>(dgtp_lo + (row0 + rr) * ld_dgtp + 8 * q) = v;
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
static std::mutex g_ats_mu;
static size_t g_ats_limit[2][64];

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

static int ats_warps(int c, size_t per_warp_bytes) {
    int w = 4;
    while (w > 1 && (size_t)w * per_warp_bytes > 200 * 1024) w >>= 1;
    return w;
}

int attention_series_fwd(const float* gtp, long long Et, int c, rpg_bf16* y, int ldy, rpg_bf16* y_lo, cudaStream_t s) {
    if (c % 16 || c < 16 || c > 256 || ldy % 8) return set_error(RPG_E_UNSUPPORTED, "attention_fwd (series): c must be a multiple of 16 in [16,256], ldy of 8");
    const size_t per_warp = (size_t)32 * (3 * c + 4) * sizeof(float);
    const int warps = ats_warps(c, per_warp);
    const size_t smem = warps * per_warp;
    if (smem > 227 * 1024) return set_error(RPG_E_UNSUPPORTED, "attention_fwd (series): row too wide for shared memory");
    ats_smem_limit(attention_series_fwd_kernel, 0, smem);
    const long long nblk = (Et + 31) / 32;
    long long grid = (nblk + warps - 1) / warps;
    if (grid > 148 * 8) grid = 148 * 8;
    ProfScope prof(RPG_PROF_ATTENTION_FWD, (double)Et * c * (12.0 + 2.0 + (y_lo ? 2.0 : 0.0)), s, 0.0);
    launch_pdl(attention_series_fwd_kernel, dim3((unsigned)grid), dim3(warps * 32), smem, s, gtp, Et, c,
               reinterpret_cast<bf16*>(y), ldy, reinterpret_cast<bf16*>(y_lo), force_exact());
    return check_launch("attention_series_fwd_kernel");
}

int attention_series_bwd(const float* gtp, const float* dyn, int ld_dyn, const rpg_graph_t* graph, long long Et, int c,
                         rpg_bf16* dgtp, int ld_dgtp, rpg_bf16* dgtp_lo, cudaStream_t s) {
    if (c % 16 || c < 16 || c > 256 || ld_dgtp % 8 || ld_dyn % 4)
        return set_error(RPG_E_UNSUPPORTED, "attention_bwd (series): c must be a multiple of 16 in [16,256]");
    const size_t per_warp = ((size_t)32 * (3 * c + 4) + 5 * c) * sizeof(float);
    const int warps = ats_warps(c, per_warp);
    const size_t smem = warps * per_warp;
    if (smem > 227 * 1024) return set_error(RPG_E_UNSUPPORTED, "attention_bwd (series): row too wide for shared memory");
    ats_smem_limit(attention_series_bwd_kernel, 1, smem);
    const long long nblk = (Et + 31) / 32;
    long long grid = (nblk + warps - 1) / warps;
    if (grid > 148 * 8) grid = 148 * 8;
    ProfScope prof(RPG_PROF_ATTENTION_BWD, (double)Et * c * (12.0 + 6.0 + (dgtp_lo ? 6.0 : 0.0)) + (double)graph->G * graph->N * c * 4.0, s, 0.0);
    launch_pdl(attention_series_bwd_kernel, dim3((unsigned)grid), dim3(warps * 32), smem, s, gtp, dyn, ld_dyn, graph->dst,
               graph->Ep, graph->N, Et, c, reinterpret_cast<bf16*>(dgtp), ld_dgtp, reinterpret_cast<bf16*>(dgtp_lo), force_exact());
    return check_launch("attention_series_bwd_kernel");
}

}  // namespace rpg
