// Internal declarations shared by the translation units of librpg_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>

#include "../../include/rpg.h"

struct CUtensorMap_st;      // <cuda.h>'s tensor-map descriptor (only rpg_gemm*.cu include that header)

namespace rpg {

// Records a message for rpg_last_error_string() (thread-local) and returns `code`.
int set_error(int code, const char* msg);
// cudaGetLastError() after a launch; returns 0 or the cudaError_t (recorded with the kernel name).
int check_launch(const char* what);

int gemm_launch(const rpg_gemm_t* g, cudaStream_t stream);
int set_gemm_cluster(int cl);
// Optional per-launch CUDA-event timing of the GEMM kernel (bench.py roofline leg); off by default.
int profile_begin();
int profile_end(double* nt_ms, double* tn_ms, int* nt_launches, int* tn_launches, double* nt_flops, double* tn_flops);

// Per-launch records for every kernel class (rpg_profile_records).  prof_open records the start event and returns a
// slot (or -1 when no window is open); prof_close records the end event.  ProfScope brackets one launch.
bool prof_active();
int prof_open(int cls, double flops, double bytes, int M, int N, int K, cudaStream_t s);
void prof_close(int slot, cudaStream_t s);
int profile_records(rpg_prof_rec_t* out, int max_records, int* n_records);
struct ProfScope {
    int slot;
    cudaStream_t s;
    ProfScope(int cls, double bytes, cudaStream_t s_, double flops = 0.0) : slot(prof_open(cls, flops, bytes, 0, 0, 0, s_)), s(s_) {}
    ~ProfScope() { if (slot >= 0) prof_close(slot, s); }
};

// Series form of the channel attention (rpg_attention.cu); the exp2 kernels of rpg_aux.cu remain for RPG_ATT_SERIES=0,
// for callers that want the row statistics (`aux`) and for c that is not a multiple of 16.
bool attention_series_enabled();
// gtp: the projections [Et, 3c], fp32 or (gtp_bf16 != 0) bf16
int attention_series_fwd(const void* gtp, int gtp_bf16, long long Et, int c, rpg_bf16* y, int ldy, rpg_bf16* y_lo, cudaStream_t s);
int attention_series_bwd(const void* gtp, int gtp_bf16, const float* dyn, int ld_dyn, const rpg_graph_t* graph, long long Et,
                         int c, rpg_bf16* dgtp, int ld_dgtp, rpg_bf16* dgtp_lo, cudaStream_t s);

// Grouped weight-gradient GEMMs on CTA pairs (rpg_gemm_tn.cu): C_i[M, N] partials = A_i^T B_i for up to TN_GROUP_MAX
// independent problems in one persistent launch.  part: fp32 [splits][M][N] (+ colsum [splits][M] when given).
constexpr int TN_GROUP_MAX = 12;
struct TnDesc {
    const rpg_bf16* A; int lda; int M;
    const rpg_bf16* B; int ldb; int N;
    long long R;
    float* part; int splits;
    float* colsum;
};
bool tn_group_supported(const rpg_bf16* A, int lda, int M, const rpg_bf16* B, int ldb, int N, long long R);
int tn_group_splits(int M, int N, long long R, int sm_count);
int tn_group_launch(const TnDesc* d, int n, cudaStream_t stream);

// cuTensorMapEncodeTiled with this library's fixed choices (rpg_gemm.cu): rank 2 or 3, bf16 or fp32 elements, 128-byte swizzle; dims / byte
// strides / box innermost first.
int tmap_encode(::CUtensorMap_st* tm, int f32, int rank, const void* base, const uint64_t* dims, const uint64_t* strides,
                const uint32_t* box);

// rpg_colsum_bf16 with an optional second output that receives the same sums (rpg_aux.cu)
int colsum_bf16_2(const rpg_bf16* v, int ldv, int64_t rows, int cols, const float* row_w, int row_w_mod, float* out,
                  float* out2, int accumulate, float* scratch, cudaStream_t s);

inline cudaStream_t as_stream(rpg_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch: every kernel of the library starts with griddepcontrol.launch_dependents +
// griddepcontrol.wait (rpg_ptx.cuh: pdl_prologue), so the next kernel's CTAs are scheduled -- and run their prologue --
// while this one drains; nothing touches global memory before the wait.  RPG_PDL=0 turns the launch attribute off.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl_enabled() && !prof_active()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace rpg
