// Shared helpers for the ha2g_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HA2G_API extern "C" __attribute__((visibility("default")))

// Launch-status convention of the C ABI: 0 = ok, otherwise the cudaError_t of the failed launch.
#define HA2G_RETURN_LAST() \
    do { cudaError_t e__ = cudaPeekAtLastError(); return (int)e__; } while (0)

// The scratch arena of the stream a launcher was called on (gemm_tc2.cu): the one registered for that stream with
// ha2g_set_workspace_lane, else the default arena of ha2g_set_workspace; nullptr when it is missing or smaller than `need`.
// Launchers on one stream may all use their arena from offset 0 (stream order keeps their uses apart); launchers on
// different streams run concurrently, which is why every side stream owns an arena.
unsigned char* ha2g_ws(size_t need_bytes, cudaStream_t stream);
// The top quarter of the same arena, reserved for the partial tiles of deterministic split-K reductions (so that a GEMM's
// packed operands at the bottom and its partials never overlap).
unsigned char* ha2g_ws_top(size_t need_bytes, cudaStream_t stream);
// C[m][n] = (accumulate ? C[m][n] : 0) + bias[n] + part[0][m][n] + part[1][m][n] + ...  (index order: deterministic)
int ha2g_splitk_reduce(const float* part, int nz, int M, int N, float* C, int ldc, const float* bias, int accumulate,
                       cudaStream_t stream);

static inline int ha2g_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Grid sizing for grid-stride element-wise kernels: a multiple of the SM count (148 on B200),
// capped so tiny problems do not launch idle CTAs.
static inline int ha2g_ew_grid(int64_t n, int threads = 256, int per_thread = 4) {
    int64_t want = (n + (int64_t)threads * per_thread - 1) / ((int64_t)threads * per_thread);
    const int64_t cap = 148 * 8;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

// Exactly one lane of a CONVERGED warp.  tcgen05.mma / cp.async.bulk are uniform-datapath instructions: guarded by
// elect.sync (with warp-uniform operands) they issue directly from uniform registers; guarded by `lane == 0` the
// compiler wraps every one of them in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall (~145 cycles per instruction,
// measured on the GRU recurrence: 4 170 -> ~600 cycles for 60 MMAs).
__device__ __forceinline__ bool ha2g_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// threadIdx.x / 32 as a provably warp-uniform value (so that role branches are uniform branches)
__device__ __forceinline__ int ha2g_warp_id() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

__device__ __forceinline__ float ha2g_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum for blockDim.x <= 1024 (result valid in every thread).
__device__ __forceinline__ float block_sum(float v, float* sh /* >= 33 floats */) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        float t = lane < nw ? sh[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) sh[32] = t;
    }
    __syncthreads();
    return sh[32];
}
__device__ __forceinline__ double block_sum_d(double v, double* sh /* >= 33 doubles */) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = lane < nw ? sh[lane] : 0.0;
        t = warp_sum_d(t);
        if (lane == 0) sh[32] = t;
    }
    __syncthreads();
    return sh[32];
}

// activation codes shared by gemm epilogues and element-wise kernels
enum { HA2G_ACT_NONE = 0, HA2G_ACT_RELU = 1, HA2G_ACT_LRELU = 2, HA2G_ACT_ELU = 3, HA2G_ACT_SIGMOID = 4, HA2G_ACT_TANH = 5,
       HA2G_ACT_LRELU02 = 6 /* LeakyReLU(0.2): the FGD auto-encoder, scripts/model/motion_ae.py:21,26 */,
       HA2G_ACT_LRELU03 = 7 /* LeakyReLU(0.3): the baseline WavEncoder, scripts/model/multimodal_context_net.py:15,18,21 */ };

__device__ __forceinline__ float ha2g_act(float x, int act) {
    switch (act) {
        case HA2G_ACT_RELU: return x > 0.f ? x : 0.f;
        case HA2G_ACT_LRELU: return x > 0.f ? x : 0.01f * x;
        case HA2G_ACT_LRELU02: return x > 0.f ? x : 0.2f * x;
        case HA2G_ACT_LRELU03: return x > 0.f ? x : 0.3f * x;
        case HA2G_ACT_ELU: return x > 0.f ? x : expm1f(x);
        case HA2G_ACT_SIGMOID: return ha2g_sigmoid(x);
        case HA2G_ACT_TANH: return tanhf(x);
        default: return x;
    }
}
// derivative expressed with the activation OUTPUT y (and, for lrelu/elu, sign information in y)
__device__ __forceinline__ float ha2g_act_grad_from_out(float y, int act) {
    switch (act) {
        case HA2G_ACT_RELU: return y > 0.f ? 1.f : 0.f;
        case HA2G_ACT_LRELU: return y > 0.f ? 1.f : 0.01f;
        case HA2G_ACT_LRELU02: return y > 0.f ? 1.f : 0.2f;
        case HA2G_ACT_LRELU03: return y > 0.f ? 1.f : 0.3f;
        case HA2G_ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
        case HA2G_ACT_SIGMOID: return y * (1.f - y);
        case HA2G_ACT_TANH: return 1.f - y * y;
        default: return 1.f;
    }
}
