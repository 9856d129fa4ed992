// fp32 GEMM (parity path): C[M,N] (+)= act(op(A)[M,K] * op(B)[K,N] + bias[N]).
//
// Used for every dense projection on the HA2G step in exact-fp32 mode: Linear layers
// (hierarchy_net.py:44,89-93,117-119,218-219; ResNetSE34V2.py:36,40,44,60-61), the GRU input
// projections and weight gradients (K9/K11 in SURVEY.md), TCN k=2 convs as [x(t-d) | x(t)] GEMMs.
// 128x64x16 tiles, 256 threads, 8x4 register micro-tile, guarded loads (any M,N,K, any stride),
// optional split-K (per-slice partial planes + an ordered reduction: deterministic) for the skinny weight-gradient shapes.
#include "common.cuh"
#include <cstdio>

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;

template <bool TA, bool TB>
__global__ void __launch_bounds__(NT) gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                      float* __restrict__ C, const float* __restrict__ bias,
                                                      int M, int N, int K, int lda, int ldb, int ldc,
                                                      int act, int accumulate, int k_per_split,
                                                      int kseg_len, int kseg_stride, float* __restrict__ part) {
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    const int ty = tid / 16, tx = tid % 16;  // 16 x 16 thread grid -> 8 rows x 4 cols each

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float ra[8], rb[4];
    // reduction index -> physical row of a [K][*]-stored operand (segmented views, e.g. (batch, t<T-1))
    auto krow = [&](int gk) -> size_t {
        return kseg_len > 0 ? (size_t)(gk / kseg_len) * kseg_stride + (gk % kseg_len) : (size_t)gk;
    };
    auto load_tiles = [&](int k0) {
        // A tile: BM x BK = 2048 elements, 8 per thread
        if (TA) {  // A stored [K][M]: m contiguous
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int e = tid + i * NT;
                int m = e % BM, k = e / BM;
                int gm = m0 + m, gk = k0 + k;
                ra[i] = (gm < M && gk < kend) ? A[krow(gk) * lda + gm] : 0.f;
            }
        } else {  // A stored [M][K]: k contiguous
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int e = tid + i * NT;
                int k = e % BK, m = e / BK;
                int gm = m0 + m, gk = k0 + k;
                ra[i] = (gm < M && gk < kend) ? A[(size_t)gm * lda + gk] : 0.f;
            }
        }
        if (TB) {  // op(B)[k][n] = B[n][k]: k contiguous
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int e = tid + i * NT;
                int k = e % BK, n = e / BK;
                int gn = n0 + n, gk = k0 + k;
                rb[i] = (gn < N && gk < kend) ? B[(size_t)gn * ldb + gk] : 0.f;
            }
        } else {  // B stored [K][N]: n contiguous
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int e = tid + i * NT;
                int n = e % BN, k = e / BN;
                int gn = n0 + n, gk = k0 + k;
                rb[i] = (gn < N && gk < kend) ? B[krow(gk) * ldb + gn] : 0.f;
            }
        }
    };
    auto store_tiles = [&](int buf) {
        if (TA) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { int e = tid + i * NT; As[buf][e / BM][e % BM] = ra[i]; }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) { int e = tid + i * NT; As[buf][e % BK][e / BK] = ra[i]; }
        }
        if (TB) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { int e = tid + i * NT; Bs[buf][e % BK][e / BK] = rb[i]; }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) { int e = tid + i * NT; Bs[buf][e / BN][e % BN] = rb[i]; }
        }
    };

    int buf = 0;
    if (kbeg < kend) {
        load_tiles(kbeg);
        store_tiles(0);
    }
    __syncthreads();
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        const bool has_next = (k0 + BK) < kend;
        if (has_next) load_tiles(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[8], b[4];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (has_next) store_tiles(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    // split-K: every K slice stores its tile into its own plane of `part`; ha2g_splitk_reduce adds the planes in order
    const bool split = gridDim.z > 1;
    float* pz = split ? part + (size_t)blockIdx.z * M * N : nullptr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int gm = m0 + ty * 8 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j];
            float* dst = C + (size_t)gm * ldc + gn;
            if (split) {
                pz[(size_t)gm * N + gn] = v;
            } else {
                if (bias != nullptr) v += bias[gn];
                v = ha2g_act(v, act);
                if (accumulate) v += *dst;
                *dst = v;
            }
        }
    }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ part, int nz, int M, int N, float* __restrict__ C, int ldc,
                                     const float* __restrict__ bias, int accumulate) {
    const int64_t n = (int64_t)M * N;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(e / N), c = (int)(e % N);
        float t = 0.f;
        for (int z = 0; z < nz; ++z) t += part[(size_t)z * n + e];
        if (bias != nullptr) t += bias[c];
        float* dst = C + (size_t)m * ldc + c;
        *dst = accumulate ? *dst + t : t;
    }
}

}  // namespace

int ha2g_splitk_reduce(const float* part, int nz, int M, int N, float* C, int ldc, const float* bias, int accumulate,
                       cudaStream_t stream) {
    const int64_t n = (int64_t)M * N;
    if (n <= 0) return 0;
    splitk_reduce_kernel<<<ha2g_ew_grid(n, 256, 2), 256, 0, stream>>>(part, nz, M, N, C, ldc, bias, accumulate);
    HA2G_RETURN_LAST();
}

// C-ABI.  Row-major everywhere.  transA: A is stored [K][M] (lda >= M) else [M][K] (lda >= K);
// transB: B is stored [N][K] (ldb >= K) else [K][N] (ldb >= N).  split_k > 1 requires act == 0 and a
// C that already holds the value to accumulate onto (zeros for a plain product).
//
// ha2g_gemm_f32_kseg: the reduction index kk of a [K][*]-stored operand (A when transA, B when !transB)
// maps to physical row (kk / kseg_len) * kseg_stride + kk % kseg_len  (kseg_len = 0: identity).
HA2G_API int ha2g_gemm_f32_kseg(const float* A, const float* B, float* C, const float* bias, int M, int N, int K,
                                int lda, int ldb, int ldc, int transA, int transB, int act, int accumulate,
                                int split_k, int kseg_len, int kseg_stride, cudaStream_t stream) {
    if (M <= 0 || N <= 0) return 0;
    if (split_k < 1) split_k = 1;
    if (split_k > 1 && act != 0) return (int)cudaErrorInvalidValue;
    if (split_k > 1) accumulate = 1;  // split-K always means C += A*B
    int k_per = ((K + split_k - 1) / split_k + BK - 1) / BK * BK;
    if (k_per < BK) k_per = BK;
    int nz = K > 0 ? (K + k_per - 1) / k_per : 1;
    dim3 grid(ha2g_div_up(N, BN), ha2g_div_up(M, BM), nz);
    float* part = nullptr;
    if (nz > 1) {
        part = reinterpret_cast<float*>(ha2g_ws_top((size_t)nz * M * N * sizeof(float), stream));
        if (part == nullptr) return (int)cudaErrorMemoryAllocation;
    }
#define LAUNCH(TA_, TB_) \
    gemm_f32_kernel<TA_, TB_><<<grid, NT, 0, stream>>>(A, B, C, bias, M, N, K, lda, ldb, ldc, act, accumulate, k_per, \
                                                       kseg_len, kseg_stride, part)
    if (transA) { if (transB) LAUNCH(true, true); else LAUNCH(true, false); }
    else { if (transB) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
    if (nz > 1) return ha2g_splitk_reduce(part, nz, M, N, C, ldc, bias, accumulate, stream);
    HA2G_RETURN_LAST();
}

HA2G_API int ha2g_gemm_f32(const float* A, const float* B, float* C, const float* bias, int M, int N, int K,
                           int lda, int ldb, int ldc, int transA, int transB, int act, int accumulate, int split_k,
                           cudaStream_t stream) {
    return ha2g_gemm_f32_kseg(A, B, C, bias, M, N, K, lda, ldb, ldc, transA, transB, act, accumulate, split_k, 0, 0,
                              stream);
}

// ---- implementation switch ---------------------------------------------------------------------------------------
// ha2g_gemm / ha2g_gemm_kseg are what the rest of the library calls: the bulk-copy-fed tcgen05 bf16x3 GEMM over packed
// operands (gemm_tc2.cu) for every problem big enough to amortise the packing pass, the exact SIMT kernel above for the
// small ones.  ha2g_set_gemm_impl(0) forces the SIMT kernel everywhere (exact-fp32 parity runs), 3 the tensor-core kernel
// everywhere; ha2g_set_gemm_terms(1) switches the tensor-core kernel from the fp32-accurate 3-term split to plain bf16.
extern "C" int ha2g_gemm_tc2(const float*, const float*, float*, const float*, int, int, int, int, int, int, int, int, int,
                             int, int, int, int, int, cudaStream_t);
static int g_gemm_impl = 1;
static int g_gemm_terms = 3;
static int g_gemm_log = 0;
HA2G_API int ha2g_set_gemm_log(int on) { g_gemm_log = on; return 0; }
static int gemm_impl() { return g_gemm_impl; }
static int gemm_terms() { return g_gemm_terms; }
HA2G_API int ha2g_set_gemm_impl(int impl) {
    if (impl != 0 && impl != 1 && impl != 3) return (int)cudaErrorInvalidValue;
    g_gemm_impl = impl;
    return 0;
}
HA2G_API int ha2g_set_gemm_terms(int terms) {
    g_gemm_terms = terms == 1 ? 1 : 3;
    return 0;
}
HA2G_API int ha2g_gemm_kseg(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda,
                            int ldb, int ldc, int transA, int transB, int act, int accumulate, int split_k, int kseg_len,
                            int kseg_stride, cudaStream_t stream) {
    const int impl = gemm_impl();
    const double flops = 2.0 * (double)M * (double)N * (double)K;
    if (g_gemm_log)   // ha2g_set_gemm_log(1): one line per call on stderr (tools/gemm_shapes.py aggregates them)
        fprintf(stderr, "ha2g_gemm M=%d N=%d K=%d tA=%d tB=%d act=%d acc=%d split=%d kseg=%d\n", M, N, K, transA, transB, act,
                accumulate, split_k, kseg_len);
    // Tensor-core path (bf16 hi + lo operands: 2^-18 operand representation, ~2e-6 output error) for problems that amortise
    // the packing pass: at least two 128-row tiles of output rows, or a long reduction (the weight-gradient GEMMs: few output
    // rows, K = batch x time).  Small problems stay on the exact fp32 SIMT kernel (they gain nothing from tcgen05, and the
    // B = 2..5 parity fixtures then see fp32 arithmetic end to end).
    if (impl == 3 || (impl == 1 && flops >= 6.0e7 && K >= 32 && (M >= 256 || K >= 1024)))
        return ha2g_gemm_tc2(A, B, C, bias, M, N, K, lda, ldb, ldc, transA, transB, act, accumulate, split_k, kseg_len,
                             kseg_stride, gemm_terms(), stream);
    return ha2g_gemm_f32_kseg(A, B, C, bias, M, N, K, lda, ldb, ldc, transA, transB, act, accumulate, split_k, kseg_len,
                              kseg_stride, stream);
}
HA2G_API int ha2g_gemm(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda, int ldb,
                       int ldc, int transA, int transB, int act, int accumulate, int split_k, cudaStream_t stream) {
    return ha2g_gemm_kseg(A, B, C, bias, M, N, K, lda, ldb, ldc, transA, transB, act, accumulate, split_k, 0, 0, stream);
}
