// tcgen05 tensor-core GEMM with fp32-accurate 3-term bf16 splitting ("bf16x3"):
//     C[M,N] (+)= act(op(A)[M,K] * op(B)[K,N] + bias[N]),   fp32 in HBM, fp32 accumulation in TMEM.
//
// Every fp32 operand element x is split on the fly into hi = bf16(x), lo = bf16(x - hi) while it is staged from
// HBM into shared memory (so no separate conversion pass and no bf16 copies in HBM); the product is accumulated as
// hi*hi + hi*lo + lo*hi by three tcgen05.mma (kind::f16, M=128, N=BN, K=16) per K-step into one TMEM accumulator.
// Measured error of this scheme inside the GRU recurrence: 4e-6 relative vs 2e-3 for plain bf16, which is
// what lets the tensor-core path meet the reference's fp32 parity bar (1e-3).
//
// Structure (one 128 x BN output tile per CTA, optional split-K over grid.z):
//   128 threads stage A and B tiles (BK = 32) HBM -> registers -> shared memory in the UMMA canonical K-major
//   no-swizzle layout (8x16B core matrices), for any of the four operand transposes;
//   thread 0 issues the MMAs (single-thread tcgen05 issue) and commits them to per-stage mbarriers so that staging of
//   stage s+1.. overlaps the tensor pipe; the epilogue reads TMEM with tcgen05.ld (32 lanes x 32 columns per warp),
//   applies bias/activation and writes / accumulates fp32 rows.
// Replaces ha2g_gemm_f32 (gemm.cu) for the dense projections on the step: Linear layers, GRU input projections and
// weight gradients, TCN convolutions (hierarchy_net.py:44,89-93,117-119; tcn.py:19-24).
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

constexpr int TBM = 128;  // UMMA_M
constexpr int TBK = 32;   // K per stage = 2 MMA k-steps
constexpr int TNT = 128;  // threads
constexpr int TST = 2;    // stages (2 x 32 KB at BN=128 -> 3 CTAs/SM co-resident hide the staging latency)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    long long t0 = clock64();
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && clock64() - t0 > 4000000000LL) __trap();  // ~2 s: never hang the GPU on a protocol bug
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (distance between the two 8-element K chunks) | [32,46) SBO>>4 (distance between
//   8-row groups) | [46,48) version = 1 | [61,64) layout type = 0 (no swizzle)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor: F32 accumulate, BF16 A/B, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// hi = bf16(x), lo = bf16(x - hi) for two floats at a time (packed cvt.rn.bf16x2.f32)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<uint32_t*>(&h);
    const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xffff0000u);
    __nv_bfloat162 l = __floats2bfloat162_rn(a - ah, b - bh);
    lo = *reinterpret_cast<uint32_t*>(&l);
}

// canonical K-major no-swizzle tile of R rows x TBK columns (bf16): byte offset of element (row, k)
template <int R>
__device__ __forceinline__ uint32_t canon_off(int row, int k) {
    return (uint32_t)((k >> 3) * (R * 16) + (row >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2);
}

// One operand tile (R rows x TBK k) of a matrix stored either [rows][K] (KCONTIG) or [K][rows] is staged in two phases
// so that the global loads of stage kb+1 are in flight while the tensor core works on stage kb:
//   tile_load : HBM/L2 -> registers (all loads issued back to back; rows >= nrows / k >= kend read as 0)
//   tile_store: registers -> bf16 hi/lo -> shared memory (UMMA canonical layout)
// src(row, k) = KCONTIG ? base[row*ld + k] : base[krow(k)*ld + row]
template <int R, bool KCONTIG>
struct TileRegs {
    // KCONTIG: item = (8-row group, 16-k half); lane -> row = lane % 8, float4 c = lane / 8 of the 16 k.  R/4 float4 / thread.
    // else   : item = (4 consecutive rows, 8 k);  (R/4)*(TBK/8)/TNT items per thread, 8 float4 each.
    static constexpr int NV = KCONTIG ? (R * TBK / 4) / TNT : (((R / 4) * (TBK / 8) + TNT - 1) / TNT) * 8;
    float4 v[NV];
};

template <int R, bool KCONTIG>
__device__ __forceinline__ void tile_load(TileRegs<R, KCONTIG>& t, const float* __restrict__ base, int ld, int row0,
                                          int nrows, int k0, int kend, bool vec_ok, int kseg_len, int kseg_stride) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (KCONTIG) {
#pragma unroll
        for (int it = 0; it < TileRegs<R, KCONTIG>::NV; ++it) {
            const int witem = warp + it * (TNT / 32);             // (row group, k half)
            const int row = (witem >> 1) * 8 + (lane & 7);
            const int k = (witem & 1) * 16 + (lane >> 3) * 4;
            const int grow = row0 + row, gk = k0 + k;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (grow < nrows && gk < kend) {
                const float* p = base + (size_t)grow * ld + gk;
                if (vec_ok && gk + 3 < kend) v = *reinterpret_cast<const float4*>(p);
                else {
                    v.x = p[0];
                    if (gk + 1 < kend) v.y = p[1];
                    if (gk + 2 < kend) v.z = p[2];
                    if (gk + 3 < kend) v.w = p[3];
                }
            }
            t.v[it] = v;
        }
    } else {
        constexpr int ITEMS = (R / 4) * (TBK / 8);
#pragma unroll
        for (int it = 0; it < TileRegs<R, KCONTIG>::NV / 8; ++it) {
            const int item = tid + it * TNT;
            const int rq = item % (R / 4), ko = item / (R / 4);
            const int grow = row0 + rq * 4;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const int gk = k0 + ko * 8 + kk;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (item < ITEMS && gk < kend && grow < nrows) {
                    const size_t prow = kseg_len > 0 ? (size_t)(gk / kseg_len) * kseg_stride + (gk % kseg_len) : (size_t)gk;
                    const float* p = base + prow * ld + grow;
                    if (vec_ok && grow + 3 < nrows) v = *reinterpret_cast<const float4*>(p);
                    else {
                        v.x = p[0];
                        if (grow + 1 < nrows) v.y = p[1];
                        if (grow + 2 < nrows) v.z = p[2];
                        if (grow + 3 < nrows) v.w = p[3];
                    }
                }
                t.v[it * 8 + kk] = v;
            }
        }
    }
}

template <int R, bool KCONTIG>
__device__ __forceinline__ void tile_store(const TileRegs<R, KCONTIG>& t, unsigned char* __restrict__ hi_dst,
                                           unsigned char* __restrict__ lo_dst) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (KCONTIG) {
#pragma unroll
        for (int it = 0; it < TileRegs<R, KCONTIG>::NV; ++it) {
            const int witem = warp + it * (TNT / 32);
            const int row = (witem >> 1) * 8 + (lane & 7);
            const int k = (witem & 1) * 16 + (lane >> 3) * 4;
            uint2 hi, lo;
            split2(t.v[it].x, t.v[it].y, hi.x, lo.x);
            split2(t.v[it].z, t.v[it].w, hi.y, lo.y);
            const uint32_t off = canon_off<R>(row, k);
            *reinterpret_cast<uint2*>(hi_dst + off) = hi;
            *reinterpret_cast<uint2*>(lo_dst + off) = lo;
        }
    } else {
        constexpr int ITEMS = (R / 4) * (TBK / 8);
#pragma unroll
        for (int it = 0; it < TileRegs<R, KCONTIG>::NV / 8; ++it) {
            const int item = tid + it * TNT;
            if (item >= ITEMS) break;
            const int rq = item % (R / 4), ko = item / (R / 4);
            const float4* v = &t.v[it * 8];
            // the thread owns rows 4rq..4rq+3 x 8 k: four 16-byte rows, written in a lane-rotated order (2-way instead
            // of 8-way shared-memory bank conflicts)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int i = (jj + lane) & 3;
                float e[8];
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    e[kk] = i == 0 ? v[kk].x : (i == 1 ? v[kk].y : (i == 2 ? v[kk].z : v[kk].w));
                uint4 hi, lo;
                split2(e[0], e[1], hi.x, lo.x); split2(e[2], e[3], hi.y, lo.y);
                split2(e[4], e[5], hi.z, lo.z); split2(e[6], e[7], hi.w, lo.w);
                const uint32_t off = canon_off<R>(rq * 4 + i, ko * 8);
                *reinterpret_cast<uint4*>(hi_dst + off) = hi;
                *reinterpret_cast<uint4*>(lo_dst + off) = lo;
            }
        }
    }
}

template <int BN>
struct TcSmem {
    static constexpr int A_BYTES = TBM * TBK * 2;  // one of hi / lo
    static constexpr int B_BYTES = BN * TBK * 2;
    static constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int TOTAL = TST * STAGE + 128;  // + barriers / tmem pointer
};

template <int BN, bool TA, bool TB>
__global__ void __launch_bounds__(TNT) gemm_tc_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                      float* __restrict__ C, const float* __restrict__ bias, int M, int N,
                                                      int K, int lda, int ldb, int ldc, int act, int accumulate,
                                                      int k_per_split, int vecA, int vecB, int kseg_len, int kseg_stride) {
    extern __shared__ __align__(128) unsigned char smem[];
    using S = TcSmem<BN>;
    uint64_t* bar_empty = reinterpret_cast<uint64_t*>(smem + TST * S::STAGE);  // [TST]
    uint64_t* bar_done = bar_empty + TST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * k_per_split, kend = min(K, kbeg + k_per_split);
    const int nkb = (kend - kbeg + TBK - 1) / TBK;

    if (tid == 0) {
        for (int i = 0; i < TST; ++i) mbar_init(bar_empty + i, 1);
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // TMEM allocation: one warp, BN fp32 columns (power of two >= 32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    const uint32_t idesc = make_idesc(TBM, BN);

    TileRegs<TBM, !TA> ra;
    TileRegs<BN, TB> rb;
    // op(A)[m][k]: TA ? A[k*lda + m] : A[m*lda + k]      op(B)[k][n]: TB ? B[n*ldb + k] : B[k*ldb + n]
    if (nkb > 0) {
        tile_load<TBM, !TA>(ra, A, lda, m0, M, kbeg, kend, vecA != 0, kseg_len, kseg_stride);
        tile_load<BN, TB>(rb, B, ldb, n0, N, kbeg, kend, vecB != 0, kseg_len, kseg_stride);
    }
    for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % TST;
        unsigned char* sa_hi = smem + st * S::STAGE;
        unsigned char* sa_lo = sa_hi + S::A_BYTES;
        unsigned char* sb_hi = sa_lo + S::A_BYTES;
        unsigned char* sb_lo = sb_hi + S::B_BYTES;
        if (kb >= TST) mbar_wait(bar_empty + st, ((kb / TST) - 1) & 1);  // MMAs that read this stage have completed
        tile_store<TBM, !TA>(ra, sa_hi, sa_lo);
        tile_store<BN, TB>(rb, sb_hi, sb_lo);
        fence_async_smem();  // generic-proxy smem writes -> visible to the tensor core's async proxy
        __syncthreads();
        if (kb + 1 < nkb) {  // next stage's global loads fly while the tensor core consumes this one
            const int k1 = kbeg + (kb + 1) * TBK;
            tile_load<TBM, !TA>(ra, A, lda, m0, M, k1, kend, vecA != 0, kseg_len, kseg_stride);
            tile_load<BN, TB>(rb, B, ldb, n0, N, k1, kend, vecB != 0, kseg_len, kseg_stride);
        }
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < TBK / 16; ++ks) {
                const uint32_t a_off = (uint32_t)(ks * 2 * (TBM * 16));
                const uint32_t b_off = (uint32_t)(ks * 2 * (BN * 16));
                const uint64_t da_hi = make_desc(smem_u32(sa_hi) + a_off, TBM * 16, 128);
                const uint64_t da_lo = make_desc(smem_u32(sa_lo) + a_off, TBM * 16, 128);
                const uint64_t db_hi = make_desc(smem_u32(sb_hi) + b_off, BN * 16, 128);
                const uint64_t db_lo = make_desc(smem_u32(sb_lo) + b_off, BN * 16, 128);
                umma(tmem_d, da_hi, db_hi, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                umma(tmem_d, da_hi, db_lo, idesc, 1u);
                umma(tmem_d, da_lo, db_hi, idesc, 1u);
            }
            umma_commit(bar_empty + st);            // frees this stage when the MMAs above retire
            if (kb == nkb - 1) umma_commit(bar_done);  // accumulator complete
        }
    }
    if (nkb > 0) {
        mbar_wait(bar_done, 0);
        tc_fence_after();
    }
    // ---- epilogue: thread t owns output row m0 + t (TMEM lane t) --------------------------------------------
    const int row = m0 + tid;
    const bool split = gridDim.z > 1;
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        if (nkb > 0) {
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
        if (row < M) {
            float* crow = C + (size_t)row * ldc;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int col = n0 + c0 + i;
                if (col < N) {
                    float v = __uint_as_float(r[i]);
                    if (split) {
                        if (bias != nullptr && blockIdx.z == 0) v += bias[col];
                        atomicAdd(crow + col, v);
                    } else {
                        if (bias != nullptr) v += bias[col];
                        v = ha2g_act(v, act);
                        if (accumulate) v += crow[col];
                        crow[col] = v;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(BN));
}

template <int BN>
static int launch_tc(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda, int ldb,
                     int ldc, int transA, int transB, int act, int accumulate, int split_k, int kseg_len, int kseg_stride,
                     cudaStream_t stream) {
    if (split_k < 1) split_k = 1;
    if (split_k > 1) accumulate = 1;
    int k_per = ((K + split_k - 1) / split_k + TBK - 1) / TBK * TBK;
    if (k_per < TBK) k_per = TBK;
    const int nz = K > 0 ? (K + k_per - 1) / k_per : 1;
    dim3 grid(ha2g_div_up(N, BN), ha2g_div_up(M, TBM), nz);
    const int smem = TcSmem<BN>::TOTAL;
    const int vecA = (lda % 4 == 0) && (((uintptr_t)A & 15) == 0);
    const int vecB = (ldb % 4 == 0) && (((uintptr_t)B & 15) == 0);
#define TC_LAUNCH(TA_, TB_)                                                                                              \
    do {                                                                                                                 \
        cudaError_t e_ = cudaFuncSetAttribute(gemm_tc_kernel<BN, TA_, TB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
        if (e_ != cudaSuccess) return (int)e_;                                                                           \
        gemm_tc_kernel<BN, TA_, TB_><<<grid, TNT, smem, stream>>>(A, B, C, bias, M, N, K, lda, ldb, ldc, act, accumulate, \
                                                                  k_per, vecA, vecB, kseg_len, kseg_stride);             \
    } while (0)
    if (transA) { if (transB) TC_LAUNCH(true, true); else TC_LAUNCH(true, false); }
    else { if (transB) TC_LAUNCH(false, true); else TC_LAUNCH(false, false); }
#undef TC_LAUNCH
    HA2G_RETURN_LAST();
}

}  // namespace

// Same contract as ha2g_gemm_f32_kseg (gemm.cu) -- row-major, transA: A stored [K][M]; transB: B stored [N][K];
// split_k > 1 accumulates into C with atomics; kseg_*: segmented reduction-row view of the K-leading operands --
// executed on the 5th-generation tensor cores with bf16x3 splitting.
HA2G_API int ha2g_gemm_tc_kseg(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda,
                               int ldb, int ldc, int transA, int transB, int act, int accumulate, int split_k,
                               int kseg_len, int kseg_stride, cudaStream_t stream) {
    if (M <= 0 || N <= 0) return 0;
    if (split_k > 1 && act != 0) return (int)cudaErrorInvalidValue;
    if (N > 64) return launch_tc<128>(A, B, C, bias, M, N, K, lda, ldb, ldc, transA, transB, act, accumulate, split_k, kseg_len, kseg_stride, stream);
    if (N > 32) return launch_tc<64>(A, B, C, bias, M, N, K, lda, ldb, ldc, transA, transB, act, accumulate, split_k, kseg_len, kseg_stride, stream);
    return launch_tc<32>(A, B, C, bias, M, N, K, lda, ldb, ldc, transA, transB, act, accumulate, split_k, kseg_len, kseg_stride, stream);
}
HA2G_API int ha2g_gemm_tc(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda,
                          int ldb, int ldc, int transA, int transB, int act, int accumulate, int split_k,
                          cudaStream_t stream) {
    return ha2g_gemm_tc_kseg(A, B, C, bias, M, N, K, lda, ldb, ldc, transA, transB, act, accumulate, split_k, 0, 0, stream);
}
