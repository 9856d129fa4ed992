// Bulk-copy-fed tcgen05 GEMM over pre-packed bf16 hi/lo operands (the dense-GEMM engine of the step).
//
//   pack  : fp32 matrix X[rows][K] (either memory order)  ->  P_hi, P_lo  bf16 [K/8 chunks][rows_p][8]
//           ("k-chunk-major": for one 8-wide K chunk ALL rows are contiguous at 16-byte stride, so any 128-row window
//            at any row offset is already an UMMA canonical K-major no-swizzle tile -- no tensor map needed, a plain
//            cp.async.bulk of 2 KB per chunk brings it into shared memory; rows/K are zero-padded to 128 / 32).
//   gemm  : C[M,N] (+)= act(A B^T + bias) with A = (A_hi, A_lo) [M x K], B = (B_hi, B_lo) [N x K] packed as above.
//           Warp-specialised: warp 0 = bulk-copy producer (one thread, mbarrier expect_tx), warp 1 = MMA issuer (one
//           thread; per 16-wide K step hi*hi + hi*lo + lo*hi into one TMEM accumulator, or hi*hi only in bf16 mode),
//           warps 2-5 = epilogue (tcgen05.ld 32 lanes x 32 columns, bias/activation, fp32 store; split-K slices store partial planes that are reduced in order).
//           3-stage shared-memory ring (96 KB at BN = 128 -> two CTAs per SM so one CTA's epilogue hides under the
//           other's main loop).  No thread touches operand data: the SM only issues copies and MMAs.
// This replaces the register-staged kernel of gemm_tc.cu (producer-bound at ~50 TFLOP/s: every CTA re-converted its
// operand tiles with ~4 instructions per element) for the GRU input projections / weight gradients, TCN and Linear
// layers (hierarchy_net.py:44,87-93,117-119; tcn.py:19-24).
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

constexpr int PBM = 128;   // UMMA_M
constexpr int PBK = 32;    // K per stage (4 chunks of 8)
constexpr int PNT = 192;   // 6 warps

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = s_u32(bar);
    uint32_t done = 0;
    long long t0 = clock64();
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 4000000000LL) __trap();  // never hang the GPU on a protocol bug
    }
}
__device__ __forceinline__ void mb_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s_u32(dst)), "l"(src), "r"(bytes), "r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}

__device__ __forceinline__ void split2p(float a, float b, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<uint32_t*>(&h);
    const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xffff0000u);
    __nv_bfloat162 l = __floats2bfloat162_rn(a - ah, b - bh);
    lo = *reinterpret_cast<uint32_t*>(&l);
}

// one thread per (row, 8-wide K chunk); consecutive threads -> consecutive rows (512-byte coalesced packed writes)
__global__ void pack_bf16x2_kernel(const float* __restrict__ src, int ld, int rows, int K, int kcontig, int kseg_len,
                                   int kseg_stride, uint4* __restrict__ hi, uint4* __restrict__ lo, int rows_p, int chunks_p) {
    const int64_t total = (int64_t)rows_p * chunks_p;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(e % rows_p);
        const int c = (int)(e / rows_p);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        if (row < rows) {
            const int k0 = c * 8;
            if (kcontig) {
                const float* p = src + (size_t)row * ld + k0;
                if (k0 + 7 < K && (ld & 3) == 0 && (((uintptr_t)src & 15) == 0)) {
                    const float4 a = *reinterpret_cast<const float4*>(p);
                    const float4 b = *reinterpret_cast<const float4*>(p + 4);
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) if (k0 + i < K) v[i] = p[i];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int k = k0 + i;
                    if (k < K) {
                        const size_t prow = kseg_len > 0 ? (size_t)(k / kseg_len) * kseg_stride + (k % kseg_len) : (size_t)k;
                        v[i] = src[prow * ld + row];
                    }
                }
            }
        }
        uint4 h, l;
        split2p(v[0], v[1], h.x, l.x); split2p(v[2], v[3], h.y, l.y);
        split2p(v[4], v[5], h.z, l.z); split2p(v[6], v[7], h.w, l.w);
        hi[e] = h;
        lo[e] = l;
    }
}

// Both operands of one GEMM in a single launch: CTAs [0, ctas_a) pack A, the rest pack B (same per-element code as
// pack_bf16x2_kernel; halves the launch count of the packing passes, ~950 -> ~480 per training step).
struct PackJob {
    const float* src; int ld, rows, K, kcontig, kseg_len, kseg_stride; uint4* hi; uint4* lo; int rows_p, chunks_p;
};
__device__ __forceinline__ void pack_elems(const PackJob& j, int64_t first, int64_t stride) {
    const int64_t total = (int64_t)j.rows_p * j.chunks_p;
    for (int64_t e = first; e < total; e += stride) {
        const int row = (int)(e % j.rows_p);
        const int c = (int)(e / j.rows_p);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        if (row < j.rows) {
            const int k0 = c * 8;
            if (j.kcontig) {
                const float* p = j.src + (size_t)row * j.ld + k0;
                if (k0 + 7 < j.K && (j.ld & 3) == 0 && (((uintptr_t)j.src & 15) == 0)) {
                    const float4 a = *reinterpret_cast<const float4*>(p);
                    const float4 b = *reinterpret_cast<const float4*>(p + 4);
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) if (k0 + i < j.K) v[i] = p[i];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int k = k0 + i;
                    if (k < j.K) {
                        const size_t prow = j.kseg_len > 0 ? (size_t)(k / j.kseg_len) * j.kseg_stride + (k % j.kseg_len) : (size_t)k;
                        v[i] = j.src[prow * j.ld + row];
                    }
                }
            }
        }
        uint4 h, l;
        split2p(v[0], v[1], h.x, l.x); split2p(v[2], v[3], h.y, l.y);
        split2p(v[4], v[5], h.z, l.z); split2p(v[6], v[7], h.w, l.w);
        j.hi[e] = h;
        j.lo[e] = l;
    }
}
__global__ void pack_pair_kernel(PackJob a, PackJob b, int ctas_a) {
    if ((int)blockIdx.x < ctas_a) pack_elems(a, blockIdx.x * (int64_t)blockDim.x + threadIdx.x, (int64_t)ctas_a * blockDim.x);
    else pack_elems(b, (blockIdx.x - ctas_a) * (int64_t)blockDim.x + threadIdx.x, (int64_t)(gridDim.x - ctas_a) * blockDim.x);
}

template <int BN>
struct PSmem {
    static constexpr int A_BYTES = PBM * PBK * 2;
    static constexpr int B_BYTES = BN * PBK * 2;
    static constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;
    // bytes in flight per SM are what hide the ~2 us loaded L2 latency: 192 KB either as 2 CTAs x 96 KB (BN <= 128) or as
    // one CTA with four 48 KB stages (BN = 256, twice the MMA work per byte moved)
    static constexpr int NSTAGE = BN == 256 ? 4 : (BN == 128 ? 3 : 4);
    static constexpr int TOTAL = NSTAGE * STAGE + 256;
};

template <int BN, int TERMS>
__global__ void __launch_bounds__(PNT) gemm_packed_kernel(const uint4* __restrict__ a_hi, const uint4* __restrict__ a_lo,
                                                          int rows_pa, const uint4* __restrict__ b_hi,
                                                          const uint4* __restrict__ b_lo, int rows_pb,
                                                          float* __restrict__ C, const float* __restrict__ bias, int M,
                                                          int N, int chunks_p, int ldc, int act, int accumulate,
                                                          int stages_per_split, float* __restrict__ part) {
    extern __shared__ __align__(128) unsigned char smem[];
    using S = PSmem<BN>;
    constexpr int PST = S::NSTAGE;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + PST * S::STAGE);
    uint64_t* bar_empty = bar_full + PST;
    uint64_t* bar_done = bar_empty + PST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);

    const int tid = threadIdx.x, warp = ha2g_warp_id(), lane = tid & 31;
    const int m0 = blockIdx.y * PBM, n0 = blockIdx.x * BN;
    const int total_stages = chunks_p / (PBK / 8);
    const int sb = blockIdx.z * stages_per_split;
    const int nkb = max(0, min(total_stages, sb + stages_per_split) - sb);

    if (tid == 0) {
        for (int i = 0; i < PST; ++i) { mb_init(bar_full + i, 1); mb_init(bar_empty + i, 1); }
        mb_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        // ===== producer: one thread streams packed chunks with 1-D bulk copies =====
        if (ha2g_elect_one()) {
            constexpr uint32_t bytes = (uint32_t)((TERMS == 3 ? 2 : 1) * (S::A_BYTES + S::B_BYTES));
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % PST;
                if (kb >= PST) mb_wait(bar_empty + st, ((kb / PST) - 1) & 1);
                unsigned char* sa_hi = smem + st * S::STAGE;
                unsigned char* sa_lo = sa_hi + S::A_BYTES;
                unsigned char* sb_hi = sa_lo + S::A_BYTES;
                unsigned char* sb_lo = sb_hi + S::B_BYTES;
                mb_expect_tx(bar_full + st, bytes);
#pragma unroll
                for (int c = 0; c < PBK / 8; ++c) {
                    const size_t ch = (size_t)(sb + kb) * (PBK / 8) + c;
                    bulk_g2s(sa_hi + c * (PBM * 16), a_hi + ch * rows_pa + m0, PBM * 16, bar_full + st);
                    bulk_g2s(sb_hi + c * (BN * 16), b_hi + ch * rows_pb + n0, BN * 16, bar_full + st);
                    if (TERMS == 3) {
                        bulk_g2s(sa_lo + c * (PBM * 16), a_lo + ch * rows_pa + m0, PBM * 16, bar_full + st);
                        bulk_g2s(sb_lo + c * (BN * 16), b_lo + ch * rows_pb + n0, BN * 16, bar_full + st);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (ha2g_elect_one()) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(PBM >> 4) << 24);
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % PST;
                mb_wait(bar_full + st, (kb / PST) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa_hi = s_u32(smem + st * S::STAGE);
                const uint32_t sa_lo = sa_hi + S::A_BYTES;
                const uint32_t sb_hi = sa_lo + S::A_BYTES;
                const uint32_t sb_lo = sb_hi + S::B_BYTES;
#pragma unroll
                for (int ks = 0; ks < PBK / 16; ++ks) {
                    const uint32_t a_off = (uint32_t)(ks * 2 * (PBM * 16)), b_off = (uint32_t)(ks * 2 * (BN * 16));
                    const uint64_t dah = mk_desc(sa_hi + a_off, PBM * 16, 128), dbh = mk_desc(sb_hi + b_off, BN * 16, 128);
                    mma_bf16(tmem_d, dah, dbh, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                    if (TERMS == 3) {
                        const uint64_t dal = mk_desc(sa_lo + a_off, PBM * 16, 128), dbl = mk_desc(sb_lo + b_off, BN * 16, 128);
                        mma_bf16(tmem_d, dah, dbl, idesc, 1u);
                        mma_bf16(tmem_d, dal, dbh, idesc, 1u);
                    }
                }
                mma_commit(bar_empty + st);
                if (kb == nkb - 1) mma_commit(bar_done);
            }
        }
    } else {
        // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        if (nkb > 0) {
            mb_wait(bar_done, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        // split-K: this K slice stores its plain tile into plane blockIdx.z of `part` (ha2g_splitk_reduce then adds the
        // planes in order, with the bias): no atomics, the result does not depend on CTA scheduling
        if (gridDim.z > 1) {
            C = part + (size_t)blockIdx.z * M * N;
            ldc = N; bias = nullptr; act = 0; accumulate = 0;
        }
        // The accumulator arrives one ROW per thread (TMEM lane = row).  Writing it out that way makes every store
        // instruction touch 32 different lines; instead each warp transposes its 32 x 32 block through shared memory (the
        // operand ring is idle once bar_done has fired) and writes 128 contiguous bytes per row: 8x fewer L1 wavefronts.
        constexpr int SP = 36;   // staging row pitch in floats: 16-byte aligned rows, conflict-free 128-bit accesses
        float* stg = reinterpret_cast<float*>(smem) + (size_t)(warp - 2) * 32 * SP;
        const int rr4 = lane >> 3, c4 = lane & 7;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            if (nkb > 0) {
                const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
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
            if (n0 + c0 >= N) continue;   // warp-uniform: nothing of this block is inside C
#pragma unroll
            for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<uint4*>(stg + (size_t)lane * SP + i) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
            __syncwarp();
            const bool vec = ((ldc & 3) == 0) && (n0 + c0 + 32 <= N) && ((((uintptr_t)C) & 15) == 0) &&
                             (bias == nullptr || (((uintptr_t)bias) & 15) == 0);
            if (vec) {
                const int col = n0 + c0 + c4 * 4;
                float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bias != nullptr) bv = *reinterpret_cast<const float4*>(bias + col);
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rl = it * 4 + rr4, grow = m0 + q * 32 + rl;
                    if (grow < M) {
                        float4 v = *reinterpret_cast<const float4*>(stg + (size_t)rl * SP + c4 * 4);
                        v.x = ha2g_act(v.x + bv.x, act); v.y = ha2g_act(v.y + bv.y, act);
                        v.z = ha2g_act(v.z + bv.z, act); v.w = ha2g_act(v.w + bv.w, act);
                        float4* dst = reinterpret_cast<float4*>(C + (size_t)grow * ldc + col);
                        if (accumulate) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                        *dst = v;
                    }
                }
            } else {
                const int col = n0 + c0 + lane;
                if (col < N) {
                    const float bvs = bias != nullptr ? bias[col] : 0.f;
#pragma unroll 4
                    for (int rl = 0; rl < 32; ++rl) {
                        const int grow = m0 + q * 32 + rl;
                        if (grow < M) {
                            float v = stg[(size_t)rl * SP + lane] + bvs;
                            float* dst = C + (size_t)grow * ldc + col;
                            v = ha2g_act(v, act);
                            if (accumulate) v += *dst;
                            *dst = v;
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(BN));
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

template <int BN, int TERMS>
static int launch_packed(const void* a_hi, const void* a_lo, int rows_pa, const void* b_hi, const void* b_lo, int rows_pb,
                         float* C, const float* bias, int M, int N, int chunks_p, int ldc, int act, int accumulate,
                         int split_k, cudaStream_t stream) {
    const int total_stages = chunks_p / (PBK / 8);
    if (split_k < 1) split_k = 1;
    if (split_k > total_stages) split_k = total_stages > 0 ? total_stages : 1;
    const int per = (total_stages + split_k - 1) / split_k;
    const int nz = total_stages > 0 ? (total_stages + per - 1) / per : 1;
    dim3 grid(ha2g_div_up(N, BN), ha2g_div_up(M, PBM), nz);
    const int smem = PSmem<BN>::TOTAL;
    cudaError_t e = cudaFuncSetAttribute(gemm_packed_kernel<BN, TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    float* part = nullptr;
    if (nz > 1) {
        part = reinterpret_cast<float*>(ha2g_ws_top((size_t)nz * M * N * sizeof(float), stream));
        if (part == nullptr) return (int)cudaErrorMemoryAllocation;
    }
    gemm_packed_kernel<BN, TERMS><<<grid, PNT, smem, stream>>>(
        reinterpret_cast<const uint4*>(a_hi), reinterpret_cast<const uint4*>(a_lo), rows_pa,
        reinterpret_cast<const uint4*>(b_hi), reinterpret_cast<const uint4*>(b_lo), rows_pb, C, bias, M, N, chunks_p, ldc, act,
        accumulate, per > 0 ? per : 1, part);
    if (nz > 1) return ha2g_splitk_reduce(part, nz, M, N, C, ldc, bias, accumulate, stream);
    HA2G_RETURN_LAST();
}

}  // namespace

// Packed sizes of a [rows x K] operand: rows_p = rows rounded to 256, chunks_p = K/8 rounded to a multiple of 4;
// each of hi / lo holds rows_p * chunks_p * 16 bytes.
HA2G_API int ha2g_pack_dims(int rows, int K, int* rows_p, int* chunks_p) {
    *rows_p = round_up(rows, 256);  // 256: the widest N tile reads 256 rows of the B operand
    *chunks_p = round_up(round_up(K, 8) / 8, PBK / 8);
    return 0;
}

// fp32 -> packed bf16 hi/lo.  kcontig != 0: X[row][k] = src[row*ld + k]; else X[row][k] = src[krow(k)*ld + row] with the
// optional segmented view krow(k) = (k / kseg_len) * kseg_stride + k % kseg_len (kseg_len = 0: identity).
HA2G_API int ha2g_pack_bf16x2(const float* src, int ld, int rows, int K, int kcontig, int kseg_len, int kseg_stride,
                              void* hi, void* lo, cudaStream_t stream) {
    int rows_p, chunks_p;
    ha2g_pack_dims(rows, K, &rows_p, &chunks_p);
    const int64_t total = (int64_t)rows_p * chunks_p;
    pack_bf16x2_kernel<<<ha2g_ew_grid(total, 256, 2), 256, 0, stream>>>(src, ld, rows, K, kcontig, kseg_len, kseg_stride,
                                                                        reinterpret_cast<uint4*>(hi),
                                                                        reinterpret_cast<uint4*>(lo), rows_p, chunks_p);
    HA2G_RETURN_LAST();
}

// Same packing with an explicit row padding (a multiple of the tile that will read the operand: 128 for A, the N tile --
// 64 / 128 / 256 -- for B) instead of the conservative 256: a [32 x K] operand then writes 64 rows, not 256.
HA2G_API int ha2g_pack_bf16x2_rows(const float* src, int ld, int rows, int K, int kcontig, int kseg_len, int kseg_stride,
                                   int rows_p, void* hi, void* lo, cudaStream_t stream) {
    int rp_unused, chunks_p;
    ha2g_pack_dims(rows, K, &rp_unused, &chunks_p);
    if (rows_p < rows || rows_p % 64 != 0) return (int)cudaErrorInvalidValue;
    const int64_t total = (int64_t)rows_p * chunks_p;
    pack_bf16x2_kernel<<<ha2g_ew_grid(total, 256, 2), 256, 0, stream>>>(src, ld, rows, K, kcontig, kseg_len, kseg_stride,
                                                                        reinterpret_cast<uint4*>(hi),
                                                                        reinterpret_cast<uint4*>(lo), rows_p, chunks_p);
    HA2G_RETURN_LAST();
}

// C[M,N] (+)= act(A B^T + bias) on packed operands (see ha2g_pack_bf16x2); terms = 3: fp32-accurate bf16x3, 1: plain bf16.
HA2G_API int ha2g_gemm_packed(const void* a_hi, const void* a_lo, int rows_pa, const void* b_hi, const void* b_lo,
                              int rows_pb, float* C, const float* bias, int M, int N, int chunks_p, int ldc, int act,
                              int accumulate, int split_k, int terms, cudaStream_t stream) {
    if (M <= 0 || N <= 0) return 0;
    if (split_k > 1 && act != 0) return (int)cudaErrorInvalidValue;
    if (terms == 3) {
        if (N >= 384) return launch_packed<256, 3>(a_hi, a_lo, rows_pa, b_hi, b_lo, rows_pb, C, bias, M, N, chunks_p, ldc, act, accumulate, split_k, stream);
        if (N > 64) return launch_packed<128, 3>(a_hi, a_lo, rows_pa, b_hi, b_lo, rows_pb, C, bias, M, N, chunks_p, ldc, act, accumulate, split_k, stream);
        return launch_packed<64, 3>(a_hi, a_lo, rows_pa, b_hi, b_lo, rows_pb, C, bias, M, N, chunks_p, ldc, act, accumulate, split_k, stream);
    }
    if (N >= 384) return launch_packed<256, 1>(a_hi, a_lo, rows_pa, b_hi, b_lo, rows_pb, C, bias, M, N, chunks_p, ldc, act, accumulate, split_k, stream);
    if (N > 64) return launch_packed<128, 1>(a_hi, a_lo, rows_pa, b_hi, b_lo, rows_pb, C, bias, M, N, chunks_p, ldc, act, accumulate, split_k, stream);
    return launch_packed<64, 1>(a_hi, a_lo, rows_pa, b_hi, b_lo, rows_pb, C, bias, M, N, chunks_p, ldc, act, accumulate, split_k, stream);
}

// Scratch arena for the packing passes and the deterministic reductions: the host registers ONE device buffer
// (torch-allocated) per process; calls on the same stream reuse it safely (stream order).  The launchers allocate nothing:
// without a large enough arena they return cudaErrorMemoryAllocation.  Bottom 3/4: packed operands / reduction partials of
// one launcher call; top 1/4: split-K partial planes.
// Scratch arenas: the default one (ha2g_set_workspace) and up to three more, each tied to one side stream
// (ha2g_set_workspace_lane): launchers enqueued on a side stream run concurrently with the main stream's, so their operand
// packing and partial planes must not share bytes with it.
struct WsLane { unsigned char* ptr; size_t bytes; cudaStream_t stream; };
static WsLane g_lanes[4] = {};
static inline const WsLane& ws_lane(cudaStream_t stream) {
    for (int i = 1; i < 4; ++i)
        if (g_lanes[i].ptr != nullptr && g_lanes[i].stream == stream) return g_lanes[i];
    return g_lanes[0];
}
static inline size_t ws_top_bytes(const WsLane& l) { return (l.bytes / 4) & ~(size_t)255; }
unsigned char* ha2g_ws(size_t need_bytes, cudaStream_t stream) {
    const WsLane& l = ws_lane(stream);
    return (l.ptr != nullptr && need_bytes <= l.bytes - ws_top_bytes(l)) ? l.ptr : nullptr;
}
unsigned char* ha2g_ws_top(size_t need_bytes, cudaStream_t stream) {
    const WsLane& l = ws_lane(stream);
    return (l.ptr != nullptr && need_bytes <= ws_top_bytes(l)) ? l.ptr + (l.bytes - ws_top_bytes(l)) : nullptr;
}
HA2G_API int ha2g_set_workspace(void* ptr, int64_t bytes) {
    g_lanes[0].ptr = reinterpret_cast<unsigned char*>(ptr);
    g_lanes[0].bytes = ptr != nullptr ? (size_t)bytes : 0;
    g_lanes[0].stream = nullptr;
    return 0;
}
// lane 1..3: the arena of the launchers enqueued on `stream` (ptr = nullptr removes the lane)
HA2G_API int ha2g_set_workspace_lane(int lane, void* ptr, int64_t bytes, cudaStream_t stream) {
    if (lane < 1 || lane > 3) return (int)cudaErrorInvalidValue;
    g_lanes[lane].ptr = reinterpret_cast<unsigned char*>(ptr);
    g_lanes[lane].bytes = ptr != nullptr ? (size_t)bytes : 0;
    g_lanes[lane].stream = stream;
    return 0;
}

// Drop-in with the ha2g_gemm contract: packs both operands into the scratch arena, runs the packed GEMM.
// (terms: 3 = fp32-accurate, 1 = bf16.)
HA2G_API int ha2g_gemm_tc2(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda,
                           int ldb, int ldc, int transA, int transB, int act, int accumulate, int split_k, int kseg_len,
                           int kseg_stride, int terms, cudaStream_t stream) {
    if (M <= 0 || N <= 0) return 0;
    int rpa, rpb, cp, cp2;
    ha2g_pack_dims(M, K, &rpa, &cp);
    ha2g_pack_dims(N, K, &rpb, &cp2);
    const size_t a_bytes = (size_t)rpa * cp * 16, b_bytes = (size_t)rpb * cp * 16;
    const size_t need = 2 * (a_bytes + b_bytes);
    unsigned char* ws = ha2g_ws(need, stream);
    if (ws == nullptr) return (int)cudaErrorMemoryAllocation;   // the arena (ha2g_set_workspace) is missing or too small
    unsigned char *ah = ws, *al = ws + a_bytes, *bh = ws + 2 * a_bytes, *bl = ws + 2 * a_bytes + b_bytes;
    // skinny outputs with a long reduction (e.g. the 4032->32 head projections): too few output tiles to fill 148 SMs,
    // so split K automatically (needs a zero-initialised C and no activation)
    if (split_k <= 1 && act == 0) {
        const int bn = N >= 384 ? 256 : (N > 64 ? 128 : 64);
        const int tiles = ha2g_div_up(N, bn) * ha2g_div_up(M, PBM);
        const int stages = cp / (PBK / 8);
        if (tiles * 2 <= 148 && stages >= 16) {
            int s = 148 / tiles;
            if (s > 8) s = 8;
            if (s > stages / 4) s = stages / 4;
            if (s > 1) split_k = s;   // the ordered reduction overwrites C when accumulate == 0
        }
    }
    int rc = 0;
    {   // both operands in one launch
        PackJob ja{A, lda, M, K, transA ? 0 : 1, kseg_len, kseg_stride, reinterpret_cast<uint4*>(ah), reinterpret_cast<uint4*>(al), rpa, cp};
        PackJob jb{B, ldb, N, K, transB ? 1 : 0, kseg_len, kseg_stride, reinterpret_cast<uint4*>(bh), reinterpret_cast<uint4*>(bl), rpb, cp};
        const int ca = ha2g_ew_grid((int64_t)rpa * cp, 256, 2), cb = ha2g_ew_grid((int64_t)rpb * cp, 256, 2);
        pack_pair_kernel<<<ca + cb, 256, 0, stream>>>(ja, jb, ca);
        cudaError_t pe = cudaPeekAtLastError();
        if (pe != cudaSuccess) rc = (int)pe;
    }
    if (rc == 0) rc = ha2g_gemm_packed(ah, al, rpa, bh, bl, rpb, C, bias, M, N, cp, ldc, act, accumulate, split_k, terms, stream);
    return rc;
}
