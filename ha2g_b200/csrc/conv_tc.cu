// Stride-1 NHWC convolution (forward and data gradient) on tcgen05, bf16x3 (fp32-accurate), for the ResNetSE-34 audio
// encoder (ResNetSE34V2.py:34-42,96-111; ResNetBlocks.py:12-14): 3x3/pad 1 block convolutions and the 2x2 / 3x3 un-padded
// head convolutions.  (Stride-2 convolutions, the 1-channel stem and the weight gradient stay on conv2d.cu.)
//
// Idea: pack the activation once as bf16 hi/lo in the "k-chunk-major" layout of gemm_tc2.cu over the ZERO-PADDED
// image, rows = padded pixel index q = guard + (n*Hp + hp)*Wp + wp, K = channels:  [C/8][rows][8].  In that layout
// ANY run of consecutive rows is a valid UMMA canonical K-major tile, and a filter tap (r,s) is just a row offset
// r*Wp + s.  A CTA therefore brings ONE halo block of 128 + (KH-1)*Wp + (KW-1) rows into shared memory with plain
// cp.async.bulk copies and issues the MMAs of all KH*KW taps against it by moving the start address of the A
// descriptor -- the activation is read once per tile instead of once per tap; packed weights stream through a ring.
// Outputs are produced for every padded position; the epilogue keeps the valid ones and writes fp32 NHWC.
// The data gradient is the same kernel run on the packed output gradient with flipped / transposed packed weights.
#include "common.cuh"
#include <cuda_bf16.h>
#include <cstdlib>

namespace {

constexpr int CBM = 128;    // output positions per CTA (UMMA M)
constexpr int CSUB = 4;     // K chunks (32 channels) per weight stage
constexpr int CWST = 4;     // weight ring stages
constexpr int CNT = 192;    // warps: 0 producer, 1 MMA, 2-5 epilogue

__device__ __forceinline__ uint32_t cs32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cmb_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cs32(bar)), "r"(count));
}
__device__ __forceinline__ void cmb_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = cs32(bar);
    uint32_t done = 0;
    long long t0 = clock64();
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void cmb_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cs32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cbulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(cs32(dst)), "l"(src), "r"(bytes), "r"(cs32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t cdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void csplit2(float a, float b, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<uint32_t*>(&h);
    const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xffff0000u);
    __nv_bfloat162 l = __floats2bfloat162_rn(a - ah, b - bh);
    lo = *reinterpret_cast<uint32_t*>(&l);
}
template <bool TF32>
__device__ __forceinline__ void cmma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    if (TF32)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// one 16-byte K chunk of hi / lo planes from up to 8 fp32 values:
//   bf16x3: 8 channels per chunk, hi = bf16(x), lo = bf16(x - hi)
//   tf32x3: 4 channels per chunk, hi = x with the 13 low mantissa bits cleared (a TF32 number), lo = x - hi (exact in fp32;
//           the tensor core reads its top 11 mantissa bits) -> ~2^-21 relative product error instead of 2^-16
template <bool TF32>
__device__ __forceinline__ void make_chunk(const float* v, uint4& h4, uint4& l4) {
    if (TF32) {
        uint32_t h[4];
        float l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            h[i] = __float_as_uint(v[i]) & 0xffffe000u;
            l[i] = v[i] - __uint_as_float(h[i]);
        }
        h4 = make_uint4(h[0], h[1], h[2], h[3]);
        l4 = make_uint4(__float_as_uint(l[0]), __float_as_uint(l[1]), __float_as_uint(l[2]), __float_as_uint(l[3]));
    } else {
        csplit2(v[0], v[1], h4.x, l4.x); csplit2(v[2], v[3], h4.y, l4.y);
        csplit2(v[4], v[5], h4.z, l4.z); csplit2(v[6], v[7], h4.w, l4.w);
    }
}
__device__ __forceinline__ void ccommit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(cs32(bar)) : "memory");
}

struct ConvTcGeom {
    int N, Hp, Wp, pad, KH, KW, Ho, Wo;   // padded input dims, kernel, output dims
    int guard;                            // zero rows in front of the packed activation
    int rows_pa, rows_pb;                 // packed row counts (activation, weights)
    int cs_chunks;                        // source channels / 8, rounded up to a multiple of CSUB
    int cb;                               // chunks per resident halo block (<= 16)
    int Cd;                               // destination channels (GEMM N)
    int RA;                               // halo block rows = CBM + (KH-1)*Wp + (KW-1)
};

// activation [N,H,W,C] fp32 NHWC -> packed, zero-padded by `pad`, with `guard` leading zero rows
template <bool TF32>
__global__ void pack_nhwc_kernel(const float* __restrict__ x, int N, int H, int W, int C, int pad, int guard, int rows_p,
                                 int chunks_p, uint4* __restrict__ hi, uint4* __restrict__ lo) {
    constexpr int EPC = TF32 ? 4 : 8;  // channels per 16-byte chunk
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    const int64_t total = (int64_t)rows_p * chunks_p;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(e % rows_p);
        const int c = (int)(e / rows_p);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        const int q = row - guard;
        if (q >= 0 && q < N * Hp * Wp && c * EPC < C) {
            const int wp = q % Wp, hp = (q / Wp) % Hp, n = q / (Wp * Hp);
            const int h = hp - pad, w = wp - pad;
            if (h >= 0 && h < H && w >= 0 && w < W) {
                const float* p = x + (((size_t)n * H + h) * W + w) * C + c * EPC;
                const float4 a = *reinterpret_cast<const float4*>(p);          // C % EPC == 0 for every layer on this path
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
                if (!TF32) {
                    const float4 b = *reinterpret_cast<const float4*>(p + 4);
                    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                }
            }
        }
        uint4 h4, l4;
        make_chunk<TF32>(v, h4, l4);
        hi[e] = h4;
        lo[e] = l4;
    }
}

// weights OIHW [Cout][Cin][KH][KW] -> packed [rows][K]:
//   dgrad == 0: rows = co, K index = (r*KW + s)*Cs_p + ci          (Cs = Cin)
//   dgrad != 0: rows = ci, K index = (r'*KW + s')*Cs_p + co, r' = KH-1-r, s' = KW-1-s (flipped), (Cs = Cout)
template <bool TF32>
__global__ void pack_conv_w_kernel(const float* __restrict__ w, int Cout, int Cin, int KH, int KW, int dgrad, int rows_p,
                                   int cs_chunks, uint4* __restrict__ hi, uint4* __restrict__ lo) {
    constexpr int EPC = TF32 ? 4 : 8;
    const int taps = KH * KW;
    const int64_t total = (int64_t)rows_p * taps * cs_chunks;
    const int R = dgrad ? Cin : Cout, Cs = dgrad ? Cout : Cin;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(e % rows_p);
        const int kc = (int)(e / rows_p);        // K chunk = tap * cs_chunks + c
        const int tap = kc / cs_chunks, c = kc % cs_chunks;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        if (row < R) {
            int r = tap / KW, s = tap % KW;
            if (dgrad) { r = KH - 1 - r; s = KW - 1 - s; }
#pragma unroll
            for (int i = 0; i < EPC; ++i) {
                const int cs = c * EPC + i;
                if (cs < Cs) {
                    const int co = dgrad ? cs : row, ci = dgrad ? row : cs;
                    v[i] = w[(((size_t)co * Cin + ci) * KH + r) * KW + s];
                }
            }
        }
        uint4 h4, l4;
        make_chunk<TF32>(v, h4, l4);
        hi[e] = h4;
        lo[e] = l4;
    }
}

// shared memory: A_hi [cb][RA][16B] | A_lo | W ring [CWST][hi: CSUB*BN*16 | lo] | barriers
template <int BN, bool TF32>
__global__ void __launch_bounds__(CNT) conv_tc_kernel(const uint4* __restrict__ a_hi, const uint4* __restrict__ a_lo,
                                                      const uint4* __restrict__ b_hi, const uint4* __restrict__ b_lo,
                                                      const float* __restrict__ bias, float* __restrict__ out, ConvTcGeom g) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int a_bytes = g.cb * g.RA * 16;                      // one of hi / lo
    constexpr int W_HALF = CSUB * BN * 16;
    unsigned char* sa_hi = smem;
    unsigned char* sa_lo = smem + a_bytes;
    unsigned char* sw = smem + 2 * a_bytes;
    uint64_t* bar_afull = reinterpret_cast<uint64_t*>(sw + CWST * 2 * W_HALF);
    uint64_t* bar_aempty = bar_afull + 1;
    uint64_t* bar_wfull = bar_aempty + 1;       // [CWST]
    uint64_t* bar_wempty = bar_wfull + CWST;    // [CWST]
    uint64_t* bar_done = bar_wempty + CWST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);

    const int tid = threadIdx.x, warp = ha2g_warp_id(), lane = tid & 31;
    const int q0 = g.guard + blockIdx.y * CBM;          // first padded output position (packed row index) of this tile
    const int n0 = blockIdx.x * BN;
    const int taps = g.KH * g.KW;
    const int n_ablk = g.cs_chunks / g.cb;              // halo blocks along the channel dimension
    const int subs = g.cb / CSUB;                       // weight stages per (halo block, tap)
    const int min_off = -(g.pad * g.Wp + g.pad);        // row offset of tap (0,0) relative to the output position

    if (tid == 0) {
        cmb_init(bar_afull, 1); cmb_init(bar_aempty, 1); cmb_init(bar_done, 1);
        for (int i = 0; i < CWST; ++i) { cmb_init(bar_wfull + i, 1); cmb_init(bar_wempty + i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(cs32(tmem_slot)), "r"(BN < 32 ? 32 : BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        if (ha2g_elect_one()) {
            int wit = 0;
            for (int ab = 0; ab < n_ablk; ++ab) {
                if (ab > 0) cmb_wait(bar_aempty, (ab - 1) & 1);
                cmb_expect(bar_afull, (uint32_t)(2 * a_bytes));
                for (int c = 0; c < g.cb; ++c) {
                    const size_t src = (size_t)(ab * g.cb + c) * g.rows_pa + (size_t)(q0 + min_off);
                    cbulk(sa_hi + (size_t)c * g.RA * 16, a_hi + src, (uint32_t)(g.RA * 16), bar_afull);
                    cbulk(sa_lo + (size_t)c * g.RA * 16, a_lo + src, (uint32_t)(g.RA * 16), bar_afull);
                }
                for (int tap = 0; tap < taps; ++tap) {
                    for (int sb = 0; sb < subs; ++sb, ++wit) {
                        const int st = wit % CWST;
                        if (wit >= CWST) cmb_wait(bar_wempty + st, ((wit / CWST) - 1) & 1);
                        unsigned char* w_hi = sw + (size_t)st * 2 * W_HALF;
                        unsigned char* w_lo = w_hi + W_HALF;
                        cmb_expect(bar_wfull + st, 2 * W_HALF);
#pragma unroll
                        for (int c = 0; c < CSUB; ++c) {
                            const size_t kc = (size_t)tap * g.cs_chunks + (size_t)ab * g.cb + sb * CSUB + c;
                            cbulk(w_hi + c * (BN * 16), b_hi + kc * g.rows_pb + n0, BN * 16, bar_wfull + st);
                            cbulk(w_lo + c * (BN * 16), b_lo + kc * g.rows_pb + n0, BN * 16, bar_wfull + st);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (ha2g_elect_one()) {
            const uint32_t fmt = TF32 ? 2u : 1u;  // instruction-descriptor A/B format: 1 = BF16, 2 = TF32
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(CBM >> 4) << 24);
            const uint32_t lbo_a = (uint32_t)(g.RA * 16);
            int wit = 0;
            uint32_t first = 1;
            for (int ab = 0; ab < n_ablk; ++ab) {
                cmb_wait(bar_afull, ab & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // descriptors differ only in their 14-bit start-address field (units of 16 bytes): build them once and add
                // offsets -- with N <= 64 an MMA is 16-32 cycles of tensor work, so the issue loop must not cost more
                const uint64_t da_hi0 = cdesc(cs32(sa_hi), lbo_a, 128), da_lo0 = cdesc(cs32(sa_lo), lbo_a, 128);
                const uint64_t dw0 = cdesc(cs32(sw), BN * 16, 128);
                for (int tap = 0; tap < taps; ++tap) {
                    const int roff = (tap / g.KW) * g.Wp + (tap % g.KW);   // tap row offset inside the halo block
                    for (int sb = 0; sb < subs; ++sb, ++wit) {
                        const int st = wit % CWST;
                        cmb_wait(bar_wfull + st, (wit / CWST) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t dwh = dw0 + (uint64_t)(st * ((2 * W_HALF) >> 4)), dwl = dwh + (uint64_t)(W_HALF >> 4);
                        const uint64_t a_add = (uint64_t)((sb * CSUB) * g.RA + roff);
#pragma unroll
                        for (int ks = 0; ks < CSUB / 2; ++ks) {
                            const uint64_t dah = da_hi0 + a_add + (uint64_t)(ks * 2 * g.RA), dal = da_lo0 + a_add + (uint64_t)(ks * 2 * g.RA);
                            const uint64_t dbh = dwh + (uint64_t)(ks * 2 * BN), dbl = dwl + (uint64_t)(ks * 2 * BN);
                            cmma<TF32>(tmem_d, dah, dbh, idesc, first ? 0u : 1u);
                            first = 0;
                            cmma<TF32>(tmem_d, dah, dbl, idesc, 1u);
                            cmma<TF32>(tmem_d, dal, dbh, idesc, 1u);
                        }
                        ccommit(bar_wempty + st);
                    }
                }
                ccommit(bar_aempty);
            }
            ccommit(bar_done);
        }
    } else {
        const int qd = warp & 3;
        const int q = q0 + qd * 32 + lane - g.guard;      // padded position (without guard)
        cmb_wait(bar_done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        bool valid = q >= 0 && q < g.N * g.Hp * g.Wp;
        int n = 0, ho = 0, wo = 0;
        if (valid) {
            const int wp = q % g.Wp, hp = (q / g.Wp) % g.Hp;
            n = q / (g.Wp * g.Hp);
            ho = hp - g.pad; wo = wp - g.pad;
            valid = ho >= 0 && ho < g.Ho && wo >= 0 && wo < g.Wo;
        }
        // The accumulator arrives one output PIXEL per thread.  Each warp transposes its 32 pixels x 32 channels through
        // shared memory (halo block and weight ring are idle once bar_done has fired) so that 8 consecutive lanes write the
        // 128 contiguous bytes of one pixel instead of every lane writing to a different pixel.
        constexpr int SP = 36;   // staging pitch (floats): 16-byte aligned rows, conflict-free 128-bit accesses
        float* stg = reinterpret_cast<float*>(smem) + (size_t)(warp - 2) * 32 * SP;
        long long* optr = reinterpret_cast<long long*>(smem + (size_t)4 * 32 * SP * 4) + (warp - 2) * 32;
        optr[lane] = valid ? (long long)((((size_t)n * g.Ho + ho) * g.Wo + wo) * g.Cd) : -1ll;
        __syncwarp();
        const int rr4 = lane >> 3, c4 = lane & 7;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(qd * 32) << 16) + (uint32_t)c0;
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
            if (n0 + c0 >= g.Cd) continue;   // warp-uniform
#pragma unroll
            for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<uint4*>(stg + (size_t)lane * SP + i) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
            __syncwarp();
            const int col = n0 + c0 + c4 * 4;
            if (col < g.Cd) {   // Cd % 4 == 0
                float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bias != nullptr) bv = make_float4(bias[col], bias[col + 1], bias[col + 2], bias[col + 3]);
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rl = it * 4 + rr4;
                    const long long off = optr[rl];
                    if (off >= 0) {
                        float4 v = *reinterpret_cast<const float4*>(stg + (size_t)rl * SP + c4 * 4);
                        v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                        *reinterpret_cast<float4*>(out + off + col) = v;
                    }
                }
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(BN < 32 ? 32 : BN));
}

// ---- persistent variant ------------------------------------------------------------------------------------------
// One CTA per SM walks the output tiles (128 positions each) round-robin.  Compared with one short-lived CTA per tile
// (barrier init + TMEM allocation + an exposed halo-block load latency + an exposed epilogue for ~3 us of tensor work):
//   * halo blocks are double-buffered (block = 8 channel chunks), so the next tile's activation is in flight while the
//     current one is multiplied; the weight ring simply keeps streaming across tiles;
//   * two TMEM accumulators alternate, so the epilogue of tile i (TMEM -> shared transpose -> coalesced stores) runs
//     under the MMAs of tile i+1.
// shared memory: A[2] (hi | lo) | W ring | epilogue staging + pixel offsets | barriers
constexpr int CPB = 8;      // channel chunks per resident halo block in the persistent kernel

template <int BN, bool TF32>
__global__ void __launch_bounds__(CNT, 1) conv_tc_persist_kernel(const uint4* __restrict__ a_hi, const uint4* __restrict__ a_lo,
                                                                 const uint4* __restrict__ b_hi, const uint4* __restrict__ b_lo,
                                                                 const float* __restrict__ bias, float* __restrict__ out,
                                                                 ConvTcGeom g, int tiles) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int a_bytes = g.cb * g.RA * 16;                      // one of hi / lo of one halo block
    constexpr int W_HALF = CSUB * BN * 16;
    constexpr int SP = 36;
    constexpr int TCOLS = 2 * BN < 32 ? 32 : 2 * BN;
    unsigned char* sa = smem;                                  // [2][hi | lo]
    unsigned char* sw = smem + 4 * (size_t)a_bytes;
    float* stg_base = reinterpret_cast<float*>(sw + CWST * 2 * W_HALF);
    long long* optr_base = reinterpret_cast<long long*>(stg_base + 4 * 32 * SP);
    uint64_t* bar_afull = reinterpret_cast<uint64_t*>(optr_base + 4 * 32);   // [2]
    uint64_t* bar_aempty = bar_afull + 2;       // [2]
    uint64_t* bar_wfull = bar_aempty + 2;       // [CWST]
    uint64_t* bar_wempty = bar_wfull + CWST;    // [CWST]
    uint64_t* bar_cfull = bar_wempty + CWST;    // [2] accumulator ready
    uint64_t* bar_cempty = bar_cfull + 2;       // [2] accumulator drained (4 epilogue warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_cempty + 2);

    const int tid = threadIdx.x, warp = ha2g_warp_id(), lane = tid & 31;
    const int n0 = blockIdx.x * BN;
    const int taps = g.KH * g.KW;
    const int n_ablk = g.cs_chunks / g.cb;
    const int subs = g.cb / CSUB;
    const int min_off = -(g.pad * g.Wp + g.pad);

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { cmb_init(bar_afull + i, 1); cmb_init(bar_aempty + i, 1); cmb_init(bar_cfull + i, 1); cmb_init(bar_cempty + i, 4); }
        for (int i = 0; i < CWST; ++i) { cmb_init(bar_wfull + i, 1); cmb_init(bar_wempty + i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(cs32(tmem_slot)), "r"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        if (ha2g_elect_one()) {
            int wit = 0, ai = 0;
            for (int tile = blockIdx.y; tile < tiles; tile += gridDim.y) {
                const int q0 = g.guard + tile * CBM;
                for (int ab = 0; ab < n_ablk; ++ab, ++ai) {
                    const int buf = ai & 1;
                    if (ai >= 2) cmb_wait(bar_aempty + buf, ((ai >> 1) - 1) & 1);
                    unsigned char* sa_hi = sa + (size_t)buf * 2 * a_bytes;
                    unsigned char* sa_lo = sa_hi + a_bytes;
                    cmb_expect(bar_afull + buf, (uint32_t)(2 * a_bytes));
                    for (int c = 0; c < g.cb; ++c) {
                        const size_t src = (size_t)(ab * g.cb + c) * g.rows_pa + (size_t)(q0 + min_off);
                        cbulk(sa_hi + (size_t)c * g.RA * 16, a_hi + src, (uint32_t)(g.RA * 16), bar_afull + buf);
                        cbulk(sa_lo + (size_t)c * g.RA * 16, a_lo + src, (uint32_t)(g.RA * 16), bar_afull + buf);
                    }
                    for (int tap = 0; tap < taps; ++tap) {
                        for (int sb = 0; sb < subs; ++sb, ++wit) {
                            const int st = wit % CWST;
                            if (wit >= CWST) cmb_wait(bar_wempty + st, ((wit / CWST) - 1) & 1);
                            unsigned char* w_hi = sw + (size_t)st * 2 * W_HALF;
                            unsigned char* w_lo = w_hi + W_HALF;
                            cmb_expect(bar_wfull + st, 2 * W_HALF);
#pragma unroll
                            for (int c = 0; c < CSUB; ++c) {
                                const size_t kc = (size_t)tap * g.cs_chunks + (size_t)ab * g.cb + sb * CSUB + c;
                                cbulk(w_hi + c * (BN * 16), b_hi + kc * g.rows_pb + n0, BN * 16, bar_wfull + st);
                                cbulk(w_lo + c * (BN * 16), b_lo + kc * g.rows_pb + n0, BN * 16, bar_wfull + st);
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (ha2g_elect_one()) {
            const uint32_t fmt = TF32 ? 2u : 1u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(CBM >> 4) << 24);
            const uint32_t lbo_a = (uint32_t)(g.RA * 16);
            int wit = 0, ai = 0, ti = 0;
            for (int tile = blockIdx.y; tile < tiles; tile += gridDim.y, ++ti) {
                const int cb_ = ti & 1;
                if (ti >= 2) cmb_wait(bar_cempty + cb_, ((ti >> 1) - 1) & 1);   // the epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t dcol = tmem_d + (uint32_t)(cb_ * BN);
                uint32_t first = 1;
                for (int ab = 0; ab < n_ablk; ++ab, ++ai) {
                    const int buf = ai & 1;
                    cmb_wait(bar_afull + buf, (ai >> 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa_hi = cs32(sa + (size_t)buf * 2 * a_bytes), sa_lo = sa_hi + a_bytes;
                    const uint64_t da_hi0 = cdesc(sa_hi, lbo_a, 128), da_lo0 = cdesc(sa_lo, lbo_a, 128);
                    const uint64_t dw0 = cdesc(cs32(sw), BN * 16, 128);
                    for (int tap = 0; tap < taps; ++tap) {
                        const int roff = (tap / g.KW) * g.Wp + (tap % g.KW);
                        for (int sb = 0; sb < subs; ++sb, ++wit) {
                            const int st = wit % CWST;
                            cmb_wait(bar_wfull + st, (wit / CWST) & 1);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            const uint64_t dwh = dw0 + (uint64_t)(st * ((2 * W_HALF) >> 4)), dwl = dwh + (uint64_t)(W_HALF >> 4);
                            const uint64_t a_add = (uint64_t)((sb * CSUB) * g.RA + roff);
#pragma unroll
                            for (int ks = 0; ks < CSUB / 2; ++ks) {
                                const uint64_t dah = da_hi0 + a_add + (uint64_t)(ks * 2 * g.RA), dal = da_lo0 + a_add + (uint64_t)(ks * 2 * g.RA);
                                const uint64_t dbh = dwh + (uint64_t)(ks * 2 * BN), dbl = dwl + (uint64_t)(ks * 2 * BN);
                                cmma<TF32>(dcol, dah, dbh, idesc, first ? 0u : 1u);
                                first = 0;
                                cmma<TF32>(dcol, dah, dbl, idesc, 1u);
                                cmma<TF32>(dcol, dal, dbh, idesc, 1u);
                            }
                            ccommit(bar_wempty + st);
                        }
                    }
                    ccommit(bar_aempty + buf);
                }
                ccommit(bar_cfull + cb_);
            }
        }
        __syncwarp();
    } else {
        const int qd = warp & 3;
        float* stg = stg_base + (size_t)(warp - 2) * 32 * SP;
        long long* optr = optr_base + (warp - 2) * 32;
        const int rr4 = lane >> 3, c4 = lane & 7;
        int ti = 0;
        for (int tile = blockIdx.y; tile < tiles; tile += gridDim.y, ++ti) {
            const int cb_ = ti & 1;
            const int q = tile * CBM + qd * 32 + lane;      // padded position (without guard)
            bool valid = q < g.N * g.Hp * g.Wp;
            int n = 0, ho = 0, wo = 0;
            if (valid) {
                const int wp = q % g.Wp, hp = (q / g.Wp) % g.Hp;
                n = q / (g.Wp * g.Hp);
                ho = hp - g.pad; wo = wp - g.pad;
                valid = ho >= 0 && ho < g.Ho && wo >= 0 && wo < g.Wo;
            }
            __syncwarp();
            optr[lane] = valid ? (long long)((((size_t)n * g.Ho + ho) * g.Wo + wo) * g.Cd) : -1ll;
            __syncwarp();
            cmb_wait(bar_cfull + cb_, (ti >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem_d + ((uint32_t)(qd * 32) << 16) + (uint32_t)(cb_ * BN + c0);
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
                if (c0 + 32 >= BN) {   // the accumulator is in registers: hand it back to the MMA warp (one arrival per warp)
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cs32(bar_cempty + cb_)) : "memory");
                }
                if (n0 + c0 >= g.Cd) continue;   // warp-uniform
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<uint4*>(stg + (size_t)lane * SP + i) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
                __syncwarp();
                const int col = n0 + c0 + c4 * 4;
                if (col < g.Cd) {
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (bias != nullptr) bv = make_float4(bias[col], bias[col + 1], bias[col + 2], bias[col + 3]);
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rl = it * 4 + rr4;
                        const long long off = optr[rl];
                        if (off >= 0) {
                            float4 v = *reinterpret_cast<const float4*>(stg + (size_t)rl * SP + c4 * 4);
                            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                            *reinterpret_cast<float4*>(out + off + col) = v;
                        }
                    }
                }
                __syncwarp();
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TCOLS));
}

static inline int cround(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

// Packed geometry of a stride-1 convolution input [N,H,W,C] padded by `pad` for a KHxKW kernel:
//   guard = pad*Wp + pad leading zero rows, rows_p (multiple of 128, with tail room for the last tile's halo),
//   chunks_p = C/8 rounded up to a multiple of 4.
//   prec: 0 = bf16x3 (8 channels per 16-byte chunk), 1 = tf32x3 (4 channels per chunk)
HA2G_API int ha2g_conv_tc_dims(int N, int H, int W, int C, int pad, int KH, int KW, int prec, int* guard, int* rows_p,
                               int* chunks_p) {
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    const int epc = prec ? 4 : 8;
    *guard = pad * Wp + pad;
    const int tiles = (N * Hp * Wp + CBM - 1) / CBM;
    *rows_p = cround(*guard + tiles * CBM + (KH - 1) * Wp + (KW - 1) + 1, 128);
    *chunks_p = cround((C + epc - 1) / epc, CSUB);
    return 0;
}

// x [N,H,W,C] fp32 NHWC (C % 8 == 0) -> packed bf16 hi/lo (each rows_p * chunks_p * 16 bytes, see ha2g_conv_tc_dims)
HA2G_API int ha2g_conv_tc_pack_act(const float* x, int N, int H, int W, int C, int pad, int KH, int KW, int prec, void* hi,
                                   void* lo, cudaStream_t stream) {
    if (C % (prec ? 4 : 8) != 0) return (int)cudaErrorInvalidValue;
    int guard, rows_p, chunks_p;
    ha2g_conv_tc_dims(N, H, W, C, pad, KH, KW, prec, &guard, &rows_p, &chunks_p);
    const int64_t total = (int64_t)rows_p * chunks_p;
    if (prec)
        pack_nhwc_kernel<true><<<ha2g_ew_grid(total, 256, 2), 256, 0, stream>>>(x, N, H, W, C, pad, guard, rows_p, chunks_p,
                                                                                reinterpret_cast<uint4*>(hi), reinterpret_cast<uint4*>(lo));
    else
        pack_nhwc_kernel<false><<<ha2g_ew_grid(total, 256, 2), 256, 0, stream>>>(x, N, H, W, C, pad, guard, rows_p, chunks_p,
                                                                                 reinterpret_cast<uint4*>(hi), reinterpret_cast<uint4*>(lo));
    HA2G_RETURN_LAST();
}

// w OIHW -> packed weights; dgrad == 0: rows = Cout (rows_p = round256), K = KH*KW*cs_chunks*8 with Cs = Cin;
// dgrad != 0: rows = Cin, flipped taps, Cs = Cout.  Each of hi / lo: rows_p * KH*KW*cs_chunks * 16 bytes.
HA2G_API int ha2g_conv_tc_pack_w(const float* w, int Cout, int Cin, int KH, int KW, int dgrad, int prec, void* hi, void* lo,
                                 cudaStream_t stream) {
    const int R = dgrad ? Cin : Cout, Cs = dgrad ? Cout : Cin;
    const int epc = prec ? 4 : 8;
    const int rows_p = cround(R, 256), cs_chunks = cround((Cs + epc - 1) / epc, CSUB);
    const int64_t total = (int64_t)rows_p * KH * KW * cs_chunks;
    if (prec)
        pack_conv_w_kernel<true><<<ha2g_ew_grid(total, 256, 2), 256, 0, stream>>>(w, Cout, Cin, KH, KW, dgrad, rows_p, cs_chunks,
                                                                                  reinterpret_cast<uint4*>(hi), reinterpret_cast<uint4*>(lo));
    else
        pack_conv_w_kernel<false><<<ha2g_ew_grid(total, 256, 2), 256, 0, stream>>>(w, Cout, Cin, KH, KW, dgrad, rows_p, cs_chunks,
                                                                                   reinterpret_cast<uint4*>(hi), reinterpret_cast<uint4*>(lo));
    HA2G_RETURN_LAST();
}

// out [N,Ho,Wo,Cd] fp32 NHWC = conv(packed activation, packed weights) (+ bias[Cd]); stride 1.
//   (N,H,W,Cs): un-padded source activation dims;  pad, KH, KW: of THIS correlation (for dgrad pass pad' = K-1-pad).
//   Cd: destination channels (rows of the packed weights).
HA2G_API int ha2g_conv_tc(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, const float* bias,
                          float* out, int N, int H, int W, int Cs, int Cd, int pad, int KH, int KW, int prec,
                          cudaStream_t stream) {
    if (Cd % 4 != 0) return (int)cudaErrorInvalidValue;
    ConvTcGeom g;
    g.N = N; g.pad = pad; g.KH = KH; g.KW = KW;
    g.Hp = H + 2 * pad; g.Wp = W + 2 * pad;
    g.Ho = g.Hp - KH + 1; g.Wo = g.Wp - KW + 1;
    int chunks_p;
    ha2g_conv_tc_dims(N, H, W, Cs, pad, KH, KW, prec, &g.guard, &g.rows_pa, &chunks_p);
    g.cs_chunks = chunks_p;
    g.cb = chunks_p > 16 ? 16 : chunks_p;
    if (chunks_p % g.cb != 0) return (int)cudaErrorInvalidValue;
    g.rows_pb = cround(Cd, 256);
    g.Cd = Cd;
    g.RA = CBM + (KH - 1) * g.Wp + (KW - 1);
    const int tiles = (N * g.Hp * g.Wp + CBM - 1) / CBM;
    const int bn = Cd > 64 ? 128 : (Cd > 32 ? 64 : 32);
    {   // persistent kernel: halo blocks of CPB chunks, double-buffered; falls through to the per-tile kernel if it cannot fit
        static int persist = -1;
        if (persist < 0) { const char* e = getenv("HA2G_CONV_PERSIST"); persist = (e != nullptr && e[0] == '0') ? 0 : 1; }
        ConvTcGeom gp = g;
        gp.cb = chunks_p > CPB ? CPB : chunks_p;
        const size_t smem_p = (size_t)4 * gp.cb * gp.RA * 16 + (size_t)CWST * 2 * CSUB * bn * 16 + (size_t)4 * 32 * 36 * 4 + 4 * 32 * 8 + 256;
        if (persist && chunks_p % gp.cb == 0 && smem_p <= 227 * 1024 && tiles >= 2) {
            const int nx = ha2g_div_up(Cd, bn);
            int gy = 148 / nx;
            if (gy < 1) gy = 1;
            if (gy > tiles) gy = tiles;
            dim3 pgrid(nx, gy);
#define CONV_PLAUNCH(BN_, TF_)                                                                                          \
            do {                                                                                                      \
                cudaError_t e_ = cudaFuncSetAttribute(conv_tc_persist_kernel<BN_, TF_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p); \
                if (e_ != cudaSuccess) return (int)e_;                                                                \
                conv_tc_persist_kernel<BN_, TF_><<<pgrid, CNT, smem_p, stream>>>(reinterpret_cast<const uint4*>(a_hi), reinterpret_cast<const uint4*>(a_lo), \
                                                                 reinterpret_cast<const uint4*>(b_hi), reinterpret_cast<const uint4*>(b_lo), \
                                                                 bias, out, gp, tiles);                               \
            } while (0)
            if (prec) { if (bn == 128) CONV_PLAUNCH(128, true); else if (bn == 64) CONV_PLAUNCH(64, true); else CONV_PLAUNCH(32, true); }
            else { if (bn == 128) CONV_PLAUNCH(128, false); else if (bn == 64) CONV_PLAUNCH(64, false); else CONV_PLAUNCH(32, false); }
#undef CONV_PLAUNCH
            HA2G_RETURN_LAST();
        }
    }
    const size_t smem = (size_t)2 * g.cb * g.RA * 16 + (size_t)CWST * 2 * CSUB * bn * 16 + 256;
    if (smem > 227 * 1024) return (int)cudaErrorInvalidValue;
    dim3 grid(ha2g_div_up(Cd, bn), tiles);
#define CONV_LAUNCH(BN_, TF_)                                                                                            \
    do {                                                                                                              \
        cudaError_t e_ = cudaFuncSetAttribute(conv_tc_kernel<BN_, TF_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e_ != cudaSuccess) return (int)e_;                                                                        \
        conv_tc_kernel<BN_, TF_><<<grid, CNT, smem, stream>>>(reinterpret_cast<const uint4*>(a_hi), reinterpret_cast<const uint4*>(a_lo), \
                                                         reinterpret_cast<const uint4*>(b_hi), reinterpret_cast<const uint4*>(b_lo), \
                                                         bias, out, g);                                               \
    } while (0)
    if (prec) { if (bn == 128) CONV_LAUNCH(128, true); else if (bn == 64) CONV_LAUNCH(64, true); else CONV_LAUNCH(32, true); }
    else { if (bn == 128) CONV_LAUNCH(128, false); else if (bn == 64) CONV_LAUNCH(64, false); else CONV_LAUNCH(32, false); }
#undef CONV_LAUNCH
    HA2G_RETURN_LAST();
}

// ---- weight gradient on the packed tensor-core GEMM ------------------------------------------------------------------
// dwf[(r*KW+s)*Cin + ci][co] = sum_p x[p shifted by tap (r,s)][ci] * dy[p][co]   (stride 1).
// The reduction runs over output pixels, so both GEMM operands are "K-leading"; they are packed block-wise over the pixel
// range (blocks sized to stay L2-resident between the packing pass and the GEMM): an im2col-gather writes the x operand
// [rows = KH*KW*Cin][K = pixels] directly in the packed bf16 hi/lo layout, dy is packed by ha2g_pack_bf16x2, and
// ha2g_gemm_packed (gemm_tc2.cu) accumulates each block into dwf with split-K.
extern "C" int ha2g_pack_bf16x2_rows(const float*, int, int, int, int, int, int, int, void*, void*, cudaStream_t);
extern "C" int ha2g_gemm_packed(const void*, const void*, int, const void*, const void*, int, float*, const float*, int, int,
                                int, int, int, int, int, int, cudaStream_t);
extern "C" int ha2g_pack_dims(int, int, int*, int*);

namespace {
__global__ void im2col_pack_kernel(const float* __restrict__ x, int N, int H, int W, int Cin, int KH, int KW, int pad, int Ho,
                                   int Wo, int64_t p_begin, int p_valid, int rows, int rows_p, int chunks_p,
                                   uint4* __restrict__ hi, uint4* __restrict__ lo) {
    const int64_t total = (int64_t)rows_p * chunks_p;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(e % rows_p);
        const int c = (int)(e / rows_p);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        if (row < rows) {
            const int tap = row / Cin, ci = row % Cin;
            const int r = tap / KW - pad, s = tap % KW - pad;
            // decode the first pixel of the chunk once (32-bit: N*Ho*Wo < 2^31), then walk (wo, ho, n) with carries
            const int pl0 = c * 8;
            const int p0 = (int)p_begin + pl0;
            int wo = p0 % Wo, t1 = p0 / Wo;
            int ho = t1 % Ho, n = t1 / Ho;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (pl0 + i < p_valid) {
                    const int ih = ho + r, iw = wo + s;
                    if (ih >= 0 && ih < H && iw >= 0 && iw < W) v[i] = x[(((size_t)n * H + ih) * W + iw) * Cin + ci];
                }
                if (++wo == Wo) { wo = 0; if (++ho == Ho) { ho = 0; ++n; } }
            }
        }
        uint4 h4, l4;
        make_chunk<false>(v, h4, l4);
        hi[e] = h4;
        lo[e] = l4;
    }
}
}  // namespace

static inline int wgrad_bn(int Cout) { return Cout >= 384 ? 256 : (Cout > 64 ? 128 : 64); }   // N tile of ha2g_gemm_packed

// workspace bytes for ha2g_conv_wgrad_tc (one pixel block of both packed operands)
HA2G_API int ha2g_conv_wgrad_tc_workspace(int Cin, int Cout, int KH, int KW, int64_t* bytes, int* block_pixels) {
    const int rows_pa = cround(KH * KW * Cin, 128), rows_pb = cround(Cout, wgrad_bn(Cout));
    int64_t px = ((int64_t)48 << 20) / ((int64_t)rows_pa * 4);   // packed x-operand block <= 48 MB (hi + lo)
    px = px / 32 * 32;
    if (px < 1024) px = 1024;
    if (px > (1 << 20)) px = 1 << 20;
    *block_pixels = (int)px;
    *bytes = (int64_t)(rows_pa + rows_pb) * (px / 8) * 32;
    return 0;
}

// dwf [KH*KW*Cin, Cout] += weight gradient of a stride-1 convolution (x [N,H,W,Cin], dy [N,Ho,Wo,Cout], NHWC fp32)
HA2G_API int ha2g_conv_wgrad_tc(const float* x, const float* dy, float* dwf, int N, int H, int W, int Cin, int Cout, int KH,
                                int KW, int pad, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
    const int Ho = H + 2 * pad - KH + 1, Wo = W + 2 * pad - KW + 1;
    const int64_t P = (int64_t)N * Ho * Wo;
    const int rows = KH * KW * Cin;
    const int rows_pa = cround(rows, 128), rows_pb = cround(Cout, wgrad_bn(Cout));
    int64_t need; int block_px;
    ha2g_conv_wgrad_tc_workspace(Cin, Cout, KH, KW, &need, &block_px);
    if (workspace == nullptr || workspace_bytes < need) return (int)cudaErrorInvalidValue;
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    const size_t a_plane = (size_t)rows_pa * (block_px / 8) * 16, b_plane = (size_t)rows_pb * (block_px / 8) * 16;
    unsigned char *ah = ws, *al = ws + a_plane, *bh = ws + 2 * a_plane, *bl = ws + 2 * a_plane + b_plane;
    const int bn = wgrad_bn(Cout);
    const int tiles = ha2g_div_up(Cout, bn) * ha2g_div_up(rows, 128);
    for (int64_t p0 = 0; p0 < P; p0 += block_px) {
        const int pv = (int)((P - p0) < block_px ? (P - p0) : block_px);
        int rp_unused, chunks_p;
        ha2g_pack_dims(Cout, pv, &rp_unused, &chunks_p);   // chunks_p = pixels/8 rounded up to a multiple of 4
        const int64_t total = (int64_t)rows_pa * chunks_p;
        im2col_pack_kernel<<<ha2g_ew_grid(total, 256, 2), 256, 0, stream>>>(x, N, H, W, Cin, KH, KW, pad, Ho, Wo, p0, pv, rows,
                                                                           rows_pa, chunks_p, reinterpret_cast<uint4*>(ah),
                                                                           reinterpret_cast<uint4*>(al));
        int rc = ha2g_pack_bf16x2_rows(dy + p0 * Cout, Cout, Cout, pv, 0, 0, 0, rows_pb, bh, bl, stream);
        if (rc != 0) return rc;
        const int stages = chunks_p / 4;
        int split = ha2g_div_up(296, tiles);
        if (split > stages / 4) split = stages / 4;
        if (split < 1) split = 1;
        rc = ha2g_gemm_packed(ah, al, rows_pa, bh, bl, rows_pb, dwf, nullptr, rows, Cout, chunks_p, Cout, 0, 1, split, 3, stream);
        if (rc != 0) return rc;
    }
    HA2G_RETURN_LAST();
}
