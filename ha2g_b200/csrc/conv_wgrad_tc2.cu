// Weight gradient of the stride-1 KxK "same" convolutions of the ResNetSE-34 audio encoder on tcgen05, WITHOUT im2col
// (ResNetBlocks.py:12-14 conv1/conv2 of every SE block; ResNetSE34V2.py:96-111 head convolutions with padding):
//
//     dW[co][ci][r][s] = sum_q dy[q][co] * x[q + (r-pad)*Wp + (s-pad)][ci]        q = padded pixel index (n*Hp + hp)*Wp + wp
//
// Both tensors are packed once as bf16 hi/lo planes over the zero-padded image, [C/8][rows][8 channels] (the layout of
// conv_tc.cu; dy is packed with the SAME padding so that q indexes both).  Read with MN-major UMMA descriptors, one such
// plane set is a [K = pixels] x [MN = channels] operand in which ANY run of consecutive pixel rows is a valid tile -- so a
// filter tap is only a different start row of the x operand.  A CTA walks its share of the pixel range in blocks of 64
// rows: one bulk-copy stage holds the dy block (A, M = 128 output channels, missing channel planes are zero) and ONE halo
// block of x (B, N = input-channel tile); the MMA thread issues, for every tap of its group, 4 K-steps x 3 bf16-split terms
// into that tap's own TMEM accumulator (9 taps x 32 columns for Cin = 32; one kernel row = 3 taps x <=128 columns
// otherwise).  x and dy are each read from HBM once (plus the halo overlap from L2); nothing 9x the activation is ever
// written.  Every pixel-range split stores its partial sums into its own plane of the scratch arena; ha2g_splitk_reduce
// adds the planes in order (deterministic).
// Replaces conv2d_wgrad_kernel (fp32 FMA, 1.47 ms per layer-1 convolution at B = 128) and the im2col + packed-GEMM path
// (0.65 ms per convolution, 9x activation blow-up through HBM).
#include "common.cuh"
#include <cuda_bf16.h>

extern "C" int ha2g_conv_tc_dims(int, int, int, int, int, int, int, int, int*, int*, int*);
extern "C" int ha2g_conv_tc_pack_act(const float*, int, int, int, int, int, int, int, int, void*, void*, cudaStream_t);

namespace {

constexpr int WKB = 64;      // pixel rows (GEMM K) per stage
constexpr int WST = 3;       // stages
constexpr int WNT = 192;     // warps: 0 producer, 1 MMA, 2-5 epilogue
constexpr int WM = 128;      // UMMA M = output-channel tile (16 channel planes)

struct WgGeom {
    int Q;            // padded pixels N*Hp*Wp
    int guard, rows_p, Wp, KH, KW, pad;
    int Cin, Cout;
    int TG;           // taps per CTA: KH*KW (all) or KW (one kernel row)
    int n_groups;     // KH*KW / TG
    int NT;           // input-channel tile (UMMA N)
    int RA;           // x halo rows per stage
    int nblk, per;    // pixel blocks in total / per CTA
    int a_stage, b_stage;   // bytes of one hi (or lo) operand stage
};

__device__ __forceinline__ uint32_t ws32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wmb_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ws32(bar)), "r"(count));
}
__device__ __forceinline__ void wmb_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = ws32(bar);
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
__device__ __forceinline__ void wmb_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ws32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wbulk(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(ws32(bar)) : "memory");
}
// MN-major, no swizzle: core matrix = 8 K-rows x 16 bytes (8 MN elements) stored as 128 contiguous bytes;
// LBO = byte distance between consecutive 8-row K groups, SBO = byte distance between consecutive 8-element MN chunks.
__device__ __forceinline__ uint64_t wdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void wmma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void wcommit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ws32(bar)) : "memory");
}

// grid: x = pixel-range split, y = output-channel tile (128), z = tap group + n_groups * input-channel tile
__global__ void __launch_bounds__(WNT, 1) conv_wgrad_tc2_kernel(const uint4* __restrict__ x_hi, const uint4* __restrict__ x_lo,
                                                                const uint4* __restrict__ dy_hi, const uint4* __restrict__ dy_lo,
                                                                float* __restrict__ dwf, WgGeom g, int tmem_cols) {
    extern __shared__ __align__(128) unsigned char smem[];
    // stage: A_hi | A_lo | B_hi | B_lo
    const int stage_bytes = 2 * g.a_stage + 2 * g.b_stage;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + WST * stage_bytes);
    uint64_t* bar_empty = bar_full + WST;
    uint64_t* bar_done = bar_empty + WST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = ha2g_warp_id();
    const int tg = blockIdx.z % g.n_groups, nt = blockIdx.z / g.n_groups;
    const int co0 = blockIdx.y * WM, ci0 = nt * g.NT;
    const int mc_real = min(WM / 8, (g.Cout - co0) / 8);     // real output-channel planes of the A operand
    const int NC = g.NT / 8;
    const int b0 = blockIdx.x * g.per, b1 = min(g.nblk, b0 + g.per);
    const bool all_taps = g.TG == g.KH * g.KW;
    // first x row of a block's halo, relative to its q0: all taps -> -(pad*Wp + pad); one kernel row r -> (r-pad)*Wp - pad
    const int off_min = all_taps ? -(g.pad * g.Wp + g.pad) : (tg - g.pad) * g.Wp - g.pad;

    if (tid == 0) {
        for (int i = 0; i < WST; ++i) { wmb_init(bar_full + i, 1); wmb_init(bar_empty + i, 1); }
        wmb_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ws32(tmem_slot)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // output-channel planes that do not exist (Cout < 128) stay zero in every stage: the copies never touch them
    for (int st = 0; st < WST; ++st) {
        for (int hl = 0; hl < 2; ++hl) {
            uint4* base = reinterpret_cast<uint4*>(smem + st * stage_bytes + hl * g.a_stage);
            for (int e = mc_real * WKB + tid; e < (WM / 8) * WKB; e += WNT) base[e] = make_uint4(0, 0, 0, 0);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        // ===== producer: one elected lane streams the dy block and the x halo block of every pixel block =====
        if (ha2g_elect_one()) {
            const uint32_t bytes = (uint32_t)(2 * (mc_real * WKB * 16 + NC * g.RA * 16));
            for (int b = b0; b < b1; ++b) {
                const int i = b - b0, st = i % WST;
                if (i >= WST) wmb_wait(bar_empty + st, ((i / WST) - 1) & 1);
                const uint32_t sa_hi = ws32(smem + st * stage_bytes), sa_lo = sa_hi + g.a_stage;
                const uint32_t sb_hi = sa_lo + g.a_stage, sb_lo = sb_hi + g.b_stage;
                wmb_expect(bar_full + st, bytes);
                const size_t rowA = (size_t)g.guard + (size_t)b * WKB;
                const size_t rowB = rowA + off_min;
                for (int c = 0; c < mc_real; ++c) {
                    const size_t src = (size_t)(co0 / 8 + c) * g.rows_p + rowA;
                    wbulk(sa_hi + c * (WKB * 16), dy_hi + src, WKB * 16, bar_full + st);
                    wbulk(sa_lo + c * (WKB * 16), dy_lo + src, WKB * 16, bar_full + st);
                }
                for (int c = 0; c < NC; ++c) {
                    const size_t src = (size_t)(ci0 / 8 + c) * g.rows_p + rowB;
                    wbulk(sb_hi + c * (g.RA * 16), x_hi + src, (uint32_t)(g.RA * 16), bar_full + st);
                    wbulk(sb_lo + c * (g.RA * 16), x_lo + src, (uint32_t)(g.RA * 16), bar_full + st);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer: D[tap][co][ci] += dy_block^T * x_block shifted by the tap =====
        if (ha2g_elect_one()) {
            // A and B both MN-major (bits 15, 16): the packed planes are [K rows][8 channels]
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(g.NT >> 3) << 17) | ((uint32_t)(WM >> 4) << 24);
            for (int b = b0; b < b1; ++b) {
                const int i = b - b0, st = i % WST;
                wmb_wait(bar_full + st, (i / WST) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa_hi = ws32(smem + st * stage_bytes), sa_lo = sa_hi + g.a_stage;
                const uint32_t sb_hi = sa_lo + g.a_stage, sb_lo = sb_hi + g.b_stage;
                // descriptors differ only in their start-address field (units of 16 bytes = one pixel row of a chunk plane)
                const uint64_t da_hi0 = wdesc(sa_hi, 128, WKB * 16), da_lo0 = wdesc(sa_lo, 128, WKB * 16);
                const uint64_t db_hi0 = wdesc(sb_hi, 128, (uint32_t)(g.RA * 16)), db_lo0 = wdesc(sb_lo, 128, (uint32_t)(g.RA * 16));
                for (int t = 0; t < g.TG; ++t) {
                    const int toff = all_taps ? (t / g.KW) * g.Wp + (t % g.KW) : t;   // rows from the halo start
                    const uint32_t dcol = tmem_d + (uint32_t)(t * g.NT);
#pragma unroll
                    for (int ks = 0; ks < WKB / 16; ++ks) {
                        const uint64_t dah = da_hi0 + (uint64_t)(ks * 16), dal = da_lo0 + (uint64_t)(ks * 16);
                        const uint64_t dbh = db_hi0 + (uint64_t)(toff + ks * 16), dbl = db_lo0 + (uint64_t)(toff + ks * 16);
                        wmma(dcol, dah, dbh, idesc, (i > 0 || ks > 0) ? 1u : 0u);
                        wmma(dcol, dah, dbl, idesc, 1u);
                        wmma(dcol, dal, dbh, idesc, 1u);
                    }
                }
                wcommit(bar_empty + st);
            }
            wcommit(bar_done);
        }
        __syncwarp();
    } else {
        // ===== epilogue warps 2..5: lane = output channel, columns = (tap, input channel); plain stores into this pixel split's partial plane =====
        const int q = warp & 3;
        const int co = co0 + q * 32 + lane;
        if (b1 > b0) {
            wmb_wait(bar_done, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int t = 0; t < g.TG; ++t) {
                const int tap = tg * g.TG + t;
                for (int c0 = 0; c0 < g.NT; c0 += 16) {
                    uint32_t r[16];
                    const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * g.NT + c0);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (co < g.Cout) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int ci = ci0 + c0 + i;
                            dwf[(((size_t)blockIdx.x * g.KH * g.KW + tap) * g.Cin + ci) * g.Cout + co] = __uint_as_float(r[i]);
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols));
}

static inline int wround(int x, int m) { return (x + m - 1) / m * m; }

static bool wg_config(int Cin, int Cout, int KH, int KW, int* TG, int* NT) {
    if (Cin % 16 != 0 || Cout % 8 != 0) return false;
    *NT = Cin > 128 ? 128 : Cin;
    if (Cin % *NT != 0) return false;
    if (KH * KW * *NT <= 512) *TG = KH * KW;          // all taps in one CTA (Cin = 32: 9 x 32 columns)
    else if (KW * *NT <= 512) *TG = KW;               // one kernel row per CTA
    else return false;
    return true;
}

}  // namespace

// 1 through *ok when ha2g_conv_wgrad_tc2 can serve this convolution (stride 1, "same" padding), and the workspace bytes
// (packed x and dy, hi + lo) it needs.
HA2G_API int ha2g_conv_wgrad_tc2_workspace(int N, int H, int W, int Cin, int Cout, int KH, int KW, int pad, int* ok,
                                           int64_t* bytes) {
    int TG, NT;
    *ok = 0;
    *bytes = 0;
    if (KH != 2 * pad + 1 || KW != 2 * pad + 1 || !wg_config(Cin, Cout, KH, KW, &TG, &NT)) return 0;
    int guard, rows_p, cpx, cpy;
    ha2g_conv_tc_dims(N, H, W, Cin, pad, KH, KW, 0, &guard, &rows_p, &cpx);
    ha2g_conv_tc_dims(N, H, W, Cout, pad, KH, KW, 0, &guard, &rows_p, &cpy);
    *bytes = (int64_t)rows_p * (cpx + cpy) * 32;
    *ok = 1;
    return 0;
}

// dwf [KH*KW*Cin, Cout] += weight gradient of a stride-1 "same" convolution: x [N,H,W,Cin], dy [N,H,W,Cout], NHWC fp32.
// (row (r*KW+s)*Cin + ci, column co  <->  dW[co][ci][r][s] of the OIHW parameter.)
HA2G_API int ha2g_conv_wgrad_tc2(const float* x, const float* dy, float* dwf, int N, int H, int W, int Cin, int Cout, int KH,
                                 int KW, int pad, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
    int ok; int64_t need;
    ha2g_conv_wgrad_tc2_workspace(N, H, W, Cin, Cout, KH, KW, pad, &ok, &need);
    if (!ok || workspace == nullptr || workspace_bytes < need) return (int)cudaErrorInvalidValue;
    WgGeom g{};
    wg_config(Cin, Cout, KH, KW, &g.TG, &g.NT);
    int cpx, cpy;
    ha2g_conv_tc_dims(N, H, W, Cin, pad, KH, KW, 0, &g.guard, &g.rows_p, &cpx);
    ha2g_conv_tc_dims(N, H, W, Cout, pad, KH, KW, 0, &g.guard, &g.rows_p, &cpy);
    const int Hp = H + 2 * pad;
    g.Wp = W + 2 * pad; g.KH = KH; g.KW = KW; g.pad = pad; g.Cin = Cin; g.Cout = Cout;
    g.Q = N * Hp * g.Wp;
    g.n_groups = KH * KW / g.TG;
    g.RA = WKB + (g.TG == KH * KW ? (KH - 1) * g.Wp + (KW - 1) : (KW - 1));
    g.nblk = (g.Q + WKB - 1) / WKB;
    g.a_stage = (WM / 8) * WKB * 16;
    g.b_stage = (g.NT / 8) * g.RA * 16;
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    const size_t xp = (size_t)g.rows_p * cpx * 16, yp = (size_t)g.rows_p * cpy * 16;
    unsigned char *xh = ws, *xl = ws + xp, *yh = ws + 2 * xp, *yl = ws + 2 * xp + yp;
    int rc = ha2g_conv_tc_pack_act(x, N, H, W, Cin, pad, KH, KW, 0, xh, xl, stream);
    if (rc != 0) return rc;
    rc = ha2g_conv_tc_pack_act(dy, N, H, W, Cout, pad, KH, KW, 0, yh, yl, stream);
    if (rc != 0) return rc;
    const int m_tiles = (Cout + WM - 1) / WM, n_tiles = Cin / g.NT;
    const int ctas_per_split = m_tiles * n_tiles * g.n_groups;
    int split = (148 + ctas_per_split - 1) / ctas_per_split;      // ~ one CTA per SM
    if (split > g.nblk) split = g.nblk;
    g.per = (g.nblk + split - 1) / split;
    split = (g.nblk + g.per - 1) / g.per;
    int tmem_cols = 32;
    while (tmem_cols < g.TG * g.NT) tmem_cols *= 2;
    const size_t smem = (size_t)WST * (2 * g.a_stage + 2 * g.b_stage) + 256;
    if (smem > 227 * 1024 || tmem_cols > 512) return (int)cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid(split, m_tiles, g.n_groups * n_tiles);
    // every pixel split writes its own [KH*KW*Cin, Cout] plane; the planes are added onto dwf in split order (deterministic)
    const int rows = KH * KW * Cin;
    float* part = reinterpret_cast<float*>(ha2g_ws_top((size_t)split * rows * Cout * sizeof(float), stream));
    if (part == nullptr) return (int)cudaErrorMemoryAllocation;
    conv_wgrad_tc2_kernel<<<grid, WNT, smem, stream>>>(reinterpret_cast<const uint4*>(xh), reinterpret_cast<const uint4*>(xl),
                                                       reinterpret_cast<const uint4*>(yh), reinterpret_cast<const uint4*>(yl),
                                                       part, g, tmem_cols);
    cudaError_t le = cudaPeekAtLastError();
    if (le != cudaSuccess) return (int)le;
    return ha2g_splitk_reduce(part, split, rows, Cout, dwf, Cout, nullptr, 1, stream);
}
