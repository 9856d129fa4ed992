// Persistent cluster GRU recurrence on the 5th-generation tensor cores (forward) -- the fused GRU kernel of the step.
//
// Same decomposition as gru_cluster.cu (8-CTA cluster per (direction, 16-row batch chunk); CTA `rank` owns HSP = 40
// hidden units, W_hh slice resident in shared memory for all T steps, h exchanged through distributed shared memory),
// but the per-step product  gates^T[128 x 16] = W_slice[128 x K] * h^T[K x 16]  runs on tcgen05:
//   A (stationary)  : the CTA's 3*HSP gate rows (r | z | n, zero-padded to M = 128) x K = 8*HSP*... = 320, packed ONCE per
//                     layer as bf16 hi/lo in the UMMA canonical K-major no-swizzle layout ([K/8][128 rows][8], 160 KB);
//   B (per step)    : h_{t-1} of the 16 batch rows, bf16 hi/lo, [K/8][16 rows][8] (10 KB each), double buffered; every
//                     CTA's gate epilogue writes its 5 K-chunks straight into all 8 CTAs' buffers (16-byte DSMEM stores);
//   D               : 128 lanes x 16 fp32 columns in TMEM; three MMAs per 16-wide K step (hi*hi + hi*lo + lo*hi) keep the
//                     recurrence fp32-accurate (4e-6 vs 2e-3 for plain bf16 over 34 steps x 4 layers).
// Per step: 60 x tcgen05.mma (M=128, N=16, K=16) issued by one thread -> tcgen05.commit -> 4 epilogue warps read their TMEM
// lane quarter (tcgen05.ld), transpose through 8 KB of shared memory so that one thread holds r, z, n of 8 consecutive
// units of one batch row, apply the gate math in fp32 (fp32 master copy of the CTA's own h slice), write y / saved
// gates, and push the packed bf16 h_t slice to the cluster.  One cluster barrier per step.
// Replaces nn.GRU's recurrence at scripts/model/hierarchy_net.py:144 (H = 300) and :232 (H = 64).
#include "common.cuh"
#include <cooperative_groups.h>
#include <cuda_bf16.h>

namespace cg = cooperative_groups;

namespace {

constexpr int CL = 8;        // CTAs per cluster
constexpr int NB = 16;       // batch rows per cluster task (= UMMA N)
constexpr int TM = 128;      // UMMA M (gate rows incl. padding)
constexpr int TNT = 160;     // warp 0: MMA issue + TMEM alloc; warps 1-4: epilogue

struct TcParams {
    const float* gi;       // [M,T,2,3H]
    const float* w_hh[2];  // [3H,H]
    const float* b_hh[2];  // [3H]
    float* y;              // [M,T,2H]
    float* gates;          // [M,T,2,4H] or nullptr
    int M, T, H, HSP, n_chunks;
    long long* dbg;        // optional [T][8] clock64 samples of cluster 0 / rank 0 (phase timing), nullptr otherwise
};

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbi(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(bar)), "r"(count));
}
__device__ __forceinline__ void mbw(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = su32(bar);
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
__device__ __forceinline__ uint64_t mkd(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void mma16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void split2g(float a, float b, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<uint32_t*>(&h);
    const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xffff0000u);
    __nv_bfloat162 l = __floats2bfloat162_rn(a - ah, b - bh);
    lo = *reinterpret_cast<uint32_t*>(&l);
}

// shared memory map (bytes): A_hi | A_lo | B_hi[2] | B_lo[2] | G | hown | barrier | tmem slot
struct TcLayout {
    int kc;  // K chunks = CL * HSP / 8
    size_t a_bytes, b_bytes, off_alo, off_bhi, off_blo, off_g, off_hown, off_bar, total;
    __host__ __device__ explicit TcLayout(int HSP) {
        kc = CL * HSP / 8;
        a_bytes = (size_t)kc * TM * 16;
        b_bytes = (size_t)kc * NB * 16;
        off_alo = a_bytes;
        off_bhi = 2 * a_bytes;
        off_blo = off_bhi + 2 * b_bytes;
        off_g = off_blo + 2 * b_bytes;
        off_hown = off_g + (size_t)TM * (NB + 1) * 4;
        off_bar = off_hown + (size_t)HSP * NB * 4;
        off_bar = (off_bar + 15) / 16 * 16;
        total = off_bar + 64;
    }
};

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(TNT, 1) gru_seq_fwd_tc_kernel(TcParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
    const int dir = cluster_id & 1;
    const int H = p.H, HSP = p.HSP, T = p.T, M = p.M;
    const TcLayout L(HSP);
    const int KC = L.kc;                 // K chunks (K = 8*KC)
    const int CPC = HSP / 8;             // chunks owned per CTA
    unsigned char* a_hi = smem;
    unsigned char* a_lo = smem + L.off_alo;
    unsigned char* b_hi = smem + L.off_bhi;   // [2][KC][NB][16 B]
    unsigned char* b_lo = smem + L.off_blo;
    float* G = reinterpret_cast<float*>(smem + L.off_g);         // [TM][NB+1]
    float* hown = reinterpret_cast<float*>(smem + L.off_hown);   // [HSP][NB]
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int j0 = rank * HSP;
    const float* __restrict__ W = p.w_hh[dir];
    const float* __restrict__ b_hh = p.b_hh[dir];

    // ---- one-time: pack this CTA's W_hh rows (r | z | n of units j0..j0+HSP) as bf16 hi/lo, canonical K-major ----
    for (int e = tid; e < KC * TM; e += TNT) {
        const int row = e % TM, c = e / TM;
        const int g = row / HSP, u = row % HSP, j = j0 + u;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        if (g < 3 && j < H) {
            const float* src = W + ((size_t)g * H + j) * H + c * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) if (c * 8 + i < H) v[i] = src[i];
        }
        uint4 h4, l4;
        split2g(v[0], v[1], h4.x, l4.x); split2g(v[2], v[3], h4.y, l4.y);
        split2g(v[4], v[5], h4.z, l4.z); split2g(v[6], v[7], h4.w, l4.w);
        *reinterpret_cast<uint4*>(a_hi + ((size_t)c * TM + row) * 16) = h4;
        *reinterpret_cast<uint4*>(a_lo + ((size_t)c * TM + row) * 16) = l4;
    }
    if (tid == 0) {
        mbi(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(su32(tmem_slot)), "r"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

    // epilogue roles: warps 1..4 -> TMEM lane quarter (warp & 3); gate item = (chunk cc of the CTA, batch row b)
    const int et = tid - 32;                        // 0..127 for epilogue threads
    const bool is_epi = warp >= 1;
    const int q = warp & 3;
    const int cc = is_epi ? et / NB : 0, bb = is_epi ? et % NB : 0;
    const bool has_item = is_epi && cc < CPC;
    // per-CTA constant descriptors (A stationary; B for each of the two h buffers)
    const uint64_t dah0 = mkd(su32(a_hi), TM * 16, 128), dal0 = mkd(su32(a_lo), TM * 16, 128);
    const uint64_t dbh0 = mkd(su32(b_hi), NB * 16, 128), dbl0 = mkd(su32(b_lo), NB * 16, 128);
    const uint64_t dbh1 = mkd(su32(b_hi + L.b_bytes), NB * 16, 128), dbl1 = mkd(su32(b_lo + L.b_bytes), NB * 16, 128);
    const uint64_t a_step = (uint64_t)((2 * TM * 16) >> 4), b_step = (uint64_t)((2 * NB * 16) >> 4);
    // hidden-side biases of this thread's 8 units (constant over the sequence)
    float bhr[8], bhz[8], bhn[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int j = j0 + cc * 8 + i;
        const bool okj = has_item && j < H;
        bhr[i] = okj ? b_hh[j] : 0.f; bhz[i] = okj ? b_hh[H + j] : 0.f; bhn[i] = okj ? b_hh[2 * H + j] : 0.f;
    }
    uint32_t it = 0;  // running step counter: phase parity of bar_mma
    const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0;

    for (int task = cluster_id >> 1; task < p.n_chunks; task += n_clusters >> 1) {
        const int m0 = task * NB;
        // h_{-1} = 0: packed buffer 0 (hi and lo) and the fp32 master copy
        for (int e = tid; e < KC * NB; e += TNT) {
            reinterpret_cast<uint4*>(b_hi)[e] = make_uint4(0, 0, 0, 0);
            reinterpret_cast<uint4*>(b_lo)[e] = make_uint4(0, 0, 0, 0);
        }
        for (int e = tid; e < HSP * NB; e += TNT) hown[e] = 0.f;
        asm volatile("fence.proxy.async;" ::: "memory");
        cluster.sync();
        int cur = 0;
        for (int s = 0; s < T; ++s, ++it) {
            const int t = dir == 0 ? s : T - 1 - s;
            // ---- tensor core: D[128 x 16] = W_slice * h_{t-1}^T ------------------------------------------------
            if (tid == 0) {
                if (dbg_on) p.dbg[s * 8 + 0] = clock64();
                asm volatile("fence.proxy.async;" ::: "memory");
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // descriptors differ only in their 14-bit start-address field: one 64-bit add per operand per K step
                uint64_t dah = dah0, dal = dal0;
                uint64_t dbh = cur ? dbh1 : dbh0, dbl = cur ? dbl1 : dbl0;
                mma16(tmem_d, dah, dbh, idesc, 0u);
                mma16(tmem_d, dah, dbl, idesc, 1u);
                mma16(tmem_d, dal, dbh, idesc, 1u);
#pragma unroll 4
                for (int ks = 1; ks < KC / 2; ++ks) {
                    dah += a_step; dal += a_step; dbh += b_step; dbl += b_step;
                    mma16(tmem_d, dah, dbh, idesc, 1u);
                    mma16(tmem_d, dah, dbl, idesc, 1u);
                    mma16(tmem_d, dal, dbh, idesc, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(bar_mma)) : "memory");
                if (dbg_on) p.dbg[s * 8 + 1] = clock64();
            }
            if (is_epi) {
                // x-side pre-activations for this thread's 8 units (independent of the recurrence: issued before the wait)
                float gir[8], giz[8], gin[8];
                const int b = m0 + bb;
                const int jbase = j0 + cc * 8;
                const bool live = has_item && b < M;
#pragma unroll
                for (int i = 0; i < 8; ++i) gir[i] = giz[i] = gin[i] = 0.f;
                const bool full8 = jbase + 7 < H;   // whole 8-unit chunk valid -> 128-bit accesses (rows are 16-byte aligned)
                if (live) {
                    const float* g = p.gi + (((size_t)b * T + t) * 2 + dir) * 3 * H;
                    if (full8) {
                        const float4 r0 = *reinterpret_cast<const float4*>(g + jbase), r1 = *reinterpret_cast<const float4*>(g + jbase + 4);
                        const float4 z0 = *reinterpret_cast<const float4*>(g + H + jbase), z1 = *reinterpret_cast<const float4*>(g + H + jbase + 4);
                        const float4 n0 = *reinterpret_cast<const float4*>(g + 2 * H + jbase), n1 = *reinterpret_cast<const float4*>(g + 2 * H + jbase + 4);
                        gir[0] = r0.x; gir[1] = r0.y; gir[2] = r0.z; gir[3] = r0.w; gir[4] = r1.x; gir[5] = r1.y; gir[6] = r1.z; gir[7] = r1.w;
                        giz[0] = z0.x; giz[1] = z0.y; giz[2] = z0.z; giz[3] = z0.w; giz[4] = z1.x; giz[5] = z1.y; giz[6] = z1.z; giz[7] = z1.w;
                        gin[0] = n0.x; gin[1] = n0.y; gin[2] = n0.z; gin[3] = n0.w; gin[4] = n1.x; gin[5] = n1.y; gin[6] = n1.z; gin[7] = n1.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int j = jbase + i;
                            if (j < H) { gir[i] = g[j]; giz[i] = g[H + j]; gin[i] = g[2 * H + j]; }
                        }
                    }
                }
                mbw(bar_mma, it & 1);
                if (dbg_on && tid == 32) p.dbg[s * 8 + 2] = clock64();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                {   // TMEM -> shared: lane (gate row) q*32+lane holds 16 batch columns
                    uint32_t r[16];
                    const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    float* grow = G + (size_t)(q * 32 + lane) * (NB + 1);
#pragma unroll
                    for (int i = 0; i < 16; ++i) grow[i] = __uint_as_float(r[i]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");  // the 4 epilogue warps
                if (dbg_on && tid == 32) p.dbg[s * 8 + 3] = clock64();
                if (has_item) {
                    float hnew[8], sr[8], sz[8], sn[8], shn[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int u = cc * 8 + i, j = jbase + i;
                        hnew[i] = sr[i] = sz[i] = sn[i] = shn[i] = 0.f;
                        if (live && j < H) {
                            const float hr = G[(size_t)u * (NB + 1) + bb] + bhr[i];
                            const float hz = G[(size_t)(HSP + u) * (NB + 1) + bb] + bhz[i];
                            const float hn = G[(size_t)(2 * HSP + u) * (NB + 1) + bb] + bhn[i];
                            // ex2-based sigmoid/tanh (abs error ~1e-7, far inside the 1e-3 parity bar; 3x fewer instructions
                            // than expf/tanhf on the serial critical path of the recurrence)
                            const float r = __fdividef(1.f, 1.f + __expf(-(gir[i] + hr)));
                            const float z = __fdividef(1.f, 1.f + __expf(-(giz[i] + hz)));
                            const float n = 1.f - __fdividef(2.f, 1.f + __expf(2.f * (gin[i] + r * hn)));
                            const float hp = hown[u * NB + bb];
                            hnew[i] = (1.f - z) * n + z * hp;
                            sr[i] = r; sz[i] = z; sn[i] = n; shn[i] = hn;
                        }
                        hown[u * NB + bb] = hnew[i];
                    }
                    if (live) {
                        const size_t row = (size_t)b * T + t;
                        float* yo = p.y + row * 2 * H + dir * H + jbase;
                        float* gs = p.gates != nullptr ? p.gates + (row * 2 + dir) * 4 * H + jbase : nullptr;
                        if (full8) {
                            reinterpret_cast<float4*>(yo)[0] = make_float4(hnew[0], hnew[1], hnew[2], hnew[3]);
                            reinterpret_cast<float4*>(yo)[1] = make_float4(hnew[4], hnew[5], hnew[6], hnew[7]);
                            if (gs != nullptr) {
                                reinterpret_cast<float4*>(gs)[0] = make_float4(sr[0], sr[1], sr[2], sr[3]);
                                reinterpret_cast<float4*>(gs)[1] = make_float4(sr[4], sr[5], sr[6], sr[7]);
                                reinterpret_cast<float4*>(gs + H)[0] = make_float4(sz[0], sz[1], sz[2], sz[3]);
                                reinterpret_cast<float4*>(gs + H)[1] = make_float4(sz[4], sz[5], sz[6], sz[7]);
                                reinterpret_cast<float4*>(gs + 2 * H)[0] = make_float4(sn[0], sn[1], sn[2], sn[3]);
                                reinterpret_cast<float4*>(gs + 2 * H)[1] = make_float4(sn[4], sn[5], sn[6], sn[7]);
                                reinterpret_cast<float4*>(gs + 3 * H)[0] = make_float4(shn[0], shn[1], shn[2], shn[3]);
                                reinterpret_cast<float4*>(gs + 3 * H)[1] = make_float4(shn[4], shn[5], shn[6], shn[7]);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                if (jbase + i < H) {
                                    yo[i] = hnew[i];
                                    if (gs != nullptr) { gs[i] = sr[i]; gs[H + i] = sz[i]; gs[2 * H + i] = sn[i]; gs[3 * H + i] = shn[i]; }
                                }
                            }
                        }
                    }
                    if (dbg_on && tid == 32) p.dbg[s * 8 + 4] = clock64();
                    uint4 h4, l4;
                    split2g(hnew[0], hnew[1], h4.x, l4.x); split2g(hnew[2], hnew[3], h4.y, l4.y);
                    split2g(hnew[4], hnew[5], h4.z, l4.z); split2g(hnew[6], hnew[7], h4.w, l4.w);
                    const size_t off = (size_t)(cur ^ 1) * L.b_bytes + ((size_t)(rank * CPC + cc) * NB + bb) * 16;
#pragma unroll
                    for (int dst = 0; dst < CL; ++dst) {
                        unsigned char* rh = cluster.map_shared_rank(b_hi, dst);
                        unsigned char* rl = cluster.map_shared_rank(b_lo, dst);
                        *reinterpret_cast<uint4*>(rh + off) = h4;
                        *reinterpret_cast<uint4*>(rl + off) = l4;
                    }
                    asm volatile("fence.proxy.async;" ::: "memory");
                    if (dbg_on && tid == 32) p.dbg[s * 8 + 5] = clock64();
                }
            }
            cluster.sync();  // every CTA holds the complete packed h_t in buffer cur^1
            if (dbg_on && tid == 32) p.dbg[s * 8 + 6] = clock64();
            cur ^= 1;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster.sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(32));
}

}  // namespace

// 1 through *ok if the tensor-core recurrence can serve hidden size H (gate rows 3*HSP <= 128, shared memory fit).
HA2G_API int ha2g_gru_tc_supported(int H, int* ok) {
    const int HSP = ((H + CL - 1) / CL + 7) / 8 * 8;
    const TcLayout L(HSP);
    *ok = (3 * HSP <= TM && (HSP / 8) * NB <= 128 && L.total <= 227 * 1024) ? 1 : 0;
    return 0;
}

extern "C" int ha2g_gru_seq_fwd_tc_dbg(const float*, const float*, const float*, const float*, const float*, float*, float*,
                                       int, int, int, long long*, cudaStream_t);
// All T steps of one bidirectional layer, forward, on tcgen05 (see the file header).  gi must hold x W_ih^T + b_ih.
HA2G_API int ha2g_gru_seq_fwd_tc(const float* gi, const float* w_hh_f, const float* w_hh_r, const float* b_hh_f,
                                 const float* b_hh_r, float* y, float* gates, int M, int T, int H, cudaStream_t stream) {
    return ha2g_gru_seq_fwd_tc_dbg(gi, w_hh_f, w_hh_r, b_hh_f, b_hh_r, y, gates, M, T, H, nullptr, stream);
}

// Same, with an optional device buffer dbg [T][8] of clock64() samples (cluster 0, rank 0) for phase timing.
HA2G_API int ha2g_gru_seq_fwd_tc_dbg(const float* gi, const float* w_hh_f, const float* w_hh_r, const float* b_hh_f,
                                     const float* b_hh_r, float* y, float* gates, int M, int T, int H, long long* dbg,
                                     cudaStream_t stream) {
    TcParams p{};
    p.dbg = dbg;
    p.gi = gi; p.w_hh[0] = w_hh_f; p.w_hh[1] = w_hh_r; p.b_hh[0] = b_hh_f; p.b_hh[1] = b_hh_r;
    p.y = y; p.gates = gates; p.M = M; p.T = T; p.H = H;
    p.HSP = ((H + CL - 1) / CL + 7) / 8 * 8;
    p.n_chunks = (M + NB - 1) / NB;
    const TcLayout L(p.HSP);
    cudaError_t e = cudaFuncSetAttribute(gru_seq_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    if (e != cudaSuccess) return (int)e;
    int clusters = 2 * p.n_chunks;
    if (clusters > 16) clusters = 16;
    gru_seq_fwd_tc_kernel<<<clusters * CL, TNT, L.total, stream>>>(p);
    HA2G_RETURN_LAST();
}
