// Persistent cluster GRU recurrence on tcgen05, BACKWARD (BPTT through one bidirectional layer).
//
// Same decomposition as the forward kernel (gru_cluster_tc2.cu): an 8-CTA cluster per (direction, NB-row batch chunk);
// CTA `rank` owns HSP = 40 hidden units, i.e. the 3*HSP gate rows (r | z | n) of W_hh that belong to them.  Walking the
// sequence backwards, every step does
//   A  (epilogue warps)  dh_t of the owned units = dy_t + dh_{t+1}*z_{t+1} (carried in registers) + the 8 partial products
//      that landed in shared memory; gate gradients dr, dz, dn of the owned units -> dgi / dgh (global, for the
//      weight-gradient GEMMs) and, as bf16 hi/lo, the B operand [K = 3*HSP (pad 128) x 16] of the step's MMA;
//   C  (one elected lane) partial dh_{t-1}[all units] = W_hh[own gate rows, :]^T * dgh_own: 3 M-tiles x 8 K-steps x 3
//      bf16-split terms of tcgen05.mma with the transposed W_hh slice resident in TENSOR MEMORY (384 columns);
//   D  (epilogue warps)  the [384 x NB] fp32 partial goes TMEM -> shared staging (transposed: [owning CTA][batch row][unit],
//      so that the receiver's reduction reads 128-bit words) -> 8 bulk copies (one slice per owning CTA) that complete_tx on the destination's mbarrier: a reduce-scatter over distributed shared memory with no
//      cluster barrier in the loop.  Receive and staging buffers are double-buffered; a peer can only be two half-steps
//      ahead of me after it has consumed what I sent (see the forward kernel's header for the argument).
// Everything that depends only on saved tensors (gates, y, dy) is loaded before the wait for the partials.
// Tried and measured no faster at B = 128 (tools/time_gru_tc.py, 133 us per layer as shipped): two interleaved half-tasks
// of 10 rows per cluster as in the forward kernel (138 us: the chain of a half is latency- not volume-bound here, so it
// does not get shorter with fewer rows) and a register prefetch of the saved tensors one round ahead (137 us).
// Replaces the fp32 FMA kernel gru_seq_bwd_cluster_kernel (652 us per layer at B = 128: 19 us per step) behind
// loss.backward() through nn.GRU, scripts/model/hierarchy_net.py:144 (H = 300) and :232 (H = 64).
#include "common.cuh"
#include <cooperative_groups.h>
#include <cuda_bf16.h>

namespace cg = cooperative_groups;

namespace {

constexpr int CL = 8;          // CTAs per cluster
// NB = batch rows per cluster task: 16, or 20 / 32 when 16-row chunks would need more clusters than the device keeps
// resident (15 on a B200: the 16 clusters of a 128-row batch ran as two waves, i.e. at twice the time; 7 chunks of 20 rows
// per direction fit one wave).  NN = UMMA N = NB rounded up to 16; rows NB..NN-1 of the B operand stay zero.
constexpr int TM = 128;        // UMMA M (hidden units per M-tile)
constexpr int TMEM_COLS = 512;
// warp 0: MMA issue + TMEM alloc; 4 (NB = 16) or 8 (NB = 32) epilogue warps: everything else
__host__ __device__ constexpr int epi_warps(int NB) { return NB <= 24 ? 4 : 8; }
__host__ __device__ constexpr int mma_n(int NB) { return (NB + 15) / 16 * 16; }
__host__ __device__ constexpr int block_threads(int NB) { return 32 + 32 * epi_warps(NB); }
// D tiles at columns [0, NN*n_mt), W^T slices (384 columns) from column a_col
__host__ __device__ constexpr int a_col(int NB) { return mma_n(NB) == 16 ? 64 : 128; }
constexpr size_t MIN_SMEM = 120 * 1024;   // one CTA per SM (512-column TMEM allocation)

struct BwdParams {
    const float* dy; int dy_ld, dy_dir_stride;
    const float* y;        // [M,T,2H]
    const float* gates;    // [M,T,2,4H]  r | z | n | hn
    const float* w_hh[2];  // [3H,H]
    float* dgi;            // [M,T,2,3H]
    float* dgh;            // [M,T,2,3H]
    int M, T, H, HSP, n_chunks;
    long long* dbg;
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
__device__ __forceinline__ void mb_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(mbar_cluster) : "memory");
}
__device__ __forceinline__ uint64_t mkd(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void mma16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void split2g(float a, float b, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<uint32_t*>(&h);
    const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xffff0000u);
    __nv_bfloat162 l = __floats2bfloat162_rn(a - ah, b - bh);
    lo = *reinterpret_cast<uint32_t*>(&l);
}
// 8 consecutive floats of which the first 4*n4 exist (n4 = 1: the chunk straddles H, e.g. units 296..303 of 300)
__device__ __forceinline__ void ld8(const float* p, float* v, int n4) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    if (n4 > 1) {
        const float4 b = *reinterpret_cast<const float4*>(p + 4);
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
}
__device__ __forceinline__ void st8(float* p, const float* v) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// shared memory map (bytes): B operand | barriers + tmem slot | { recv[2] | send[2] | out staging }  (the braces alias
// the one-time fp32 staging of the CTA's W_hh rows)
struct BwdLayout {
    int ksteps, n_mt, orow;
    size_t bop_bytes, recv_bytes, send_bytes, slice_bytes, off_bar, off_recv, off_send, off_out, off_w, total;
    __host__ __device__ BwdLayout(int HSP, int H, int NB) {
        ksteps = (3 * HSP + 15) / 16;
        n_mt = (CL * HSP + TM - 1) / TM;
        orow = HSP + 4;
        bop_bytes = (size_t)ksteps * 2 * 2 * mma_n(NB) * 16;
        slice_bytes = (size_t)NB * (HSP + 4) * 4;       // [NB batch rows][HSP units + 4]: the receiver reads 128-bit words
        recv_bytes = (size_t)CL * slice_bytes;          // one buffer
        send_bytes = (size_t)CL * slice_bytes;          // one buffer (rows k = dst*HSP + u)
        off_bar = bop_bytes;
        off_recv = off_bar + 64;
        off_send = off_recv + 2 * recv_bytes;
        off_out = off_send + 2 * send_bytes;
        off_w = off_recv;
        size_t end_loop = off_out + (size_t)4 * NB * orow * 4;
        size_t end_w = off_w + (size_t)ksteps * 16 * H * 4;   // K rows incl. zero padding
        total = end_loop > end_w ? end_loop : end_w;
        if (total < MIN_SMEM) total = MIN_SMEM;
    }
};

template <int NB>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(block_threads(NB), 1) gru_seq_bwd_tc2_kernel(BwdParams p) {
    constexpr int TNT = block_threads(NB), NET = 32 * epi_warps(NB), A_COL = a_col(NB), NN = mma_n(NB);
    constexpr int CW = NB / (epi_warps(NB) / 4);   // accumulator columns (batch rows) per epilogue warp: 16 or 20
    constexpr bool RPF = epi_warps(NB) == 4;       // 5 warps per CTA: up to 255 registers per thread
    extern __shared__ __align__(128) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
    const int dir = cluster_id & 1;
    const int H = p.H, HSP = p.HSP, T = p.T, M = p.M;
    const BwdLayout L(HSP, H, NB);
    const int KS = L.ksteps, NMT = L.n_mt, CPC = HSP / 8;
    unsigned char* bop = smem;                                   // [2*KS chunks][hi | lo][NN][16 B]
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* bar_recv = bar_mma + 1;                            // [2]
    uint64_t* bar_w = bar_mma + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 4);
    float* recv = reinterpret_cast<float*>(smem + L.off_recv);   // [2][CL src][NB][orow]
    float* send = reinterpret_cast<float*>(smem + L.off_send);   // [2][CL dst][NB][orow]
    float* outst = reinterpret_cast<float*>(smem + L.off_out);   // [4][NB][orow]: dr | dz | dn | dn*r
    const float* wrows = reinterpret_cast<const float*>(smem + L.off_w);   // [3*HSP][H], prologue only
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = ha2g_warp_id();
    const int j0 = rank * HSP;
    // kernel parameters into registers up front (dynamic indexing of the by-value struct would spill it to local memory
    // and make every global access a generic LD.E / ST.E behind an LDL)
    const float* __restrict__ W = dir ? p.w_hh[1] : p.w_hh[0];
    const float* __restrict__ g_dy = p.dy;
    const float* __restrict__ g_y = p.y;
    const float* __restrict__ g_gates = p.gates;
    float* __restrict__ g_dgi = p.dgi;
    float* __restrict__ g_dgh = p.dgh;
    const int dy_ld = p.dy_ld, dy_ds = p.dy_dir_stride;
    const uint32_t tx_bytes = (uint32_t)(CL * L.slice_bytes);
    const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0;
    if (dbg_on && tid == 0) p.dbg[T * 8 + 0] = clock64();

    if (tid == 0) {
        mbi(bar_mma, 1);
        mbi(bar_recv + 0, 1);
        mbi(bar_recv + 1, 1);
        mbi(bar_w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(su32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int e = tid; e < (int)(L.bop_bytes / 16); e += TNT) reinterpret_cast<uint4*>(bop)[e] = make_uint4(0, 0, 0, 0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t tmem_d = tmem_base, tmem_a = tmem_base + A_COL;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

    const int et = tid - 32;
    const bool is_epi = warp >= 1;
    const int q = warp & 3;
    // item -> thread: the chunk index runs fastest, so that a warp's 32 items cover ~6 batch rows x the CTA's 160
    // contiguous bytes per row of every global tensor (lanes across batch rows made every 128-bit load / store touch 32
    // different lines and kept the load/store unit busy for ~1 500 cycles per step)
    const int cc = is_epi ? et % CPC : 0, bb = is_epi ? et / CPC : 0;
    const bool has_item = is_epi && bb < NB;

    // ---- one-time: W_hh rows of the owned gates, global -> shared (bulk copies) -> TMEM, transposed -------------------
    {
        int nvalid = 0;
        for (int g = 0; g < 3; ++g) { const int left = H - j0; nvalid += left <= 0 ? 0 : (left < HSP ? left : HSP); }
        if (tid == 0 && nvalid > 0) mb_expect_tx(bar_w, (uint32_t)((size_t)nvalid * H * 4));
        __syncthreads();
        {
            // one row per thread: every thread issues its own copies, all rows in flight at once
            for (int r = tid; r < 3 * HSP; r += TNT) {
                const int g = r / HSP, u = r % HSP, j = j0 + u;
                if (j < H)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(su32(wrows + (size_t)r * H)), "l"(W + ((size_t)g * H + j) * H), "r"((uint32_t)(H * 4)),
                                   "r"(su32(bar_w)) : "memory");
            }
            __syncwarp();
        }
        // rows that no bulk copy fills (K padding, units beyond H) are zeroed, so the transposing pass needs no predicates
        float* wz = const_cast<float*>(wrows);
        for (int r = warp; r < KS * 16; r += TNT / 32) {
            const int g = r / HSP, u = r - g * HSP;
            if (r >= 3 * HSP || j0 + u >= H)
                for (int c = lane; c < H; c += 32) wz[(size_t)r * H + c] = 0.f;
        }
        if (nvalid > 0) mbw(bar_w, 0);
        __syncthreads();
    }
    if (dbg_on && tid == 0) p.dbg[T * 8 + 1] = clock64();
    // A operand of M-tile mt, term (hi | lo): lane = hidden unit k = mt*128 + lane, 32-bit column ks*8 + i = the bf16 pair of
    // gate rows (kk = 16 ks + 2i, +1), kk = gate*HSP + u  <->  W_hh[gate*H + j0 + u][k]
    if (is_epi && warp <= 4) {   // one warp per TMEM lane quarter
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        for (int mt = 0; mt < NMT; ++mt) {
            const int k = mt * TM + q * 32 + lane;
            const float km = k < H ? 1.f : 0.f;
            const float* col = wrows + (k < H ? k : H - 1);
#pragma unroll 2
            for (int ks = 0; ks < KS; ++ks) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int kk0 = ks * 16 + 2 * i;
                    split2g(km * col[(size_t)kk0 * H], km * col[(size_t)(kk0 + 1) * H], hi[i], lo[i]);
                }
                const uint32_t chi = tmem_a + lane_addr + (uint32_t)((mt * 2 + 0) * KS * 8 + ks * 8);
                const uint32_t clo = tmem_a + lane_addr + (uint32_t)((mt * 2 + 1) * KS * 8 + ks * 8);
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                             ::"r"(chi), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]) : "memory");
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                             ::"r"(clo), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]) : "memory");
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // wrows is dead from here on: its bytes become recv / send / out staging
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (dbg_on && tid == 0) p.dbg[T * 8 + 2] = clock64();

    const uint32_t lbo = 2 * NN * 16;
    const uint64_t dbh0 = mkd(su32(bop), lbo, 128), dbl0 = mkd(su32(bop + NN * 16), lbo, 128);
    const uint64_t b_step = (uint64_t)((2 * lbo) >> 4);
    // copy-out roles for dgi / dgh (fixed per thread)
    int co_n = 0, co_rb[2] = {0, 0}, co_f4[2] = {0, 0};
    if (is_epi) {
        const int q4 = HSP / 4;
        for (int k = 0; k < 2; ++k) {
            const int idx = et + k * NET;
            if (idx < NB * q4) { co_rb[k] = idx / q4; co_f4[k] = idx % q4; co_n = k + 1; }
        }
    }
    // D phase (fixed per thread): first accumulator column (batch row) of this warp, and where the partial of unit
    // k = mt*128 + q*32 + lane goes inside a send buffer (slice of the owning CTA k / HSP, transposed: [batch row][unit])
    const int OR = L.orow;
    const int cg16 = is_epi ? ((warp - 1) >> 2) * CW : 0;
    int sd_off[3];
#pragma unroll
    for (int mt = 0; mt < 3; ++mt) {
        const int k = mt * TM + q * 32 + lane;
        sd_off[mt] = (is_epi && mt < NMT && k < CL * HSP) ? ((k / HSP) * NB + cg16) * OR + (k % HSP) : -1;
    }
    uint32_t it = 0;                       // MMA rounds so far: phase parity of bar_mma
    uint32_t recv_ph0 = 0, recv_ph1 = 0;   // phase parities of bar_recv (tracked by every epilogue thread)

    for (int task = cluster_id >> 1; task < p.n_chunks; task += n_clusters >> 1) {
        const int m0 = task * NB;
        if (tid == 0) {   // partials of round 0 land in buffer 1 (read by round 1), of round 1 in buffer 0 (read by round 2)
            if (T >= 2) mb_expect_tx(bar_recv + 1, tx_bytes);
            if (T >= 3) mb_expect_tx(bar_recv + 0, tx_bytes);
        }
        float dhz[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dhz[i] = 0.f;
        cluster.sync();   // every CTA is out of its prologue / previous task: its buffers may be written
        if (dbg_on && tid == 0) p.dbg[T * 8 + 3] = clock64();
        for (int rd = 0; rd < T; ++rd) {
            const int s = T - 1 - rd;
            const int t = dir == 0 ? s : T - 1 - s;
            const int tp = dir == 0 ? t - 1 : t + 1;
            const int cur = rd & 1;
            if (is_epi) {
                // ---- saved tensors of this step (independent of the recurrence: in flight while the partials travel) ----
                float r[8], z[8], n[8], hn[8], hp[8], dyv[8];
                const int b = m0 + bb;
                const int jbase = j0 + cc * 8;
                const int n4 = (H - jbase) >= 8 ? 2 : ((H - jbase) >= 4 ? 1 : 0);   // valid float4 halves of the chunk (H % 4 == 0)
                const bool live = has_item && b < M && n4 > 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = z[i] = n[i] = hn[i] = hp[i] = dyv[i] = 0.f;
                const size_t row = (size_t)b * T + t;
                if (live) {
                    const float* gs = g_gates + (row * 2 + dir) * 4 * H + jbase;
                    ld8(gs, r, n4); ld8(gs + H, z, n4); ld8(gs + 2 * H, n, n4); ld8(gs + 3 * H, hn, n4);
                    if (s > 0) ld8(g_y + ((size_t)b * T + tp) * 2 * H + dir * H + jbase, hp, n4);
                    ld8(g_dy + row * dy_ld + dir * dy_ds + jbase, dyv, n4);
                }
                // next round's saved tensors into L2 now (no registers held): their loads, issued at the top of the next
                // round only ~1 400 cycles before first use, then pay an L2 hit instead of a DRAM round trip
                if (live && rd + 1 < T) {
                    const int s2 = s - 1, t2 = dir == 0 ? s2 : T - 1 - s2, tp2 = dir == 0 ? t2 - 1 : t2 + 1;
                    const size_t row2 = (size_t)b * T + t2;
                    const float* gs2 = g_gates + (row2 * 2 + dir) * 4 * H + jbase;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(gs2));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(gs2 + H));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(gs2 + 2 * H));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(gs2 + 3 * H));
                    if (s2 > 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(g_y + ((size_t)b * T + tp2) * 2 * H + dir * H + jbase));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(g_dy + row2 * dy_ld + dir * dy_ds + jbase));
                }
                if (dbg_on && tid == 32) p.dbg[rd * 8 + 0] = clock64();
                // ---- A: wait for the 8 partial products of dh_t, reduce, gate gradients ---------------------------------
                float acc[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = dhz[i];
                if (rd > 0) {
                    mbw(bar_recv + cur, cur ? recv_ph1 : recv_ph0);
                    if (cur) recv_ph1 ^= 1; else recv_ph0 ^= 1;
                    if (tid == 32 && rd + 2 <= T - 1) mb_expect_tx(bar_recv + cur, tx_bytes);
                    if (dbg_on && tid == 32) p.dbg[rd * 8 + 1] = clock64();
                    if (has_item) {
                        // batch-row-major slices: the thread's 8 units of one source are two 128-bit words (the unit-major
                        // layout of the first version cost 64 scalar loads per thread, 5-way bank-conflicted)
                        const float* rb = recv + (size_t)cur * (L.recv_bytes / 4) + (size_t)bb * L.orow + cc * 8;
#pragma unroll
                        for (int src = 0; src < CL; ++src) {
                            const float4 a0 = *reinterpret_cast<const float4*>(rb + (size_t)src * NB * L.orow);
                            const float4 a1 = *reinterpret_cast<const float4*>(rb + (size_t)src * NB * L.orow + 4);
                            acc[0] += a0.x; acc[1] += a0.y; acc[2] += a0.z; acc[3] += a0.w;
                            acc[4] += a1.x; acc[5] += a1.y; acc[6] += a1.z; acc[7] += a1.w;
                        }
                    }
                }
                float o_r[8], o_z[8], o_n[8], o_nr[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float dh = (live && i < 4 * n4) ? dyv[i] + acc[i] : 0.f;
                    const float dn = dh * (1.f - z[i]);
                    const float dz = dh * (hp[i] - n[i]);
                    const float dn_pre = dn * (1.f - n[i] * n[i]);
                    const float dz_pre = dz * z[i] * (1.f - z[i]);
                    const float dr_pre = dn_pre * hn[i] * r[i] * (1.f - r[i]);
                    o_r[i] = dr_pre; o_z[i] = dz_pre; o_n[i] = dn_pre; o_nr[i] = dn_pre * r[i];
                    dhz[i] = dh * z[i];
                }
                if (has_item) {
                    if (rd < T - 1) {   // B operand of this round's MMA: chunk (gate*CPC + cc), hi | lo, row bb
                        uint4 h4, l4;
                        split2g(o_r[0], o_r[1], h4.x, l4.x); split2g(o_r[2], o_r[3], h4.y, l4.y);
                        split2g(o_r[4], o_r[5], h4.z, l4.z); split2g(o_r[6], o_r[7], h4.w, l4.w);
                        *reinterpret_cast<uint4*>(bop + ((size_t)((0 * CPC + cc) * 2 + 0) * NN + bb) * 16) = h4;
                        *reinterpret_cast<uint4*>(bop + ((size_t)((0 * CPC + cc) * 2 + 1) * NN + bb) * 16) = l4;
                        split2g(o_z[0], o_z[1], h4.x, l4.x); split2g(o_z[2], o_z[3], h4.y, l4.y);
                        split2g(o_z[4], o_z[5], h4.z, l4.z); split2g(o_z[6], o_z[7], h4.w, l4.w);
                        *reinterpret_cast<uint4*>(bop + ((size_t)((1 * CPC + cc) * 2 + 0) * NN + bb) * 16) = h4;
                        *reinterpret_cast<uint4*>(bop + ((size_t)((1 * CPC + cc) * 2 + 1) * NN + bb) * 16) = l4;
                        split2g(o_nr[0], o_nr[1], h4.x, l4.x); split2g(o_nr[2], o_nr[3], h4.y, l4.y);
                        split2g(o_nr[4], o_nr[5], h4.z, l4.z); split2g(o_nr[6], o_nr[7], h4.w, l4.w);
                        *reinterpret_cast<uint4*>(bop + ((size_t)((2 * CPC + cc) * 2 + 0) * NN + bb) * 16) = h4;
                        *reinterpret_cast<uint4*>(bop + ((size_t)((2 * CPC + cc) * 2 + 1) * NN + bb) * 16) = l4;
                    }
                    float* o = outst + (size_t)bb * L.orow + cc * 8;
                    const size_t as = (size_t)NB * L.orow;
                    st8(o, o_r); st8(o + as, o_z); st8(o + 2 * as, o_n); st8(o + 3 * as, o_nr);
                }
                if (dbg_on && tid == 32) p.dbg[rd * 8 + 2] = clock64();
            }
            if (rd < T - 1) {
                // ---- C: partial dh_{t-1} for all units on the tensor cores ------------------------------------------
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                asm volatile("bar.sync 2, %0;" ::"n"(TNT) : "memory");   // B operand written; the previous round's D has been read
                if (warp == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (ha2g_elect_one()) {
                        for (int mt = 0; mt < NMT; ++mt) {
                            uint32_t ah = tmem_a + (uint32_t)((mt * 2 + 0) * KS * 8), al = tmem_a + (uint32_t)((mt * 2 + 1) * KS * 8);
                            uint64_t dbh = dbh0, dbl = dbl0;
                            const uint32_t dcol = tmem_d + (uint32_t)(mt * NN);
                            mma16_ts(dcol, ah, dbh, idesc, 0u);
                            mma16_ts(dcol, ah, dbl, idesc, 1u);
                            mma16_ts(dcol, al, dbh, idesc, 1u);
#pragma unroll 4
                            for (int ks = 1; ks < KS; ++ks) {
                                ah += 8; al += 8; dbh += b_step; dbl += b_step;
                                mma16_ts(dcol, ah, dbh, idesc, 1u);
                                mma16_ts(dcol, ah, dbl, idesc, 1u);
                                mma16_ts(dcol, al, dbh, idesc, 1u);
                            }
                        }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(bar_mma)) : "memory");
                    }
                    __syncwarp();
                }
            }
            if (is_epi) {
                // ---- dgi / dgh of this step to global (coalesced through the staging lines) while the MMAs run -----------
                asm volatile("bar.sync 1, %0;" ::"n"(NET) : "memory");
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (k < co_n) {
                        const int rb = co_rb[k], f4 = co_f4[k];
                        const int bg = m0 + rb, jg = j0 + f4 * 4;
                        if (bg < M && jg < H) {
                            const size_t row = (size_t)bg * T + t;
                            const float* src = outst + (size_t)rb * L.orow + f4 * 4;
                            const size_t as = (size_t)NB * L.orow;
                            const float4 vr = *reinterpret_cast<const float4*>(src), vz = *reinterpret_cast<const float4*>(src + as);
                            const float4 vn = *reinterpret_cast<const float4*>(src + 2 * as), vnr = *reinterpret_cast<const float4*>(src + 3 * as);
                            float* gi_o = g_dgi + (row * 2 + dir) * 3 * H + jg;
                            float* gh_o = g_dgh + (row * 2 + dir) * 3 * H + jg;
                            *reinterpret_cast<float4*>(gi_o) = vr; *reinterpret_cast<float4*>(gi_o + H) = vz;
                            *reinterpret_cast<float4*>(gi_o + 2 * H) = vn;
                            *reinterpret_cast<float4*>(gh_o) = vr; *reinterpret_cast<float4*>(gh_o + H) = vz;
                            *reinterpret_cast<float4*>(gh_o + 2 * H) = vnr;
                        }
                    }
                }
                if (rd < T - 1) {
                    // ---- D: partial products TMEM -> staging -> the owning CTAs ------------------------------------------
                    mbw(bar_mma, it & 1);
                    if (dbg_on && tid == 32) p.dbg[rd * 8 + 3] = clock64();
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    float* sd = send + (size_t)cur * (L.send_bytes / 4);
                    // 4 epilogue warps (registers to spare): all tiles' accumulator columns into registers first (one wait),
                    // then the transposed stores; 8 warps: tile by tile
                    constexpr int VT = RPF ? 3 : 1, VN = CW > 16 ? 20 : 16;
                    uint32_t v[VT][VN];
                    auto ld_tile = [&](int mt, uint32_t* vv) {
                        if (mt < NMT && mt * TM + q * 32 < CL * HSP) {   // warp-uniform: this lane quarter holds real units
                            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * NN + cg16);
                            asm volatile(
                                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                : "=r"(vv[0]), "=r"(vv[1]), "=r"(vv[2]), "=r"(vv[3]), "=r"(vv[4]), "=r"(vv[5]), "=r"(vv[6]), "=r"(vv[7]),
                                  "=r"(vv[8]), "=r"(vv[9]), "=r"(vv[10]), "=r"(vv[11]), "=r"(vv[12]), "=r"(vv[13]), "=r"(vv[14]), "=r"(vv[15])
                                : "r"(taddr));
                            if (CW > 16)
                                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                                             : "=r"(vv[VN - 4]), "=r"(vv[VN - 3]), "=r"(vv[VN - 2]), "=r"(vv[VN - 1]) : "r"(taddr + 16u));
                        }
                    };
                    auto st_tile = [&](int mt, const uint32_t* vv) {
                        if (sd_off[mt] >= 0) {   // unit k of tile mt belongs to CTA k / HSP: its partial goes into that CTA's slice, transposed
                            float* dcol = sd + sd_off[mt];
#pragma unroll
                            for (int i = 0; i < VN; ++i)
                                if (i < CW && cg16 + i < NB) dcol[i * OR] = __uint_as_float(vv[i]);
                        }
                    };
                    if (RPF) {
#pragma unroll
                        for (int mt = 0; mt < 3; ++mt) ld_tile(mt, v[mt % VT]);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int mt = 0; mt < 3; ++mt) st_tile(mt, v[mt % VT]);
                    } else {
#pragma unroll
                        for (int mt = 0; mt < 3; ++mt) {
                            ld_tile(mt, v[0]);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                            st_tile(mt, v[0]);
                        }
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("bar.sync 1, %0;" ::"n"(NET) : "memory");
                    if (warp == 1) {
                        if (ha2g_elect_one()) {
                            const uint32_t dst = su32(recv) + (uint32_t)((size_t)(cur ^ 1) * L.recv_bytes + (size_t)rank * L.slice_bytes);
                            const uint32_t bar = su32(bar_recv + (cur ^ 1));
#pragma unroll
                            for (uint32_t d = 0; d < CL; ++d)
                                bulk_s2c(mapa(dst, d), su32(sd) + d * (uint32_t)L.slice_bytes, (uint32_t)L.slice_bytes, mapa(bar, d));
                        }
                        __syncwarp();
                    }
                    if (dbg_on && tid == 32) p.dbg[rd * 8 + 4] = clock64();
                }
            }
            if (rd < T - 1) ++it;
        }
        if (dbg_on && tid == 0) p.dbg[T * 8 + 4] = clock64();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        cluster.sync();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

}  // namespace

template <int NB>
static bool bwd_fits(int H) {
    const int HSP = ((H + CL - 1) / CL + 7) / 8 * 8;
    const BwdLayout L(HSP, H, NB);
    return H % 4 == 0 && (HSP / 8) * NB <= 32 * epi_warps(NB) && NB * (HSP / 4) <= 2 * 32 * epi_warps(NB) &&
           a_col(NB) + L.n_mt * 2 * L.ksteps * 8 <= TMEM_COLS && L.n_mt * mma_n(NB) <= a_col(NB) && L.total <= 227 * 1024;
}

// 1 through *ok if the tensor-core backward recurrence can serve hidden size H.
HA2G_API int ha2g_gru_tc2_bwd_supported(int H, int* ok) {
    *ok = (bwd_fits<16>(H) && bwd_fits<20>(H) && bwd_fits<32>(H)) ? 1 : 0;
    return 0;
}

extern "C" int ha2g_gru_max_clusters(int* n);

template <int NB>
static int launch_bwd(BwdParams& p, cudaStream_t stream) {
    p.n_chunks = (p.M + NB - 1) / NB;
    const BwdLayout L(p.HSP, p.H, NB);
    cudaError_t e = cudaFuncSetAttribute(gru_seq_bwd_tc2_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    if (e != cudaSuccess) return (int)e;
    int cap = 16;
    ha2g_gru_max_clusters(&cap);   // resident 8-CTA clusters (one CTA per SM, like the forward kernel)
    cap &= ~1;
    if (cap < 2) cap = 2;
    int clusters = 2 * p.n_chunks;
    if (clusters > cap) clusters = cap;
    gru_seq_bwd_tc2_kernel<NB><<<clusters * CL, block_threads(NB), L.total, stream>>>(p);
    HA2G_RETURN_LAST();
}

extern "C" int ha2g_gru_seq_bwd_tc2_dbg(const float*, int, int, const float*, const float*, const float*, const float*, float*,
                                        float*, int, int, int, int, long long*, cudaStream_t);
// All T steps of one bidirectional layer, backward, on tcgen05: writes dgi, dgh [M,T,2,3H] (gru.cu documents them) from
// dy (element (m,t,dir,j) at dy[(m*T+t)*dy_ld + dir*dy_dir_stride + j]), y [M,T,2H] and the saved gates [M,T,2,4H].
HA2G_API int ha2g_gru_seq_bwd_tc2(const float* dy, int dy_ld, int dy_dir_stride, const float* y, const float* gates,
                                  const float* w_hh_f, const float* w_hh_r, float* dgi, float* dgh, int M, int T, int H,
                                  cudaStream_t stream) {
    return ha2g_gru_seq_bwd_tc2_dbg(dy, dy_ld, dy_dir_stride, y, gates, w_hh_f, w_hh_r, dgi, dgh, M, T, H, 0, nullptr, stream);
}

// Same, with an explicit rows-per-cluster choice (nb = 16 / 20 / 32; 0 = automatic: the smallest whose chunks are all resident)
// and an optional device buffer dbg [T+1][8] of clock64() samples (cluster 0, rank 0): per round 0 = saved tensors
// requested, 1 = partials arrived, 2 = gate gradients + B operand written, 3 = MMAs done, 4 = partials sent;
// row T: 0 = kernel entry, 1 = W rows in shared memory, 2 = W^T in tensor memory, 3 = loop starts, 4 = loop done.
HA2G_API int ha2g_gru_seq_bwd_tc2_dbg(const float* dy, int dy_ld, int dy_dir_stride, const float* y, const float* gates,
                                      const float* w_hh_f, const float* w_hh_r, float* dgi, float* dgh, int M, int T, int H,
                                      int nb, long long* dbg, cudaStream_t stream) {
    if (M <= 0 || T <= 0) return 0;
    BwdParams p{};
    p.dbg = dbg;
    p.dy = dy; p.dy_ld = dy_ld; p.dy_dir_stride = dy_dir_stride; p.y = y; p.gates = gates;
    p.w_hh[0] = w_hh_f; p.w_hh[1] = w_hh_r; p.dgi = dgi; p.dgh = dgh; p.M = M; p.T = T; p.H = H;
    p.HSP = ((H + CL - 1) / CL + 7) / 8 * 8;
    if (nb == 0) {
        int cap = 16;
        ha2g_gru_max_clusters(&cap);
        nb = (M + 15) / 16 <= cap / 2 ? 16 : ((M + 19) / 20 <= cap / 2 ? 20 : 32);
    }
    if (nb == 16) return launch_bwd<16>(p, stream);
    if (nb == 20) return launch_bwd<20>(p, stream);
    if (nb == 32) return launch_bwd<32>(p, stream);
    return (int)cudaErrorInvalidValue;
}
