// Persistent cluster GRU recurrence, second generation (forward) -- the fused GRU kernel of the step.
//
// Decomposition: an 8-CTA cluster per (direction, NB-row batch chunk), NB = 16 / 32 / 48 chosen from the batch so that the
// 16 resident clusters cover it in one wave (M <= 128 / <= 256 / larger: the 3B-row cascade of the training step runs at
// NB = 48); CTA `rank` owns HSP = 40 hidden units = 3*HSP gate rows (r | z | n, zero-padded to M = 128) of W_hh for all T
// steps, and every step computes
//     gates^T[128 x NB] = W_slice[128 x K] * h_{t-1}^T[K x NB]          (K = 8*HSP = 320, bf16x3 split, fp32 accumulate)
// on tcgen05: the same 60 tcgen05.mma per step carry NB columns, the fixed per-step latency (mbarrier wake-ups, TMEM
// round trip, DSMEM flight) is amortised over NB rows, and with NB > 16 eight epilogue warps (two per TMEM lane quarter,
// each taking half of the columns) keep the gate math at one 8-unit item per thread.
// History, each item taken from the phase timing of the first kernel (tools/time_gru_tc.py:
// 10 600 cycles per step = MMA issue 4 170 + gate math/stores 3 230 + DSMEM push 1 950 + cluster barrier 910):
//   * W_slice lives in TENSOR MEMORY (tcgen05.st once per layer, 320 columns: hi | lo) and is the MMA's A operand
//     straight from TMEM.  With A in shared memory every one of the 60 MMAs of a step re-read 4 KB of weights through
//     the 128 B/clk shared-memory port (240 KB per step = the 4 170 cycles); now a step reads only h (30 KB).
//   * h_t travels by bulk asynchronous copies (cp.async.bulk shared::cta -> shared::cluster, 2.5 KB per destination,
//     8 per CTA per step, issued by 8 lanes) that complete_tx on the DESTINATION's mbarrier.  The MMA thread of each CTA
//     waits on its own mbarrier for 8 x 2.5 KB: no generic-proxy remote stores, no proxy fence over remote traffic, no
//     cluster barrier inside the time loop.  Double-buffered h and staging make the reuse distances safe (see below).
//   * y / saved-gate stores to global memory are issued AFTER the h push, off the recurrence's critical path.
// h buffer layout (per buffer): [K/8 chunks][hi | lo][16 rows][8 bf16] = UMMA canonical K-major no-swizzle core matrices
// with LBO = 512 B; a rank's slice (its 5 chunks, hi and lo) is one contiguous 2 560 B range = one bulk copy.
// Reuse distances: step s reads buffer s&1 and its epilogue fills buffer (s+1)&1 of all CTAs.  A peer can only be
// writing h_{s+1} into my buffer s&1 after it has received MY h_s, which I send after my step-s MMA has completed --
// so no copy ever lands in a buffer an MMA is still reading; the same argument two steps apart covers the staging pair.
// Replaces nn.GRU's recurrence at scripts/model/hierarchy_net.py:144 (H = 300) and :232 (H = 64).
#include "common.cuh"
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace {

constexpr int CL = 8;          // CTAs per cluster
constexpr int TM = 128;        // UMMA M (gate rows incl. padding)
constexpr int TMEM_COLS = 512; // D at columns [0,NB), W_slice hi at [64, 64+K/2), lo right after
constexpr int A_COL = 64;
// NB = batch rows per cluster task (= UMMA N); epilogue warps: 4 at NB = 16, 8 above (warp 0 issues the MMAs)
__host__ __device__ constexpr int epi_warps(int NB) { return NB == 16 ? 4 : 8; }
__host__ __device__ constexpr int block_threads(int NB) { return 32 + 32 * epi_warps(NB); }
constexpr size_t MIN_SMEM = 120 * 1024;  // > half of the SM's shared memory: one CTA per SM, so the 512-column TMEM
                                         // allocation can never wait on a co-resident CTA of the same cluster

struct Tc2Params {
    const float* gi;       // [M,T,2,3H]
    const float* w_hh[2];  // [3H,H]
    const float* b_hh[2];  // [3H]
    float* y;              // [M,T,2H]
    float* gates;          // [M,T,2,4H] or nullptr
    int M, T, H, HSP, n_chunks;
    int M_gates;           // gates are stored for batch rows < M_gates only (the rows whose BPTT will run)
    long long* dbg;        // optional [T+1][8] clock64 samples of cluster 0 / rank 0 (phase timing), nullptr otherwise
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
// wait with acquire at CLUSTER scope: the bytes were written by st.async from the other CTAs of the cluster
__device__ __forceinline__ void mbw_cluster(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = su32(bar);
    uint32_t done = 0;
    long long t0 = clock64();
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
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
// 16 bytes to another CTA's shared memory; completes 16 bytes of the transaction count of that CTA's mbarrier
__device__ __forceinline__ void st_async16(uint32_t dst_cluster, const uint4& v, uint32_t mbar_cluster) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(dst_cluster), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar_cluster) : "memory");
}
__device__ __forceinline__ uint64_t mkd(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]
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

// shared memory map (bytes): barriers | tmem slot | { h[2] | G | hown | out staging }, the braces aliasing the one-time
// fp32 staging of the CTA's W_hh rows (consumed into tensor memory before the first task starts)
struct Tc2Layout {
    int kc;  // K chunks = CL * HSP / 8
    size_t b_bytes, slice_bytes, off_h, off_g, off_hown, off_out, off_bar, off_w, total;
    int orow;   // floats per (array, batch row) line of the output staging (HSP + 4: conflict-free 128-bit stores)
    __host__ __device__ Tc2Layout(int HSP, int H, int NB) {
        kc = CL * HSP / 8;
        b_bytes = (size_t)kc * 2 * NB * 16;
        slice_bytes = (size_t)(HSP / 8) * 2 * NB * 16;
        off_bar = 0;
        off_h = 128;
        off_w = 128;
        off_g = off_h + 2 * b_bytes;
        off_hown = off_g + (size_t)TM * (NB + 1) * 4;
        off_out = (off_hown + (size_t)HSP * NB * 4 + 15) / 16 * 16;
        orow = HSP + 4;
        const size_t loop_end = off_out + (size_t)5 * NB * orow * 4;   // y | r | z | n | hn lines of one step
        const size_t w_end = off_w + (size_t)3 * HSP * H * 4;          // fp32 W_hh rows of this CTA, bulk-copied once
        total = loop_end > w_end ? loop_end : w_end;
        total = (total + 127) / 128 * 128;
        if (total < MIN_SMEM) total = MIN_SMEM;
    }
};

template <int NB>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(block_threads(NB), 1) gru_seq_fwd_tc2_kernel(Tc2Params p) {
    constexpr int NEW = epi_warps(NB);          // epilogue warps
    constexpr int TNT = block_threads(NB);
    constexpr int NET = 32 * NEW;               // epilogue threads
    constexpr int CPW = NB / (NEW / 4);         // accumulator columns handled per epilogue warp
    extern __shared__ __align__(128) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
    const int dir = cluster_id & 1;
    const int H = p.H, HSP = p.HSP, T = p.T, M = p.M;
    const Tc2Layout L(HSP, H, NB);
    const int KC = L.kc;                 // K chunks (K = 8*KC)
    const int CPC = HSP / 8;             // chunks owned per CTA
    unsigned char* hbuf = smem + L.off_h;               // [2][KC][2][NB][16 B]
    float* G = reinterpret_cast<float*>(smem + L.off_g);         // [TM][NB+1]
    float* hown = reinterpret_cast<float*>(smem + L.off_hown);   // [HSP][NB]
    float* outst = reinterpret_cast<float*>(smem + L.off_out);   // [5][NB][orow]
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* bar_full = bar_mma + 1;                   // [2]: h buffer b complete (8 x slice_bytes landed)
    uint64_t* bar_w = bar_mma + 3;                      // one-time: the W_hh rows have landed in shared memory
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 4);
    const float* wrows = reinterpret_cast<const float*>(smem + L.off_w);   // [3*HSP][H]
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = ha2g_warp_id();   // provably warp-uniform
    const int j0 = rank * HSP;
    // kernel parameters into registers up front: indexing the by-value struct dynamically (p.w_hh[dir]) would spill it to
    // local memory and turn every global access below into a generic LD.E / ST.E behind an LDL
    const float* __restrict__ W = dir ? p.w_hh[1] : p.w_hh[0];
    const float* __restrict__ b_hh = dir ? p.b_hh[1] : p.b_hh[0];
    const float* __restrict__ g_gi = p.gi;
    float* __restrict__ g_y = p.y;
    float* __restrict__ g_gates = p.gates;
    long long* g_dbg = p.dbg;
    const int M_gates = p.M_gates;
    const uint32_t tx_bytes = (uint32_t)(CL * L.slice_bytes);

    const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0;
    if (dbg_on && tid == 0) p.dbg[p.T * 8 + 0] = clock64();
    if (tid == 0) {
        mbi(bar_mma, 1);
        mbi(bar_full + 0, 1);
        mbi(bar_full + 1, 1);
        mbi(bar_w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(su32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t tmem_d = tmem_base;
    const uint32_t tmem_ahi = tmem_base + A_COL, tmem_alo = tmem_base + A_COL + (uint32_t)KC * 4;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

    // epilogue roles: warp -> TMEM lane quarter (warp & 3) and, with 8 warps, column half (warp - 1) / 4;
    // gate item = (chunk cc of the CTA, batch row b)
    const int et = tid - 32;                        // 0..NET-1 for epilogue threads
    const bool is_epi = warp >= 1;
    const int q = warp & 3;
    const int col0 = is_epi ? ((warp - 1) >> 2) * CPW : 0;   // first accumulator column of this warp
    const int cc = is_epi ? et / NB : 0, bb = is_epi ? et % NB : 0;
    const bool has_item = is_epi && cc < CPC;

    // ---- one-time: this CTA's W_hh rows (r | z | n of units j0..j0+HSP) as bf16 hi/lo into tensor memory -------------
    // global -> shared: one bulk copy per gate row (H*4 contiguous bytes), all in flight at once
    {
        int nvalid = 0;   // rows of this CTA that exist (j < H); identical formula in every thread
        for (int g = 0; g < 3; ++g) { const int left = H - j0; nvalid += left <= 0 ? 0 : (left < HSP ? left : HSP); }
        if (tid == 0 && nvalid > 0) mb_expect_tx(bar_w, (uint32_t)((size_t)nvalid * H * 4));
        __syncthreads();
        {   // rows split evenly over the 5 warps; one elected lane per warp issues its rows from uniform registers
            const int per_warp = (3 * HSP + TNT / 32 - 1) / (TNT / 32);
            const int r0 = warp * per_warp, r1 = min(3 * HSP, r0 + per_warp);
            if (ha2g_elect_one()) {
                for (int r = r0; r < r1; ++r) {
                    const int g = r / HSP, u = r % HSP, j = j0 + u;
                    if (j < H)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(su32(wrows + (size_t)r * H)), "l"(W + ((size_t)g * H + j) * H), "r"((uint32_t)(H * 4)),
                                       "r"(su32(bar_w)) : "memory");
                }
            }
            __syncwarp();
        }
        if (nvalid > 0) mbw(bar_w, 0);
        if (dbg_on && tid == 0) g_dbg[p.T * 8 + 1] = clock64();
    }
    // shared -> TMEM: lane = gate row, 32-bit column c*4+i = the bf16 pair (k = 8c+2i, 8c+2i+1): the A-operand layout of
    // kind::f16 with A in tensor memory
    if (is_epi && warp <= 4) {   // one warp per TMEM lane quarter
        const int row = q * 32 + lane;
        const int g = row / HSP, u = row % HSP, j = j0 + u;
        const bool valid = g < 3 && j < H;
        const float* src = wrows + (size_t)(valid ? row : 0) * H;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        for (int c = 0; c < KC; ++c) {
            float v[8];
            if (valid && c * 8 + 7 < H) {
                const float4 a = *reinterpret_cast<const float4*>(src + c * 8), b = *reinterpret_cast<const float4*>(src + c * 8 + 4);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = (valid && c * 8 + i < H) ? src[c * 8 + i] : 0.f;
            }
            uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
            split2g(v[0], v[1], h0, l0); split2g(v[2], v[3], h1, l1);
            split2g(v[4], v[5], h2, l2); split2g(v[6], v[7], h3, l3);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                         ::"r"(tmem_ahi + lane_addr + (uint32_t)c * 4), "r"(h0), "r"(h1), "r"(h2), "r"(h3) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                         ::"r"(tmem_alo + lane_addr + (uint32_t)c * 4), "r"(l0), "r"(l1), "r"(l2), "r"(l3) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (dbg_on && tid == 0) g_dbg[p.T * 8 + 2] = clock64();
    // B descriptors: chunk stride (K direction, LBO) = 2*NB*16 (hi and lo of a chunk are adjacent), 8-row groups 128 B apart
    const uint32_t lbo = 2 * NB * 16;
    const uint64_t dbh0 = mkd(su32(hbuf), lbo, 128), dbl0 = mkd(su32(hbuf + NB * 16), lbo, 128);
    const uint64_t dbh1 = mkd(su32(hbuf + L.b_bytes), lbo, 128), dbl1 = mkd(su32(hbuf + L.b_bytes + NB * 16), lbo, 128);
    const uint64_t b_step = (uint64_t)((2 * lbo) >> 4);   // one K = 16 step = two chunks
    // hidden-side biases of this thread's 8 units (constant over the sequence)
    float bhr[8], bhz[8], bhn[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int j = j0 + cc * 8 + i;
        const bool okj = has_item && j < H;
        bhr[i] = okj ? b_hh[j] : 0.f; bhz[i] = okj ? b_hh[H + j] : 0.f; bhn[i] = okj ? b_hh[2 * H + j] : 0.f;
    }
    // copy-out roles (fixed per thread; NB * HSP/4 <= 2 * NET float4 per line group)
    int co_n = 0, co_rb[2] = {0, 0}, co_f4[2] = {0, 0};
    if (is_epi) {
        const int q4 = HSP / 4;
        for (int k = 0; k < 2; ++k) {
            const int idx = et + k * NET;
            if (idx < NB * q4) { co_rb[k] = idx / q4; co_f4[k] = idx % q4; co_n = k + 1; }
        }
    }
    uint32_t it = 0;             // running step counter: phase parity of bar_mma
    uint32_t full_ph0 = 0, full_ph1 = 0;  // phase parities of bar_full (MMA thread only)

    for (int task = cluster_id >> 1; task < p.n_chunks; task += n_clusters >> 1) {
        const int m0 = task * NB;
        // h_{-1} = 0: buffer 0 (hi and lo) and the fp32 master copy
        for (int e = tid; e < KC * 2 * NB; e += TNT) reinterpret_cast<uint4*>(hbuf)[e] = make_uint4(0, 0, 0, 0);
        for (int e = tid; e < HSP * NB; e += TNT) hown[e] = 0.f;
        asm volatile("fence.proxy.async;" ::: "memory");
        if (tid == 0) {   // h_0 lands in buffer 1 (consumed by step 1), h_1 in buffer 0 (consumed by step 2)
            if (T >= 2) mb_expect_tx(bar_full + 1, tx_bytes);
            if (T >= 3) mb_expect_tx(bar_full + 0, tx_bytes);
        }
        cluster.sync();
        if (dbg_on && tid == 0) g_dbg[p.T * 8 + 3] = clock64();
        // x-side pre-activations of this thread's 8 units for one step (independent of the recurrence).  They are loaded
        // one step AHEAD, before the y / gate copy-out of the current step fills the load/store queue, so that they are
        // in registers long before the gate math needs them.
        float gir[8], giz[8], gin[8];
        auto load_gi = [&](int s_next) {
#pragma unroll
            for (int i = 0; i < 8; ++i) gir[i] = giz[i] = gin[i] = 0.f;
            const int b = m0 + bb;
            const int jbase = j0 + cc * 8;
            if (!(has_item && b < M) || s_next >= T) return;
            const int tn = dir == 0 ? s_next : T - 1 - s_next;
            const float* g = g_gi + (((size_t)b * T + tn) * 2 + dir) * 3 * H;
            if (jbase + 7 < H) {   // whole 8-unit chunk valid -> 128-bit accesses (rows are 16-byte aligned)
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
        };
        if (is_epi) load_gi(0);
        for (int s = 0; s < T; ++s, ++it) {
            const int t = dir == 0 ? s : T - 1 - s;
            const int cur = s & 1;
            // ---- tensor core: D[128 x 16] = W_slice * h_{t-1}^T ------------------------------------------------
            if (warp == 0) {   // whole warp, converged; one elected lane issues
                if (dbg_on && lane == 0) g_dbg[s * 8 + 7] = clock64();
                // (measured: splitting the wait into per-rank-pair barriers so that the MMAs start under the arrival of the
                // remaining slices LOST 4-14 % -- the slices land together, and every extra acquire wait costs a CCTL.IVALL)
                if (s > 0) {
                    mbw_cluster(bar_full + cur, cur ? full_ph1 : full_ph0);
                    if (cur) full_ph1 ^= 1; else full_ph0 ^= 1;
                }
                if (dbg_on && lane == 0) g_dbg[s * 8 + 0] = clock64();
                // st.async data (generic proxy of the peers) -> visible to the tensor core's async proxy
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (ha2g_elect_one()) {
                    if (s > 0 && s + 2 <= T - 1) mb_expect_tx(bar_full + cur, tx_bytes);   // h_{s+1} will land here
                    uint32_t ah = tmem_ahi, al = tmem_alo;
                    uint64_t dbh = cur ? dbh1 : dbh0, dbl = cur ? dbl1 : dbl0;
                    mma16_ts(tmem_d, ah, dbh, idesc, 0u);
                    mma16_ts(tmem_d, ah, dbl, idesc, 1u);
                    mma16_ts(tmem_d, al, dbh, idesc, 1u);
#pragma unroll 4
                    for (int ks = 1; ks < KC / 2; ++ks) {
                        ah += 8; al += 8; dbh += b_step; dbl += b_step;
                        mma16_ts(tmem_d, ah, dbh, idesc, 1u);
                        mma16_ts(tmem_d, ah, dbl, idesc, 1u);
                        mma16_ts(tmem_d, al, dbh, idesc, 1u);
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(bar_mma)) : "memory");
                }
                __syncwarp();
                if (dbg_on && lane == 0) g_dbg[s * 8 + 1] = clock64();
            }
            if (is_epi) {
                // (the x-side pre-activations gir/giz/gin of this step were prefetched before the previous step's copy-out)
                const int b = m0 + bb;
                const int jbase = j0 + cc * 8;
                const bool live = has_item && b < M;
                mbw(bar_mma, it & 1);
                if (dbg_on && tid == 32) g_dbg[s * 8 + 2] = clock64();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                {   // TMEM -> shared: lane (gate row) q*32+lane, this warp's CPW batch columns, 8 at a time
                    uint32_t r[CPW];
                    const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)col0;
#pragma unroll
                    for (int c8 = 0; c8 < CPW / 8; ++c8)
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                            : "=r"(r[c8 * 8 + 0]), "=r"(r[c8 * 8 + 1]), "=r"(r[c8 * 8 + 2]), "=r"(r[c8 * 8 + 3]),
                              "=r"(r[c8 * 8 + 4]), "=r"(r[c8 * 8 + 5]), "=r"(r[c8 * 8 + 6]), "=r"(r[c8 * 8 + 7])
                            : "r"(taddr + (uint32_t)(c8 * 8)));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    float* grow = G + (size_t)(q * 32 + lane) * (NB + 1) + col0;
#pragma unroll
                    for (int i = 0; i < CPW; ++i) grow[i] = __uint_as_float(r[i]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                asm volatile("bar.sync 1, %0;" ::"n"(NET) : "memory");  // the epilogue warps
                if (dbg_on && tid == 32) g_dbg[s * 8 + 3] = clock64();
                float hnew[8], sr[8], sz[8], sn[8], shn[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) hnew[i] = sr[i] = sz[i] = sn[i] = shn[i] = 0.f;
                if (has_item) {
                    // branch-free over the 8 units (a per-unit `if` serialises the eight ~150-cycle dependency chains)
                    float hr8[8], hz8[8], hn8[8], hp8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int u = cc * 8 + i;
                        hr8[i] = G[(size_t)u * (NB + 1) + bb] + bhr[i];
                        hz8[i] = G[(size_t)(HSP + u) * (NB + 1) + bb] + bhz[i];
                        hn8[i] = G[(size_t)(2 * HSP + u) * (NB + 1) + bb] + bhn[i];
                        hp8[i] = hown[u * NB + bb];
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        // ex2-based sigmoid/tanh (abs error ~1e-7, far inside the 1e-3 parity bar; 3x fewer instructions
                        // than expf/tanhf on the serial critical path of the recurrence)
                        const float r = __fdividef(1.f, 1.f + __expf(-(gir[i] + hr8[i])));
                        const float z = __fdividef(1.f, 1.f + __expf(-(giz[i] + hz8[i])));
                        const float n = 1.f - __fdividef(2.f, 1.f + __expf(2.f * (gin[i] + r * hn8[i])));
                        const bool ok = live && (jbase + i < H);
                        hnew[i] = ok ? (1.f - z) * n + z * hp8[i] : 0.f;
                        sr[i] = r; sz[i] = z; sn[i] = n; shn[i] = hn8[i];
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) hown[(cc * 8 + i) * NB + bb] = hnew[i];
                }
                if (dbg_on && tid == 32) g_dbg[s * 8 + 4] = clock64();
                if (s < T - 1) {
                    // ---- push h_t: every item thread sends its packed 8-unit chunk (hi and lo, 16 bytes each) straight from
                    // registers to all 8 CTAs with st.async, which also signs the bytes off on the DESTINATION's mbarrier:
                    // no staging buffer, no proxy fence, no barrier, and one DSMEM hop of latency instead of a trip
                    // through the bulk-copy engine (measured: ~1.1 us -> see tools/time_gru_tc.py)
                    if (has_item) {
                        uint4 h4, l4;
                        split2g(hnew[0], hnew[1], h4.x, l4.x); split2g(hnew[2], hnew[3], h4.y, l4.y);
                        split2g(hnew[4], hnew[5], h4.z, l4.z); split2g(hnew[6], hnew[7], h4.w, l4.w);
                        const uint32_t dst_hi = su32(hbuf) + (uint32_t)((size_t)(cur ^ 1) * L.b_bytes +
                                                                       ((size_t)((rank * CPC + cc) * 2 + 0) * NB + bb) * 16);
                        const uint32_t dst_lo = dst_hi + NB * 16;
                        const uint32_t bar = su32(bar_full + (cur ^ 1));
#pragma unroll
                        for (uint32_t d = 0; d < CL; ++d) {
                            const uint32_t rbar = mapa(bar, d);
                            st_async16(mapa(dst_hi, d), h4, rbar);
                            st_async16(mapa(dst_lo, d), l4, rbar);
                        }
                    }
                    if (dbg_on && tid == 32) g_dbg[s * 8 + 5] = clock64();
                }
                load_gi(s + 1);   // next step's x-side pre-activations: issued before the stores below
                // ---- y and the saved gates: off the critical path (the next step's MMA is already being fed), staged
                // through shared memory so that the global stores are 160-byte runs instead of 32 scattered sectors per
                // instruction (the scattered version kept the LSU busy for ~1 000 cycles per step)
                {
                    const int OR = L.orow;
                    const int narr = g_gates != nullptr ? 5 : 1;
                    if (has_item) {
                        float* o = outst + (size_t)bb * OR + cc * 8;
                        reinterpret_cast<float4*>(o)[0] = make_float4(hnew[0], hnew[1], hnew[2], hnew[3]);
                        reinterpret_cast<float4*>(o)[1] = make_float4(hnew[4], hnew[5], hnew[6], hnew[7]);
                        if (narr == 5) {
                            const size_t as = (size_t)NB * OR;
                            reinterpret_cast<float4*>(o + as)[0] = make_float4(sr[0], sr[1], sr[2], sr[3]);
                            reinterpret_cast<float4*>(o + as)[1] = make_float4(sr[4], sr[5], sr[6], sr[7]);
                            reinterpret_cast<float4*>(o + 2 * as)[0] = make_float4(sz[0], sz[1], sz[2], sz[3]);
                            reinterpret_cast<float4*>(o + 2 * as)[1] = make_float4(sz[4], sz[5], sz[6], sz[7]);
                            reinterpret_cast<float4*>(o + 3 * as)[0] = make_float4(sn[0], sn[1], sn[2], sn[3]);
                            reinterpret_cast<float4*>(o + 3 * as)[1] = make_float4(sn[4], sn[5], sn[6], sn[7]);
                            reinterpret_cast<float4*>(o + 4 * as)[0] = make_float4(shn[0], shn[1], shn[2], shn[3]);
                            reinterpret_cast<float4*>(o + 4 * as)[1] = make_float4(shn[4], shn[5], shn[6], shn[7]);
                        }
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(NET) : "memory");
                    // copy-out: thread -> up to two fixed (batch row, float4) positions of every line group
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        if (k < co_n) {
                            const int rb = co_rb[k], f4 = co_f4[k];
                            const int bg = m0 + rb, jg = j0 + f4 * 4;
                            if (bg < M && jg < H) {               // H % 4 == 0: a float4 is entirely valid or entirely padding
                                const size_t row = (size_t)bg * T + t;
                                const float* src = outst + (size_t)rb * OR + f4 * 4;
                                *reinterpret_cast<float4*>(g_y + row * 2 * H + dir * H + jg) = *reinterpret_cast<const float4*>(src);
                                if (narr == 5 && bg < M_gates) {
                                    float* gd = g_gates + (row * 2 + dir) * 4 * H + jg;
#pragma unroll
                                    for (int a = 1; a < 5; ++a)
                                        *reinterpret_cast<float4*>(gd + (size_t)(a - 1) * H) =
                                            *reinterpret_cast<const float4*>(src + (size_t)a * NB * OR);
                                }
                            }
                        }
                    }
                }
                if (dbg_on && tid == 32) g_dbg[s * 8 + 6] = clock64();
            }
        }
        if (dbg_on && tid == 0) g_dbg[p.T * 8 + 4] = clock64();
        // every CTA is past its last MMA (and hence past every copy into its buffers) before buffers are re-zeroed / freed
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        cluster.sync();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

}  // namespace

// batch rows per cluster task for a batch of M rows: the smallest NB whose chunks fit the 16 resident clusters in one wave
static int pick_nb(int M) { return M <= 128 ? 16 : (M <= 256 ? 32 : 48); }

// 1 through *ok if the second-generation tensor-core recurrence can serve hidden size H (gate rows 3*HSP <= 128, the
// W_hh slice fits the TMEM columns behind the accumulator, one 8-unit item per epilogue thread at every NB).
HA2G_API int ha2g_gru_tc2_supported(int H, int* ok) {
    const int HSP = ((H + CL - 1) / CL + 7) / 8 * 8;
    const int kc = CL * HSP / 8;
    bool good = 3 * HSP <= TM && kc % 2 == 0 && A_COL + kc * 8 <= TMEM_COLS && H % 4 == 0;
    for (int nb = 16; nb <= 48 && good; nb += 16) {
        const Tc2Layout L(HSP, H, nb);
        good = (HSP / 8) * nb <= 32 * epi_warps(nb) && nb * (HSP / 4) <= 2 * 32 * epi_warps(nb) && L.total <= 227 * 1024;
    }
    *ok = good ? 1 : 0;
    return 0;
}

template <int NB>
static int launch_tc2(Tc2Params& p, cudaStream_t stream) {
    p.n_chunks = (p.M + NB - 1) / NB;
    const Tc2Layout L(p.HSP, p.H, NB);
    cudaError_t e = cudaFuncSetAttribute(gru_seq_fwd_tc2_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    if (e != cudaSuccess) return (int)e;
    int clusters = 2 * p.n_chunks;
    if (clusters > 16) clusters = 16;
    gru_seq_fwd_tc2_kernel<NB><<<clusters * CL, block_threads(NB), L.total, stream>>>(p);
    HA2G_RETURN_LAST();
}

extern "C" int ha2g_gru_seq_fwd_tc2_dbg(const float*, const float*, const float*, const float*, const float*, float*, float*,
                                        int, int, int, int, int, long long*, cudaStream_t);
// All T steps of one bidirectional layer, forward, on tcgen05 with W_hh resident in tensor memory (see the file header).
// gi must hold x W_ih^T + b_ih.  gates (nullable) receives r | z | n | hn for the batch rows < M_gates only (the rows
// whose backward pass will run; M_gates = M saves all).  Replaces the recurrence of nn.GRU at
// scripts/model/hierarchy_net.py:144 / :232.
HA2G_API int ha2g_gru_seq_fwd_tc2(const float* gi, const float* w_hh_f, const float* w_hh_r, const float* b_hh_f,
                                  const float* b_hh_r, float* y, float* gates, int M, int M_gates, int T, int H,
                                  cudaStream_t stream) {
    return ha2g_gru_seq_fwd_tc2_dbg(gi, w_hh_f, w_hh_r, b_hh_f, b_hh_r, y, gates, M, M_gates, T, H, 0, nullptr, stream);
}

// Same, with an explicit rows-per-cluster choice (nb = 16 / 32 / 48; 0 = automatic) and an optional device buffer dbg
// [T+1][8] of clock64() samples (cluster 0, rank 0) for phase timing:
// 7 = MMA thread reaches the h-arrival wait, 0 = h arrived / MMA issue starts, 1 = MMAs issued + committed,
// 2 = epilogue woken by the commit, 3 = accumulator transposed through shared memory, 4 = gate math done,
// 5 = h_t pushed to the peers, 6 = y / gates stored; row T: 0 = kernel entry, 1 = W_hh rows in shared memory,
// 2 = W_hh in tensor memory, 3 = time loop starts, 4 = time loop done.
HA2G_API int ha2g_gru_seq_fwd_tc2_dbg(const float* gi, const float* w_hh_f, const float* w_hh_r, const float* b_hh_f,
                                      const float* b_hh_r, float* y, float* gates, int M, int M_gates, int T, int H, int nb,
                                      long long* dbg, cudaStream_t stream) {
    if (M <= 0 || T <= 0) return 0;
    Tc2Params p{};
    p.dbg = dbg;
    p.gi = gi; p.w_hh[0] = w_hh_f; p.w_hh[1] = w_hh_r; p.b_hh[0] = b_hh_f; p.b_hh[1] = b_hh_r;
    p.y = y; p.gates = gates; p.M = M; p.T = T; p.H = H;
    p.M_gates = gates != nullptr ? (M_gates < M ? M_gates : M) : 0;
    p.HSP = ((H + CL - 1) / CL + 7) / 8 * 8;
    if (nb == 0) nb = pick_nb(M);
    if (nb == 16) return launch_tc2<16>(p, stream);
    if (nb == 32) return launch_tc2<32>(p, stream);
    if (nb == 48) return launch_tc2<48>(p, stream);
    return (int)cudaErrorInvalidValue;
}
