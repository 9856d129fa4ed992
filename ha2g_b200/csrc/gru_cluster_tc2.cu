// Persistent cluster GRU recurrence, second generation (forward) -- the fused GRU kernel of the step.
//
// Decomposition: an 8-CTA cluster per (direction, row chunk of the batch).  A B200 keeps 15 such clusters resident (a
// cluster lives inside one GPC; measured with cudaOccupancyMaxActiveClusters and tools/gru_waves.py), i.e. 7 per
// direction: the rows per chunk are the smallest of 16 / 20 / 32 / 48 / 56 / 64 whose chunks fit that one wave (the
// 3B = 384-row cascade of the training step: 56 rows; a 128-row batch: 20).  Until round 2 the launcher assumed 16 resident
// clusters: the 16th waited for a whole sequence and every recurrence of the step ran at twice its one-wave time.
// Chunks of 20 rows and more run as NH = 2 independent half-tasks of NB rows each inside the cluster (own accumulator
// columns, h buffers, mbarriers and five epilogue warps per half; the MMA warp serves them alternately): the per-step
// chain MMA -> TMEM read -> gate math -> all-gather of one half hides under the other's (384 rows: 199 -> 152 us).
// CTA `rank` owns HSP = 40 hidden units = 3*HSP gate rows (r | z | n, zero-padded to M = 128) of W_hh for all T steps, and
// every step computes
//     gates^T[128 x NN] = W_slice[128 x K] * h_{t-1}^T[K x NN]          (K = 8*HSP = 320, bf16x3 split, fp32 accumulate)
// on tcgen05 (NN = NB rounded up to the UMMA N granularity of 16; accumulator columns >= NB are never read): the same 60
// tcgen05.mma per step carry NB columns, the fixed per-step latency (mbarrier wake-ups, TMEM round trip, DSMEM flight)
// is amortised over NB rows, and 4 / 8 / 10 epilogue warps keep the gate math at one 8-unit item per thread.
// History, each item taken from phase timings (tools/time_gru_tc.py; first kernel: 10 600 cycles per step =
// MMA issue 4 170 + gate math/stores 3 230 + DSMEM push 1 950 + cluster barrier 910):
//   * W_slice lives in TENSOR MEMORY (tcgen05.st once per layer, 320 columns: hi | lo) and is the MMA's A operand
//     straight from TMEM: 57 dependent MMAs of N = 48 take 1 470 cycles (tools/ubench/mma_chain.cu: 26 cycles each,
//     the 128*N/256 floor; with A in shared memory 46).
//   * h_t travels by bulk asynchronous copies (cp.async.bulk shared::cta -> shared::cluster): every item thread writes
//     its packed chunk into the CTA's own slice of the next h buffer, one lane hands that slice to the copy engine once
//     per peer, and each copy completes its bytes on the DESTINATION's mbarrier.  The MMA thread waits on its own
//     mbarrier for the 7 peer slices plus a local arrival for the own one: no cluster barrier inside the time loop, and
//     no remote traffic through the load/store unit (an intermediate version pushed with st.async from registers: fine
//     at NB = 16, but at NB = 48 the 61 KB per step occupied the LSU for ~3 000 cycles per step).
//   * y / saved-gate stores to global memory are issued AFTER the h push, off the recurrence's critical path.
// h buffer layout (per buffer): [K/8 chunks][hi | lo][NR rows][8 bf16] = UMMA canonical K-major no-swizzle core matrices
// (NR = NB rounded up to 8) with LBO = 2*NR*16 B; a rank's slice (its 5 chunks, hi and lo) is one contiguous range = one bulk copy.
// Reuse distances: step s reads buffer s&1 and its epilogue fills buffer (s+1)&1 of all CTAs.  A peer can only be
// writing h_{s+1} into my buffer s&1 after it has received MY h_s, which I send after my step-s MMA has completed --
// so no copy ever lands in a buffer an MMA is still reading; and my own slice of buffer (s+1)&1, the source of the
// copies of step s, is next written at step s+2, after every peer's h_{s+1} (hence their receipt of my h_s) has arrived.
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
// NB = batch rows per cluster task, NN = UMMA N (>= NB, multiple of 16; accumulator columns >= NB are never read).
// Epilogue warps (warp 0 issues the MMAs): one (8-unit chunk, batch row) gate item per thread -> 4 / 8 / 10 warps.
// NH = 2: the cluster task is split into two independent halves of NB rows each, 5 epilogue warps per half (see header).
__host__ __device__ constexpr int epi_warps(int NB, int NH = 1) { return NH == 2 ? 10 : (NB <= 24 ? 4 : (NB <= 48 ? 8 : 10)); }
__host__ __device__ constexpr int mma_n(int NB) { return (NB + 15) / 16 * 16; }
// accumulator columns per epilogue warp: the NB columns are split over the 1 - 3 warps that share a TMEM lane quarter, in
// multiples of 8; this is the largest share (quarters served by 2 of the 10 warps)
__host__ __device__ constexpr int cols_per_warp(int NB, int NH = 1) {
    return NH == 2 ? (NB + 7) / 8 * 8 : ((NB + epi_warps(NB) / 4 - 1) / (epi_warps(NB) / 4) + 7) / 8 * 8;
}
__host__ __device__ constexpr int block_threads(int NB, int NH = 1) { return 32 + 32 * epi_warps(NB, NH); }
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
    int gpt;               // > 0: every task takes gpt of the gate-storing rows and fills up with the others (see row_of)
    int xflags;            // timing experiments only (tools/time_gru_tc.py): 1 = no gi loads, 2 = no global stores, 4 = no staging
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
__device__ __forceinline__ void mb_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(su32(bar)) : "memory");
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

// shared memory map (bytes): barriers | tmem slot | { h[NH][2] | G[NH] | hown[NH] | biases }, the braces aliasing the one-time
// fp32 staging of the CTA's W_hh rows (consumed into tensor memory before the first task starts)
struct Tc2Layout {
    int kc;  // K chunks = CL * HSP / 8
    size_t b_bytes, slice_bytes, off_h, off_g, off_hown, off_bias, off_bar, off_w, total;
    int orow;   // floats per batch row of hown (HSP + 4: 128-bit accesses)
    int grow;   // floats per batch row of the transposed accumulator (3*HSP + 4)
    size_t g_stride, hown_stride;   // per half (NH = 2): each half has its own G / hown / h buffers
    __host__ __device__ Tc2Layout(int HSP, int H, int NB, int NH = 1) {
        const int NR = (NB + 7) / 8 * 8;   // rows per (chunk, hi | lo) block of the B operand: whole 8-row core matrices
        kc = CL * HSP / 8;
        b_bytes = (size_t)kc * 2 * NR * 16;
        slice_bytes = (size_t)(HSP / 8) * 2 * NR * 16;
        off_bar = 0;
        off_h = 128;
        off_w = 128;
        off_g = off_h + (size_t)NH * 2 * b_bytes;
        orow = HSP + 4;
        grow = 3 * HSP + 4;
        const size_t gsz = (size_t)NB * grow * 4, hsz = (size_t)NB * orow * 4;
        g_stride = gsz;
        hown_stride = hsz;
        off_hown = off_g + NH * gsz;
        size_t loop_end = off_hown + NH * hsz;
        off_bias = loop_end;                                           // hidden-side biases of the CTA's units: [3][HSP]
        loop_end += (size_t)3 * HSP * 4;
        const size_t w_end = off_w + (size_t)3 * HSP * H * 4;          // fp32 W_hh rows of this CTA, bulk-copied once
        total = loop_end > w_end ? loop_end : w_end;
        total = (total + 127) / 128 * 128;
        if (total < MIN_SMEM) total = MIN_SMEM;
    }
};

template <int NB, int NH>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(block_threads(NB, NH), 1) gru_seq_fwd_tc2_kernel(Tc2Params p) {
    constexpr int NEW = epi_warps(NB, NH);      // epilogue warps
    constexpr int TNT = block_threads(NB, NH);
    constexpr int GW = NEW / NH;                // epilogue warps per half
    constexpr int NET = 32 * GW;                // epilogue threads per half (one named barrier per half)
    constexpr int NN = mma_n(NB);               // UMMA N (multiple of 16)
    // rows per (chunk, hi | lo) block of the B operand.  When NR < NN (NB = 20, 56) the MMA's last 8-row group reads the
    // first rows of the NEXT block: finite or not, they only reach accumulator columns >= NR, which nobody reads.
    constexpr int NR = (NB + 7) / 8 * 8;
    constexpr int CPW = cols_per_warp(NB, NH);  // most accumulator columns any epilogue warp handles
    extern __shared__ __align__(128) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
    const int dir = cluster_id & 1;
    const int H = p.H, HSP = p.HSP, T = p.T, M = p.M;
    const Tc2Layout L(HSP, H, NB, NH);
    const int KC = L.kc;                 // K chunks (K = 8*KC)
    const int CPC = HSP / 8;             // chunks owned per CTA
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = ha2g_warp_id();   // provably warp-uniform
    const bool is_epi = warp >= 1;
    const int grp = is_epi ? (warp - 1) / GW : 0;       // which half of the task this epilogue warp serves
    unsigned char* hbuf_all = smem + L.off_h;           // [NH][2][KC][2][NR][16 B]
    unsigned char* hbuf = hbuf_all + (size_t)grp * 2 * L.b_bytes;
    float* G = reinterpret_cast<float*>(smem + L.off_g + grp * L.g_stride);          // [NB][grow]: W_hh h of batch row b, gate rows r | z | n
    float* hown = reinterpret_cast<float*>(smem + L.off_hown + grp * L.hown_stride); // [NB][orow]: fp32 master copy of the CTA's units of h
    uint64_t* bar_mma_all = reinterpret_cast<uint64_t*>(smem + L.off_bar);   // [NH]: accumulator of half g complete
    uint64_t* bar_full_all = bar_mma_all + 2;           // [NH][2]: h buffer b of half g complete (7 peer slices + own)
    uint64_t* bar_mma = bar_mma_all + grp;
    uint64_t* bar_full = bar_full_all + 2 * grp;
    uint64_t* bar_w = bar_mma_all + 6;                  // one-time: the W_hh rows have landed in shared memory
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma_all + 7);
    const float* wrows = reinterpret_cast<const float*>(smem + L.off_w);   // [3*HSP][H]
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
    const int xflags = p.xflags;
    const uint32_t tx_bytes = (uint32_t)((CL - 1) * L.slice_bytes);   // the 7 peer slices; the own one is a plain arrival

    const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0;
    if (dbg_on && tid == 0) p.dbg[p.T * 8 + 0] = clock64();
    if (tid == 0) {
        for (int g = 0; g < NH; ++g) {
            mbi(bar_mma_all + g, 1);
            mbi(bar_full_all + 2 * g + 0, 2);   // arrive.expect_tx of the MMA thread + the arrival that signs off the CTA's own slice
            mbi(bar_full_all + 2 * g + 1, 2);
        }
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
    const uint32_t tmem_d = tmem_base + (uint32_t)(grp * NN);   // the half's accumulator: columns [grp*NN, grp*NN + NN)
    const uint32_t tmem_ahi = tmem_base + A_COL, tmem_alo = tmem_base + A_COL + (uint32_t)KC * 4;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

    // epilogue roles: warp -> TMEM lane quarter (warp & 3) and, with 8 warps, column half (warp - 1) / 4;
    // gate item = (chunk cc of the CTA, batch row b)
    const int et = tid - 32 - grp * NET;            // 0..NET-1 for the epilogue threads of a half
    const int q = warp & 3;
    // accumulator columns of this warp: the warps of its half that share TMEM lane quarter q split the NB columns
    int nwq = 0, iwq = 0;
    for (int w = 1 + grp * GW; w < 1 + (grp + 1) * GW; ++w)
        if ((w & 3) == q) { if (w < warp) ++iwq; ++nwq; }
    if (nwq == 0) nwq = 1;
    const int cpw = ((NB + nwq - 1) / nwq + 7) / 8 * 8;
    const int col0 = is_epi ? iwq * cpw : 0;        // first accumulator column of this warp
    // item -> thread: the chunk index runs fastest, so that a warp's 32 items cover ~6 batch rows x the CTA's 160
    // contiguous bytes per row of every global tensor (lanes across batch rows made every 128-bit load / store touch 32
    // different lines and kept the load/store unit busy for ~1 500 cycles per step)
    const int cc = is_epi ? et % CPC : 0, bb = is_epi ? et / CPC : 0;
    const bool has_item = is_epi && bb < NB;

    // ---- one-time: this CTA's W_hh rows (r | z | n of units j0..j0+HSP) as bf16 hi/lo into tensor memory -------------
    // global -> shared: one bulk copy per gate row (H*4 contiguous bytes), all in flight at once
    {
        int nvalid = 0;   // rows of this CTA that exist (j < H); identical formula in every thread
        for (int g = 0; g < 3; ++g) { const int left = H - j0; nvalid += left <= 0 ? 0 : (left < HSP ? left : HSP); }
        if (tid == 0 && nvalid > 0) mb_expect_tx(bar_w, (uint32_t)((size_t)nvalid * H * 4));
        __syncthreads();
        {   // rows split evenly over the 5 warps; one elected lane per warp issues its rows from uniform registers
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
        if (nvalid > 0) mbw(bar_w, 0);
        if (dbg_on && tid == 0) g_dbg[p.T * 8 + 1] = clock64();
    }
    // shared -> TMEM: lane = gate row, 32-bit column c*4+i = the bf16 pair (k = 8c+2i, 8c+2i+1): the A-operand layout of
    // kind::f16 with A in tensor memory
    if (is_epi) {   // a warp can only write its own TMEM lane quarter: the warps that share a quarter split the K chunks
        int nwa = 0, iwa = 0;
        for (int w = 1; w <= NEW; ++w)
            if ((w & 3) == q) { if (w < warp) ++iwa; ++nwa; }
        const int row = q * 32 + lane;
        const int g = row / HSP, u = row % HSP, j = j0 + u;
        const bool valid = g < 3 && j < H;
        const float* src = wrows + (size_t)(valid ? row : 0) * H;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        for (int c = iwa; c < KC; c += nwa) {
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
    // B descriptors: chunk stride (K direction, LBO) = 2*NR*16 (hi and lo of a chunk are adjacent), 8-row groups 128 B apart
    const uint32_t lbo = 2 * NR * 16;
    // (MMA warp: the descriptors of half 0; half g's buffers are 2*b_bytes further on)
    const uint64_t dbh0 = mkd(su32(hbuf_all), lbo, 128), dbl0 = mkd(su32(hbuf_all + NR * 16), lbo, 128);
    const uint64_t dbh1 = mkd(su32(hbuf_all + L.b_bytes), lbo, 128), dbl1 = mkd(su32(hbuf_all + L.b_bytes + NR * 16), lbo, 128);
    const uint64_t half_step = (uint64_t)((2 * L.b_bytes) >> 4);
    const uint64_t b_step = (uint64_t)((2 * lbo) >> 4);   // one K = 16 step = two chunks
    // hidden-side biases of the CTA's units (constant over the sequence) in shared memory, [gate][HSP]: re-read as
    // broadcast 128-bit words every step rather than held in 24 registers per thread (the W_hh staging is dead by now)
    float* bsm = reinterpret_cast<float*>(smem + L.off_bias);
    for (int e = tid; e < 3 * HSP; e += TNT) {
        const int g = e / HSP, j = j0 + e % HSP;
        bsm[e] = j < H ? b_hh[g * H + j] : 0.f;
    }
    // (made visible to the epilogue warps by the cluster.sync() that opens the first task)
    uint32_t it = 0;             // running step counter: phase parity of bar_mma
    uint32_t full_ph = 0;        // phase parities of bar_full[g][b], bit 2*g + b (MMA thread only)

    for (int task = cluster_id >> 1; task < p.n_chunks; task += n_clusters >> 1) {
        // batch row of this thread's item.  Saving the gates costs 4x the store traffic of y, and only the first M_gates
        // rows save them: with consecutive rows per task a third of the clusters would carry all of it (and set the kernel
        // time), so every task takes gpt gate-saving rows and R - gpt of the others.
        int brow;
        {
            const int R = NH * NB, i = grp * NB + bb;
            if (p.gpt == 0) brow = task * R + i;
            else if (i < p.gpt) brow = (task * p.gpt + i < M_gates) ? task * p.gpt + i : M;
            else brow = M_gates + task * (R - p.gpt) + (i - p.gpt);
            if (brow > M) brow = M;
        }
        // h_{-1} = 0: buffer 0 (hi and lo) and the fp32 master copy; buffer 1 too, for the rows NB..NR-1 nobody writes
        for (int e = tid; e < NH * 2 * KC * 2 * NR; e += TNT) reinterpret_cast<uint4*>(hbuf_all)[e] = make_uint4(0, 0, 0, 0);
        {
            float* hz = reinterpret_cast<float*>(smem + L.off_hown);
            for (int e = tid; e < NH * NB * L.orow; e += TNT) hz[e] = 0.f;
        }
        asm volatile("fence.proxy.async;" ::: "memory");
        if (tid == 0) {   // h_0 lands in buffer 1 (consumed by step 1), h_1 in buffer 0 (consumed by step 2)
            for (int g = 0; g < NH; ++g) {
                if (T >= 2) mb_expect_tx(bar_full_all + 2 * g + 1, tx_bytes);
                if (T >= 3) mb_expect_tx(bar_full_all + 2 * g + 0, tx_bytes);
            }
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
            const int b = brow;
            const int jbase = j0 + cc * 8;
            if (!(has_item && b < M) || s_next >= T || (xflags & 1)) return;
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
                // NH = 2: the halves are independent recurrences that share the weights in tensor memory.  While half g's
                // epilogue warps do their gate math and its h travels, the tensor core works on the other half: the
                // per-step chain (MMA -> TMEM read -> gate math -> all-gather) of one half hides under the other's.
#pragma unroll
                for (int g = 0; g < NH; ++g) {
                    uint64_t* bfull = bar_full_all + 2 * g + cur;
                    if (dbg_on && lane == 0 && g == 0) g_dbg[s * 8 + 7] = clock64();
                    // (measured: splitting the wait into per-rank-pair barriers so that the MMAs start under the arrival of
                    // the remaining slices LOST 4-14 % -- the slices land together)
                    if (s > 0) {
                        mbw(bfull, (full_ph >> (2 * g + cur)) & 1u);
                        full_ph ^= 1u << (2 * g + cur);
                    }
                    if (dbg_on && lane == 0 && g == 0) g_dbg[s * 8 + 0] = clock64();
                    // (the peers' slices were written by the bulk-copy engine, the own one behind a proxy fence: all visible
                    // to the tensor core's async proxy once the barrier has completed)
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (ha2g_elect_one()) {
                        if (s > 0 && s + 2 <= T - 1) mb_expect_tx(bfull, tx_bytes);   // h_{s+1} will land here
                        const uint32_t td = tmem_base + (uint32_t)(g * NN);
                        uint32_t ah = tmem_ahi, al = tmem_alo;
                        uint64_t dbh = (cur ? dbh1 : dbh0) + (uint64_t)g * half_step, dbl = (cur ? dbl1 : dbl0) + (uint64_t)g * half_step;
                        mma16_ts(td, ah, dbh, idesc, 0u);
                        mma16_ts(td, ah, dbl, idesc, 1u);
                        mma16_ts(td, al, dbh, idesc, 1u);
#pragma unroll 4
                        for (int ks = 1; ks < KC / 2; ++ks) {
                            ah += 8; al += 8; dbh += b_step; dbl += b_step;
                            mma16_ts(td, ah, dbh, idesc, 1u);
                            mma16_ts(td, ah, dbl, idesc, 1u);
                            mma16_ts(td, al, dbh, idesc, 1u);
                        }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(bar_mma_all + g)) : "memory");
                    }
                    __syncwarp();
                    if (dbg_on && lane == 0 && g == 0) g_dbg[s * 8 + 1] = clock64();
                }
            }
            if (is_epi) {
                // (the x-side pre-activations gir/giz/gin of this step were prefetched before the previous step's copy-out)
                const int b = brow;
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
                        if (c8 * 8 < cpw && col0 + c8 * 8 < NB)
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                            : "=r"(r[c8 * 8 + 0]), "=r"(r[c8 * 8 + 1]), "=r"(r[c8 * 8 + 2]), "=r"(r[c8 * 8 + 3]),
                              "=r"(r[c8 * 8 + 4]), "=r"(r[c8 * 8 + 5]), "=r"(r[c8 * 8 + 6]), "=r"(r[c8 * 8 + 7])
                            : "r"(taddr + (uint32_t)(c8 * 8)));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    // transposed: a warp writes 32 consecutive gate rows of one batch row per instruction (conflict-free),
                    // and the gate math below reads its 8 units of a gate as two 128-bit words
                    const int grow_i = q * 32 + lane;
                    if (grow_i < 3 * HSP) {
                        float* gcol = G + (size_t)col0 * L.grow + grow_i;
#pragma unroll
                        for (int i = 0; i < CPW; ++i)
                            if (i < cpw && col0 + i < NB) gcol[(size_t)i * L.grow] = __uint_as_float(r[i]);
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(NET) : "memory");  // the epilogue warps
                if (dbg_on && tid == 32) g_dbg[s * 8 + 3] = clock64();
                float hnew[8], sr[8], sz[8], sn[8], shn[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) hnew[i] = sr[i] = sz[i] = sn[i] = shn[i] = 0.f;
                if (has_item) {
                    // branch-free over the 8 units (a per-unit `if` serialises the eight ~150-cycle dependency chains)
                    float hr8[8], hz8[8], hn8[8], hp8[8];
                    {
                        const float* gb = G + (size_t)bb * L.grow + cc * 8;
                        const float4 r0 = *reinterpret_cast<const float4*>(gb), r1 = *reinterpret_cast<const float4*>(gb + 4);
                        const float4 z0 = *reinterpret_cast<const float4*>(gb + HSP), z1 = *reinterpret_cast<const float4*>(gb + HSP + 4);
                        const float4 n0 = *reinterpret_cast<const float4*>(gb + 2 * HSP), n1 = *reinterpret_cast<const float4*>(gb + 2 * HSP + 4);
                        const float* hb = hown + (size_t)bb * L.orow + cc * 8;
                        const float4 p0 = *reinterpret_cast<const float4*>(hb), p1 = *reinterpret_cast<const float4*>(hb + 4);
                        hr8[0] = r0.x; hr8[1] = r0.y; hr8[2] = r0.z; hr8[3] = r0.w; hr8[4] = r1.x; hr8[5] = r1.y; hr8[6] = r1.z; hr8[7] = r1.w;
                        hz8[0] = z0.x; hz8[1] = z0.y; hz8[2] = z0.z; hz8[3] = z0.w; hz8[4] = z1.x; hz8[5] = z1.y; hz8[6] = z1.z; hz8[7] = z1.w;
                        hn8[0] = n0.x; hn8[1] = n0.y; hn8[2] = n0.z; hn8[3] = n0.w; hn8[4] = n1.x; hn8[5] = n1.y; hn8[6] = n1.z; hn8[7] = n1.w;
                        hp8[0] = p0.x; hp8[1] = p0.y; hp8[2] = p0.z; hp8[3] = p0.w; hp8[4] = p1.x; hp8[5] = p1.y; hp8[6] = p1.z; hp8[7] = p1.w;
                        const float* bb8 = bsm + cc * 8;
#pragma unroll
                        for (int i = 0; i < 8; ++i) { hr8[i] += bb8[i]; hz8[i] += bb8[HSP + i]; hn8[i] += bb8[2 * HSP + i]; }
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
                    {
                        float* hb = hown + (size_t)bb * L.orow + cc * 8;
                        reinterpret_cast<float4*>(hb)[0] = make_float4(hnew[0], hnew[1], hnew[2], hnew[3]);
                        reinterpret_cast<float4*>(hb)[1] = make_float4(hnew[4], hnew[5], hnew[6], hnew[7]);
                    }
                }
                if (dbg_on && tid == 32) g_dbg[s * 8 + 4] = clock64();
                if (s < T - 1) {
                    // ---- push h_t: every item thread writes its packed 8-unit chunk (hi and lo, 16 bytes each) into the CTA's
                    // OWN slice of the next h buffer; one lane then hands the slice to the bulk-copy engine, once per peer,
                    // each copy signing its bytes off on the DESTINATION's mbarrier, and signs the own slice off locally.
                    // (First version: 16 st.async per thread straight from registers -- lowest latency, but the 61 KB per
                    // step left the SM through the load/store unit at ~21 B/clk and kept it busy for ~3 000 cycles per step,
                    // stalling the y / gate stores and the next step's loads behind it; the copy engine moves the same
                    // bytes without occupying the LSU.)
                    unsigned char* nxt = hbuf + (size_t)(cur ^ 1) * L.b_bytes;
                    if (has_item) {
                        uint4 h4, l4;
                        split2g(hnew[0], hnew[1], h4.x, l4.x); split2g(hnew[2], hnew[3], h4.y, l4.y);
                        split2g(hnew[4], hnew[5], h4.z, l4.z); split2g(hnew[6], hnew[7], h4.w, l4.w);
                        unsigned char* d = nxt + ((size_t)((rank * CPC + cc) * 2 + 0) * NR + bb) * 16;
                        *reinterpret_cast<uint4*>(d) = h4;
                        *reinterpret_cast<uint4*>(d + NR * 16) = l4;
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    }
                    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(NET) : "memory");
                    if (warp == 1 + grp * GW) {   // the first warp of the half
                        if (ha2g_elect_one()) {
                            const uint32_t src = su32(nxt + (size_t)rank * L.slice_bytes);
                            const uint32_t bar = su32(bar_full + (cur ^ 1));
#pragma unroll
                            for (uint32_t d = 1; d < CL; ++d) {   // nearest-rank-first order differs per CTA: no hot destination
                                const uint32_t dst = ((uint32_t)rank + d) & (CL - 1);
                                bulk_s2c(mapa(src, dst), src, (uint32_t)L.slice_bytes, mapa(bar, dst));
                            }
                            mb_arrive(bar_full + (cur ^ 1));
                        }
                        __syncwarp();
                    }
                    if (dbg_on && tid == 32) g_dbg[s * 8 + 5] = clock64();
                }
                load_gi(s + 1);   // next step's x-side pre-activations: issued before the stores below
                // ---- y and the saved gates: off the critical path (the next step's MMA is already being fed), straight from
                // registers -- with the chunk-fastest item mapping the five lanes of a batch row write 160 contiguous bytes per
                // array, which is what the shared-memory staging of the earlier versions was there to achieve
                if (live && !(xflags & 6)) {
                    const int n4 = (H - jbase) >= 8 ? 2 : ((H - jbase) >= 4 ? 1 : 0);   // valid float4 halves (H % 4 == 0)
                    const size_t row = (size_t)b * T + t;
                    if (n4 > 0) {
                        float* yd = g_y + row * 2 * H + dir * H + jbase;
                        reinterpret_cast<float4*>(yd)[0] = make_float4(hnew[0], hnew[1], hnew[2], hnew[3]);
                        if (n4 > 1) reinterpret_cast<float4*>(yd)[1] = make_float4(hnew[4], hnew[5], hnew[6], hnew[7]);
                        if (g_gates != nullptr && b < M_gates) {
                            float* gd = g_gates + (row * 2 + dir) * 4 * H + jbase;
                            reinterpret_cast<float4*>(gd)[0] = make_float4(sr[0], sr[1], sr[2], sr[3]);
                            reinterpret_cast<float4*>(gd + H)[0] = make_float4(sz[0], sz[1], sz[2], sz[3]);
                            reinterpret_cast<float4*>(gd + 2 * H)[0] = make_float4(sn[0], sn[1], sn[2], sn[3]);
                            reinterpret_cast<float4*>(gd + 3 * H)[0] = make_float4(shn[0], shn[1], shn[2], shn[3]);
                            if (n4 > 1) {
                                reinterpret_cast<float4*>(gd)[1] = make_float4(sr[4], sr[5], sr[6], sr[7]);
                                reinterpret_cast<float4*>(gd + H)[1] = make_float4(sz[4], sz[5], sz[6], sz[7]);
                                reinterpret_cast<float4*>(gd + 2 * H)[1] = make_float4(sn[4], sn[5], sn[6], sn[7]);
                                reinterpret_cast<float4*>(gd + 3 * H)[1] = make_float4(shn[4], shn[5], shn[6], shn[7]);
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


// How many 8-CTA clusters of the recurrence kernel the device keeps resident at once (a cluster must sit inside one GPC,
// so this is well below SMs / 8 rounded down on a floor-swept part); cached per process.  Every configuration asks for
// more than half an SM's shared memory, i.e. one CTA per SM, so one query serves them all.
static int max_resident_clusters() {
    static int cached = 0;
    if (cached > 0) return cached;
    const Tc2Layout L(40, 300, 48, 1);
    if (cudaFuncSetAttribute(gru_seq_fwd_tc2_kernel<48, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total) != cudaSuccess) return 16;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16 * CL); cfg.blockDim = dim3(block_threads(48, 1)); cfg.dynamicSmemBytes = L.total;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gru_seq_fwd_tc2_kernel<48, 1>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 16; }
    cached = n;
    return n;
}

static int g_xflags = 0;
// timing experiments only: see Tc2Params::xflags (results are wrong with any flag set)
HA2G_API int ha2g_gru_fwd_xflags(int flags) { g_xflags = flags; return 0; }

HA2G_API int ha2g_gru_max_clusters(int* n) { *n = max_resident_clusters(); return 0; }

// Rows per cluster task for a batch of M rows, as the code the _dbg launcher takes: the smallest task whose row chunks,
// for both directions, are all resident at once.  A B200 keeps 15 of these 8-CTA clusters resident (measured,
// tools/gru_waves.py): a 16th cluster waits for a whole sequence, i.e. doubles the kernel time -- so 7 chunks per
// direction is the limit of one wave.  Tasks of 20 rows and more run as two interleaved halves (200 + rows per half):
// measured faster at every size (128 rows: 87 vs 106 us; 224: 88 vs 115; 384: 152 vs 199).
static int pick_nb(int M) {
    const int per_dir = max_resident_clusters() / 2;
    for (int nb : {16, 20, 32, 48, 56, 64})
        if ((M + nb - 1) / nb <= per_dir) return nb >= 20 ? 200 + nb / 2 : nb;
    return 232;
}

// 1 through *ok if the second-generation tensor-core recurrence can serve hidden size H (gate rows 3*HSP <= 128, the
// W_hh slice fits the TMEM columns behind the accumulator, one 8-unit item per epilogue thread in every configuration).
HA2G_API int ha2g_gru_tc2_supported(int H, int* ok) {
    const int HSP = ((H + CL - 1) / CL + 7) / 8 * 8;
    const int kc = CL * HSP / 8;
    bool good = 3 * HSP <= TM && kc % 2 == 0 && A_COL + kc * 8 <= TMEM_COLS && H % 4 == 0;
    const int cfgs[11][2] = {{16, 1}, {20, 1}, {32, 1}, {48, 1}, {56, 1}, {64, 1}, {10, 2}, {16, 2}, {24, 2}, {28, 2}, {32, 2}};
    for (int i = 0; i < 11 && good; ++i) {
        const int nb = cfgs[i][0], nh = cfgs[i][1];
        const Tc2Layout L(HSP, H, nb, nh);
        const int threads_half = 32 * epi_warps(nb, nh) / nh;
        good = (HSP / 8) * nb <= threads_half && nh * mma_n(nb) <= A_COL && L.total <= 227 * 1024;
    }
    *ok = good ? 1 : 0;
    return 0;
}

template <int NB, int NH>
static int launch_tc2(Tc2Params& p, cudaStream_t stream) {
    p.n_chunks = (p.M + NH * NB - 1) / (NH * NB);
    p.gpt = 0;
    if (p.M_gates > 0 && p.M_gates < p.M) {   // spread the gate-saving rows evenly over the tasks when the split works out
        const int R = NH * NB, a = (p.M_gates + p.n_chunks - 1) / p.n_chunks;
        if (a < R && (long long)p.n_chunks * (R - a) >= p.M - p.M_gates) p.gpt = a;
    }
    const Tc2Layout L(p.HSP, p.H, NB, NH);
    cudaError_t e = cudaFuncSetAttribute(gru_seq_fwd_tc2_kernel<NB, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    if (e != cudaSuccess) return (int)e;
    int clusters = 2 * p.n_chunks;
    int cap = max_resident_clusters() & ~1;   // the same number of clusters per direction
    if (cap < 2) cap = 2;
    if (clusters > cap) clusters = cap;
    gru_seq_fwd_tc2_kernel<NB, NH><<<clusters * CL, block_threads(NB, NH), L.total, stream>>>(p);
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

// Same, with an explicit rows-per-cluster choice (nb = 16 / 20 / 32 / 48 / 56 / 64 as one task, 210 / 216 / 224 / 228 / 232 = two
// interleaved halves of 10 / 16 / 24 / 28 / 32 rows; 0 = automatic) and an optional device buffer dbg
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
    p.xflags = g_xflags;
    p.gi = gi; p.w_hh[0] = w_hh_f; p.w_hh[1] = w_hh_r; p.b_hh[0] = b_hh_f; p.b_hh[1] = b_hh_r;
    p.y = y; p.gates = gates; p.M = M; p.T = T; p.H = H;
    p.M_gates = gates != nullptr ? (M_gates < M ? M_gates : M) : 0;
    p.HSP = ((H + CL - 1) / CL + 7) / 8 * 8;
    if (nb == 0) nb = pick_nb(M);
    if (nb == 16) return launch_tc2<16, 1>(p, stream);
    if (nb == 20) return launch_tc2<20, 1>(p, stream);
    if (nb == 32) return launch_tc2<32, 1>(p, stream);
    if (nb == 48) return launch_tc2<48, 1>(p, stream);
    if (nb == 56) return launch_tc2<56, 1>(p, stream);
    if (nb == 64) return launch_tc2<64, 1>(p, stream);
    if (nb == 210) return launch_tc2<10, 2>(p, stream);
    if (nb == 216) return launch_tc2<16, 2>(p, stream);
    if (nb == 224) return launch_tc2<24, 2>(p, stream);
    if (nb == 228) return launch_tc2<28, 2>(p, stream);
    if (nb == 232) return launch_tc2<32, 2>(p, stream);
    return (int)cudaErrorInvalidValue;
}
