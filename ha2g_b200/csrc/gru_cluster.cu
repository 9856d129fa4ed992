// Persistent thread-block-cluster GRU recurrence (K10/K11 of SURVEY.md), exact fp32.
//
// One launch runs ALL T time steps of one bidirectional layer.  A cluster of 8 CTAs serves one
// (direction, 16-row batch chunk) task at a time; CTA `rank` owns the hidden units [rank*HS, rank*HS+HS) and keeps
// its W_hh slice (rows r|z|n of those units, all K = H columns) resident in shared memory for the whole sequence:
// W_hh is read from HBM/L2 ONCE per layer instead of once per time step.  Each step
//   forward : every CTA computes the three gate pre-activations of its units from the full h_{t-1} held in its own
//             shared memory, applies the gate math, writes h_t to y, and pushes its h_t slice into the shared memory
//             of all 8 CTAs with distributed-shared-memory (DSMEM) 128-bit stores; one cluster barrier per step.
//   backward: every CTA turns dh_t of its units into gate gradients (dgi/dgh to global for the weight-gradient
//             GEMMs), multiplies its gate-gradient slice with the same resident W_hh rows to get a partial
//             dh_{t-1} for ALL units, and reduce-scatters the partials to the owning CTAs through DSMEM.
// Replaces the per-step launches of gru.cu (78 us/step at B=128, latency bound: profiles/r01_a_*) for the
// nn.GRU call sites scripts/model/hierarchy_net.py:144 (H=300) and :232 (H=64).
#include "common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

constexpr int CL = 8;       // CTAs per cluster (portable maximum)
constexpr int NB = 16;      // batch rows per cluster task
constexpr int NQ = NB / 4;  // float4 groups of batch rows
constexpr int GC_NT = 160;  // threads per CTA (>= NQ * HS for H = 300: 4 * 38 = 152)

struct GruClusterParams {
    const float* gi;       // [M,T,2,3H]
    const float* w_hh[2];  // [3H,H] per direction
    const float* b_hh[2];  // [3H]
    float* y;              // [M,T,2H]
    float* gates;          // [M,T,2,4H] or nullptr
    // backward only
    const float* dy;
    int dy_ld, dy_dir_stride;
    float* dgi;
    float* dgh;
    int M, T, H, HS, n_chunks;
};

__device__ __forceinline__ void load_w_slice(float4* Wt, const float* __restrict__ W, int H, int HS, int j0, int nu) {
    // Wt[k][u] = { W[r-row j0+u][k], W[z-row][k], W[n-row][k], 0 };  e enumerates (u,k) with k fastest (coalesced reads)
    for (int e = threadIdx.x; e < H * HS; e += blockDim.x) {
        int k = e % H, u = e / H;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (u < nu) {
            int j = j0 + u;
            v.x = W[(size_t)j * H + k];
            v.y = W[(size_t)(H + j) * H + k];
            v.z = W[(size_t)(2 * H + j) * H + k];
        }
        Wt[(size_t)k * HS + u] = v;
    }
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(GC_NT, 1) gru_seq_fwd_cluster_kernel(GruClusterParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
    const int dir = cluster_id & 1;
    const int H = p.H, HS = p.HS, T = p.T, M = p.M;
    const int j0 = rank * HS;
    const int nu = max(0, min(HS, H - j0));
    float4* Wt = reinterpret_cast<float4*>(smem_raw);               // [H][HS]
    float* hbuf = reinterpret_cast<float*>(Wt + (size_t)H * HS);    // [2][H][NB]
    const int tid = threadIdx.x;
    const int u = tid % HS, q = tid / HS;
    const bool active = q < NQ;
    const bool owner = active && u < nu;
    const int j = j0 + u;
    const float* __restrict__ b_hh = p.b_hh[dir];

    load_w_slice(Wt, p.w_hh[dir], H, HS, j0, nu);
    float bh_r = 0.f, bh_z = 0.f, bh_n = 0.f;
    if (owner) { bh_r = b_hh[j]; bh_z = b_hh[H + j]; bh_n = b_hh[2 * H + j]; }

    for (int task = cluster_id >> 1; task < p.n_chunks; task += n_clusters >> 1) {
        const int m0 = task * NB;
        for (int e = tid; e < H * NB; e += blockDim.x) hbuf[e] = 0.f;  // h_{-1} = 0 (buffer 0)
        cluster.sync();  // W resident, buffers initialised, previous task's remote traffic drained
        int cur = 0;
        for (int s = 0; s < T; ++s) {
            const int t = dir == 0 ? s : T - 1 - s;
            const float* hb = hbuf + (size_t)cur * H * NB;
            float* hn_local = hbuf + (size_t)(cur ^ 1) * H * NB;
            // x-side pre-activations of this step (independent of the recurrence: issued before the k loop)
            float gr[4], gz[4], gn[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int b = m0 + q * 4 + i;
                gr[i] = gz[i] = gn[i] = 0.f;
                if (owner && b < M) {
                    const float* g = p.gi + (((size_t)b * T + t) * 2 + dir) * 3 * H;
                    gr[i] = g[j]; gz[i] = g[H + j]; gn[i] = g[2 * H + j];
                }
            }
            float ar[4] = {0.f, 0.f, 0.f, 0.f}, az[4] = {0.f, 0.f, 0.f, 0.f}, an[4] = {0.f, 0.f, 0.f, 0.f};
            if (active && s > 0) {
                const float4* wp = Wt + u;
                const float* hp = hb + q * 4;
#pragma unroll 4
                for (int k = 0; k < H; ++k) {
                    const float4 w = wp[(size_t)k * HS];
                    const float4 h4 = *reinterpret_cast<const float4*>(hp + (size_t)k * NB);
                    ar[0] = fmaf(w.x, h4.x, ar[0]); ar[1] = fmaf(w.x, h4.y, ar[1]); ar[2] = fmaf(w.x, h4.z, ar[2]); ar[3] = fmaf(w.x, h4.w, ar[3]);
                    az[0] = fmaf(w.y, h4.x, az[0]); az[1] = fmaf(w.y, h4.y, az[1]); az[2] = fmaf(w.y, h4.z, az[2]); az[3] = fmaf(w.y, h4.w, az[3]);
                    an[0] = fmaf(w.z, h4.x, an[0]); an[1] = fmaf(w.z, h4.y, an[1]); an[2] = fmaf(w.z, h4.z, an[2]); an[3] = fmaf(w.z, h4.w, an[3]);
                }
            }
            if (owner) {
                float hnew[4];
                const float4 hprev4 = *reinterpret_cast<const float4*>(hb + (size_t)j * NB + q * 4);
                const float hprev[4] = {hprev4.x, hprev4.y, hprev4.z, hprev4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int b = m0 + q * 4 + i;
                    hnew[i] = 0.f;
                    if (b < M) {
                        const float hr = ar[i] + bh_r, hz = az[i] + bh_z, hn = an[i] + bh_n;
                        const float r = ha2g_sigmoid(gr[i] + hr);
                        const float z = ha2g_sigmoid(gz[i] + hz);
                        const float n = tanhf(gn[i] + r * hn);
                        hnew[i] = (1.f - z) * n + z * hprev[i];
                        const size_t row = (size_t)b * T + t;
                        p.y[row * 2 * H + dir * H + j] = hnew[i];
                        if (p.gates != nullptr) {
                            float* gs = p.gates + (row * 2 + dir) * 4 * H;
                            gs[j] = r; gs[H + j] = z; gs[2 * H + j] = n; gs[3 * H + j] = hn;
                        }
                    }
                }
                const float4 v = make_float4(hnew[0], hnew[1], hnew[2], hnew[3]);
                const size_t off = (size_t)(cur ^ 1) * H * NB + (size_t)j * NB + q * 4;
#pragma unroll
                for (int dst = 0; dst < CL; ++dst) {
                    float* remote = cluster.map_shared_rank(hbuf, dst);
                    *reinterpret_cast<float4*>(remote + off) = v;
                }
            }
            (void)hn_local;
            cluster.sync();  // release/acquire: every CTA now holds the complete h_t in buffer cur^1
            cur ^= 1;
        }
    }
    cluster.sync();  // no CTA leaves while a peer could still address its shared memory
}

// Backward.  Shared memory: Wt [H][HS] float4 | recv [CL][HS][NB] | dg [HS][3][NB]
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(GC_NT, 1) gru_seq_bwd_cluster_kernel(GruClusterParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
    const int dir = cluster_id & 1;
    const int H = p.H, HS = p.HS, T = p.T, M = p.M;
    const int j0 = rank * HS;
    const int nu = max(0, min(HS, H - j0));
    float4* Wt = reinterpret_cast<float4*>(smem_raw);                       // [H][HS]
    float* recv = reinterpret_cast<float*>(Wt + (size_t)H * HS);            // [CL][HS][NB]
    float* dg = recv + (size_t)CL * HS * NB;                                // [HS][3][NB]
    const int tid = threadIdx.x;
    const int u = tid % HS, q = tid / HS;
    const bool active = q < NQ;
    const bool owner = active && u < nu;
    const int j = j0 + u;

    load_w_slice(Wt, p.w_hh[dir], H, HS, j0, nu);

    for (int task = cluster_id >> 1; task < p.n_chunks; task += n_clusters >> 1) {
        const int m0 = task * NB;
        for (int e = tid; e < CL * HS * NB; e += blockDim.x) recv[e] = 0.f;
        for (int e = tid; e < HS * 3 * NB; e += blockDim.x) dg[e] = 0.f;
        float dhz[4] = {0.f, 0.f, 0.f, 0.f};  // direct term dh_total * z carried to the previous time step
        cluster.sync();
        for (int s = T - 1; s >= 0; --s) {
            const int t = dir == 0 ? s : T - 1 - s;
            const int tp = dir == 0 ? t - 1 : t + 1;
            // ---- phase A: gate gradients of the owned units ----------------------------------------------
            if (owner) {
                float4 acc = make_float4(dhz[0], dhz[1], dhz[2], dhz[3]);
#pragma unroll
                for (int src = 0; src < CL; ++src) {
                    const float4 v = *reinterpret_cast<const float4*>(recv + ((size_t)src * HS + u) * NB + q * 4);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
                const float dh_rec[4] = {acc.x, acc.y, acc.z, acc.w};
                float o_r[4], o_z[4], o_n[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int b = m0 + q * 4 + i;
                    o_r[i] = o_z[i] = o_n[i] = 0.f;
                    dhz[i] = 0.f;
                    if (b < M) {
                        const size_t row = (size_t)b * T + t;
                        const float* gs = p.gates + (row * 2 + dir) * 4 * H;
                        const float r = gs[j], z = gs[H + j], n = gs[2 * H + j], hn = gs[3 * H + j];
                        const float hp = s > 0 ? p.y[((size_t)b * T + tp) * 2 * H + dir * H + j] : 0.f;
                        const float dh = p.dy[row * p.dy_ld + dir * p.dy_dir_stride + j] + dh_rec[i];
                        const float dn = dh * (1.f - z);
                        const float dz = dh * (hp - n);
                        const float dn_pre = dn * (1.f - n * n);
                        const float dz_pre = dz * z * (1.f - z);
                        const float dr_pre = dn_pre * hn * r * (1.f - r);
                        float* gi_o = p.dgi + (row * 2 + dir) * 3 * H;
                        float* gh_o = p.dgh + (row * 2 + dir) * 3 * H;
                        gi_o[j] = dr_pre; gi_o[H + j] = dz_pre; gi_o[2 * H + j] = dn_pre;
                        gh_o[j] = dr_pre; gh_o[H + j] = dz_pre; gh_o[2 * H + j] = dn_pre * r;
                        o_r[i] = dr_pre; o_z[i] = dz_pre; o_n[i] = dn_pre * r;
                        dhz[i] = dh * z;
                    }
                }
                float* d0 = dg + ((size_t)u * 3) * NB + q * 4;
                *reinterpret_cast<float4*>(d0) = make_float4(o_r[0], o_r[1], o_r[2], o_r[3]);
                *reinterpret_cast<float4*>(d0 + NB) = make_float4(o_z[0], o_z[1], o_z[2], o_z[3]);
                *reinterpret_cast<float4*>(d0 + 2 * NB) = make_float4(o_n[0], o_n[1], o_n[2], o_n[3]);
            }
            cluster.sync();  // (A) every CTA consumed recv; dg visible CTA-wide
            // ---- phase C: partial dh_{prev}[b][k] over the owned gate rows, for ALL k; scatter to owners ------
            if (s > 0) {
                for (int k = tid; k < H; k += blockDim.x) {
                    float acc[NB];
#pragma unroll
                    for (int i = 0; i < NB; ++i) acc[i] = 0.f;
                    const float4* wp = Wt + (size_t)k * HS;
                    for (int uu = 0; uu < nu; ++uu) {
                        const float4 w = wp[uu];
                        const float* d = dg + (size_t)uu * 3 * NB;
#pragma unroll
                        for (int i4 = 0; i4 < NQ; ++i4) {
                            const float4 a = *reinterpret_cast<const float4*>(d + i4 * 4);
                            const float4 bq = *reinterpret_cast<const float4*>(d + NB + i4 * 4);
                            const float4 c = *reinterpret_cast<const float4*>(d + 2 * NB + i4 * 4);
                            acc[i4 * 4 + 0] = fmaf(w.x, a.x, fmaf(w.y, bq.x, fmaf(w.z, c.x, acc[i4 * 4 + 0])));
                            acc[i4 * 4 + 1] = fmaf(w.x, a.y, fmaf(w.y, bq.y, fmaf(w.z, c.y, acc[i4 * 4 + 1])));
                            acc[i4 * 4 + 2] = fmaf(w.x, a.z, fmaf(w.y, bq.z, fmaf(w.z, c.z, acc[i4 * 4 + 2])));
                            acc[i4 * 4 + 3] = fmaf(w.x, a.w, fmaf(w.y, bq.w, fmaf(w.z, c.w, acc[i4 * 4 + 3])));
                        }
                    }
                    const int dst = k / HS, ud = k % HS;
                    float* remote = cluster.map_shared_rank(recv, dst) + ((size_t)rank * HS + ud) * NB;
#pragma unroll
                    for (int i4 = 0; i4 < NQ; ++i4)
                        *reinterpret_cast<float4*>(remote + i4 * 4) =
                            make_float4(acc[i4 * 4], acc[i4 * 4 + 1], acc[i4 * 4 + 2], acc[i4 * 4 + 3]);
                }
            }
            cluster.sync();  // (B) all partial slices delivered
        }
    }
    cluster.sync();
}

static size_t fwd_smem(int H, int HS) { return (size_t)H * HS * 16 + (size_t)2 * H * NB * 4; }
static size_t bwd_smem(int H, int HS) { return (size_t)H * HS * 16 + (size_t)CL * HS * NB * 4 + (size_t)HS * 3 * NB * 4; }

static int cluster_grid(int n_chunks) {
    int clusters = 2 * n_chunks;       // one per (direction, chunk) when they fit
    if (clusters > 16) clusters = 16;  // 148 SMs hold at most 16-18 resident 8-CTA clusters; the rest would queue
    return clusters * CL;
}

}  // namespace

// 1 if the persistent cluster kernels can serve hidden size H (shared-memory fit), else 0 -> caller uses gru.cu's
// per-step path.  (Returned through *ok; the int return value is the usual status.)
HA2G_API int ha2g_gru_cluster_supported(int H, int* ok) {
    const int HS = (H + CL - 1) / CL;
    *ok = (NQ * HS <= GC_NT && fwd_smem(H, HS) <= 227 * 1024 && bwd_smem(H, HS) <= 227 * 1024) ? 1 : 0;
    return 0;
}

// All T steps of one bidirectional layer, forward.  gi [M,T,2,3H] must already hold x W_ih^T + b_ih.
HA2G_API int ha2g_gru_seq_fwd_cluster(const float* gi, const float* w_hh_f, const float* w_hh_r, const float* b_hh_f,
                                      const float* b_hh_r, float* y, float* gates, int M, int T, int H,
                                      cudaStream_t stream) {
    GruClusterParams p{};
    p.gi = gi; p.w_hh[0] = w_hh_f; p.w_hh[1] = w_hh_r; p.b_hh[0] = b_hh_f; p.b_hh[1] = b_hh_r;
    p.y = y; p.gates = gates; p.M = M; p.T = T; p.H = H;
    p.HS = (H + CL - 1) / CL;
    p.n_chunks = (M + NB - 1) / NB;
    const size_t smem = fwd_smem(H, p.HS);
    cudaError_t e = cudaFuncSetAttribute(gru_seq_fwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    gru_seq_fwd_cluster_kernel<<<cluster_grid(p.n_chunks), GC_NT, smem, stream>>>(p);
    HA2G_RETURN_LAST();
}

// All T steps of one bidirectional layer, backward: writes dgi, dgh [M,T,2,3H] (see gru.cu for their meaning).
HA2G_API int ha2g_gru_seq_bwd_cluster(const float* dy, int dy_ld, int dy_dir_stride, const float* y, const float* gates,
                                      const float* w_hh_f, const float* w_hh_r, float* dgi, float* dgh, int M, int T, int H,
                                      cudaStream_t stream) {
    GruClusterParams p{};
    p.dy = dy; p.dy_ld = dy_ld; p.dy_dir_stride = dy_dir_stride;
    p.y = const_cast<float*>(y); p.gates = const_cast<float*>(gates);
    p.w_hh[0] = w_hh_f; p.w_hh[1] = w_hh_r; p.dgi = dgi; p.dgh = dgh; p.M = M; p.T = T; p.H = H;
    p.HS = (H + CL - 1) / CL;
    p.n_chunks = (M + NB - 1) / NB;
    const size_t smem = bwd_smem(H, p.HS);
    cudaError_t e = cudaFuncSetAttribute(gru_seq_bwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    gru_seq_bwd_cluster_kernel<<<cluster_grid(p.n_chunks), GC_NT, smem, stream>>>(p);
    HA2G_RETURN_LAST();
}
