// Bidirectional GRU layer, exact-fp32 path (K9/K10/K11 of SURVEY.md).
//
// Semantics = torch.nn.GRU(batch_first=True, bidirectional=True) for ONE layer, h0 = 0
// (reference call sites: scripts/model/hierarchy_net.py:87-88,144 generator H=300;
// :213-214,232 discriminator H=64).  Gate order r,z,n;  n = tanh(gi_n + r * (W_hn h + b_hn)).
//
// Layouts (row-major, fp32):
//   x      [M, T, I]            layer input (batch first)
//   gi     [M, T, 2, 3H]        x W_ih^T + b_ih for both directions (dir 0 = forward, 1 = reverse)
//   y      [M, T, 2H]           hidden states, forward half then reverse half (what nn.GRU returns)
//   gates  [M, T, 2, 4H]        r | z | n | hn(=W_hn h + b_hn) saved for BPTT (nullptr: inference)
//   dgi    [M, T, 2, 3H]        gradient wrt gi          (input side:  dr_pre | dz_pre | dn_pre)
//   dgh    [M, T, 2, 3H]        gradient wrt W_hh h+b_hh (hidden side: dr_pre | dz_pre | dn_pre * r)
//
// One launch per time step covers both directions (grid.z); h_{t-1} is read straight out of y.
#include "common.cuh"

extern "C" int ha2g_gemm(const float*, const float*, float*, const float*, int, int, int, int, int, int, int, int,
                             int, int, int, cudaStream_t);
extern "C" int ha2g_col_sum(const float* x, int rows, int cols, int ld, float* out, cudaStream_t stream);
extern "C" int ha2g_gemm_kseg(const float*, const float*, float*, const float*, int, int, int, int, int, int, int,
                                  int, int, int, int, int, int, cudaStream_t);

#include <cstdlib>
#include <cstring>

// The recurrences run on the tcgen05 cluster kernels (gru_cluster_tc2.cu forward, gru_cluster_tc2_bwd.cu backward) for
// every hidden size they can serve (H = 300 generators, H = 64 discriminator); the per-step fp32 kernels of this file
// are the exact fallback for any other H.
extern "C" int ha2g_gru_tc2_supported(int H, int* ok);
extern "C" int ha2g_gru_seq_fwd_tc2(const float*, const float*, const float*, const float*, const float*, float*, float*, int,
                                    int, int, int, cudaStream_t);
extern "C" int ha2g_gru_tc2_bwd_supported(int H, int* ok);
extern "C" int ha2g_gru_seq_bwd_tc2(const float*, int, int, const float*, const float*, const float*, const float*, float*,
                                    float*, int, int, int, cudaStream_t);
static bool use_tc2_bwd(int H) {
    int ok = 0;
    ha2g_gru_tc2_bwd_supported(H, &ok);
    return ok != 0;
}
static bool use_tc2_recurrence(int H) {
    int ok = 0;
    ha2g_gru_tc2_supported(H, &ok);
    return ok != 0;
}

namespace {

constexpr int GR_ROWS = 32;   // batch rows per CTA
constexpr int GR_HID = 16;    // hidden units per CTA (x3 gates = 48 columns)
constexpr int GR_BK = 20;     // K chunk (300 = 15 * 20)
constexpr int GR_NT = 96;     // 8 row-groups(4 rows) x 12 col-groups(4 cols)

// step s of the recurrence: forward direction handles t = s, reverse handles t = T-1-s.
__global__ void __launch_bounds__(GR_NT) gru_step_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ w_hh_f,
                                                             const float* __restrict__ w_hh_r,
                                                             const float* __restrict__ b_hh_f,
                                                             const float* __restrict__ b_hh_r, float* __restrict__ y,
                                                             float* __restrict__ gates, int M, int M_gates, int T, int H,
                                                             int s) {
    __shared__ __align__(16) float Hs[GR_BK][GR_ROWS + 4];
    __shared__ __align__(16) float Ws[GR_BK][3 * GR_HID + 4];
    __shared__ float Gs[GR_ROWS][3 * GR_HID + 1];

    const int dir = blockIdx.z;
    const int t = dir == 0 ? s : T - 1 - s;
    const int tp = dir == 0 ? t - 1 : t + 1;  // time index of h_{prev}
    const float* __restrict__ w_hh = dir == 0 ? w_hh_f : w_hh_r;
    const float* __restrict__ b_hh = dir == 0 ? b_hh_f : b_hh_r;
    const int j0 = blockIdx.x * GR_HID;
    const int m0 = blockIdx.y * GR_ROWS;
    const int tid = threadIdx.x;
    const int cg = tid % 12, rg = tid / 12;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    if (s > 0) {
        const size_t ystride = (size_t)T * 2 * H;
        for (int k0 = 0; k0 < H; k0 += GR_BK) {
            // h_prev tile: GR_ROWS x GR_BK
            for (int e = tid; e < GR_ROWS * GR_BK; e += GR_NT) {
                int k = e % GR_BK, r = e / GR_BK;
                int gm = m0 + r, gk = k0 + k;
                Hs[k][r] = (gm < M && gk < H) ? y[(size_t)gm * ystride + (size_t)tp * 2 * H + dir * H + gk] : 0.f;
            }
            // W_hh tile: 48 gate columns x GR_BK   (column c -> gate c/16, hidden j0 + c%16)
            for (int e = tid; e < 3 * GR_HID * GR_BK; e += GR_NT) {
                int k = e % GR_BK, c = e / GR_BK;
                int g = c / GR_HID, j = j0 + c % GR_HID, gk = k0 + k;
                Ws[k][c] = (j < H && gk < H) ? w_hh[(size_t)(g * H + j) * H + gk] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < GR_BK; ++kk) {
                const float4 a = *reinterpret_cast<const float4*>(&Hs[kk][rg * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Ws[kk][cg * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w};
                const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) Gs[rg * 4 + i][cg * 4 + j] = acc[i][j];
    __syncthreads();

    // gate epilogue: one (row, hidden) pair per iteration
    for (int e = tid; e < GR_ROWS * GR_HID; e += GR_NT) {
        int jj = e % GR_HID, r = e / GR_HID;
        int gm = m0 + r, j = j0 + jj;
        if (gm >= M || j >= H) continue;
        const size_t row = (size_t)gm * T + t;
        const float* g = gi + (row * 2 + dir) * 3 * H;
        float hr = Gs[r][jj] + b_hh[j];
        float hz = Gs[r][GR_HID + jj] + b_hh[H + j];
        float hn = Gs[r][2 * GR_HID + jj] + b_hh[2 * H + j];
        float rr = ha2g_sigmoid(g[j] + hr);
        float zz = ha2g_sigmoid(g[H + j] + hz);
        float nn = tanhf(g[2 * H + j] + rr * hn);
        float hp = s > 0 ? y[((size_t)gm * T + tp) * 2 * H + dir * H + j] : 0.f;
        float hnew = (1.f - zz) * nn + zz * hp;
        y[row * 2 * H + dir * H + j] = hnew;
        if (gates != nullptr && gm < M_gates) {
            float* gs = gates + (row * 2 + dir) * 4 * H;
            gs[j] = rr; gs[H + j] = zz; gs[2 * H + j] = nn; gs[3 * H + j] = hn;
        }
    }
}

// BPTT element-wise part for step s (processed in decreasing s).
//   dh_total = dy[m,t,dir] + dh_rec[m,dir]  ->  dgi, dgh at (m,t,dir);  dh_rec <- dh_total * z
__global__ void gru_gates_bwd_kernel(const float* __restrict__ dy, int dy_ld, int dy_dir_stride,
                                     const float* __restrict__ y, const float* __restrict__ gates,
                                     float* __restrict__ dgi, float* __restrict__ dgh, float* __restrict__ dh_rec, int M,
                                     int T, int H, int s) {
    const int64_t n = (int64_t)M * 2 * H;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int j = (int)(e % H);
        int dir = (int)((e / H) % 2);
        int m = (int)(e / (2 * H));
        int t = dir == 0 ? s : T - 1 - s;
        int tp = dir == 0 ? t - 1 : t + 1;
        const size_t row = (size_t)m * T + t;
        const float* gs = gates + (row * 2 + dir) * 4 * H;
        float r = gs[j], z = gs[H + j], nn = gs[2 * H + j], hn = gs[3 * H + j];
        float hp = s > 0 ? y[((size_t)m * T + tp) * 2 * H + dir * H + j] : 0.f;
        float dh = dy[row * dy_ld + dir * dy_dir_stride + j] + dh_rec[e];
        float dn = dh * (1.f - z);
        float dz = dh * (hp - nn);
        float dn_pre = dn * (1.f - nn * nn);
        float dz_pre = dz * z * (1.f - z);
        float dr_pre = dn_pre * hn * r * (1.f - r);
        float* gi_o = dgi + (row * 2 + dir) * 3 * H;
        float* gh_o = dgh + (row * 2 + dir) * 3 * H;
        gi_o[j] = dr_pre; gi_o[H + j] = dz_pre; gi_o[2 * H + j] = dn_pre;
        gh_o[j] = dr_pre; gh_o[H + j] = dz_pre; gh_o[2 * H + j] = dn_pre * r;
        dh_rec[e] = dh * z;
    }
}

}  // namespace

#define HA2G_CHECK(call) do { int e_ = (call); if (e_ != 0) return e_; } while (0)

// Forward of ONE bidirectional nn.GRU layer (replaces the cuDNN RNN call behind `self.gru(in_data, None)`,
// scripts/model/hierarchy_net.py:144 generator / :232 discriminator): gi = x W_ih^T + b_ih (two GEMMs),
// then T fused "h W_hh^T + gates" step launches covering both directions.
//   x [M,T,I];  w_ih_* [3H,I], w_hh_* [3H,H], b_* [3H] (gate order r,z,n; *_f forward, *_r reverse direction)
//   gi [M,T,2,3H] scratch/out;  y [M,T,2H] out;  gates [M_gates,T,2,4H] out: saved for the FIRST M_gates batch rows only
//   (the rows whose backward pass will run; nullptr = inference, nothing saved)
HA2G_API int ha2g_gru_layer_fwd(const float* x, int I, const float* w_ih_f, const float* w_ih_r, const float* b_ih_f,
                                const float* b_ih_r, const float* w_hh_f, const float* w_hh_r, const float* b_hh_f,
                                const float* b_hh_r, float* gi, float* y, float* gates, int M, int M_gates, int T, int H,
                                cudaStream_t stream) {
    const int MT = M * T;
    if (w_ih_r == w_ih_f + (size_t)3 * H * I && b_ih_r == b_ih_f + 3 * H) {
        // both directions' input weights are adjacent (one [6H, I] matrix): ONE projection GEMM with N = 6H
        HA2G_CHECK(ha2g_gemm(x, w_ih_f, gi, b_ih_f, MT, 6 * H, I, I, I, 6 * H, 0, 1, 0, 0, 1, stream));
    } else {
        HA2G_CHECK(ha2g_gemm(x, w_ih_f, gi, b_ih_f, MT, 3 * H, I, I, I, 6 * H, 0, 1, 0, 0, 1, stream));
        HA2G_CHECK(ha2g_gemm(x, w_ih_r, gi + 3 * H, b_ih_r, MT, 3 * H, I, I, I, 6 * H, 0, 1, 0, 0, 1, stream));
    }
    if (use_tc2_recurrence(H)) return ha2g_gru_seq_fwd_tc2(gi, w_hh_f, w_hh_r, b_hh_f, b_hh_r, y, gates, M, M_gates, T, H, stream);
    dim3 grid(ha2g_div_up(H, GR_HID), ha2g_div_up(M, GR_ROWS), 2);
    for (int s = 0; s < T; ++s) {
        gru_step_fwd_kernel<<<grid, GR_NT, 0, stream>>>(gi, w_hh_f, w_hh_r, b_hh_f, b_hh_r, y, gates, M, M_gates, T, H, s);
    }
    HA2G_RETURN_LAST();
}

// Backward of one bidirectional layer.
//   dy: gradient wrt y; element (m,t,dir,j) at dy[(m*T+t)*dy_ld + dir*dy_dir_stride + j]
//       (dy_ld = 2H, dy_dir_stride = H for a [M,T,2H] gradient; dy_ld = H, dy_dir_stride = 0 when the
//        consumer summed the two directions, hierarchy_net.py:145).
//   Outputs: dx [M,T,I] (overwritten), and ACCUMULATED (+=) into dw_ih_*, dw_hh_*, db_ih_*, db_hh_*.
//   Scratch: dgi, dgh [M,T,2,3H]; dh_rec [M,2,H].
HA2G_API int ha2g_gru_layer_bwd(const float* dy, int dy_ld, int dy_dir_stride, const float* x, int I, const float* y,
                       const float* gates, const float* w_ih_f, const float* w_ih_r, const float* w_hh_f,
                       const float* w_hh_r, float* dgi, float* dgh, float* dh_rec, float* dx, float* dw_ih_f,
                       float* dw_ih_r, float* dw_hh_f, float* dw_hh_r, float* db_ih_f, float* db_ih_r, float* db_hh_f,
                       float* db_hh_r, int M, int T, int H, cudaStream_t stream) {
    const int MT = M * T;
    cudaError_t ce = cudaMemsetAsync(dh_rec, 0, sizeof(float) * (size_t)M * 2 * H, stream);
    if (ce != cudaSuccess) return (int)ce;
    const int ew_grid = ha2g_ew_grid((int64_t)M * 2 * H, 256, 1);
    const bool clustered = use_tc2_bwd(H);
    if (clustered)
        HA2G_CHECK(ha2g_gru_seq_bwd_tc2(dy, dy_ld, dy_dir_stride, y, gates, w_hh_f, w_hh_r, dgi, dgh, M, T, H, stream));
    for (int s = T - 1; s >= 0 && !clustered; --s) {
        gru_gates_bwd_kernel<<<ew_grid, 256, 0, stream>>>(dy, dy_ld, dy_dir_stride, y, gates, dgi, dgh, dh_rec, M, T, H, s);
        if (s > 0) {
            // dh_rec[m,dir,:] += dgh[m,t,dir,:] * W_hh_dir        ([M,3H] x [3H,H])
            const int tf = s, tr = T - 1 - s;
            // (skinny: M x H outputs over K = 3H; split-K spreads it over the SMs until the cluster kernel takes over)
            const int rsplit = H >= 256 ? 8 : 1;
            HA2G_CHECK(ha2g_gemm(dgh + ((size_t)tf * 2 + 0) * 3 * H, w_hh_f, dh_rec, nullptr, M, H, 3 * H,
                                     T * 6 * H, H, 2 * H, 0, 0, 0, 1, rsplit, stream));
            HA2G_CHECK(ha2g_gemm(dgh + ((size_t)tr * 2 + 1) * 3 * H, w_hh_r, dh_rec + H, nullptr, M, H, 3 * H,
                                     T * 6 * H, H, 2 * H, 0, 0, 0, 1, rsplit, stream));
        }
    }
    const int split = 4;
    // dW_ih_dir += dgi[:,dir]^T x            ([3H, MT] x [MT, I])
    if (dw_ih_r == dw_ih_f + (size_t)3 * H * I) {
        // the two directions' gradients are adjacent: one [6H x MT] x [MT x I] GEMM, x packed once
        HA2G_CHECK(ha2g_gemm(dgi, x, dw_ih_f, nullptr, 6 * H, I, MT, 6 * H, I, I, 1, 0, 0, 1, split, stream));
    } else {
        HA2G_CHECK(ha2g_gemm(dgi, x, dw_ih_f, nullptr, 3 * H, I, MT, 6 * H, I, I, 1, 0, 0, 1, split, stream));
        HA2G_CHECK(ha2g_gemm(dgi + 3 * H, x, dw_ih_r, nullptr, 3 * H, I, MT, 6 * H, I, I, 1, 0, 0, 1, split, stream));
    }
    // dW_hh_f += sum_{m,t>=1} dgh[m,t,0]^T y[m,t-1,0:H];  dW_hh_r += sum_{m,t<=T-2} dgh[m,t,1]^T y[m,t+1,H:2H]
    if (T > 1) {
        HA2G_CHECK(ha2g_gemm_kseg(dgh + (size_t)6 * H, y, dw_hh_f, nullptr, 3 * H, H, M * (T - 1), 6 * H, 2 * H, H, 1,
                                      0, 0, 1, split, T - 1, T, stream));
        HA2G_CHECK(ha2g_gemm_kseg(dgh + 3 * H, y + (size_t)2 * H + H, dw_hh_r, nullptr, 3 * H, H, M * (T - 1), 6 * H,
                                      2 * H, H, 1, 0, 0, 1, split, T - 1, T, stream));
    }
    // biases: column sums over all (m,t) rows
    if (db_ih_r == db_ih_f + 3 * H) {
        HA2G_CHECK(ha2g_col_sum(dgi, MT, 6 * H, 6 * H, db_ih_f, stream));
    } else {
        HA2G_CHECK(ha2g_col_sum(dgi, MT, 3 * H, 6 * H, db_ih_f, stream));
        HA2G_CHECK(ha2g_col_sum(dgi + 3 * H, MT, 3 * H, 6 * H, db_ih_r, stream));
    }
    if (db_hh_r == db_hh_f + 3 * H) {
        HA2G_CHECK(ha2g_col_sum(dgh, MT, 6 * H, 6 * H, db_hh_f, stream));
    } else {
        HA2G_CHECK(ha2g_col_sum(dgh, MT, 3 * H, 6 * H, db_hh_f, stream));
        HA2G_CHECK(ha2g_col_sum(dgh + 3 * H, MT, 3 * H, 6 * H, db_hh_r, stream));
    }
    // dx = dgi_f W_ih_f + dgi_r W_ih_r            ([MT,3H] x [3H,I])
    if (dx != nullptr) {
        if (w_ih_r == w_ih_f + (size_t)3 * H * I) {
            // both directions' input weights are adjacent (one [6H, I] matrix): ONE GEMM over K = 6H, dgi packed once
            HA2G_CHECK(ha2g_gemm(dgi, w_ih_f, dx, nullptr, MT, I, 6 * H, 6 * H, I, I, 0, 0, 0, 0, 1, stream));
        } else {
            HA2G_CHECK(ha2g_gemm(dgi, w_ih_f, dx, nullptr, MT, I, 3 * H, 6 * H, I, I, 0, 0, 0, 0, 1, stream));
            HA2G_CHECK(ha2g_gemm(dgi + 3 * H, w_ih_r, dx, nullptr, MT, I, 3 * H, 6 * H, I, I, 0, 0, 0, 1, 1, stream));
        }
    }
    HA2G_RETURN_LAST();
}
