// Train/eval BatchNorm over a [rows, C] channels-last matrix (BatchNorm2d on NHWC maps, BatchNorm1d on
// [B*T, C] sequences) with the reference's op orderings fused in:
//   pre_relu = 1 :  y = BN(ReLU(x))          conv -> ReLU -> BN  (ResNetBlocks.py:24-26, ResNetSE34V2.py:127-129,157-159)
//   post_act     :  y = act(BN(x))           BN -> LeakyReLU     (hierarchy_net.py:205-210)
// Statistics are accumulated in double (sum, sum of squares) so E[x^2]-E[x]^2 is safe for the
// [-80,0] dB spectrogram stem.  HBM-bound: x is read once for stats and once for apply.
#include "common.cuh"

namespace {

// partial column sums of f(x) and f(x)^2;  blockDim = (32, 8)
__global__ void bn_stats_kernel(const float* __restrict__ x, int64_t rows, int C, int pre_relu,
                                double* __restrict__ sums /* [2][C] */, int rows_per_cta) {
    __shared__ double sh[2][8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
    const int64_t r1 = min(rows, r0 + rows_per_cta);
    float s = 0.f, q = 0.f;
    double ds = 0.0, dq = 0.0;
    if (c < C) {
        int cnt = 0;
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
            float v = x[r * C + c];
            if (pre_relu) v = fmaxf(v, 0.f);
            s += v; q += v * v;
            if (++cnt == 64) { ds += s; dq += q; s = 0.f; q = 0.f; cnt = 0; }
        }
        ds += s; dq += q;
    }
    sh[0][threadIdx.y][threadIdx.x] = ds;
    sh[1][threadIdx.y][threadIdx.x] = dq;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { a += sh[0][i][threadIdx.x]; b += sh[1][i][threadIdx.x]; }
        double* dst = sums + (size_t)blockIdx.y * 2 * C;   // this row chunk's partial [2][C]
        dst[c] = a;
        dst[C + c] = b;
    }
}

// sums[2C] = sum over the row-chunk partials in a FIXED pattern (8 interleaved lanes over the chunks, combined in lane
// order): deterministic, no atomics anywhere in the reduction.  blockDim = (32 columns, 8 chunk lanes)
__global__ void bn_combine_kernel(const double* __restrict__ part, int nparts, int C2, double* __restrict__ sums) {
    __shared__ double sh[8][33];
    const int i = blockIdx.x * 32 + threadIdx.x;
    double t = 0.0;
    if (i < C2)
        for (int p = threadIdx.y; p < nparts; p += 8) t += part[(size_t)p * C2 + i];
    sh[threadIdx.y][threadIdx.x] = t;
    __syncthreads();
    if (threadIdx.y == 0 && i < C2) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) a += sh[k][threadIdx.x];
        sums[i] = a;
    }
}

// mean / invstd from the sums; running-stat update (momentum, unbiased var) as nn.BatchNorm does in train mode
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int64_t rows, int C, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double m = sums[c] / (double)rows;
    double var = sums[C + c] / (double)rows - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean != nullptr) {
        double unb = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var, int C,
                                     float eps, float* __restrict__ mean, float* __restrict__ invstd) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    mean[c] = running_mean[c];
    invstd[c] = 1.0f / sqrtf(running_var[c] + eps);
}

__global__ void bn_apply_kernel(const float* __restrict__ x, int64_t n, int C, int pre_relu, const float* __restrict__ mean,
                                const float* __restrict__ invstd, const float* __restrict__ gamma,
                                const float* __restrict__ beta, int post_act, float* __restrict__ y) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        float v = x[i];
        if (pre_relu) v = fmaxf(v, 0.f);
        v = (v - mean[c]) * invstd[c] * gamma[c] + beta[c];
        y[i] = ha2g_act(v, post_act);
    }
}

// pass 1 of backward: dbeta[c] = sum g, dgamma[c] = sum g * xhat   with g = dy * post_act'(y)
__global__ void bn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                                     int64_t rows, int C, int pre_relu, int post_act, const float* __restrict__ mean,
                                     const float* __restrict__ invstd, double* __restrict__ sums /* [2][C] */,
                                     int rows_per_cta) {
    __shared__ double sh[2][8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
    const int64_t r1 = min(rows, r0 + rows_per_cta);
    double ds = 0.0, dq = 0.0;
    if (c < C) {
        float s = 0.f, q = 0.f;
        int cnt = 0;
        const float mu = mean[c], is = invstd[c];
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
            float g = dy[r * C + c];
            if (post_act) g *= ha2g_act_grad_from_out(y[r * C + c], post_act);
            float v = x[r * C + c];
            if (pre_relu) v = fmaxf(v, 0.f);
            s += g; q += g * (v - mu) * is;
            if (++cnt == 64) { ds += s; dq += q; s = 0.f; q = 0.f; cnt = 0; }
        }
        ds += s; dq += q;
    }
    sh[0][threadIdx.y][threadIdx.x] = ds;
    sh[1][threadIdx.y][threadIdx.x] = dq;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { a += sh[0][i][threadIdx.x]; b += sh[1][i][threadIdx.x]; }
        double* dst = sums + (size_t)blockIdx.y * 2 * C;
        dst[c] = a;
        dst[C + c] = b;
    }
}

// pass 2: dx = gamma*invstd*(g - dbeta/R - xhat*dgamma/R) [* (x>0) if pre_relu];  dgamma/dbeta (+=) written by block 0
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                                    int64_t rows, int C, int pre_relu, int post_act, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, const float* __restrict__ gamma,
                                    const double* __restrict__ sums, float* __restrict__ dx, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta) {
    const int64_t n = rows * C;
    const double invR = 1.0 / (double)rows;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        float g = dy[i];
        if (post_act) g *= ha2g_act_grad_from_out(y[i], post_act);
        float xv = x[i];
        float mask = 1.f;
        if (pre_relu) { mask = xv > 0.f ? 1.f : 0.f; xv = fmaxf(xv, 0.f); }
        float xhat = (xv - mean[c]) * invstd[c];
        float db = (float)(sums[c] * invR), dg = (float)(sums[C + c] * invR);
        dx[i] = gamma[c] * invstd[c] * (g - db - xhat * dg) * mask;
    }
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            dbeta[c] += (float)sums[c];
            dgamma[c] += (float)sums[C + c];
        }
    }
}


// ---- 128-bit variants of the two apply passes (C % 4 == 0, C <= BN_VEC_MAXC, 16-byte aligned tensors) -----------------
// Per-channel coefficients are computed once per CTA into shared memory as floats (the scalar kernels redo a double
// multiply per ELEMENT for the backward coefficients); each thread then streams float4s with its channel quad fixed by a
// single modulo per float4.
constexpr int BN_VEC_MAXC = 1024;

__global__ void __launch_bounds__(256) bn_apply_vec_kernel(const float4* __restrict__ x, int64_t n4, int C, int pre_relu,
                                                           const float* __restrict__ mean, const float* __restrict__ invstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           int post_act, float4* __restrict__ y) {
    __shared__ float sc[BN_VEC_MAXC], sm[BN_VEC_MAXC], sb[BN_VEC_MAXC];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        sc[c] = invstd[c] * gamma[c];
        sm[c] = mean[c];
        sb[c] = beta[c];
    }
    __syncthreads();
    const int C4 = C >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) << 2;
        float4 v = x[i];
        if (pre_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        // (v - mean) * (invstd * gamma) + beta: the mean is subtracted first, as in the scalar kernel (no cancellation)
        v.x = ha2g_act(fmaf(v.x - sm[c], sc[c], sb[c]), post_act);
        v.y = ha2g_act(fmaf(v.y - sm[c + 1], sc[c + 1], sb[c + 1]), post_act);
        v.z = ha2g_act(fmaf(v.z - sm[c + 2], sc[c + 2], sb[c + 2]), post_act);
        v.w = ha2g_act(fmaf(v.w - sm[c + 3], sc[c + 3], sb[c + 3]), post_act);
        y[i] = v;
    }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_vec_kernel(const float4* __restrict__ dy, const float4* __restrict__ x,
                                                               const float4* __restrict__ y, int64_t n4, int64_t rows, int C,
                                                               int pre_relu, int post_act, const float* __restrict__ mean,
                                                               const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                               const double* __restrict__ sums, float4* __restrict__ dx,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta) {
    // dx = gamma*invstd*(g - db - xhat*dg)*mask,  xhat = (xv - mean)*invstd  (same evaluation order as the scalar kernel)
    __shared__ float cA[BN_VEC_MAXC], cM[BN_VEC_MAXC], cI[BN_VEC_MAXC], cDb[BN_VEC_MAXC], cDg[BN_VEC_MAXC];
    const double invR = 1.0 / (double)rows;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        cA[c] = gamma[c] * invstd[c];
        cM[c] = mean[c];
        cI[c] = invstd[c];
        cDb[c] = (float)(sums[c] * invR);
        cDg[c] = (float)(sums[C + c] * invR);
    }
    __syncthreads();
    const int C4 = C >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) << 2;
        float4 g = dy[i];
        if (post_act) {
            const float4 yo = y[i];
            g.x *= ha2g_act_grad_from_out(yo.x, post_act); g.y *= ha2g_act_grad_from_out(yo.y, post_act);
            g.z *= ha2g_act_grad_from_out(yo.z, post_act); g.w *= ha2g_act_grad_from_out(yo.w, post_act);
        }
        const float4 xv = x[i];
        float4 o;
#define BN_BWD_1(f, k)                                                            \
        {                                                                         \
            float xx = xv.f, m = 1.f;                                             \
            if (pre_relu) { m = xx > 0.f ? 1.f : 0.f; xx = fmaxf(xx, 0.f); }      \
            o.f = cA[c + k] * (g.f - cDb[c + k] - (xx - cM[c + k]) * cI[c + k] * cDg[c + k]) * m; \
        }
        BN_BWD_1(x, 0) BN_BWD_1(y, 1) BN_BWD_1(z, 2) BN_BWD_1(w, 3)
#undef BN_BWD_1
        dx[i] = o;
    }
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            dbeta[c] += (float)sums[c];
            dgamma[c] += (float)sums[C + c];
        }
    }
}

static inline bool bn_vec_ok(int C, const void* a, const void* b, const void* c, const void* d) {
    return C % 4 == 0 && C <= BN_VEC_MAXC &&
           (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d) & 15) == 0;
}


// 128-bit column reductions: a CTA owns a row range; thread = (row slot, channel quad).  MODE 0: sums of f(x), f(x)^2
// (forward statistics);  MODE 1: sums of g and g*xhat (backward).  Partial sums are kept in float for 64 rows at a time
// and flushed into double accumulators, exactly like the scalar kernels.
template <int MODE>
__global__ void __launch_bounds__(256) bn_reduce_vec_kernel(const float4* __restrict__ x, const float4* __restrict__ dy,
                                                            const float4* __restrict__ y, int64_t rows, int C, int pre_relu,
                                                            int post_act, const float* __restrict__ mean,
                                                            const float* __restrict__ invstd, double* __restrict__ sums,
                                                            int rows_per_cta) {
    __shared__ double sh[2][256][4];
    const int C4 = C >> 2;
    const int slots = 256 / C4;
    const int q = threadIdx.x % C4, slot = threadIdx.x / C4;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + (int64_t)rows_per_cta);
    double ds[4] = {0.0, 0.0, 0.0, 0.0}, dq[4] = {0.0, 0.0, 0.0, 0.0};
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), is = mu;
    if (MODE == 1) { mu = *reinterpret_cast<const float4*>(mean + q * 4); is = *reinterpret_cast<const float4*>(invstd + q * 4); }
    float s[4] = {0.f, 0.f, 0.f, 0.f}, qq[4] = {0.f, 0.f, 0.f, 0.f};
    int cnt = 0;
    for (int64_t r = r0 + slot; r < r1; r += slots) {
        const int64_t i = r * C4 + q;
        float4 v = x[i];
        if (pre_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        if (MODE == 0) {
            s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
            qq[0] += v.x * v.x; qq[1] += v.y * v.y; qq[2] += v.z * v.z; qq[3] += v.w * v.w;
        } else {
            float4 g = dy[i];
            if (post_act) {
                const float4 yo = y[i];
                g.x *= ha2g_act_grad_from_out(yo.x, post_act); g.y *= ha2g_act_grad_from_out(yo.y, post_act);
                g.z *= ha2g_act_grad_from_out(yo.z, post_act); g.w *= ha2g_act_grad_from_out(yo.w, post_act);
            }
            s[0] += g.x; s[1] += g.y; s[2] += g.z; s[3] += g.w;
            qq[0] += g.x * (v.x - mu.x) * is.x; qq[1] += g.y * (v.y - mu.y) * is.y;
            qq[2] += g.z * (v.z - mu.z) * is.z; qq[3] += g.w * (v.w - mu.w) * is.w;
        }
        if (++cnt == 64) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { ds[k] += s[k]; dq[k] += qq[k]; s[k] = 0.f; qq[k] = 0.f; }
            cnt = 0;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { sh[0][threadIdx.x][k] = ds[k] + s[k]; sh[1][threadIdx.x][k] = dq[k] + qq[k]; }
    __syncthreads();
    if (slot == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double a = 0.0, b = 0.0;
            for (int j = 0; j < slots; ++j) { a += sh[0][j * C4 + q][k]; b += sh[1][j * C4 + q][k]; }
            double* dst = sums + (size_t)blockIdx.x * 2 * C;
            dst[q * 4 + k] = a;
            dst[C + q * 4 + k] = b;
        }
    }
}
static inline bool bn_red_vec_ok(int C, const void* a, const void* b, const void* c) {
    return C >= 8 && C <= 1024 && (C & (C - 1)) == 0 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
}
static inline void bn_red_grid(int64_t rows, int& ctas, int& rows_per) {
    int64_t want = 148 * 4;
    int64_t rp = (rows + want - 1) / want;
    if (rp < 64) rp = 64;
    rows_per = (int)rp;
    ctas = (int)((rows + rp - 1) / rp);
}

static inline void bn_grid(int64_t rows, int C, dim3& grid, int& rows_per) {
    int gx = ha2g_div_up(C, 32);
    int want_y = ha2g_div_up(148 * 4, gx);
    int64_t rp = (rows + want_y - 1) / want_y;
    if (rp < 64) rp = 64;
    rows_per = (int)rp;
    grid = dim3(gx, ha2g_div_up(rows, rp));
}

}  // namespace

// Forward.  training != 0: batch statistics (+ running-stat update when running_mean != nullptr);
// training == 0: running statistics.  mean/invstd [C] are outputs (saved for backward);
// sums_scratch: 2*C doubles.  y may alias x only when pre_relu == 0 and no backward is needed.
HA2G_API int ha2g_bn_fwd(const float* x, int64_t rows, int C, int pre_relu, int post_act, int training,
                         const float* gamma, const float* beta, float* running_mean, float* running_var, float eps,
                         float momentum, float* mean, float* invstd, double* sums_scratch, float* y, cudaStream_t stream) {
    if (rows <= 0) return 0;
    if (training) {
        dim3 grid; int rows_per, nparts;
        const bool vec = bn_red_vec_ok(C, x, x, x);
        if (vec) bn_red_grid(rows, nparts, rows_per);
        else { bn_grid(rows, C, grid, rows_per); nparts = (int)grid.y; }
        double* part = reinterpret_cast<double*>(ha2g_ws((size_t)nparts * 2 * C * sizeof(double), stream));
        if (part == nullptr) return (int)cudaErrorMemoryAllocation;
        if (vec)
            bn_reduce_vec_kernel<0><<<nparts, 256, 0, stream>>>(reinterpret_cast<const float4*>(x), nullptr, nullptr, rows, C, pre_relu,
                                                                0, nullptr, nullptr, part, rows_per);
        else
            bn_stats_kernel<<<grid, dim3(32, 8), 0, stream>>>(x, rows, C, pre_relu, part, rows_per);
        bn_combine_kernel<<<ha2g_div_up(2 * C, 32), dim3(32, 8), 0, stream>>>(part, nparts, 2 * C, sums_scratch);
        bn_finalize_kernel<<<ha2g_div_up(C, 128), 128, 0, stream>>>(sums_scratch, rows, C, eps, momentum, mean, invstd,
                                                                    running_mean, running_var);
    } else {
        bn_eval_stats_kernel<<<ha2g_div_up(C, 128), 128, 0, stream>>>(running_mean, running_var, C, eps, mean, invstd);
    }
    const int64_t n = rows * C;
    if (bn_vec_ok(C, x, y, x, y))
        bn_apply_vec_kernel<<<ha2g_ew_grid(n / 4, 256, 4), 256, 0, stream>>>(reinterpret_cast<const float4*>(x), n / 4, C, pre_relu,
                                                                            mean, invstd, gamma, beta, post_act,
                                                                            reinterpret_cast<float4*>(y));
    else
        bn_apply_kernel<<<ha2g_ew_grid(n), 256, 0, stream>>>(x, n, C, pre_relu, mean, invstd, gamma, beta, post_act, y);
    HA2G_RETURN_LAST();
}

// Backward of the training-mode forward.  dgamma/dbeta are ACCUMULATED (+=).  y is only read when post_act != 0.
HA2G_API int ha2g_bn_bwd(const float* dy, const float* x, const float* y, int64_t rows, int C, int pre_relu, int post_act,
                         const float* gamma, const float* mean, const float* invstd, double* sums_scratch, float* dx,
                         float* dgamma, float* dbeta, cudaStream_t stream) {
    if (rows <= 0) return 0;
    dim3 grid; int rows_per, nparts;
    const bool rvec = bn_red_vec_ok(C, x, dy, post_act ? (const void*)y : (const void*)x) && bn_red_vec_ok(C, mean, invstd, x);
    if (rvec) bn_red_grid(rows, nparts, rows_per);
    else { bn_grid(rows, C, grid, rows_per); nparts = (int)grid.y; }
    double* part = reinterpret_cast<double*>(ha2g_ws((size_t)nparts * 2 * C * sizeof(double), stream));
    if (part == nullptr) return (int)cudaErrorMemoryAllocation;
    if (rvec)
        bn_reduce_vec_kernel<1><<<nparts, 256, 0, stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(dy),
                                                            reinterpret_cast<const float4*>(y), rows, C, pre_relu, post_act, mean,
                                                            invstd, part, rows_per);
    else
        bn_bwd_reduce_kernel<<<grid, dim3(32, 8), 0, stream>>>(dy, x, y, rows, C, pre_relu, post_act, mean, invstd, part, rows_per);
    bn_combine_kernel<<<ha2g_div_up(2 * C, 32), dim3(32, 8), 0, stream>>>(part, nparts, 2 * C, sums_scratch);
    if (bn_vec_ok(C, dy, x, post_act ? (const void*)y : (const void*)x, dx))
        bn_bwd_apply_vec_kernel<<<ha2g_ew_grid(rows * C / 4, 256, 4), 256, 0, stream>>>(
            reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(y),
            rows * C / 4, rows, C, pre_relu, post_act, mean, invstd, gamma, sums_scratch, reinterpret_cast<float4*>(dx), dgamma,
            dbeta);
    else
        bn_bwd_apply_kernel<<<ha2g_ew_grid(rows * C), 256, 0, stream>>>(dy, x, y, rows, C, pre_relu, post_act, mean, invstd,
                                                                        gamma, sums_scratch, dx, dgamma, dbeta);
    HA2G_RETURN_LAST();
}
