// Loss kernels of the HA2G step (K15-K18 of SURVEY.md).  Every forward kernel also emits the
// un-scaled gradient of its scalar wrt its input, so backward is one multiply by the upstream
// device scalar (no host sync anywhere).
//   Huber                 train_hierarchy_expressive.py:312-318
//   GAN log losses        :224, :322
//   KLD                   :410
//   diversity regulariser :396-406
//   physical (bone angle) :426-449 (expressive, 42 bones + 2 palm normals) / train_hierarchy.py:242-262
//   contrastive           SoftmaxContrastiveLoss, train_hierarchy.py:54-68 / ..._expressive.py:108-121
#include "common.cuh"
#include <math_constants.h>

namespace {

// Deterministic grid-wide accumulation of one value per CTA into a device scalar: every CTA stores its partial, takes an
// integer ticket, and the LAST CTA to arrive adds the partials in a fixed pattern (lane-strided, then a shuffle tree) --
// the result does not depend on the order in which the CTAs ran.  One scratch serves every loss kernel: they are issued
// on one stream (or one captured chain), never concurrently.
constexpr int RED_MAX = 4096;
__device__ float g_red_part[RED_MAX];
__device__ unsigned int g_red_ticket;
__device__ __forceinline__ void grid_accumulate(float block_val /* valid in thread 0 */, float* __restrict__ loss) {
    __shared__ int s_last;
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const unsigned nblk = gridDim.x * gridDim.y;
    const unsigned bid = blockIdx.x + blockIdx.y * gridDim.x;
    if (nblk == 1) {
        if (tid == 0) *loss += block_val;
        return;
    }
    if (tid == 0) {
        g_red_part[bid] = block_val;
        __threadfence();
        s_last = atomicAdd(&g_red_ticket, 1u) == nblk - 1 ? 1 : 0;
    }
    __syncthreads();
    if (s_last && tid < 32) {
        __threadfence();
        float t = 0.f;
        for (unsigned i = tid; i < nblk; i += 32) t += reinterpret_cast<volatile float*>(g_red_part)[i];
        t = warp_sum(t);
        if (tid == 0) {
            *loss += t;
            g_red_ticket = 0u;
        }
    }
}

// loss += beta * mean(smooth_l1((o-t)/beta));  grad[i] = clamp((o-t)/beta, -1, 1) / n
__global__ void huber_kernel(const float* __restrict__ o, const float* __restrict__ t, float* __restrict__ grad, int64_t n,
                             float beta, float* __restrict__ loss) {
    __shared__ float sh[33];
    float acc = 0.f;
    const float invb = 1.f / beta, invn = 1.f / (float)n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float x = o[i] * invb - t[i] * invb;  // matches (o/beta - t/beta) of the reference
        float ax = fabsf(x);
        acc += ax < 1.f ? 0.5f * x * x : ax - 0.5f;
        if (grad != nullptr) grad[i] = fminf(fmaxf(x, -1.f), 1.f) * invn;
    }
    acc = block_sum(acc, sh);
    grid_accumulate(acc * beta * invn, loss);
}

// mode 0: loss += -mean(log(x + 1e-8)),     grad = -1/(n (x+1e-8))
// mode 1: loss += -mean(log(1 - x + 1e-8)), grad = +1/(n (1-x+1e-8))
__global__ void log_loss_kernel(const float* __restrict__ x, float* __restrict__ grad, int64_t n, int mode,
                                float* __restrict__ loss) {
    __shared__ float sh[33];
    float acc = 0.f;
    const float invn = 1.f / (float)n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = x[i];
        if (mode == 0) {
            float a = v + 1e-8f;
            acc -= logf(a);
            if (grad != nullptr) grad[i] = -invn / a;
        } else {
            float a = 1.f - v + 1e-8f;
            acc -= logf(a);
            if (grad != nullptr) grad[i] = invn / a;
        }
    }
    acc = block_sum(acc, sh);
    grid_accumulate(acc * invn, loss);
}

// kld = -0.5 * mean(1 + lv - mu^2 - exp(lv));  dmu = mu/n;  dlv = -0.5 (1 - exp(lv))/n
__global__ void kld_kernel(const float* __restrict__ mu, const float* __restrict__ lv, float* __restrict__ dmu,
                           float* __restrict__ dlv, int64_t n, float* __restrict__ loss) {
    __shared__ float sh[33];
    float acc = 0.f;
    const float invn = 1.f / (float)n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float m = mu[i], l = lv[i], e = expf(l);
        acc += 1.f + l - m * m - e;
        dmu[i] = m * invn;
        dlv[i] = -0.5f * (1.f - e) * invn;
    }
    acc = block_sum(acc, sh);
    grid_accumulate(-0.5f * acc * invn, loss);
}

// one CTA per sample b:  pose_l1 = sum_{t,d} beta*smooth_l1((o-r)/beta);  z_l1 = mean_j |z - zr|
// div_b = max(-pose_l1/(z_l1+1e-5), -1000);  loss += div_b / B;  grad[b,:] = d div_b/d o / B
__global__ void div_reg_kernel(const float* __restrict__ o, const float* __restrict__ r, const float* __restrict__ z,
                               const float* __restrict__ zr, float* __restrict__ grad, int B, int TD, int Z, float beta,
                               float* __restrict__ loss) {
    __shared__ float sh[33];
    const int b = blockIdx.x;
    const float invb = 1.f / beta;
    float acc = 0.f;
    for (int e = threadIdx.x; e < TD; e += blockDim.x) {
        float x = o[(size_t)b * TD + e] * invb - r[(size_t)b * TD + e] * invb;
        float ax = fabsf(x);
        acc += (ax < 1.f ? 0.5f * x * x : ax - 0.5f) * beta;
    }
    const float pose = block_sum(acc, sh);
    float za = 0.f;
    for (int e = threadIdx.x; e < Z; e += blockDim.x) za += fabsf(z[(size_t)b * Z + e] - zr[(size_t)b * Z + e]);
    const float zl1 = block_sum(za, sh) / (float)Z;
    const float den = zl1 + 1.0e-5f;
    const float val = -pose / den;
    const bool live = val >= -1000.f;
    const float coef = live ? -1.f / (den * (float)B) : 0.f;
    for (int e = threadIdx.x; e < TD; e += blockDim.x) {
        float x = o[(size_t)b * TD + e] * invb - r[(size_t)b * TD + e] * invb;
        grad[(size_t)b * TD + e] = coef * fminf(fmaxf(x, -1.f), 1.f);
    }
    grid_accumulate(fmaxf(val, -1000.f) / (float)B, loss);
}

__global__ void scale_by_scalar_kernel(const float* __restrict__ x, const float* __restrict__ s, float alpha,
                                       float* __restrict__ out, int64_t n, int accumulate) {
    const float f = alpha * (s != nullptr ? *s : 1.f);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = accumulate ? out[i] + f * x[i] : f * x[i];
}

// ---- physical loss: one warp per (b,t) row -------------------------------------------------------------
constexpr int PHY_MAXV = 44;
constexpr int PHY_MAXP = 48;
__constant__ int c_phy_pairs[2][PHY_MAXP][2];
__constant__ float c_phy_avg[2][PHY_MAXP];
__constant__ float c_phy_var[2][PHY_MAXP];
__constant__ float c_phy_mean[2][3 * 42];

// variant 0: gesture (9 bones, 4 pairs, no palms), 1: expressive (42 bones + 2 palms, 41 pairs)
__global__ void physical_kernel(const float* __restrict__ out, float* __restrict__ grad, int64_t rows, int variant,
                                int nb, int npairs, float* __restrict__ loss) {
    constexpr int WPB = 4;
    __shared__ float raw[WPB][PHY_MAXV * 3];
    __shared__ float unit[WPB][PHY_MAXV * 3];
    __shared__ float nrm[WPB][PHY_MAXV];
    __shared__ float dun[WPB][PHY_MAXV * 3];
    __shared__ float wl[WPB];
    const int w = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int64_t row = (int64_t)blockIdx.x * WPB + w;
    const bool active = row < rows;
    const int D = nb * 3;
    const int nv = variant == 1 ? nb + 2 : nb;
    float lsum = 0.f;
    if (active) {
        for (int e = lane; e < D; e += 32) raw[w][e] = out[row * D + e] + c_phy_mean[variant][e];
        for (int e = lane; e < nv * 3; e += 32) dun[w][e] = 0.f;
    }
    __syncwarp();
    if (active && variant == 1 && lane < 2) {
        // palm normals: cross(bone 11, bone 17) -> vec 42; cross(bone 28, bone 34) -> vec 43
        const int a = lane == 0 ? 11 : 28, b = lane == 0 ? 17 : 34;
        const float ax = raw[w][a * 3], ay = raw[w][a * 3 + 1], az = raw[w][a * 3 + 2];
        const float bx = raw[w][b * 3], by = raw[w][b * 3 + 1], bz = raw[w][b * 3 + 2];
        raw[w][(nb + lane) * 3 + 0] = ay * bz - az * by;
        raw[w][(nb + lane) * 3 + 1] = az * bx - ax * bz;
        raw[w][(nb + lane) * 3 + 2] = ax * by - ay * bx;
    }
    __syncwarp();
    if (active) {
        for (int v = lane; v < nv; v += 32) {
            float x = raw[w][v * 3], y = raw[w][v * 3 + 1], z = raw[w][v * 3 + 2];
            float n = fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
            nrm[w][v] = n;
            unit[w][v * 3] = x / n; unit[w][v * 3 + 1] = y / n; unit[w][v * 3 + 2] = z / n;
        }
    }
    __syncwarp();
    const float lo = (float)(-1.0 + 1e-7), hi = (float)(1.0 - 1e-7);
    const float inv_rows = 1.f / (float)rows;
    if (active) {
        for (int p = lane; p < npairs; p += 32) {
            const int i0 = c_phy_pairs[variant][p][0], i1 = c_phy_pairs[variant][p][1];
            const float* u0 = &unit[w][i0 * 3];
            const float* u1 = &unit[w][i1 * 3];
            float ip = u0[0] * u1[0] + u0[1] * u1[1] + u0[2] * u1[2];
            const bool inside = ip >= lo && ip <= hi;
            ip = fminf(fmaxf(ip, lo), hi);
            const float ang = acosf(ip) / CUDART_PI_F;
            const float dlt = ang - c_phy_avg[variant][p];
            const float iv = 1.f / (2.f * c_phy_var[variant][p]);
            lsum += dlt * dlt * iv;
            if (grad != nullptr && inside) {
                // d/d ip: 2*dlt*iv * (1/pi) * (-1/sqrt(1-ip^2)) / rows
                const float gip = 2.f * dlt * iv * (-1.f / (CUDART_PI_F * sqrtf(1.f - ip * ip))) * inv_rows;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    atomicAdd(&dun[w][i0 * 3 + c], gip * u1[c]);
                    atomicAdd(&dun[w][i1 * 3 + c], gip * u0[c]);
                }
            }
        }
    }
    lsum = warp_sum(lsum);
    if (lane == 0) wl[w] = active ? lsum : 0.f;
    __syncwarp();
    if (grad != nullptr && active) {
        // normalize backward: draw = (du - (du . u) u) / max(|raw|, eps)   (in place in dun)
        for (int v = lane; v < nv; v += 32) {
            float dx = dun[w][v * 3], dy = dun[w][v * 3 + 1], dz = dun[w][v * 3 + 2];
            float ux = unit[w][v * 3], uy = unit[w][v * 3 + 1], uz = unit[w][v * 3 + 2];
            float dot = dx * ux + dy * uy + dz * uz;
            float inv = 1.f / nrm[w][v];
            dun[w][v * 3] = (dx - dot * ux) * inv; dun[w][v * 3 + 1] = (dy - dot * uy) * inv; dun[w][v * 3 + 2] = (dz - dot * uz) * inv;
        }
        __syncwarp();
        if (variant == 1 && lane < 2) {
            // c = a x b:  da += b x dc,  db += dc x a
            const int a = lane == 0 ? 11 : 28, b = lane == 0 ? 17 : 34, cidx = nb + lane;
            const float ax = raw[w][a * 3], ay = raw[w][a * 3 + 1], az = raw[w][a * 3 + 2];
            const float bx = raw[w][b * 3], by = raw[w][b * 3 + 1], bz = raw[w][b * 3 + 2];
            const float cx = dun[w][cidx * 3], cy = dun[w][cidx * 3 + 1], cz = dun[w][cidx * 3 + 2];
            atomicAdd(&dun[w][a * 3 + 0], by * cz - bz * cy);
            atomicAdd(&dun[w][a * 3 + 1], bz * cx - bx * cz);
            atomicAdd(&dun[w][a * 3 + 2], bx * cy - by * cx);
            atomicAdd(&dun[w][b * 3 + 0], cy * az - cz * ay);
            atomicAdd(&dun[w][b * 3 + 1], cz * ax - cx * az);
            atomicAdd(&dun[w][b * 3 + 2], cx * ay - cy * ax);
        }
        __syncwarp();
        for (int e = lane; e < D; e += 32) grad[row * D + e] = dun[w][e];
    }
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x == 0)
        for (int i = 0; i < WPB; ++i) t += wl[i];
    grid_accumulate(t * inv_rows, loss);
}

// ---- contrastive --------------------------------------------------------------------------------------
constexpr int CC = 32;      // feature width (nOut of both encoders)
constexpr int CT = 128;     // rows per CTA / tile

// xn = x / max(|x|, 1e-12), norms saved.   one warp per row, lane = channel
__global__ void l2norm_rows_kernel(const float* __restrict__ x, float* __restrict__ xn, float* __restrict__ nrm, int64_t N) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (row >= N) return;
    float v = x[row * CC + lane];
    float s = warp_sum(v * v);
    float n = fmaxf(sqrtf(s), 1e-12f);
    xn[row * CC + lane] = v / n;
    if (lane == 0) nrm[row] = n;
}
// dx = (dn - (dn . n) n) / norm, with dn = the sum of `planes` partial planes [planes][N][32] added in index order (the
// column splits of contrastive_pair_kernel<1> each write their own plane: no atomics)
__global__ void l2norm_rows_bwd_kernel(const float* __restrict__ dn, int planes, const float* __restrict__ xn,
                                       const float* __restrict__ nrm, float* __restrict__ dx, int64_t N) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (row >= N) return;
    float g = 0.f;
    for (int p = 0; p < planes; ++p) g += dn[((size_t)p * N + row) * CC + lane];
    const float n = xn[row * CC + lane];
    float dot = warp_sum(g * n);
    dx[row * CC + lane] = (g - dot * n) / nrm[row];
}

__device__ __forceinline__ float contrastive_logit(float d, int variant) {
    return variant == 0 ? fmaxf(1.0f / (d + 1e-8f), 1e-8f) : 1.0f / d;
}

// MODE 0: forward partial (max, sumexp) of row i over the column split -> part[i][split][2]; diag logit -> diag[i]
// MODE 1: gradient wrt the OWNED rows (rows of `own`), other side streamed.  own_is_a: owned rows are a (loss rows i)
//         da_i += sum_j c_ij (a_i - b_j)      c_ij = gs * (p_ij - delta_ij) * dl/dD / D,   p_ij = exp(l_ij - lse_i)
//         (own_is_a == 0): db_j += sum_i c_ij (b_j - a_i)
template <int MODE>
__global__ void __launch_bounds__(CT, 4) contrastive_pair_kernel(const float* __restrict__ own, const float* __restrict__ oth,
                                                              const float* __restrict__ lse, float* __restrict__ part,
                                                              float* __restrict__ diag, float* __restrict__ dgrad,
                                                              const float* __restrict__ gscale, int N, int variant,
                                                              int own_is_a, int cols_per_split, int Noth, int off) {
    // rectangular form (data-parallel global batch): `own` has N rows, `oth` has Noth rows; the loss rows are always the
    // rows of a, whose label is column (row + off) of b.  Square single-GPU case: Noth == N, off == 0.
    // every lane of a warp reads the SAME tile row (its own row of `own` lives in registers): unpadded rows, 128-bit
    // broadcast loads -- 8 LDS.128 per pair instead of 32 LDS.32
    __shared__ __align__(16) float tile[CT][CC];
    __shared__ float tlse[CT];
    const int i = blockIdx.x * CT + threadIdx.x;
    const bool valid = i < N;
    float me[CC];
#pragma unroll
    for (int c = 0; c < CC; ++c) me[c] = valid ? own[(size_t)i * CC + c] : 0.f;
    const int j_beg = blockIdx.y * cols_per_split, j_end = min(Noth, j_beg + cols_per_split);
    const int Na = own_is_a ? N : Noth;   // number of loss rows (rows of a)
    float m = -CUDART_INF_F, s = 0.f;
    float acc[CC];
    float my_lse = 0.f, gs = 0.f;
    if (MODE == 1) {
#pragma unroll
        for (int c = 0; c < CC; ++c) acc[c] = 0.f;
        gs = (gscale != nullptr ? *gscale : 1.f) / (float)Na;
        if (own_is_a && valid) my_lse = lse[i];
    }
    for (int j0 = j_beg; j0 < j_end; j0 += CT) {
        __syncthreads();
        for (int e = threadIdx.x; e < CT * CC / 4; e += CT) {   // 16-byte coalesced copies ([N,32] fp32 rows are 128 B)
            const int r = e / (CC / 4), c4 = e % (CC / 4);
            const int j = j0 + r;
            reinterpret_cast<float4*>(&tile[r][0])[c4] =
                j < j_end ? reinterpret_cast<const float4*>(oth + (size_t)j * CC)[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (MODE == 1 && !own_is_a) {
            int j = j0 + threadIdx.x;
            tlse[threadIdx.x] = j < j_end ? lse[j] : 0.f;
        }
        __syncthreads();
        if (!valid) continue;
        const int cnt = min(CT, j_end - j0);
        for (int r = 0; r < cnt; ++r) {
            float t[CC];
#pragma unroll
            for (int c4 = 0; c4 < CC / 4; ++c4) {
                const float4 v = reinterpret_cast<const float4*>(&tile[r][0])[c4];
                t[c4 * 4] = v.x; t[c4 * 4 + 1] = v.y; t[c4 * 4 + 2] = v.z; t[c4 * 4 + 3] = v.w;
            }
            float d2 = 0.f;
#pragma unroll
            for (int c = 0; c < CC; ++c) { float df = me[c] - t[c]; d2 = fmaf(df, df, d2); }
            const float d = sqrtf(d2);
            const float l = contrastive_logit(d, variant);
            const int j = j0 + r;
            const bool is_diag = own_is_a ? (j == i + off) : (j + off == i);
            if (MODE == 0) {
                if (is_diag) diag[i] = l;
                if (l > m) { s = s * expf(m - l) + 1.f; m = l; }
                else s += expf(l - m);
            } else {
                const float row_lse = own_is_a ? my_lse : tlse[r];
                float p = expf(l - row_lse);
                if (is_diag) p -= 1.f;
                float dld;
                if (variant == 0) { float t = d + 1e-8f; dld = -1.f / (t * t); }
                else dld = -1.f / (d * d);
                float cf = d > 0.f ? gs * p * dld / d : 0.f;
#pragma unroll
                for (int c = 0; c < CC; ++c) acc[c] = fmaf(cf, me[c] - t[c], acc[c]);
            }
        }
    }
    if (!valid) return;
    if (MODE == 0) {
        part[((size_t)i * gridDim.y + blockIdx.y) * 2 + 0] = m;
        part[((size_t)i * gridDim.y + blockIdx.y) * 2 + 1] = s;
    } else {
#pragma unroll
        for (int c = 0; c < CC; ++c) dgrad[((size_t)blockIdx.y * N + i) * CC + c] = acc[c];
    }
}

// lse_i from the split partials; loss += mean_i (lse_i - l_ii)
__global__ void contrastive_finalize_kernel(const float* __restrict__ part, const float* __restrict__ diag, int N, int S,
                                            float* __restrict__ lse, float* __restrict__ loss) {
    __shared__ float sh[33];
    float acc = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        float m = -CUDART_INF_F;
        for (int k = 0; k < S; ++k) m = fmaxf(m, part[((size_t)i * S + k) * 2]);
        float s = 0.f;
        for (int k = 0; k < S; ++k) s += part[((size_t)i * S + k) * 2 + 1] * expf(part[((size_t)i * S + k) * 2] - m);
        float l = m + logf(s);
        lse[i] = l;
        acc += l - diag[i];
    }
    acc = block_sum(acc, sh);
    grid_accumulate(acc / (float)N, loss);
}

}  // namespace

// loss (device scalar, ACCUMULATED) += beta*mean(smooth_l1(o/beta, t/beta));  grad (nullable) = d/d o, unscaled
HA2G_API int ha2g_huber(const float* o, const float* t, float* grad, int64_t n, float beta, float* loss,
                        cudaStream_t stream) {
    huber_kernel<<<ha2g_ew_grid(n), 256, 0, stream>>>(o, t, grad, n, beta, loss);
    HA2G_RETURN_LAST();
}
// mode 0: -mean(log(x+1e-8)); mode 1: -mean(log(1-x+1e-8));  loss ACCUMULATED
HA2G_API int ha2g_log_loss(const float* x, float* grad, int64_t n, int mode, float* loss, cudaStream_t stream) {
    log_loss_kernel<<<ha2g_ew_grid(n), 256, 0, stream>>>(x, grad, n, mode, loss);
    HA2G_RETURN_LAST();
}
HA2G_API int ha2g_kld(const float* mu, const float* logvar, float* dmu, float* dlogvar, int64_t n, float* loss,
                      cudaStream_t stream) {
    kld_kernel<<<ha2g_ew_grid(n), 256, 0, stream>>>(mu, logvar, dmu, dlogvar, n, loss);
    HA2G_RETURN_LAST();
}
// o, r: [B,TD];  z, zr: [B,Z];  grad [B,TD] = d loss / d o (unscaled);  loss ACCUMULATED
HA2G_API int ha2g_div_reg(const float* o, const float* r, const float* z, const float* zr, float* grad, int B, int TD,
                          int Z, float beta, float* loss, cudaStream_t stream) {
    if (B > RED_MAX) return (int)cudaErrorInvalidValue;
    div_reg_kernel<<<B, 256, 0, stream>>>(o, r, z, zr, grad, B, TD, Z, beta, loss);
    HA2G_RETURN_LAST();
}
// out = alpha * (*s) * x  (s nullable => 1);  accumulate != 0: out += ...
HA2G_API int ha2g_scale_by_scalar(const float* x, const float* s, float alpha, float* out, int64_t n, int accumulate,
                                  cudaStream_t stream) {
    if (n <= 0) return 0;
    scale_by_scalar_kernel<<<ha2g_ew_grid(n), 256, 0, stream>>>(x, s, alpha, out, n, accumulate);
    HA2G_RETURN_LAST();
}
// Upload the physical-loss tables (host pointers): pairs [npairs][2], avg/var [npairs], mean_dir_vec [3*nb].
HA2G_API int ha2g_physical_set_tables(int variant, const int* pairs, const float* avg, const float* var, int npairs,
                                      const float* mean_dir_vec, int nb) {
    if (variant < 0 || variant > 1 || npairs > PHY_MAXP || nb > 42) return (int)cudaErrorInvalidValue;
    cudaError_t e;
    e = cudaMemcpyToSymbol(c_phy_pairs, pairs, sizeof(int) * 2 * npairs, sizeof(int) * 2 * PHY_MAXP * variant);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpyToSymbol(c_phy_avg, avg, sizeof(float) * npairs, sizeof(float) * PHY_MAXP * variant);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpyToSymbol(c_phy_var, var, sizeof(float) * npairs, sizeof(float) * PHY_MAXP * variant);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpyToSymbol(c_phy_mean, mean_dir_vec, sizeof(float) * 3 * nb, sizeof(float) * 3 * 42 * variant);
    return (int)e;
}
// out [rows, 3*nb]; loss ACCUMULATED += sum_p mean_rows((angle_p - avg_p)^2 / (2 var_p)); grad nullable, unscaled
HA2G_API int ha2g_physical(const float* out, float* grad, int64_t rows, int variant, int nb, int npairs, float* loss,
                           cudaStream_t stream) {
    if (ha2g_div_up(rows, 4) > RED_MAX) return (int)cudaErrorInvalidValue;
    physical_kernel<<<ha2g_div_up(rows, 4), 128, 0, stream>>>(out, grad, rows, variant, nb, npairs, loss);
    HA2G_RETURN_LAST();
}
extern "C" int ha2g_contrastive_fwd_rect(const float*, const float*, float*, float*, float*, float*, float*, float*, float*,
                                         int, int, int, int, float*, cudaStream_t);
extern "C" int ha2g_contrastive_bwd_rect(const float*, const float*, const float*, const float*, const float*, const float*,
                                         float*, float*, float*, float*, int, int, int, int, cudaStream_t);
// Rectangular SoftmaxContrastiveLoss for the data-parallel global batch (the reference computes the loss over the whole
// DataParallel batch, scripts/train_expressive.py:184-197 + train_hierarchy_expressive.py:244-249): a [Na,32] are this
// rank's rows, b [Nb,32] the all-gathered columns, row i's positive is column i + off (off = rank * Na).
// loss ACCUMULATED += mean over the Na local rows.  part: [Na * ceil(Nb/128) * 2] floats (upper bound), diag/lse/na: [Na],
// an [Na,32], bn [Nb,32], nb [Nb].
HA2G_API int ha2g_contrastive_fwd_rect(const float* a, const float* b, float* an, float* bn, float* na, float* nb,
                                       float* lse, float* part, float* diag, int Na, int Nb, int off, int variant,
                                       float* loss, cudaStream_t stream) {
    const int rows_ctas = ha2g_div_up(Na, CT);
    int S = (148 * 4 + rows_ctas - 1) / rows_ctas;      // 4 resident CTAs per SM (launch bounds of the pair kernel)
    const int max_s = ha2g_div_up(Nb, CT);
    if (S > max_s) S = max_s;
    if (S < 1) S = 1;
    const int cols = ((Nb + S - 1) / S + CT - 1) / CT * CT;
    l2norm_rows_kernel<<<ha2g_div_up(Na, 8), 256, 0, stream>>>(a, an, na, Na);
    l2norm_rows_kernel<<<ha2g_div_up(Nb, 8), 256, 0, stream>>>(b, bn, nb, Nb);
    dim3 grid(rows_ctas, ha2g_div_up(Nb, cols));
    contrastive_pair_kernel<0><<<grid, CT, 0, stream>>>(an, bn, nullptr, part, diag, nullptr, nullptr, Na, variant, 1, cols, Nb, off);
    contrastive_finalize_kernel<<<ha2g_div_up(Na, 256), 256, 0, stream>>>(part, diag, Na, grid.y, lse, loss);
    HA2G_RETURN_LAST();
}
// Backward of the rectangular loss: da [Na,32] and db [Nb,32] (the gradient wrt ALL gathered columns: the caller
// reduce-scatters it to the owning ranks) = gscale * d loss / d a, b.  dan [Na,32], dbn [Nb,32]: zero-initialised scratch.
HA2G_API int ha2g_contrastive_bwd_rect(const float* an, const float* bn, const float* na, const float* nb, const float* lse,
                                       const float* gscale, float* dan, float* dbn, float* da, float* db, int Na, int Nb,
                                       int off, int variant, cudaStream_t stream) {
    auto splits = [](int rows, int oth) {
        const int rc = (rows + CT - 1) / CT;
        int S = (148 * 4 + rc - 1) / rc;
        const int mx = (oth + CT - 1) / CT;
        if (S > mx) S = mx;
        if (S < 1) S = 1;
        return ((oth + S - 1) / S + CT - 1) / CT * CT;
    };
    const int cols_a = splits(Na, Nb), cols_b = splits(Nb, Na);
    dim3 grid_a(ha2g_div_up(Na, CT), ha2g_div_up(Nb, cols_a)), grid_b(ha2g_div_up(Nb, CT), ha2g_div_up(Na, cols_b));
    // per-column-split gradient planes in the scratch arena (dan / dbn of the signature are no longer needed)
    (void)dan; (void)dbn;
    const size_t pa = (size_t)grid_a.y * Na * CC, pb = (size_t)grid_b.y * Nb * CC;
    float* plane_a = reinterpret_cast<float*>(ha2g_ws((pa + pb) * sizeof(float)));
    if (plane_a == nullptr) return (int)cudaErrorMemoryAllocation;
    float* plane_b = plane_a + pa;
    contrastive_pair_kernel<1><<<grid_a, CT, 0, stream>>>(an, bn, lse, nullptr, nullptr, plane_a, gscale, Na, variant, 1, cols_a, Nb, off);
    contrastive_pair_kernel<1><<<grid_b, CT, 0, stream>>>(bn, an, lse, nullptr, nullptr, plane_b, gscale, Nb, variant, 0, cols_b, Na, off);
    l2norm_rows_bwd_kernel<<<ha2g_div_up(Na, 8), 256, 0, stream>>>(plane_a, (int)grid_a.y, an, na, da, Na);
    l2norm_rows_bwd_kernel<<<ha2g_div_up(Nb, 8), 256, 0, stream>>>(plane_b, (int)grid_b.y, bn, nb, db, Nb);
    HA2G_RETURN_LAST();
}
// Streaming SoftmaxContrastiveLoss forward (replaces criterion(text_feat, feat_*) at
// scripts/train_eval/train_hierarchy_expressive.py:244-249; class at train_hierarchy.py:23-68): L2-normalise rows,
// logits = 1/pairwise-distance (variant 0 gesture: 1/(D+1e-8) clamped at 1e-8; variant 1 expressive: 1/D),
// cross-entropy against the diagonal -- tiled with an online log-sum-exp, the N x N x 32 tensor is never built.
// a, b: [N,32].  Outputs: an, bn [N,32], na, nb [N] (normalised rows + norms), lse [N]; loss ACCUMULATED.
// scratch: part [N * ceil(N/128) * 2] floats (upper bound), diag [N].
HA2G_API int ha2g_contrastive_fwd(const float* a, const float* b, float* an, float* bn, float* na, float* nb, float* lse,
                                  float* part, float* diag, int N, int variant, float* loss, cudaStream_t stream) {
    return ha2g_contrastive_fwd_rect(a, b, an, bn, na, nb, lse, part, diag, N, N, 0, variant, loss, stream);
}
// Contrastive backward: da, db [N,32] = gscale * d loss/d a, d loss/d b.  dan, dbn: zero-initialised scratch [N,32].
HA2G_API int ha2g_contrastive_bwd(const float* an, const float* bn, const float* na, const float* nb, const float* lse,
                                  const float* gscale, float* dan, float* dbn, float* da, float* db, int N, int variant,
                                  cudaStream_t stream) {
    return ha2g_contrastive_bwd_rect(an, bn, na, nb, lse, gscale, dan, dbn, da, db, N, N, 0, variant, stream);
}

namespace {
// FGD evaluator side metrics (scripts/model/embedding_space_evaluator.py:75-99), one CTA per sample b:
//   rec[b] = mean_{t,d} |recon - pose| + mean_{t<T-1,d} |(recon[t+1]-recon[t]) - (pose[t+1]-pose[t])|
//   cs[b]  = sum_{t,j} (1 - cos(recon[t,3j:3j+3], pose[t,3j:3j+3]))      (torch.cosine_similarity: eps 1e-8 on each norm)
__global__ void recon_metrics_kernel(const float* __restrict__ recon, const float* __restrict__ pose, int T, int D,
                                     float* __restrict__ rec, float* __restrict__ cs) {
    __shared__ float sh[33];
    const int b = blockIdx.x;
    const float* r = recon + (size_t)b * T * D;
    const float* p = pose + (size_t)b * T * D;
    float a = 0.f, d = 0.f, c = 0.f;
    for (int e = threadIdx.x; e < T * D; e += blockDim.x) {
        a += fabsf(r[e] - p[e]);
        if (e < (T - 1) * D) d += fabsf((r[e + D] - r[e]) - (p[e + D] - p[e]));
    }
    const int J = D / 3;
    for (int e = threadIdx.x; e < T * J; e += blockDim.x) {
        const float* rv = r + (size_t)e * 3;
        const float* pv = p + (size_t)e * 3;
        const float dot = rv[0] * pv[0] + rv[1] * pv[1] + rv[2] * pv[2];
        const float nr = fmaxf(sqrtf(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]), 1e-8f);
        const float np = fmaxf(sqrtf(pv[0] * pv[0] + pv[1] * pv[1] + pv[2] * pv[2]), 1e-8f);
        c += 1.f - dot / (nr * np);
    }
    a = block_sum(a, sh);
    d = block_sum(d, sh);
    c = block_sum(c, sh);
    if (threadIdx.x == 0) {
        rec[b] = a / (float)(T * D) + d / (float)((T - 1) * D);
        cs[b] = c;
    }
}
}  // namespace

// Per-sample reconstruction / cosine errors of the FGD auto-encoder's output against its input: recon, pose [B,T,D]
// (D = 3 * joints) -> rec [B], cs [B]   (EmbeddingSpaceEvaluator.push_samples, embedding_space_evaluator.py:75-99)
HA2G_API int ha2g_recon_metrics(const float* recon, const float* pose, int B, int T, int D, float* rec, float* cs,
                                cudaStream_t stream) {
    if (B <= 0) return 0;
    if (D % 3 != 0 || T < 2) return (int)cudaErrorInvalidValue;
    recon_metrics_kernel<<<B, 256, 0, stream>>>(recon, pose, T, D, rec, cs);
    HA2G_RETURN_LAST();
}
