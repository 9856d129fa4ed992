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
// SoftmaxContrastiveLoss on the tensor cores.  Rows are L2-normalised, so the pairwise distances come from ONE Gram
// matrix  G = an bn^T  (K = 32: the packed tcgen05 GEMM of gemm_tc2.cu, bf16x3):  D_ij^2 = |an_i|^2 + |bn_j|^2 - 2 G_ij.
// Where that difference cancels (D^2 < 0.25: close pairs, e.g. trained positives) the entry is recomputed from the direct
// differences, so the logits 1/D keep the reference's behaviour near coincident rows.  Everything after the Gram matrix
// is row-local:   logits + online log-sum-exp (one CTA per row)  ->  loss;
// backward:       C_ij = d loss / d D_ij / D_ij  written over the logits (one CTA per row, row sums on the way),
//                 column sums (ordered two-stage reduction),  X = C bn  and  Y = C^T an  (two more tcgen05 GEMMs, K = N),
//                 da_i = rs_i an_i - X_i,  db_j = cs_j bn_j - Y_j,  then the L2-normalisation backward.
// One pass over the N x N matrix per stage instead of three 32-wide SIMT sweeps; no atomics anywhere (deterministic).
constexpr int CC = 32;      // feature width (nOut of both encoders)
constexpr float CLOSE_D2 = 0.25f;

// xn = x / max(|x|, 1e-12), norms and |xn|^2 saved.   one warp per row, lane = channel
__global__ void l2norm_rows_kernel(const float* __restrict__ x, float* __restrict__ xn, float* __restrict__ nrm,
                                   float* __restrict__ sq, int64_t N) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (row >= N) return;
    float v = x[row * CC + lane];
    float s = warp_sum(v * v);
    float n = fmaxf(sqrtf(s), 1e-12f);
    const float u = v / n;
    xn[row * CC + lane] = u;
    const float q = warp_sum(u * u);
    if (lane == 0) { nrm[row] = n; sq[row] = q; }
}

__device__ __forceinline__ float contrastive_logit(float d, int variant) {
    return variant == 0 ? fmaxf(1.0f / (d + 1e-8f), 1e-8f) : 1.0f / d;
}

// One CTA per loss row i: G[i, :] (Gram entries) -> logits (in place), lse[i], rowloss[i] = lse_i - l_{i, i+off}
__global__ void __launch_bounds__(256) contrastive_rows_kernel(float* __restrict__ G, const float* __restrict__ an,
                                                               const float* __restrict__ bn, const float* __restrict__ sqa,
                                                               const float* __restrict__ sqb, float* __restrict__ lse,
                                                               float* __restrict__ rowloss, int Nb, int off, int variant) {
    __shared__ float arow[CC];
    __shared__ float shm[8], shs[8];
    __shared__ float sdiag;
    const int i = blockIdx.x, tid = threadIdx.x;
    if (tid < CC) arow[tid] = an[(size_t)i * CC + tid];
    __syncthreads();
    const float qa = sqa[i];
    float* g = G + (size_t)i * Nb;
    float m = -CUDART_INF_F, s = 0.f;
    for (int j = tid; j < Nb; j += 256) {
        float d2 = qa + sqb[j] - 2.f * g[j];
        if (d2 < CLOSE_D2) {   // cancellation zone: direct differences, as the reference computes every pair
            const float* br = bn + (size_t)j * CC;
            d2 = 0.f;
#pragma unroll
            for (int c = 0; c < CC; ++c) { const float df = arow[c] - br[c]; d2 = fmaf(df, df, d2); }
        }
        const float l = contrastive_logit(sqrtf(fmaxf(d2, 0.f)), variant);
        g[j] = l;
        if (j == i + off) sdiag = l;
        if (l > m) { s = s * expf(m - l) + 1.f; m = l; }
        else s += expf(l - m);
    }
    // combine (max, sum-exp) over the block in a fixed order: lanes by shuffle tree, then the 8 warps in index order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, m, o), so = __shfl_xor_sync(0xffffffffu, s, o);
        const float mm = fmaxf(m, mo);
        const float sa = (m == -CUDART_INF_F) ? 0.f : s * expf(m - mm), sb = (mo == -CUDART_INF_F) ? 0.f : so * expf(mo - mm);
        // both halves of a pair must compute the identical value: order the operands by lane parity of the exchange
        const bool low = ((threadIdx.x & o) == 0);
        s = low ? sa + sb : sb + sa;
        m = mm;
    }
    if ((tid & 31) == 0) { shm[tid >> 5] = m; shs[tid >> 5] = s; }
    __syncthreads();
    if (tid == 0) {
        float mm = -CUDART_INF_F;
        for (int w = 0; w < 8; ++w) mm = fmaxf(mm, shm[w]);
        float ss = 0.f;
        for (int w = 0; w < 8; ++w) ss += (shm[w] == -CUDART_INF_F) ? 0.f : shs[w] * expf(shm[w] - mm);
        const float l = mm + logf(ss);
        lse[i] = l;
        rowloss[i] = l - sdiag;
    }
}

// loss += mean(rowloss) with a fixed summation pattern (one CTA)
__global__ void __launch_bounds__(1024) contrastive_mean_kernel(const float* __restrict__ rowloss, int N, float* __restrict__ loss) {
    __shared__ float sh[33];
    float acc = 0.f;
    for (int i = threadIdx.x; i < N; i += 1024) acc += rowloss[i];
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) *loss += acc / (float)N;
}

// One CTA per loss row i: logits -> C_ij = gs (p_ij - delta_ij) (dl/dD) / D   (in place), rs[i] = sum_j C_ij
__global__ void __launch_bounds__(256) contrastive_coef_kernel(float* __restrict__ G, const float* __restrict__ lse,
                                                               const float* __restrict__ gscale, float* __restrict__ rs,
                                                               int Na, int Nb, int off, int variant) {
    __shared__ float sh[33];
    const int i = blockIdx.x, tid = threadIdx.x;
    const float gs = (gscale != nullptr ? *gscale : 1.f) / (float)Na;
    const float li = lse[i];
    float* g = G + (size_t)i * Nb;
    float acc = 0.f;
    for (int j = tid; j < Nb; j += 256) {
        const float l = g[j];
        float p = expf(l - li);
        if (j == i + off) p -= 1.f;
        // D and dl/dD from the stored logit: expressive l = 1/D; gesture l = 1/(D + 1e-8)  (its clamp at 1e-8 is never
        // active for D <= 2).  dl/dD = -l^2 in both; coincident rows (D = 0) contribute no gradient, as before.
        const float d = variant == 0 ? 1.0f / l - 1e-8f : 1.0f / l;
        const float cf = (d > 0.f && isfinite(l)) ? gs * p * (-l * l) / d : 0.f;
        g[j] = cf;
        acc += cf;
    }
    acc = block_sum(acc, sh);
    if (tid == 0) rs[i] = acc;
}

// dn = sc[row] * xn[row] - P[row];  dx = (dn - (dn . xn) xn) / norm        (one warp per row, lane = channel)
__global__ void contrastive_finish_kernel(const float* __restrict__ sc, const float* __restrict__ xn, const float* __restrict__ P,
                                          const float* __restrict__ nrm, float* __restrict__ dx, int64_t N) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (row >= N) return;
    const float n = xn[row * CC + lane];
    const float g = sc[row] * n - P[row * CC + lane];
    const float dot = warp_sum(g * n);
    dx[row * CC + lane] = (g - dot * n) / nrm[row];
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
extern "C" int ha2g_gemm(const float*, const float*, float*, const float*, int, int, int, int, int, int, int, int, int, int, int,
                         cudaStream_t);
extern "C" int ha2g_col_sum(const float* x, int rows, int cols, int ld, float* out, cudaStream_t stream);

// Rectangular SoftmaxContrastiveLoss forward (also the data-parallel global-batch form: the reference computes the loss
// over the whole DataParallel batch, scripts/train_expressive.py:184-197 + train_hierarchy_expressive.py:244-249;
// class at train_hierarchy.py:23-68 / train_hierarchy_expressive.py:108-121): a [Na,32] are this rank's rows, b [Nb,32]
// the (all-gathered) columns, row i's positive is column i + off (off = rank * Na; square single-GPU case: Nb = Na,
// off = 0).  Logits = 1/pairwise distance of the L2-normalised rows (variant 0 gesture: 1/(D+1e-8) clamped at 1e-8;
// 1 expressive: 1/D), cross-entropy against the positives; the N x N x 32 tensor of the reference is never built.
// Outputs: an, bn (normalised rows), na, nb (norms), sqa, sqb (|an|^2, |bn|^2), lse [Na], G [Na,Nb] (the logits, kept for
// backward), rowloss [Na] scratch; loss ACCUMULATED += mean over the Na local rows.
HA2G_API int ha2g_contrastive_fwd_rect(const float* a, const float* b, float* an, float* bn, float* na, float* nb,
                                       float* sqa, float* sqb, float* lse, float* G, float* rowloss, int Na, int Nb, int off,
                                       int variant, float* loss, cudaStream_t stream) {
    if (Na <= 0 || Nb <= 0) return 0;
    l2norm_rows_kernel<<<ha2g_div_up(Na, 8), 256, 0, stream>>>(a, an, na, sqa, Na);
    l2norm_rows_kernel<<<ha2g_div_up(Nb, 8), 256, 0, stream>>>(b, bn, nb, sqb, Nb);
    int rc = ha2g_gemm(an, bn, G, nullptr, Na, Nb, CC, CC, CC, Nb, 0, 1, 0, 0, 1, stream);   // G = an bn^T
    if (rc != 0) return rc;
    contrastive_rows_kernel<<<Na, 256, 0, stream>>>(G, an, bn, sqa, sqb, lse, rowloss, Nb, off, variant);
    contrastive_mean_kernel<<<1, 1024, 0, stream>>>(rowloss, Na, loss);
    HA2G_RETURN_LAST();
}
// Backward of the rectangular loss: da [Na,32], db [Nb,32] (the gradient wrt ALL columns: under data parallelism the
// caller reduce-scatters it to the owning ranks) = gscale * d loss / d a, b.  G holds the forward's logits and is
// overwritten; scratch: X [Na,32], Y [Nb,32], rs [Na], cs [Nb].
HA2G_API int ha2g_contrastive_bwd_rect(const float* an, const float* bn, const float* na, const float* nb, const float* lse,
                                       const float* gscale, float* G, float* X, float* Y, float* rs, float* cs, float* da,
                                       float* db, int Na, int Nb, int off, int variant, cudaStream_t stream) {
    if (Na <= 0 || Nb <= 0) return 0;
    contrastive_coef_kernel<<<Na, 256, 0, stream>>>(G, lse, gscale, rs, Na, Nb, off, variant);
    cudaError_t ce = cudaMemsetAsync(cs, 0, sizeof(float) * (size_t)Nb, stream);
    if (ce != cudaSuccess) return (int)ce;
    int rc = ha2g_col_sum(G, Na, Nb, Nb, cs, stream);                                           // cs_j = sum_i C_ij
    if (rc == 0) rc = ha2g_gemm(G, bn, X, nullptr, Na, CC, Nb, Nb, CC, CC, 0, 0, 0, 0, 1, stream);   // X = C bn
    if (rc == 0) rc = ha2g_gemm(G, an, Y, nullptr, Nb, CC, Na, Nb, CC, CC, 1, 0, 0, 0, 1, stream);   // Y = C^T an
    if (rc != 0) return rc;
    contrastive_finish_kernel<<<ha2g_div_up(Na, 8), 256, 0, stream>>>(rs, an, X, na, da, Na);
    contrastive_finish_kernel<<<ha2g_div_up(Nb, 8), 256, 0, stream>>>(cs, bn, Y, nb, db, Nb);
    HA2G_RETURN_LAST();
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
