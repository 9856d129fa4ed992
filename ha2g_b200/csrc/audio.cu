// Audio-encoder kernels that are not convolutions (K4, K5, K6 of SURVEY.md), all NHWC / channels-last:
//   * squeeze-excite block tail  (ResNetBlocks.py:29-37,81-96): GAP -> FC -> ReLU -> FC -> sigmoid ->
//     channel scale -> + residual -> ReLU, forward and backward
//   * PixelShuffle (ResNetSE34V2.py:165-166,177-178) and the head flatten
//     (B,C,F,T).reshape(B,C*F,T).transpose(1,2) (ResNetSE34V2.py:160-162) as index remaps
//   * speaker-conditioned softmax blend of the three feature levels (ResNetSE34V2.py:202-212)
#include "common.cuh"

namespace {

// ---- SE -------------------------------------------------------------------------------------------
// gap[n,c] = mean_hw u[n,hw,c].   grid (C/32, N), block (32, 8)
__global__ void se_gap_kernel(const float* __restrict__ u, float* __restrict__ gap, int HW, int C) {
    __shared__ float sh[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int n = blockIdx.y;
    float s = 0.f;
    if (c < C) {
        const float* base = u + (size_t)n * HW * C + c;
        for (int p = threadIdx.y; p < HW; p += 8) s += base[(size_t)p * C];
    }
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
        gap[(size_t)n * C + c] = t / (float)HW;
    }
}

// per sample: h = relu(W1 gap + b1) [R];  s = sigmoid(W2 h + b2) [C].   grid N, block C (<= 256)
__global__ void se_fc_fwd_kernel(const float* __restrict__ gap, const float* __restrict__ w1, const float* __restrict__ b1,
                                 const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ hbuf,
                                 float* __restrict__ sbuf, int C, int R) {
    __shared__ float g[256];
    __shared__ float h[32];
    const int n = blockIdx.x, t = threadIdx.x;
    if (t < C) g[t] = gap[(size_t)n * C + t];
    __syncthreads();
    if (t < R) {
        float a = b1[t];
        for (int c = 0; c < C; ++c) a = fmaf(w1[t * C + c], g[c], a);
        a = fmaxf(a, 0.f);
        h[t] = a;
        hbuf[(size_t)n * R + t] = a;
    }
    __syncthreads();
    if (t < C) {
        float a = b2[t];
        for (int r = 0; r < R; ++r) a = fmaf(w2[t * R + r], h[r], a);
        sbuf[(size_t)n * C + t] = ha2g_sigmoid(a);
    }
}

// out = relu(u * s[n,c] + res)
__global__ void se_apply_kernel(const float* __restrict__ u, const float* __restrict__ s, const float* __restrict__ res,
                                float* __restrict__ out, int64_t total, int HW, int C) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        int64_t n = i / ((int64_t)HW * C);
        out[i] = fmaxf(fmaf(u[i], s[n * C + c], res[i]), 0.f);
    }
}

// g = dout * (out > 0);  dres = g;  ds[n,c] = sum_hw g * u.     grid (C/32, N), block (32, 8)
__global__ void se_bwd_a_kernel(const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ u,
                                float* __restrict__ dres, float* __restrict__ ds, int HW, int C) {
    __shared__ float sh[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int n = blockIdx.y;
    float s = 0.f;
    if (c < C) {
        const size_t base = (size_t)n * HW * C + c;
        for (int p = threadIdx.y; p < HW; p += 8) {
            size_t i = base + (size_t)p * C;
            float g = out[i] > 0.f ? dout[i] : 0.f;
            dres[i] = g;
            s = fmaf(g, u[i], s);
        }
    }
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
        ds[(size_t)n * C + c] = t;
    }
}

// per sample FC backward: dgap, plus the pre-activation gradients dz2 [N,C] / dz1 [N,R] for the weight-gradient pass
// (se_fc_wgrad_kernel sums them over the samples in index order: no atomics).  grid N, block C
__global__ void se_fc_bwd_kernel(const float* __restrict__ ds, const float* __restrict__ sbuf, const float* __restrict__ hbuf,
                                 const float* __restrict__ gap, const float* __restrict__ w1, const float* __restrict__ w2,
                                 float* __restrict__ dz2_out, float* __restrict__ dz1_out, float* __restrict__ dgap, int C,
                                 int R) {
    __shared__ float dz2[256];
    __shared__ float dz1[32];
    __shared__ float h[32];
    __shared__ float g[256];
    const int n = blockIdx.x, t = threadIdx.x;
    if (t < R) h[t] = hbuf[(size_t)n * R + t];
    if (t < C) {
        float s = sbuf[(size_t)n * C + t];
        dz2[t] = ds[(size_t)n * C + t] * s * (1.f - s);
        g[t] = gap[(size_t)n * C + t];
    }
    __syncthreads();
    if (t < C) dz2_out[(size_t)n * C + t] = dz2[t];
    if (t < R) {
        float a = 0.f;
        for (int c = 0; c < C; ++c) a = fmaf(w2[c * R + t], dz2[c], a);
        a = h[t] > 0.f ? a : 0.f;
        dz1[t] = a;
        dz1_out[(size_t)n * R + t] = a;
    }
    __syncthreads();
    if (t < C) {
        float a = 0.f;
        for (int r = 0; r < R; ++r) a = fmaf(w1[r * C + t], dz1[r], a);
        dgap[(size_t)n * C + t] = a;
    }
}
// dw2[c][r] += sum_n dz2[n,c] h[n,r];  db2[c] += sum_n dz2[n,c];  dw1[r][c] += sum_n dz1[n,r] gap[n,c];  db1[r] += sum_n dz1[n,r]
// One CTA per channel c; thread (r, lane) with r in [0, R] (r == R: the bias columns) sums the samples n = lane, lane+8, ...
// and the 8 lanes are combined in lane order: deterministic, 8 x fewer serial iterations than one thread per (c, r).
__global__ void se_fc_wgrad_kernel(const float* __restrict__ dz2, const float* __restrict__ dz1, const float* __restrict__ hbuf,
                                   const float* __restrict__ gap, float* __restrict__ dw1, float* __restrict__ db1,
                                   float* __restrict__ dw2, float* __restrict__ db2, int N, int C, int R) {
    __shared__ float sa[8][33], sb[8][33];
    const int c = blockIdx.x, r = threadIdx.x, lane = threadIdx.y;   // blockDim = (R + 1, 8)
    float a = 0.f, b = 0.f;
    if (r == R) {
        for (int n = lane; n < N; n += 8) {
            a += dz2[(size_t)n * C + c];
            if (c < R) b += dz1[(size_t)n * R + c];
        }
    } else {
        for (int n = lane; n < N; n += 8) {
            a = fmaf(dz2[(size_t)n * C + c], hbuf[(size_t)n * R + r], a);
            b = fmaf(dz1[(size_t)n * R + r], gap[(size_t)n * C + c], b);
        }
    }
    sa[lane][r] = a;
    sb[lane][r] = b;
    __syncthreads();
    if (lane == 0) {
        float ta = 0.f, tb = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { ta += sa[k][r]; tb += sb[k][r]; }
        if (r == R) {
            db2[c] += ta;
            if (c < R) db1[c] += tb;
        } else {
            dw2[c * R + r] += ta;
            dw1[r * C + c] += tb;
        }
    }
}
// acc[i] = scale * (part[0][i] + part[1][i] + ...) in index order
__global__ void se_combine_kernel(const float* __restrict__ part, int nparts, int n, float scale, float* __restrict__ acc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float t = 0.f;
    for (int p = 0; p < nparts; ++p) t += part[(size_t)p * n + i];
    acc[i] = t * scale;
}

// du = (dout * (out>0)) * s[n,c] + dgap[n,c] / HW
__global__ void se_bwd_b_kernel(const float* __restrict__ dres /* = dout*(out>0) */, const float* __restrict__ s,
                                const float* __restrict__ dgap, float* __restrict__ du, int64_t total, int HW, int C) {
    const float inv = 1.f / (float)HW;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        int64_t n = i / ((int64_t)HW * C);
        du[i] = fmaf(dres[i], s[n * C + c], dgap[n * C + c] * inv);
    }
}


// ---- 128-bit variants of the SE streaming kernels (C % 4 == 0, 16-byte aligned maps) -------------------------------------
// Reductions (GAP, ds): a CTA owns a pixel range of one sample; thread = (pixel slot, channel quad) streams float4s, the
// block reduces over its pixel slots in shared memory and stores ONE partial per channel into its own plane
// acc[split][n][c] (se_combine_kernel adds the planes in order) -- so a 128 x 70 x 32 map is spread over N x splits CTAs
// instead of N, every access is 16 bytes, and the result does not depend on CTA scheduling.
constexpr int SE_NT = 256;

template <int MODE>   // 0: partial sums of u;   1: g = dout*(out>0) -> dres, partial sums of g*u
__global__ void __launch_bounds__(SE_NT) se_reduce_vec_kernel(const float4* __restrict__ u, const float4* __restrict__ dout,
                                                              const float4* __restrict__ out, float4* __restrict__ dres,
                                                              float* __restrict__ acc, int HW, int C, int px_per_cta) {
    __shared__ float4 sh[SE_NT];
    const int C4 = C >> 2;
    const int slots = SE_NT / C4;                 // pixels handled per block iteration (C4 divides 256: C = 8..256, 2^k)
    const int q = threadIdx.x % C4, slot = threadIdx.x / C4;
    const int n = blockIdx.y;
    const int p0 = blockIdx.x * px_per_cta, p1 = min(HW, p0 + px_per_cta);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (slot < slots) {
        const size_t base = (size_t)n * HW * C4 + q;
        for (int p = p0 + slot; p < p1; p += slots) {
            const size_t i = base + (size_t)p * C4;
            const float4 uv = u[i];
            if (MODE == 0) {
                a.x += uv.x; a.y += uv.y; a.z += uv.z; a.w += uv.w;
            } else {
                const float4 o = out[i], d = dout[i];
                float4 g;
                g.x = o.x > 0.f ? d.x : 0.f; g.y = o.y > 0.f ? d.y : 0.f; g.z = o.z > 0.f ? d.z : 0.f; g.w = o.w > 0.f ? d.w : 0.f;
                dres[i] = g;
                a.x = fmaf(g.x, uv.x, a.x); a.y = fmaf(g.y, uv.y, a.y); a.z = fmaf(g.z, uv.z, a.z); a.w = fmaf(g.w, uv.w, a.w);
            }
        }
    }
    sh[threadIdx.x] = a;
    __syncthreads();
    if (slot == 0) {
        for (int k = 1; k < slots; ++k) {
            const float4 b = sh[k * C4 + q];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        *reinterpret_cast<float4*>(acc + ((size_t)blockIdx.x * gridDim.y + n) * C + q * 4) = a;
    }
}

template <int MODE>   // 0: out = relu(u*s + res);   1: du = dres*s + dgap/HW
__global__ void __launch_bounds__(256) se_stream_vec_kernel(const float4* __restrict__ a, const float* __restrict__ s,
                                                            const float4* __restrict__ b, const float* __restrict__ dgap,
                                                            float4* __restrict__ o, int64_t total4, int HW, int C) {
    const int C4 = C >> 2;
    const int64_t per_n = (int64_t)HW * C4;
    const float inv = 1.f / (float)HW;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) << 2;
        const int64_t n = i / per_n;
        const float4 sv = *reinterpret_cast<const float4*>(s + n * C + c);
        const float4 av = a[i];
        float4 r;
        if (MODE == 0) {
            const float4 bv = b[i];
            r.x = fmaxf(fmaf(av.x, sv.x, bv.x), 0.f); r.y = fmaxf(fmaf(av.y, sv.y, bv.y), 0.f);
            r.z = fmaxf(fmaf(av.z, sv.z, bv.z), 0.f); r.w = fmaxf(fmaf(av.w, sv.w, bv.w), 0.f);
        } else {
            const float4 gv = *reinterpret_cast<const float4*>(dgap + n * C + c);
            r.x = fmaf(av.x, sv.x, gv.x * inv); r.y = fmaf(av.y, sv.y, gv.y * inv);
            r.z = fmaf(av.z, sv.z, gv.z * inv); r.w = fmaf(av.w, sv.w, gv.w * inv);
        }
        o[i] = r;
    }
}

static inline bool se_vec_ok(int C, const void* a, const void* b, const void* c, const void* d) {
    return C >= 8 && C <= 256 && (C & (C - 1)) == 0 &&
           (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d) & 15) == 0;
}
static inline void se_reduce_grid(int N, int HW, int& splits, int& px) {
    splits = (148 * 4 + N - 1) / N;
    if (splits > (HW + 63) / 64) splits = (HW + 63) / 64;
    if (splits < 1) splits = 1;
    px = (HW + splits - 1) / splits;
    splits = (HW + px - 1) / px;
}

// ---- index remaps -------------------------------------------------------------------------------------
// PixelShuffle(r), NHWC:  out[n, h*r+i, w*r+j, c] = in[n, h, w, c*r*r + i*r + j].   inverse: swap roles.
__global__ void pixel_shuffle_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t N, int H, int W, int Cout,
                                     int r, int inverse) {
    const int Ho = H * r, Wo = W * r;
    const int64_t total = N * Ho * Wo * Cout;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(e % Cout);
        int64_t p = e / Cout;
        int wo = (int)(p % Wo), ho = (int)((p / Wo) % Ho);
        int64_t n = p / ((int64_t)Wo * Ho);
        int h = ho / r, i = ho % r, w = wo / r, j = wo % r;
        int64_t in_idx = ((n * H + h) * W + w) * ((int64_t)Cout * r * r) + (int64_t)c * r * r + i * r + j;
        if (inverse) dst[in_idx] = src[e]; else dst[e] = src[in_idx];
    }
}
// head flatten: dst[n][t][c*F + f] = src[n][f][t][c]   (src NHWC with H=F, W=T).  inverse: swap roles.
__global__ void head_flatten_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t N, int F, int T, int C,
                                    int inverse) {
    const int64_t total = N * F * T * C;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        // e indexes the flattened side [n][t][c*F+f]
        int f = (int)(e % F);
        int c = (int)((e / F) % C);
        int t = (int)((e / ((int64_t)F * C)) % T);
        int64_t n = e / ((int64_t)F * C * T);
        int64_t s = ((n * F + f) * T + t) * C + c;
        if (inverse) dst[s] = src[e]; else dst[e] = src[s];
    }
}

// ---- speaker blend ------------------------------------------------------------------------------------
// logits [B,3,L] -> weight = softmax over dim 1;  blend[i][b,t,:] = sum_k weight[b,k,i] * feat_k[b,t,:]
__global__ void blend_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ f0, const float* __restrict__ f1,
                                 const float* __restrict__ f2, float* __restrict__ weight, float* __restrict__ blend, int B,
                                 int TC, int L) {
    __shared__ float w[3 * 8];
    const int b = blockIdx.x;
    if (threadIdx.x < L) {
        int i = threadIdx.x;
        float a0 = logits[(b * 3 + 0) * L + i], a1 = logits[(b * 3 + 1) * L + i], a2 = logits[(b * 3 + 2) * L + i];
        float m = fmaxf(a0, fmaxf(a1, a2));
        float e0 = expf(a0 - m), e1 = expf(a1 - m), e2 = expf(a2 - m);
        float inv = 1.f / (e0 + e1 + e2);
        w[0 * L + i] = e0 * inv; w[1 * L + i] = e1 * inv; w[2 * L + i] = e2 * inv;
        weight[(b * 3 + 0) * L + i] = e0 * inv;
        weight[(b * 3 + 1) * L + i] = e1 * inv;
        weight[(b * 3 + 2) * L + i] = e2 * inv;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < TC; e += blockDim.x) {
        size_t idx = (size_t)b * TC + e;
        float x0 = f0[idx], x1 = f1[idx], x2 = f2[idx];
        for (int i = 0; i < L; ++i)
            blend[((size_t)i * B + b) * TC + e] = x0 * w[i] + x1 * w[L + i] + x2 * w[2 * L + i];
    }
}
// dfeat_k[b,t,:] = sum_i weight[b,k,i]*dblend[i][b,t,:];  dlogits via softmax backward with
// dW[b,k,i] = dweight[b,k,i] + sum_{t,c} dblend[i][b,t,c]*feat_k[b,t,c]
__global__ void blend_bwd_kernel(const float* __restrict__ weight, const float* __restrict__ f0, const float* __restrict__ f1,
                                 const float* __restrict__ f2, const float* __restrict__ dweight /* may be null */,
                                 const float* __restrict__ dblend, float* __restrict__ df0, float* __restrict__ df1,
                                 float* __restrict__ df2, float* __restrict__ dlogits, int B, int TC, int L) {
    __shared__ float w[24];
    __shared__ float dwacc[24];
    __shared__ float sh[33];
    const int b = blockIdx.x;
    if (threadIdx.x < 3 * L) w[threadIdx.x] = weight[b * 3 * L + threadIdx.x];
    __syncthreads();
    float part[24];
    for (int q = 0; q < 3 * L; ++q) part[q] = 0.f;
    for (int e = threadIdx.x; e < TC; e += blockDim.x) {
        size_t idx = (size_t)b * TC + e;
        float x0 = f0[idx], x1 = f1[idx], x2 = f2[idx];
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int i = 0; i < L; ++i) {
            float d = dblend[((size_t)i * B + b) * TC + e];
            a0 = fmaf(w[i], d, a0); a1 = fmaf(w[L + i], d, a1); a2 = fmaf(w[2 * L + i], d, a2);
            part[i] = fmaf(d, x0, part[i]); part[L + i] = fmaf(d, x1, part[L + i]); part[2 * L + i] = fmaf(d, x2, part[2 * L + i]);
        }
        df0[idx] = a0; df1[idx] = a1; df2[idx] = a2;
    }
    for (int q = 0; q < 3 * L; ++q) {
        float t = block_sum(part[q], sh);
        if (threadIdx.x == 0) dwacc[q] = t + (dweight != nullptr ? dweight[b * 3 * L + q] : 0.f);
    }
    __syncthreads();
    if (threadIdx.x < L) {
        int i = threadIdx.x;
        float dot = w[i] * dwacc[i] + w[L + i] * dwacc[L + i] + w[2 * L + i] * dwacc[2 * L + i];
        for (int k = 0; k < 3; ++k) dlogits[(b * 3 + k) * L + i] = w[k * L + i] * (dwacc[k * L + i] - dot);
    }
}

}  // namespace

// SE tail forward.  u, res, out: [N,HW,C];  gap, s: [N,C];  h: [N,R]  (R = C/8);  C <= 256, R <= 32.
HA2G_API int ha2g_se_fwd(const float* u, const float* res, const float* w1, const float* b1, const float* w2,
                         const float* b2, float* gap, float* h, float* s, float* out, int N, int HW, int C, int R,
                         cudaStream_t stream) {
    if (C > 256 || R > 32) return (int)cudaErrorInvalidValue;
    const bool vec = se_vec_ok(C, u, res, out, s);
    if (vec) {
        int splits, px;
        se_reduce_grid(N, HW, splits, px);
        float* part = reinterpret_cast<float*>(ha2g_ws((size_t)splits * N * C * sizeof(float), stream));
        if (part == nullptr) return (int)cudaErrorMemoryAllocation;
        se_reduce_vec_kernel<0><<<dim3(splits, N), SE_NT, 0, stream>>>(reinterpret_cast<const float4*>(u), nullptr, nullptr, nullptr,
                                                                      part, HW, C, px);
        se_combine_kernel<<<ha2g_div_up(N * C, 256), 256, 0, stream>>>(part, splits, N * C, 1.f / (float)HW, gap);
    } else {
        se_gap_kernel<<<dim3(ha2g_div_up(C, 32), N), dim3(32, 8), 0, stream>>>(u, gap, HW, C);
    }
    se_fc_fwd_kernel<<<N, 256, 0, stream>>>(gap, w1, b1, w2, b2, h, s, C, R);
    int64_t total = (int64_t)N * HW * C;
    if (vec)
        se_stream_vec_kernel<0><<<ha2g_ew_grid(total / 4, 256, 4), 256, 0, stream>>>(reinterpret_cast<const float4*>(u), s,
                                                                                    reinterpret_cast<const float4*>(res), nullptr,
                                                                                    reinterpret_cast<float4*>(out), total / 4, HW, C);
    else
        se_apply_kernel<<<ha2g_ew_grid(total), 256, 0, stream>>>(u, s, res, out, total, HW, C);
    HA2G_RETURN_LAST();
}
// SE tail backward.  dres (= dout*(out>0)) and du are outputs [N,HW,C]; dw1,db1,dw2,db2 ACCUMULATED; ds,dgap scratch [N,C].
HA2G_API int ha2g_se_bwd(const float* dout, const float* out, const float* u, const float* gap, const float* h,
                         const float* s, const float* w1, const float* w2, float* dres, float* du, float* ds, float* dgap,
                         float* dw1, float* db1, float* dw2, float* db2, int N, int HW, int C, int R,
                         cudaStream_t stream) {
    if (C > 256 || R > 32) return (int)cudaErrorInvalidValue;
    const bool vec = se_vec_ok(C, dout, out, u, dres) && se_vec_ok(C, du, s, dgap, du);
    if (vec) {
        int splits, px;
        se_reduce_grid(N, HW, splits, px);
        float* part = reinterpret_cast<float*>(ha2g_ws((size_t)splits * N * C * sizeof(float), stream));
        if (part == nullptr) return (int)cudaErrorMemoryAllocation;
        se_reduce_vec_kernel<1><<<dim3(splits, N), SE_NT, 0, stream>>>(reinterpret_cast<const float4*>(u),
                                                                      reinterpret_cast<const float4*>(dout),
                                                                      reinterpret_cast<const float4*>(out),
                                                                      reinterpret_cast<float4*>(dres), part, HW, C, px);
        se_combine_kernel<<<ha2g_div_up(N * C, 256), 256, 0, stream>>>(part, splits, N * C, 1.f, ds);
    } else {
        se_bwd_a_kernel<<<dim3(ha2g_div_up(C, 32), N), dim3(32, 8), 0, stream>>>(dout, out, u, dres, ds, HW, C);
    }
    float* dz2 = reinterpret_cast<float*>(ha2g_ws((size_t)N * (C + R) * sizeof(float), stream));
    if (dz2 == nullptr) return (int)cudaErrorMemoryAllocation;
    float* dz1 = dz2 + (size_t)N * C;
    se_fc_bwd_kernel<<<N, 256, 0, stream>>>(ds, s, h, gap, w1, w2, dz2, dz1, dgap, C, R);
    se_fc_wgrad_kernel<<<C, dim3(R + 1, 8), 0, stream>>>(dz2, dz1, h, gap, dw1, db1, dw2, db2, N, C, R);
    int64_t total = (int64_t)N * HW * C;
    if (vec)
        se_stream_vec_kernel<1><<<ha2g_ew_grid(total / 4, 256, 4), 256, 0, stream>>>(reinterpret_cast<const float4*>(dres), s, nullptr,
                                                                                    dgap, reinterpret_cast<float4*>(du), total / 4, HW, C);
    else
        se_bwd_b_kernel<<<ha2g_ew_grid(total), 256, 0, stream>>>(dres, s, dgap, du, total, HW, C);
    HA2G_RETURN_LAST();
}
// nn.PixelShuffle(r) on NHWC: src [N,H,W,Cout*r*r] -> dst [N,H*r,W*r,Cout]  (inverse != 0: the other way)
HA2G_API int ha2g_pixel_shuffle(const float* src, float* dst, int64_t N, int H, int W, int Cout, int r, int inverse,
                                cudaStream_t stream) {
    int64_t total = N * H * r * W * r * Cout;
    pixel_shuffle_kernel<<<ha2g_ew_grid(total), 256, 0, stream>>>(src, dst, N, H, W, Cout, r, inverse);
    HA2G_RETURN_LAST();
}
// src [N,F,T,C] (NHWC) -> dst [N,T,C*F]  (feature index c*F+f, ResNetSE34V2.py:160-162); inverse != 0: back
HA2G_API int ha2g_head_flatten(const float* src, float* dst, int64_t N, int F, int T, int C, int inverse,
                               cudaStream_t stream) {
    int64_t total = N * F * T * C;
    head_flatten_kernel<<<ha2g_ew_grid(total), 256, 0, stream>>>(src, dst, N, F, T, C, inverse);
    HA2G_RETURN_LAST();
}
// speaker blend forward: logits [B,3,L], f* [B,TC] -> weight [B,3,L], blend [L,B,TC]   (L <= 8)
HA2G_API int ha2g_blend_fwd(const float* logits, const float* f0, const float* f1, const float* f2, float* weight,
                            float* blend, int B, int TC, int L, cudaStream_t stream) {
    if (L > 8) return (int)cudaErrorInvalidValue;
    blend_fwd_kernel<<<B, 256, 0, stream>>>(logits, f0, f1, f2, weight, blend, B, TC, L);
    HA2G_RETURN_LAST();
}
HA2G_API int ha2g_blend_bwd(const float* weight, const float* f0, const float* f1, const float* f2, const float* dweight,
                            const float* dblend, float* df0, float* df1, float* df2, float* dlogits, int B, int TC, int L,
                            cudaStream_t stream) {
    if (L > 8) return (int)cudaErrorInvalidValue;
    blend_bwd_kernel<<<B, 256, 0, stream>>>(weight, f0, f1, f2, dweight, dblend, df0, df1, df2, dlogits, B, TC, L);
    HA2G_RETURN_LAST();
}

// ---- stride-2 convolutions as stride-1 convolutions (so that they run on the tcgen05 conv kernels) -------------------
namespace {
// y[n][i][j][(a*2+b)*C + c] = x[n][2i+a][2j+b][c] (0 outside);  inverse: x[n][h][w][c] = y[n][h/2][w/2][((h&1)*2+(w&1))*C + c]
__global__ void space_to_depth2_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int H, int W, int C,
                                       int H2, int W2, int inverse) {
    const int C4 = C >> 2;
    if (!inverse) {
        const int64_t total = (int64_t)N * H2 * W2 * 4 * C4;
        for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
            const int c4 = (int)(e % C4);
            int64_t t = e / C4;
            const int ab = (int)(t % 4); t /= 4;
            const int j = (int)(t % W2); t /= W2;
            const int i = (int)(t % H2);
            const int n = (int)(t / H2);
            const int h = 2 * i + (ab >> 1), w = 2 * j + (ab & 1);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (h < H && w < W) v = reinterpret_cast<const float4*>(src)[(((int64_t)n * H + h) * W + w) * C4 + c4];
            reinterpret_cast<float4*>(dst)[e] = v;
        }
    } else {
        const int64_t total = (int64_t)N * H * W * C4;
        for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
            const int c4 = (int)(e % C4);
            int64_t t = e / C4;
            const int w = (int)(t % W); t /= W;
            const int h = (int)(t % H);
            const int n = (int)(t / H);
            const int ab = ((h & 1) << 1) | (w & 1);
            reinterpret_cast<float4*>(dst)[e] =
                reinterpret_cast<const float4*>(src)[((((int64_t)n * H2 + (h >> 1)) * W2 + (w >> 1)) * 4 + ab) * C4 + c4];
        }
    }
}
// 3x3 stride-2 pad-1 weight w[co][c][r][s] <-> 2x2 stride-1 pad-1 weight over the space-to-depth input
// w2[co][(a*2+b)*C + c][p][q], with r = 2p + a - 1, s = 2q + b - 1 (combinations with r or s = -1 are structural zeros).
__global__ void conv_s2_weight_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout, int C, int inverse) {
    if (!inverse) {
        const int64_t total = (int64_t)Cout * 4 * C * 4;
        for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
            const int pq = (int)(e % 4);
            int64_t t = e / 4;
            const int c = (int)(t % C); t /= C;
            const int ab = (int)(t % 4);
            const int co = (int)(t / 4);
            const int r = 2 * (pq >> 1) + (ab >> 1) - 1, s = 2 * (pq & 1) + (ab & 1) - 1;
            dst[e] = (r >= 0 && s >= 0) ? src[(((int64_t)co * C + c) * 3 + r) * 3 + s] : 0.f;
        }
    } else {
        const int64_t total = (int64_t)Cout * C * 9;
        for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
            const int s = (int)(e % 3);
            int64_t t = e / 3;
            const int r = (int)(t % 3); t /= 3;
            const int c = (int)(t % C);
            const int co = (int)(t / C);
            const int p = (r + 1) >> 1, a = (r + 1) & 1, q = (s + 1) >> 1, b = (s + 1) & 1;
            dst[e] = src[((((int64_t)co * 4 + (a * 2 + b)) * C + c) * 2 + p) * 2 + q];
        }
    }
}
// y[n][i][j][:] = x[n][2i][2j][:];  inverse: x = 0 except x[n][2i][2j][:] = y[n][i][j][:]
__global__ void subsample2_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int H, int W, int C, int H2,
                                  int W2, int inverse) {
    const int C4 = C >> 2;
    if (!inverse) {
        const int64_t total = (int64_t)N * H2 * W2 * C4;
        for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
            const int c4 = (int)(e % C4);
            int64_t t = e / C4;
            const int j = (int)(t % W2); t /= W2;
            const int i = (int)(t % H2);
            const int n = (int)(t / H2);
            reinterpret_cast<float4*>(dst)[e] = reinterpret_cast<const float4*>(src)[(((int64_t)n * H + 2 * i) * W + 2 * j) * C4 + c4];
        }
    } else {
        const int64_t total = (int64_t)N * H * W * C4;
        for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
            const int c4 = (int)(e % C4);
            int64_t t = e / C4;
            const int w = (int)(t % W); t /= W;
            const int h = (int)(t % H);
            const int n = (int)(t / H);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!(h & 1) && !(w & 1)) v = reinterpret_cast<const float4*>(src)[(((int64_t)n * H2 + (h >> 1)) * W2 + (w >> 1)) * C4 + c4];
            reinterpret_cast<float4*>(dst)[e] = v;
        }
    }
}
}  // namespace

// Space-to-depth by 2 of an NHWC tensor (C % 4 == 0): [N,H,W,C] -> [N,ceil(H/2),ceil(W/2),4C] (inverse != 0: the adjoint,
// which is the exact inverse on the un-padded positions).  With ha2g_conv_s2_weight it turns the stride-2 3x3 convolutions
// of ResNetSE-34 (ResNetBlocks.py:12 with stride 2, ResNetSE34V2.py:97-99) into stride-1 2x2 convolutions for conv_tc.cu.
HA2G_API int ha2g_space_to_depth2(const float* src, float* dst, int N, int H, int W, int C, int inverse, cudaStream_t stream) {
    if (C % 4 != 0) return (int)cudaErrorInvalidValue;
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
    const int64_t total = inverse ? (int64_t)N * H * W * (C / 4) : (int64_t)N * H2 * W2 * C;
    space_to_depth2_kernel<<<ha2g_ew_grid(total, 256, 2), 256, 0, stream>>>(src, dst, N, H, W, C, H2, W2, inverse);
    HA2G_RETURN_LAST();
}
// w [Cout,C,3,3] -> w2 [Cout,4C,2,2] (inverse != 0: gradient of w from the gradient of w2).
HA2G_API int ha2g_conv_s2_weight(const float* src, float* dst, int Cout, int C, int inverse, cudaStream_t stream) {
    const int64_t total = inverse ? (int64_t)Cout * C * 9 : (int64_t)Cout * C * 16;
    conv_s2_weight_kernel<<<ha2g_ew_grid(total, 256, 2), 256, 0, stream>>>(src, dst, Cout, C, inverse);
    HA2G_RETURN_LAST();
}
// Every second pixel of an NHWC tensor (the 1x1 stride-2 downsample convolutions, ResNetSE34V2.py:70-75): [N,H,W,C] ->
// [N,ceil(H/2),ceil(W/2),C]; inverse != 0: the adjoint (zeros at the skipped pixels).
HA2G_API int ha2g_subsample2(const float* src, float* dst, int N, int H, int W, int C, int inverse, cudaStream_t stream) {
    if (C % 4 != 0) return (int)cudaErrorInvalidValue;
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
    const int64_t total = inverse ? (int64_t)N * H * W * (C / 4) : (int64_t)N * H2 * W2 * (C / 4);
    subsample2_kernel<<<ha2g_ew_grid(total, 256, 2), 256, 0, stream>>>(src, dst, N, H, W, C, H2, W2, inverse);
    HA2G_RETURN_LAST();
}
