// NHWC implicit-GEMM 2-D convolution, exact-fp32 path (K2 of SURVEY.md): forward, data gradient and
// weight gradient for the ResNetSE-34 audio encoder (ResNetSE34V2.py:27,34-42,96-111; ResNetBlocks.py:12-14).
//
// Layouts: activations NHWC ([N,H,W,C] row-major, C innermost -> 128-bit coalesced channel loads);
// weights are repacked once per step from the checkpoint layout OIHW into
//     wf [KH*KW*Cin , Cout]   row (r*KW+s)*Cin + ci     (forward / wgrad GEMM-B layout)
//     wb [KH*KW*Cout, Cin ]   row (r*KW+s)*Cout + co    (dgrad GEMM-B layout)
// GEMM view:  fwd   : [N*Ho*Wo, KH*KW*Cin ] x wf -> y      (A gathered on the fly, never materialised)
//             dgrad : [N*H*W  , KH*KW*Cout] x wb -> dx     (A = dy gathered through the transposed map)
//             wgrad : A^T [KH*KW*Cin, N*Ho*Wo] x dy -> dwf (split over pixels, ordered reduction of partial planes)
// Requirements of the generic kernels: Cin % 16 == 0 (fwd, wgrad), Cout % 16 == 0 (dgrad).  The
// 1-channel stem has its own direct kernels.
#include "common.cuh"

namespace {

struct ConvGeom {
    int N, H, W, Cin, Ho, Wo, Cout, KH, KW, stride, pad;
};

constexpr int CBM = 128, CBK = 16, CNT = 256;

// ---- fwd / dgrad ---------------------------------------------------------------------------------
// DGRAD == false: output pixel (n,oh,ow), tap (r,s) reads x[n, oh*stride-pad+r, ow*stride-pad+s, :]
// DGRAD == true : "output" pixel is an INPUT pixel (n,ih,iw); tap (r,s) reads dy[n,(ih+pad-r)/stride,(iw+pad-s)/stride,:]
//                 when both divisions are exact and in range.
template <int BN, bool DGRAD>
__global__ void __launch_bounds__(CNT) conv2d_igemm_kernel(const float* __restrict__ src, const float* __restrict__ wpk,
                                                           const float* __restrict__ bias, float* __restrict__ dst,
                                                           ConvGeom g, int act) {
    constexpr int TM = (CBM * BN) / (CNT * 4);  // rows per thread (4 cols per thread): 8 for BN=64, 4 for BN=32
    constexpr int TXN = BN / 4;                 // threads along N
    __shared__ __align__(16) float As[2][CBK][CBM + 4];
    __shared__ __align__(16) float Bs[2][CBK][BN + 4];

    // GEMM dims in "destination" space
    const int dH = DGRAD ? g.H : g.Ho, dW = DGRAD ? g.W : g.Wo;      // destination map
    const int sH = DGRAD ? g.Ho : g.H, sW = DGRAD ? g.Wo : g.W;      // source map
    const int Cs = DGRAD ? g.Cout : g.Cin;                            // source channels (K per tap)
    const int Cd = DGRAD ? g.Cin : g.Cout;                            // destination channels (GEMM N)
    const int64_t M = (int64_t)g.N * dH * dW;
    const int K = g.KH * g.KW * Cs;

    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.y * CBM;
    const int n0 = blockIdx.x * BN;
    const int tx = tid % TXN, ty = tid / TXN;

    // A loader: 128 rows x 16 k = 512 float4; thread handles rows (tid/4) and (tid/4 + 64), k4 = (tid%4)*4
    const int lk = (tid % 4) * 4;
    int ln[2], lh[2], lw[2];
    bool lvalid[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int64_t m = m0 + tid / 4 + i * 64;
        lvalid[i] = m < M;
        int64_t mm = lvalid[i] ? m : 0;
        lw[i] = (int)(mm % dW);
        lh[i] = (int)((mm / dW) % dH);
        ln[i] = (int)(mm / ((int64_t)dW * dH));
    }
    // B loader: 16 x BN floats = 4*BN float4
    float4 ra[2], rb;
    const bool b_active = tid < 4 * BN;
    const int bk = tid / (BN / 4), bn4 = (tid % (BN / 4)) * 4;

    auto load = [&](int k0) {
        const int tap = k0 / Cs, c0 = k0 % Cs;
        const int r = tap / g.KW, s = tap % g.KW;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lvalid[i]) {
                int sh, sw;
                bool ok;
                if (!DGRAD) {
                    sh = lh[i] * g.stride - g.pad + r;
                    sw = lw[i] * g.stride - g.pad + s;
                    ok = sh >= 0 && sh < sH && sw >= 0 && sw < sW;
                } else {
                    int th = lh[i] + g.pad - r, tw = lw[i] + g.pad - s;
                    ok = th >= 0 && tw >= 0 && (th % g.stride) == 0 && (tw % g.stride) == 0;
                    sh = th / g.stride; sw = tw / g.stride;
                    ok = ok && sh < sH && sw < sW;
                }
                if (ok) v = *reinterpret_cast<const float4*>(src + (((int64_t)ln[i] * sH + sh) * sW + sw) * Cs + c0 + lk);
            }
            ra[i] = v;
        }
        if (b_active) {
            int gn = n0 + bn4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gn < Cd) v = *reinterpret_cast<const float4*>(wpk + (int64_t)(k0 + bk) * Cd + gn);  // Cd % 4 == 0
            rb = v;
        }
    };
    auto store = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int row = tid / 4 + i * 64;
            As[buf][lk + 0][row] = ra[i].x; As[buf][lk + 1][row] = ra[i].y;
            As[buf][lk + 2][row] = ra[i].z; As[buf][lk + 3][row] = ra[i].w;
        }
        if (b_active) *reinterpret_cast<float4*>(&Bs[buf][bk][bn4]) = rb;
    };

    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    int buf = 0;
    load(0);
    store(0);
    __syncthreads();
    for (int k0 = 0; k0 < K; k0 += CBK) {
        const bool has_next = k0 + CBK < K;
        if (has_next) load(k0 + CBK);
#pragma unroll
        for (int kk = 0; kk < CBK; ++kk) {
            float a[TM];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 t = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + i]);
                a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
            }
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (has_next) store(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    const int gn = n0 + tx * 4;
    if (gn < Cd) {
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias != nullptr) bv = *reinterpret_cast<const float4*>(bias + gn);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            int64_t m = m0 + ty * TM + i;
            if (m >= M) continue;
            float4 o;
            o.x = ha2g_act(acc[i][0] + bv.x, act); o.y = ha2g_act(acc[i][1] + bv.y, act);
            o.z = ha2g_act(acc[i][2] + bv.z, act); o.w = ha2g_act(acc[i][3] + bv.w, act);
            *reinterpret_cast<float4*>(dst + m * Cd + gn) = o;
        }
    }
}

// ---- wgrad ---------------------------------------------------------------------------------------
// part[split][(r,s,ci), co] = sum_{pixels in this split} x_gather[p,(r,s,ci)] * dy[p,co]   (`dwf` points at the partial
// planes, one per pixel split; ha2g_splitk_reduce adds them in order: deterministic)
template <int BN>
__global__ void __launch_bounds__(CNT) conv2d_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           float* __restrict__ dwf, ConvGeom g, int64_t pix_per_split) {
    constexpr int TM = (CBM * BN) / (CNT * 4);
    constexpr int TXN = BN / 4;
    __shared__ __align__(16) float As[2][CBK][CBM + 4];  // [pixel][kdim]
    __shared__ __align__(16) float Bs[2][CBK][BN + 4];   // [pixel][co]

    const int Kd = g.KH * g.KW * g.Cin;
    const int64_t P = (int64_t)g.N * g.Ho * g.Wo;
    const int tid = threadIdx.x;
    const int kd0 = blockIdx.y * CBM;
    const int n0 = blockIdx.x * BN;
    const int64_t p_beg = (int64_t)blockIdx.z * pix_per_split;
    const int64_t p_end = min(P, p_beg + pix_per_split);
    const int tx = tid % TXN, ty = tid / TXN;

    // A loader: 16 pixels x 128 kdim = 512 float4: thread -> pixel (tid/32) and (tid/32 + 8), kd4 = (tid%32)*4
    const int akd = kd0 + (tid % 32) * 4;
    const bool a_kvalid = akd < Kd;
    const int atap = a_kvalid ? akd / g.Cin : 0, aci = a_kvalid ? akd % g.Cin : 0;
    const int ar = atap / g.KW, as_ = atap % g.KW;
    // B loader: 16 pixels x BN co
    const bool b_active = tid < 4 * BN;
    const int bp = tid / (BN / 4), bn4 = (tid % (BN / 4)) * 4;

    float4 ra[2], rb;
    auto load = [&](int64_t p0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int64_t p = p0 + tid / 32 + i * 8;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a_kvalid && p < p_end) {
                int ow = (int)(p % g.Wo);
                int oh = (int)((p / g.Wo) % g.Ho);
                int n = (int)(p / ((int64_t)g.Wo * g.Ho));
                int ih = oh * g.stride - g.pad + ar, iw = ow * g.stride - g.pad + as_;
                if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
                    v = *reinterpret_cast<const float4*>(x + (((int64_t)n * g.H + ih) * g.W + iw) * g.Cin + aci);
            }
            ra[i] = v;
        }
        if (b_active) {
            int64_t p = p0 + bp;
            int gn = n0 + bn4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < p_end && gn < g.Cout) v = *reinterpret_cast<const float4*>(dy + p * g.Cout + gn);
            rb = v;
        }
    };
    auto store = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(&As[buf][tid / 32 + i * 8][(tid % 32) * 4]) = ra[i];
        if (b_active) *reinterpret_cast<float4*>(&Bs[buf][bp][bn4]) = rb;
    };

    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    int buf = 0;
    if (p_beg < p_end) {
        load(p_beg);
        store(0);
    }
    __syncthreads();
    for (int64_t p0 = p_beg; p0 < p_end; p0 += CBK) {
        const bool has_next = p0 + CBK < p_end;
        if (has_next) load(p0 + CBK);
#pragma unroll
        for (int kk = 0; kk < CBK; ++kk) {
            float a[TM];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 t = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + i]);
                a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
            }
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (has_next) store(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int kd = kd0 + ty * TM + i;
        if (kd >= Kd) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tx * 4 + j;
            if (gn < g.Cout) dwf[((int64_t)blockIdx.z * Kd + kd) * g.Cout + gn] = acc[i][j];
        }
    }
}

// ---- weight repack OIHW <-> wf / wb ------------------------------------------------------------------
// mode 0: w -> wf ; 1: w -> wb ; 2: dwf -> dw (assign)
__global__ void conv2d_pack_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout, int Cin, int KH,
                                   int KW, int mode) {
    const int64_t n = (int64_t)Cout * Cin * KH * KW;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        // e indexes OIHW
        int s = (int)(e % KW);
        int r = (int)((e / KW) % KH);
        int ci = (int)((e / ((int64_t)KW * KH)) % Cin);
        int co = (int)(e / ((int64_t)KW * KH * Cin));
        int64_t f = ((int64_t)(r * KW + s) * Cin + ci) * Cout + co;
        int64_t b = ((int64_t)(r * KW + s) * Cout + co) * Cin + ci;
        if (mode == 0) dst[f] = src[e];
        else if (mode == 1) dst[b] = src[e];
        else dst[e] = src[f];
    }
}

// ---- 1-input-channel stem (ResNetSE34V2.py:27): direct kernels ------------------------------------
// y[n,h,w,co] = b[co] + sum_{r,s} x[n,h+r-1,w+s-1] * w[co,0,r,s]      (3x3, pad 1, stride 1)
__global__ void stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                float* __restrict__ y, int N, int H, int W, int Cout) {
    const int64_t n = (int64_t)N * H * W * Cout;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int co = (int)(e % Cout);
        int64_t p = e / Cout;
        int ww = (int)(p % W), hh = (int)((p / W) % H);
        int64_t nn = p / ((int64_t)W * H);
        float acc = b[co];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            int ih = hh + r - 1;
            if (ih < 0 || ih >= H) continue;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                int iw = ww + s - 1;
                if (iw < 0 || iw >= W) continue;
                acc = fmaf(x[(nn * H + ih) * W + iw], w[co * 9 + r * 3 + s], acc);
            }
        }
        y[e] = acc;
    }
}
// part[cta][i][co] (i < 9: dw tap, i == 9: db) = sums over this CTA's pixel range; reduced in CTA order by
// stem_wgrad_reduce_kernel.  Cout <= 32.
__global__ void stem_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ part,
                                  int N, int H, int W, int Cout, int64_t pix_per_cta) {
    // blockDim = (32 channels, 8 pixel lanes)
    __shared__ float sh[8][10][33];
    const int co = threadIdx.x;
    const int64_t P = (int64_t)N * H * W;
    const int64_t p0 = (int64_t)blockIdx.x * pix_per_cta, p1 = min(P, p0 + pix_per_cta);
    float acc[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = 0.f;
    if (co < Cout) {
        for (int64_t p = p0 + threadIdx.y; p < p1; p += 8) {
            int ww = (int)(p % W), hh = (int)((p / W) % H);
            int64_t nn = p / ((int64_t)W * H);
            float g = dy[p * Cout + co];
            acc[9] += g;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                int ih = hh + r - 1;
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    int iw = ww + s - 1;
                    float xv = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? x[(nn * H + ih) * W + iw] : 0.f;
                    acc[r * 3 + s] = fmaf(xv, g, acc[r * 3 + s]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) sh[threadIdx.y][i][threadIdx.x] = acc[i];
    __syncthreads();
    if (threadIdx.y == 0 && co < Cout) {
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            float t = 0.f;
#pragma unroll
            for (int l = 0; l < 8; ++l) t += sh[l][i][threadIdx.x];
            part[((size_t)blockIdx.x * 10 + i) * 32 + co] = t;
        }
    }
}
// dw[co,0,r,s] += sum_cta part[cta][r*3+s][co];  db[co] += sum_cta part[cta][9][co]
__global__ void stem_wgrad_reduce_kernel(const float* __restrict__ part, int nparts, float* __restrict__ dw,
                                         float* __restrict__ db, int Cout) {
    const int co = threadIdx.x, i = blockIdx.x;   // 32 threads x 10 CTAs
    if (co >= Cout) return;
    float t = 0.f;
    for (int p = 0; p < nparts; ++p) t += part[((size_t)p * 10 + i) * 32 + co];
    if (i < 9) dw[co * 9 + i] += t;
    else db[co] += t;
}

static inline ConvGeom make_geom(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad) {
    ConvGeom g;
    g.N = N; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout; g.KH = KH; g.KW = KW; g.stride = stride; g.pad = pad;
    g.Ho = (H + 2 * pad - KH) / stride + 1;
    g.Wo = (W + 2 * pad - KW) / stride + 1;
    return g;
}

}  // namespace

// y[N,Ho,Wo,Cout] = act(conv(x[N,H,W,Cin], wf) + bias)      (nn.Conv2d forward; Cin % 16 == 0, Cout % 4 == 0)
HA2G_API int ha2g_conv2d_fwd(const float* x, const float* wf, const float* bias, float* y, int N, int H, int W, int Cin,
                             int Cout, int KH, int KW, int stride, int pad, int act, cudaStream_t stream) {
    if (Cin % 16 != 0 || Cout % 4 != 0) return (int)cudaErrorInvalidValue;
    ConvGeom g = make_geom(N, H, W, Cin, Cout, KH, KW, stride, pad);
    int64_t M = (int64_t)N * g.Ho * g.Wo;
    if (Cout > 32) {
        dim3 grid(ha2g_div_up(Cout, 64), ha2g_div_up(M, CBM));
        conv2d_igemm_kernel<64, false><<<grid, CNT, 0, stream>>>(x, wf, bias, y, g, act);
    } else {
        dim3 grid(ha2g_div_up(Cout, 32), ha2g_div_up(M, CBM));
        conv2d_igemm_kernel<32, false><<<grid, CNT, 0, stream>>>(x, wf, bias, y, g, act);
    }
    HA2G_RETURN_LAST();
}

// dx[N,H,W,Cin] = conv_transpose(dy[N,Ho,Wo,Cout], wb)       (Cout % 16 == 0, Cin % 4 == 0)
HA2G_API int ha2g_conv2d_dgrad(const float* dy, const float* wb, float* dx, int N, int H, int W, int Cin, int Cout,
                               int KH, int KW, int stride, int pad, cudaStream_t stream) {
    if (Cout % 16 != 0 || Cin % 4 != 0) return (int)cudaErrorInvalidValue;
    ConvGeom g = make_geom(N, H, W, Cin, Cout, KH, KW, stride, pad);
    int64_t M = (int64_t)N * H * W;
    if (Cin > 32) {
        dim3 grid(ha2g_div_up(Cin, 64), ha2g_div_up(M, CBM));
        conv2d_igemm_kernel<64, true><<<grid, CNT, 0, stream>>>(dy, wb, nullptr, dx, g, 0);
    } else {
        dim3 grid(ha2g_div_up(Cin, 32), ha2g_div_up(M, CBM));
        conv2d_igemm_kernel<32, true><<<grid, CNT, 0, stream>>>(dy, wb, nullptr, dx, g, 0);
    }
    HA2G_RETURN_LAST();
}

// dwf[KH*KW*Cin, Cout] += x^T_gathered * dy     (dwf must be initialised; split over pixels; partial planes reduced in order)
HA2G_API int ha2g_conv2d_wgrad(const float* x, const float* dy, float* dwf, int N, int H, int W, int Cin, int Cout,
                               int KH, int KW, int stride, int pad, cudaStream_t stream) {
    if (Cin % 16 != 0 || Cout % 4 != 0) return (int)cudaErrorInvalidValue;
    ConvGeom g = make_geom(N, H, W, Cin, Cout, KH, KW, stride, pad);
    int64_t P = (int64_t)N * g.Ho * g.Wo;
    int Kd = KH * KW * Cin;
    int bn = Cout > 32 ? 64 : 32;
    int tiles = ha2g_div_up(Cout, bn) * ha2g_div_up(Kd, CBM);
    int splits = ha2g_div_up(148 * 2, tiles);
    int64_t per = (P + splits - 1) / splits;
    per = (per + CBK - 1) / CBK * CBK;
    if (per < 256) per = 256;
    splits = (int)((P + per - 1) / per);
    dim3 grid(ha2g_div_up(Cout, bn), ha2g_div_up(Kd, CBM), splits);
    float* part = reinterpret_cast<float*>(ha2g_ws_top((size_t)splits * Kd * Cout * sizeof(float), stream));
    if (part == nullptr) return (int)cudaErrorMemoryAllocation;
    if (bn == 64) conv2d_wgrad_kernel<64><<<grid, CNT, 0, stream>>>(x, dy, part, g, per);
    else conv2d_wgrad_kernel<32><<<grid, CNT, 0, stream>>>(x, dy, part, g, per);
    return ha2g_splitk_reduce(part, splits, Kd, Cout, dwf, Cout, nullptr, 1, stream);
}

// OIHW checkpoint layout <-> GEMM layouts.  mode 0: w->wf, 1: w->wb, 2: dwf->dw
HA2G_API int ha2g_conv2d_pack(const float* src, float* dst, int Cout, int Cin, int KH, int KW, int mode,
                              cudaStream_t stream) {
    int64_t n = (int64_t)Cout * Cin * KH * KW;
    conv2d_pack_kernel<<<ha2g_ew_grid(n), 256, 0, stream>>>(src, dst, Cout, Cin, KH, KW, mode);
    HA2G_RETURN_LAST();
}

// 3x3 / pad 1 / 1 input channel stem: x [N,H,W], w [Cout,1,3,3] (checkpoint layout), y [N,H,W,Cout]
HA2G_API int ha2g_stem_conv_fwd(const float* x, const float* w, const float* b, float* y, int N, int H, int W, int Cout,
                                cudaStream_t stream) {
    int64_t n = (int64_t)N * H * W * Cout;
    stem_fwd_kernel<<<ha2g_ew_grid(n), 256, 0, stream>>>(x, w, b, y, N, H, W, Cout);
    HA2G_RETURN_LAST();
}
// dw [Cout,1,3,3] += ..., db [Cout] += ...   (Cout <= 32)
HA2G_API int ha2g_stem_conv_wgrad(const float* x, const float* dy, float* dw, float* db, int N, int H, int W, int Cout,
                                  cudaStream_t stream) {
    if (Cout > 32) return (int)cudaErrorInvalidValue;
    int64_t P = (int64_t)N * H * W;
    int ctas = 148 * 4;
    int64_t per = (P + ctas - 1) / ctas;
    if (per < 64) per = 64;
    ctas = (int)((P + per - 1) / per);
    float* part = reinterpret_cast<float*>(ha2g_ws_top((size_t)ctas * 10 * 32 * sizeof(float), stream));
    if (part == nullptr) return (int)cudaErrorMemoryAllocation;
    stem_wgrad_kernel<<<ctas, dim3(32, 8), 0, stream>>>(x, dy, part, N, H, W, Cout, per);
    stem_wgrad_reduce_kernel<<<10, 32, 0, stream>>>(part, ctas, dw, db, Cout);
    HA2G_RETURN_LAST();
}
