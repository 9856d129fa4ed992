// Log-mel front-end on the GPU (K1 of SURVEY.md): replaces the CPU/librosa call
//   librosa.feature.melspectrogram(y, sr=16000, n_fft=1024, hop_length=512, power=2) -> power_to_db(ref=np.max) -> fp16
// (scripts/utils/data_utils.py:34-38, used by the synthesize scripts and at dataset-build time).
// One CTA per tile of 8 STFT frames: samples staged once in shared memory (bulk copy), two real frames per complex
// 1024-point FFT, table-driven window / twiddles, |.|^2 -> 128 Slaney mel bands (sparse triangular rows of the
// host-computed basis) -> per-clip max.  A second kernel applies 10 log10, the per-clip reference and the 80 dB floor and
// rounds through fp16.  HBM-bound by construction: every audio sample is read once, 128 floats written per frame.
#include "common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int NFFT = 1024, HOP = 512, NBIN = 513, NMEL = 128, LOG2N = 10;
constexpr int FT = 8;                        // STFT frames per CTA = 4 complex FFTs (two real frames each)
constexpr int NP = FT / 2;                   // frame pairs
constexpr int TILE = (FT + 1) * HOP;         // audio samples a tile of FT overlapping frames covers
constexpr int PWS = NBIN + 3;                // padded power-spectrum row

__device__ __forceinline__ int bitrev10(int x) { return (int)(__brev((unsigned)x) >> (32 - LOG2N)); }
__device__ __forceinline__ uint32_t msu32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct MelSmem {
    float tile[TILE];                        // the clip's samples under the FT frames (reflect-padded at the clip edges)
    float re[NP][NFFT], im[NP][NFFT];        // NP complex FFTs in place
    float twc[NFFT / 2], tws[NFFT / 2];      // exp(-2 pi i j / 1024)
    float pw[FT][PWS];                       // |X_f[k]|^2, k = 0..512
    float shmax[8];
    unsigned long long bar;
};

// Second generation of the STFT + mel kernel.  One CTA = FT = 8 consecutive frames of one clip:
//   * the 4 608 samples under them are staged ONCE in shared memory (the frames overlap by 50 %: the first generation
//     read every sample twice from L2, one CTA per frame) -- by a single bulk asynchronous copy (cp.async.bulk, 18 KB,
//     mbarrier completion) when the tile is interior and 16-byte aligned, by coalesced loads with reflection otherwise;
//   * real-input trick: frames (2p, 2p+1) are the real and imaginary part of ONE 1024-point complex FFT, and
//     X_a[k] = (Z[k] + conj Z[N-k]) / 2, X_b[k] = (Z[k] - conj Z[N-k]) / 2i  -- half the butterflies per frame;
//   * Hann window and twiddles come from host-computed tables (float64 -> float32) instead of 1 536 sincospif / cospif
//     evaluations per frame; 8 butterflies per thread between barriers instead of 2;
//   * the 128 x 8 mel powers of the tile leave as 8 consecutive frames per mel row (full 32-byte sectors) instead of
//     one 4-byte store per sector.
__global__ void __launch_bounds__(256) mel_power_kernel(const float* __restrict__ audio, int64_t n_samples, int n_frames,
                                                        const float* __restrict__ basis /* [128][513] */,
                                                        const int* __restrict__ band_start, const int* __restrict__ band_len,
                                                        const float* __restrict__ tables /* hann[1024] | cos[512] | sin[512] */,
                                                        float* __restrict__ melpow /* [B][128][n_frames] */,
                                                        unsigned int* __restrict__ clip_max /* [B] float bits */) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MelSmem& S = *reinterpret_cast<MelSmem*>(smem_raw);
    const int f0 = blockIdx.x * FT, b = blockIdx.y, tid = threadIdx.x;
    const float* y = audio + (size_t)b * n_samples;
    const int64_t g0 = (int64_t)f0 * HOP - NFFT / 2;           // global sample under tile[0]
    const bool interior = g0 >= 0 && g0 + TILE <= n_samples;
    const bool bulk = interior && ((reinterpret_cast<uintptr_t>(y + g0) & 15) == 0);
    if (bulk) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(msu32(&S.bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(msu32(&S.bar)), "r"((uint32_t)(TILE * 4)) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(msu32(S.tile)), "l"(y + g0), "r"((uint32_t)(TILE * 4)), "r"(msu32(&S.bar)) : "memory");
        }
    } else {
        for (int i = tid; i < TILE; i += 256) {
            int64_t j = g0 + i;
            if (j < 0) j = -j;                                   // reflect padding (librosa center=True, pad_mode='reflect')
            if (j >= n_samples) j = 2 * (n_samples - 1) - j;
            if (j < 0) j = 0;                                    // (only reachable past the last real frame)
            S.tile[i] = y[j];
        }
    }
    for (int i = tid; i < NFFT / 2; i += 256) { S.twc[i] = tables[NFFT + i]; S.tws[i] = tables[NFFT + NFFT / 2 + i]; }
    if (bulk) {
        __syncthreads();   // barrier initialised before anyone polls it
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(msu32(&S.bar)) : "memory");
    }
    __syncthreads();
    // windowed frames -> bit-reversed complex inputs: re = frame 2p, im = frame 2p+1
    for (int e = tid; e < NP * NFFT; e += 256) {
        const int p = e >> LOG2N, n = e & (NFFT - 1);
        const float w = __ldg(tables + n);
        const int r = bitrev10(n);
        const int fa = 2 * p, fb = fa + 1;
        S.re[p][r] = (f0 + fa < n_frames) ? S.tile[fa * HOP + n] * w : 0.f;
        S.im[p][r] = (f0 + fb < n_frames) ? S.tile[fb * HOP + n] * w : 0.f;
    }
    __syncthreads();
    for (int s = 1; s <= LOG2N; ++s) {
        const int half = 1 << (s - 1), tstep = NFFT >> s;
#pragma unroll
        for (int q = 0; q < NP * (NFFT / 2) / 256; ++q) {
            const int e = tid + q * 256;
            const int p = e >> (LOG2N - 1), bfly = e & (NFFT / 2 - 1);
            const int k = bfly & (half - 1);
            const int i0 = ((bfly >> (s - 1)) << s) + k, i1 = i0 + half;
            const float c = S.twc[k * tstep], sn = S.tws[k * tstep];
            const float xr = S.re[p][i1], xi = S.im[p][i1];
            const float tr = c * xr - sn * xi, ti = c * xi + sn * xr;
            const float ur = S.re[p][i0], ui = S.im[p][i0];
            S.re[p][i0] = ur + tr; S.im[p][i0] = ui + ti;
            S.re[p][i1] = ur - tr; S.im[p][i1] = ui - ti;
        }
        __syncthreads();
    }
    // power spectra of the two real frames packed in each complex FFT
    for (int e = tid; e < NP * NBIN; e += 256) {
        const int p = e / NBIN, k = e % NBIN;
        const int nk = (NFFT - k) & (NFFT - 1);
        const float zr = S.re[p][k], zi = S.im[p][k], wr = S.re[p][nk], wi = S.im[p][nk];
        const float ar = zr + wr, ai = zi - wi;        // 2 X_a
        const float br = zi + wi, bi = wr - zr;        // 2 X_b
        S.pw[2 * p][k] = 0.25f * (ar * ar + ai * ai);
        S.pw[2 * p + 1][k] = 0.25f * (br * br + bi * bi);
    }
    __syncthreads();
    // 128 mel bands x 8 frames: thread -> (band, 4 consecutive frames)
    const int m = tid & (NMEL - 1), fq = (tid >> 7) * 4;
    const int st = band_start[m], ln = band_len[m];
    const float* row = basis + (size_t)m * NBIN + st;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < ln; ++k) {
        const float w = __ldg(row + k);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fmaf(w, S.pw[fq + j][st + k], acc[j]);
    }
    float mx = 0.f;
    float* dst = melpow + ((size_t)b * NMEL + m) * n_frames + f0 + fq;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (f0 + fq + j < n_frames) { dst[j] = acc[j]; mx = fmaxf(mx, acc[j]); }
    mx = warp_max(mx);
    if ((tid & 31) == 0) S.shmax[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t = fmaxf(t, S.shmax[i]);
        atomicMax(clip_max + b, __float_as_uint(t));  // non-negative floats order like their bit patterns (order-free: exact)
    }
}

// out[b][m][f] = fp16_round( max(10 log10(max(amin,S)) - 10 log10(max(amin,ref_b)), -top_db) ),  f < n_out
__global__ void mel_db_kernel(const float* __restrict__ melpow, const unsigned int* __restrict__ clip_max, int n_frames,
                              int n_out, int64_t total, float* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int f = (int)(i % n_out);
        int64_t bm = i / n_out;
        int64_t b = bm / NMEL;
        float ref = fmaxf(__uint_as_float(clip_max[b]), 1e-10f);
        float s = fmaxf(melpow[bm * n_frames + f], 1e-10f);
        float db = 10.0f * log10f(s) - 10.0f * log10f(ref);
        db = fmaxf(db, -80.0f);
        out[i] = __half2float(__float2half_rn(db));
    }
}

}  // namespace

// audio [B, n_samples] fp32 (16 kHz) -> out [B, 128, n_out] fp32 holding fp16-rounded dB values (n_out <= n_frames =
// 1 + n_samples/512).  basis [128][513], band_start/band_len [128]: Slaney mel filterbank rows and their non-zero spans.
// scratch: melpow [B*128*n_frames] floats, clip_max [B] uint32.
// tables [2048]: periodic Hann window [1024] | cos(2 pi j / 1024) [512] | -sin(2 pi j / 1024) [512].
HA2G_API int ha2g_logmel(const float* audio, int B, int64_t n_samples, const float* basis, const int* band_start,
                         const int* band_len, const float* tables, float* melpow, unsigned int* clip_max, float* out,
                         int n_out, cudaStream_t stream) {
    const int n_frames = 1 + (int)(n_samples / HOP);
    if (n_out > n_frames || n_samples < NFFT / 2 + 1) return (int)cudaErrorInvalidValue;
    cudaError_t ce = cudaMemsetAsync(clip_max, 0, sizeof(unsigned int) * B, stream);
    if (ce != cudaSuccess) return (int)ce;
    cudaError_t fe = cudaFuncSetAttribute(mel_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MelSmem));
    if (fe != cudaSuccess) return (int)fe;
    mel_power_kernel<<<dim3(ha2g_div_up(n_frames, FT), B), 256, sizeof(MelSmem), stream>>>(
        audio, n_samples, n_frames, basis, band_start, band_len, tables, melpow, clip_max);
    const int64_t total = (int64_t)B * NMEL * n_out;
    mel_db_kernel<<<ha2g_ew_grid(total), 256, 0, stream>>>(melpow, clip_max, n_frames, n_out, total, out);
    HA2G_RETURN_LAST();
}
