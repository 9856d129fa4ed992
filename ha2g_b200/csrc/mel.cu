// Log-mel front-end on the GPU (K1 of SURVEY.md): replaces the CPU/librosa call
//   librosa.feature.melspectrogram(y, sr=16000, n_fft=1024, hop_length=512, power=2) -> power_to_db(ref=np.max) -> fp16
// (scripts/utils/data_utils.py:34-38, used by the synthesize scripts and at dataset-build time).
// One CTA per STFT frame: reflect-padded, Hann-windowed 1024-sample frame -> in-shared-memory radix-2 FFT ->
// |.|^2 -> 128 Slaney mel bands (sparse triangular rows of the host-computed basis) -> per-clip max.
// A second kernel applies 10 log10, the per-clip reference and the 80 dB floor and rounds through fp16.
// HBM-bound: each audio sample is read twice (50 % frame overlap, second read hits L2), 128 floats written per frame.
#include "common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int NFFT = 1024, HOP = 512, NBIN = 513, NMEL = 128, LOG2N = 10;

__device__ __forceinline__ int bitrev10(int x) { return (int)(__brev((unsigned)x) >> (32 - LOG2N)); }

__global__ void __launch_bounds__(256) mel_power_kernel(const float* __restrict__ audio, int64_t n_samples, int n_frames,
                                                        const float* __restrict__ basis /* [128][513] */,
                                                        const int* __restrict__ band_start, const int* __restrict__ band_len,
                                                        float* __restrict__ melpow /* [B][128][n_frames] */,
                                                        unsigned int* __restrict__ clip_max /* [B] float bits */) {
    __shared__ float re[NFFT], im[NFFT];
    __shared__ float twc[NFFT / 2], tws[NFFT / 2];
    __shared__ float pw[NBIN + 3];
    __shared__ float shmax[33];
    const int frame = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const float* y = audio + (size_t)b * n_samples;
    for (int i = tid; i < NFFT / 2; i += blockDim.x) {
        float s, c;
        sincospif(-2.0f * (float)i / (float)NFFT, &s, &c);
        twc[i] = c; tws[i] = s;
    }
    for (int i = tid; i < NFFT; i += blockDim.x) {
        int64_t j = (int64_t)frame * HOP + i - NFFT / 2;
        if (j < 0) j = -j;
        if (j >= n_samples) j = 2 * (n_samples - 1) - j;
        float w = 0.5f - 0.5f * cospif(2.0f * (float)i / (float)NFFT);
        int r = bitrev10(i);
        re[r] = y[j] * w;
        im[r] = 0.f;
    }
    __syncthreads();
    for (int s = 1; s <= LOG2N; ++s) {
        const int m = 1 << s, half = m >> 1, tstep = NFFT / m;
        for (int bfly = tid; bfly < NFFT / 2; bfly += blockDim.x) {
            int k = bfly & (half - 1);
            int i0 = ((bfly >> (s - 1)) << s) + k, i1 = i0 + half;
            float c = twc[k * tstep], sn = tws[k * tstep];
            float xr = re[i1], xi = im[i1];
            float tr = c * xr - sn * xi, ti = c * xi + sn * xr;
            float ur = re[i0], ui = im[i0];
            re[i0] = ur + tr; im[i0] = ui + ti;
            re[i1] = ur - tr; im[i1] = ui - ti;
        }
        __syncthreads();
    }
    for (int k = tid; k < NBIN; k += blockDim.x) pw[k] = re[k] * re[k] + im[k] * im[k];
    __syncthreads();
    float mx = 0.f;
    if (tid < NMEL) {
        const int st = band_start[tid], ln = band_len[tid];
        const float* row = basis + (size_t)tid * NBIN + st;
        float acc = 0.f;
        for (int k = 0; k < ln; ++k) acc = fmaf(row[k], pw[st + k], acc);
        melpow[((size_t)b * NMEL + tid) * n_frames + frame] = acc;
        mx = acc;
    }
    mx = warp_max(mx);
    if ((tid & 31) == 0) shmax[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t = fmaxf(t, shmax[i]);
        atomicMax(clip_max + b, __float_as_uint(t));  // non-negative floats order like their bit patterns
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
HA2G_API int ha2g_logmel(const float* audio, int B, int64_t n_samples, const float* basis, const int* band_start,
                         const int* band_len, float* melpow, unsigned int* clip_max, float* out, int n_out,
                         cudaStream_t stream) {
    const int n_frames = 1 + (int)(n_samples / HOP);
    if (n_out > n_frames || n_samples < NFFT / 2 + 1) return (int)cudaErrorInvalidValue;
    cudaError_t ce = cudaMemsetAsync(clip_max, 0, sizeof(unsigned int) * B, stream);
    if (ce != cudaSuccess) return (int)ce;
    mel_power_kernel<<<dim3(n_frames, B), 256, 0, stream>>>(audio, n_samples, n_frames, basis, band_start, band_len, melpow,
                                                            clip_max);
    const int64_t total = (int64_t)B * NMEL * n_out;
    mel_db_kernel<<<ha2g_ew_grid(total), 256, 0, stream>>>(melpow, clip_max, n_frames, n_out, total, out);
    HA2G_RETURN_LAST();
}
