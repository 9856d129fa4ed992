// Multi-tensor Adam (K19 of SURVEY.md): one launch updates every parameter of one optimizer.
// Semantics = torch.optim.Adam(betas=(b1,b2), eps, weight_decay=0, amsgrad=False) as configured at
// scripts/train_expressive.py:212-230 (lr 5e-4, betas (0.5, 0.999); discriminator lr x0.2):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
#include "common.cuh"

namespace {
constexpr int ADAM_CHUNK = 16384;  // elements per CTA-chunk

// One chunk of one tensor.  128-bit accesses when the four arrays are 16-byte aligned (chunk offsets are multiples of
// 16384 elements, so alignment of the bases is all that matters): 7 x 16 bytes per thread-iteration instead of 7 x 4.
__device__ __forceinline__ void adam_chunk(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                           float* __restrict__ v, int64_t off, int64_t end, float b1, float b2, float eps,
                                           float step_size, float bc2_sqrt) {
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    int64_t i = off;
    if (vec) {
        const int64_t n4 = (end - off) >> 2;
        for (int64_t q = threadIdx.x; q < n4; q += blockDim.x) {
            const int64_t e = off + (q << 2);
            const float4 g4 = *reinterpret_cast<const float4*>(g + e);
            float4 m4 = *reinterpret_cast<const float4*>(m + e), v4 = *reinterpret_cast<const float4*>(v + e);
            float4 p4 = *reinterpret_cast<const float4*>(p + e);
#define ADAM_1(c)                                                                  \
            m4.c = b1 * m4.c + (1.f - b1) * g4.c;                                  \
            v4.c = b2 * v4.c + (1.f - b2) * g4.c * g4.c;                           \
            p4.c -= step_size * (m4.c / (sqrtf(v4.c) / bc2_sqrt + eps));
            ADAM_1(x) ADAM_1(y) ADAM_1(z) ADAM_1(w)
#undef ADAM_1
            *reinterpret_cast<float4*>(m + e) = m4;
            *reinterpret_cast<float4*>(v + e) = v4;
            *reinterpret_cast<float4*>(p + e) = p4;
        }
        i = off + (n4 << 2);
    }
    for (i += threadIdx.x; i < end; i += blockDim.x) {
        const float gi = g[i];
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    }
}

// table[k] = {p, g, m, v} device addresses of tensor k;  chunk c -> (tensor chunk_tensor[c], offset chunk_off[c])
__global__ void __launch_bounds__(256) adam_multi_kernel(const int64_t* __restrict__ table, const int64_t* __restrict__ sizes,
                                                         const int* __restrict__ chunk_tensor,
                                                         const int64_t* __restrict__ chunk_off, float lr, float b1, float b2,
                                                         float eps, float bc1, float bc2_sqrt) {
    const int k = chunk_tensor[blockIdx.x];
    const int64_t off = chunk_off[blockIdx.x];
    float* __restrict__ p = reinterpret_cast<float*>(table[4 * k + 0]);
    const float* __restrict__ g = reinterpret_cast<const float*>(table[4 * k + 1]);
    float* __restrict__ m = reinterpret_cast<float*>(table[4 * k + 2]);
    float* __restrict__ v = reinterpret_cast<float*>(table[4 * k + 3]);
    const int64_t n = sizes[k];
    const int64_t end = min(n, off + ADAM_CHUNK);
    adam_chunk(p, g, m, v, off, end, b1, b2, eps, lr / bc1, bc2_sqrt);
}
// CUDA-graph variant: the step count and the hyper-parameters live in device memory, so one captured launch stays
// valid for every replay.  state[0] = step count (incremented here); hyper = {lr, b1, b2, eps}; the bias corrections
// are written to hyper[4..5] for the update kernel that follows on the same stream.
__global__ void adam_tick_kernel(int* __restrict__ step, double* __restrict__ hyper) {
    const int s = ++(*step);
    hyper[4] = 1.0 - pow(hyper[1], (double)s);
    hyper[5] = sqrt(1.0 - pow(hyper[2], (double)s));
}
__global__ void __launch_bounds__(256) adam_multi_dev_kernel(const int64_t* __restrict__ table, const int64_t* __restrict__ sizes,
                                                             const int* __restrict__ chunk_tensor,
                                                             const int64_t* __restrict__ chunk_off,
                                                             const double* __restrict__ hyper) {
    const float lr = (float)hyper[0], b1 = (float)hyper[1], b2 = (float)hyper[2], eps = (float)hyper[3];
    const float bc1 = (float)hyper[4], bc2_sqrt = (float)hyper[5];
    const int k = chunk_tensor[blockIdx.x];
    const int64_t off = chunk_off[blockIdx.x];
    float* __restrict__ p = reinterpret_cast<float*>(table[4 * k + 0]);
    const float* __restrict__ g = reinterpret_cast<const float*>(table[4 * k + 1]);
    float* __restrict__ m = reinterpret_cast<float*>(table[4 * k + 2]);
    float* __restrict__ v = reinterpret_cast<float*>(table[4 * k + 3]);
    const int64_t n = sizes[k];
    const int64_t end = min(n, off + ADAM_CHUNK);
    adam_chunk(p, g, m, v, off, end, b1, b2, eps, lr / bc1, bc2_sqrt);
}
}  // namespace

// table: [ntensors*4] int64 device addresses (p,g,m,v), sizes: [ntensors] int64, chunk maps built by the host
// (chunk = 16384 elements).  step = the (1-based) step count AFTER increment, shared by all tensors.
HA2G_API int ha2g_adam_multi(const int64_t* table, const int64_t* sizes, const int* chunk_tensor, const int64_t* chunk_off,
                             int nchunks, float lr, float b1, float b2, float eps, int step, cudaStream_t stream) {
    if (nchunks <= 0) return 0;
    const double bc1 = 1.0 - pow((double)b1, (double)step);
    const double bc2 = 1.0 - pow((double)b2, (double)step);
    adam_multi_kernel<<<nchunks, 256, 0, stream>>>(table, sizes, chunk_tensor, chunk_off, lr, b1, b2, eps, (float)bc1,
                                                   (float)sqrt(bc2));
    HA2G_RETURN_LAST();
}

// Same update with the step counter (int, device, incremented by this call) and hyper = {lr, b1, b2, eps, -, -, -, -}
// (8 doubles, device; slots 4..5 are scratch for the bias corrections) read from device memory: the launch can be
// captured once into a CUDA graph and replayed while the host only rewrites `hyper` when a schedule changes lr.
HA2G_API int ha2g_adam_multi_dev(const int64_t* table, const int64_t* sizes, const int* chunk_tensor,
                                 const int64_t* chunk_off, int nchunks, double* hyper, int* step, cudaStream_t stream) {
    if (nchunks <= 0) return 0;
    adam_tick_kernel<<<1, 1, 0, stream>>>(step, hyper);
    adam_multi_dev_kernel<<<nchunks, 256, 0, stream>>>(table, sizes, chunk_tensor, chunk_off, hyper);
    HA2G_RETURN_LAST();
}
