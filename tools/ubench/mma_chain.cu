// Micro-benchmark behind the GRU recurrence's MMA phase: cycles per tcgen05.mma (M = 128, K = 16, kind::f16) in a chain of
// 57 MMAs (the recurrence's per-step count) as a function of N, of where A lives (tensor memory / shared memory) and of
// how many accumulators the chain alternates between.   nvcc -gencode arch=compute_100a,code=sm_100a -o mma_chain mma_chain.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mkd(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (uint64_t)((lbo >> 4) & 0x3FFF) << 16 | (uint64_t)((sbo >> 4) & 0x3FFF) << 32 | (uint64_t)1 << 46;
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbw(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(su32(bar)), "r"(parity) : "memory");
}

struct Cfg { int N, ts, nacc, nmma, same_a; };

template <int NACC>
__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long* out /* [reps][2] */, int reps) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int wu = __shfl_sync(0xffffffffu, warp, 0);
    for (int i = tid; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(su32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = slot;
    {   // zero the whole tensor memory (A operand region and accumulators)
        const uint32_t la = tb + ((uint32_t)(warp * 32) << 16);
        for (int col = 0; col < 512; col += 4)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" ::"r"(la + col), "r"(0u) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // B: [chunk][N rows][16 B] (LBO = N*16 between K chunks, 8-row groups 128 B apart), A in smem: [chunk][128 rows][16 B]
    const uint32_t b_lbo = c.N * 16, a_lbo = 128 * 16;
    unsigned char* a_smem = smem + 96 * 1024;
    uint32_t ph = 0;
    for (int r = 0; r < reps; ++r) {
        if (wu == 0) {   // whole warp, converged; one elected lane issues (no per-instruction election loop in the SASS)
          long long t0 = 0, t1 = 0; unsigned long long g0 = 0, g1 = 0;
          if (elect_one()) {
            t0 = clock64();
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
            // fully unrolled: every index below is a compile-time constant, the issue loop is MMA instructions plus one
            // 64-bit add per descriptor (as in the recurrence kernel)
            const uint64_t db0 = mkd(su32(smem), b_lbo, 128), da0 = mkd(su32(a_smem), a_lbo, 128);
            const uint64_t bstep = (uint64_t)((2 * b_lbo) >> 4), astep = (uint64_t)((2 * a_lbo) >> 4);
            const uint32_t a0 = tb + 256, dstep = (uint32_t)c.N, amul = c.same_a ? 0u : 8u;
            if (c.ts) {
#pragma unroll
                for (int i = 0; i < 57; ++i)
                    mma_ts(tb + (uint32_t)(i % NACC) * dstep, a0 + (uint32_t)(i % 19) * amul, db0 + (uint64_t)(i % 19) * bstep, idesc, 1u);
            } else {
#pragma unroll
                for (int i = 0; i < 57; ++i)
                    mma_ss(tb + (uint32_t)(i % NACC) * dstep, da0 + (uint64_t)(i % 19) * astep, db0 + (uint64_t)(i % 19) * bstep, idesc, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(&bar)) : "memory");
            t1 = clock64();
          }
          __syncwarp();
          mbw(&bar, ph);
          if (elect_one()) {
            long long t2 = clock64();
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
            out[r * 3 + 0] = t1 - t0;
            out[r * 3 + 1] = t2 - t0;
            out[r * 3 + 2] = (long long)(g1 - g0);
          }
        }
        ph ^= 1;
        __syncthreads();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
}

int main() {
    const int reps = 20;
    long long* d;
    cudaMalloc(&d, reps * 3 * sizeof(long long));
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(bench<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("| N | A in | accumulators | MMAs | issue ticks | issue+complete ticks | ns | ticks/MMA | ns/MMA |\n|---|---|---|---|---|---|---|---|---|\n");
    const int Ns[] = {16, 32, 48, 64, 96, 128, 192, 256};
    for (int same_a = 0; same_a < 2; ++same_a)
    for (int ts = 1; ts >= 0; --ts)
        for (int N : Ns)
            for (int nacc = 1; nacc <= 4; nacc *= 2) {
                if (nacc * N > 256) continue;
                if (same_a && (nacc > 1 || !ts)) continue;
                Cfg c{N, ts, nacc, 57, same_a};
                if (nacc == 1) bench<1><<<1, 128, 200 * 1024>>>(c, d, reps);
                else if (nacc == 2) bench<2><<<1, 128, 200 * 1024>>>(c, d, reps);
                else bench<4><<<1, 128, 200 * 1024>>>(c, d, reps);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("failed N=%d ts=%d: %s\n", N, ts, cudaGetErrorString(cudaGetLastError())); return 1; }
                long long h[reps * 3];
                cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                long long bi = 1LL << 60, bc = 1LL << 60, bn = 1LL << 60;
                for (int r = 2; r < reps; ++r) { if (h[r * 3 + 1] < bc) { bc = h[r * 3 + 1]; bi = h[r * 3]; } if (h[r * 3 + 2] < bn) bn = h[r * 3 + 2]; }
                printf("| %d | %s%s | %d | %d | %lld | %lld | %lld | %.1f | %.1f |\n", N, ts ? "TMEM" : "smem", same_a ? " (same A tile)" : "", nacc, c.nmma, bi, bc, bn,
                       (double)bc / c.nmma, (double)bn / c.nmma);
            }
    return 0;
}
