// HBM-bound glue kernels of the HA2G step: activations, column-block copies (concat / split /
// broadcast), embedding gather + scatter-add, reparameterize, TCN weight-norm and causal shift,
// cascade pre_seq build.  All are grid-stride, coalesced along the innermost (channel) axis.
#include "common.cuh"

namespace {

__global__ void col_sum_kernel(const float* __restrict__ x, int rows, int cols, int ld, float* __restrict__ dst,
                               int rows_per_cta, int add) {
    // blockDim = (32, 8): 32 columns x 8 row lanes.  CTA (bx, by) owns rows [by*rows_per_cta, ...) of 32 columns and writes
    // ONE value per column: dst[by*cols + c] (add == 0: partial, reduced in order by col_sum_reduce_kernel) or, when the
    // grid has a single row chunk, dst[c] += t.  No atomics: the result does not depend on CTA scheduling.
    __shared__ float sh[8][33];
    int c = blockIdx.x * 32 + threadIdx.x;
    int r0 = blockIdx.y * rows_per_cta;
    int r1 = min(rows, r0 + rows_per_cta);
    float s = 0.f;
    if (c < cols)
        for (int r = r0 + threadIdx.y; r < r1; r += 8) s += x[(size_t)r * ld + c];
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
        if (add) dst[c] += t;
        else dst[(size_t)blockIdx.y * cols + c] = t;
    }
}
// 32 columns x 32 row lanes, ALL rows: out[c] += sum (4 loads in flight per lane; lanes combined in a fixed order)
__global__ void __launch_bounds__(1024) col_sum_owner_kernel(const float* __restrict__ x, int rows, int cols, int ld,
                                                             float* __restrict__ out) {
    __shared__ float sh[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (c < cols) {
        int r = threadIdx.y;
        for (; r + 96 < rows; r += 128) {
            s0 += x[(size_t)r * ld + c]; s1 += x[(size_t)(r + 32) * ld + c];
            s2 += x[(size_t)(r + 64) * ld + c]; s3 += x[(size_t)(r + 96) * ld + c];
        }
        for (; r < rows; r += 32) s0 += x[(size_t)r * ld + c];
    }
    sh[threadIdx.y][threadIdx.x] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) t += sh[i][threadIdx.x];
        out[c] += t;
    }
}
// out[c] += part[0][c] + part[1][c] + ... in index order
__global__ void col_sum_reduce_kernel(const float* __restrict__ part, int ny, int cols, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float t = 0.f;
    for (int y = 0; y < ny; ++y) t += part[(size_t)y * cols + c];
    out[c] += t;
}

__global__ void act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, int act) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = ha2g_act(x[i], act);
}
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx,
                               int64_t n, int act) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dx[i] = dy[i] * ha2g_act_grad_from_out(y[i], act);
}
// y = act(a + b)
__global__ void add_act_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
                                   int64_t n, int act) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = ha2g_act(a[i] + b[i], act);
}
// out = alpha * a + beta * b  (b may be nullptr)
__global__ void axpby_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                             int64_t n, float alpha, float beta) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = alpha * a[i] + (b != nullptr ? beta * b[i] : 0.f);
}
// y = x * mask * scale   (dropout forward and backward)
__global__ void mul_mask_kernel(const float* __restrict__ x, const float* __restrict__ mask, float scale,
                                float* __restrict__ y, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = x[i] * mask[i] * scale;
}

__global__ void embedding_fwd_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx,
                                     float* __restrict__ out, int64_t n_idx, int dim) {
    const int64_t n = n_idx * dim;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / dim;
        int c = (int)(i % dim);
        out[i] = table[idx[r] * dim + c];
    }
}
// Deterministic scatter-add of nn.Embedding's dense gradient.  Rows that share an index are summed by ONE CTA in
// ascending row order (fixed 8-way interleave, combined in lane order), so the result is independent of scheduling.
//   pass 1: head[r] = 1 iff no earlier row has the same index
//   pass 2: CTA (r, column chunk) of a head row compacts the member rows {r'' >= r : idx[r''] == idx[r]} in order into
//           shared memory and accumulates dout over them
__global__ void __launch_bounds__(256) embedding_heads_kernel(const int64_t* __restrict__ idx, int n,
                                                              unsigned char* __restrict__ head) {
    __shared__ int64_t tile[256];
    const int r = blockIdx.x * 256 + threadIdx.x;
    const int64_t me = r < n ? idx[r] : -1;
    int dup = 0;
    for (int k0 = 0; k0 <= blockIdx.x * 256; k0 += 256) {   // only rows before this block's last row matter
        __syncthreads();
        tile[threadIdx.x] = k0 + threadIdx.x < n ? idx[k0 + threadIdx.x] : -2;
        __syncthreads();
        const int lim = min(256, r - k0);                    // compare against rows k < r only
#pragma unroll 8
        for (int k = 0; k < 256; ++k) dup |= (k < lim && tile[k] == me) ? 1 : 0;
    }
    if (r < n) head[r] = dup ? 0 : 1;
}
__global__ void __launch_bounds__(256) embedding_bwd_kernel(const float* __restrict__ dout, const int64_t* __restrict__ idx,
                                                            const unsigned char* __restrict__ head,
                                                            float* __restrict__ dtable, int n, int dim) {
    extern __shared__ int members[];          // [n] worst case
    __shared__ int warp_cnt[8];
    __shared__ float part[8][32];
    const int r = blockIdx.x;
    if (!head[r]) return;
    const int64_t me = idx[r];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int count = 0;
    for (int base = r; base < n; base += 256) {
        const int k = base + tid;
        const bool hit = k < n && idx[k] == me;
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) warp_cnt[w] = __popc(bal);
        __syncthreads();
        int off = count;
        for (int i = 0; i < w; ++i) off += warp_cnt[i];
        if (hit) members[off + __popc(bal & ((1u << lane) - 1u))] = k;
        for (int i = 0; i < 8; ++i) count += warp_cnt[i];
        __syncthreads();
    }
    const int c = blockIdx.y * 32 + lane;
    float acc = 0.f;
    if (c < dim)
        for (int k = w; k < count; k += 8) acc += dout[(size_t)members[k] * dim + c];
    part[w][lane] = acc;
    __syncthreads();
    if (w == 0 && c < dim) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][lane];
        dtable[me * dim + c] += t;
    }
}

// dst[g*dst_ld + dst_off + c] += sum over the `group` consecutive source rows of group g (in order): the z gradient of
// concat_seq, where z was broadcast over T
__global__ void sum_row_groups_kernel(const float* __restrict__ src, int src_ld, int src_off, float* __restrict__ dst,
                                      int dst_ld, int dst_off, int64_t groups, int group, int ncols) {
    const int64_t n = groups * ncols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = i / ncols;
        const int c = (int)(i % ncols);
        float t = 0.f;
        for (int k = 0; k < group; ++k) t += src[(g * group + k) * src_ld + src_off + c];
        dst[g * dst_ld + dst_off + c] += t;
    }
}

// dst[(r / dst_div) * dst_ld + dst_off + c] (op)= src[(r / src_div) * src_ld + src_off + c]
__global__ void copy_cols_kernel(const float* __restrict__ src, int src_ld, int src_off, int src_div,
                                 float* __restrict__ dst, int dst_ld, int dst_off, int dst_div, int64_t rows, int ncols,
                                 int mode) {
    const int64_t n = rows * ncols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / ncols;
        int c = (int)(i % ncols);
        float v = src[(r / src_div) * src_ld + src_off + c];
        float* d = dst + (r / dst_div) * dst_ld + dst_off + c;
        if (mode == 0) *d = v;
        else *d += v;
    }
}

__global__ void gather_cols_kernel(const float* __restrict__ src, int src_ld, const int* __restrict__ idx, int n_idx,
                                   float* __restrict__ dst, int dst_ld, int64_t rows) {
    const int64_t n = rows * n_idx;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / n_idx;
        int c = (int)(i % n_idx);
        dst[r * dst_ld + c] = src[r * src_ld + idx[c]];
    }
}

__global__ void reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ logvar,
                                   const float* __restrict__ eps, float* __restrict__ z, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        z[i] = mu[i] + eps[i] * expf(0.5f * logvar[i]);
}
__global__ void reparam_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ logvar,
                                   const float* __restrict__ eps, float* __restrict__ dmu, float* __restrict__ dlogvar,
                                   int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        dmu[i] = dz[i];
        dlogvar[i] = dz[i] * eps[i] * 0.5f * expf(0.5f * logvar[i]);
    }
}

// weight_norm (dim=0) of a Conv1d weight v[O][I][Kw] -> wcat[O][Kw*I] with wcat[o][k*I+i] = g[o]*v[o][i][k]/||v[o]||
__global__ void tcn_weight_fwd_kernel(const float* __restrict__ g, const float* __restrict__ v, float* __restrict__ wcat,
                                      float* __restrict__ norm_out, int I, int Kw) {
    __shared__ float sh[33];
    const int o = blockIdx.x;
    const int n = I * Kw;
    const float* vo = v + (size_t)o * n;
    float s = 0.f;
    for (int e = threadIdx.x; e < n; e += blockDim.x) s += vo[e] * vo[e];
    s = block_sum(s, sh);
    const float nrm = sqrtf(s);
    if (threadIdx.x == 0) norm_out[o] = nrm;
    const float sc = g[o] / nrm;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        int i = e / Kw, k = e % Kw;
        wcat[(size_t)o * n + k * I + i] = vo[e] * sc;
    }
}
// dg[o] += sum(dW*v)/||v||;  dv[o] += g/||v|| * (dW - (sum(dW*v)/||v||^2) * v)
__global__ void tcn_weight_bwd_kernel(const float* __restrict__ dwcat, const float* __restrict__ g,
                                      const float* __restrict__ v, const float* __restrict__ norm,
                                      float* __restrict__ dg, float* __restrict__ dv, int I, int Kw) {
    __shared__ float sh[33];
    const int o = blockIdx.x;
    const int n = I * Kw;
    const float* vo = v + (size_t)o * n;
    float s = 0.f;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        int i = e / Kw, k = e % Kw;
        s += dwcat[(size_t)o * n + k * I + i] * vo[e];
    }
    s = block_sum(s, sh);
    const float nrm = norm[o];
    if (threadIdx.x == 0) dg[o] += s / nrm;
    const float a = g[o] / nrm, bcoef = s / (nrm * nrm);
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        int i = e / Kw, k = e % Kw;
        dv[(size_t)o * n + e] += a * (dwcat[(size_t)o * n + k * I + i] - bcoef * vo[e]);
    }
}

// out[b,t] = [ x[b,t-d] (zeros when t<d) | x[b,t] ]      x [B,T,C] -> out [B,T,2C]
__global__ void shift_concat_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t B, int T, int C,
                                        int d) {
    const int64_t n = B * T * 2 * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int c2 = (int)(i % (2 * C));
        int64_t bt = i / (2 * C);
        int t = (int)(bt % T);
        float v;
        if (c2 < C) v = t >= d ? x[(bt - d) * C + c2] : 0.f;
        else v = x[bt * C + (c2 - C)];
        out[i] = v;
    }
}
// 128-bit variants (C % 4 == 0, 16-byte aligned tensors): one float4 per thread-iteration, index math per float4
__global__ void shift_concat_fwd_vec_kernel(const float4* __restrict__ x, float4* __restrict__ out, int64_t B, int T, int C4,
                                            int d) {
    const int64_t n = B * T * 2 * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c2 = (int)(i % (2 * C4));
        const int64_t bt = i / (2 * C4);
        const int t = (int)(bt % T);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c2 < C4) { if (t >= d) v = x[(bt - d) * C4 + c2]; }
        else v = x[bt * C4 + (c2 - C4)];
        out[i] = v;
    }
}
__global__ void shift_concat_bwd_vec_kernel(const float4* __restrict__ dout, float4* __restrict__ dx, int64_t B, int T,
                                            int C4, int d) {
    const int64_t n = B * T * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        const int64_t bt = i / C4;
        const int t = (int)(bt % T);
        float4 v = dout[bt * 2 * C4 + C4 + c];
        if (t + d < T) {
            const float4 w = dout[(bt + d) * 2 * C4 + c];
            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
        dx[i] = v;
    }
}
// dx[b,t] = dout[b,t][C:2C] + (t+d < T ? dout[b,t+d][0:C] : 0)
__global__ void shift_concat_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx, int64_t B, int T, int C,
                                        int d) {
    const int64_t n = B * T * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        int64_t bt = i / C;
        int t = (int)(bt % T);
        float v = dout[bt * 2 * C + C + c];
        if (t + d < T) v += dout[(bt + d) * 2 * C + c];
        dx[i] = v;
    }
}

// cascade glue (train_hierarchy_expressive.py:252-262): pre[b,t,:] for one level.
//   t <  n_pre: pre[.., :d] = target_k, flag = 1
//   t >= n_pre: pre[.., dst_idx[i]] = prev_out[.., src_idx[i]], everything else 0
__global__ void pre_seq_fwd_kernel(const float* __restrict__ target_k, const float* __restrict__ prev_out, int dp,
                                   const int* __restrict__ slot_src /* [d+1]: src col in prev_out or -1 */,
                                   float* __restrict__ pre, int64_t B, int T, int d, int n_pre) {
    const int w = d + 1;
    const int64_t n = B * T * w;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % w);
        int64_t bt = i / w;
        int t = (int)(bt % T);
        float v = 0.f;
        if (t < n_pre) v = c < d ? target_k[bt * d + c] : 1.f;
        else if (prev_out != nullptr) {
            int s = slot_src[c];
            if (s >= 0) v = prev_out[bt * dp + s];
        }
        pre[i] = v;
    }
}
// dprev[b,t,s] = (t >= n_pre && src_slot[s] >= 0) ? dpre[b,t,src_slot[s]] : 0     (overwrites dprev)
__global__ void pre_seq_bwd_kernel(const float* __restrict__ dpre, int w,
                                   const int* __restrict__ src_slot /* [dp]: dst col in pre or -1 */,
                                   float* __restrict__ dprev, int64_t B, int T, int dp, int n_pre) {
    const int64_t n = B * T * dp;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int s = (int)(i % dp);
        int64_t bt = i / dp;
        int t = (int)(bt % T);
        float v = 0.f;
        if (t >= n_pre) {
            int c = src_slot[s];
            if (c >= 0) v = dpre[bt * w + c];
        }
        dprev[i] = v;
    }
}

}  // namespace

#define EW_LAUNCH(kern, n, ...) \
    do { if ((n) > 0) kern<<<ha2g_ew_grid((n)), 256, 0, stream>>>(__VA_ARGS__); HA2G_RETURN_LAST(); } while (0)

// out[c] += sum_r x[r*ld + c]   (out must be initialised by the caller).  Deterministic: up to 1 024 rows one CTA of
// 32 x 32 threads owns 32 columns over ALL rows (one launch); beyond that row chunks (enough CTAs to cover the SMs) write
// partial sums into the scratch arena and a second kernel adds them in chunk order.
HA2G_API int ha2g_col_sum(const float* x, int rows, int cols, int ld, float* out, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return 0;
    int gx = ha2g_div_up(cols, 32);
    if (rows <= 1024) {
        col_sum_owner_kernel<<<gx, dim3(32, 32), 0, stream>>>(x, rows, cols, ld, out);
        HA2G_RETURN_LAST();
    }
    int want_y = ha2g_div_up(148 * 4, gx);
    int rows_per = ha2g_div_up(rows, want_y);
    if (rows_per < 64) rows_per = 64;
    const int gy = ha2g_div_up(rows, rows_per);
    dim3 grid(gx, gy);
    float* part = reinterpret_cast<float*>(ha2g_ws((size_t)gy * cols * sizeof(float), stream));
    if (part == nullptr) return (int)cudaErrorMemoryAllocation;
    col_sum_kernel<<<grid, dim3(32, 8), 0, stream>>>(x, rows, cols, ld, part, rows_per, 0);
    col_sum_reduce_kernel<<<ha2g_div_up(cols, 128), 128, 0, stream>>>(part, gy, cols, out);
    HA2G_RETURN_LAST();
}
HA2G_API int ha2g_act_fwd(const float* x, float* y, int64_t n, int act, cudaStream_t stream) {
    EW_LAUNCH(act_fwd_kernel, n, x, y, n, act);
}
HA2G_API int ha2g_act_bwd(const float* dy, const float* y, float* dx, int64_t n, int act, cudaStream_t stream) {
    EW_LAUNCH(act_bwd_kernel, n, dy, y, dx, n, act);
}
HA2G_API int ha2g_add_act_fwd(const float* a, const float* b, float* y, int64_t n, int act, cudaStream_t stream) {
    EW_LAUNCH(add_act_fwd_kernel, n, a, b, y, n, act);
}
HA2G_API int ha2g_axpby(const float* a, const float* b, float* out, int64_t n, float alpha, float beta,
                        cudaStream_t stream) {
    EW_LAUNCH(axpby_kernel, n, a, b, out, n, alpha, beta);
}
HA2G_API int ha2g_mul_mask(const float* x, const float* mask, float scale, float* y, int64_t n, cudaStream_t stream) {
    EW_LAUNCH(mul_mask_kernel, n, x, mask, scale, y, n);
}
// out[r,:] = table[idx[r],:]      (nn.Embedding forward, hierarchy_net.py:49,80,117)
HA2G_API int ha2g_embedding_fwd(const float* table, const int64_t* idx, float* out, int64_t n_idx, int dim,
                                cudaStream_t stream) {
    EW_LAUNCH(embedding_fwd_kernel, n_idx * dim, table, idx, out, n_idx, dim);
}
// head[r] = 1 iff row r is the first occurrence of idx[r] (one byte per row): the work list of ha2g_embedding_bwd.  It
// depends on the indices only, so the seven text encoders of a step (same token batch) share one call.
HA2G_API int ha2g_embedding_heads(const int64_t* idx, int64_t n_idx, unsigned char* head, cudaStream_t stream) {
    if (n_idx <= 0) return 0;
    embedding_heads_kernel<<<ha2g_div_up(n_idx, 256), 256, 0, stream>>>(idx, (int)n_idx, head);
    HA2G_RETURN_LAST();
}
// dtable[idx[r],:] += dout[r,:]   (dense gradient like nn.Embedding(sparse=False)); deterministic, see the kernels.
// head: from ha2g_embedding_heads for the same idx.
HA2G_API int ha2g_embedding_bwd(const float* dout, const int64_t* idx, const unsigned char* head, float* dtable,
                                int64_t n_idx, int dim, cudaStream_t stream) {
    if (n_idx <= 0 || dim <= 0) return 0;
    const int n = (int)n_idx;
    const size_t smem = (size_t)n * sizeof(int);
    if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(embedding_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    embedding_bwd_kernel<<<dim3(n, ha2g_div_up(dim, 32)), 256, smem, stream>>>(dout, idx, head, dtable, n, dim);
    HA2G_RETURN_LAST();
}
// mode 0: assign, 1: +=, 2: dst row g = r / dst_div accumulates its dst_div consecutive source rows in order (src_div == 1)
HA2G_API int ha2g_copy_cols(const float* src, int src_ld, int src_off, int src_div, float* dst, int dst_ld, int dst_off,
                            int dst_div, int64_t rows, int ncols, int mode, cudaStream_t stream) {
    if (mode == 2) {
        if (src_div != 1 || dst_div < 1 || rows % dst_div != 0) return (int)cudaErrorInvalidValue;
        const int64_t groups = rows / dst_div;
        EW_LAUNCH(sum_row_groups_kernel, groups * ncols, src, src_ld, src_off, dst, dst_ld, dst_off, groups, dst_div, ncols);
    }
    EW_LAUNCH(copy_cols_kernel, rows * ncols, src, src_ld, src_off, src_div, dst, dst_ld, dst_off, dst_div, rows, ncols,
              mode);
}
HA2G_API int ha2g_gather_cols(const float* src, int src_ld, const int* idx, int n_idx, float* dst, int dst_ld,
                              int64_t rows, cudaStream_t stream) {
    EW_LAUNCH(gather_cols_kernel, rows * n_idx, src, src_ld, idx, n_idx, dst, dst_ld, rows);
}
// embedding_net.py:10-13
HA2G_API int ha2g_reparam_fwd(const float* mu, const float* logvar, const float* eps, float* z, int64_t n,
                              cudaStream_t stream) {
    EW_LAUNCH(reparam_fwd_kernel, n, mu, logvar, eps, z, n);
}
HA2G_API int ha2g_reparam_bwd(const float* dz, const float* logvar, const float* eps, float* dmu, float* dlogvar,
                              int64_t n, cudaStream_t stream) {
    EW_LAUNCH(reparam_bwd_kernel, n, dz, logvar, eps, dmu, dlogvar, n);
}
// tcn.py:19-24 weight_norm(Conv1d): g [O], v [O][I][Kw] -> wcat [O][Kw*I], norm [O]
HA2G_API int ha2g_tcn_weight_fwd(const float* g, const float* v, float* wcat, float* norm, int O, int I, int Kw,
                                 cudaStream_t stream) {
    tcn_weight_fwd_kernel<<<O, 256, 0, stream>>>(g, v, wcat, norm, I, Kw);
    HA2G_RETURN_LAST();
}
HA2G_API int ha2g_tcn_weight_bwd(const float* dwcat, const float* g, const float* v, const float* norm, float* dg,
                                 float* dv, int O, int I, int Kw, cudaStream_t stream) {
    tcn_weight_bwd_kernel<<<O, 256, 0, stream>>>(dwcat, g, v, norm, dg, dv, I, Kw);
    HA2G_RETURN_LAST();
}
// causal dilated k=2 conv input staging (tcn.py:19-21 padding + Chomp1d)
HA2G_API int ha2g_shift_concat_fwd(const float* x, float* out, int64_t B, int T, int C, int d, cudaStream_t stream) {
    if (C % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
        const int64_t n4 = B * T * 2 * (C / 4);
        if (n4 <= 0) return 0;
        shift_concat_fwd_vec_kernel<<<ha2g_ew_grid(n4, 256, 2), 256, 0, stream>>>(reinterpret_cast<const float4*>(x),
                                                                               reinterpret_cast<float4*>(out), B, T, C / 4, d);
        HA2G_RETURN_LAST();
    }
    EW_LAUNCH(shift_concat_fwd_kernel, B * T * 2 * C, x, out, B, T, C, d);
}
HA2G_API int ha2g_shift_concat_bwd(const float* dout, float* dx, int64_t B, int T, int C, int d, cudaStream_t stream) {
    if (C % 4 == 0 && ((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0) {
        const int64_t n4 = B * T * (C / 4);
        if (n4 <= 0) return 0;
        shift_concat_bwd_vec_kernel<<<ha2g_ew_grid(n4, 256, 2), 256, 0, stream>>>(reinterpret_cast<const float4*>(dout),
                                                                               reinterpret_cast<float4*>(dx), B, T, C / 4, d);
        HA2G_RETURN_LAST();
    }
    EW_LAUNCH(shift_concat_bwd_kernel, B * T * C, dout, dx, B, T, C, d);
}
HA2G_API int ha2g_pre_seq_fwd(const float* target_k, const float* prev_out, int dp, const int* slot_src, float* pre,
                              int64_t B, int T, int d, int n_pre, cudaStream_t stream) {
    EW_LAUNCH(pre_seq_fwd_kernel, B * T * (d + 1), target_k, prev_out, dp, slot_src, pre, B, T, d, n_pre);
}
HA2G_API int ha2g_pre_seq_bwd(const float* dpre, int w, const int* src_slot, float* dprev, int64_t B, int T, int dp,
                              int n_pre, cudaStream_t stream) {
    EW_LAUNCH(pre_seq_bwd_kernel, B * T * dp, dpre, w, src_slot, dprev, B, T, dp, n_pre);
}

namespace {
// x [B,T,C] -> out [B,To,Kw*C] with out[b,t,k*C+c] = x[b,t+k,c]   (valid Conv1d as a GEMM; To = T-Kw+1)
__global__ void unfold1d_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t B, int T, int C, int Kw) {
    const int To = T - Kw + 1, W = Kw * C;
    const int64_t n = B * To * W;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int kc = (int)(i % W);
        int64_t bt = i / W;
        int t = (int)(bt % To);
        int64_t b = bt / To;
        int k = kc / C, c = kc % C;
        out[i] = x[(b * T + t + k) * C + c];
    }
}
// dx[b,t,c] = sum_k dout[b,t-k,k*C+c] over valid t-k in [0,To)
__global__ void unfold1d_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx, int64_t B, int T, int C, int Kw) {
    const int To = T - Kw + 1, W = Kw * C;
    const int64_t n = B * T * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        int64_t bt = i / C;
        int t = (int)(bt % T);
        int64_t b = bt / T;
        float v = 0.f;
        for (int k = 0; k < Kw; ++k) {
            int to = t - k;
            if (to >= 0 && to < To) v += dout[(b * To + to) * W + k * C + c];
        }
        dx[i] = v;
    }
}
// inverse == 0: wp[o][k*I+i] = w[o][i][k];  inverse != 0: w[o][i][k] = wp[o][k*I+i]
__global__ void pack_conv1d_w_kernel(const float* __restrict__ src, float* __restrict__ dst, int O, int I, int Kw, int inverse) {
    const int64_t n = (int64_t)O * I * Kw;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int k = (int)(e % Kw);
        int i = (int)((e / Kw) % I);
        int64_t o = e / ((int64_t)Kw * I);
        int64_t packed = o * Kw * I + (int64_t)k * I + i;
        if (inverse) dst[e] = src[packed]; else dst[packed] = src[e];
    }
}
}  // namespace

// Conv1d (no padding) input staging for the discriminator front-end (hierarchy_net.py:202-211)
HA2G_API int ha2g_unfold1d_fwd(const float* x, float* out, int64_t B, int T, int C, int Kw, cudaStream_t stream) {
    EW_LAUNCH(unfold1d_fwd_kernel, B * (T - Kw + 1) * Kw * C, x, out, B, T, C, Kw);
}
HA2G_API int ha2g_unfold1d_bwd(const float* dout, float* dx, int64_t B, int T, int C, int Kw, cudaStream_t stream) {
    EW_LAUNCH(unfold1d_bwd_kernel, B * T * C, dout, dx, B, T, C, Kw);
}
// Conv1d weight [O][I][Kw] <-> GEMM layout [O][Kw*I]
HA2G_API int ha2g_pack_conv1d_w(const float* src, float* dst, int O, int I, int Kw, int inverse, cudaStream_t stream) {
    EW_LAUNCH(pack_conv1d_w_kernel, (int64_t)O * I * Kw, src, dst, O, I, Kw, inverse);
}

namespace {
// perm[rank(j)] = j with rank(j) = #{k : key[k] < key[j] or (key[k] == key[j] and k < j)}: a uniformly random
// permutation when the keys are i.i.d. uniform draws.  O(n^2) compares; n is the batch size (<= a few thousand).
__global__ void __launch_bounds__(256) rank_perm_kernel(const float* __restrict__ keys, int64_t* __restrict__ perm, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float kj = keys[j];
    int r = 0;
    for (int k = 0; k < n; ++k) {
        const float kk = keys[k];
        r += (kk < kj || (kk == kj && k < j)) ? 1 : 0;
    }
    perm[r] = (int64_t)j;
}
}  // namespace

// torch.randperm(n) for the mismatched-speaker pass (scripts/train_eval/train_hierarchy_expressive.py:328) from n uniform
// keys: stream-ordered, no host synchronisation, capturable into a CUDA graph.  Bit-exact index contract: perm is a
// permutation of 0..n-1 (int64).
HA2G_API int ha2g_rank_perm(const float* keys, int64_t* perm, int n, cudaStream_t stream) {
    if (n <= 0) return 0;
    rank_perm_kernel<<<ha2g_div_up(n, 256), 256, 0, stream>>>(keys, perm, n);
    HA2G_RETURN_LAST();
}

namespace {
// Philox4x32-10 (Salmon et al., SC'11): counter-based, so a dropout mask is a pure function of
// (seed, step, call id, element index) and the backward pass regenerates it instead of reading a stored mask.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// state = {seed, step}: one step counter tick per training step makes every (step, call, element) triple unique
__global__ void rng_tick_kernel(unsigned long long* state) { state[1] += 1ull; }

__global__ void __launch_bounds__(256) dropout_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n,
                                                      uint32_t threshold, float scale,
                                                      const unsigned long long* __restrict__ state, uint32_t call_id) {
    const unsigned long long seed = state[0], step = state[1];
    const uint32_t k0 = (uint32_t)seed ^ (uint32_t)(step * 0x9E3779B97F4A7C15ull >> 32), k1 = (uint32_t)(seed >> 32) ^ (uint32_t)step;
    const int64_t n4 = (n + 3) >> 2;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n4; q += (int64_t)gridDim.x * blockDim.x) {
        uint32_t r[4];
        philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), call_id, 0x2B992DDFu, k0, k1, r);
        const int64_t i = q << 2;
        if (i + 3 < n && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
            const float4 v = *reinterpret_cast<const float4*>(x + i);
            float4 o;
            o.x = r[0] >= threshold ? v.x * scale : 0.f; o.y = r[1] >= threshold ? v.y * scale : 0.f;
            o.z = r[2] >= threshold ? v.z * scale : 0.f; o.w = r[3] >= threshold ? v.w * scale : 0.f;
            *reinterpret_cast<float4*>(y + i) = o;
        } else {
            for (int j = 0; j < 4 && i + j < n; ++j) y[i + j] = r[j] >= threshold ? x[i + j] * scale : 0.f;
        }
    }
}
}  // namespace

// Advance the dropout stream by one training step (state = {seed, step} as two uint64 in device memory).
HA2G_API int ha2g_rng_tick(uint64_t* state, cudaStream_t stream) {
    rng_tick_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<unsigned long long*>(state));
    HA2G_RETURN_LAST();
}
// nn.Dropout(p) forward AND backward (tcn.py:22,27, hierarchy_net.py:43,88 inter-layer GRU dropout): y = x * keep / (1-p)
// with keep = Philox(seed, step, call_id, element) >= p * 2^32.  The backward pass calls it on the incoming gradient with
// the same call_id: the mask is regenerated, never stored.
HA2G_API int ha2g_dropout(const float* x, float* y, int64_t n, float p, const uint64_t* state, uint32_t call_id,
                          cudaStream_t stream) {
    if (n <= 0) return 0;
    double t = (double)p * 4294967296.0;
    if (t < 0.0) t = 0.0;
    if (t > 4294967295.0) t = 4294967295.0;
    dropout_kernel<<<ha2g_ew_grid((n + 3) / 4, 256, 2), 256, 0, stream>>>(x, y, n, (uint32_t)t, 1.f / (1.f - p),
                                                                          reinterpret_cast<const unsigned long long*>(state), call_id);
    HA2G_RETURN_LAST();
}

namespace {
// one thread per clip: the reference's sequential loop (later words overwrite earlier ones on the same frame)
__global__ void place_words_kernel(const int64_t* __restrict__ word_id, const double* __restrict__ word_start,
                                   const int* __restrict__ word_off, const double* __restrict__ clip_start,
                                   const double* __restrict__ clip_end, int B, int n_frames, int64_t* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int64_t* row = out + (size_t)b * n_frames;
    for (int f = 0; f < n_frames; ++f) row[f] = 0;   // 0 = PAD
    const double st = clip_start[b];
    const double frame_duration = (clip_end[b] - st) / (double)n_frames;
    for (int w = word_off[b]; w < word_off[b + 1]; ++w) {
        int idx = (int)floor((word_start[w] - st) / frame_duration);
        if (idx < 0) idx = 0;
        if (idx < n_frames) row[idx] = word_id[w];
    }
}
}  // namespace

// extended_word_seq of a whole batch on the device (SpeechMotionDataset.extend_word_seq,
// scripts/data_loader/lmdb_data_loader_expressive.py:116-141, timing-preserving branch): clip b owns the words
// [word_off[b], word_off[b+1]); word w lands on frame max(0, floor((word_start[w] - clip_start[b]) / frame_duration)) when that
// is < n_frames, with frame_duration = (clip_end[b] - clip_start[b]) / n_frames evaluated in float64 like the reference's
// Python arithmetic -- bit-exact index contract.  out [B, n_frames] int64.
HA2G_API int ha2g_place_words(const int64_t* word_id, const double* word_start, const int* word_off, const double* clip_start,
                              const double* clip_end, int B, int n_frames, int64_t* out, cudaStream_t stream) {
    if (B <= 0) return 0;
    place_words_kernel<<<ha2g_div_up(B, 128), 128, 0, stream>>>(word_id, word_start, word_off, clip_start, clip_end, B, n_frames, out);
    HA2G_RETURN_LAST();
}

namespace {
// x [W, T, D]: for every window w >= 1 and overlap frame j < n: x[w, j] = x[w-1, T-n+j] * (n-j)/(n+1) + x[w, j] * (j+1)/(n+1).
// In place: the tails read (frames T-n.. of window w-1) are never written (T > 2n).
__global__ void crossfade_kernel(float* __restrict__ x, int W, int T, int D, int n) {
    const int64_t total = (int64_t)(W - 1) * n * D;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int d = (int)(i % D);
        const int j = (int)((i / D) % n);
        const int64_t w = 1 + i / ((int64_t)D * n);
        const float prev = x[((w - 1) * T + (T - n + j)) * D + d];
        float* cur = x + (w * T + j) * D + d;
        *cur = prev * (float)(n - j) / (float)(n + 1) + *cur * (float)(j + 1) / (float)(n + 1);
    }
}
}  // namespace

// Linear cross-fade between consecutive inference windows over their n overlapping frames, all windows at once
// (scripts/synthesize_expressive_hierarchy.py:195-203: out_seq[j] = last[j] * (n - j) / (n + 1) + out_seq[j] * (j + 1) / (n + 1)).
// x [W, T, D] holds the raw window outputs; in place.  Requires T > 2 n.
HA2G_API int ha2g_crossfade(float* x, int W, int T, int D, int n, cudaStream_t stream) {
    if (W <= 1 || n <= 0) return 0;
    if (T <= 2 * n) return (int)cudaErrorInvalidValue;
    const int64_t total = (int64_t)(W - 1) * n * D;
    crossfade_kernel<<<ha2g_ew_grid(total), 256, 0, stream>>>(x, W, T, D, n);
    HA2G_RETURN_LAST();
}

namespace {
// out[b,to,k*C+c] = x[b, to*stride - pad + k, c] (0 outside [0,T))
__global__ void unfold1d_strided_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t B, int T, int C,
                                            int Kw, int stride, int pad, int To) {
    const int W = Kw * C;
    const int64_t n = B * To * W;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int kc = (int)(i % W);
        const int64_t bt = i / W;
        const int to = (int)(bt % To);
        const int64_t b = bt / To;
        const int k = kc / C, c = kc % C;
        const int t = to * stride - pad + k;
        out[i] = (t >= 0 && t < T) ? x[(b * T + t) * C + c] : 0.f;
    }
}
// dx[b,t,c] = sum over (to, k) with to*stride - pad + k == t of dout[b,to,k*C+c]   (gather form: no atomics)
__global__ void unfold1d_strided_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx, int64_t B, int T, int C,
                                            int Kw, int stride, int pad, int To) {
    const int W = Kw * C;
    const int64_t n = B * T * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t bt = i / C;
        const int t = (int)(bt % T);
        const int64_t b = bt / T;
        float v = 0.f;
        // k = t + pad - to*stride in [0, Kw)  <=>  to in [ceil((t + pad - Kw + 1) / stride), floor((t + pad) / stride)]
        const int hi = min(To - 1, (t + pad) / stride);
        int lo = t + pad - Kw + 1;
        lo = lo <= 0 ? 0 : (lo + stride - 1) / stride;
        for (int to = lo; to <= hi; ++to) {
            const int k = t + pad - to * stride;
            v += dout[(b * To + to) * W + k * C + c];
        }
        dx[i] = v;
    }
}
}  // namespace

// Strided, zero-padded Conv1d input staging (im2col) for the baseline WavEncoder (multimodal_context_net.py:13-22):
// x [B,T,C] -> out [B,To,Kw*C], To = (T + 2*pad - Kw)/stride + 1
HA2G_API int ha2g_unfold1d_strided_fwd(const float* x, float* out, int64_t B, int T, int C, int Kw, int stride, int pad,
                                       int To, cudaStream_t stream) {
    EW_LAUNCH(unfold1d_strided_fwd_kernel, B * To * Kw * C, x, out, B, T, C, Kw, stride, pad, To);
}
HA2G_API int ha2g_unfold1d_strided_bwd(const float* dout, float* dx, int64_t B, int T, int C, int Kw, int stride, int pad,
                                       int To, cudaStream_t stream) {
    EW_LAUNCH(unfold1d_strided_bwd_kernel, B * T * C, dout, dx, B, T, C, Kw, stride, pad, To);
}
