// SURVEY.md section 8(f) rank 1 -- the step that directly follows backward in the reference's loop:
//   torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1); optimizer.step()   (train.py:115-118, AdamW train.py:268)
// as two multi-tensor launches over all ~280 parameter tensors instead of ~1400 small ATen kernels and a host sync:
//   1. global sum of squares of every gradient            (one fp32 accumulator, atomics per CTA)
//   2. clip coefficient min(1, max_norm / (norm + 1e-6)) read from device memory + decoupled-weight-decay Adam update
// Chunk tables (tensor index, element offset) are built once on the host; every CTA owns one chunk.
#include "common.cuh"

namespace effdet {

__global__ void __launch_bounds__(256) multi_sumsq_kernel(const unsigned long long* __restrict__ g_ptrs,
                                                          const long long* __restrict__ numels,
                                                          const int* __restrict__ chunk_tensor,
                                                          const long long* __restrict__ chunk_off, int chunk,
                                                          float* __restrict__ norm_sq) {
    __shared__ float red[8];
    const int t = chunk_tensor[blockIdx.x];
    const long long off = chunk_off[blockIdx.x];
    const float* g = reinterpret_cast<const float*>(g_ptrs[t]);
    long long n = numels[t] - off;
    if (n > chunk) n = chunk;
    float s = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = __ldg(g + off + i);
        s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int i = 0; i < 8; ++i) tot += red[i];
        atomicAdd(norm_sq, tot);
    }
}

__global__ void __launch_bounds__(256) multi_clip_adamw_kernel(
    const unsigned long long* __restrict__ p_ptrs, const unsigned long long* __restrict__ g_ptrs,
    const unsigned long long* __restrict__ m_ptrs, const unsigned long long* __restrict__ v_ptrs,
    const long long* __restrict__ numels, const int* __restrict__ chunk_tensor, const long long* __restrict__ chunk_off,
    int chunk, const float* __restrict__ norm_sq, float max_norm, float lr, float beta1, float beta2, float eps,
    float weight_decay, float bias_c1, float bias_c2, int write_clipped_grad) {
    const int t = chunk_tensor[blockIdx.x];
    const long long off = chunk_off[blockIdx.x];
    float* p = reinterpret_cast<float*>(p_ptrs[t]) + off;
    float* g = reinterpret_cast<float*>(g_ptrs[t]) + off;
    float* m = reinterpret_cast<float*>(m_ptrs[t]) + off;
    float* v = reinterpret_cast<float*>(v_ptrs[t]) + off;
    long long n = numels[t] - off;
    if (n > chunk) n = chunk;
    // clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
    float coef = 1.f;
    if (max_norm > 0.f) coef = fminf(max_norm / (sqrtf(__ldg(norm_sq)) + 1e-6f), 1.f);
    const float step_size = lr / bias_c1;
    const float inv_sqrt_c2 = rsqrtf(bias_c2);
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float gi = g[i] * coef;
        float pi = p[i];
        pi *= (1.f - lr * weight_decay);                       // decoupled weight decay (AdamW)
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        const float denom = sqrtf(vi) * inv_sqrt_c2 + eps;
        pi -= step_size * (mi / denom);
        p[i] = pi;
        m[i] = mi;
        v[i] = vi;
        if (write_clipped_grad) g[i] = gi;
    }
}

}  // namespace effdet

using namespace effdet;

extern "C" int effdet_multi_sumsq(const uint64_t* g_ptrs, const int64_t* numels, const int32_t* chunk_tensor,
                                  const int64_t* chunk_off, int nchunks, int chunk, float* norm_sq, int device,
                                  effdet_stream_t stream) {
    EFFDET_REQUIRE(g_ptrs && numels && chunk_tensor && chunk_off && norm_sq && nchunks > 0 && chunk > 0,
                   "multi_sumsq: bad arguments");
    EFFDET_DEVICE(device);
    multi_sumsq_kernel<<<nchunks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const unsigned long long*>(g_ptrs), reinterpret_cast<const long long*>(numels), chunk_tensor,
        reinterpret_cast<const long long*>(chunk_off), chunk, norm_sq);
    return launch_status("multi_sumsq_kernel");
}

extern "C" int effdet_multi_clip_adamw(const uint64_t* p_ptrs, const uint64_t* g_ptrs, const uint64_t* m_ptrs,
                                       const uint64_t* v_ptrs, const int64_t* numels, const int32_t* chunk_tensor,
                                       const int64_t* chunk_off, int nchunks, int chunk, const float* norm_sq,
                                       float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay,
                                       float bias_c1, float bias_c2, int write_clipped_grad, int device,
                                       effdet_stream_t stream) {
    EFFDET_REQUIRE(p_ptrs && g_ptrs && m_ptrs && v_ptrs && numels && chunk_tensor && chunk_off && norm_sq && nchunks > 0 &&
                       chunk > 0,
                   "multi_clip_adamw: bad arguments");
    EFFDET_REQUIRE(bias_c1 > 0.f && bias_c2 > 0.f, "multi_clip_adamw: bias corrections must be positive");
    EFFDET_DEVICE(device);
    multi_clip_adamw_kernel<<<nchunks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const unsigned long long*>(p_ptrs), reinterpret_cast<const unsigned long long*>(g_ptrs),
        reinterpret_cast<const unsigned long long*>(m_ptrs), reinterpret_cast<const unsigned long long*>(v_ptrs),
        reinterpret_cast<const long long*>(numels), chunk_tensor, reinterpret_cast<const long long*>(chunk_off), chunk,
        norm_sq, max_norm, lr, beta1, beta2, eps, weight_decay, bias_c1, bias_c2, write_clipped_grad);
    return launch_status("multi_clip_adamw_kernel");
}
