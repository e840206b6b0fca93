// Squeeze-excite gate of the MBConv block (models/efficientnet.py:58-65,90-94) and its backward.
//   forward : s_pre[b,j] = W1[j,:] . mean[b,:] + b1[j];   gate[b,c] = sigmoid(W2[c,:] . swish(s_pre[b,:]) + b2[c])
//   backward: dp2 = dgate * gate * (1 - gate);  dW2[c,j] = sum_b dp2[b,c] * swish(s_pre[b,j]);  db2 = sum_b dp2
//             dp1[b,j] = (sum_c dp2[b,c] W2[c,j]) * swish'(s_pre[b,j]);  dW1[j,c] = sum_b dp1[b,j] * mean[b,c];  db1 = sum_b dp1
//             dmean[b,c] = sum_j dp1[b,j] W1[j,c]
// These are tiny GEMV / outer-product problems (C <= a few thousand, S = C_in/4, B = batch): pure latency.  Round 1 ran
// one CTA per SAMPLE with C*S global atomics per sample for the weight gradients (45 us per block at bs 32, 106 us at
// bs 4 where only four CTAs existed).  Here every phase is spread over (sample, channel-slice) CTAs, one warp per dot
// product, and the weight gradients are summed over the batch inside a thread -- no atomics at all.
// W1 = _se_reduce.weight [S,C], W2 = _se_expand.weight [C,S].
#include "common.cuh"

namespace effdet {

// s_pre[b,j]: one warp per (b, j)
__global__ void __launch_bounds__(256) se_squeeze_kernel(const float* __restrict__ mean, const float* __restrict__ w1,
                                                         const float* __restrict__ b1, float* __restrict__ s_pre, int C, int S) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= S) return;
    const float* m = mean + (long long)b * C;
    const float* w = w1 + (long long)j * C;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(__ldg(m + c), __ldg(w + c), acc);
    acc = warp_sum(acc);
    if (lane == 0) s_pre[(long long)b * S + j] = acc + __ldg(b1 + j);
}

// gate[b,c]: one thread per (b, c); swish(s_pre[b,:]) staged in shared memory
__global__ void __launch_bounds__(256) se_excite_kernel(const float* __restrict__ s_pre, const float* __restrict__ w2,
                                                        const float* __restrict__ b2, float* __restrict__ gate, int C, int S) {
    extern __shared__ float sw[];
    const int b = blockIdx.y;
    for (int j = threadIdx.x; j < S; j += blockDim.x) sw[j] = swishf_(__ldg(s_pre + (long long)b * S + j));
    __syncthreads();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float acc = __ldg(b2 + c);
    const float* w = w2 + (long long)c * S;
    for (int j = 0; j < S; ++j) acc = fmaf(sw[j], __ldg(w + j), acc);
    gate[(long long)b * C + c] = sigmoidf_(acc);
}

// dp2[b,c] = dgate * g * (1 - g)   and   dp1[b,j] = (sum_c dp2[b,c] W2[c,j]) * swish'(s_pre[b,j]):  one CTA per sample
// computes dp2 into shared + global memory, then one warp per j reduces over c.
__global__ void __launch_bounds__(256) se_bwd_dp_kernel(const float* __restrict__ dgate, const float* __restrict__ gate,
                                                        const float* __restrict__ s_pre, const float* __restrict__ w2,
                                                        float* __restrict__ dp2, float* __restrict__ dp1, int C, int S) {
    extern __shared__ float d2[];        // [C]
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float g = __ldg(gate + (long long)b * C + c);
        const float d = __ldg(dgate + (long long)b * C + c) * g * (1.f - g);
        d2[c] = d;
        dp2[(long long)b * C + c] = d;
    }
    __syncthreads();
    // dp1[j] = sum_c d2[c] * W2[c,j]: lanes run over j (rows of W2 are read coalesced), warps over c, then a
    // cross-warp reduction through shared memory
    __shared__ float part[8][33];
    for (int j0 = 0; j0 < S; j0 += 32) {
        const int j = j0 + lane;
        float acc = 0.f;
        if (j < S)
            for (int c = warp; c < C; c += 8) acc = fmaf(d2[c], __ldg(w2 + (long long)c * S + j), acc);
        part[warp][lane] = acc;
        __syncthreads();
        if (warp == 0 && j < S) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += part[w][lane];
            dp1[(long long)b * S + j] = t * swish_gradf_(__ldg(s_pre + (long long)b * S + j));
        }
        __syncthreads();
    }
}

// dmean[b,c] = sum_j dp1[b,j] W1[j,c]: one thread per (b, c), W1 rows read coalesced
__global__ void __launch_bounds__(128) se_bwd_dmean_kernel(const float* __restrict__ dp1, const float* __restrict__ w1,
                                                           float* __restrict__ dmean, int C, int S) {
    extern __shared__ float p1[];                      // dp1[b,:]
    const int b = blockIdx.y;
    for (int j = threadIdx.x; j < S; j += blockDim.x) p1[j] = __ldg(dp1 + (long long)b * S + j);
    __syncthreads();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float acc = 0.f;
    for (int j = 0; j < S; ++j) acc = fmaf(p1[j], __ldg(w1 + (long long)j * C + c), acc);
    dmean[(long long)b * C + c] = acc;
}

// the sums over the batch, one thread per (c, j):  dW2[c,j] += sum_b dp2[b,c] sw[b,j];  dW1[j,c] += sum_b dp1[b,j] mean[b,c];
// db2[c] += sum_b dp2[b,c] (threads with j == 0);  db1[j] += sum_b dp1[b,j] (threads with c == 0).  No atomics: every
// output element has exactly one writer.
__global__ void __launch_bounds__(128) se_bwd_dw_kernel(const float* __restrict__ dp2, const float* __restrict__ dp1,
                                                        const float* __restrict__ mean, const float* __restrict__ s_pre,
                                                        float* __restrict__ dw1, float* __restrict__ db1,
                                                        float* __restrict__ dw2, float* __restrict__ db2, int B, int C, int S) {
    const int j = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a2 = 0.f, a1 = 0.f, sb2 = 0.f, sb1 = 0.f;
    for (int b = 0; b < B; ++b) {
        const float d2 = __ldg(dp2 + (long long)b * C + c);
        const float d1 = __ldg(dp1 + (long long)b * S + j);
        a2 = fmaf(d2, swishf_(__ldg(s_pre + (long long)b * S + j)), a2);
        a1 = fmaf(d1, __ldg(mean + (long long)b * C + c), a1);
        sb2 += d2;
        sb1 += d1;
    }
    dw2[(long long)c * S + j] += a2;
    dw1[(long long)j * C + c] += a1;
    if (j == 0) db2[c] += sb2;
    if (c == 0) db1[j] += sb1;
}

}  // namespace effdet

using namespace effdet;

extern "C" int effdet_se_gate_fwd(const float* mean, const float* w1, const float* b1, const float* w2, const float* b2,
                                  float* s_pre, float* gate, int B, int C, int S, int device, effdet_stream_t stream) {
    EFFDET_REQUIRE(mean && w1 && b1 && w2 && b2 && s_pre && gate, "se_gate_fwd: null tensor");
    EFFDET_REQUIRE(B > 0 && B <= 65535 && C > 0 && S > 0 && (size_t)S * 4 <= 48 * 1024, "se_gate_fwd: bad shape C=%d S=%d", C, S);
    EFFDET_DEVICE(device);
    cudaStream_t st = (cudaStream_t)stream;
    se_squeeze_kernel<<<dim3(cdiv(S, 8), B), 256, 0, st>>>(mean, w1, b1, s_pre, C, S);
    int s = launch_status("se_squeeze_kernel");
    if (s) return s;
    se_excite_kernel<<<dim3(cdiv(C, 256), B), 256, (size_t)S * sizeof(float), st>>>(s_pre, w2, b2, gate, C, S);
    return launch_status("se_excite_kernel");
}

extern "C" int effdet_se_gate_bwd(const float* dgate, const float* mean, const float* s_pre, const float* gate,
                                  const float* w1, const float* w2, float* dmean, float* dw1, float* db1, float* dw2,
                                  float* db2, float* ws, int B, int C, int S, int device, effdet_stream_t stream) {
    EFFDET_REQUIRE(dgate && mean && s_pre && gate && w1 && w2 && dmean && dw1 && db1 && dw2 && db2 && ws,
                   "se_gate_bwd: null tensor");
    EFFDET_REQUIRE(B > 0 && B <= 65535 && C > 0 && S > 0 && S <= 65535 && (size_t)C * 4 <= 200 * 1024 && (size_t)S * 4 <= 48 * 1024,
                   "se_gate_bwd: bad shape");
    EFFDET_DEVICE(device);
    cudaStream_t st = (cudaStream_t)stream;
    float* dp2 = ws;                       // [B,C]
    float* dp1 = ws + (size_t)B * C;       // [B,S]
    const size_t sm1 = (size_t)C * sizeof(float);
    if (sm1 > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(se_bwd_dp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
        if (e != cudaSuccess) return fail(EFFDET_ERR_LAUNCH, "se_gate_bwd: smem opt-in: %s", cudaGetErrorString(e));
    }
    se_bwd_dp_kernel<<<B, 256, sm1, st>>>(dgate, gate, s_pre, w2, dp2, dp1, C, S);
    int s = launch_status("se_bwd_dp_kernel");
    if (s) return s;
    se_bwd_dmean_kernel<<<dim3(cdiv(C, 128), B), 128, (size_t)S * sizeof(float), st>>>(dp1, w1, dmean, C, S);
    s = launch_status("se_bwd_dmean_kernel");
    if (s) return s;
    se_bwd_dw_kernel<<<dim3(cdiv(C, 128), S), 128, 0, st>>>(dp2, dp1, mean, s_pre, dw1, db1, dw2, db2, B, C, S);
    return launch_status("se_bwd_dw_kernel");
}
