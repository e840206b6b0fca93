// Inference post-processing on the device: box decode + clip + per-anchor class max + score
// threshold + sort + greedy NMS, for image 0 (the reference is batch-1 at inference).
// Reference: models/module.py:24-49 (BBoxTransform), :57-67 (ClipBoxes),
//            models/efficientdet.py:70-86 (threshold, nms, gather), torchvision.ops.nms.
// Bit-exactness rules (the keep-set must equal torchvision's): fp32 IoU from separately rounded
// ops (no FMA contraction), suppress iff IoU > thr compared in double, candidates ordered by
// score descending with ties broken by lower anchor index (== stable sort of the masked list).
#include "common.cuh"

namespace effdet {

__device__ __forceinline__ uint32_t float_order(float f) {  // monotone float -> uint32
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// one warp per anchor (coalesced over classes); anchors >= A emit sentinel keys
__global__ void __launch_bounds__(256) detect_candidates_kernel(const float* __restrict__ cls, const float* __restrict__ reg,
                                                                const float* __restrict__ anchors, float* __restrict__ boxes,
                                                                float* __restrict__ scores, int32_t* __restrict__ classes,
                                                                uint64_t* __restrict__ keys, int32_t* __restrict__ count,
                                                                int A, int K, int npad, float img_w, float img_h,
                                                                float threshold) {
    const int lane = threadIdx.x & 31;
    const int a = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (a >= npad) return;
    if (a >= A) {
        if (lane == 0) keys[a] = ~0ull;
        return;
    }
    float best = -INFINITY;
    int arg = 0x7fffffff;
    for (int k = lane; k < K; k += 32) {
        const float v = __ldg(cls + (long long)a * K + k);
        if (v > best) { best = v; arg = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ov > best || (ov == best && oa < arg)) { best = ov; arg = oa; }
    }
    if (lane != 0) return;
    const float4 an = ldg4(anchors + (long long)a * 4);
    const float4 d = ldg4(reg + (long long)a * 4);
    const float w = __fsub_rn(an.z, an.x), h = __fsub_rn(an.w, an.y);
    const float cx = __fadd_rn(an.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(an.y, __fmul_rn(0.5f, h));
    const float dx = __fadd_rn(__fmul_rn(d.x, 0.1f), 0.f), dy = __fadd_rn(__fmul_rn(d.y, 0.1f), 0.f);
    const float dw = __fadd_rn(__fmul_rn(d.z, 0.2f), 0.f), dh = __fadd_rn(__fmul_rn(d.w, 0.2f), 0.f);
    const float pcx = __fadd_rn(cx, __fmul_rn(dx, w)), pcy = __fadd_rn(cy, __fmul_rn(dy, h));
    const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
    float x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw)), y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
    float x2 = __fadd_rn(pcx, __fmul_rn(0.5f, pw)), y2 = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
    x1 = fmaxf(x1, 0.f); y1 = fmaxf(y1, 0.f);
    x2 = fminf(x2, img_w); y2 = fminf(y2, img_h);
    st4(boxes + (long long)a * 4, make_float4(x1, y1, x2, y2));
    scores[a] = best;
    classes[a] = arg;
    if (best > threshold) {
        keys[a] = ((uint64_t)(~float_order(best)) << 32) | (uint32_t)a;
        atomicAdd(count, 1);
    } else {
        keys[a] = ~0ull;
    }
}

// ---- bitonic sort of 64-bit keys (ascending); n is a power of two >= 2048 or handled as one chunk ----
constexpr int kChunk = 2048;

__device__ __forceinline__ void cmpswap(uint64_t& a, uint64_t& b, bool asc) {
    if ((a > b) == asc) { const uint64_t t = a; a = b; b = t; }
}

// sorts every aligned chunk (stages k = 2 .. chunk) with the direction given by the global index
__global__ void __launch_bounds__(1024) bitonic_chunk_sort_kernel(uint64_t* __restrict__ keys, int chunk) {
    extern __shared__ uint64_t sk[];
    const int base = blockIdx.x * chunk;
    for (int i = threadIdx.x; i < chunk; i += blockDim.x) sk[i] = keys[base + i];
    __syncthreads();
    for (int k = 2; k <= chunk; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < chunk / 2; t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const bool asc = (((base + i) & k) == 0);
                cmpswap(sk[i], sk[i | j], asc);
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < chunk; i += blockDim.x) keys[base + i] = sk[i];
}

__global__ void __launch_bounds__(256) bitonic_global_step_kernel(uint64_t* __restrict__ keys, int n, int k, int j) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n / 2) return;
    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    const bool asc = ((i & k) == 0);
    uint64_t a = keys[i], b = keys[i | j];
    if ((a > b) == asc) { keys[i] = b; keys[i | j] = a; }
}

// finishes stage k inside each chunk: j = chunk/2 .. 1
__global__ void __launch_bounds__(1024) bitonic_chunk_merge_kernel(uint64_t* __restrict__ keys, int chunk, int k) {
    extern __shared__ uint64_t sk[];
    const int base = blockIdx.x * chunk;
    for (int i = threadIdx.x; i < chunk; i += blockDim.x) sk[i] = keys[base + i];
    __syncthreads();
    for (int j = chunk >> 1; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < chunk / 2; t += blockDim.x) {
            const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
            const bool asc = (((base + i) & k) == 0);
            cmpswap(sk[i], sk[i | j], asc);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < chunk; i += blockDim.x) keys[base + i] = sk[i];
}

__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, double thr) {
    const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
    const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
    const float width = fmaxf(__fsub_rn(right, left), 0.f), height = fmaxf(__fsub_rn(bottom, top), 0.f);
    // disjoint boxes (the overwhelming majority of pairs): inter == 0 -> IoU is 0 (or 0/0 = NaN): never > thr for thr >= 0,
    // so the exact division below is skipped; the result is unchanged
    if ((width <= 0.f || height <= 0.f) && thr >= 0.0) return false;
    const float inter = __fmul_rn(width, height);
    const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float sb = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter));
    return (double)ovr > thr;
}

// mask[i][cb] bit j set  <=>  sorted box (64*cb + j) is suppressed by sorted box i  (upper triangle only)
__global__ void __launch_bounds__(64) nms_mask_kernel(const float* __restrict__ boxes, const uint64_t* __restrict__ keys,
                                                      int n, double thr, uint64_t* __restrict__ mask) {
    const int rb = blockIdx.y, cb = blockIdx.x;
    if (rb > cb) return;
    const int col_blocks = (n + 63) / 64;
    __shared__ float4 cbx[64];
    const int cj = cb * 64 + threadIdx.x;
    if (cj < n) cbx[threadIdx.x] = ldg4(boxes + (long long)(uint32_t)keys[cj] * 4);
    __syncthreads();
    const int i = rb * 64 + threadIdx.x;
    if (i >= n) return;
    const float4 me = ldg4(boxes + (long long)(uint32_t)keys[i] * 4);
    const int csize = min(64, n - cb * 64);
    uint64_t bits = 0;
    for (int j = (rb == cb ? threadIdx.x + 1 : 0); j < csize; ++j)
        if (iou_gt(me, cbx[j], thr)) bits |= 1ull << j;
    mask[(long long)i * col_blocks + cb] = bits;
}

// Greedy scan over the sorted candidates, one CTA; the `removed` bitmap lives in shared memory.  Candidates are
// consumed 64 at a time: warp 0 resolves the block's internal dependencies from the 64 diagonal mask words (a
// sequential walk over 64 bits, registers + shuffles only), then every thread ORs the rows of the block's survivors
// into its slice of `removed`.  Two CTA barriers per 64 candidates instead of two per kept box: the scan used to be
// 255 ms for the 55 k survivors of a random-weight D7 image.  Result identical to the one-at-a-time scan.
__global__ void __launch_bounds__(1024) nms_scan_kernel(const uint64_t* __restrict__ mask, const uint64_t* __restrict__ keys,
                                                        int n, int32_t* __restrict__ keep_idx, int32_t* __restrict__ nkeep) {
    extern __shared__ uint64_t removed[];
    __shared__ uint64_t keep_bits;
    const int col_blocks = (n + 63) / 64;
    const int lane = threadIdx.x & 31;
    for (int j = threadIdx.x; j < col_blocks; j += blockDim.x) removed[j] = 0;
    __syncthreads();
    int kept = 0;
    for (int nb = 0; nb < col_blocks; ++nb) {
        if (threadIdx.x < 32) {
            const int i0 = nb * 64 + lane, i1 = i0 + 32;
            const uint64_t d0 = i0 < n ? mask[(long long)i0 * col_blocks + nb] : 0ull;     // bits j > i inside the block
            const uint64_t d1 = i1 < n ? mask[(long long)i1 * col_blocks + nb] : 0ull;
            uint64_t rem = removed[nb], kb = 0;
            const int valid = min(64, n - nb * 64);
#pragma unroll 1
            for (int b = 0; b < valid; ++b) {
                const uint64_t row = __shfl_sync(0xffffffffu, b < 32 ? d0 : d1, b & 31);
                if (!((rem >> b) & 1ull)) {
                    kb |= 1ull << b;
                    rem |= row;
                }
            }
            if (lane == 0) keep_bits = kb;
        }
        __syncthreads();
        const uint64_t kb = keep_bits;
        if (threadIdx.x < 64 && ((kb >> threadIdx.x) & 1ull))
            keep_idx[kept + __popcll(kb & ((1ull << threadIdx.x) - 1ull))] = (int32_t)(uint32_t)keys[nb * 64 + threadIdx.x];
        for (int j = nb + 1 + threadIdx.x; j < col_blocks; j += blockDim.x) {
            uint64_t acc = 0, bits = kb;
            while (bits) {
                const int b = __ffsll((long long)bits) - 1;
                bits &= bits - 1;
                acc |= mask[(long long)(nb * 64 + b) * col_blocks + j];
            }
            removed[j] |= acc;
        }
        kept += __popcll(kb);
        __syncthreads();
    }
    if (threadIdx.x == 0) nkeep[0] = kept;
}

__global__ void gather_detections_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                         const int32_t* __restrict__ classes, const int32_t* __restrict__ keep_idx, int nkeep,
                                         float* __restrict__ out_scores, long long* __restrict__ out_classes,
                                         float* __restrict__ out_boxes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nkeep) return;
    const int a = keep_idx[i];
    out_scores[i] = scores[a];
    out_classes[i] = (long long)classes[a];
    st4(out_boxes + (long long)i * 4, ldg4(boxes + (long long)a * 4));
}

}  // namespace effdet

using namespace effdet;

extern "C" int effdet_detect_candidates(const float* cls, const float* reg, const float* anchors, float* boxes,
                                        float* scores, int32_t* classes, uint64_t* keys, int32_t* count, int A, int K,
                                        int npad, float img_w, float img_h, float threshold, int device,
                                        effdet_stream_t stream) {
    EFFDET_REQUIRE(cls && reg && anchors && boxes && scores && classes && keys && count, "detect_candidates: null tensor");
    EFFDET_REQUIRE(A > 0 && K > 0 && npad >= A && (npad & (npad - 1)) == 0, "detect_candidates: npad must be a power of two >= A");
    EFFDET_REQUIRE(aligned16(reg) && aligned16(anchors) && aligned16(boxes), "detect_candidates: alignment");
    EFFDET_DEVICE(device);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return fail(EFFDET_ERR_LAUNCH, "detect_candidates: memset: %s", cudaGetErrorString(e));
    detect_candidates_kernel<<<cdiv(npad, 8), 256, 0, st>>>(cls, reg, anchors, boxes, scores, classes, keys, count, A, K, npad,
                                                           img_w, img_h, threshold);
    int s = launch_status("detect_candidates_kernel");
    if (s) return s;
    // sort ascending: best candidates first, sentinels last
    const int chunk = npad < kChunk ? npad : kChunk;
    if (chunk >= 2) {
        bitonic_chunk_sort_kernel<<<npad / chunk, 1024, chunk * sizeof(uint64_t), st>>>(keys, chunk);
        if ((s = launch_status("bitonic_chunk_sort_kernel"))) return s;
    }
    for (int k = chunk * 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j >= chunk; j >>= 1) {
            bitonic_global_step_kernel<<<cdiv(npad / 2, 256), 256, 0, st>>>(keys, npad, k, j);
            if ((s = launch_status("bitonic_global_step_kernel"))) return s;
        }
        bitonic_chunk_merge_kernel<<<npad / chunk, 1024, chunk * sizeof(uint64_t), st>>>(keys, chunk, k);
        if ((s = launch_status("bitonic_chunk_merge_kernel"))) return s;
    }
    return EFFDET_OK;
}

extern "C" int effdet_nms(const float* boxes, const uint64_t* keys, int n, double iou_threshold, uint64_t* mask_ws,
                          int32_t* keep_idx, int32_t* nkeep, int device, effdet_stream_t stream) {
    EFFDET_REQUIRE(boxes && keys && mask_ws && keep_idx && nkeep && n > 0, "nms: bad arguments");
    EFFDET_DEVICE(device);
    cudaStream_t st = (cudaStream_t)stream;
    const int col_blocks = cdiv(n, 64);
    EFFDET_REQUIRE(col_blocks <= 65535, "nms: too many candidates (%d)", n);
    const size_t smem = (size_t)col_blocks * sizeof(uint64_t);
    EFFDET_REQUIRE(smem <= 200 * 1024, "nms: too many candidates for the scan bitmap (%d)", n);
    nms_mask_kernel<<<dim3(col_blocks, col_blocks), 64, 0, st>>>(boxes, keys, n, iou_threshold, mask_ws);
    int s = launch_status("nms_mask_kernel");
    if (s) return s;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(EFFDET_ERR_LAUNCH, "nms: smem opt-in: %s", cudaGetErrorString(e));
    }
    nms_scan_kernel<<<1, 1024, smem, st>>>(mask_ws, keys, n, keep_idx, nkeep);
    return launch_status("nms_scan_kernel");
}

extern "C" int effdet_gather_detections(const float* boxes, const float* scores, const int32_t* classes,
                                        const int32_t* keep_idx, int nkeep, float* out_scores, int64_t* out_classes,
                                        float* out_boxes, int device, effdet_stream_t stream) {
    EFFDET_REQUIRE(boxes && scores && classes && keep_idx && out_scores && out_classes && out_boxes && nkeep > 0,
                   "gather_detections: bad arguments");
    EFFDET_REQUIRE(aligned16(boxes) && aligned16(out_boxes), "gather_detections: alignment");
    EFFDET_DEVICE(device);
    gather_detections_kernel<<<cdiv(nkeep, 256), 256, 0, (cudaStream_t)stream>>>(boxes, scores, classes, keep_idx, nkeep,
                                                                               out_scores, (long long*)out_classes,
                                                                               out_boxes);
    return launch_status("gather_detections_kernel");
}
