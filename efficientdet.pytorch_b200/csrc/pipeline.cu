// The steps either side of the hot path, on the device:
//   input side  : Normalizer + Augmenter (horizontal flip) + zero-pad to the common size + collater + `.cuda().float()`
//                 datasets/augmentation.py:69-91,111-150, train.py:105-106  (cv2.resize is NOT reproduced: images enter at
//                 their final resolution; everything after the decode is otherwise done here)
//   output side : score threshold + top-`max_detections` by score + per-class split of eval.py:108-128
// Arithmetic follows the reference's dtypes exactly so the results are bit-identical to NumPy:
//   Normalizer computes (img.astype(float32) - mean[float64]) / std[float64] in float64 and train.py casts to float32 on
//   the device; collater writes float64 annotations into a float32 tensor; eval.py divides float32 boxes by float32(scale).
#include "common.cuh"

namespace effdet {

// one thread per output pixel (x fastest): reads 3 interleaved bytes, writes the three channel planes
__global__ void __launch_bounds__(256) normalize_pad_kernel(const uint8_t* __restrict__ pix, const int64_t* __restrict__ offs,
                                                            const int32_t* __restrict__ hw, const uint8_t* __restrict__ flip,
                                                            float* __restrict__ out, int S, double m0, double m1, double m2,
                                                            double s0, double s1, double s2) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= S) return;
    const int h = hw[2 * b], w = hw[2 * b + 1];
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;                              // np.zeros((S, S, 3)) padding
    if (y < h && x < w) {
        const int sx = (flip && flip[b]) ? (w - 1 - x) : x;          // image[:, ::-1, :]
        const uint8_t* p = pix + offs[b] + ((long long)y * w + sx) * 3;
        // (float32(u8) - mean) / std in float64 (exact subtraction, IEEE division), then the .float() of train.py:105
        v0 = (float)__ddiv_rn(__dsub_rn((double)(float)p[0], m0), s0);
        v1 = (float)__ddiv_rn(__dsub_rn((double)(float)p[1], m1), s1);
        v2 = (float)__ddiv_rn(__dsub_rn((double)(float)p[2], m2), s2);
    }
    const long long plane = (long long)S * S;
    float* o = out + (long long)b * 3 * plane + (long long)y * S + x;
    o[0] = v0;
    o[plane] = v1;
    o[2 * plane] = v2;
}

// annotations: rows [n_b, 5] float64 (x1, y1, x2, y2, label) concatenated over the batch -> float32 [B, G, 5], -1 padded.
// Resizer scales the box by `scale` (float64), Augmenter mirrors x about the image width BEFORE the resize.
__global__ void collate_annots_kernel(const double* __restrict__ rows, const int32_t* __restrict__ row_off,
                                      const double* __restrict__ scale, const uint8_t* __restrict__ flip,
                                      const int32_t* __restrict__ width, float* __restrict__ out, int B, int G) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * G) return;
    const int b = i / G, g = i - b * G;
    const int n = row_off[b + 1] - row_off[b];
    float* o = out + (long long)i * 5;
    if (g >= n) {
        o[0] = o[1] = o[2] = o[3] = o[4] = -1.f;
        return;
    }
    const double* r = rows + (long long)(row_off[b] + g) * 5;
    double x1 = r[0], y1 = r[1], x2 = r[2], y2 = r[3];
    if (flip && flip[b]) {                                            // annots[:, 0] = cols - x2 ; annots[:, 2] = cols - x1
        const double cols = (double)width[b];
        const double nx1 = __dsub_rn(cols, x2), nx2 = __dsub_rn(cols, x1);
        x1 = nx1;
        x2 = nx2;
    }
    const double sc = scale ? scale[b] : 1.0;                         // annots[:, :4] *= scale
    o[0] = (float)__dmul_rn(x1, sc);
    o[1] = (float)__dmul_rn(y1, sc);
    o[2] = (float)__dmul_rn(x2, sc);
    o[3] = (float)__dmul_rn(y2, sc);
    o[4] = (float)r[4];
}

// eval.py:108-128 for one image.  One CTA.  Input rows are the model's detections (any order); a row is selected when
// score > thr and its rank by descending score (ties: lower input index first) is < max_det; selected rows are grouped by
// label (ascending), in score order inside a label -- exactly the per-label arrays eval.py builds.
__global__ void __launch_bounds__(256) eval_select_kernel(const float* __restrict__ scores, const int64_t* __restrict__ labels,
                                                          const float* __restrict__ boxes, int n, float scale, float thr,
                                                          int max_det, int num_classes, float* __restrict__ out_dets,
                                                          int32_t* __restrict__ out_labels, int32_t* __restrict__ class_off,
                                                          int32_t* __restrict__ count) {
    extern __shared__ int32_t sm[];
    int32_t* sel = sm;                    // [max_det] input index of the detection with rank r
    int32_t* cls_cnt = sm + max_det;      // [num_classes + 1]
    __shared__ int nsel;
    for (int i = threadIdx.x; i < max_det; i += blockDim.x) sel[i] = -1;
    for (int i = threadIdx.x; i <= num_classes; i += blockDim.x) cls_cnt[i] = 0;
    if (threadIdx.x == 0) nsel = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float s = scores[i];
        if (!(s > thr)) continue;
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const float t = scores[j];
            rank += (t > thr) && (t > s || (t == s && j < i));
        }
        if (rank < max_det) {
            sel[rank] = i;
            atomicAdd(&nsel, 1);
        }
    }
    __syncthreads();
    const int m = nsel;
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
        const int lab = (int)labels[sel[r]];
        if (lab >= 0 && lab < num_classes) atomicAdd(&cls_cnt[lab + 1], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int c = 0; c < num_classes; ++c) cls_cnt[c + 1] += cls_cnt[c];
        count[0] = m;
    }
    __syncthreads();
    for (int c = threadIdx.x; c <= num_classes; c += blockDim.x) class_off[c] = cls_cnt[c];
    // stable placement: position of rank r inside its label group = number of lower ranks with the same label
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
        const int i = sel[r];
        const int lab = (int)labels[i];
        if (lab < 0 || lab >= num_classes) continue;
        int pos = 0;
        for (int q = 0; q < r; ++q) pos += ((int)labels[sel[q]] == lab);
        const int dst = cls_cnt[lab] + pos;
        const float4 bx = *reinterpret_cast<const float4*>(boxes + (long long)i * 4);
        float* o = out_dets + (long long)dst * 5;
        o[0] = __fdiv_rn(bx.x, scale);                                // boxes /= scale   (float32 array / float32 scalar)
        o[1] = __fdiv_rn(bx.y, scale);
        o[2] = __fdiv_rn(bx.z, scale);
        o[3] = __fdiv_rn(bx.w, scale);
        o[4] = scores[i];
        out_labels[dst] = lab;
    }
}

}  // namespace effdet

using namespace effdet;

extern "C" int effdet_normalize_pad(const uint8_t* pixels, const int64_t* offsets, const int32_t* hw, const uint8_t* flip,
                                    float* out_nchw, int B, int S, const double* mean3, const double* std3, int device,
                                    effdet_stream_t stream) {
    EFFDET_REQUIRE(pixels && offsets && hw && out_nchw && mean3 && std3 && B > 0 && B <= 65535 && S > 0 && S <= 65535,
                   "normalize_pad: bad arguments");
    EFFDET_DEVICE(device);
    dim3 grid(cdiv(S, 256), S, B);
    normalize_pad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pixels, offsets, hw, flip, out_nchw, S, mean3[0], mean3[1], mean3[2],
                                                                std3[0], std3[1], std3[2]);
    return launch_status("normalize_pad_kernel");
}

extern "C" int effdet_collate_annots(const double* rows, const int32_t* row_off, const double* scale, const uint8_t* flip,
                                     const int32_t* width, float* out, int B, int G, int device, effdet_stream_t stream) {
    EFFDET_REQUIRE(row_off && out && B > 0 && G > 0 && (!flip || width), "collate_annots: bad arguments");
    EFFDET_DEVICE(device);
    collate_annots_kernel<<<cdiv((long long)B * G, 128), 128, 0, (cudaStream_t)stream>>>(rows, row_off, scale, flip, width, out, B, G);
    return launch_status("collate_annots_kernel");
}

extern "C" int effdet_eval_select(const float* scores, const int64_t* labels, const float* boxes, int n, float scale,
                                  float score_threshold, int max_det, int num_classes, float* out_dets, int32_t* out_labels,
                                  int32_t* class_offsets, int32_t* count, int device, effdet_stream_t stream) {
    EFFDET_REQUIRE(out_dets && out_labels && class_offsets && count && n >= 0 && max_det > 0 && num_classes > 0 &&
                       (n == 0 || (scores && labels && boxes)),
                   "eval_select: bad arguments");
    EFFDET_REQUIRE((size_t)(max_det + num_classes + 1) * 4 <= 48 * 1024, "eval_select: max_det + num_classes too large");
    EFFDET_REQUIRE(aligned16(boxes), "eval_select: boxes must be 16-byte aligned");
    EFFDET_REQUIRE(scale > 0.f, "eval_select: scale must be positive");
    EFFDET_DEVICE(device);
    eval_select_kernel<<<1, 256, (size_t)(max_det + num_classes + 1) * 4, (cudaStream_t)stream>>>(
        scores, labels, boxes, n, scale, score_threshold, max_det, num_classes, out_dets, out_labels, class_offsets, count);
    return launch_status("eval_select_kernel");
}
