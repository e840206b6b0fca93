// EfficientNet stem: 3x3 stride-2 conv (3 -> C0) straight from the NCHW image, static TF-"SAME"
// pad (left,right,top,bottom) = (0,1,0,1), fused eval-BN affine + swish, NHWC output.
// Reference: models/efficientnet.py:140-143,193 ; pad rule models/utils.py:126-149.
// HBM-bound (AI ~ 10 FLOP/B): one pass over the image, one write of z (kept for backward) and y.
#include "common.cuh"

namespace effdet {

__global__ void __launch_bounds__(256) stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                       float* __restrict__ z, float* __restrict__ y, int B, int H, int W,
                                                       int C0, int Ho, int Wo) {
    extern __shared__ __align__(16) float ws[];  // [27][C0], tap-major
    for (int i = threadIdx.x; i < 27 * C0; i += blockDim.x) {
        const int co = i % C0, tap = i / C0;  // tap = ci*9 + ky*3 + kx
        ws[i] = __ldg(w + co * 27 + tap);
    }
    __syncthreads();
    const int cvecs = C0 / 4;
    const long long total = (long long)B * Ho * Wo * cvecs;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % cvecs);
    long long pix = idx / cvecs;
    const int ox = (int)(pix % Wo);
    pix /= Wo;
    const int oy = (int)(pix % Ho);
    const int b = (int)(pix / Ho);
    float4 acc = f4zero();
    const float* xb = x + (long long)b * 3 * H * W;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = 2 * oy + ky;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = 2 * ox + kx;
                float v = 0.f;
                if (iy < H && ix < W) v = __ldg(xb + ((long long)ci * H + iy) * W + ix);
                const float4 wv = *reinterpret_cast<const float4*>(&ws[(ci * 9 + ky * 3 + kx) * C0 + cv * 4]);
                acc = f4fma(make_float4(v, v, v, v), wv, acc);
            }
        }
    }
    const long long o = (((long long)b * Ho + oy) * Wo + ox) * C0 + cv * 4;
    st4(z + o, acc);
    if (y == nullptr) return;
    float4 u = f4fma(acc, ldg4(scale + cv * 4), ldg4(shift + cv * 4));
    st4(y + o, make_float4(swishf_(u.x), swishf_(u.y), swishf_(u.z), swishf_(u.w)));
}

// One thread per output pixel, all C0 = 4*NV channels in registers: the 27 input taps are fetched once per pixel
// (not once per 4-channel group) and the weights are warp-broadcast 128-bit shared loads.
template <int NV>
__global__ void __launch_bounds__(128) stem_fwd_px_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ scale, const float* __restrict__ shift,
                                                          float* __restrict__ z, float* __restrict__ y, int B, int H, int W,
                                                          int Ho, int Wo) {
    constexpr int C0 = NV * 4;
    __shared__ __align__(16) float ws[27 * C0];
    __shared__ __align__(16) float sc_s[C0];
    __shared__ __align__(16) float sh_s[C0];
    for (int i = threadIdx.x; i < 27 * C0; i += blockDim.x) {
        const int co = i % C0, tap = i / C0;
        ws[i] = __ldg(w + co * 27 + tap);
    }
    for (int i = threadIdx.x; i < C0; i += blockDim.x) { sc_s[i] = __ldg(scale + i); sh_s[i] = __ldg(shift + i); }
    __syncthreads();
    const long long npix = (long long)B * Ho * Wo;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const int ox = (int)(pix % Wo);
    const long long r = pix / Wo;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    float v[27];
    const float* xb = x + (long long)b * 3 * H * W;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int iy = 2 * oy + ky, ix = 2 * ox + kx;
                v[ci * 9 + ky * 3 + kx] = (iy < H && ix < W) ? __ldg(xb + ((long long)ci * H + iy) * W + ix) : 0.f;
            }
    float4 acc[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] = f4zero();
#pragma unroll
    for (int tap = 0; tap < 27; ++tap) {
        const float4 xv = make_float4(v[tap], v[tap], v[tap], v[tap]);
#pragma unroll
        for (int j = 0; j < NV; ++j) acc[j] = f4fma(xv, *reinterpret_cast<const float4*>(&ws[tap * C0 + j * 4]), acc[j]);
    }
    float* zo = z + pix * C0;
    float* yo = y + pix * C0;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        st4(zo + j * 4, acc[j]);
        if (y == nullptr) continue;
        const float4 u = f4fma(acc[j], *reinterpret_cast<const float4*>(&sc_s[j * 4]), *reinterpret_cast<const float4*>(&sh_s[j * 4]));
        st4(yo + j * 4, make_float4(swishf_(u.x), swishf_(u.y), swishf_(u.z), swishf_(u.w)));
    }
}

constexpr int kStemP = 64;     // pixels staged per iteration
constexpr int kStemMaxT = 2;   // taps per thread upper bound: 27 <= 2 * (256 / (C0/4)) for C0 <= 72

// dw[co][tap] += sum_pixels x[pixel + tap] * dz[pixel][co]  -- a 27 x C0 GEMM over ~2 M pixels, shared-memory bound:
// a thread owns 4 output channels (one 128-bit shared load of dz per pixel) and 1-2 taps (32-bit broadcast loads of
// the im2col row), i.e. 4 FMAs per 2 shared loads; the previous mapping (1 channel x 4 taps) paid 2 loads per FMA
// and ran at 0.07 of the HBM roofline.
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                                         float* __restrict__ dw, int B, int H, int W, int C0, int Ho,
                                                         int Wo) {
    extern __shared__ __align__(16) float sm[];
    float* xs = sm;                    // [kStemP][28]  (27 taps, padded)
    float* ds = sm + kStemP * 28;      // [kStemP][C0]
    const int t = threadIdx.x;
    const int cvs = C0 / 4;            // channel vectors
    const int ntg = 256 / cvs;         // tap groups
    const int cv = t % cvs, tg = t / cvs;
    const bool worker = tg < ntg;
    float4 acc[kStemMaxT];
#pragma unroll
    for (int j = 0; j < kStemMaxT; ++j) acc[j] = f4zero();
    const long long npix = (long long)B * Ho * Wo;
    for (long long p0 = (long long)blockIdx.x * kStemP; p0 < npix; p0 += (long long)gridDim.x * kStemP) {
        for (int i = t; i < kStemP * 27; i += 256) {
            const int pp = i / 27, tap = i - pp * 27;
            const long long pix = p0 + pp;
            float v = 0.f;
            if (pix < npix) {
                const int ox = (int)(pix % Wo);
                const long long r = pix / Wo;
                const int oy = (int)(r % Ho);
                const int b = (int)(r / Ho);
                const int ci = tap / 9, ky = (tap % 9) / 3, kx = tap % 3;
                const int iy = 2 * oy + ky, ix = 2 * ox + kx;
                if (iy < H && ix < W) v = __ldg(x + (((long long)b * 3 + ci) * H + iy) * W + ix);
            }
            xs[pp * 28 + tap] = v;
        }
        for (int i = t; i < kStemP * C0 / 4; i += 256) {
            const int pp = i / (C0 / 4), c4 = i - pp * (C0 / 4);
            const long long pix = p0 + pp;
            float4 v = f4zero();
            if (pix < npix) v = ldg4(dz + pix * C0 + c4 * 4);
            *reinterpret_cast<float4*>(&ds[pp * C0 + c4 * 4]) = v;
        }
        __syncthreads();
        if (worker) {
#pragma unroll 8
            for (int pp = 0; pp < kStemP; ++pp) {
                const float4 g = *reinterpret_cast<const float4*>(&ds[pp * C0 + cv * 4]);
#pragma unroll
                for (int j = 0; j < kStemMaxT; ++j) {
                    const int tap = tg + j * ntg;
                    if (tap < 27) acc[j] = f4fma(make_float4(xs[pp * 28 + tap], xs[pp * 28 + tap], xs[pp * 28 + tap], xs[pp * 28 + tap]), g, acc[j]);
                }
            }
        }
        __syncthreads();
    }
    if (worker) {
#pragma unroll
        for (int j = 0; j < kStemMaxT; ++j) {
            const int tap = tg + j * ntg;
            if (tap < 27) {
                float* o = dw + (cv * 4) * 27 + tap;
                atomicAdd(o, acc[j].x); atomicAdd(o + 27, acc[j].y); atomicAdd(o + 54, acc[j].z); atomicAdd(o + 81, acc[j].w);
            }
        }
    }
}

}  // namespace effdet

using namespace effdet;

extern "C" int effdet_stem_fwd(const float* x_nchw, const float* w_oihw, const float* scale, const float* shift,
                               float* z, float* y, int B, int H, int W, int C0, int device, effdet_stream_t stream) {
    EFFDET_REQUIRE(x_nchw && w_oihw && scale && shift && z, "stem_fwd: null tensor");
    EFFDET_REQUIRE(C0 % 4 == 0 && C0 > 0 && C0 <= 256, "stem_fwd: C0=%d unsupported", C0);
    EFFDET_REQUIRE(B > 0 && H >= 2 && W >= 2, "stem_fwd: bad shape");
    EFFDET_REQUIRE(aligned16(z) && aligned16(y) && aligned16(scale) && aligned16(shift), "stem_fwd: alignment");
    EFFDET_DEVICE(device);
    const int Ho = (H + 1 - 3) / 2 + 1, Wo = (W + 1 - 3) / 2 + 1;
    const long long npix = (long long)B * Ho * Wo;
    cudaStream_t st = (cudaStream_t)stream;
    switch (C0) {      // stem widths of EfficientNet-B0..B7: 32, 32, 32, 40, 48, 48, 56, 64
        case 32: stem_fwd_px_kernel<8><<<cdiv(npix, 128), 128, 0, st>>>(x_nchw, w_oihw, scale, shift, z, y, B, H, W, Ho, Wo); break;
        case 40: stem_fwd_px_kernel<10><<<cdiv(npix, 128), 128, 0, st>>>(x_nchw, w_oihw, scale, shift, z, y, B, H, W, Ho, Wo); break;
        case 48: stem_fwd_px_kernel<12><<<cdiv(npix, 128), 128, 0, st>>>(x_nchw, w_oihw, scale, shift, z, y, B, H, W, Ho, Wo); break;
        case 56: stem_fwd_px_kernel<14><<<cdiv(npix, 128), 128, 0, st>>>(x_nchw, w_oihw, scale, shift, z, y, B, H, W, Ho, Wo); break;
        case 64: stem_fwd_px_kernel<16><<<cdiv(npix, 128), 128, 0, st>>>(x_nchw, w_oihw, scale, shift, z, y, B, H, W, Ho, Wo); break;
        default: {
            const long long total = npix * (C0 / 4);
            stem_fwd_kernel<<<cdiv(total, 256), 256, 27 * C0 * sizeof(float), st>>>(x_nchw, w_oihw, scale, shift, z, y, B, H, W, C0,
                                                                                     Ho, Wo);
        }
    }
    return launch_status("stem_fwd_kernel");
}

extern "C" int effdet_stem_wgrad(const float* x_nchw, const float* dz, float* dw_oihw, int B, int H, int W, int C0,
                                 int device, effdet_stream_t stream) {
    EFFDET_REQUIRE(x_nchw && dz && dw_oihw, "stem_wgrad: null tensor");
    EFFDET_REQUIRE(C0 % 4 == 0 && C0 >= 8 && C0 <= 64, "stem_wgrad: C0=%d unsupported (8..64)", C0);
    EFFDET_REQUIRE(aligned16(dz), "stem_wgrad: alignment");
    EFFDET_DEVICE(device);
    const int Ho = (H + 1 - 3) / 2 + 1, Wo = (W + 1 - 3) / 2 + 1;
    const long long npix = (long long)B * Ho * Wo;
    int blocks = cdiv(npix, kStemP);
    if (blocks > 148 * 8) blocks = 148 * 8;
    const size_t smem = (size_t)kStemP * (28 + C0) * sizeof(float);
    stem_wgrad_kernel<<<blocks, 256, smem, (cudaStream_t)stream>>>(x_nchw, dz, dw_oihw, B, H, W, C0, Ho, Wo);
    return launch_status("stem_wgrad_kernel");
}
