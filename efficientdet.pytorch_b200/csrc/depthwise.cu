// Depthwise k x k convolution of the MBConv block (k in {3,5}, stride in {1,2}), NHWC, with the
// reference's static asymmetric TF-"SAME" padding, fused eval-BN affine + swish epilogue.
// Reference: models/efficientnet.py:52-56,87 ; pads models/utils.py:126-149.
// Pure bandwidth kernels: 128-bit loads along C, each thread produces a strip of 4 outputs so
// every input column is fetched once per (row, strip) instead of once per tap.
#include "common.cuh"

namespace effdet {

constexpr int kTW = 4;  // outputs along x per thread

// 16-byte asynchronous global->shared copy; src_bytes = 0 zero-fills the destination (halo / tail)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// BWD = true turns the same strip kernel into the stride-1 data gradient: taps are read 180-degree rotated, the
// epilogue is a plain store (the caller passes dz as x and the mirrored pads)
template <int K, int S, bool BWD>
__global__ void __launch_bounds__(256) dw_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                                     float* __restrict__ z, float* __restrict__ y, int B, int H, int W,
                                                     int C, int pad_t, int pad_l, int Ho, int Wo) {
    const int cvecs = C / 4;
    const int wgroups = (Wo + kTW - 1) / kTW;
    const long long total = (long long)B * Ho * wgroups * cvecs;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % cvecs);
    long long r = idx / cvecs;
    const int og = (int)(r % wgroups);
    r /= wgroups;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const int ox0 = og * kTW;
    constexpr int NCOL = (kTW - 1) * S + K;
    float4 acc[kTW];
#pragma unroll
    for (int i = 0; i < kTW; ++i) acc[i] = f4zero();
    const float* xb = x + (long long)b * H * W * C + cv * 4;
    const float* wc = w + cv * 4;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
        const int iy = oy * S + ky - pad_t;
        if (iy < 0 || iy >= H) continue;
        float4 wrow[K];
#pragma unroll
        for (int kx = 0; kx < K; ++kx)
            wrow[kx] = BWD ? ldg4(wc + ((K - 1 - ky) * K + (K - 1 - kx)) * C) : ldg4(wc + (ky * K + kx) * C);
        const float* xr = xb + (long long)iy * W * C;
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
            const int ix = ox0 * S + j - pad_l;
            float4 v = f4zero();
            if (ix >= 0 && ix < W) v = ldg4(xr + (long long)ix * C);
#pragma unroll
            for (int i = 0; i < kTW; ++i) {
                const int kx = j - i * S;
                if (kx >= 0 && kx < K) acc[i] = f4fma(v, wrow[kx], acc[i]);
            }
        }
    }
    float4 sc = f4zero(), sh = f4zero();
    if (!BWD) { sc = ldg4(scale + cv * 4); sh = ldg4(shift + cv * 4); }
#pragma unroll
    for (int i = 0; i < kTW; ++i) {
        const int ox = ox0 + i;
        if (ox >= Wo) break;
        const long long o = (((long long)b * Ho + oy) * Wo + ox) * C + cv * 4;
        if (BWD) {
            st4(y + o, acc[i]);
        } else {
            st4(z + o, acc[i]);
            const float4 u = f4fma(acc[i], sc, sh);
            st4(y + o, make_float4(swishf_(u.x), swishf_(u.y), swishf_(u.z), swishf_(u.w)));
        }
    }
}

// dx[b,iy,ix,c] = sum_{ky,kx} dz[b,(iy+pt-ky)/S,(ix+pl-kx)/S,c] * w[ky][kx][c]  (where divisible, in range)
template <int K, int S>
__global__ void __launch_bounds__(256) dw_bwd_data_kernel(const float* __restrict__ dz, const float* __restrict__ w,
                                                          float* __restrict__ dx, int B, int H, int W, int C, int pad_t,
                                                          int pad_l, int Ho, int Wo) {
    const int cvecs = C / 4;
    const long long total = (long long)B * H * W * cvecs;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % cvecs);
    long long r = idx / cvecs;
    const int ix = (int)(r % W);
    r /= W;
    const int iy = (int)(r % H);
    const int b = (int)(r / H);
    float4 acc = f4zero();
    const float* gb = dz + (long long)b * Ho * Wo * C + cv * 4;
    const float* wc = w + cv * 4;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
        const int ty = iy + pad_t - ky;
        if (ty < 0 || (ty % S) != 0) continue;
        const int oy = ty / S;
        if (oy >= Ho) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
            const int tx = ix + pad_l - kx;
            if (tx < 0 || (tx % S) != 0) continue;
            const int ox = tx / S;
            if (ox >= Wo) continue;
            acc = f4fma(ldg4(gb + ((long long)oy * Wo + ox) * C), ldg4(wc + (ky * K + kx) * C), acc);
        }
    }
    st4(dx + idx * 4, acc);
}

// dw[c][ky][kx] += sum_{b,oy,ox} x[b,oy*S+ky-pt,ox*S+kx-pl,c] * dz[b,oy,ox,c]
// Each thread owns 4 channels and walks strips of kTW consecutive outputs: per kernel row it fetches the
// (kTW-1)*S+K input columns once and reuses them for all K taps of all kTW outputs (2.4x fewer loads than one
// fetch per tap for k5).  Per-thread partial sums live in registers, one shared-memory reduction + atomics at the end.
template <int K, int S>
__global__ void __launch_bounds__(256) dw_bwd_weight_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                                            float* __restrict__ dw, int B, int H, int W, int C, int pad_t,
                                                            int pad_l, int Ho, int Wo, int rows_per_block) {
    __shared__ float4 red[256];
    constexpr int NCOL = (kTW - 1) * S + K;
    const int cvecs = C / 4;
    const RowPack rp = rowpack(cvecs, blockIdx.y);
    float4 acc[K * K];
#pragma unroll
    for (int i = 0; i < K * K; ++i) acc[i] = f4zero();
    const int wgroups = (Wo + kTW - 1) / kTW;
    const long long ngroups = (long long)B * Ho * wgroups;
    if (rp.active) {
        const long long r_begin = (long long)blockIdx.x * rows_per_block;
        const long long r_end = min(ngroups, r_begin + rows_per_block);
        for (long long gi = r_begin + rp.tr; gi < r_end; gi += rp.rows) {
            const int og = (int)(gi % wgroups);
            const long long q = gi / wgroups;
            const int oy = (int)(q % Ho);
            const int b = (int)(q / Ho);
            const int ox0 = og * kTW;
            float4 g[kTW];
            const float* gp = dz + (((long long)b * Ho + oy) * Wo + ox0) * C + rp.cv * 4;
#pragma unroll
            for (int t = 0; t < kTW; ++t) g[t] = (ox0 + t < Wo) ? ldg4(gp + (long long)t * C) : f4zero();
            const float* xb = x + (long long)b * H * W * C + rp.cv * 4;
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
                const int iy = oy * S + ky - pad_t;
                if (iy < 0 || iy >= H) continue;
                const float* xr = xb + (long long)iy * W * C;
#pragma unroll
                for (int j = 0; j < NCOL; ++j) {
                    const int ix = ox0 * S + j - pad_l;
                    float4 v = f4zero();
                    if (ix >= 0 && ix < W) v = ldg4(xr + (long long)ix * C);
#pragma unroll
                    for (int t = 0; t < kTW; ++t) {
                        const int kx = j - t * S;
                        if (kx >= 0 && kx < K) acc[ky * K + kx] = f4fma(v, g[t], acc[ky * K + kx]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int tap = 0; tap < K * K; ++tap) {
        red[threadIdx.x] = acc[tap];
        __syncthreads();
        if (rp.tr == 0 && rp.cv < cvecs) {
            float4 s = f4zero();
            for (int rr = 0; rr < rp.rows; ++rr) s = f4add(s, red[rr * rp.cvb + rp.tc]);
            float* o = dw + (long long)(rp.cv * 4) * (K * K) + tap;
            atomicAdd(o, s.x); atomicAdd(o + K * K, s.y); atomicAdd(o + 2 * K * K, s.z); atomicAdd(o + 3 * K * K, s.w);
        }
        __syncthreads();
    }
}

// Shared-memory tiled variant of the weight gradient (used for the large feature maps): a CTA stages a
// TH x TW output tile of dz and the matching input tile (with halo) for 32 channels, then every thread owns one
// (tap, 4 channels) pair and sweeps the tile with 128-bit shared loads -- global memory is read once per tile
// instead of once per tap.  CTAs are persistent over tiles; one round of atomics per CTA at the end.
template <int K, int S, int TH, int TWD>
__global__ void __launch_bounds__(256) dw_bwd_weight_tiled_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                                                  float* __restrict__ dw, int B, int H, int W, int C,
                                                                  int pad_t, int pad_l, int Ho, int Wo, int tiles_x,
                                                                  int tiles_y) {
    constexpr int IH = (TH - 1) * S + K, IW = (TWD - 1) * S + K;
    constexpr int KK = K * K;
    constexpr int NPG = 32 / KK;              // pixel groups sharing the 32 (tap, group) slots of a CTA
    extern __shared__ __align__(16) float sm[];
    float4* xs = reinterpret_cast<float4*>(sm);                 // [IH][IW][8]
    float4* gs = xs + IH * IW * 8;                              // [TH][TWD][8]
    const int t = threadIdx.x;
    const int cv_l = t & 7, slot = t >> 3;
    const int tap = slot % KK, pg = slot / KK;
    const bool worker = pg < NPG;
    const int ky = tap / K, kx = tap - ky * K;
    const int cvecs = C / 4;
    const int cv0 = blockIdx.y * 8;
    const bool cv_ok = cv0 + cv_l < cvecs;
    const long long ntiles = (long long)B * tiles_y * tiles_x;
    float4 acc = f4zero();
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tx = (int)(tile % tiles_x);
        const long long q = tile / tiles_x;
        const int ty = (int)(q % tiles_y);
        const int b = (int)(q / tiles_y);
        const int oy0 = ty * TH, ox0 = tx * TWD;
        const int iy0 = oy0 * S - pad_t, ix0 = ox0 * S - pad_l;
        // stage both tiles with cp.async: every copy of the tile is in flight at once (zero fill = padding)
        for (int i = t; i < IH * IW * 8; i += 256) {
            const int c8 = i & 7;
            const int p = i >> 3;
            const int r = p / IW, cc = p - r * IW;
            const int iy = iy0 + r, ix = ix0 + cc;
            const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W && cv0 + c8 < cvecs;
            const float* src = ok ? x + (((long long)b * H + iy) * W + ix) * C + (cv0 + c8) * 4 : x;
            cp_async16(&xs[i], src, ok ? 16 : 0);
        }
        for (int i = t; i < TH * TWD * 8; i += 256) {
            const int c8 = i & 7;
            const int p = i >> 3;
            const int r = p / TWD, cc = p - r * TWD;
            const int oy = oy0 + r, ox = ox0 + cc;
            const bool ok = oy < Ho && ox < Wo && cv0 + c8 < cvecs;
            const float* src = ok ? dz + (((long long)b * Ho + oy) * Wo + ox) * C + (cv0 + c8) * 4 : dz;
            cp_async16(&gs[i], src, ok ? 16 : 0);
        }
        cp_async_wait_all();
        __syncthreads();
        if (worker) {
            for (int p = pg; p < TH * TWD; p += NPG) {
                const int r = p / TWD, cc = p - r * TWD;
                acc = f4fma(xs[((r * S + ky) * IW + cc * S + kx) * 8 + cv_l], gs[p * 8 + cv_l], acc);
            }
        }
        __syncthreads();
    }
    if (worker && cv_ok) {
        float* o = dw + (long long)((cv0 + cv_l) * 4) * KK + tap;
        atomicAdd(o, acc.x); atomicAdd(o + KK, acc.y); atomicAdd(o + 2 * KK, acc.z); atomicAdd(o + 3 * KK, acc.w);
    }
}

__global__ void pack_dw_weight_kernel(const float* __restrict__ w, float* __restrict__ o, int C, int kk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // output index [tap][c]
    if (i >= C * kk) return;
    const int c = i % C, tap = i / C;
    o[i] = __ldg(w + c * kk + tap);
}

}  // namespace effdet

using namespace effdet;

#define DW_DISPATCH(KERNEL, ...)                                                                      \
    if (k == 3 && stride == 1) KERNEL<3, 1> __VA_ARGS__;                                              \
    else if (k == 3 && stride == 2) KERNEL<3, 2> __VA_ARGS__;                                         \
    else if (k == 5 && stride == 1) KERNEL<5, 1> __VA_ARGS__;                                         \
    else KERNEL<5, 2> __VA_ARGS__;

static int dw_check(const char* who, int B, int H, int W, int C, int k, int stride, int pad_t, int pad_l, int Ho, int Wo) {
    EFFDET_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "%s: bad shape (C must be a multiple of 4)", who);
    EFFDET_REQUIRE((k == 3 || k == 5) && (stride == 1 || stride == 2), "%s: k=%d stride=%d unsupported", who, k, stride);
    EFFDET_REQUIRE(pad_t >= 0 && pad_l >= 0 && pad_t < k && pad_l < k && Ho > 0 && Wo > 0, "%s: bad padding/output", who);
    EFFDET_REQUIRE((Ho - 1) * stride + k - pad_t <= H + k && (Wo - 1) * stride + k - pad_l <= W + k, "%s: output too large", who);
    return EFFDET_OK;
}

extern "C" int effdet_dwconv_fwd(const float* x, const float* w_kkc, const float* scale, const float* shift, float* z,
                                 float* y, int B, int H, int W, int C, int k, int stride, int pad_t, int pad_l, int Ho,
                                 int Wo, int device, effdet_stream_t stream) {
    EFFDET_REQUIRE(x && w_kkc && scale && shift && z && y, "dwconv_fwd: null tensor");
    EFFDET_REQUIRE(aligned16(x) && aligned16(w_kkc) && aligned16(scale) && aligned16(shift) && aligned16(z) && aligned16(y),
                   "dwconv_fwd: alignment");
    int s = dw_check("dwconv_fwd", B, H, W, C, k, stride, pad_t, pad_l, Ho, Wo);
    if (s) return s;
    EFFDET_DEVICE(device);
    const long long total = (long long)B * Ho * cdiv(Wo, kTW) * (C / 4);
    cudaStream_t st = (cudaStream_t)stream;
    if (k == 3 && stride == 1) dw_fwd_kernel<3, 1, false><<<cdiv(total, 256), 256, 0, st>>>(x, w_kkc, scale, shift, z, y, B, H, W, C, pad_t, pad_l, Ho, Wo);
    else if (k == 3) dw_fwd_kernel<3, 2, false><<<cdiv(total, 256), 256, 0, st>>>(x, w_kkc, scale, shift, z, y, B, H, W, C, pad_t, pad_l, Ho, Wo);
    else if (stride == 1) dw_fwd_kernel<5, 1, false><<<cdiv(total, 256), 256, 0, st>>>(x, w_kkc, scale, shift, z, y, B, H, W, C, pad_t, pad_l, Ho, Wo);
    else dw_fwd_kernel<5, 2, false><<<cdiv(total, 256), 256, 0, st>>>(x, w_kkc, scale, shift, z, y, B, H, W, C, pad_t, pad_l, Ho, Wo);
    return launch_status("dw_fwd_kernel");
}

extern "C" int effdet_dwconv_bwd_data(const float* dz, const float* w_kkc, float* dx, int B, int H, int W, int C, int k,
                                      int stride, int pad_t, int pad_l, int Ho, int Wo, int device,
                                      effdet_stream_t stream) {
    EFFDET_REQUIRE(dz && w_kkc && dx, "dwconv_bwd_data: null tensor");
    EFFDET_REQUIRE(aligned16(dz) && aligned16(w_kkc) && aligned16(dx), "dwconv_bwd_data: alignment");
    int s = dw_check("dwconv_bwd_data", B, H, W, C, k, stride, pad_t, pad_l, Ho, Wo);
    if (s) return s;
    EFFDET_DEVICE(device);
    cudaStream_t st = (cudaStream_t)stream;
    if (stride == 1 && Ho == H && Wo == W) {
        // stride 1: the data gradient is the same strip convolution with rotated taps and mirrored pads
        const long long strips = (long long)B * H * cdiv(W, kTW) * (C / 4);
        const int pt = k - 1 - pad_t, pl = k - 1 - pad_l;
        if (k == 3) dw_fwd_kernel<3, 1, true><<<cdiv(strips, 256), 256, 0, st>>>(dz, w_kkc, nullptr, nullptr, nullptr, dx, B, H, W, C, pt, pl, H, W);
        else dw_fwd_kernel<5, 1, true><<<cdiv(strips, 256), 256, 0, st>>>(dz, w_kkc, nullptr, nullptr, nullptr, dx, B, H, W, C, pt, pl, H, W);
        return launch_status("dw_fwd_kernel<bwd>");
    }
    const long long total = (long long)B * H * W * (C / 4);
    DW_DISPATCH(dw_bwd_data_kernel, <<<cdiv(total, 256), 256, 0, st>>>(dz, w_kkc, dx, B, H, W, C, pad_t, pad_l, Ho, Wo))
    return launch_status("dw_bwd_data_kernel");
}

extern "C" int effdet_dwconv_bwd_weight(const float* x, const float* dz, float* dw_c1kk, int B, int H, int W, int C,
                                        int k, int stride, int pad_t, int pad_l, int Ho, int Wo, int device,
                                        effdet_stream_t stream) {
    EFFDET_REQUIRE(x && dz && dw_c1kk, "dwconv_bwd_weight: null tensor");
    EFFDET_REQUIRE(aligned16(x) && aligned16(dz), "dwconv_bwd_weight: alignment");
    int s = dw_check("dwconv_bwd_weight", B, H, W, C, k, stride, pad_t, pad_l, Ho, Wo);
    if (s) return s;
    EFFDET_DEVICE(device);
    cudaStream_t st = (cudaStream_t)stream;
#define EFFDET_DW_TILED(K_, S_, TH_, TW_)                                                                                  \
        do {                                                                                                              \
            constexpr int IH = (TH_ - 1) * S_ + K_, IW = (TW_ - 1) * S_ + K_;                                             \
            const size_t smem = (size_t)(IH * IW + TH_ * TW_) * 8 * sizeof(float4);                                       \
            const int tiles_x = cdiv(Wo, TW_), tiles_y = cdiv(Ho, TH_);                                                   \
            cudaError_t e = cudaFuncSetAttribute(dw_bwd_weight_tiled_kernel<K_, S_, TH_, TW_>,                            \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
            if (e != cudaSuccess) return fail(EFFDET_ERR_LAUNCH, "dwconv_bwd_weight: smem opt-in: %s", cudaGetErrorString(e)); \
            const long long ntiles = (long long)B * tiles_x * tiles_y;                                                    \
            const int chunks = cdiv(C / 4, 8);                                                                            \
            long long gx = (148 * 2 + chunks - 1) / chunks;                                                               \
            if (gx > ntiles) gx = ntiles;                                                                                 \
            if (gx < 1) gx = 1;                                                                                           \
            dw_bwd_weight_tiled_kernel<K_, S_, TH_, TW_><<<dim3((unsigned)gx, chunks), 256, smem, st>>>(                  \
                x, dz, dw_c1kk, B, H, W, C, pad_t, pad_l, Ho, Wo, tiles_x, tiles_y);                                      \
        } while (0)
    if (Ho >= 16 && Wo >= 16) {
        // tiled kernel: TH x TW = 16x16 outputs (stride 1) or 8x16 (stride 2), 32 channels per CTA
        if (k == 3 && stride == 1) EFFDET_DW_TILED(3, 1, 16, 16);
        else if (k == 3 && stride == 2) EFFDET_DW_TILED(3, 2, 8, 16);
        else if (k == 5 && stride == 1) EFFDET_DW_TILED(5, 1, 16, 16);
        else EFFDET_DW_TILED(5, 2, 8, 16);
        return launch_status("dw_bwd_weight_tiled_kernel");
    }
    if (Ho >= 4 && Wo >= 4) {
        // late stages (8x8 / 4x4 maps, many channels): one 8x8 tile per image and 32-channel chunk
        if (k == 3 && stride == 1) EFFDET_DW_TILED(3, 1, 8, 8);
        else if (k == 3 && stride == 2) EFFDET_DW_TILED(3, 2, 8, 8);
        else if (k == 5 && stride == 1) EFFDET_DW_TILED(5, 1, 8, 8);
        else EFFDET_DW_TILED(5, 2, 8, 8);
        return launch_status("dw_bwd_weight_tiled_kernel");
    }
#undef EFFDET_DW_TILED
    const int cvecs = C / 4;
    const long long npix = (long long)B * Ho * cdiv(Wo, kTW);   // strips of kTW outputs
    const int rows = rowpack_rows(cvecs);
    long long rpb = (npix + 148 * 2 - 1) / (148 * 2);
    if (rpb < (long long)rows * 2) rpb = (long long)rows * 2;
    dim3 grid(cdiv(npix, rpb), rowpack_chunks(cvecs));
    DW_DISPATCH(dw_bwd_weight_kernel, <<<grid, 256, 0, st>>>(x, dz, dw_c1kk, B, H, W, C, pad_t, pad_l, Ho, Wo, (int)rpb))
    return launch_status("dw_bwd_weight_kernel");
}

extern "C" int effdet_pack_dw_weight(const float* w_c1kk, float* w_kkc, int C, int k, int device,
                                     effdet_stream_t stream) {
    EFFDET_REQUIRE(w_c1kk && w_kkc && C > 0 && (k == 3 || k == 5), "pack_dw_weight: bad arguments");
    EFFDET_DEVICE(device);
    pack_dw_weight_kernel<<<cdiv((long long)C * k * k, 256), 256, 0, (cudaStream_t)stream>>>(w_c1kk, w_kkc, C, k * k);
    return launch_status("pack_dw_weight_kernel");
}
