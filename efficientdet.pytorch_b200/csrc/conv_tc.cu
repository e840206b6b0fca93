// tcgen05 tensor-core implicit-GEMM convolution for the dense 3x3 (and 1x1) layers of head and
// neck: forward, data gradient (same kernel on the rotated/transposed weight pack) and weight
// gradient.  95 % of the model's FLOPs live here.
//
// Precision: operands stay fp32 in HBM (module API parity) and are split on the fly into
// bf16 hi + bf16 lo (x = hi + lo to 16 mantissa bits); every product is evaluated as
//   hi*hi + lo*hi + hi*lo      (three kind::f16 MMAs, fp32 accumulation in TMEM)
// which is ~2^-16 per product -- far inside north_star's 1e-3 where single-pass TF32 is not
//  -- at 2/3 the cost of 3xTF32.
//
// Structure per CTA (192 threads, one 128 x BN output tile, accumulator in TMEM):
//   warps 0-3  producers: gather the im2col A tile (128 pixels x 64 channels of one tap) with
//              128-bit loads, split to bf16 hi/lo, store into the canonical K-major SWIZZLE_128B
//              layout; afterwards the same warps run the epilogue (tcgen05.ld, bias / ReLU /
//              sigmoid / ReLU-mask / residual, 128-bit stores straight into the caller's layout)
//   warp 4     TMA: weight tiles (pre-split bf16 planes) -> smem, mbarrier complete_tx
//   warp 5     one elected thread issues tcgen05.mma; tcgen05.commit frees the smem stage
// Weight gradient: both operands are produced by the gather warps (pixels are the GEMM-K
// dimension, so the natural NHWC rows are MN-major operands), split-K over pixel ranges with
// fp32 atomics into the OIHW gradient.
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace effdet {

constexpr int kTcThreads = 192;        // weight-gradient kernels: 4 gather/epilogue warps + TMA + MMA
constexpr int kTcProducers = 128;
constexpr int kFwdThreads = 320;       // forward/dgrad kernel: 8 gather/epilogue warps + TMA + MMA
constexpr int kFwdProducers = 256;
constexpr int kTileM = 128;     // pixels per CTA (fwd/dgrad) or output channels per CTA (wgrad)
constexpr int kTileK = 64;      // bf16 elements per 128-byte swizzled row

// ---------------------------------------------------------------------------------------------
// forward / data-gradient kernel
// ---------------------------------------------------------------------------------------------
template <int BN, int STAGES>
struct FwdSmem {
    static constexpr int kA = kTileM * 128;   // bytes of one bf16 plane of the A tile
    static constexpr int kB = BN * 128;       // bytes of one bf16 plane of the B tile
    static constexpr int kStage = 2 * kA + 2 * kB;
    static constexpr int kChan = 3 * BN * 4;  // bias | scale | shift of this CTA's output channels
    static constexpr int kBytes = STAGES * kStage + 1024 /*alignment slack*/ + 256 /*barriers*/ + kChan;
};

// MB = true adds the MBConv-only pieces (BN+swish / SE gate on the input, raw-output save, BN affine, drop-connect scale)
template <int BN, int STAGES, bool MB>
__device__ __forceinline__ void conv_tc_body(const CUtensorMap& wmap, const effdet_conv_args& p, const int M, const int HW,
                                             const int kblocks, const int m0, const int n0) {
    using S = FwdSmem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned AND still a shared-space pointer (LDS/STS, not generic LD/ST)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::kStage);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
    float* chan = reinterpret_cast<float*>(smem + STAGES * S::kStage + 256);   // [3][BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int taps = p.ksize * p.ksize, pad = p.ksize / 2;
    const int KT = taps * kblocks;
    for (int i = threadIdx.x; i < BN; i += kFwdThreads) {
        const int n = n0 + i;
        const bool ok = n < p.Cout;
        chan[i] = (ok && p.bias) ? __ldg(p.bias + n) : 0.f;
        chan[BN + i] = (MB && ok && p.scale) ? __ldg(p.scale + n) : 1.f;
        chan[2 * BN + i] = (MB && ok && p.shift) ? __ldg(p.shift + n) : 0.f;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], kFwdProducers + 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc<BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ---------------- producers: im2col gather + bf16 split -----------------------------------
        const int t = threadIdx.x;
        const int j = t & 7;                 // 16-byte chunk (8 channels) within the 64-channel row
        long long base[4];
        int oyx[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = i * 32 + (t >> 3);
            const int m = m0 + r;
            if (m < M) {
                const int b = m / HW;
                const int pix = m - b * HW;
                const int oy = pix / p.W;
                oyx[i] = (oy << 16) | (pix - oy * p.W);
                base[i] = (long long)b * p.x_bstride;
            } else {
                oyx[i] = -1;
                base[i] = 0;
            }
        }
        // gather one stage worth of fp32 operands into registers (4 rows x 32 bytes per thread)
        auto load_stage = [&](int kt, float4 (&v)[8]) {
            const int tap = kt / kblocks;
            const int c = (kt - tap * kblocks) * kTileK + j * 8;
            const int ky = tap / p.ksize - pad, kx = tap % p.ksize - pad;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                v[2 * i] = f4zero();
                v[2 * i + 1] = f4zero();
                if (oyx[i] >= 0 && c < p.Cin) {
                    const int iy = (oyx[i] >> 16) + ky, ix = (oyx[i] & 0xffff) + kx;
                    if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
                        const float* src = p.x + base[i] + ((long long)iy * p.W + ix) * p.Cin + c;
                        v[2 * i] = ldg4(src);
                        if (c + 4 < p.Cin) v[2 * i + 1] = ldg4(src + 4);
                    }
                }
            }
        };
        // input prologue of the MBConv project conv (applied when the stage is converted, so the loads of the next
        // stage stay in flight): eval-BN + swish of the raw depthwise output, then the squeeze-excite gate.
        // Zero-filled elements (rows beyond M, channels beyond Cin) must stay zero, hence the validity tests.
        auto prologue = [&](int kt, float4 (&v)[8]) {
            const int tap = kt / kblocks;
            const int c = (kt - tap * kblocks) * kTileK + j * 8;
            if (c >= p.Cin) return;
            const bool hi_ok = c + 4 < p.Cin;
            float4 s0 = make_float4(1.f, 1.f, 1.f, 1.f), s1 = s0, h0 = f4zero(), h1 = f4zero();
            if (p.in_scale) {
                s0 = ldg4(p.in_scale + c); h0 = ldg4(p.in_shift + c);
                if (hi_ok) { s1 = ldg4(p.in_scale + c + 4); h1 = ldg4(p.in_shift + c + 4); }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (oyx[i] < 0) continue;
                if (p.in_scale) {
                    const float4 u0 = f4fma(v[2 * i], s0, h0), u1 = f4fma(v[2 * i + 1], s1, h1);
                    v[2 * i] = make_float4(swishf_(u0.x), swishf_(u0.y), swishf_(u0.z), swishf_(u0.w));
                    v[2 * i + 1] = hi_ok ? make_float4(swishf_(u1.x), swishf_(u1.y), swishf_(u1.z), swishf_(u1.w)) : f4zero();
                }
                if (p.a_scale) {                               // squeeze-excite gate on the input (per image, channel)
                    const float* gp = p.a_scale + (base[i] / p.x_bstride) * p.Cin + c;
                    v[2 * i] = f4mul(v[2 * i], ldg4(gp));
                    if (hi_ok) v[2 * i + 1] = f4mul(v[2 * i + 1], ldg4(gp + 4));
                }
            }
        };
        // split to bf16 hi/lo and publish the stage to the MMA warp
        auto store_stage = [&](int kt, float4 (&v)[8]) {
            const int s = kt % STAGES;
            const uint32_t ph = (kt / STAGES) & 1;
            if (MB && (p.in_scale || p.a_scale)) prologue(kt, v);
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* a_hi = smem + s * S::kStage;
            uint8_t* a_lo = a_hi + S::kA;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = i * 32 + (t >> 3);
                uint4 hi, lo;
                split8(v[2 * i], v[2 * i + 1], hi, lo);
                const int off = r * 128 + ((j ^ (r & 7)) << 4);
                *reinterpret_cast<uint4*>(a_hi + off) = hi;
                *reinterpret_cast<uint4*>(a_lo + off) = lo;
            }
            fence_proxy_async();
            mbar_arrive(&full_bar[s]);
        };
        // software pipeline: the loads of stage kt+1 are in flight while stage kt is converted and stored
        float4 va[8], vb[8];
        load_stage(0, va);
        for (int kt = 0; kt < KT; kt += 2) {
            if (kt + 1 < KT) load_stage(kt + 1, vb);
            store_stage(kt, va);
            if (kt + 1 < KT) {
                if (kt + 2 < KT) load_stage(kt + 2, va);
                store_stage(kt + 1, vb);
            }
        }
        // ---------------- epilogue ------------------------------------------------------------------
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int quarter = warp & 3, half = warp >> 2;     // TMEM lane quarter, column half of this warp
        const int m = m0 + quarter * 32 + lane;
        const bool row_ok = m < M;
        int b = 0;
        long long pix = 0;
        if (row_ok) {
            b = m / HW;
            pix = m - b * HW;
        }
        const float rs = (MB && row_ok && p.row_scale) ? __ldg(p.row_scale + b) : 1.f;
        const int ncols = min(BN, p.Cout - n0);
        const int nchunks = (ncols + 31) >> 5;
        const int c_begin = half ? (nchunks + 1) >> 1 : 0, c_end = half ? nchunks : (nchunks + 1) >> 1;
        const long long ybase = (long long)b * p.y_bstride + pix * p.Cout;
        const long long rbase = (long long)b * p.r_bstride + pix * p.Cout;
        const long long mbase = (long long)b * p.m_bstride + pix * p.Cout;
#pragma unroll 1
        for (int cc = c_begin; cc < c_end; ++cc) {
            uint32_t acc[32];
            tmem_ld32_issue(tmem_base + ((uint32_t)(quarter * 32) << 16) + cc * 32, acc);
            // streaming operands of the epilogue go in flight while the TMEM load completes
            float4 rv[8], mv[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int n = n0 + cc * 32 + q * 4;
                rv[q] = f4zero();
                mv[q] = make_float4(1.f, 1.f, 1.f, 1.f);
                if (row_ok && n < p.Cout) {
                    if (p.residual) rv[q] = ldg4(p.residual + rbase + n);
                    if (p.mask_src) mv[q] = ldg4(p.mask_src + mbase + n);
                }
            }
            tmem_ld32_wait(acc);
            if (!row_ok) continue;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int nl = cc * 32 + q * 4;
                const int n = n0 + nl;
                if (n >= p.Cout) break;
                float4 v = make_float4(__uint_as_float(acc[q * 4]), __uint_as_float(acc[q * 4 + 1]),
                                       __uint_as_float(acc[q * 4 + 2]), __uint_as_float(acc[q * 4 + 3]));
                v = f4add(v, *reinterpret_cast<const float4*>(chan + nl));
                if (MB && p.z) st4(p.z + ybase + n, v);
                if (MB) v = f4fma(v, *reinterpret_cast<const float4*>(chan + BN + nl), *reinterpret_cast<const float4*>(chan + 2 * BN + nl));
                if (p.act == EFFDET_ACT_RELU) {
                    v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
                } else if (p.act == EFFDET_ACT_SIGMOID) {
                    v = make_float4(sigmoidf_(v.x), sigmoidf_(v.y), sigmoidf_(v.z), sigmoidf_(v.w));
                } else if (p.act == EFFDET_ACT_SWISH) {
                    v = make_float4(swishf_(v.x), swishf_(v.y), swishf_(v.z), swishf_(v.w));
                }
                if (MB && p.row_scale) v = f4scale(v, rs);
                v = f4add(v, rv[q]);
                if (p.mask_src)
                    v = make_float4(mv[q].x > 0.f ? v.x : 0.f, mv[q].y > 0.f ? v.y : 0.f, mv[q].z > 0.f ? v.z : 0.f,
                                    mv[q].w > 0.f ? v.w : 0.f);
                st4(p.y + ybase + n, v);
            }
        }
        tc_fence_before();
    } else if (warp == 8) {
        // ---------------- TMA: weight tiles (hi plane, lo plane) ---------------------------------------
        if (lane == 0) {
            for (int kt = 0; kt < KT; ++kt) {
                const int s = kt % STAGES;
                const uint32_t ph = (kt / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* b_hi = smem + s * S::kStage + 2 * S::kA;
                mbar_arrive_expect_tx(&full_bar[s], 2 * S::kB);
                tma_load_3d(b_hi, &wmap, &full_bar[s], kt * kTileK, n0, 0);
                tma_load_3d(b_hi + S::kB, &wmap, &full_bar[s], kt * kTileK, n0, 1);
            }
        }
    } else {
        // ---------------- MMA issue ----------------------------------------------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(kTileM, BN, 0, 0);
            for (int kt = 0; kt < KT; ++kt) {
                const int s = kt % STAGES;
                const uint32_t ph = (kt / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + s * S::kStage);
                const uint32_t a_lo = a_hi + S::kA;
                const uint32_t b_hi = a_hi + 2 * S::kA;
                const uint32_t b_lo = b_hi + S::kB;
#pragma unroll
                for (int k = 0; k < kTileK / 16; ++k) {
                    const uint64_t dah = umma_desc(a_hi + k * 32, 16, 1024), dal = umma_desc(a_lo + k * 32, 16, 1024);
                    const uint64_t dbh = umma_desc(b_hi + k * 32, 16, 1024), dbl = umma_desc(b_lo + k * 32, 16, 1024);
                    umma_bf16(tmem_base, dal, dbh, idesc, (kt | k) != 0);
                    umma_bf16(tmem_base, dah, dbl, idesc, 1);
                    umma_bf16(tmem_base, dah, dbh, idesc, 1);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(accum_bar);
        }
    }
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc<BN>(tmem_base);
    }
}

template <int BN, int STAGES, bool MB>
__global__ void __launch_bounds__(kFwdThreads, (STAGES == 1 ? 2 : 1))
conv_tc_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ effdet_conv_args p, const int M, const int HW,
               const int kblocks) {
    conv_tc_body<BN, STAGES, MB>(wmap, p, M, HW, kblocks, blockIdx.x * kTileM, blockIdx.y * BN);
}

// Several pyramid levels that share one weight tensor (RetinaHead runs the same convs on P3..P7,
// models/retinahead.py:131-132) in ONE launch: the M tiles of all levels are concatenated so the small
// levels (a handful of CTAs each) ride along with the large ones instead of paying their own latency-bound launch.
constexpr int kMaxLevels = 8;
struct ConvMultiArgs {
    effdet_conv_args lv[kMaxLevels];
    int tile_begin[kMaxLevels + 1];
    int nlevels;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kFwdThreads, 1)
conv_tc_multi_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ ConvMultiArgs ma, const int kblocks) {
    const int tile = blockIdx.x;
    int l = 0;
    while (l + 1 < ma.nlevels && tile >= ma.tile_begin[l + 1]) ++l;
    const effdet_conv_args& p = ma.lv[l];
    conv_tc_body<BN, STAGES, false>(wmap, p, p.B * p.H * p.W, p.H * p.W, kblocks, (tile - ma.tile_begin[l]) * kTileM,
                                    blockIdx.y * BN);
}

// ---------------------------------------------------------------------------------------------
// Persistent variant of the multi-level kernel (the RetinaHead towers: 95 % of the model's FLOPs).
// The single-tile kernel above spends ~25 % of a tile's time outside the MMA main loop (TMEM allocation and barrier set-up,
// pipeline ramp, and above all the epilogue: 128 x 256 fp32 through thread-per-row stores while the tensor pipe idles).
// Here a CTA owns its SM for the whole launch and walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...; the roles are
// decoupled and the accumulator is double-buffered in TMEM (2 x BN columns = all 512 for BN = 256), so the epilogue of
// tile i drains accumulator i&1 while the MMA warp is already accumulating tile i+1 into the other one:
//   warps 0-7   im2col gather producers (as above), running ahead into the next tile as stages free up
//   warp 8      weight TMA            warp 9   MMA issuer
//   warps 10-13 epilogue: tcgen05.ld -> bias / ReLU / sigmoid / residual / ReLU-mask -> stores
// ---------------------------------------------------------------------------------------------
constexpr int kPersistThreads = 448;

template <int BN, int STAGES>
__global__ void __launch_bounds__(kPersistThreads, 1)
conv_tc_persist_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ ConvMultiArgs ma, const int kblocks,
                       const int ntn, const int total_tiles) {
    using S = FwdSmem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned AND still a shared-space pointer (LDS/STS, not generic LD/ST)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::kStage);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* chan = reinterpret_cast<float*>(smem + STAGES * S::kStage + 256);   // [BN] bias of the current n-tile

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const effdet_conv_args& p0 = ma.lv[0];
    const int taps = p0.ksize * p0.ksize, pad = p0.ksize / 2;
    const int KT = taps * kblocks;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], kFwdProducers + 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 4);
        }
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc<2 * BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile -> (level, first pixel row, first output channel)
    auto decode = [&](int tile, int& l, int& m0, int& n0) {
        const int mt = tile / ntn;
        n0 = (tile - mt * ntn) * BN;
        l = 0;
        while (l + 1 < ma.nlevels && mt >= ma.tile_begin[l + 1]) ++l;
        m0 = (mt - ma.tile_begin[l]) * kTileM;
    };

    if (warp < 8) {
        // ---------------- producers ------------------------------------------------------------------------------------
        const int t = threadIdx.x;
        const int j = t & 7;
        uint32_t it = 0;                                   // stage counter, continues across tiles
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int l, m0, n0;
            decode(tile, l, m0, n0);
            const effdet_conv_args& p = ma.lv[l];
            const int M = p.B * p.H * p.W, HW = p.H * p.W;
            long long base[4];
            int oyx[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = i * 32 + (t >> 3);
                const int m = m0 + r;
                if (m < M) {
                    const int b = m / HW;
                    const int pix = m - b * HW;
                    const int oy = pix / p.W;
                    oyx[i] = (oy << 16) | (pix - oy * p.W);
                    base[i] = (long long)b * p.x_bstride;
                } else {
                    oyx[i] = -1;
                    base[i] = 0;
                }
            }
            auto load_stage = [&](int kt, float4 (&v)[8]) {
                const int tap = kt / kblocks;
                const int c = (kt - tap * kblocks) * kTileK + j * 8;
                const int ky = tap / p.ksize - pad, kx = tap % p.ksize - pad;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    v[2 * i] = f4zero();
                    v[2 * i + 1] = f4zero();
                    if (oyx[i] >= 0 && c < p.Cin) {
                        const int iy = (oyx[i] >> 16) + ky, ix = (oyx[i] & 0xffff) + kx;
                        if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
                            const float* src = p.x + base[i] + ((long long)iy * p.W + ix) * p.Cin + c;
                            v[2 * i] = ldg4(src);
                            if (c + 4 < p.Cin) v[2 * i + 1] = ldg4(src + 4);
                        }
                    }
                }
            };
            auto store_stage = [&](int kt, const float4 (&v)[8]) {
                const uint32_t g = it + kt;
                const int s = g % STAGES;
                const uint32_t ph = (g / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* a_hi = smem + s * S::kStage;
                uint8_t* a_lo = a_hi + S::kA;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = i * 32 + (t >> 3);
                    uint4 hi, lo;
                    split8(v[2 * i], v[2 * i + 1], hi, lo);
                    const int off = r * 128 + ((j ^ (r & 7)) << 4);
                    *reinterpret_cast<uint4*>(a_hi + off) = hi;
                    *reinterpret_cast<uint4*>(a_lo + off) = lo;
                }
                fence_proxy_async();
                mbar_arrive(&full_bar[s]);
            };
            float4 va[8], vb[8];
            load_stage(0, va);
            for (int kt = 0; kt < KT; kt += 2) {
                if (kt + 1 < KT) load_stage(kt + 1, vb);
                store_stage(kt, va);
                if (kt + 1 < KT) {
                    if (kt + 2 < KT) load_stage(kt + 2, va);
                    store_stage(kt + 1, vb);
                }
            }
            it += KT;
        }
    } else if (warp == 8) {
        // ---------------- TMA: weight tiles of the tile's n-range (hi plane, lo plane) --------------------------------------
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int l, m0, n0;
                decode(tile, l, m0, n0);
                for (int kt = 0; kt < KT; ++kt, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* b_hi = smem + s * S::kStage + 2 * S::kA;
                    mbar_arrive_expect_tx(&full_bar[s], 2 * S::kB);
                    tma_load_3d(b_hi, &wmap, &full_bar[s], kt * kTileK, n0, 0);
                    tma_load_3d(b_hi + S::kB, &wmap, &full_bar[s], kt * kTileK, n0, 1);
                }
            }
        }
    } else if (warp == 9) {
        // ---------------- MMA issue -----------------------------------------------------------------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(kTileM, BN, 0, 0);
            uint32_t it = 0, iu = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iu) {
                const uint32_t acc = iu & 1, pacc = (iu >> 1) & 1;
                mbar_wait(&acc_empty[acc], pacc ^ 1);              // the epilogue warps have drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem_base + acc * BN;
                for (int kt = 0; kt < KT; ++kt, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + s * S::kStage);
                    const uint32_t a_lo = a_hi + S::kA;
                    const uint32_t b_hi = a_hi + 2 * S::kA;
                    const uint32_t b_lo = b_hi + S::kB;
#pragma unroll
                    for (int k = 0; k < kTileK / 16; ++k) {
                        const uint64_t dah = umma_desc(a_hi + k * 32, 16, 1024), dal = umma_desc(a_lo + k * 32, 16, 1024);
                        const uint64_t dbh = umma_desc(b_hi + k * 32, 16, 1024), dbl = umma_desc(b_lo + k * 32, 16, 1024);
                        umma_bf16(d, dal, dbh, idesc, (kt | k) != 0);
                        umma_bf16(d, dah, dbl, idesc, 1);
                        umma_bf16(d, dah, dbh, idesc, 1);
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&acc_full[acc]);
            }
        }
    } else {
        // ---------------- epilogue warps ------------------------------------------------------------------------------------
        const int etid = threadIdx.x - 320;
        const int quarter = warp & 3;
        uint32_t iu = 0;
        int chan_n0 = -1;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iu) {
            int l, m0, n0;
            decode(tile, l, m0, n0);
            const effdet_conv_args& p = ma.lv[l];
            const int M = p.B * p.H * p.W, HW = p.H * p.W;
            if (n0 != chan_n0) {                                   // bias slice of this n-tile (shared by all levels)
                named_bar_sync(1, 128);
                for (int i = etid; i < BN; i += 128) chan[i] = (n0 + i < p.Cout && p.bias) ? __ldg(p.bias + n0 + i) : 0.f;
                named_bar_sync(1, 128);
                chan_n0 = n0;
            }
            const uint32_t acc = iu & 1, pacc = (iu >> 1) & 1;
            mbar_wait(&acc_full[acc], pacc);
            tc_fence_after();
            const int m = m0 + quarter * 32 + lane;
            const bool row_ok = m < M;
            int b = 0;
            long long pix = 0;
            if (row_ok) {
                b = m / HW;
                pix = m - (long long)b * HW;
            }
            const int ncols = min(BN, p.Cout - n0);
            const int nchunks = (ncols + 31) >> 5;
            const long long ybase = (long long)b * p.y_bstride + pix * p.Cout;
            const long long rbase = (long long)b * p.r_bstride + pix * p.Cout;
            const long long mbase = (long long)b * p.m_bstride + pix * p.Cout;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int cc = 0; cc < nchunks; ++cc) {
                uint32_t a32[32];
                tmem_ld32_issue(taddr + cc * 32, a32);
                float4 rv[8], mv[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int n = n0 + cc * 32 + q * 4;
                    rv[q] = f4zero();
                    mv[q] = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (row_ok && n < p.Cout) {
                        if (p.residual) rv[q] = ldg4(p.residual + rbase + n);
                        if (p.mask_src) mv[q] = ldg4(p.mask_src + mbase + n);
                    }
                }
                tmem_ld32_wait(a32);
                if (cc == nchunks - 1) {                           // accumulator drained: the MMA warp may reuse it
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[acc]);
                }
                if (!row_ok) continue;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int nl = cc * 32 + q * 4;
                    const int n = n0 + nl;
                    if (n >= p.Cout) break;
                    float4 v = make_float4(__uint_as_float(a32[q * 4]), __uint_as_float(a32[q * 4 + 1]),
                                           __uint_as_float(a32[q * 4 + 2]), __uint_as_float(a32[q * 4 + 3]));
                    v = f4add(v, *reinterpret_cast<const float4*>(chan + nl));
                    if (p.act == EFFDET_ACT_RELU) {
                        v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
                    } else if (p.act == EFFDET_ACT_SIGMOID) {
                        v = make_float4(sigmoidf_(v.x), sigmoidf_(v.y), sigmoidf_(v.z), sigmoidf_(v.w));
                    } else if (p.act == EFFDET_ACT_SWISH) {
                        v = make_float4(swishf_(v.x), swishf_(v.y), swishf_(v.z), swishf_(v.w));
                    }
                    v = f4add(v, rv[q]);
                    if (p.mask_src)
                        v = make_float4(mv[q].x > 0.f ? v.x : 0.f, mv[q].y > 0.f ? v.y : 0.f, mv[q].z > 0.f ? v.z : 0.f,
                                        mv[q].w > 0.f ? v.w : 0.f);
                    st4(p.y + ybase + n, v);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc<2 * BN>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------
// weight-gradient kernel:  D[n, c | tap] += sum_pixels dy[pixel, n] * x[pixel + tap, c]
//   GEMM-M = 128 output channels, GEMM-N = BC input channels, GEMM-K = pixels (64 per stage);
//   both operands are NHWC rows (channels contiguous) = MN-major SWIZZLE_128B operands:
//   one 64-channel group of a stage = 64 pixel-rows x 128 B, groups LBO = 8192 B apart,
//   8-pixel groups SBO = 1024 B apart.
// ---------------------------------------------------------------------------------------------
template <int BC, int STAGES>
struct WgSmem {
    static constexpr int kA = kTileK * 128 * (kTileM / 64);   // one plane of the dy tile: 2 channel groups
    static constexpr int kB = kTileK * 128 * (BC / 64);       // one plane of the x tile
    static constexpr int kStage = 2 * kA + 2 * kB;
    static constexpr int kBytes = STAGES * kStage + 1024 + 256;
};

// gather `ngroups` channel groups (64 channels each) of 64 pixel rows into swizzled bf16 planes
template <int NGROUPS>
__device__ __forceinline__ void wg_produce(const float* __restrict__ src, const long long bstride, const int C, const int c0,
                                           const int H, const int W, const int HW, const int M, const int mbase, const int dy,
                                           const int dx, uint8_t* hi_plane, uint8_t* lo_plane, const int t) {
    // item = (pixel row r, group g, chunk j): 64 * NGROUPS * 8 items over 128 threads
    constexpr int ITEMS = 64 * NGROUPS * 8 / kTcProducers;
    const int j = t & 7;
#pragma unroll 4
    for (int it = 0; it < ITEMS; ++it) {
        const int idx = it * kTcProducers + t;
        const int rg = idx >> 3;                   // r * NGROUPS + g  (g fastest so a warp reads contiguous channels)
        const int g = rg % NGROUPS, r = rg / NGROUPS;
        const int m = mbase + r;
        const int c = c0 + g * 64 + j * 8;
        float4 v0 = f4zero(), v1 = f4zero();
        if (m < M && c < C) {
            const int b = m / HW;
            const int pix = m - b * HW;
            const int oy = pix / W;
            const int iy = oy + dy, ix = pix - oy * W + dx;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
                const float* q = src + (long long)b * bstride + ((long long)iy * W + ix) * C + c;
                v0 = ldg4(q);
                if (c + 4 < C) v1 = ldg4(q + 4);
            }
        }
        uint4 hi, lo;
        split8(v0, v1, hi, lo);
        const int off = g * (kTileK * 128) + r * 128 + ((j ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(hi_plane + off) = hi;
        *reinterpret_cast<uint4*>(lo_plane + off) = lo;
    }
}

template <int BC, int STAGES>
__global__ void __launch_bounds__(kTcThreads, 1)
wgrad_tc_kernel(const effdet_wgrad_args p, const int M, const int HW, const int chunks_per_split, const int ctiles) {
    using S = WgSmem<BC, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned AND still a shared-space pointer (LDS/STS, not generic LD/ST)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::kStage);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ct = blockIdx.x % ctiles, nt = blockIdx.x / ctiles;
    const int c0 = ct * BC, n0 = nt * kTileM;
    const int tap = blockIdx.y;
    const int pad = p.ksize / 2;
    const int dy = tap / p.ksize - pad, dx = tap % p.ksize - pad;
    const int nchunks = (M + kTileK - 1) / kTileK;
    const int ch_begin = blockIdx.z * chunks_per_split;
    const int ch_end = min(nchunks, ch_begin + chunks_per_split);
    const int KT = ch_end - ch_begin;      // >= 1 by construction of the grid

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], kTcProducers);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc<BC>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        const int t = threadIdx.x;
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % STAGES;
            const uint32_t ph = (kt / STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* a_hi = smem + s * S::kStage;
            uint8_t* b_hi = a_hi + 2 * S::kA;
            const int mbase = (ch_begin + kt) * kTileK;
            wg_produce<kTileM / 64>(p.dy, p.dy_bstride, p.Cout, n0, p.H, p.W, HW, M, mbase, 0, 0, a_hi, a_hi + S::kA, t);
            wg_produce<BC / 64>(p.x, p.x_bstride, p.Cin, c0, p.H, p.W, HW, M, mbase, dy, dx, b_hi, b_hi + S::kB, t);
            fence_proxy_async();
            mbar_arrive(&full_bar[s]);
        }
        // epilogue: row = output channel n, columns = input channels c -> atomics into OIHW
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int n = n0 + warp * 32 + lane;
        const int kk = p.ksize * p.ksize;
#pragma unroll 1
        for (int cc = 0; cc < BC / 32; ++cc) {
            uint32_t acc[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + cc * 32, acc);
            if (n >= p.Cout) continue;
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int c = c0 + cc * 32 + q;
                if (c < p.Cin) atomicAdd(p.dw + ((long long)n * p.Cin + c) * kk + tap, __uint_as_float(acc[q]));
            }
        }
        tc_fence_before();
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(kTileM, BC, 1, 1);
            constexpr uint32_t LBO = kTileK * 128, SBO = 1024;
            for (int kt = 0; kt < KT; ++kt) {
                const int s = kt % STAGES;
                const uint32_t ph = (kt / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + s * S::kStage);
                const uint32_t a_lo = a_hi + S::kA;
                const uint32_t b_hi = a_hi + 2 * S::kA;
                const uint32_t b_lo = b_hi + S::kB;
#pragma unroll
                for (int k = 0; k < kTileK / 16; ++k) {
                    const uint32_t ko = k * 2 * SBO;     // 16 pixels = two 8-row groups
                    const uint64_t dah = umma_desc(a_hi + ko, LBO, SBO), dal = umma_desc(a_lo + ko, LBO, SBO);
                    const uint64_t dbh = umma_desc(b_hi + ko, LBO, SBO), dbl = umma_desc(b_lo + ko, LBO, SBO);
                    umma_bf16(tmem_base, dal, dbh, idesc, (kt | k) != 0);
                    umma_bf16(tmem_base, dah, dbl, idesc, 1);
                    umma_bf16(tmem_base, dah, dbh, idesc, 1);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(accum_bar);
        }
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc<BC>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------
// weight-gradient kernel, TMA-fed: the operands were pre-split into bf16 hi/lo planes
// [2][B][H][W][Cpad] by split_planes_kernel, so the gather warps disappear: one thread issues
// 5-D tensor-map loads (channel group, x, y, image, plane) whose out-of-bounds zero fill IS the
// convolution's zero padding (the tap shift is just a coordinate offset), one thread issues the
// MMAs, four warps drain TMEM into the OIHW gradient with atomics.  A stage covers a box of
// kstage = Wb*Hb*Bb pixels (<= 64, multiple of 16).
// ---------------------------------------------------------------------------------------------
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

int conv_tc_kpad(int k) { return (k + kTileK - 1) / kTileK * kTileK; }

template <int BC, int STAGES>
__global__ void __launch_bounds__(kTcThreads, 1)
wgrad_tc2_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, const effdet_wgrad_args p,
                 const WgGeom g, const int chunks_per_split, const int ctiles) {
    using S = WgSmem<BC, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned AND still a shared-space pointer (LDS/STS, not generic LD/ST)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::kStage);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ct = blockIdx.x % ctiles, nt = blockIdx.x / ctiles;
    const int c0 = ct * BC, n0 = nt * kTileM;
    const int tap = blockIdx.y;
    const int pad = p.ksize / 2;
    const int dy = tap / p.ksize - pad, dx = tap % p.ksize - pad;
    const int nchunks = g.nbx * g.nby * g.nbb;
    const int ch_begin = blockIdx.z * chunks_per_split;
    const int ch_end = min(nchunks, ch_begin + chunks_per_split);
    const int KT = ch_end - ch_begin;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc<BC>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr int GROUP = kTileK * 128;            // smem slot of one 64-channel group of one plane

    if (warp == 4) {
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)(2 * (kTileM / 64 + BC / 64) * g.kstage * 128);
            for (int kt = 0; kt < KT; ++kt) {
                const int s = kt % STAGES;
                const uint32_t ph = (kt / STAGES) & 1;
                int ch = ch_begin + kt;
                const int bx = ch % g.nbx;
                ch /= g.nbx;
                const int by = ch % g.nby;
                const int bb = ch / g.nby;
                const int x0 = bx * g.Wb, y0 = by * g.Hb, b0 = bb * g.Bb;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_arrive_expect_tx(&full_bar[s], bytes);
                uint8_t* a_hi = smem + s * S::kStage;
                uint8_t* b_hi = a_hi + 2 * S::kA;
#pragma unroll
                for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
                    for (int q = 0; q < kTileM / 64; ++q)
                        tma_load_5d(a_hi + pl * S::kA + q * GROUP, &map_dy, &full_bar[s], n0 + q * 64, x0, y0, b0, pl);
#pragma unroll
                    for (int q = 0; q < BC / 64; ++q)
                        tma_load_5d(b_hi + pl * S::kB + q * GROUP, &map_x, &full_bar[s], c0 + q * 64, x0 + dx, y0 + dy, b0, pl);
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(kTileM, BC, 1, 1);
            constexpr uint32_t LBO = GROUP, SBO = 1024;
            const int ksteps = g.kstage / 16;
            for (int kt = 0; kt < KT; ++kt) {
                const int s = kt % STAGES;
                const uint32_t ph = (kt / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + s * S::kStage);
                const uint32_t a_lo = a_hi + S::kA;
                const uint32_t b_hi = a_hi + 2 * S::kA;
                const uint32_t b_lo = b_hi + S::kB;
                for (int k = 0; k < ksteps; ++k) {
                    const uint32_t ko = k * 2 * SBO;
                    const uint64_t dah = umma_desc(a_hi + ko, LBO, SBO), dal = umma_desc(a_lo + ko, LBO, SBO);
                    const uint64_t dbh = umma_desc(b_hi + ko, LBO, SBO), dbl = umma_desc(b_lo + ko, LBO, SBO);
                    umma_bf16(tmem_base, dal, dbh, idesc, (kt | k) != 0);
                    umma_bf16(tmem_base, dah, dbl, idesc, 1);
                    umma_bf16(tmem_base, dah, dbh, idesc, 1);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(accum_bar);
        }
    } else {
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int n = n0 + warp * 32 + lane;
        const int kk = p.ksize * p.ksize;
#pragma unroll 1
        for (int cc = 0; cc < BC / 32; ++cc) {
            uint32_t acc[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + cc * 32, acc);
            if (n >= p.Cout) continue;
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int c = c0 + cc * 32 + q;
                if (c < p.Cin) atomicAdd(p.dw + ((long long)n * p.Cin + c) * kk + tap, __uint_as_float(acc[q]));
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc<BC>(tmem_base);
    }
}

// Multi-level variant: the pixel chunks of several pyramid levels (same weights, e.g. the five RetinaHead levels)
// form one long GEMM-K dimension, so one launch covers them all; each chunk looks up its level's tensor maps and
// pixel-box geometry.
constexpr int kWgMaxLevels = 8;
struct WgMaps {
    CUtensorMap dy[kWgMaxLevels];
    CUtensorMap x[kWgMaxLevels];
};
struct WgMultiArgs {
    WgGeom g[kWgMaxLevels];
    int chunk_begin[kWgMaxLevels + 1];
    int nlevels;
    float* dw;
    int Cin, Cout, ksize;
};

template <int BC, int STAGES>
__global__ void __launch_bounds__(kTcThreads, 1)
wgrad_tc2_multi_kernel(const __grid_constant__ WgMaps maps, const __grid_constant__ WgMultiArgs a, const int chunks_per_split,
                       const int ctiles) {
    using S = WgSmem<BC, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned AND still a shared-space pointer (LDS/STS, not generic LD/ST)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::kStage);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ct = blockIdx.x % ctiles, nt = blockIdx.x / ctiles;
    const int c0 = ct * BC, n0 = nt * kTileM;
    const int tap = blockIdx.y;
    const int pad = a.ksize / 2;
    const int dy = tap / a.ksize - pad, dx = tap % a.ksize - pad;
    const int nchunks = a.chunk_begin[a.nlevels];
    const int ch_begin = blockIdx.z * chunks_per_split;
    const int ch_end = min(nchunks, ch_begin + chunks_per_split);
    const int KT = ch_end - ch_begin;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc<BC>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr int GROUP = kTileK * 128;

    if (warp == 4) {
        if (lane == 0) {
            int l = 0;
            for (int kt = 0; kt < KT; ++kt) {
                const int s = kt % STAGES;
                const uint32_t ph = (kt / STAGES) & 1;
                const int chg = ch_begin + kt;
                while (l + 1 < a.nlevels && chg >= a.chunk_begin[l + 1]) ++l;
                const WgGeom& g = a.g[l];
                int ch = chg - a.chunk_begin[l];
                const int bx = ch % g.nbx;
                ch /= g.nbx;
                const int by = ch % g.nby;
                const int bb = ch / g.nby;
                const int x0 = bx * g.Wb, y0 = by * g.Hb, b0 = bb * g.Bb;
                const uint32_t bytes = (uint32_t)(2 * (kTileM / 64 + BC / 64) * g.kstage * 128);
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_arrive_expect_tx(&full_bar[s], bytes);
                uint8_t* a_hi = smem + s * S::kStage;
                uint8_t* b_hi = a_hi + 2 * S::kA;
#pragma unroll
                for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
                    for (int q = 0; q < kTileM / 64; ++q)
                        tma_load_5d(a_hi + pl * S::kA + q * GROUP, &maps.dy[l], &full_bar[s], n0 + q * 64, x0, y0, b0, pl);
#pragma unroll
                    for (int q = 0; q < BC / 64; ++q)
                        tma_load_5d(b_hi + pl * S::kB + q * GROUP, &maps.x[l], &full_bar[s], c0 + q * 64, x0 + dx, y0 + dy, b0, pl);
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(kTileM, BC, 1, 1);
            constexpr uint32_t LBO = GROUP, SBO = 1024;
            int l = 0;
            for (int kt = 0; kt < KT; ++kt) {
                const int s = kt % STAGES;
                const uint32_t ph = (kt / STAGES) & 1;
                const int chg = ch_begin + kt;
                while (l + 1 < a.nlevels && chg >= a.chunk_begin[l + 1]) ++l;
                const int ksteps = a.g[l].kstage / 16;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + s * S::kStage);
                const uint32_t a_lo = a_hi + S::kA;
                const uint32_t b_hi = a_hi + 2 * S::kA;
                const uint32_t b_lo = b_hi + S::kB;
                for (int k = 0; k < ksteps; ++k) {
                    const uint32_t ko = k * 2 * SBO;
                    const uint64_t dah = umma_desc(a_hi + ko, LBO, SBO), dal = umma_desc(a_lo + ko, LBO, SBO);
                    const uint64_t dbh = umma_desc(b_hi + ko, LBO, SBO), dbl = umma_desc(b_lo + ko, LBO, SBO);
                    umma_bf16(tmem_base, dal, dbh, idesc, (kt | k) != 0);
                    umma_bf16(tmem_base, dah, dbl, idesc, 1);
                    umma_bf16(tmem_base, dah, dbh, idesc, 1);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(accum_bar);
        }
    } else {
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int n = n0 + warp * 32 + lane;
        const int kk = a.ksize * a.ksize;
#pragma unroll 1
        for (int cc = 0; cc < BC / 32; ++cc) {
            uint32_t acc[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + cc * 32, acc);
            if (n >= a.Cout) continue;
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int c = c0 + cc * 32 + q;
                if (c < a.Cin) atomicAdd(a.dw + ((long long)n * a.Cin + c) * kk + tap, __uint_as_float(acc[q]));
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc<BC>(tmem_base);
    }
}

// fp32 [B][HW][C] (image stride bstride) -> bf16 planes [2][B*HW][Cpad], zero padded channels
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ x, long long bstride, const float* __restrict__ a_scale,
                                                           __nv_bfloat16* __restrict__ out, int B, int HW, int C, int Cpad,
                                                           const float* __restrict__ in_scale = nullptr,
                                                           const float* __restrict__ in_shift = nullptr) {
    const int cv = Cpad / 8;
    const long long total = (long long)B * HW * cv;
    const long long plane = (long long)B * HW * Cpad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i % cv);
        const long long row = i / cv;
        const int b = (int)(row / HW);
        const long long pix = row - (long long)b * HW;
        const int c = j * 8;
        float4 v0 = f4zero(), v1 = f4zero();
        if (c < C) {
            const float* q = x + (long long)b * bstride + pix * C + c;
            v0 = ldg4(q);
            if (c + 4 < C) v1 = ldg4(q + 4);
            if (in_scale) {                        // operand = swish(bn(x)) of a raw conv output (pre-activation only in HBM)
                const float4 u0 = f4fma(v0, ldg4(in_scale + c), ldg4(in_shift + c));
                v0 = make_float4(swishf_(u0.x), swishf_(u0.y), swishf_(u0.z), swishf_(u0.w));
                if (c + 4 < C) {
                    const float4 u1 = f4fma(v1, ldg4(in_scale + c + 4), ldg4(in_shift + c + 4));
                    v1 = make_float4(swishf_(u1.x), swishf_(u1.y), swishf_(u1.z), swishf_(u1.w));
                }
            }
            if (a_scale) {
                v0 = f4mul(v0, ldg4(a_scale + (long long)b * C + c));
                if (c + 4 < C) v1 = f4mul(v1, ldg4(a_scale + (long long)b * C + c + 4));
            }
        }
        uint4 hi, lo;
        split8(v0, v1, hi, lo);
        *reinterpret_cast<uint4*>(out + row * Cpad + c) = hi;
        *reinterpret_cast<uint4*>(out + plane + row * Cpad + c) = lo;
    }
}

// split + per-channel column sums in one pass: the weight gradient's dy operand is split into planes AND reduced
// into the bias gradient (dbias[n] += sum over pixels) while it streams by, so no second read of dy.
__global__ void __launch_bounds__(256) split_planes_colsum_kernel(const float* __restrict__ x, long long bstride,
                                                                  __nv_bfloat16* __restrict__ out, float* __restrict__ colsum,
                                                                  int B, int HW, int C, int Cpad, int rows_per_block) {
    __shared__ float red[256 * 8];
    const int cv8 = Cpad / 8;
    const int cvb = cv8 < 256 ? cv8 : 256;
    const int rows = 256 / cvb;
    const int tr = threadIdx.x / cvb, tc = threadIdx.x - tr * cvb;
    const int j = blockIdx.y * cvb + tc;
    const bool active = tr < rows && j < cv8;
    const long long nrows = (long long)B * HW;
    const long long plane = nrows * Cpad;
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.f;
    if (active) {
        const int c = j * 8;
        const long long r_begin = (long long)blockIdx.x * rows_per_block;
        const long long r_end = min(nrows, r_begin + rows_per_block);
        for (long long row = r_begin + tr; row < r_end; row += rows) {
            const int b = (int)(row / HW);
            const long long pix = row - (long long)b * HW;
            float4 v0 = f4zero(), v1 = f4zero();
            if (c < C) {
                const float* q = x + (long long)b * bstride + pix * C + c;
                v0 = ldg4(q);
                if (c + 4 < C) v1 = ldg4(q + 4);
            }
            uint4 hi, lo;
            split8(v0, v1, hi, lo);
            *reinterpret_cast<uint4*>(out + row * Cpad + c) = hi;
            *reinterpret_cast<uint4*>(out + plane + row * Cpad + c) = lo;
            s[0] += v0.x; s[1] += v0.y; s[2] += v0.z; s[3] += v0.w;
            s[4] += v1.x; s[5] += v1.y; s[6] += v1.z; s[7] += v1.w;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = s[i];
    __syncthreads();
    if (tr == 0 && j < cv8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float acc = 0.f;
            for (int r = 0; r < rows; ++r) acc += red[(r * cvb + tc) * 8 + i];
            const int c = j * 8 + i;
            if (c < C) atomicAdd(colsum + c, acc);
        }
    }
}

static int split_dy_launch(const effdet_wgrad_args* a, int cout_pad, cudaStream_t st) {
    const int HW = a->H * a->W;
    if (!a->dbias) {
        int blocks = cdiv((long long)a->B * HW * (cout_pad / 8), 256);
        if (blocks > 148 * 16) blocks = 148 * 16;
        split_planes_kernel<<<blocks, 256, 0, st>>>(a->dy, a->dy_bstride, nullptr, (__nv_bfloat16*)a->ws_dy, a->B, HW, a->Cout, cout_pad);
        return launch_status("split_planes_kernel");
    }
    const int cv8 = cout_pad / 8;
    const int cvb = cv8 < 256 ? cv8 : 256;
    const int rows = 256 / cvb;
    const long long nrows = (long long)a->B * HW;
    long long rpb = (nrows + 148 * 4 - 1) / (148 * 4);
    if (rpb < (long long)rows * 8) rpb = (long long)rows * 8;
    dim3 grid(cdiv(nrows, rpb), cdiv(cv8, cvb));
    split_planes_colsum_kernel<<<grid, 256, 0, st>>>(a->dy, a->dy_bstride, (__nv_bfloat16*)a->ws_dy, a->dbias, a->B, HW, a->Cout,
                                                    cout_pad, (int)rpb);
    return launch_status("split_planes_colsum_kernel");
}

// ---------------------------------------------------------------------------------------------
// weight pre-split: OIHW fp32 -> bf16 planes [2][rows][taps][Kpad] (K-major, zero padded)
//   forward pack : rows = Cout, k = Cin,  W[n][c][tap]
//   dgrad pack   : rows = Cin,  k = Cout, W[n][c][taps-1-tap]   (180-degree rotation, transpose)
// ---------------------------------------------------------------------------------------------
__global__ void pack_weight_tc_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int Cin, int taps,
                                      int kpad, int dgrad) {
    const int rows = dgrad ? Cin : Cout;
    const int kdim = dgrad ? Cout : Cin;
    const long long plane = (long long)rows * taps * kpad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % kpad);
        const long long r2 = i / kpad;
        const int tap = (int)(r2 % taps);
        const int row = (int)(r2 / taps);
        float v = 0.f;
        if (k < kdim) {
            const int n = dgrad ? k : row, c = dgrad ? row : k;
            const int st = dgrad ? (taps - 1 - tap) : tap;
            v = __ldg(w + ((long long)n * Cin + c) * taps + st);
        }
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        out[i] = h;
        out[plane + i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------


bool conv_tc_eligible(const effdet_conv_args* a) {
    if (a->w_tc == nullptr || a->Cin % 4 || a->Cout % 4 || a->Cout < 16) return false;
    // measured (profiles/r01_bench_full_breakdown.json): the narrowest 1x1 layers are faster on the CUDA cores
    if (a->ksize == 1 && (a->Cin < 24 || (a->Cin <= 32 && a->Cout <= 16))) return false;
    return true;
}

int conv_tc_launch(const effdet_conv_args* a, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(EFFDET_ERR_UNSUPPORTED, "conv2d(tc): cuTensorMapEncodeTiled unavailable");
    const long long Mll = (long long)a->B * a->H * a->W;
    const int M = (int)Mll, HW = a->H * a->W;
    const int taps = a->ksize * a->ksize;
    const int kpad = conv_tc_kpad(a->Cin);
    const int kblocks = kpad / kTileK;
    const int KT = taps * kblocks;
    const int BN = a->Cout <= 64 ? 64 : (a->Cout <= 128 ? 128 : 256);
    CUtensorMap map;
    const cuuint64_t gdim[3] = {(cuuint64_t)taps * kpad, (cuuint64_t)a->Cout, 2};
    const cuuint64_t gstr[2] = {(cuuint64_t)taps * kpad * 2, (cuuint64_t)a->Cout * taps * kpad * 2};
    const cuuint32_t box[3] = {(cuuint32_t)kTileK, (cuuint32_t)BN, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(a->w_tc), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(EFFDET_ERR_LAUNCH, "conv2d(tc): cuTensorMapEncodeTiled failed (%d)", (int)r);
    dim3 grid(cdiv(M, kTileM), cdiv(a->Cout, BN));
    const bool mb = a->a_scale || a->z || a->scale || a->row_scale || a->in_scale;
#define EFFDET_TC_LAUNCH1(BN_, ST_, MB_)                                                                                   \
    do {                                                                                                                  \
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN_, ST_, MB_>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                             FwdSmem<BN_, ST_>::kBytes);                                                  \
        if (e != cudaSuccess) return fail(EFFDET_ERR_LAUNCH, "conv2d(tc): smem opt-in: %s", cudaGetErrorString(e));       \
        conv_tc_kernel<BN_, ST_, MB_><<<grid, kFwdThreads, FwdSmem<BN_, ST_>::kBytes, st>>>(map, *a, M, HW, kblocks);      \
    } while (0)
#define EFFDET_TC_LAUNCH(BN_, ST_)                                                                                         \
    do {                                                                                                                  \
        if (mb) EFFDET_TC_LAUNCH1(BN_, ST_, true);                                                                        \
        else EFFDET_TC_LAUNCH1(BN_, ST_, false);                                                                          \
    } while (0)
    // short reductions (1x1 convs of the backbone): single-stage instances so 2-3 CTAs share an SM and hide each
    // other's prologue / epilogue; long reductions: deep pipelines, one CTA per SM
    if (KT <= 2) {
        if (BN == 64) EFFDET_TC_LAUNCH(64, 1);
        else if (BN == 128) EFFDET_TC_LAUNCH(128, 1);
        else EFFDET_TC_LAUNCH(256, 1);
    } else {
        if (BN == 64) EFFDET_TC_LAUNCH(64, 4);
        else if (BN == 128) EFFDET_TC_LAUNCH(128, 3);
        else EFFDET_TC_LAUNCH(256, 2);
    }
#undef EFFDET_TC_LAUNCH
#undef EFFDET_TC_LAUNCH1
    return launch_status("conv_tc_kernel");
}


int conv_tc_multi_launch(const effdet_conv_args* levels, int nlevels, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(EFFDET_ERR_UNSUPPORTED, "conv2d_multi(tc): cuTensorMapEncodeTiled unavailable");
    const effdet_conv_args* a = &levels[0];
    const int taps = a->ksize * a->ksize;
    const int kpad = conv_tc_kpad(a->Cin);
    const int kblocks = kpad / kTileK;
    const int BN = a->Cout <= 64 ? 64 : (a->Cout <= 128 ? 128 : 256);
    CUtensorMap map;
    const cuuint64_t gdim[3] = {(cuuint64_t)taps * kpad, (cuuint64_t)a->Cout, 2};
    const cuuint64_t gstr[2] = {(cuuint64_t)taps * kpad * 2, (cuuint64_t)a->Cout * taps * kpad * 2};
    const cuuint32_t box[3] = {(cuuint32_t)kTileK, (cuuint32_t)BN, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(a->w_tc), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(EFFDET_ERR_LAUNCH, "conv2d_multi(tc): cuTensorMapEncodeTiled failed (%d)", (int)r);
    ConvMultiArgs ma;
    memset(&ma, 0, sizeof(ma));
    ma.nlevels = nlevels;
    int tiles = 0;
    for (int l = 0; l < nlevels; ++l) {
        ma.lv[l] = levels[l];
        ma.tile_begin[l] = tiles;
        tiles += cdiv((long long)levels[l].B * levels[l].H * levels[l].W, kTileM);
    }
    for (int l = nlevels; l <= kMaxLevels; ++l) ma.tile_begin[l] = tiles;
    static const bool persist = [] {
        const char* v = getenv("EFFDET_B200_PERSIST");
        return !(v && v[0] == '0');
    }();
    if (persist) {
        const int ntn = cdiv(a->Cout, BN);
        const int total = tiles * ntn;
        const int pgrid = total < 148 ? total : 148;
#define EFFDET_TCP_LAUNCH(BN_, ST_)                                                                                        \
    do {                                                                                                                  \
        cudaError_t e = cudaFuncSetAttribute(conv_tc_persist_kernel<BN_, ST_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             FwdSmem<BN_, ST_>::kBytes);                                                  \
        if (e != cudaSuccess) return fail(EFFDET_ERR_LAUNCH, "conv2d_multi(tc): smem opt-in: %s", cudaGetErrorString(e)); \
        conv_tc_persist_kernel<BN_, ST_><<<pgrid, kPersistThreads, FwdSmem<BN_, ST_>::kBytes, st>>>(map, ma, kblocks, ntn, total); \
    } while (0)
        if (BN == 64) EFFDET_TCP_LAUNCH(64, 4);
        else if (BN == 128) EFFDET_TCP_LAUNCH(128, 3);
        else EFFDET_TCP_LAUNCH(256, 2);
#undef EFFDET_TCP_LAUNCH
        return launch_status("conv_tc_persist_kernel");
    }
    dim3 grid(tiles, cdiv(a->Cout, BN));
#define EFFDET_TCM_LAUNCH(BN_, ST_)                                                                                        \
    do {                                                                                                                  \
        cudaError_t e = cudaFuncSetAttribute(conv_tc_multi_kernel<BN_, ST_>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                             FwdSmem<BN_, ST_>::kBytes);                                                  \
        if (e != cudaSuccess) return fail(EFFDET_ERR_LAUNCH, "conv2d_multi(tc): smem opt-in: %s", cudaGetErrorString(e)); \
        conv_tc_multi_kernel<BN_, ST_><<<grid, kFwdThreads, FwdSmem<BN_, ST_>::kBytes, st>>>(map, ma, kblocks);            \
    } while (0)
    if (BN == 64) EFFDET_TCM_LAUNCH(64, 4);
    else if (BN == 128) EFFDET_TCM_LAUNCH(128, 3);
    else EFFDET_TCM_LAUNCH(256, 2);
#undef EFFDET_TCM_LAUNCH
    return launch_status("conv_tc_multi_kernel");
}

bool wgrad_tc_eligible(const effdet_wgrad_args* a) {
    if (a->precision != 1 || a->Cin % 4 || a->Cout % 4 || a->Cin < 16 || a->Cout < 16) return false;
    WgGeom g;
    const bool tma_ok = (a->ws_x || a->x_planes) && (a->ws_dy || a->dy_planes) && wg_geometry(a->B, a->H, a->W, &g);
    return tma_ok || (!a->a_scale && !a->in_scale && !a->dy_planes && !a->x_planes);     // the gather-producer fallback has no input prologue
}

bool wg_geometry(int B, int H, int W, WgGeom* g) {
    const int Wb = W <= 64 ? W : 64;
    if (W % Wb) return false;
    int Hb = 1;
    for (int h = 1; h <= H && Wb * h <= 64; ++h)
        if (H % h == 0) Hb = h;
    int Bb = 64 / (Wb * Hb);
    if (Bb < 1) Bb = 1;
    if (Bb > B) Bb = B;
    const int ks = Wb * Hb * Bb;
    if (ks < 16 || ks % 16) return false;
    g->Wb = Wb; g->Hb = Hb; g->Bb = Bb;
    g->nbx = W / Wb; g->nby = H / Hb; g->nbb = (B + Bb - 1) / Bb;
    g->kstage = ks;
    return true;
}

// 5-D tensor map (channel, x, y, image, plane) over bf16 hi/lo planes [2][B][H][W][pitch]; C = channel extent (<= pitch):
// channels beyond C and pixels outside the image are zero-filled by the hardware
int planes_map(EncodeTiledFn enc, CUtensorMap* map, void* base, int B, int H, int W, int C, int pitch, const WgGeom& g) {
    const cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, 2};
    const cuuint64_t gstr[4] = {(cuuint64_t)pitch * 2, (cuuint64_t)W * pitch * 2, (cuuint64_t)H * W * pitch * 2,
                                (cuuint64_t)B * H * W * pitch * 2};
    const cuuint32_t box[5] = {64, (cuuint32_t)g.Wb, (cuuint32_t)g.Hb, (cuuint32_t)g.Bb, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(EFFDET_ERR_LAUNCH, "wgrad(tc): cuTensorMapEncodeTiled failed (%d)", (int)r);
    return EFFDET_OK;
}

// TMA-fed weight gradient; returns 1 when the geometry has no legal pixel box (caller falls back)
static int wgrad_tc2_launch(const effdet_wgrad_args* a, cudaStream_t st) {
    WgGeom g;
    EncodeTiledFn enc = encode_fn();
    if (!enc || !(a->ws_x || a->x_planes) || !(a->ws_dy || a->dy_planes) || !wg_geometry(a->B, a->H, a->W, &g)) return 1;
    const int HW = a->H * a->W;
    const int cin_pad = conv_tc_kpad(a->Cin), cout_pad = conv_tc_kpad(a->Cout);
    int s = EFFDET_OK;
    if (!a->x_planes) {
        int blocks = cdiv((long long)a->B * HW * (cin_pad / 8), 256);
        if (blocks > 148 * 16) blocks = 148 * 16;
        split_planes_kernel<<<blocks, 256, 0, st>>>(a->x, a->x_bstride, a->a_scale, (__nv_bfloat16*)a->ws_x, a->B, HW, a->Cin, cin_pad,
                                                    a->in_scale, a->in_shift);
        s = launch_status("split_planes_kernel");
        if (s) return s;
    }
    CUtensorMap mdy, mx;
    if (a->dy_planes) {        // dy arrives pre-split (unpadded pitch; TMA zero-fills the channels beyond Cout)
        if ((s = planes_map(enc, &mdy, const_cast<void*>(a->dy_planes), a->B, a->H, a->W, a->Cout, (a->Cout + 7) / 8 * 8, g))) return s;
    } else {
        if ((s = split_dy_launch(a, cout_pad, st))) return s;      // also accumulates the bias gradient
        if ((s = planes_map(enc, &mdy, a->ws_dy, a->B, a->H, a->W, cout_pad, cout_pad, g))) return s;
    }
    if (a->x_planes) {
        if ((s = planes_map(enc, &mx, const_cast<void*>(a->x_planes), a->B, a->H, a->W, a->Cin, (a->Cin + 7) / 8 * 8, g))) return s;
    } else {
        if ((s = planes_map(enc, &mx, a->ws_x, a->B, a->H, a->W, cin_pad, cin_pad, g))) return s;
    }
    const int taps = a->ksize * a->ksize;
    const int BC = a->Cin > 64 ? 256 : 64;
    const int ctiles = cdiv(a->Cin, BC), ntiles = cdiv(a->Cout, kTileM);
    const int nchunks = g.nbx * g.nby * g.nbb;
    // split-K so that the grid is as close as possible to (but not above) two full waves of 148 CTAs
    int splits = (148 * 2) / (ctiles * ntiles * taps);
    if (splits < 1) splits = 1;
    if (splits > cdiv(nchunks, 4)) splits = cdiv(nchunks, 4);
    int cps = cdiv(nchunks, splits);
    splits = cdiv(nchunks, cps);
    dim3 grid(ctiles * ntiles, taps, splits);
    cudaError_t e;
    if (BC == 256) {
        constexpr int ST = 2;
        e = cudaFuncSetAttribute(wgrad_tc2_kernel<256, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem<256, ST>::kBytes);
        if (e != cudaSuccess) return fail(EFFDET_ERR_LAUNCH, "wgrad(tc): smem opt-in: %s", cudaGetErrorString(e));
        wgrad_tc2_kernel<256, ST><<<grid, kTcThreads, WgSmem<256, ST>::kBytes, st>>>(mdy, mx, *a, g, cps, ctiles);
    } else {
        constexpr int ST = 4;
        e = cudaFuncSetAttribute(wgrad_tc2_kernel<64, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem<64, ST>::kBytes);
        if (e != cudaSuccess) return fail(EFFDET_ERR_LAUNCH, "wgrad(tc): smem opt-in: %s", cudaGetErrorString(e));
        wgrad_tc2_kernel<64, ST><<<grid, kTcThreads, WgSmem<64, ST>::kBytes, st>>>(mdy, mx, *a, g, cps, ctiles);
    }
    return launch_status("wgrad_tc2_kernel");
}

// all levels in one launch; returns 1 when some level cannot use the TMA path (caller falls back per level)
int wgrad_tc2_multi_launch(const effdet_wgrad_args* levels, int nlevels, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    if (!enc || nlevels > kWgMaxLevels) return 1;
    WgMaps maps;
    WgMultiArgs ma;
    memset(&ma, 0, sizeof(ma));
    const effdet_wgrad_args* a0 = &levels[0];
    const int cin_pad = conv_tc_kpad(a0->Cin), cout_pad = conv_tc_kpad(a0->Cout);
    int chunks = 0;
    for (int l = 0; l < nlevels; ++l) {
        const effdet_wgrad_args* a = &levels[l];
        if (!(a->ws_x || a->x_planes) || !(a->ws_dy || a->dy_planes) || a->a_scale || a->in_scale ||
            !wg_geometry(a->B, a->H, a->W, &ma.g[l]))
            return 1;
        ma.chunk_begin[l] = chunks;
        chunks += ma.g[l].nbx * ma.g[l].nby * ma.g[l].nbb;
    }
    for (int l = nlevels; l <= kWgMaxLevels; ++l) ma.chunk_begin[l] = chunks;
    ma.nlevels = nlevels;
    ma.dw = a0->dw;
    ma.Cin = a0->Cin; ma.Cout = a0->Cout; ma.ksize = a0->ksize;
    for (int l = 0; l < nlevels; ++l) {
        const effdet_wgrad_args* a = &levels[l];
        const int HW = a->H * a->W;
        int s = EFFDET_OK;
        if (a->x_planes) {          // operands that already live as planes: no split pass at all
            if ((s = planes_map(enc, &maps.x[l], const_cast<void*>(a->x_planes), a->B, a->H, a->W, a->Cin, (a->Cin + 7) / 8 * 8, ma.g[l])))
                return s;
        } else {
            int blocks = cdiv((long long)a->B * HW * (cin_pad / 8), 256);
            if (blocks > 148 * 16) blocks = 148 * 16;
            split_planes_kernel<<<blocks, 256, 0, st>>>(a->x, a->x_bstride, nullptr, (__nv_bfloat16*)a->ws_x, a->B, HW, a->Cin, cin_pad);
            s = launch_status("split_planes_kernel");
            if (s) return s;
            if ((s = planes_map(enc, &maps.x[l], a->ws_x, a->B, a->H, a->W, cin_pad, cin_pad, ma.g[l]))) return s;
        }
        if (a->dy_planes) {
            if ((s = planes_map(enc, &maps.dy[l], const_cast<void*>(a->dy_planes), a->B, a->H, a->W, a->Cout, (a->Cout + 7) / 8 * 8,
                                ma.g[l])))
                return s;
        } else {
            if ((s = split_dy_launch(a, cout_pad, st))) return s;  // also accumulates the bias gradient
            if ((s = planes_map(enc, &maps.dy[l], a->ws_dy, a->B, a->H, a->W, cout_pad, cout_pad, ma.g[l]))) return s;
        }
    }
    for (int l = nlevels; l < kWgMaxLevels; ++l) { maps.dy[l] = maps.dy[0]; maps.x[l] = maps.x[0]; }
    const int taps = a0->ksize * a0->ksize;
    const int BC = a0->Cin > 64 ? 256 : 64;
    const int ctiles = cdiv(a0->Cin, BC), ntiles = cdiv(a0->Cout, kTileM);
    int splits = (148 * 2) / (ctiles * ntiles * taps);
    if (splits < 1) splits = 1;
    if (splits > cdiv(chunks, 4)) splits = cdiv(chunks, 4);
    int cps = cdiv(chunks, splits);
    splits = cdiv(chunks, cps);
    cudaError_t e;
    dim3 grid(ctiles * ntiles, taps, splits);
    if (BC == 256) {
        constexpr int ST = 2;
        e = cudaFuncSetAttribute(wgrad_tc2_multi_kernel<256, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem<256, ST>::kBytes);
        if (e != cudaSuccess) return fail(EFFDET_ERR_LAUNCH, "wgrad_multi(tc): smem opt-in: %s", cudaGetErrorString(e));
        wgrad_tc2_multi_kernel<256, ST><<<grid, kTcThreads, WgSmem<256, ST>::kBytes, st>>>(maps, ma, cps, ctiles);
    } else {
        constexpr int ST = 4;
        e = cudaFuncSetAttribute(wgrad_tc2_multi_kernel<64, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem<64, ST>::kBytes);
        if (e != cudaSuccess) return fail(EFFDET_ERR_LAUNCH, "wgrad_multi(tc): smem opt-in: %s", cudaGetErrorString(e));
        wgrad_tc2_multi_kernel<64, ST><<<grid, kTcThreads, WgSmem<64, ST>::kBytes, st>>>(maps, ma, cps, ctiles);
    }
    return launch_status("wgrad_tc2_multi_kernel");
}

int wgrad_tc_launch(const effdet_wgrad_args* a, cudaStream_t st, bool* dbias_done) {
    *dbias_done = false;
    {
        const int r = wgrad_tc2_launch(a, st);
        if (r == 0) *dbias_done = true;     // the TMA path folds the bias gradient into its dy split pass
        if (r <= 0) return r;       // launched (0) or failed (<0); 1 = geometry unsupported -> gather kernel
    }
    const long long Mll = (long long)a->B * a->H * a->W;
    const int M = (int)Mll, HW = a->H * a->W;
    const int taps = a->ksize * a->ksize;
    const int BC = a->Cin > 64 ? 256 : 64;
    const int ctiles = cdiv(a->Cin, BC), ntiles = cdiv(a->Cout, kTileM);
    const int nchunks = cdiv(M, kTileK);
    int splits = cdiv(148 * 2, ctiles * ntiles * taps);
    if (splits < 1) splits = 1;
    if (splits > cdiv(nchunks, 8)) splits = cdiv(nchunks, 8);
    int cps = cdiv(nchunks, splits);
    splits = cdiv(nchunks, cps);
    dim3 grid(ctiles * ntiles, taps, splits);
    cudaError_t e;
    if (BC == 256) {
        constexpr int ST = 2;
        e = cudaFuncSetAttribute(wgrad_tc_kernel<256, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem<256, ST>::kBytes);
        if (e != cudaSuccess) return fail(EFFDET_ERR_LAUNCH, "wgrad(tc): smem opt-in: %s", cudaGetErrorString(e));
        wgrad_tc_kernel<256, ST><<<grid, kTcThreads, WgSmem<256, ST>::kBytes, st>>>(*a, M, HW, cps, ctiles);
    } else {
        constexpr int ST = 4;
        e = cudaFuncSetAttribute(wgrad_tc_kernel<64, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem<64, ST>::kBytes);
        if (e != cudaSuccess) return fail(EFFDET_ERR_LAUNCH, "wgrad(tc): smem opt-in: %s", cudaGetErrorString(e));
        wgrad_tc_kernel<64, ST><<<grid, kTcThreads, WgSmem<64, ST>::kBytes, st>>>(*a, M, HW, cps, ctiles);
    }
    return launch_status("wgrad_tc_kernel");
}

}  // namespace effdet

using namespace effdet;

extern "C" int effdet_conv_tc_kpad(int channels) { return conv_tc_kpad(channels); }
extern "C" int effdet_wgrad_tc_geometry_ok(int B, int H, int W) {
    WgGeom g;
    return wg_geometry(B, H, W, &g) ? 1 : 0;
}

namespace effdet {   // defined in conv_simt.cu
int colsum_launch(const float* x, float* out, long long M, int N, long long HW, long long bstride, int device,
                  effdet_stream_t stream);
}

extern "C" int effdet_conv2d_wgrad_multi(const effdet_wgrad_args* levels, int nlevels, int device, effdet_stream_t stream) {
    EFFDET_REQUIRE(levels && nlevels >= 1, "conv2d_wgrad_multi: no levels");
    bool same = true;
    for (int l = 0; l < nlevels; ++l) {
        const effdet_wgrad_args* a = &levels[l];
        EFFDET_REQUIRE((a->x || a->x_planes) && (a->dy || a->dy_planes) && a->dw, "conv2d_wgrad_multi: null tensor");
        EFFDET_REQUIRE(!(a->dy_planes && a->dbias), "conv2d_wgrad_multi: dy_planes excludes dbias (the producer supplies the column sums)");
        same = same && a->dw == levels[0].dw && a->dbias == levels[0].dbias && a->Cin == levels[0].Cin &&
               a->Cout == levels[0].Cout && a->ksize == levels[0].ksize && a->precision == 1 && wgrad_tc_eligible(a);
    }
    if (same && nlevels > 1) {
        EFFDET_DEVICE(device);
        const int r = wgrad_tc2_multi_launch(levels, nlevels, (cudaStream_t)stream);
        if (r < 0) return r;
        if (r == 0) return EFFDET_OK;          // bias gradient was fused into the dy split pass
    }
    for (int l = 0; l < nlevels; ++l) {
        const int s = effdet_conv2d_wgrad(&levels[l], device, stream);
        if (s) return s;
    }
    return EFFDET_OK;
}

extern "C" int effdet_conv2d_multi(const effdet_conv_args* levels, int nlevels, int device, effdet_stream_t stream) {
    EFFDET_REQUIRE(levels && nlevels >= 1 && nlevels <= kMaxLevels, "conv2d_multi: 1..%d levels", kMaxLevels);
    bool tc = true;
    for (int l = 0; l < nlevels; ++l) {
        const effdet_conv_args* a = &levels[l];
        EFFDET_REQUIRE(a->x && a->w && a->y, "conv2d_multi: null tensor");
        EFFDET_REQUIRE(a->Cin == levels[0].Cin && a->Cout == levels[0].Cout && a->ksize == levels[0].ksize &&
                           a->act == levels[0].act && a->w == levels[0].w && a->w_tc == levels[0].w_tc &&
                           a->bias == levels[0].bias,
                       "conv2d_multi: all levels must share weights, bias, channels and activation");
        const bool mb = a->a_scale || a->z || a->scale || a->row_scale || a->in_scale;
        tc = tc && conv_tc_eligible(a) && !mb && (long long)a->B * a->H * a->W < (1ll << 31);
        EFFDET_REQUIRE(aligned16(a->x) && aligned16(a->y) && aligned16(a->residual) && aligned16(a->mask_src) &&
                           a->x_bstride % 4 == 0 && a->y_bstride % 4 == 0 && a->r_bstride % 4 == 0 && a->m_bstride % 4 == 0,
                       "conv2d_multi: alignment");
    }
    if (tc && nlevels > 1) {
        EFFDET_DEVICE(device);
        return conv_tc_multi_launch(levels, nlevels, (cudaStream_t)stream);
    }
    for (int l = 0; l < nlevels; ++l) {          // exact-fp32 mode / unsupported shapes: one launch per level
        int s = effdet_conv2d(&levels[l], device, stream);
        if (s) return s;
    }
    return EFFDET_OK;
}

extern "C" int effdet_pack_conv_weight_tc(const float* w_oihw, void* w_fwd, void* w_dgrad, int Cout, int Cin, int ksize,
                                          int device, effdet_stream_t stream) {
    EFFDET_REQUIRE(w_oihw && w_fwd && Cout > 0 && Cin > 0 && (ksize == 1 || ksize == 3), "pack_conv_weight_tc: bad arguments");
    EFFDET_DEVICE(device);
    const int taps = ksize * ksize;
    cudaStream_t st = (cudaStream_t)stream;
    {
        const int kpad = conv_tc_kpad(Cin);
        const long long plane = (long long)Cout * taps * kpad;
        int blocks = cdiv(plane, 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        pack_weight_tc_kernel<<<blocks, 256, 0, st>>>(w_oihw, (__nv_bfloat16*)w_fwd, Cout, Cin, taps, kpad, 0);
        int s = launch_status("pack_weight_tc_kernel");
        if (s) return s;
    }
    if (w_dgrad) {
        const int kpad = conv_tc_kpad(Cout);
        const long long plane = (long long)Cin * taps * kpad;
        int blocks = cdiv(plane, 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        pack_weight_tc_kernel<<<blocks, 256, 0, st>>>(w_oihw, (__nv_bfloat16*)w_dgrad, Cout, Cin, taps, kpad, 1);
        return launch_status("pack_weight_tc_kernel");
    }
    return EFFDET_OK;
}
