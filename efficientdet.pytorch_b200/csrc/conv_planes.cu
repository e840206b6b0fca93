// Dense 3x3 / 1x1 convolution whose activations live in HBM as bf16 hi/lo planes (x ~= hi + lo, the tensor-core operand
// format) -- the RetinaHead towers end to end (models/retinahead.py:67-132; 95 % of the model's FLOPs).
//
// Round 1 kept fp32 activations: every conv re-split its input on the fly with eight gather warps (LDG + cvt + st.shared)
// and every weight gradient needed a separate split pass over both operands (~190 launches, ~3 ms per step).  With the
// planes format the im2col gather IS a TMA load: a 5-D tensor map (channel, x, y, image, plane) with a pixel box of
// Wb x Hb x Bb pixels delivers, for one tap, 64 channels of 16..64 pixels as consecutive 128-byte rows in the canonical
// K-major SWIZZLE_128B layout; the tap shift is a coordinate offset and the hardware's out-of-bounds zero fill is the
// convolution's zero padding.  The kernel is the canonical Blackwell GEMM:
//   warp 0      TMA producer (activation boxes hi/lo + weight tile hi/lo -> mbarrier complete_tx), ring of stages
//   warp 1      tcgen05.mma issuer, 3 MMAs per K16 (lo*hi, hi*lo, hi*hi), two TMEM accumulators (2 x 256 columns)
//   warps 2-5   epilogue: tcgen05.ld -> bias / ReLU / sigmoid / ReLU-mask / residual -> bf16 hi/lo planes (the next
//               layer's operand) and / or fp32 (head outputs, data gradient w.r.t. the BiFPN features); optional
//               per-channel column sums of what was stored (= the bias gradient of the producing layer) reduced by warp
//               shuffles, one atomic per column per warp
// persistent over (pixel tile, channel tile) units of all pyramid levels that share the weights.
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace effdet {

constexpr int kPlMaxLevels = 8;
constexpr int kPlThreads = 192;
constexpr int kPlA = 128 * 128;            // one plane of the activation tile: 128 pixel rows x 64 channels (bf16)
// BN = output channels per tile (TMEM: 2 x BN columns); the ring depth is what fits next to it
template <int BN>
struct PlCfg {
    static constexpr int kStages = BN == 256 ? 2 : (BN == 128 ? 3 : 4);
    static constexpr int kB = BN * 128;    // one plane of the weight tile: BN output channels x 64 input channels
    static constexpr int kStage = 2 * kPlA + 2 * kB;
    static constexpr int kSmem = kStages * kStage + 1024 + 256 + BN * 4;
};

struct PlLevel {
    int B, H, W;
    WgGeom g;
    int tile_begin;                        // first pixel tile of this level
    int nboxes;                            // pixel boxes of this level
    float* y;                              // fp32 output [B][H*W][Cout] with image stride y_bstride, or NULL
    long long y_bstride;
    __nv_bfloat16* y_planes;               // bf16 hi/lo output planes [2][B*H*W][opitch], or NULL
    const __nv_bfloat16* mask_planes;      // ReLU-backward mask source [2][B*H*W][opitch] (value > 0 keeps the gradient), or NULL
    const float* residual;                 // fp32 [B][H*W][Cout] added last, or NULL
    long long r_bstride;
};
struct PlArgs {
    PlLevel lv[kPlMaxLevels];
    int nlevels, total_tiles, ntn;
    int Cin, Cout, ksize, act, kblocks, opitch;
    const float* bias;                     // [Cout] or NULL
    float* colsum;                         // [Cout] += column sums of the stored values, or NULL
};
struct PlMaps {
    CUtensorMap x[kPlMaxLevels];
};

template <int BN>
__global__ void __launch_bounds__(kPlThreads, 1)
conv_planes_kernel(const __grid_constant__ PlMaps maps, const __grid_constant__ CUtensorMap wmap, const __grid_constant__ PlArgs P) {
    constexpr int kPlBN = BN, kPlStages = PlCfg<BN>::kStages, kPlB = PlCfg<BN>::kB, kPlStage = PlCfg<BN>::kStage;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned AND still a shared-space pointer (LDS/STS, not generic LD/ST)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kPlStages * kPlStage);
    uint64_t* empty_bar = full_bar + kPlStages;
    uint64_t* acc_full = empty_bar + kPlStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* chan = reinterpret_cast<float*>(smem + kPlStages * kPlStage + 256);      // bias of the current channel tile

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int taps = P.ksize * P.ksize, pad = P.ksize / 2;
    const int KT = taps * P.kblocks;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kPlStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 4);
        }
        fence_barrier_init();
        tma_prefetch_desc(&wmap);
    }
    if (warp == 1) tmem_alloc<2 * kPlBN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // unit -> (level, first box of the pixel tile, first output channel)
    auto decode = [&](int unit, int& l, int& box0, int& n0) {
        const int mt = unit / P.ntn;
        n0 = (unit - mt * P.ntn) * kPlBN;
        l = 0;
        while (l + 1 < P.nlevels && mt >= P.lv[l + 1].tile_begin) ++l;
        box0 = (mt - P.lv[l].tile_begin) * (128 / P.lv[l].g.kstage);
    };

    if (warp == 0) {
        // ---------------- TMA producer ------------------------------------------------------------------------------------
        if (lane == 0) {
            uint32_t it = 0;
            for (int unit = blockIdx.x; unit < P.total_tiles; unit += gridDim.x) {
                int l, box0, n0;
                decode(unit, l, box0, n0);
                const PlLevel& L = P.lv[l];
                const int ks = L.g.kstage, nbox = 128 / ks;
                const int nvalid = min(nbox, L.nboxes - box0);
                for (int kt = 0; kt < KT; ++kt, ++it) {
                    const int s = it % kPlStages;
                    const uint32_t ph = (it / kPlStages) & 1;
                    const int tap = kt / P.kblocks, kb = kt - tap * P.kblocks;
                    const int dy = tap / P.ksize - pad, dx = tap % P.ksize - pad;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* a_hi = smem + s * kPlStage;
                    uint8_t* b_hi = a_hi + 2 * kPlA;
                    mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(2 * nvalid * ks * 128 + 2 * kPlB));
                    for (int q = 0; q < nvalid; ++q) {
                        int ch = box0 + q;
                        const int bx = ch % L.g.nbx;
                        ch /= L.g.nbx;
                        const int by = ch % L.g.nby;
                        const int bb = ch / L.g.nby;
                        const int x0 = bx * L.g.Wb + dx, y0 = by * L.g.Hb + dy, b0 = bb * L.g.Bb;
                        tma_load_5d(a_hi + q * ks * 128, &maps.x[l], &full_bar[s], kb * 64, x0, y0, b0, 0);
                        tma_load_5d(a_hi + kPlA + q * ks * 128, &maps.x[l], &full_bar[s], kb * 64, x0, y0, b0, 1);
                    }
                    tma_load_3d(b_hi, &wmap, &full_bar[s], kt * 64, n0, 0);
                    tma_load_3d(b_hi + kPlB, &wmap, &full_bar[s], kt * 64, n0, 1);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer --------------------------------------------------------------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(128, kPlBN, 0, 0);
            uint32_t it = 0, iu = 0;
            for (int unit = blockIdx.x; unit < P.total_tiles; unit += gridDim.x, ++iu) {
                const uint32_t acc = iu & 1, pacc = (iu >> 1) & 1;
                mbar_wait(&acc_empty[acc], pacc ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * kPlBN;
                for (int kt = 0; kt < KT; ++kt, ++it) {
                    const int s = it % kPlStages;
                    const uint32_t ph = (it / kPlStages) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + s * kPlStage);
                    const uint32_t a_lo = a_hi + kPlA;
                    const uint32_t b_hi = a_hi + 2 * kPlA;
                    const uint32_t b_lo = b_hi + kPlB;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
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
        // ---------------- epilogue warps ----------------------------------------------------------------------------------
        const int etid = threadIdx.x - 64;
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;                         // row of the tile owned by this thread
        uint32_t iu = 0;
        int chan_n0 = -1;
        for (int unit = blockIdx.x; unit < P.total_tiles; unit += gridDim.x, ++iu) {
            int l, box0, n0;
            decode(unit, l, box0, n0);
            const PlLevel& L = P.lv[l];
            if (n0 != chan_n0) {
                named_bar_sync(1, 128);
                for (int i = etid; i < kPlBN; i += 128) chan[i] = (n0 + i < P.Cout && P.bias) ? __ldg(P.bias + n0 + i) : 0.f;
                named_bar_sync(1, 128);
                chan_n0 = n0;
            }
            // pixel of row r: box q of the tile, position i inside the box (x fastest, then y, then image)
            const int ks = L.g.kstage;
            const int q = r / ks, i = r - q * ks;
            int ch = box0 + q;
            bool row_ok = ch < L.nboxes;
            const int bx = ch % L.g.nbx;
            ch /= L.g.nbx;
            const int by = ch % L.g.nby;
            const int bb = ch / L.g.nby;
            const int wh = L.g.Wb * L.g.Hb;
            const int bi = i / wh, rem = i - bi * wh;
            const int yy = rem / L.g.Wb, xx = rem - yy * L.g.Wb;
            const int b = bb * L.g.Bb + bi, y = by * L.g.Hb + yy, x = bx * L.g.Wb + xx;
            row_ok = row_ok && b < L.B;
            const long long pixb = (long long)y * L.W + x;                     // pixel inside its image
            const long long pix = (long long)b * L.H * L.W + pixb;             // pixel in the planes
            const long long plane = (long long)L.B * L.H * L.W * P.opitch;
            const uint32_t acc = iu & 1, pacc = (iu >> 1) & 1;
            mbar_wait(&acc_full[acc], pacc);
            tc_fence_after();
            const int ncols = min(kPlBN, P.Cout - n0);
            const int nchunks = (ncols + 31) >> 5;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kPlBN;
#pragma unroll 1
            for (int cc = 0; cc < nchunks; ++cc) {
                uint32_t a32[32];
                tmem_ld32(taddr + cc * 32, a32);
                if (cc == nchunks - 1) {                           // accumulator drained
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[acc]);
                }
                const int nb = n0 + cc * 32;                        // first channel of the chunk
                float v[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    float t = __uint_as_float(a32[k]) + chan[cc * 32 + k];
                    if (P.act == EFFDET_ACT_RELU) t = fmaxf(t, 0.f);
                    else if (P.act == EFFDET_ACT_SIGMOID) t = sigmoidf_(t);
                    v[k] = t;
                }
                if (row_ok && L.residual) {
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4) {
                        if (nb + k4 * 4 >= P.Cout) break;
                        const float4 rv = ldg4(L.residual + (long long)b * L.r_bstride + pixb * P.Cout + nb + k4 * 4);
                        v[k4 * 4] += rv.x; v[k4 * 4 + 1] += rv.y; v[k4 * 4 + 2] += rv.z; v[k4 * 4 + 3] += rv.w;
                    }
                }
                if (row_ok && L.mask_planes) {                     // gradient passes where the forward activation was > 0
                    const __nv_bfloat16* mh = L.mask_planes + pix * P.opitch + nb;
#pragma unroll
                    for (int k8 = 0; k8 < 4; ++k8) {
                        if (nb + k8 * 8 >= P.Cout) break;
                        const uint4 hv = __ldg(reinterpret_cast<const uint4*>(mh + k8 * 8));
                        const uint4 lv = __ldg(reinterpret_cast<const uint4*>(mh + plane + k8 * 8));
                        const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const uint32_t hb = (hw[e >> 1] >> ((e & 1) * 16)) & 0xffffu, lb = (lw[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
                            // bf16 bits: positive and non-zero  <=>  sign clear and magnitude bits set
                            const bool pos = (hb & 0x7fffu) ? !(hb & 0x8000u) : ((lb & 0x7fffu) && !(lb & 0x8000u));
                            if (!pos) v[k8 * 8 + e] = 0.f;
                        }
                    }
                }
                if (!row_ok) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) v[k] = 0.f;
                }
                if (row_ok && L.y) {
                    float* yo = L.y + (long long)b * L.y_bstride + pixb * P.Cout + nb;
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4) {
                        if (nb + k4 * 4 >= P.Cout) break;
                        st4(yo + k4 * 4, make_float4(v[k4 * 4], v[k4 * 4 + 1], v[k4 * 4 + 2], v[k4 * 4 + 3]));
                    }
                }
                if (row_ok && L.y_planes) {
                    __nv_bfloat16* ph = L.y_planes + pix * P.opitch + nb;
#pragma unroll
                    for (int k8 = 0; k8 < 4; ++k8) {
                        if (nb + k8 * 8 >= P.Cout) break;
                        uint4 hi, lo;
                        split8(make_float4(v[k8 * 8], v[k8 * 8 + 1], v[k8 * 8 + 2], v[k8 * 8 + 3]),
                               make_float4(v[k8 * 8 + 4], v[k8 * 8 + 5], v[k8 * 8 + 6], v[k8 * 8 + 7]), hi, lo);
                        *reinterpret_cast<uint4*>(ph + k8 * 8) = hi;
                        *reinterpret_cast<uint4*>(ph + plane + k8 * 8) = lo;
                    }
                }
                if (P.colsum) {
                    // warp transpose-reduce: afterwards v[0] of lane j is the sum over the warp's 32 rows of column j
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
                        const bool upper = (lane & off) != 0;
#pragma unroll
                        for (int k = 0; k < off; ++k) {
                            const float send = upper ? v[k] : v[k + off];
                            const float keep = upper ? v[k + off] : v[k];
                            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                        }
                    }
                    if (nb + lane < P.Cout) atomicAdd(P.colsum + nb + lane, v[0]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<2 * kPlBN>(tmem_base);
    }
}

// fp32 [B][HW][C] (image stride bstride) -> bf16 hi/lo planes [2][B*HW][pitch]; optionally multiplied by p*(1-p) of a
// second tensor (sigmoid backward, models/retinahead.py:121) and optionally reduced into per-channel column sums (the
// bias gradient) on the way -- one read of the gradient instead of three passes
__global__ void __launch_bounds__(256) to_planes_kernel(const float* __restrict__ x, long long x_bstride, const float* __restrict__ prob,
                                                        long long p_bstride, __nv_bfloat16* __restrict__ out, float* __restrict__ colsum,
                                                        int B, int HW, int C, int pitch, int rows_per_block) {
    __shared__ float red[256 * 8];
    const int cv8 = pitch / 8;
    const int cvb = cv8 < 256 ? cv8 : 256;
    const int rows = 256 / cvb;
    const int tr = threadIdx.x / cvb, tc = threadIdx.x - tr * cvb;
    const int j = blockIdx.y * cvb + tc;
    const bool active = tr < rows && j < cv8;
    const long long nrows = (long long)B * HW;
    const long long plane = nrows * pitch;
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
                const float* q = x + (long long)b * x_bstride + pix * C + c;
                v0 = ldg4(q);
                if (c + 4 < C) v1 = ldg4(q + 4);
                if (prob) {
                    const float* pp = prob + (long long)b * p_bstride + pix * C + c;
                    const float4 p0 = ldg4(pp);
                    v0 = make_float4(v0.x * p0.x * (1.f - p0.x), v0.y * p0.y * (1.f - p0.y), v0.z * p0.z * (1.f - p0.z),
                                     v0.w * p0.w * (1.f - p0.w));
                    if (c + 4 < C) {
                        const float4 p1 = ldg4(pp + 4);
                        v1 = make_float4(v1.x * p1.x * (1.f - p1.x), v1.y * p1.y * (1.f - p1.y), v1.z * p1.z * (1.f - p1.z),
                                         v1.w * p1.w * (1.f - p1.w));
                    }
                }
            }
            uint4 hi, lo;
            split8(v0, v1, hi, lo);
            *reinterpret_cast<uint4*>(out + row * pitch + c) = hi;
            *reinterpret_cast<uint4*>(out + plane + row * pitch + c) = lo;
            s[0] += v0.x; s[1] += v0.y; s[2] += v0.z; s[3] += v0.w;
            s[4] += v1.x; s[5] += v1.y; s[6] += v1.z; s[7] += v1.w;
        }
    }
    if (colsum == nullptr) return;
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = s[i];
    __syncthreads();
    if (tr == 0 && j < cv8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float acc = 0.f;
            for (int rr = 0; rr < rows; ++rr) acc += red[(rr * cvb + tc) * 8 + i];
            const int c = j * 8 + i;
            if (c < C) atomicAdd(colsum + c, acc);
        }
    }
}

}  // namespace effdet

using namespace effdet;

extern "C" int effdet_to_planes(const float* x, int64_t x_bstride, const float* prob, int64_t p_bstride, void* planes,
                                float* colsum, int B, int HW, int C, int device, effdet_stream_t stream) {
    EFFDET_REQUIRE(x && planes && B > 0 && HW > 0 && C > 0 && C % 4 == 0, "to_planes: bad arguments");
    EFFDET_REQUIRE(aligned16(x) && aligned16(prob) && aligned16(planes) && x_bstride % 4 == 0 && p_bstride % 4 == 0,
                   "to_planes: alignment");
    EFFDET_DEVICE(device);
    const int pitch = (C + 7) / 8 * 8;
    const int cv8 = pitch / 8;
    const int cvb = cv8 < 256 ? cv8 : 256;
    const int rows = 256 / cvb;
    const long long nrows = (long long)B * HW;
    long long rpb = (nrows + 148 * 4 - 1) / (148 * 4);
    if (rpb < (long long)rows * 8) rpb = (long long)rows * 8;
    dim3 grid(cdiv(nrows, rpb), cdiv(cv8, cvb));
    to_planes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, x_bstride, prob, p_bstride, (__nv_bfloat16*)planes, colsum, B, HW, C,
                                                            pitch, (int)rpb);
    return launch_status("to_planes_kernel");
}

extern "C" int effdet_conv_planes_multi(const effdet_conv_planes_args* levels, int nlevels, int device, effdet_stream_t stream) {
    EFFDET_REQUIRE(levels && nlevels >= 1 && nlevels <= kPlMaxLevels, "conv_planes_multi: 1..%d levels", kPlMaxLevels);
    const effdet_conv_planes_args* a0 = &levels[0];
    EFFDET_REQUIRE(a0->w_tc && (a0->ksize == 1 || a0->ksize == 3) && a0->Cin % 4 == 0 && a0->Cout % 4 == 0 && a0->Cin >= 8 &&
                       a0->Cout >= 8,
                   "conv_planes_multi: needs the bf16 weight pack, k in {1,3}, channels %% 4 == 0");
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(EFFDET_ERR_UNSUPPORTED, "conv_planes_multi: cuTensorMapEncodeTiled unavailable");
    EFFDET_DEVICE(device);
    PlMaps maps;
    PlArgs P;
    memset(&P, 0, sizeof(P));
    const int opitch = (a0->Cout + 7) / 8 * 8, ipitch = (a0->Cin + 7) / 8 * 8;
    int tiles = 0;
    for (int l = 0; l < nlevels; ++l) {
        const effdet_conv_planes_args* a = &levels[l];
        EFFDET_REQUIRE(a->x_planes && (a->y || a->y_planes), "conv_planes_multi: null tensor");
        EFFDET_REQUIRE(a->Cin == a0->Cin && a->Cout == a0->Cout && a->ksize == a0->ksize && a->act == a0->act && a->w_tc == a0->w_tc &&
                           a->bias == a0->bias && a->colsum == a0->colsum,
                       "conv_planes_multi: all levels must share weights, bias, channels and activation");
        EFFDET_REQUIRE(aligned16(a->x_planes) && aligned16(a->y) && aligned16(a->y_planes) && aligned16(a->mask_planes) &&
                           aligned16(a->residual) && a->y_bstride % 4 == 0 && a->r_bstride % 4 == 0,
                       "conv_planes_multi: alignment");
        PlLevel& L = P.lv[l];
        L.B = a->B; L.H = a->H; L.W = a->W;
        if (!wg_geometry(a->B, a->H, a->W, &L.g))
            return fail(EFFDET_ERR_UNSUPPORTED, "conv_planes_multi: a %dx%dx%d map has no legal pixel box (check effdet_wgrad_tc_geometry_ok)",
                        a->B, a->H, a->W);
        L.nboxes = L.g.nbx * L.g.nby * L.g.nbb;
        L.tile_begin = tiles;
        tiles += cdiv(L.nboxes, 128 / L.g.kstage);
        L.y = a->y; L.y_bstride = a->y_bstride;
        L.y_planes = (__nv_bfloat16*)a->y_planes;
        L.mask_planes = (const __nv_bfloat16*)a->mask_planes;
        L.residual = a->residual; L.r_bstride = a->r_bstride;
        int s = planes_map(enc, &maps.x[l], const_cast<void*>(a->x_planes), a->B, a->H, a->W, a->Cin, ipitch, L.g);
        if (s) return s;
    }
    for (int l = nlevels; l < kPlMaxLevels; ++l) {
        maps.x[l] = maps.x[0];
        P.lv[l].tile_begin = tiles;
    }
    const int taps = a0->ksize * a0->ksize;
    const int kpad = conv_tc_kpad(a0->Cin);
    const int BN = a0->Cout <= 64 ? 64 : (a0->Cout <= 128 ? 128 : 256);
    CUtensorMap wmap;
    {
        const cuuint64_t gdim[3] = {(cuuint64_t)taps * kpad, (cuuint64_t)a0->Cout, 2};
        const cuuint64_t gstr[2] = {(cuuint64_t)taps * kpad * 2, (cuuint64_t)a0->Cout * taps * kpad * 2};
        const cuuint32_t box[3] = {64, (cuuint32_t)BN, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(a0->w_tc), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(EFFDET_ERR_LAUNCH, "conv_planes_multi: tensor map of the weights failed (%d)", (int)r);
    }
    P.nlevels = nlevels;
    P.ntn = cdiv(a0->Cout, BN);
    P.total_tiles = tiles * P.ntn;
    P.Cin = a0->Cin; P.Cout = a0->Cout; P.ksize = a0->ksize; P.act = a0->act;
    P.kblocks = kpad / 64;
    P.opitch = opitch;
    P.bias = a0->bias;
    P.colsum = a0->colsum;
    const int grid = P.total_tiles < 148 ? P.total_tiles : 148;
#define EFFDET_PL_LAUNCH(BN_)                                                                                              \
    do {                                                                                                                  \
        cudaError_t e = cudaFuncSetAttribute(conv_planes_kernel<BN_>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                             PlCfg<BN_>::kSmem);                                                          \
        if (e != cudaSuccess) return fail(EFFDET_ERR_LAUNCH, "conv_planes_multi: smem opt-in: %s", cudaGetErrorString(e)); \
        conv_planes_kernel<BN_><<<grid, kPlThreads, PlCfg<BN_>::kSmem, (cudaStream_t)stream>>>(maps, wmap, P);             \
    } while (0)
    if (BN == 64) EFFDET_PL_LAUNCH(64);
    else if (BN == 128) EFFDET_PL_LAUNCH(128);
    else EFFDET_PL_LAUNCH(256);
#undef EFFDET_PL_LAUNCH
    return launch_status("conv_planes_kernel");
}
