// Pointwise (1x1) convolution of the backbone / laterals as a persistent, warp-specialised tcgen05 GEMM.
//
//   y[M, Cout] = epilogue( prologue(x[M, Cin]) * W[Cin, Cout] )        M = B*H*W pixels (NHWC rows)
//
// These layers are HBM-bound (arithmetic intensity 7..80 FLOP/B): what matters is that every SM keeps
// tens of KB of loads and stores in flight and that nothing is serialised behind anything else.  Round 1 ran them on
// the 3x3 implicit-GEMM kernel: one tile per CTA, the gather warps doubled as epilogue warps, so load, MMA and store
// phases of a CTA never overlapped (0.15 of the HBM roofline).  Here a CTA is persistent over (m-tile, n-tile) units and
// five roles run concurrently, decoupled by mbarrier rings:
//
//   warp 0      TMA producer : fp32 activation boxes [128 rows x 32 channels] (SWIZZLE_128B, out-of-bounds rows / channels
//                              zero-filled by the hardware) + the bf16 hi/lo weight tile of the k-block -> ring of NS stages
//   NC warps    converters   : fp32 tile -> optional BN+swish (+ squeeze-excite gate) prologue -> bf16 hi + lo planes in the
//                              canonical K-major SWIZZLE_128B layout (x = hi + lo to 16 mantissa bits); NC = 8, or 16 when
//                              the prologue evaluates a swish per element (the elementwise work of the whole SM sits in
//                              these warps: with 4 of them the kernel was converter-latency-bound at 0.1-0.2 of the roofline)
//   warp 1      MMA issuer   : 3 x tcgen05.mma kind::f16 per K16 (lo*hi, hi*lo, hi*hi), fp32 accumulation in TMEM;
//                              two accumulators (2 x 128 columns) so unit u+1 accumulates while unit u drains
//   4 warps     epilogue     : tcgen05.ld -> bias / BN affine / drop-connect / residual -> either 128-byte-row swizzled
//                              staging + TMA tile store (plain outputs: every store is a full line), or direct stores
//                              (small Cout with residual / raw-output save)
//
// Reference ops replaced: MBConvBlock expand / project convs and their data gradients (models/efficientnet.py:85,96-104),
// BIFPN lateral convs (models/bifpn.py:96-105).
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace effdet {

constexpr int kPwMaxOut = 6;               // staging buffers of the TMA-store epilogue (stores in flight per SM)
constexpr int kPwA32Half = 128 * 128;      // bytes of one fp32 half-box: 128 rows x 32 floats
constexpr int kPwA16 = 2 * 128 * 128;      // bytes of one bf16 stage: hi plane + lo plane, 128 rows x 64 bf16 each
constexpr int kPwOut = 128 * 128;          // bytes of one staging buffer: 128 rows x 32 floats
constexpr int kPwMaxStages = 6;
constexpr int kPwBarBytes = 512;           // 20 mbarriers + tmem slot
constexpr int kPwChanBytes = 3 * 128 * 4;  // bias | scale | shift of the current n-tile

struct PwParams {
    effdet_conv_args a;
    int M, HW;
    int KB;          // k-blocks of 64 input channels
    int BN;          // output channels per n-tile (multiple of 32, <= 128)
    int ntn;         // n-tiles
    int units;       // m-tiles * n-tiles
    int NS;          // ring stages
    int a32_halves;  // fp32 half-boxes per stage (1 when Cin <= 32)
    int tma_store;   // epilogue through shared memory + TMA tile stores
    int nout;        // staging buffers (tma_store only)
    int bres;        // the whole packed weight (ntn x KB tiles) stays resident in shared memory: loaded once per CTA, the
                     // ring then carries activations only and is deeper (more loads in flight per SM)
    int planes;      // the A operand arrives pre-split as bf16 hi/lo planes [2][M][Cin]: TMA writes the MMA operand
                     // layout directly, no converter work (data gradients of the expand convs: the fused depthwise
                     // backward emits dz0 in this form for this kernel and for the weight gradient)
};

__device__ __forceinline__ void split4(const float4 v, uint2& hi, uint2& lo) {
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - __low2float(h0), v.y - __high2float(h0));
    const __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - __low2float(h1), v.w - __high2float(h1));
    hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}

template <int N>
__device__ __forceinline__ void tma_store_wait_read_upto(int pending) {       // run-time "at most `pending` groups still reading"
    if (pending >= N) tma_store_wait_read<N>();
    else if constexpr (N > 0) tma_store_wait_read_upto<N - 1>(pending);
}

template <int NC>
__global__ void __launch_bounds__((NC + 6) * 32, 1)
pw_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_y, const __grid_constant__ PwParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned AND still a shared-space pointer (LDS/STS, not generic LD/ST)
    const effdet_conv_args& p = P.a;
    const int a32_bytes = P.planes ? kPwA16 : P.a32_halves * kPwA32Half;
    const int b_plane = P.BN * 128;
    const int stage_bytes = a32_bytes + (P.bres ? 0 : 2 * b_plane);
    uint8_t* ring = smem;
    uint8_t* bres = ring + P.NS * stage_bytes;                         // [ntn][KB][hi | lo] weight tiles (resident mode)
    uint8_t* a16 = bres + (P.bres ? P.ntn * P.KB * 2 * b_plane : 0);
    uint8_t* outst = a16 + (P.planes ? 0 : 2 * kPwA16);
    uint64_t* ld_full = reinterpret_cast<uint64_t*>(outst + P.nout * kPwOut);
    uint64_t* ld_empty = ld_full + kPwMaxStages;
    uint64_t* a16_full = ld_empty + kPwMaxStages;
    uint64_t* a16_empty = a16_full + 2;
    uint64_t* acc_full = a16_empty + 2;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* b_full = acc_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);
    float* chan = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ld_full) + kPwBarBytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kPwMaxStages; ++s) {
            mbar_init(&ld_full[s], 1);
            mbar_init(&ld_empty[s], P.planes ? 1 : NC + 1);   // the converter warps + the MMA commit
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&a16_full[s], NC);
            mbar_init(&a16_empty[s], 1);
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 4);
        }
        mbar_init(b_full, 1);
        fence_barrier_init();
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        if (P.tma_store) tma_prefetch_desc(&map_y);
    }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer ----------------------------------------------
        if (lane == 0) {
            const uint32_t btx = P.bres ? 0u : (uint32_t)(2 * b_plane);
            if (P.bres) {                                          // all weight tiles once, on their own barrier
                mbar_arrive_expect_tx(b_full, (uint32_t)(P.ntn * P.KB * 2 * b_plane));
                for (int nt = 0; nt < P.ntn; ++nt)
                    for (int kb = 0; kb < P.KB; ++kb) {
                        uint8_t* bt = bres + (nt * P.KB + kb) * 2 * b_plane;
                        tma_load_3d(bt, &map_b, b_full, kb * 64, nt * P.BN, 0);
                        tma_load_3d(bt + b_plane, &map_b, b_full, kb * 64, nt * P.BN, 1);
                    }
            }
            uint32_t it = 0;
            for (int u = blockIdx.x; u < P.units; u += gridDim.x) {
                const int mt = u / P.ntn, nt = u - mt * P.ntn;
                const int m0 = mt * 128, n0 = nt * P.BN;
                for (int kb = 0; kb < P.KB; ++kb, ++it) {
                    const int s = it % P.NS;
                    const uint32_t ph = (it / P.NS) & 1;
                    mbar_wait(&ld_empty[s], ph ^ 1);
                    uint8_t* st = ring + s * stage_bytes;
                    if (P.planes) {
                        mbar_arrive_expect_tx(&ld_full[s], (uint32_t)kPwA16 + btx);
                        tma_load_3d(st, &map_a, &ld_full[s], kb * 64, m0, 0);
                        tma_load_3d(st + kPwA16 / 2, &map_a, &ld_full[s], kb * 64, m0, 1);
                    } else {
                        const int halves = (p.Cin - kb * 64 > 32) ? 2 : 1;
                        mbar_arrive_expect_tx(&ld_full[s], (uint32_t)(halves * kPwA32Half) + btx);
                        tma_load_2d(st, &map_a, &ld_full[s], kb * 64, m0);
                        if (halves == 2) tma_load_2d(st + kPwA32Half, &map_a, &ld_full[s], kb * 64 + 32, m0);
                    }
                    if (!P.bres) {
                        tma_load_3d(st + a32_bytes, &map_b, &ld_full[s], kb * 64, n0, 0);
                        tma_load_3d(st + a32_bytes + b_plane, &map_b, &ld_full[s], kb * 64, n0, 1);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer ------------------------------------------------
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(128, P.BN, 0, 0);
            uint32_t it = 0, iu = 0;
            if (P.bres) mbar_wait(b_full, 0);                       // resident weights landed
            for (int u = blockIdx.x; u < P.units; u += gridDim.x, ++iu) {
                const int nt_u = u % P.ntn;
                const uint32_t acc = iu & 1, pacc = (iu >> 1) & 1;
                mbar_wait(&acc_empty[acc], pacc ^ 1);          // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem_base + acc * 128;
                for (int kb = 0; kb < P.KB; ++kb, ++it) {
                    const int s = it % P.NS;
                    const uint32_t ph = (it / P.NS) & 1;
                    const uint32_t sa = it & 1, pha = (it >> 1) & 1;
                    mbar_wait(&ld_full[s], ph);                // weight tile (and, in planes mode, the A planes) landed
                    if (!P.planes) mbar_wait(&a16_full[sa], pha);      // converters published the bf16 planes
                    tc_fence_after();
                    const int valid = p.Cin - kb * 64;
                    const int ksteps = valid >= 64 ? 4 : (valid + 15) >> 4;
                    const uint32_t a_hi = P.planes ? smem_u32(ring + s * stage_bytes) : smem_u32(a16 + sa * kPwA16);
                    const uint32_t a_lo = a_hi + kPwA16 / 2;
                    const uint32_t b_hi = P.bres ? smem_u32(bres + (nt_u * P.KB + kb) * 2 * b_plane)
                                                 : smem_u32(ring + s * stage_bytes + a32_bytes);
                    const uint32_t b_lo = b_hi + b_plane;
                    for (int k = 0; k < ksteps; ++k) {
                        const uint64_t dah = umma_desc(a_hi + k * 32, 16, 1024), dal = umma_desc(a_lo + k * 32, 16, 1024);
                        const uint64_t dbh = umma_desc(b_hi + k * 32, 16, 1024), dbl = umma_desc(b_lo + k * 32, 16, 1024);
                        umma_bf16(d, dal, dbh, idesc, (kb | k) != 0);
                        umma_bf16(d, dah, dbl, idesc, 1);
                        umma_bf16(d, dah, dbh, idesc, 1);
                    }
                    if (!P.planes) umma_commit(&a16_empty[sa]);
                    umma_commit(&ld_empty[s]);
                }
                umma_commit(&acc_full[acc]);
            }
        }
    } else if (warp < 2 + NC) {
        // ------------------------------------------------ converters ------------------------------------------------
        constexpr int RP = 4 * NC;                             // rows per pass (8 threads per row)
        const int tid = threadIdx.x - 64;
        const int j = tid & 7, rbase = tid >> 3;
        const bool pro = p.in_scale != nullptr || p.a_scale != nullptr;
        uint32_t it = 0;
        for (int u = blockIdx.x; u < (P.planes ? 0 : P.units); u += gridDim.x) {      // planes mode: nothing to convert
            const int mt = u / P.ntn;
            const int m0 = mt * 128;
            const int b_first = m0 / P.HW;
            const bool one_image = (min(m0 + 127, P.M - 1) / P.HW) == b_first;   // the usual case: HW >> 128
            for (int kb = 0; kb < P.KB; ++kb, ++it) {
                const int s = it % P.NS;
                const uint32_t ph = (it / P.NS) & 1;
                const uint32_t sa = it & 1, pha = (it >> 1) & 1;
                const int valid = p.Cin - kb * 64;
                const int nh = valid > 32 ? 2 : 1;
                mbar_wait(&ld_full[s], ph);
                mbar_wait(&a16_empty[sa], pha ^ 1);
                const uint8_t* a32 = ring + s * stage_bytes;
                uint8_t* hi_pl = a16 + sa * kPwA16;
                uint8_t* lo_pl = hi_pl + kPwA16 / 2;
                for (int h = 0; h < nh; ++h) {
                    const int c = kb * 64 + h * 32 + 4 * j;
                    const bool col_ok = c < p.Cin;
                    float4 isc = make_float4(1.f, 1.f, 1.f, 1.f), ish = f4zero(), gate = isc;
                    if (p.in_scale && col_ok) { isc = ldg4(p.in_scale + c); ish = ldg4(p.in_shift + c); }
                    if (p.a_scale && col_ok && one_image) gate = ldg4(p.a_scale + (long long)b_first * p.Cin + c);
#pragma unroll
                    for (int i = 0; i < 128 / RP; ++i) {
                        const int r = rbase + RP * i;
                        float4 v = *reinterpret_cast<const float4*>(a32 + h * kPwA32Half + r * 128 + ((j ^ (r & 7)) << 4));
                        if (pro) {
                            if (!col_ok) {
                                v = f4zero();                  // the swish of a zero-filled column would not be zero
                            } else {
                                if (p.in_scale) {
                                    const float4 q = f4fma(v, isc, ish);
                                    v = make_float4(fswish(q.x), fswish(q.y), fswish(q.z), fswish(q.w));
                                }
                                if (p.a_scale) {
                                    if (!one_image) {
                                        const int m = m0 + r;
                                        gate = ldg4(p.a_scale + (long long)((m < P.M ? m : P.M - 1) / P.HW) * p.Cin + c);
                                    }
                                    v = f4mul(v, gate);
                                }
                            }
                        }
                        uint2 hi, lo;
                        split4(v, hi, lo);
                        const int off = r * 128 + (((h * 4 + (j >> 1)) ^ (r & 7)) << 4) + (j & 1) * 8;
                        *reinterpret_cast<uint2*>(hi_pl + off) = hi;
                        *reinterpret_cast<uint2*>(lo_pl + off) = lo;
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&a16_full[sa]);
                    mbar_arrive(&ld_empty[s]);
                }
            }
        }
    } else {
        // ------------------------------------------------ epilogue --------------------------------------------------
        const int etid = threadIdx.x - (2 + NC) * 32;
        const int quarter = warp & 3;                          // TMEM lane quarter this warp may read
        const int r = quarter * 32 + lane;                     // row of the tile owned by this thread
        uint32_t iu = 0, nstore = 0;
        for (int u = blockIdx.x; u < P.units; u += gridDim.x, ++iu) {
            const int mt = u / P.ntn, nt = u - mt * P.ntn;
            const int m0 = mt * 128, n0 = nt * P.BN;
            named_bar_sync(1, 128);                            // everybody is done with the previous unit's vectors
            for (int i = etid; i < P.BN; i += 128) {
                const int n = n0 + i;
                const bool ok = n < p.Cout;
                chan[i] = (ok && p.bias) ? __ldg(p.bias + n) : 0.f;
                chan[128 + i] = (ok && p.scale) ? __ldg(p.scale + n) : 1.f;
                chan[256 + i] = (ok && p.shift) ? __ldg(p.shift + n) : 0.f;
            }
            named_bar_sync(1, 128);
            const uint32_t acc = iu & 1, pacc = (iu >> 1) & 1;
            mbar_wait(&acc_full[acc], pacc);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 128;
            const int ncols = min(P.BN, p.Cout - n0);
            const int nchunks = (ncols + 31) >> 5;
            const int m = m0 + r;
            const bool row_ok = m < P.M;
            int b = 0;
            long long pix = 0;
            if (row_ok) { b = m / P.HW; pix = m - (long long)b * P.HW; }
            const float rs = (row_ok && p.row_scale) ? __ldg(p.row_scale + b) : 1.f;
            const long long ybase = (long long)b * p.y_bstride + pix * p.Cout;
            const long long rbase_ = (long long)b * p.r_bstride + pix * p.Cout;
            const long long mbase = (long long)b * p.m_bstride + pix * p.Cout;
#pragma unroll 1
            for (int cc = 0; cc < nchunks; ++cc) {
                uint32_t v32[32];
                tmem_ld32(taddr + cc * 32, v32);
                if (cc == nchunks - 1) {                       // accumulator drained: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[acc]);
                }
                if (P.tma_store) {
                    // every warp stages and stores ITS 32 rows (a 4 KB, 1024-byte aligned slice of the staging buffer) on its
                    // own: no CTA-wide barrier in the store path, the four warps drift freely (bulk groups are per thread)
                    uint8_t* buf = outst + (nstore % P.nout) * kPwOut;
                    if (lane == 0) tma_store_wait_read_upto<kPwMaxOut - 1>(P.nout - 1);   // this warp's store that last used the slot
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float4 v = make_float4(__uint_as_float(v32[q * 4]), __uint_as_float(v32[q * 4 + 1]),
                                               __uint_as_float(v32[q * 4 + 2]), __uint_as_float(v32[q * 4 + 3]));
                        v = f4add(v, *reinterpret_cast<const float4*>(chan + cc * 32 + q * 4));
                        *reinterpret_cast<float4*>(buf + r * 128 + ((q ^ (r & 7)) << 4)) = v;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&map_y, buf + quarter * 4096, n0 + cc * 32, m0 + quarter * 32);
                        tma_store_commit();
                    }
                    ++nstore;
                } else if (row_ok) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int nl = cc * 32 + q * 4;
                        const int n = n0 + nl;
                        if (n >= p.Cout) break;
                        float4 v = make_float4(__uint_as_float(v32[q * 4]), __uint_as_float(v32[q * 4 + 1]),
                                               __uint_as_float(v32[q * 4 + 2]), __uint_as_float(v32[q * 4 + 3]));
                        v = f4add(v, *reinterpret_cast<const float4*>(chan + nl));
                        if (p.z) st4(p.z + ybase + n, v);
                        v = f4fma(v, *reinterpret_cast<const float4*>(chan + 128 + nl), *reinterpret_cast<const float4*>(chan + 256 + nl));
                        if (p.act == EFFDET_ACT_RELU) {
                            v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
                        } else if (p.act == EFFDET_ACT_SIGMOID) {
                            v = make_float4(sigmoidf_(v.x), sigmoidf_(v.y), sigmoidf_(v.z), sigmoidf_(v.w));
                        } else if (p.act == EFFDET_ACT_SWISH) {
                            v = make_float4(swishf_(v.x), swishf_(v.y), swishf_(v.z), swishf_(v.w));
                        }
                        if (p.row_scale) v = f4scale(v, rs);
                        if (p.residual) v = f4add(v, ldg4(p.residual + rbase_ + n));
                        if (p.mask_src) {
                            const float4 mk = ldg4(p.mask_src + mbase + n);
                            v = make_float4(mk.x > 0.f ? v.x : 0.f, mk.y > 0.f ? v.y : 0.f, mk.z > 0.f ? v.z : 0.f,
                                            mk.w > 0.f ? v.w : 0.f);
                        }
                        st4(p.y + ybase + n, v);
                    }
                }
            }
        }
        if (P.tma_store && lane == 0) tma_store_wait_read<0>();     // shared memory stays valid until the last store has read it
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

static bool pw_enabled() {
    static const bool on = [] {
        const char* v = getenv("EFFDET_B200_PW");
        return !(v && v[0] == '0');
    }();
    return on;
}

bool pw_gemm_eligible(const effdet_conv_args* a) {
    if (!pw_enabled() || a->ksize != 1 || a->w_tc == nullptr || a->Cin % 4 || a->Cout % 4 || a->Cin < 8 || a->Cout < 8) return false;
    const long long HW = (long long)a->H * a->W;
    if (a->x_planes) return a->Cin % 8 == 0 && !a->in_scale && !a->a_scale && (long long)a->B * HW < (1ll << 31);
    if (a->x_bstride != HW * a->Cin) return false;              // x must be one dense [M, Cin] matrix for the 2-D tensor map
    if ((long long)a->B * HW >= (1ll << 31)) return false;
    return true;
}

int pw_gemm_launch(const effdet_conv_args* a, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(EFFDET_ERR_UNSUPPORTED, "conv2d(pw): cuTensorMapEncodeTiled unavailable");
    PwParams P;
    memset(&P, 0, sizeof(P));
    P.a = *a;
    P.HW = a->H * a->W;
    P.M = a->B * P.HW;
    const int kpad = conv_tc_kpad(a->Cin);
    P.KB = kpad / 64;
    P.ntn = cdiv(a->Cout, 128);
    P.BN = cdiv(cdiv(a->Cout, P.ntn), 32) * 32;
    P.ntn = cdiv(a->Cout, P.BN);
    const int mtiles = cdiv(P.M, 128);
    P.units = mtiles * P.ntn;
    P.a32_halves = a->Cin > 32 ? 2 : 1;
    P.planes = a->x_planes ? 1 : 0;
    P.tma_store = (!a->z && !a->scale && !a->row_scale && !a->residual && !a->mask_src && a->act == EFFDET_ACT_NONE &&
                   a->y_bstride == (long long)P.HW * a->Cout)
                      ? 1
                      : 0;
    const int a_stage = P.planes ? kPwA16 : P.a32_halves * kPwA32Half;
    const int b_all = P.ntn * P.KB * P.BN * 256;                        // every weight tile of the layer (hi + lo)
    P.bres = b_all <= 64 * 1024 ? 1 : 0;
    const int stage_bytes = a_stage + (P.bres ? 0 : P.BN * 256);
    const int fixed = (P.planes ? 0 : 2 * kPwA16) + kPwBarBytes + kPwChanBytes + 1024 + (P.bres ? b_all : 0);
    const int budget = 227 * 1024 - fixed;
    // shared memory split: loads in flight (ring stages) vs stores in flight (staging buffers of the TMA-store epilogue)
    int ns, nout = 0;
    if (P.tma_store) {
        nout = 3;
        ns = (budget - nout * kPwOut) / stage_bytes;
        if (ns < 2) { nout = 2; ns = (budget - nout * kPwOut) / stage_bytes; }
        if (ns > 4) {                                               // plenty of room: split the rest between both sides
            ns = 4;
            nout = (budget - ns * stage_bytes) / kPwOut;
            if (nout > kPwMaxOut) nout = kPwMaxOut;
        }
    } else {
        ns = budget / stage_bytes;
    }
    if (ns > kPwMaxStages) ns = kPwMaxStages;
    if (ns < 2) return fail(EFFDET_ERR_UNSUPPORTED, "conv2d(pw): shared memory budget");
    P.NS = ns;
    P.nout = nout;
    const size_t smem = (size_t)fixed + (size_t)ns * stage_bytes + (size_t)nout * kPwOut;

    CUtensorMap map_a, map_b, map_y;
    if (P.planes) {
        const cuuint64_t gdim[3] = {(cuuint64_t)a->Cin, (cuuint64_t)P.M, 2};
        const cuuint64_t gstr[2] = {(cuuint64_t)a->Cin * 2, (cuuint64_t)P.M * a->Cin * 2};
        const cuuint32_t box[3] = {64, 128, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(a->x_planes), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(EFFDET_ERR_LAUNCH, "conv2d(pw): tensor map of the x planes failed (%d)", (int)r);
    } else {
        const cuuint64_t gdim[2] = {(cuuint64_t)a->Cin, (cuuint64_t)P.M};
        const cuuint64_t gstr[1] = {(cuuint64_t)a->Cin * 4};
        const cuuint32_t box[2] = {32, 128};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a->x), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(EFFDET_ERR_LAUNCH, "conv2d(pw): tensor map of x failed (%d)", (int)r);
    }
    {
        const cuuint64_t gdim[3] = {(cuuint64_t)kpad, (cuuint64_t)a->Cout, 2};
        const cuuint64_t gstr[2] = {(cuuint64_t)kpad * 2, (cuuint64_t)a->Cout * kpad * 2};
        const cuuint32_t box[3] = {64, (cuuint32_t)P.BN, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(a->w_tc), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(EFFDET_ERR_LAUNCH, "conv2d(pw): tensor map of the weights failed (%d)", (int)r);
    }
    if (P.tma_store) {
        const cuuint64_t gdim[2] = {(cuuint64_t)a->Cout, (cuuint64_t)P.M};
        const cuuint64_t gstr[1] = {(cuuint64_t)a->Cout * 4};
        const cuuint32_t box[2] = {32, 32};              // one epilogue warp's rows
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&map_y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a->y, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(EFFDET_ERR_LAUNCH, "conv2d(pw): tensor map of y failed (%d)", (int)r);
    } else {
        map_y = map_a;
    }
    const int grid = P.units < 148 ? P.units : 148;
    cudaError_t e;
    if (a->in_scale) {          // a swish per staged element: 16 converter warps
        e = cudaFuncSetAttribute(pw_gemm_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return fail(EFFDET_ERR_LAUNCH, "conv2d(pw): smem opt-in: %s", cudaGetErrorString(e));
        pw_gemm_kernel<16><<<grid, (16 + 6) * 32, smem, st>>>(map_a, map_b, map_y, P);
    } else {
        e = cudaFuncSetAttribute(pw_gemm_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return fail(EFFDET_ERR_LAUNCH, "conv2d(pw): smem opt-in: %s", cudaGetErrorString(e));
        pw_gemm_kernel<8><<<grid, (8 + 6) * 32, smem, st>>>(map_a, map_b, map_y, P);
    }
    return launch_status("pw_gemm_kernel");
}

}  // namespace effdet
