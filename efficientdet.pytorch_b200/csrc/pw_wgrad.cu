// Weight gradient of the 1x1 convolutions of the MBConv block (models/efficientnet.py:85 expand, :96 project; reached
// through autograd's cuDNN bwd-filter in the reference), straight from the fp32 tensors:
//     dW[n][c] += sum_m dy[m][n] * xt[m][c],      xt = swish(x*in_scale+in_shift) * a_scale[image]   (both optional)
// The TMA-fed kernel in conv_tc.cu wants both operands pre-split into bf16 hi/lo planes; for these layers that cost a
// split pass over x (with the BN+swish+SE-gate prologue) and one over dy -- each a full read + write of an expanded
// activation -- before the GEMM read them a third time.  Here sixteen converter warps read the fp32 rows with 128-bit
// coalesced loads, apply the prologue in registers, split to bf16 hi/lo and write the operand tiles in the
// SWIZZLE_128B "MN-major" layout (row = pixel, 128 bytes = 64 channels) the tensor-map loads would have produced; one
// thread issues the bf16x3 tcgen05 MMAs (GEMM-K = pixels, M = output channels, N = input channels, accumulator in
// TMEM); the CTA then adds its [128 x N] partial to dW with vector reductions.  Every operand byte is read from HBM
// once: algorithmic bytes = 4*M*(Cin + Cout) (dy pre-split by the depthwise backward kernel: same 4 bytes / element).
#include "tc_ptx.cuh"

#include <cstdlib>
#include <cstring>

namespace effdet {

constexpr int kWgConv = 12;                     // converter warps (12 * 32 threads * ~150 registers: room for the double-buffered loads)
constexpr int kWgCT = kWgConv * 32;             // converter threads
constexpr int kWgThreads = kWgCT + 32;          // + the MMA warp
constexpr int kWgUnits = 4;                     // 8-channel units per converter thread and stage: K * (octs of x + dy) <= 1536

struct PwWgParams {
    const float* x;
    const float* dy;
    const uint16_t* dy_planes;      // [2][M][Cout] bf16 or NULL
    const float* in_scale;
    const float* in_shift;
    const float* a_scale;
    float* dw;
    int M, HW, Cin, Cout;
    int ntn, NX, TM;                // input-channel tiles, their width (multiple of 16, <= 128), output channels per tile (<= 128)
    int K;                          // pixels per stage: 64, or 128 for narrow tiles (more bytes in flight per thread)
    int nchunks, cps;               // K-pixel chunks, chunks per split
    int NS, stage_bytes;            // ring depth, bytes per stage
    int a_plane, b_plane;           // bytes of one dy / x plane of a stage (1 or 2 groups of 64 channels)
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1) pw_wgrad_kernel(const __grid_constant__ PwWgParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ctl = smem + P.NS * P.stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctl);
    uint64_t* empty_bar = full_bar + 4;
    uint64_t* accum_bar = empty_bar + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
    float* chan = reinterpret_cast<float*>(ctl + 128);          // in_scale | in_shift of this tile's channels

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tn = blockIdx.x % P.ntn, tm = blockIdx.x / P.ntn;
    const int c0 = tn * P.NX, n0 = tm * P.TM;
    const int ncur = min(P.NX, P.Cin - c0);                     // real input channels of this tile (multiple of 8)
    const int nmma = (ncur + 15) & ~15;
    const int mcur = min(P.TM, P.Cout - n0);                    // real output channels
    const int ch_begin = blockIdx.y * P.cps;
    const int KT = min(P.nchunks, ch_begin + P.cps) - ch_begin;
    const int group = P.K * 128;                                // bytes of one 64-channel group of one plane

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.NS; ++s) {
            mbar_init(&full_bar[s], kWgConv);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == kWgConv) tmem_alloc<128>(tmem_slot);
    if (P.in_scale)
        for (int i = threadIdx.x; i < ncur; i += kWgThreads) {
            chan[i] = __ldg(P.in_scale + c0 + i);
            chan[128 + i] = __ldg(P.in_shift + c0 + i);
        }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kWgConv) {
        // this thread's units: the same (pixel row, channel octet) of every stage.  kind 0 = x, 1 = dy, 2 = none
        const int nxo = ncur >> 3, nyo = mcur >> 3;
        const int UX = P.K * nxo, U = UX + P.K * nyo;
        const bool planes = P.dy_planes != nullptr;
        int kind[kWgUnits], prow[kWgUnits], ooff[kWgUnits];
        const char* src[kWgUnits];
        uint32_t dst[kWgUnits];
        size_t step_x = (size_t)P.K * P.Cin * 4, step_y = (size_t)P.K * P.Cout * (planes ? 2 : 4);
#pragma unroll
        for (int i = 0; i < kWgUnits; ++i) {
            const int u = threadIdx.x + i * kWgCT;
            kind[i] = 2; prow[i] = 0; ooff[i] = 0; src[i] = nullptr; dst[i] = 0;
            if (u < UX) {
                const int p = u / nxo, o = u - p * nxo;
                kind[i] = 0; prow[i] = p; ooff[i] = o * 8;
                src[i] = reinterpret_cast<const char*>(P.x + ((size_t)ch_begin * P.K + p) * P.Cin + c0 + o * 8);
                dst[i] = 2 * P.a_plane + (o >> 3) * group + p * 128 + (((o & 7) ^ (p & 7)) << 4);
            } else if (u < U) {
                const int v = u - UX;
                const int p = v / nyo, o = v - p * nyo;
                kind[i] = 1; prow[i] = p;
                const size_t e = ((size_t)ch_begin * P.K + p) * P.Cout + n0 + o * 8;
                src[i] = planes ? reinterpret_cast<const char*>(P.dy_planes + e) : reinterpret_cast<const char*>(P.dy + e);
                dst[i] = (o >> 3) * group + p * 128 + (((o & 7) ^ (p & 7)) << 4);
            }
        }
        const size_t plane_bytes = (size_t)P.M * P.Cout * 2;    // hi -> lo plane of dy
        float4 va[2][kWgUnits], vb[2][kWgUnits];
        auto load = [&](int kt, float4 (&a)[kWgUnits], float4 (&b)[kWgUnits]) {
            const int m0 = (ch_begin + kt) * P.K;
#pragma unroll
            for (int i = 0; i < kWgUnits; ++i) {
                a[i] = b[i] = f4zero();
                if (kind[i] != 2 && m0 + prow[i] < P.M) {
                    const char* s = src[i] + (size_t)kt * (kind[i] == 0 ? step_x : step_y);
                    a[i] = __ldg(reinterpret_cast<const float4*>(s));
                    b[i] = __ldg(reinterpret_cast<const float4*>(s + ((kind[i] == 1 && planes) ? plane_bytes : 16)));
                }
            }
        };
        auto convert = [&](int kt, float4 (&a)[kWgUnits], float4 (&b)[kWgUnits]) {
            const int s = kt % P.NS;
            const int m0 = (ch_begin + kt) * P.K;
            if (lane == 0) mbar_wait(&empty_bar[s], ((kt / P.NS) & 1) ^ 1);   // the MMAs that last read the slot are done
            __syncwarp();
            uint8_t* st = smem + (size_t)s * P.stage_bytes;
#pragma unroll
            for (int i = 0; i < kWgUnits; ++i) {
                if (kind[i] == 2) continue;
                uint4 hi, lo;
                if (kind[i] == 0) {
                    float4 xa = a[i], xb = b[i];
                    if (m0 + prow[i] < P.M) {
                        if (P.in_scale) {
                            const float* cs = chan + ooff[i];
                            xa = f4fma(xa, *reinterpret_cast<const float4*>(cs), *reinterpret_cast<const float4*>(cs + 128));
                            xb = f4fma(xb, *reinterpret_cast<const float4*>(cs + 4), *reinterpret_cast<const float4*>(cs + 132));
                            xa = make_float4(fswish(xa.x), fswish(xa.y), fswish(xa.z), fswish(xa.w));
                            xb = make_float4(fswish(xb.x), fswish(xb.y), fswish(xb.z), fswish(xb.w));
                        }
                        if (P.a_scale) {
                            const float* g = P.a_scale + (size_t)((m0 + prow[i]) / P.HW) * P.Cin + c0 + ooff[i];
                            xa = f4mul(xa, ldg4(g));
                            xb = f4mul(xb, ldg4(g + 4));
                        }
                    }
                    split8(xa, xb, hi, lo);
                    *reinterpret_cast<uint4*>(st + dst[i]) = hi;
                    *reinterpret_cast<uint4*>(st + dst[i] + P.b_plane) = lo;
                } else {
                    if (planes) {
                        hi = *reinterpret_cast<const uint4*>(&a[i]);
                        lo = *reinterpret_cast<const uint4*>(&b[i]);
                    } else {
                        split8(a[i], b[i], hi, lo);
                    }
                    *reinterpret_cast<uint4*>(st + dst[i]) = hi;
                    *reinterpret_cast<uint4*>(st + dst[i] + P.a_plane) = lo;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[s]);
        };
        // software pipeline: the loads of chunk kt+1 are in flight while chunk kt is converted
        load(0, va[0], vb[0]);
        for (int kt = 0; kt < KT; kt += 2) {
            if (kt + 1 < KT) load(kt + 1, va[1], vb[1]);
            convert(kt, va[0], vb[0]);
            if (kt + 1 < KT) {
                if (kt + 2 < KT) load(kt + 2, va[0], vb[0]);
                convert(kt + 1, va[1], vb[1]);
            }
        }
        // epilogue: TMEM lane = output channel, column = input channel; warp w drains lane quarter w % 4, column groups w / 4, w / 4 + 3
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int quarter = warp & 3;
#pragma unroll 1
        for (int col = (warp >> 2) * 32; col < nmma; col += (kWgConv / 4) * 32) {    // warp-uniform
            uint32_t acc[32];
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + col, acc);
            const int r = quarter * 32 + lane;
            if (r < mcur) {
                float* row = P.dw + (size_t)(n0 + r) * P.Cin + c0 + col;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (col + q * 4 < ncur)
                        red_add_v4(row + q * 4, __uint_as_float(acc[q * 4]), __uint_as_float(acc[q * 4 + 1]),
                                   __uint_as_float(acc[q * 4 + 2]), __uint_as_float(acc[q * 4 + 3]));
            }
        }
        tc_fence_before();
    } else if (lane == 0) {
        const uint32_t idesc = umma_idesc(128, nmma, 1, 1);
        // dy tiles narrower than 65 channels keep ONE group in shared memory: the second half of the M = 128 operand
        // aliases the first (LBO 0); its accumulator rows are never read
        const uint32_t LBO_A = P.a_plane > group ? group : 0, LBO_B = group, SBO = 1024;
        const int ksteps = P.K / 16;
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % P.NS;
            const uint32_t ph = (kt / P.NS) & 1;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(smem + (size_t)s * P.stage_bytes);
            const uint32_t a_lo = a_hi + P.a_plane;
            const uint32_t b_hi = a_hi + 2 * P.a_plane;
            const uint32_t b_lo = b_hi + P.b_plane;
            for (int k = 0; k < ksteps; ++k) {
                const uint32_t ko = k * 2 * SBO;                 // 16 pixels = two 8-row groups
                const uint64_t dah = umma_desc(a_hi + ko, LBO_A, SBO), dal = umma_desc(a_lo + ko, LBO_A, SBO);
                const uint64_t dbh = umma_desc(b_hi + ko, LBO_B, SBO), dbl = umma_desc(b_lo + ko, LBO_B, SBO);
                umma_bf16(tmem_base, dal, dbh, idesc, (kt | k) != 0);
                umma_bf16(tmem_base, dah, dbl, idesc, 1);
                umma_bf16(tmem_base, dah, dbh, idesc, 1);
            }
            umma_commit(&empty_bar[s]);
        }
        umma_commit(accum_bar);
    }
    __syncthreads();
    if (warp == kWgConv) {
        tc_fence_after();
        tmem_dealloc<128>(tmem_base);
    }
}

static bool pw_wgrad_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("EFFDET_B200_PWWG");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on;
}

bool pw_wgrad_eligible(const effdet_wgrad_args* a) {
    if (!pw_wgrad_enabled() || a->ksize != 1 || a->precision != 1 || a->dbias || a->x_planes || !a->x) return false;
    if (a->Cin % 8 || a->Cout % 8 || a->Cin < 8 || a->Cout < 8) return false;
    const long long HW = (long long)a->H * a->W;
    if (a->x_bstride != HW * a->Cin) return false;
    if (!a->dy_planes && (!a->dy || a->dy_bstride != HW * a->Cout)) return false;
    return (long long)a->B * HW < (1ll << 31) - 64;
}

int pw_wgrad_launch(const effdet_wgrad_args* a, cudaStream_t st) {
    PwWgParams P;
    memset(&P, 0, sizeof(P));
    P.x = a->x;
    P.dy = a->dy;
    P.dy_planes = reinterpret_cast<const uint16_t*>(a->dy_planes);
    P.in_scale = a->in_scale;
    P.in_shift = a->in_shift;
    P.a_scale = a->a_scale;
    P.dw = a->dw;
    P.HW = a->H * a->W;
    P.M = a->B * P.HW;
    P.Cin = a->Cin;
    P.Cout = a->Cout;
    // balanced tiles, (x, dy) channels per tile <= (128, 64) or (64, 128) -- whichever re-reads fewer bytes: every
    // output-channel tile reads x again, every input-channel tile reads dy again
    int best = -1;
    long long best_cost = 0;
    for (int opt = 0; opt < 2; ++opt) {
        const int nx_max = opt ? 64 : 128, tm_max = opt ? 128 : 64;
        const int ntn = cdiv(a->Cin, nx_max), ntm = cdiv(a->Cout, tm_max);
        const long long cost = (long long)ntm * a->Cin + (long long)ntn * a->Cout;
        if (best < 0 || cost < best_cost) { best = opt; best_cost = cost; }
    }
    const int nx_max = best ? 64 : 128, tm_max = best ? 128 : 64;
    P.ntn = cdiv(a->Cin, nx_max);
    P.NX = cdiv(cdiv(a->Cin, P.ntn), 16) * 16;
    P.ntn = cdiv(a->Cin, P.NX);
    int ntm = cdiv(a->Cout, tm_max);
    P.TM = cdiv(cdiv(a->Cout, ntm), 8) * 8;
    ntm = cdiv(a->Cout, P.TM);
    const int tiles = P.ntn * ntm;
    const int octs = (P.NX < a->Cin ? P.NX : a->Cin) / 8 + P.TM / 8;
    P.K = octs <= 12 ? 128 : 64;                                 // K * octs <= 1536 units = 4 per converter thread
    P.nchunks = cdiv(P.M, P.K);
    int splits = 148 / tiles;
    if (splits < 1) splits = 1;
    if (splits > P.nchunks) splits = P.nchunks;
    P.cps = cdiv(P.nchunks, splits);
    splits = cdiv(P.nchunks, P.cps);
    const int group = P.K * 128;
    P.a_plane = cdiv(P.TM, 64) * group;
    P.b_plane = cdiv(P.NX, 64) * group;
    P.stage_bytes = 2 * P.a_plane + 2 * P.b_plane;
    P.NS = (200 * 1024) / P.stage_bytes;
    if (P.NS > 4) P.NS = 4;
    const size_t smem = (size_t)P.NS * P.stage_bytes + 128 + 2 * 128 * sizeof(float) + 1024;
    cudaError_t e = cudaFuncSetAttribute(pw_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(EFFDET_ERR_LAUNCH, "wgrad(pw): smem opt-in: %s", cudaGetErrorString(e));
    pw_wgrad_kernel<<<dim3(tiles, splits), kWgThreads, smem, st>>>(P);
    return launch_status("pw_wgrad_kernel");
}

}  // namespace effdet
