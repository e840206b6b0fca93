// Shared helpers for the effdet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/effdet_b200.h"

namespace effdet {

// ---- error plumbing: no exceptions cross the C ABI; the message is thread-local -------------
char* err_buf();
int fail(int code, const char* fmt, ...);
void count_launch(int n = 1);

// cudaPeekAtLastError after a launch; never synchronises the device.
inline int launch_status(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(EFFDET_ERR_LAUNCH, "%s: %s", what, cudaGetErrorString(e));
    }
    count_launch();
    return EFFDET_OK;
}

// Makes `device` current for the duration of an entry point and restores the caller's current device on return
// (a call on a tensor of another GPU must not leak a device switch into the calling thread).
struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    int set(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev == device) return EFFDET_OK;
        cudaError_t e = cudaSetDevice(device);
        if (e != cudaSuccess) return fail(EFFDET_ERR_DEVICE, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
        changed = prev >= 0;
        return EFFDET_OK;
    }
    ~DeviceGuard() {
        if (changed) cudaSetDevice(prev);
    }
};

#define EFFDET_REQUIRE(cond, ...)                                            \
    do {                                                                     \
        if (!(cond)) return effdet::fail(EFFDET_ERR_ARG, __VA_ARGS__);       \
    } while (0)
#define EFFDET_DEVICE(dev)                                                   \
    effdet::DeviceGuard _effdet_device_guard;                                \
    do {                                                                     \
        int _s = _effdet_device_guard.set(dev);                              \
        if (_s) return _s;                                                   \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- device helpers ----------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float swishf_(float x) { return x * sigmoidf_(x); }
// d/dx [x*sigmoid(x)] = s*(1 + x*(1-s))      (reference: models/utils.py:38-42)
__device__ __forceinline__ float swish_gradf_(float x) {
    float s = sigmoidf_(x);
    return s * (1.0f + x * (1.0f - s));
}

// Fast sigmoid / swish for the HBM-bound kernels that evaluate an activation per element they move: ex2.approx with
// the log2(e) scaling + rcp.approx, ~1e-6 relative (|x| * 2^-24 from the exponent scaling + 3 ulp), 4 instructions
// instead of ~25 (with branches) for expf + IEEE division.
__device__ __forceinline__ float fsigmoid(const float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));      // no slow path: 2 MUFU + 2 FP ops, branch-free
    return r;
}
__device__ __forceinline__ float fswish(const float x) { return x * fsigmoid(x); }
__device__ __forceinline__ float fswish_grad(const float x) {
    const float s = fsigmoid(x);
    return s * (1.0f + x * (1.0f - s));
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4fma(float4 a, float4 b, float4 c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 f4scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// "Row-packed" thread mapping for NHWC element-wise kernels with per-channel reductions.
// A 256-thread block covers `rows` consecutive rows x `cvb` float4 channel-vectors at a time
// (rows*cvb <= 256), so global accesses are fully contiguous and every thread keeps a fixed
// channel vector across iterations (register accumulation of per-channel sums).
struct RowPack {
    int cvb;     // channel vectors per block (<= blockDim.x)
    int rows;    // rows processed per iteration
    int tr, tc;  // this thread's row lane and channel-vector lane
    int cv;      // global channel-vector index
    bool active;
};
__device__ __forceinline__ RowPack rowpack(int cvecs, int chunk) {
    RowPack r;
    r.cvb = cvecs < (int)blockDim.x ? cvecs : (int)blockDim.x;
    r.rows = (int)blockDim.x / r.cvb;
    r.tr = (int)threadIdx.x / r.cvb;
    r.tc = (int)threadIdx.x - r.tr * r.cvb;
    r.cv = chunk * r.cvb + r.tc;
    r.active = (r.tr < r.rows) && (r.cv < cvecs);
    return r;
}
static inline int rowpack_chunks(int cvecs, int threads = 256) { return cvecs <= threads ? 1 : cdiv(cvecs, threads); }
static inline int rowpack_rows(int cvecs, int threads = 256) { return cvecs >= threads ? 1 : threads / cvecs; }

}  // namespace effdet
