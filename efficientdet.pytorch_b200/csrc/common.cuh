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

inline int use_device(int device) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(EFFDET_ERR_DEVICE, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    return EFFDET_OK;
}

#define EFFDET_REQUIRE(cond, ...)                                            \
    do {                                                                     \
        if (!(cond)) return effdet::fail(EFFDET_ERR_ARG, __VA_ARGS__);       \
    } while (0)
#define EFFDET_DEVICE(dev)                                                   \
    do {                                                                     \
        int _s = effdet::use_device(dev);                                    \
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
