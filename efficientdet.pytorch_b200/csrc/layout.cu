// Layout plumbing at the module boundary: callers hold logical NCHW tensors, every kernel in
// this library is NHWC.  Tiled 32x32 shared-memory transposes, coalesced on both sides.
#include "common.cuh"

namespace effdet {

// in: [B][R][S] -> out: [B][S][R]
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int S) {
    __shared__ float tile[32][33];
    const long long boff = (long long)blockIdx.z * R * S;
    const int s0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, s = s0 + tx;
        if (r < R && s < S) tile[i][tx] = __ldg(in + boff + (long long)r * S + s);
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int s = s0 + i, r = r0 + tx;
        if (r < R && s < S) out[boff + (long long)s * R + r] = tile[tx][i];
    }
}

}  // namespace effdet

using namespace effdet;

static int transpose_launch(const char* who, const float* x, float* y, int B, int R, int S, int device, effdet_stream_t stream) {
    EFFDET_REQUIRE(x && y && B > 0 && R > 0 && S > 0 && B <= 65535, "%s: bad arguments", who);
    EFFDET_DEVICE(device);
    EFFDET_REQUIRE(cdiv(R, 32) <= 65535, "%s: dimension too large", who);
    transpose_kernel<<<dim3(cdiv(S, 32), cdiv(R, 32), B), 256, 0, (cudaStream_t)stream>>>(x, y, R, S);
    return launch_status("transpose_kernel");
}

extern "C" int effdet_nchw_to_nhwc(const float* x, float* y, int B, int C, int H, int W, int device, effdet_stream_t stream) {
    return transpose_launch("nchw_to_nhwc", x, y, B, C, H * W, device, stream);
}
extern "C" int effdet_nhwc_to_nchw(const float* x, float* y, int B, int C, int H, int W, int device, effdet_stream_t stream) {
    return transpose_launch("nhwc_to_nchw", x, y, B, H * W, C, device, stream);
}
