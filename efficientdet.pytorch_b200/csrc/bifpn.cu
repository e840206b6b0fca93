// BiFPN fast-normalised weighted fusion with its resampling fused in (nearest x2 upsample of the
// coarser map, or 2x2/2 max-pool of the finer map), forward and backward.
// Reference: BiFPNModule.forward, models/bifpn.py:172-203:
//   w = relu(w_raw); w /= (sum_over_rows w) + eps                       (:177-180)
//   node = (w0*a + w1*resample(b) [+ w2*c]) / (w0 + w1 [+ w2] + eps)    (:189-190,195-196,200-201)
// i.e. the weights are normalised twice; both normalisations are differentiated here.
// One read of each input, one write of the fused map: 4*B*C*(s^2 + s^2/4 + s^2) bytes for an
// up-node, 4*B*C*(s^2 + 4 s^2 + [s^2] + s^2) for a pool-node.
#include "common.cuh"

#include <cuda_bf16.h>

namespace effdet {

struct FuseCoef { float n0, n1, n2, D; };

__device__ __forceinline__ FuseCoef fuse_coef(const float* __restrict__ w, int stride, int nin, float eps) {
    const float r0 = fmaxf(__ldg(w), 0.f), r1 = fmaxf(__ldg(w + stride), 0.f);
    const float r2 = nin == 3 ? fmaxf(__ldg(w + 2 * stride), 0.f) : 0.f;
    const float E = (nin == 3 ? (r0 + r1) + r2 : r0 + r1) + eps;
    FuseCoef c;
    c.n0 = r0 / E; c.n1 = r1 / E; c.n2 = nin == 3 ? r2 / E : 0.f;
    c.D = (nin == 3 ? (c.n0 + c.n1) + c.n2 : c.n0 + c.n1) + eps;
    return c;
}

// first maximum in row-major window order (what F.max_pool2d's backward routes to)
__device__ __forceinline__ void max4(const float4 v[4], float4& m, int4& arg) {
    m = v[0];
    arg = make_int4(0, 0, 0, 0);
#pragma unroll
    for (int i = 1; i < 4; ++i) {
        if (v[i].x > m.x) { m.x = v[i].x; arg.x = i; }
        if (v[i].y > m.y) { m.y = v[i].y; arg.y = i; }
        if (v[i].z > m.z) { m.z = v[i].z; arg.z = i; }
        if (v[i].w > m.w) { m.w = v[i].w; arg.w = i; }
    }
}

__global__ void __launch_bounds__(256) fuse_fwd_kernel(const effdet_fuse_args p) {
    const int cvecs = p.C / 4;
    const long long total = (long long)p.B * p.H * p.W * cvecs;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int nin = p.c ? 3 : 2;
    const FuseCoef k = fuse_coef(p.w, p.w_stride, nin, p.eps);
    const int cv = (int)(idx % cvecs);
    long long r = idx / cvecs;
    const int x = (int)(r % p.W);
    r /= p.W;
    const int y = (int)(r % p.H);
    const int b = (int)(r / p.H);
    const float4 a = ldg4(p.a + idx * 4);
    float4 bb;
    if (p.mode == EFFDET_FUSE_UP) {
        const int Hb = p.H / 2, Wb = p.W / 2;
        bb = ldg4(p.b + (((long long)b * Hb + y / 2) * Wb + x / 2) * p.C + cv * 4);
    } else {
        const int Wb = p.W * 2;
        const float* q = p.b + (((long long)b * p.H * 2 + 2 * y) * Wb + 2 * x) * p.C + cv * 4;
        float4 v[4] = {ldg4(q), ldg4(q + p.C), ldg4(q + (long long)Wb * p.C), ldg4(q + (long long)Wb * p.C + p.C)};
        int4 arg;
        max4(v, bb, arg);
    }
    float4 s = make_float4(k.n0 * a.x + k.n1 * bb.x, k.n0 * a.y + k.n1 * bb.y, k.n0 * a.z + k.n1 * bb.z,
                           k.n0 * a.w + k.n1 * bb.w);
    if (nin == 3) {
        const float4 c = ldg4(p.c + idx * 4);
        s = make_float4(s.x + k.n2 * c.x, s.y + k.n2 * c.y, s.z + k.n2 * c.z, s.w + k.n2 * c.w);
    }
    const float4 o = make_float4(s.x / k.D, s.y / k.D, s.z / k.D, s.w / k.D);
    if (p.out) st4(p.out + idx * 4, o);
    if (p.out_planes) {          // the node conv's operand format: bf16 hi/lo planes [2][B*H*W][pitch], o ~= hi + lo
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
        const __nv_bfloat162 l0 = __floats2bfloat162_rn(o.x - __low2float(h0), o.y - __high2float(h0));
        const __nv_bfloat162 l1 = __floats2bfloat162_rn(o.z - __low2float(h1), o.w - __high2float(h1));
        const int pitch = (p.C + 7) / 8 * 8;
        __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(p.out_planes);
        const long long e = (idx / cvecs) * pitch + cv * 4;
        *reinterpret_cast<uint2*>(pl + e) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
        *reinterpret_cast<uint2*>(pl + (long long)p.B * p.H * p.W * pitch + e) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
    }
}

__device__ __forceinline__ void put4(float* dst, float4 v, int acc) {
    if (acc) v = f4add(v, *reinterpret_cast<const float4*>(dst));
    st4(dst, v);
}
__device__ __forceinline__ float hsum4(float4 v) { return (v.x + v.y) + (v.z + v.w); }

__device__ __forceinline__ void block_add3(float s0, float s1, float s2, float* scratch) {
    __shared__ float red[3][8];
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; red[2][warp] = s2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i];
        atomicAdd(scratch + threadIdx.x, s);
    }
}

// UP node: threads tile the COARSE grid of b; each owns the 2x2 block of fine pixels of a/out.
__global__ void __launch_bounds__(256) fuse_bwd_up_kernel(const effdet_fuse_bwd_args p) {
    const int cvecs = p.C / 4;
    const int Hb = p.H / 2, Wb = p.W / 2;
    const long long total = (long long)p.B * Hb * Wb * cvecs;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (idx < total) {
        const int nin = p.c ? 3 : 2;
        const FuseCoef k = fuse_coef(p.w, p.w_stride, nin, p.eps);
        const float c0 = k.n0 / k.D, c1 = k.n1 / k.D, c2 = k.n2 / k.D;
        const int cv = (int)(idx % cvecs);
        long long r = idx / cvecs;
        const int xb = (int)(r % Wb);
        r /= Wb;
        const int yb = (int)(r % Hb);
        const int b = (int)(r / Hb);
        const float4 bv = ldg4(p.b + idx * 4);
        float4 gsum = f4zero();
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const long long o = (((long long)b * p.H + 2 * yb + dy) * p.W + 2 * xb + dx) * p.C + cv * 4;
                const float4 g = ldg4(p.dout + o);
                gsum = f4add(gsum, g);
                s0 += hsum4(f4mul(g, ldg4(p.a + o)));
                put4(p.da + o, f4scale(g, c0), p.acc_a);
                if (nin == 3) {
                    s2 += hsum4(f4mul(g, ldg4(p.c + o)));
                    put4(p.dc + o, f4scale(g, c2), p.acc_c);
                }
            }
        s1 = hsum4(f4mul(gsum, bv));
        put4(p.db + idx * 4, f4scale(gsum, c1), p.acc_b);
    }
    block_add3(s0, s1, s2, p.scratch);
}

// POOL node: threads tile the grid of a/out; each owns the 2x2 window of the finer map b.
__global__ void __launch_bounds__(256) fuse_bwd_pool_kernel(const effdet_fuse_bwd_args p) {
    const int cvecs = p.C / 4;
    const long long total = (long long)p.B * p.H * p.W * cvecs;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (idx < total) {
        const int nin = p.c ? 3 : 2;
        const FuseCoef k = fuse_coef(p.w, p.w_stride, nin, p.eps);
        const float c0 = k.n0 / k.D, c1 = k.n1 / k.D, c2 = k.n2 / k.D;
        const int cv = (int)(idx % cvecs);
        long long r = idx / cvecs;
        const int x = (int)(r % p.W);
        r /= p.W;
        const int y = (int)(r % p.H);
        const int b = (int)(r / p.H);
        const float4 g = ldg4(p.dout + idx * 4);
        s0 = hsum4(f4mul(g, ldg4(p.a + idx * 4)));
        put4(p.da + idx * 4, f4scale(g, c0), p.acc_a);
        if (nin == 3) {
            s2 = hsum4(f4mul(g, ldg4(p.c + idx * 4)));
            put4(p.dc + idx * 4, f4scale(g, c2), p.acc_c);
        }
        const int Wb = p.W * 2;
        const long long o00 = (((long long)b * p.H * 2 + 2 * y) * Wb + 2 * x) * p.C + cv * 4;
        const long long offs[4] = {o00, o00 + p.C, o00 + (long long)Wb * p.C, o00 + (long long)Wb * p.C + p.C};
        float4 v[4] = {ldg4(p.b + offs[0]), ldg4(p.b + offs[1]), ldg4(p.b + offs[2]), ldg4(p.b + offs[3])};
        float4 m;
        int4 arg;
        max4(v, m, arg);
        s1 = hsum4(f4mul(g, m));
        const float4 gb = f4scale(g, c1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 d = make_float4(arg.x == i ? gb.x : 0.f, arg.y == i ? gb.y : 0.f, arg.z == i ? gb.z : 0.f,
                                         arg.w == i ? gb.w : 0.f);
            put4(p.db + offs[i], d, p.acc_b);
        }
    }
    block_add3(s0, s1, s2, p.scratch);
}

// scalar chain rule through both normalisations; T_j = sum(dout * in_j) arrives in scratch
__global__ void fuse_bwd_weights_kernel(const float* __restrict__ w, int stride, int nin, float eps,
                                        const float* __restrict__ scratch, float* __restrict__ dw) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float wr[3], r[3], n[3], T[3], dn[3];
    float E = 0.f;
    for (int j = 0; j < nin; ++j) { wr[j] = w[j * stride]; r[j] = fmaxf(wr[j], 0.f); E += r[j]; T[j] = scratch[j]; }
    E += eps;
    float D = 0.f;
    for (int j = 0; j < nin; ++j) { n[j] = r[j] / E; D += n[j]; }
    D += eps;
    float nT = 0.f;
    for (int j = 0; j < nin; ++j) nT += n[j] * T[j];
    float dnr = 0.f;
    for (int j = 0; j < nin; ++j) { dn[j] = T[j] / D - nT / (D * D); dnr += dn[j] * r[j]; }
    for (int j = 0; j < nin; ++j) {
        const float dr = dn[j] / E - dnr / (E * E);
        if (wr[j] > 0.f) dw[j * stride] += dr;
    }
}

}  // namespace effdet

using namespace effdet;

extern "C" int effdet_bifpn_fuse_fwd(const effdet_fuse_args* a, int device, effdet_stream_t stream) {
    EFFDET_REQUIRE(a && a->a && a->b && a->w && (a->out || a->out_planes), "bifpn_fuse_fwd: null tensor");
    EFFDET_REQUIRE(aligned16(a->out_planes), "bifpn_fuse_fwd: alignment");
    EFFDET_REQUIRE(a->C % 4 == 0 && a->B > 0 && a->H > 0 && a->W > 0, "bifpn_fuse_fwd: bad shape");
    EFFDET_REQUIRE(a->mode == EFFDET_FUSE_POOL || (a->H % 2 == 0 && a->W % 2 == 0), "bifpn_fuse_fwd: up-node needs even H,W");
    EFFDET_REQUIRE(aligned16(a->a) && aligned16(a->b) && aligned16(a->c) && aligned16(a->out), "bifpn_fuse_fwd: alignment");
    EFFDET_DEVICE(device);
    const long long total = (long long)a->B * a->H * a->W * (a->C / 4);
    fuse_fwd_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(*a);
    return launch_status("fuse_fwd_kernel");
}

extern "C" int effdet_bifpn_fuse_bwd(const effdet_fuse_bwd_args* a, int device, effdet_stream_t stream) {
    EFFDET_REQUIRE(a && a->dout && a->a && a->b && a->w && a->da && a->db && a->dw && a->scratch, "bifpn_fuse_bwd: null tensor");
    EFFDET_REQUIRE((a->c == nullptr) == (a->dc == nullptr), "bifpn_fuse_bwd: c and dc must come together");
    EFFDET_REQUIRE(a->C % 4 == 0 && a->B > 0 && a->H > 0 && a->W > 0, "bifpn_fuse_bwd: bad shape");
    EFFDET_REQUIRE(a->mode == EFFDET_FUSE_POOL || (a->H % 2 == 0 && a->W % 2 == 0), "bifpn_fuse_bwd: up-node needs even H,W");
    EFFDET_REQUIRE(aligned16(a->dout) && aligned16(a->a) && aligned16(a->b) && aligned16(a->c) && aligned16(a->da) &&
                       aligned16(a->db) && aligned16(a->dc), "bifpn_fuse_bwd: alignment");
    EFFDET_DEVICE(device);
    cudaStream_t st = (cudaStream_t)stream;
    if (a->mode == EFFDET_FUSE_UP) {
        const long long total = (long long)a->B * (a->H / 2) * (a->W / 2) * (a->C / 4);
        fuse_bwd_up_kernel<<<cdiv(total, 256), 256, 0, st>>>(*a);
    } else {
        const long long total = (long long)a->B * a->H * a->W * (a->C / 4);
        fuse_bwd_pool_kernel<<<cdiv(total, 256), 256, 0, st>>>(*a);
    }
    int s = launch_status("fuse_bwd_kernel");
    if (s) return s;
    fuse_bwd_weights_kernel<<<1, 32, 0, st>>>(a->w, a->w_stride, a->c ? 3 : 2, a->eps, a->scratch, a->dw);
    return launch_status("fuse_bwd_weights_kernel");
}
