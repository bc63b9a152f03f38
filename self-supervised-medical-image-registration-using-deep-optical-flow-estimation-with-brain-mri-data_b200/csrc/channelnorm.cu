// ChannelNorm forward/backward (replaces channelnorm_cuda; call sites flownet2/models.py:141-180).
//
//   fwd: out[b,0,p] = sqrt(sum_c x[b,c,p]^2)            algorithmic bytes 4*B*HW*(C+1)
//   bwd: gI[b,c,p]  = gO[b,0,p] * x[b,c,p] / (out[b,0,p] + 1e-8)        4*B*HW*(2C+2)
//
// Pure HBM streaming.  Each thread owns VEC consecutive pixels (128-bit loads and
// stores when HW % 4 == 0 and the bases are 16-byte aligned), loops the channels,
// and the grid is a whole number of waves over the 148 SMs (grid-stride).
#include "ofr_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kCtasPerSm = 8;
// every launch covers the whole tensor with one vector per thread (no grid-stride tail): the block scheduler keeps
// 8 x 256 threads resident per SM, each with its loads in flight

template <int VEC>
struct Pix;
template <>
struct Pix<4> {
    using T = float4;
    static __device__ __forceinline__ T ld(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
    static __device__ __forceinline__ void st(float* p, T v) { __stcs(reinterpret_cast<float4*>(p), v); }
};
template <>
struct Pix<1> {
    using T = float;
    static __device__ __forceinline__ T ld(const float* p) { return __ldcs(p); }
    static __device__ __forceinline__ void st(float* p, T v) { __stcs(p, v); }
};

template <int VEC>
__global__ void __launch_bounds__(kThreads) chnorm_fwd_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                              int C, long long hw, long long total_vec) {
    const long long hwv = hw / VEC;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_vec;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / hwv, p = (i - b * hwv) * VEC;
        const float* xp = x + b * C * hw + p;
        if constexpr (VEC == 4) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int c = 0; c < C; ++c) {
                float4 v = Pix<4>::ld(xp + (long long)c * hw);
                s.x = fmaf(v.x, v.x, s.x);
                s.y = fmaf(v.y, v.y, s.y);
                s.z = fmaf(v.z, v.z, s.z);
                s.w = fmaf(v.w, v.w, s.w);
            }
            Pix<4>::st(out + b * hw + p, make_float4(sqrtf(s.x), sqrtf(s.y), sqrtf(s.z), sqrtf(s.w)));
        } else {
            float s = 0.f;
#pragma unroll 4
            for (int c = 0; c < C; ++c) {
                float v = Pix<1>::ld(xp + (long long)c * hw);
                s = fmaf(v, v, s);
            }
            Pix<1>::st(out + b * hw + p, sqrtf(s));
        }
    }
}

template <int VEC>
__global__ void __launch_bounds__(kThreads) chnorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ out,
                                                              const float* __restrict__ gO, float* __restrict__ gI,
                                                              int C, long long hw, long long total_vec) {
    const long long hwv = hw / VEC;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_vec;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / hwv, p = (i - b * hwv) * VEC;
        const float* xp = x + b * C * hw + p;
        float* gp = gI + b * C * hw + p;
        if constexpr (VEC == 4) {
            const float4 o = Pix<4>::ld(out + b * hw + p);
            const float4 g = Pix<4>::ld(gO + b * hw + p);
            float4 r;  // gO / (out + 1e-8), one division per pixel instead of per element
            r.x = g.x / (o.x + 1e-8f);
            r.y = g.y / (o.y + 1e-8f);
            r.z = g.z / (o.z + 1e-8f);
            r.w = g.w / (o.w + 1e-8f);
#pragma unroll 4
            for (int c = 0; c < C; ++c) {
                float4 v = Pix<4>::ld(xp + (long long)c * hw);
                Pix<4>::st(gp + (long long)c * hw, make_float4(v.x * r.x, v.y * r.y, v.z * r.z, v.w * r.w));
            }
        } else {
            const float r = Pix<1>::ld(gO + b * hw + p) / (Pix<1>::ld(out + b * hw + p) + 1e-8f);
#pragma unroll 4
            for (int c = 0; c < C; ++c) Pix<1>::st(gp + (long long)c * hw, Pix<1>::ld(xp + (long long)c * hw) * r);
        }
    }
}

}  // namespace

extern "C" int ofr_chnorm_fwd(const float* x, float* out, int B, int C, int H, int W, ofr_stream_t stream) {
    if (!x || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0) return OFR_EINVAL;
    const long long hw = (long long)H * W;
    cudaStream_t st = (cudaStream_t)stream;
    if (hw % 4 == 0 && ofr::aligned16(x) && ofr::aligned16(out)) {
        const long long n = (long long)B * hw / 4;
        chnorm_fwd_kernel<4><<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, st>>>(x, out, C, hw, n);
    } else {
        const long long n = (long long)B * hw;
        chnorm_fwd_kernel<1><<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, st>>>(x, out, C, hw, n);
    }
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

extern "C" int ofr_chnorm_bwd(const float* x, const float* out, const float* gO, float* gI, int B, int C, int H,
                              int W, ofr_stream_t stream) {
    if (!x || !out || !gO || !gI || B <= 0 || C <= 0 || H <= 0 || W <= 0) return OFR_EINVAL;
    const long long hw = (long long)H * W;
    cudaStream_t st = (cudaStream_t)stream;
    if (hw % 4 == 0 && ofr::aligned16(x) && ofr::aligned16(out) && ofr::aligned16(gO) && ofr::aligned16(gI)) {
        const long long n = (long long)B * hw / 4;
        chnorm_bwd_kernel<4><<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, st>>>(x, out, gO, gI, C, hw, n);
    } else {
        const long long n = (long long)B * hw;
        chnorm_bwd_kernel<1><<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, st>>>(x, out, gO, gI, C, hw, n);
    }
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}
