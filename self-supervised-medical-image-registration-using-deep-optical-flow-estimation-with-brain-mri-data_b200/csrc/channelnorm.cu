// ChannelNorm forward/backward (replaces channelnorm_cuda; call sites flownet2/models.py:141-180).
//
//   fwd: out[b,0,p] = sqrt(sum_c x[b,c,p]^2)            algorithmic bytes 4*B*HW*(C+1)
//   bwd: gI[b,c,p]  = gO[b,0,p] * x[b,c,p] / (out[b,0,p] + 1e-8)        4*B*HW*(2C+2)
//
// Pure HBM streaming.  Each thread owns kVPT vectors of VEC consecutive pixels (128-bit loads and
// stores when HW % 4 == 0 and the bases are 16-byte aligned) and loops the channels.
#include "ofr_common.cuh"

namespace {

constexpr int kThreads = 256;

template <int VEC>
struct Pix;
template <>
struct Pix<4> {
    using T = float4;
    static __device__ __forceinline__ T ld(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
    static __device__ __forceinline__ void st(float* p, T v) { __stcs(reinterpret_cast<float4*>(p), v); }
    static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ T fma(T a, T b, T c) {
        return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
    }
    static __device__ __forceinline__ T sqrt(T a) { return make_float4(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z), sqrtf(a.w)); }
    static __device__ __forceinline__ T mul(T a, T b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
    static __device__ __forceinline__ T sub(T a, T b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
    static __device__ __forceinline__ T neg(T a) { return make_float4(-a.x, -a.y, -a.z, -a.w); }
    static __device__ __forceinline__ T ratio(T g, T o) {
        return make_float4(g.x / (o.x + 1e-8f), g.y / (o.y + 1e-8f), g.z / (o.z + 1e-8f), g.w / (o.w + 1e-8f));
    }
};
template <>
struct Pix<1> {
    using T = float;
    static __device__ __forceinline__ T ld(const float* p) { return __ldcs(p); }
    static __device__ __forceinline__ void st(float* p, T v) { __stcs(p, v); }
    static __device__ __forceinline__ T zero() { return 0.f; }
    static __device__ __forceinline__ T fma(T a, T b, T c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ T sqrt(T a) { return sqrtf(a); }
    static __device__ __forceinline__ T mul(T a, T b) { return a * b; }
    static __device__ __forceinline__ T sub(T a, T b) { return a - b; }
    static __device__ __forceinline__ T neg(T a) { return -a; }
    static __device__ __forceinline__ T ratio(T g, T o) { return g / (o + 1e-8f); }
};

constexpr int kVPT = 4;  // vectors per thread: their loads are all issued before the first use (bytes in flight per SM
                         // = resident threads x kVPT x 16 B; one vector per thread left the HBM pipe half empty)

// thread -> vectors (blockIdx.x * kVPT + j) * kThreads + threadIdx.x, j < kVPT: every load / store of a warp is one
// contiguous 512-byte run
// x2 != nullptr: the norm of the difference x - x2 (`x1 - resample(...)` then ChannelNorm at flownet2/models.py:140-141,
// 152-153, 166-167, 179-180) without materialising the difference
template <int VEC, bool DIFF>
__global__ void __launch_bounds__(kThreads) chnorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ x2,
                                                              float* __restrict__ out, int C, long long hw,
                                                              long long total_vec) {
    const long long hwv = hw / VEC;
    const long long i0 = (long long)blockIdx.x * kVPT * kThreads + threadIdx.x;
    const float* xp[kVPT];
    float* op[kVPT];
    bool ok[kVPT];
#pragma unroll
    for (int j = 0; j < kVPT; ++j) {
        const long long i = i0 + (long long)j * kThreads;
        ok[j] = i < total_vec;
        const long long ic = ok[j] ? i : 0;
        const long long b = ic / hwv, p = (ic - b * hwv) * VEC;
        xp[j] = x + b * C * hw + p;
        op[j] = out + b * hw + p;
    }
    const long long d2 = DIFF ? x2 - x : 0;  // element distance between the two inputs
    typename Pix<VEC>::T s[kVPT];
#pragma unroll
    for (int j = 0; j < kVPT; ++j) s[j] = Pix<VEC>::zero();
    for (int c = 0; c < C; ++c) {
        typename Pix<VEC>::T v[kVPT];
#pragma unroll
        for (int j = 0; j < kVPT; ++j) v[j] = ok[j] ? Pix<VEC>::ld(xp[j]) : Pix<VEC>::zero();
        if (DIFF) {
#pragma unroll
            for (int j = 0; j < kVPT; ++j)
                if (ok[j]) v[j] = Pix<VEC>::sub(v[j], Pix<VEC>::ld(xp[j] + d2));
        }
#pragma unroll
        for (int j = 0; j < kVPT; ++j) {
            s[j] = Pix<VEC>::fma(v[j], v[j], s[j]);
            xp[j] += hw;
        }
    }
#pragma unroll
    for (int j = 0; j < kVPT; ++j)
        if (ok[j]) Pix<VEC>::st(op[j], Pix<VEC>::sqrt(s[j]));
}

// x2 != nullptr: backward of the norm of x - x2; gI gets the gradient of x, gI2 (optional) its negative for x2
template <int VEC, bool DIFF>
__global__ void __launch_bounds__(kThreads) chnorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ x2,
                                                              const float* __restrict__ out,
                                                              const float* __restrict__ gO, float* __restrict__ gI,
                                                              float* __restrict__ gI2, int C, long long hw,
                                                              long long total_vec) {
    const long long hwv = hw / VEC;
    const long long i0 = (long long)blockIdx.x * kVPT * kThreads + threadIdx.x;
    const float* xp[kVPT];
    float* gp[kVPT];
    bool ok[kVPT];
    typename Pix<VEC>::T r[kVPT];  // gO / (out + 1e-8), one division per pixel instead of per element
#pragma unroll
    for (int j = 0; j < kVPT; ++j) {
        const long long i = i0 + (long long)j * kThreads;
        ok[j] = i < total_vec;
        const long long ic = ok[j] ? i : 0;
        const long long b = ic / hwv, p = (ic - b * hwv) * VEC;
        xp[j] = x + b * C * hw + p;
        gp[j] = (DIFF && !gI) ? gI2 + b * C * hw + p : gI + b * C * hw + p;  // DIFF: gI and / or gI2
        r[j] = Pix<VEC>::ratio(Pix<VEC>::ld(gO + b * hw + p), Pix<VEC>::ld(out + b * hw + p));
    }
    const long long d2 = DIFF ? x2 - x : 0, dg = (DIFF && gI && gI2) ? gI2 - gI : 0;
    for (int c = 0; c < C; ++c) {
        typename Pix<VEC>::T v[kVPT];
#pragma unroll
        for (int j = 0; j < kVPT; ++j) v[j] = ok[j] ? Pix<VEC>::ld(xp[j]) : Pix<VEC>::zero();
        if (DIFF) {
#pragma unroll
            for (int j = 0; j < kVPT; ++j)
                if (ok[j]) v[j] = Pix<VEC>::sub(v[j], Pix<VEC>::ld(xp[j] + d2));
        }
#pragma unroll
        for (int j = 0; j < kVPT; ++j) {
            const typename Pix<VEC>::T gq = Pix<VEC>::mul(v[j], r[j]);
            if (!DIFF) {
                if (ok[j]) Pix<VEC>::st(gp[j], gq);
            } else if (ok[j]) {
                if (gI) Pix<VEC>::st(gp[j], gq);
                if (gI2) Pix<VEC>::st(gI ? gp[j] + dg : gp[j], Pix<VEC>::neg(gq));
            }
            xp[j] += hw;
            gp[j] += hw;
        }
    }
}

}  // namespace

extern "C" int ofr_chnorm_fwd(const float* x, float* out, int B, int C, int H, int W, ofr_stream_t stream) {
    return ofr_chnorm_diff_fwd(x, nullptr, out, B, C, H, W, stream);
}

extern "C" int ofr_chnorm_diff_fwd(const float* x, const float* x2, float* out, int B, int C, int H, int W,
                                   ofr_stream_t stream) {
    if (!x || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0) return OFR_EINVAL;
    const long long hw = (long long)H * W;
    cudaStream_t st = (cudaStream_t)stream;
    if (hw % 4 == 0 && ofr::aligned16(x) && ofr::aligned16(out) && (!x2 || ofr::aligned16(x2))) {
        const long long n = (long long)B * hw / 4;
        if (x2) chnorm_fwd_kernel<4, true><<<(unsigned)((n + kThreads * kVPT - 1) / (kThreads * kVPT)), kThreads, 0, st>>>(x, x2, out, C, hw, n);
        else chnorm_fwd_kernel<4, false><<<(unsigned)((n + kThreads * kVPT - 1) / (kThreads * kVPT)), kThreads, 0, st>>>(x, x2, out, C, hw, n);
    } else {
        const long long n = (long long)B * hw;
        if (x2) chnorm_fwd_kernel<1, true><<<(unsigned)((n + kThreads * kVPT - 1) / (kThreads * kVPT)), kThreads, 0, st>>>(x, x2, out, C, hw, n);
        else chnorm_fwd_kernel<1, false><<<(unsigned)((n + kThreads * kVPT - 1) / (kThreads * kVPT)), kThreads, 0, st>>>(x, x2, out, C, hw, n);
    }
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

extern "C" int ofr_chnorm_bwd(const float* x, const float* out, const float* gO, float* gI, int B, int C, int H,
                              int W, ofr_stream_t stream) {
    if (!gI) return OFR_EINVAL;
    return ofr_chnorm_diff_bwd(x, nullptr, out, gO, gI, nullptr, B, C, H, W, stream);
}

extern "C" int ofr_chnorm_diff_bwd(const float* x, const float* x2, const float* out, const float* gO, float* gI,
                                   float* gI2, int B, int C, int H, int W, ofr_stream_t stream) {
    if (!x || !out || !gO || (!gI && !gI2) || (gI2 && !x2) || B <= 0 || C <= 0 || H <= 0 || W <= 0) return OFR_EINVAL;
    const long long hw = (long long)H * W;
    cudaStream_t st = (cudaStream_t)stream;
    const bool al = ofr::aligned16(x) && ofr::aligned16(out) && ofr::aligned16(gO) && (!gI || ofr::aligned16(gI)) &&
                    (!gI2 || ofr::aligned16(gI2)) && (!x2 || ofr::aligned16(x2));
    if (hw % 4 == 0 && al) {
        const long long n = (long long)B * hw / 4;
        if (x2) chnorm_bwd_kernel<4, true><<<(unsigned)((n + kThreads * kVPT - 1) / (kThreads * kVPT)), kThreads, 0, st>>>(x, x2, out, gO, gI, gI2, C, hw, n);
        else chnorm_bwd_kernel<4, false><<<(unsigned)((n + kThreads * kVPT - 1) / (kThreads * kVPT)), kThreads, 0, st>>>(x, x2, out, gO, gI, gI2, C, hw, n);
    } else {
        const long long n = (long long)B * hw;
        if (x2) chnorm_bwd_kernel<1, true><<<(unsigned)((n + kThreads * kVPT - 1) / (kThreads * kVPT)), kThreads, 0, st>>>(x, x2, out, gO, gI, gI2, C, hw, n);
        else chnorm_bwd_kernel<1, false><<<(unsigned)((n + kThreads * kVPT - 1) / (kThreads * kVPT)), kThreads, 0, st>>>(x, x2, out, gO, gI, gI2, C, hw, n);
    }
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}
