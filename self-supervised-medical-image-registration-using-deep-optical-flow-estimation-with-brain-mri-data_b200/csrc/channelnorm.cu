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

// blockIdx.y = image, thread -> vectors (blockIdx.x * kVPT + j) * kThreads + threadIdx.x of that image, j < kVPT: every
// load / store of a warp is one contiguous 512-byte run.  What made the first version slow (tools/micro/stream_variants.cu,
// 50 MB one-channel forward under back-to-back launches: 12.9 us against 9.5 us for the same loads and stores without
// them): (i) a flat index split by 64-bit divisions in front of the first load (-1.6 us without), (ii) a run-time channel
// loop around PREDICATED loads, which the compiler serialises behind branches (-1.5 us with the channel count a template
// constant or the loads unpredicated).  So: no division (2-D grid), CC = compile-time channel count for the model's
// C = 1, 2, 3 (0 = run-time loop), and tail vectors are loaded from a clamped (valid) index -- only the stores are
// predicated.
// x2 != nullptr: the norm of the difference x - x2 (`x1 - resample(...)` then ChannelNorm at flownet2/models.py:140-141,
// 152-153, 166-167, 179-180) without materialising the difference
template <int VEC, bool DIFF, int CC>
__global__ void __launch_bounds__(kThreads) chnorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ x2,
                                                              float* __restrict__ out, int Crt, long long hw, int B) {
    const int C = CC ? CC : Crt;
    const long long hwv = hw / VEC;
    const long long i0 = (long long)blockIdx.x * kVPT * kThreads + threadIdx.x;
    long long p[kVPT];
    bool ok[kVPT];
#pragma unroll
    for (int j = 0; j < kVPT; ++j) {
        const long long i = i0 + (long long)j * kThreads;
        ok[j] = i < hwv;
        p[j] = (ok[j] ? i : hwv - 1) * VEC;
    }
    const long long d2 = DIFF ? x2 - x : 0;  // element distance between the two inputs
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const float* xb = x + (long long)b * C * hw;
        typename Pix<VEC>::T s[kVPT];
#pragma unroll
        for (int j = 0; j < kVPT; ++j) s[j] = Pix<VEC>::zero();
#pragma unroll
        for (int c = 0; c < (CC ? CC : 1); ++c) {       // CC > 0: unrolled; CC == 0: one pass of the run-time loop below
            for (int cr = 0; cr < (CC ? 1 : C); ++cr) {
                const float* xc = xb + (long long)(CC ? c : cr) * hw;
                typename Pix<VEC>::T v[kVPT];
#pragma unroll
                for (int j = 0; j < kVPT; ++j) v[j] = Pix<VEC>::ld(xc + p[j]);
                if (DIFF) {
#pragma unroll
                    for (int j = 0; j < kVPT; ++j) v[j] = Pix<VEC>::sub(v[j], Pix<VEC>::ld(xc + p[j] + d2));
                }
#pragma unroll
                for (int j = 0; j < kVPT; ++j) s[j] = Pix<VEC>::fma(v[j], v[j], s[j]);
            }
        }
        float* ob = out + (long long)b * hw;
#pragma unroll
        for (int j = 0; j < kVPT; ++j)
            if (ok[j]) Pix<VEC>::st(ob + p[j], Pix<VEC>::sqrt(s[j]));
    }
}

// x2 != nullptr: backward of the norm of x - x2; gI gets the gradient of x, gI2 (optional) its negative for x2
template <int VEC, bool DIFF, int CC>
__global__ void __launch_bounds__(kThreads) chnorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ x2,
                                                              const float* __restrict__ out,
                                                              const float* __restrict__ gO, float* __restrict__ gI,
                                                              float* __restrict__ gI2, int Crt, long long hw, int B) {
    const int C = CC ? CC : Crt;
    const long long hwv = hw / VEC;
    const long long i0 = (long long)blockIdx.x * kVPT * kThreads + threadIdx.x;
    long long p[kVPT];
    bool ok[kVPT];
#pragma unroll
    for (int j = 0; j < kVPT; ++j) {
        const long long i = i0 + (long long)j * kThreads;
        ok[j] = i < hwv;
        p[j] = (ok[j] ? i : hwv - 1) * VEC;
    }
    const long long d2 = DIFF ? x2 - x : 0;
    float* g1 = (DIFF && !gI) ? gI2 : gI;                       // DIFF: gI and / or gI2
    const bool both = DIFF && gI && gI2;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        typename Pix<VEC>::T r[kVPT];  // gO / (out + 1e-8), one division per pixel instead of per element
#pragma unroll
        for (int j = 0; j < kVPT; ++j)
            r[j] = Pix<VEC>::ratio(Pix<VEC>::ld(gO + (long long)b * hw + p[j]), Pix<VEC>::ld(out + (long long)b * hw + p[j]));
        const long long base = (long long)b * C * hw;
#pragma unroll
        for (int c = 0; c < (CC ? CC : 1); ++c) {
            for (int cr = 0; cr < (CC ? 1 : C); ++cr) {
                const long long off = base + (long long)(CC ? c : cr) * hw;
                typename Pix<VEC>::T v[kVPT];
#pragma unroll
                for (int j = 0; j < kVPT; ++j) v[j] = Pix<VEC>::ld(x + off + p[j]);
                if (DIFF) {
#pragma unroll
                    for (int j = 0; j < kVPT; ++j) v[j] = Pix<VEC>::sub(v[j], Pix<VEC>::ld(x + off + p[j] + d2));
                }
#pragma unroll
                for (int j = 0; j < kVPT; ++j) {
                    const typename Pix<VEC>::T gq = Pix<VEC>::mul(v[j], r[j]);
                    if (ok[j]) {
                        if (!DIFF) {
                            Pix<VEC>::st(gI + off + p[j], gq);
                        } else {
                            Pix<VEC>::st(g1 + off + p[j], (gI ? gq : Pix<VEC>::neg(gq)));
                            if (both) Pix<VEC>::st(gI2 + off + p[j], Pix<VEC>::neg(gq));
                        }
                    }
                }
            }
        }
    }
}

// x: vectors of one image / (kThreads * kVPT); y: images (a CTA walks several when B exceeds the grid limit)
dim3 grid_for_image(long long vec_per_image, int B) {
    const long long gx = (vec_per_image + kThreads * kVPT - 1) / (kThreads * kVPT);
    return dim3((unsigned)gx, (unsigned)(B < 65535 ? B : 65535));
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
    const bool v4 = hw % 4 == 0 && ofr::aligned16(x) && ofr::aligned16(out) && (!x2 || ofr::aligned16(x2));
    const dim3 grid = grid_for_image(v4 ? hw / 4 : hw, B);
#define OFR_CHN_FWD(VEC, DIFF, CC) chnorm_fwd_kernel<VEC, DIFF, CC><<<grid, kThreads, 0, st>>>(x, x2, out, C, hw, B)
#define OFR_CHN_FWD_C(VEC, DIFF)                                                                          \
    do {                                                                                                  \
        if (C == 1) OFR_CHN_FWD(VEC, DIFF, 1); else if (C == 2) OFR_CHN_FWD(VEC, DIFF, 2);                \
        else if (C == 3) OFR_CHN_FWD(VEC, DIFF, 3); else OFR_CHN_FWD(VEC, DIFF, 0);                       \
    } while (0)
    if (v4) {
        if (x2) OFR_CHN_FWD_C(4, true); else OFR_CHN_FWD_C(4, false);
    } else {
        if (x2) OFR_CHN_FWD(1, true, 0); else OFR_CHN_FWD(1, false, 0);
    }
#undef OFR_CHN_FWD_C
#undef OFR_CHN_FWD
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
    const bool v4 = hw % 4 == 0 && al;
    const dim3 grid = grid_for_image(v4 ? hw / 4 : hw, B);
#define OFR_CHN_BWD(VEC, DIFF, CC) chnorm_bwd_kernel<VEC, DIFF, CC><<<grid, kThreads, 0, st>>>(x, x2, out, gO, gI, gI2, C, hw, B)
#define OFR_CHN_BWD_C(VEC, DIFF)                                                                          \
    do {                                                                                                  \
        if (C == 1) OFR_CHN_BWD(VEC, DIFF, 1); else if (C == 2) OFR_CHN_BWD(VEC, DIFF, 2);                \
        else if (C == 3) OFR_CHN_BWD(VEC, DIFF, 3); else OFR_CHN_BWD(VEC, DIFF, 0);                       \
    } while (0)
    if (v4) {
        if (x2) OFR_CHN_BWD_C(4, true); else OFR_CHN_BWD_C(4, false);
    } else {
        if (x2) OFR_CHN_BWD(1, true, 0); else OFR_CHN_BWD(1, false, 0);
    }
#undef OFR_CHN_BWD_C
#undef OFR_CHN_BWD
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}
