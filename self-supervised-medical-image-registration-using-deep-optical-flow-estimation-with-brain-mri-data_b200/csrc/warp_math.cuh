// Sampling-position arithmetic shared by the warp kernels (resample2d.cu) and the fused registration loss (loss.cu):
// both must form the float32 sampling positions, tap offsets and bilinear weights in exactly the same operation order
// (which is PyTorch's, spelled with round-to-nearest intrinsics so the compiler cannot contract or reassociate it).
#pragma once
#include "ofr_common.cuh"

namespace ofr_warp {

struct WarpGeom {
    int B, C, Hin, Win, Hf, Wf;
    float fx, fy;  // STN: float32(2/W), float32(2/H)
    float fs;      // the flow is multiplied by fs first (1 = plain call; PWCNet.py:214 passes flo * 0.625 ...)
};

struct Samp {
    float wx0, wx1, wy0, wy1;  // bilinear weights of the left/right and top/bottom taps
    int x0, y0;                // floor of the sampling position, clamped to [-2, W] / [-2, H]
    bool live;                 // false: contributes nothing (PWC mask) -> output 0, no gradient
};

__device__ __forceinline__ int safe_floor_int(float f, int hi) {
    // clamp in float first so huge / NaN flows cannot overflow the int conversion
    f = fminf(fmaxf(f, -2.f), (float)hi);
    return (int)f;
}

template <int MODE>
__device__ __forceinline__ Samp make_samp(int x, int y, float u, float v, const WarpGeom& g) {
    Samp s;
    float ix, iy;
    u = __fmul_rn(u, g.fs);  // exact identity for fs = 1
    v = __fmul_rn(v, g.fs);
    if (MODE == OFR_WARP_FLOWNET2 || MODE == OFR_WARP_FLOWNET2_NEAREST) {
        ix = __fadd_rn((float)x, u);
        iy = __fadd_rn((float)y, v);
    } else if (MODE == OFR_WARP_STN) {
        // (x / 2 == x * 0.5 exactly in binary floating point)
        float gx = __fsub_rn(__fmul_rn(__fadd_rn(u, (float)x), g.fx), 1.f);
        float gy = __fsub_rn(__fmul_rn(__fadd_rn(v, (float)y), g.fy), 1.f);
        ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), (float)(g.Win - 1));
        iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), (float)(g.Hin - 1));
    } else {  // PWC
        const float dx = (float)max(g.Win - 1, 1), dy = (float)max(g.Hin - 1, 1);
        float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, __fadd_rn((float)x, u)), dx), 1.f);
        float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, __fadd_rn((float)y, v)), dy), 1.f);
        ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)g.Win), 1.f), 0.5f);
        iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)g.Hin), 1.f), 0.5f);
    }
    const float fxl = floorf(ix), fyl = floorf(iy);
    if (MODE == OFR_WARP_FLOWNET2 || MODE == OFR_WARP_FLOWNET2_NEAREST) {
        s.wx1 = __fsub_rn(ix, fxl);
        s.wy1 = __fsub_rn(iy, fyl);
        s.wx0 = __fsub_rn(1.f, s.wx1);
        s.wy0 = __fsub_rn(1.f, s.wy1);
    } else {  // grid_sample: nw = (ix_se - ix) * (iy_se - iy) ...
        s.wx0 = __fsub_rn(__fadd_rn(fxl, 1.f), ix);
        s.wx1 = __fsub_rn(ix, fxl);
        s.wy0 = __fsub_rn(__fadd_rn(fyl, 1.f), iy);
        s.wy1 = __fsub_rn(iy, fyl);
    }
    s.x0 = safe_floor_int(fxl, g.Win);
    s.y0 = safe_floor_int(fyl, g.Hin);
    s.live = true;
    if (MODE == OFR_WARP_PWC) {
        const bool xl = (unsigned)s.x0 < (unsigned)g.Win, xr = (unsigned)(s.x0 + 1) < (unsigned)g.Win;
        const bool yt = (unsigned)s.y0 < (unsigned)g.Hin, yb = (unsigned)(s.y0 + 1) < (unsigned)g.Hin;
        float m = 0.f;  // the warped ones-image, accumulated nw, ne, sw, se like grid_sample
        if (xl && yt) m = __fadd_rn(m, __fmul_rn(s.wx0, s.wy0));
        if (xr && yt) m = __fadd_rn(m, __fmul_rn(s.wx1, s.wy0));
        if (xl && yb) m = __fadd_rn(m, __fmul_rn(s.wx0, s.wy1));
        if (xr && yb) m = __fadd_rn(m, __fmul_rn(s.wx1, s.wy1));
        s.live = m >= 0.9999f;
    }
    return s;
}

// Tap column/row for tap offset d in {0,1}: clamped (FLOWNET2) or -1 when outside (zeros padding).
template <int MODE>
__device__ __forceinline__ int tap_index(int i0, int d, int n) {
    int t = i0 + d;
    if (MODE == OFR_WARP_FLOWNET2 || MODE == OFR_WARP_FLOWNET2_NEAREST) return min(max(t, 0), n - 1);
    return ((unsigned)t < (unsigned)n) ? t : -1;
}

// The four taps of a sample as element offsets into one image plane (-1: outside, reads as 0).
struct TapOff {
    int tl, tr, bl, br;
};

template <int MODE>
__device__ __forceinline__ TapOff tap_offsets(int x0, int y0, int H, int W) {
    const int xl = tap_index<MODE>(x0, 0, W), xr = tap_index<MODE>(x0, 1, W);
    const int yt = tap_index<MODE>(y0, 0, H), yb = tap_index<MODE>(y0, 1, H);
    TapOff t;
    t.tl = (xl >= 0 && yt >= 0) ? yt * W + xl : -1;
    t.tr = (xr >= 0 && yt >= 0) ? yt * W + xr : -1;
    t.bl = (xl >= 0 && yb >= 0) ? yb * W + xl : -1;
    t.br = (xr >= 0 && yb >= 0) ? yb * W + xr : -1;
    return t;
}

// A pointer the compiler must keep as one 64-bit register pair: element accesses p[int offset] then cost one IMAD.WIDE.
// (Left alone, nvcc folds kernel-argument pointers plus image/plane offsets into a 64-bit element index and rebuilds
// every address with IADD3 + LEA.HI.X.SX32 + LEA + LEA.HI.X -- four issue slots per load in kernels that are issue-bound.)
template <class T>
__device__ __forceinline__ T* opaque_ptr(T* p) {
    unsigned long long v = reinterpret_cast<unsigned long long>(p);
    asm volatile("" : "+l"(v));
    return reinterpret_cast<T*>(v);
}

__device__ __forceinline__ float fetch(const float* __restrict__ img, int off) { return off >= 0 ? __ldg(img + off) : 0.f; }

// Forward taps: four ALWAYS-LOADABLE element offsets and the four bilinear products.  Border mode (FlowNet2): the taps
// are clamped into the image.  Zeros mode (stn, PWC): the addresses are clamped too, and the 1-D weight of an
// out-of-range column / row is zeroed first, so the products of the taps grid_sample skips are exactly 0 and the sum
// nw, ne, sw, se below equals its sum over the valid taps bit for bit (x + 0 = x) -- no predicated loads, no -1
// offsets.  (The image is assumed finite: 0 * Inf would not be skipped.)
struct FwdTaps {
    int tl, tr, bl, br;
    float w00, w01, w10, w11;
};

template <int MODE>
__device__ __forceinline__ FwdTaps fwd_taps(const Samp& s, int H, int W) {
    FwdTaps t;
    const int x1 = s.x0 + 1, y1 = s.y0 + 1;
    const int xl = min(max(s.x0, 0), W - 1), xr = min(max(x1, 0), W - 1);
    const int yt = min(max(s.y0, 0), H - 1) * W, yb = min(max(y1, 0), H - 1) * W;
    float wx0 = s.wx0, wx1 = s.wx1, wy0 = s.wy0, wy1 = s.wy1;
    if (MODE == OFR_WARP_STN || MODE == OFR_WARP_PWC) {
        wx0 = (unsigned)s.x0 < (unsigned)W ? wx0 : 0.f;
        wx1 = (unsigned)x1 < (unsigned)W ? wx1 : 0.f;
        wy0 = (unsigned)s.y0 < (unsigned)H ? wy0 : 0.f;
        wy1 = (unsigned)y1 < (unsigned)H ? wy1 : 0.f;
        if (MODE == OFR_WARP_PWC && !s.live) wx0 = wx1 = 0.f;  // masked pixel: every product 0 -> the output is 0
    }
    t.tl = yt + xl, t.tr = yt + xr, t.bl = yb + xl, t.br = yb + xr;
    t.w00 = __fmul_rn(wx0, wy0);
    t.w01 = __fmul_rn(wx1, wy0);
    t.w10 = __fmul_rn(wx0, wy1);
    t.w11 = __fmul_rn(wx1, wy1);
    return t;
}

// the bilinear sum in grid_sample's order (nw, ne, sw, se)
__device__ __forceinline__ float tap_sum(const float* __restrict__ img, const FwdTaps& t) {
    float a = __fmul_rn(t.w00, __ldg(img + t.tl));
    a = fmaf(t.w01, __ldg(img + t.tr), a);
    a = fmaf(t.w10, __ldg(img + t.bl), a);
    a = fmaf(t.w11, __ldg(img + t.br), a);
    return a;
}

}  // namespace ofr_warp
