// Bilinear warps: Resample2d (flownet2), opticalFlowReg.stn and PWCDCNet.warp in one kernel family.
//
// Reference semantics:
//   FLOWNET2 : flownet2/models.py:139,151,163,176  sample (x+u, y+v), taps clamped, fractional weights unclamped
//   STN      : models.py:256-268                   g=(p+f)*(2/W,2/H)-1, grid_sample(zeros, align_corners=True)
//   PWC      : PWC/models/PWCNet.py:143-179        g=2(p+f)/(W-1)-1, grid_sample(zeros, align_corners=False),
//                                                  output zeroed where the warped ones-mask < 0.9999
// The coordinate arithmetic is spelled with explicit round-to-nearest intrinsics in the
// same operation order PyTorch uses, so that the float32 sampling positions agree with
// the reference bit for bit (a 1-ulp difference in a normalised coordinate is 1.5e-5 px at W=512).
//
// Forward  : one thread per 4 consecutive output pixels (128-bit flow loads / output stores when
//            Wf % 4 == 0), channel chunks on blockIdx.y.  Algorithmic bytes 4*B*Hf*Wf*(2C+2).
// Backward : gI2 (flow) is a per-source gather + analytic derivative, channels split over the CTA's
//            y-dimension and combined through shared memory in a fixed order.
//            gI1 (image) is the adjoint scatter.  Upstream used float atomicAdd (non-deterministic);
//            here every source pixel is binned by the top-left cell of its 2x2 footprint
//            (integer atomics only: counts are order independent), the bins are turned into a CSR
//            by an exclusive scan, and each *target* pixel then gathers from the <= 4 cells that can
//            reach it, visiting the entries of a cell in ascending source index.  The floating-point
//            summation order is therefore a function of the data only: bitwise reproducible.
//            Cells holding more than kHeavy entries are sorted once by a CTA-wide bitonic network so
//            the gather never does quadratic work on degenerate flows.
#include "ofr_common.cuh"

namespace {

constexpr int kHeavy = 32;       // cells with more entries than this are pre-sorted
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

struct WarpGeom {
    int B, C, Hin, Win, Hf, Wf;
    float fx, fy;  // STN: float32(2/W), float32(2/H)
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
    if (MODE == OFR_WARP_FLOWNET2 || MODE == OFR_WARP_FLOWNET2_NEAREST) {
        ix = __fadd_rn((float)x, u);
        iy = __fadd_rn((float)y, v);
    } else if (MODE == OFR_WARP_STN) {
        float gx = __fsub_rn(__fmul_rn(__fadd_rn(u, (float)x), g.fx), 1.f);
        float gy = __fsub_rn(__fmul_rn(__fadd_rn(v, (float)y), g.fy), 1.f);
        ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(g.Win - 1));
        iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)(g.Hin - 1));
    } else {  // PWC
        const float dx = (float)max(g.Win - 1, 1), dy = (float)max(g.Hin - 1, 1);
        float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, __fadd_rn((float)x, u)), dx), 1.f);
        float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, __fadd_rn((float)y, v)), dy), 1.f);
        ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)g.Win), 1.f), 2.f);
        iy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)g.Hin), 1.f), 2.f);
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
        const bool xl = s.x0 >= 0 && s.x0 < g.Win, xr = s.x0 + 1 >= 0 && s.x0 + 1 < g.Win;
        const bool yt = s.y0 >= 0 && s.y0 < g.Hin, yb = s.y0 + 1 >= 0 && s.y0 + 1 < g.Hin;
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
    return (t >= 0 && t < n) ? t : -1;
}

template <int MODE>
struct Taps {
    int xl, xr, yt, yb;
    __device__ __forceinline__ Taps(const Samp& s, const WarpGeom& g) {
        xl = tap_index<MODE>(s.x0, 0, g.Win);
        xr = tap_index<MODE>(s.x0, 1, g.Win);
        yt = tap_index<MODE>(s.y0, 0, g.Hin);
        yb = tap_index<MODE>(s.y0, 1, g.Hin);
    }
    __device__ __forceinline__ float fetch(const float* __restrict__ img, int y, int x, int W) const {
        return (x >= 0 && y >= 0) ? __ldg(img + (long long)y * W + x) : 0.f;
    }
};

template <int MODE>
__device__ __forceinline__ float bilerp(const float* __restrict__ img, const Samp& s, const Taps<MODE>& t, int W) {
    float acc = __fmul_rn(__fmul_rn(s.wx0, s.wy0), t.fetch(img, t.yt, t.xl, W));
    acc = fmaf(__fmul_rn(s.wx1, s.wy0), t.fetch(img, t.yt, t.xr, W), acc);
    acc = fmaf(__fmul_rn(s.wx0, s.wy1), t.fetch(img, t.yb, t.xl, W), acc);
    acc = fmaf(__fmul_rn(s.wx1, s.wy1), t.fetch(img, t.yb, t.xr, W), acc);
    return acc;
}

// ------------------------------------------------------------------------------------------------ forward
constexpr int kFwdThreads = 256;
constexpr int kFwdChunk = 8;  // channels per blockIdx.y slice

template <int MODE, int PX>
__global__ void __launch_bounds__(kFwdThreads, 3) resample_fwd_kernel(const float* __restrict__ in1,
                                                                      const float* __restrict__ flow,
                                                                      float* __restrict__ out, WarpGeom g) {
    // one work item (PX consecutive pixels of a row) per thread: every CTA issues its flow loads at once and the
    // block scheduler, not a grid-stride loop, keeps the SMs full -> more bytes in flight per SM
    const long long hwf = (long long)g.Hf * g.Wf, hwi = (long long)g.Hin * g.Win;
    const long long total = (long long)g.B * hwf / PX;
    const int c0 = blockIdx.y * kFwdChunk, c1 = min(c0 + kFwdChunk, g.C);
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long pix = i * PX;
    const int b = (int)(pix / hwf);
    const int rem = (int)(pix - (long long)b * hwf);
    const int y = rem / g.Wf, x = rem - y * g.Wf;
    const float* fu = flow + (long long)b * 2 * hwf + rem;
    float u[PX], v[PX];
    if (PX == 4) {
        float4 a = ofr::ldg4(fu), c = ofr::ldg4(fu + hwf);
        u[0] = a.x, u[1 % PX] = a.y, u[2 % PX] = a.z, u[3 % PX] = a.w;
        v[0] = c.x, v[1 % PX] = c.y, v[2 % PX] = c.z, v[3 % PX] = c.w;
    } else {
        u[0] = __ldg(fu);
        v[0] = __ldg(fu + hwf);
    }
    Samp s[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) s[k] = make_samp<MODE>(x + k, y, u[k], v[k], g);
    for (int c = c0; c < c1; ++c) {
        const float* img = in1 + ((long long)b * g.C + c) * hwi;
        float r[PX];
#pragma unroll
        for (int k = 0; k < PX; ++k) {
            if (MODE == OFR_WARP_FLOWNET2_NEAREST) {
                // nearest tap: floor(x + u + 0.5), clamped
                const float xn = floorf(__fadd_rn(__fadd_rn((float)(x + k), u[k]), 0.5f));
                const float yn = floorf(__fadd_rn(__fadd_rn((float)y, v[k]), 0.5f));
                const int xi = min(max(safe_floor_int(xn, g.Win), 0), g.Win - 1);
                const int yi = min(max(safe_floor_int(yn, g.Hin), 0), g.Hin - 1);
                r[k] = __ldg(img + (long long)yi * g.Win + xi);
            } else {
                Taps<MODE> t(s[k], g);
                r[k] = s[k].live ? bilerp<MODE>(img, s[k], t, g.Win) : 0.f;
            }
        }
        float* o = out + ((long long)b * g.C + c) * hwf + rem;
        if (PX == 4)
            ofr::stcs4(o, make_float4(r[0], r[1 % PX], r[2 % PX], r[3 % PX]));
        else
            o[0] = r[0];
    }
}

// ------------------------------------------------------------------------------------------------ backward
// Cell of a source pixel: top-left corner of its footprint, kx in [-1, Win-1], ky in [-1, Hin-1];
// -1 when the source reaches no target at all.
template <int MODE>
__device__ __forceinline__ int cell_of(const Samp& s, const WarpGeom& g) {
    if (!s.live) return -1;
    int kx, ky;
    if (MODE == OFR_WARP_FLOWNET2) {
        kx = min(max(s.x0, -1), g.Win - 1);
        ky = min(max(s.y0, -1), g.Hin - 1);
    } else {
        if (s.x0 < -1 || s.x0 > g.Win - 1 || s.y0 < -1 || s.y0 > g.Hin - 1) return -1;
        kx = s.x0;
        ky = s.y0;
    }
    return (ky + 1) * (g.Win + 1) + (kx + 1);
}

// Per-source pass: flow gradient (optional) and cell binning (optional).
// blockDim = (32, CG): 32 consecutive source pixels, CG channel groups.
template <int MODE, int CG, bool WANT_G2, bool WANT_BIN>
__global__ void __launch_bounds__(32 * CG) resample_bwd_src_kernel(const float* __restrict__ in1,
                                                                   const float* __restrict__ flow,
                                                                   const float* __restrict__ gO,
                                                                   float* __restrict__ gI2, int* __restrict__ counts,
                                                                   int* __restrict__ key, int* __restrict__ rank,
                                                                   WarpGeom g, float sx, float sy) {
    __shared__ float red[2][CG][32];
    const long long hwf = (long long)g.Hf * g.Wf, hwi = (long long)g.Hin * g.Win;
    const long long total = (long long)g.B * hwf;
    const long long pix = blockIdx.x * 32ll + threadIdx.x;
    const bool in_range = pix < total;
    int b = 0, rem = 0, y = 0, x = 0;
    Samp s;
    s.live = false;
    if (in_range) {
        b = (int)(pix / hwf);
        rem = (int)(pix - (long long)b * hwf);
        y = rem / g.Wf;
        x = rem - y * g.Wf;
        const float* fu = flow + (long long)b * 2 * hwf + rem;
        s = make_samp<MODE>(x, y, __ldg(fu), __ldg(fu + hwf), g);
    }
    if (WANT_BIN && threadIdx.y == 0 && in_range) {
        const int cell = cell_of<MODE>(s, g);
        int r = 0;
        if (cell >= 0) r = atomicAdd(counts + (long long)b * (g.Hin + 1) * (g.Win + 1) + cell, 1);
        key[pix] = cell;
        rank[pix] = r;
    }
    if (!WANT_G2) return;
    float ax = 0.f, ay = 0.f;
    if (in_range && s.live) {
        Taps<MODE> t(s, g);
#pragma unroll 8
        for (int c = threadIdx.y; c < g.C; c += CG) {
            const float* img = in1 + ((long long)b * g.C + c) * hwi;
            const float go = __ldg(gO + ((long long)b * g.C + c) * hwf + rem);
            const float tl = t.fetch(img, t.yt, t.xl, g.Win), tr = t.fetch(img, t.yt, t.xr, g.Win);
            const float bl = t.fetch(img, t.yb, t.xl, g.Win), br = t.fetch(img, t.yb, t.xr, g.Win);
            // d out / d ix = wy0*(TR-TL) + wy1*(BR-BL);  d out / d iy = wx0*(BL-TL) + wx1*(BR-TR)
            ax = fmaf(go, fmaf(s.wy1, br - bl, s.wy0 * (tr - tl)), ax);
            ay = fmaf(go, fmaf(s.wx1, br - tr, s.wx0 * (bl - tl)), ay);
        }
    }
    if (CG > 1) {
        red[0][threadIdx.y][threadIdx.x] = ax;
        red[1][threadIdx.y][threadIdx.x] = ay;
        __syncthreads();
        if (threadIdx.y == 0) {
            ax = red[0][0][threadIdx.x];
            ay = red[1][0][threadIdx.x];
#pragma unroll
            for (int k = 1; k < CG; ++k) {  // fixed order -> deterministic
                ax += red[0][k][threadIdx.x];
                ay += red[1][k][threadIdx.x];
            }
        }
    }
    if (threadIdx.y == 0 && in_range) {
        float* o = gI2 + (long long)b * 2 * hwf + rem;
        o[0] = ax * sx;
        o[hwf] = ay * sy;
    }
}

// Few-channel variant (C < 8: the full-resolution FlowNet2 / stn warps, C = 1): one thread owns 4 consecutive
// source pixels, 128-bit loads of flow and gO, 128-bit store of the flow gradient; channels looped in the thread.
template <int MODE, bool WANT_G2, bool WANT_BIN>
__global__ void __launch_bounds__(256, 3) resample_bwd_src4_kernel(const float* __restrict__ in1,
                                                                   const float* __restrict__ flow,
                                                                   const float* __restrict__ gO,
                                                                   float* __restrict__ gI2, int* __restrict__ counts,
                                                                   int* __restrict__ key, int* __restrict__ rank,
                                                                   WarpGeom g, float sx, float sy) {
    const long long hwf = (long long)g.Hf * g.Wf, hwi = (long long)g.Hin * g.Win;
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= (long long)g.B * hwf / 4) return;
    const long long pix = i * 4;
    const int b = (int)(pix / hwf);
    const int rem = (int)(pix - (long long)b * hwf);
    const int y = rem / g.Wf, x = rem - y * g.Wf;
    const float* fu = flow + (long long)b * 2 * hwf + rem;
    const float4 fu4 = ofr::ldg4(fu), fv4 = ofr::ldg4(fu + hwf);
    const float u[4] = {fu4.x, fu4.y, fu4.z, fu4.w}, v[4] = {fv4.x, fv4.y, fv4.z, fv4.w};
    Samp s[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) s[k] = make_samp<MODE>(x + k, y, u[k], v[k], g);
    if (WANT_BIN) {
        int cells[4], ranks[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            cells[k] = cell_of<MODE>(s[k], g);
            ranks[k] = 0;
            if (cells[k] >= 0) ranks[k] = atomicAdd(counts + (long long)b * (g.Hin + 1) * (g.Win + 1) + cells[k], 1);
        }
        *reinterpret_cast<int4*>(key + pix) = make_int4(cells[0], cells[1], cells[2], cells[3]);
        *reinterpret_cast<int4*>(rank + pix) = make_int4(ranks[0], ranks[1], ranks[2], ranks[3]);
    }
    if (!WANT_G2) return;
    float ax[4] = {0.f, 0.f, 0.f, 0.f}, ay[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < g.C; ++c) {
        const float* img = in1 + ((long long)b * g.C + c) * hwi;
        const float4 go4 = ofr::ldg4(gO + ((long long)b * g.C + c) * hwf + rem);
        const float go[4] = {go4.x, go4.y, go4.z, go4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!s[k].live) continue;
            Taps<MODE> t(s[k], g);
            const float tl = t.fetch(img, t.yt, t.xl, g.Win), tr = t.fetch(img, t.yt, t.xr, g.Win);
            const float bl = t.fetch(img, t.yb, t.xl, g.Win), br = t.fetch(img, t.yb, t.xr, g.Win);
            ax[k] = fmaf(go[k], fmaf(s[k].wy1, br - bl, s[k].wy0 * (tr - tl)), ax[k]);
            ay[k] = fmaf(go[k], fmaf(s[k].wx1, br - tr, s[k].wx0 * (bl - tl)), ay[k]);
        }
    }
    float* o = gI2 + (long long)b * 2 * hwf + rem;
    ofr::stcs4(o, make_float4(ax[0] * sx, ax[1] * sx, ax[2] * sx, ax[3] * sx));
    ofr::stcs4(o + hwf, make_float4(ay[0] * sy, ay[1] * sy, ay[2] * sy, ay[3] * sy));
}

// ---- exclusive scan of the cell counts (three small kernels)
__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) >= d) v += n;
    }
    return v;
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const int* __restrict__ counts,
                                                                   int* __restrict__ blocksum, int n) {
    __shared__ int wsum[kScanThreads / 32];
    const int base = blockIdx.x * kScanTile;
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int i = base + k * kScanThreads + threadIdx.x;
        if (i < n) s += counts[i];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < kScanThreads / 32; ++k) t += wsum[k];
        blocksum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) scan_top_kernel(int* __restrict__ blocksum, int nblk) {
    __shared__ int wsum[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nblk; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nblk ? blocksum[i] : 0;
        int incl = warp_incl_scan(v);
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = wsum[threadIdx.x];
            int wi = warp_incl_scan(w);
            wsum[threadIdx.x] = wi - w;  // exclusive prefix of the warp totals
        }
        __syncthreads();
        const int carry = carry_s;
        const int excl = carry + wsum[threadIdx.x >> 5] + incl - v;
        if (i < nblk) blocksum[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int* __restrict__ counts,
                                                                  const int* __restrict__ blocksum,
                                                                  int* __restrict__ offsets, int n) {
    __shared__ int wsum[kScanThreads / 32];
    // each thread owns kScanItems consecutive elements so the scan order is the array order
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems];
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? counts[base + k] : 0;
        s += v[k];
    }
    const int incl = warp_incl_scan(s);
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    int wpre = 0;
    for (int k = 0; k < (int)(threadIdx.x >> 5); ++k) wpre += wsum[k];
    int run = blocksum[blockIdx.x] + wpre + incl - s;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) offsets[base + k] = run;
        run += v[k];
    }
}

// ---- CSR fill.  An entry is a 16-byte record {source index within its image, (x0,y0) packed, wx1, wy1}: everything
// the per-target gather needs about a source, so the gather never touches the flow again.
__device__ __forceinline__ int pack_xy(int x0, int y0) { return ((y0 + 2) << 16) | (x0 + 2); }
__device__ __forceinline__ void unpack_xy(int p, int& x0, int& y0) {
    x0 = (p & 0xffff) - 2;
    y0 = (p >> 16) - 2;
}

template <int MODE>
__global__ void __launch_bounds__(256) resample_fill_kernel(const float* __restrict__ flow,
                                                            const int* __restrict__ key, const int* __restrict__ rank,
                                                            const int* __restrict__ counts,
                                                            const int* __restrict__ offsets, int4* __restrict__ entries,
                                                            int* __restrict__ heavy, int* __restrict__ heavy_n,
                                                            WarpGeom g) {
    const int hwf = g.Hf * g.Wf;
    const long long total = (long long)g.B * hwf;
    const long long pix = blockIdx.x * 256ll + threadIdx.x;
    if (pix >= total) return;
    const int k = key[pix];
    if (k < 0) return;
    const int b = (int)(pix / hwf);
    const int src = (int)(pix - (long long)b * hwf);
    const int cell = b * (g.Hin + 1) * (g.Win + 1) + k;
    const int r = rank[pix];
    const int sy = src / g.Wf, sx = src - sy * g.Wf;
    const float* fl = flow + (long long)b * 2 * hwf + src;
    const Samp s = make_samp<MODE>(sx, sy, __ldg(fl), __ldg(fl + hwf), g);
    entries[offsets[cell] + r] = make_int4(src, pack_xy(s.x0, s.y0), __float_as_int(s.wx1), __float_as_int(s.wy1));
    if (r == 0 && counts[cell] > kHeavy) heavy[atomicAdd(heavy_n, 1)] = cell;
}

// ---- sort the records of every cell with 2..kHeavy entries by source index (ascending): one thread per cell,
// insertion sort on a private copy.  After this pass and the heavy pass every cell is in canonical order, so the
// per-target gather streams its records without any searching.
__global__ void __launch_bounds__(128) resample_sort_small_kernel(const int* __restrict__ counts,
                                                                  const int* __restrict__ offsets,
                                                                  int4* __restrict__ entries, int ncell) {
    const int cell = blockIdx.x * 128 + threadIdx.x;
    if (cell >= ncell) return;
    const int n = counts[cell];
    if (n < 2 || n > kHeavy) return;
    int4* a = entries + offsets[cell];
    if (n == 2) {  // by far the most common crowded cell
        const int4 r0 = a[0], r1 = a[1];
        if (r0.x > r1.x) {
            a[0] = r1;
            a[1] = r0;
        }
        return;
    }
    int4 v[kHeavy];
    for (int i = 0; i < n; ++i) v[i] = a[i];
    for (int i = 1; i < n; ++i) {
        const int4 key = v[i];
        int j = i - 1;
        while (j >= 0 && v[j].x > key.x) {
            v[j + 1] = v[j];
            --j;
        }
        v[j + 1] = key;
    }
    for (int i = 0; i < n; ++i) a[i] = v[i];
}

// ---- sort the entries of crowded cells (ascending).  One CTA per heavy cell, all-ascending bitonic
// network ("flip" first step) so a virtual +inf padding beyond n never has to move.
__global__ void __launch_bounds__(512) resample_sort_heavy_kernel(const int* __restrict__ heavy,
                                                                  const int* __restrict__ heavy_n,
                                                                  const int* __restrict__ counts,
                                                                  const int* __restrict__ offsets,
                                                                  int4* __restrict__ entries) {
    const int nh = *heavy_n;
    for (int h = blockIdx.x; h < nh; h += gridDim.x) {
        const int cell = heavy[h];
        const int n = counts[cell];
        int4* a = entries + offsets[cell];
        int P = 1;
        while (P < n) P <<= 1;
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < n; i += blockDim.x) {
                    const int l = (j == (k >> 1)) ? (i ^ (k - 1)) : (i ^ j);
                    if (l > i && l < n) {
                        const int4 vi = a[i], vl = a[l];
                        if (vi.x > vl.x) {
                            a[i] = vl;
                            a[l] = vi;
                        }
                    }
                }
                __syncthreads();
            }
        }
    }
}

// ---- per-target gather of the adjoint, one CTA per tile of targets.
// A CTA owns kTY x kTX targets of one image and ALL channels.  The records that can reach the tile live in the
// (kTY+1) x (kTX+1) block of cells around it; each row of that block is one contiguous run of the CSR, so the CTA
// copies the runs into shared memory with coalesced 16-byte loads ONCE and every thread then walks its <= 4 cells
// out of shared memory (no dependent global round trips), channel chunk after channel chunk.  Only the gradient
// gathers go to global memory; lanes are consecutive targets, so for smooth flows they coalesce.
// Visiting order per target: cell (ty-1,tx-1), (ty-1,tx), (ty,tx-1), (ty,tx), ascending source index inside a cell
// (the cells were sorted by resample_sort_small/heavy_kernel) -- a function of the data only: bitwise reproducible.
constexpr int kTY = 8, kTX = 32, kTgtThreads = kTY * kTX;
constexpr int kTgtChunk = 16;      // channels accumulated per sweep over a thread's records
constexpr int kRecCap = 1024;      // records staged per tile (16 KB); denser tiles read their records from global
constexpr int kListCap = 8;        // decoded (source, weight) pairs kept per target (16 KB per CTA)

template <int MODE>
__global__ void __launch_bounds__(kTgtThreads, 4) resample_bwd_tgt_kernel(const float* __restrict__ gO,
                                                                       const int* __restrict__ counts,
                                                                       const int* __restrict__ offsets,
                                                                       const int4* __restrict__ entries,
                                                                       float* __restrict__ gI1, WarpGeom g,
                                                                       int tiles_x, int tiles_y, int chunks_per_cta) {
    __shared__ int4 srec[kRecCap];
    __shared__ int2 elist[kListCap][kTgtThreads];
    __shared__ int soff[kTY + 1][kTX + 2];  // CSR offset of every cell of the block (+1 column: end of the run)
    __shared__ int spos[kTY + 2];           // where the run of each cell row starts in srec

    int bid = blockIdx.x;
    const int txi = bid % tiles_x;
    bid /= tiles_x;
    const int tyi = bid % tiles_y;
    const int b = bid / tiles_y;
    const int tx0 = txi * kTX, ty0 = tyi * kTY;
    const int cw = g.Win + 1, ch = g.Hin + 1;
    const long long cell_base = (long long)b * ch * cw;
    const long long hwf = (long long)g.Hf * g.Wf, hwi = (long long)g.Hin * g.Win;

    // biased cell coordinates of the block: rows ty0 .. ty0+kTY, columns tx0 .. tx0+kTX  (cell (ky,kx) is stored at
    // (ky+1, kx+1), and target (ty,tx) reads cells (ty-1..ty, tx-1..tx))
    for (int i = threadIdx.x; i < (kTY + 1) * (kTX + 2); i += kTgtThreads) {
        const int r = i / (kTX + 2), c = i - r * (kTX + 2);
        const int cy = ty0 + r, cx = tx0 + c;
        int v = 0;
        // one past the last cell of a row is the first cell of the next row (or offsets[ncell] = total): same array
        if (cy < ch) v = __ldg(offsets + cell_base + (long long)cy * cw + min(cx, cw));
        soff[r][c] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int at = 0;
        for (int r = 0; r <= kTY; ++r) {
            spos[r] = at;
            const int cy = ty0 + r;
            if (cy < ch) {
                const int cend = min(kTX + 1, cw - tx0);  // first column index after the block (clipped to the row)
                at += soff[r][cend] - soff[r][0];
            }
        }
        spos[kTY + 1] = at;
    }
    __syncthreads();
    const int total = spos[kTY + 1];
    const bool staged = total <= kRecCap;
    if (staged) {
        for (int r = 0; r <= kTY; ++r) {
            const int n = spos[r + 1] - spos[r];
            const int4* src = entries + soff[r][0];
            for (int i = threadIdx.x; i < n; i += kTgtThreads) srec[spos[r] + i] = __ldg(src + i);
        }
    }
    __syncthreads();

    const int lx = threadIdx.x % kTX, ly = threadIdx.x / kTX;
    const int tx = tx0 + lx, ty = ty0 + ly;
    const bool live = tx < g.Win && ty < g.Hin;
    // the four cells of this target: block rows ly, ly+1 ; block columns lx, lx+1
    int start[4], cnt[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = ly + (q >> 1), c = lx + (q & 1);
        const bool in = live && (ty0 + r) < ch && (tx0 + c) < cw;
        const int o = soff[r][c];
        cnt[q] = in ? soff[r][c + 1] - o : 0;
        start[q] = staged ? spos[r] + (o - soff[r][0]) : o;
    }
    const int s1 = cnt[0], s2 = s1 + cnt[1], s3 = s2 + cnt[2], tot = s3 + cnt[3];
    const int trem = ty * g.Win + tx;

    // e-th record of this target in visiting order -> (source index, weight).  The weight is the sum, in tap order
    // TL, TR, BL, BR, of the taps of that source that land on this target (exactly one away from clamped borders).
    auto entry = [&](int e, int& src, float& w) {
        const int q = (e >= s1) + (e >= s2) + (e >= s3);  // no dynamically indexed arrays
        const int base = q == 0 ? start[0] : q == 1 ? start[1] - s1 : q == 2 ? start[2] - s2 : start[3] - s3;
        const int4 rec = staged ? srec[base + e] : __ldg(entries + base + e);
        Samp s;
        unpack_xy(rec.y, s.x0, s.y0);
        s.wx1 = __int_as_float(rec.z);
        s.wy1 = __int_as_float(rec.w);
        s.wx0 = __fsub_rn(1.f, s.wx1);
        s.wy0 = __fsub_rn(1.f, s.wy1);
        const Taps<MODE> t(s, g);
        float acc = 0.f;
        if (t.xl == tx && t.yt == ty) acc = __fadd_rn(acc, __fmul_rn(s.wx0, s.wy0));
        if (t.xr == tx && t.yt == ty) acc = __fadd_rn(acc, __fmul_rn(s.wx1, s.wy0));
        if (t.xl == tx && t.yb == ty) acc = __fadd_rn(acc, __fmul_rn(s.wx0, s.wy1));
        if (t.xr == tx && t.yb == ty) acc = __fadd_rn(acc, __fmul_rn(s.wx1, s.wy1));
        src = rec.x;
        w = acc;
    };
    // decode the first kListCap records once; the channel sweeps below then only do gather + FMA
    const int nlist = min(tot, kListCap);
    for (int e = 0; e < nlist; ++e) {
        int src;
        float w;
        entry(e, src, w);
        elist[e][threadIdx.x] = make_int2(src, __float_as_int(w));
    }
    // (each thread reads back only what it wrote itself: no barrier needed)

    const int cbeg = blockIdx.y * chunks_per_cta * kTgtChunk;
    const int cend = min(g.C, cbeg + chunks_per_cta * kTgtChunk);
    for (int c0 = cbeg; c0 < cend; c0 += kTgtChunk) {
        const int nc = min(kTgtChunk, g.C - c0);
        const float* go = gO + ((long long)b * g.C + c0) * hwf;
        float acc[kTgtChunk];
#pragma unroll
        for (int k = 0; k < kTgtChunk; ++k) acc[k] = 0.f;
        for (int e = 0; e < tot; ++e) {
            int src;
            float w;
            if (e < kListCap) {
                const int2 v = elist[e][threadIdx.x];
                src = v.x;
                w = __int_as_float(v.y);
            } else {
                entry(e, src, w);  // crowded targets (> kListCap records): decode on the fly
            }
            const float* gp = go + src;
            float gv[kTgtChunk];
            if (nc == kTgtChunk) {  // full chunk: no per-channel predicates
#pragma unroll
                for (int k = 0; k < kTgtChunk; ++k) gv[k] = __ldg(gp + (size_t)k * (size_t)hwf);
            } else {
#pragma unroll
                for (int k = 0; k < kTgtChunk; ++k) gv[k] = (k < nc) ? __ldg(gp + (size_t)k * (size_t)hwf) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < kTgtChunk; ++k) acc[k] = fmaf(w, gv[k], acc[k]);
        }
        if (live) {
#pragma unroll
            for (int k = 0; k < kTgtChunk; ++k)
                if (k < nc) gI1[((long long)b * g.C + c0 + k) * hwi + trem] = acc[k];
        }
    }
}

// nearest mode: the flow gets no gradient; the image gradient is still a (deterministic) gather,
// but no reference call site uses it -> not implemented.

struct WsLayout {
    size_t counts, misc, offsets, blocksum, key, rank, entries, heavy, total;
    int ncell, nblk;
    long long npix;
};

WsLayout ws_layout(int B, int Hin, int Win, int Hf, int Wf) {
    WsLayout L;
    L.ncell = B * (Hin + 1) * (Win + 1);
    L.npix = (long long)B * Hf * Wf;
    L.nblk = (L.ncell + 1 + kScanTile - 1) / kScanTile;  // the scan runs over ncell+1 items: offsets[ncell] = total
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = ofr::align_up(o + bytes, 256);
        return at;
    };
    L.counts = take(sizeof(int) * ((size_t)L.ncell + 1));
    L.misc = take(256);
    L.offsets = take(sizeof(int) * ((size_t)L.ncell + 1));
    L.blocksum = take(sizeof(int) * (size_t)L.nblk);
    L.key = take(sizeof(int) * (size_t)L.npix);
    L.rank = take(sizeof(int) * (size_t)L.npix);
    L.entries = take(sizeof(int4) * (size_t)L.npix);
    L.heavy = take(sizeof(int) * (size_t)(L.npix / kHeavy + 1));
    L.total = o;
    return L;
}

bool geom_ok(int B, int C, int Hin, int Win, int Hf, int Wf) {
    if (B <= 0 || C <= 0 || Hin <= 0 || Win <= 0 || Hf <= 0 || Wf <= 0) return false;
    // index arithmetic inside one image is 32-bit
    if ((long long)Hf * Wf >= (1ll << 30) || (long long)(Hin + 1) * (Win + 1) >= (1ll << 30)) return false;
    if ((long long)B * (Hin + 1) * (Win + 1) >= (1ll << 31)) return false;
    if (Hin > 32000 || Win > 32000 || Hf > 32000 || Wf > 32000) return false;  // (x0, y0) are packed in 16 bits each
    return true;
}

WarpGeom make_geom(int B, int C, int Hin, int Win, int Hf, int Wf) {
    WarpGeom g;
    g.B = B, g.C = C, g.Hin = Hin, g.Win = Win, g.Hf = Hf, g.Wf = Wf;
    g.fx = (float)(2.0 / Win);  // models.py:264 builds a FloatTensor from the python floats 2/w, 2/h
    g.fy = (float)(2.0 / Hin);
    return g;
}

template <int MODE>
int launch_fwd(const float* in1, const float* flow, float* out, const WarpGeom& g, cudaStream_t st) {
    const long long npix = (long long)g.B * g.Hf * g.Wf;
    const unsigned gy = (unsigned)((g.C + kFwdChunk - 1) / kFwdChunk);
    const bool vec = (g.Wf % 4 == 0) && ofr::aligned16(flow) && ofr::aligned16(out);
    if (vec) {
        dim3 grid((unsigned)((npix / 4 + kFwdThreads - 1) / kFwdThreads), gy);
        resample_fwd_kernel<MODE, 4><<<grid, kFwdThreads, 0, st>>>(in1, flow, out, g);
    } else {
        dim3 grid((unsigned)((npix + kFwdThreads - 1) / kFwdThreads), gy);
        resample_fwd_kernel<MODE, 1><<<grid, kFwdThreads, 0, st>>>(in1, flow, out, g);
    }
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

template <int MODE, int CG>
int launch_src(const float* in1, const float* flow, const float* gO, float* gI2, int* counts, int* key, int* rank,
               const WarpGeom& g, float sx, float sy, bool want_bin, cudaStream_t st) {
    const long long npix = (long long)g.B * g.Hf * g.Wf;
    dim3 block(32, CG), grid((unsigned)((npix + 31) / 32));
    if (gI2 && want_bin)
        resample_bwd_src_kernel<MODE, CG, true, true><<<grid, block, 0, st>>>(in1, flow, gO, gI2, counts, key, rank, g, sx, sy);
    else if (gI2)
        resample_bwd_src_kernel<MODE, CG, true, false><<<grid, block, 0, st>>>(in1, flow, gO, gI2, counts, key, rank, g, sx, sy);
    else
        resample_bwd_src_kernel<MODE, CG, false, true><<<grid, block, 0, st>>>(in1, flow, gO, gI2, counts, key, rank, g, sx, sy);
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

template <int MODE>
int run_bwd(const float* in1, const float* flow, const float* gO, float* gI1, float* gI2, const WarpGeom& g,
            void* ws, cudaStream_t st) {
    // chain factors d(sample position)/d(flow), multiplied in the order autograd applies them
    float sx = 1.f, sy = 1.f;
    if (MODE == OFR_WARP_STN) {
        sx = (0.5f * (float)(g.Win - 1)) * g.fx;
        sy = (0.5f * (float)(g.Hin - 1)) * g.fy;
    } else if (MODE == OFR_WARP_PWC) {
        sx = (0.5f * (float)g.Win) / (float)max(g.Win - 1, 1) * 2.f;
        sy = (0.5f * (float)g.Hin) / (float)max(g.Hin - 1, 1) * 2.f;
    }
    const bool want_bin = gI1 != nullptr;
    WsLayout L = ws_layout(g.B, g.Hin, g.Win, g.Hf, g.Wf);
    char* w = (char*)ws;
    int* counts = want_bin ? (int*)(w + L.counts) : nullptr;
    int* misc = want_bin ? (int*)(w + L.misc) : nullptr;
    int* offsets = want_bin ? (int*)(w + L.offsets) : nullptr;
    int* blocksum = want_bin ? (int*)(w + L.blocksum) : nullptr;
    int* key = want_bin ? (int*)(w + L.key) : nullptr;
    int* rank = want_bin ? (int*)(w + L.rank) : nullptr;
    int4* entries = want_bin ? (int4*)(w + L.entries) : nullptr;
    int* heavy = want_bin ? (int*)(w + L.heavy) : nullptr;
    if (want_bin) OFR_CUDA(cudaMemsetAsync(w + L.counts, 0, L.offsets - L.counts, st));  // counts + misc

    int rc;
    const bool vec4 = g.C < 8 && (g.Wf % 4 == 0) && ofr::aligned16(flow) && ofr::aligned16(gO) &&
                      (!gI2 || ofr::aligned16(gI2));
    if (vec4) {
        const long long n4 = (long long)g.B * g.Hf * g.Wf / 4;
        const unsigned grid = (unsigned)((n4 + 255) / 256);
        if (gI2 && want_bin)
            resample_bwd_src4_kernel<MODE, true, true><<<grid, 256, 0, st>>>(in1, flow, gO, gI2, counts, key, rank, g, sx, sy);
        else if (gI2)
            resample_bwd_src4_kernel<MODE, true, false><<<grid, 256, 0, st>>>(in1, flow, gO, gI2, counts, key, rank, g, sx, sy);
        else
            resample_bwd_src4_kernel<MODE, false, true><<<grid, 256, 0, st>>>(in1, flow, gO, gI2, counts, key, rank, g, sx, sy);
        rc = ofr::after_launch();
    } else {
        // Channel split across the CTA only as far as needed to fill the machine: this kernel is a chain of dependent
        // memory round trips, so its time is (warps / resident warps) x chain latency -- fewer, longer-running warps
        // with more independent loads each are better than many short ones once the SMs are full.
        const long long pix_warps = ((long long)g.B * g.Hf * g.Wf + 31) / 32;
        const long long want = 6ll * 32 * ofr::sm_count();  // ~6 waves of 32 resident warps per SM
        if (g.C >= 8 && pix_warps * 4 < want)
            rc = launch_src<MODE, 8>(in1, flow, gO, gI2, counts, key, rank, g, sx, sy, want_bin, st);
        else if (g.C >= 4 && pix_warps * 2 < want)
            rc = launch_src<MODE, 4>(in1, flow, gO, gI2, counts, key, rank, g, sx, sy, want_bin, st);
        else if (g.C >= 2 && pix_warps < want)
            rc = launch_src<MODE, 2>(in1, flow, gO, gI2, counts, key, rank, g, sx, sy, want_bin, st);
        else
            rc = launch_src<MODE, 1>(in1, flow, gO, gI2, counts, key, rank, g, sx, sy, want_bin, st);
    }
    if (rc != OFR_OK || !want_bin) return rc;

    scan_reduce_kernel<<<L.nblk, kScanThreads, 0, st>>>(counts, blocksum, L.ncell + 1);
    OFR_LAUNCH_CHECK();
    scan_top_kernel<<<1, 1024, 0, st>>>(blocksum, L.nblk);
    OFR_LAUNCH_CHECK();
    scan_apply_kernel<<<L.nblk, kScanThreads, 0, st>>>(counts, blocksum, offsets, L.ncell + 1);
    OFR_LAUNCH_CHECK();
    resample_fill_kernel<MODE><<<(unsigned)((L.npix + 255) / 256), 256, 0, st>>>(flow, key, rank, counts, offsets,
                                                                                entries, heavy, misc, g);
    OFR_LAUNCH_CHECK();
    resample_sort_small_kernel<<<(unsigned)((L.ncell + 127) / 128), 128, 0, st>>>(counts, offsets, entries, L.ncell);
    OFR_LAUNCH_CHECK();
    resample_sort_heavy_kernel<<<ofr::sm_count(), 512, 0, st>>>(heavy, misc, counts, offsets, entries);
    OFR_LAUNCH_CHECK();
    const int tiles_x = (g.Win + kTX - 1) / kTX, tiles_y = (g.Hin + kTY - 1) / kTY;
    // all channels in one CTA when there are enough tiles to fill the GPU; otherwise split the channel chunks over
    // blockIdx.y (each CTA re-stages the tile's records, which is cheap)
    const int nchunks = (g.C + kTgtChunk - 1) / kTgtChunk;
    const long long ntile = (long long)tiles_x * tiles_y * g.B;
    int split = (int)((4ll * ofr::sm_count() + ntile - 1) / ntile);
    split = split < 1 ? 1 : (split > nchunks ? nchunks : split);
    const int chunks_per_cta = (nchunks + split - 1) / split;
    dim3 grid((unsigned)ntile, (unsigned)((nchunks + chunks_per_cta - 1) / chunks_per_cta));
    resample_bwd_tgt_kernel<MODE><<<grid, kTgtThreads, 0, st>>>(gO, counts, offsets, entries, gI1, g, tiles_x, tiles_y,
                                                                chunks_per_cta);
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

}  // namespace

extern "C" int ofr_resample_fwd(const float* in1, const float* flow, float* out, int B, int C, int Hin, int Win,
                                int Hf, int Wf, int mode, ofr_stream_t stream) {
    if (!in1 || !flow || !out || !geom_ok(B, C, Hin, Win, Hf, Wf)) return OFR_EINVAL;
    if ((mode == OFR_WARP_STN || mode == OFR_WARP_PWC) && (Hin != Hf || Win != Wf)) return OFR_EINVAL;
    const WarpGeom g = make_geom(B, C, Hin, Win, Hf, Wf);
    cudaStream_t st = (cudaStream_t)stream;
    switch (mode) {
        case OFR_WARP_FLOWNET2: return launch_fwd<OFR_WARP_FLOWNET2>(in1, flow, out, g, st);
        case OFR_WARP_STN: return launch_fwd<OFR_WARP_STN>(in1, flow, out, g, st);
        case OFR_WARP_PWC: return launch_fwd<OFR_WARP_PWC>(in1, flow, out, g, st);
        case OFR_WARP_FLOWNET2_NEAREST: return launch_fwd<OFR_WARP_FLOWNET2_NEAREST>(in1, flow, out, g, st);
        default: return OFR_EINVAL;
    }
}

extern "C" size_t ofr_resample_bwd_workspace(int B, int C, int Hin, int Win, int Hf, int Wf, int mode) {
    (void)mode;
    if (!geom_ok(B, C, Hin, Win, Hf, Wf)) return 0;
    return ws_layout(B, Hin, Win, Hf, Wf).total;
}

extern "C" int ofr_resample_bwd(const float* in1, const float* flow, const float* gO, float* gI1, float* gI2, int B,
                                int C, int Hin, int Win, int Hf, int Wf, int mode, void* ws, size_t ws_bytes,
                                ofr_stream_t stream) {
    if (!in1 || !flow || !gO || !geom_ok(B, C, Hin, Win, Hf, Wf)) return OFR_EINVAL;
    if (!gI1 && !gI2) return OFR_OK;
    if ((mode == OFR_WARP_STN || mode == OFR_WARP_PWC) && (Hin != Hf || Win != Wf)) return OFR_EINVAL;
    if (mode == OFR_WARP_FLOWNET2_NEAREST) return OFR_ENOTSUP;
    if (gI1) {
        if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255u) || ws_bytes < ws_layout(B, Hin, Win, Hf, Wf).total)
            return OFR_EWORKSPACE;
    }
    const WarpGeom g = make_geom(B, C, Hin, Win, Hf, Wf);
    cudaStream_t st = (cudaStream_t)stream;
    switch (mode) {
        case OFR_WARP_FLOWNET2: return run_bwd<OFR_WARP_FLOWNET2>(in1, flow, gO, gI1, gI2, g, ws, st);
        case OFR_WARP_STN: return run_bwd<OFR_WARP_STN>(in1, flow, gO, gI1, gI2, g, ws, st);
        case OFR_WARP_PWC: return run_bwd<OFR_WARP_PWC>(in1, flow, gO, gI1, gI2, g, ws, st);
        default: return OFR_EINVAL;
    }
}
