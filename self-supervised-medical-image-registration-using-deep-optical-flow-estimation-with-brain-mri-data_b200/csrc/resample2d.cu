// Bilinear warps: Resample2d (flownet2), opticalFlowReg.stn and PWCDCNet.warp in one kernel family.
//
// Reference semantics:
//   FLOWNET2 : flownet2/models.py:139,151,163,176  sample (x+u, y+v), taps clamped, fractional weights unclamped
//   STN      : models.py:256-268                   g=(p+f)*(2/W,2/H)-1, grid_sample(zeros, align_corners=True)
//   PWC      : PWC/models/PWCNet.py:143-179        g=2(p+f)/(W-1)-1, grid_sample(zeros, align_corners=False),
//                                                  output zeroed where the warped ones-mask < 0.9999
// The coordinate arithmetic is spelled with explicit round-to-nearest intrinsics in the same operation order PyTorch
// uses, so that the float32 sampling positions agree with the reference bit for bit (a 1-ulp difference in a
// normalised coordinate is 1.5e-5 px at W=512).
//
// All kernels index with (tile x, row, image) from the block / thread ids: no integer division per pixel.
//
// Forward  : one thread per 4 consecutive output pixels of a row (128-bit flow loads / output stores when Wf % 4 == 0),
//            the four tap offsets and weights are formed once and reused for every channel of the thread's channel
//            chunk.  Algorithmic bytes 4*B*Hf*Wf*(2C+2).
// Backward : gI2 (flow) is a per-source gather + analytic derivative.
//            gI1 (image) is the adjoint scatter.  Upstream used float atomicAdd (non-deterministic); here it is a
//            per-target gather in a data-defined order (north-star rule 3):
//              1. resample_prep_kernel writes one 16-byte record per source pixel {offset, (x0,y0), wx1, wy1} in source-tile
//                 major order and the bounding box of the footprint cells of every 8 x 32 source tile;
//              2. resample_bwd_fused_kernel: one CTA per 8 x 32 tile of TARGETS.  It finds the source tiles whose
//                 bounding box touches the tile (ascending), bins their records by footprint cell in SHARED memory
//                 (integer atomics only: counts and slots are order independent once each cell is sorted by source offset), and
//                 every target thread then walks its <= 4 cells in a fixed order, gathering gO for 16 channels at a
//                 time.  The same thread also forms the flow gradient of the source pixel at its own position, so gO
//                 and in1 are read once.  When the channels are split over several CTAs (small maps), the CTAs of one
//                 tile form a thread-block cluster and the partial flow gradients are combined through distributed
//                 shared memory in rank order.
//            Floating-point summation order is a function of the data only: bitwise reproducible.  Tiles receiving
//            more records than fit in shared memory (strongly compressive flows) take a slower scan path with the
//            same summation order.
#include <cooperative_groups.h>

#include <climits>

#include "ofr_common.cuh"
#include "warp_math.cuh"

namespace cg = cooperative_groups;

namespace {

using namespace ofr_warp;

// ------------------------------------------------------------------------------------------------ forward
constexpr int kFwdThreads = 256;
constexpr int kFwdChunk = 8;  // channels per blockIdx.z slice
constexpr int kPX = 4;        // pixels per thread, stacked in a column

// blockDim = (bx, by), bx * by = 256, bx <= 32: the lanes of a warp are CONSECUTIVE pixels of 32/bx rows, so every flow
// load, tap gather and output store of a warp covers whole runs of consecutive pixels (1-2 cache lines per request for
// smooth flows; four pixels per thread along x would spread every request over four lines and make the L1 tag stage the
// bottleneck).  A thread owns kPX vertically adjacent pixels: rows y0 .. y0+kPX-1, whose tap rows overlap in L1.
// blockIdx.z = image * nchunks + channel chunk.
// pixels (x, y0 .. y0+kPX-1) of image b, channels [c0, c1)
// Two bodies.  warp_fwd_body: taps outside the image carry offset -1 and are not loaded (predicated gathers) -- best when
// a thread loops over many channels (PWC pyramid: the predicates are formed once, masked pixels load nothing).
// warp_fwd_body_lean: every tap is loaded from a clamped address and out-of-range taps have zero weight -- fewer
// instructions per pixel, best for the one-channel images (stn of the moving image, FlowNet2's Resample2d), which ncu
// shows issue-bound (80 % issue slots), not memory-bound.
template <int MODE>
__device__ __forceinline__ void warp_fwd_body(const float* __restrict__ in1, const float* __restrict__ flow,
                                              float* __restrict__ out, const WarpGeom& g, int x, int y0, int b, int c0,
                                              int c1) {
    const int hwf = g.Hf * g.Wf, hwi = g.Hin * g.Win;
    const int rem = y0 * g.Wf + x;
    bool ok[kPX];
#pragma unroll
    for (int k = 0; k < kPX; ++k) ok[k] = y0 + k < g.Hf;
    {
        const float* fu = flow + (size_t)b * 2 * hwf + rem;
        float u[kPX], v[kPX];
#pragma unroll
        for (int k = 0; k < kPX; ++k) {
            u[k] = ok[k] ? __ldg(fu + k * g.Wf) : 0.f;
            v[k] = ok[k] ? __ldg(fu + hwf + k * g.Wf) : 0.f;
        }
        TapOff t[kPX];
        float w00[kPX], w01[kPX], w10[kPX], w11[kPX];
        bool live[kPX];
#pragma unroll
        for (int k = 0; k < kPX; ++k) {
            const Samp s = make_samp<MODE>(x, y0 + k, u[k], v[k], g);
            live[k] = s.live;
            if (MODE == OFR_WARP_FLOWNET2_NEAREST) {  // nearest tap: floor(x + u + 0.5), clamped
                const float xn = floorf(__fadd_rn(__fadd_rn((float)x, __fmul_rn(u[k], g.fs)), 0.5f));
                const float yn = floorf(__fadd_rn(__fadd_rn((float)(y0 + k), __fmul_rn(v[k], g.fs)), 0.5f));
                const int xi = min(max(safe_floor_int(xn, g.Win), 0), g.Win - 1);
                const int yi = min(max(safe_floor_int(yn, g.Hin), 0), g.Hin - 1);
                t[k].tl = yi * g.Win + xi;
            } else {
                t[k] = tap_offsets<MODE>(s.x0, s.y0, g.Hin, g.Win);
                w00[k] = __fmul_rn(s.wx0, s.wy0);
                w01[k] = __fmul_rn(s.wx1, s.wy0);
                w10[k] = __fmul_rn(s.wx0, s.wy1);
                w11[k] = __fmul_rn(s.wx1, s.wy1);
            }
        }
        const float* img = in1 + ((size_t)b * g.C + c0) * hwi;
        float* o = out + ((size_t)b * g.C + c0) * hwf + rem;
        // (a compile-time trip count of kFwdChunk with 4 channels unrolled spills at 64 registers: 33.5 -> 135 us at level 2)
#pragma unroll 2
        for (int c = c0; c < c1; ++c, img = img + hwi, o = o + hwf) {
            float r[kPX];
#pragma unroll
            for (int k = 0; k < kPX; ++k) {
                if (MODE == OFR_WARP_FLOWNET2_NEAREST) {
                    r[k] = __ldg(img + t[k].tl);
                } else {
                    float acc = __fmul_rn(w00[k], fetch(img, t[k].tl));
                    acc = fmaf(w01[k], fetch(img, t[k].tr), acc);
                    acc = fmaf(w10[k], fetch(img, t[k].bl), acc);
                    acc = fmaf(w11[k], fetch(img, t[k].br), acc);
                    r[k] = live[k] ? acc : 0.f;
                }
            }
#pragma unroll
            for (int k = 0; k < kPX; ++k)
                if (ok[k]) __stcs(o + k * g.Wf, r[k]);
        }
    }
}


// ONE: exactly one channel (stn of the moving image, FlowNet2's Resample2d): no channel loop at all
template <int MODE, bool ONE = false>
__device__ __forceinline__ void warp_fwd_body_lean(const float* __restrict__ in1, const float* __restrict__ flow,
                                              float* __restrict__ out, const WarpGeom& g, int x, int y0, int b, int c0,
                                              int c1) {
    const int hwf = g.Hf * g.Wf, hwi = g.Hin * g.Win;
    // rows past the bottom edge repeat the last row (loads need no predicate); only their stores are dropped
    int ro[kPX];
#pragma unroll
    for (int k = 0; k < kPX; ++k) ro[k] = min(y0 + k, g.Hf - 1) * g.Wf + x;
    const float* fu = opaque_ptr(flow + (size_t)b * 2 * hwf);
    const float* fv = opaque_ptr(fu + hwf);
    float u[kPX], v[kPX];
#pragma unroll
    for (int k = 0; k < kPX; ++k) {
        u[k] = __ldg(fu + ro[k]);
        v[k] = __ldg(fv + ro[k]);
    }
    FwdTaps t[kPX];
#pragma unroll
    for (int k = 0; k < kPX; ++k) {
        const Samp s = make_samp<MODE>(x, y0 + k, u[k], v[k], g);
        if (MODE == OFR_WARP_FLOWNET2_NEAREST) {  // nearest tap: floor(x + u + 0.5), clamped
            const float xn = floorf(__fadd_rn(__fadd_rn((float)x, __fmul_rn(u[k], g.fs)), 0.5f));
            const float yn = floorf(__fadd_rn(__fadd_rn((float)(y0 + k), __fmul_rn(v[k], g.fs)), 0.5f));
            const int xi = min(max(safe_floor_int(xn, g.Win), 0), g.Win - 1);
            const int yi = min(max(safe_floor_int(yn, g.Hin), 0), g.Hin - 1);
            t[k].tl = yi * g.Win + xi;
        } else {
            t[k] = fwd_taps<MODE>(s, g.Hin, g.Win);
        }
    }
    const float* img = opaque_ptr(in1 + ((size_t)b * g.C + c0) * hwi);
    float* o = opaque_ptr(out + ((size_t)b * g.C + c0) * hwf);
#pragma unroll 2
    for (int c = c0; c < (ONE ? c0 + 1 : c1); ++c, img = opaque_ptr(img + hwi), o = opaque_ptr(o + hwf)) {
        float r[kPX];
#pragma unroll
        for (int k = 0; k < kPX; ++k) {
            if (MODE == OFR_WARP_FLOWNET2_NEAREST) {
                r[k] = __ldg(img + t[k].tl);
            } else {
                r[k] = tap_sum(img, t[k]);  // (PWC: the weights of a masked pixel are all 0)
            }
        }
#pragma unroll
        for (int k = 0; k < kPX; ++k)
            if (y0 + k < g.Hf) __stcs(o + ro[k], r[k]);
    }
}


// Cross-CTA L2 prefetch (flow-gradient kernel only).  A thread's work is a chain of dependent DRAM round trips (flow ->
// tap addresses -> image), and with ~1000 resident threads per SM of four pixels each the kernels are bound by that
// latency (ncu: long_scoreboard 5.6 per issue at ~50 % of the HBM peak).  The CTA that handles image b pulls the rows image
// b + pf_z will need -- same (x, y) position, about 1.5 waves of CTAs further in launch order -- into the L2: one
// `prefetch.global.L2` per 128-byte row segment (lane 0 of each warp).  Measured (B = 24, 512 x 512, one channel):
// flow-gradient kernel 53.7 -> 49.8 us; the FORWARD kernel got slower with it (29.1 -> 34.8 us, PWC level 2 33.5 ->
// 40.2 us: its own loads then queue behind the prefetches), so it does not use it.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int MODE, bool LEAN>
__global__ void __launch_bounds__(kFwdThreads, 4) resample_fwd_kernel(const float* __restrict__ in1,
                                                                      const float* __restrict__ flow,
                                                                      float* __restrict__ out, WarpGeom g, int nchunks) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * kPX;
    if (x >= g.Wf || y0 >= g.Hf) return;
    for (int z = blockIdx.z; z < g.B * nchunks; z += gridDim.z) {
        int b = z, chunk = 0;
        if (nchunks > 1) {
            b = z / nchunks;
            chunk = z - b * nchunks;
        }
        const int c0 = chunk * kFwdChunk;
        if (LEAN && g.C == 1)
            warp_fwd_body_lean<MODE, true>(in1, flow, out, g, x, y0, b, 0, 1);
        else if (LEAN)
            warp_fwd_body_lean<MODE>(in1, flow, out, g, x, y0, b, c0, min(c0 + kFwdChunk, g.C));
        else
            warp_fwd_body<MODE>(in1, flow, out, g, x, y0, b, c0, min(c0 + kFwdChunk, g.C));
    }
}

// ------------------------------------------------------------------------------------------------ backward: flow only
// gI2 only (the image needs no gradient: opticalFlowReg.stn on the moving image, FlowNet2's Resample2d on the input
// images).  Same thread mapping as the forward; channels looped in the thread.
// CC: compile-time channel count (1: the one-channel images of stn / FlowNet2's Resample2d -- a run-time channel loop
// around predicated gathers is serialised behind branches, tools/micro/stream_variants.cu); 0 = run-time loop
template <int MODE, int CC = 0>
__device__ __forceinline__ void warp_bwd_flow_body(const float* __restrict__ in1, const float* __restrict__ flow,
                                                   const float* __restrict__ gO, float* __restrict__ gI2,
                                                   const WarpGeom& g, float sx, float sy, int x, int y0, int b) {
    const int hwf = g.Hf * g.Wf, hwi = g.Hin * g.Win;
    const int rem = y0 * g.Wf + x;
    bool ok[kPX];
#pragma unroll
    for (int k = 0; k < kPX; ++k) ok[k] = y0 + k < g.Hf;
    {
        const float* fu = flow + (size_t)b * 2 * hwf + rem;
        float u[kPX], v[kPX];
#pragma unroll
        for (int k = 0; k < kPX; ++k) {
            u[k] = ok[k] ? __ldg(fu + k * g.Wf) : 0.f;
            v[k] = ok[k] ? __ldg(fu + hwf + k * g.Wf) : 0.f;
        }
        TapOff t[kPX];
        Samp s[kPX];
#pragma unroll
        for (int k = 0; k < kPX; ++k) {
            s[k] = make_samp<MODE>(x, y0 + k, u[k], v[k], g);
            t[k] = tap_offsets<MODE>(s[k].x0, s[k].y0, g.Hin, g.Win);
            if (!s[k].live || !ok[k]) t[k].tl = t[k].tr = t[k].bl = t[k].br = -1;
        }
        float ax[kPX], ay[kPX];
#pragma unroll
        for (int k = 0; k < kPX; ++k) ax[k] = ay[k] = 0.f;
        const int nc = CC ? CC : g.C;
        const float* img = in1 + (size_t)b * nc * hwi;
        const float* gp = gO + (size_t)b * nc * hwf + rem;
#pragma unroll
        for (int c = 0; c < nc; ++c, img = img + hwi, gp = gp + hwf) {
#pragma unroll
            for (int k = 0; k < kPX; ++k) {
                const float go = ok[k] ? __ldg(gp + k * g.Wf) : 0.f;
                const float tl = fetch(img, t[k].tl), tr = fetch(img, t[k].tr);
                const float bl = fetch(img, t[k].bl), br = fetch(img, t[k].br);
                // d out / d ix = wy0*(TR-TL) + wy1*(BR-BL);  d out / d iy = wx0*(BL-TL) + wx1*(BR-TR)
                ax[k] = fmaf(go, fmaf(s[k].wy1, br - bl, s[k].wy0 * (tr - tl)), ax[k]);
                ay[k] = fmaf(go, fmaf(s[k].wx1, br - tr, s[k].wx0 * (bl - tl)), ay[k]);
            }
        }
        float* o = gI2 + (size_t)b * 2 * hwf + rem;
#pragma unroll
        for (int k = 0; k < kPX; ++k)
            if (ok[k]) {
                __stcs(o + k * g.Wf, ax[k] * sx * g.fs);
                __stcs(o + hwf + k * g.Wf, ay[k] * sy * g.fs);
            }
    }
}

template <int MODE>
__global__ void __launch_bounds__(kFwdThreads, 3) resample_bwd_flow_kernel(const float* __restrict__ in1,
                                                                           const float* __restrict__ flow,
                                                                           const float* __restrict__ gO,
                                                                           float* __restrict__ gI2, WarpGeom g,
                                                                           float sx, float sy, int pf_z) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * kPX;
    if (x >= g.Wf || y0 >= g.Hf) return;
#ifndef OFR_NO_XPF
    if (pf_z > 0 && blockDim.x == 32 && threadIdx.x == 0 && (int)blockIdx.z + pf_z < g.B && g.C <= 8) {  // (see resample_fwd_kernel)
        const int bp = blockIdx.z + pf_z, hwf = g.Hf * g.Wf, hwi = g.Hin * g.Win;
        const bool same = g.Hin == g.Hf && g.Win == g.Wf;
#pragma unroll
        for (int k = 0; k < kPX; ++k) {
            const int ro = min(y0 + k, g.Hf - 1) * g.Wf + x;
            prefetch_l2(flow + (size_t)bp * 2 * hwf + ro);
            prefetch_l2(flow + (size_t)bp * 2 * hwf + hwf + ro);
            for (int c = 0; c < g.C; ++c) {
                prefetch_l2(gO + ((size_t)bp * g.C + c) * hwf + ro);
                if (same) prefetch_l2(in1 + ((size_t)bp * g.C + c) * hwi + ro);
            }
        }
    }
#endif
    if (g.C == 1) {
        for (int b = blockIdx.z; b < g.B; b += gridDim.z) warp_bwd_flow_body<MODE, 1>(in1, flow, gO, gI2, g, sx, sy, x, y0, b);
    } else {
        for (int b = blockIdx.z; b < g.B; b += gridDim.z) warp_bwd_flow_body<MODE, 0>(in1, flow, gO, gI2, g, sx, sy, x, y0, b);
    }
}

// ------------------------------------------------------------------------------------------------ multi-scale stn
// The per-scale stn warps of opticalFlowReg.forward (models.py:277-281: one stn(flow_i, frame) per pyramid scale) in ONE
// launch each way: most scales are a few thousand pixels, where a launch per scale is pure latency.  CTA = 32 x 32
// pixels of one image of one scale; the per-thread work is the single-scale kernels' bodies.
struct StnScale {
    const float* frame;
    const float* flow;
    float* out;
    const float* g_out;
    float* g_flow;
    WarpGeom g;
    float sx, sy;
    int tiles_x, tiles_y, cta_begin;
};
struct StnParams {
    int n, cta_end;
    StnScale sc[OFR_STN_MAX_SCALES];
};

template <bool BACKWARD>
__global__ void __launch_bounds__(kFwdThreads, 3) stn_multi_kernel(const __grid_constant__ StnParams p) {
    int s = 0;
    while (s + 1 < p.n && (int)blockIdx.x >= p.sc[s + 1].cta_begin) ++s;
    const StnScale& sc = p.sc[s];
    int t = blockIdx.x - sc.cta_begin;
    const int per = sc.tiles_x * sc.tiles_y;
    const int b = t / per;
    t -= b * per;
    const int ty = t / sc.tiles_x, tx = t - ty * sc.tiles_x;
    const int x = tx * 32 + (threadIdx.x & 31);
    const int y0 = (ty * (kFwdThreads / 32) + (threadIdx.x >> 5)) * kPX;
    if (x >= sc.g.Wf || y0 >= sc.g.Hf) return;
    if (BACKWARD)
        if (sc.g.C == 1)
            warp_bwd_flow_body<OFR_WARP_STN, 1>(sc.frame, sc.flow, sc.g_out, sc.g_flow, sc.g, sc.sx, sc.sy, x, y0, b);
        else
            warp_bwd_flow_body<OFR_WARP_STN, 0>(sc.frame, sc.flow, sc.g_out, sc.g_flow, sc.g, sc.sx, sc.sy, x, y0, b);
    else
        if (sc.g.C == 1)
            warp_fwd_body_lean<OFR_WARP_STN, true>(sc.frame, sc.flow, sc.out, sc.g, x, y0, b, 0, 1);
        else
            warp_fwd_body_lean<OFR_WARP_STN>(sc.frame, sc.flow, sc.out, sc.g, x, y0, b, 0, sc.g.C);
}

// ------------------------------------------------------------------------------------------------ backward: image (+ flow)
constexpr int kTX = 32, kTY = 8, kTileThreads = kTX * kTY;  // tile of sources (prep) / targets (gather)
constexpr int kCW = kTX + 1, kCH = kTY + 1, kCells = kCW * kCH;  // footprint cells that can reach a target tile
constexpr int kChunk = 16;         // channels accumulated per sweep over a thread's records
constexpr int kRecCap = 1024;      // records binned per target tile in shared memory (16 KB)
constexpr int kListCap = 8;        // decoded (source, weight) pairs kept per target (16 KB per CTA)
constexpr int kMaxSplit = 8;       // channel split = cluster size

__device__ __forceinline__ int pack_xy(int x0, int y0) { return ((y0 + 2) << 16) | (x0 + 2); }
__device__ __forceinline__ void unpack_xy(int p, int& x0, int& y0) {
    x0 = (p & 0xffff) - 2;
    y0 = (p >> 16) - 2;
}

// Footprint cell of a record = top-left corner of its 2x2 footprint, kx in [-1, Win-1], ky in [-1, Hin-1].  With
// clamped taps (FLOWNET2) positions outside the image still land on the border row / column.
template <int MODE>
__device__ __forceinline__ void cell_of(int x0, int y0, const WarpGeom& g, int& kx, int& ky) {
    if (MODE == OFR_WARP_FLOWNET2) {
        kx = min(max(x0, -1), g.Win - 1);
        ky = min(max(y0, -1), g.Hin - 1);
    } else {
        kx = x0;
        ky = y0;
    }
}

// One record per source pixel, stored source-tile major: recs[(image, tile)][256].  {source offset inside a plane,
// packed (x0,y0) or -1 when the source reaches no target, wx1, wy1}.
template <int MODE>
__global__ void __launch_bounds__(kTileThreads) resample_prep_kernel(const float* __restrict__ flow,
                                                                     int4* __restrict__ recs, int4* __restrict__ bbox,
                                                                     WarpGeom g) {
    __shared__ int red[4][kTileThreads / 32];
    const int tid = threadIdx.x, lx = tid & 31, ly = tid >> 5;
    const int sx = blockIdx.x * kTX + lx, sy = blockIdx.y * kTY + ly, b = blockIdx.z;
    const int tile = blockIdx.y * gridDim.x + blockIdx.x, nst = gridDim.x * gridDim.y;
    int4 rec = make_int4(sy * g.Wf + sx, -1, 0, 0);
    int kx0 = INT_MAX, kx1 = INT_MIN, ky0 = INT_MAX, ky1 = INT_MIN;
    if (sx < g.Wf && sy < g.Hf) {
        const int hwf = g.Hf * g.Wf;
        const float* fl = flow + (size_t)b * 2 * hwf + sy * g.Wf + sx;
        const Samp s = make_samp<MODE>(sx, sy, __ldg(fl), __ldg(fl + hwf), g);
        bool reach = s.live;
        if (MODE != OFR_WARP_FLOWNET2)
            reach = reach && s.x0 >= -1 && s.x0 <= g.Win - 1 && s.y0 >= -1 && s.y0 <= g.Hin - 1;
        if (reach) {
            rec.y = pack_xy(s.x0, s.y0);
            rec.z = __float_as_int(s.wx1);
            rec.w = __float_as_int(s.wy1);
            int kx, ky;
            cell_of<MODE>(s.x0, s.y0, g, kx, ky);
            kx0 = kx1 = kx;
            ky0 = ky1 = ky;
        }
    }
    recs[((size_t)b * nst + tile) * kTileThreads + tid] = rec;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        kx0 = min(kx0, __shfl_xor_sync(0xffffffffu, kx0, d));
        kx1 = max(kx1, __shfl_xor_sync(0xffffffffu, kx1, d));
        ky0 = min(ky0, __shfl_xor_sync(0xffffffffu, ky0, d));
        ky1 = max(ky1, __shfl_xor_sync(0xffffffffu, ky1, d));
    }
    if (lx == 0) red[0][ly] = kx0, red[1][ly] = kx1, red[2][ly] = ky0, red[3][ly] = ky1;
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 1; k < kTileThreads / 32; ++k) {
            kx0 = min(kx0, red[0][k]);
            kx1 = max(kx1, red[1][k]);
            ky0 = min(ky0, red[2][k]);
            ky1 = max(ky1, red[3][k]);
        }
        bbox[(size_t)b * nst + tile] = make_int4(kx0, kx1, ky0, ky1);  // empty tile: kx0 > kx1
    }
}

// blockIdx = (target tile x, target tile y, image * csplit + split); launched with cluster dims (1, 1, csplit).
// HW != 0: the plane size Hf*Wf == Hin*Win is this compile-time constant, so the channel offsets of the 16-channel
// gathers are instruction immediates (with a run-time stride a third of the kernel's issue slots were 64-bit address
// arithmetic).  Instantiated for the power-of-two maps of the PWC-Net pyramid.
#ifndef OFR_WARP_TRACE
#define OFR_WARP_TRACE 0   // 1: clock64() marks of thread 0 of one CTA of resample_bwd_fused_kernel, read back with
#endif                     //    ofr_debug_warp_trace (tools/diag/warp_bwd_trace.py; profiles/r02_warp_bwd_study.txt)
#if OFR_WARP_TRACE
__device__ long long ofr_warp_trace_buf[16];
#define OFR_WT(i)                                                                                                  \
    do {                                                                                                           \
        if (threadIdx.x == 0 && blockIdx.x == 1 && blockIdx.y == 7 && blockIdx.z == 5) ofr_warp_trace_buf[i] = clock64(); \
    } while (0)
#else
#define OFR_WT(i)
#endif

template <int MODE, bool WANT1, bool WANT2, int HW>
__global__ void __launch_bounds__(kTileThreads, 3) resample_bwd_fused_kernel(
    const float* __restrict__ in1, const float* __restrict__ gO, const int4* __restrict__ recs,
    const int4* __restrict__ bbox, float* __restrict__ gI1, float* __restrict__ gI2, WarpGeom g, int stx, int sty,
    int csplit, int chunks_per_cta, float gsx, float gsy) {
    __shared__ int4 srec[kRecCap];
    __shared__ int2 elist[kListCap][kTileThreads];
    __shared__ int ccount[kCells + 1], coff[kCells + 1];
    __shared__ int tlist[kTileThreads];
    __shared__ int warp_n[kTileThreads / 32];
    __shared__ float part[2][kTileThreads];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int lx = lane, ly = wid;
    const int b = blockIdx.z / csplit, split = blockIdx.z - b * csplit;
    const int tx0 = blockIdx.x * kTX, ty0 = blockIdx.y * kTY;
    const int tx = tx0 + lx, ty = ty0 + ly;
    const bool live = tx < g.Win && ty < g.Hin;
    const int hwf = HW ? HW : g.Hf * g.Wf, hwi = HW ? HW : g.Hin * g.Win;
    const int nst = stx * sty;
    const int cxmin = tx0 - 1, cymin = ty0 - 1;  // cell (ky, kx) is binned at [ky - cymin][kx - cxmin]
    const int4* recs_b = recs + (size_t)b * nst * kTileThreads;
    const int4* bbox_b = bbox + (size_t)b * nst;

    // Visit, in ascending tile order, every source tile whose footprint cells touch this target tile; `body(tile, rec)`
    // is called by all threads with the record of "their" source of that tile.
    auto for_each_source = [&](auto&& body) {
        for (int base = 0; base < nst; base += kTileThreads) {
            const int i = base + tid;
            bool hit = false;
            if (i < nst) {
                const int4 bb = __ldg(bbox_b + i);
                hit = bb.x <= cxmin + kTX && bb.y >= cxmin && bb.z <= cymin + kTY && bb.w >= cymin;
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) warp_n[wid] = __popc(m);
            __syncthreads();
            int pre = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < kTileThreads / 32; ++w) {
                const int n = warp_n[w];
                pre += (w < wid) ? n : 0;
                tot += n;
            }
            if (hit) tlist[pre + __popc(m & ((1u << lane) - 1u))] = i;
            __syncthreads();
            // four tiles' records in flight at once: the loop is a chain of L2 round trips otherwise
            for (int j = 0; j < tot; j += 4) {
                int t[4];
                int4 r[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    t[k] = tlist[min(j + k, tot - 1)];
                    r[k] = __ldg(recs_b + (size_t)t[k] * kTileThreads + tid);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (j + k < tot) body(t[k], r[k]);
            }
            __syncthreads();
        }
    };
    // bin index of a record inside this tile's cell block, or -1
    auto bin_of = [&](const int4& r) {
        if (r.y < 0) return -1;
        int x0, y0, kx, ky;
        unpack_xy(r.y, x0, y0);
        cell_of<MODE>(x0, y0, g, kx, ky);
        const int cx = kx - cxmin, cy = ky - cymin;
        return ((unsigned)cx < (unsigned)kCW && (unsigned)cy < (unsigned)kCH) ? cy * kCW + cx : -1;
    };
    // weight with which the source of record r lands on this thread's target: sum, in tap order TL, TR, BL, BR, of the
    // taps that hit (exactly one away from clamped borders)
    auto weight_of = [&](const int4& r) {
        int x0, y0;
        unpack_xy(r.y, x0, y0);
        const float wx1 = __int_as_float(r.z), wy1 = __int_as_float(r.w);
        const float wx0 = __fsub_rn(1.f, wx1), wy0 = __fsub_rn(1.f, wy1);
        const int xl = tap_index<MODE>(x0, 0, g.Win), xr = tap_index<MODE>(x0, 1, g.Win);
        const int yt = tap_index<MODE>(y0, 0, g.Hin), yb = tap_index<MODE>(y0, 1, g.Hin);
        float acc = 0.f;
        if (xl == tx && yt == ty) acc = __fadd_rn(acc, __fmul_rn(wx0, wy0));
        if (xr == tx && yt == ty) acc = __fadd_rn(acc, __fmul_rn(wx1, wy0));
        if (xl == tx && yb == ty) acc = __fadd_rn(acc, __fmul_rn(wx0, wy1));
        if (xr == tx && yb == ty) acc = __fadd_rn(acc, __fmul_rn(wx1, wy1));
        return acc;
    };

    // Pull the tile's own pixels of one 16-channel chunk of gO (and of in1 for the flow gradient) into the L1 ahead of
    // use: for smooth flows almost every gather of the chunk then hits L1 instead of paying an L2 round trip per entry.
    // The first chunk is prefetched here, behind the binning prologue; chunk c+1 at the start of chunk c.
    const int cbeg = split * chunks_per_cta * kChunk;
    const int cend = min(g.C, cbeg + chunks_per_cta * kChunk);
    auto prefetch_chunk = [&](int c0) {
        if (!live || c0 >= cend) return;
        const int nc = min(kChunk, g.C - c0);
        const float* pg = gO + ((size_t)b * g.C + c0) * hwf + ty * g.Wf + tx;
        const float* pi = in1 + ((size_t)b * g.C + c0) * hwi + ty * g.Win + tx;
        if (lx % 8 == 0) {  // one request per 32-byte sector
#pragma unroll
            for (int k = 0; k < kChunk; ++k)
                if (k < nc) {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(pg + k * hwf));
                    if (WANT2) asm volatile("prefetch.global.L1 [%0];" ::"l"(pi + k * hwi));
                }
        }
    };
    OFR_WT(0);
    prefetch_chunk(cbeg);

    int total = 0;
    int start[4] = {0, 0, 0, 0}, cnt[4] = {0, 0, 0, 0}, mycell[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) mycell[q] = (ly + (q >> 1)) * kCW + lx + (q & 1);
    if (WANT1) {
        for (int i = tid; i <= kCells; i += kTileThreads) ccount[i] = 0;
        __syncthreads();
        for_each_source([&](int, const int4& r) {
            const int bin = bin_of(r);
            if (bin >= 0) atomicAdd(&ccount[bin], 1);
        });
        OFR_WT(1);
        // exclusive scan of the kCells counts (two cells per thread)
        {
            const int i0 = 2 * tid, i1 = 2 * tid + 1;
            const int v0 = i0 < kCells ? ccount[i0] : 0, v1 = i1 < kCells ? ccount[i1] : 0;
            int incl = v0 + v1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += n;
            }
            if (lane == 31) warp_n[wid] = incl;
            __syncthreads();
            int pre = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < kTileThreads / 32; ++w) {
                const int n = warp_n[w];
                pre += (w < wid) ? n : 0;
                tot += n;
            }
            const int excl = pre + incl - v0 - v1;
            if (i0 < kCells) coff[i0] = excl;
            if (i1 < kCells) coff[i1] = excl + v0;
            if (tid == 0) coff[kCells] = tot;
            total = tot;
            __syncthreads();
            if (i0 < kCells) ccount[i0] = 0;  // reused as the fill cursor
            if (i1 < kCells) ccount[i1] = 0;
            __syncthreads();
        }
    }
    OFR_WT(2);
    const bool staged = WANT1 && total <= kRecCap;
    int tot_e = 0, s1 = 0, s2 = 0, s3 = 0;
    if (staged) {
        for_each_source([&](int, const int4& r) {
            const int bin = bin_of(r);
            if (bin >= 0) srec[coff[bin] + atomicAdd(&ccount[bin], 1)] = r;
        });
        OFR_WT(3);
        // canonical order inside every cell: ascending source offset (one thread per cell, insertion sort in shared memory)
        for (int cell = tid; cell < kCells; cell += kTileThreads) {
            const int o = coff[cell], n = coff[cell + 1] - o;
            for (int i = 1; i < n; ++i) {
                const int4 key = srec[o + i];
                int j = i - 1;
                while (j >= 0 && srec[o + j].x > key.x) {
                    srec[o + j + 1] = srec[o + j];
                    --j;
                }
                srec[o + j + 1] = key;
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            start[q] = coff[mycell[q]];
            cnt[q] = live ? coff[mycell[q] + 1] - start[q] : 0;
        }
        s1 = cnt[0], s2 = s1 + cnt[1], s3 = s2 + cnt[2], tot_e = s3 + cnt[3];
    }
    // e-th record of this target in visiting order (cells TL, TR, BL, BR; ascending source offset inside a cell)
    auto entry = [&](int e, int& src, float& w) {
        const int q = (e >= s1) + (e >= s2) + (e >= s3);  // no dynamically indexed arrays
        const int base = q == 0 ? start[0] : q == 1 ? start[1] - s1 : q == 2 ? start[2] - s2 : start[3] - s3;
        const int4 r = srec[base + e];
        src = r.x;
        w = weight_of(r);
    };
    OFR_WT(4);
    if (staged) {  // decode the first kListCap records once; the channel sweeps then only gather + FMA
        const int nlist = min(tot_e, kListCap);
        for (int e = 0; e < nlist; ++e) {
            int src;
            float w;
            entry(e, src, w);
            elist[e][tid] = make_int2(src, __float_as_int(w));
        }
    }

    // the source pixel at this thread's own position (flow gradient)
    TapOff ot = {-1, -1, -1, -1};
    float owx0 = 0.f, owx1 = 0.f, owy0 = 0.f, owy1 = 0.f;
    const int orem = ty * g.Wf + tx;
    if (WANT2 && live) {
        const int4 r = __ldg(recs_b + (size_t)(blockIdx.y * stx + blockIdx.x) * kTileThreads + tid);
        if (r.y >= 0) {
            int x0, y0;
            unpack_xy(r.y, x0, y0);
            ot = tap_offsets<MODE>(x0, y0, g.Hin, g.Win);
            owx1 = __int_as_float(r.z), owy1 = __int_as_float(r.w);
            owx0 = __fsub_rn(1.f, owx1), owy0 = __fsub_rn(1.f, owy1);
        }
    }
    float ax = 0.f, ay = 0.f;
    OFR_WT(5);
    int wt = 6;  // trace slot of the next mark (unused without OFR_WARP_TRACE)
    (void)wt;

    for (int c0 = cbeg; c0 < cend; c0 += kChunk) {
        const int nc = min(kChunk, g.C - c0);
        const float* go = gO + ((size_t)b * g.C + c0) * hwf;
        prefetch_chunk(c0 + kChunk);
        if (WANT1) {
            float acc[kChunk];
#pragma unroll
            for (int k = 0; k < kChunk; ++k) acc[k] = 0.f;
            auto gather = [&](int src, float w) {
                const float* gp = go + src;  // bumped plane by plane: one IMAD.WIDE per load
                float gv[kChunk];
                if (nc == kChunk) {  // full chunk: no per-channel predicates
#pragma unroll
                    for (int k = 0; k < kChunk; ++k) gv[k] = __ldg(gp + k * hwf);
                } else {
#pragma unroll
                    for (int k = 0; k < kChunk; ++k) gv[k] = (k < nc) ? __ldg(gp + k * hwf) : 0.f;
                }
#pragma unroll
                for (int k = 0; k < kChunk; ++k) acc[k] = fmaf(w, gv[k], acc[k]);
            };
            if (staged) {
                for (int e = 0; e < tot_e; ++e) {
                    int src;
                    float w;
                    if (e < kListCap) {
                        const int2 v = elist[e][tid];
                        src = v.x;
                        w = __int_as_float(v.y);
                    } else {
                        entry(e, src, w);  // crowded targets (> kListCap records): decode on the fly
                    }
                    gather(src, w);
                }
            } else {
                // too many records for shared memory: scan the contributing source tiles, once per cell of the target,
                // in the same order the staged path visits them
                for (int q = 0; q < 4; ++q) {
                    const int want = live ? mycell[q] : -2;
                    for_each_source([&](int, const int4& r) {
                        __syncthreads();  // previous tile fully consumed
                        srec[tid] = r;
                        __syncthreads();
                        for (int j = 0; j < kTileThreads; ++j) {
                            const int4 rj = srec[j];
                            if (bin_of(rj) == want) gather(rj.x, weight_of(rj));
                        }
                    });
                }
            }
            if (live) {
                float* o = gI1 + ((size_t)b * g.C + c0) * hwi + ty * g.Win + tx;
#pragma unroll
                for (int k = 0; k < kChunk; ++k)
                    if (k < nc) o[k * hwi] = acc[k];
            }
        }
        OFR_WT(wt);
        ++wt;
        if (WANT2 && live) {
            // four tap pointers and the own-pixel pointer, bumped plane by plane (invalid taps never dereferenced)
            const float* img = in1 + ((size_t)b * g.C + c0) * hwi;
            const float *ptl = img + max(ot.tl, 0), *ptr = img + max(ot.tr, 0);
            const float *pbl = img + max(ot.bl, 0), *pbr = img + max(ot.br, 0);
            const float* pg = go + orem;
#pragma unroll 4
            for (int k = 0; k < nc; ++k) {
                const float gv = __ldg(pg + k * hwf);
                const float tl = ot.tl >= 0 ? __ldg(ptl + k * hwi) : 0.f, tr = ot.tr >= 0 ? __ldg(ptr + k * hwi) : 0.f;
                const float bl = ot.bl >= 0 ? __ldg(pbl + k * hwi) : 0.f, br = ot.br >= 0 ? __ldg(pbr + k * hwi) : 0.f;
                ax = fmaf(gv, fmaf(owy1, br - bl, owy0 * (tr - tl)), ax);
                ay = fmaf(gv, fmaf(owx1, br - tr, owx0 * (bl - tl)), ay);
            }
        }
        OFR_WT(wt);
        ++wt;
    }
    if (WANT2) {
        if (csplit == 1) {
            if (live) {
                float* o = gI2 + (size_t)b * 2 * hwf + orem;
                o[0] = ax * gsx * g.fs;
                o[hwf] = ay * gsy * g.fs;
            }
        } else {
            // combine the channel splits of this tile through distributed shared memory, in rank order
            cg::cluster_group cluster = cg::this_cluster();
            part[0][tid] = ax;
            part[1][tid] = ay;
            cluster.sync();
            if (split == 0 && live) {
                float sxv = 0.f, syv = 0.f;
                for (int r = 0; r < csplit; ++r) {
                    const float* p = cluster.map_shared_rank(&part[0][0], r);
                    sxv += p[tid];
                    syv += p[kTileThreads + tid];
                }
                float* o = gI2 + (size_t)b * 2 * hwf + orem;
                o[0] = sxv * gsx * g.fs;
                o[hwf] = syv * gsy * g.fs;
            }
            cluster.sync();  // remote shared memory must outlive the reads
        }
    }
}

struct WsLayout {
    size_t recs, bbox, total;
    int stx, sty;
};

WsLayout ws_layout(int B, int Hf, int Wf) {
    WsLayout L;
    L.stx = (Wf + kTX - 1) / kTX;
    L.sty = (Hf + kTY - 1) / kTY;
    const size_t nst = (size_t)L.stx * L.sty;
    L.recs = 0;
    L.bbox = ofr::align_up(sizeof(int4) * (size_t)B * nst * kTileThreads, 256);
    L.total = L.bbox + ofr::align_up(sizeof(int4) * (size_t)B * nst, 256);
    return L;
}

bool geom_ok(int B, int C, int Hin, int Win, int Hf, int Wf) {
    if (B <= 0 || C <= 0 || Hin <= 0 || Win <= 0 || Hf <= 0 || Wf <= 0) return false;
    // index arithmetic inside one image plane is 32-bit
    if ((long long)Hf * Wf >= (1ll << 30) || (long long)Hin * Win >= (1ll << 30)) return false;
    if (Hin > 32000 || Win > 32000 || Hf > 32000 || Wf > 32000) return false;  // (x0, y0) are packed in 16 bits each
    if (B > 65535) return false;
    return true;
}

WarpGeom make_geom(int B, int C, int Hin, int Win, int Hf, int Wf, float fs) {
    WarpGeom g;
    g.fs = fs;
    g.B = B, g.C = C, g.Hin = Hin, g.Win = Win, g.Hf = Hf, g.Wf = Wf;
    g.fx = (float)(2.0 / Win);  // models.py:264 builds a FloatTensor from the python floats 2/w, 2/h
    g.fy = (float)(2.0 / Hin);
    return g;
}

// (bx, by) with bx * by = 256 threads: bx = the smallest power of two covering a row, at most 32 (one warp = whole runs
// of consecutive pixels); a CTA covers by * kPX rows
// slices ahead (in blockIdx.z) whose rows a CTA pulls into the L2: about 1.5 waves of resident CTAs further in launch order
int prefetch_distance(const dim3& grid, int ctas_per_sm) {
    const long long per_z = (long long)grid.x * grid.y;
    const long long ahead = (3ll * ofr::sm_count() * ctas_per_sm / 2 + per_z - 1) / per_z;
    return (int)(ahead < 1 ? 1 : ahead);
}

void block_shape(int W, int H, dim3& block, dim3& grid, unsigned gz) {
    int bx = 8;
    while (bx < W && bx < 32) bx <<= 1;
    int by = kFwdThreads / bx;
    const int need = (H + kPX - 1) / kPX;  // small maps: no thread rows beyond the image
    if (by > need) by = need;
    block = dim3(bx, by);
    grid = dim3((W + bx - 1) / bx, (H + by * kPX - 1) / (by * kPX), gz);
}

template <int MODE>
int launch_fwd(const float* in1, const float* flow, float* out, const WarpGeom& g, cudaStream_t st) {
    const int nchunks = (g.C + kFwdChunk - 1) / kFwdChunk;
    const long long gz = (long long)g.B * nchunks;
    dim3 block, grid;
    block_shape(g.Wf, g.Hf, block, grid, (unsigned)(gz < 65535 ? gz : 65535));
    if (g.C <= 2)
        resample_fwd_kernel<MODE, true><<<grid, block, 0, st>>>(in1, flow, out, g, nchunks);
    else
        resample_fwd_kernel<MODE, false><<<grid, block, 0, st>>>(in1, flow, out, g, nchunks);
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

template <int MODE>
int launch_flow_only(const float* in1, const float* flow, const float* gO, float* gI2, const WarpGeom& g, float sx,
                     float sy, cudaStream_t st) {
    dim3 block, grid;
    block_shape(g.Wf, g.Hf, block, grid, (unsigned)g.B);
    resample_bwd_flow_kernel<MODE><<<grid, block, 0, st>>>(in1, flow, gO, gI2, g, sx, sy, prefetch_distance(grid, 3));
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

template <int MODE, bool WANT1, bool WANT2, int HW = 0>
int launch_fused(const float* in1, const float* gO, const int4* recs, const int4* bbox, float* gI1, float* gI2,
                 const WarpGeom& g, const WsLayout& L, float sx, float sy, cudaStream_t st) {
    const int ttx = (g.Win + kTX - 1) / kTX, tty = (g.Hin + kTY - 1) / kTY;
    const int nchunks = (g.C + kChunk - 1) / kChunk;
    // split the channel chunks over a cluster only as far as needed to give every SM a few CTAs
    const long long ntile = (long long)ttx * tty * g.B;
    // (a divisor of the chunk count, so no rank of the cluster idles: 6 chunks at 32x32 split 6 ways, not 4)
    int csplit = 1;
    for (int cs = 1; cs <= kMaxSplit && cs <= nchunks; ++cs) {
        if (nchunks % cs) continue;
        csplit = cs;
        if (ntile * cs >= 3ll * ofr::sm_count()) break;
    }
    const int chunks_per_cta = nchunks / csplit;
    if ((long long)g.B * csplit > 65535 || tty > 65535) return OFR_ENOTSUP;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ttx, tty, g.B * csplit);
    cfg.blockDim = dim3(kTileThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = csplit;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    OFR_CUDA(cudaLaunchKernelEx(&cfg, resample_bwd_fused_kernel<MODE, WANT1, WANT2, HW>, in1, gO, recs, bbox, gI1, gI2, g,
                                L.stx, L.sty, csplit, chunks_per_cta, sx, sy));
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
    if (!gI1) return launch_flow_only<MODE>(in1, flow, gO, gI2, g, sx, sy, st);

    const WsLayout L = ws_layout(g.B, g.Hf, g.Wf);
    int4* recs = (int4*)((char*)ws + L.recs);
    int4* bbox = (int4*)((char*)ws + L.bbox);
    resample_prep_kernel<MODE><<<dim3(L.stx, L.sty, g.B), kTileThreads, 0, st>>>(flow, recs, bbox, g);
    OFR_LAUNCH_CHECK();
    // the flow gradient rides along when source and target grids coincide (always for stn / PWC)
    const bool same = g.Hf == g.Hin && g.Wf == g.Win;
    if (gI2 && same) {
        if (MODE == OFR_WARP_PWC) {  // the pyramid levels of PWC-Net: plane size known at compile time
            switch (g.Hf * g.Wf) {
                case 128 * 128: return launch_fused<MODE, true, true, 128 * 128>(in1, gO, recs, bbox, gI1, gI2, g, L, sx, sy, st);
                case 64 * 64: return launch_fused<MODE, true, true, 64 * 64>(in1, gO, recs, bbox, gI1, gI2, g, L, sx, sy, st);
                case 32 * 32: return launch_fused<MODE, true, true, 32 * 32>(in1, gO, recs, bbox, gI1, gI2, g, L, sx, sy, st);
                case 16 * 16: return launch_fused<MODE, true, true, 16 * 16>(in1, gO, recs, bbox, gI1, gI2, g, L, sx, sy, st);
                case 8 * 8: return launch_fused<MODE, true, true, 8 * 8>(in1, gO, recs, bbox, gI1, gI2, g, L, sx, sy, st);
                default: break;
            }
        }
        return launch_fused<MODE, true, true>(in1, gO, recs, bbox, gI1, gI2, g, L, sx, sy, st);
    }
    int rc = launch_fused<MODE, true, false>(in1, gO, recs, bbox, gI1, gI2, g, L, sx, sy, st);
    if (rc == OFR_OK && gI2) rc = launch_flow_only<MODE>(in1, flow, gO, gI2, g, sx, sy, st);
    return rc;
}

}  // namespace

extern "C" int ofr_resample_fwd(const float* in1, const float* flow, float* out, int B, int C, int Hin, int Win,
                                int Hf, int Wf, int mode, ofr_stream_t stream) {
    return ofr_resample_fwd_scaled(in1, flow, out, B, C, Hin, Win, Hf, Wf, mode, 1.0f, stream);
}

extern "C" int ofr_resample_fwd_scaled(const float* in1, const float* flow, float* out, int B, int C, int Hin, int Win,
                                       int Hf, int Wf, int mode, float flow_scale, ofr_stream_t stream) {
    if (!in1 || !flow || !out || !geom_ok(B, C, Hin, Win, Hf, Wf)) return OFR_EINVAL;
    if ((mode == OFR_WARP_STN || mode == OFR_WARP_PWC) && (Hin != Hf || Win != Wf)) return OFR_EINVAL;
    const WarpGeom g = make_geom(B, C, Hin, Win, Hf, Wf, flow_scale);
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
    return ws_layout(B, Hf, Wf).total;
}

extern "C" int ofr_resample_bwd(const float* in1, const float* flow, const float* gO, float* gI1, float* gI2, int B,
                                int C, int Hin, int Win, int Hf, int Wf, int mode, void* ws, size_t ws_bytes,
                                ofr_stream_t stream) {
    return ofr_resample_bwd_scaled(in1, flow, gO, gI1, gI2, B, C, Hin, Win, Hf, Wf, mode, 1.0f, ws, ws_bytes, stream);
}

extern "C" int ofr_resample_bwd_scaled(const float* in1, const float* flow, const float* gO, float* gI1, float* gI2,
                                       int B, int C, int Hin, int Win, int Hf, int Wf, int mode, float flow_scale,
                                       void* ws, size_t ws_bytes, ofr_stream_t stream) {
    if (!in1 || !flow || !gO || !geom_ok(B, C, Hin, Win, Hf, Wf)) return OFR_EINVAL;
    if (!gI1 && !gI2) return OFR_OK;
    if ((mode == OFR_WARP_STN || mode == OFR_WARP_PWC) && (Hin != Hf || Win != Wf)) return OFR_EINVAL;
    if (mode == OFR_WARP_FLOWNET2_NEAREST) return OFR_ENOTSUP;
    if (gI1) {
        if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255u) || ws_bytes < ws_layout(B, Hf, Wf).total)
            return OFR_EWORKSPACE;
    }
    const WarpGeom g = make_geom(B, C, Hin, Win, Hf, Wf, flow_scale);
    cudaStream_t st = (cudaStream_t)stream;
    switch (mode) {
        case OFR_WARP_FLOWNET2: return run_bwd<OFR_WARP_FLOWNET2>(in1, flow, gO, gI1, gI2, g, ws, st);
        case OFR_WARP_STN: return run_bwd<OFR_WARP_STN>(in1, flow, gO, gI1, gI2, g, ws, st);
        case OFR_WARP_PWC: return run_bwd<OFR_WARP_PWC>(in1, flow, gO, gI1, gI2, g, ws, st);
        default: return OFR_EINVAL;
    }
}

namespace {

int stn_multi(const ofr_stn_scale_t* scales, int n, int B, int C, bool backward, cudaStream_t st) {
    if (!scales || n < 1 || n > OFR_STN_MAX_SCALES || B <= 0 || C <= 0) return OFR_EINVAL;
    StnParams p;
    p.n = n;
    long long at = 0;
    for (int i = 0; i < n; ++i) {
        const ofr_stn_scale_t& s = scales[i];
        if (!s.frame || !s.flow || !geom_ok(B, C, s.h, s.w, s.h, s.w)) return OFR_EINVAL;
        if (backward ? (!s.g_out || !s.g_flow) : !s.out) return OFR_EINVAL;
        StnScale& d = p.sc[i];
        d.frame = s.frame, d.flow = s.flow, d.out = s.out, d.g_out = s.g_out, d.g_flow = s.g_flow;
        d.g = make_geom(B, C, s.h, s.w, s.h, s.w, 1.0f);
        d.sx = (0.5f * (float)(s.w - 1)) * d.g.fx;  // chain factors of the stn grid, as in run_bwd
        d.sy = (0.5f * (float)(s.h - 1)) * d.g.fy;
        d.tiles_x = (s.w + 31) / 32;
        d.tiles_y = (s.h + (kFwdThreads / 32) * kPX - 1) / ((kFwdThreads / 32) * kPX);
        d.cta_begin = (int)at;
        at += (long long)d.tiles_x * d.tiles_y * B;
        if (at >= (1ll << 31)) return OFR_ENOTSUP;
    }
    for (int i = n; i < OFR_STN_MAX_SCALES; ++i) p.sc[i] = p.sc[n - 1];
    p.cta_end = (int)at;
    if (backward)
        stn_multi_kernel<true><<<(unsigned)at, kFwdThreads, 0, st>>>(p);
    else
        stn_multi_kernel<false><<<(unsigned)at, kFwdThreads, 0, st>>>(p);
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

}  // namespace

extern "C" int ofr_stn_multi_fwd(const ofr_stn_scale_t* scales, int n_scales, int B, int C, ofr_stream_t stream) {
    return stn_multi(scales, n_scales, B, C, false, (cudaStream_t)stream);
}

extern "C" int ofr_stn_multi_bwd(const ofr_stn_scale_t* scales, int n_scales, int B, int C, ofr_stream_t stream) {
    return stn_multi(scales, n_scales, B, C, true, (cudaStream_t)stream);
}

#if OFR_WARP_TRACE
extern "C" int ofr_debug_warp_trace(long long* host) {
    return (int)cudaMemcpyFromSymbol(host, ofr_warp_trace_buf, sizeof(ofr_warp_trace_buf));
}
#endif
