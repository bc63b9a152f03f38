// Fused multi-scale self-supervised loss (replaces the eager OFEloss, loss.py:66-84, and its terms
// photometric_loss :9-13, smoothness_loss :23-30, charbonnier :33-35, correlation_loss :52-64).
//
// Forward  : ONE launch streams every scale once.  A CTA owns a strip of rows of one image of one scale; a thread
//            owns a column of 4-pixel groups (128-bit loads of warped / flow) and walks kRPT consecutive rows, so
//            the x-resampling taps are formed once, the lower flow row of one step is the upper row of the next,
//            and no index needs a division.  Per pixel it resamples the full-resolution fixed image to the scale (bilinear,
//            align_corners=False -- the F.interpolate of loss.py:11,54 folded in) and accumulates the seven
//            sums the three terms need:
//              sum charb(f-w), sum w, sum f, sum w^2, sum f^2, sum w*f, sum (charb(dv)+charb(dh))/2
//            plus min/max of w and f (the reference's "centred tensor is identically zero" branch,
//            loss.py:58-60, fires exactly when a tensor is constant).  The 4 pixels of a thread are summed
//            in float32, everything above that in float64.  CTA partials go to the workspace; a second
//            one-CTA launch adds them in a fixed order, forms the per-scale terms and the weighted totals
//            (weights 0.05*(i+1), float64, as in loss.py:71).  No atomics: bitwise reproducible.
//            Algorithmic bytes per scale: 4*B*h*w*(1 warped + 2 flow) + the taps of `fixed`
//            (<= 4*B*Hs*Ws, L2-resident across scales).
// Backward : ONE launch over every scale with the same thread mapping, using the moments kept in `state`.  The
//            charbonnier derivative of every flow difference is evaluated once and used for both pixels it couples
//            (the vertical one is carried down the thread's rows): the kernel is bound by the MUFU pipe
//            (two special-function evaluations per derivative), so halving the evaluations is what counts.
//            d/d warped gets the photometric and NCC terms, d/d flow the smoothness term; `fixed`
//            is an input image and receives no gradient (it never requires one in the reference).
#include "ofr_common.cuh"
#include "warp_math.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxCtasPerScale = 8192;  // CTAs per scale (bounds the partials workspace); beyond it a CTA loops over items
constexpr int kNumPart = 11;           // partial slots per CTA, see enum

enum { P_PHOTO = 0, P_SW, P_SF, P_SWW, P_SFF, P_SWF, P_SMOOTH, P_MINW, P_MAXW, P_MINF, P_MAXF };
// state slots per scale (OFR_LOSS_STATE_DOUBLES = 16)
enum { S_PHOTO = 0, S_SW, S_SF, S_SWW, S_SFF, S_SWF, S_SMOOTH, S_MINW, S_MAXW, S_MINF, S_MAXF,
       S_N = 11, S_DEGENERATE = 12, S_TERM_P = 13, S_TERM_C = 14, S_TERM_S = 15 };

struct ScaleDev {
    const float* warped;
    const float* flow;
    float* g_warped;
    float* g_flow;
    int h, w;
    int rows_per_cta;  // work item = (image, strip of rows_per_cta rows)
    int strips;        // strips per image
    int tx_shift;      // the CTA's 256 threads are (1 << tx_shift) pixel-group columns x (256 >> tx_shift) row lanes
    int vec4;          // 128-bit accesses: w % 4 == 0 and every tensor of the scale 16-byte aligned (else scalar)
    int items;         // strips * B work items; the scale's CTAs stride over them
    // fused registration loss (ofr_regloss_*): warped = stn(flow, moving) is formed inside the kernels
    const float* moving;     // [B,1,h,w] moving image resized to this scale, or nullptr (plain OFEloss)
    float* warped_out;       // forward: the warped image is also written here (the backward reads it back)
    ofr_warp::WarpGeom wg;   // stn geometry of the scale
    float sx, sy;            // chain factors d(sample position)/d(flow) of the stn grid
};

struct LossParams {
    const float* fixed;
    int B, Hs, Ws, n;
    double lamb_da, gamma, zeta;
    ScaleDev sc[OFR_LOSS_MAX_SCALES];
    int cta_begin[OFR_LOSS_MAX_SCALES + 1];  // CTA index range of each scale in the multi-scale launches
    int grp_begin[OFR_LOSS_MAX_SCALES + 1];  // final reduction: range of thread groups that fold each scale's CTA partials
};

__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// (x^2 + (1e-9)^2)^0.25 in float32 like the reference; MUFU square roots (<= 2 ulp), far inside the 1e-5 budget.
// The arguments are >= 1e-18 (never subnormal), so the .ftz forms -- one MUFU each, no range fix-up code -- are exact
// replacements of the non-ftz ones.
__device__ __forceinline__ float charb(float x) { return sqrt_approx(sqrt_approx(fmaf(x, x, 1e-18f))); }
// d/dx = 0.5 * x * (x^2 + eps^2)^-0.75
__device__ __forceinline__ float charb_grad(float x) {
    const float s = fmaf(x, x, 1e-18f);
    const float r = rsqrt_approx(s);          // s^-0.5
    // (explicitly rounded product: the sums it enters must not contract it into an FMA differently from kernel to
    // kernel -- the fused registration loss is bit-identical to the separate warp and loss kernels)
    return __fmul_rn(0.5f * x * r, sqrt_approx(r));  // s^-0.5 * s^-0.25
}

// F.interpolate(fixed, (h,w), mode='bilinear', align_corners=False) evaluated at one output pixel,
// with ATen's index/weight arithmetic (area_pixel_compute_source_index).
struct AxisTap {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ AxisTap axis_tap(float scale, int dst, int n) {
    float f = __fsub_rn(__fmul_rn(scale, (float)dst + 0.5f), 0.5f);
    f = f < 0.f ? 0.f : f;
    AxisTap t;
    t.i0 = min((int)f, n - 1);
    t.i1 = t.i0 + (t.i0 < n - 1);
    t.l1 = f - (float)t.i0;
    t.l0 = 1.f - t.l1;
    return t;
}
__device__ __forceinline__ float resize_at(const float* __restrict__ r0, const float* __restrict__ r1,
                                           const AxisTap& ty, const AxisTap& tx) {
    const float top = __fadd_rn(__fmul_rn(tx.l0, __ldg(r0 + tx.i0)), __fmul_rn(tx.l1, __ldg(r0 + tx.i1)));
    const float bot = __fadd_rn(__fmul_rn(tx.l0, __ldg(r1 + tx.i0)), __fmul_rn(tx.l1, __ldg(r1 + tx.i1)));
    return __fadd_rn(__fmul_rn(ty.l0, top), __fmul_rn(ty.l1, bot));
}

__device__ __forceinline__ int scale_of_cta(const LossParams& p, int cta) {
    int s = 0;
    while (s + 1 < p.n && cta >= p.cta_begin[s + 1]) ++s;
    return s;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ float warp_minf(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}
__device__ __forceinline__ float warp_maxf(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}

// Work decomposition shared by forward and backward: CTA -> (scale, image b, row strip); threads sweep the
// strip as (row, group of VEC pixels).  VEC = 4 when w % 4 == 0 (always true for the reference's scales >= 4).
struct Strip {
    int s, b, y0, y1;
};
__device__ __forceinline__ Strip strip_of_item(const LossParams& p, int s, int item) {
    Strip st;
    st.s = s;
    const ScaleDev& sc = p.sc[s];
    st.b = item / sc.strips;
    st.y0 = (item - st.b * sc.strips) * sc.rows_per_cta;
    st.y1 = min(st.y0 + sc.rows_per_cta, sc.h);
    return st;
}

template <int VEC>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[VEC]) {
    if (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x, v[1 % VEC] = t.y, v[2 % VEC] = t.z, v[3 % VEC] = t.w;
    } else {
        v[0] = __ldg(p);
    }
}
template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const float (&v)[VEC]) {
    if (VEC == 4)
        __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]));
    else
        p[0] = v[0];
}

template <int VEC>
__device__ __forceinline__ void zero_vec(float (&v)[VEC]) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) v[k] = 0.f;
}

constexpr int kRPT = 4;  // consecutive rows walked by one thread

// Thread mapping shared by forward and backward: thread -> (pixel-group column gx, row lane ry); it visits the
// pixel groups gx, gx + tx_n, ... of rows [ya, yb) = the ry-th run of kRPT rows of the CTA's strip.
struct Walk {
    int gx, tx_n, ya, yb;
};
__device__ __forceinline__ Walk walk_of_thread(const ScaleDev& sc, const Strip& st) {
    Walk wk;
    wk.tx_n = 1 << sc.tx_shift;
    wk.gx = threadIdx.x & (wk.tx_n - 1);
    const int ry = threadIdx.x >> sc.tx_shift;
    wk.ya = st.y0 + ry * kRPT;
    wk.yb = min(wk.ya + kRPT, st.y1);
    return wk;
}

// fixed resampled to the scale at pixels x..x+VEC-1 of row y (x taps precomputed in tx[])
template <int VEC>
__device__ __forceinline__ void fixed_row(const float* __restrict__ fixed, int Ws, bool ident, int x, int y,
                                          const AxisTap& ty, const AxisTap (&tx)[VEC], float (&fv)[VEC]) {
    if (ident) {  // scale == full resolution: weights are exactly (1, 0), the resize is the identity
        load_vec<VEC>(fixed + (size_t)y * Ws + x, fv);
    } else {
        const float* r0 = fixed + (size_t)ty.i0 * Ws;
        const float* r1 = fixed + (size_t)ty.i1 * Ws;
#pragma unroll
        for (int k = 0; k < VEC; ++k) fv[k] = resize_at(r0, r1, ty, tx[k]);
    }
}

// value of stn(flow, moving) at one pixel: the operation order of resample2d.cu's stn forward (bit-identical)
__device__ __forceinline__ float stn_value(const float* __restrict__ mv, int x, int y, float u, float v,
                                           const ofr_warp::WarpGeom& wg) {
    using namespace ofr_warp;
    const Samp s = make_samp<OFR_WARP_STN>(x, y, u, v, wg);
    const float a = tap_sum(mv, fwd_taps<OFR_WARP_STN>(s, wg.Hin, wg.Win));
    return a;
}

template <int VEC, bool FUSED>
__device__ __forceinline__ void partial_body(const LossParams& p, const Strip& st, float (&mm)[4], double (&acc)[7]) {
    const ScaleDev sc = p.sc[st.s];
    const int h = sc.h, w = sc.w, gw = w / VEC, hw = h * w;
    const float sh = (float)p.Hs / (float)h, sw = (float)p.Ws / (float)w;
    const bool ident = VEC == 4 && h == p.Hs && w == p.Ws && (p.Ws % 4) == 0 && (reinterpret_cast<uintptr_t>(p.fixed) & 15u) == 0;
    const float* fixed = p.fixed + (size_t)st.b * p.Hs * p.Ws;
    const float* wp = sc.warped + (size_t)st.b * hw;
    const float* fu = sc.flow + (size_t)st.b * 2 * hw;
    const float* fvp = fu + hw;
    const float* mv = FUSED ? sc.moving + (size_t)st.b * hw : nullptr;
    float* wo = FUSED ? sc.warped_out + (size_t)st.b * hw : nullptr;
    const Walk wk = walk_of_thread(sc, st);
    if (wk.ya >= wk.yb) return;
    for (int g = wk.gx; g < gw; g += wk.tx_n) {
        const int x = g * VEC;
        AxisTap tx[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) tx[k] = axis_tap(sw, x + k, p.Ws);
        const bool has_right = x + VEC < w;
        float u[VEC], v[VEC];
        int o = wk.ya * w + x;
        load_vec<VEC>(fu + o, u);
        load_vec<VEC>(fvp + o, v);
        for (int y = wk.ya; y < wk.yb; ++y, o += w) {
            float wv[VEC], fx[VEC], ud[VEC], vd[VEC];
            if (FUSED) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) wv[k] = stn_value(mv, x + k, y, u[k], v[k], sc.wg);
                store_vec<VEC>(wo + o, wv);
            } else {
                load_vec<VEC>(wp + o, wv);
            }
            if (y + 1 < h) {
                load_vec<VEC>(fu + o + w, ud);
                load_vec<VEC>(fvp + o + w, vd);
            } else {  // the last row is differenced against 0 (loss.py:25)
                zero_vec<VEC>(ud);
                zero_vec<VEC>(vd);
            }
            const float ur = has_right ? __ldg(fu + o + VEC) : 0.f;  // right neighbour of the last pixel (0 past the edge)
            const float vr = has_right ? __ldg(fvp + o + VEC) : 0.f;
            fixed_row<VEC>(fixed, p.Ws, ident, x, y, axis_tap(sh, y, p.Hs), tx, fx);
            float f32[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) f32[k] = 0.f;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const float fv = fx[k], a = wv[k];
                f32[P_PHOTO] += charb(__fsub_rn(fv, a));
                f32[P_SW] += a;
                f32[P_SF] += fv;
                f32[P_SWW] = fmaf(a, a, f32[P_SWW]);
                f32[P_SFF] = fmaf(fv, fv, f32[P_SFF]);
                f32[P_SWF] = fmaf(a, fv, f32[P_SWF]);
                mm[0] = fminf(mm[0], a), mm[1] = fmaxf(mm[1], a), mm[2] = fminf(mm[2], fv), mm[3] = fmaxf(mm[3], fv);
                const float un = (k + 1 < VEC) ? u[(k + 1) % VEC] : ur;
                const float vn = (k + 1 < VEC) ? v[(k + 1) % VEC] : vr;
                const float sm = __fadd_rn(charb(__fsub_rn(u[k], ud[k])), charb(__fsub_rn(u[k], un))) +
                                 __fadd_rn(charb(__fsub_rn(v[k], vd[k])), charb(__fsub_rn(v[k], vn)));
                f32[P_SMOOTH] = fmaf(sm, 0.5f, f32[P_SMOOTH]);
            }
#pragma unroll
            for (int k = 0; k < 7; ++k) acc[k] += (double)f32[k];
#pragma unroll
            for (int k = 0; k < VEC; ++k) u[k] = ud[k], v[k] = vd[k];  // the lower row is the next step's upper row
        }
    }
}

template <bool FUSED>
__global__ void __launch_bounds__(kThreads, FUSED ? 3 : 4) loss_partial_kernel(const __grid_constant__ LossParams p,
                                                                            double* __restrict__ partials) {
    __shared__ double red[kNumPart][kThreads / 32];
    const int s = scale_of_cta(p, blockIdx.x);
    const int c0 = blockIdx.x - p.cta_begin[s], nc = p.cta_begin[s + 1] - p.cta_begin[s];
    double acc[7];
    float mm[4] = {3.0e38f, -3.0e38f, 3.0e38f, -3.0e38f};
#pragma unroll
    for (int k = 0; k < 7; ++k) acc[k] = 0.0;
    auto run = [&](int item) {
        const Strip st = strip_of_item(p, s, item);
        if (p.sc[s].vec4)
            partial_body<4, FUSED>(p, st, mm, acc);
        else
            partial_body<1, FUSED>(p, st, mm, acc);
    };
    // one item per CTA (straight-line code: the loop form costs 7-15 % here) unless strips * B > kMaxCtasPerScale
    if (p.sc[s].items <= nc) {
        run(c0);
    } else {
#pragma unroll 1
        for (int item = c0; item < p.sc[s].items; item += nc) run(item);
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const double v = warp_sum(acc[k]);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
    }
    {
        const float a = warp_minf(mm[0]), b = warp_maxf(mm[1]), c = warp_minf(mm[2]), d = warp_maxf(mm[3]);
        if ((threadIdx.x & 31) == 0) {
            red[P_MINW][threadIdx.x >> 5] = a;
            red[P_MAXW][threadIdx.x >> 5] = b;
            red[P_MINF][threadIdx.x >> 5] = c;
            red[P_MAXF][threadIdx.x >> 5] = d;
        }
    }
    __syncthreads();
    if (threadIdx.x < kNumPart) {
        const int k = threadIdx.x;
        double v = red[k][0];
        for (int j = 1; j < kThreads / 32; ++j) {
            const double t = red[k][j];
            v = (k == P_MINW || k == P_MINF) ? fmin(v, t) : (k == P_MAXW || k == P_MAXF) ? fmax(v, t) : v + t;
        }
        partials[(long long)blockIdx.x * kNumPart + k] = v;
    }
}

#ifndef OFR_LOSS_TRACE
#define OFR_LOSS_TRACE 0   // 1: clock64() marks of thread 0 of loss_final_kernel, read with ofr_debug_loss_trace
#endif
#if OFR_LOSS_TRACE
__device__ long long ofr_loss_trace_buf[8];
#define OFR_LT(i) do { if (threadIdx.x == 0) ofr_loss_trace_buf[i] = clock64(); } while (0)
#else
#define OFR_LT(i)
#endif

// One CTA.  The CTA partials of a scale are a dense [partial][kNumPart] array of doubles; the kFinalGroups groups of
// kNumPart threads are dealt to the scales in proportion to their partial counts (fill_params).  Group j of a scale's G
// folds partials j, j+G, ... in ascending order, lane k of the group owning component k: consecutive threads read
// consecutive doubles (the first version read them at an 88-byte stride with 11 accumulators per thread in 64 registers
// and spent 26 000 of the kernel's 31 000 cycles here).  One thread per (scale, component) then adds the group results
// in group order.  Everything is a fixed function of the launch shape: deterministic.
constexpr int kFinalGroups = 1024 / kNumPart;          // 93
constexpr int kFinalBatch = 8;                         // loads in flight per thread (12, 16, 24: no faster)
__device__ __forceinline__ double part_identity(int k) {
    return (k == P_MINW || k == P_MINF) ? 1e300 : (k == P_MAXW || k == P_MAXF) ? -1e300 : 0.0;
}
// the component index is a run-time value here: all three candidates are formed and one is selected, so that the
// compiler emits selects (a `k == ...` chain became an indirect branch through a jump table per element: 480 cycles each)
struct PartOp {
    bool is_min, is_max;
    __device__ explicit PartOp(int k) : is_min(k == P_MINW || k == P_MINF), is_max(k == P_MAXW || k == P_MAXF) {}
    __device__ __forceinline__ double operator()(double a, double b) const {
        const double sum = a + b, lo = fmin(a, b), hi = fmax(a, b);
        return is_min ? lo : is_max ? hi : sum;
    }
};
__global__ void __launch_bounds__(1024) loss_final_kernel(const __grid_constant__ LossParams p,
                                                          const double* __restrict__ partials,
                                                          double* __restrict__ state, double* __restrict__ out4) {
    __shared__ double red[kFinalGroups][kNumPart];
    OFR_LT(0);
    {
        const int g = threadIdx.x / kNumPart, k = threadIdx.x - g * kNumPart;
        if (g < p.grp_begin[p.n]) {
            int s = 0;  // scale of the group: grp_begin ascends (fixed offsets into the parameters: no dependent look-ups)
#pragma unroll
            for (int i = 1; i < OFR_LOSS_MAX_SCALES; ++i) s += (i < p.n && g >= p.grp_begin[i]) ? 1 : 0;
            const int G = p.grp_begin[s + 1] - p.grp_begin[s], cend = p.cta_begin[s + 1];
            const PartOp fold(k);
            const double ident = part_identity(k);
            double v = ident;
            for (int c = p.cta_begin[s] + (g - p.grp_begin[s]); c < cend; c += kFinalBatch * G) {
                double x[kFinalBatch];
#pragma unroll
                for (int q = 0; q < kFinalBatch; ++q)
                    x[q] = (c + q * G < cend) ? partials[(long long)(c + q * G) * kNumPart + k] : ident;
#pragma unroll
                for (int q = 0; q < kFinalBatch; ++q) v = fold(v, x[q]);
            }
            red[g][k] = v;
        }
    }
    OFR_LT(1);
    __syncthreads();
    OFR_LT(2);
    const int s = threadIdx.x >> 7, t = threadIdx.x & 127;  // from here on: 128 threads per scale
    if (s < p.n && t < kNumPart) {
        const int k = t;
        const PartOp fold(k);
        double x = red[p.grp_begin[s]][k];  // (four interleaved chains were not faster than this one: 5 000 against 3 900 cycles)
        for (int g = p.grp_begin[s] + 1; g < p.grp_begin[s + 1]; ++g) x = fold(x, red[g][k]);
        state[s * OFR_LOSS_STATE_DOUBLES + k] = x;
    }
    __syncthreads();
    // per-scale terms: one thread per scale (float64 sqrt / divisions are long serial chains), then thread 0 adds them
    __shared__ double terms[OFR_LOSS_MAX_SCALES][3];
    OFR_LT(3);
    if (s < p.n && t == 0) {
        const int i = s;
        double* st = state + i * OFR_LOSS_STATE_DOUBLES;
        const double N = (double)p.B * p.sc[i].h * p.sc[i].w, Bd = (double)p.B;
        const double mw = st[S_SW] / N, mf = st[S_SF] / N;
        const double sxx = st[S_SWW] - N * mw * mw, syy = st[S_SFF] - N * mf * mf, sxy = st[S_SWF] - N * mw * mf;
        // loss.py:58-60: a centred tensor that is identically zero (a constant image) short-circuits to corr = 1
        const bool degenerate = (st[S_MINW] == st[S_MAXW]) || (st[S_MINF] == st[S_MAXF]) || !(sxx > 0.0) || !(syy > 0.0);
        const double corr = degenerate ? 1.0 : (1.0 / Bd) * sxy / (sqrt(sxx) * sqrt(syy));
        const double photo = st[S_PHOTO] / Bd, ncc = 1.0 - corr, smooth = st[S_SMOOTH] / Bd;
        st[S_N] = N;
        st[S_DEGENERATE] = degenerate ? 1.0 : 0.0;
        st[S_TERM_P] = photo;
        st[S_TERM_C] = ncc;
        st[S_TERM_S] = smooth;
        const double wi = 0.05 * (double)(i + 1);
        terms[i][0] = wi * photo;
        terms[i][1] = wi * ncc;
        terms[i][2] = wi * smooth;
    }
    OFR_LT(4);
    __syncthreads();
    OFR_LT(5);
    if (threadIdx.x != 0) return;
    double tp = 0.0, tc = 0.0, ts = 0.0;
    for (int i = 0; i < p.n; ++i) {  // ascending, like the reference's loop (loss.py:74-80)
        tp += terms[i][0];
        tc += terms[i][1];
        ts += terms[i][2];
    }
    const double inv_n = 1.0 / (double)p.n;
    const double P = inv_n * p.gamma * tp, C = inv_n * p.zeta * tc, S = inv_n * p.lamb_da * ts;
    out4[0] = P;
    out4[1] = C;
    out4[2] = S;
    out4[3] = P + S + C;  // loss.py:83 adds in this order
    OFR_LT(6);
}

template <int VEC>
__device__ __forceinline__ void bwd_body(const LossParams& p, const Strip& st, float cp, float cs, float k1, float k2,
                                         float mf, float mw) {
    const ScaleDev sc = p.sc[st.s];
    const int h = sc.h, w = sc.w, gw = w / VEC, hw = h * w;
    const float sh = (float)p.Hs / (float)h, sw = (float)p.Ws / (float)w;
    const bool ident = VEC == 4 && h == p.Hs && w == p.Ws && (p.Ws % 4) == 0 && (reinterpret_cast<uintptr_t>(p.fixed) & 15u) == 0;
    const float* fixed = p.fixed + (size_t)st.b * p.Hs * p.Ws;
    const Walk wk = walk_of_thread(sc, st);
    if (wk.ya >= wk.yb) return;
    // (One pass over the rows doing all three gradients at once needed 80+ registers with spills and ran 25 % slower.)
    for (int g = wk.gx; g < gw; g += wk.tx_n) {
        const int x = g * VEC;
        if (sc.g_warped) {
            AxisTap tx[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) tx[k] = axis_tap(sw, x + k, p.Ws);
            const float* wp = sc.warped + (size_t)st.b * hw;
            float* gp = sc.g_warped + (size_t)st.b * hw;
            int o = wk.ya * w + x;
            for (int y = wk.ya; y < wk.yb; ++y, o += w) {
                float wv[VEC], fx[VEC], gq[VEC];
                load_vec<VEC>(wp + o, wv);
                fixed_row<VEC>(fixed, p.Ws, ident, x, y, axis_tap(sh, y, p.Hs), tx, fx);
#pragma unroll
                for (int k = 0; k < VEC; ++k)  // photometric: -cp * psi(f - w);  NCC: k1 * (f - mean f) - k2 * (w - mean w)
                    gq[k] = fmaf(-cp, charb_grad(__fsub_rn(fx[k], wv[k])), fmaf(k1, fx[k] - mf, -k2 * (wv[k] - mw)));
                store_vec<VEC>(gp + o, gq);
            }
        }
        if (sc.g_flow) {
            const bool has_right = x + VEC < w, has_left = x > 0;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float* f = sc.flow + ((size_t)st.b * 2 + c) * hw;
                float* gp = sc.g_flow + ((size_t)st.b * 2 + c) * hw;
                int o = wk.ya * w + x;
                float v[VEC], up[VEC];  // up[k] = psi(v[y-1][k] - v[y][k]), carried from the previous row
                load_vec<VEC>(f + o, v);
                if (wk.ya > 0) {
                    float t[VEC];
                    load_vec<VEC>(f + o - w, t);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) up[k] = charb_grad(__fsub_rn(t[k], v[k]));
                } else {
                    zero_vec<VEC>(up);
                }
                for (int y = wk.ya; y < wk.yb; ++y, o += w) {
                    float dn[VEC], gq[VEC], hd[VEC + 1];
                    if (y + 1 < h) {
                        load_vec<VEC>(f + o + w, dn);
                    } else {
                        zero_vec<VEC>(dn);
                    }
                    const float rt = has_right ? __ldg(f + o + VEC) : 0.f;
                    // horizontal derivatives: hd[k] = psi(v[k-1] - v[k]) for k = 0..VEC (v[-1] = left, v[VEC] = right)
                    hd[0] = has_left ? charb_grad(__fsub_rn(__ldg(f + o - 1), v[0])) : 0.f;
#pragma unroll
                    for (int k = 1; k <= VEC; ++k)
                        hd[k] = charb_grad(__fsub_rn(v[k - 1], (k < VEC) ? v[k % VEC] : rt));
#pragma unroll
                    for (int k = 0; k < VEC; ++k) {
                        const float vd = charb_grad(__fsub_rn(v[k], dn[k]));  // also the next row's `up`
                        gq[k] = cs * (((vd + hd[k + 1]) - up[k]) - hd[k]);
                        up[k] = vd;
                        v[k] = dn[k];
                    }
                    store_vec<VEC>(gp + o, gq);
                }
            }
        }
    }
}

// Backward of the fused registration loss: d total / d flow = smoothness part + (d total / d warped) chained through
// stn(flow, moving).  Per thread: the gradient w.r.t. the warped pixels of its rows is formed first (as in bwd_body)
// and turned into the two flow-gradient contributions of the warp (the arithmetic of resample2d.cu's stn backward,
// kept in registers), then the smoothness sweeps add theirs: the flow gradient is written once, and neither the
// gradient of the warped image nor the two partial flow gradients exist in memory.
template <int VEC>
__device__ __forceinline__ void bwd_body_fused(const LossParams& p, const Strip& st, float cp, float cs, float k1,
                                               float k2, float mf, float mw) {
    using namespace ofr_warp;
    const ScaleDev sc = p.sc[st.s];
    const int h = sc.h, w = sc.w, gw = w / VEC, hw = h * w;
    const float sh = (float)p.Hs / (float)h, sw = (float)p.Ws / (float)w;
    const bool ident = VEC == 4 && h == p.Hs && w == p.Ws && (p.Ws % 4) == 0 && (reinterpret_cast<uintptr_t>(p.fixed) & 15u) == 0;
    const float* fixed = ofr_warp::opaque_ptr(p.fixed + (size_t)st.b * p.Hs * p.Ws);
    const float* wp = ofr_warp::opaque_ptr(sc.warped + (size_t)st.b * hw);
    const float* mv = ofr_warp::opaque_ptr(sc.moving + (size_t)st.b * hw);
    const float* f0 = ofr_warp::opaque_ptr(sc.flow + (size_t)st.b * 2 * hw);
    float* gfp = ofr_warp::opaque_ptr(sc.g_flow + (size_t)st.b * 2 * hw);
    const Walk wk = walk_of_thread(sc, st);
    if (wk.ya >= wk.yb) return;
    for (int g = wk.gx; g < gw; g += wk.tx_n) {
        const int x = g * VEC;
        AxisTap tx[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) tx[k] = axis_tap(sw, x + k, p.Ws);
        float gwarp[2][kRPT][VEC];  // flow-gradient contributions through the warp, rows ya .. ya+kRPT-1
#pragma unroll
        for (int r = 0; r < kRPT; ++r) {
            const int y = wk.ya + r;
            if (y < wk.yb) {
                const int o = y * w + x;
                float wv[VEC], fx[VEC], u[VEC], v[VEC];
                load_vec<VEC>(wp + o, wv);
                load_vec<VEC>(f0 + o, u);
                load_vec<VEC>(f0 + hw + o, v);
                fixed_row<VEC>(fixed, p.Ws, ident, x, y, axis_tap(sh, y, p.Hs), tx, fx);
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    // d total / d warped: photometric -cp * psi(f - w);  NCC k1 * (f - mean f) - k2 * (w - mean w)
                    const float go = fmaf(-cp, charb_grad(__fsub_rn(fx[k], wv[k])), fmaf(k1, fx[k] - mf, -k2 * (wv[k] - mw)));
                    const Samp s = make_samp<OFR_WARP_STN>(x + k, y, u[k], v[k], sc.wg);
                    const TapOff t = tap_offsets<OFR_WARP_STN>(s.x0, s.y0, h, w);
                    const float tl = fetch(mv, t.tl), tr = fetch(mv, t.tr), bl = fetch(mv, t.bl), br = fetch(mv, t.br);
                    // d out / d ix = wy0*(TR-TL) + wy1*(BR-BL);  d out / d iy = wx0*(BL-TL) + wx1*(BR-TR)
                    const float ax = fmaf(go, fmaf(s.wy1, br - bl, s.wy0 * (tr - tl)), 0.f);
                    const float ay = fmaf(go, fmaf(s.wx1, br - tr, s.wx0 * (bl - tl)), 0.f);
                    gwarp[0][r][k] = __fmul_rn(ax, sc.sx);
                    gwarp[1][r][k] = __fmul_rn(ay, sc.sy);
                }
            }
        }
        const bool has_right = x + VEC < w, has_left = x > 0;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float* f = ofr_warp::opaque_ptr(f0 + c * hw);
            float* gp = ofr_warp::opaque_ptr(gfp + c * hw);
            float v[VEC], up[VEC];  // up[k] = psi(v[y-1][k] - v[y][k]), carried from the previous row
            {
                const int o = wk.ya * w + x;
                load_vec<VEC>(f + o, v);
                if (wk.ya > 0) {
                    float t[VEC];
                    load_vec<VEC>(f + o - w, t);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) up[k] = charb_grad(__fsub_rn(t[k], v[k]));
                } else {
                    zero_vec<VEC>(up);
                }
            }
#pragma unroll
            for (int r = 0; r < kRPT; ++r) {
                const int y = wk.ya + r;
                if (y < wk.yb) {
                    const int o = y * w + x;
                    float dn[VEC], gq[VEC], hd[VEC + 1];
                    if (y + 1 < h) {
                        load_vec<VEC>(f + o + w, dn);
                    } else {
                        zero_vec<VEC>(dn);
                    }
                    const float rt = has_right ? __ldg(f + o + VEC) : 0.f;
                    hd[0] = has_left ? charb_grad(__fsub_rn(__ldg(f + o - 1), v[0])) : 0.f;
#pragma unroll
                    for (int k = 1; k <= VEC; ++k)
                        hd[k] = charb_grad(__fsub_rn(v[k - 1], (k < VEC) ? v[k % VEC] : rt));
#pragma unroll
                    for (int k = 0; k < VEC; ++k) {
                        const float vd = charb_grad(__fsub_rn(v[k], dn[k]));
                        // autograd adds the two flow gradients as float32 tensors: product, then one rounded add
                        gq[k] = __fadd_rn(__fmul_rn(cs, ((vd + hd[k + 1]) - up[k]) - hd[k]), gwarp[c][r][k]);
                        up[k] = vd;
                        v[k] = dn[k];
                    }
                    store_vec<VEC>(gp + o, gq);
                }
            }
        }
    }
}

// per-scale coefficients of the backward: upstream gradients x loss.py's weights
struct BwdCoef {
    float cp, cs, k1, k2, mf, mw;
};
__device__ __forceinline__ BwdCoef bwd_coef(const LossParams& p, int s, const double* __restrict__ grad4,
                                            const double* __restrict__ state) {
    const double* sd = state + s * OFR_LOSS_STATE_DOUBLES;
    const double wi = 0.05 * (double)(s + 1), inv_n = 1.0 / (double)p.n, Bd = (double)p.B;
    const double gP = grad4[0] + grad4[3], gC = grad4[1] + grad4[3], gS = grad4[2] + grad4[3];
    BwdCoef c;
    c.cp = (float)(gP * inv_n * p.gamma * wi / Bd);            // d total / d photometric-sum
    c.cs = (float)(gS * inv_n * p.lamb_da * wi / (2.0 * Bd));  // d total / d smoothness-sum
    // NCC: d(1 - corr)/dw_i = -(1/B) * [ vy_i / sqrt(sxx*syy) - sxy * vx_i / (sxx^1.5 * syy^0.5) ]
    const double N = sd[S_N], mw = sd[S_SW] / N, mf = sd[S_SF] / N;
    const double sxx = sd[S_SWW] - N * mw * mw, syy = sd[S_SFF] - N * mf * mf, sxy = sd[S_SWF] - N * mw * mf;
    const bool degenerate = sd[S_DEGENERATE] != 0.0;
    const double cc = degenerate ? 0.0 : -(gC * inv_n * p.zeta * wi / Bd);
    c.k1 = degenerate ? 0.f : (float)(cc / (sqrt(sxx) * sqrt(syy)));
    c.k2 = degenerate ? 0.f : (float)(cc * sxy / (sxx * sqrt(sxx) * sqrt(syy)));
    c.mf = (float)mf, c.mw = (float)mw;
    return c;
}

// (four CTAs per SM: 64 registers, as before the item loop was added -- at 80 registers the kernel ran 15 % slower)
__global__ void __launch_bounds__(kThreads, 4) loss_bwd_kernel(const __grid_constant__ LossParams p,
                                                            const double* __restrict__ grad4,
                                                            const double* __restrict__ state) {
    const int s = scale_of_cta(p, blockIdx.x);
    const int c0 = blockIdx.x - p.cta_begin[s], nc = p.cta_begin[s + 1] - p.cta_begin[s];
    const BwdCoef c = bwd_coef(p, s, grad4, state);
    auto run = [&](int item) {
        const Strip st = strip_of_item(p, s, item);
        if (p.sc[s].vec4)
            bwd_body<4>(p, st, c.cp, c.cs, c.k1, c.k2, c.mf, c.mw);
        else
            bwd_body<1>(p, st, c.cp, c.cs, c.k1, c.k2, c.mf, c.mw);
    };
    if (p.sc[s].items <= nc) {
        run(c0);
    } else {
#pragma unroll 1
        for (int item = c0; item < p.sc[s].items; item += nc) run(item);
    }
}

__global__ void __launch_bounds__(kThreads, 2) regloss_bwd_kernel(const __grid_constant__ LossParams p,
                                                                  const double* __restrict__ grad4,
                                                                  const double* __restrict__ state) {
    const int s = scale_of_cta(p, blockIdx.x);
    const int c0 = blockIdx.x - p.cta_begin[s], nc = p.cta_begin[s + 1] - p.cta_begin[s];
    const BwdCoef c = bwd_coef(p, s, grad4, state);
    auto run = [&](int item) {
        const Strip st = strip_of_item(p, s, item);
        if (p.sc[s].vec4)
            bwd_body_fused<4>(p, st, c.cp, c.cs, c.k1, c.k2, c.mf, c.mw);
        else
            bwd_body_fused<1>(p, st, c.cp, c.cs, c.k1, c.k2, c.mf, c.mw);
    };
    if (p.sc[s].items <= nc) {
        run(c0);
    } else {
#pragma unroll 1
        for (int item = c0; item < p.sc[s].items; item += nc) run(item);
    }
}

int fill_params(LossParams& p, const float* fixed, int B, int Hs, int Ws, const ofr_loss_scale_t* scales, int n,
                double lamb_da, double gamma, double zeta) {
    if (!fixed || !scales || B <= 0 || Hs <= 0 || Ws <= 0 || n < 1 || n > OFR_LOSS_MAX_SCALES) return OFR_EINVAL;
    p.fixed = fixed;
    p.B = B, p.Hs = Hs, p.Ws = Ws, p.n = n;
    p.lamb_da = lamb_da, p.gamma = gamma, p.zeta = zeta;
    int at = 0;
    for (int i = 0; i < OFR_LOSS_MAX_SCALES; ++i) {
        p.cta_begin[i] = at;
        if (i < n) {
            const ofr_loss_scale_t& s = scales[i];
            if (!s.warped || !s.flow || s.h <= 0 || s.w <= 0) return OFR_EINVAL;
            if ((long long)s.h * s.w >= (1ll << 30)) return OFR_EINVAL;
            // thread shape: tx_n = power of two covering the pixel groups of a row (at most 256), the other
            // 256 / tx_n thread lanes each walk kRPT consecutive rows -> rows per CTA
            // 128-bit path only when every tensor of the scale can take it (a contiguous view with an odd storage
            // offset is legal input: it runs on the scalar path instead of faulting)
            const bool vec4 = s.w % 4 == 0 && ofr::aligned16(s.warped) && ofr::aligned16(s.flow) &&
                              ofr::aligned16(s.g_warped) && ofr::aligned16(s.g_flow);
            const int gw = vec4 ? s.w / 4 : s.w;
            int tx_shift = 0;
            while ((1 << tx_shift) < gw && tx_shift < 8) ++tx_shift;
            int rows = (kThreads >> tx_shift) * kRPT;
            if (rows > s.h) rows = (s.h + kRPT - 1) / kRPT * kRPT;
            int strips = (s.h + rows - 1) / rows;
            const long long items = (long long)strips * B;
            if (items >= (1ll << 31)) return OFR_EINVAL;
            p.sc[i] = ScaleDev{s.warped, s.flow, s.g_warped, s.g_flow, s.h, s.w, rows, strips, tx_shift, vec4 ? 1 : 0, (int)items};
            at += (int)(items < kMaxCtasPerScale ? items : kMaxCtasPerScale);
        } else {
            p.sc[i] = ScaleDev{nullptr, nullptr, nullptr, nullptr, 0, 0, 1, 0, 0, 0, 0};
        }
    }
    p.cta_begin[OFR_LOSS_MAX_SCALES] = at;
    for (int i = n; i <= OFR_LOSS_MAX_SCALES; ++i) p.cta_begin[i] = at;
    // final reduction: the thread groups are dealt to the scales in proportion to their CTA partials, at least one each
    int groups[OFR_LOSS_MAX_SCALES], sum = 0;
    for (int i = 0; i < n; ++i) {
        const long long share = (long long)kFinalGroups * (p.cta_begin[i + 1] - p.cta_begin[i]) / (at > 0 ? at : 1);
        groups[i] = share < 1 ? 1 : (int)share;
        sum += groups[i];
    }
    while (sum > kFinalGroups) {  // the floors were lifted to 1 for small scales: take the excess from the largest share
        int big = 0;
        for (int i = 1; i < n; ++i)
            if (groups[i] > groups[big]) big = i;
        --groups[big], --sum;
    }
    p.grp_begin[0] = 0;
    for (int i = 0; i < OFR_LOSS_MAX_SCALES; ++i) p.grp_begin[i + 1] = p.grp_begin[i] + (i < n ? groups[i] : 0);
    return OFR_OK;
}

size_t ws_bytes_for(int total_ctas) { return ofr::align_up(sizeof(double) * kNumPart * (size_t)total_ctas, 256); }

}  // namespace

#if OFR_LOSS_TRACE
extern "C" int ofr_debug_loss_trace(long long* host) {
    return (int)cudaMemcpyFromSymbol(host, ofr_loss_trace_buf, sizeof(ofr_loss_trace_buf));
}
#endif

extern "C" size_t ofr_loss_workspace(int n_scales) {
    if (n_scales < 1 || n_scales > OFR_LOSS_MAX_SCALES) return 0;
    // upper bound that does not depend on the batch: fill_params never uses more than kMaxCtasPerScale CTAs per scale
    // (larger batches / images make every CTA walk several (image, strip) items).
    return ws_bytes_for(kMaxCtasPerScale * n_scales);
}

extern "C" int ofr_loss_fwd(const float* fixed, int B, int Hs, int Ws, const ofr_loss_scale_t* scales, int n_scales,
                            double lamb_da, double gamma, double zeta, double* out4, double* state, void* ws,
                            size_t ws_bytes, ofr_stream_t stream) {
    LossParams p;
    int rc = fill_params(p, fixed, B, Hs, Ws, scales, n_scales, lamb_da, gamma, zeta);
    if (rc != OFR_OK) return rc;
    if (!out4 || !state) return OFR_EINVAL;
    if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255u) || ws_bytes < ws_bytes_for(p.cta_begin[n_scales])) return OFR_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    double* partials = (double*)ws;
    loss_partial_kernel<false><<<p.cta_begin[n_scales], kThreads, 0, st>>>(p, partials);
    OFR_LAUNCH_CHECK();
    loss_final_kernel<<<1, 1024, 0, st>>>(p, partials, state, out4);
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

extern "C" int ofr_loss_bwd(const float* fixed, int B, int Hs, int Ws, const ofr_loss_scale_t* scales, int n_scales,
                            double lamb_da, double gamma, double zeta, const double* grad4, const double* state,
                            ofr_stream_t stream) {
    LossParams p;
    int rc = fill_params(p, fixed, B, Hs, Ws, scales, n_scales, lamb_da, gamma, zeta);
    if (rc != OFR_OK) return rc;
    if (!grad4 || !state) return OFR_EINVAL;
    loss_bwd_kernel<<<p.cta_begin[n_scales], kThreads, 0, (cudaStream_t)stream>>>(p, grad4, state);
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

// ------------------------------------------------------------------------------------ fused registration loss
namespace {

// scales of the fused call -> the plain loss's parameters + the stn geometry of every scale (resample2d.cu's make_geom
// and run_bwd spell the same float32 factors)
int fill_regloss(LossParams& p, const float* fixed, int B, int Hs, int Ws, const ofr_regloss_scale_t* scales, int n,
                 double lamb_da, double gamma, double zeta, bool backward) {
    if (!scales || n < 1 || n > OFR_LOSS_MAX_SCALES) return OFR_EINVAL;
    ofr_loss_scale_t plain[OFR_LOSS_MAX_SCALES];
    for (int i = 0; i < n; ++i) {
        const ofr_regloss_scale_t& s = scales[i];
        if (!s.moving || !s.warped || (backward && !s.g_flow)) return OFR_EINVAL;
        plain[i] = ofr_loss_scale_t{s.warped, s.flow, nullptr, s.g_flow, s.h, s.w};
    }
    int rc = fill_params(p, fixed, B, Hs, Ws, plain, n, lamb_da, gamma, zeta);
    if (rc != OFR_OK) return rc;
    for (int i = 0; i < n; ++i) {
        const ofr_regloss_scale_t& s = scales[i];
        ScaleDev& d = p.sc[i];
        d.moving = s.moving;
        d.warped_out = s.warped;
        ofr_warp::WarpGeom& g = d.wg;
        g.B = B, g.C = 1, g.Hin = g.Hf = s.h, g.Win = g.Wf = s.w;
        g.fs = 1.f;
        g.fx = (float)(2.0 / s.w);  // models.py:264 builds a FloatTensor from the python floats 2/w, 2/h
        g.fy = (float)(2.0 / s.h);
        d.sx = (0.5f * (float)(s.w - 1)) * g.fx;
        d.sy = (0.5f * (float)(s.h - 1)) * g.fy;
    }
    return OFR_OK;
}

}  // namespace

extern "C" int ofr_regloss_fwd(const float* fixed, int B, int Hs, int Ws, const ofr_regloss_scale_t* scales,
                               int n_scales, double lamb_da, double gamma, double zeta, double* out4, double* state,
                               void* ws, size_t ws_bytes, ofr_stream_t stream) {
    LossParams p;
    int rc = fill_regloss(p, fixed, B, Hs, Ws, scales, n_scales, lamb_da, gamma, zeta, false);
    if (rc != OFR_OK) return rc;
    if (!out4 || !state) return OFR_EINVAL;
    if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255u) || ws_bytes < ws_bytes_for(p.cta_begin[n_scales])) return OFR_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    double* partials = (double*)ws;
    loss_partial_kernel<true><<<p.cta_begin[n_scales], kThreads, 0, st>>>(p, partials);
    OFR_LAUNCH_CHECK();
    loss_final_kernel<<<1, 1024, 0, st>>>(p, partials, state, out4);
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

extern "C" int ofr_regloss_bwd(const float* fixed, int B, int Hs, int Ws, const ofr_regloss_scale_t* scales,
                               int n_scales, double lamb_da, double gamma, double zeta, const double* grad4,
                               const double* state, ofr_stream_t stream) {
    LossParams p;
    int rc = fill_regloss(p, fixed, B, Hs, Ws, scales, n_scales, lamb_da, gamma, zeta, true);
    if (rc != OFR_OK) return rc;
    if (!grad4 || !state) return OFR_EINVAL;
    regloss_bwd_kernel<<<p.cta_begin[n_scales], kThreads, 0, (cudaStream_t)stream>>>(p, grad4, state);
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}
