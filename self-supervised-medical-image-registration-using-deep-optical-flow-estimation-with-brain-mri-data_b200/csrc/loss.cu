// Fused multi-scale self-supervised loss (replaces the eager OFEloss, loss.py:66-84, and its terms
// photometric_loss :9-13, smoothness_loss :23-30, charbonnier :33-35, correlation_loss :52-64).
//
// Forward  : ONE launch streams every scale once.  Per pixel it resamples the full-resolution fixed image
//            to the scale (bilinear, align_corners=False -- the F.interpolate of loss.py:11,54 folded in),
//            and accumulates, in float64, the seven sums that the three terms need:
//              sum charb(f-w), sum w, sum f, sum w^2, sum f^2, sum w*f, sum (charb(dv)+charb(dh))/2
//            plus min/max of w and f (the reference's "centred tensor is identically zero" branch,
//            loss.py:58-60, fires exactly when a tensor is constant).  CTA partials go to the workspace;
//            a second one-CTA launch adds them in a fixed order, forms the per-scale terms and the
//            weighted totals (weights 0.05*(i+1), float64, as in loss.py:71).  No atomics anywhere:
//            bitwise reproducible.  Algorithmic bytes per scale: 4*B*h*w*(1 warped + 2 flow) + the taps
//            of `fixed` (<= 4*B*Hs*Ws, L2-resident across scales).
// Backward : ONE launch, elementwise over every scale, using the moments kept in `state`.
//            d/d warped gets the photometric and NCC terms, d/d flow the smoothness term; `fixed`
//            is an input image and receives no gradient (it never requires one in the reference).
#include "ofr_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxCtasPerScale = 296;  // two waves of the 148 SMs
constexpr int kPixPerCta = 2048;
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
};

struct LossParams {
    const float* fixed;
    int B, Hs, Ws, n;
    double lamb_da, gamma, zeta;
    ScaleDev sc[OFR_LOSS_MAX_SCALES];
    int cta_begin[OFR_LOSS_MAX_SCALES + 1];  // CTA index range of each scale in the multi-scale launches
};

__device__ __forceinline__ float charb(float x) {  // (x^2 + (1e-9)^2)^0.25, float32 like the reference
    return sqrtf(sqrtf(fmaf(x, x, 1e-18f)));
}
__device__ __forceinline__ float charb_grad(float x) {  // d/dx = 0.5 * x * (x^2 + eps^2)^-0.75
    const float s = fmaf(x, x, 1e-18f);
    const float r = sqrtf(s);
    return 0.5f * x / (r * sqrtf(r));
}

// F.interpolate(fixed, (h,w), mode='bilinear', align_corners=False) evaluated at one output pixel,
// with ATen's index/weight arithmetic (area_pixel_compute_source_index).
struct Resizer {
    float sh, sw;
    int Hs, Ws;
    __device__ __forceinline__ Resizer(int Hs_, int Ws_, int h, int w) : Hs(Hs_), Ws(Ws_) {
        sh = (float)Hs_ / (float)h;
        sw = (float)Ws_ / (float)w;
    }
    __device__ __forceinline__ float at(const float* __restrict__ img, int y, int x) const {
        float fy = __fsub_rn(__fmul_rn(sh, (float)y + 0.5f), 0.5f);
        float fx = __fsub_rn(__fmul_rn(sw, (float)x + 0.5f), 0.5f);
        fy = fy < 0.f ? 0.f : fy;
        fx = fx < 0.f ? 0.f : fx;
        const int y0 = min((int)fy, Hs - 1), x0 = min((int)fx, Ws - 1);
        const int y1 = y0 + (y0 < Hs - 1), x1 = x0 + (x0 < Ws - 1);
        const float ly = fy - (float)y0, lx = fx - (float)x0;
        const float hy = 1.f - ly, hx = 1.f - lx;
        const float* r0 = img + (long long)y0 * Ws;
        const float* r1 = img + (long long)y1 * Ws;
        const float top = __fadd_rn(__fmul_rn(hx, __ldg(r0 + x0)), __fmul_rn(lx, __ldg(r0 + x1)));
        const float bot = __fadd_rn(__fmul_rn(hx, __ldg(r1 + x0)), __fmul_rn(lx, __ldg(r1 + x1)));
        return __fadd_rn(__fmul_rn(hy, top), __fmul_rn(ly, bot));
    }
};

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
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}

__global__ void __launch_bounds__(kThreads) loss_partial_kernel(const __grid_constant__ LossParams p,
                                                                double* __restrict__ partials) {
    __shared__ double red[kNumPart][kThreads / 32];
    const int s = scale_of_cta(p, blockIdx.x);
    const ScaleDev sc = p.sc[s];
    const int cta = blockIdx.x - p.cta_begin[s], nctas = p.cta_begin[s + 1] - p.cta_begin[s];
    const int h = sc.h, w = sc.w;
    const long long hw = (long long)h * w, total = (long long)p.B * hw;
    const Resizer rz(p.Hs, p.Ws, h, w);

    double a[kNumPart];
#pragma unroll
    for (int k = 0; k < kNumPart; ++k) a[k] = 0.0;
    a[P_MINW] = a[P_MINF] = 1e300;
    a[P_MAXW] = a[P_MAXF] = -1e300;

    for (long long i = (long long)cta * kThreads + threadIdx.x; i < total; i += (long long)nctas * kThreads) {
        const int b = (int)(i / hw);
        const int rem = (int)(i - (long long)b * hw);
        const int y = rem / w, x = rem - y * w;
        const float wv = __ldg(sc.warped + i);
        const float fv = rz.at(p.fixed + (long long)b * p.Hs * p.Ws, y, x);
        a[P_PHOTO] += (double)charb(__fsub_rn(fv, wv));
        a[P_SW] += (double)wv;
        a[P_SF] += (double)fv;
        a[P_SWW] += (double)wv * (double)wv;
        a[P_SFF] += (double)fv * (double)fv;
        a[P_SWF] += (double)wv * (double)fv;
        a[P_MINW] = fmin(a[P_MINW], (double)wv);
        a[P_MAXW] = fmax(a[P_MAXW], (double)wv);
        a[P_MINF] = fmin(a[P_MINF], (double)fv);
        a[P_MAXF] = fmax(a[P_MAXF], (double)fv);
        const float* fl = sc.flow + (long long)b * 2 * hw + rem;
        float sm = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float v = __ldg(fl + c * hw);
            const float dn = (y + 1 < h) ? __ldg(fl + c * hw + w) : 0.f;
            const float rt = (x + 1 < w) ? __ldg(fl + c * hw + 1) : 0.f;
            sm += __fadd_rn(charb(__fsub_rn(v, dn)), charb(__fsub_rn(v, rt)));
        }
        a[P_SMOOTH] += (double)(sm * 0.5f);
    }
#pragma unroll
    for (int k = 0; k < kNumPart; ++k) {
        double v = (k == P_MINW || k == P_MINF) ? warp_min(a[k]) : (k == P_MAXW || k == P_MAXF) ? warp_max(a[k]) : warp_sum(a[k]);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
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

// One CTA: warp s reduces scale s (lane k = slot k, fixed CTA order), then thread 0 forms the terms.
__global__ void __launch_bounds__(32 * OFR_LOSS_MAX_SCALES) loss_final_kernel(const __grid_constant__ LossParams p,
                                                                              const double* __restrict__ partials,
                                                                              double* __restrict__ state,
                                                                              double* __restrict__ out4) {
    const int s = threadIdx.x >> 5, k = threadIdx.x & 31;
    if (s < p.n && k < kNumPart) {
        const int c0 = p.cta_begin[s], c1 = p.cta_begin[s + 1];
        double v = partials[(long long)c0 * kNumPart + k];
        for (int c = c0 + 1; c < c1; ++c) {
            const double t = partials[(long long)c * kNumPart + k];
            v = (k == P_MINW || k == P_MINF) ? fmin(v, t) : (k == P_MAXW || k == P_MAXF) ? fmax(v, t) : v + t;
        }
        state[s * OFR_LOSS_STATE_DOUBLES + k] = v;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    double tp = 0.0, tc = 0.0, ts = 0.0;
    for (int i = 0; i < p.n; ++i) {
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
        tp += wi * photo;
        tc += wi * ncc;
        ts += wi * smooth;
    }
    const double inv_n = 1.0 / (double)p.n;
    const double P = inv_n * p.gamma * tp, C = inv_n * p.zeta * tc, S = inv_n * p.lamb_da * ts;
    out4[0] = P;
    out4[1] = C;
    out4[2] = S;
    out4[3] = P + S + C;  // loss.py:83 adds in this order
}

__global__ void __launch_bounds__(kThreads) loss_bwd_kernel(const __grid_constant__ LossParams p,
                                                            const double* __restrict__ grad4,
                                                            const double* __restrict__ state) {
    const int s = scale_of_cta(p, blockIdx.x);
    const ScaleDev sc = p.sc[s];
    const int cta = blockIdx.x - p.cta_begin[s], nctas = p.cta_begin[s + 1] - p.cta_begin[s];
    const int h = sc.h, w = sc.w;
    const long long hw = (long long)h * w, total = (long long)p.B * hw;
    const Resizer rz(p.Hs, p.Ws, h, w);
    const double* st = state + s * OFR_LOSS_STATE_DOUBLES;

    const double wi = 0.05 * (double)(s + 1), inv_n = 1.0 / (double)p.n, Bd = (double)p.B;
    const double gP = grad4[0] + grad4[3], gC = grad4[1] + grad4[3], gS = grad4[2] + grad4[3];
    const float cp = (float)(gP * inv_n * p.gamma * wi / Bd);          // d total / d photometric-sum
    const float cs = (float)(gS * inv_n * p.lamb_da * wi / (2.0 * Bd));  // d total / d smoothness-sum
    // NCC: d(1 - corr)/dw_i = -(1/B) * [ vy_i / sqrt(sxx*syy) - sxy * vx_i / (sxx^1.5 * syy^0.5) ]
    const double N = st[S_N], mw = st[S_SW] / N, mf = st[S_SF] / N;
    const double sxx = st[S_SWW] - N * mw * mw, syy = st[S_SFF] - N * mf * mf, sxy = st[S_SWF] - N * mw * mf;
    const bool degenerate = st[S_DEGENERATE] != 0.0;
    const double cc = degenerate ? 0.0 : -(gC * inv_n * p.zeta * wi / Bd);
    const double k1 = degenerate ? 0.0 : cc / (sqrt(sxx) * sqrt(syy));
    const double k2 = degenerate ? 0.0 : cc * sxy / (sxx * sqrt(sxx) * sqrt(syy));

    for (long long i = (long long)cta * kThreads + threadIdx.x; i < total; i += (long long)nctas * kThreads) {
        const int b = (int)(i / hw);
        const int rem = (int)(i - (long long)b * hw);
        const int y = rem / w, x = rem - y * w;
        if (sc.g_warped) {
            const float wv = __ldg(sc.warped + i);
            const float fv = rz.at(p.fixed + (long long)b * p.Hs * p.Ws, y, x);
            const float gph = -cp * charb_grad(__fsub_rn(fv, wv));
            const double gn = k1 * ((double)fv - mf) - k2 * ((double)wv - mw);
            sc.g_warped[i] = gph + (float)gn;
        }
        if (sc.g_flow) {
            const float* fl = sc.flow + (long long)b * 2 * hw + rem;
            float* gf = sc.g_flow + (long long)b * 2 * hw + rem;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float* f = fl + c * hw;
                const float v = __ldg(f);
                const float dn = (y + 1 < h) ? __ldg(f + w) : 0.f;
                const float rt = (x + 1 < w) ? __ldg(f + 1) : 0.f;
                float g = charb_grad(__fsub_rn(v, dn)) + charb_grad(__fsub_rn(v, rt));
                if (y > 0) g -= charb_grad(__fsub_rn(__ldg(f - w), v));
                if (x > 0) g -= charb_grad(__fsub_rn(__ldg(f - 1), v));
                gf[c * hw] = cs * g;
            }
        }
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
            p.sc[i] = ScaleDev{s.warped, s.flow, s.g_warped, s.g_flow, s.h, s.w};
            long long pix = (long long)B * s.h * s.w;
            long long nc = (pix + kPixPerCta - 1) / kPixPerCta;
            at += (int)(nc < 1 ? 1 : (nc > kMaxCtasPerScale ? kMaxCtasPerScale : nc));
        } else {
            p.sc[i] = ScaleDev{nullptr, nullptr, nullptr, nullptr, 0, 0};
        }
    }
    p.cta_begin[OFR_LOSS_MAX_SCALES] = at;
    for (int i = n; i <= OFR_LOSS_MAX_SCALES; ++i) p.cta_begin[i] = at;
    return OFR_OK;
}

}  // namespace

extern "C" size_t ofr_loss_workspace(int n_scales) {
    if (n_scales < 1 || n_scales > OFR_LOSS_MAX_SCALES) return 0;
    return ofr::align_up(sizeof(double) * kNumPart * (size_t)kMaxCtasPerScale * n_scales, 256);
}

extern "C" int ofr_loss_fwd(const float* fixed, int B, int Hs, int Ws, const ofr_loss_scale_t* scales, int n_scales,
                            double lamb_da, double gamma, double zeta, double* out4, double* state, void* ws,
                            size_t ws_bytes, ofr_stream_t stream) {
    LossParams p;
    int rc = fill_params(p, fixed, B, Hs, Ws, scales, n_scales, lamb_da, gamma, zeta);
    if (rc != OFR_OK) return rc;
    if (!out4 || !state) return OFR_EINVAL;
    if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255u) || ws_bytes < ofr_loss_workspace(n_scales)) return OFR_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    double* partials = (double*)ws;
    loss_partial_kernel<<<p.cta_begin[n_scales], kThreads, 0, st>>>(p, partials);
    OFR_LAUNCH_CHECK();
    loss_final_kernel<<<1, 32 * OFR_LOSS_MAX_SCALES, 0, st>>>(p, partials, state, out4);
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
