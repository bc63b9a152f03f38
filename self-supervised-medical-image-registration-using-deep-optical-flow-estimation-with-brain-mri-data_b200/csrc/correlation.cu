// Correlation cost volume, forward and backward (replaces correlation_cuda; ctor sites
// flownet2/networks/FlowNetC.py:28,31 and PWC/models/PWCNet.py:69, calls FlowNetC.py:88, PWCNet.py:200-259).
//
//   out[b,(tj+r)*D+(ti+r),y,x] = 1/(k*k*C) * sum_{c,h,w<k} P1[b,c,y1+h,x1+w] * P2[b,c,y1+tj*s2+h, x1+ti*s2+w]
//   P = input zero-padded by `pad`;  y1 = y*s1 + md;  r = md/s2;  D = 2r+1.
//
// Two families of kernels:
//  * tiled (k=1, s1=1 -- every call site of the reference): a CTA owns a TH x TW tile of output pixels of
//    one image and a group of TJB displacement rows.  Per chunk of CK channels it stages the f1 tile and
//    the f2 tile + displacement halo in shared memory, de-interleaved by x mod s2 so that the values one
//    thread needs are contiguous.  A thread owns M pixels along x (stride s2) x all D x-displacements of
//    one displacement row: a Toeplitz register tile acc[i][j] += a[i] * b[i+j] fed by 128-bit LDS, with
//    the tj lanes of a warp sharing (broadcasting) the same `a` vector.  No atomics, no split-K: the
//    channel sum is a fixed-order loop -> bitwise reproducible.
//  * generic (any k, s1, s2, pad): direct kernels, one thread per output element.  Correct for every
//    parameter set the Python module accepts; only reached by configurations the reference never uses.
#include "corr_internal.cuh"

namespace {

using ofr::CorrGeom;

bool make_geom(CorrGeom& g, int B, int C, int H, int W, int pad, int k, int md, int s1, int s2) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || pad < 0 || k < 1 || (k & 1) == 0 || md < 0 || s1 < 1 || s2 < 1)
        return false;
    const int kr = (k - 1) / 2, border = md + kr;
    const int ph = H + 2 * pad - 2 * border, pw = W + 2 * pad - 2 * border;
    if (ph < 1 || pw < 1) return false;
    g.B = B, g.C = C, g.H = H, g.W = W, g.pad = pad, g.k = k, g.md = md, g.s1 = s1, g.s2 = s2;
    g.r = md / s2;
    g.D = 2 * g.r + 1;
    g.Hout = (ph + s1 - 1) / s1;
    g.Wout = (pw + s1 - 1) / s1;
    g.scale = 1.0f / (float)(k * k * C);
    g.slope = 1.0f;
    return true;
}

// ------------------------------------------------------------------------------------------------ generic
__device__ __forceinline__ float rd(const float* __restrict__ img, int y, int x, int H, int W) {
    return (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(img + (long long)y * W + x) : 0.f;
}

__global__ void __launch_bounds__(256) corr_fwd_generic_kernel(const float* __restrict__ f1,
                                                               const float* __restrict__ f2,
                                                               float* __restrict__ out, CorrGeom g) {
    const long long total = (long long)g.B * g.D * g.D * g.Hout * g.Wout;
    const long long hw = (long long)g.H * g.W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long t = i;
        const int x = (int)(t % g.Wout);
        t /= g.Wout;
        const int y = (int)(t % g.Hout);
        t /= g.Hout;
        const int d = (int)(t % (g.D * g.D));
        const int b = (int)(t / (g.D * g.D));
        const int tj = d / g.D - g.r, ti = d % g.D - g.r;
        const int y1 = y * g.s1 + g.md - g.pad, x1 = x * g.s1 + g.md - g.pad;  // unpadded coordinates
        const int y2 = y1 + tj * g.s2, x2 = x1 + ti * g.s2;
        float acc = 0.f;
        for (int c = 0; c < g.C; ++c) {
            const float* a = f1 + ((long long)b * g.C + c) * hw;
            const float* bb = f2 + ((long long)b * g.C + c) * hw;
            for (int h = 0; h < g.k; ++h)
                for (int w = 0; w < g.k; ++w)
                    acc = fmaf(rd(a, y1 + h, x1 + w, g.H, g.W), rd(bb, y2 + h, x2 + w, g.H, g.W), acc);
        }
        out[i] = ofr::corr_out(acc, g);
    }
}

// gI1[b,c,yi,xi] = scale * sum_{h,w} sum_d gO[b,d,y,x] * P2[c, yp + tj*s2, xp + ti*s2]
//   with (y,x) the output pixel whose window element (h,w) is this input pixel: y*s1 + md + h = yi + pad.
// gI2[b,c,yj,xj] = scale * sum_{h,w} sum_d gO[b,d,y,x] * P1[c, yp - tj*s2, xp - ti*s2]
//   with y*s1 + md + tj*s2 + h = yj + pad.
template <bool SECOND>
__global__ void __launch_bounds__(256) corr_bwd_generic_kernel(const float* __restrict__ other,
                                                               const float* __restrict__ gO,
                                                               float* __restrict__ gI, CorrGeom g) {
    const long long hw = (long long)g.H * g.W, ohw = (long long)g.Hout * g.Wout;
    const long long total = (long long)g.B * g.C * hw;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long t = i;
        const int xi = (int)(t % g.W);
        t /= g.W;
        const int yi = (int)(t % g.H);
        t /= g.H;
        const int c = (int)(t % g.C);
        const int b = (int)(t / g.C);
        const float* oth = other + ((long long)b * g.C + c) * hw;
        const float* go = gO + (long long)b * g.D * g.D * ohw;
        float acc = 0.f;
        for (int tj = -g.r; tj <= g.r; ++tj) {
            for (int ti = -g.r; ti <= g.r; ++ti) {
                const float* gd = go + (long long)((tj + g.r) * g.D + (ti + g.r)) * ohw;
                for (int h = 0; h < g.k; ++h) {
                    // padded row of the window origin in image 1
                    const int yo = yi + g.pad - h - g.md - (SECOND ? tj * g.s2 : 0);
                    if (yo < 0 || yo % g.s1) continue;
                    const int y = yo / g.s1;
                    if (y >= g.Hout) continue;
                    for (int w = 0; w < g.k; ++w) {
                        const int xo = xi + g.pad - w - g.md - (SECOND ? ti * g.s2 : 0);
                        if (xo < 0 || xo % g.s1) continue;
                        const int x = xo / g.s1;
                        if (x >= g.Wout) continue;
                        const int yy = SECOND ? yi - tj * g.s2 : yi + tj * g.s2;
                        const int xx = SECOND ? xi - ti * g.s2 : xi + ti * g.s2;
                        acc = fmaf(__ldg(gd + (long long)y * g.Wout + x), rd(oth, yy, xx, g.H, g.W), acc);
                    }
                }
            }
        }
        gI[i] = acc * g.scale;
    }
}

// ------------------------------------------------------------------------------------------------ tiled forward
// Template parameters: S2 stride2, R = md/s2, M pixels per thread, TH tile rows, NS strips per tile row
// (tile width TW = NS*M, NS % S2 == 0), TJB displacement rows per CTA, CK channels per shared-memory chunk.
template <int S2, int R, int M, int TH, int NS, int TJB, int CK>
struct FwdCfg {
    static constexpr int D = 2 * R + 1;
    static constexpr int TW = NS * M;             // tile width in pixels
    static constexpr int U = TW / S2;             // pixels per x-residue class in a tile row
    static constexpr int U2 = U + 2 * R;          // ... plus displacement halo
    // row pitch of an f2 residue row: multiple of 4 (128-bit LDS) and == 4*odd (mod 32) so that the
    // rows read by the tj lanes of a warp start in different bank groups
    static constexpr int pitch2() {
        int p = (U2 + 3) / 4 * 4;
        while ((p / 4) % 2 == 0 || p % 32 == 0) p += 4;
        return p;
    }
    static constexpr int P2 = pitch2();
    static constexpr int R2 = TH + (TJB - 1) * S2;  // f2 rows a CTA needs
    static constexpr int XW2 = TW + 2 * R * S2;     // f2 columns (global x extent) a CTA needs
    static constexpr int F1_FLOATS = CK * TH * S2 * U;
    static constexpr int F2_FLOATS = CK * R2 * S2 * P2;
    static constexpr int THREADS = TJB * NS * TH;
    static constexpr size_t SMEM = sizeof(float) * (F1_FLOATS + F2_FLOATS);
    static_assert(NS % S2 == 0, "strips must cover every x residue class equally");
    static_assert(M % 4 == 0 && U % 4 == 0 && TW % 4 == 0 && XW2 % 4 == 0, "128-bit paths need multiples of 4");
    static_assert((M + 2 * R) % 4 == 0, "b window must be a whole number of float4");
    static_assert(THREADS <= 1024, "CTA too large");
};

template <class Cfg, int S2, int R, int M, int TH, int NS, int TJB, int CK>
__global__ void __launch_bounds__(Cfg::THREADS)
    corr_fwd_tiled_kernel(const float* __restrict__ f1, const float* __restrict__ f2, float* __restrict__ out,
                          CorrGeom g, int off /* md - pad */, int tiles_x, int tiles_y, int tj_groups) {
    constexpr int D = Cfg::D, U = Cfg::U, P2 = Cfg::P2, R2 = Cfg::R2, XW2 = Cfg::XW2, TW = Cfg::TW;
    extern __shared__ __align__(16) float smem[];
    float* f1s = smem;                    // [CK][TH][S2][U]
    float* f2s = smem + Cfg::F1_FLOATS;   // [CK][R2][S2][P2]

    int bid = blockIdx.x;
    const int txi = bid % tiles_x;
    bid /= tiles_x;
    const int tyi = bid % tiles_y;
    bid /= tiles_y;
    const int tjg = bid % tj_groups;
    const int b = bid / tj_groups;

    const int x0 = txi * TW, y0 = tyi * TH;  // output tile origin
    const int tj0 = tjg * TJB;               // first displacement-row index (0..D-1) of this CTA

    const int tid = threadIdx.x;
    const int tjl = tid % TJB;
    const int s = (tid / TJB) % NS;
    const int row = tid / (TJB * NS);
    const int pi = s % S2, us = s / S2;

    float acc[M][D];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) acc[i][j] = 0.f;

    const long long hw = (long long)g.H * g.W;
    const float* f1b = f1 + (long long)b * g.C * hw;
    const float* f2b = f2 + (long long)b * g.C * hw;
    // global origin of the staged tiles (input coordinates)
    const int gy1 = y0 + off, gx1 = x0 + off;
    const int gy2 = y0 + off + (tj0 - R) * S2, gx2 = x0 + off - R * S2;

    for (int c0 = 0; c0 < g.C; c0 += CK) {
        // ---- stage f1 tile: CK x TH x TW, 128-bit global loads, de-interleave by x mod S2
        for (int e = tid; e < CK * TH * (TW / 4); e += Cfg::THREADS) {
            const int q = e % (TW / 4);
            const int rr = (e / (TW / 4)) % TH;
            const int cc = e / ((TW / 4) * TH);
            const int gy = gy1 + rr, gx = gx1 + 4 * q, c = c0 + cc;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < g.C && gy >= 0 && gy < g.H && gx >= 0 && gx + 3 < g.W) {
                v = ofr::ldg4(f1b + (long long)c * hw + (long long)gy * g.W + gx);
            } else if (c < g.C && gy >= 0 && gy < g.H) {  // ragged edge: element-wise
                const float* p = f1b + (long long)c * hw + (long long)gy * g.W;
                v.x = (gx >= 0 && gx < g.W) ? __ldg(p + gx) : 0.f;
                v.y = (gx + 1 >= 0 && gx + 1 < g.W) ? __ldg(p + gx + 1) : 0.f;
                v.z = (gx + 2 >= 0 && gx + 2 < g.W) ? __ldg(p + gx + 2) : 0.f;
                v.w = (gx + 3 >= 0 && gx + 3 < g.W) ? __ldg(p + gx + 3) : 0.f;
            }
            float* dst = f1s + (cc * TH + rr) * S2 * U;
            if (S2 == 1) {
                *reinterpret_cast<float4*>(dst + 4 * q) = v;
            } else if (S2 == 2) {
                *reinterpret_cast<float2*>(dst + 2 * q) = make_float2(v.x, v.z);
                *reinterpret_cast<float2*>(dst + U + 2 * q) = make_float2(v.y, v.w);
            } else {
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) dst[((4 * q + k) % S2) * U + (4 * q + k) / S2] = vv[k];
            }
        }
        // ---- stage f2 tile + halo: CK x R2 x XW2
        for (int e = tid; e < CK * R2 * (XW2 / 4); e += Cfg::THREADS) {
            const int q = e % (XW2 / 4);
            const int rr = (e / (XW2 / 4)) % R2;
            const int cc = e / ((XW2 / 4) * R2);
            const int gy = gy2 + rr, gx = gx2 + 4 * q, c = c0 + cc;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < g.C && gy >= 0 && gy < g.H && gx >= 0 && gx + 3 < g.W) {
                v = ofr::ldg4(f2b + (long long)c * hw + (long long)gy * g.W + gx);
            } else if (c < g.C && gy >= 0 && gy < g.H && gx + 3 >= 0 && gx < g.W) {
                const float* p = f2b + (long long)c * hw + (long long)gy * g.W;
                v.x = (gx >= 0 && gx < g.W) ? __ldg(p + gx) : 0.f;
                v.y = (gx + 1 >= 0 && gx + 1 < g.W) ? __ldg(p + gx + 1) : 0.f;
                v.z = (gx + 2 >= 0 && gx + 2 < g.W) ? __ldg(p + gx + 2) : 0.f;
                v.w = (gx + 3 >= 0 && gx + 3 < g.W) ? __ldg(p + gx + 3) : 0.f;
            }
            float* dst = f2s + (cc * R2 + rr) * S2 * P2;
            if (S2 == 1) {
                *reinterpret_cast<float4*>(dst + 4 * q) = v;
            } else if (S2 == 2) {
                *reinterpret_cast<float2*>(dst + 2 * q) = make_float2(v.x, v.z);
                *reinterpret_cast<float2*>(dst + P2 + 2 * q) = make_float2(v.y, v.w);
            } else {
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) dst[((4 * q + k) % S2) * P2 + (4 * q + k) / S2] = vv[k];
            }
        }
        __syncthreads();

        // ---- Toeplitz register tile over the chunk's channels
        const float* ap = f1s + (row * S2 + pi) * U + us * M;
        const float* bp = f2s + ((row + tjl * S2) * S2 + pi) * P2 + us * M;
#pragma unroll 2
        for (int cc = 0; cc < CK; ++cc) {
            float a[M], bv[M + 2 * R];
#pragma unroll
            for (int i = 0; i < M; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(ap + cc * TH * S2 * U + i);
                a[i] = v.x, a[i + 1] = v.y, a[i + 2] = v.z, a[i + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < M + 2 * R; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(bp + cc * R2 * S2 * P2 + i);
                bv[i] = v.x, bv[i + 1] = v.y, bv[i + 2] = v.z, bv[i + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) acc[i][j] = fmaf(a[i], bv[i + j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue: out[b, tj*D + j, y, x_i]
    const int tj = tj0 + tjl;
    const int y = y0 + row;
    if (tj < D && y < g.Hout) {
        const long long ohw = (long long)g.Hout * g.Wout;
        float* ob = out + ((long long)b * D * D + (long long)tj * D) * ohw + (long long)y * g.Wout;
        const int xb = x0 + pi + S2 * us * M;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            float* o = ob + (long long)j * ohw;
            if (S2 == 1 && xb + M <= g.Wout) {
#pragma unroll
                for (int i = 0; i < M; i += 4)
                    ofr::stcs4(o + xb + i, make_float4(ofr::corr_out(acc[i][j], g), ofr::corr_out(acc[i + 1][j], g),
                                                       ofr::corr_out(acc[i + 2][j], g), ofr::corr_out(acc[i + 3][j], g)));
            } else {
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    const int x = xb + S2 * i;
                    if (x < g.Wout) o[x] = ofr::corr_out(acc[i][j], g);
                }
            }
        }
    }
}

template <int S2, int R, int M, int TH, int NS, int TJB, int CK>
int launch_tiled(const float* f1, const float* f2, float* out, const CorrGeom& g, cudaStream_t st) {
    using Cfg = FwdCfg<S2, R, M, TH, NS, TJB, CK>;
    auto kern = corr_fwd_tiled_kernel<Cfg, S2, R, M, TH, NS, TJB, CK>;
    OFR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    const int tiles_x = (g.Wout + Cfg::TW - 1) / Cfg::TW, tiles_y = (g.Hout + TH - 1) / TH;
    const int tj_groups = (Cfg::D + TJB - 1) / TJB;
    const long long nblk = (long long)tiles_x * tiles_y * tj_groups * g.B;
    if (nblk >= (1ll << 31)) return OFR_ENOTSUP;
    kern<<<(unsigned)nblk, Cfg::THREADS, Cfg::SMEM, st>>>(f1, f2, out, g, g.md - g.pad, tiles_x, tiles_y, tj_groups);
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

// ------------------------------------------------------------------------------------------------ polyphase split
// stride2 = 2 (FlowNetC): pixel (y,x) only meets f2 pixels (y + 2tj, x + 2ti) of its own parity class, so the operator
// factors into four independent stride-1 correlations (R = md/2) on the class images [B][4][C][H/2][W/2].  The split /
// merge passes are pure streaming copies (128-bit loads, 64-bit stores) and cost a few percent of the FMA-bound
// correlation they enable; they replace upstream's NHWC repack of both inputs.
__global__ void __launch_bounds__(256) polyphase_split_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                              int B, int C, int H, int W) {
    const int w4 = W / 4, h2 = H / 2, w2 = W / 2;
    const long long total = (long long)B * C * H * w4;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        long long t = i;
        const int q = (int)(t % w4);
        t /= w4;
        const int y = (int)(t % H);
        t /= H;
        const int c = (int)(t % C);
        const int b = (int)(t / C);
        const float4 v = ofr::ldg4(src + (((long long)b * C + c) * H + y) * W + 4 * q);
        const int py = y & 1;
        float* d0 = dst + ((((long long)b * 4 + py * 2 + 0) * C + c) * h2 + (y >> 1)) * w2 + 2 * q;
        float* d1 = dst + ((((long long)b * 4 + py * 2 + 1) * C + c) * h2 + (y >> 1)) * w2 + 2 * q;
        *reinterpret_cast<float2*>(d0) = make_float2(v.x, v.z);
        *reinterpret_cast<float2*>(d1) = make_float2(v.y, v.w);
    }
}

__global__ void __launch_bounds__(256) polyphase_merge_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                              int B, int C, int H, int W) {
    const int w4 = W / 4, h2 = H / 2, w2 = W / 2;
    const long long total = (long long)B * C * H * w4;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        long long t = i;
        const int q = (int)(t % w4);
        t /= w4;
        const int y = (int)(t % H);
        t /= H;
        const int c = (int)(t % C);
        const int b = (int)(t / C);
        const int py = y & 1;
        const float2 e = __ldg(reinterpret_cast<const float2*>(
            src + ((((long long)b * 4 + py * 2 + 0) * C + c) * h2 + (y >> 1)) * w2 + 2 * q));
        const float2 o = __ldg(reinterpret_cast<const float2*>(
            src + ((((long long)b * 4 + py * 2 + 1) * C + c) * h2 + (y >> 1)) * w2 + 2 * q));
        ofr::stcs4(dst + (((long long)b * C + c) * H + y) * W + 4 * q, make_float4(e.x, o.x, e.y, o.y));
    }
}

int polyphase(bool split, const float* src, float* dst, int B, int C, int H, int W, cudaStream_t st) {
    const long long n = (long long)B * C * H * (W / 4);
    if (split)
        polyphase_split_kernel<<<ofr::grid_for(n, 256, 8), 256, 0, st>>>(src, dst, B, C, H, W);
    else
        polyphase_merge_kernel<<<ofr::grid_for(n, 256, 8), 256, 0, st>>>(src, dst, B, C, H, W);
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

// Can this geometry run as four stride-1 correlations on class images?  (instantiated for R = 10 only)
bool polyphase_ok(const CorrGeom& g) {
    return g.k == 1 && g.s1 == 1 && g.s2 == 2 && g.pad == g.md && g.md == 20 && (g.H % 2) == 0 && (g.W % 8) == 0;
}
CorrGeom class_geom(const CorrGeom& g) {
    CorrGeom c = g;
    c.B = g.B * 4, c.H = g.H / 2, c.W = g.W / 2;
    c.pad = c.md = g.md / 2, c.s2 = 1;
    c.r = c.md, c.D = g.D, c.Hout = c.H, c.Wout = c.W;
    return c;
}
size_t poly_ws_bytes(const CorrGeom& g, bool backward) {
    const size_t feat = ofr::align_up(sizeof(float) * (size_t)g.B * g.C * g.H * g.W, 256);
    const size_t vol = ofr::align_up(sizeof(float) * (size_t)g.B * g.D * g.D * g.H * g.W, 256);
    return backward ? 4 * feat + vol : 2 * feat + vol;
}

}  // namespace

extern "C" int ofr_corr_out_shape(int H, int W, int pad, int k, int md, int s1, int s2, int* D, int* Hout,
                                  int* Wout) {
    CorrGeom g;
    if (!make_geom(g, 1, 1, H, W, pad, k, md, s1, s2)) return OFR_EINVAL;
    if (D) *D = g.D;
    if (Hout) *Hout = g.Hout;
    if (Wout) *Wout = g.Wout;
    return OFR_OK;
}

extern "C" size_t ofr_corr_workspace(int B, int C, int H, int W, int pad, int k, int md, int s1, int s2, int backward) {
    CorrGeom g;
    if (!make_geom(g, B, C, H, W, pad, k, md, s1, s2) || !polyphase_ok(g)) return 0;
    if (ofr::corr_umma_applicable(class_geom(g), backward != 0)) return 0;  // tensor-core path: everything in place
    return poly_ws_bytes(g, backward != 0);
}

extern "C" int ofr_corr_fwd(const float* f1, const float* f2, float* out, int B, int C, int H, int W, int pad, int k,
                            int md, int s1, int s2, void* ws, size_t ws_bytes, ofr_stream_t stream) {
    return ofr_corr_fwd_act(f1, f2, out, B, C, H, W, pad, k, md, s1, s2, 1.0f, ws, ws_bytes, stream);
}

extern "C" int ofr_corr_fwd_act(const float* f1, const float* f2, float* out, int B, int C, int H, int W, int pad,
                                int k, int md, int s1, int s2, float negative_slope, void* ws, size_t ws_bytes,
                                ofr_stream_t stream) {
    CorrGeom g;
    if (!f1 || !f2 || !out || !make_geom(g, B, C, H, W, pad, k, md, s1, s2)) return OFR_EINVAL;
    if (!(negative_slope >= 0.f)) return OFR_EINVAL;
    g.slope = negative_slope;
    cudaStream_t st = (cudaStream_t)stream;
    if (polyphase_ok(g)) {  // tensor-core kernel: reads f1 / f2 and writes the cost volume in place (stride 2), no scratch
        const int rc = ofr::corr_fwd_umma(f1, f2, out, class_geom(g), st, true, true);
        if (rc != OFR_ENOTSUP) return rc;
    }
    if (polyphase_ok(g) && ws && !(reinterpret_cast<uintptr_t>(ws) & 255u) && ws_bytes >= poly_ws_bytes(g, false) &&
        ofr::aligned16(f1) && ofr::aligned16(f2) && ofr::aligned16(out)) {
        // FlowNetC: Correlation(20,1,20,1,2) (FlowNetC.py:31) = 4 x stride-1 correlation (R = 10) on the class images
        const size_t feat = ofr::align_up(sizeof(float) * (size_t)B * C * H * W, 256);
        float* f1p = (float*)ws;
        float* f2p = (float*)((char*)ws + feat);
        float* outp = (float*)((char*)ws + 2 * feat);
        const CorrGeom cg = class_geom(g);
        int rc = polyphase(true, f1, f1p, B, C, H, W, st);
        if (rc == OFR_OK) rc = polyphase(true, f2, f2p, B, C, H, W, st);
        if (rc != OFR_OK) return rc;
        rc = ofr::corr_fwd_tma(f1p, f2p, outp, cg, st);
        if (rc == OFR_OK) return polyphase(false, outp, out, B, g.D * g.D, H, W, st);
        if (rc != OFR_ENOTSUP) return rc;
    }
    const bool vec_ok = k == 1 && s1 == 1 && (W % 4 == 0) && ((md - pad) % 4 == 0) && ((g.r * s2) % 4 == 0) &&
                        ofr::aligned16(f1) && ofr::aligned16(f2) && ofr::aligned16(out) && (g.Wout % 4 == 0);
    {  // PWC-Net: Correlation(4,1,4,1,1) (PWCNet.py:69) -> TMA-staged kernels
        const int rc = ofr::corr_fwd_tma(f1, f2, out, g, st);
        if (rc != OFR_ENOTSUP) return rc;
    }
    if (vec_ok && s2 == 1 && g.r == 4) {  // same geometry without TMA (misaligned bases): register-tiled kernel
        if (g.Wout >= 64) return launch_tiled<1, 4, 8, 4, 8, 9, 8>(f1, f2, out, g, st);   // 4 x 64 tile
        if (g.Wout >= 32) return launch_tiled<1, 4, 8, 4, 4, 9, 8>(f1, f2, out, g, st);   // 4 x 32 tile
        if (g.Wout >= 16) return launch_tiled<1, 4, 8, 4, 2, 9, 8>(f1, f2, out, g, st);   // 4 x 16 tile
        return launch_tiled<1, 4, 4, 4, 2, 9, 8>(f1, f2, out, g, st);                     // 4 x 8 tile
    }
    if (vec_ok && s2 == 2 && g.r == 10) {  // FlowNetC: Correlation(20,1,20,1,2)  (FlowNetC.py:31)
        return launch_tiled<2, 10, 4, 8, 8, 3, 16>(f1, f2, out, g, st);                   // 8 x 32 tile, 3 tj rows
    }
    const long long total = (long long)B * g.D * g.D * g.Hout * g.Wout;
    corr_fwd_generic_kernel<<<ofr::grid_for(total, 256, 8), 256, 0, st>>>(f1, f2, out, g);
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

extern "C" int ofr_corr_bwd(const float* f1, const float* f2, const float* gO, float* gI1, float* gI2, int B, int C,
                            int H, int W, int pad, int k, int md, int s1, int s2, void* ws, size_t ws_bytes,
                            ofr_stream_t stream) {
    CorrGeom g;
    if (!f1 || !f2 || !gO || !make_geom(g, B, C, H, W, pad, k, md, s1, s2)) return OFR_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (!gI1 && !gI2) return OFR_OK;
    if (polyphase_ok(g)) {  // tensor-core kernel: f1 / f2 / gO read and both gradients written in place, no scratch
        const int rc = ofr::corr_bwd_umma(f1, f2, gO, gI1, gI2, class_geom(g), st, true, true);
        if (rc != OFR_ENOTSUP) return rc;
    }
    if (polyphase_ok(g) && ws && !(reinterpret_cast<uintptr_t>(ws) & 255u) && ws_bytes >= poly_ws_bytes(g, true) &&
        ofr::aligned16(f1) && ofr::aligned16(f2) && ofr::aligned16(gO) && (!gI1 || ofr::aligned16(gI1)) &&
        (!gI2 || ofr::aligned16(gI2))) {
        const size_t feat = ofr::align_up(sizeof(float) * (size_t)B * C * H * W, 256);
        const size_t vol = ofr::align_up(sizeof(float) * (size_t)B * g.D * g.D * H * W, 256);
        char* w = (char*)ws;
        float* f1p = (float*)w;
        float* f2p = (float*)(w + feat);
        float* gOp = (float*)(w + 2 * feat);
        float* g1p = (float*)(w + 2 * feat + vol);
        float* g2p = (float*)(w + 3 * feat + vol);
        const CorrGeom cg = class_geom(g);
        int rc = OFR_OK;
        if (gI2) rc = polyphase(true, f1, f1p, B, C, H, W, st);
        if (rc == OFR_OK && gI1) rc = polyphase(true, f2, f2p, B, C, H, W, st);
        if (rc == OFR_OK) rc = polyphase(true, gO, gOp, B, g.D * g.D, H, W, st);
        if (rc != OFR_OK) return rc;
        rc = ofr::corr_bwd_tma(f1p, f2p, gOp, gI1 ? g1p : nullptr, gI2 ? g2p : nullptr, cg, st);
        if (rc == OFR_OK) {
            if (gI1) rc = polyphase(false, g1p, gI1, B, C, H, W, st);
            if (rc == OFR_OK && gI2) rc = polyphase(false, g2p, gI2, B, C, H, W, st);
            return rc;
        }
        if (rc != OFR_ENOTSUP) return rc;
    }
    {
        const int rc = ofr::corr_bwd_tma(f1, f2, gO, gI1, gI2, g, st);
        if (rc != OFR_ENOTSUP) return rc;
    }
    const long long total = (long long)B * C * H * W;
    if (gI1) {
        corr_bwd_generic_kernel<false><<<ofr::grid_for(total, 256, 8), 256, 0, st>>>(f2, gO, gI1, g);
        OFR_LAUNCH_CHECK();
    }
    if (gI2) {
        corr_bwd_generic_kernel<true><<<ofr::grid_for(total, 256, 8), 256, 0, st>>>(f1, gO, gI2, g);
        OFR_LAUNCH_CHECK();
    }
    return OFR_OK;
}
