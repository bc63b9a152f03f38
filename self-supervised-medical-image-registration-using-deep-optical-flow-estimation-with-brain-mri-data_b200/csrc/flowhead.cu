// Flow heads of the PWC-Net decoder (decoder-side ops adjacent to the cost volume), sm_100a.
//
//   predict_flow  = nn.Conv2d(C, 2, kernel_size=3, stride=1, padding=1)             PWC/models/PWCNet.py:35-36, 79..123
//   deconv/upfeat = nn.ConvTranspose2d(C, 2, kernel_size=4, stride=2, padding=1)    PWC/models/PWCNet.py:38-39, 80..125
//
// Both map hundreds of channels (C = 529..661 behind the DenseNet blocks, 32 for dc_conv7, 2 for the flow upsamplers)
// to the TWO channels of a flow field.  A GEMM-shaped convolution kernel has nothing to tile over with N = 2: cuDNN
// takes 0.8 ms forward / 1.9 ms backward for predict_flow2 at B = 24, 128 x 128 (an 892 MB input: 137 us at the HBM
// peak).  Here the two output channels live in registers and the operator is what it looks like, a streaming
// reduction over channels: a thread owns a 2-row x 4-pixel patch of the map, walks the channels of its CTA's channel
// range, and does 144 (3x3) / 256 (4x4 transposed) FMAs per channel on a 4 x 6 neighbourhood it loads as four 128-bit
// vectors plus the two columns beside them.  fp32 FMA throughout (cuDNN's default for these layers is TF32).
//
//   forward       : channel ranges are split over blockIdx.y when the map alone has too few patches to fill the GPU;
//                   the partial sums go to the workspace and a second kernel adds them in split order (+ bias).
//   input gradient: same patch; the 4 x 6 x 2 gradient neighbourhood is loaded once and every channel is 144 / 256 FMAs
//                   and two 128-bit stores.
//   weight gradient: a thread keeps the 18 / 32 sums of TWO channels in registers and walks a strided subset of the
//                   patches; the CTA adds its threads in a fixed tree, CTAs write partials, a second kernel adds them in
//                   order.  No atomics anywhere: bitwise reproducible.
#include "ofr_common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kWPad = 20;   // 18 weights of one (channel, 3x3, 2 outputs) padded to five 128-bit words
constexpr int kTPad = 32;   // 32 weights of one (channel, 4x4, 2 outputs)
constexpr int kMaxCper = 160;  // channels per CTA (weights staged in shared memory)

struct Geo {
    int B, C, H, W;  // input map of the convolution (= output map for the 3x3; the transposed conv writes 2H x 2W)
    int rp, gx;      // row pairs per image, 4-pixel groups per row
    long long items; // B * rp * gx patches
};

__host__ __device__ inline Geo make_geo(int B, int C, int H, int W) {
    Geo g;
    g.B = B, g.C = C, g.H = H, g.W = W;
    g.rp = (H + 1) / 2, g.gx = W / 4;
    g.items = (long long)B * g.rp * g.gx;
    return g;
}

struct Patch {
    int b, y0, x0;
    bool rv[4];   // rows y0-1 .. y0+2 inside the map
    bool lv, cv;  // column x0-1 / x0+4 inside the map
};

__device__ __forceinline__ Patch patch_of(const Geo& g, long long item) {
    Patch p;
    const int gi = (int)(item % g.gx);
    const long long t = item / g.gx;
    const int r = (int)(t % g.rp);
    p.b = (int)(t / g.rp);
    p.y0 = 2 * r, p.x0 = 4 * gi;
#pragma unroll
    for (int i = 0; i < 4; ++i) p.rv[i] = (unsigned)(p.y0 - 1 + i) < (unsigned)g.H;
    p.lv = p.x0 > 0, p.cv = p.x0 + 4 < g.W;
    return p;
}

// 4 rows x 6 columns around the patch from one plane (zero outside the map)
__device__ __forceinline__ void load_patch(const float* __restrict__ plane, const Geo& g, const Patch& p, float (&v)[4][6]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float* row = plane + (long long)(p.y0 - 1 + i) * g.W + p.x0;
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        float l = 0.f, r = 0.f;
        if (p.rv[i]) {
            c = ofr::ldg4(row);
            if (p.lv) l = __ldg(row - 1);
            if (p.cv) r = __ldg(row + 4);
        }
        v[i][0] = l, v[i][1] = c.x, v[i][2] = c.y, v[i][3] = c.z, v[i][4] = c.w, v[i][5] = r;
    }
}

// ------------------------------------------------------------------------------------------------ 3x3, C -> 2
// shared-memory weights of the CTA's channel range: ws[c][o*9 + ky*3 + kx]
__device__ __forceinline__ void stage_w3(float* ws, const float* __restrict__ w, int C, int c0, int nc) {
    for (int i = threadIdx.x; i < nc * 18; i += blockDim.x) {
        const int c = i / 18, j = i - c * 18, o = j / 9, k = j - o * 9;
        ws[c * kWPad + j] = __ldg(w + ((long long)o * C + c0 + c) * 9 + k);
    }
}

// (95 / 149 registers for this kernel and deconv4_fwd_kernel; capped at 80 / 128 for six / four CTAs per SM both get slower:
// predict_flow2 forward 338 -> 399 us, upfeat3 forward 172 -> 230 us)
__global__ void __launch_bounds__(kThreads) conv3_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ out,
                                                             Geo g, int cper, int direct) {
    __shared__ __align__(16) float ws[kMaxCper * kWPad];
    const int c0 = blockIdx.y * cper, nc = min(cper, g.C - c0);
    const long long item = (long long)blockIdx.x * kThreads + threadIdx.x;
    const bool valid = item < g.items;
    const Patch p = patch_of(g, valid ? item : 0);
    const long long hw = (long long)g.H * g.W;
    const float* plane = x + ((long long)p.b * g.C + c0) * hw;
    // the first channel's loads are in flight while the weights are staged; afterwards channel c+1 is loaded before
    // the arithmetic of channel c (two register buffers, loop unrolled by two so that no copy is needed)
    float va[4][6], vb[4][6];
    if (valid) load_patch(plane, g, p, va);
    stage_w3(ws, w, g.C, c0, nc);
    __syncthreads();
    if (!valid) return;
    float acc[2][2][4];
#pragma unroll
    for (int o = 0; o < 2; ++o)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[o][r][i] = 0.f;
    auto fma_channel = [&](const float (&v)[4][6], int c) {
        float wv[kWPad];
#pragma unroll
        for (int i = 0; i < kWPad; i += 4) {
            const float4 t = *reinterpret_cast<const float4*>(ws + c * kWPad + i);
            wv[i] = t.x, wv[i + 1] = t.y, wv[i + 2] = t.z, wv[i + 3] = t.w;
        }
#pragma unroll
        for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[o][r][i] = fmaf(wv[o * 9 + ky * 3 + kx], v[r + ky][i + kx], acc[o][r][i]);
    };
    for (int c = 0; c < nc; c += 2) {
        if (c + 1 < nc) load_patch(plane + (long long)(c + 1) * hw, g, p, vb);
        fma_channel(va, c);
        if (c + 1 < nc) {
            if (c + 2 < nc) load_patch(plane + (long long)(c + 2) * hw, g, p, va);
            fma_channel(vb, c + 1);
        }
    }
    // direct: the only channel split -> finished values (+ bias); else partial sums [split][B][2][H][W]
    float* dst = out + ((direct ? 0 : (long long)blockIdx.y * g.B) + p.b) * 2 * hw;
#pragma unroll
    for (int o = 0; o < 2; ++o) {
        const float bo = direct && bias ? __ldg(bias + o) : 0.f;
#pragma unroll
        for (int r = 0; r < 2; ++r)
            if (p.y0 + r < g.H)
                *reinterpret_cast<float4*>(dst + o * hw + (long long)(p.y0 + r) * g.W + p.x0) =
                    make_float4(acc[o][r][0] + bo, acc[o][r][1] + bo, acc[o][r][2] + bo, acc[o][r][3] + bo);
    }
}

// out[b][o][.] = bias[o] + sum over splits (ascending) of part[s][b][o][.]
__global__ void split_sum_kernel(const float* __restrict__ part, const float* __restrict__ bias, float* __restrict__ out,
                                 long long n4, long long plane4, int ks) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const int o = (int)((i / plane4) & 1);
    const float bo = bias ? __ldg(bias + o) : 0.f;
    float4 s = make_float4(bo, bo, bo, bo);
#pragma unroll 4
    for (int k = 0; k < ks; ++k) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(part) + (long long)k * n4 + i);
        s.x += t.x, s.y += t.y, s.z += t.z, s.w += t.w;
    }
    reinterpret_cast<float4*>(out)[i] = s;
}

// gX[b][c][y][x] = sum_o sum_k w[o][c][ky][kx] * gO[b][o][y + 1 - ky][x + 1 - kx]
__global__ void __launch_bounds__(kThreads) conv3_bwd_x_kernel(const float* __restrict__ gO, const float* __restrict__ w,
                                                               float* __restrict__ gX, Geo g, int cper) {
    __shared__ __align__(16) float ws[kMaxCper * kWPad];
    const int c0 = blockIdx.y * cper, nc = min(cper, g.C - c0);
    const long long item = (long long)blockIdx.x * kThreads + threadIdx.x;
    const bool valid = item < g.items;
    const Patch p = patch_of(g, valid ? item : 0);
    const long long hw = (long long)g.H * g.W;
    float gv[2][4][6];  // in flight while the weights are staged
    if (valid)
#pragma unroll
        for (int o = 0; o < 2; ++o) load_patch(gO + ((long long)p.b * 2 + o) * hw, g, p, gv[o]);
    stage_w3(ws, w, g.C, c0, nc);
    __syncthreads();
    if (!valid) return;
    float* dst = gX + ((long long)p.b * g.C + c0) * hw + (long long)p.y0 * g.W + p.x0;
    const bool r1 = p.y0 + 1 < g.H;
    for (int c = 0; c < nc; ++c, dst += hw) {
        float wv[kWPad];
#pragma unroll
        for (int i = 0; i < kWPad; i += 4) {
            const float4 t = *reinterpret_cast<const float4*>(ws + c * kWPad + i);
            wv[i] = t.x, wv[i + 1] = t.y, wv[i + 2] = t.z, wv[i + 3] = t.w;
        }
        float acc[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[r][i] = 0.f;
#pragma unroll
        for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            acc[r][i] = fmaf(wv[o * 9 + ky * 3 + kx], gv[o][r + 2 - ky][i + 2 - kx], acc[r][i]);
        ofr::stcs4(dst, make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]));
        if (r1) ofr::stcs4(dst + g.W, make_float4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]));
    }
}

// warp-wide sum of N per-lane values in a fixed butterfly; every lane ends with the total
template <int N>
__device__ __forceinline__ void warp_sum(float (&v)[N]) {
#pragma unroll
    for (int k = 0; k < N; ++k)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], d);
}

// one row of the 4 x 6 neighbourhood: 4 pixels at x0 plus the columns beside them (zero outside the map)
__device__ __forceinline__ void load_row6(const float* __restrict__ plane, int y, int x0, int H, int W, bool lv, bool cv, float (&v)[6]) {
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    float l = 0.f, r = 0.f;
    if ((unsigned)y < (unsigned)H) {
        const float* row = plane + (long long)y * W + x0;
        c = ofr::ldg4(row);
        if (lv) l = __ldg(row - 1);
        if (cv) r = __ldg(row + 4);
    }
    v[0] = l, v[1] = c.x, v[2] = c.y, v[3] = c.z, v[4] = c.w, v[5] = r;
}

// row pairs a lane walks down in the 3x3 weight-gradient kernel: a quarter of the map's, between 2 and 16
__host__ __device__ inline int walk_rows(int rp) { return rp >= 64 ? 16 : rp >= 8 ? rp / 4 : rp >= 2 ? 2 : 1; }

// partial[split][c][o*9 + k] = sum over the split's patches of gO[o][p] * x[c][p + k - 1].  A WARP owns a channel pair;
// its lanes own 32 neighbouring 4-pixel columns and walk walk_rows() row pairs DOWN them, so of the four rows a patch needs
// two are still in registers from the patch above (x is read once, 16 instead of 28 loads per 288 FMAs), and the rows
// of the next patch are loaded before the arithmetic of this one.  The four warps of a CTA own four pairs and walk the
// SAME columns: the gradient rows they all need come out of the L1 after the first warp has fetched them.  Pair 0 also
// sums gO itself (bias gradient) into pbias[split][2].
// (166 registers; capped at 128 for four CTAs per SM it spills and the backward of predict_flow2 goes from 556 to 781 us)
__global__ void __launch_bounds__(kThreads) conv3_bwd_w_kernel(const float* __restrict__ x, const float* __restrict__ gO,
                                                               float* __restrict__ partial, float* __restrict__ pbias,
                                                               Geo g, int splits) {
    const int lane = threadIdx.x & 31, pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int c0 = pair * 2;
    if (c0 >= g.C) return;  // whole warp
    const bool two = c0 + 1 < g.C;
    const long long hw = (long long)g.H * g.W;
    const int wrows = walk_rows(g.rp), nrb = (g.rp + wrows - 1) / wrows;
    const long long ncols = (long long)g.B * nrb * g.gx;
    float acc[38];  // [ch][o][ky][kx] (36) + the two bias sums
#pragma unroll
    for (int k = 0; k < 38; ++k) acc[k] = 0.f;
    for (long long col = (long long)blockIdx.y * 32 + lane; col < ncols; col += (long long)splits * 32) {
        const int gi = (int)(col % g.gx);
        const long long t = col / g.gx;
        const int rb = (int)(t % nrb), b = (int)(t / nrb);
        const int r0 = rb * wrows, r1 = min(g.rp, r0 + wrows), x0 = 4 * gi;
        const bool lv = x0 > 0, cv = x0 + 4 < g.W;
        const float* xa = x + ((long long)b * g.C + c0) * hw;
        const float* xb = two ? xa + hw : xa;
        const float* ga = gO + (long long)b * 2 * hw;
        float v[2][4][6], gv[2][2][4];
        auto load_g = [&](int y0, float (&d)[2][2][4]) {
#pragma unroll
            for (int o = 0; o < 2; ++o)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (y0 + r < g.H) q = ofr::ldg4(ga + o * hw + (long long)(y0 + r) * g.W + x0);
                    d[o][r][0] = q.x, d[o][r][1] = q.y, d[o][r][2] = q.z, d[o][r][3] = q.w;
                }
        };
        // rows 2*r0 - 1 .. 2*r0 + 2 of both channels and the first gradient rows
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            load_row6(xa, 2 * r0 - 1 + i, x0, g.H, g.W, lv, cv, v[0][i]);
            load_row6(xb, 2 * r0 - 1 + i, x0, g.H, g.W, lv, cv, v[1][i]);
        }
        load_g(2 * r0, gv);
        for (int r = r0; r < r1; ++r) {
            // the next patch's two new rows and gradient rows, in flight during this patch's arithmetic
            float nv[2][2][6], ng[2][2][4];
            const bool more = r + 1 < r1;
            if (more) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    load_row6(xa, 2 * r + 3 + i, x0, g.H, g.W, lv, cv, nv[0][i]);
                    load_row6(xb, 2 * r + 3 + i, x0, g.H, g.W, lv, cv, nv[1][i]);
                }
                load_g(2 * r + 2, ng);
            }
            if (pair == 0) {
#pragma unroll
                for (int o = 0; o < 2; ++o)
#pragma unroll
                    for (int q = 0; q < 2; ++q)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[36 + o] += gv[o][q][i];
            }
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                if (ch == 1 && !two) break;
#pragma unroll
                for (int o = 0; o < 2; ++o)
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                            for (int q = 0; q < 2; ++q)
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    acc[ch * 18 + o * 9 + ky * 3 + kx] = fmaf(gv[o][q][i], v[ch][q + ky][i + kx], acc[ch * 18 + o * 9 + ky * 3 + kx]);
            }
            if (more) {
#pragma unroll
                for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        v[ch][0][j] = v[ch][2][j], v[ch][1][j] = v[ch][3][j];
                        v[ch][2][j] = nv[ch][0][j], v[ch][3][j] = nv[ch][1][j];
                    }
#pragma unroll
                for (int o = 0; o < 2; ++o)
#pragma unroll
                    for (int q = 0; q < 2; ++q)
#pragma unroll
                        for (int i = 0; i < 4; ++i) gv[o][q][i] = ng[o][q][i];
            }
        }
    }
    warp_sum<38>(acc);
    if (lane == 0) {
        float* dst = partial + ((long long)blockIdx.y * g.C + c0) * 18;
#pragma unroll
        for (int k = 0; k < 18; ++k) dst[k] = acc[k];
        if (two)
#pragma unroll
            for (int k = 0; k < 18; ++k) dst[18 + k] = acc[18 + k];
        if (pair == 0) pbias[blockIdx.y * 2] = acc[36], pbias[blockIdx.y * 2 + 1] = acc[37];
    }
}

// gW[o][c][k] = sum over the splits of partial[s][c][o*K + k];  gB[o] = sum of pbias[s][o].  A warp per output element:
// lane l adds splits l, l+32, ... in ascending order, then a fixed butterfly -- the order is a function of the shapes only.
__global__ void wgrad_sum_kernel(const float* __restrict__ partial, const float* __restrict__ pbias, float* __restrict__ gW,
                                 float* __restrict__ gB, int C, int K, int splits, int transposed) {
    const int lane = threadIdx.x & 31;
    const int i = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    const int n = C * 2 * K;
    if (i >= n + 2) return;  // whole warp
    const bool is_bias = i >= n;
    if (is_bias && !gB) return;
    const float* src = is_bias ? pbias + (i - n) : partial + i;
    const long long stride = is_bias ? 2 : (long long)n;
    float s = 0.f;
#pragma unroll 4
    for (int q = lane; q < splits; q += 32) s += __ldg(src + q * stride);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane != 0) return;
    if (is_bias) {
        gB[i - n] = s;
    } else {
        const int c = i / (2 * K), j = i - c * 2 * K, o = j / K, k = j - o * K;
        // Conv2d weight [2][C][K]; ConvTranspose2d weight [C][2][K]
        gW[transposed ? ((long long)c * 2 + o) * K + k : ((long long)o * C + c) * K + k] = s;
    }
}

// ------------------------------------------------------------------------------------------------ 4x4 stride 2 transposed, C -> 2
// out[b][o][2y - 1 + ky][2x - 1 + kx] += x[b][c][y][x] * w[c][o][ky][kx].  Output row 2yy + py takes (input row, ky) =
// (yy, 1), (yy - 1, 3) for py = 0 and (yy, 2), (yy + 1, 0) for py = 1; columns alike.
__device__ __forceinline__ void stage_w4(float* ws, const float* __restrict__ w, int c0, int nc) {
    for (int i = threadIdx.x; i < nc * 32; i += blockDim.x) ws[i] = __ldg(w + (long long)c0 * 32 + i);  // [c][o][ky][kx] as stored
}
__device__ __forceinline__ constexpr int tap_k(int parity, int t) { return parity == 0 ? (t == 0 ? 1 : 3) : (t == 0 ? 2 : 0); }
__device__ __forceinline__ constexpr int tap_d(int parity, int t) { return t == 0 ? 0 : (parity == 0 ? -1 : 1); }

__global__ void __launch_bounds__(kThreads) deconv4_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ bias, float* __restrict__ out,
                                                               Geo g, int cper, int direct) {
    __shared__ __align__(16) float ws[kMaxCper * kTPad];
    const int c0 = blockIdx.y * cper, nc = min(cper, g.C - c0);
    const long long item = (long long)blockIdx.x * kThreads + threadIdx.x;
    const bool valid = item < g.items;
    const Patch p = patch_of(g, valid ? item : 0);
    const long long hw = (long long)g.H * g.W;
    const float* plane = x + ((long long)p.b * g.C + c0) * hw;
    float va[4][6], vb[4][6];  // two-stage prefetch as in conv3_fwd_kernel
    if (valid) load_patch(plane, g, p, va);
    stage_w4(ws, w, c0, nc);
    __syncthreads();
    if (!valid) return;
    float acc[2][2][2][4][2];  // [o][r][py][i][px]
#pragma unroll
    for (int k = 0; k < 64; ++k) (&acc[0][0][0][0][0])[k] = 0.f;
    auto fma_channel = [&](const float (&v)[4][6], int c) {
        float wv[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            const float4 t = *reinterpret_cast<const float4*>(ws + c * kTPad + i);
            wv[i] = t.x, wv[i + 1] = t.y, wv[i + 2] = t.z, wv[i + 3] = t.w;
        }
#pragma unroll
        for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int py = 0; py < 2; ++py)
#pragma unroll
                    for (int ty = 0; ty < 2; ++ty)
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int px = 0; px < 2; ++px)
#pragma unroll
                                for (int tx = 0; tx < 2; ++tx)
                                    acc[o][r][py][i][px] = fmaf(v[r + 1 + tap_d(py, ty)][i + 1 + tap_d(px, tx)],
                                                                wv[o * 16 + tap_k(py, ty) * 4 + tap_k(px, tx)], acc[o][r][py][i][px]);
    };
    for (int c = 0; c < nc; c += 2) {
        if (c + 1 < nc) load_patch(plane + (long long)(c + 1) * hw, g, p, vb);
        fma_channel(va, c);
        if (c + 1 < nc) {
            if (c + 2 < nc) load_patch(plane + (long long)(c + 2) * hw, g, p, va);
            fma_channel(vb, c + 1);
        }
    }
    const long long hw2 = 4 * hw;
    const int W2 = 2 * g.W;
    float* dst = out + ((direct ? 0 : (long long)blockIdx.y * g.B) + p.b) * 2 * hw2;
#pragma unroll
    for (int o = 0; o < 2; ++o) {
        const float bo = direct && bias ? __ldg(bias + o) : 0.f;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (p.y0 + r >= g.H) continue;
#pragma unroll
            for (int py = 0; py < 2; ++py) {
                float* row = dst + o * hw2 + (long long)(2 * (p.y0 + r) + py) * W2 + 2 * p.x0;
                const float(&a)[4][2] = acc[o][r][py];
                *reinterpret_cast<float4*>(row) = make_float4(a[0][0] + bo, a[0][1] + bo, a[1][0] + bo, a[1][1] + bo);
                *reinterpret_cast<float4*>(row + 4) = make_float4(a[2][0] + bo, a[2][1] + bo, a[3][0] + bo, a[3][1] + bo);
            }
        }
    }
}

// one input row x 4 pixels per thread for the two backward kernels: the 4 x 10 x 2 neighbourhood of the output gradient
struct RowPatch {
    int b, y, x0;
};
__device__ __forceinline__ RowPatch rowpatch_of(const Geo& g, long long item) {
    RowPatch p;
    const int gi = (int)(item % g.gx);
    const long long t = item / g.gx;
    p.y = (int)(t % g.H);
    p.b = (int)(t / g.H);
    p.x0 = 4 * gi;
    return p;
}
// rows 2y-1 .. 2y+2, columns 2x0-1 .. 2x0+8 of one plane of gO [2H][2W] (zero outside)
__device__ __forceinline__ void load_gpatch(const float* __restrict__ plane, const Geo& g, const RowPatch& p, float (&v)[4][10]) {
    const int H2 = 2 * g.H, W2 = 2 * g.W, X0 = 2 * p.x0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int Y = 2 * p.y - 1 + i;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
        float l = 0.f, r = 0.f;
        if ((unsigned)Y < (unsigned)H2) {
            const float* row = plane + (long long)Y * W2 + X0;
            a = ofr::ldg4(row), c = ofr::ldg4(row + 4);
            if (X0 > 0) l = __ldg(row - 1);
            if (X0 + 8 < W2) r = __ldg(row + 8);
        }
        v[i][0] = l, v[i][1] = a.x, v[i][2] = a.y, v[i][3] = a.z, v[i][4] = a.w;
        v[i][5] = c.x, v[i][6] = c.y, v[i][7] = c.z, v[i][8] = c.w, v[i][9] = r;
    }
}

// gX[b][c][y][x] = sum_o sum_k gO[b][o][2y - 1 + ky][2x - 1 + kx] * w[c][o][ky][kx]
__global__ void __launch_bounds__(kThreads) deconv4_bwd_x_kernel(const float* __restrict__ gO, const float* __restrict__ w,
                                                                 float* __restrict__ gX, Geo g, long long items, int cper) {
    __shared__ __align__(16) float ws[kMaxCper * kTPad];
    const int c0 = blockIdx.y * cper, nc = min(cper, g.C - c0);
    const long long item = (long long)blockIdx.x * kThreads + threadIdx.x;
    const bool valid = item < items;
    const RowPatch p = rowpatch_of(g, valid ? item : 0);
    const long long hw = (long long)g.H * g.W;
    float gv[2][4][10];  // in flight while the weights are staged
    if (valid)
#pragma unroll
        for (int o = 0; o < 2; ++o) load_gpatch(gO + ((long long)p.b * 2 + o) * 4 * hw, g, p, gv[o]);
    stage_w4(ws, w, c0, nc);
    __syncthreads();
    if (!valid) return;
    float* dst = gX + ((long long)p.b * g.C + c0) * hw + (long long)p.y * g.W + p.x0;
    for (int c = 0; c < nc; ++c, dst += hw) {
        float wv[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            const float4 t = *reinterpret_cast<const float4*>(ws + c * kTPad + i);
            wv[i] = t.x, wv[i + 1] = t.y, wv[i + 2] = t.z, wv[i + 3] = t.w;
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int ky = 0; ky < 4; ++ky)
#pragma unroll
                for (int kx = 0; kx < 4; ++kx)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[i] = fmaf(gv[o][ky][2 * i + kx], wv[o * 16 + ky * 4 + kx], acc[i]);
        ofr::stcs4(dst, make_float4(acc[0], acc[1], acc[2], acc[3]));
    }
}

// one row of the 4 x 10 output-gradient neighbourhood of a row patch (zero outside the 2H x 2W map)
__device__ __forceinline__ void load_grow10(const float* __restrict__ plane, int Y, int X0, int H2, int W2, float (&v)[10]) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
    float l = 0.f, r = 0.f;
    if ((unsigned)Y < (unsigned)H2) {
        const float* row = plane + (long long)Y * W2 + X0;
        a = ofr::ldg4(row), c = ofr::ldg4(row + 4);
        if (X0 > 0) l = __ldg(row - 1);
        if (X0 + 8 < W2) r = __ldg(row + 8);
    }
    v[0] = l, v[1] = a.x, v[2] = a.y, v[3] = a.z, v[4] = a.w, v[5] = c.x, v[6] = c.y, v[7] = c.z, v[8] = c.w, v[9] = r;
}

// partial[split][c][o*16 + k] = sum over the split's row patches of x[c][y][x] * gO[o][2y - 1 + ky][2x - 1 + kx]; warps and
// lanes as in conv3_bwd_w_kernel: a lane walks walk_rows(H) input rows down its 4-pixel column, and of the four gradient
// rows 2y-1 .. 2y+2 it needs two are still in registers from the row above
// three CTAs per SM: 200 -> 168 registers at the price of a 32-byte spill; upfeat3's backward 422 -> 350 us (four CTAs, 128
// registers: 505 us)
__global__ void __launch_bounds__(kThreads, 3) deconv4_bwd_w_kernel(const float* __restrict__ x, const float* __restrict__ gO,
                                                                 float* __restrict__ partial, float* __restrict__ pbias,
                                                                 Geo g, int splits) {
    const int lane = threadIdx.x & 31, pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int c0 = pair * 2;
    if (c0 >= g.C) return;  // whole warp
    const bool two = c0 + 1 < g.C;
    const long long hw = (long long)g.H * g.W;
    const int H2 = 2 * g.H, W2 = 2 * g.W;
    const int wrows = walk_rows(g.H), nrb = (g.H + wrows - 1) / wrows;
    const long long ncols = (long long)g.B * nrb * g.gx;
    float acc[66];  // [ch][o][ky][kx] (64) + the two bias sums
#pragma unroll
    for (int k = 0; k < 66; ++k) acc[k] = 0.f;
    for (long long col = (long long)blockIdx.y * 32 + lane; col < ncols; col += (long long)splits * 32) {
        const int gi = (int)(col % g.gx);
        const long long t = col / g.gx;
        const int rb = (int)(t % nrb), b = (int)(t / nrb);
        const int y0 = rb * wrows, y1 = min(g.H, y0 + wrows), x0 = 4 * gi, X0 = 2 * x0;
        const float* xa = x + ((long long)b * g.C + c0) * hw + x0;
        const float* g0 = gO + (long long)b * 2 * 4 * hw;
        const float* g1 = g0 + 4 * hw;
        float gv[2][4][10];
#pragma unroll
        for (int i = 0; i < 2; ++i) {  // rows 2*y0 - 1, 2*y0; the walk loads the other two
            load_grow10(g0, 2 * y0 - 1 + i, X0, H2, W2, gv[0][i]);
            load_grow10(g1, 2 * y0 - 1 + i, X0, H2, W2, gv[1][i]);
        }
        for (int y = y0; y < y1; ++y) {
#pragma unroll
            for (int i = 2; i < 4; ++i) {
                load_grow10(g0, 2 * y - 1 + i, X0, H2, W2, gv[0][i]);
                load_grow10(g1, 2 * y - 1 + i, X0, H2, W2, gv[1][i]);
            }
            const float4 t0 = ofr::ldg4(xa + (long long)y * g.W);
            const float4 t1 = two ? ofr::ldg4(xa + hw + (long long)y * g.W) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float xv[2][4] = {{t0.x, t0.y, t0.z, t0.w}, {t1.x, t1.y, t1.z, t1.w}};
            if (pair == 0) {  // the patch's own outputs: rows 2y, 2y+1 (patch rows 1, 2), columns 2x0 .. 2x0+7 (1 .. 8)
#pragma unroll
                for (int o = 0; o < 2; ++o)
#pragma unroll
                    for (int r = 1; r < 3; ++r)
#pragma unroll
                        for (int i = 1; i < 9; ++i) acc[64 + o] += gv[o][r][i];
            }
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                if (ch == 1 && !two) break;
#pragma unroll
                for (int o = 0; o < 2; ++o)
#pragma unroll
                    for (int ky = 0; ky < 4; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 4; ++kx)
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                acc[ch * 32 + o * 16 + ky * 4 + kx] = fmaf(xv[ch][i], gv[o][ky][2 * i + kx], acc[ch * 32 + o * 16 + ky * 4 + kx]);
            }
#pragma unroll
            for (int o = 0; o < 2; ++o)
#pragma unroll
                for (int j = 0; j < 10; ++j) gv[o][0][j] = gv[o][2][j], gv[o][1][j] = gv[o][3][j];
        }
    }
    warp_sum<66>(acc);
    if (lane == 0) {
        float* dst = partial + ((long long)blockIdx.y * g.C + c0) * 32;
#pragma unroll
        for (int k = 0; k < 32; ++k) dst[k] = acc[k];
        if (two)
#pragma unroll
            for (int k = 0; k < 32; ++k) dst[32 + k] = acc[32 + k];
        if (pair == 0) pbias[blockIdx.y * 2] = acc[64], pbias[blockIdx.y * 2 + 1] = acc[65];
    }
}

// ------------------------------------------------------------------------------------------------ launch plans
struct Plan {
    Geo g;
    int cper, ks;      // forward / input gradient: channels per CTA, channel splits
    int wsplits;       // 3x3 weight gradient: subsets of the column groups
    int wsplits4;      // transposed weight gradient: subsets of the row-patch groups
    int wwarps;        // weight gradient: channel pairs (warps) per CTA
    unsigned xblocks;  // ceil(items / kThreads)
};

bool make_plan(Plan& p, int B, int C, int H, int W) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || (W % 4) != 0) return false;
    p.g = make_geo(B, C, H, W);
    if (p.g.items >= (1ll << 31) * (long long)kThreads) return false;
    p.xblocks = (unsigned)((p.g.items + kThreads - 1) / kThreads);
    // channel splits: enough CTAs for ~4 per SM, at most kMaxCper channels per CTA, at least 8
    const long long want = 4ll * ofr::sm_count();
    int ks = (int)((want + p.xblocks - 1) / p.xblocks);
    const int min_ks = (C + kMaxCper - 1) / kMaxCper, max_ks = (C + 7) / 8;
    if (ks < min_ks) ks = min_ks;
    if (ks > max_ks) ks = max_ks;
    if (ks < 1) ks = 1;
    p.cper = (C + ks - 1) / ks;
    p.ks = (C + p.cper - 1) / p.cper;
    if (p.ks > 65535) return false;
    // weight gradient: a warp per channel pair, wwarps (<= 4) pairs per CTA, wsplits subsets of the lanes' work units (3x3:
    // groups of 32 columns of walk_rows() row pairs; transposed: the same with walk_rows(H) input rows); ~32 warps per SM
    const int pairs = (C + 1) / 2;
    p.wwarps = pairs < kThreads / 32 ? pairs : kThreads / 32;
    const int pair_ctas = (pairs + p.wwarps - 1) / p.wwarps;
    const long long want_ws = (32ll * ofr::sm_count() / p.wwarps + pair_ctas - 1) / pair_ctas;
    const int wr = walk_rows(p.g.rp);
    const long long groups3 = ((long long)B * ((p.g.rp + wr - 1) / wr) * p.g.gx + 31) / 32;
    const int wr4 = walk_rows(H);
    const long long groups4 = ((long long)B * ((H + wr4 - 1) / wr4) * p.g.gx + 31) / 32;
    auto clamp_ws = [](long long ws, long long groups) { return (int)(ws > groups ? groups : ws < 1 ? 1 : ws > 65535 ? 65535 : ws); };
    p.wsplits = clamp_ws(want_ws, groups3);
    p.wsplits4 = clamp_ws(want_ws, groups4);
    return true;
}

size_t fwd_ws_bytes(const Plan& p, long long out_elems) { return p.ks > 1 ? sizeof(float) * (size_t)p.ks * out_elems : 0; }
size_t wgrad_ws_bytes(const Plan& p, int K) {
    const size_t sp = K == 9 ? p.wsplits : p.wsplits4;
    return ofr::align_up(sizeof(float) * (sp * p.g.C * 2 * K + sp * 2), 256);
}

}  // namespace

extern "C" size_t ofr_flowconv3_workspace(int B, int C, int H, int W, int backward) {
    Plan p;
    if (!make_plan(p, B, C, H, W)) return 0;
    const size_t f = fwd_ws_bytes(p, (long long)B * 2 * H * W), b = wgrad_ws_bytes(p, 9);
    return ofr::align_up(backward ? b : f, 256);
}

extern "C" int ofr_flowconv3_fwd(const float* x, const float* w, const float* bias, float* out, int B, int C, int H, int W,
                                 void* ws, size_t ws_bytes, ofr_stream_t stream) {
    Plan p;
    if (!x || !w || !out) return OFR_EINVAL;
    if (!make_plan(p, B, C, H, W) || !ofr::aligned16(x) || !ofr::aligned16(out)) return OFR_ENOTSUP;
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)B * 2 * H * W;
    if (p.ks == 1) {
        conv3_fwd_kernel<<<dim3(p.xblocks, 1), kThreads, 0, st>>>(x, w, bias, out, p.g, p.cper, 1);
        OFR_LAUNCH_CHECK();
        return OFR_OK;
    }
    if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255u) || ws_bytes < fwd_ws_bytes(p, n)) return OFR_EWORKSPACE;
    conv3_fwd_kernel<<<dim3(p.xblocks, p.ks), kThreads, 0, st>>>(x, w, bias, (float*)ws, p.g, p.cper, 0);
    OFR_LAUNCH_CHECK();
    split_sum_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>((const float*)ws, bias, out, n / 4, (long long)H * W / 4, p.ks);
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

extern "C" int ofr_flowconv3_bwd(const float* x, const float* w, const float* gO, float* gX, float* gW, float* gB, int B,
                                 int C, int H, int W, void* ws, size_t ws_bytes, ofr_stream_t stream) {
    Plan p;
    if (!x || !w || !gO) return OFR_EINVAL;
    if (!make_plan(p, B, C, H, W) || !ofr::aligned16(x) || !ofr::aligned16(gO) || (gX && !ofr::aligned16(gX))) return OFR_ENOTSUP;
    cudaStream_t st = (cudaStream_t)stream;
    if (gX) {
        conv3_bwd_x_kernel<<<dim3(p.xblocks, p.ks), kThreads, 0, st>>>(gO, w, gX, p.g, p.cper);
        OFR_LAUNCH_CHECK();
    }
    if (gW) {
        if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255u) || ws_bytes < wgrad_ws_bytes(p, 9)) return OFR_EWORKSPACE;
        float* partial = (float*)ws;
        float* pbias = partial + (size_t)p.wsplits * C * 18;
        conv3_bwd_w_kernel<<<dim3(((C + 1) / 2 + p.wwarps - 1) / p.wwarps, p.wsplits), 32 * p.wwarps, 0, st>>>(x, gO, partial, pbias, p.g, p.wsplits);
        OFR_LAUNCH_CHECK();
        wgrad_sum_kernel<<<(C * 18 + 2 + 7) / 8, 256, 0, st>>>(partial, pbias, gW, gB, C, 9, p.wsplits, 0);
        OFR_LAUNCH_CHECK();
    }
    return OFR_OK;
}

extern "C" size_t ofr_flowdeconv4_workspace(int B, int C, int H, int W, int backward) {
    Plan p;
    if (!make_plan(p, B, C, H, W)) return 0;
    const size_t f = fwd_ws_bytes(p, (long long)B * 2 * 4 * H * W), b = wgrad_ws_bytes(p, 16);
    return ofr::align_up(backward ? b : f, 256);
}

extern "C" int ofr_flowdeconv4_fwd(const float* x, const float* w, const float* bias, float* out, int B, int C, int H, int W,
                                   void* ws, size_t ws_bytes, ofr_stream_t stream) {
    Plan p;
    if (!x || !w || !out) return OFR_EINVAL;
    if (!make_plan(p, B, C, H, W) || !ofr::aligned16(x) || !ofr::aligned16(out) || !ofr::aligned16(w)) return OFR_ENOTSUP;
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)B * 2 * 4 * H * W;
    if (p.ks == 1) {
        deconv4_fwd_kernel<<<dim3(p.xblocks, 1), kThreads, 0, st>>>(x, w, bias, out, p.g, p.cper, 1);
        OFR_LAUNCH_CHECK();
        return OFR_OK;
    }
    if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255u) || ws_bytes < fwd_ws_bytes(p, n)) return OFR_EWORKSPACE;
    deconv4_fwd_kernel<<<dim3(p.xblocks, p.ks), kThreads, 0, st>>>(x, w, bias, (float*)ws, p.g, p.cper, 0);
    OFR_LAUNCH_CHECK();
    split_sum_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>((const float*)ws, bias, out, n / 4, (long long)H * W, p.ks);
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

extern "C" int ofr_flowdeconv4_bwd(const float* x, const float* w, const float* gO, float* gX, float* gW, float* gB, int B,
                                   int C, int H, int W, void* ws, size_t ws_bytes, ofr_stream_t stream) {
    Plan p;
    if (!x || !w || !gO) return OFR_EINVAL;
    if (!make_plan(p, B, C, H, W) || !ofr::aligned16(x) || !ofr::aligned16(gO) || (gX && !ofr::aligned16(gX))) return OFR_ENOTSUP;
    cudaStream_t st = (cudaStream_t)stream;
    const long long items = (long long)B * H * (W / 4);  // one input row x 4 pixels per thread
    const unsigned xb = (unsigned)((items + kThreads - 1) / kThreads);
    if (gX) {
        deconv4_bwd_x_kernel<<<dim3(xb, p.ks), kThreads, 0, st>>>(gO, w, gX, p.g, items, p.cper);
        OFR_LAUNCH_CHECK();
    }
    if (gW) {
        if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255u) || ws_bytes < wgrad_ws_bytes(p, 16)) return OFR_EWORKSPACE;
        float* partial = (float*)ws;
        float* pbias = partial + (size_t)p.wsplits4 * C * 32;
        deconv4_bwd_w_kernel<<<dim3(((C + 1) / 2 + p.wwarps - 1) / p.wwarps, p.wsplits4), 32 * p.wwarps, 0, st>>>(x, gO, partial, pbias, p.g, p.wsplits4);
        OFR_LAUNCH_CHECK();
        wgrad_sum_kernel<<<(C * 32 + 2 + 7) / 8, 256, 0, st>>>(partial, pbias, gW, gB, C, 16, p.wsplits4, 1);
        OFR_LAUNCH_CHECK();
    }
    return OFR_OK;
}
