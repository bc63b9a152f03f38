// TMA-staged correlation kernels for the PWC-Net configuration Correlation(4,1,4,1,1) (PWC/models/PWCNet.py:69):
// k=1, stride1=1, stride2=1, pad == md == 4, D = 9, 81 output channels.
//
// Forward  (corr_fwd_tma_kernel): a CTA owns a TH x TW tile of one image.  One elected thread streams chunks of
//   CK channels through an NST-deep ring of shared-memory stages with cp.async.bulk.tensor (TMA): the f1 tile and
//   the f2 tile + 4-pixel displacement halo; out-of-image halo elements are zero-filled by the TMA unit, which is
//   exactly the zero padding of the operator.  Each thread owns 8 pixels along x and the 9 x-displacements of ONE
//   displacement row tj: the Toeplitz register tile acc[i][j] += a[i] * b[i+j] (72 FMAs per 6 128-bit LDS).  The 9
//   tj-threads of a strip are adjacent lanes, so their `a` loads are one broadcast.  Channel sum = fixed-order loop.
// Backward (corr_bwd_tma_kernel): the same Toeplitz tile, transposed.  A CTA owns an 8 x TW tile and 32 channels.
//   The feature tile + halo of those 32 channels stays resident in shared memory; the 81 gradient planes stream
//   through a 2-stage TMA ring, 9 planes (one displacement row) per stage.  A thread owns 8 pixels x 8 channels:
//   acc[i][k] += g[ti][i] * b_k[i+ti].  Lanes that differ only in channel group broadcast the g loads.
//     gI1[c,p] = s * sum_t gO[t,p]      * f2[c,p+t]              (planes staged in place)
//     gI2[c,q] = s * sum_t gO[-t,q+t]   * f1[c,q+t]              (plane -t staged with its box shifted by t:
//                                                                  the same inner loop serves both gradients)
//   No atomics, fixed summation order -> bitwise reproducible.
#include "corr_internal.cuh"
#include "tma.cuh"

namespace {

using ofr::CorrGeom;

constexpr int kR = 4, kD = 9, kM = 8;

__host__ __device__ constexpr int pitch_for(int cols) {  // multiple of 4, == 4*odd (mod 32)
    int p = (cols + 3) / 4 * 4;
    while ((p / 4) % 2 == 0 || p % 32 == 0) p += 4;
    return p;
}

// ------------------------------------------------------------------------------------------------ forward
template <int TH, int NS, int CK, int NST>
struct FwdT {
    static constexpr int TW = NS * kM;
    static constexpr int P2 = pitch_for(TW + 2 * kR);
    static constexpr int R2 = TH + 2 * kR;
    static constexpr int F1 = CK * TH * TW;   // floats per stage
    static constexpr int F2 = CK * R2 * P2;
    static constexpr int STAGE = F1 + F2;
    static constexpr int THREADS = kD * NS * TH;
    static constexpr size_t SMEM = sizeof(float) * STAGE * NST + 8 * NST + 1024;
    static_assert((F1 * 4) % 128 == 0 && (F2 * 4) % 128 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(THREADS <= 1024, "CTA too large");
};

template <int TH, int NS, int CK, int NST>
__global__ void __launch_bounds__(FwdT<TH, NS, CK, NST>::THREADS, 1)
    corr_fwd_tma_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                        float* __restrict__ out, CorrGeom g, int tiles_x, int tiles_y) {
    using T = FwdT<TH, NS, CK, NST>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // round the base up to 1024 B with an offset (not an integer cast) so the pointer stays in the shared window: LDS, not LD
    float* smem = reinterpret_cast<float*>(smem_raw + ((1024u - (ofr::smem_u32(smem_raw) & 1023u)) & 1023u));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T::STAGE * NST);

    int bid = blockIdx.x;
    const int txi = bid % tiles_x;
    bid /= tiles_x;
    const int tyi = bid % tiles_y;
    const int b = bid / tiles_y;
    const int x0 = txi * T::TW, y0 = tyi * TH;

    const int tid = threadIdx.x;
    const int tjl = tid % kD;
    const int s = (tid / kD) % NS;
    const int row = tid / (kD * NS);

    const int nchunks = (g.C + CK - 1) / CK;
    auto issue = [&](int chunk, int st) {
        float* dst = smem + st * T::STAGE;
        ofr::mbar_expect_tx(&bars[st], T::STAGE * 4);
        ofr::tma_load_4d(dst, &map1, &bars[st], x0, y0, chunk * CK, b);
        ofr::tma_load_4d(dst + T::F1, &map2, &bars[st], x0 - kR, y0 - kR, chunk * CK, b);
    };
    if (tid == 0) {
        ofr::tma_prefetch_desc(&map1);
        ofr::tma_prefetch_desc(&map2);
        for (int i = 0; i < NST; ++i) ofr::mbar_init(&bars[i], 1);
        ofr::mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0)
        for (int i = 0; i < NST && i < nchunks; ++i) issue(i, i);

    float acc[kM][kD];
#pragma unroll
    for (int i = 0; i < kM; ++i)
#pragma unroll
        for (int j = 0; j < kD; ++j) acc[i][j] = 0.f;

    for (int ck = 0; ck < nchunks; ++ck) {
        const int st = ck % NST;
        ofr::mbar_wait(&bars[st], (ck / NST) & 1);
        const float* ap = smem + st * T::STAGE + row * T::TW + s * kM;
        const float* bp = smem + st * T::STAGE + T::F1 + (row + tjl) * T::P2 + s * kM;
#pragma unroll 1
        for (int cc = 0; cc < CK; ++cc) {
            float a[kM], bv[kM + 2 * kR];
#pragma unroll
            for (int i = 0; i < kM; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(ap + cc * TH * T::TW + i);
                a[i] = v.x, a[i + 1] = v.y, a[i + 2] = v.z, a[i + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < kM + 2 * kR; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(bp + cc * T::R2 * T::P2 + i);
                bv[i] = v.x, bv[i + 1] = v.y, bv[i + 2] = v.z, bv[i + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < kM; ++i)
#pragma unroll
                for (int j = 0; j < kD; ++j) acc[i][j] = fmaf(a[i], bv[i + j], acc[i][j]);
        }
        __syncthreads();  // every thread is done reading stage st
        if (tid == 0 && ck + NST < nchunks) {
            ofr::fence_proxy_async();
            issue(ck + NST, st);
        }
    }

    const int y = y0 + row;
    if (y < g.Hout) {
        const long long ohw = (long long)g.Hout * g.Wout;
        float* ob = out + ((long long)b * kD * kD + (long long)tjl * kD) * ohw + (long long)y * g.Wout;
        const int xb = x0 + s * kM;
#pragma unroll
        for (int j = 0; j < kD; ++j) {
            float* o = ob + (long long)j * ohw;
            if (xb + kM <= g.Wout) {
                ofr::stcs4(o + xb, make_float4(acc[0][j] * g.scale, acc[1][j] * g.scale, acc[2][j] * g.scale, acc[3][j] * g.scale));
                ofr::stcs4(o + xb + 4, make_float4(acc[4][j] * g.scale, acc[5][j] * g.scale, acc[6][j] * g.scale, acc[7][j] * g.scale));
            } else {
#pragma unroll
                for (int i = 0; i < kM; ++i)
                    if (xb + i < g.Wout) o[xb + i] = acc[i][j] * g.scale;
            }
        }
    }
}

template <int TH, int NS, int CK, int NST>
int launch_fwd(const float* f1, const float* f2, float* out, const CorrGeom& g, cudaStream_t st) {
    using T = FwdT<TH, NS, CK, NST>;
    CUtensorMap m1, m2;
    if (!ofr::make_tmap_nchw(&m1, f1, g.B, g.C, g.H, g.W, T::TW, TH, CK)) return OFR_ENOTSUP;
    if (!ofr::make_tmap_nchw(&m2, f2, g.B, g.C, g.H, g.W, T::P2, T::R2, CK)) return OFR_ENOTSUP;
    auto kern = corr_fwd_tma_kernel<TH, NS, CK, NST>;
    OFR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM));
    const int tiles_x = (g.Wout + T::TW - 1) / T::TW, tiles_y = (g.Hout + TH - 1) / TH;
    kern<<<(unsigned)(tiles_x * tiles_y * g.B), T::THREADS, T::SMEM, st>>>(m1, m2, out, g, tiles_x, tiles_y);
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

// ------------------------------------------------------------------------------------------------ backward
constexpr int kBTH = 8;    // tile rows
constexpr int kCB = 32;    // channels per CTA
constexpr int kCG = 4;     // channel groups (lanes); each thread owns kCB / kCG = 8 channels, c = k*4 + cg
constexpr int kCN = kCB / kCG;

template <int NS, bool SHIFTED>
struct BwdT {
    static constexpr int TW = NS * kM;
    static constexpr int P2 = pitch_for(TW + 2 * kR);
    static constexpr int R2 = kBTH + 2 * kR + 1;  // +1 row: makes the channel pitch == 4*odd (mod 32) -> cg lanes hit distinct banks
    static constexpr int FT = kCB * R2 * P2;      // resident feature tile, floats
    // TMA needs a 16-byte aligned box origin along x, so the shifted planes of the second gradient are staged as
    // aligned boxes widened by the halo and the +-4 pixel shift is applied when the registers are filled.
    static constexpr int GW = SHIFTED ? TW + 2 * kR : TW;  // row pitch of a staged gradient plane
    static constexpr int GP = kBTH * GW;                    // one plane tile, floats
    static constexpr int GS = kD * GP;                      // one gradient stage (9 planes), floats
    static constexpr int THREADS = kBTH * NS * kCG;
    static constexpr size_t SMEM = sizeof(float) * (FT + 2 * GS) + 8 * 3 + 1024;
    static_assert((FT * 4) % 128 == 0 && (GP * 4) % 128 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(((R2 * P2) / 4) % 2 == 1, "channel pitch must be 4*odd (mod 32)");
};

template <int NS, bool SHIFTED>
__global__ void __launch_bounds__(BwdT<NS, SHIFTED>::THREADS, 1)
    corr_bwd_tma_kernel(const __grid_constant__ CUtensorMap mapF, const __grid_constant__ CUtensorMap mapG,
                        float* __restrict__ gI, CorrGeom g, int tiles_x, int tiles_y, int cblocks) {
    using T = BwdT<NS, SHIFTED>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Fs = reinterpret_cast<float*>(smem_raw + ((1024u - (ofr::smem_u32(smem_raw) & 1023u)) & 1023u));
    float* Gs = Fs + T::FT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Gs + 2 * T::GS);  // [0],[1]: gradient stages, [2]: feature tile

    int bid = blockIdx.x;
    const int txi = bid % tiles_x;
    bid /= tiles_x;
    const int tyi = bid % tiles_y;
    bid /= tiles_y;
    const int cbi = bid % cblocks;
    const int b = bid / cblocks;
    const int x0 = txi * T::TW, y0 = tyi * kBTH, c0 = cbi * kCB;

    const int tid = threadIdx.x;
    const int cg = tid % kCG;
    const int s = (tid / kCG) % NS;
    const int row = tid / (kCG * NS);

    auto issue_g = [&](int tji, int st) {
        float* dst = Gs + st * T::GS;
        ofr::mbar_expect_tx(&bars[st], T::GS * 4);
#pragma unroll
        for (int tii = 0; tii < kD; ++tii) {
            if (SHIFTED)  // plane of displacement -t, rows shifted by tj here, columns shifted by ti at register fill
                ofr::tma_load_4d(dst + tii * T::GP, &mapG, &bars[st], x0 - kR, y0 + tji - kR,
                                 (2 * kR - tji) * kD + (2 * kR - tii), b);
            else
                ofr::tma_load_4d(dst + tii * T::GP, &mapG, &bars[st], x0, y0, tji * kD + tii, b);
        }
    };
    if (tid == 0) {
        ofr::tma_prefetch_desc(&mapF);
        ofr::tma_prefetch_desc(&mapG);
        ofr::mbar_init(&bars[0], 1);
        ofr::mbar_init(&bars[1], 1);
        ofr::mbar_init(&bars[2], 1);
        ofr::mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        ofr::mbar_expect_tx(&bars[2], T::FT * 4);
        ofr::tma_load_4d(Fs, &mapF, &bars[2], x0 - kR, y0 - kR, c0, b);
        issue_g(0, 0);
        issue_g(1, 1);
    }

    float acc[kM][kCN];
#pragma unroll
    for (int i = 0; i < kM; ++i)
#pragma unroll
        for (int k = 0; k < kCN; ++k) acc[i][k] = 0.f;

    ofr::mbar_wait(&bars[2], 0);
    for (int tji = 0; tji < kD; ++tji) {
        const int st = tji & 1;
        ofr::mbar_wait(&bars[st], (tji >> 1) & 1);
        const float* gp = Gs + st * T::GS + row * T::GW + s * kM;
        float gv[kD][kM];
#pragma unroll
        for (int tii = 0; tii < kD; ++tii) {
            if (SHIFTED) {  // value for target x sits at column x - (x0 - R) + ti = s*8 + i + tii of the widened box
                float w[kM + 2 * kR];
#pragma unroll
                for (int i = 0; i < kM + 2 * kR; i += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(gp + tii * T::GP + i);
                    w[i] = v.x, w[i + 1] = v.y, w[i + 2] = v.z, w[i + 3] = v.w;
                }
#pragma unroll
                for (int i = 0; i < kM; ++i) gv[tii][i] = w[i + tii];
            } else {
                const float4 v0 = *reinterpret_cast<const float4*>(gp + tii * T::GP);
                const float4 v1 = *reinterpret_cast<const float4*>(gp + tii * T::GP + 4);
                gv[tii][0] = v0.x, gv[tii][1] = v0.y, gv[tii][2] = v0.z, gv[tii][3] = v0.w;
                gv[tii][4] = v1.x, gv[tii][5] = v1.y, gv[tii][6] = v1.z, gv[tii][7] = v1.w;
            }
        }
        const float* fp = Fs + (long long)cg * T::R2 * T::P2 + (row + tji) * T::P2 + s * kM;
#pragma unroll
        for (int k = 0; k < kCN; ++k) {
            float bv[kM + 2 * kR];
#pragma unroll
            for (int i = 0; i < kM + 2 * kR; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(fp + k * kCG * T::R2 * T::P2 + i);
                bv[i] = v.x, bv[i + 1] = v.y, bv[i + 2] = v.z, bv[i + 3] = v.w;
            }
#pragma unroll
            for (int tii = 0; tii < kD; ++tii)
#pragma unroll
                for (int i = 0; i < kM; ++i) acc[i][k] = fmaf(gv[tii][i], bv[i + tii], acc[i][k]);
        }
        __syncthreads();
        if (tid == 0 && tji + 2 < kD) {
            ofr::fence_proxy_async();
            issue_g(tji + 2, st);
        }
    }

    const int y = y0 + row, xb = x0 + s * kM;
    if (y < g.H) {
        const long long hw = (long long)g.H * g.W;
#pragma unroll
        for (int k = 0; k < kCN; ++k) {
            const int c = c0 + k * kCG + cg;
            if (c >= g.C) continue;
            float* o = gI + ((long long)b * g.C + c) * hw + (long long)y * g.W;
            if (xb + kM <= g.W) {
                *reinterpret_cast<float4*>(o + xb) = make_float4(acc[0][k] * g.scale, acc[1][k] * g.scale, acc[2][k] * g.scale, acc[3][k] * g.scale);
                *reinterpret_cast<float4*>(o + xb + 4) = make_float4(acc[4][k] * g.scale, acc[5][k] * g.scale, acc[6][k] * g.scale, acc[7][k] * g.scale);
            } else {
#pragma unroll
                for (int i = 0; i < kM; ++i)
                    if (xb + i < g.W) o[xb + i] = acc[i][k] * g.scale;
            }
        }
    }
}

template <int NS, bool SHIFTED>
int launch_bwd(const float* feat, const float* gO, float* gI, const CorrGeom& g, cudaStream_t st) {
    using T = BwdT<NS, SHIFTED>;
    CUtensorMap mF, mG;
    if (!ofr::make_tmap_nchw(&mF, feat, g.B, g.C, g.H, g.W, T::P2, T::R2, kCB)) return OFR_ENOTSUP;
    if (!ofr::make_tmap_nchw(&mG, gO, g.B, kD * kD, g.Hout, g.Wout, T::GW, kBTH, 1)) return OFR_ENOTSUP;
    auto kern = corr_bwd_tma_kernel<NS, SHIFTED>;
    OFR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM));
    const int tiles_x = (g.W + T::TW - 1) / T::TW, tiles_y = (g.H + kBTH - 1) / kBTH;
    const int cblocks = (g.C + kCB - 1) / kCB;
    kern<<<(unsigned)(tiles_x * tiles_y * cblocks * g.B), T::THREADS, T::SMEM, st>>>(mF, mG, gI, g, tiles_x, tiles_y, cblocks);
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

bool applicable(const CorrGeom& g, const void* a, const void* b, const void* c) {
    return g.k == 1 && g.s1 == 1 && g.s2 == 1 && g.md == kR && g.pad == kR && (g.W % 4) == 0 && ofr::aligned16(a) &&
           ofr::aligned16(b) && ofr::aligned16(c) && ofr::encode_tiled_fn() != nullptr &&
           (long long)g.B * 4096 * ((g.W + 7) / 8) * ((g.H + 7) / 8) < (1ll << 31);
}

}  // namespace

namespace ofr {

int corr_fwd_tma(const float* f1, const float* f2, float* out, const CorrGeom& g, cudaStream_t st) {
    if (!applicable(g, f1, f2, out)) return OFR_ENOTSUP;
    if (g.W > 32) return launch_fwd<8, 8, 8, 3>(f1, f2, out, g, st);   // 8 x 64 tile, 576 threads
    if (g.W > 16) return launch_fwd<8, 4, 8, 3>(f1, f2, out, g, st);   // 8 x 32 tile, 288 threads
    if (g.W > 8) return launch_fwd<8, 2, 8, 4>(f1, f2, out, g, st);    // 8 x 16 tile, 144 threads
    return launch_fwd<4, 1, 8, 4>(f1, f2, out, g, st);                 // 4 x 8 tile,   36 threads
}

int corr_bwd_tma(const float* f1, const float* f2, const float* gO, float* gI1, float* gI2, const CorrGeom& g,
                 cudaStream_t st) {
    if (!applicable(g, f1, f2, gO) || (gI1 && !ofr::aligned16(gI1)) || (gI2 && !ofr::aligned16(gI2))) return OFR_ENOTSUP;
    int rc = OFR_OK;
    if (gI1) {
        if (g.W > 32) rc = launch_bwd<8, false>(f2, gO, gI1, g, st);
        else if (g.W > 16) rc = launch_bwd<4, false>(f2, gO, gI1, g, st);
        else if (g.W > 8) rc = launch_bwd<2, false>(f2, gO, gI1, g, st);
        else rc = launch_bwd<1, false>(f2, gO, gI1, g, st);
        if (rc != OFR_OK) return rc;
    }
    if (gI2) {
        if (g.W > 32) rc = launch_bwd<8, true>(f1, gO, gI2, g, st);
        else if (g.W > 16) rc = launch_bwd<4, true>(f1, gO, gI2, g, st);
        else if (g.W > 8) rc = launch_bwd<2, true>(f1, gO, gI2, g, st);
        else rc = launch_bwd<1, true>(f1, gO, gI2, g, st);
    }
    return rc;
}

}  // namespace ofr
