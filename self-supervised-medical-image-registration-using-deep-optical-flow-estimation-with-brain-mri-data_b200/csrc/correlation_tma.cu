// TMA-staged stride-1 correlation kernels (k = 1, stride1 = 1, stride2 = 1, pad == md == R, D = 2R+1).
//   R = 4  : the PWC-Net configuration Correlation(4,1,4,1,1)               (PWC/models/PWCNet.py:69)
//   R = 10 : the FlowNetC configuration Correlation(20,1,20,1,2) after the 2x2 polyphase split of correlation.cu:
//            with stride2 = 2 a pixel only ever meets pixels of its own (y mod 2, x mod 2) class, so on each of the
//            four half-resolution class images the operator IS a stride-1 correlation with R = 10
//            (flownet2/networks/FlowNetC.py:31).
//
// Forward  (corr_fwd_tma_kernel): persistent CTAs (one per SM) walk the (image, tile) list.  One elected thread
//   streams chunks of CK channels through an NST-deep ring of shared-memory stages with cp.async.bulk.tensor (TMA):
//   the f1 tile and the f2 tile + R-pixel displacement halo; out-of-image halo elements are zero-filled by the TMA
//   unit, which is exactly the zero padding of the operator.  The ring runs ahead across tile boundaries.
//   Thread tile: M pixels along x  x  D x-displacements of ONE displacement row tj: acc[i][j] += a[i] * b[i+j].
//   Shared-memory bandwidth is the scarce resource (a 128-bit LDS costs >= 4 wavefronts however many lanes share an
//   address), so the lane mapping makes every quarter-warp bank-conflict free: lanes = 8 tj (fastest) x 4
//   (strip,row) pairs and the f2 row pitch is 4*odd (mod 32) floats, so the 8 rows of a quarter-warp start in 8
//   different bank groups.  Channel sum = fixed-order loop (no split-K, no atomics).  Epilogue: staged through shared
//   memory and written with asynchronous TMA tensor stores (whole 128-byte lines, image-edge clipping for free).
// Backward (corr_bwd_tma_kernel): the same Toeplitz tile, transposed.  A CTA owns a TH x TW tile and 32 channels.
//   The feature tile + halo of those channels stays resident in shared memory; the D*D gradient planes stream
//   through a 2-stage TMA ring, D planes (one displacement row) per stage.  A thread owns 8 pixels x CN channels:
//   acc[i][k] += g[ti][i] * b_k[i+ti].
//     gI1[c,p] = s * sum_t gO[t,p]      * f2[c,p+t]              (planes staged in place)
//     gI2[c,q] = s * sum_t gO[-t,q+t]   * f1[c,q+t]              (plane -t staged as an aligned, widened box with
//                                                                  its rows shifted by tj; the x shift is applied
//                                                                  when the registers are filled)
//   No atomics, fixed summation order -> bitwise reproducible.
#include <cooperative_groups.h>

#include <cstdlib>

#include "corr_internal.cuh"
#include "tma.cuh"

namespace cg = cooperative_groups;

#ifndef OFR_FWD_TRACE
#define OFR_FWD_TRACE 0   // 1: clock64() timeline of CTA 0 (threads 0 and 32) per ring iteration of the forward kernel,
#endif                    //    read back with ofr_debug_fwd_trace (tools/diag/fwd_trace.py; profiles/r02_corr_fwd_trace.txt)
#if OFR_FWD_TRACE
__device__ long long ofr_fwd_trace_buf[2][64][6];
extern "C" int ofr_debug_fwd_trace(long long* host) {
    return (int)cudaMemcpyFromSymbol(host, ofr_fwd_trace_buf, sizeof(ofr_fwd_trace_buf));
}
#define OFR_TR(slot)                                                                                                  \
    do {                                                                                                              \
        if (blockIdx.x == 0 && (tid == 0 || tid == 32) && q < 64) ofr_fwd_trace_buf[tid >> 5][q][slot] = clock64(); \
    } while (0)
#else
#define OFR_TR(slot)
#endif

namespace {

using ofr::CorrGeom;

__host__ __device__ constexpr int pitch_for(int cols) {  // multiple of 4, == 4*odd (mod 32)
    int p = (cols + 3) / 4 * 4;
    while ((p / 4) % 2 == 0 || p % 32 == 0) p += 4;
    return p;
}

// ------------------------------------------------------------------------------------------------ forward
template <int R, int M, int TH, int NS, int CK, int NST>
struct FwdT {
    static constexpr int D = 2 * R + 1;
    static constexpr int TW = NS * M;
    static constexpr int XO = (4 - R % 4) % 4;  // TMA box origins must be 16-byte aligned along x: start XO pixels further left
    static constexpr int BL = (M + 2 * R + XO + 3) / 4 * 4;  // floats of an f2 row one thread reads
    static constexpr int P1 = pitch_for(TW);
    static constexpr int P2 = pitch_for(TW - M + BL);
    static constexpr int R2 = TH + 2 * R;
    static constexpr int F1 = CK * TH * P1;   // floats per stage
    static constexpr int F2 = CK * R2 * P2;
    static constexpr int STAGE = F1 + F2;
    // D = 9: 8 tj lanes per (row,strip) + one extra warp for tj = 8 (no idle lanes).  Otherwise ceil(D/8) groups of
    // 8 tj lanes, the lanes with tj >= D idle.
    static constexpr bool kNine = (D == 9);
    static constexpr int TJG = (D + 7) / 8;
    static constexpr int T1 = 8 * TH * NS;
    static constexpr int THREADS = kNine ? D * TH * NS : 8 * TJG * TH * NS;
    static constexpr int NWARPS = (THREADS + 31) / 32;
    static constexpr int WSTAGE = 32 * (M + 4) * 4;  // bytes of one per-warp epilogue staging buffer (covers both lane layouts)
    static constexpr int CPR = NS * M / 4;           // 16-byte chunks per run
    static constexpr int EPI = M / 4;                // chunks every lane writes back per step: 32/NS runs * CPR / 32 lanes
    static constexpr size_t SMEM = sizeof(float) * (STAGE * NST) + (size_t)NWARPS * 2 * WSTAGE + 8 * NST + 1024;
    static_assert(32 / NS * (NS * M + 4) <= 32 * (M + 4) && TH * (NS * M + 4) <= 32 * (M + 4) && TH * NS <= 32 &&
                      NS <= 4 && (NS & (NS - 1)) == 0 && (TH * NS) % 4 == 0 && T1 % 32 == 0,
                  "staging layout");
    static_assert((F1 * 4) % 128 == 0 && (F2 * 4) % 128 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(M % 4 == 0 && THREADS <= 1024, "bad tile");
};

template <int R, int M, int TH, int NS, int CK, int NST>
__global__ void __maxnreg__(224)
    corr_fwd_tma_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                        float* __restrict__ out, CorrGeom g, int tiles_x, int tiles_y) {
    using T = FwdT<R, M, TH, NS, CK, NST>;
    constexpr int D = T::D;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // round the base up to 1024 B with an offset (not an integer cast) so the pointer stays in the shared window: LDS, not LD
    float* smem = reinterpret_cast<float*>(smem_raw + ((1024u - (ofr::smem_u32(smem_raw) & 1023u)) & 1023u));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T::STAGE * NST + T::NWARPS * (2 * T::WSTAGE / 4));

    const int tid = threadIdx.x;
    int tj, s, row;
    if (T::kNine) {
        if (tid < T::T1) {
            tj = tid & 7;
            const int q = tid >> 3;
            s = q % NS;
            row = q / NS;
        } else {  // the ninth displacement row: lanes = rows (fastest) x strips
            tj = 8;
            const int q = tid - T::T1;
            row = q % TH;
            s = q / TH;
        }
    } else {
        const int q = tid >> 3;
        const int combo = q % (TH * NS);
        tj = (q / (TH * NS)) * 8 + (tid & 7);
        s = combo % NS;
        row = combo / NS;
    }
    const bool active = tj < D;
    const int tjc = active ? tj : D - 1;  // idle lanes read valid shared memory and discard the result

    // epilogue staging: one private area per warp, two buffers of WSTAGE bytes, organised in runs of NS*M pixels
    // (+16 bytes of padding so the 8 tj lanes of a quarter-warp store to 8 different bank groups).
    //   regular warps : lanes = 8 tj x 4 (strip,row) pairs; run r = (tj = tj_base + (r & 7), row = row_base + (r >> 3))
    //   tj = D-1 warp of the D = 9 layout : lanes = rows x strips; run r = row r
    const int lane = tid & 31;
    const bool grouped = !(T::kNine && tid >= T::T1);
    const int qb = (tid >> 5) * 4;  // first (strip,row) pair of the warp
    const int tj_base = T::kNine ? 0 : (qb / (TH * NS)) * 8;
    const int row_base = T::kNine ? qb / NS : (qb % (TH * NS)) / NS;
    const int nruns = grouped ? 32 / NS : TH;
    const int slot = grouped ? (((lane & 7) + 8 * ((lane >> 3) / NS)) * (NS * M + 4) + ((lane >> 3) % NS) * M)
                             : row * (NS * M + 4) + s * M;
    float* wstage = smem + T::STAGE * NST + (tid >> 5) * (2 * T::WSTAGE / 4);
    const unsigned wmask = __activemask();
    int jcount = 0;

    const int nchunks = (g.C + CK - 1) / CK;
    const int ntiles = tiles_x * tiles_y * g.B;
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int total = my_tiles * nchunks;  // ring iterations of this CTA

    auto tile_origin = [&](int k, int& x0, int& y0, int& b) {  // k-th tile of this CTA
        int t = blockIdx.x + k * gridDim.x;
        x0 = (t % tiles_x) * T::TW;
        t /= tiles_x;
        y0 = (t % tiles_y) * TH;
        b = t / tiles_y;
    };
    auto issue = [&](int q) {
        int x0, y0, b;
        tile_origin(q / nchunks, x0, y0, b);
        const int c0 = (q % nchunks) * CK, st = q % NST;
        float* dst = smem + st * T::STAGE;
        ofr::mbar_expect_tx(&bars[st], T::STAGE * 4);
        ofr::tma_load_4d(dst, &map1, &bars[st], x0, y0, c0, b);
        ofr::tma_load_4d(dst + T::F1, &map2, &bars[st], x0 - R - T::XO, y0 - R, c0, b);
    };
    if (tid == 0) {
        ofr::tma_prefetch_desc(&map1);
        ofr::tma_prefetch_desc(&map2);
        for (int i = 0; i < NST; ++i) ofr::mbar_init(&bars[i], 1);
        ofr::mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0)
        for (int i = 0; i < NST && i < total; ++i) issue(i);

    float acc[M][D];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) acc[i][j] = 0.f;

    int ck = 0, tile_k = 0;
    for (int q = 0; q < total; ++q) {
        const int st = q % NST;
        OFR_TR(0);
        ofr::mbar_wait(&bars[st], (q / NST) & 1);
        OFR_TR(1);
        const float* ap = smem + st * T::STAGE + row * T::P1 + s * M;
        const float* bp = smem + st * T::STAGE + T::F1 + (row + tjc) * T::P2 + s * M;
#pragma unroll 1
        for (int cc = 0; cc < CK; ++cc) {
            float a[M], bv[T::BL];
#pragma unroll
            for (int i = 0; i < M; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(ap + cc * TH * T::P1 + i);
                a[i] = v.x, a[i + 1] = v.y, a[i + 2] = v.z, a[i + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < T::BL; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(bp + cc * T::R2 * T::P2 + i);
                bv[i] = v.x, bv[i + 1] = v.y, bv[i + 2] = v.z, bv[i + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) acc[i][j] = fmaf(a[i], bv[i + j + T::XO], acc[i][j]);
        }
        OFR_TR(2);
        __syncthreads();  // every thread is done reading stage st
        OFR_TR(3);
        if (tid == 0 && q + NST < total) {
            ofr::fence_proxy_async();
            issue(q + NST);
        }
        OFR_TR(4);
        if (++ck == nchunks) {
            // Tile finished.  Warp-private epilogue (no CTA barrier, so warps drift apart and overlap their phases):
            // for each x-displacement j the lanes park their M-pixel runs in the warp's staging buffer so that the NS
            // strips of one (plane, row) are contiguous, the warp reads the buffer back as whole 16-byte chunks in run
            // order, and every STG.128 of the warp then covers whole 128-byte lines (evict-first).  Two staging
            // buffers alternate, so one __syncwarp per step orders everything.
            // (Register -> global STG.128 straight from the accumulator layout scatters every store over 16-32 sectors
            // and cost 36 % of the kernel; a CTA-wide staged TMA-store epilogue cost more in barriers; 256-byte bulk
            // copies per run were slower than both.)
            int x0, y0, b;
            tile_origin(tile_k, x0, y0, b);
            const size_t hw = (size_t)g.Hout * g.Wout;
            float* dst[T::EPI];
            const float* src[T::EPI];
#pragma unroll
            for (int k = 0; k < T::EPI; ++k) {
                const int idx = k * (grouped ? 32 : TH * NS) + lane, r = idx / T::CPR, c = idx % T::CPR;
                const int tj_r = grouped ? tj_base + (r & 7) : D - 1, row_r = grouped ? row_base + (r >> 3) : r;
                const int y = y0 + row_r, x = x0 + c * 4;
                const bool ok = r < nruns && tj_r < D && y < g.Hout && x < g.Wout;
                dst[k] = ok ? out + ((size_t)b * D * D + (size_t)tj_r * D) * hw + (size_t)y * g.Wout + x : nullptr;
                src[k] = wstage + r * (NS * M + 4) + c * 4;
            }
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const int boff = (jcount & 1) * (T::WSTAGE / 4);
                float* buf = wstage + boff + slot;
#pragma unroll
                for (int i = 0; i < M; i += 4)
                    *reinterpret_cast<float4*>(buf + i) = make_float4(ofr::corr_out(acc[i][j], g), ofr::corr_out(acc[i + 1][j], g),
                                                                      ofr::corr_out(acc[i + 2][j], g), ofr::corr_out(acc[i + 3][j], g));
                __syncwarp(wmask);
#pragma unroll
                for (int k = 0; k < T::EPI; ++k) {
                    if (dst[k]) {
                        __stcs(reinterpret_cast<float4*>(dst[k]), *reinterpret_cast<const float4*>(src[k] + boff));
                        dst[k] += hw;
                    }
                }
                ++jcount;
            }
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) acc[i][j] = 0.f;
            ck = 0;
            ++tile_k;
            OFR_TR(5);
        }
    }
}

// Small maps (too few tiles to fill the machine: the 8x8 ... 32x32 levels of the PWC pyramid): the channel chunks of ONE
// tile are split over the KS CTAs of a thread-block cluster.  Each CTA runs the same ring / register tile over its
// chunk range, parks its partial cost volume in its own shared memory, and after a cluster barrier every CTA adds the
// KS partials of its share of the output planes through distributed shared memory in rank order (deterministic) and
// streams them out with whole-line stores.
template <int R, int M, int TH, int NS, int CK, int NST>
__global__ void __maxnreg__(224)
    corr_fwd_split_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                          float* __restrict__ out, CorrGeom g, int tiles_x, int tiles_y, int ks) {
    using T = FwdT<R, M, TH, NS, CK, NST>;
    constexpr int D = T::D;
    constexpr int NOUT = D * D * TH * T::TW;  // floats of one partial cost-volume tile
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* smem = reinterpret_cast<float*>(smem_raw + ((1024u - (ofr::smem_u32(smem_raw) & 1023u)) & 1023u));
    float* part = smem + T::STAGE * NST;  // [D*D][TH][TW]
    uint64_t* bars = reinterpret_cast<uint64_t*>(part + NOUT);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();

    const int tid = threadIdx.x;
    int tj, s, row;
    if (T::kNine) {
        if (tid < T::T1) {
            tj = tid & 7;
            const int q = tid >> 3;
            s = q % NS;
            row = q / NS;
        } else {
            tj = 8;
            const int q = tid - T::T1;
            row = q % TH;
            s = q / TH;
        }
    } else {
        const int q = tid >> 3;
        const int combo = q % (TH * NS);
        tj = (q / (TH * NS)) * 8 + (tid & 7);
        s = combo % NS;
        row = combo / NS;
    }
    const bool active = tj < D;
    const int tjc = active ? tj : D - 1;

    int t = blockIdx.x / ks;
    const int x0 = (t % tiles_x) * T::TW;
    t /= tiles_x;
    const int y0 = (t % tiles_y) * TH, b = t / tiles_y;
    const int nchunks = (g.C + CK - 1) / CK, per = (nchunks + ks - 1) / ks;
    const int cbeg = min(rank * per, nchunks), total = min(cbeg + per, nchunks) - cbeg;  // this rank's chunks

    auto issue = [&](int q) {
        const int c0 = (cbeg + q) * CK, st = q % NST;
        float* dst = smem + st * T::STAGE;
        ofr::mbar_expect_tx(&bars[st], T::STAGE * 4);
        ofr::tma_load_4d(dst, &map1, &bars[st], x0, y0, c0, b);
        ofr::tma_load_4d(dst + T::F1, &map2, &bars[st], x0 - R - T::XO, y0 - R, c0, b);
    };
    if (tid == 0) {
        ofr::tma_prefetch_desc(&map1);
        ofr::tma_prefetch_desc(&map2);
        for (int i = 0; i < NST; ++i) ofr::mbar_init(&bars[i], 1);
        ofr::mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0)
        for (int i = 0; i < NST && i < total; ++i) issue(i);

    float acc[M][D];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) acc[i][j] = 0.f;

    for (int q = 0; q < total; ++q) {
        const int st = q % NST;
        ofr::mbar_wait(&bars[st], (q / NST) & 1);
        const float* ap = smem + st * T::STAGE + row * T::P1 + s * M;
        const float* bp = smem + st * T::STAGE + T::F1 + (row + tjc) * T::P2 + s * M;
#pragma unroll 1
        for (int cc = 0; cc < CK; ++cc) {
            float a[M], bv[T::BL];
#pragma unroll
            for (int i = 0; i < M; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(ap + cc * TH * T::P1 + i);
                a[i] = v.x, a[i + 1] = v.y, a[i + 2] = v.z, a[i + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < T::BL; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(bp + cc * T::R2 * T::P2 + i);
                bv[i] = v.x, bv[i + 1] = v.y, bv[i + 2] = v.z, bv[i + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) acc[i][j] = fmaf(a[i], bv[i + j + T::XO], acc[i][j]);
        }
        if (q + NST < total) {
            __syncthreads();  // every thread is done reading stage st
            if (tid == 0) {
                ofr::fence_proxy_async();
                issue(q + NST);
            }
        }
    }
    // partial tile -> own shared memory (plane-major, dense rows)
    if (active) {
#pragma unroll
        for (int j = 0; j < D; ++j) {
            float* dst = part + (tj * D + j) * (TH * T::TW) + row * T::TW + s * M;
#pragma unroll
            for (int i = 0; i < M; i += 4)
                *reinterpret_cast<float4*>(dst + i) = make_float4(acc[i][j], acc[i + 1][j], acc[i + 2][j], acc[i + 3][j]);
        }
    }
    cluster.sync();
    // every CTA reduces the 16-byte chunks v = rank*THREADS + tid (mod ks*THREADS) of the tile over all ranks, in order
    const size_t hw = (size_t)g.Hout * g.Wout;
    constexpr int TW4 = T::TW / 4;
    for (int v = rank * T::THREADS + tid; v < NOUT / 4; v += ks * T::THREADS) {
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < ks; ++r) {
            const float4 x = *reinterpret_cast<const float4*>(cluster.map_shared_rank(part, r) + 4 * v);
            sum.x += x.x, sum.y += x.y, sum.z += x.z, sum.w += x.w;
        }
        const int plane = v / (TH * TW4), rem = v - plane * (TH * TW4);
        const int ry = rem / TW4, cx = (rem - ry * TW4) * 4;
        const int y = y0 + ry, x = x0 + cx;
        if (y < g.Hout && x < g.Wout)
            __stcs(reinterpret_cast<float4*>(out + ((size_t)b * D * D + plane) * hw + (size_t)y * g.Wout + x),
                   make_float4(ofr::corr_out(sum.x, g), ofr::corr_out(sum.y, g), ofr::corr_out(sum.z, g), ofr::corr_out(sum.w, g)));
    }
    cluster.sync();  // remote shared memory must outlive the reads
}

template <int R, int M, int TH, int NS, int CK, int NST>
int launch_fwd_split(const float* f1, const float* f2, float* out, const CorrGeom& g, int ks, cudaStream_t st) {
    using T = FwdT<R, M, TH, NS, CK, NST>;
    CUtensorMap m1, m2;
    if (!ofr::make_tmap_nchw(&m1, f1, g.B, g.C, g.H, g.W, T::P1, TH, CK)) return OFR_ENOTSUP;
    if (!ofr::make_tmap_nchw(&m2, f2, g.B, g.C, g.H, g.W, T::P2, T::R2, CK)) return OFR_ENOTSUP;
    auto kern = corr_fwd_split_kernel<R, M, TH, NS, CK, NST>;
    constexpr size_t smem = sizeof(float) * (T::STAGE * NST + T::D * T::D * TH * T::TW) + 8 * NST + 1024;
    static_assert(smem <= 227 * 1024, "split tile does not fit in shared memory");
    OFR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles_x = (g.Wout + T::TW - 1) / T::TW, tiles_y = (g.Hout + TH - 1) / TH;
    const long long ntiles = (long long)tiles_x * tiles_y * g.B;
    if (ntiles * ks >= (1ll << 31)) return OFR_ENOTSUP;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ntiles * ks));
    cfg.blockDim = dim3(T::THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ks;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    OFR_CUDA(cudaLaunchKernelEx(&cfg, kern, m1, m2, out, g, tiles_x, tiles_y, ks));
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

// channel split for a geometry with `ntiles` tiles and `nchunks` channel chunks: 1 (no split), 2, 4 or 8
int split_for(long long ntiles, int nchunks) {
    // measured on the PWC pyramid (B = 24): 8x8 maps 23.7 -> 10.0 us at ks = 4, 16x16 maps 23.4 -> 18.9 us at ks = 2,
    // 32x32 maps get slower with any split (the 83 KB partial tile dominates) -> split only while the grid stays
    // within about two thirds of the SMs
    int ks = 1;
    while (ks < 8 && ks * 2 <= nchunks && ntiles * ks * 2 * 10 <= 7ll * ofr::sm_count()) ks *= 2;
    return ks;
}

template <int R, int M, int TH, int NS, int CK, int NST>
int launch_fwd(const float* f1, const float* f2, float* out, const CorrGeom& g, cudaStream_t st) {
    using T = FwdT<R, M, TH, NS, CK, NST>;
    CUtensorMap m1, m2;
    if (!ofr::make_tmap_nchw(&m1, f1, g.B, g.C, g.H, g.W, T::P1, TH, CK)) return OFR_ENOTSUP;
    if (!ofr::make_tmap_nchw(&m2, f2, g.B, g.C, g.H, g.W, T::P2, T::R2, CK)) return OFR_ENOTSUP;
    auto kern = corr_fwd_tma_kernel<R, M, TH, NS, CK, NST>;
    OFR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM));
    const int tiles_x = (g.Wout + T::TW - 1) / T::TW, tiles_y = (g.Hout + TH - 1) / TH;
    const long long ntiles = (long long)tiles_x * tiles_y * g.B;
    const int ctas_per_sm = (int)(227 * 1024 / T::SMEM) < 1 ? 1 : (int)(227 * 1024 / T::SMEM);
    long long grid = (long long)ofr::sm_count() * (ctas_per_sm > 4 ? 4 : ctas_per_sm);
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, T::THREADS, T::SMEM, st>>>(m1, m2, out, g, tiles_x, tiles_y);
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

// ------------------------------------------------------------------------------------------------ backward
// Build knobs of the backward, kept for A/B runs (tools/build_variant.py); level-2 times, DESIGN.md 4.1:
//   OFR_BWD_BM=4    4-pixel strips = twice the warps at 96 registers: 171.5 us against 169.3 (before the two below)
//   OFR_BWD_BOX=0   one TMA box per gradient plane (9 copy instructions per step) instead of 1-2: 169.3 us against 155.5
//   OFR_BWD_EARLY=0 gradient stage released at the end of its step instead of once it is in registers: 155.5 against 152.8
#ifndef OFR_BWD_BM
#define OFR_BWD_BM 8
#endif
#ifndef OFR_BWD_EARLY
#define OFR_BWD_EARLY 1
#endif
#ifndef OFR_BWD_BOX
#define OFR_BWD_BOX 1
#endif
constexpr int kBM = OFR_BWD_BM;     // pixels per thread along x

// CB channels per CTA = CG channel groups (warps) x CN channels per thread; the thread's channels are c = k*CG + cg.
// TIC = x-displacements handled per register sweep (D must be a multiple of TIC).
template <int R, int TH, int NS, int CG, int TIC, bool SHIFTED, int NSTG, int CB>
struct BwdT {
    static constexpr int D = 2 * R + 1;
    static constexpr int CN = CB / CG;
    static constexpr int TW = NS * kBM;
    static constexpr int XO = (4 - R % 4) % 4;  // TMA box origins must be 16-byte aligned along x: start XO pixels further left
    static constexpr int P2 = pitch_for(TW + 2 * R + XO + 3);
    static constexpr int R2 = TH + 2 * R;
    static constexpr int FT = CB * R2 * P2;      // resident feature tile, floats
    // TMA needs a 16-byte aligned box origin along x, so each shifted plane of the second gradient is staged as a box
    // whose origin is rounded down to a multiple of 4 pixels (4 extra columns) and the residual 0..3 pixel shift is
    // applied when the registers are filled.
    // row pitch of a staged gradient plane: 4*odd (mod 32) floats, so the 8 rows read by a quarter-warp hit 8 bank groups
    static constexpr int GW = pitch_for(SHIFTED ? TW + 4 : TW);
    static constexpr int GP = TH * GW;            // one plane tile, floats
    static constexpr int GS = D * GP;             // one gradient stage (D planes), floats
    static constexpr int THREADS = TH * NS * CG;
    static constexpr int FH = R2 / 2;             // the feature tile arrives as two boxes of FH rows, each with its own barrier
    // One displacement row of the gradient in as few TMA boxes as possible (a copy instruction queues behind the CTA's
    // shared-memory loads, so nine of them cost the issuing thread a good part of a step): the D planes of the
    // unshifted direction are ONE box.  The shifted direction's planes are consecutive in memory in DESCENDING
    // displacement order and may share a 16-byte aligned box origin while their column shifts differ by at most 4
    // (the box is 4 columns wider than the tile): groups of four planes, the last group of five, each one box whose
    // planes land in reverse order.
    static constexpr bool GROUPED = OFR_BWD_BOX && XO == 0 && D % 4 == 1 && D >= 5;
    __host__ __device__ static constexpr int group_lo(int tii) { return tii >= D - 5 ? D - 5 : tii & ~3; }
    __host__ __device__ static constexpr int group_hi(int tii) { return tii >= D - 5 ? D - 1 : (tii & ~3) + 3; }
    __host__ __device__ static constexpr int slot(int tii) {   // stage slot that holds the plane of x-displacement tii
        return GROUPED && SHIFTED ? group_lo(tii) + (group_hi(tii) - tii) : tii;
    }
    __host__ __device__ static constexpr int shift(int tii) {  // columns between the box origin and the plane's first pixel
        return GROUPED ? tii - group_lo(tii) : (tii + XO) & 3;
    }
    static constexpr size_t SMEM = sizeof(float) * (FT + NSTG * GS) + 8 * (NSTG + 2) + 1024;
    static_assert((FT * 4) % 128 == 0 && (GP * 4) % 128 == 0 && (CB * FH * P2 * 4) % 128 == 0,
                  "TMA destinations must stay 128-byte aligned");
    static_assert(D % TIC == 0 && TH % 8 == 0 && CB % CG == 0 && R2 % 2 == 0 && NSTG >= 2 && NSTG <= D, "bad tile");
    static_assert(TH <= FH, "the first displacement row must only need the upper half of the feature tile");
    static_assert(SMEM <= 227 * 1024, "tile does not fit in shared memory");
};

// registers <- the gradient values of the x-displacements [T0, T0+TIC) of the staged displacement row
template <class T, int TIC, bool SHIFTED, int T0>
__device__ __forceinline__ void bwd_load_g(const float* __restrict__ gp, float (&gv)[TIC][kBM]) {
    if (SHIFTED) {
        // plane tii was staged with its box origin shift(tii) columns left of x0 - R + tii (a multiple of 4 pixels): the
        // value for target column s*8 + i sits at box column s*8 + i + shift(tii); aligned 128-bit loads, then select
#pragma unroll
        for (int t = 0; t < TIC; ++t) {
            const int sh = T::shift(T0 + t), a0 = sh & ~3, rem = sh & 3;  // compile-time after unrolling
            float w[kBM + 4];
#pragma unroll
            for (int i = 0; i < kBM + 4; i += 4) {
                if (i < kBM || rem) {
                    const float4 v = *reinterpret_cast<const float4*>(gp + T::slot(T0 + t) * T::GP + a0 + i);
                    w[i] = v.x, w[i + 1] = v.y, w[i + 2] = v.z, w[i + 3] = v.w;
                }
            }
#pragma unroll
            for (int i = 0; i < kBM; ++i) gv[t][i] = w[i + rem];
        }
    } else {
#pragma unroll
        for (int t = 0; t < TIC; ++t)
#pragma unroll
            for (int i = 0; i < kBM; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(gp + T::slot(T0 + t) * T::GP + i);
                gv[t][i] = v.x, gv[t][i + 1] = v.y, gv[t][i + 2] = v.z, gv[t][i + 3] = v.w;
            }
    }
}

// channels [K0, K1) of the thread: feature window x the gradient values in registers
template <class T, int CG, int TIC, int T0, int K0, int K1>
__device__ __forceinline__ void bwd_fma(const float* __restrict__ fp, const float (&gv)[TIC][kBM],
                                        float (&acc)[kBM][T::CN]) {
    constexpr int B0 = (T0 + T::XO) & ~3;                                // aligned start of the feature window
    constexpr int BL = ((T0 + T::XO + TIC + kBM - 1) - B0 + 3) / 4 * 4;  // floats to load
#pragma unroll
    for (int k = K0; k < K1; ++k) {
        float bv[BL];
#pragma unroll
        for (int i = 0; i < BL; i += 4) {
            const float4 v = *reinterpret_cast<const float4*>(fp + k * CG * T::FH * T::P2 + B0 + i);
            bv[i] = v.x, bv[i + 1] = v.y, bv[i + 2] = v.z, bv[i + 3] = v.w;
        }
#pragma unroll
        for (int t = 0; t < TIC; ++t)
#pragma unroll
            for (int i = 0; i < kBM; ++i) acc[i][k] = fmaf(gv[t][i], bv[i + T0 + t + T::XO - B0], acc[i][k]);
    }
}

// one sweep over the x-displacements [T0, T0+TIC) of the staged displacement row
template <class T, int R, int CG, int TIC, bool SHIFTED, int T0>
__device__ __forceinline__ void bwd_sweep(const float* __restrict__ gp, const float* __restrict__ fp,
                                          float (&acc)[kBM][T::CN]) {
    float gv[TIC][kBM];
    bwd_load_g<T, TIC, SHIFTED, T0>(gp, gv);
    bwd_fma<T, CG, TIC, T0, 0, T::CN>(fp, gv, acc);
}

template <class T, int R, int CG, int TIC, bool SHIFTED, int T0>
struct SweepAll {
    static __device__ __forceinline__ void run(const float* gp, const float* fp, float (&acc)[kBM][T::CN]) {
        bwd_sweep<T, R, CG, TIC, SHIFTED, T0>(gp, fp, acc);
        if constexpr (T0 + TIC < T::D) SweepAll<T, R, CG, TIC, SHIFTED, T0 + TIC>::run(gp, fp, acc);
    }
};

template <int R, int TH, int NS, int CG, int TIC, bool SHIFTED, int NSTG, int CB>
__device__ __forceinline__ void corr_bwd_tma_body(const CUtensorMap& mapF, const CUtensorMap& mapG,
                                                  const CUtensorMap& mapGq, float* __restrict__ gI, const CorrGeom& g, int tiles_x, int tiles_y,
                                                  int cblocks, int ndir) {
    using T = BwdT<R, TH, NS, CG, TIC, SHIFTED, NSTG, CB>;
    constexpr int D = T::D;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Fs = reinterpret_cast<float*>(smem_raw + ((1024u - (ofr::smem_u32(smem_raw) & 1023u)) & 1023u));
    float* Gs = Fs + T::FT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Gs + NSTG * T::GS);  // [0, NSTG): gradient stages, then the two feature halves
    uint64_t* fbar = bars + NSTG;

    auto decode = [&](int bid, int& x0, int& y0, int& c0, int& b) {
        // channel blocks fastest: the CTAs that stream the same gradient planes run together (L2 hits)
        const int cbi = bid % cblocks;
        bid /= cblocks;
        const int txi = bid % tiles_x;
        bid /= tiles_x;
        const int tyi = bid % tiles_y;
        b = bid / tiles_y;
        x0 = txi * T::TW, y0 = tyi * TH, c0 = cbi * CB;
    };
    int x0, y0, c0, b;
    decode(blockIdx.x / ndir, x0, y0, c0, b);

    // lanes: rows (fastest) x strips of ONE channel group; warps: channel groups.  A quarter-warp reads 8 different rows
    // of the feature / gradient tiles, whose pitches are 4*odd (mod 32): conflict-free 128-bit LDS; and the 32-byte
    // pieces a warp stores for one channel are the whole rows of the tile (full sectors, 8 lines per store).
    const int tid = threadIdx.x;
    const int row = tid % TH;
    const int s = (tid / TH) % NS;
    const int cg = tid / (TH * NS);

    // one displacement row (D planes) of the gradient into stage st
    auto issue_g = [&](int tji, int st) {
        float* dst = Gs + st * T::GS;
        ofr::mbar_expect_tx(&bars[st], T::GS * 4);
        if constexpr (T::GROUPED) {
            if constexpr (!SHIFTED) {
                ofr::tma_load_4d(dst, &mapG, &bars[st], x0, y0, tji * D, b);  // mapG: boxes of D planes
            } else {
                const int pl0 = (2 * R - tji) * D + 2 * R, cy = y0 + tji - R;
#pragma unroll 1
                for (int lo = 0; lo < D - 5; lo += 4)  // mapGq: boxes of 4 planes; mapG: the last box, 5 planes
                    ofr::tma_load_4d(dst + lo * T::GP, &mapGq, &bars[st], x0 - R + lo, cy, pl0 - (lo + 3), b);
                ofr::tma_load_4d(dst + (D - 5) * T::GP, &mapG, &bars[st], x0 - R + D - 5, cy, pl0 - (D - 1), b);
            }
            return;
        }
#pragma unroll 1
        for (int tii = 0; tii < D; ++tii) {
            // SHIFTED: plane of displacement -t, rows shifted by tj here, columns shifted by ti at register fill
            const int cx = SHIFTED ? x0 - R - T::XO + ((tii + T::XO) & ~3) : x0;
            const int cy = SHIFTED ? y0 + tji - R : y0;
            const int pl = SHIFTED ? (2 * R - tji) * D + (2 * R - tii) : tji * D + tii;
            ofr::tma_load_4d(dst + tii * T::GP, &mapG, &bars[st], cx, cy, pl, b);
        }
    };
    if (tid == 0) {
        ofr::tma_prefetch_desc(&mapF);
        ofr::tma_prefetch_desc(&mapG);
        ofr::tma_prefetch_desc(&mapGq);
        for (int i = 0; i < NSTG + 2; ++i) ofr::mbar_init(&bars[i], 1);
        ofr::mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        // feature tile in two halves: the first displacement row only needs the upper one
        ofr::mbar_expect_tx(&fbar[0], CB * T::FH * T::P2 * 4);
        ofr::tma_load_4d(Fs, &mapF, &fbar[0], x0 - R - T::XO, y0 - R, c0, b);
        issue_g(0, 0);
        ofr::mbar_expect_tx(&fbar[1], CB * T::FH * T::P2 * 4);
        ofr::tma_load_4d(Fs + CB * T::FH * T::P2, &mapF, &fbar[1], x0 - R - T::XO, y0 - R + T::FH, c0, b);
        for (int i = 1; i < NSTG; ++i) issue_g(i, i);
        // (Pulling later displacement rows / the next CTA's tiles into the L2 with cp.async.bulk.prefetch made the
        // kernel 10-20 % slower: the prefetches queue in the same TMA unit in front of the loads that matter.)
    }

    float acc[kBM][T::CN];
#pragma unroll
    for (int i = 0; i < kBM; ++i)
#pragma unroll
        for (int k = 0; k < T::CN; ++k) acc[i][k] = 0.f;

    ofr::mbar_wait(&fbar[0], 0);
    for (int tji = 0; tji < D; ++tji) {
        const int st = tji % NSTG;
        if (tji == 1) ofr::mbar_wait(&fbar[1], 0);  // rows row + tji, tji >= 1, reach into the lower half
        ofr::mbar_wait(&bars[st], (tji / NSTG) & 1);
        const float* gp = Gs + st * T::GS + row * T::GW + s * kBM;
        // feature tile layout: [half][channel][FH rows][P2] (two TMA boxes)
        const int rr = row + tji, half = rr >= T::FH ? 1 : 0;
        const float* fp = Fs + half * (CB * T::FH * T::P2) + cg * (T::FH * T::P2) + (rr - half * T::FH) * T::P2 + s * kBM;
        if constexpr (OFR_BWD_EARLY && TIC == D) {
            // the whole displacement row fits in registers: the stage is free again once every thread has loaded it
            // (the first channel's arithmetic consumes all of it, so the loads have landed at the barrier), and its
            // refill gets almost two steps instead of one to arrive
            float gv[TIC][kBM];
            bwd_load_g<T, TIC, SHIFTED, 0>(gp, gv);
            bwd_fma<T, CG, TIC, 0, 0, 1>(fp, gv, acc);
            if (tji + NSTG < D) {
                __syncthreads();
                if (tid == 0) {
                    ofr::fence_proxy_async();
                    issue_g(tji + NSTG, st);
                }
            }
            bwd_fma<T, CG, TIC, 0, 1, T::CN>(fp, gv, acc);
        } else {
            SweepAll<T, R, CG, TIC, SHIFTED, 0>::run(gp, fp, acc);
            if (tji + NSTG < D) {
                __syncthreads();  // every thread is done reading stage st
                if (tid == 0) {
                    ofr::fence_proxy_async();
                    issue_g(tji + NSTG, st);
                }
            }
        }
    }

    const int y = y0 + row, xb = x0 + s * kBM;
    if (y < g.H) {
        const long long hw = (long long)g.H * g.W;
#pragma unroll
        for (int k = 0; k < T::CN; ++k) {
            const int c = c0 + k * CG + cg;
            if (c >= g.C) continue;
            float* o = gI + ((long long)b * g.C + c) * hw + (long long)y * g.W;
            if (xb + kBM <= g.W) {
#pragma unroll
                for (int i = 0; i < kBM; i += 4)
                    __stcs(reinterpret_cast<float4*>(o + xb + i), make_float4(acc[i][k] * g.scale, acc[i + 1][k] * g.scale, acc[i + 2][k] * g.scale, acc[i + 3][k] * g.scale));
            } else {
#pragma unroll
                for (int i = 0; i < kBM; ++i)
                    if (xb + i < g.W) o[xb + i] = acc[i][k] * g.scale;
            }
        }
    }
}

// Both gradients in ONE launch: the lowest bit of blockIdx.x picks the direction (0: gI1 from f2, 1: gI2 from f1), so
// the two CTAs that stream the same gradient planes are neighbours in launch order and the second read is an L2 hit
// (with the direction on blockIdx.y ncu still showed the planes read twice from DRAM); on small maps the fixed
// launch / first-load latency is paid once instead of twice.  `first` = the direction when only one gradient is wanted.
template <int R, int TH, int NS, int CG, int TIC, int NSTG, int CB>
__global__ void __launch_bounds__(BwdT<R, TH, NS, CG, TIC, true, NSTG, CB>::THREADS, 1)
    corr_bwd_tma_kernel(const __grid_constant__ CUtensorMap mapF2, const __grid_constant__ CUtensorMap mapG0,
                        const __grid_constant__ CUtensorMap mapF1, const __grid_constant__ CUtensorMap mapG1,
                        const __grid_constant__ CUtensorMap mapG1q, float* __restrict__ gI1, float* __restrict__ gI2, CorrGeom g, int tiles_x, int tiles_y,
                        int cblocks, int ndir, int first) {
    if ((int)(blockIdx.x % ndir) + first == 0)
        corr_bwd_tma_body<R, TH, NS, CG, TIC, false, NSTG, CB>(mapF2, mapG0, mapG0, gI1, g, tiles_x, tiles_y, cblocks, ndir);
    else
        corr_bwd_tma_body<R, TH, NS, CG, TIC, true, NSTG, CB>(mapF1, mapG1, mapG1q, gI2, g, tiles_x, tiles_y, cblocks, ndir);
}

template <int R, int TH, int NS, int CG, int TIC, int NSTG, int CB>
int launch_bwd(const float* f1, const float* f2, const float* gO, float* gI1, float* gI2, const CorrGeom& g,
               cudaStream_t st) {
    using T0 = BwdT<R, TH, NS, CG, TIC, false, NSTG, CB>;
    using T1 = BwdT<R, TH, NS, CG, TIC, true, NSTG, CB>;
    CUtensorMap mF2, mG0, mF1, mG1, mG1q;
    if (!ofr::make_tmap_nchw(&mF2, f2, g.B, g.C, g.H, g.W, T0::P2, T0::FH, CB)) return OFR_ENOTSUP;
    if (!ofr::make_tmap_nchw(&mF1, f1, g.B, g.C, g.H, g.W, T1::P2, T1::FH, CB)) return OFR_ENOTSUP;
    if (!ofr::make_tmap_nchw(&mG0, gO, g.B, T0::D * T0::D, g.Hout, g.Wout, T0::GW, TH, T0::GROUPED ? T0::D : 1))
        return OFR_ENOTSUP;
    if (!ofr::make_tmap_nchw(&mG1, gO, g.B, T1::D * T1::D, g.Hout, g.Wout, T1::GW, TH, T1::GROUPED ? 5 : 1))
        return OFR_ENOTSUP;
    if (!ofr::make_tmap_nchw(&mG1q, gO, g.B, T1::D * T1::D, g.Hout, g.Wout, T1::GW, TH, T1::GROUPED ? 4 : 1))
        return OFR_ENOTSUP;
    auto kern = corr_bwd_tma_kernel<R, TH, NS, CG, TIC, NSTG, CB>;
    constexpr size_t smem = T0::SMEM > T1::SMEM ? T0::SMEM : T1::SMEM;
    OFR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles_x = (g.W + T0::TW - 1) / T0::TW, tiles_y = (g.H + TH - 1) / TH;
    const int cblocks = (g.C + CB - 1) / CB;
    const long long nblk = (long long)tiles_x * tiles_y * cblocks * g.B;
    if (nblk >= (1ll << 30)) return OFR_ENOTSUP;
    const int ndir = (gI1 ? 1 : 0) + (gI2 ? 1 : 0);
    if (ndir == 0) return OFR_OK;
    kern<<<(unsigned)(nblk * ndir), T0::THREADS, smem, st>>>(mF2, mG0, mF1, mG1, mG1q, gI1, gI2, g, tiles_x, tiles_y,
                                                              cblocks,
                                                              ndir, gI1 ? 0 : 1);
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

int dispatch_bwd4(const float* f1, const float* f2, const float* gO, float* gI1, float* gI2, const CorrGeom& g,
                  cudaStream_t st) {
    // 8 x 32 tiles: 112 KB of shared memory per CTA -> two CTAs per SM, so one CTA's feature-tile load and epilogue
    // overlap the other's arithmetic
    // (16 x 32 tiles with 16 channels per CTA -- same shared memory, half the halo, twice the arithmetic per stage --
    // measured 8 % slower; a 3-deep ring at one CTA per SM 30 % slower.)
    if (g.W > 16) return launch_bwd<4, 8, 32 / kBM, 4, 9, 2, 32>(f1, f2, gO, gI1, gI2, g, st);
    // small maps: the whole gradient tile (all 9 displacement rows) is in flight at once -- latency, not bandwidth --
    // and there are too few tiles to fill the machine, so the channels are spread over more warps (4 / 2 per thread)
    if (g.W > 8) return launch_bwd<4, 8, 16 / kBM, 8, 9, 9, 32>(f1, f2, gO, gI1, gI2, g, st);  // (16 groups: 21.6 -> 24.1 us)
    return launch_bwd<4, 8, 8 / kBM, 16, 9, 9, 32>(f1, f2, gO, gI1, gI2, g, st);
}

bool applicable(const CorrGeom& g, int R, const void* a, const void* b, const void* c) {
    return g.k == 1 && g.s1 == 1 && g.s2 == 1 && g.md == R && g.pad == R && (g.W % 4) == 0 && ofr::aligned16(a) &&
           ofr::aligned16(b) && ofr::aligned16(c) && ofr::encode_tiled_fn() != nullptr;
}

}  // namespace

namespace ofr {

int corr_fwd_tma(const float* f1, const float* f2, float* out, const CorrGeom& g, cudaStream_t st) {
    if (applicable(g, 4, f1, f2, out)) {
        // 7 rows: 7 warps own tj = 0..7 and one warp owns tj = 8 -> 8 warps; a 9-warp CTA would be allocated registers
        // for 12 warps and could not keep the 144 accumulators of the 16-pixel strip
        // (8 x 32 tiles at two CTAs per SM and 8-pixel strips with 288 threads both measured slower.)
        if (g.W > 32) return launch_fwd<4, 16, 7, 4, 8, 3>(f1, f2, out, g, st);   // 7 x 64 tile, 252 threads
        const int nchunks = (g.C + 7) / 8;
        const long long tiles_y8 = (g.Hout + 7) / 8;
        // Small maps are latency bound (one to three warps per scheduler): narrower strips = more threads per tile.
        // Measured at B = 24: 32x32 maps 27.0 -> 21.9 us with 8-pixel strips, 16x16 maps 19.1 -> 12.9 us and 8x8 maps
        // 10.3 -> 9.1 us with 4-pixel strips.
        const int ks = split_for(tiles_y8 * g.B, nchunks);
        if (g.W > 16) {  // 8 x 32 tile, 8-pixel strips, 288 threads
            if (ks > 1) return launch_fwd_split<4, 8, 8, 4, 8, 3>(f1, f2, out, g, ks, st);
            return launch_fwd<4, 8, 8, 4, 8, 3>(f1, f2, out, g, st);
        }
        if (g.W > 8) {   // 8 x 16 tile, 4-pixel strips, 288 threads
            if (ks > 1) return launch_fwd_split<4, 4, 8, 4, 8, 4>(f1, f2, out, g, ks, st);
            return launch_fwd<4, 4, 8, 4, 8, 4>(f1, f2, out, g, st);
        }
        // 8 x 8 tile, 4-pixel strips, 144 threads
        if (ks > 1) return launch_fwd_split<4, 4, 8, 2, 8, 4>(f1, f2, out, g, ks, st);
        return launch_fwd<4, 4, 8, 2, 8, 4>(f1, f2, out, g, st);
    }
    if (applicable(g, 10, f1, f2, out)) {
        // the 21 x 21 window is FMA-bound (59 FLOP/B): Gram tiles on the tensor cores (3xTF32, TMEM accumulators)
        const int rc = ofr::corr_fwd_umma(f1, f2, out, g, st);
        if (rc != OFR_ENOTSUP) return rc;
        return launch_fwd<10, 8, 4, 2, 8, 3>(f1, f2, out, g, st);                 // 4 x 16 tile, 192 threads (3 x 8 tj lanes)
    }
    return OFR_ENOTSUP;
}

int corr_bwd_tma(const float* f1, const float* f2, const float* gO, float* gI1, float* gI2, const CorrGeom& g,
                 cudaStream_t st) {
    if ((gI1 && !ofr::aligned16(gI1)) || (gI2 && !ofr::aligned16(gI2))) return OFR_ENOTSUP;
    if (applicable(g, 4, f1, f2, gO)) {
        return dispatch_bwd4(f1, f2, gO, gI1, gI2, g, st);
    }
    if (applicable(g, 10, f1, f2, gO)) {
        {  // FlowNetC window: banded GEMMs on the tensor cores (correlation_umma.cu)
            const int rc = ofr::corr_bwd_umma(f1, f2, gO, gI1, gI2, g, st);
            if (rc != OFR_ENOTSUP) return rc;
        }
        // R = 10 (class images of FlowNetC, 16 x 16 or 32 x 32): 8 x 16 tiles, 8 channel groups x 4 channels, 7
        // x-displacements per register sweep
        return launch_bwd<10, 8, 16 / kBM, 8, 7, 2, 32>(f1, f2, gO, gI1, gI2, g, st);
    }
    return OFR_ENOTSUP;
}

}  // namespace ofr
