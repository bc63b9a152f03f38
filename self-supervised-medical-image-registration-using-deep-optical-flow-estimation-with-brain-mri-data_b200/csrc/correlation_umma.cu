// Correlation forward for the FlowNetC window on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Scope: the stride-1 class-image problem the polyphase split of Correlation(20,1,20,1,2) produces
// (flownet2/networks/FlowNetC.py:31,88): k = 1, s1 = s2 = 1, pad = md = 10, window 21 x 21 = 441 displacements, C = 256.
// This is the configuration SURVEY.md 8(a1) marks FMA-bound (59 FLOP/B); the register-tiled FMA kernel
// (correlation_tma.cu, R = 10) reaches 34 % of the FP32 pipe because the 21 x 21 halo is 13x the tile.
//
// Formulation: for a tile of 128 pixels p (8 rows x 16) the cost volume is a band of the Gram matrix
//     S[p, q] = sum_c f1[c, p] * f2[c, q],          out[(dy, dx), p] = S[p, p + (dy, dx)] / C
// over the 28 x 36 neighbourhood q of the tile.  S is computed as four GEMMs M = 128, N = 256 (7 neighbourhood rows of
// 36 columns = 252, padded), K = C in stages of 16 channels:
//   * operands are written by the CTA's threads into the canonical K-major no-swizzle layout (8 x 16 B core matrices),
//     out-of-image neighbours as zeros (= the operator's zero padding), each value split into a TF32 "hi" part
//     (cvt.rna) and the exact remainder "lo"; the product is hi*hi + hi*lo + lo*hi (3xTF32) accumulated in FP32 in TMEM:
//     measured max error 4e-6 on sums of magnitude 8 (tools/micro/tf32_umma.cu), i.e. < 1e-7 after the 1/C scaling --
//     inside the 1e-5 bound of the operator; one TF32 product alone (7e-4) is not;
//   * one elected thread issues tcgen05.mma (cta_group::1, kind::tf32); tcgen05.commit on an mbarrier per stage buffer
//     tells the threads when a buffer may be overwritten; the global loads of stage s+1 are in flight while stage s is
//     written and multiplied;
//   * epilogue: warps 0-3 read their 32 TMEM lanes (= pixels) one neighbourhood row at a time (36 columns), pick the
//     21-wide band through a per-warp shared-memory transposition (the band offset depends on the pixel's column, a
//     register index must not) and store whole 64-byte row segments of each displacement plane.
// Two CTAs fit an SM (100 KB of shared memory, 256 of the 512 TMEM columns each): one CTA's operand staging and
// epilogue overlap the other's MMAs.  44 % of S is wanted and the split costs 3 MMAs: 6.9x the useful MACs on a pipe
// measured at 1096 TFLOP/s (tf32, N = 256) against 72 TFLOP/s of FP32 FMA.
#include "corr_internal.cuh"

namespace {

constexpr int UM = 128;          // pixels per tile = MMA M
constexpr int UN = 256;          // neighbourhood positions per chunk (252 used) = MMA N
constexpr int UKT = 16;          // channels per stage (two K = 8 MMA steps)
constexpr int UR = 10;           // displacement radius
constexpr int UD = 2 * UR + 1;   // 21
constexpr int UTH = 8, UTW = 16; // tile
constexpr int UNW = UTW + 2 * UR;           // 36 neighbourhood columns
constexpr int UROWS = 7;                    // neighbourhood rows per chunk
constexpr int UCHUNKS = (UTH + 2 * UR) / UROWS;  // 4
constexpr int UTHREADS = 256;
constexpr int A_FLOATS = UM * UKT, B_FLOATS = UN * UKT;
constexpr int STAGE_FLOATS = 2 * (A_FLOATS + B_FLOATS);  // hi + lo
constexpr int EBUF_FLOATS = 4 * UD * 32;
constexpr size_t USMEM = sizeof(float) * (2 * STAGE_FLOATS + EBUF_FLOATS) + 1024;
static_assert(UCHUNKS * UROWS == UTH + 2 * UR && UROWS * UNW <= UN, "chunking");

// Where the class images live.  full = 0: class layout [B*4][C][H][W] (the polyphase split has been run).  full = 1: the
// operator's own tensors [B][C][2H][2W]; class q = 2*py + px of image b is read in place with stride 2 (pixel (y, x) of
// the class image is (2y + py, 2x + px)), so no split pass and no second copy of the inputs exists.  The four classes
// of an image are neighbours in launch order: the half-used sectors of one class are L2 hits for the next.
struct SrcMap {
    int full;      // inputs in place (stride 2)
    int dst_full;  // outputs written in place too (no merge pass)
};
struct SrcView {
    size_t base;   // element offset of (channel 0, pixel (0,0)) of this class image
    int cs, rs, xs;  // channel / row / column strides in elements
};
__device__ __forceinline__ SrcView src_view(const SrcMap& m, int b4, int planes, int H, int W) {
    SrcView v;
    if (m.full) {
        const int b = b4 >> 2, py = (b4 >> 1) & 1, px = b4 & 1;
        v.cs = 4 * H * W, v.rs = 4 * W, v.xs = 2;
        v.base = (size_t)b * planes * v.cs + py * 2 * W + px;
    } else {
        v.cs = H * W, v.rs = W, v.xs = 1;
        v.base = (size_t)b4 * planes * v.cs;
    }
    return v;
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, no swizzle: element (row r, 16-byte chunk j along K) at j * LBO + r * 16
// (core matrix = 8 rows x 16 B contiguous, SBO = 128 B between 8-row groups)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version of sm_100
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

// wait for the phase with the given parity; a barrier that never completes (it cannot, unless the MMA was rejected)
// traps after ~2 s instead of hanging the device
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(s32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    uint32_t hb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
    hi = __uint_as_float(hb);
    lo = x - hi;  // exact
}

__global__ void __launch_bounds__(UTHREADS, 2) corr_fwd_umma_kernel(const float* __restrict__ f1,
                                                                   const float* __restrict__ f2,
                                                                   float* __restrict__ out, ofr::CorrGeom g,
                                                                   int tiles_x, int tiles_y, SrcMap sm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* stages = reinterpret_cast<float*>(smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u));
    float* ebuf = stages + 2 * STAGE_FLOATS;
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int bid = blockIdx.x;
    const int txi = bid % tiles_x;
    bid /= tiles_x;
    const int tyi = bid % tiles_y;
    const int b = bid / tiles_y;
    const int x0 = txi * UTW, y0 = tyi * UTH;
    const int H = g.H, W = g.W, C = g.C;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)),
                     "r"((uint32_t)UN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    // D = F32 (bit 4), A = B = TF32 (2 at bits 7, 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(UN >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
    constexpr uint32_t LBO_A = UM * 16, LBO_B = UN * 16;

    // staging tasks: one task = one row of an operand (a pixel) x 4 consecutive channels = one 16-byte chunk
    //   A: 128 pixels x 4 chunks = 512 tasks (2 per thread);  B: 256 positions x 4 chunks = 1024 tasks (4 per thread)
    const SrcView sv = src_view(sm, b, C, H, W);
    const SrcView dv = src_view(SrcMap{sm.dst_full, 0}, b, UD * UD, H, W);
    const int cs = sv.cs;
    const float* f1b = f1 + sv.base;
    const float* f2b = f2 + sv.base;
    int a_src[2], a_dst[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int t = tid + UTHREADS * i, m = t & (UM - 1), kq = t >> 7;
        a_src[i] = kq * 4 * cs + (y0 + (m >> 4)) * sv.rs + (x0 + (m & 15)) * sv.xs;
        a_dst[i] = kq * (UM * 4) + m * 4;
    }
    int b_dst[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = tid + UTHREADS * i, j = t & (UN - 1), kq = t >> 8;
        b_dst[i] = kq * (UN * 4) + j * 4;
    }

    const int nstages = C / UKT;
    uint32_t it = 0;  // stages issued so far (buffer = it & 1)
    for (int n = 0; n < UCHUNKS; ++n) {
        // a chunk whose 7 neighbourhood rows all lie outside the image is all zero padding: no GEMM, the epilogue
        // writes the zeros (top / bottom tiles: 2 of the 16 chunks of a 32-row class image)
        const bool chunk_live = y0 - UR + n * UROWS + UROWS > 0 && y0 - UR + n * UROWS < H;
        int b_src[4];  // element offset of the task's neighbour inside one channel plane, or -1 (zero padding)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int t = tid + UTHREADS * i, j = t & (UN - 1), kq = t >> 8;
            const int rc = j / UNW, cc = j - rc * UNW;
            const int yy = y0 - UR + n * UROWS + rc, xx = x0 - UR + cc;
            const bool in = j < UROWS * UNW && (unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W;
            b_src[i] = in ? kq * 4 * cs + yy * sv.rs + xx * sv.xs : -1;
        }
        float pa[2][4], pb[4][4];
        auto load = [&](int s) {
            const float* p1 = f1b + (size_t)s * UKT * cs;
            const float* p2 = f2b + (size_t)s * UKT * cs;
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) pa[i][c] = __ldg(p1 + a_src[i] + c * cs);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) pb[i][c] = b_src[i] >= 0 ? __ldg(p2 + b_src[i] + c * cs) : 0.f;
        };
        if (chunk_live) load(0);
        for (int s = 0; chunk_live && s < nstages; ++s, ++it) {
            const uint32_t buf = it & 1u;
            float* sAh = stages + buf * STAGE_FLOATS;
            float* sAl = sAh + A_FLOATS;
            float* sBh = sAl + A_FLOATS;
            float* sBl = sBh + B_FLOATS;
            if (it >= 2) mbar_wait_or_trap(&bars[buf], ((it >> 1) - 1u) & 1u);  // the MMAs that read this buffer are done
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float4 h, l;
                split_tf32(pa[i][0], h.x, l.x);
                split_tf32(pa[i][1], h.y, l.y);
                split_tf32(pa[i][2], h.z, l.z);
                split_tf32(pa[i][3], h.w, l.w);
                *reinterpret_cast<float4*>(sAh + a_dst[i]) = h;
                *reinterpret_cast<float4*>(sAl + a_dst[i]) = l;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 h, l;
                split_tf32(pb[i][0], h.x, l.x);
                split_tf32(pb[i][1], h.y, l.y);
                split_tf32(pb[i][2], h.z, l.z);
                split_tf32(pb[i][3], h.w, l.w);
                *reinterpret_cast<float4*>(sBh + b_dst[i]) = h;
                *reinterpret_cast<float4*>(sBl + b_dst[i]) = l;
            }
            if (s + 1 < nstages) load(s + 1);  // in flight while this stage is multiplied
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int ks = 0; ks < UKT / 8; ++ks) {
                    const uint64_t dAh = umma_desc(s32(sAh) + ks * 2 * LBO_A, LBO_A), dAl = umma_desc(s32(sAl) + ks * 2 * LBO_A, LBO_A);
                    const uint64_t dBh = umma_desc(s32(sBh) + ks * 2 * LBO_B, LBO_B), dBl = umma_desc(s32(sBl) + ks * 2 * LBO_B, LBO_B);
                    umma_tf32(tmem_d, dAh, dBh, idesc, (s | ks) ? 1u : 0u);
                    umma_tf32(tmem_d, dAh, dBl, idesc, 1u);
                    umma_tf32(tmem_d, dAl, dBh, idesc, 1u);
                }
                // arrives on the buffer's barrier when these (and all earlier) MMAs have completed
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bars[buf]))
                             : "memory");
            }
        }
        // the chunk's accumulators are complete when the last stage's commit has arrived
        if (chunk_live) mbar_wait_or_trap(&bars[(it - 1u) & 1u], ((it - 1u) >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp < 4) {
            const int yp = 2 * warp + (lane >> 4), xp = lane & 15;
            float* eb = ebuf + warp * (UD * 32);
            const uint32_t trow = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
            for (int rc = 0; rc < UROWS; ++rc) {
                const int dy = n * UROWS + rc - yp;
                const bool valid = (unsigned)dy < (unsigned)UD;
                uint32_t v[UNW];
                const uint32_t ta = trow + rc * UNW;
                if (!chunk_live) {  // (CTA-uniform)
                    if (valid) {
                        float* o = out + dv.base + (size_t)(dy * UD) * dv.cs + (y0 + yp) * dv.rs + (x0 + xp) * dv.xs;
#pragma unroll
                        for (int dx = 0; dx < UD; ++dx) __stcs(o + (size_t)dx * dv.cs, 0.f);
                    }
                    continue;
                }
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(ta)
                    : "memory");
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35])
                             : "r"(ta + 32)
                             : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // band of this pixel: neighbourhood columns xp .. xp + 20 -> displacement column dx = j - xp
#pragma unroll
                for (int j = 0; j < UNW; ++j) {
                    const int dx = j - xp;
                    if (valid && (unsigned)dx < (unsigned)UD) eb[dx * 32 + lane] = ofr::corr_out(__uint_as_float(v[j]), g);
                }
                __syncwarp();
                if (valid) {
                    float* o = out + dv.base + (size_t)(dy * UD) * dv.cs + (y0 + yp) * dv.rs + (x0 + xp) * dv.xs;
#pragma unroll
                    for (int dx = 0; dx < UD; ++dx) __stcs(o + (size_t)dx * dv.cs, eb[dx * 32 + lane]);
                }
                __syncwarp();
            }
        }
        // the next chunk's first MMA overwrites the accumulators: every reader must be done
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)UN) : "memory");
}


// ------------------------------------------------------------------------------------------------ backward
// Both gradients are the same GEMM with the roles of the operands exchanged.  For a tile of 128 target pixels t and its
// 28 x 36 neighbourhood n:
//     gI1[c, t] = s * sum_n G1[t, n] * f2[c, n],   G1[t, n] = gO[plane(n - t)][t]
//     gI2[c, t] = s * sum_n G2[t, n] * f1[c, n],   G2[t, n] = gO[plane(t - n)][n]        (n inside the image)
// M = 128 targets, N = C channels (<= 256: the accumulators are C TMEM columns), K = the neighbourhood, walked one row
// (36 positions, padded to 40) at a time as two stages of 24 + 16.  The banded operand G is built in shared memory: the
// gradient rows a neighbourhood row needs (8 tile rows x 21 x-displacements x 16 pixels, whole 64-byte segments of the
// planes) are loaded into a small buffer R[yp][dx][xp], and A[t][cc] = R[yp][cc - xp][xp] for 0 <= cc - xp <= 20 (bank
// conflict free: -15 * xp mod 32 is a permutation).  Features (B operand) are read as rows of 24 / 16 consecutive
// pixels per channel.  Everything else (3xTF32 split, one MMA thread, commit barriers, register prefetch of the next
// stage) is the forward's.  Epilogue: TMEM lanes = pixels, columns = channels -> 64-byte row segments per channel.
constexpr int BK0 = 24, BK1 = 16;                    // K of the two stages of a neighbourhood row (36 -> 40 positions)
constexpr int BTHREADS = 512;
constexpr int BA_FLOATS = UM * BK0;                  // A stage (sized for the larger half)
constexpr int BB_FLOATS = (256 + 1) * BK0;           // B stage (chunk stride padded by one row, see LBO_B)
constexpr int BSTAGE_FLOATS = 2 * (BA_FLOATS + BB_FLOATS);
constexpr int BR_FLOATS = UTH * UD * UTW;            // 2688
constexpr int BR_PER_THREAD = (BR_FLOATS + BTHREADS - 1) / BTHREADS;  // 6
constexpr int BB_TASKS = 256 * (BK0 / 4) / BTHREADS;                  // 3 feature tasks per thread and stage
constexpr size_t BSMEM = sizeof(float) * (2 * BSTAGE_FLOATS + BR_FLOATS) + 1024;
static_assert(BSMEM <= 227 * 1024, "backward stages do not fit");

template <int DIR>
__device__ __forceinline__ void corr_bwd_umma_body(const float* __restrict__ feat, const float* __restrict__ gO,
                                                   float* __restrict__ gI, const ofr::CorrGeom& g, int tiles_x,
                                                   int tiles_y, int bid, const SrcMap& sm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* stages = reinterpret_cast<float*>(smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u));
    float* Rs = stages + 2 * BSTAGE_FLOATS;
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int txi = bid % tiles_x;
    bid /= tiles_x;
    const int tyi = bid % tiles_y;
    const int b = bid / tiles_y;
    const int x0 = txi * UTW, y0 = tyi * UTH;
    const int H = g.H, W = g.W, C = g.C;
    const uint32_t ncols = C <= 32 ? 32u : C <= 64 ? 64u : C <= 128 ? 128u : 256u;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
    // B chunks are (C + 1) rows apart: a warp stores the 6 (4) chunks of ~5 (8) consecutive channels at once, and with a
    // stride of exactly C rows all chunks of a channel fall on the same four banks (6-way conflict on every store)
    const uint32_t LBO_A = UM * 16, LBO_B = (uint32_t)(C + 1) * 16;

    const SrcView fv = src_view(sm, b, C, H, W), gv = src_view(sm, b, UD * UD, H, W);
    const float* fb = feat + fv.base;
    const float* gb = gO + gv.base;

    // ---- per-thread constants of the staging tasks (everything that does not depend on the neighbourhood row) ----
    // R: element idx = (yp * 21 + dxx) * 16 + xp of the gradient buffer of a row
    int r_off[BR_PER_THREAD], r_yp[BR_PER_THREAD];  // r_off < 0: never loaded
#pragma unroll
    for (int i = 0; i < BR_PER_THREAD; ++i) {
        const int idx = tid + BTHREADS * i;
        const int xp = idx & 15, rest = idx >> 4, yp = rest / UD, dxx = rest - yp * UD;
        r_yp[i] = yp;
        if (idx >= BR_FLOATS) {
            r_off[i] = -1;
        } else if (DIR == 0) {  // gO[(dyy * 21 + dxx)][y0 + yp][x0 + xp]
            r_off[i] = dxx * gv.cs + (y0 + yp) * gv.rs + (x0 + xp) * gv.xs;
        } else {                // gO[(20 - dyy) * 21 + 20 - dxx][y0 - 10 + rr][x0 - 10 + xp + dxx]
            const int xx = x0 - UR + xp + dxx;
            r_off[i] = (unsigned)xx < (unsigned)W ? (UD - 1 - dxx) * gv.cs + xx * gv.xs : -1;
        }
    }
    // B: task = (channel c, chunk kq of 4 neighbourhood columns); per half h its own decomposition
    int b_off[2][BB_TASKS], b_dst[2][BB_TASKS];
    unsigned b_msk[2][BB_TASKS];  // bit e: column of element e is inside the image (and the task exists)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int nk = h ? BK1 / 4 : BK0 / 4, cbase = h ? BK0 : 0;
#pragma unroll
        for (int i = 0; i < BB_TASKS; ++i) {
            // 8 consecutive tasks = 8 consecutive channels of one chunk: a quarter-warp's 16-byte stores then fill one
            // 128-byte core-matrix row group (chunk-fastest order put two tasks of a quarter-warp on the same banks)
            const int t = tid + BTHREADS * i;
            const int grp = t >> 3, cblk = grp / nk, kq = grp - cblk * nk, c = cblk * 8 + (t & 7);
            const int xx0 = x0 - UR + cbase + kq * 4;
            unsigned m = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (c < C && cbase + kq * 4 + e < UNW && (unsigned)(xx0 + e) < (unsigned)W) m |= 1u << e;
            b_msk[h][i] = m;
            b_off[h][i] = c * fv.cs + xx0 * fv.xs;
            b_dst[h][i] = c < C ? kq * ((C + 1) * 4) + c * 4 : -1;
        }
    }

    float pr[BR_PER_THREAD];
    auto load_r = [&](int rr) {
#pragma unroll
        for (int i = 0; i < BR_PER_THREAD; ++i) {
            const int dyy = rr - r_yp[i];
            float v = 0.f;
            if (r_off[i] >= 0 && (unsigned)dyy < (unsigned)UD) {
                if (DIR == 0) {
                    v = __ldg(gb + (size_t)(dyy * UD) * gv.cs + r_off[i]);
                } else {
                    const int yy = y0 - UR + rr;
                    if ((unsigned)yy < (unsigned)H)
                        v = __ldg(gb + (size_t)((UD - 1 - dyy) * UD) * gv.cs + yy * gv.rs + r_off[i]);
                }
            }
            pr[i] = v;
        }
    };
    float pb[2][BB_TASKS][4];  // set h holds the features of stage (row, h)
    auto load_b = [&](int rr, int h, float (&dst)[BB_TASKS][4]) {
        const int yy = y0 - UR + rr;
        const bool row_in = (unsigned)yy < (unsigned)H;
        const float* rowp = fb + yy * fv.rs;
#pragma unroll
        for (int i = 0; i < BB_TASKS; ++i) {
            const unsigned m = row_in ? b_msk[h][i] : 0u;
#pragma unroll
            for (int e = 0; e < 4; ++e) dst[i][e] = (m >> e) & 1u ? __ldg(rowp + b_off[h][i] + e * fv.xs) : 0.f;
        }
    };

    // neighbourhood rows outside the image multiply zero features (gI1) / zero gradients (gI2): only the rows inside
    // are walked (18 of the 28 for the top and bottom tiles of a 32-row class image)
    const int r_lo = max(0, UR - y0), nrows = min(UTH + 2 * UR, H + UR - y0);
    load_r(r_lo);
    load_b(r_lo, 0, pb[0]);
    load_b(r_lo, 1, pb[1]);
    uint32_t it = 0;
#pragma unroll 1
    for (int rr = r_lo; rr < nrows; ++rr) {
#pragma unroll
        for (int h = 0; h < 2; ++h, ++it) {
            const uint32_t buf = it & 1u;
            const int kt = h ? BK1 : BK0, nk = kt / 4, cbase = h ? BK0 : 0;
            float* sAh = stages + buf * BSTAGE_FLOATS;
            float* sAl = sAh + BA_FLOATS;
            float* sBh = sAl + BA_FLOATS;
            float* sBl = sBh + BB_FLOATS;
            if (it >= 2) mbar_wait_or_trap(&bars[buf], ((it >> 1) - 1u) & 1u);
            if (h == 0) {
                // (the previous row's A tiles were built before the last __syncthreads: R is free)
#pragma unroll
                for (int i = 0; i < BR_PER_THREAD; ++i) {
                    const int idx = tid + BTHREADS * i;
                    if (idx < BR_FLOATS) Rs[idx] = pr[i];
                }
                __syncthreads();
                if (rr + 1 < nrows) load_r(rr + 1);  // two stages ahead of its use
            }
            // A: task = (target t, chunk kq): 128 * nk tasks
            for (int t = tid; t < UM * nk; t += BTHREADS) {
                const int m = t & (UM - 1), kq = t >> 7;
                const int yp = m >> 4, xp = m & 15;
                const float* rp = Rs + yp * (UD * UTW) + xp;
                const int d0 = cbase + kq * 4 - xp;
                float4 hv, lv;
                split_tf32((unsigned)(d0 + 0) < (unsigned)UD ? rp[(d0 + 0) * UTW] : 0.f, hv.x, lv.x);
                split_tf32((unsigned)(d0 + 1) < (unsigned)UD ? rp[(d0 + 1) * UTW] : 0.f, hv.y, lv.y);
                split_tf32((unsigned)(d0 + 2) < (unsigned)UD ? rp[(d0 + 2) * UTW] : 0.f, hv.z, lv.z);
                split_tf32((unsigned)(d0 + 3) < (unsigned)UD ? rp[(d0 + 3) * UTW] : 0.f, hv.w, lv.w);
                *reinterpret_cast<float4*>(sAh + kq * (UM * 4) + m * 4) = hv;
                *reinterpret_cast<float4*>(sAl + kq * (UM * 4) + m * 4) = lv;
            }
            // B: the prefetched features of this stage
#pragma unroll
            for (int i = 0; i < BB_TASKS; ++i) {
                if (b_dst[h][i] >= 0) {
                    float4 hv, lv;
                    split_tf32(pb[h][i][0], hv.x, lv.x);
                    split_tf32(pb[h][i][1], hv.y, lv.y);
                    split_tf32(pb[h][i][2], hv.z, lv.z);
                    split_tf32(pb[h][i][3], hv.w, lv.w);
                    *reinterpret_cast<float4*>(sBh + b_dst[h][i]) = hv;
                    *reinterpret_cast<float4*>(sBl + b_dst[h][i]) = lv;
                }
            }
            if (rr + 1 < nrows) load_b(rr + 1, h, pb[h]);  // the same half of the next row: two stages ahead
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int ks = 0; ks < kt / 8; ++ks) {
                    const uint64_t dAh = umma_desc(s32(sAh) + ks * 2 * LBO_A, LBO_A), dAl = umma_desc(s32(sAl) + ks * 2 * LBO_A, LBO_A);
                    const uint64_t dBh = umma_desc(s32(sBh) + ks * 2 * LBO_B, LBO_B), dBl = umma_desc(s32(sBl) + ks * 2 * LBO_B, LBO_B);
                    umma_tf32(tmem_d, dAh, dBh, idesc, (it | (uint32_t)ks) ? 1u : 0u);
                    umma_tf32(tmem_d, dAh, dBl, idesc, 1u);
                    umma_tf32(tmem_d, dAl, dBh, idesc, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bars[buf]))
                             : "memory");
            }
        }
    }
    mbar_wait_or_trap(&bars[(it - 1u) & 1u], ((it - 1u) >> 1) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
        const int yp = 2 * warp + (lane >> 4), xp = lane & 15;
        const uint32_t trow = tmem_d + ((uint32_t)(warp * 32) << 16);
        const SrcView dv = src_view(SrcMap{sm.dst_full, 0}, b, C, H, W);
        float* o = gI + dv.base + (y0 + yp) * dv.rs + (x0 + xp) * dv.xs;
        for (int c0 = 0; c0 < C; c0 += 16) {
            uint32_t v[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(trow + c0)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) __stcs(o + (size_t)(c0 + j) * dv.cs, __uint_as_float(v[j]) * g.scale);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(ncols) : "memory");
}

// direction on the lowest bit of blockIdx.x: the two CTAs that read the same gradient planes are neighbours
__global__ void __launch_bounds__(BTHREADS, 1) corr_bwd_umma_kernel(const float* __restrict__ f1,
                                                                   const float* __restrict__ f2,
                                                                   const float* __restrict__ gO, float* __restrict__ gI1,
                                                                   float* __restrict__ gI2, ofr::CorrGeom g, int tiles_x,
                                                                   int tiles_y, int ndir, int first, SrcMap sm) {
    const int dir = (int)(blockIdx.x % ndir) + first;
    if (dir == 0)
        corr_bwd_umma_body<0>(f2, gO, gI1, g, tiles_x, tiles_y, blockIdx.x / ndir, sm);
    else
        corr_bwd_umma_body<1>(f1, gO, gI2, g, tiles_x, tiles_y, blockIdx.x / ndir, sm);
}

}  // namespace

namespace ofr {

// Class-image correlation (k = 1, s1 = s2 = 1, pad = md = 10) on the tensor cores.  OFR_ENOTSUP (nothing launched) when
// the geometry does not tile (the caller then uses the FMA kernels).  No process-global switch: the choice is a pure
// function of the geometry.
// host-side predicate: will corr_fwd_umma / corr_bwd_umma take this class geometry?  (ofr_corr_workspace uses it: the
// tensor-core kernels address the operator's tensors in place and need no scratch)
bool corr_umma_applicable(const CorrGeom& g, bool backward) {
    if (!(g.k == 1 && g.s1 == 1 && g.s2 == 1 && g.md == UR && g.pad == UR)) return false;
    if ((g.W % UTW) || (g.H % UTH) || (g.C % 16) || g.C < 16) return false;
    if (backward && g.C > 256) return false;
    if ((long long)g.C * g.H * g.W * 4 >= (1ll << 30) || (long long)UD * UD * g.H * g.W * 4 >= (1ll << 30)) return false;
    return true;
}

int corr_fwd_umma(const float* f1, const float* f2, float* out, const CorrGeom& g, cudaStream_t st, bool full_src,
                  bool full_dst) {
    if (!corr_umma_applicable(g, false)) return OFR_ENOTSUP;
    if (!(g.k == 1 && g.s1 == 1 && g.s2 == 1 && g.md == UR && g.pad == UR)) return OFR_ENOTSUP;
    if ((g.W % UTW) || (g.H % UTH) || (g.C % UKT) || g.C < UKT) return OFR_ENOTSUP;
    if ((long long)g.C * g.H * g.W >= (1ll << 30)) return OFR_ENOTSUP;
    const int tiles_x = g.W / UTW, tiles_y = g.H / UTH;
    const long long nblk = (long long)tiles_x * tiles_y * g.B;
    if (nblk >= (1ll << 31)) return OFR_ENOTSUP;
    OFR_CUDA(cudaFuncSetAttribute(corr_fwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)USMEM));
    if ((long long)g.C * g.H * g.W * 4 >= (1ll << 30)) return OFR_ENOTSUP;
    corr_fwd_umma_kernel<<<(unsigned)nblk, UTHREADS, USMEM, st>>>(f1, f2, out, g, tiles_x, tiles_y, SrcMap{full_src ? 1 : 0, full_dst ? 1 : 0});
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

int corr_bwd_umma(const float* f1, const float* f2, const float* gO, float* gI1, float* gI2, const CorrGeom& g,
                  cudaStream_t st, bool full_src, bool full_dst) {
    if (!corr_umma_applicable(g, true)) return OFR_ENOTSUP;
    if (!(g.k == 1 && g.s1 == 1 && g.s2 == 1 && g.md == UR && g.pad == UR)) return OFR_ENOTSUP;
    if ((g.W % UTW) || (g.H % UTH) || (g.C % 16) || g.C < 16 || g.C > 256) return OFR_ENOTSUP;
    if ((long long)g.C * g.H * g.W * 4 >= (1ll << 30) || (long long)UD * UD * g.H * g.W * 4 >= (1ll << 30)) return OFR_ENOTSUP;
    const int ndir = (gI1 ? 1 : 0) + (gI2 ? 1 : 0);
    if (ndir == 0) return OFR_OK;
    const int tiles_x = g.W / UTW, tiles_y = g.H / UTH;
    const long long nblk = (long long)tiles_x * tiles_y * g.B * ndir;
    if (nblk >= (1ll << 31)) return OFR_ENOTSUP;
    OFR_CUDA(cudaFuncSetAttribute(corr_bwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BSMEM));
    corr_bwd_umma_kernel<<<(unsigned)nblk, BTHREADS, BSMEM, st>>>(f1, f2, gO, gI1, gI2, g, tiles_x, tiles_y, ndir,
                                                                  gI1 ? 0 : 1, SrcMap{full_src ? 1 : 0, full_dst ? 1 : 0});
    OFR_LAUNCH_CHECK();
    return OFR_OK;
}

}  // namespace ofr
