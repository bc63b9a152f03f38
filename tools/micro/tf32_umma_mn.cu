// MN-major operand layouts for tcgen05 kind::tf32 (operands as TMA delivers NCHW pixels: the pixel index contiguous).
// Derived from tf32_umma.cu; checks which of the descriptor's two byte offsets is the K-group stride in the no-swizzle
// MN-major layout  element (mn, k) at (mn/4) * 128 + (k/8) * (R/4) * 128 + (k%8) * 16 + (mn%4) * 4  bytes.
// Original header:
// tcgen05 kind::tf32 microbenchmark for the planned tensor-core recast of the FlowNetC correlation:
//   (1) accuracy of 1xTF32 and of the 3xTF32 split (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM) against float64,
//   (2) MMA throughput per SM for M = 128, N in {32, 64, 128, 256}, K = 32 per stage, operands read from shared memory.
// One CTA per SM, 128 threads; operands are written by the threads in the canonical K-major no-swizzle layout
// (8 rows x 16 bytes core matrices), one elected thread issues the MMAs, tcgen05.commit -> mbarrier, tcgen05.ld epilogue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tf32_umma tf32_umma.cu && ./tf32_umma
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

constexpr int M = 128;
constexpr int KT = 32;  // K per stage (4 MMA K-steps of 8)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: element (row r, 16-byte chunk j along K) at j * (R * 16) + r * 16 bytes
// -> core matrix (8 rows x 16 B) contiguous, SBO (next 8 rows) = 128 B, LBO (next chunk along K) = R * 16 B
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    return d;                // layout_type 0 = no swizzle, base_offset 0
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

__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    while (clock64() - t0 < 4000000000ll) {  // ~2 s: a wrong descriptor must not hang the GPU box
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}

// SPLIT = 1: one MMA per K-step on the raw operands; SPLIT = 3: hi*hi + hi*lo + lo*hi
template <int N, int SPLIT, int VARIANT>
__global__ void __launch_bounds__(128) umma_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                   float* __restrict__ D, int iters, int* __restrict__ fail) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sAh = reinterpret_cast<float*>(smem);
    float* sAl = sAh + M * KT;
    float* sBh = sAl + M * KT;
    float* sBl = sBh + N * KT;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    // operands -> canonical layout, split into tf32 hi (round to nearest) and lo = x - hi (exact)
    auto put = [&](float* hi, float* lo, const float* src, int R) {
        for (int e = tid; e < R * KT; e += 128) {
            const int r = e / KT, k = e - r * KT;
            const float x = src[(size_t)blockIdx.x * 0 + e];
            float h;
            if (SPLIT == 1) {
                h = x;
            } else {
                uint32_t hb;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
                h = __uint_as_float(hb);
            }
            const int off = VARIANT == 0 ? (k >> 2) * (R * 4) + r * 4 + (k & 3)                       // K-major
                                         : (r >> 2) * 32 + (k >> 3) * (R * 8) + (k & 7) * 4 + (r & 3);  // MN-major (floats)
            hi[off] = h;
            lo[off] = x - h;
        }
    };
    put(sAh, sAl, A, M);
    put(sBh, sBl, B, N);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // make the generic-proxy writes of the operands visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    constexpr uint32_t NCOLS = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;  // power of two >= 32
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(NCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    // instruction descriptor: D = F32 (bit 4), A = B = TF32 (2 at bits 7 and 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24) |
                               (VARIANT ? (1u << 15) | (1u << 16) : 0u);  // a_major / b_major = MN
    if (tid == 0) {
        // K-major: LBO = next 16-byte chunk along K, SBO = next 8 rows.  MN-major: K-group (8 k) stride and MN-group (4 mn)
        // stride; VARIANT 1 puts the K-group stride in LBO, VARIANT 2 in SBO
        const uint32_t kgA = VARIANT ? M * 32 : M * 16, kgB = VARIANT ? N * 32 : N * 16;   // bytes per MMA K-step / K chunk
        const uint32_t lboA = VARIANT == 2 ? 128 : kgA, lboB = VARIANT == 2 ? 128 : kgB;
        const uint32_t sboA = VARIANT == 2 ? kgA : 128, sboB = VARIANT == 2 ? kgB : 128;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int ks = 0; ks < KT / 8; ++ks) {
                const uint32_t offA = VARIANT ? ks * kgA : ks * 2 * kgA, offB = VARIANT ? ks * kgB : ks * 2 * kgB;
                const uint64_t dAh = make_desc(smem_u32(sAh) + offA, lboA, sboA), dAl = make_desc(smem_u32(sAl) + offA, lboA, sboA);
                const uint64_t dBh = make_desc(smem_u32(sBh) + offB, lboB, sboB), dBl = make_desc(smem_u32(sBl) + offB, lboB, sboB);
                umma_tf32(tmem_d, dAh, dBh, idesc, (it | ks) ? 1u : 0u);
                if (SPLIT == 3) {
                    umma_tf32(tmem_d, dAh, dBl, idesc, 1u);
                    umma_tf32(tmem_d, dAl, dBh, idesc, 1u);
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                     : "memory");
    }
    const bool ok = mbar_wait_bounded(&bar, 0);
    if (!ok && tid == 0) atomicAdd(fail, 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w owns TMEM lanes 32w .. 32w+31 (rows of D); 32 columns per load
    if (ok && D != nullptr && blockIdx.x == 0) {
        for (int c0 = 0; c0 < NCOLS; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c0 + j < N) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(NCOLS) : "memory");
}


// ---------------------------------------------------------------------------------------------------------------------
// Layout probe: A = all ones (layout independent), B shared memory word i holds a code of i; ONE MMA (K = 8) with the given
// descriptor fields; D[0][n] = sum of the codes of the 8 words the tensor core reads for column n.  Two codes (i & 1023,
// i >> 10) identify the words.  (Then the roles are exchanged to probe A.)
template <int N>
__global__ void __launch_bounds__(128) probe_kernel(float* __restrict__ D, uint32_t lbo, uint32_t sbo, uint32_t majorA,
                                                    uint32_t majorB, int code, int probeA, int* __restrict__ fail) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sA = reinterpret_cast<float*>(smem);
    float* sB = sA + M * KT;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * KT; i += 128) sA[i] = probeA ? (float)(code ? (i >> 10) : (i & 1023)) : 1.f;
    for (int i = tid; i < N * KT; i += 128) sB[i] = probeA ? 1.f : (float)(code ? (i >> 10) : (i & 1023));
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24) |
                           (majorA << 15) | (majorB << 16);
    if (tid == 0) {
        umma_tf32(tmem_d, make_desc(smem_u32(sA), lbo, sbo), make_desc(smem_u32(sB), lbo, sbo), idesc, 0u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    const bool ok = mbar_wait_bounded(&bar, 0);
    if (!ok && tid == 0) atomicAdd(fail, 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok) {
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(256u) : "memory");
}

void probe(float* dD, int* dFail, uint32_t lbo, uint32_t sbo, uint32_t majA, uint32_t majB, int probeA) {
    constexpr int N = 256;
    const size_t smem = sizeof(float) * (M * KT + N * KT);
    CK(cudaFuncSetAttribute(probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<float> lo((size_t)M * N), hi((size_t)M * N);
    for (int code = 0; code < 2; ++code) {
        probe_kernel<N><<<1, 128, smem>>>(dD, lbo, sbo, majA, majB, code, probeA, dFail);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy((code ? hi : lo).data(), dD, sizeof(float) * M * N, cudaMemcpyDeviceToHost));
    }
    printf("probe %s: LBO %u SBO %u majorA %u majorB %u -> (sum of word index & 1023, sum of word index >> 10) for %s index 0..11, 16, 17, 64, 255:\n   ",
           probeA ? "A" : "B", lbo, sbo, majA, majB, probeA ? "row" : "column");
    const int idx[16] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 16, 17, 64, probeA ? 127 : 255};
    for (int q = 0; q < 16; ++q) {
        const size_t at = probeA ? (size_t)idx[q] * N : (size_t)idx[q];
        printf(" %d:(%.0f,%.0f)", idx[q], lo[at], hi[at]);
    }
    printf("\n");
}

template <int N, int SPLIT, int VARIANT>
void run(const float* dA, const float* dB, float* dD, int* dFail, const std::vector<float>& hA, const std::vector<float>& hB,
         int sms) {
    auto kern = umma_kernel<N, SPLIT, VARIANT>;
    const size_t smem = sizeof(float) * 2 * (M * KT + N * KT);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // accuracy: one pass of K = 32
    CK(cudaMemset(dFail, 0, sizeof(int)));
    kern<<<1, 128, smem>>>(dA, dB, dD, 1, dFail);
    CK(cudaDeviceSynchronize());
    int fail = 0;
    CK(cudaMemcpy(&fail, dFail, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<float> hD((size_t)M * N);
    CK(cudaMemcpy(hD.data(), dD, sizeof(float) * M * N, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < KT; ++k) ref += (double)hA[m * KT + k] * (double)hB[n * KT + k];
            maxerr = fmax(maxerr, fabs(ref - (double)hD[(size_t)m * N + n]));
            maxref = fmax(maxref, fabs(ref));
        }
    // throughput: one CTA per SM, many stages on resident operands
    const int iters = 4000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    kern<<<sms, 128, smem>>>(dA, dB, nullptr, iters, dFail);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    kern<<<sms, 128, smem>>>(dA, dB, nullptr, iters, dFail);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&fail, dFail, sizeof(int), cudaMemcpyDeviceToHost));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double mmas = (double)sms * iters * (KT / 8) * SPLIT;
    const double flops = mmas * 2.0 * M * N * 8;
    printf("variant %d  N=%3d  %dxTF32  max|err| %.3e (max|ref| %.2f, rel %.2e)  %8.3f ms  %8.1f TFLOP/s issued  %8.1f TFLOP/s useful%s\n", VARIANT, N,
           SPLIT, maxerr, maxref, maxerr / maxref, ms, flops / ms * 1e-9, flops / SPLIT / ms * 1e-9,
           fail ? "  [mbarrier wait timed out]" : "");
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    std::vector<float> hA(M * KT), hB(256 * KT);
    srand(6);
    auto rnd = [] { return (float)((rand() / (double)RAND_MAX) * 2.0 - 1.0); };
    for (auto& x : hA) x = rnd();
    for (auto& x : hB) x = rnd();
    float *dA, *dB, *dD;
    int* dFail;
    CK(cudaMalloc(&dA, sizeof(float) * hA.size()));
    CK(cudaMalloc(&dB, sizeof(float) * hB.size()));
    CK(cudaMalloc(&dD, sizeof(float) * M * 256));
    CK(cudaMalloc(&dFail, sizeof(int)));
    CK(cudaMemcpy(dA, hA.data(), sizeof(float) * hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), sizeof(float) * hB.size(), cudaMemcpyHostToDevice));
    printf("tcgen05.mma kind::tf32, M=128, K=%d per stage, %d SMs, operands in shared memory (K-major, no swizzle)\n", KT, sms);
    // layout probes: K-major reference first (known good), then MN-major with both assignments of the two offsets
    probe(dD, dFail, 256 * 16, 128, 0, 0, 0);
    probe(dD, dFail, 256 * 32, 128, 0, 1, 0);
    probe(dD, dFail, 128, 256 * 32, 0, 1, 0);
    probe(dD, dFail, 128 * 16, 128, 0, 0, 1);
    probe(dD, dFail, 128 * 32, 128, 1, 0, 1);
    probe(dD, dFail, 128, 128 * 32, 1, 0, 1);
    for (int v = 0; v < 1; ++v) {
        run<256, 3, 0>(dA, dB, dD, dFail, hA, hB, sms);
        run<256, 3, 1>(dA, dB, dD, dFail, hA, hB, sms);
        run<256, 3, 2>(dA, dB, dD, dFail, hA, hB, sms);
        run<240, 3, 1>(dA, dB, dD, dFail, hA, hB, sms);
        run<240, 3, 2>(dA, dB, dD, dFail, hA, hB, sms);
    }
    return 0;
}
