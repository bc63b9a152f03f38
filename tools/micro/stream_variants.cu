// Which ingredients of a small (50 MB) streaming kernel cost bandwidth on B200?  out[i] = sqrt(x[i]^2) over 24 x 512 x 512
// floats (ChannelNorm forward, C = 1: 25 MB read + 25 MB written), variants of block size, vectors per thread, cache hints,
// square root and index arithmetic; rotating buffer sets larger than the L2, CUDA events around 24 back-to-back launches.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_variants stream_variants.cu && ./stream_variants
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template <int HINT> __device__ __forceinline__ float4 ld4(const float* p) {
    if (HINT == 1) return __ldcs(reinterpret_cast<const float4*>(p));
    if (HINT == 2) return __ldg(reinterpret_cast<const float4*>(p));
    return *reinterpret_cast<const float4*>(p);
}
template <int HINT> __device__ __forceinline__ void st4(float* p, float4 v) {
    if (HINT == 1) __stcs(reinterpret_cast<float4*>(p), v);
    else *reinterpret_cast<float4*>(p) = v;
}
template <int SQ> __device__ __forceinline__ float sq(float a) {
    if (SQ == 0) return sqrtf(a);
    if (SQ == 1) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
    return fabsf(a);  // (no square root at all: upper bound)
}

// THREADS per block, VPT vectors per thread (block-strided), HINT 0 plain / 1 .cs / 2 ldg+plain store, SQ, DIV: 64-bit
// division per vector like the generic (B, C, HW) indexing, PERSIST: grid-stride loop over a grid of SMs * k blocks
template <int THREADS, int VPT, int HINT, int SQ, bool DIV, bool PERSIST>
__global__ void __launch_bounds__(THREADS) k(const float* __restrict__ x, float* __restrict__ out, long long nvec, long long hwv) {
    for (long long base = (long long)blockIdx.x * VPT * THREADS; base < nvec; base += PERSIST ? (long long)gridDim.x * VPT * THREADS : nvec) {
        float4 v[VPT];
        long long idx[VPT];
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            long long i = base + (long long)j * THREADS + threadIdx.x;
            if (DIV) { const long long b = i / hwv; i = b * hwv + (i - b * hwv); }
            idx[j] = i;
        }
#pragma unroll
        for (int j = 0; j < VPT; ++j) v[j] = idx[j] < nvec ? ld4<HINT>(x + idx[j] * 4) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            float4 s = make_float4(sq<SQ>(v[j].x * v[j].x), sq<SQ>(v[j].y * v[j].y), sq<SQ>(v[j].z * v[j].z), sq<SQ>(v[j].w * v[j].w));
            if (idx[j] < nvec) st4<HINT>(out + idx[j] * 4, s);
        }
    }
}

// the production kernel's structure (channelnorm.cu): predicated tail, pointer arrays, runtime channel loop
template <int VPT>
__global__ void __launch_bounds__(256) prod(const float* __restrict__ x, float* __restrict__ out, int C, long long hw, long long total_vec) {
    const long long hwv = hw / 4;
    const long long i0 = (long long)blockIdx.x * VPT * 256 + threadIdx.x;
    const float* xp[VPT];
    float* op[VPT];
    bool ok[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        const long long i = i0 + (long long)j * 256;
        ok[j] = i < total_vec;
        const long long ic = ok[j] ? i : 0;
        const long long b = ic / hwv, p = (ic - b * hwv) * 4;
        xp[j] = x + b * C * hw + p;
        op[j] = out + b * hw + p;
    }
    float4 s[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) s[j] = make_float4(0, 0, 0, 0);
    for (int c = 0; c < C; ++c) {
        float4 v[VPT];
#pragma unroll
        for (int j = 0; j < VPT; ++j) v[j] = ok[j] ? __ldcs(reinterpret_cast<const float4*>(xp[j])) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            s[j] = make_float4(fmaf(v[j].x, v[j].x, s[j].x), fmaf(v[j].y, v[j].y, s[j].y), fmaf(v[j].z, v[j].z, s[j].z), fmaf(v[j].w, v[j].w, s[j].w));
            xp[j] += hw;
        }
    }
#pragma unroll
    for (int j = 0; j < VPT; ++j)
        if (ok[j]) __stcs(reinterpret_cast<float4*>(op[j]), make_float4(sqrtf(s[j].x), sqrtf(s[j].y), sqrtf(s[j].z), sqrtf(s[j].w)));
}
// bisect of the production structure: FLAT (no division), NOPRED (no tail predicate), CFIX (C = 1 at compile time), PLAIN
template <int VPT, bool FLAT, bool NOPRED, bool CFIX, bool PLAIN>
__global__ void __launch_bounds__(256) prod2(const float* __restrict__ x, float* __restrict__ out, int Crt, long long hw, long long total_vec) {
    const int C = CFIX ? 1 : Crt;
    const long long hwv = hw / 4;
    const long long i0 = (long long)blockIdx.x * VPT * 256 + threadIdx.x;
    const float* xp[VPT];
    float* op[VPT];
    bool ok[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        const long long i = i0 + (long long)j * 256;
        ok[j] = NOPRED || i < total_vec;
        const long long ic = ok[j] ? i : 0;
        if (FLAT) {
            xp[j] = x + ic * 4;
            op[j] = out + ic * 4;
        } else {
            const long long b = ic / hwv, p = (ic - b * hwv) * 4;
            xp[j] = x + b * C * hw + p;
            op[j] = out + b * hw + p;
        }
    }
    float4 s[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) s[j] = make_float4(0, 0, 0, 0);
    for (int c = 0; c < C; ++c) {
        float4 v[VPT];
#pragma unroll
        for (int j = 0; j < VPT; ++j) v[j] = ok[j] ? (PLAIN ? *reinterpret_cast<const float4*>(xp[j]) : __ldcs(reinterpret_cast<const float4*>(xp[j]))) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            s[j] = make_float4(fmaf(v[j].x, v[j].x, s[j].x), fmaf(v[j].y, v[j].y, s[j].y), fmaf(v[j].z, v[j].z, s[j].z), fmaf(v[j].w, v[j].w, s[j].w));
            xp[j] += hw;
        }
    }
#pragma unroll
    for (int j = 0; j < VPT; ++j)
        if (ok[j]) {
            const float4 r = make_float4(sqrtf(s[j].x), sqrtf(s[j].y), sqrtf(s[j].z), sqrtf(s[j].w));
            if (PLAIN) *reinterpret_cast<float4*>(op[j]) = r; else __stcs(reinterpret_cast<float4*>(op[j]), r);
        }
}
// ATen-like elementwise: 128 threads, 2 vectors per thread, plain accesses
__global__ void __launch_bounds__(128) add1(const float* __restrict__ x, float* __restrict__ out, long long nvec) {
    const long long base = (long long)blockIdx.x * 256 + threadIdx.x;
    float4 v[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) if (base + j * 128 < nvec) v[j] = *reinterpret_cast<const float4*>(x + (base + j * 128) * 4);
#pragma unroll
    for (int j = 0; j < 2; ++j) if (base + j * 128 < nvec) *reinterpret_cast<float4*>(out + (base + j * 128) * 4) = make_float4(v[j].x + 1.f, v[j].y + 1.f, v[j].z + 1.f, v[j].w + 1.f);
}

constexpr int NSETS = 8;
float *X[NSETS], *Y[NSETS];
const long long N = 24ll * 512 * 512, NV = N / 4;

template <int THREADS, int VPT, int HINT, int SQ, bool DIV, bool PERSIST>
void run(const char* name, int sms) {
    const long long per = (long long)THREADS * VPT;
    unsigned grid = PERSIST ? (unsigned)(sms * (2048 / THREADS)) : (unsigned)((NV + per - 1) / per);
    auto kern = k<THREADS, VPT, HINT, SQ, DIV, PERSIST>;
    for (int i = 0; i < NSETS; ++i) kern<<<grid, THREADS>>>(X[i], Y[i], NV, 512 * 512 / 4);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 24; ++i) kern<<<grid, THREADS>>>(X[i % NSETS], Y[i % NSETS], NV, 512 * 512 / 4);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double us = best * 1e3 / 24;
    printf("%-58s grid %6u  %7.2f us  %7.1f GB/s\n", name, grid, us, 2.0 * N * 4 / us * 1e-3);
}

template <class F>
void run_graph(const char* name, F launch) {
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    for (int i = 0; i < NSETS; ++i) launch(i, st);
    CK(cudaStreamSynchronize(st));
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
    for (int i = 0; i < 24; ++i) launch(i % NSETS, st);
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0, st)); CK(cudaGraphLaunch(ge, st)); CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    const double us = best * 1e3 / 24;
    printf("%-58s [graph]      %7.2f us  %7.1f GB/s\n", name, us, 2.0 * N * 4 / us * 1e-3);
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    for (int i = 0; i < NSETS; ++i) { CK(cudaMalloc(&X[i], N * 4)); CK(cudaMalloc(&Y[i], N * 4)); CK(cudaMemset(X[i], 0x3f, N * 4)); }
    //   THREADS VPT HINT SQ DIV PERSIST
    run<256, 4, 1, 0, true, false>("A  256thr vpt4 .cs  ieee-sqrt div        (current)", sms);
    run<256, 4, 1, 0, false, false>("B  256thr vpt4 .cs  ieee-sqrt flat", sms);
    run<256, 4, 0, 0, false, false>("C  256thr vpt4 plain ieee-sqrt flat", sms);
    run<256, 4, 2, 0, false, false>("C2 256thr vpt4 ldg   ieee-sqrt flat", sms);
    run<256, 4, 0, 1, false, false>("D  256thr vpt4 plain approx-sqrt flat", sms);
    run<256, 4, 0, 2, false, false>("D2 256thr vpt4 plain fabs flat", sms);
    run<128, 4, 0, 0, false, false>("E  128thr vpt4 plain ieee-sqrt flat", sms);
    run<128, 1, 0, 0, false, false>("F  128thr vpt1 plain ieee-sqrt flat", sms);
    run<256, 1, 0, 0, false, false>("F2 256thr vpt1 plain ieee-sqrt flat", sms);
    run<256, 2, 0, 0, false, false>("F3 256thr vpt2 plain ieee-sqrt flat", sms);
    run<256, 8, 0, 0, false, false>("G  256thr vpt8 plain ieee-sqrt flat", sms);
    run<256, 4, 0, 0, false, true>("H  256thr vpt4 plain ieee-sqrt flat persistent", sms);
    run<512, 4, 0, 0, false, false>("I  512thr vpt4 plain ieee-sqrt flat", sms);
    run<256, 4, 1, 1, false, false>("J  256thr vpt4 .cs  approx-sqrt flat", sms);
    run_graph("production structure vpt4 (C=1)", [](int i, cudaStream_t st) { prod<4><<<(unsigned)((NV + 1023) / 1024), 256, 0, st>>>(X[i], Y[i], 1, 512 * 512, NV); });
    run_graph("production structure vpt2 (C=1)", [](int i, cudaStream_t st) { prod<2><<<(unsigned)((NV + 511) / 512), 256, 0, st>>>(X[i], Y[i], 1, 512 * 512, NV); });
    run_graph("prod2 flat", [](int i, cudaStream_t st) { prod2<4, true, false, false, false><<<(unsigned)((NV + 1023) / 1024), 256, 0, st>>>(X[i], Y[i], 1, 512 * 512, NV); });
    run_graph("prod2 flat nopred", [](int i, cudaStream_t st) { prod2<4, true, true, false, false><<<(unsigned)((NV + 1023) / 1024), 256, 0, st>>>(X[i], Y[i], 1, 512 * 512, NV); });
    run_graph("prod2 flat cfix", [](int i, cudaStream_t st) { prod2<4, true, false, true, false><<<(unsigned)((NV + 1023) / 1024), 256, 0, st>>>(X[i], Y[i], 1, 512 * 512, NV); });
    run_graph("prod2 flat cfix nopred", [](int i, cudaStream_t st) { prod2<4, true, true, true, false><<<(unsigned)((NV + 1023) / 1024), 256, 0, st>>>(X[i], Y[i], 1, 512 * 512, NV); });
    run_graph("prod2 div cfix", [](int i, cudaStream_t st) { prod2<4, false, false, true, false><<<(unsigned)((NV + 1023) / 1024), 256, 0, st>>>(X[i], Y[i], 1, 512 * 512, NV); });
    run_graph("prod2 flat cfix plain", [](int i, cudaStream_t st) { prod2<4, true, false, true, true><<<(unsigned)((NV + 1023) / 1024), 256, 0, st>>>(X[i], Y[i], 1, 512 * 512, NV); });
    run_graph("prod2 flat plain", [](int i, cudaStream_t st) { prod2<4, true, false, false, true><<<(unsigned)((NV + 1023) / 1024), 256, 0, st>>>(X[i], Y[i], 1, 512 * 512, NV); });
    run_graph("ATen-like add1 128thr vpt2 plain", [](int i, cudaStream_t st) { add1<<<(unsigned)((NV + 255) / 256), 128, 0, st>>>(X[i], Y[i], NV); });
    run_graph("B 256thr vpt4 .cs ieee flat", [](int i, cudaStream_t st) { k<256, 4, 1, 0, false, false><<<(unsigned)((NV + 1023) / 1024), 256, 0, st>>>(X[i], Y[i], NV, 65536); });
    run_graph("F3 256thr vpt2 plain ieee flat", [](int i, cudaStream_t st) { k<256, 2, 0, 0, false, false><<<(unsigned)((NV + 511) / 512), 256, 0, st>>>(X[i], Y[i], NV, 65536); });
    run_graph("D2 256thr vpt4 plain fabs flat", [](int i, cudaStream_t st) { k<256, 4, 0, 2, false, false><<<(unsigned)((NV + 1023) / 1024), 256, 0, st>>>(X[i], Y[i], NV, 65536); });
    run_graph("128thr vpt2 plain fabs flat", [](int i, cudaStream_t st) { k<128, 2, 0, 2, false, false><<<(unsigned)((NV + 255) / 256), 128, 0, st>>>(X[i], Y[i], NV, 65536); });
    return 0;
}
