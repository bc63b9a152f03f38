// FP32 FMA peak microbenchmark: scalar FFMA vs packed fma.rn.f32x2 (FFMA2), register operands only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_peak fma_peak.cu && ./fma_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float x, float y) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    float b0 = x, b1 = y, b2 = x + 1.f, b3 = y + 1.f;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    a[i] = fmaf(a[i], b0, b1);
                    a[i + 1] = fmaf(a[i + 1], b1, b2);
                    a[i + 2] = fmaf(a[i + 2], b2, b3);
                    a[i + 3] = fmaf(a[i + 3], b3, b0);
                }
            }
        } else if (MODE == 1) {  // accumulate form d = a*b + d with distinct a, b registers (like a register-tiled GEMM)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    a[i] = fmaf(b0, b1, a[i]);
                    a[i + 1] = fmaf(b0, b2, a[i + 1]);
                    a[i + 2] = fmaf(b0, b3, a[i + 2]);
                    a[i + 3] = fmaf(b1, b2, a[i + 3]);
                }
            }
        } else {  // packed: 2 FMAs per instruction
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    unsigned long long d, p, q;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a[i]), "f"(a[i + 1]));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(b0), "f"(b1));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(b2), "f"(b3));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(p), "l"(q));
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(d));
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* out, int ctas) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<ctas, 256>>>(out, 100, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<ctas, 256>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)ctas * 256 * iters * 64.0;
    printf("%-28s ctas/SM %d  %.3f ms  %.2f TFLOP/s\n", name, ctas / 148, ms, 2 * fma / ms / 1e9);
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    for (int c : {2, 4, 8}) {
        run<0>("ffma a=a*b+c", out, 148 * c);
        run<1>("ffma d=a*b+d", out, 148 * c);
        run<2>("fma.f32x2 d=a*b+d", out, 148 * c);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
