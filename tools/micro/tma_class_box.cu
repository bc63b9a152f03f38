// Micro-test for the TMA operand path of the tcgen05 correlation: one 5-D tiled tensor map over a
// float32 [G][H][W] tensor (G = batch x channels merged) that delivers, for one polyphase class (py, px) of the image,
// a box [cblk][rows][xb][c8][xi4] -- 4 class pixels (16 B) innermost, then 8 channels (= one 128-byte MN-major core matrix),
// then x blocks, class rows, channel blocks -- straight into the canonical no-swizzle MN-major UMMA layout.  Checks
// elementStrides = 2 on x and y (class pixel (y, x) = input pixel (2y + py, 2x + px)), the aliasing x = xi + 8 * xb, and
// zero fill outside the image.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_class_box tma_class_box.cu -lcuda && ./tma_class_box
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int G = 32, H = 24, W = 32;          // input tensor (class images 12 x 16)
constexpr int NXB = 5, NROW = 3, NCB = 2;      // box: 5 x-blocks (20 class px), 3 class rows, 2 channel blocks (16 channels)
constexpr int BOX_FLOATS = NCB * NROW * NXB * 8 * 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void box_kernel(const __grid_constant__ CUtensorMap map, float* out, int c0, int c1, int c2, int c3, int c4) {
    __shared__ __align__(1024) float buf[BOX_FLOATS];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < BOX_FLOATS; i += blockDim.x) buf[i] = -777.f;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BOX_FLOATS * 4) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
            ::"r"(smem_u32(buf)), "l"(&map), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
    }
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        if (ok) break;
        if (clock64() - t0 > 2000000000ll) { if (threadIdx.x == 0) out[0] = -999.f; return; }
    }
    for (int i = threadIdx.x; i < BOX_FLOATS; i += blockDim.x) out[i] = buf[i];
}

int main(int argc, char** argv) {
    const int cfg = argc > 1 ? atoi(argv[1]) : 0;   // 0: full (strides on x and y); 1: no element strides; 2: y only; 3: x only
    std::vector<float> h((size_t)G * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;   // value = linear index: (g, y, x) recoverable
    float *d, *dout;
    CK(cudaMalloc(&d, h.size() * 4));
    CK(cudaMalloc(&dout, BOX_FLOATS * 4));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
    // dims: d0 = x inside a block of 8 input pixels, d1 = channel inside a block of 8, d2 = x block, d3 = y, d4 = channel block
    const cuuint64_t dims[5] = {8, 8, W / 8, H, G / 8};
    const cuuint64_t strides[4] = {(cuuint64_t)H * W * 4, 8 * 4, (cuuint64_t)W * 4, (cuuint64_t)8 * H * W * 4};
    const cuuint32_t sx = (cfg == 0 || cfg == 3) ? 2 : 1, sy = (cfg == 0 || cfg == 2) ? 2 : 1;
    const cuuint32_t box[5] = {4 * sx, 8, NXB, sy * NROW, NCB};
    const cuuint32_t estr[5] = {sx, 1, 1, sy, 1};
    printf("config %d: element strides x %u y %u\n", cfg, sx, sy);
    CUtensorMap map;
    CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("cuTensorMapEncodeTiled rc = %d\n", (int)rc);
    if (rc != CUDA_SUCCESS) return 1;
    int nbad_total = 0;
    // cases: (py, px, first x block, first class row, first channel block)
    const int cases[4][5] = {{0, 0, 0, 0, 0}, {1, 1, 1, 2, 1}, {1, 0, -1, -1, 0}, {0, 1, 2, 10, 2}};
    for (auto& cs : cases) {
        const int py = cs[0], px = cs[1], xb0 = cs[2], r0 = cs[3], cb0 = cs[4];
        box_kernel<<<1, 128>>>(map, dout, sx == 2 ? px : 0, 0, xb0, sy == 2 ? 2 * r0 + py : r0, cb0);
        CK(cudaDeviceSynchronize());
        std::vector<float> o(BOX_FLOATS);
        CK(cudaMemcpy(o.data(), dout, BOX_FLOATS * 4, cudaMemcpyDeviceToHost));
        int nbad = 0;
        for (int cb = 0; cb < NCB; ++cb)
            for (int r = 0; r < NROW; ++r)
                for (int xb = 0; xb < NXB; ++xb)
                    for (int c8 = 0; c8 < 8; ++c8)
                        for (int xi = 0; xi < 4; ++xi) {
                            const int g = (cb0 + cb) * 8 + c8, y = sy == 2 ? 2 * (r0 + r) + py : r0 + r, xblk = xb0 + xb,
                                      x = xblk * 8 + (sx == 2 ? 2 * xi + px : xi);
                            const bool in = g >= 0 && g < G && y >= 0 && y < H && xblk >= 0 && xblk < W / 8;
                            const float want = in ? (float)(((size_t)g * H + y) * W + x) : 0.f;
                            const float got = o[(((cb * NROW + r) * NXB + xb) * 8 + c8) * 4 + xi];
                            if (got != want) {
                                if (nbad < 6) printf("  mismatch cb %d r %d xb %d c8 %d xi %d: got %.0f want %.0f\n", cb, r, xb, c8, xi, got, want);
                                ++nbad;
                            }
                        }
        printf("case py=%d px=%d xb0=%d r0=%d cb0=%d: %s (%d mismatches of %d)\n", py, px, xb0, r0, cb0, nbad ? "FAIL" : "PASS", nbad, BOX_FLOATS);
        nbad_total += nbad;
    }
    printf(nbad_total ? "TMA class box: FAIL\n" : "TMA class box: PASS\n");
    return 0;
}
