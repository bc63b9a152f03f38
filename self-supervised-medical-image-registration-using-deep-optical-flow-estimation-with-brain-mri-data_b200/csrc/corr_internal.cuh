// Shared between correlation.cu (generic + register-tiled kernels, ABI entry points) and correlation_tma.cu.
#pragma once
#include "ofr_common.cuh"

namespace ofr {

struct CorrGeom {
    int B, C, H, W;
    int pad, k, md, s1, s2;
    int r, D, Hout, Wout;
    float scale;  // 1 / (k*k*C)
    float slope;  // fused LeakyReLU on the forward output (PWCNet.py:201, FlowNetC.py:89); 1 = none
};

// output element: scaled sum, then the optional LeakyReLU
__device__ __forceinline__ float corr_out(float acc, const CorrGeom& g) {
    const float v = acc * g.scale;
    return v > 0.f ? v : v * g.slope;
}

// TMA-staged kernels for Correlation(4,1,4,1,1).  Return OFR_ENOTSUP (nothing launched) when the geometry or the
// pointer alignment does not fit; the caller then falls back to the generic kernels.
int corr_fwd_tma(const float* f1, const float* f2, float* out, const CorrGeom& g, cudaStream_t st);
int corr_bwd_tma(const float* f1, const float* f2, const float* gO, float* gI1, float* gI2, const CorrGeom& g,
                 cudaStream_t st);

// tcgen05 / TMEM forward for the class-image problem of the FlowNetC window (pad = md = 10 after the polyphase split):
// correlation_umma.cu.  OFR_ENOTSUP when the geometry does not tile.
// `g` is the class geometry (B*4 class images of H/2 x W/2).  full_src: the inputs (f1, f2, gO) are the operator's own
// [B][C][H][W] tensors and are read in place with stride 2; the outputs are always written in the class layout.
bool corr_umma_applicable(const CorrGeom& class_geom, bool backward);
int corr_fwd_umma(const float* f1, const float* f2, float* out, const CorrGeom& g, cudaStream_t st, bool full_src = false,
                  bool full_dst = false);  // full_dst: `out` is the operator's own [B][D*D][H][W] tensor (no merge pass)
int corr_bwd_umma(const float* f1, const float* f2, const float* gO, float* gI1, float* gI2, const CorrGeom& g,
                  cudaStream_t st, bool full_src = false, bool full_dst = false);

}  // namespace ofr
