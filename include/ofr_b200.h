/*
 * ofr_b200.h -- C ABI of the B200 (sm_100a) optical-flow registration hot path.
 *
 * One shared library (libofr_b200.so) exports every entry point below with C
 * linkage.  Signatures carry only plain pointers, ints, doubles and an opaque
 * stream handle (a cudaStream_t / CUstream passed as void*): no torch types, no
 * C++ types.  This is the boundary a maintainer of the reference binds instead
 * of the pybind11 modules `correlation_cuda`, `resample2d_cuda`,
 * `channelnorm_cuda` of NVIDIA/flownet2-pytorch that the reference imports
 * (flownet2/networks/FlowNetC.py:8, flownet2/models.py:10-11,
 * PWC/models/PWCNet.py:13) and instead of the eager-PyTorch code of
 * models.py:256-268 (stn), PWC/models/PWCNet.py:143-179 (warp) and
 * loss.py:9-84 (OFEloss).  INTEGRATION.md shows the ctypes binding.
 *
 * Contract:
 *  - every tensor is float32, NCHW, contiguous, resident on the current device;
 *  - the caller owns all memory (inputs, outputs, gradients, workspaces); the
 *    library never allocates device memory, never frees, never keeps a pointer;
 *  - outputs are fully overwritten (no "pre-zeroed" assumption);
 *  - calls are asynchronous on `stream`, re-entrant, hold no mutable global
 *    state, never synchronise the host and are CUDA-graph capturable;
 *  - return value: 0 on success; OFR_E* (< 1000) for argument errors detected
 *    on the host before any launch; 1000 + cudaError_t for CUDA failures.
 *    Nothing throws across the ABI.  There is no CPU fallback.
 */
#ifndef OFR_B200_H
#define OFR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFR_ABI_VERSION 3

typedef void* ofr_stream_t; /* cudaStream_t */

enum {
    OFR_OK = 0,
    OFR_EINVAL = 1,     /* null pointer / non-positive dimension / bad enum        */
    OFR_ENOTSUP = 2,    /* legal but unimplemented parameter combination           */
    OFR_EWORKSPACE = 3, /* workspace missing, too small or misaligned (256 B)      */
    OFR_ECUDA = 1000    /* 1000 + cudaError_t                                      */
};

/* Version of this ABI (OFR_ABI_VERSION the library was built with). */
int ofr_abi_version(void);
/* Static, human readable text for a return code (never NULL). */
const char* ofr_error_string(int code);
/* Number of kernels this library has launched in the calling process since load
 * (monotonic, relaxed atomic; used by bench.py for its "gpu_launches" claim). */
uint64_t ofr_launch_count(void);

/* ------------------------------------------------------------------ Correlation
 * Replaces Correlation(pad_size, kernel_size, max_displacement, stride1, stride2,
 * corr_multiply).forward/backward -- ctor sites flownet2/networks/FlowNetC.py:28,31
 * and PWC/models/PWCNet.py:69; call sites FlowNetC.py:88, PWCNet.py:200-259;
 * alias FlowNetS/util.py:58-72 (spatial_correlation_sample(...)/C).
 *
 *   out[b,(tj+r)*D+(ti+r),y,x] = 1/(k*k*C) * sum_{c,h,w<k}
 *        P1[b,c,y1+h,x1+w] * P2[b,c,y1+tj*s2+h, x1+ti*s2+w]
 *   P = input zero-padded by `pad`, y1 = y*s1+md, r = md/s2, D = 2r+1,
 *   Hout = ceil((H + 2*pad - 2*(md + (k-1)/2)) / s1).
 * corr_multiply is accepted by the Python module and ignored (always a product).
 */
int ofr_corr_out_shape(int H, int W, int pad, int k, int md, int s1, int s2,
                       int* D, int* Hout, int* Wout);
/* Bytes of scratch the fast path of this geometry wants (0 = none).  backward != 0 sizes the backward call.
 * The stride2 = 2 configuration (FlowNetC) runs as four stride-1 correlations on the 2x2 polyphase components of
 * the inputs.  Shapes that tile (class images multiples of 8 x 16, C a multiple of 16, C <= 256 for the backward) run
 * on the tcgen05 tensor cores, address every tensor in place and need no scratch (0 is returned); the FP32 FMA
 * fallback needs room for the split copies.  Passing ws = NULL (or too little) is legal: the call then takes a
 * slower kernel that needs no scratch. */
size_t ofr_corr_workspace(int B, int C, int H, int W, int pad, int k, int md, int s1, int s2, int backward);
int ofr_corr_fwd(const float* f1, const float* f2, float* out,
                 int B, int C, int H, int W,
                 int pad, int k, int md, int s1, int s2,
                 void* ws, size_t ws_bytes, ofr_stream_t stream);
/* Same with LeakyReLU(negative_slope) fused into the store of the cost volume -- the activation that follows every
 * correlation of the reference (PWC/models/PWCNet.py:201,216,230,244,260; flownet2/networks/FlowNetC.py:89;
 * FlowNetS/util.py:72).  negative_slope = 1 is ofr_corr_fwd.  The backward of the activation is the caller's
 * (mask the upstream gradient with the sign of the output, then ofr_corr_bwd). */
int ofr_corr_fwd_act(const float* f1, const float* f2, float* out,
                     int B, int C, int H, int W,
                     int pad, int k, int md, int s1, int s2, float negative_slope,
                     void* ws, size_t ws_bytes, ofr_stream_t stream);
/* gI1 / gI2 may each be NULL to skip that gradient. */
int ofr_corr_bwd(const float* f1, const float* f2, const float* gO,
                 float* gI1, float* gI2,
                 int B, int C, int H, int W,
                 int pad, int k, int md, int s1, int s2,
                 void* ws, size_t ws_bytes, ofr_stream_t stream);

/* ------------------------------------------------------------------ Resample / warps
 * One kernel family, three sampling conventions:
 *  OFR_WARP_FLOWNET2 : Resample2d(kernel_size=1, bilinear=True) -- flownet2/models.py:52,139..176.
 *                      sample (x+u, y+v) in pixels, taps clamped to the image.
 *  OFR_WARP_FLOWNET2_NEAREST : Resample2d(bilinear=False), nearest tap, clamped.
 *  OFR_WARP_STN      : opticalFlowReg.stn -- models.py:256-268 (after its F.interpolate):
 *                      g=(p+flow)*(2/W,2/H)-1, grid_sample(zeros, align_corners=True).
 *  OFR_WARP_PWC      : PWCDCNet.warp -- PWC/models/PWCNet.py:143-179:
 *                      g=2(p+flow)/max(W-1,1)-1, grid_sample(zeros, align_corners=False),
 *                      result zeroed where the warped ones-mask < 0.9999.
 * in1 [B,C,Hin,Win], flow [B,2,Hf,Wf], out [B,C,Hf,Wf].  STN and PWC need Hin==Hf, Win==Wf.
 * Images must be finite: with C <= 2 the zeros-padding modes (STN, PWC) load border-clamped taps and give the taps
 * grid_sample would skip a weight of exactly 0, so an Inf/NaN border pixel would leak into samples that fall outside
 * (0 * Inf); the C > 2 kernels predicate the loads instead.  Flows may be anything (huge / NaN positions are clamped).
 * kernel_size > 1 of upstream's Resample2d is not provided (no reference call site uses it; upstream's own kernel
 * reads out of bounds for it): the Python module raises NotImplementedError.
 */
enum {
    OFR_WARP_FLOWNET2 = 0,
    OFR_WARP_STN = 1,
    OFR_WARP_PWC = 2,
    OFR_WARP_FLOWNET2_NEAREST = 3
};
int ofr_resample_fwd(const float* in1, const float* flow, float* out,
                     int B, int C, int Hin, int Win, int Hf, int Wf,
                     int mode, ofr_stream_t stream);
/* Same with the flow multiplied by flow_scale first (float32 product, as `flo * 0.625` etc. at the PWC-Net call
 * sites PWC/models/PWCNet.py:214,228,242,258). */
int ofr_resample_fwd_scaled(const float* in1, const float* flow, float* out,
                            int B, int C, int Hin, int Win, int Hf, int Wf, int mode, float flow_scale,
                            ofr_stream_t stream);
/* Bytes of scratch ofr_resample_bwd needs when gI1 != NULL (0 when only gI2 is wanted). */
size_t ofr_resample_bwd_workspace(int B, int C, int Hin, int Win, int Hf, int Wf, int mode);
/* Deterministic backward.  gI1 [B,C,Hin,Win] and/or gI2 [B,2,Hf,Wf]; either may be NULL.
 * The adjoint scatter into gI1 is run as a cell-binned gather in a fixed order
 * (bitwise reproducible), not with floating-point atomics.  `ws` must be 256-byte
 * aligned; its contents on entry are ignored. */
int ofr_resample_bwd(const float* in1, const float* flow, const float* gO,
                     float* gI1, float* gI2,
                     int B, int C, int Hin, int Win, int Hf, int Wf, int mode,
                     void* ws, size_t ws_bytes, ofr_stream_t stream);
/* Backward of ofr_resample_fwd_scaled: gI2 is the gradient of the UNSCALED flow. */
int ofr_resample_bwd_scaled(const float* in1, const float* flow, const float* gO,
                     float* gI1, float* gI2,
                     int B, int C, int Hin, int Win, int Hf, int Wf, int mode, float flow_scale,
                     void* ws, size_t ws_bytes, ofr_stream_t stream);

/* ------------------------------------------------------------------ multi-scale stn
 * The loop `for flow in flows: stn(flow, frame)` of opticalFlowReg.forward (models.py:277-281) in one launch
 * each way.  frame_i is the frame already resized to flow_i's size ([B,C,h,w]); out_i = stn(flow_i, frame_i).
 * Backward writes the flow gradients only (the frame is an input image and needs none). */
#define OFR_STN_MAX_SCALES 8
typedef struct {
    const float* frame;  /* [B,C,h,w] */
    const float* flow;   /* [B,2,h,w] */
    float* out;          /* forward:  [B,C,h,w]                 */
    const float* g_out;  /* backward: [B,C,h,w]                 */
    float* g_flow;       /* backward: [B,2,h,w]                 */
    int h, w;
} ofr_stn_scale_t;
int ofr_stn_multi_fwd(const ofr_stn_scale_t* scales, int n_scales, int B, int C, ofr_stream_t stream);
int ofr_stn_multi_bwd(const ofr_stn_scale_t* scales, int n_scales, int B, int C, ofr_stream_t stream);

/* ------------------------------------------------------------------ ChannelNorm
 * Replaces ChannelNorm().forward/backward -- flownet2/models.py:40,141..180.
 *   out[b,0,y,x] = sqrt(sum_c x[b,c,y,x]^2);   gI = gO * x / (out + 1e-8).
 */
int ofr_chnorm_fwd(const float* x, float* out, int B, int C, int H, int W, ofr_stream_t stream);
int ofr_chnorm_bwd(const float* x, const float* out, const float* gO, float* gI,
                   int B, int C, int H, int W, ofr_stream_t stream);
/* ChannelNorm of a difference, ||x - x2||, without materialising it: the `x[:, :3] - resampled` -> channelnorm pairs of
 * flownet2/models.py:140-141,152-153,166-167,179-180.  Backward: gI = gradient of x, gI2 = gradient of x2 (= -gI);
 * either may be NULL. */
int ofr_chnorm_diff_fwd(const float* x, const float* x2, float* out, int B, int C, int H, int W, ofr_stream_t stream);
int ofr_chnorm_diff_bwd(const float* x, const float* x2, const float* out, const float* gO, float* gI, float* gI2,
                        int B, int C, int H, int W, ofr_stream_t stream);

/* ------------------------------------------------------------------ Self-supervised loss
 * Replaces OFEloss and its three terms -- loss.py:9-13 (photometric), :23-35
 * (smoothness, Charbonnier), :52-64 (global NCC), :66-84 (multi-scale weights).
 * `fixed` is the full-resolution fixed image [B,1,Hs,Ws]; its bilinear
 * (align_corners=False) resize to each scale is done inside the kernel.
 * All reductions accumulate in float64 in a fixed order (bitwise reproducible).
 * Any batch / image size is accepted (beyond 8192 (image, row-strip) items per scale a CTA walks several items).
 * Scales whose width is not a multiple of 4, or whose tensors are not 16-byte aligned (a contiguous view with an odd
 * storage offset), take scalar instead of 128-bit accesses -- same results.
 * NCC of a constant image (documented divergence, pinned by tests/golden/ncc_const.npz): loss.py:58-60 returns
 * corr = 1 (loss 0) only if `x - mean(x)` is identically zero in float32, i.e. for constants whose float32 mean is
 * exact (0, 0.5, ...); for other constants (0.7, 0.3 on some reduction orders) float32 rounding noise of torch.mean
 * decides and the reference returns a loss of ~1.  This library returns corr = 1 (loss 0, zero gradient) for every
 * image with min == max, and whenever a centred second moment is not positive in float64.
 */
#define OFR_LOSS_MAX_SCALES 8
#define OFR_LOSS_STATE_DOUBLES 16 /* per scale, see loss.cu for the slot layout */
typedef struct {
    const float* warped; /* [B,1,h,w]                       */
    const float* flow;   /* [B,2,h,w]                       */
    float* g_warped;     /* backward only: [B,1,h,w] or NULL */
    float* g_flow;       /* backward only: [B,2,h,w] or NULL */
    int h, w;
} ofr_loss_scale_t;
size_t ofr_loss_workspace(int n_scales);
/* out4 (device, 4 doubles) <- (photometric, ncc, smoothness, total) exactly as OFEloss
 * returns them; state (device, n*OFR_LOSS_STATE_DOUBLES doubles) <- moments kept for backward. */
int ofr_loss_fwd(const float* fixed, int B, int Hs, int Ws,
                 const ofr_loss_scale_t* scales, int n_scales,
                 double lamb_da, double gamma, double zeta,
                 double* out4, double* state,
                 void* ws, size_t ws_bytes, ofr_stream_t stream);
/* grad4 (device, 4 doubles) = upstream gradients of (photometric, ncc, smoothness, total). */
int ofr_loss_bwd(const float* fixed, int B, int Hs, int Ws,
                 const ofr_loss_scale_t* scales, int n_scales,
                 double lamb_da, double gamma, double zeta,
                 const double* grad4, const double* state, ofr_stream_t stream);

/* Fused registration loss: the `[self.stn(flow, moving) for flow in flow_predictions]` of
 * opticalFlowReg.forward (models.py:277) followed by OFEloss (train.py:56 -> loss.py:66-84) in ONE pass per
 * direction.  `moving` is the moving image already resized to the scale (the F.interpolate of models.py:258 stays
 * with the caller).  Forward writes the warped image of every scale to `warped` (the model returns them,
 * models.py:289; the backward reads them back) and the four loss values to out4; backward writes
 * d total / d flow = smoothness gradient + photometric/NCC gradients chained through the warp to `g_flow` -- the
 * gradient of the warped images is never materialised.  Bit-identical to ofr_stn_multi_fwd + ofr_loss_fwd and to
 * ofr_loss_bwd + ofr_stn_multi_bwd + the float32 add autograd performs.  Workspace: ofr_loss_workspace. */
typedef struct {
    const float* moving; /* [B,1,h,w]                                     */
    const float* flow;   /* [B,2,h,w]                                     */
    float* warped;       /* [B,1,h,w]  written by fwd, read by bwd        */
    float* g_flow;       /* backward only: [B,2,h,w]                      */
    int h, w;
} ofr_regloss_scale_t;
int ofr_regloss_fwd(const float* fixed, int B, int Hs, int Ws,
                    const ofr_regloss_scale_t* scales, int n_scales,
                    double lamb_da, double gamma, double zeta,
                    double* out4, double* state,
                    void* ws, size_t ws_bytes, ofr_stream_t stream);
int ofr_regloss_bwd(const float* fixed, int B, int Hs, int Ws,
                    const ofr_regloss_scale_t* scales, int n_scales,
                    double lamb_da, double gamma, double zeta,
                    const double* grad4, const double* state, ofr_stream_t stream);

/* ------------------------------------------------------------------ Flow heads of the decoder
 * Replace the cuDNN calls behind the two-output-channel layers of PWC/models/PWCNet.py:
 *   predict_flow(in_planes) = nn.Conv2d(in_planes, 2, kernel_size=3, stride=1, padding=1)          -- :30-31, 81..121, 134
 *   deconv(in_planes, 2)    = nn.ConvTranspose2d(in_planes, 2, kernel_size=4, stride=2, padding=1) -- :33-34, 82..124
 * Weights in PyTorch's layouts: Conv2d [2][C][3][3], ConvTranspose2d [C][2][4][4]; bias [2] or NULL.  fp32 arithmetic.
 * W % 4 == 0 and 16-byte aligned x / out / gO / gX, else OFR_ENOTSUP (the caller keeps its own convolution).
 *   flowconv3:   x [B,C,H,W] -> out [B,2,H,W]          flowdeconv4: x [B,C,H,W] -> out [B,2,2H,2W]
 * Backward: gX (gradient of x), gW + gB (gradients of weight and bias; gB may be NULL) -- gX or gW may be NULL.
 * Workspace: `*_workspace(..., backward)` bytes, 256-byte aligned (0 = none needed).  Deterministic (no atomics). */
size_t ofr_flowconv3_workspace(int B, int C, int H, int W, int backward);
int ofr_flowconv3_fwd(const float* x, const float* w, const float* bias, float* out, int B, int C, int H, int W,
                      void* ws, size_t ws_bytes, ofr_stream_t stream);
int ofr_flowconv3_bwd(const float* x, const float* w, const float* gO, float* gX, float* gW, float* gB,
                      int B, int C, int H, int W, void* ws, size_t ws_bytes, ofr_stream_t stream);
size_t ofr_flowdeconv4_workspace(int B, int C, int H, int W, int backward);
int ofr_flowdeconv4_fwd(const float* x, const float* w, const float* bias, float* out, int B, int C, int H, int W,
                        void* ws, size_t ws_bytes, ofr_stream_t stream);
int ofr_flowdeconv4_bwd(const float* x, const float* w, const float* gO, float* gX, float* gW, float* gB,
                        int B, int C, int H, int W, void* ws, size_t ws_bytes, ofr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OFR_B200_H */
