// Shared declarations of libssd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/ssd_b200.h"

namespace ssdb {

void set_error(const char* fmt, ...);
extern long long g_launches;          // kernels launched by this library

#define SSDB_CUDA(call)                                                              \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            ::ssdb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,          \
                              cudaGetErrorString(e__));                              \
            return SSDB_ECUDA;                                                       \
        }                                                                            \
    } while (0)

#define SSDB_LAUNCH_CHECK()                                                          \
    do {                                                                             \
        ++::ssdb::g_launches;                                                        \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            ::ssdb::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,      \
                              cudaGetErrorString(e__));                              \
            return SSDB_ECUDA;                                                       \
        }                                                                            \
    } while (0)

#define SSDB_REQUIRE(cond, msg)                                                      \
    do {                                                                             \
        if (!(cond)) {                                                               \
            ::ssdb::set_error("%s:%d: %s (%s)", __FILE__, __LINE__, msg, #cond);     \
            return SSDB_EINVAL;                                                      \
        }                                                                            \
    } while (0)

// Geometry of one convolution (TF semantics; NHWC activations, HWIO filter).
struct ConvGeom {
    int B, H, W, Cin;        // input
    int Ho, Wo, Cout;        // output (Cout = channel stride of y / dz and of the HWIO filter)
    int k, stride, dil;
    int pad_t, pad_l;        // zeros before the first row / column
};

// How a forward conv stores its result.
struct ConvEpilogue {
    const float* bias = nullptr;   // [Cout] or null
    int relu = 0;
    // head scatter: y is the [B, A, V] output tensor; channel n = j*V + v goes to
    // row (anchor_base + j*Ho*Wo + pixel), column v   (ssdvgg.py:63,356-366)
    int scatter = 0;
    int V = 0;                      // C+5
    int n_valid = 0;                // box types * V (channels >= n_valid are padding)
    int anchor_base = 0;
    int A = 0;
    // fused input pre-processing for conv1_1 (Cin == 3): x[c] = raw[swap ? 2-c : c] - mean[c]
    int preprocess = 0;
    int swap_rb = 0;
    float mean[3] = {0, 0, 0};
    // round the stored activation to tf32 (round-to-nearest) so that the tensor-core kernels that
    // read it later see exactly representable operands (the MMA itself truncates)
    int round_tf32 = 0;
};

#ifdef __CUDACC__
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
#endif

// ---- SIMT (CUDA-core) implicit GEMM: every shape, used for tails / stride 2 / Cin=3 ----
int conv_simt_fprop(const ConvGeom& g, const float* x, const float* w, const ConvEpilogue& ep,
                    float* y, cudaStream_t st);
int conv_simt_dgrad(const ConvGeom& g, const float* dz, const float* w, const float* mask_x,
                    int beta, int round_out, float* dx, cudaStream_t st);
// partial: workspace of at least conv_simt_wgrad_ws(g) floats
size_t conv_simt_wgrad_ws(const ConvGeom& g);
int conv_simt_wgrad(const ConvGeom& g, const float* x, const float* dz, const ConvEpilogue& ep,
                    float* dw, float* partial, cudaStream_t st);
// db[n] = sum over pixels of dz[p][n]  (deterministic two-stage), partial >= 1184*Cout floats
int bias_grad(const float* dz, long long pixels, int Cout, float* db, float* partial, cudaStream_t st);

// ---- tcgen05 / TMEM / TMA implicit GEMM (tf32 operands, fp32 accumulate) ----
bool conv_tc_supported_fprop(const ConvGeom& g);
bool conv_tc_supported_dgrad(const ConvGeom& g);
bool conv_tc_supported_wgrad(const ConvGeom& g);
// w_t: per-tap transposed filter [k*k][CoutPad][Cin] (K-major B operand), see pack_filter_t
int conv_tc_fprop(const ConvGeom& g, const float* x, const float* w_t, int cout_pad,
                  const ConvEpilogue& ep, float* y, cudaStream_t st);
int conv_tc_dgrad(const ConvGeom& g, const float* dz, const float* w_hwio, const float* mask_x,
                  int beta, int round_out, float* dx, cudaStream_t st);
size_t conv_tc_wgrad_ws(const ConvGeom& g);
// db != null: the bias gradient is produced by the same kernel (all-ones slot)
int conv_tc_wgrad(const ConvGeom& g, const float* x, const float* dz, float* dw, float* db, float* partial,
                  cudaStream_t st);
int pack_filter_t(const float* w_hwio, int taps, int Cin, int Cout, int cout_pad, float* w_t,
                  cudaStream_t st);

// ---- pools, L2 norm ----
int maxpool_fwd(const float* x, int B, int H, int W, int C, int k, int stride, int pad_t, int pad_l,
                int Ho, int Wo, float* y, cudaStream_t st);
// dx = (beta*dx + routed dy) * (x > 0)
int maxpool_bwd(const float* x, const float* dy, int B, int H, int W, int C, int k, int stride,
                int pad_t, int pad_l, int Ho, int Wo, int beta, int relu_mask, int round_out, float* dx, cudaStream_t st);
// overlapping pools (3x3 stride 1): forward records the winning window cell (1 byte per element), backward routes by it
int maxpool_fwd_arg(const float* x, int B, int H, int W, int C, int k, int stride, int pad_t, int pad_l, int Ho, int Wo,
                    float* y, unsigned char* arg, cudaStream_t st);
int maxpool_bwd_arg(const float* x, const float* dy, const unsigned char* arg, int B, int H, int W, int C, int k, int stride, int pad_t,
                    int pad_l, int Ho, int Wo, int beta, int relu_mask, int round_out, float* dx, cudaStream_t st);
// non-overlapping 2x2/s2 pools: one code byte per output element (winning cell + "winner > 0"), so the backward reads no activations
int maxpool2x2_fwd_code(const float* x, int B, int H, int W, int C, int Ho, int Wo, float* y, unsigned char* code, cudaStream_t st);
int maxpool2x2_bwd_code(const float* dy, const unsigned char* code, int B, int H, int W, int C, int Ho, int Wo, int relu_mask,
                        int round_out, float* dx, cudaStream_t st);
int l2norm_fwd(const float* x, const float* scale, long long pixels, int C, int round_out, float* y, cudaStream_t st);
int l2norm_bwd(const float* x, const float* scale, const float* dy, long long pixels, int C, int beta,
               int round_out, float* dx, float* dscale, float* partial, cudaStream_t st);

// ---- conv1_1 (Cin = 3): explicit 3x3 patch matrix so that the layer runs as a 1x1 tensor-core conv ----
// patches[B*S*S][32]: 27 pre-processed (mean-subtracted, optionally R/B-swapped) taps in (kh, kw, c) order + 5 zeros,
// tf32-rounded; SAME padding (zeros outside the image, after pre-processing, as in the graph)
int conv1_im2col(const float* images, int B, int S, int swap_rb, const float mean[3], float* patches, cudaStream_t st);
// w32[32][Cout] <- w[27][Cout] (rows 27..31 zero)
int conv1_pad_filter(const float* w27, int Cout, float* w32, cudaStream_t st);

// ---- head layout helpers ----
// dz[B,H,W,Npad] (NHWC, zero padded channels) <- grad[B,A,V]
int head_grad_gather(const float* grad, int B, int A, int V, int anchor_base, int HW, int nbox,
                     int Npad, int round_out, float* dz, cudaStream_t st);
// dst[i] = tf32_rn(src[i])
int round_tf32_copy(const float* src, float* dst, long long n, cudaStream_t st);
int softmax_result(const float* output, long long rows, int C, float* result, cudaStream_t st);

// ---- optimizer ----
// g = g*post_scale + wd*w (filters only); v = mu*v + g; w -= lr*v   over [begin,end) of the flat buffer
int sgd_momentum(float* w, float* g, float* v, long long n, const unsigned char* decay_mask_per_block,
                 float lr, float mu, float wd, float post_scale, cudaStream_t st);
// out[0] = sum over decayed elements of w^2/2 (deterministic)
int l2_sum(const float* w, long long n, const unsigned char* decay_mask_per_block, float* partial,
           float* out, cudaStream_t st);
constexpr int OPT_BLOCK = 1024;   // elements per decay-mask entry (tensors are padded to this)

// ---- loss / match / detect (see loss.cu, detect.cu) ----
int multibox_loss_launch(const float* output, const float* labels, const double* gt, const int* gt_count,
                         int G, const double* anchors_prop, int B, int A, int C, float grad_scale,
                         float* losses_out, float* grad_out, float* result_out, int* match_out,
                         void* ws, cudaStream_t st);
// workspace of multibox_loss_launch: zero it once when allocated
size_t multibox_loss_ws_bytes(int B, int A);
int match_anchors_launch(const double* gt, const int* gt_count, int B, int G, const double* anchors_prop,
                         int A, int C, int* match_out, float* labels_out, cudaStream_t st);
int decode_nms_launch(const float* pred, int B, int A, int C, const double* anchors_prop, float conf_thr,
                      int cap, double iou_thr, int* dets_out, int* counts_out, void* scratch,
                      size_t scratch_bytes, cudaStream_t st);
size_t decode_nms_scratch_bytes(int B, int A, int cap);
int nms_only_host(const int* boxes, const int* cls, const float* conf, int n, int nclass, double iou_thr, int* keep_out, int* count_out);

}  // namespace ssdb
