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

// Lazily created per-DEVICE state (shared-memory attributes, lookup tables, scratch): a process may drive several GPUs, and
// a pointer or a function attribute set up on one device means nothing on another.  `static PerDevice<T> x;` + `x.get()`.
constexpr int MAX_DEVICES = 64;
inline int cur_device() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < MAX_DEVICES) ? d : 0; }
template <class T>
struct PerDevice {
    T v[MAX_DEVICES] = {};
    T& get() { return v[cur_device()]; }
};

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
    // fused 2x2 / stride-2 max pool (tensor-core fprop in split mode only, see conv_tc_fprop_can_pool): the kernel writes the
    // pooled activation [B, Ho/2, Wo/2, Cout] and the pool's code bytes INSTEAD of y
    float* pool_dst = nullptr;
    unsigned char* pool_code = nullptr;
    // round the stored activation to tf32 (round-to-nearest) so that the tensor-core kernels that
    // read it later see exactly representable operands (the MMA itself truncates)
    int round_tf32 = 0;
};

// Storage format of activation / activation-gradient tensors (and of the packed filter copies the tensor-core kernels read).
//   ACT_F32  plain float32 NHWC (SIMT mode; tf32 mode stores tf32-rounded float32 values in it)
//   ACT_S32  "split": every aligned group of 32 consecutive elements (32 channels of one pixel: channel counts are
//            multiples of 32) occupies the same 128 bytes as in ACT_F32, but holds 32 bf16 HIGH parts followed by 32
//            bf16 LOW parts:  hi = bf16_rn(v), lo = bf16_rn(v - hi), value = hi + lo (16 significand bits, exact in
//            float32).  A 128-byte row is then at once a K-major SWIZZLE_128B operand row of 64 bf16 (tcgen05
//            kind::f16: the products hi*hi + lo*hi + hi*lo are three K = 16 MMAs per 16 channels that differ only in
//            the 32-byte K offsets of their descriptors) and, seen through a TMA map with a 64-byte inner box, two
//            MN-major SWIZZLE_64B blocks (wgrad).  Same HBM footprint and traffic as float32, ~2^-17 operand error
//            instead of tf32's 2^-11.
enum ActFmt { ACT_F32 = 0, ACT_S32 = 1 };

#ifdef __CUDACC__
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// two floats -> packed bf16x2 (round to nearest even); `a` lands in the low half
__device__ __forceinline__ uint32_t bf16x2_rn(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ float bf16_lo_f(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi_f(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }

// split two values into their (hi, lo) bf16 pairs
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = bf16x2_rn(a, b);
    lo = bf16x2_rn(a - bf16_lo_f(hi), b - bf16_hi_f(hi));
}

// byte address of element `e` (flat element index; 32-element groups are aligned because channel counts are
// multiples of 32) of an ACT_S32 tensor: the HIGH part; the LOW part is 64 bytes further
__device__ __forceinline__ const unsigned char* s32_addr(const float* t, long long e) {
    return reinterpret_cast<const unsigned char*>(t) + (e >> 5) * 128 + (e & 31) * 2;
}
__device__ __forceinline__ unsigned char* s32_addr(float* t, long long e) {
    return reinterpret_cast<unsigned char*>(t) + (e >> 5) * 128 + (e & 31) * 2;
}

// four consecutive elements starting at flat element index e (e % 4 == 0)
template <int FMT>
__device__ __forceinline__ float4 act_ld4(const float* t, long long e) {
    if (FMT == ACT_F32) return *reinterpret_cast<const float4*>(t + e);
    const unsigned char* p = s32_addr(t, e);
    const uint2 h = *reinterpret_cast<const uint2*>(p);
    const uint2 l = *reinterpret_cast<const uint2*>(p + 64);
    return make_float4(bf16_lo_f(h.x) + bf16_lo_f(l.x), bf16_hi_f(h.x) + bf16_hi_f(l.x),
                       bf16_lo_f(h.y) + bf16_lo_f(l.y), bf16_hi_f(h.y) + bf16_hi_f(l.y));
}
// only the sign / zero test of four elements (ReLU mask): the HIGH parts decide (hi == 0 <=> value == 0)
template <int FMT>
__device__ __forceinline__ float4 act_ld4_sign(const float* t, long long e) {
    if (FMT == ACT_F32) return *reinterpret_cast<const float4*>(t + e);
    const uint2 h = *reinterpret_cast<const uint2*>(s32_addr(t, e));
    return make_float4(bf16_lo_f(h.x), bf16_hi_f(h.x), bf16_lo_f(h.y), bf16_hi_f(h.y));
}
template <int FMT>
__device__ __forceinline__ void act_st4(float* t, long long e, float4 v) {
    if (FMT == ACT_F32) { *reinterpret_cast<float4*>(t + e) = v; return; }
    uint2 h, l;
    split2(v.x, v.y, h.x, l.x);
    split2(v.z, v.w, h.y, l.y);
    unsigned char* p = s32_addr(t, e);
    *reinterpret_cast<uint2*>(p) = h;
    *reinterpret_cast<uint2*>(p + 64) = l;
}
template <int FMT>
__device__ __forceinline__ float act_ld1(const float* t, long long e) {
    if (FMT == ACT_F32) return t[e];
    const unsigned char* p = s32_addr(t, e);
    const uint32_t h = *reinterpret_cast<const unsigned short*>(p), l = *reinterpret_cast<const unsigned short*>(p + 64);
    return __uint_as_float(h << 16) + __uint_as_float(l << 16);
}
#endif

// ---- SIMT (CUDA-core) implicit GEMM: every shape, used for tails / stride 2 / Cin=3 ----
// `fmt` (ActFmt) is the storage format of the activation-like tensors (x, y, dz, dx, mask); filters, biases, the head
// output (scatter) and the raw image (Cin = 3) are always plain float32.
int conv_simt_fprop(const ConvGeom& g, const float* x, const float* w, int fmt, const ConvEpilogue& ep,
                    float* y, cudaStream_t st);
int conv_simt_dgrad(const ConvGeom& g, const float* dz, const float* w, int fmt, const float* mask_x,
                    int beta, int round_out, float* dx, cudaStream_t st);
// partial: workspace of at least conv_simt_wgrad_ws(g) floats
size_t conv_simt_wgrad_ws(const ConvGeom& g);
int conv_simt_wgrad(const ConvGeom& g, const float* x, const float* dz, int fmt, const ConvEpilogue& ep,
                    float* dw, float* partial, cudaStream_t st);
// db[n] = sum over pixels of dz[p][n]  (deterministic two-stage), partial >= 1184*Cout floats
int bias_grad(const float* dz, int fmt, long long pixels, int Cout, float* db, float* partial, cudaStream_t st);

// ---- tcgen05 / TMEM / TMA implicit GEMM ----
// fmt = ACT_F32: tf32 operands (activations stored tf32-rounded), ACT_S32: split bf16 operands (3 MMAs per product);
// fp32 accumulation in tensor memory either way
bool conv_tc_supported_fprop(const ConvGeom& g);
bool conv_tc_fprop_can_pool(const ConvGeom& g, int fmt);
bool conv_tc_supported_dgrad(const ConvGeom& g);
bool conv_tc_supported_wgrad(const ConvGeom& g, int fmt);
// w_t: per-tap transposed filter [k*k][CoutPad][Cin] (K-major B operand) in format fmt, see pack_filter_t
int conv_tc_fprop(const ConvGeom& g, const float* x, const float* w_t, int cout_pad, int fmt,
                  const ConvEpilogue& ep, float* y, cudaStream_t st);
// w_hwio: the HWIO filter in format fmt (tf32-rounded floats, or split along Cout)
int conv_tc_dgrad(const ConvGeom& g, const float* dz, const float* w_hwio, int fmt, const float* mask_x,
                  int beta, int round_out, float* dx, cudaStream_t st);
size_t conv_tc_wgrad_ws(const ConvGeom& g, int fmt);
// db != null: the bias gradient is produced by the same kernel (all-ones slot); dw / db are plain float32
int conv_tc_wgrad(const ConvGeom& g, const float* x, const float* dz, int fmt, float* dw, float* db, float* partial,
                  cudaStream_t st);
int pack_filter_t(const float* w_hwio, int taps, int Cin, int Cout, int cout_pad, int fmt, float* w_t,
                  cudaStream_t st);

// ---- pools, L2 norm (activation tensors in format fmt) ----
int maxpool_fwd(const float* x, int fmt, int B, int H, int W, int C, int k, int stride, int pad_t, int pad_l,
                int Ho, int Wo, float* y, cudaStream_t st);
// dx = (beta*dx + routed dy) * (x > 0)
int maxpool_bwd(const float* x, const float* dy, int fmt, int B, int H, int W, int C, int k, int stride,
                int pad_t, int pad_l, int Ho, int Wo, int beta, int relu_mask, int round_out, float* dx, cudaStream_t st);
// overlapping pools (3x3 stride 1): forward records the winning window cell (1 byte per element), backward routes by it
int maxpool_fwd_arg(const float* x, int fmt, int B, int H, int W, int C, int k, int stride, int pad_t, int pad_l, int Ho, int Wo,
                    float* y, unsigned char* arg, cudaStream_t st);
int maxpool_bwd_arg(const float* x, const float* dy, const unsigned char* arg, int fmt, int B, int H, int W, int C, int k, int stride, int pad_t,
                    int pad_l, int Ho, int Wo, int beta, int relu_mask, int round_out, float* dx, cudaStream_t st);
// non-overlapping 2x2/s2 pools: one code byte per output element (winning cell + "winner > 0"), so the backward reads no activations
int maxpool2x2_fwd_code(const float* x, int fmt, int B, int H, int W, int C, int Ho, int Wo, float* y, unsigned char* code, cudaStream_t st);
int maxpool2x2_bwd_code(const float* dy, const unsigned char* code, int fmt, int B, int H, int W, int C, int Ho, int Wo, int relu_mask,
                        int round_out, float* dx, cudaStream_t st);
int l2norm_fwd(const float* x, const float* scale, int fmt, long long pixels, int C, int round_out, float* y, cudaStream_t st);
int l2norm_bwd(const float* x, const float* scale, const float* dy, int fmt, long long pixels, int C, int beta,
               int round_out, float* dx, float* dscale, float* partial, cudaStream_t st);

// ---- conv1_1 (Cin = 3): explicit 3x3 patch matrix so that the layer runs as a 1x1 tensor-core conv ----
// patches[B*S*S][32]: 27 pre-processed (mean-subtracted, optionally R/B-swapped) taps in (kh, kw, c) order + 5 zeros,
// in format fmt (tf32-rounded floats / split); SAME padding (zeros outside the image, after pre-processing, as in the graph)
int conv1_im2col(const float* images, int B, int S, int swap_rb, const float mean[3], int fmt, float* patches, cudaStream_t st);
// w32[32][Cout] <- w[27][Cout] (rows 27..31 zero)
int conv1_pad_filter(const float* w27, int Cout, float* w32, cudaStream_t st);

// ---- head layout helpers ----
// dz[B,H,W,Npad] (NHWC, zero padded channels, format fmt) <- grad[B,A,V]
int head_grad_gather(const float* grad, int B, int A, int V, int anchor_base, int HW, int nbox,
                     int Npad, int fmt, int round_out, float* dz, cudaStream_t st);
// dst[i] = tf32_rn(src[i])
int round_tf32_copy(const float* src, float* dst, long long n, cudaStream_t st);
// plain float32 <-> ACT_S32 (n % 32 == 0)
int split_copy(const float* src, float* dst_s32, long long n, cudaStream_t st);
int unsplit_copy(const float* src_s32, float* dst, long long n, cudaStream_t st);
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
