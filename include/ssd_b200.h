/*
 * ssd_b200.h -- C ABI of libssd_b200.so, the B200 (sm_100a) engine behind the
 * SSD-VGG hot path of ljanyst/ssd-tensorflow.
 *
 * The reference has no native/FFI layer of its own: its hot path is the Python
 * surface of ssdvgg.py / ssdutils.py / transforms.py sitting on TensorFlow 1.x.
 * Each entry point below therefore cites the reference *Python* interface it
 * replaces (file:line in the reference tree); INTEGRATION.md shows the ctypes
 * stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain C types only: pointers, sizes, scalars.  No torch / C++ types.
 *   - "_dev" pointers are CUDA device pointers on the current device, "_host"
 *     pointers are host memory (pinned memory makes the copies asynchronous).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Device entry points are asynchronous on that stream; host entry points
 *     synchronise the stream before returning.
 *   - every function returns 0 on success, a negative SSDB_E* code on failure;
 *     ssdb_last_error() returns a thread-local message for the last failure.
 *   - a handle is not thread-safe: one handle per GPU per host thread.
 *   - tensors are float32, NHWC / [B, A, C+5] row-major unless stated.
 *     C = number of object classes (20), C+1 logits with background LAST,
 *     row = C+1 scores | 4 box offsets  (ssdvgg.py:106-107,365-372).
 */
#ifndef SSD_B200_H
#define SSD_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSDB_OK          0
#define SSDB_EINVAL     -1   /* bad argument                                   */
#define SSDB_ECUDA      -2   /* CUDA runtime / driver error                    */
#define SSDB_ENOGPU     -3   /* no sm_100 device: there is NO CPU fallback     */
#define SSDB_ESTATE     -4   /* call order violated (e.g. backward w/o forward)*/
#define SSDB_ENOTFOUND  -5   /* unknown preset / tensor name                   */

/* conv implementation selector for the ssdb_op_conv* test hooks */
#define SSDB_CONV_AUTO   0   /* what the engine would pick for the shape (split tensor-core kernels where they apply) */
#define SSDB_CONV_SIMT   1   /* CUDA-core implicit GEMM (odd shapes, tails)    */
#define SSDB_CONV_TC     2   /* tcgen05 / TMEM / TMA implicit GEMM, tf32 operands (comparison mode)   */
#define SSDB_CONV_TC_SPLIT 3 /* tcgen05 / TMEM / TMA implicit GEMM, split bf16 operands: every fp32
                                operand is a (hi, lo) bf16 pair and every product three kind::f16 MMAs
                                (hi*hi + lo*hi + hi*lo) -- the engine's default: fp32-grade results      */

typedef struct ssdb_net ssdb_net;

int         ssdb_version(void);
const char* ssdb_last_error(void);
/* 0 when a CUDA device with compute capability 10.x is current, else SSDB_ENOGPU */
int         ssdb_device_ok(void);

/* ------------------------------------------------------------------------
 * Box path (stateless).  Replaces ssdutils.py / transforms.py NumPy code.
 * anchors_prop: [A,4] float64 (cx, cy, w, h) proportional, in anchor order
 *   map -> box type -> row -> col   (ssdutils.get_anchors_for_preset :76-117)
 * ------------------------------------------------------------------------ */

/* Anchor matching + dense label creation for a batch of ground-truth sets.
 * Replaces LabelCreatorTransform.__call__ (transforms.py:72-114), which calls
 * ssdutils.compute_overlap (:155-170), jaccard_overlap (:138-152) and
 * compute_location (:173-179); quantisation per utils.prop2abs (utils.py:100-108).
 *   gt         [B, G, 5] float64 rows (labelid, cx, cy, w, h); rows >= gt_count[b] ignored
 *   match_out  [B, A] int32: GT row owning the anchor, -1 = background   (may be NULL)
 *   labels_out [B, A, C+5] float32 dense label tensor                    (may be NULL) */
int ssdb_match_anchors(const double* gt_dev, const int* gt_count_dev, int B, int G,
                       const double* anchors_prop_dev, int A, int C,
                       int* match_out_dev, float* labels_out_dev, void* stream);
int ssdb_match_anchors_host(const double* gt_host, const int* gt_count_host, int B, int G,
                            const double* anchors_prop_host, int A, int C,
                            int* match_out_host, float* labels_out_host);

/* Fused decode + top-k + class-wise greedy NMS, one image per CTA.
 * Replaces ssdutils.decode_boxes (:192-229) [arg-max class over the C object
 * classes, top-`cap` by confidence, stop below conf_thr, decode_location
 * (:182-189), utils.normalize_box (utils.py:118-135)] followed by
 * ssdutils.suppress_overlaps / non_maximum_suppression (:232-318) at iou_thr.
 *   pred       [B, A, C+5] float32 (softmax scores | offsets), read-only
 *   cap        max candidates per image; cap <= 0 means "no cap" (detections_cap=None)
 *   dets_out   [B, cap_eff, 8] int32 rows: conf (float32 bits), labelid, xmin, xmax,
 *              ymin, ymax (1000-grid, after normalize_box), anchor index, position of the
 *              detection in the confidence-sorted candidate list
 *              cap_eff = cap if cap > 0 else A.  Output order = the reference's:
 *              classes by first appearance, confidence-descending inside a class.
 *   counts_out [B, 2] int32: (kept detections, candidates that entered NMS)
 * conf_thr is compared in float32 and iou_thr in float64, as NumPy does in the reference.
 * Ties in confidence are broken by lower anchor index (NumPy leaves it unspecified).
 * Two launches (anchor-tile scan, then one CTA per image); the 5-byte-per-anchor workspace between them is a
 * grow-only buffer kept per stream by the library (allocated on first use / when B*A grows, never per call). */
int ssdb_decode_nms(const float* pred_dev, int B, int A, int C, const double* anchors_prop_dev,
                    float conf_thr, int cap, double iou_thr,
                    int* dets_out_dev, int* counts_out_dev, void* stream);
int ssdb_decode_nms_host(const float* pred_host, int B, int A, int C, const double* anchors_prop_host,
                         float conf_thr, int cap, double iou_thr,
                         int* dets_out_host, int* counts_out_host);

/* Class-wise greedy NMS over already-decoded boxes: ssdutils.suppress_overlaps /
 * non_maximum_suppression (:232-318) when a caller runs them separately from decode_boxes.
 *   boxes_abs [n,4] int32 (xmin,xmax,ymin,ymax) exactly as utils.prop2abs(box, 1000x1000) yields
 *   labelid   [n] int32 in [0, nclass)    conf [n] float32
 *   keep_out  [n] int32: indices of the kept boxes in the reference's output order (classes by
 *             first appearance in the input list, confidence-descending inside a class)
 *   count_out [1] int32 */
int ssdb_nms_host(const int* boxes_abs_host, const int* labelid_host, const float* conf_host, int n, int nclass,
                  double iou_thr, int* keep_out_host, int* count_out_host);

/* ------------------------------------------------------------------------
 * Multibox loss (stateless).  Replaces SSDVGG.build_optimizer's loss graph
 * (ssdvgg.py:380-580): softmax-CE + smooth-L1 (:68-71), per-image 3:1 hard
 * negative mining via top_k (:463-501), per-image normalisation by the number
 * of positives (:513-521,552-560), batch mean.
 *   output     [B, A, C+5] raw head output (logits | offsets)
 *   labels     [B, A, C+5] dense labels (net.labels placeholder, ssdvgg.py:378)
 *   losses_out [2] float32: confidence_loss, localization_loss
 *   grad_out   [B, A, C+5] d(conf+loc)/d(output) * grad_scale      (may be NULL)
 *   result_out [B, A, C+5] softmax(logits) | offsets = net.result (:368-372) (may be NULL)
 * ssdb_multibox_loss_gt is the fused variant: anchor matching happens inside the
 * loss kernels from raw ground truth (no dense label tensor in HBM), in exact integer arithmetic.
 * Three streaming launches (anchor tiles, per-image selection, gradient tiles; the fused variant adds two small
 * matching launches); their ~6-byte-per-anchor workspace is a grow-only buffer kept by the library (one caller
 * thread per process, like the reference's single session). */
int ssdb_multibox_loss(const float* output_dev, const float* labels_dev, int B, int A, int C,
                       float grad_scale, float* losses_out_dev, float* grad_out_dev,
                       float* result_out_dev, void* stream);
int ssdb_multibox_loss_gt(const float* output_dev, const double* gt_dev, const int* gt_count_dev,
                          int B, int G, const double* anchors_prop_dev, int A, int C,
                          float grad_scale, float* losses_out_dev, float* grad_out_dev,
                          float* result_out_dev, int* match_out_dev, void* stream);

/* ------------------------------------------------------------------------
 * Layer test hooks (stateless): one convolution through a chosen kernel.
 * TF semantics of tf.nn.conv2d / atrous_conv2d + bias_add + relu as used by
 * conv_map / classifier (ssdvgg.py:42-65,260) -- NHWC x, HWIO w.
 *   pad_t/pad_l: zeros before the first row/col (TF SAME: total//2); the
 *   output size fixes the padding after.   All pointers device.
 * ------------------------------------------------------------------------ */
int ssdb_op_conv_fprop(int impl, const float* x, const float* w_hwio, const float* bias,
                       int B, int H, int W, int Cin, int Cout, int k, int stride, int dil,
                       int pad_t, int pad_l, int Ho, int Wo, int relu, float* y, void* stream);
/* dx = conv_transpose(dz, w);  if mask_x != NULL: dx *= (mask_x > 0);  beta in {0,1}: dx += old */
int ssdb_op_conv_dgrad(int impl, const float* dz, const float* w_hwio, const float* mask_x,
                       int B, int H, int W, int Cin, int Cout, int k, int stride, int dil,
                       int pad_t, int pad_l, int Ho, int Wo, int beta, float* dx, void* stream);
/* dw[k,k,Cin,Cout] = x (*) dz ;  db[Cout] = sum dz   (db may be NULL) */
int ssdb_op_conv_wgrad(int impl, const float* x, const float* dz,
                       int B, int H, int W, int Cin, int Cout, int k, int stride, int dil,
                       int pad_t, int pad_l, int Ho, int Wo, float* dw, float* db, void* stream);

/* Layer micro-benchmark hook: device time (ms, CUDA events, 2 warm-up launches, mean of `iters`) of ONE convolution kernel
 * of the engine on synthetic data already in the engine's operand format (activations half zeros like a ReLU output,
 * gradients dense), i.e. the bare kernel as a training step launches it.  kind: 0 fprop (+bias+ReLU), 1 dgrad
 * (with_mask: ReLU mask of x; beta: accumulate), 2 wgrad (+bias gradient).  impl as above (AUTO = split tensor-core). */
int ssdb_op_conv_bench(int kind, int impl, int B, int H, int W, int Cin, int Cout, int k, int stride, int dil,
                       int pad_t, int pad_l, int Ho, int Wo, int with_mask, int beta, int iters, float* ms_out);

/* ------------------------------------------------------------------------
 * Network engine.  Replaces SSDVGG (ssdvgg.py:87-649) + the tf.Session that
 * runs it (train.py:166,262-266; infer.py:211,225-227).
 * ------------------------------------------------------------------------ */

/* SSDVGG(session, preset).build_from_vgg(vgg_dir, num_classes) (ssdvgg.py:89-118):
 * builds the layer plan for `preset` ("vgg300" | "vgg512", ssdutils.py:36-62),
 * allocates parameters / gradients / momentum (flat float32 buffers) and the
 * activation workspace for up to `max_batch` images.  Parameters start at zero:
 * load them with ssdb_set_tensor.  flags: 0, or SSDB_FLAG_INFERENCE. */
#define SSDB_FLAG_INFERENCE 1u   /* frozen model (export_model.py:62-72 / detect.py:90-112): forward + detection only.  No
                                    gradient, momentum, label or loss buffers (about half the memory), training entry points
                                    return SSDB_EINVAL, and ssdb_forward_detect_host replays the whole device side -- ~60 forward
                                    launches, softmax, decode + NMS -- as ONE CUDA graph per (batch, threshold, cap, IoU) */
int ssdb_create(const char* preset, int num_classes, int max_batch, unsigned flags, ssdb_net** out);
int ssdb_destroy(ssdb_net* net);

int ssdb_num_anchors(const ssdb_net* net);
int ssdb_image_size(const ssdb_net* net);
/* Trainable tensors under the reference's variable names ("conv1_1/filter",
 * "mod_conv6/biases", "classifiers/classifier0_1/filter", "l2_norm_conv4_3/scale";
 * ssdvgg.py:44,47,82,602-622).  Shapes are the reference's (HWIO filters). */
int ssdb_num_tensors(const ssdb_net* net);
int ssdb_tensor_info(const ssdb_net* net, int index, char* name_out, int name_cap,
                     int* rank_out, int shape_out[4]);
/* which: 0 = parameter, 1 = gradient (after a backward), 2 = momentum accumulator */
int ssdb_get_tensor(ssdb_net* net, const char* name, int which, float* host_out, long long count);
int ssdb_set_tensor(ssdb_net* net, const char* name, int which, const float* host_in, long long count);
/* The flat device buffers (for the data-parallel all-reduce on gradients only). */
int ssdb_flat_buffer(ssdb_net* net, int which, void** dev_ptr_out, long long* count_out);
/* Gradient buckets for a data-parallel all-reduce that overlaps the backward.  The flat gradient buffer is cut into
 * contiguous ranges that become final in the order the backward writes them -- [mod_conv6 .. end) (conv6 / conv7, extra
 * layers, scale, classifiers) first, then [conv4_1 .. mod_conv6), then [conv1_1 .. conv4_1) -- and the engine records a
 * CUDA event on the step's stream when a range is complete.  ssdb_grad_buckets fills up to `cap` (begin, end) float
 * offsets and returns the number of buckets; ssdb_wait_grad_bucket makes `stream` wait for bucket `bucket` of the LAST
 * ssdb_train_step (cudaStreamWaitEvent), so the caller can start that range's ncclAllReduce on a side stream while the
 * remaining dgrad / wgrad kernels run. */
int ssdb_grad_buckets(const ssdb_net* net, int cap, long long* begin_out, long long* end_out);
int ssdb_wait_grad_bucket(ssdb_net* net, int bucket, void* stream);
/* Input pre-processing that the reference leaves to the third-party VGG graph
 * (ssdvgg.py:190-207): y[c] = x[swap_rb ? 2-c : c] - mean[c].  Default: swap, ImageNet means. */
int ssdb_set_preprocess(ssdb_net* net, int swap_rb, const float mean[3]);

/* sess.run(net.result, {image_input: x}) (infer.py:225-227): forward only.
 *   images  [B, S, S, 3] float32, raw 0..255 as cv2 delivers (infer.py:51-52)
 *   result  [B, A, C+5] softmax scores | offsets (may be NULL: keep on device) */
int ssdb_forward(ssdb_net* net, const float* images_dev, int B, float* result_dev, void* stream);
int ssdb_forward_host(ssdb_net* net, const float* images_host, int B, float* result_host);

/* sess.run(net.logits, ...) (ssdvgg.py:365-366): the raw head output of the LAST forward / train / eval call,
 * [B, A, C+5] = C+1 class logits (pre-softmax) | 4 box offsets, copied to host memory. */
int ssdb_read_output_host(ssdb_net* net, int B, float* output_host);

/* sess.run([net.result, net.losses, net.optimizer], {image_input, labels})
 * (train.py:262-266): forward + multibox loss + backward + Momentum update
 * (ssdvgg.py:375-599; accum = mu*accum + g, var -= lr*accum, L2 on filters).
 *   labels      [B, A, C+5] dense labels, OR NULL with gt/gt_count given (fused match)
 *   losses_out  [4] float32: total, localization, confidence, l2  (net.losses)
 *   result      pre-update forward result (may be NULL)
 *   apply_update 0: stop after backward (gradients in the flat buffer, scaled by
 *               grad_scale) so the caller can all-reduce them, then ssdb_apply_update. */
int ssdb_train_step(ssdb_net* net, const float* images_dev, const float* labels_dev,
                    const double* gt_dev, const int* gt_count_dev, int G, int B,
                    float lr, float momentum, float weight_decay, float grad_scale,
                    int apply_update, float* losses_out_dev, float* result_dev, void* stream);
int ssdb_train_step_host(ssdb_net* net, const float* images_host, const float* labels_host,
                         int B, float lr, float momentum, float weight_decay,
                         float* losses_out_host, float* result_host);
/* Data-parallel flavour of ssdb_train_step_host: forward + loss + backward with the same overlapped host copies, but NO
 * update -- the caller all-reduces ssdb_flat_buffer(net, 1) across ranks and then calls ssdb_apply_update(1/world). */
int ssdb_train_step_host_noupdate(ssdb_net* net, const float* images_host, const float* labels_host, int B,
                                  float weight_decay, float* losses_out_host, float* result_host);
/* The same training step fed with RAW ground truth instead of the dense label tensor: anchor matching
 * (LabelCreatorTransform, transforms.py:72-114, which the reference runs in its data-loader workers and ships to the
 * device as 873 KB of labels per image, train.py:257-266) happens inside the fused match + loss kernels, in exact
 * integer arithmetic; the host -> device traffic per image drops from 1.95 MB to 1.08 MB.
 *   gt        [B, G, 5] float64 rows (labelid, cx, cy, w, h), proportional coordinates like utils.Box; G <= 128
 *   gt_count  [B] int32 valid rows per image (0 allowed: an all-background image contributes zero loss, ssdvgg.py:417-419)
 *   apply_update 1: full step; 0: stop after the backward, gradients in the flat buffer (data-parallel callers
 *             all-reduce, then ssdb_apply_update); -1: forward + loss only (the validation pass, train.py:291-294)
 *   match_out [B, A] int32 owner GT row per anchor, -1 = background (may be NULL)
 * Label ids outside [0, C) are rejected with SSDB_EINVAL (the reference would raise an IndexError). */
int ssdb_train_step_host_gt(ssdb_net* net, const float* images_host, const double* gt_host, const int* gt_count_host,
                            int G, int B, float lr, float momentum, float weight_decay, int apply_update,
                            float* losses_out_host, float* result_host, int* match_out_host);

/* The data-parallel host-fed step in two halves.  _begin enqueues upload, forward, loss, backward and the result download on
 * the engine's streams and returns WITHOUT waiting (labels_host, or NULL with gt_host / gt_count_host / G); the caller
 * overlaps the bucketed gradient all-reduce (ssdb_grad_buckets / ssdb_wait_grad_bucket) and ssdb_apply_update on its own
 * stream with the rest of the backward, then calls _end, which waits for the engine's streams and returns net.losses.
 * The host buffers must stay valid until _end returns. */
int ssdb_train_step_host_begin(ssdb_net* net, const float* images_host, const float* labels_host, const double* gt_host,
                               const int* gt_count_host, int G, int B, float weight_decay, float* result_host,
                               int* match_out_host);
int ssdb_train_step_host_end(ssdb_net* net, float* losses_out_host);

/* infer.py:225-235 / detect.py:103-112 as ONE call: sess.run(net.result) followed per image by decode_boxes +
 * suppress_overlaps.  The result tensor stays in device memory, the fused decode + top-k + class-wise NMS kernels
 * (ssdb_decode_nms) read it there, and only the detections come back: [B, cap_eff, 8] int32 rows + [B, 2] counts in
 * the layout of ssdb_decode_nms (cap <= 0: no cap).  result_host (may be NULL) additionally receives net.result. */
int ssdb_forward_detect_host(ssdb_net* net, const float* images_host, int B, float conf_thr, int cap, double iou_thr,
                             int* dets_out_host, int* counts_out_host, float* result_host);

/* validation pass (train.py:291-294): forward + loss, no backward */
int ssdb_eval_step(ssdb_net* net, const float* images_dev, const float* labels_dev, int B,
                   float weight_decay, float* losses_out_dev, float* result_dev, void* stream);
/* w-gradient post-scale (1/world after an all-reduce) + weight decay + momentum + update */
int ssdb_apply_update(ssdb_net* net, float lr, float momentum, float weight_decay,
                      float grad_post_scale, void* stream);

/* page-locked host memory for feeds / fetches (cudaHostAlloc): makes the host entry points' copies asynchronous DMA */
int ssdb_pinned_alloc(long long bytes, void** host_ptr_out);
int ssdb_pinned_free(void* host_ptr);

/* CRC32C (Castagnoli) of a HOST buffer, continuing from `crc` (0 to start).  Not on the compute path: it is the
 * checksum TensorFlow's checkpoint-V2 "tensor bundle" files carry per table block and per tensor
 * (tf.train.Saver, train.py:208,336-343; the VGG saved-model, ssdvgg.py:190-207), used by tf_bundle.py to read / write
 * those files without TensorFlow when payloads are hundreds of megabytes. */
unsigned int ssdb_crc32c(unsigned int crc, const void* data_host, size_t bytes);

/* Test / diagnosis hooks: an intermediate tensor of the last step as plain float32 NHWC on the host.
 * name = an op of the plan ("conv1_1" .. "conv11_2", "pool1" .. "pool4", "mod_pool5", "l2_norm_conv4_3") for the activation it
 * wrote, "grad:<op>" for the gradient with respect to it, "output" / "output_grad" for the raw head output [B, A, C+5] and
 * its gradient.  ssdb_debug_shape: (H, W, C) of an op's activation. */
int ssdb_debug_read(ssdb_net* net, const char* name, int B, float* host_out, long long count);
int ssdb_debug_shape(const ssdb_net* net, const char* name, int shape_out[3]);

/* number of kernels this library launched since load (bench.py's gpu_launches) */
long long ssdb_launch_count(void);
/* per-kernel-family device time of the LAST ssdb_profile_step (ms): fills up to cap
 * (name, ms, launches) triples; returns the number of families */
int ssdb_profile_step(ssdb_net* net, const float* images_dev, const float* labels_dev, int B,
                      char names_out[][32], float* ms_out, int* launches_out, int cap);

#ifdef __cplusplus
}
#endif
#endif /* SSD_B200_H */
