// Anchor matching and the fused multibox loss.
//
// match:  LabelCreatorTransform (reference transforms.py:47-114) over
//         compute_overlap / jaccard_overlap / compute_location
//         (ssdutils.py:138-179) with utils.prop2abs quantisation (utils.py:100-108).
// loss:   the loss graph of SSDVGG.build_optimizer (ssdvgg.py:380-580) and its
//         gradient w.r.t. the head output: softmax-CE, smooth-L1, per-image 3:1
//         hard-negative mining (top_k replaced by an exact radix select of the
//         k-th largest negative CE), per-image normalisation, batch mean.
//
// One CTA per image.  All box arithmetic uses explicit round-to-nearest
// intrinsics (no FMA contraction) so the integer grid coordinates and IoU
// comparisons are bit-identical to the reference's float64 NumPy arithmetic.
#include "common.cuh"

namespace ssdb {
namespace {

constexpr int LT = 1024;          // threads per CTA
constexpr int MAX_G = 128;        // ground-truth boxes per image

struct IBox { int x0, x1, y0, y1; };

// utils.prop2abs on the 1000x1000 grid, float64, int() truncation
__device__ __forceinline__ IBox prop2abs_1000(double cx, double cy, double w, double h) {
    double hw = __ddiv_rn(__dmul_rn(w, 1000.0), 2.0);
    double hh = __ddiv_rn(__dmul_rn(h, 1000.0), 2.0);
    double px = __dmul_rn(cx, 1000.0);
    double py = __dmul_rn(cy, 1000.0);
    IBox b;
    b.x0 = (int)__dsub_rn(px, hw); b.x1 = (int)__dadd_rn(px, hw);
    b.y0 = (int)__dsub_rn(py, hh); b.y1 = (int)__dadd_rn(py, hh);
    return b;
}

// jaccard_overlap: inclusive-pixel IoU, integer-valued float64 operands, one rounded division
__device__ __forceinline__ double iou_incl(const IBox& a, const IBox& b) {
    long long area_a = (long long)(a.x1 - a.x0 + 1) * (a.y1 - a.y0 + 1);
    long long area_b = (long long)(b.x1 - b.x0 + 1) * (b.y1 - b.y0 + 1);
    int iw = min(a.x1, b.x1) - max(a.x0, b.x0) + 1; iw = iw < 0 ? 0 : iw;
    int ih = min(a.y1, b.y1) - max(a.y0, b.y0) + 1; ih = ih < 0 ? 0 : ih;
    long long inter = (long long)iw * ih;
    long long uni = area_a + area_b - inter;
    return __ddiv_rn((double)inter, (double)uni);
}

// compute_location: float64 encode, stored as float32
__device__ __forceinline__ void encode_loc(const double* gt5, const double* anc4, float* o) {
    o[0] = (float)__dmul_rn(__ddiv_rn(__dsub_rn(gt5[1], anc4[0]), anc4[2]), 10.0);
    o[1] = (float)__dmul_rn(__ddiv_rn(__dsub_rn(gt5[2], anc4[1]), anc4[3]), 10.0);
    o[2] = (float)__dmul_rn(log(__ddiv_rn(gt5[3], anc4[2])), 5.0);
    o[3] = (float)__dmul_rn(log(__ddiv_rn(gt5[4], anc4[3])), 5.0);
}

struct MatchShared {
    IBox gt[MAX_G];
    double best_iou[MAX_G];
    int best_idx[MAX_G];
    double red_iou[LT / 32];
    int red_idx[LT / 32];
};

// Block-cooperative matching for one image.  On return `owner(a)` is available through
// match_one(); best_idx/best_iou hold each GT's arg-max anchor (first maximum).
__device__ void match_prepare(MatchShared& ms, const double* gt, int G, const double* anchors, int A) {
    const int tid = threadIdx.x;
    if (tid < G) ms.gt[tid] = prop2abs_1000(gt[tid * 5 + 1], gt[tid * 5 + 2], gt[tid * 5 + 3], gt[tid * 5 + 4]);
    __syncthreads();
    for (int g = 0; g < G; ++g) {
        double bi = -1.0; int bx = 0x7fffffff;
        IBox gb = ms.gt[g];
        for (int a = tid; a < A; a += LT) {
            IBox ab = prop2abs_1000(anchors[a * 4 + 0], anchors[a * 4 + 1], anchors[a * 4 + 2], anchors[a * 4 + 3]);
            double v = iou_incl(gb, ab);
            if (v > bi) { bi = v; bx = a; }     // ascending a per thread: first maximum kept
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            double oi = __shfl_xor_sync(0xffffffffu, bi, o);
            int ox = __shfl_xor_sync(0xffffffffu, bx, o);
            if (oi > bi || (oi == bi && ox < bx)) { bi = oi; bx = ox; }
        }
        if ((tid & 31) == 0) { ms.red_iou[tid >> 5] = bi; ms.red_idx[tid >> 5] = bx; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < LT / 32; ++w)
                if (ms.red_iou[w] > bi || (ms.red_iou[w] == bi && ms.red_idx[w] < bx)) { bi = ms.red_iou[w]; bx = ms.red_idx[w]; }
            ms.best_iou[g] = bi; ms.best_idx[g] = bx;
        }
        __syncthreads();
    }
}

// owner GT of anchor a (-1 = background): pass 1 over all IoU > 0.5 (strictly higher wins,
// earlier GT keeps ties), then pass 2 with a fresh score table over the GTs whose arg-max is a.
__device__ __forceinline__ int match_one(const MatchShared& ms, int G, const double* anchors, int a) {
    IBox ab = prop2abs_1000(anchors[a * 4 + 0], anchors[a * 4 + 1], anchors[a * 4 + 2], anchors[a * 4 + 3]);
    int owner = -1; double score = -1.0;
    for (int g = 0; g < G; ++g) {
        double v = iou_incl(ms.gt[g], ab);
        if (v > 0.5 && v > score) { score = v; owner = g; }
    }
    bool any2 = false; double score2 = -1.0;
    for (int g = 0; g < G; ++g) {
        if (ms.best_idx[g] != a || !(ms.best_iou[g] > 0.5)) continue;
        if (any2 && score2 >= ms.best_iou[g]) continue;
        any2 = true; score2 = ms.best_iou[g]; owner = g;
    }
    return owner;
}

__global__ void __launch_bounds__(LT) match_kernel(const double* __restrict__ gt, const int* __restrict__ gt_count, int G,
                                                    const double* __restrict__ anchors, int A, int C,
                                                    int* __restrict__ match_out, float* __restrict__ labels_out) {
    __shared__ MatchShared ms;
    const int b = blockIdx.x;
    const int g_n = min(gt_count[b], G);
    const double* gtb = gt + (long long)b * G * 5;
    match_prepare(ms, gtb, g_n, anchors, A);
    const int V = C + 5;
    for (int a = threadIdx.x; a < A; a += LT) {
        int owner = match_one(ms, g_n, anchors, a);
        if (match_out) match_out[(long long)b * A + a] = owner;
        if (labels_out) {
            float* row = labels_out + ((long long)b * A + a) * V;
            for (int c = 0; c < V; ++c) row[c] = 0.f;
            if (owner < 0) row[C] = 1.f;
            else {
                row[(int)gtb[owner * 5]] = 1.f;
                encode_loc(gtb + owner * 5, anchors + a * 4, row + C + 1);
            }
        }
    }
}

__device__ __forceinline__ unsigned int order_key(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float block_sum(float v, float* red) {
    __syncthreads();
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = threadIdx.x < LT / 32 ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) red[0] = t;
    }
    __syncthreads();
    t = red[0];
    return t;
}

__device__ __forceinline__ int block_sum_int(int v, int* red) {
    __syncthreads();
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < LT / 32 ? red[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) red[0] = t;
    }
    __syncthreads();
    t = red[0];
    return t;
}

constexpr int MAXV = 32;   // C + 5 <= 32

// dynamic shared: float ce[A]; signed char kind[A] (0 positive, 1 negative, 2 selected negative); signed char own[A]
template <bool GT_MODE>
__global__ void __launch_bounds__(LT) multibox_loss_kernel(
    const float* __restrict__ output, const float* __restrict__ labels, const double* __restrict__ gt,
    const int* __restrict__ gt_count, int G, const double* __restrict__ anchors, int B, int A, int C, float grad_scale,
    float* __restrict__ losses_out, float* __restrict__ grad_out, float* __restrict__ result_out, int* __restrict__ match_out,
    float* __restrict__ per_image, unsigned int* __restrict__ counter) {
    extern __shared__ __align__(16) unsigned char dyn[];
    float* ce = reinterpret_cast<float*>(dyn);
    signed char* kind = reinterpret_cast<signed char*>(ce + A);
    signed char* own = kind + A;
    __shared__ MatchShared ms;
    __shared__ float redf[LT / 32];
    __shared__ int redi[LT / 32];
    __shared__ int hist[256];
    __shared__ int sel_bin, sel_rem;
    __shared__ int scan[LT];

    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    const int V = C + 5, NC = C + 1;
    const float* outb = output + (long long)b * A * V;
    const float* labb = GT_MODE ? nullptr : labels + (long long)b * A * V;
    const double* gtb = GT_MODE ? gt + (long long)b * G * 5 : nullptr;
    int g_n = 0;
    if (GT_MODE) { g_n = min(gt_count[b], G); match_prepare(ms, gtb, g_n, anchors, A); }

    // ---- phase 2: per-anchor CE / smooth-L1, positive sums ----
    float pos_sum = 0.f, loc_sum = 0.f; int pos_cnt = 0, neg_cnt = 0;
    for (int a = tid; a < A; a += LT) {
        float z[MAXV];
        const float* zr = outb + (long long)a * V;
#pragma unroll 5
        for (int c = 0; c < V; ++c) z[c] = zr[c];
        float m = z[0];
        for (int c = 1; c < NC; ++c) m = fmaxf(m, z[c]);
        float s = 0.f;
        for (int c = 0; c < NC; ++c) s += expf(z[c] - m);
        float lse = m + logf(s);
        float cev, l1 = 0.f; bool pos;
        if (GT_MODE) {
            int owner = match_one(ms, g_n, anchors, a);
            own[a] = (signed char)owner;
            if (match_out) match_out[(long long)b * A + a] = owner;
            pos = owner >= 0;
            int cls = pos ? (int)gtb[owner * 5] : C;
            cev = lse - z[cls];
            if (pos) {
                float t[4]; encode_loc(gtb + owner * 5, anchors + a * 4, t);
#pragma unroll
                for (int i = 0; i < 4; ++i) { float d = z[NC + i] - t[i]; float ad = fabsf(d); l1 += ad < 1.f ? 0.5f * d * d : ad - 0.5f; }
            }
        } else {
            const float* yr = labb + (long long)a * V;
            float dot = 0.f, sy = 0.f;
            for (int c = 0; c < NC; ++c) { float y = yr[c]; dot += y * z[c]; sy += y; }
            cev = sy * lse - dot;
            pos = yr[C] == 0.f;
            if (pos) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { float d = z[NC + i] - yr[NC + i]; float ad = fabsf(d); l1 += ad < 1.f ? 0.5f * d * d : ad - 0.5f; }
            }
        }
        ce[a] = cev;
        kind[a] = pos ? 0 : 1;
        if (pos) { pos_sum += cev; loc_sum += l1; ++pos_cnt; } else ++neg_cnt;
    }
    pos_sum = block_sum(pos_sum, redf);
    loc_sum = block_sum(loc_sum, redf);
    pos_cnt = block_sum_int(pos_cnt, redi);
    neg_cnt = block_sum_int(neg_cnt, redi);

    // ---- phase 3: k-th largest negative CE (radix select), sum of the top k ----
    const int k = min(neg_cnt, 3 * pos_cnt);
    float neg_sum = 0.f;
    if (k > 0) {
        unsigned int prefix = 0, mask = 0; int remaining = k;
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            for (int i = tid; i < 256; i += LT) hist[i] = 0;
            __syncthreads();
            for (int a = tid; a < A; a += LT) {
                if (kind[a] != 1) continue;
                unsigned int key = order_key(ce[a]);
                if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1);
            }
            __syncthreads();
            if (tid == 0) {
                int cum = 0, bin = 255;
                for (; bin >= 0; --bin) { if (cum + hist[bin] >= remaining) break; cum += hist[bin]; }
                sel_bin = bin; sel_rem = remaining - cum;
            }
            __syncthreads();
            prefix |= ((unsigned int)sel_bin) << shift; mask |= 255u << shift; remaining = sel_rem;
            __syncthreads();
        }
        // prefix = key of the k-th largest; `remaining` ties (key == prefix) are taken, lowest index first
        const int per = (A + LT - 1) / LT;
        const int a_lo = tid * per, a_hi = min(A, a_lo + per);
        int ties = 0;
        for (int a = a_lo; a < a_hi; ++a) if (kind[a] == 1 && order_key(ce[a]) == prefix) ++ties;
        scan[tid] = ties;
        __syncthreads();
        // exclusive scan of scan[] (Hillis-Steele over LT entries)
        for (int o = 1; o < LT; o <<= 1) {
            int v = tid >= o ? scan[tid - o] : 0;
            __syncthreads();
            scan[tid] += v;
            __syncthreads();
        }
        int rank = scan[tid] - ties;
        float part = 0.f;
        for (int a = a_lo; a < a_hi; ++a) {
            if (kind[a] != 1) continue;
            unsigned int key = order_key(ce[a]);
            bool take = key > prefix;
            if (key == prefix) { take = rank < remaining; ++rank; }
            if (take) { kind[a] = 2; part += ce[a]; }
        }
        neg_sum = block_sum(part, redf);
    }
    __syncthreads();

    // ---- per-image losses ----
    const float inv_pos = pos_cnt > 0 ? 1.f / (float)pos_cnt : 0.f;
    if (tid == 0) {
        per_image[b * 2 + 0] = pos_cnt > 0 ? (pos_sum + neg_sum) / (float)pos_cnt : 0.f;
        per_image[b * 2 + 1] = pos_cnt > 0 ? loc_sum / (float)pos_cnt : 0.f;
    }

    // ---- phase 4: gradient w.r.t. the head output, and net.result ----
    if (grad_out || result_out) {
        const float gs = grad_scale * inv_pos / (float)B;
        for (int a = tid; a < A; a += LT) {
            float z[MAXV];
            const float* zr = outb + (long long)a * V;
#pragma unroll 5
            for (int c = 0; c < V; ++c) z[c] = zr[c];
            float m = z[0];
            for (int c = 1; c < NC; ++c) m = fmaxf(m, z[c]);
            float s = 0.f;
            for (int c = 0; c < NC; ++c) s += expf(z[c] - m);
            float inv = 1.f / s;
            float p[MAXV];
            for (int c = 0; c < NC; ++c) p[c] = expf(z[c] - m) * inv;
            if (result_out) {
                float* rr = result_out + ((long long)b * A + a) * V;
                for (int c = 0; c < NC; ++c) rr[c] = p[c];
                for (int c = NC; c < V; ++c) rr[c] = z[c];
            }
            if (grad_out) {
                float* gr = grad_out + ((long long)b * A + a) * V;
                const int kd = kind[a];
                if (kd == 1) { for (int c = 0; c < V; ++c) gr[c] = 0.f; continue; }
                if (GT_MODE) {
                    int owner = own[a];
                    int cls = owner >= 0 ? (int)gtb[owner * 5] : C;
                    for (int c = 0; c < NC; ++c) gr[c] = (p[c] - (c == cls ? 1.f : 0.f)) * gs;
                    if (owner >= 0) {
                        float t[4]; encode_loc(gtb + owner * 5, anchors + a * 4, t);
                        for (int i = 0; i < 4; ++i) { float d = z[NC + i] - t[i]; gr[NC + i] = (fabsf(d) < 1.f ? d : (d > 0.f ? 1.f : -1.f)) * gs; }
                    } else for (int i = 0; i < 4; ++i) gr[NC + i] = 0.f;
                } else {
                    const float* yr = labb + (long long)a * V;
                    float sy = 0.f;
                    for (int c = 0; c < NC; ++c) sy += yr[c];
                    for (int c = 0; c < NC; ++c) gr[c] = (sy * p[c] - yr[c]) * gs;
                    if (kd == 0) { for (int i = 0; i < 4; ++i) { float d = z[NC + i] - yr[NC + i]; gr[NC + i] = (fabsf(d) < 1.f ? d : (d > 0.f ? 1.f : -1.f)) * gs; } }
                    else for (int i = 0; i < 4; ++i) gr[NC + i] = 0.f;
                }
            }
        }
    }

    // ---- batch mean by the last CTA, summed in image order (deterministic) ----
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned int done = atomicAdd(counter, 1u);
        if (done == (unsigned int)B - 1) {
            __threadfence();
            float c = 0.f, l = 0.f;
            const volatile float* pi = per_image;
            for (int i = 0; i < B; ++i) { c += pi[i * 2]; l += pi[i * 2 + 1]; }
            losses_out[0] = c / (float)B;
            losses_out[1] = l / (float)B;
            *counter = 0;
        }
    }
}

}  // namespace

int match_anchors_launch(const double* gt, const int* gt_count, int B, int G, const double* anchors_prop, int A, int C,
                         int* match_out, float* labels_out, cudaStream_t st) {
    SSDB_REQUIRE(G >= 1 && G <= MAX_G, "G must be in [1,128]");
    SSDB_REQUIRE(B >= 1 && A >= 1 && C >= 1, "bad sizes");
    match_kernel<<<B, LT, 0, st>>>(gt, gt_count, G, anchors_prop, A, C, match_out, labels_out);
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

int multibox_loss_launch(const float* output, const float* labels, const double* gt, const int* gt_count, int G,
                         const double* anchors_prop, int B, int A, int C, float grad_scale, float* losses_out,
                         float* grad_out, float* result_out, int* match_out, float* per_image_ws,
                         unsigned int* counter_ws, cudaStream_t st) {
    SSDB_REQUIRE(C + 5 <= MAXV, "too many classes");
    SSDB_REQUIRE(B >= 1 && A >= 1, "bad sizes");
    size_t sh = (size_t)A * 4 + (size_t)A * 2 + 16;
    if (labels) {
        static bool attr0 = false;
        if (!attr0) { SSDB_CUDA(cudaFuncSetAttribute(multibox_loss_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr0 = true; }
        multibox_loss_kernel<false><<<B, LT, sh, st>>>(output, labels, nullptr, nullptr, 0, nullptr, B, A, C, grad_scale,
                                                        losses_out, grad_out, result_out, nullptr, per_image_ws, counter_ws);
    } else {
        SSDB_REQUIRE(gt && gt_count && anchors_prop && G >= 1 && G <= MAX_G, "ground truth required");
        static bool attr1 = false;
        if (!attr1) { SSDB_CUDA(cudaFuncSetAttribute(multibox_loss_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr1 = true; }
        multibox_loss_kernel<true><<<B, LT, sh, st>>>(output, nullptr, gt, gt_count, G, anchors_prop, B, A, C, grad_scale,
                                                       losses_out, grad_out, result_out, match_out, per_image_ws, counter_ws);
    }
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

}  // namespace ssdb
