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
// Two implementations of the loss live here:
//   v2 (default): three streaming kernels.  `loss_rows_kernel` (grid = anchor tiles x images) pulls a
//       contiguous tile of 256 anchor rows of the head output (and of the dense labels) into shared
//       memory with one TMA bulk copy each, computes CE / smooth-L1 / softmax thread-per-row out of
//       shared memory (row stride 25 words: conflict free), writes `net.result` back with a bulk
//       store and leaves per-anchor CE + kind and per-tile partial sums in a workspace;
//       `loss_select_kernel` (one CTA per image) does the exact radix select of the k-th largest
//       negative CE on that 44 KB workspace slice; `loss_grad_kernel` (tiles x images) writes the
//       gradient tile by tile (zeros for unselected negatives, recomputed rows for the rest).
//       HBM traffic = the algorithmic bytes: read output + labels once, write result + gradient once.
//   v1 (SSDB_LOSS=v1): the original one-CTA-per-image kernel, kept as the on-device cross-check.
// All box arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction) so the integer grid
// coordinates and IoU comparisons are bit-identical to the reference's float64 NumPy arithmetic.
#include <cstdlib>

#include "bulk.cuh"
#include "common.cuh"
#include "select.cuh"

namespace ssdb {
namespace {

constexpr int LT = 1024;          // threads per CTA
constexpr int MAX_G = 128;        // ground-truth boxes per image

struct IBox { int x0, x1, y0, y1; };

// label id of a ground-truth row as an index: ids outside [0, C) (or NaN) are treated as background -- the host entry points
// reject them (SSDB_EINVAL, like the reference's IndexError at transforms.py:107); for device-pointer callers this keeps
// every access in bounds
__device__ __forceinline__ int gt_class(const double* row, int C) {
    const double id = row[0];
    return (id >= 0.0 && id < (double)C) ? (int)id : C;
}

// utils.prop2abs on the 1000x1000 grid, float64, int() truncation
__device__ __forceinline__ IBox prop2abs_1000(double cx, double cy, double w, double h) {
    double hw = __dmul_rn(__dmul_rn(w, 1000.0), 0.5);          // x * 0.5 == x / 2 exactly; float64 division is slow
    double hh = __dmul_rn(__dmul_rn(h, 1000.0), 0.5);
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
    if (inter == 0) return 0.0;                       // 0 / uni == 0.0 exactly: skip the float64 division for disjoint boxes
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
__device__ __forceinline__ int match_one_box(const MatchShared& ms, int G, const IBox& ab, int a) {
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
__device__ __forceinline__ int match_one(const MatchShared& ms, int G, const double* anchors, int a) {
    IBox ab = prop2abs_1000(anchors[a * 4 + 0], anchors[a * 4 + 1], anchors[a * 4 + 2], anchors[a * 4 + 3]);
    return match_one_box(ms, G, ab, a);
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
                row[gt_class(gtb + owner * 5, C)] = 1.f;
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
    extern __shared__ __align__(128) unsigned char dyn[];
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
            int cls = pos ? gt_class(gtb + owner * 5, C) : C;
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
                    int cls = owner >= 0 ? gt_class(gtb + owner * 5, C) : C;
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


// ---- exact integer form of the matching (v2) ----
// Coordinates live on the 1000x1000 grid, so intersection I and union U are integers below 2^21.  Two IoUs compare
// exactly by cross-multiplication (products < 2^42), and I/U > 0.5 <=> 2I > U: distinct rationals with such small
// denominators are more than 2^-42 apart, far more than the rounding of the reference's float64 division, so every
// comparison below decides exactly like the reference's `iou > threshold` / `score > stored` on rounded float64 values
//  -- without a single float64 instruction.
struct IU { int i, u; };
__device__ __forceinline__ IU iou_int(const IBox& a, const IBox& b) {
    const int area_a = (a.x1 - a.x0 + 1) * (a.y1 - a.y0 + 1), area_b = (b.x1 - b.x0 + 1) * (b.y1 - b.y0 + 1);
    int iw = min(a.x1, b.x1) - max(a.x0, b.x0) + 1; iw = iw < 0 ? 0 : iw;
    int ih = min(a.y1, b.y1) - max(a.y0, b.y0) + 1; ih = ih < 0 ? 0 : ih;
    IU r; r.i = iw * ih; r.u = area_a + area_b - r.i;
    return r;
}
__device__ __forceinline__ bool iu_greater(const IU& a, const IU& b) {      // a.i/a.u > b.i/b.u   (u > 0)
    return (long long)a.i * b.u > (long long)b.i * a.u;
}

struct MatchSharedInt {
    IBox gt[MAX_G];
    IU best[MAX_G];          // IoU of each GT with its arg-max anchor
    int best_idx[MAX_G];
};

// owner GT of anchor a (-1 = background), same two-pass rule as match_one_box
__device__ __forceinline__ int match_one_int(const MatchSharedInt& ms, int G, const IBox& ab, int a) {
    int owner = -1; IU score{-1, 1};
    for (int g = 0; g < G; ++g) {
        const IU v = iou_int(ms.gt[g], ab);
        if (2 * v.i > v.u && iu_greater(v, score)) { score = v; owner = g; }
    }
    bool any2 = false; IU score2{-1, 1};
    for (int g = 0; g < G; ++g) {
        const IU bg = ms.best[g];
        if (ms.best_idx[g] != a || !(2 * bg.i > bg.u)) continue;
        if (any2 && !iu_greater(bg, score2)) continue;
        any2 = true; score2 = bg; owner = g;
    }
    return owner;
}

// =====================================================================================
// v2: tiled streaming implementation
// =====================================================================================
constexpr int RT = 256;            // anchor rows per tile = threads per CTA of the streaming kernels

struct TilePart { float pos_sum, loc_sum; int pos_cnt, neg_cnt; };

struct LossWs {
    float* ce;               // [B*A]  per-anchor cross entropy
    unsigned char* kind;     // [B*A]  0 positive, 1 negative, 2 selected negative
    signed char* own;        // [B*A]  owner GT (fused-match mode)
    TilePart* part;          // [B*S]
    float* gs;               // [B]    gradient scale of the image: grad_scale / (B * pos_num)
    float* per_image;        // [B*2]
    unsigned int* counter;   // [1]    self-resetting
    IU* best_iou;            // [B*MAX_G] (intersection, union) of each GT with its arg-max anchor
    int* best_idx;           // [B*MAX_G]
    IBox* anc_abs;           // [A]
};

__host__ __device__ inline size_t ws_align(size_t v) { return (v + 255) / 256 * 256; }

// anchors on the 1000x1000 grid, once per launch instead of once per (GT, anchor) pair
__global__ void __launch_bounds__(256) anchor_abs_kernel(const double* __restrict__ anchors, int A, IBox* __restrict__ out) {
    int a = blockIdx.x * 256 + threadIdx.x;
    if (a < A) out[a] = prop2abs_1000(anchors[a * 4 + 0], anchors[a * 4 + 1], anchors[a * 4 + 2], anchors[a * 4 + 3]);
}

// arg-max anchor of every GT box (first maximum = lowest anchor index), one CTA per (GT, image); integer-exact
__global__ void __launch_bounds__(RT) match_best_kernel(const double* __restrict__ gt, const int* __restrict__ gt_count, int G,
                                                         const IBox* __restrict__ anc_abs, int A, IU* __restrict__ best_iou,
                                                         int* __restrict__ best_idx) {
    const int g = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    if (g >= min(gt_count[b], G)) return;
    __shared__ IU red_iou[RT / 32];
    __shared__ int red_idx[RT / 32];
    const double* r = gt + ((long long)b * G + g) * 5;
    const IBox gb = prop2abs_1000(r[1], r[2], r[3], r[4]);
    IU bi{-1, 1}; int bx = 0x7fffffff;
    for (int a = tid; a < A; a += RT) {
        const IU v = iou_int(gb, anc_abs[a]);
        if (iu_greater(v, bi)) { bi = v; bx = a; }        // ascending a per thread: first maximum kept
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        IU oi; oi.i = __shfl_xor_sync(0xffffffffu, bi.i, o); oi.u = __shfl_xor_sync(0xffffffffu, bi.u, o);
        const int ox = __shfl_xor_sync(0xffffffffu, bx, o);
        if (iu_greater(oi, bi) || (!iu_greater(bi, oi) && ox < bx)) { bi = oi; bx = ox; }
    }
    if ((tid & 31) == 0) { red_iou[tid >> 5] = bi; red_idx[tid >> 5] = bx; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < RT / 32; ++w) {
            const IU oi = red_iou[w]; const int ox = red_idx[w];
            if (iu_greater(oi, bi) || (!iu_greater(bi, oi) && ox < bx)) { bi = oi; bx = ox; }
        }
        best_iou[b * MAX_G + g] = bi; best_idx[b * MAX_G + g] = bx;
    }
}

using bulk::tile_load_wait;

// VT = compile-time row width (C + 5), 0 = generic (<= MAXV)
template <bool GT_MODE, int VT>
__global__ void __launch_bounds__(RT) loss_rows_kernel(
    const float* __restrict__ output, const float* __restrict__ labels, const double* __restrict__ gt,
    const int* __restrict__ gt_count, int G, const double* __restrict__ anchors, int A, int C, int S, int use_bulk, LossWs ws,
    float* __restrict__ result_out, int* __restrict__ match_out) {
    extern __shared__ __align__(128) unsigned char dyn[];
    constexpr int VMAX = VT ? VT : MAXV;
    const int V = VT ? VT : C + 5, NC = V - 4;
    float* zt = reinterpret_cast<float*>(dyn);        // [RT*V] head output tile, becomes the result tile
    float* yt = zt + RT * V;                          // [RT*V] dense label tile (not in fused-match mode)
    __shared__ __align__(8) unsigned long long bar;
    __shared__ MatchSharedInt ms;
    __shared__ float red[RT / 32][4];

    const int tid = threadIdx.x, s = blockIdx.x, b = blockIdx.y;
    const int a0 = s * RT, rows = min(RT, A - a0), nfl = rows * V;
    const long long tile_off = ((long long)b * A + a0) * V;
    bulk::tile_load<RT>(zt, output + tile_off, yt, GT_MODE ? nullptr : labels + tile_off, nfl, use_bulk, &bar);
    const double* gtb = GT_MODE ? gt + (long long)b * G * 5 : nullptr;
    int g_n = 0;
    if (GT_MODE) {
        g_n = min(gt_count[b], G);
        if (tid < g_n) {
            ms.gt[tid] = prop2abs_1000(gtb[tid * 5 + 1], gtb[tid * 5 + 2], gtb[tid * 5 + 3], gtb[tid * 5 + 4]);
            ms.best[tid] = ws.best_iou[b * MAX_G + tid];
            ms.best_idx[tid] = ws.best_idx[b * MAX_G + tid];
        }
    }
    tile_load_wait(use_bulk, &bar);

    float pos_sum = 0.f, loc_sum = 0.f; int pos_cnt = 0, neg_cnt = 0;
    if (tid < rows) {
        const int a = a0 + tid;
        float* zr = zt + tid * V;
        float z[VMAX];
#pragma unroll
        for (int c = 0; c < VMAX; ++c) z[c] = c < V ? zr[c] : 0.f;
        float m = z[0];
#pragma unroll
        for (int c = 1; c < VMAX; ++c) if (c < NC) m = fmaxf(m, z[c]);
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < VMAX; ++c) if (c < NC) sum += expf(z[c] - m);
        const float lse = m + logf(sum);
        const float inv = 1.f / sum;
        float cev, l1 = 0.f; bool pos;
        if (GT_MODE) {
            const int owner = match_one_int(ms, g_n, ws.anc_abs[a], a);
            ws.own[(long long)b * A + a] = (signed char)owner;
            if (match_out) match_out[(long long)b * A + a] = owner;
            pos = owner >= 0;
            const int cls = pos ? gt_class(gtb + owner * 5, C) : C;
            cev = lse - zr[cls];
            if (pos) {
                float t[4]; encode_loc(gtb + owner * 5, anchors + a * 4, t);
#pragma unroll
                for (int i = 0; i < 4; ++i) { float d = zr[NC + i] - t[i]; float ad = fabsf(d); l1 += ad < 1.f ? 0.5f * d * d : ad - 0.5f; }
            }
        } else {
            const float* yr = yt + tid * V;
            float dot = 0.f, sy = 0.f;
#pragma unroll
            for (int c = 0; c < VMAX; ++c) if (c < NC) { float y = yr[c]; dot += y * z[c]; sy += y; }
            cev = sy * lse - dot;
            pos = yr[C] == 0.f;
            if (pos) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { float d = zr[NC + i] - yr[NC + i]; float ad = fabsf(d); l1 += ad < 1.f ? 0.5f * d * d : ad - 0.5f; }
            }
        }
        ws.ce[(long long)b * A + a] = cev;
        ws.kind[(long long)b * A + a] = pos ? 0 : 1;
        if (pos) { pos_sum = cev; loc_sum = l1; pos_cnt = 1; } else neg_cnt = 1;
        // the tile becomes net.result: softmax over the class columns, offsets unchanged (ssdvgg.py:365-372)
#pragma unroll
        for (int c = 0; c < VMAX; ++c) if (c < NC) zr[c] = expf(z[c] - m) * inv;
    }
    // per-tile partial sums (fixed order: deterministic)
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        pos_sum += __shfl_xor_sync(0xffffffffu, pos_sum, o); loc_sum += __shfl_xor_sync(0xffffffffu, loc_sum, o);
        pos_cnt += __shfl_xor_sync(0xffffffffu, pos_cnt, o); neg_cnt += __shfl_xor_sync(0xffffffffu, neg_cnt, o);
    }
    if ((tid & 31) == 0) { red[tid >> 5][0] = pos_sum; red[tid >> 5][1] = loc_sum; red[tid >> 5][2] = __int_as_float(pos_cnt); red[tid >> 5][3] = __int_as_float(neg_cnt); }
    if (result_out) bulk::tile_store<RT>(result_out + tile_off, zt, nfl, use_bulk); else __syncthreads();
    if (tid == 0) {
        TilePart tp{0.f, 0.f, 0, 0};
        for (int w = 0; w < RT / 32; ++w) { tp.pos_sum += red[w][0]; tp.loc_sum += red[w][1]; tp.pos_cnt += __float_as_int(red[w][2]); tp.neg_cnt += __float_as_int(red[w][3]); }
        ws.part[b * S + s] = tp;
    }
}

struct NegKey {          // participants of the hard-negative selection: negatives, ordered by cross entropy
    const float* ce; const signed char* kind;
    __device__ __forceinline__ bool operator()(int a, unsigned& key) const { key = order_key(ce[a]); return kind[a] == 1; }
};

// one CTA per image: exact k-th largest negative CE, marks the selected negatives, per-image losses, batch mean
#define SSDB_TRACE_PT(k) do { if (trace && blockIdx.x == 0 && threadIdx.x == 0) trace[k] = clock64(); } while (0)
__global__ void __launch_bounds__(LT) loss_select_kernel(LossWs ws, int B, int A, int S, float grad_scale, float* __restrict__ losses_out,
                                                         long long* __restrict__ trace) {
    extern __shared__ __align__(128) unsigned char dyn[];
    float* ce = reinterpret_cast<float*>(dyn);
    signed char* kind = reinterpret_cast<signed char*>(ce + A);
    __shared__ float redf[LT / 32];
    __shared__ int whist[(LT / 32) * 256];
    __shared__ int tot_bins[256];
    __shared__ int sel[2];
    __shared__ int wsum[32];
    __shared__ TilePart tot;
    const int tid = threadIdx.x, b = blockIdx.x;
    const float* gce = ws.ce + (long long)b * A;
    unsigned char* gkind = ws.kind + (long long)b * A;
    SSDB_TRACE_PT(0);
    for (int a = tid; a < A; a += LT) { ce[a] = gce[a]; kind[a] = (signed char)gkind[a]; }
    if (tid < 32) {
        // per-tile partial sums in tile order (lane-strided, then a fixed shuffle tree: deterministic)
        TilePart t{0.f, 0.f, 0, 0};
        for (int s = tid; s < S; s += 32) { const TilePart p = ws.part[b * S + s]; t.pos_sum += p.pos_sum; t.loc_sum += p.loc_sum; t.pos_cnt += p.pos_cnt; t.neg_cnt += p.neg_cnt; }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            t.pos_sum += __shfl_xor_sync(0xffffffffu, t.pos_sum, o); t.loc_sum += __shfl_xor_sync(0xffffffffu, t.loc_sum, o);
            t.pos_cnt += __shfl_xor_sync(0xffffffffu, t.pos_cnt, o); t.neg_cnt += __shfl_xor_sync(0xffffffffu, t.neg_cnt, o);
        }
        if (tid == 0) tot = t;
    }
    __syncthreads();
    const float pos_sum = tot.pos_sum, loc_sum = tot.loc_sum; const int pos_cnt = tot.pos_cnt, neg_cnt = tot.neg_cnt;
    const int k = min(neg_cnt, 3 * pos_cnt);
    float neg_sum = 0.f;
    SSDB_TRACE_PT(1);
    if (k > 0) {
        unsigned int prefix; int remaining;
        radix_select_kth<LT>(A, k, NegKey{ce, kind}, whist, tot_bins, sel, prefix, remaining);
        SSDB_TRACE_PT(2);
        // prefix = key of the k-th largest; `remaining` ties (key == prefix) are taken, lowest index first (tf.nn.top_k order)
        const int per = (A + LT - 1) / LT;
        const int a_lo = tid * per, a_hi = min(A, a_lo + per);
        int ties = 0;
        for (int a = a_lo; a < a_hi; ++a) if (kind[a] == 1 && order_key(ce[a]) == prefix) ++ties;
        int rank = block_excl_scan<LT>(ties, wsum);
        SSDB_TRACE_PT(3);
        float part = 0.f;
        for (int a = a_lo; a < a_hi; ++a) {
            if (kind[a] != 1) continue;
            unsigned int key = order_key(ce[a]);
            bool take = key > prefix;
            if (key == prefix) { take = rank < remaining; ++rank; }
            if (take) { gkind[a] = 2; part += ce[a]; }
        }
        neg_sum = block_sum(part, redf);
    }
    SSDB_TRACE_PT(4);
    if (tid == 0) {
        ws.per_image[b * 2 + 0] = pos_cnt > 0 ? (pos_sum + neg_sum) / (float)pos_cnt : 0.f;
        ws.per_image[b * 2 + 1] = pos_cnt > 0 ? loc_sum / (float)pos_cnt : 0.f;
        ws.gs[b] = grad_scale * (pos_cnt > 0 ? 1.f / (float)pos_cnt : 0.f) / (float)B;
    }
    // ---- batch mean by the last CTA, summed in image order (deterministic) ----
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned int done = atomicAdd(ws.counter, 1u);
        if (done == (unsigned int)B - 1) {
            __threadfence();
            float c = 0.f, l = 0.f;
            const volatile float* pi = ws.per_image;
            for (int i = 0; i < B; ++i) { c += pi[i * 2]; l += pi[i * 2 + 1]; }
            losses_out[0] = c / (float)B;
            losses_out[1] = l / (float)B;
            *ws.counter = 0;
        }
    }
    SSDB_TRACE_PT(5);
}

// gradient w.r.t. the head output, tile by tile: zeros for unselected negatives, recomputed rows for the rest
template <bool GT_MODE, int VT>
__global__ void __launch_bounds__(RT) loss_grad_kernel(
    const float* __restrict__ output, const float* __restrict__ labels, const double* __restrict__ gt, int G,
    const double* __restrict__ anchors, int A, int C, int use_bulk, LossWs ws, float* __restrict__ grad_out) {
    extern __shared__ __align__(128) unsigned char dyn[];
    constexpr int VMAX = VT ? VT : MAXV;
    const int V = VT ? VT : C + 5, NC = V - 4;
    float* gtile = reinterpret_cast<float*>(dyn);     // [RT*V]
    const int tid = threadIdx.x, s = blockIdx.x, b = blockIdx.y;
    const int a0 = s * RT, rows = min(RT, A - a0), nfl = rows * V;
    const long long tile_off = ((long long)b * A + a0) * V;
    if (tid < rows) {
        const int a = a0 + tid;
        float* gr = gtile + tid * V;
        const int kd = ws.kind[(long long)b * A + a];
        if (kd == 1) {
#pragma unroll
            for (int c = 0; c < VMAX; ++c) if (c < V) gr[c] = 0.f;
        } else {
            const float gs = ws.gs[b];
            const float* zr = output + tile_off + (long long)tid * V;
            float z[VMAX];
#pragma unroll
            for (int c = 0; c < VMAX; ++c) z[c] = c < V ? zr[c] : 0.f;
            float m = z[0];
#pragma unroll
            for (int c = 1; c < VMAX; ++c) if (c < NC) m = fmaxf(m, z[c]);
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < VMAX; ++c) if (c < NC) sum += expf(z[c] - m);
            const float inv = 1.f / sum;
            if (GT_MODE) {
                const int owner = ws.own[(long long)b * A + a];
                const double* gtb = gt + (long long)b * G * 5;
                const int cls = owner >= 0 ? gt_class(gtb + owner * 5, C) : C;
#pragma unroll
                for (int c = 0; c < VMAX; ++c) if (c < NC) gr[c] = (expf(z[c] - m) * inv - (c == cls ? 1.f : 0.f)) * gs;
                if (owner >= 0) {
                    float t[4]; encode_loc(gtb + owner * 5, anchors + a * 4, t);
#pragma unroll
                    for (int i = 0; i < 4; ++i) { float d = zr[NC + i] - t[i]; gr[NC + i] = (fabsf(d) < 1.f ? d : (d > 0.f ? 1.f : -1.f)) * gs; }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) gr[NC + i] = 0.f;
                }
            } else {
                const float* yr = labels + tile_off + (long long)tid * V;
                float y[VMAX];
#pragma unroll
                for (int c = 0; c < VMAX; ++c) y[c] = c < V ? yr[c] : 0.f;
                float sy = 0.f;
#pragma unroll
                for (int c = 0; c < VMAX; ++c) if (c < NC) sy += y[c];
#pragma unroll
                for (int c = 0; c < VMAX; ++c) if (c < NC) gr[c] = (sy * (expf(z[c] - m) * inv) - y[c]) * gs;
                if (kd == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) { float d = zr[NC + i] - yr[NC + i]; gr[NC + i] = (fabsf(d) < 1.f ? d : (d > 0.f ? 1.f : -1.f)) * gs; }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) gr[NC + i] = 0.f;
                }
            }
        }
    }
    bulk::tile_store<RT>(grad_out + tile_off, gtile, nfl, use_bulk);
}

template <typename K>
int opt_in_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) SSDB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return SSDB_OK;
}

LossWs carve_ws(void* base, int B, int A) {
    const int S = (A + RT - 1) / RT;
    unsigned char* p = reinterpret_cast<unsigned char*>(base);
    LossWs w;
    auto take = [&](size_t bytes) { unsigned char* r = p; p += ws_align(bytes); return r; };
    w.counter = reinterpret_cast<unsigned int*>(take(256));
    w.ce = reinterpret_cast<float*>(take((size_t)B * A * 4));
    w.kind = reinterpret_cast<unsigned char*>(take((size_t)B * A));
    w.own = reinterpret_cast<signed char*>(take((size_t)B * A));
    w.part = reinterpret_cast<TilePart*>(take((size_t)B * S * sizeof(TilePart)));
    w.gs = reinterpret_cast<float*>(take((size_t)B * 4));
    w.per_image = reinterpret_cast<float*>(take((size_t)B * 8));
    w.best_iou = reinterpret_cast<IU*>(take((size_t)B * MAX_G * 8));
    w.best_idx = reinterpret_cast<int*>(take((size_t)B * MAX_G * 4));
    w.anc_abs = reinterpret_cast<IBox*>(take((size_t)A * sizeof(IBox)));
    return w;
}

template <bool GT_MODE, int VT>
int loss_v2(const float* output, const float* labels, const double* gt, const int* gt_count, int G, const double* anchors, int B, int A,
            int C, float grad_scale, float* losses_out, float* grad_out, float* result_out, int* match_out, void* ws_base, cudaStream_t st) {
    const int V = C + 5, S = (A + RT - 1) / RT;
    LossWs ws = carve_ws(ws_base, B, A);
    const bool row_ok = ((size_t)A * V * 4) % 16 == 0;
    const int bulk_rows = row_ok && bulk::aligned16(output) && (GT_MODE || bulk::aligned16(labels)) && (!result_out || bulk::aligned16(result_out));
    const int bulk_grad = row_ok && (!grad_out || bulk::aligned16(grad_out));
    if (GT_MODE) {
        anchor_abs_kernel<<<(A + 255) / 256, 256, 0, st>>>(anchors, A, ws.anc_abs);
        SSDB_LAUNCH_CHECK();
        match_best_kernel<<<dim3(G, B), RT, 0, st>>>(gt, gt_count, G, ws.anc_abs, A, ws.best_iou, ws.best_idx);
        SSDB_LAUNCH_CHECK();
    }
    const size_t sh_rows = (size_t)RT * V * 4 * (GT_MODE ? 1 : 2), sh_sel = (size_t)A * 5 + 16, sh_grad = (size_t)RT * V * 4;
    static PerDevice<size_t> sel_pd;    // per template instantiation; 227 KB per CTA minus the select kernel's static shared memory
    size_t& sel_max = sel_pd.get();
    if (!sel_max) {
        int rc = opt_in_smem(loss_rows_kernel<GT_MODE, VT>, (size_t)RT * MAXV * 8); if (rc) return rc;
        cudaFuncAttributes fa;
        SSDB_CUDA(cudaFuncGetAttributes(&fa, loss_select_kernel));
        size_t m = 232448 - fa.sharedSizeBytes;
        SSDB_CUDA(cudaFuncSetAttribute(loss_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m));
        sel_max = m;
    }
    SSDB_REQUIRE(sh_sel <= sel_max, "anchor count too large for the select kernel");
    loss_rows_kernel<GT_MODE, VT><<<dim3(S, B), RT, sh_rows, st>>>(output, labels, gt, gt_count, G, anchors, A, C, S, bulk_rows, ws,
                                                                   result_out, match_out);
    SSDB_LAUNCH_CHECK();
    static PerDevice<long long*> trace_pd;
    long long*& trace_dev = trace_pd.get();
    const bool tracing = getenv("SSDB_TRACE") != nullptr;
    if (tracing) {
        if (!trace_dev) SSDB_CUDA(cudaMalloc(&trace_dev, 16 * sizeof(long long)));
        SSDB_CUDA(cudaMemsetAsync(trace_dev, 0, 16 * sizeof(long long), st));
    }
    loss_select_kernel<<<B, LT, sh_sel, st>>>(ws, B, A, S, grad_scale, losses_out, tracing ? trace_dev : nullptr);
    SSDB_LAUNCH_CHECK();
    if (tracing) {
        long long h[16];
        SSDB_CUDA(cudaStreamSynchronize(st));
        SSDB_CUDA(cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "ssdb trace loss_select_kernel (cycles since entry):");
        for (int k = 1; k <= 5; ++k) fprintf(stderr, " p%d=%lld", k, h[k] ? h[k] - h[0] : -1);
        fprintf(stderr, "\n");
    }
    if (grad_out) {
        loss_grad_kernel<GT_MODE, VT><<<dim3(S, B), RT, sh_grad, st>>>(output, labels, gt, G, anchors, A, C, bulk_grad, ws, grad_out);
        SSDB_LAUNCH_CHECK();
    }
    return SSDB_OK;
}

bool use_v1() {       // read per call so that a test can run both implementations in one process
    const char* e = getenv("SSDB_LOSS");
    return e && e[0] == 'v' && e[1] == '1';
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

size_t multibox_loss_ws_bytes(int B, int A) {
    const int S = (A + RT - 1) / RT;
    return 256 + ws_align((size_t)B * A * 4) + 2 * ws_align((size_t)B * A) + ws_align((size_t)B * S * sizeof(TilePart)) + ws_align((size_t)B * 4) +
           ws_align((size_t)B * 8) + ws_align((size_t)B * MAX_G * 8) + ws_align((size_t)B * MAX_G * 4) + ws_align((size_t)A * sizeof(IBox));
}

// ws: multibox_loss_ws_bytes(B, A) bytes, zeroed once when allocated (the batch counter resets itself)
int multibox_loss_launch(const float* output, const float* labels, const double* gt, const int* gt_count, int G,
                         const double* anchors_prop, int B, int A, int C, float grad_scale, float* losses_out,
                         float* grad_out, float* result_out, int* match_out, void* ws, cudaStream_t st) {
    SSDB_REQUIRE(C + 5 <= MAXV, "too many classes");
    SSDB_REQUIRE(B >= 1 && A >= 1 && ws, "bad sizes");
    if (!labels) SSDB_REQUIRE(gt && gt_count && anchors_prop && G >= 1 && G <= MAX_G, "ground truth required");
    if (!use_v1()) {
        if (labels) return C == 20 ? loss_v2<false, 25>(output, labels, nullptr, nullptr, 0, nullptr, B, A, C, grad_scale, losses_out, grad_out, result_out, nullptr, ws, st)
                                   : loss_v2<false, 0>(output, labels, nullptr, nullptr, 0, nullptr, B, A, C, grad_scale, losses_out, grad_out, result_out, nullptr, ws, st);
        return C == 20 ? loss_v2<true, 25>(output, nullptr, gt, gt_count, G, anchors_prop, B, A, C, grad_scale, losses_out, grad_out, result_out, match_out, ws, st)
                       : loss_v2<true, 0>(output, nullptr, gt, gt_count, G, anchors_prop, B, A, C, grad_scale, losses_out, grad_out, result_out, match_out, ws, st);
    }
    LossWs w = carve_ws(ws, B, A);
    size_t sh = (size_t)A * 4 + (size_t)A * 2 + 16;
    if (labels) {
        static PerDevice<bool> attr0_pd;
        bool& attr0 = attr0_pd.get();
        if (!attr0) { SSDB_CUDA(cudaFuncSetAttribute(multibox_loss_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr0 = true; }
        multibox_loss_kernel<false><<<B, LT, sh, st>>>(output, labels, nullptr, nullptr, 0, nullptr, B, A, C, grad_scale,
                                                        losses_out, grad_out, result_out, nullptr, w.per_image, w.counter);
    } else {
        static PerDevice<bool> attr1_pd;
        bool& attr1 = attr1_pd.get();
        if (!attr1) { SSDB_CUDA(cudaFuncSetAttribute(multibox_loss_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr1 = true; }
        multibox_loss_kernel<true><<<B, LT, sh, st>>>(output, nullptr, gt, gt_count, G, anchors_prop, B, A, C, grad_scale,
                                                       losses_out, grad_out, result_out, match_out, w.per_image, w.counter);
    }
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

}  // namespace ssdb
