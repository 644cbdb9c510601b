// Fused decode_boxes + class-wise greedy NMS.
//
// Restates on the GPU, bit-exactly, what the reference does per image in Python:
//   ssdutils.decode_boxes (ssdutils.py:192-229): arg-max over the C object classes,
//     top-`cap` anchors by confidence, stop below the threshold,
//   ssdutils.decode_location (:182-189) + utils.normalize_box / prop2abs / abs2prop
//     (utils.py:85-135): offsets -> box on the 1000x1000 grid.  Under NumPy 2 the
//     reference computes the centre in float32 and the size in float64 and the
//     `centre - half` subtraction in float32; the same op order is spelled out with
//     round-to-nearest intrinsics here (no FMA contraction),
//   ssdutils.suppress_overlaps / non_maximum_suppression (:232-318): boxes are
//     re-quantised through prop2abs in float64 (not always a round trip), greedy NMS
//     per class with inclusive-pixel IoU > thr in float64, output grouped by class in
//     order of first appearance.
//
// Two kernels (default; SSDB_NMS=v1 keeps everything in the per-image kernel as the cross-check):
//   decode_scan_kernel (grid = anchor tiles x images): the only large traffic, the HBM read of
//     pred [B, A, C+5].  A contiguous tile of 256 anchor rows comes into shared memory with one TMA
//     bulk copy; thread-per-row arg-max out of shared memory (stride 25 words: conflict free);
//     one ordered confidence key + the arg-max class per anchor go to a 5-byte/anchor workspace.
//   decode_nms_kernel (one CTA per image): keys -> exact radix select of the cap-th largest
//     confidence (warp-private histograms, finished by counting once 16 bits are fixed) -> compaction
//     -> sort of <= cap 64-bit keys (confidence desc, anchor index asc; rank sort for cap <= 256) ->
//     decode (the reference's float32 / float64 op order) -> greedy NMS -> output order -> write.
//     For cap <= 256 the greedy pass is lazy and per class: one warp per class, only kept boxes
//     test their later class mates (no suppression matrix, no per-box CTA barrier), and the output
//     order comes from popcounts of the per-class kept masks.
//   Measured on B200 (clock64 phase trace, SSDB_TRACE=1): the full pairwise suppression matrix (a
//   divergent pair loop, ~36 K of 77 K cycles) was the cost that mattered; the lazy pass takes ~7 K.
#include <cstdlib>
#include <vector>

#include "bulk.cuh"
#include "common.cuh"
#include "select.cuh"

namespace ssdb {
namespace {

constexpr int DT = 1024;
constexpr int RT = 256;              // anchor rows per tile / threads per CTA of the scan kernel
constexpr int MAXV = 72;             // C <= 64
constexpr int BITS_P_MAX = 256;      // suppression bit matrix up to this many candidates
constexpr int SMEM_P_MAX = 1024;     // candidate lists up to this size live in shared memory
constexpr int CAND_WORDS = 11;       // cls, box[4], nms[4], conf bits, anchor

__device__ __forceinline__ unsigned int okey(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float okey_inv(unsigned int k) {
    unsigned int u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

struct DetArgs {
    const float* pred; const double* anchors; int B, A, C;
    float conf_thr; int cap; double iou_thr;
    int* dets; int* counts;
    unsigned long long* g_keys; int* g_cand; int P;   // global scratch when P > SMEM_P_MAX
    int cap_eff;
    const unsigned int* ckey_in;                      // [B*A] keys from decode_scan_kernel, or null (v1: scan in this kernel)
    const unsigned char* cls_in;                      // [B*A] arg-max class from decode_scan_kernel (with ckey_in)
    int fast;                                         // 1: private-histogram select, shuffle scan, rank sort, bit-matrix NMS
    const double* half_over_1000;                     // [2000] (k / 2) / 1000, or null (v1: divide in the kernel)
    long long* trace; int trace_block;                // SSDB_TRACE=1: clock64() of CTA SSDB_TRACE_BLOCK at the phase boundaries (bring-up only)
};
#define SSDB_TRACE_PT(k) do { if (p.trace && blockIdx.x == p.trace_block && threadIdx.x == 0) p.trace[k] = clock64(); } while (0)

struct ConfKey {         // participants of the top-cap selection: anchors at or above the confidence threshold
    const unsigned int* ckey;
    __device__ __forceinline__ bool operator()(int a, unsigned& key) const { key = ckey[a]; return key != 0u; }
};

// VT = compile-time row width (C + 5), 0 = generic
template <int VT>
__global__ void __launch_bounds__(RT) decode_scan_kernel(const float* __restrict__ pred, int A, int C, float conf_thr, int use_bulk,
                                                          unsigned int* __restrict__ ckey_out, unsigned char* __restrict__ cls_out) {
    extern __shared__ __align__(128) unsigned char dyn[];
    float* zt = reinterpret_cast<float*>(dyn);
    __shared__ __align__(8) unsigned long long bar;
    const int V = VT ? VT : C + 5;
    const int tid = threadIdx.x, s = blockIdx.x, b = blockIdx.y;
    const int a0 = s * RT, rows = min(RT, A - a0), nfl = rows * V;
    const long long tile_off = ((long long)b * A + a0) * V;
    bulk::tile_load<RT>(zt, pred + tile_off, nullptr, nullptr, nfl, use_bulk, &bar);
    bulk::tile_load_wait(use_bulk, &bar);
    if (tid < rows) {
        const float* r = zt + tid * V;
        float best = r[0]; int cls = 0;
        if (VT) {
#pragma unroll
            for (int c = 1; c < (VT ? VT - 5 : 1); ++c) { float v = r[c]; if (v > best) { best = v; cls = c; } }
        } else {
            for (int c = 1; c < C; ++c) { float v = r[c]; if (v > best) { best = v; cls = c; } }
        }
        const bool ok = !(best < conf_thr);
        ckey_out[(long long)b * A + a0 + tid] = ok ? okey(best) : 0u;
        cls_out[(long long)b * A + a0 + tid] = (unsigned char)cls;      // np.argmax: first maximum
    }
}

__global__ void __launch_bounds__(DT) decode_nms_kernel(DetArgs p) {
    extern __shared__ __align__(128) unsigned char dyn[];
    unsigned int* ckey = reinterpret_cast<unsigned int*>(dyn);                 // [A] ordered confidence keys (0 = not a candidate)
    unsigned long long* keys;                                                  // [P] sort keys
    int* cand;                                                                 // [CAND_WORDS][P]
    const int P = p.P;
    const int b = blockIdx.x, tid = threadIdx.x;
    if (P <= SMEM_P_MAX) {
        size_t off = ((size_t)p.A * 4 + 15) / 16 * 16;
        keys = reinterpret_cast<unsigned long long*>(dyn + off);                // [P] (+ [P] sorted copy when P <= BITS_P_MAX)
        cand = reinterpret_cast<int*>(dyn + off + (size_t)P * 8 * (P <= BITS_P_MAX ? 2 : 1));
    } else {
        keys = p.g_keys + (size_t)b * P;
        cand = p.g_cand + (size_t)b * P * CAND_WORDS;
    }
    __shared__ int hist[256];
    __shared__ int sel_bin, sel_rem, n_gt;
    __shared__ int scan[DT];
    __shared__ int redi[DT / 32];
    __shared__ int first_pos[64];
    __shared__ int whist[(DT / 32) * 256];
    __shared__ int sel[2];
    __shared__ int wsum[32];
    __shared__ unsigned int cmask[64 * 8];       // per class: which of the (<= 256) sorted candidates belong to it
    __shared__ unsigned int rmask[64 * 8];       // per class: candidates removed by the greedy pass
    __shared__ __align__(4) unsigned char ccls[BITS_P_MAX];   // class of each sorted candidate (255 = empty slot)
    __shared__ int kcnt[64], kbase[64];          // per class: kept boxes, and kept boxes of the classes that come earlier in the output

    const int V = p.C + 5, A = p.A, C = p.C;
    SSDB_TRACE_PT(0);
    const float* pb = p.pred + (size_t)b * A * V;

    // ---- pass 1: arg-max class and confidence per anchor ----
    int nvalid = 0;
    if (p.ckey_in) {
        const unsigned int* gk = p.ckey_in + (size_t)b * A;
        for (int a = tid; a < A; a += DT) { unsigned int k = gk[a]; ckey[a] = k; nvalid += k != 0u; }
    } else
    for (int a = tid; a < A; a += DT) {
        const float* r = pb + (size_t)a * V;
        float best = r[0];
        for (int c = 1; c < C; ++c) { float v = r[c]; if (v > best) best = v; }
        bool ok = !(best < p.conf_thr);
        ckey[a] = ok ? okey(best) : 0u;       // okey() of any float is never 0 except for -NaN patterns; fine
        nvalid += ok;
    }
    // block sum
#pragma unroll
    for (int o = 16; o; o >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
    if ((tid & 31) == 0) redi[tid >> 5] = nvalid;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int i = 0; i < DT / 32; ++i) t += redi[i]; redi[0] = t; }
    __syncthreads();
    nvalid = redi[0];
    const int n = min(nvalid, p.cap_eff);
    SSDB_TRACE_PT(1);
    __syncthreads();

    if (n > 0) {
        // ---- exact n-th largest confidence key (radix select, 4 x 8 bits) ----
        unsigned int prefix = 0, mask = 0; int remaining = n;
        if (p.fast) radix_select_kth<DT>(A, n, ConfKey{ckey}, whist, hist, sel, prefix, remaining);
        else
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            for (int i = tid; i < 256; i += DT) hist[i] = 0;
            __syncthreads();
            for (int a = tid; a < A; a += DT) {
                unsigned int k = ckey[a];
                if (k != 0u && (k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255], 1);
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
        SSDB_TRACE_PT(2);
        // ---- compaction: keys above the pivot anywhere, `remaining` ties by lowest anchor index ----
        if (tid == 0) n_gt = 0;
        for (int i = tid; i < P; i += DT) keys[i] = 0ull;
        __syncthreads();
        const int per = (A + DT - 1) / DT;
        const int a_lo = tid * per, a_hi = min(A, a_lo + per);
        int ties = 0, above = 0;
        for (int a = a_lo; a < a_hi; ++a) { const unsigned int k = ckey[a]; ties += k == prefix; above += k > prefix; }
        int rank, gslot = -1;
        if (p.fast) {
            // one scan for both counts (each total <= A < 65536): slots come from prefix sums, no shared-memory atomics
            const int packed = block_excl_scan<DT>((above << 16) | ties, wsum);
            rank = packed & 0xffff; gslot = packed >> 16;
        } else {
            scan[tid] = ties;
            __syncthreads();
            for (int o = 1; o < DT; o <<= 1) {
                int v = tid >= o ? scan[tid - o] : 0;
                __syncthreads();
                scan[tid] += v;
                __syncthreads();
            }
            rank = scan[tid] - ties;
        }
        const int base_ties = n - remaining;       // number of keys strictly above the pivot
        for (int a = a_lo; a < a_hi; ++a) {
            unsigned int k = ckey[a];
            if (k == 0u) continue;
            int slot = -1;
            if (k > prefix) slot = gslot >= 0 ? gslot++ : atomicAdd(&n_gt, 1);
            else if (k == prefix) { if (rank < remaining) slot = base_ties + rank; ++rank; }
            if (slot >= 0) keys[slot] = ((unsigned long long)k << 32) | (unsigned long long)(0xffffffffu - (unsigned int)a);
        }
        __syncthreads();
        SSDB_TRACE_PT(3);
        if (p.fast && P <= BITS_P_MAX) {
            // ---- rank sort: keys are unique (they carry the anchor index), so #greater = final position; 4 lanes per key ----
            unsigned long long* sorted = keys + P;
            const int i = tid >> 2, q = tid & 3;
            const unsigned long long me = i < n ? keys[i] : 0ull;
            int cnt = 0;
            if (i < n) for (int j = q; j < n; j += 4) cnt += keys[j] > me;
            cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
            if (i < n && q == 0) sorted[cnt] = me;
            __syncthreads();
            keys = sorted;
        } else
        // ---- bitonic sort, descending ----
        for (int k2 = 2; k2 <= P; k2 <<= 1) {
            for (int j = k2 >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < P; i += DT) {
                    int l = i ^ j;
                    if (l > i) {
                        unsigned long long x = keys[i], y = keys[l];
                        bool desc = (i & k2) == 0;
                        if (desc ? (x < y) : (x > y)) { keys[i] = y; keys[l] = x; }
                    }
                }
                __syncthreads();
            }
        }
        SSDB_TRACE_PT(4);
        // ---- decode the n candidates ----
        for (int i = tid; i < C && i < 64; i += DT) first_pos[i] = 0x7fffffff;
        for (int i = tid; i < 64 * 8; i += DT) { cmask[i] = 0u; rmask[i] = 0u; }
        for (int i = tid; i < BITS_P_MAX; i += DT) ccls[i] = 255;
        __syncthreads();
        for (int i = tid; i < n; i += DT) {
            unsigned long long kk = keys[i];
            int a = (int)(0xffffffffu - (unsigned int)(kk & 0xffffffffull));
            const float* r = pb + (size_t)a * V;
            float o0 = r[C + 1], o1 = r[C + 2], o2 = r[C + 3], o3 = r[C + 4];
            int cls = 0;
            if (p.cls_in) cls = p.cls_in[(size_t)b * A + a];
            else { float best = r[0]; for (int c = 1; c < C; ++c) { float v = r[c]; if (v > best) { best = v; cls = c; } } }
            o0 = o0 > 100.f ? 100.f : o0; o1 = o1 > 100.f ? 100.f : o1;
            o2 = o2 > 100.f ? 100.f : o2; o3 = o3 > 100.f ? 100.f : o3;
            const double ax = p.anchors[a * 4 + 0], ay = p.anchors[a * 4 + 1], aw = p.anchors[a * 4 + 2], ah = p.anchors[a * 4 + 3];
            float x = __fadd_rn(__fmul_rn(__fdiv_rn(o0, 10.f), (float)aw), (float)ax);
            float y = __fadd_rn(__fmul_rn(__fdiv_rn(o1, 10.f), (float)ah), (float)ay);
            double w = __dmul_rn(exp((double)__fdiv_rn(o2, 5.f)), aw);
            double h = __dmul_rn(exp((double)__fdiv_rn(o3, 5.f)), ah);
            float px = __fmul_rn(x, 1000.f), py = __fmul_rn(y, 1000.f);
            float hw = (float)__dmul_rn(__dmul_rn(w, 1000.0), 0.5);          // x * 0.5 == x / 2 exactly
            float hh = (float)__dmul_rn(__dmul_rn(h, 1000.0), 0.5);
            int x0 = (int)__fsub_rn(px, hw), x1 = (int)__fadd_rn(px, hw);
            int y0 = (int)__fsub_rn(py, hh), y1 = (int)__fadd_rn(py, hh);
            x0 = max(x0, 0); x1 = min(x1, 999); y0 = max(y0, 0); y1 = min(y1, 999);
            x0 = min(x0, x1); y0 = min(y0, y1);
            // abs2prop (float64) then the NMS stage's prop2abs (float64)
            double bw = (double)(x1 - x0), bh = (double)(y1 - y0);
            // (x0 + bw/2) / 1000 and bw / 1000: the numerators are half-integers in [0, 999.5], so the four float64 divisions
            // come from a 2000-entry table of (k / 2) / 1000 computed on the host with the same IEEE division
            double pcx, pcy, sw, sh;
            if (p.half_over_1000) {
                pcx = __ldg(p.half_over_1000 + (x0 + x1)); pcy = __ldg(p.half_over_1000 + (y0 + y1));
                sw = __ldg(p.half_over_1000 + 2 * (x1 - x0)); sh = __ldg(p.half_over_1000 + 2 * (y1 - y0));
            } else {
                pcx = __ddiv_rn(__dadd_rn((double)x0, __ddiv_rn(bw, 2.0)), 1000.0);
                pcy = __ddiv_rn(__dadd_rn((double)y0, __ddiv_rn(bh, 2.0)), 1000.0);
                sw = __ddiv_rn(bw, 1000.0); sh = __ddiv_rn(bh, 1000.0);
            }
            double hw2 = __dmul_rn(__dmul_rn(sw, 1000.0), 0.5), hh2 = __dmul_rn(__dmul_rn(sh, 1000.0), 0.5);
            double cx2 = __dmul_rn(pcx, 1000.0), cy2 = __dmul_rn(pcy, 1000.0);
            cand[0 * P + i] = cls;
            cand[1 * P + i] = x0; cand[2 * P + i] = x1; cand[3 * P + i] = y0; cand[4 * P + i] = y1;
            cand[5 * P + i] = (int)__dsub_rn(cx2, hw2); cand[6 * P + i] = (int)__dadd_rn(cx2, hw2);
            cand[7 * P + i] = (int)__dsub_rn(cy2, hh2); cand[8 * P + i] = (int)__dadd_rn(cy2, hh2);
            cand[9 * P + i] = (int)__float_as_uint(okey_inv((unsigned int)(kk >> 32)));
            cand[10 * P + i] = a;
            if (cls < 64) atomicMin(&first_pos[cls], i);
            if (P <= BITS_P_MAX) { atomicOr(&cmask[cls * 8 + (i >> 5)], 1u << (i & 31)); ccls[i] = (unsigned char)cls; }
        }
        __syncthreads();
        SSDB_TRACE_PT(5);
        // ---- greedy NMS in confidence order; alive flags reuse ckey[] ----
        unsigned int* alive = ckey;
        if (p.fast && P <= BITS_P_MAX) {
            // Lazy greedy NMS, one WARP per class (classes never interact).  Only a KEPT box ever suppresses anything, and
            // only ~6% of the candidates are kept, so no suppression matrix is built: the warp compacts its class's members
            // (confidence order) into a list, one position per lane and 32-step; the next alive position is kept, the lanes
            // test their own later alive positions against it and mark them dead in a private bit mask.
            // IoU > thr without float64 on the common path: areas fit 32 bits, inter / uni converts exactly to fp32 and
            // thr_f * uni is within 1.2e-7 (relative) of thr * uni; only a pair within 1e-6 of the threshold takes the
            // reference's rounded float64 division.
            {
                const int lane = tid & 31;
                const float thr_f = (float)p.iou_thr;
                unsigned char* ml = reinterpret_cast<unsigned char*>(cand + CAND_WORDS * P) + (tid >> 5) * P;   // this warp's member list (P bytes)
                for (int c = tid >> 5; c < C; c += DT / 32) {
                    // the class's candidates in confidence order, compacted: position q of the list lives in lane q % 32
                    int cnt = 0;
                    for (int w = 0; w < 8; ++w) {
                        const unsigned int m = cmask[c * 8 + w];
                        if ((m >> lane) & 1u) ml[cnt + __popc(m & ((1u << lane) - 1u))] = (unsigned char)(w * 32 + lane);
                        cnt += __popc(m);
                    }
                    __syncwarp();
                    const int K = (cnt + 31) >> 5;
                    unsigned int dead = 0u;                       // bit k: list position lane + 32k was suppressed
                    for (int k = 0; k < K; ++k) {
                        unsigned int todo = __ballot_sync(0xffffffffu, lane + 32 * k < cnt);
                        while (true) {                                                 // warp-uniform
                            todo &= ~__ballot_sync(0xffffffffu, (dead >> k) & 1u);
                            if (!todo) break;
                            const int src = __ffs(todo) - 1;                           // next alive member: it is kept
                            todo &= 0xfffffffeu << src;
                            const int pos = 32 * k + src, i = ml[pos];
                            const int ix0 = cand[5 * P + i], ix1 = cand[6 * P + i], iy0 = cand[7 * P + i], iy1 = cand[8 * P + i];
                            const int area_i = (ix1 - ix0 + 1) * (iy1 - iy0 + 1);     // <= 10^6
                            for (int k2 = k; k2 < K; ++k2) {
                                const int q = lane + 32 * k2;
                                if (q <= pos || q >= cnt || ((dead >> k2) & 1u)) continue;
                                const int j = ml[q];
                                const int jx0 = cand[5 * P + j], jx1 = cand[6 * P + j], jy0 = cand[7 * P + j], jy1 = cand[8 * P + j];
                                int iw = min(ix1, jx1) - max(ix0, jx0) + 1; iw = iw < 0 ? 0 : iw;
                                int ih = min(iy1, jy1) - max(iy0, jy0) + 1; ih = ih < 0 ? 0 : ih;
                                const int inter = iw * ih;
                                const int uni = area_i + (jx1 - jx0 + 1) * (jy1 - jy0 + 1) - inter;
                                const float fi = (float)inter, lim = thr_f * (float)uni;
                                bool hit;
                                if (inter == 0) hit = 0.0 > p.iou_thr;
                                else if (fi > lim * 1.000001f && lim >= 0.f) hit = true;
                                else if (fi < lim * 0.999999f) hit = false;
                                else hit = __ddiv_rn((double)inter, (double)uni) > p.iou_thr;
                                if (hit) dead |= 1u << k2;
                            }
                        }
                    }
                    for (int k = 0; k < K; ++k) {
                        const int q = lane + 32 * k;
                        if (q < cnt && ((dead >> k) & 1u)) { const int j = ml[q]; atomicOr(&rmask[c * 8 + (j >> 5)], 1u << (j & 31)); }
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
            SSDB_TRACE_PT(8);
            // ---- output order = classes by first appearance, confidence order inside a class: with the per-class kept
            //      masks this is a handful of popcounts per candidate (no pass over the other candidates) ----
            if (tid < 64) {
                int sk = 0;
                if (tid < C)
#pragma unroll
                    for (int w = 0; w < 8; ++w) sk += __popc(cmask[tid * 8 + w] & ~rmask[tid * 8 + w]);
                kcnt[tid] = sk;
            }
            __syncthreads();
            if (tid < 64) {
                int base = 0;
                if (tid < C) { const int fp = first_pos[tid]; for (int c2 = 0; c2 < C; ++c2) if (first_pos[c2] < fp) base += kcnt[c2]; }
                kbase[tid] = base;
            }
            if (tid == 0) { int t = 0; for (int c2 = 0; c2 < C; ++c2) t += kcnt[c2]; p.counts[b * 2] = t; p.counts[b * 2 + 1] = n; }
            __syncthreads();
            SSDB_TRACE_PT(9);
            for (int i = tid; i < n; i += DT) {
                const int ci = ccls[i], w = i >> 5;
                const unsigned int aw = cmask[ci * 8 + w] & ~rmask[ci * 8 + w];
                if (!((aw >> (i & 31)) & 1u)) continue;
                int rnk = kbase[ci] + __popc(aw & ((1u << (i & 31)) - 1u));
                for (int w2 = 0; w2 < w; ++w2) rnk += __popc(cmask[ci * 8 + w2] & ~rmask[ci * 8 + w2]);
                int* o = p.dets + ((size_t)b * p.cap_eff + rnk) * 8;
                o[0] = cand[9 * P + i]; o[1] = ci;
                o[2] = cand[1 * P + i]; o[3] = cand[2 * P + i]; o[4] = cand[3 * P + i]; o[5] = cand[4 * P + i];
                o[6] = cand[10 * P + i]; o[7] = i;
            }
        } else {
        for (int i = tid; i < n; i += DT) alive[i] = 1u;
        __syncthreads();
        for (int i = 0; i < n; ++i) {
            if (!alive[i]) continue;                    // uniform: written only before a barrier
            const int ci = cand[0 * P + i];
            const int ix0 = cand[5 * P + i], ix1 = cand[6 * P + i], iy0 = cand[7 * P + i], iy1 = cand[8 * P + i];
            const long long area_i = (long long)(ix1 - ix0 + 1) * (iy1 - iy0 + 1);
            for (int j = i + 1 + tid; j < n; j += DT) {
                if (!alive[j] || cand[0 * P + j] != ci) continue;
                int jx0 = cand[5 * P + j], jx1 = cand[6 * P + j], jy0 = cand[7 * P + j], jy1 = cand[8 * P + j];
                int iw = min(ix1, jx1) - max(ix0, jx0) + 1; iw = iw < 0 ? 0 : iw;
                int ih = min(iy1, jy1) - max(iy0, jy0) + 1; ih = ih < 0 ? 0 : ih;
                long long inter = (long long)iw * ih;
                long long uni = area_i + (long long)(jx1 - jx0 + 1) * (jy1 - jy0 + 1) - inter;
                if (__ddiv_rn((double)inter, (double)uni) > p.iou_thr) alive[j] = 0u;
            }
            __syncthreads();
        }
        SSDB_TRACE_PT(9);
        // ---- output rank: classes by first appearance, confidence order inside a class ----
        int kept = 0;
        for (int i = tid; i < n; i += DT) kept += alive[i] ? 1 : 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
        __syncthreads();
        if ((tid & 31) == 0) redi[tid >> 5] = kept;
        __syncthreads();
        if (tid == 0) { int t = 0; for (int i = 0; i < DT / 32; ++i) t += redi[i]; p.counts[b * 2] = t; p.counts[b * 2 + 1] = n; }
        for (int i = tid; i < n; i += DT) {
            if (!alive[i]) continue;
            const int ci = cand[0 * P + i];
            const int fi = ci < 64 ? first_pos[ci] : 0;
            int rnk = 0;
            for (int j = 0; j < n; ++j) {
                if (!alive[j]) continue;
                const int cj = cand[0 * P + j];
                const int fj = cj < 64 ? first_pos[cj] : 0;
                if (fj < fi || (fj == fi && j < i)) ++rnk;
            }
            int* o = p.dets + ((size_t)b * p.cap_eff + rnk) * 8;
            o[0] = cand[9 * P + i]; o[1] = ci;
            o[2] = cand[1 * P + i]; o[3] = cand[2 * P + i]; o[4] = cand[3 * P + i]; o[5] = cand[4 * P + i];
            o[6] = cand[10 * P + i]; o[7] = i;
        }
        }
    } else {
        if (tid == 0) { p.counts[b * 2] = 0; p.counts[b * 2 + 1] = 0; }
    }
    __syncthreads();
    SSDB_TRACE_PT(10);
}


// Stand-alone class-wise greedy NMS over already-decoded boxes (the reference calls
// suppress_overlaps / non_maximum_suppression on decode_boxes' output, ssdutils.py:232-318).
// boxes: [n,4] int (xmin,xmax,ymin,ymax) as prop2abs yields them; one CTA; global scratch.
__global__ void __launch_bounds__(DT) nms_only_kernel(const int* __restrict__ boxes, const int* __restrict__ cls,
                                                       const float* __restrict__ conf, int n, int P, double iou_thr,
                                                       unsigned long long* __restrict__ keys, unsigned int* __restrict__ alive,
                                                       int* __restrict__ first_pos, int nclass, int* __restrict__ keep_out,
                                                       int* __restrict__ count_out) {
    const int tid = threadIdx.x;
    __shared__ int redi[DT / 32];
    for (int i = tid; i < P; i += DT)
        keys[i] = i < n ? (((unsigned long long)okey(conf[i]) << 32) | (unsigned long long)(0xffffffffu - (unsigned int)i)) : 0ull;
    for (int i = tid; i < nclass; i += DT) first_pos[i] = 0x7fffffff;
    __syncthreads();
    for (int i = tid; i < n; i += DT) atomicMin(&first_pos[cls[i]], i);      // class order = first appearance in the INPUT list
    for (int k2 = 2; k2 <= P; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P; i += DT) {
                int l = i ^ j;
                if (l > i) {
                    unsigned long long x = keys[i], y = keys[l];
                    bool desc = (i & k2) == 0;
                    if (desc ? (x < y) : (x > y)) { keys[i] = y; keys[l] = x; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < n; i += DT) alive[i] = 1u;
    __syncthreads();
    for (int i = 0; i < n; ++i) {
        if (!alive[i]) continue;
        const int a = (int)(0xffffffffu - (unsigned int)(keys[i] & 0xffffffffull));
        const int ci = cls[a];
        const int ix0 = boxes[a * 4], ix1 = boxes[a * 4 + 1], iy0 = boxes[a * 4 + 2], iy1 = boxes[a * 4 + 3];
        const long long area_i = (long long)(ix1 - ix0 + 1) * (iy1 - iy0 + 1);
        for (int j = i + 1 + tid; j < n; j += DT) {
            if (!alive[j]) continue;
            const int bj = (int)(0xffffffffu - (unsigned int)(keys[j] & 0xffffffffull));
            if (cls[bj] != ci) continue;
            int jx0 = boxes[bj * 4], jx1 = boxes[bj * 4 + 1], jy0 = boxes[bj * 4 + 2], jy1 = boxes[bj * 4 + 3];
            int iw = min(ix1, jx1) - max(ix0, jx0) + 1; iw = iw < 0 ? 0 : iw;
            int ih = min(iy1, jy1) - max(iy0, jy0) + 1; ih = ih < 0 ? 0 : ih;
            long long inter = (long long)iw * ih;
            long long uni = area_i + (long long)(jx1 - jx0 + 1) * (jy1 - jy0 + 1) - inter;
            if (__ddiv_rn((double)inter, (double)uni) > iou_thr) alive[j] = 0u;
        }
        __syncthreads();
    }
    int kept = 0;
    for (int i = tid; i < n; i += DT) kept += alive[i] ? 1 : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
    if ((tid & 31) == 0) redi[tid >> 5] = kept;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int i = 0; i < DT / 32; ++i) t += redi[i]; count_out[0] = t; }
    for (int i = tid; i < n; i += DT) {
        if (!alive[i]) continue;
        const int a = (int)(0xffffffffu - (unsigned int)(keys[i] & 0xffffffffull));
        const int fi = first_pos[cls[a]];
        int rnk = 0;
        for (int j = 0; j < n; ++j) {
            if (!alive[j]) continue;
            const int bj = (int)(0xffffffffu - (unsigned int)(keys[j] & 0xffffffffull));
            const int fj = first_pos[cls[bj]];
            if (fj < fi || (fj == fi && j < i)) ++rnk;
        }
        keep_out[rnk] = a;
    }
}

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

bool nms_v1() {        // read per call so that a test can run both implementations in one process
    const char* e = getenv("SSDB_NMS");
    return e && e[0] == 'v' && e[1] == '1';
}

size_t align256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

// scratch = [B*A] u32 confidence keys, then (cap > 1024 only) sort keys + candidate records in global memory
size_t decode_nms_scratch_bytes(int B, int A, int cap) {
    int cap_eff = (cap > 0 && cap < A) ? cap : A;
    int P = next_pow2(cap_eff);
    size_t keys = align256((size_t)B * A * 4) + align256((size_t)B * A);
    if (P <= SMEM_P_MAX) return keys;
    return keys + (size_t)B * P * (8 + 4 * CAND_WORDS);
}

int decode_nms_launch(const float* pred, int B, int A, int C, const double* anchors_prop, float conf_thr, int cap,
                      double iou_thr, int* dets_out, int* counts_out, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    SSDB_REQUIRE(B >= 1 && A >= 1 && C >= 1 && C <= 64, "bad sizes");
    SSDB_REQUIRE(scratch && scratch_bytes >= decode_nms_scratch_bytes(B, A, cap), "scratch too small");
    DetArgs p;
    p.pred = pred; p.anchors = anchors_prop; p.B = B; p.A = A; p.C = C; p.conf_thr = conf_thr; p.cap = cap; p.iou_thr = iou_thr;
    p.dets = dets_out; p.counts = counts_out;
    p.cap_eff = (cap > 0 && cap < A) ? cap : A;
    p.P = next_pow2(p.cap_eff);
    size_t sh = ((size_t)A * 4 + 15) / 16 * 16;
    p.g_keys = nullptr; p.g_cand = nullptr; p.ckey_in = nullptr; p.cls_in = nullptr; p.fast = nms_v1() ? 0 : 1;
    static PerDevice<double*> table_pd;
    double*& table_dev = table_pd.get();
    if (!table_dev) {
        std::vector<double> t(2000);
        for (int k = 0; k < 2000; ++k) { volatile double half = (double)k / 2.0; t[k] = half / 1000.0; }
        SSDB_CUDA(cudaMalloc(&table_dev, t.size() * sizeof(double)));
        SSDB_CUDA(cudaMemcpy(table_dev, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    p.half_over_1000 = p.fast ? table_dev : nullptr;
    unsigned char* sc = reinterpret_cast<unsigned char*>(scratch);
    const size_t keys_bytes = align256((size_t)B * A * 4) + align256((size_t)B * A);
    if (p.P <= SMEM_P_MAX) {
        sh += (size_t)p.P * (8 + 4 * CAND_WORDS);
        if (p.P <= BITS_P_MAX) sh += (size_t)p.P * 8 + (size_t)p.P * 8 * 4;      // sorted keys + suppression matrix rows of 8 words
    } else {
        p.g_keys = reinterpret_cast<unsigned long long*>(sc + keys_bytes);
        p.g_cand = reinterpret_cast<int*>(sc + keys_bytes + (size_t)B * p.P * 8);
    }
    static PerDevice<size_t> dyn_pd;      // 227 KB per CTA minus the kernel's static shared memory
    size_t& dyn_max = dyn_pd.get();
    if (!dyn_max) {
        cudaFuncAttributes fa;
        SSDB_CUDA(cudaFuncGetAttributes(&fa, decode_nms_kernel));
        size_t m = 232448 - fa.sharedSizeBytes;
        SSDB_CUDA(cudaFuncSetAttribute(decode_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m));
        SSDB_CUDA(cudaFuncSetAttribute(decode_scan_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, RT * MAXV * 4));
        dyn_max = m;
    }
    SSDB_REQUIRE(sh <= dyn_max, "anchor count too large for one CTA");
    if (!nms_v1()) {
        const int V = C + 5, S = (A + RT - 1) / RT;
        const int use_bulk = ((size_t)A * V * 4) % 16 == 0 && bulk::aligned16(pred);
        unsigned int* ckey = reinterpret_cast<unsigned int*>(sc);
        unsigned char* ccls = sc + align256((size_t)B * A * 4);
        if (C == 20) decode_scan_kernel<25><<<dim3(S, B), RT, (size_t)RT * V * 4, st>>>(pred, A, C, conf_thr, use_bulk, ckey, ccls);
        else decode_scan_kernel<0><<<dim3(S, B), RT, (size_t)RT * V * 4, st>>>(pred, A, C, conf_thr, use_bulk, ckey, ccls);
        SSDB_LAUNCH_CHECK();
        p.ckey_in = ckey; p.cls_in = ccls;
    }
    static PerDevice<long long*> trace_pd;
    long long*& trace_dev = trace_pd.get();
    const bool tracing = getenv("SSDB_TRACE") != nullptr;
    p.trace = nullptr; p.trace_block = getenv("SSDB_TRACE_BLOCK") ? atoi(getenv("SSDB_TRACE_BLOCK")) : 0;
    if (tracing) {
        if (!trace_dev) SSDB_CUDA(cudaMalloc(&trace_dev, 64 * sizeof(long long)));
        SSDB_CUDA(cudaMemsetAsync(trace_dev, 0, 64 * sizeof(long long), st));
        p.trace = trace_dev;
    }
    decode_nms_kernel<<<B, DT, sh, st>>>(p);
    SSDB_LAUNCH_CHECK();
    if (tracing) {
        long long h[64];
        SSDB_CUDA(cudaStreamSynchronize(st));
        SSDB_CUDA(cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "ssdb trace decode_nms_kernel (cycles since entry):");
        for (int k = 1; k <= 10; ++k) fprintf(stderr, " p%d=%lld", k, h[k] ? h[k] - h[0] : -1);
        fprintf(stderr, "\n");
    }
    return SSDB_OK;
}


int nms_only_host(const int* boxes, const int* cls, const float* conf, int n, int nclass, double iou_thr, int* keep_out, int* count_out) {
    SSDB_REQUIRE(n >= 1 && nclass >= 1 && boxes && cls && conf && keep_out && count_out, "bad arguments");
    int P = next_pow2(n);
    unsigned char* d = nullptr;
    size_t off_cls = (size_t)n * 16, off_conf = off_cls + (size_t)n * 4, off_keys = (off_conf + (size_t)n * 4 + 7) / 8 * 8;
    size_t off_alive = off_keys + (size_t)P * 8, off_first = off_alive + (size_t)n * 4, off_keep = off_first + (size_t)nclass * 4;
    size_t total = off_keep + (size_t)n * 4 + 4;
    SSDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&d), total));
    SSDB_CUDA(cudaMemcpy(d, boxes, (size_t)n * 16, cudaMemcpyHostToDevice));
    SSDB_CUDA(cudaMemcpy(d + off_cls, cls, (size_t)n * 4, cudaMemcpyHostToDevice));
    SSDB_CUDA(cudaMemcpy(d + off_conf, conf, (size_t)n * 4, cudaMemcpyHostToDevice));
    nms_only_kernel<<<1, DT>>>(reinterpret_cast<int*>(d), reinterpret_cast<int*>(d + off_cls), reinterpret_cast<float*>(d + off_conf), n, P,
                               iou_thr, reinterpret_cast<unsigned long long*>(d + off_keys), reinterpret_cast<unsigned int*>(d + off_alive),
                               reinterpret_cast<int*>(d + off_first), nclass, reinterpret_cast<int*>(d + off_keep),
                               reinterpret_cast<int*>(d + off_keep + (size_t)n * 4));
    SSDB_LAUNCH_CHECK();
    SSDB_CUDA(cudaMemcpy(count_out, d + off_keep + (size_t)n * 4, 4, cudaMemcpyDeviceToHost));
    SSDB_CUDA(cudaMemcpy(keep_out, d + off_keep, (size_t)n * 4, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return SSDB_OK;
}

}  // namespace ssdb
