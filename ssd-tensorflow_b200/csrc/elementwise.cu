// Bandwidth-bound layer kernels of the SSD-VGG graph: max pools (the VGG 2x2/s2
// SAME pools and mod_pool5 3x3/s1, reference ssdvgg.py:190-207,234), the L2
// normalisation of conv4_3 (ssdvgg.py:80-84), head-gradient re-layout, the
// result softmax (ssdvgg.py:368-372) and the Momentum update (ssdvgg.py:585-588).
// All NHWC float32, vectorised float4 along channels, deterministic reductions.
#include "common.cuh"

namespace ssdb {
namespace {

template <int FMT>
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, int B, int H, int W, int C4, int k, int s,
                                   int pt, int pl, int Ho, int Wo, float* __restrict__ y) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * Ho * Wo * C4;
    if (i >= total) return;
    int c = (int)(i % C4); long long r = i / C4;
    int ox = (int)(r % Wo); r /= Wo;
    int oy = (int)(r % Ho); int b = (int)(r / Ho);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int dy = 0; dy < k; ++dy) {
        int iy = oy * s + dy - pt;
        if (iy < 0 || iy >= H) continue;
        for (int dx = 0; dx < k; ++dx) {
            int ix = ox * s + dx - pl;
            if (ix < 0 || ix >= W) continue;
            float4 v = act_ld4<FMT>(x, ((((long long)b * H + iy) * W + ix) * C4 + c) * 4);
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
    }
    act_st4<FMT>(y, i * 4, m);
}

// forward of a pool that also records, per output element, which window cell won (first maximum, row-major): 4 bytes per
// float4 group.  The backward of overlapping windows (mod_pool5, 3x3 stride 1) then tests 9 bytes per input element
// instead of re-scanning 81 neighbours.
template <int FMT>
__global__ void maxpool_fwd_arg_kernel(const float* __restrict__ x, int B, int H, int W, int C4, int k, int s,
                                       int pt, int pl, int Ho, int Wo, float* __restrict__ y, uchar4* __restrict__ arg) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * Ho * Wo * C4;
    if (i >= total) return;
    int c = (int)(i % C4); long long r = i / C4;
    int ox = (int)(r % Wo); r /= Wo;
    int oy = (int)(r % Ho); int b = (int)(r / Ho);
    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    unsigned char a[4] = {255, 255, 255, 255};
    for (int dy = 0; dy < k; ++dy) {
        int iy = oy * s + dy - pt;
        if (iy < 0 || iy >= H) continue;
        for (int dx = 0; dx < k; ++dx) {
            int ix = ox * s + dx - pl;
            if (ix < 0 || ix >= W) continue;
            float4 v = act_ld4<FMT>(x, ((((long long)b * H + iy) * W + ix) * C4 + c) * 4);
            float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) if (vv[q] > m[q] || a[q] == 255) { m[q] = vv[q]; a[q] = (unsigned char)(dy * k + dx); }
        }
    }
    act_st4<FMT>(y, i * 4, make_float4(m[0], m[1], m[2], m[3]));
    arg[i] = make_uchar4(a[0], a[1], a[2], a[3]);
}

template <int FMT>
__global__ void maxpool_bwd_arg_kernel(const float* __restrict__ x, const float* __restrict__ dy, const uchar4* __restrict__ arg, int B,
                                       int H, int W, int C4, int k, int s, int pt, int pl, int Ho, int Wo, int beta, int relu_mask,
                                       int round_out, float* __restrict__ dx) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * H * W * C4;
    if (i >= total) return;
    int c = (int)(i % C4); long long r = i / C4;
    int ix = (int)(r % W); r /= W;
    int iy = (int)(r % H); int b = (int)(r / H);
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    for (int wy = 0; wy < k; ++wy) {
        int ny = iy + pt - wy;                       // window row oy with oy*s + wy - pt == iy
        if (ny < 0 || ny % s) continue;
        int oy = ny / s; if (oy >= Ho) continue;
        for (int wx = 0; wx < k; ++wx) {
            int nx = ix + pl - wx;
            if (nx < 0 || nx % s) continue;
            int ox = nx / s; if (ox >= Wo) continue;
            long long o = (((long long)b * Ho + oy) * Wo + ox) * C4 + c;
            uchar4 a = arg[o];
            float4 gy = act_ld4<FMT>(dy, o * 4);
            unsigned char cell = (unsigned char)(wy * k + wx);
            if (a.x == cell) g[0] += gy.x;
            if (a.y == cell) g[1] += gy.y;
            if (a.z == cell) g[2] += gy.z;
            if (a.w == cell) g[3] += gy.w;
        }
    }
    if (beta) { float4 o = act_ld4<FMT>(dx, i * 4); g[0] += o.x; g[1] += o.y; g[2] += o.z; g[3] += o.w; }
    if (relu_mask) {
        float4 me = act_ld4_sign<FMT>(x, i * 4);
        g[0] = me.x > 0.f ? g[0] : 0.f; g[1] = me.y > 0.f ? g[1] : 0.f; g[2] = me.z > 0.f ? g[2] : 0.f; g[3] = me.w > 0.f ? g[3] : 0.f;
    }
    if (FMT == ACT_F32 && round_out) { g[0] = tf32_rn(g[0]); g[1] = tf32_rn(g[1]); g[2] = tf32_rn(g[2]); g[3] = tf32_rn(g[3]); }
    act_st4<FMT>(dx, i * 4, make_float4(g[0], g[1], g[2], g[3]));
}

// one thread per input element group (float4 of channels): sum dy over the windows whose
// first maximum (row-major scan) is this element
template <int FMT>
__global__ void maxpool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int B, int H, int W,
                                   int C4, int k, int s, int pt, int pl, int Ho, int Wo, int beta, int relu_mask, int round_out,
                                   float* __restrict__ dx) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * H * W * C4;
    if (i >= total) return;
    int c = (int)(i % C4); long long r = i / C4;
    int ix = (int)(r % W); r /= W;
    int iy = (int)(r % H); int b = (int)(r / H);
    float4 me = act_ld4<FMT>(x, i * 4);
    float mev[4] = {me.x, me.y, me.z, me.w};
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    // windows (oy,ox) with oy*s - pt <= iy <= oy*s - pt + k - 1
    int oy_lo = (iy + pt - (k - 1) + s - 1); oy_lo = oy_lo < 0 ? 0 : oy_lo / s;
    int oy_hi = (iy + pt) / s; if (oy_hi > Ho - 1) oy_hi = Ho - 1;
    int ox_lo = (ix + pl - (k - 1) + s - 1); ox_lo = ox_lo < 0 ? 0 : ox_lo / s;
    int ox_hi = (ix + pl) / s; if (ox_hi > Wo - 1) ox_hi = Wo - 1;
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
        for (int ox = ox_lo; ox <= ox_hi; ++ox) {
            // is (iy,ix) the first max of window (oy,ox)?
            bool first[4] = {true, true, true, true};
            for (int wy = 0; wy < k; ++wy) {
                int yy = oy * s + wy - pt;
                if (yy < 0 || yy >= H) continue;
                for (int wx = 0; wx < k; ++wx) {
                    int xx = ox * s + wx - pl;
                    if (xx < 0 || xx >= W) continue;
                    if (yy == iy && xx == ix) continue;
                    float4 o = act_ld4<FMT>(x, ((((long long)b * H + yy) * W + xx) * C4 + c) * 4);
                    float ov[4] = {o.x, o.y, o.z, o.w};
                    bool before = (yy < iy) || (yy == iy && xx < ix);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (before ? (ov[q] >= mev[q]) : (ov[q] > mev[q])) first[q] = false;
                    }
                }
            }
            float4 gy = act_ld4<FMT>(dy, ((((long long)b * Ho + oy) * Wo + ox) * C4 + c) * 4);
            float gv[4] = {gy.x, gy.y, gy.z, gy.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) if (first[q]) g[q] += gv[q];
        }
    }
    if (beta) { float4 o = act_ld4<FMT>(dx, i * 4); g[0] += o.x; g[1] += o.y; g[2] += o.z; g[3] += o.w; }
    if (relu_mask) {
#pragma unroll
        for (int q = 0; q < 4; ++q) g[q] = mev[q] > 0.f ? g[q] : 0.f;
    }
    if (FMT == ACT_F32 && round_out) { g[0] = tf32_rn(g[0]); g[1] = tf32_rn(g[1]); g[2] = tf32_rn(g[2]); g[3] = tf32_rn(g[3]); }
    act_st4<FMT>(dx, i * 4, make_float4(g[0], g[1], g[2], g[3]));
}

// 2x2 stride-2 pools with a code byte per output element: bits 0-1 = winning cell (first maximum, row-major), bit 2 =
// winner > 0 (the ReLU mask of the producing conv at the only cell that can receive gradient).  The backward then reads
// dy + 1 byte instead of the four input activations: 21 instead of 36 bytes per window element.
template <int FMT>
__global__ void maxpool2x2_fwd_code_kernel(const float* __restrict__ x, int B, int H, int W, int C4, int Ho, int Wo,
                                           float* __restrict__ y, uchar4* __restrict__ code) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * Ho * Wo * C4;
    if (i >= total) return;
    int c = (int)(i % C4); long long r = i / C4;
    int ox = (int)(r % Wo); r /= Wo;
    int oy = (int)(r % Ho); int b = (int)(r / Ho);
    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int a[4] = {0, 0, 0, 0};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        int iy = oy * 2 + (t >> 1), ix = ox * 2 + (t & 1);
        if (iy >= H || ix >= W) continue;
        float4 q = act_ld4<FMT>(x, ((((long long)b * H + iy) * W + ix) * C4 + c) * 4);
        float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) if (t == 0 || v[e] > m[e]) { m[e] = v[e]; a[e] = t; }
    }
    act_st4<FMT>(y, i * 4, make_float4(m[0], m[1], m[2], m[3]));
    unsigned char cb[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) cb[e] = (unsigned char)(a[e] | (m[e] > 0.f ? 4 : 0));
    code[i] = make_uchar4(cb[0], cb[1], cb[2], cb[3]);
}

// dx = routed dy (* winner > 0 when relu_mask); overwrites dx (callers use it only when nothing was accumulated before)
template <int FMT>
__global__ void maxpool2x2_bwd_code_kernel(const float* __restrict__ dy, const uchar4* __restrict__ code, int B, int H, int W, int C4,
                                           int Ho, int Wo, int relu_mask, int round_out, float* __restrict__ dx) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * Ho * Wo * C4;
    if (i >= total) return;
    int c = (int)(i % C4); long long r = i / C4;
    int ox = (int)(r % Wo); r /= Wo;
    int oy = (int)(r % Ho); int b = (int)(r / Ho);
    float4 gy = act_ld4<FMT>(dy, i * 4);
    uchar4 cd = code[i];
    float gv[4] = {gy.x, gy.y, gy.z, gy.w};
    const unsigned char cv[4] = {cd.x, cd.y, cd.z, cd.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        if (relu_mask && !(cv[e] & 4)) gv[e] = 0.f;
        if (FMT == ACT_F32 && round_out) gv[e] = tf32_rn(gv[e]);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        int iy = oy * 2 + (t >> 1), ix = ox * 2 + (t & 1);
        if (iy >= H || ix >= W) continue;
        float4 g = make_float4((cv[0] & 3) == t ? gv[0] : 0.f, (cv[1] & 3) == t ? gv[1] : 0.f, (cv[2] & 3) == t ? gv[2] : 0.f,
                               (cv[3] & 3) == t ? gv[3] : 0.f);
        act_st4<FMT>(dx, ((((long long)b * H + iy) * W + ix) * C4 + c) * 4, g);
    }
}

// 2x2 stride-2 windows do not overlap: one thread per (window, 4 channels) reads its (up to) 4 inputs once,
// routes dy to the first maximum and writes all 4 gradients.  pad_before is 0 for these pools
// (TF SAME on even sizes; 75 -> 38 pads after), so windows only clip at the bottom / right edge.
template <int FMT>
__global__ void maxpool2x2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int B, int H, int W, int C4,
                                      int Ho, int Wo, int beta, int relu_mask, int round_out, float* __restrict__ dx) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * Ho * Wo * C4;
    if (i >= total) return;
    int c = (int)(i % C4); long long r = i / C4;
    int ox = (int)(r % Wo); r /= Wo;
    int oy = (int)(r % Ho); int b = (int)(r / Ho);
    float4 gy = act_ld4<FMT>(dy, i * 4);
    float gv[4] = {gy.x, gy.y, gy.z, gy.w};
    float v[4][4]; long long idx[4]; bool ok[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        int iy = oy * 2 + (t >> 1), ix = ox * 2 + (t & 1);
        ok[t] = iy < H && ix < W;
        idx[t] = (((long long)b * H + iy) * W + ix) * C4 + c;
        float4 q = ok[t] ? act_ld4<FMT>(x, idx[t] * 4) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        v[t][0] = q.x; v[t][1] = q.y; v[t][2] = q.z; v[t][3] = q.w;
    }
    int arg[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int a = 0; float m = v[0][q];
#pragma unroll
        for (int t = 1; t < 4; ++t) if (v[t][q] > m) { m = v[t][q]; a = t; }
        arg[q] = a;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        if (!ok[t]) continue;
        float g[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) g[q] = arg[q] == t ? gv[q] : 0.f;
        if (beta) { float4 o = act_ld4<FMT>(dx, idx[t] * 4); g[0] += o.x; g[1] += o.y; g[2] += o.z; g[3] += o.w; }
        if (relu_mask) {
#pragma unroll
            for (int q = 0; q < 4; ++q) g[q] = v[t][q] > 0.f ? g[q] : 0.f;
        }
        if (FMT == ACT_F32 && round_out) { g[0] = tf32_rn(g[0]); g[1] = tf32_rn(g[1]); g[2] = tf32_rn(g[2]); g[3] = tf32_rn(g[3]); }
        act_st4<FMT>(dx, idx[t] * 4, make_float4(g[0], g[1], g[2], g[3]));
    }
}

// one warp per pixel
template <int FMT>
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ scale, long long pixels, int C,
                                  int round_out, float* __restrict__ y) {
    long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (p >= pixels) return;
    float ss = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
        float4 v = act_ld4<FMT>(x, p * C + c);
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    float r = rsqrtf(fmaxf(ss, 1e-12f));
    for (int c = lane * 4; c < C; c += 128) {
        float4 v = act_ld4<FMT>(x, p * C + c);
        float4 s = *reinterpret_cast<const float4*>(scale + c);
        float4 o = make_float4(v.x * r * s.x, v.y * r * s.y, v.z * r * s.z, v.w * r * s.w);
        if (FMT == ACT_F32 && round_out) { o.x = tf32_rn(o.x); o.y = tf32_rn(o.y); o.z = tf32_rn(o.z); o.w = tf32_rn(o.w); }
        act_st4<FMT>(y, p * C + c, o);
    }
}

// dx_c = (s_c g_c r - x_c r^3 sum_k(g_k s_k x_k)) masked by x>0 ; dscale partial per block
template <int FMT>
__global__ void l2norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ dy,
                                  long long pixels, int C, int beta, int round_out, float* __restrict__ dx, float* __restrict__ partial) {
    extern __shared__ float sh[];   // [warps][C] dscale partials
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    float* mine = sh + (long long)warp * C;
    for (int c = lane; c < C; c += 32) mine[c] = 0.f;
    __syncwarp();
    for (long long p = (long long)blockIdx.x * nwarp + warp; p < pixels; p += (long long)gridDim.x * nwarp) {
        float ss = 0.f, dot = 0.f;
        for (int c = lane * 4; c < C; c += 128) {
            float4 v = act_ld4<FMT>(x, p * C + c);
            float4 g = act_ld4<FMT>(dy, p * C + c);
            float4 s = *reinterpret_cast<const float4*>(scale + c);
            ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            dot += g.x * s.x * v.x + g.y * s.y * v.y + g.z * s.z * v.z + g.w * s.w * v.w;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, o); dot += __shfl_xor_sync(0xffffffffu, dot, o); }
        bool clamped = ss < 1e-12f;
        float r = rsqrtf(fmaxf(ss, 1e-12f));
        float r3dot = clamped ? 0.f : r * r * r * dot;
        for (int c = lane * 4; c < C; c += 128) {
            float4 v = act_ld4<FMT>(x, p * C + c);
            float4 g = act_ld4<FMT>(dy, p * C + c);
            float4 s = *reinterpret_cast<const float4*>(scale + c);
            float vv[4] = {v.x, v.y, v.z, v.w}, gg[4] = {g.x, g.y, g.z, g.w}, sv[4] = {s.x, s.y, s.z, s.w};
            float o[4];
            float4 old = beta ? act_ld4<FMT>(dx, p * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            float ov[4] = {old.x, old.y, old.z, old.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float d = sv[q] * gg[q] * r - vv[q] * r3dot + ov[q];
                o[q] = vv[q] > 0.f ? d : 0.f;
                if (FMT == ACT_F32 && round_out) o[q] = tf32_rn(o[q]);
                mine[c + q] += gg[q] * vv[q] * r;
            }
            act_st4<FMT>(dx, p * C + c, make_float4(o[0], o[1], o[2], o[3]));
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nwarp; ++w) s += sh[(long long)w * C + c];
        partial[(long long)blockIdx.x * C + c] = s;
    }
}

__global__ void reduce_rows_kernel(const float* __restrict__ partial, int rows, int C, float* __restrict__ out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += partial[(long long)r * C + c];
    out[c] = s;
}

template <int FMT>
__global__ void head_grad_gather_kernel(const float* __restrict__ grad, int B, int A, int V, int anchor_base, int HW,
                                        int nbox, int Npad, int round_out, float* __restrict__ dz) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;        // four consecutive channels of one pixel
    long long total = (long long)B * HW * Npad;
    if (i >= total) return;
    int n0 = (int)(i % Npad); long long r = i / Npad;
    int pix = (int)(r % HW); int b = (int)(r / HW);
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int n = n0 + q;
        v[q] = 0.f;
        if (n < nbox * V) {
            int j = n / V, e = n - j * V;
            v[q] = grad[((long long)b * A + anchor_base + (long long)j * HW + pix) * V + e];
        }
        if (FMT == ACT_F32 && round_out) v[q] = tf32_rn(v[q]);
    }
    act_st4<FMT>(dz, i, make_float4(v[0], v[1], v[2], v[3]));
}

__global__ void round_tf32_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = tf32_rn(src[i]);
}

__global__ void split_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n4) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) act_st4<ACT_S32>(dst, i * 4, reinterpret_cast<const float4*>(src)[i]);
}
__global__ void unsplit_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n4) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) reinterpret_cast<float4*>(dst)[i] = act_ld4<ACT_S32>(src, i * 4);
}

// one thread per pixel builds the 27 taps (kh, kw, c) + 5 zeros = one 128-byte row; the warp then transposes its 32 rows
// through 4 KB of shared memory (16-byte pieces XOR-swizzled by row) so that every store instruction writes four complete
// rows -- one thread writing its own row put 16 bytes into each of 32 rows per instruction (1.9 TB/s); eight threads per
// pixel, each computing one piece, was slower still (index arithmetic: 1.1 ms).
template <int FMT>
__global__ void __launch_bounds__(128)
conv1_im2col_kernel(const float* __restrict__ img, int B, int S, int swap_rb, float m0, float m1, float m2, float* __restrict__ patches) {
    __shared__ uint4 stage[4][32 * 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long total = (long long)B * S * S;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) - lane;     // first pixel of this warp
    const long long i = warp0 + lane;
    uint4 piece[8];
    if (i < total) {
        int x = (int)(i % S); long long r = i / S;
        int y = (int)(r % S); int b = (int)(r / S);
        float v[32];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            int iy = y + t / 3 - 1, ix = x + t % 3 - 1;
            bool ok = iy >= 0 && iy < S && ix >= 0 && ix < S;
            const float* px = img + (((long long)b * S + (ok ? iy : 0)) * S + (ok ? ix : 0)) * 3;
            float p0 = __ldg(px), p1 = __ldg(px + 1), p2 = __ldg(px + 2);
            float c0 = (swap_rb ? p2 : p0) - m0, c1 = p1 - m1, c2 = (swap_rb ? p0 : p2) - m2;
            if (FMT == ACT_F32) { c0 = tf32_rn(c0); c1 = tf32_rn(c1); c2 = tf32_rn(c2); }
            v[t * 3 + 0] = ok ? c0 : 0.f; v[t * 3 + 1] = ok ? c1 : 0.f; v[t * 3 + 2] = ok ? c2 : 0.f;
        }
#pragma unroll
        for (int t = 27; t < 32; ++t) v[t] = 0.f;
        if (FMT == ACT_F32) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                piece[q] = make_uint4(__float_as_uint(v[q * 4]), __float_as_uint(v[q * 4 + 1]), __float_as_uint(v[q * 4 + 2]), __float_as_uint(v[q * 4 + 3]));
        } else {
            // the pixel's 128-byte row: 32 bf16 high parts, then 32 bf16 low parts
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) split2(v[2 * q], v[2 * q + 1], hi[q], lo[q]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                piece[q] = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
                piece[4 + q] = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) stage[warp][lane * 8 + (q ^ (lane & 7))] = piece[q];
    }
    __syncwarp();
    const int q = lane & 7;
    uint4* out = reinterpret_cast<uint4*>(patches);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int row = k * 4 + (lane >> 3);
        if (warp0 + row < total) out[(warp0 + row) * 8 + q] = stage[warp][row * 8 + (q ^ (row & 7))];
    }
}

__global__ void conv1_pad_filter_kernel(const float* __restrict__ w27, int Cout, float* __restrict__ w32) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 32 * Cout) return;
    w32[i] = i < 27 * Cout ? w27[i] : 0.f;
}

// one thread per row: softmax over the first C+1 columns, copy the 4 offsets
__global__ void softmax_result_kernel(const float* __restrict__ out, long long rows, int C, float* __restrict__ res) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    int V = C + 5, nc = C + 1;
    const float* z = out + r * V;
    float m = z[0];
    for (int c = 1; c < nc; ++c) m = fmaxf(m, z[c]);
    float s = 0.f;
    for (int c = 0; c < nc; ++c) s += expf(z[c] - m);
    float inv = 1.f / s;
    float* o = res + r * V;
    for (int c = 0; c < nc; ++c) o[c] = expf(z[c] - m) * inv;
    for (int c = nc; c < V; ++c) o[c] = z[c];
}

__global__ void sgd_kernel(float* __restrict__ w, float* __restrict__ g, float* __restrict__ v, long long n,
                           const unsigned char* __restrict__ decay, float lr, float mu, float wd, float post) {
    long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long i = i4 * 4;
    if (i >= n) return;
    float d = decay[i / OPT_BLOCK] ? wd : 0.f;
    float4 wv = reinterpret_cast<float4*>(w)[i4];
    float4 gv = reinterpret_cast<float4*>(g)[i4];
    float4 vv = reinterpret_cast<float4*>(v)[i4];
    float ww[4] = {wv.x, wv.y, wv.z, wv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w}, mm[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float gr = gg[q] * post + d * ww[q];
        mm[q] = mu * mm[q] + gr;
        ww[q] = ww[q] - lr * mm[q];
    }
    reinterpret_cast<float4*>(w)[i4] = make_float4(ww[0], ww[1], ww[2], ww[3]);
    reinterpret_cast<float4*>(v)[i4] = make_float4(mm[0], mm[1], mm[2], mm[3]);
}

__global__ void l2_sum_stage1(const float* __restrict__ w, long long n, const unsigned char* __restrict__ decay,
                              float* __restrict__ partial) {
    __shared__ float sh[32];
    float s = 0.f;
    for (long long blk = blockIdx.x; blk * OPT_BLOCK < n; blk += gridDim.x) {
        if (!decay[blk]) continue;
        long long base = blk * OPT_BLOCK;
        for (int j = threadIdx.x; j < OPT_BLOCK && base + j < n; j += blockDim.x) { float t = w[base + j]; s += t * t; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
        partial[blockIdx.x] = t;
    }
}

__global__ void l2_sum_stage2(const float* __restrict__ partial, int nb, float* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < nb; ++i) t += partial[i];
        out[0] = 0.5f * t;
    }
}

}  // namespace

#define FMT_LAUNCH(kernel, fmt, grid, block, shm, st, ...)                                      \
    do {                                                                                         \
        if ((fmt) == ACT_S32) kernel<ACT_S32><<<grid, block, shm, st>>>(__VA_ARGS__);           \
        else kernel<ACT_F32><<<grid, block, shm, st>>>(__VA_ARGS__);                            \
        SSDB_LAUNCH_CHECK();                                                                     \
    } while (0)

int maxpool_fwd(const float* x, int fmt, int B, int H, int W, int C, int k, int stride, int pad_t, int pad_l, int Ho, int Wo,
                float* y, cudaStream_t st) {
    SSDB_REQUIRE(C % 4 == 0, "channels must be a multiple of 4");
    long long total = (long long)B * Ho * Wo * (C / 4);
    FMT_LAUNCH(maxpool_fwd_kernel, fmt, (unsigned)((total + 255) / 256), 256, 0, st, x, B, H, W, C / 4, k, stride, pad_t, pad_l, Ho, Wo, y);
    return SSDB_OK;
}

int maxpool_fwd_arg(const float* x, int fmt, int B, int H, int W, int C, int k, int stride, int pad_t, int pad_l, int Ho, int Wo,
                    float* y, unsigned char* arg, cudaStream_t st) {
    SSDB_REQUIRE(C % 4 == 0 && k * k < 255, "unsupported pool");
    long long total = (long long)B * Ho * Wo * (C / 4);
    FMT_LAUNCH(maxpool_fwd_arg_kernel, fmt, (unsigned)((total + 255) / 256), 256, 0, st, x, B, H, W, C / 4, k, stride, pad_t, pad_l, Ho, Wo, y,
               reinterpret_cast<uchar4*>(arg));
    return SSDB_OK;
}

int maxpool_bwd_arg(const float* x, const float* dy, const unsigned char* arg, int fmt, int B, int H, int W, int C, int k, int stride, int pad_t,
                    int pad_l, int Ho, int Wo, int beta, int relu_mask, int round_out, float* dx, cudaStream_t st) {
    SSDB_REQUIRE(C % 4 == 0, "channels must be a multiple of 4");
    long long total = (long long)B * H * W * (C / 4);
    FMT_LAUNCH(maxpool_bwd_arg_kernel, fmt, (unsigned)((total + 255) / 256), 256, 0, st, x, dy, reinterpret_cast<const uchar4*>(arg), B, H, W, C / 4, k,
               stride, pad_t, pad_l, Ho, Wo, beta, relu_mask, round_out, dx);
    return SSDB_OK;
}

int maxpool2x2_fwd_code(const float* x, int fmt, int B, int H, int W, int C, int Ho, int Wo, float* y, unsigned char* code, cudaStream_t st) {
    SSDB_REQUIRE(C % 4 == 0 && Ho == (H + 1) / 2 && Wo == (W + 1) / 2, "2x2/s2 SAME pool with pad_before 0 expected");
    long long total = (long long)B * Ho * Wo * (C / 4);
    FMT_LAUNCH(maxpool2x2_fwd_code_kernel, fmt, (unsigned)((total + 255) / 256), 256, 0, st, x, B, H, W, C / 4, Ho, Wo, y, reinterpret_cast<uchar4*>(code));
    return SSDB_OK;
}

int maxpool2x2_bwd_code(const float* dy, const unsigned char* code, int fmt, int B, int H, int W, int C, int Ho, int Wo, int relu_mask,
                        int round_out, float* dx, cudaStream_t st) {
    SSDB_REQUIRE(C % 4 == 0 && Ho == (H + 1) / 2 && Wo == (W + 1) / 2, "2x2/s2 SAME pool with pad_before 0 expected");
    long long total = (long long)B * Ho * Wo * (C / 4);
    FMT_LAUNCH(maxpool2x2_bwd_code_kernel, fmt, (unsigned)((total + 255) / 256), 256, 0, st, dy, reinterpret_cast<const uchar4*>(code), B, H, W, C / 4, Ho,
               Wo, relu_mask, round_out, dx);
    return SSDB_OK;
}

int maxpool_bwd(const float* x, const float* dy, int fmt, int B, int H, int W, int C, int k, int stride, int pad_t, int pad_l,
                int Ho, int Wo, int beta, int relu_mask, int round_out, float* dx, cudaStream_t st) {
    SSDB_REQUIRE(C % 4 == 0, "channels must be a multiple of 4");
    if (k == 2 && stride == 2 && pad_t == 0 && pad_l == 0) {
        long long tw = (long long)B * Ho * Wo * (C / 4);
        FMT_LAUNCH(maxpool2x2_bwd_kernel, fmt, (unsigned)((tw + 255) / 256), 256, 0, st, x, dy, B, H, W, C / 4, Ho, Wo, beta, relu_mask, round_out, dx);
        return SSDB_OK;
    }
    long long total = (long long)B * H * W * (C / 4);
    FMT_LAUNCH(maxpool_bwd_kernel, fmt, (unsigned)((total + 255) / 256), 256, 0, st, x, dy, B, H, W, C / 4, k, stride, pad_t, pad_l, Ho, Wo,
               beta, relu_mask, round_out, dx);
    return SSDB_OK;
}

int l2norm_fwd(const float* x, const float* scale, int fmt, long long pixels, int C, int round_out, float* y, cudaStream_t st) {
    SSDB_REQUIRE(C % 4 == 0, "channels must be a multiple of 4");
    long long threads = pixels * 32;
    FMT_LAUNCH(l2norm_fwd_kernel, fmt, (unsigned)((threads + 255) / 256), 256, 0, st, x, scale, pixels, C, round_out, y);
    return SSDB_OK;
}

int l2norm_bwd(const float* x, const float* scale, const float* dy, int fmt, long long pixels, int C, int beta, int round_out, float* dx,
               float* dscale, float* partial, cudaStream_t st) {
    SSDB_REQUIRE(C % 4 == 0 && C <= 1024, "unsupported channel count");
    int nb = 296;   // 2 blocks per SM
    size_t sh = (size_t)8 * C * sizeof(float);
    FMT_LAUNCH(l2norm_bwd_kernel, fmt, nb, 256, sh, st, x, scale, dy, pixels, C, beta, round_out, dx, partial);
    reduce_rows_kernel<<<(C + 255) / 256, 256, 0, st>>>(partial, nb, C, dscale);
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

int head_grad_gather(const float* grad, int B, int A, int V, int anchor_base, int HW, int nbox, int Npad, int fmt, int round_out, float* dz,
                     cudaStream_t st) {
    SSDB_REQUIRE(Npad % 4 == 0, "padded head channels must be a multiple of 4");
    long long total = (long long)B * HW * Npad / 4;
    FMT_LAUNCH(head_grad_gather_kernel, fmt, (unsigned)((total + 255) / 256), 256, 0, st, grad, B, A, V, anchor_base, HW, nbox, Npad, round_out, dz);
    return SSDB_OK;
}

int round_tf32_copy(const float* src, float* dst, long long n, cudaStream_t st) {
    round_tf32_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, dst, n);
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

int split_copy(const float* src, float* dst_s32, long long n, cudaStream_t st) {
    SSDB_REQUIRE(n % 32 == 0, "split copies work on whole 32-element groups");
    split_copy_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(src, dst_s32, n / 4);
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

int unsplit_copy(const float* src_s32, float* dst, long long n, cudaStream_t st) {
    SSDB_REQUIRE(n % 32 == 0, "split copies work on whole 32-element groups");
    unsplit_copy_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(src_s32, dst, n / 4);
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

int conv1_im2col(const float* images, int B, int S, int swap_rb, const float mean[3], int fmt, float* patches, cudaStream_t st) {
    long long total = (long long)B * S * S;
    FMT_LAUNCH(conv1_im2col_kernel, fmt, (unsigned)((total + 127) / 128), 128, 0, st, images, B, S, swap_rb, mean[0], mean[1], mean[2], patches);
    return SSDB_OK;
}

int conv1_pad_filter(const float* w27, int Cout, float* w32, cudaStream_t st) {
    conv1_pad_filter_kernel<<<(32 * Cout + 255) / 256, 256, 0, st>>>(w27, Cout, w32);
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

int softmax_result(const float* output, long long rows, int C, float* result, cudaStream_t st) {
    softmax_result_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(output, rows, C, result);
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

int sgd_momentum(float* w, float* g, float* v, long long n, const unsigned char* decay, float lr, float mu, float wd,
                 float post_scale, cudaStream_t st) {
    SSDB_REQUIRE(n % 4 == 0, "flat buffer length must be a multiple of 4");
    long long n4 = n / 4;
    sgd_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(w, g, v, n, decay, lr, mu, wd, post_scale);
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

int l2_sum(const float* w, long long n, const unsigned char* decay, float* partial, float* out, cudaStream_t st) {
    int nb = 592;
    l2_sum_stage1<<<nb, 256, 0, st>>>(w, n, decay, partial);
    SSDB_LAUNCH_CHECK();
    l2_sum_stage2<<<1, 32, 0, st>>>(partial, nb, out);
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

}  // namespace ssdb
