// CUDA-core implicit-GEMM convolution (fprop / dgrad / wgrad) for every shape of
// the SSD-VGG graph (reference ssdvgg.py:42-65,231-332 -- tf.nn.conv2d,
// atrous_conv2d, bias_add, relu and their TF gradients).
//
// The engine uses these kernels where the tcgen05 path does not apply: conv1_1
// (Cin = 3, K = 27, bandwidth bound), the stride-2 extras and the tiny tail
// layers (M < one MMA tile).  They are also the on-device cross-check for the
// tensor-core kernels in tests.  64x64x16 tiles, 256 threads, 4x4 micro-tile.
#include "common.cuh"

namespace ssdb {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

enum Mode { FPROP = 0, DGRAD = 1, WGRAD = 2 };

struct SimtArgs {
    ConvGeom g;
    const float* a;        // fprop: x, dgrad: dz, wgrad: x
    const float* b;        // fprop: w, dgrad: w,  wgrad: dz
    float* out;            // fprop: y, dgrad: dx, wgrad: partial
    const float* aux;      // fprop: bias, dgrad: mask_x
    int relu, beta, round_out;
    int scatter, V, n_valid, anchor_base, A;
    int preprocess, swap_rb;
    float mean0, mean1, mean2;
    long long M, K;        // GEMM M and reduction length
    int N;
    long long red_per_split;   // wgrad: pixels per z-slice
};

__device__ __forceinline__ float prep(const SimtArgs& p, const float* px, int c) {
    int cs = p.swap_rb ? 2 - c : c;
    float m = c == 0 ? p.mean0 : (c == 1 ? p.mean1 : p.mean2);
    return px[cs] - m;
}

// value(s) of the im2col matrix: input pixel for output position (b,oy,ox), tap, channel c..c+VEC-1
template <int VEC, int FMT>
__device__ __forceinline__ void gather_x(const SimtArgs& p, int b, int oy, int ox, int tap, int c, float* v) {
    const ConvGeom& g = p.g;
    int kh = tap / g.k, kw = tap - kh * g.k;
    int iy = oy * g.stride + kh * g.dil - g.pad_t;
    int ix = ox * g.stride + kw * g.dil - g.pad_l;
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = 0.f;
    if (iy < 0 || iy >= g.H || ix < 0 || ix >= g.W) return;
    const long long e = (((long long)b * g.H + iy) * g.W + ix) * g.Cin;
    if (VEC == 4) {
        float4 t = act_ld4<FMT>(p.a, e + c);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        v[0] = p.preprocess ? prep(p, p.a + e, c) : act_ld1<FMT>(p.a, e + c);    // the raw image (preprocess) is plain float32
    }
}

template <int MODE, int VEC, int FMT>
__global__ void __launch_bounds__(NT) conv_simt_kernel(SimtArgs p) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const ConvGeom& g = p.g;
    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    long long red_begin = 0, red_end = p.K;
    if (MODE == WGRAD) {
        red_begin = (long long)blockIdx.z * p.red_per_split;
        red_end = min(p.K, red_begin + p.red_per_split);
    }

    // per-thread fixed coordinates
    int a_b = 0, a_y = 0, a_x = 0;       // FPROP / DGRAD: pixel of this thread's A row
    bool a_row_ok = false;
    if (MODE == FPROP || MODE == DGRAD) {
        long long m = m0 + (t >> 2);
        a_row_ok = m < p.M;
        if (a_row_ok) {
            int hw = (MODE == FPROP) ? g.Ho * g.Wo : g.H * g.W;
            int wd = (MODE == FPROP) ? g.Wo : g.W;
            a_b = (int)(m / hw);
            int r = (int)(m - (long long)a_b * hw);
            a_y = r / wd; a_x = r - a_y * wd;
        }
    }

    for (long long k0 = red_begin; k0 < red_end; k0 += BK) {
        // ---------------- A tile ----------------
        if (MODE == FPROP) {
            int a_k = (t & 3) * 4;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            long long kk = k0 + a_k;
            if (a_row_ok) {
                if (VEC == 4) {
                    if (kk < p.K) { int tap = (int)(kk / g.Cin); int c = (int)(kk - (long long)tap * g.Cin); gather_x<4, FMT>(p, a_b, a_y, a_x, tap, c, v); }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (kk + i < p.K) { int tap = (int)((kk + i) / g.Cin); int c = (int)(kk + i - (long long)tap * g.Cin); gather_x<1, FMT>(p, a_b, a_y, a_x, tap, c, &v[i]); }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) As[a_k + i][t >> 2] = v[i];
        } else if (MODE == DGRAD) {
            int a_k = (t & 3) * 4;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            long long kk = k0 + a_k;
            if (a_row_ok && kk < p.K) {
                int tap = (int)(kk / g.Cout); int n = (int)(kk - (long long)tap * g.Cout);
                int kh = tap / g.k, kw = tap - kh * g.k;
                int ny = a_y + g.pad_t - kh * g.dil, nx = a_x + g.pad_l - kw * g.dil;
                if (ny >= 0 && nx >= 0 && (ny % g.stride) == 0 && (nx % g.stride) == 0) {
                    int oy = ny / g.stride, ox = nx / g.stride;
                    if (oy < g.Ho && ox < g.Wo) {
                        float4 tv = act_ld4<FMT>(p.a, (((long long)a_b * g.Ho + oy) * g.Wo + ox) * g.Cout + n);
                        v[0] = tv.x; v[1] = tv.y; v[2] = tv.z; v[3] = tv.w;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) As[a_k + i][t >> 2] = v[i];
        } else {  // WGRAD: A[m = (tap,c)][p]
            int a_p = t >> 4, a_m = (t & 15) * 4;
            long long pix = k0 + a_p;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (pix < red_end) {
                int hw = g.Ho * g.Wo;
                int b = (int)(pix / hw); int r = (int)(pix - (long long)b * hw);
                int oy = r / g.Wo, ox = r - oy * g.Wo;
                long long mm = m0 + a_m;
                if (VEC == 4) {
                    if (mm < p.M) { int tap = (int)(mm / g.Cin); int c = (int)(mm - (long long)tap * g.Cin); gather_x<4, FMT>(p, b, oy, ox, tap, c, v); }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (mm + i < p.M) { int tap = (int)((mm + i) / g.Cin); int c = (int)(mm + i - (long long)tap * g.Cin); gather_x<1, FMT>(p, b, oy, ox, tap, c, &v[i]); }
                }
            }
            *reinterpret_cast<float4*>(&As[a_p][a_m]) = make_float4(v[0], v[1], v[2], v[3]);
        }
        // ---------------- B tile ----------------
        if (MODE == FPROP) {
            int b_k = t >> 4, b_n = (t & 15) * 4;
            long long kk = k0 + b_k;
            float4 tv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kk < p.K && n0 + b_n < p.N) tv = *reinterpret_cast<const float4*>(p.b + kk * g.Cout + n0 + b_n);
            *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = tv;
        } else if (MODE == DGRAD) {
            int b_c = t >> 2, b_k = (t & 3) * 4;
            long long kk = k0 + b_k;
            float4 tv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kk < p.K && n0 + b_c < p.N) {
                int tap = (int)(kk / g.Cout); int n = (int)(kk - (long long)tap * g.Cout);
                tv = *reinterpret_cast<const float4*>(p.b + ((long long)tap * g.Cin + n0 + b_c) * g.Cout + n);
            }
            Bs[b_k + 0][b_c] = tv.x; Bs[b_k + 1][b_c] = tv.y; Bs[b_k + 2][b_c] = tv.z; Bs[b_k + 3][b_c] = tv.w;
        } else {
            int b_p = t >> 4, b_n = (t & 15) * 4;
            long long pix = k0 + b_p;
            float4 tv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pix < red_end && n0 + b_n < p.N) tv = act_ld4<FMT>(p.b, pix * g.Cout + n0 + b_n);
            *reinterpret_cast<float4*>(&Bs[b_p][b_n]) = tv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float a4[4] = {av.x, av.y, av.z, av.w};
            float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---------------- epilogue ----------------
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        long long m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
        if (MODE == FPROP) {
            if (p.scatter) {
                int hw = g.Ho * g.Wo;
                int b = (int)(m / hw); int pix = (int)(m - (long long)b * hw);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int n = n0 + tx * 4 + j;
                    if (n >= p.n_valid) continue;
                    int bt = n / p.V, v = n - bt * p.V;
                    float r = acc[i][j] + (p.aux ? p.aux[n] : 0.f);
                    p.out[((long long)b * p.A + p.anchor_base + (long long)bt * hw + pix) * p.V + v] = r;
                }
            } else {
                int n = n0 + tx * 4;
                if (n < p.N) {
                    float r[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        r[j] = acc[i][j] + (p.aux ? p.aux[n + j] : 0.f);
                        if (p.relu) r[j] = fmaxf(r[j], 0.f);
                        if (FMT == ACT_F32 && p.round_out) r[j] = tf32_rn(r[j]);
                    }
                    act_st4<FMT>(p.out, m * g.Cout + n, make_float4(r[0], r[1], r[2], r[3]));
                }
            }
        } else if (MODE == DGRAD) {
            int n = n0 + tx * 4;
            if (n < p.N) {
                const long long e = m * g.Cin + n;
                float r[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
                if (p.beta) { float4 o = act_ld4<FMT>(p.out, e); r[0] += o.x; r[1] += o.y; r[2] += o.z; r[3] += o.w; }
                if (p.aux) {
                    float4 mk = act_ld4_sign<FMT>(p.aux, e);
                    r[0] = mk.x > 0.f ? r[0] : 0.f; r[1] = mk.y > 0.f ? r[1] : 0.f;
                    r[2] = mk.z > 0.f ? r[2] : 0.f; r[3] = mk.w > 0.f ? r[3] : 0.f;
                }
                if (FMT == ACT_F32 && p.round_out) { r[0] = tf32_rn(r[0]); r[1] = tf32_rn(r[1]); r[2] = tf32_rn(r[2]); r[3] = tf32_rn(r[3]); }
                act_st4<FMT>(p.out, e, make_float4(r[0], r[1], r[2], r[3]));
            }
        } else {
            int n = n0 + tx * 4;
            if (n < p.N) {
                float* dst = p.out + ((long long)blockIdx.z * p.M + m) * g.Cout + n;
                *reinterpret_cast<float4*>(dst) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            }
        }
    }
}

// out[i] = sum_z partial[z][i]
__global__ void reduce_splits_kernel(const float* __restrict__ partial, long long n, int splits, float* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partial[(long long)z * n + i];
    out[i] = s;
}

// stage 1 of the bias gradient (column sums of dz[pixels][C]): a block owns a contiguous slab of
// pixel rows; threads are (C/4 float4 columns) x (256/(C/4) row lanes), 4 independent 16-byte loads
// in flight per thread, shared-memory reduction over the row lanes, one partial row per block.
template <int FMT>
__global__ void __launch_bounds__(256) bias_grad_stage1(const float* __restrict__ dz, long long pixels, int C,
                                                         long long rows_per_block, float* __restrict__ partial) {
    __shared__ float4 red[256];
    const int c4n = C >> 2;                       // float4 columns (<= 256)
    const int lanes = 256 / c4n;                  // row lanes per block
    const int col = threadIdx.x % c4n, lane = threadIdx.x / c4n;
    float4 acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < lanes) {
        const long long r0 = (long long)blockIdx.x * rows_per_block;
        const long long r1 = min(pixels, r0 + rows_per_block);
        long long r = r0 + lane;
        for (; r + 3LL * lanes < r1; r += 4LL * lanes) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float4 v = act_ld4<FMT>(dz, ((r + (long long)u * lanes) * c4n + col) * 4);
                acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
            }
        }
        for (; r < r1; r += lanes) {
            float4 v = act_ld4<FMT>(dz, (r * c4n + col) * 4);
            acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w;
        }
    }
    float4 s = make_float4(acc[0].x + acc[1].x + acc[2].x + acc[3].x, acc[0].y + acc[1].y + acc[2].y + acc[3].y,
                           acc[0].z + acc[1].z + acc[2].z + acc[3].z, acc[0].w + acc[1].w + acc[2].w + acc[3].w);
    red[threadIdx.x] = s;
    __syncthreads();
    if (lane == 0 && threadIdx.x < c4n) {
        for (int l = 1; l < lanes; ++l) { float4 o = red[l * c4n + col]; s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w; }
        reinterpret_cast<float4*>(partial + (long long)blockIdx.x * C)[col] = s;
    }
}

int wgrad_splits(const ConvGeom& g) {
    long long pixels = (long long)g.B * g.Ho * g.Wo;
    long long tiles = ((long long)g.k * g.k * g.Cin + BM - 1) / BM * ((g.Cout + BN - 1) / BN);
    long long want = (148LL * 4 + tiles - 1) / tiles;          // ~4 CTAs per SM
    long long maxs = (pixels + 4 * BK - 1) / (4 * BK);
    long long s = want < 1 ? 1 : want;
    if (s > maxs) s = maxs;
    if (s > 512) s = 512;
    if (s < 1) s = 1;
    return (int)s;
}

void fill_common(SimtArgs& p, const ConvGeom& g) {
    p.g = g; p.relu = 0; p.beta = 0; p.round_out = 0; p.scatter = 0; p.V = 0; p.n_valid = 0; p.anchor_base = 0; p.A = 0;
    p.preprocess = 0; p.swap_rb = 0; p.mean0 = p.mean1 = p.mean2 = 0.f; p.aux = nullptr; p.red_per_split = 0;
}

}  // namespace

int conv_simt_fprop(const ConvGeom& g, const float* x, const float* w, int fmt, const ConvEpilogue& ep, float* y, cudaStream_t st) {
    SSDB_REQUIRE(g.Cout % 4 == 0, "Cout must be a multiple of 4");
    SSDB_REQUIRE(fmt == ACT_F32 || ep.scatter || g.Cout % 32 == 0, "split activations need channel counts that are multiples of 32");
    SimtArgs p; fill_common(p, g);
    p.a = x; p.b = w; p.out = y; p.aux = ep.bias; p.relu = ep.relu; p.round_out = ep.round_tf32;
    p.scatter = ep.scatter; p.V = ep.V; p.n_valid = ep.n_valid; p.anchor_base = ep.anchor_base; p.A = ep.A;
    p.preprocess = ep.preprocess; p.swap_rb = ep.swap_rb; p.mean0 = ep.mean[0]; p.mean1 = ep.mean[1]; p.mean2 = ep.mean[2];
    p.M = (long long)g.B * g.Ho * g.Wo; p.K = (long long)g.k * g.k * g.Cin; p.N = g.Cout;
    dim3 grid((unsigned)((p.M + BM - 1) / BM), (unsigned)((p.N + BN - 1) / BN));
    if (g.Cin % 4 == 0 && !ep.preprocess) {
        if (fmt == ACT_S32) conv_simt_kernel<FPROP, 4, ACT_S32><<<grid, NT, 0, st>>>(p);
        else conv_simt_kernel<FPROP, 4, ACT_F32><<<grid, NT, 0, st>>>(p);
    } else {
        if (fmt == ACT_S32) conv_simt_kernel<FPROP, 1, ACT_S32><<<grid, NT, 0, st>>>(p);
        else conv_simt_kernel<FPROP, 1, ACT_F32><<<grid, NT, 0, st>>>(p);
    }
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

int conv_simt_dgrad(const ConvGeom& g, const float* dz, const float* w, int fmt, const float* mask_x, int beta, int round_out, float* dx, cudaStream_t st) {
    SSDB_REQUIRE(g.Cout % 4 == 0 && g.Cin % 4 == 0, "channels must be multiples of 4");
    SimtArgs p; fill_common(p, g);
    p.a = dz; p.b = w; p.out = dx; p.aux = mask_x; p.beta = beta; p.round_out = round_out;
    p.M = (long long)g.B * g.H * g.W; p.K = (long long)g.k * g.k * g.Cout; p.N = g.Cin;
    dim3 grid((unsigned)((p.M + BM - 1) / BM), (unsigned)((p.N + BN - 1) / BN));
    if (fmt == ACT_S32) conv_simt_kernel<DGRAD, 4, ACT_S32><<<grid, NT, 0, st>>>(p);
    else conv_simt_kernel<DGRAD, 4, ACT_F32><<<grid, NT, 0, st>>>(p);
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

size_t conv_simt_wgrad_ws(const ConvGeom& g) {
    return (size_t)wgrad_splits(g) * g.k * g.k * g.Cin * g.Cout;
}

int conv_simt_wgrad(const ConvGeom& g, const float* x, const float* dz, int fmt, const ConvEpilogue& ep, float* dw, float* partial, cudaStream_t st) {
    SSDB_REQUIRE(g.Cout % 4 == 0, "Cout must be a multiple of 4");
    SimtArgs p; fill_common(p, g);
    p.a = x; p.b = dz; p.out = partial;
    p.preprocess = ep.preprocess; p.swap_rb = ep.swap_rb; p.mean0 = ep.mean[0]; p.mean1 = ep.mean[1]; p.mean2 = ep.mean[2];
    p.M = (long long)g.k * g.k * g.Cin; p.K = (long long)g.B * g.Ho * g.Wo; p.N = g.Cout;
    int splits = wgrad_splits(g);
    long long per = (p.K + splits - 1) / splits;
    per = (per + BK - 1) / BK * BK;
    splits = (int)((p.K + per - 1) / per);
    p.red_per_split = per;
    dim3 grid((unsigned)((p.M + BM - 1) / BM), (unsigned)((p.N + BN - 1) / BN), (unsigned)splits);
    if (g.Cin % 4 == 0 && !ep.preprocess) {
        if (fmt == ACT_S32) conv_simt_kernel<WGRAD, 4, ACT_S32><<<grid, NT, 0, st>>>(p);
        else conv_simt_kernel<WGRAD, 4, ACT_F32><<<grid, NT, 0, st>>>(p);
    } else {
        if (fmt == ACT_S32) conv_simt_kernel<WGRAD, 1, ACT_S32><<<grid, NT, 0, st>>>(p);
        else conv_simt_kernel<WGRAD, 1, ACT_F32><<<grid, NT, 0, st>>>(p);
    }
    SSDB_LAUNCH_CHECK();
    long long n = p.M * g.Cout;
    reduce_splits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partial, n, splits, dw);
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

int bias_grad(const float* dz, int fmt, long long pixels, int Cout, float* db, float* partial, cudaStream_t st) {
    SSDB_REQUIRE(Cout % 4 == 0 && Cout <= 1024, "bias gradient needs Cout % 4 == 0 and Cout <= 1024");
    long long want = 148 * 8;
    long long rows_per_block = (pixels + want - 1) / want;
    if (rows_per_block < 32) rows_per_block = 32;
    int nb = (int)((pixels + rows_per_block - 1) / rows_per_block);
    if (nb < 1) nb = 1;
    if (fmt == ACT_S32) bias_grad_stage1<ACT_S32><<<nb, 256, 0, st>>>(dz, pixels, Cout, rows_per_block, partial);
    else bias_grad_stage1<ACT_F32><<<nb, 256, 0, st>>>(dz, pixels, Cout, rows_per_block, partial);
    SSDB_LAUNCH_CHECK();
    reduce_splits_kernel<<<(Cout + 255) / 256, 256, 0, st>>>(partial, Cout, nb, db);
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

}  // namespace ssdb
