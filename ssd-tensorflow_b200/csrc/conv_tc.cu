// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (tf32 operands, fp32
// accumulation in tensor memory).  Replaces the cuDNN/Eigen convolutions behind
// tf.nn.conv2d / atrous_conv2d in the reference graph (ssdvgg.py:42-65,231-332)
// and their TF gradients (MomentumOptimizer.minimize, ssdvgg.py:585-588).
//
// fprop and dgrad are the same contraction -- "for every filter tap, gather the
// source pixels at a tap-dependent shift and contract over source channels" -- so
// one kernel serves both:
//     fprop: src = x  [B,H,W,Cin],   dst = y  [B,Ho,Wo,Cout], shift = tap*dil - pad
//     dgrad: src = dz [B,Ho,Wo,Cout],dst = dx [B,H,W,Cin],    shift = pad - tap*dil
//
// Mapping to the hardware
//   * M tile = 128 destination pixels arranged as a TW x TH x TN box (x, y, image)
//     chosen per layer so that TW*TH*TN <= 128 wastes the fewest rows.  A 4-D TMA
//     tensor map over the NHWC source loads the shifted box for one tap and one
//     32-channel slab: [rows][32 x fp32] = 128-byte rows, SWIZZLE_128B, which is
//     exactly the K-major UMMA operand layout; padding comes from TMA zero fill.
//   * B operand = filter rows (destination channels) x 32 source channels, K-major,
//     from a 2-D tensor map over [taps*rows][Csrc]  (fprop: per-tap transposed copy
//     of the HWIO filter; dgrad: the HWIO filter itself).
//   * tcgen05.mma.cta_group::1.kind::tf32, M=128, N<=256, K=8 per instruction, 4 per
//     32-channel slab, accumulating over taps x slabs into one of two TMEM buffers.
//   * warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue
//     (tcgen05.ld -> bias/ReLU/mask -> global).  4-stage smem ring, persistent CTAs.
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace ssdb {
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;                    // fp32 elements = 128 bytes = one swizzle row
constexpr int MAX_N = 256;
constexpr int MAX_STAGES = 4;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;   // 16 KB per M tile
constexpr int B_BYTES = MAX_N * BLOCK_K * 4;     // 32 KB
// A unit is `mtu` (1 or 2) M tiles sharing one B tile.  mtu = 1: 4 stages x 48 KB, two TMEM accumulator buffers (the
// epilogue overlaps the next tile).  mtu = 2: 3 stages x 64 KB, 8 MMAs per stage, 1.5x less L2 -> smem traffic per FLOP,
// but with N = 256 both accumulators fill TMEM and the epilogue is exposed.  Measured on B200: mtu = 2 pays off in the
// wgrad kernel for N >= 128 (long units), not in fprop / dgrad, whose mid layers already run at ~88% of the tf32 peak.
constexpr int RING_BYTES = 192 * 1024;           // 4 x 48 KB, 3 x 64 KB, or 2 x 88 KB (row-window mode, N = 128, two M tiles)
constexpr int ACC_STRIDE = 256;                  // TMEM columns per accumulator buffer
constexpr int EPI_WARPS_MAX = 8;
constexpr int EPI_STAGE_BYTES = EPI_WARPS_MAX * 32 * 128;   // per-warp 32 x 32 fp32 transpose buffers of the epilogue
constexpr int SMEM_BYTES = RING_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_STAGE_BYTES;
__host__ __device__ constexpr int stage_bytes_for(int mtu) { return mtu * A_BYTES + B_BYTES; }
__host__ __device__ constexpr int stages_for(int mtu) { return mtu == 1 ? 4 : 3; }
constexpr int NUM_THREADS = 192;                 // TMA warp + MMA warp + 4 epilogue warps (wgrad kernels, two-CTA fprop / dgrad)
constexpr int TC_THREADS = 320;                  // fprop / dgrad: 8 epilogue warps, two per TMEM lane quarter, alternating 32-column chunks.
                                                 // With N = 256 and two M tiles both accumulators fill TMEM and the epilogue is exposed
                                                 // (dgrad: + mask reads): twice the warps drain it in half the time.
constexpr int TMEM_COLS = 512;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) { printf("ssdb conv_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// one lane of a converged warp; the compiler knows the branch is taken by a single thread (no per-lane serialisation
// loop around the uniform-datapath instructions UTMALDG / UTCHMMA that `lane == 0` would get)
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xFFFFFFFF;\n\t@P1 mov.s32 %0, 1;\n\t}" : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f16 with bf16 operands (K = 16 per instruction): the split-operand mode
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- 2-CTA clusters (cta_group::2): helpers shared by the pair kernels
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_5d_c2(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tc_commit_c2(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_c2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tma_load_4d_c2(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d_c2(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"):
// 8-row groups 1024 B apart (SBO), rows 128 B apart, 16-byte units.
// Measured on B200 (round-1 probe, recorded in DESIGN.md): the 128-byte swizzle is applied to the ABSOLUTE shared-memory address, so a
// descriptor may start at any 128-byte row of a TMA-written box and use any row-multiple SBO; the base-offset field must
// stay 0 (setting it to the start's row phase gives garbage).
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr, uint32_t sbo_bytes = 1024, int use_base_offset = 0) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;                      // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;   // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
    if (use_base_offset) d |= (uint64_t)((smem_addr >> 7) & 7) << 49; // base offset: swizzle phase of a start that is not 1024-byte aligned
    d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
    return d;
}

// ------------------------------------------------------------------ kernel
struct TcArgs {
    // destination tile box and grid
    int TW, TH, TN;                 // rows of a tile = TW*TH*TN <= 128
    int tiles_x, tiles_y, tiles_n;  // destination tiling
    int n_tiles, block_n;           // channel tiling of the destination
    int Hd, Wd, Bn, Cd;             // destination extents (Cd = stored channel stride)
    int cd_valid;                   // channels actually written
    // Filter taps come in groups that share ONE source box: plain mode = one tap per group (box = the destination box
    // shifted by the tap); row-window mode (3x3, dilation 1, TW = 8) = the three taps of a filter row share a box that is
    // two pixels wider, and tap t reads the rows starting `g_win` pixels into it (an MMA descriptor with a 128-byte row
    // offset and SBO = box width * 128) -- one third of the A traffic from L2.
    int ngroups;
    signed char g_dy[9], g_dx[9];   // source shift of the group's box (added to dst*sstride)
    unsigned char g_nt[9];          // taps in the group (1..3)
    unsigned char g_win[9][3];      // row offset of each tap's window inside the box
    unsigned char g_w[9][3];        // filter tap index of each tap (row block of the B tensor map)
    int a_rows;                     // rows of one source box (TWbox*TH*TN)
    int a_slot;                     // bytes reserved per source box in a stage (multiple of 1024)
    int a_sbo;                      // bytes between 8-row groups of the A operand
    int b_tiles;                    // B tiles per stage = max taps per group
    int stage_bytes, stages;
    int use_bo;                     // bring-up: set the descriptor base-offset field for unaligned window starts
    int sstride;                    // fprop with stride s: source pixel = dst pixel * s + shift
    int dscale, dpy, dpx;           // dgrad of a strided conv: this launch writes dst pixels (i*dscale+dpy, j*dscale+dpx)
    int cblocks;                    // source channels / 32
    int rows_per_tap;               // rows of the B tensor map per tap
    int mode;                       // 0 fprop, 1 dgrad
    int mtu;                        // M tiles per unit (1 or 2)
    int ring_bytes, tmem_cols, acc_stride;   // RING_BYTES / 512 / 256, or the two-CTAs-per-SM configuration (launch_tc)
    float* dst;
    const float* bias;              // fprop
    const float* mask;              // dgrad: ReLU mask source (same shape as dst) or null
    int relu, beta, round_out;
    int scatter, V, n_valid, anchor_base, A;
    // split-operand mode (ACT_S32 storage, common.cuh): source, filter, mask and destination hold bf16 (hi | lo) pairs per
    // 32-channel group; every 32-channel slab is contracted by six kind::f16 MMAs (hi*hi, lo*hi, hi*lo; K = 16 each)
    // instead of four kind::tf32 ones.  split_terms (bring-up): 1 = hi*hi only, 2 = + lo*hi, 3 = all.
    int split, split_terms;
    // fused 2x2 / stride-2 max pool (fprop, split mode, row-window tiles with TH even): the epilogue writes the POOLED
    // activation + the pool's code bytes instead of the full-resolution one (which nothing else reads)
    int pool; float* pool_dst; unsigned char* pool_code; int Hp, Wp;
    // resident-filter window mode (3x3, stride 1, split operands, one N tile, whole filter <= 144 KB: conv1_2 both ways).
    // The filter -- 9 taps x cblocks tiles of block_n x 128 B -- is loaded ONCE per persistent CTA and stays in shared memory;
    // a stage holds only ONE source box of (TW+2) x (TH+2) pixels x 32 channels that serves all nine taps (tap t reads the rows
    // starting r_win[t] pixels into it).  L2 -> SM bytes per 128 pixels: 2 x 23 KB instead of 6 x (16.6 + 24) KB.
    int resb, res_bytes;
    unsigned char r_win[9];
    // pair mode (conv_tc_kernel<.., PAIR = true>, clusters of two CTAs): ONE tcgen05.mma.cta_group::2 M256 x N per step -- each CTA
    // supplies its own M tiles (A) and HALF of the filter rows of the N tile (B), so a CTA fills mtu * 16 + N / 2 * 128 B per
    // 32-channel slab instead of mtu * 16 KB + N * 128 B and reads half of the B bytes per MMA
    int c2;
};

// epilogue of one 128-row tile: TMEM -> registers -> (per-warp shared-memory transpose) -> global.
// A thread owns one TMEM lane = one destination pixel; writing its 32 channels directly would make every store
// instruction touch 32 different cache lines in 16-byte pieces (measured: 32 sectors per request, partial-sector writes,
// the bottleneck of the short-K layers).  Each warp therefore stages its 32 pixel x 32 channel chunk in 4 KB of shared
// memory (16-byte slots XOR-swizzled by row) and writes it back with 8 lanes per pixel: every instruction moves four
// complete 128-byte rows, and the bias / ReLU / tf32 rounding (fprop) or beta / ReLU-mask (dgrad) reads are coalesced too.
template <int FMT, bool SCATTER>
__device__ __forceinline__ void epilogue_tile(const TcArgs& p, int mt, int nt, uint32_t t_row, int row, int lane, float4* stage,
                                              int chunk0, int chunk_step) {
    const int rows_valid = p.TW * p.TH * p.TN;
    const int tx = mt % p.tiles_x; const int r1 = mt / p.tiles_x;
    const int ty = r1 % p.tiles_y; const int tn = r1 / p.tiles_y;
    const int lx = row % p.TW; const int r2 = row / p.TW;
    const int ly = r2 % p.TH; const int ln = r2 / p.TH;
    const int xi = tx * p.TW + lx, yi = ty * p.TH + ly, n = tn * p.TN + ln;       // tile-grid coordinates
    const int x = xi * p.dscale + p.dpx, y = yi * p.dscale + p.dpy;              // destination pixel
    const bool ok = row < rows_valid && x < p.Wd && y < p.Hd && n < p.Bn;
    const long long pix = ok ? ((long long)n * p.Hd + y) * p.Wd + x : -1;
    for (int c0 = chunk0 * 32; c0 < p.block_n; c0 += 32 * chunk_step) {      // this warp's share of the 32-column chunks
        uint32_t r[32];
        __syncwarp();
        tc_ld32(t_row + (uint32_t)c0, r);
        tc_wait_ld();
        const int ch0 = nt * p.block_n + c0;
        if (SCATTER) {
            if (ok) {
                const int hw = p.Hd * p.Wd;
                const int pimg = y * p.Wd + x;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int ch = ch0 + j;
                    if (ch < p.n_valid) {
                        const int bt = ch / p.V, v = ch - bt * p.V;
                        float val = __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + ch) : 0.f);
                        p.dst[((long long)n * p.A + p.anchor_base + (long long)bt * hw + pimg) * p.V + v] = val;
                    }
                }
            }
            continue;
        }
        // stage: row = lane, eight 16-byte slots, slot index XOR (lane & 7)
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
            stage[lane * 8 + (j4 ^ (lane & 7))] = make_float4(__uint_as_float(r[j4 * 4]), __uint_as_float(r[j4 * 4 + 1]),
                                                              __uint_as_float(r[j4 * 4 + 2]), __uint_as_float(r[j4 * 4 + 3]));
        __syncwarp();
        if (FMT == ACT_S32 && p.pool) {
            // Fused max pool.  TW = 8 and TH even: the warp's 32 rows are 4 line slots of 8 pixels, slots (0,1) and (2,3) are
            // vertical neighbours of one image -> 2 x 4 complete 2x2 windows.  A lane owns one window and 8 channels: bias +
            // ReLU, round each cell to the stored (hi + lo) value like the stand-alone pool sees it, first maximum in row-major
            // order, code byte = winning cell | (winner > 0) << 2 (maxpool2x2_fwd_code_kernel's format).
            const int q = lane & 3, wp = lane >> 2;                  // channel octet, window 0..7
            const int r00 = (wp >> 2) * 16 + (wp & 3) * 2;            // staged row of the window's top-left pixel
            const int chs = ch0 + q * 8;
            float bv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) bv[j] = 0.f;
            if (p.bias && chs < p.cd_valid) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + chs)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias + chs + 4));
                bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
            }
            float best[8]; int arg[8];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int rr = r00 + (t >> 1) * 8 + (t & 1);
                const float4 v0 = stage[rr * 8 + ((2 * q) ^ (rr & 7))], v1 = stage[rr * 8 + ((2 * q + 1) ^ (rr & 7))];
                float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    float a = v[j] + bv[j], b = v[j + 1] + bv[j + 1];
                    if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                    uint32_t h, l;
                    split2(a, b, h, l);
                    v[j] = bf16_lo_f(h) + bf16_lo_f(l); v[j + 1] = bf16_hi_f(h) + bf16_hi_f(l);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) if (t == 0 || v[j] > best[j]) { best[j] = v[j]; arg[j] = t; }
            }
            // the window's coordinates: the same row -> pixel map as above, for the top-left staged row
            const int R = (row & ~31) + r00;
            const int wlx = R % p.TW, w2 = R / p.TW;
            const int wxi = tx * p.TW + wlx, wyi = ty * p.TH + (w2 % p.TH), wn = tn * p.TN + w2 / p.TH;
            if (R < rows_valid && wxi < p.Wd && wyi < p.Hd && wn < p.Bn && chs < p.cd_valid) {
                const long long e = (((long long)wn * p.Hp + (wyi >> 1)) * p.Wp + (wxi >> 1)) * p.Cd + chs;
                uint4 h, l;
                split2(best[0], best[1], h.x, l.x); split2(best[2], best[3], h.y, l.y); split2(best[4], best[5], h.z, l.z); split2(best[6], best[7], h.w, l.w);
                unsigned char* a = s32_addr(p.pool_dst, e);
                *reinterpret_cast<uint4*>(a) = h;
                *reinterpret_cast<uint4*>(a + 64) = l;
                uint32_t c0w = 0, c1w = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    c0w |= (uint32_t)(arg[j] | (best[j] > 0.f ? 4 : 0)) << (8 * j);
                    c1w |= (uint32_t)(arg[4 + j] | (best[4 + j] > 0.f ? 4 : 0)) << (8 * j);
                }
                *reinterpret_cast<uint2*>(p.pool_code + e) = make_uint2(c0w, c1w);
            }
            continue;
        }
        if (FMT == ACT_S32) {
            // split storage: 4 lanes per pixel, 8 channels each = one 16-byte piece of the pixel's 64 high-part bytes and
            // one of its 64 low-part bytes; an instruction moves 8 rows x 64 contiguous bytes.  Every mask / old-value load
            // of the chunk is issued as RAW bits before anything consumes them (a load -> convert chain per row would
            // expose the memory latency once per row: measured 4.6 us per chunk).
            const int q = lane & 3, sub8 = lane >> 2;
            const int chs = ch0 + q * 8;
            const bool chok8 = chs < p.cd_valid;
            long long pr8[4];
#pragma unroll
            for (int it = 0; it < 4; ++it) pr8[it] = __shfl_sync(0xffffffffu, pix, it * 8 + sub8);
            float bv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) bv[j] = 0.f;
            if (p.mode == 0 && p.bias && chok8) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + chs)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias + chs + 4));
                bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
            }
            uint4 mraw[4], ohi[4], olo[4];
            const bool want_mask = p.mode == 1 && p.mask != nullptr, want_old = p.mode == 1 && p.beta;
            if (want_mask) {
#pragma unroll
                for (int it = 0; it < 4; ++it) {      // rows that are not stored read row 0 (a valid address) instead of branching
                    const long long e = (pr8[it] >= 0 && chok8) ? pr8[it] * p.Cd + chs : 0;
                    mraw[it] = __ldg(reinterpret_cast<const uint4*>(s32_addr(p.mask, e)));
                }
            }
            if (want_old) {
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const long long e = (pr8[it] >= 0 && chok8) ? pr8[it] * p.Cd + chs : 0;
                    const unsigned char* a = s32_addr(p.dst, e);
                    ohi[it] = *reinterpret_cast<const uint4*>(a);
                    olo[it] = *reinterpret_cast<const uint4*>(a + 64);
                }
            }
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int rr = it * 8 + sub8;
                const float4 v0 = stage[rr * 8 + ((2 * q) ^ (rr & 7))], v1 = stage[rr * 8 + ((2 * q + 1) ^ (rr & 7))];
                float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                if (p.mode == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { v[j] += bv[j]; if (p.relu) v[j] = fmaxf(v[j], 0.f); }
                } else {
                    if (want_old) {
                        const uint32_t oh[4] = {ohi[it].x, ohi[it].y, ohi[it].z, ohi[it].w}, ol[4] = {olo[it].x, olo[it].y, olo[it].z, olo[it].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            v[2 * j] += bf16_lo_f(oh[j]) + bf16_lo_f(ol[j]);
                            v[2 * j + 1] += bf16_hi_f(oh[j]) + bf16_hi_f(ol[j]);
                        }
                    }
                    if (want_mask) {
                        const uint32_t mh[4] = {mraw[it].x, mraw[it].y, mraw[it].z, mraw[it].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {        // x > 0  <=>  its high part is a positive bf16 (sign clear, not zero)
                            if (!((mh[j] & 0x8000u) == 0u && (mh[j] & 0x7fffu) != 0u)) v[2 * j] = 0.f;
                            if (!((mh[j] & 0x80000000u) == 0u && (mh[j] & 0x7fff0000u) != 0u)) v[2 * j + 1] = 0.f;
                        }
                    }
                }
                if (pr8[it] < 0 || !chok8) continue;
                uint4 h, l;
                split2(v[0], v[1], h.x, l.x); split2(v[2], v[3], h.y, l.y); split2(v[4], v[5], h.z, l.z); split2(v[6], v[7], h.w, l.w);
                unsigned char* a = s32_addr(p.dst, pr8[it] * p.Cd + chs);
                *reinterpret_cast<uint4*>(a) = h;
                *reinterpret_cast<uint4*>(a + 64) = l;
            }
            continue;
        }
        const int c4 = lane & 7, sub = lane >> 3;
        const int ch = ch0 + c4 * 4;
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.mode == 0 && p.bias && ch < p.cd_valid) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + ch));
        long long prs[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) prs[it] = __shfl_sync(0xffffffffu, pix, it * 4 + sub);   // all lanes, before any divergence
        const bool chok = ch < p.cd_valid;
        if (p.mode == 0) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + sub;
                float4 v = stage[rr * 8 + (c4 ^ (rr & 7))];
                if (prs[it] < 0 || !chok) continue;
                v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
                if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                if (FMT == ACT_F32 && p.round_out) { v.x = tf32_rn(v.x); v.y = tf32_rn(v.y); v.z = tf32_rn(v.z); v.w = tf32_rn(v.w); }
                act_st4<FMT>(p.dst, prs[it] * p.Cd + ch, v);
            }
        } else {
            // dgrad: issue every mask / old-value load of the chunk first (8 independent cache lines per lane in flight),
            // then combine -- a load -> use chain per row would expose the memory latency eight times
            float4 m4[8], o4[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const bool okr = prs[it] >= 0 && chok;
                m4[it] = (p.mask && okr) ? act_ld4_sign<FMT>(p.mask, prs[it] * p.Cd + ch) : make_float4(1.f, 1.f, 1.f, 1.f);
                o4[it] = (p.beta && okr) ? act_ld4<FMT>(p.dst, prs[it] * p.Cd + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + sub;
                float4 v = stage[rr * 8 + (c4 ^ (rr & 7))];
                if (prs[it] < 0 || !chok) continue;
                v.x += o4[it].x; v.y += o4[it].y; v.z += o4[it].z; v.w += o4[it].w;
                v.x = m4[it].x > 0.f ? v.x : 0.f; v.y = m4[it].y > 0.f ? v.y : 0.f; v.z = m4[it].z > 0.f ? v.z : 0.f; v.w = m4[it].w > 0.f ? v.w : 0.f;
                if (FMT == ACT_F32 && p.round_out) { v.x = tf32_rn(v.x); v.y = tf32_rn(v.y); v.z = tf32_rn(v.z); v.w = tf32_rn(v.w); }
                act_st4<FMT>(p.dst, prs[it] * p.Cd + ch, v);
            }
        }
    }
}

// FMT: operand / storage format (ACT_F32 = tf32 MMAs, ACT_S32 = split bf16 MMAs); SCATTER: the head epilogue (fprop only)
// PAIR: the cta_group::2 variant (TcArgs::c2; must be launched as clusters of two CTAs -- a kernel that contains cta_group::2
// instructions cannot be launched without a cluster, so it is its own instantiation)
template <int FMT, bool SCATTER, bool PAIR = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_src, const __grid_constant__ CUtensorMap map_w, const TcArgs p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                 // SWIZZLE_128B needs 1024-byte alignment
    const uint32_t bars = base + (uint32_t)p.ring_bytes;
    // barrier layout: full[MAX_STAGES] empty[MAX_STAGES] tfull[2] tempty[2] then tmem pointer
    const uint32_t full0 = bars, empty0 = bars + 8 * MAX_STAGES, tfull0 = bars + 16 * MAX_STAGES, tempty0 = tfull0 + 16;
    const uint32_t tmem_slot = tempty0 + 16;
    const uint32_t bres0 = tmem_slot + 8;                          // resident-filter mode: "the filter has landed"
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
    const int STAGES = p.stages;
    const int STAGE_BYTES = p.stage_bytes;
    const uint32_t b_off = (uint32_t)(p.mtu * p.a_slot);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nepi = (int)(blockDim.x >> 5) - 2;                  // epilogue warps: 4 (two CTAs per SM) or 8
    constexpr bool c2 = PAIR;
    const uint32_t rank = c2 ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, (uint32_t)(c2 ? 2 * nepi : nepi)); }
        mbar_init(bres0, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_src) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) {
        if (c2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (c2) cluster_sync_all();                                   // the peer's barriers exist before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // a unit = mtu consecutive M tiles (the second may not exist) x one N tile
    const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
    const int m_pairs = (m_tiles + p.mtu - 1) / p.mtu;
    const int total_units = m_pairs * p.n_tiles;
    // pair mode: a pair unit = two consecutive units of the SAME N tile, one per CTA (an odd count: the peer recomputes the last one)
    const int pair_units = ((m_pairs + 1) >> 1) * p.n_tiles;
    const int u_first = c2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int u_step = c2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int u_end = c2 ? pair_units : total_units;
    const uint32_t hb_bytes = c2 ? ((uint32_t)p.block_n * 64u) : ((uint32_t)p.block_n * 128u);     // bytes of this CTA's part of a B tile
    const uint32_t a_bytes = (uint32_t)p.a_rows * 128u;
    const uint32_t b_bytes = (uint32_t)p.block_n * 128u;
    const int noff = (p.block_n + 31) & ~31;                       // TMEM column offset of the second tile's accumulator
    const int acc_stages = (p.mtu * noff <= p.acc_stride) ? 2 : 1;   // two accumulator buffers when a unit fits in one of them

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (p.resb) {
            if (elect_one_sync()) {
                const uint32_t ring = base + (uint32_t)p.res_bytes;
                mbar_expect_tx(bres0, (uint32_t)p.res_bytes);
                for (int t = 0; t < 9; ++t)
                    for (int cb = 0; cb < p.cblocks; ++cb)
                        tma_load_2d(base + (uint32_t)(t * p.cblocks + cb) * b_bytes, &map_w, bres0, cb * BLOCK_K, t * p.rows_per_tap);
                int stage = 0; uint32_t phase = 0;
                const int dy = p.g_dy[0], dx = p.g_dx[0];
                for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
                    const int tx = u % p.tiles_x; const int r1 = u / p.tiles_x;
                    const int x0 = tx * p.TW, y0 = (r1 % p.tiles_y) * p.TH, n0 = (r1 / p.tiles_y) * p.TN;
                    for (int cb = 0; cb < p.cblocks; ++cb) {
                        mbar_wait(empty0 + 8 * stage, phase ^ 1);
                        const uint32_t fb = full0 + 8 * stage;
                        mbar_expect_tx(fb, a_bytes);
                        tma_load_4d(ring + stage * STAGE_BYTES, &map_src, fb, cb * BLOCK_K, x0 + dx, y0 + dy, n0);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        } else if (elect_one_sync()) {
            int stage = 0; uint32_t phase = 0;
            for (int u = u_first; u < u_end; u += u_step) {
                int mp, nt;
                if (c2) { nt = u % p.n_tiles; mp = 2 * (u / p.n_tiles) + (int)rank; if (mp >= m_pairs) mp = m_pairs - 1; }
                else { mp = u / p.n_tiles; nt = u - mp * p.n_tiles; }
                int x0[2], y0[2], n0[2];
                // pair mode always moves mtu tiles (a missing second tile re-reads the last one; its rows are never stored)
                const bool two = p.mtu == 2 && (c2 || 2 * mp + 1 < m_tiles);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    int mt = p.mtu * mp + j; if (c2 && mt >= m_tiles) mt = m_tiles - 1;
                    const int tx = mt % p.tiles_x; const int r1 = mt / p.tiles_x;
                    x0[j] = tx * p.TW; y0[j] = (r1 % p.tiles_y) * p.TH; n0[j] = (r1 / p.tiles_y) * p.TN;
                }
                for (int gi = 0; gi < p.ngroups; ++gi) {
                    const int dy = p.g_dy[gi], dx = p.g_dx[gi], ntap = p.g_nt[gi];
                    for (int cb = 0; cb < p.cblocks; ++cb) {
                        mbar_wait(empty0 + 8 * stage, phase ^ 1);
                        const uint32_t sa = base + stage * STAGE_BYTES;
                        const uint32_t bytes = (two ? 2u : 1u) * a_bytes + (uint32_t)ntap * hb_bytes;
                        if (!c2) {
                            const uint32_t fb = full0 + 8 * stage;
                            mbar_expect_tx(fb, bytes);
                            tma_load_4d(sa, &map_src, fb, cb * BLOCK_K, x0[0] * p.sstride + dx, y0[0] * p.sstride + dy, n0[0]);
                            if (two) tma_load_4d(sa + p.a_slot, &map_src, fb, cb * BLOCK_K, x0[1] * p.sstride + dx, y0[1] * p.sstride + dy, n0[1]);
                            for (int t = 0; t < ntap; ++t)
                                tma_load_2d(sa + b_off + (uint32_t)t * b_bytes, &map_w, fb, cb * BLOCK_K, (int)p.g_w[gi][t] * p.rows_per_tap + nt * p.block_n);
                        } else {
                            // every load of the pair signals the LEADER's barrier, which expects the bytes of both CTAs
                            const uint32_t fb = mapa_u32(full0 + 8 * stage, 0);
                            if (leader) mbar_expect_tx(full0 + 8 * stage, 2u * bytes);
                            tma_load_4d_c2(sa, &map_src, fb, cb * BLOCK_K, x0[0] * p.sstride + dx, y0[0] * p.sstride + dy, n0[0]);
                            if (two) tma_load_4d_c2(sa + p.a_slot, &map_src, fb, cb * BLOCK_K, x0[1] * p.sstride + dx, y0[1] * p.sstride + dy, n0[1]);
                            for (int t = 0; t < ntap; ++t)        // this CTA's half of the filter rows of the N tile
                                tma_load_2d_c2(sa + b_off + (uint32_t)t * hb_bytes, &map_w, fb, cb * BLOCK_K,
                                               (int)p.g_w[gi][t] * p.rows_per_tap + nt * p.block_n + (int)rank * (p.block_n >> 1));
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one_sync()) {
            // instruction descriptor: fp32 accumulate, A / B format tf32 (2) or bf16 (1), both K-major, N, M
            const uint32_t fmt = FMT == ACT_S32 ? 1u : 2u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
            const uint32_t idesc2 = (idesc & ~(0x1fu << 24)) | ((uint32_t)(256 >> 4) << 24);      // the pair's M = 256
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            if (FMT == ACT_S32 && p.resb) {
                const uint32_t ring = base + (uint32_t)p.res_bytes;
                mbar_wait(bres0, 0);                                    // the whole filter has landed
                tc_fence_after();
                for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
                    mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d0 = tmem_base + (uint32_t)(acc * p.acc_stride);
                    uint32_t started = 0;
                    for (int cb = 0; cb < p.cblocks; ++cb) {
                        mbar_wait(full0 + 8 * stage, phase);
                        tc_fence_after();
                        const uint32_t sa = ring + stage * STAGE_BYTES;
                        for (int t = 0; t < 9; ++t) {
                            const uint64_t a0 = make_kmajor_desc(sa + (uint32_t)p.r_win[t] * 128u, (uint32_t)p.a_sbo, p.use_bo);
                            const uint64_t bd = make_kmajor_desc(base + (uint32_t)(t * p.cblocks + cb) * b_bytes);
#pragma unroll
                            for (int c = 0; c < 6; ++c) {
                                if (c >= 2 * p.split_terms) break;
                                const int ka = (c < 4) ? c : c - 4;                     // 0 1 | 2 3 | 0 1
                                const int kb = (c < 2) ? c : c - 2;                     // 0 1 | 0 1 | 2 3
                                tc_mma_bf16(d0, a0 + (uint64_t)(ka * 2), bd + (uint64_t)(kb * 2), idesc, started);
                                started = 1;
                            }
                        }
                        tc_commit(empty0 + 8 * stage);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(tfull0 + 8 * acc);
                    if (++acc == acc_stages) { acc = 0; acc_phase ^= 1; }
                }
            } else if (!c2 || leader)
            for (int u = u_first; u < u_end; u += u_step) {
                const int mp = u / p.n_tiles;
                const bool two = p.mtu == 2 && (c2 || 2 * mp + 1 < m_tiles);
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(acc * p.acc_stride);
                const uint32_t d1 = d0 + (uint32_t)noff;
                uint32_t started = 0;
                for (int gi = 0; gi < p.ngroups; ++gi) {
                    const int ntap = p.g_nt[gi];
                    for (int cb = 0; cb < p.cblocks; ++cb) {
                        mbar_wait(full0 + 8 * stage, phase);
                        tc_fence_after();
                        const uint32_t sa = base + stage * STAGE_BYTES;
                        for (int t = 0; t < ntap; ++t) {
                            const uint32_t woff = (uint32_t)p.g_win[gi][t] * 128u;
                            const uint64_t a0 = make_kmajor_desc(sa + woff, (uint32_t)p.a_sbo, p.use_bo);
                            const uint64_t a1 = make_kmajor_desc(sa + (uint32_t)p.a_slot + woff, (uint32_t)p.a_sbo, p.use_bo);
                            const uint64_t bd = make_kmajor_desc(sa + b_off + (uint32_t)t * hb_bytes);
                            if (FMT == ACT_S32) {
                                // a 128-byte row = 16 hi | 16 hi | 16 lo | 16 lo (bf16) of 32 channels: K offsets 0, 1 = high parts,
                                // 2, 3 = low parts (32 bytes = +2 in 16-byte units each).  x*w ~ xh*wh + xl*wh + xh*wl.
#pragma unroll
                                for (int c = 0; c < 6; ++c) {
                                    if (c >= 2 * p.split_terms) break;
                                    const int ka = (c < 2) ? c : (c < 4 ? c : c - 4);       // 0 1 | 2 3 | 0 1
                                    const int kb = (c < 2) ? c : (c < 4 ? c - 2 : c - 2);   // 0 1 | 0 1 | 2 3
                                    if (c2) {
                                        tc_mma_bf16_c2(d0, a0 + (uint64_t)(ka * 2), bd + (uint64_t)(kb * 2), idesc2, started);
                                        if (two) tc_mma_bf16_c2(d1, a1 + (uint64_t)(ka * 2), bd + (uint64_t)(kb * 2), idesc2, started);
                                    } else {
                                        tc_mma_bf16(d0, a0 + (uint64_t)(ka * 2), bd + (uint64_t)(kb * 2), idesc, started);
                                        if (two) tc_mma_bf16(d1, a1 + (uint64_t)(ka * 2), bd + (uint64_t)(kb * 2), idesc, started);
                                    }
                                    started = 1;
                                }
                            } else {
#pragma unroll
                                for (int k = 0; k < BLOCK_K / 8; ++k) {
                                    // advance 8 tf32 = 32 bytes inside the 128-byte swizzle row: +2 in 16-byte units
                                    tc_mma_tf32(d0, a0 + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, started);
                                    if (two) tc_mma_tf32(d1, a1 + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, started);
                                    started = 1;
                                }
                            }
                        }
                        if (c2) tc_commit_c2(empty0 + 8 * stage, 3);  // ... in both CTAs
                        else tc_commit(empty0 + 8 * stage);         // frees the smem slot when these MMAs retire
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
                if (c2) tc_commit_c2(tfull0 + 8 * acc, 3);
                else tc_commit(tfull0 + 8 * acc);               // accumulators complete -> epilogue
                if (++acc == acc_stages) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (4 or 8 warps: a warp may only read the TMEM lane quarter warp % 4) =====================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int ew = warp - 2;                                      // 0 .. nepi - 1
        const int chunk0 = ew >> 2, chunk_step = nepi >> 2;           // with 8 warps the two warps of a quarter alternate chunks
        float4* stage = reinterpret_cast<float4*>(smem_raw + (bars + 256 - raw)) + ew * 256;
        int acc = 0; uint32_t acc_phase = 0;
        const uint32_t tempty_l = c2 ? mapa_u32(tempty0, 0) : tempty0;       // pair mode: the leader's MMA warp waits for both CTAs
        for (int u = u_first; u < u_end; u += u_step) {
            int mp, nt;
            if (c2) { nt = u % p.n_tiles; mp = 2 * (u / p.n_tiles) + (int)rank; }
            else { mp = u / p.n_tiles; nt = u - mp * p.n_tiles; }
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t)(acc * p.acc_stride) + ((uint32_t)(quarter * 32) << 16);
            for (int j = 0; j < p.mtu; ++j)              // one copy of the epilogue code for both tiles of a unit
                if (mp < m_pairs && p.mtu * mp + j < m_tiles)
                    epilogue_tile<FMT, SCATTER>(p, p.mtu * mp + j, nt, t_row + (uint32_t)(j * noff), row, lane, stage, chunk0, chunk_step);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (c2) mbar_arrive_cluster(tempty_l + 8 * acc); else mbar_arrive(tempty0 + 8 * acc); }
            if (++acc == acc_stages) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (c2) cluster_sync_all();                                   // no commit / remote arrive is still on its way to a CTA that exits
    if (warp == 1) {
        tc_fence_after();
        if (c2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// MN-major descriptor: the matrix dimension that is NOT contracted (channels) is the contiguous one.
// For 32-bit (tf32) operands the only MN-major shared-memory layout the tensor core accepts is
// SWIZZLE_128B_BASE32B (layout type 1): 128-byte rows, 32-byte chunks XOR-swizzled with (row & 3),
// written by TMA with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  A 32-channel x P-pixel block is
// [P rows][128 B]; blocks of 32 channels are `lbo_bytes` apart (leading byte offset) and groups of
// 4 pixel rows 512 B apart (stride byte offset); one K=8 instruction spans two such groups.
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}

// ------------------------------------------------------------------ wgrad kernel
//   dW[tap][c][n] = sum over pixels of  x[pixel + shift(tap)][c] * dz[pixel][n]      db[n] = sum over pixels of dz[pixel][n]
// GEMM view: M = 128 "filter rows" = 4 slots of 32 input channels, each slot = (tap, 32-channel
// block) so that layers with Cin = 64 still fill the MMA; N = up to 256 output channels; the
// contraction runs over pixels, P (a multiple of 8, <= 64) per pipeline stage.  Both operands are
// MN-major: the same [pixels][32 ch] boxes as in fprop, in the 32-byte-atom swizzle that tf32
// MN-major operands require, with a descriptor that says "transposed".  The tensor maps are 5-D
// (c32, x, y, image, channel block) so ONE TMA instruction brings all channel blocks of an operand.
// The bias gradient rides along as one extra slot whose A block is all ones.
// The pixel range is split over CTAs; partial filters go to a workspace and are summed in a fixed
// order afterwards (deterministic).
struct WgArgs {
    int PW, PH, PN, P;              // pixel box, P = PW*PH*PN
    int ptx, pty, ptn;              // pixel tiling of the dz map
    int cblocks, taps, kdim;        // Cin/32, k*k, k
    int slots, m_tiles;             // taps*cblocks (+1 with the bias slot), ceil(slots/8): a unit owns 8 slots = two 128-row MMA tiles
    int bias_slot;                  // slot index of the all-ones block, or -1
    int load_blocks;                // channel blocks per x TMA load = min(4, cblocks)
    int n_tiles, block_n;           // Cout tiling
    int Cin, Cout;                  // Cout = stored channel stride of dz and of the HWIO filter
    int off0, offstep;              // source shift per axis for tap index t: off0 + t*offstep
    int sstride;                    // conv stride (x coordinates = pixel*sstride + shift)
    int splits, tiles_per_split;    // pixel-tile ranges
    int mtu;                        // 128-row MMA tiles per unit (1 or 2): a unit owns 4*mtu slots
    long long psize;                // floats per split in the workspace = taps*Cin*Cout + Cout
    float* partial;                 // [splits][psize]
    const float* ones;              // >= 64*32 floats of 1.0f
    int debug;                      // bring-up switches (SSDB_WG_DEBUG): 1 skip x loads, 2 skip dz loads, 4 skip MMAs
};

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dz, const WgArgs p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bars = base + RING_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * MAX_STAGES, tfull0 = bars + 16 * MAX_STAGES, tempty0 = tfull0 + 16;
    const uint32_t tmem_slot = tempty0 + 16;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int STAGES = stages_for(p.mtu);
    const int STAGE_BYTES = stage_bytes_for(p.mtu);
    const int spu = 4 * p.mtu;                                  // slots per unit

    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dz) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int units = p.m_tiles * p.n_tiles * p.splits;
    const int pix_tiles = p.ptx * p.pty * p.ptn;
    const uint32_t blk_bytes = (uint32_t)p.P * 128u;          // one 32-channel column block
    const int nblk_b = p.block_n / 32;
    const uint32_t b_off = (uint32_t)spu * blk_bytes;          // B blocks follow the mtu x 4 A blocks

    if (warp == 0) {
        if (elect_one_sync()) {
            int stage = 0; uint32_t phase = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                // units that share a pixel range are neighbours (m tile fastest): the CTAs running at the same time stream
                // the same x / dz region, which is then read from HBM once and served from L2 to the others
                const int mt = u % p.m_tiles; const int r1 = u / p.m_tiles;
                const int nt = r1 % p.n_tiles; const int sp = r1 / p.n_tiles;
                const int q0 = sp * p.tiles_per_split;
                const int q1 = min(pix_tiles, q0 + p.tiles_per_split);
                int na = p.slots - mt * spu; na = na > spu ? spu : na;      // valid A blocks of this unit
                for (int q = q0; q < q1; ++q) {
                    const int qx = q % p.ptx; const int r2 = q / p.ptx;
                    const int qy = r2 % p.pty; const int qn = r2 / p.pty;
                    const int x0 = qx * p.PW, y0 = qy * p.PH, n0 = qn * p.PN;
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t fb = full0 + 8 * stage;
                    const uint32_t ea = (p.debug & 1) ? 0u : (uint32_t)na * blk_bytes, eb = (p.debug & 2) ? 0u : (uint32_t)nblk_b * blk_bytes;
                    if (ea + eb) mbar_expect_tx(fb, ea + eb); else mbar_arrive(fb);
                    const uint32_t sa = base + stage * STAGE_BYTES;
                    for (int j = 0; j < na && !(p.debug & 1); j += p.load_blocks) {
                        const int slot = mt * spu + j;
                        if (slot == p.bias_slot) { bulk_load_1d(sa + (uint32_t)j * blk_bytes, p.ones, blk_bytes, fb); break; }
                        const int tap = slot / p.cblocks, cb = slot - tap * p.cblocks;
                        const int kh = tap / p.kdim, kw = tap - kh * p.kdim;
                        tma_load_5d(sa + (uint32_t)j * blk_bytes, &map_x, fb, 0, x0 * p.sstride + p.off0 + kw * p.offstep,
                                    y0 * p.sstride + p.off0 + kh * p.offstep, n0, cb);
                    }
                    if (!(p.debug & 2)) tma_load_5d(sa + b_off, &map_dz, fb, 0, x0, y0, n0, nt * nblk_b);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int mt = u % p.m_tiles; const int sp = (u / p.m_tiles) / p.n_tiles;
                const int q0 = sp * p.tiles_per_split;
                const int q1 = min(pix_tiles, q0 + p.tiles_per_split);
                const bool two = p.mtu == 2 && p.slots - mt * 8 > 4;     // the second 128-row tile has at least one valid slot
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d0 = tmem_base, d1 = tmem_base + (uint32_t)MAX_N;      // both accumulators live at once: one TMEM stage
                for (int q = q0; q < q1; ++q) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = base + stage * STAGE_BYTES;
                    const uint64_t a0 = make_mnmajor_desc(sa, blk_bytes);
                    const uint64_t a1 = make_mnmajor_desc(sa + 4u * blk_bytes, blk_bytes);
                    const uint64_t bd = make_mnmajor_desc(sa + b_off, blk_bytes);
                    const int ksteps = p.P / 8;
                    for (int k = 0; k < ksteps && !(p.debug & 4); ++k) {     // 8 pixel rows = 1024 bytes = +64 in 16-byte units
                        tc_mma_tf32(d0, a0 + (uint64_t)(k * 64), bd + (uint64_t)(k * 64), idesc, (q > q0 || k > 0) ? 1u : 0u);
                        if (two) tc_mma_tf32(d1, a1 + (uint64_t)(k * 64), bd + (uint64_t)(k * 64), idesc, (q > q0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(empty0 + 8 * stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull0 + 8 * acc);
                acc_phase ^= 1;
            }
        }
    } else {
        const int quarter = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        const long long wsize = (long long)p.taps * p.Cin * p.Cout;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int mt = u % p.m_tiles; const int r1 = u / p.m_tiles;
            const int nt = r1 % p.n_tiles; const int sp = r1 / p.n_tiles;
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tc_fence_after();
            for (int half = 0; half < 2; ++half) {
                const int slot = mt * spu + half * 4 + quarter;
                if (half >= p.mtu || mt * spu + half * 4 >= p.slots) break;   // warp-uniform: no (valid) second tile
                const bool is_bias = slot == p.bias_slot;
                const bool ok = slot < p.slots && (!is_bias || lane == 0);
                const int tap = (ok && !is_bias) ? slot / p.cblocks : 0, cb = (ok && !is_bias) ? slot - tap * p.cblocks : 0;
                float* drow = p.partial + (long long)sp * p.psize +
                              (is_bias ? wsize : ((long long)tap * p.Cin + cb * 32 + lane) * p.Cout);
                const uint32_t t_row = tmem_base + (uint32_t)(half * MAX_N) + ((uint32_t)(quarter * 32) << 16);
                for (int c0 = 0; c0 < p.block_n; c0 += 32) {
                    uint32_t r[32];
                    __syncwarp();
                    tc_ld32(t_row + (uint32_t)c0, r);
                    tc_wait_ld();
                    const int ch0 = nt * p.block_n + c0;
                    if (ok) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (ch0 + j < p.Cout)
                                *reinterpret_cast<float4*>(drow + ch0 + j) =
                                    make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
            acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ window wgrad kernel (3x3, dilation 1, stride 1, SAME, N <= 128)
// The layers with few output channels make the plain wgrad kernel L2-bound: every (tap, channel block) slot reloads a
// shifted copy of the same x pixels.  Here ONE x box per channel block -- 11 pixels wide and PH + 2 lines tall for an
// 8 x PH block of output pixels -- serves all nine taps: the MMA reads tap (kh, kw) as the row window that starts kw
// pixels into line j + kh of the box.  The A descriptor of a 128-row M tile = 4 windows x 32 channels with a
// leading-dimension offset of ONE ROW (128 bytes): overlapping windows, no extra traffic (the 4th window is a dummy whose
// rows are dropped).  One K = 8 MMA consumes one 8-pixel line; lines are 11 rows apart in the x box and 8 rows apart in
// the dz box.  All 3 x cpu accumulators of a unit (+1 for the bias gradient: A = a resident block of ones with leading
// offset 0) live in TMEM at once; dz is read once per channel group, x once per pixel.
struct WgRwArgs {
    int PH;                         // output lines per stage (8 pixels each); the x box has PH + 2 lines
    int ptx, pty, ptn;              // pixel tiling of the dz map (x in steps of 8, one image per box)
    int cblocks, cpu, cgroups;      // Cin/32, channel blocks per unit, cblocks/cpu
    int block_n, n_tiles;           // output channels per unit (<= 128) and Cout / block_n
    int Cin, Cout, pad;
    int splits, tiles_per_split;
    int want_bias;
    int stage_bytes, stages;
    int terms;                      // split kernel bring-up: 1 = hi*hi only, 2 = + lo*hi, 3 = all products
    long long psize;
    float* partial;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_wgrad_rw_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dz, const WgRwArgs p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bars = base + RING_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * MAX_STAGES, tfull0 = bars + 16 * MAX_STAGES, tempty0 = tfull0 + 16;
    const uint32_t tmem_slot = tempty0 + 16;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
    const uint32_t ones_addr = bars + 1024u;                                  // 8 KB of 1.0f inside the epilogue staging area
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull0, 1); mbar_init(tempty0, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dz) : "memory");
    }
    {   // resident block of ones (the bias-gradient A operand); made visible to the tensor core's async proxy
        float* ones = reinterpret_cast<float*>(smem_raw + (ones_addr - raw));
        for (int i = threadIdx.x; i < 2048; i += NUM_THREADS) ones[i] = 1.0f;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int utypes = p.cgroups * p.n_tiles;                     // (channel group, N tile) pairs of one pixel range run side by side
    const int units = utypes * p.splits;
    const int pix_tiles = p.ptx * p.pty * p.ptn;
    const int nblk_b = p.block_n / 32;
    const uint32_t xbox_bytes = (uint32_t)(11 * (p.PH + 2)) * 128u;   // one channel block of x: PH + 2 lines of 11 pixels
    const uint32_t zblk_bytes = (uint32_t)(8 * p.PH) * 128u;          // one channel block of dz: PH lines of 8 pixels
    const uint32_t z_off = ((uint32_t)p.cpu * xbox_bytes + 1023u) & ~1023u;   // dz blocks follow the x boxes (1 KB aligned)
    const int STAGES = p.stages;
    const int noff = (p.block_n + 31) & ~31;

    if (warp == 0) {
        if (elect_one_sync()) {
            int stage = 0; uint32_t phase = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                // unit type fastest: the units of ONE pixel range run side by side and share x / dz through L2
                const int ut = u % utypes, sp = u / utypes;
                const int cg = ut % p.cgroups, nt = ut / p.cgroups;
                const int q0 = sp * p.tiles_per_split;
                const int q1 = min(pix_tiles, q0 + p.tiles_per_split);
                for (int q = q0; q < q1; ++q) {
                    const int qx = q % p.ptx; const int r2 = q / p.ptx;
                    const int qy = r2 % p.pty; const int n0 = r2 / p.pty;
                    const int x0 = qx * 8, y0 = qy * p.PH;
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t fb = full0 + 8 * stage;
                    const uint32_t sa = base + stage * p.stage_bytes;
                    mbar_expect_tx(fb, (uint32_t)p.cpu * xbox_bytes + (uint32_t)nblk_b * zblk_bytes);
                    tma_load_5d(sa, &map_x, fb, 0, x0 - p.pad, y0 - p.pad, n0, cg * p.cpu);
                    tma_load_5d(sa + z_off, &map_dz, fb, 0, x0, y0, n0, nt * nblk_b);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
            int stage = 0; uint32_t phase = 0; uint32_t acc_phase = 0;
            const uint64_t ones_desc = make_mnmajor_desc(ones_addr, 0u);
            const uint32_t a_blk = xbox_bytes >> 4;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int ut = u % utypes, sp = u / utypes;
                const int cg = ut % p.cgroups;
                const bool do_bias = p.want_bias && cg == 0;
                const int q0 = sp * p.tiles_per_split;
                const int q1 = min(pix_tiles, q0 + p.tiles_per_split);
                mbar_wait(tempty0, acc_phase ^ 1);
                tc_fence_after();
                for (int q = q0; q < q1; ++q) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = base + stage * p.stage_bytes;
                    // descriptors differ only in the start-address field (16-byte units)
                    const uint64_t b0 = make_mnmajor_desc(sa + z_off, zblk_bytes);
                    const uint64_t a0 = make_mnmajor_desc(sa, 128u);             // windows kw = 0..3 are 128 bytes apart
                    for (int j = 0; j < p.PH; ++j) {                             // one 8-pixel line = one K = 8 MMA per accumulator
                        const uint64_t bd = b0 + (uint64_t)(j * 64);
                        const uint32_t accum = (q > q0 || j > 0) ? 1u : 0u;
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh)
                            for (int t = 0; t < p.cpu; ++t)                      // x line j + kh, 11 rows (88 x 16 B) per line
                                tc_mma_tf32(tmem_base + (uint32_t)((kh * p.cpu + t) * noff), a0 + (uint64_t)(t * a_blk + (j + kh) * 88), bd, idesc, accum);
                        if (do_bias) tc_mma_tf32(tmem_base + (uint32_t)(3 * p.cpu * noff), ones_desc + (uint64_t)(j * 64), bd, idesc, accum);
                    }
                    tc_commit(empty0 + 8 * stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull0);
                acc_phase ^= 1;
            }
        }
    } else {
        const int quarter = warp & 3;
        uint32_t acc_phase = 0;
        const long long wsize = 9LL * p.Cin * p.Cout;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int ut = u % utypes, sp = u / utypes;
            const int cg = ut % p.cgroups, nt = ut / p.cgroups;
            const bool do_bias = p.want_bias && cg == 0;
            mbar_wait(tfull0, acc_phase);
            tc_fence_after();
            const int nacc = 3 * p.cpu + (do_bias ? 1 : 0);
            for (int ai = 0; ai < nacc; ++ai) {
                const bool is_bias = ai == 3 * p.cpu;
                const int kh = ai / p.cpu, t = ai - kh * p.cpu;
                // lane quarter = window kw (3 = dummy); bias accumulator: all rows equal, lane 0 of quarter 0 writes
                const bool ok = is_bias ? (quarter == 0 && lane == 0) : (quarter < 3);
                const int c = (cg * p.cpu + t) * 32 + lane;
                float* drow = p.partial + (long long)sp * p.psize +
                              (is_bias ? wsize : ((long long)(kh * 3 + quarter) * p.Cin + c) * p.Cout);
                const uint32_t t_row = tmem_base + (uint32_t)(ai * noff) + ((uint32_t)(quarter * 32) << 16);
                for (int c0 = 0; c0 < p.block_n; c0 += 32) {
                    uint32_t r[32];
                    __syncwarp();
                    tc_ld32(t_row + (uint32_t)c0, r);
                    tc_wait_ld();
                    const int ch0 = nt * p.block_n + c0;
                    if (ok) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (ch0 + j < p.Cout)
                                *reinterpret_cast<float4*>(drow + ch0 + j) =
                                    make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0);
            acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ split-operand (ACT_S32) wgrad kernels
// Same decomposition as the two kernels above, but x and dz are ACT_S32 tensors (bf16 hi | lo pairs per 32-channel group)
// and the MMAs are kind::f16.  A TMA map with a 64-byte inner box (32 bf16) and a 64-byte-strided outer dimension
// q = 2 * channel block + part brings every 32-channel block as TWO [pixels][64 B] sub-blocks (high parts, then low
// parts) in the SWIZZLE_64B layout -- the MN-major canonical layout for 16-bit operands: 32 channels contiguous, pixel rows
// 64 bytes apart, groups of 8 pixel rows SBO apart, 32-channel blocks LBO apart.  The high (or low) parts of consecutive
// channel blocks are every second sub-block, so "4 slots x 32 channels" is one descriptor with LBO = 2 sub-blocks, started
// at the first high or the first low sub-block.  dW ~ xh'zh + xl'zh + xh'zl: three K = 16 (pixel) MMAs into the SAME
// accumulator, which therefore has the layout of the tf32 kernels and the same epilogue.  16-bit MN-major operands run at
// the full MMA rate (the 32-bit ones above at half), so three bf16 MMAs per 16 pixels replace two half-rate tf32 ones.
__device__ __forceinline__ uint64_t make_mn64_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;                      // SWIZZLE_64B
    return d;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_wgrad_s_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dz, const WgArgs p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bars = base + RING_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * MAX_STAGES, tfull0 = bars + 16 * MAX_STAGES, tempty0 = tfull0 + 16;
    const uint32_t tmem_slot = tempty0 + 16;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int STAGES = stages_for(p.mtu);
    const int STAGE_BYTES = stage_bytes_for(p.mtu);
    const int spu = 4 * p.mtu;                                  // slots per unit

    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dz) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int units = p.m_tiles * p.n_tiles * p.splits;
    const int pix_tiles = p.ptx * p.pty * p.ptn;
    const uint32_t sub_bytes = (uint32_t)p.P * 64u;           // one part (hi or lo) of a 32-channel block
    const uint32_t blk_bytes = 2u * sub_bytes;                // one 32-channel block: hi sub-block, lo sub-block
    const int nblk_b = p.block_n / 32;
    const uint32_t b_off = (uint32_t)spu * blk_bytes;          // B blocks follow the mtu x 4 A blocks

    if (warp == 0) {
        if (elect_one_sync()) {
            int stage = 0; uint32_t phase = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int mt = u % p.m_tiles; const int r1 = u / p.m_tiles;
                const int nt = r1 % p.n_tiles; const int sp = r1 / p.n_tiles;
                const int q0 = sp * p.tiles_per_split;
                const int q1 = min(pix_tiles, q0 + p.tiles_per_split);
                int na = p.slots - mt * spu; na = na > spu ? spu : na;      // valid A blocks of this unit
                for (int q = q0; q < q1; ++q) {
                    const int qx = q % p.ptx; const int r2 = q / p.ptx;
                    const int qy = r2 % p.pty; const int qn = r2 / p.pty;
                    const int x0 = qx * p.PW, y0 = qy * p.PH, n0 = qn * p.PN;
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t fb = full0 + 8 * stage;
                    mbar_expect_tx(fb, (uint32_t)(na + nblk_b) * blk_bytes);
                    const uint32_t sa = base + stage * STAGE_BYTES;
                    for (int j = 0; j < na; j += p.load_blocks) {
                        const int slot = mt * spu + j;
                        if (slot == p.bias_slot) {           // high parts = 1.0, low parts = 0
                            bulk_load_1d(sa + (uint32_t)j * blk_bytes, p.ones, sub_bytes, fb);
                            bulk_load_1d(sa + (uint32_t)j * blk_bytes + sub_bytes, p.ones + 2048, sub_bytes, fb);
                            break;
                        }
                        const int tap = slot / p.cblocks, cb = slot - tap * p.cblocks;
                        const int kh = tap / p.kdim, kw = tap - kh * p.kdim;
                        tma_load_5d(sa + (uint32_t)j * blk_bytes, &map_x, fb, 0, x0 * p.sstride + p.off0 + kw * p.offstep,
                                    y0 * p.sstride + p.off0 + kh * p.offstep, n0, 2 * cb);
                    }
                    tma_load_5d(sa + b_off, &map_dz, fb, 0, x0, y0, n0, 2 * nt * nblk_b);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            // bf16 A / B, fp32 accumulate, both operands MN-major ("transposed")
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            const int terms = p.debug > 0 && p.debug <= 3 ? p.debug : 3;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int mt = u % p.m_tiles; const int sp = (u / p.m_tiles) / p.n_tiles;
                const int q0 = sp * p.tiles_per_split;
                const int q1 = min(pix_tiles, q0 + p.tiles_per_split);
                const bool two = p.mtu == 2 && p.slots - mt * 8 > 4;
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d0 = tmem_base, d1 = tmem_base + (uint32_t)MAX_N;
                for (int q = q0; q < q1; ++q) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = base + stage * STAGE_BYTES;
                    // [0] = high parts, [1] = low parts: the same blocks, one sub-block further
                    uint64_t a0[2], a1[2], bd[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        a0[h] = make_mn64_desc(sa + (uint32_t)h * sub_bytes, blk_bytes, 512u);
                        a1[h] = make_mn64_desc(sa + 4u * blk_bytes + (uint32_t)h * sub_bytes, blk_bytes, 512u);
                        bd[h] = make_mn64_desc(sa + b_off + (uint32_t)h * sub_bytes, blk_bytes, 512u);
                    }
                    const int ksteps = p.P / 16;
                    for (int k = 0; k < ksteps; ++k) {     // 16 pixel rows = 1024 bytes = +64 in 16-byte units
                        const uint64_t ko = (uint64_t)(k * 64);
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            if (c >= terms) break;
                            const int ha = c == 1 ? 1 : 0, hb = c == 2 ? 1 : 0;      // hi*hi, lo*hi, hi*lo
                            const uint32_t accum = (q > q0 || k > 0 || c > 0) ? 1u : 0u;
                            tc_mma_bf16(d0, a0[ha] + ko, bd[hb] + ko, idesc, accum);
                            if (two) tc_mma_bf16(d1, a1[ha] + ko, bd[hb] + ko, idesc, accum);
                        }
                    }
                    tc_commit(empty0 + 8 * stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull0 + 8 * acc);
                acc_phase ^= 1;
            }
        }
    } else {
        const int quarter = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        const long long wsize = (long long)p.taps * p.Cin * p.Cout;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int mt = u % p.m_tiles; const int r1 = u / p.m_tiles;
            const int nt = r1 % p.n_tiles; const int sp = r1 / p.n_tiles;
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tc_fence_after();
            for (int half = 0; half < 2; ++half) {
                const int slot = mt * spu + half * 4 + quarter;
                if (half >= p.mtu || mt * spu + half * 4 >= p.slots) break;   // warp-uniform: no (valid) second tile
                const bool is_bias = slot == p.bias_slot;
                const bool ok = slot < p.slots && (!is_bias || lane == 0);
                const int tap = (ok && !is_bias) ? slot / p.cblocks : 0, cb = (ok && !is_bias) ? slot - tap * p.cblocks : 0;
                float* drow = p.partial + (long long)sp * p.psize +
                              (is_bias ? wsize : ((long long)tap * p.Cin + cb * 32 + lane) * p.Cout);
                const uint32_t t_row = tmem_base + (uint32_t)(half * MAX_N) + ((uint32_t)(quarter * 32) << 16);
                for (int c0 = 0; c0 < p.block_n; c0 += 32) {
                    uint32_t r[32];
                    __syncwarp();
                    tc_ld32(t_row + (uint32_t)c0, r);
                    tc_wait_ld();
                    const int ch0 = nt * p.block_n + c0;
                    if (ok) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (ch0 + j < p.Cout)
                                *reinterpret_cast<float4*>(drow + ch0 + j) =
                                    make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
            acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// window variant (see conv_tc_wgrad_rw_kernel): the x box of a channel block arrives as a high and a low sub-box of
// [PH + 2 lines][11 pixels][64 B]; one K = 16 MMA consumes TWO 8-pixel lines (the second 8-row group is one line further:
// SBO = 11 rows in the x box, 8 rows in the dz box), windows kw = 0..3 are one row (64 bytes) apart.
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_wgrad_rw_s_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dz, const WgRwArgs p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bars = base + RING_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * MAX_STAGES, tfull0 = bars + 16 * MAX_STAGES, tempty0 = tfull0 + 16;
    const uint32_t tmem_slot = tempty0 + 16;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
    const uint32_t ones_addr = bars + 1024u;                                  // 2 KB of bf16 1.0 inside the epilogue staging area
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull0, 1); mbar_init(tempty0, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dz) : "memory");
    }
    {   // resident block of bf16 ones (16 pixel rows x 64 bytes are read, 2 KB are filled)
        uint32_t* ones = reinterpret_cast<uint32_t*>(smem_raw + (ones_addr - raw));
        for (int i = threadIdx.x; i < 512; i += NUM_THREADS) ones[i] = 0x3f803f80u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int utypes = p.cgroups * p.n_tiles;
    const int units = utypes * p.splits;
    const int pix_tiles = p.ptx * p.pty * p.ptn;
    const int nblk_b = p.block_n / 32;
    const uint32_t xsub = (uint32_t)(11 * (p.PH + 2)) * 64u;          // one part of one channel block of x
    const uint32_t zsub = (uint32_t)(8 * p.PH) * 64u;                 // one part of one channel block of dz
    const uint32_t z_off = ((uint32_t)(2 * p.cpu) * xsub + 1023u) & ~1023u;
    const int STAGES = p.stages;
    const int noff = (p.block_n + 31) & ~31;

    if (warp == 0) {
        if (elect_one_sync()) {
            int stage = 0; uint32_t phase = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int ut = u % utypes, sp = u / utypes;
                const int cg = ut % p.cgroups, nt = ut / p.cgroups;
                const int q0 = sp * p.tiles_per_split;
                const int q1 = min(pix_tiles, q0 + p.tiles_per_split);
                for (int q = q0; q < q1; ++q) {
                    const int qx = q % p.ptx; const int r2 = q / p.ptx;
                    const int qy = r2 % p.pty; const int n0 = r2 / p.pty;
                    const int x0 = qx * 8, y0 = qy * p.PH;
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t fb = full0 + 8 * stage;
                    const uint32_t sa = base + stage * p.stage_bytes;
                    mbar_expect_tx(fb, (uint32_t)(2 * p.cpu) * xsub + (uint32_t)(2 * nblk_b) * zsub);
                    tma_load_5d(sa, &map_x, fb, 0, x0 - p.pad, y0 - p.pad, n0, 2 * cg * p.cpu);
                    tma_load_5d(sa + z_off, &map_dz, fb, 0, x0, y0, n0, 2 * nt * nblk_b);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
            int stage = 0; uint32_t phase = 0; uint32_t acc_phase = 0;
            const uint64_t ones_desc = make_mn64_desc(ones_addr, 0u, 512u);
            const int terms = p.terms;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int ut = u % utypes, sp = u / utypes;
                const int cg = ut % p.cgroups;
                const bool do_bias = p.want_bias && cg == 0;
                const int q0 = sp * p.tiles_per_split;
                const int q1 = min(pix_tiles, q0 + p.tiles_per_split);
                mbar_wait(tempty0, acc_phase ^ 1);
                tc_fence_after();
                for (int q = q0; q < q1; ++q) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = base + stage * p.stage_bytes;
                    // dz: 32-channel blocks 2 sub-blocks apart, line pairs contiguous (8 rows per line: SBO = 512)
                    const uint64_t bz[2] = {make_mn64_desc(sa + z_off, 2u * zsub, 512u), make_mn64_desc(sa + z_off + zsub, 2u * zsub, 512u)};
                    for (int jp = 0; jp < p.PH / 2; ++jp) {                      // two 8-pixel lines = one K = 16 MMA per accumulator and term
                        const uint32_t first = (q > q0 || jp > 0) ? 1u : 0u;
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh)
                            for (int t = 0; t < p.cpu; ++t) {
                                // x: windows one row (64 B) apart, lines 11 rows apart
                                const uint32_t xa = sa + (uint32_t)(2 * t) * xsub + (uint32_t)((2 * jp + kh) * 11) * 64u;
                                const uint64_t ax[2] = {make_mn64_desc(xa, 64u, 704u), make_mn64_desc(xa + xsub, 64u, 704u)};
                                const uint32_t d = tmem_base + (uint32_t)((kh * p.cpu + t) * noff);
                                const uint64_t zo = (uint64_t)(jp * 64);         // 16 rows x 64 B = 1024 B
#pragma unroll
                                for (int c = 0; c < 3; ++c) {
                                    if (c >= terms) break;
                                    const int ha = c == 1 ? 1 : 0, hb = c == 2 ? 1 : 0;
                                    tc_mma_bf16(d, ax[ha], bz[hb] + zo, idesc, (first || c > 0) ? 1u : 0u);
                                }
                            }
                        if (do_bias) {
                            const uint32_t d = tmem_base + (uint32_t)(3 * p.cpu * noff);
                            tc_mma_bf16(d, ones_desc, bz[0] + (uint64_t)(jp * 64), idesc, first);
                            tc_mma_bf16(d, ones_desc, bz[1] + (uint64_t)(jp * 64), idesc, 1u);
                        }
                    }
                    tc_commit(empty0 + 8 * stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull0);
                acc_phase ^= 1;
            }
        }
    } else {
        const int quarter = warp & 3;
        uint32_t acc_phase = 0;
        const long long wsize = 9LL * p.Cin * p.Cout;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int ut = u % utypes, sp = u / utypes;
            const int cg = ut % p.cgroups, nt = ut / p.cgroups;
            const bool do_bias = p.want_bias && cg == 0;
            mbar_wait(tfull0, acc_phase);
            tc_fence_after();
            const int nacc = 3 * p.cpu + (do_bias ? 1 : 0);
            for (int ai = 0; ai < nacc; ++ai) {
                const bool is_bias = ai == 3 * p.cpu;
                const int kh = ai / p.cpu, t = ai - kh * p.cpu;
                const bool ok = is_bias ? (quarter == 0 && lane == 0) : (quarter < 3);
                const int c = (cg * p.cpu + t) * 32 + lane;
                float* drow = p.partial + (long long)sp * p.psize +
                              (is_bias ? wsize : ((long long)(kh * 3 + quarter) * p.Cin + c) * p.Cout);
                const uint32_t t_row = tmem_base + (uint32_t)(ai * noff) + ((uint32_t)(quarter * 32) << 16);
                for (int c0 = 0; c0 < p.block_n; c0 += 32) {
                    uint32_t r[32];
                    __syncwarp();
                    tc_ld32(t_row + (uint32_t)c0, r);
                    tc_wait_ld();
                    const int ch0 = nt * p.block_n + c0;
                    if (ok) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (ch0 + j < p.Cout)
                                *reinterpret_cast<float4*>(drow + ch0 + j) =
                                    make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0);
            acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ window wgrad, roles swapped ("rw2"): 3x3, stride 1, SAME, Cout % 128 == 0
// The plain split kernel moves 16 operand blocks per 16 pixels for two 128 x 256 accumulators and is bound by the L2 -> SM
// rate (measured: tensor pipe 62 %); the window kernel above pays 25 % dummy MMA rows and stops at N = 128.  Here dz is the
// M operand (128 output channels = 4 blocks) and x the N operand: the three kw taps of a filter row are three 32-channel
// blocks ONE PIXEL apart in the same x box (LBO = dil * 64 bytes), so N = 96 with no dummy window, and the three kh taps are
// three accumulators fed from lines j + kh * dil of that box.  Per 64 output pixels a unit loads one (8 + 2 dil) x (PH + 2 dil)
// x box of ONE channel block (hi + lo) and the dz blocks of its 128 output channels: 45 KB per 1728 MMA cycles = 26 B / cycle,
// well under the L2 rate, with every MMA row and column useful.  The bias gradient is one more N = 16 accumulator whose B
// operand is a resident block of ones.  Accumulator lanes are OUTPUT channels, so the epilogue stores are coalesced along Cout.
struct WgR2Args {
    int PH, dil;                    // output lines per stage (even); dilation = padding
    int ptx, pty, ptn;              // pixel tiling of the dz map (x in steps of 8, one image per box)
    int cblocks, mblocks;           // Cin / 32, Cout / 128
    int Cin, Cout;
    int splits, tiles_per_split;
    int want_bias, terms;
    int stage_bytes, stages;
    int c2;                         // run on SM pairs (conv_tc_wgrad_r2c2_kernel): Cout % 256 == 0
    long long psize;
    float* partial;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_wgrad_r2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dz, const WgR2Args p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bars = base + RING_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * MAX_STAGES, tfull0 = bars + 16 * MAX_STAGES, tempty0 = tfull0 + 16;
    const uint32_t tmem_slot = tempty0 + 16;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
    const uint32_t ones_addr = bars + 1024u;                                  // 2 KB of bf16 1.0 inside the epilogue staging area
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull0, 1); mbar_init(tempty0, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dz) : "memory");
    }
    {
        uint32_t* ones = reinterpret_cast<uint32_t*>(smem_raw + (ones_addr - raw));
        for (int i = threadIdx.x; i < 512; i += NUM_THREADS) ones[i] = 0x3f803f80u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int utypes = p.cblocks * p.mblocks;                     // (channel block of x, 128 output channels) pairs of one pixel range run side by side
    const int units = utypes * p.splits;
    const int pix_tiles = p.ptx * p.pty * p.ptn;
    const int bw = 8 + 2 * p.dil, bh = p.PH + 2 * p.dil;          // x box in pixels
    const uint32_t xsub = (uint32_t)(bw * bh) * 64u;              // one part (hi or lo) of the x box
    const uint32_t zsub = (uint32_t)(8 * p.PH) * 64u;             // one part of one 32-channel block of dz
    const uint32_t z_off = (2u * xsub + 1023u) & ~1023u;
    const int STAGES = p.stages;
    constexpr uint32_t ACC_N = 96;                                // 3 kw windows x 32 input channels

    if (warp == 0) {
        if (elect_one_sync()) {
            int stage = 0; uint32_t phase = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int ut = u % utypes, sp = u / utypes;
                const int cb = ut % p.cblocks, mt = ut / p.cblocks;
                const int q0 = sp * p.tiles_per_split;
                const int q1 = min(pix_tiles, q0 + p.tiles_per_split);
                for (int q = q0; q < q1; ++q) {
                    const int qx = q % p.ptx; const int r2 = q / p.ptx;
                    const int qy = r2 % p.pty; const int n0 = r2 / p.pty;
                    const int x0 = qx * 8, y0 = qy * p.PH;
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t fb = full0 + 8 * stage;
                    const uint32_t sa = base + stage * p.stage_bytes;
                    mbar_expect_tx(fb, 2u * xsub + 8u * zsub);
                    tma_load_5d(sa, &map_x, fb, 0, x0 - p.dil, y0 - p.dil, n0, 2 * cb);
                    tma_load_5d(sa + z_off, &map_dz, fb, 0, x0, y0, n0, 8 * mt);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            // bf16 A / B, fp32 accumulate, both MN-major; M = 128 output channels, N = 96 (taps) or 16 (bias)
            const uint32_t idesc_common = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BLOCK_M >> 4) << 24);
            const uint32_t idesc = idesc_common | ((ACC_N >> 3) << 17);
            const uint32_t idesc_bias = idesc_common | ((16u >> 3) << 17);
            int stage = 0; uint32_t phase = 0; uint32_t acc_phase = 0;
            const uint64_t ones_desc = make_mn64_desc(ones_addr, 0u, 512u);
            const int terms = p.terms;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int ut = u % utypes, sp = u / utypes;
                const int cb = ut % p.cblocks;
                const bool do_bias = p.want_bias && cb == 0;
                const int q0 = sp * p.tiles_per_split;
                const int q1 = min(pix_tiles, q0 + p.tiles_per_split);
                mbar_wait(tempty0, acc_phase ^ 1);
                tc_fence_after();
                for (int q = q0; q < q1; ++q) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = base + stage * p.stage_bytes;
                    // dz (M operand): 32-channel blocks two sub-blocks apart, a line pair = 16 contiguous rows (SBO = 512)
                    const uint64_t az[2] = {make_mn64_desc(sa + z_off, 2u * zsub, 512u), make_mn64_desc(sa + z_off + zsub, 2u * zsub, 512u)};
                    for (int jp = 0; jp < p.PH / 2; ++jp) {
                        const uint32_t first = (q > q0 || jp > 0) ? 1u : 0u;
                        const uint64_t zo = (uint64_t)(jp * 64);             // 16 rows x 64 B = 1024 B
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
                            // x (N operand): windows kw = 0, 1, 2 are dil pixels (dil * 64 B) apart, the second 8-row group is the next line
                            const uint32_t xa = sa + (uint32_t)((2 * jp + kh * p.dil) * bw) * 64u;
                            const uint64_t bx[2] = {make_mn64_desc(xa, (uint32_t)p.dil * 64u, (uint32_t)bw * 64u),
                                                    make_mn64_desc(xa + xsub, (uint32_t)p.dil * 64u, (uint32_t)bw * 64u)};
                            const uint32_t d = tmem_base + (uint32_t)kh * ACC_N;
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                if (c >= terms) break;
                                const int ha = c == 1 ? 1 : 0, hb = c == 2 ? 1 : 0;      // (dz, x): hi*hi, lo*hi, hi*lo
                                tc_mma_bf16(d, az[ha] + zo, bx[hb], idesc, (first || c > 0) ? 1u : 0u);
                            }
                        }
                        if (do_bias) {
                            const uint32_t d = tmem_base + 3u * ACC_N;
                            tc_mma_bf16(d, az[0] + zo, ones_desc, idesc_bias, first);
                            tc_mma_bf16(d, az[1] + zo, ones_desc, idesc_bias, 1u);
                        }
                    }
                    tc_commit(empty0 + 8 * stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull0);
                acc_phase ^= 1;
            }
        }
    } else {
        const int quarter = warp & 3;
        uint32_t acc_phase = 0;
        const long long wsize = 9LL * p.Cin * p.Cout;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int ut = u % utypes, sp = u / utypes;
            const int cb = ut % p.cblocks, mt = ut / p.cblocks;
            const bool do_bias = p.want_bias && cb == 0;
            mbar_wait(tfull0, acc_phase);
            tc_fence_after();
            const int m = mt * 128 + quarter * 32 + lane;                   // this thread's output channel
            float* pbase = p.partial + (long long)sp * p.psize;
            const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
            for (int kh = 0; kh < 3; ++kh)
                for (int w = 0; w < 3; ++w) {
                    uint32_t r[32];
                    __syncwarp();
                    tc_ld32(t_lane + (uint32_t)(kh * ACC_N + w * 32), r);
                    tc_wait_ld();
                    float* dst = pbase + ((long long)(kh * 3 + w) * p.Cin + cb * 32) * p.Cout + m;
#pragma unroll
                    for (int j = 0; j < 32; ++j) dst[(long long)j * p.Cout] = __uint_as_float(r[j]);   // 32 lanes = 32 consecutive output channels
                }
            if (do_bias) {
                uint32_t r[32];
                __syncwarp();
                tc_ld32(t_lane + 3u * ACC_N, r);                            // 16 identical columns (+ 16 unused ones)
                tc_wait_ld();
                pbase[wsize + m] = __uint_as_float(r[0]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0);
            acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ rw2 wgrad on an SM PAIR (cta_group::2): Cout % 256 == 0
// The kernel above is bound by the shared-memory port: an M128 x N96 x K16 MMA reads 4 KB of dz and 3 KB of x windows per 48 tensor
// cycles and the TMA fills (45 KB per 36 MMAs) use the same 128 B/cycle (model: 1728 / 2320 cycles = 74 %, measured 66-71 %).
// Here a cluster of two CTAs on the two SMs of a TPC issues ONE M256 x N96 x K16 MMA per step: each SM supplies its own 128
// output channels of dz (the A operand, same shared-memory offsets in both CTAs) and HALF of the N operand -- channels 16*rank ..
// 16*rank+15 of the x block for all three kw windows (SWIZZLE_32B boxes, MN-major blocks of 16 channels one pixel = 32 B apart) --
// so per SM and MMA 4 + 1.5 KB are read and 38 KB filled per stage: 1846 cycles of port time for 1728 of tensor time.
// Protocol (the CUTLASS 2-SM scheme): both producers aim their TMA loads at the LEADER's full barrier (cp.async.bulk.tensor
// .cta_group::2, expect_tx for both halves by the leader), the leader's MMA warp issues tcgen05.mma.cta_group::2 and releases the
// stage / publishes the accumulators with tcgen05.commit .multicast::cluster to BOTH CTAs, every epilogue warp of both CTAs reads
// its own TMEM and arrives on the leader's accumulator-free barrier.
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
// MN-major, SWIZZLE_32B: 16-element (32-byte) rows, blocks of 16 elements LBO apart, 8-row groups SBO apart
__device__ __forceinline__ uint64_t make_mn32_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;                      // SWIZZLE_32B
    return d;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv_tc_wgrad_r2c2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dz, const WgR2Args p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bars = base + RING_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * MAX_STAGES, tfull0 = bars + 16 * MAX_STAGES, tempty0 = tfull0 + 16;
    const uint32_t tmem_slot = tempty0 + 16;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
    const uint32_t ones_addr = bars + 1024u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull0, 1); mbar_init(tempty0, 8);                 // 4 epilogue warps of each CTA
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dz) : "memory");
    }
    {
        uint32_t* ones = reinterpret_cast<uint32_t*>(smem_raw + (ones_addr - raw));
        for (int i = threadIdx.x; i < 512; i += NUM_THREADS) ones[i] = 0x3f803f80u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                              // the peer's barriers exist before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int mpairs = p.mblocks / 2;
    const int utypes = p.cblocks * mpairs;                         // (channel block of x, 256 output channels)
    const int units = utypes * p.splits;
    const int npairs = (int)(gridDim.x >> 1), pair = (int)(blockIdx.x >> 1);
    const int pix_tiles = p.ptx * p.pty * p.ptn;
    const int bw = 8 + 2 * p.dil, bh = p.PH + 2 * p.dil;
    const uint32_t xsub = (uint32_t)(bw * bh) * 32u;              // one part (hi or lo) of this CTA's 16 channels of the x box
    const uint32_t zsub = (uint32_t)(8 * p.PH) * 64u;
    const uint32_t z_off = (2u * xsub + 1023u) & ~1023u;
    const int STAGES = p.stages;
    constexpr uint32_t ACC_N = 96;

    if (warp == 0) {
        if (elect_one_sync()) {
            int stage = 0; uint32_t phase = 0;
            for (int u = pair; u < units; u += npairs) {
                const int ut = u % utypes, sp = u / utypes;
                const int cb = ut % p.cblocks, mt = 2 * (ut / p.cblocks) + (int)rank;
                const int q0 = sp * p.tiles_per_split;
                const int q1 = min(pix_tiles, q0 + p.tiles_per_split);
                for (int q = q0; q < q1; ++q) {
                    const int qx = q % p.ptx; const int r2 = q / p.ptx;
                    const int qy = r2 % p.pty; const int n0 = r2 / p.pty;
                    const int x0 = qx * 8, y0 = qy * p.PH;
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t fb = mapa_u32(full0 + 8 * stage, 0);          // the leader's barrier collects both halves
                    if (leader) mbar_expect_tx(full0 + 8 * stage, 2u * (2u * xsub + 8u * zsub));
                    const uint32_t sa = base + stage * p.stage_bytes;
                    tma_load_5d_c2(sa, &map_x, fb, 16 * (int)rank, x0 - p.dil, y0 - p.dil, n0, 2 * cb);
                    tma_load_5d_c2(sa + z_off, &map_dz, fb, 0, x0, y0, n0, 8 * mt);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && elect_one_sync()) {
            const uint32_t idesc_common = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(256 >> 4) << 24);
            const uint32_t idesc = idesc_common | ((ACC_N >> 3) << 17);
            const uint32_t idesc_bias = idesc_common | ((32u >> 3) << 17);
            int stage = 0; uint32_t phase = 0; uint32_t acc_phase = 0;
            const uint64_t ones_desc = make_mn64_desc(ones_addr, 0u, 512u);
            const int terms = p.terms;
            for (int u = pair; u < units; u += npairs) {
                const int ut = u % utypes, sp = u / utypes;
                const int cb = ut % p.cblocks;
                const bool do_bias = p.want_bias && cb == 0;
                const int q0 = sp * p.tiles_per_split;
                const int q1 = min(pix_tiles, q0 + p.tiles_per_split);
                mbar_wait(tempty0, acc_phase ^ 1);
                tc_fence_after();
                for (int q = q0; q < q1; ++q) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = base + stage * p.stage_bytes;
                    const uint64_t az[2] = {make_mn64_desc(sa + z_off, 2u * zsub, 512u), make_mn64_desc(sa + z_off + zsub, 2u * zsub, 512u)};
                    for (int jp = 0; jp < p.PH / 2; ++jp) {
                        const uint32_t first = (q > q0 || jp > 0) ? 1u : 0u;
                        const uint64_t zo = (uint64_t)(jp * 64);
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
                            const uint32_t xa = sa + (uint32_t)((2 * jp + kh * p.dil) * bw) * 32u;
                            const uint64_t bx[2] = {make_mn32_desc(xa, (uint32_t)p.dil * 32u, (uint32_t)bw * 32u),
                                                    make_mn32_desc(xa + xsub, (uint32_t)p.dil * 32u, (uint32_t)bw * 32u)};
                            const uint32_t d = tmem_base + (uint32_t)kh * ACC_N;
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                if (c >= terms) break;
                                const int ha = c == 1 ? 1 : 0, hb = c == 2 ? 1 : 0;
                                tc_mma_bf16_c2(d, az[ha] + zo, bx[hb], idesc, (first || c > 0) ? 1u : 0u);
                            }
                        }
                        if (do_bias) {
                            const uint32_t d = tmem_base + 3u * ACC_N;
                            tc_mma_bf16_c2(d, az[0] + zo, ones_desc, idesc_bias, first);
                            tc_mma_bf16_c2(d, az[1] + zo, ones_desc, idesc_bias, 1u);
                        }
                    }
                    tc_commit_c2(empty0 + 8 * stage, 3);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit_c2(tfull0, 3);
                acc_phase ^= 1;
            }
        }
    } else {
        const int quarter = warp & 3;
        uint32_t acc_phase = 0;
        const long long wsize = 9LL * p.Cin * p.Cout;
        const uint32_t tempty_leader = mapa_u32(tempty0, 0);
        for (int u = pair; u < units; u += npairs) {
            const int ut = u % utypes, sp = u / utypes;
            const int cb = ut % p.cblocks, mt = 2 * (ut / p.cblocks) + (int)rank;
            const bool do_bias = p.want_bias && cb == 0;
            mbar_wait(tfull0, acc_phase);
            tc_fence_after();
            const int m = mt * 128 + quarter * 32 + lane;                   // this thread's output channel
            float* pbase = p.partial + (long long)sp * p.psize;
            const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
            for (int kh = 0; kh < 3; ++kh)
                for (int h = 0; h < 2; ++h)                                 // accumulator columns: [half of the pair][window][16 channels]
                    for (int w = 0; w < 3; ++w) {
                        uint32_t r[16];
                        __syncwarp();
                        tc_ld16(t_lane + (uint32_t)(kh * ACC_N + h * 48 + w * 16), r);
                        tc_wait_ld();
                        float* dst = pbase + ((long long)(kh * 3 + w) * p.Cin + cb * 32 + h * 16) * p.Cout + m;
#pragma unroll
                        for (int j = 0; j < 16; ++j) dst[(long long)j * p.Cout] = __uint_as_float(r[j]);
                    }
            if (do_bias) {
                uint32_t r[16];
                __syncwarp();
                tc_ld16(t_lane + 3u * ACC_N, r);
                tc_wait_ld();
                pbase[wsize + m] = __uint_as_float(r[0]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader);
            acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                              // no commit / remote arrive is still on its way to a CTA that exits
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// out[i] = sum_z partial[z][i]; the first nw floats go to dw, the remaining (bias) ones to db
__global__ void reduce_partials_kernel(const float* __restrict__ partial, long long psize, long long nw, int splits,
                                       float* __restrict__ dw, float* __restrict__ db) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= psize) return;
    if (i >= nw && db == nullptr) return;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < splits; ++z) {
        float4 v = *reinterpret_cast<const float4*>(partial + (long long)z * psize + i);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    if (i < nw) *reinterpret_cast<float4*>(dw + i) = s;
    else *reinterpret_cast<float4*>(db + (i - nw)) = s;
}

// per-tap transpose: w_t[tap][n][c] = w[tap][c][n] (n < Cout), zero rows for n >= Cout
// fmt = ACT_F32: tf32-rounded floats; ACT_S32: bf16 (hi | lo) pairs per 32 input channels (Cin % 32 == 0)
__global__ void pack_filter_t_kernel(const float* __restrict__ w, int taps, int Cin, int Cout, int cout_pad, int fmt, float* __restrict__ wt) {
    __shared__ float tile[32][33];
    const int tap = blockIdx.z;
    const int c0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, n = n0 + threadIdx.x;
        float v = (c < Cin && n < Cout) ? w[((long long)tap * Cin + c) * Cout + n] : 0.f;
        tile[i][threadIdx.x] = fmt == ACT_F32 ? tf32_rn(v) : v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int n = n0 + i, c = c0 + threadIdx.x;
        if (n < cout_pad && c < Cin) {
            const long long e = ((long long)tap * cout_pad + n) * Cin + c;
            const float v = tile[threadIdx.x][i];
            if (fmt == ACT_F32) wt[e] = v;
            else {
                const float hi = bf16_lo_f(bf16x2_rn(v, 0.f));
                unsigned char* q = s32_addr(wt, e);
                *reinterpret_cast<unsigned short*>(q) = (unsigned short)(__float_as_uint(hi) >> 16);
                *reinterpret_cast<unsigned short*>(q + 64) = (unsigned short)(bf16x2_rn(v - hi, 0.f) & 0xffffu);
            }
        }
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int encode_act_map(CUtensorMap* m, const float* ptr, int B, int H, int W, int C, int TW, int TH, int TN,
                   CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, int estride = 1) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return SSDB_ECUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(TW * estride), (cuuint32_t)(TH * estride), (cuuint32_t)TN};
    cuuint32_t es[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activation %dx%dx%dx%d box %d,%d,%d) failed: %d", B, H, W, C, TW, TH, TN, (int)r); return SSDB_ECUDA; }
    return SSDB_OK;
}

int encode_w_map(CUtensorMap* m, const float* ptr, long long rows, int K, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return SSDB_ECUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(filter rows %lld K %d box %d) failed: %d", rows, K, box_rows, (int)r); return SSDB_ECUDA; }
    return SSDB_OK;
}

struct TileGeom { int TW, TH, TN; double eff; };

// pick the TW x TH x TN destination box (<= 128 pixels) that wastes the fewest MMA rows
TileGeom pick_tile(int B, int H, int W) {
    TileGeom best{1, 1, 1, 0.0};
    const double total = (double)B * H * W;
    for (int tw = 1; tw <= W && tw <= 128; ++tw) {
        for (int th = 1; th <= H && tw * th <= 128; ++th) {
            int tn_max = 128 / (tw * th);
            if (tn_max > B) tn_max = B;
            for (int tn = 1; tn <= tn_max; ++tn) {
                long long tiles = (long long)((W + tw - 1) / tw) * ((H + th - 1) / th) * ((B + tn - 1) / tn);
                double eff = total / (tiles * 128.0);
                // prefer higher efficiency; tie -> wider rows (longer contiguous TMA runs)
                if (eff > best.eff + 1e-9 || (eff > best.eff - 1e-9 && tw > best.TW)) best = TileGeom{tw, th, tn, eff};
            }
        }
    }
    return best;
}

int num_sms() {
    static PerDevice<int> n_pd;
    int& n = n_pd.get();
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
    return n;
}

// Two CTAs per SM for the short-K, narrow-N layers (conv1_1 / conv1_2 / conv2_1 dgrad: N tile <= 64).  Their units are
// dominated by the epilogue (128 x 64 outputs per ~0.6 us of MMA), and one CTA's four epilogue warps cannot keep enough
// bytes in flight; a second resident CTA doubles the epilogue / TMA parallelism.  Each CTA then gets half the ring (2-3
// stages), a 256-column TMEM allocation (two 128-column accumulator buffers) and one M tile per unit.
constexpr int DUAL_RING_BYTES = 92 * 1024;
constexpr int DUAL_SMEM_BYTES = DUAL_RING_BYTES + 1024 + 256 + 4 * 32 * 128;         // 109.25 KB (4 epilogue warps): two fit in 227 KB

int launch_tc(const CUtensorMap& ms, const CUtensorMap& mw, const TcArgs& a_in, cudaStream_t st) {
    static PerDevice<bool> attr_pd;
    bool& attr = attr_pd.get();
    if (!attr) {
        SSDB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<ACT_F32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        SSDB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<ACT_F32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        SSDB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<ACT_S32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        SSDB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<ACT_S32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        SSDB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<ACT_S32, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr = true;
    }
    TcArgs a = a_in;
    a.ring_bytes = RING_BYTES; a.tmem_cols = TMEM_COLS; a.acc_stride = ACC_STRIDE;
    long long m_tiles = (long long)a.tiles_x * a.tiles_y * a.tiles_n;
    int smem = SMEM_BYTES, ctas_per_sm = 1;
    // measured on B200 (vgg300, batch 64; SSDB_TC_DUAL=1 forces it for every N <= 64 layer, =0 disables it):
    //   conv1_1 fprop 0.92 -> 0.67 ms, conv1_2 dgrad 1.74 -> 1.00 ms   (epilogue-bound: K blocks per unit <= 2, or a dgrad
    //   with its mask read and <= 18 K blocks);   conv1_2 fprop 0.87 -> 0.97, conv2_1 dgrad 0.43 -> 0.48 (MMA-heavier units
    //   miss the deeper ring more than they gain from the second CTA) -> those keep one CTA per SM.
    static int dual_mode = -1;
    if (dual_mode < 0) { const char* ov = getenv("SSDB_TC_DUAL"); dual_mode = ov ? (atoi(ov) ? 1 : 0) : 2; }
    int kblocks = 0;
    for (int g = 0; g < a.ngroups; ++g) kblocks += a.g_nt[g] * a.cblocks;
    // (split mode, 8 epilogue warps: conv1_2 dgrad is faster with ONE CTA and two tiles per unit -- 1.37 -> 1.25 ms -- so only the
    //  K <= 64 layer conv1_1 keeps two CTAs there: 0.406 -> 0.387 ms)
    const bool dual_wanted = dual_mode == 1 || (dual_mode == 2 && (kblocks <= 2 || (!a.split && a.mode == 1 && a.mask && kblocks <= 18)));
    if (dual_wanted && a.block_n <= 64 && !a.scatter && !a.resb) {
        const int sb1 = (a.a_slot + a.b_tiles * a.block_n * 128 + 1023) / 1024 * 1024;     // stage with one M tile per unit
        if (2 * sb1 <= DUAL_RING_BYTES && m_tiles * a.n_tiles >= 4LL * num_sms()) {
            a.mtu = 1; a.stage_bytes = sb1;
            a.stages = DUAL_RING_BYTES / sb1; if (a.stages > MAX_STAGES) a.stages = MAX_STAGES;
            a.ring_bytes = DUAL_RING_BYTES; a.tmem_cols = 256; a.acc_stride = 128;
            smem = DUAL_SMEM_BYTES; ctas_per_sm = 2;
        }
    }
    long long total = ((m_tiles + a.mtu - 1) / a.mtu) * a.n_tiles;
    const long long slots = (long long)num_sms() * ctas_per_sm;
    int grid = (int)(total < slots ? total : slots);
    const bool scatter = a.mode == 0 && a.scatter;
    static int wide = -1;
    if (wide < 0) { const char* ov = getenv("SSDB_TC_EPI8"); wide = ov ? (atoi(ov) ? 1 : 0) : 1; }
    const int threads = (ctas_per_sm == 2 || !wide) ? NUM_THREADS : TC_THREADS;
    if (a.c2 && ctas_per_sm == 1) {
        // clusters of two CTAs (the two SMs of a TPC); a pair owns two consecutive units of one N tile
        const long long m_pairs = (m_tiles + a.mtu - 1) / a.mtu;
        const long long pairs = ((m_pairs + 1) / 2) * a.n_tiles, pslots = num_sms() / 2;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(2 * (pairs < pslots ? pairs : pslots))); cfg.blockDim = dim3((unsigned)threads);
        cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        ++::ssdb::g_launches;
        SSDB_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<ACT_S32, false, true>, ms, mw, a));
        return SSDB_OK;
    }
    if (a.split) {
        if (scatter) conv_tc_kernel<ACT_S32, true><<<grid, threads, smem, st>>>(ms, mw, a);
        else conv_tc_kernel<ACT_S32, false><<<grid, threads, smem, st>>>(ms, mw, a);
    } else {
        if (scatter) conv_tc_kernel<ACT_F32, true><<<grid, threads, smem, st>>>(ms, mw, a);
        else conv_tc_kernel<ACT_F32, false><<<grid, threads, smem, st>>>(ms, mw, a);
    }
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

// One or two M tiles per unit?  Two tiles share one B tile (1.5x less L2 -> smem traffic per FLOP, measured 15-20% on
// long-K layers), but halve the number of units (wave quantisation on the small maps) and, when both accumulators fill
// TMEM (N = 256), expose the epilogue -- heavier in dgrad (mask / old-value reads).  Cost model fitted to B200 timings.
// Split mode (refit on B200, 8 epilogue warps, tools/layer_bench.py with SSDB_TC_MTU=1 / 2): a one-tile unit moves 48 KB
// per 32-channel slab and sits at the L2 -> SM rate (~1130 cycles), a two-tile unit moves 64 KB and is MMA-bound (12 MMAs,
// 1536 cycles): a two-tile unit costs 1.36x a one-tile unit, plus an exposed-epilogue term that the batched-load epilogue
// made small.  Two tiles then win everywhere except where halving the unit count costs a wave (conv8_x at batch 64).
int tc_mtu(int block_n, long long m_tiles, int n_tiles, int kblocks, bool dgrad, bool split) {
    if (const char* ov = getenv("SSDB_TC_MTU")) { int v = atoi(ov); if (v == 1 || v == 2) return v; }
    const int sms = num_sms();
    const long long w1 = (m_tiles * n_tiles + sms - 1) / sms;
    const long long w2 = (((m_tiles + 1) / 2) * n_tiles + sms - 1) / sms;
    const int noff = (block_n + 31) & ~31;
    const double e = (2 * noff <= ACC_STRIDE) ? 0.0 : (split ? (dgrad ? 3.0 : 1.0) : (dgrad ? 15.0 : 4.0));
    const double c2 = 2.0 * ((split ? 0.70 : 0.85) + e / (double)(kblocks > 0 ? kblocks : 1));
    return (double)w2 * c2 < (double)w1 ? 2 : 1;
}

// fill the tap-group table: plain (one tap per group) or row-window (3 taps of a filter row per group, TW must be 8)
void fill_groups(TcArgs& a, int k, const int* dy, const int* dx, bool row_window) {
    if (k > 3) k = 3;                                            // callers require k <= 3 (the tables hold 9 taps)
    const int taps = k * k;
    if (!row_window) {
        a.ngroups = taps;
        for (int t = 0; t < taps && t < 9; ++t) {
            a.g_dy[t] = (signed char)dy[t]; a.g_dx[t] = (signed char)dx[t]; a.g_nt[t] = 1; a.g_win[t][0] = 0; a.g_w[t][0] = (unsigned char)t;
        }
        a.a_rows = a.TW * a.TH * a.TN; a.a_slot = A_BYTES; a.a_sbo = 1024; a.b_tiles = 1;
    } else {
        a.ngroups = k;
        for (int kh = 0; kh < k; ++kh) {
            int dmin = dx[kh * k];
            for (int kw = 1; kw < k; ++kw) if (dx[kh * k + kw] < dmin) dmin = dx[kh * k + kw];
            a.g_dy[kh] = (signed char)dy[kh * k]; a.g_dx[kh] = (signed char)dmin; a.g_nt[kh] = (unsigned char)k;
            for (int kw = 0; kw < k; ++kw) { a.g_win[kh][kw] = (unsigned char)(dx[kh * k + kw] - dmin); a.g_w[kh][kw] = (unsigned char)(kh * k + kw); }
        }
        int boxw = a.TW + k - 1;
        if (const char* ov = getenv("SSDB_KW3_BOXW")) { int v = atoi(ov); if (v >= boxw && v <= 32) boxw = v; }
        a.a_rows = boxw * a.TH * a.TN; a.a_slot = (a.a_rows * 128 + 1023) / 1024 * 1024; a.a_sbo = boxw * 128; a.b_tiles = k;
    }
    a.use_bo = 0;
    if (const char* ov = getenv("SSDB_KW3_BO")) a.use_bo = atoi(ov) ? 1 : 0;
    a.stage_bytes = a.mtu * a.a_slot + a.b_tiles * a.block_n * 128;
    a.stage_bytes = (a.stage_bytes + 1023) / 1024 * 1024;
    a.stages = RING_BYTES / a.stage_bytes; if (a.stages > MAX_STAGES) a.stages = MAX_STAGES;
}

// Resident-filter window mode (TcArgs::resb): a 3x3 / stride-1 / dilation-1 layer in split mode whose WHOLE filter (one N tile)
// fits in shared memory beside at least two full-window source boxes -- conv1_2, fprop and dgrad (144 KB of filter + 2 x 23 KB).
// Measured before (ncu, conv1_2 fprop, batch 64): 9 GB of L2 -> SM traffic per launch for a 1.5 GB input, 37 % of it the
// filter re-loaded for every unit, the rest the input read once per filter row; the N = 64 MMAs already use 96 of the 128 B/cycle
// of the shared-memory port for their operand reads, and the TMA fills competed for the rest.  Tiles are 8 x 16 pixels of ONE
// image (the box carries a halo line above and below, so boxes of two images cannot be stacked with a constant row pitch).
bool try_resident_filter(TcArgs& a, int k, const int* dy, const int* dx, int B, int Hd, int Wd, bool split, double cur_eff) {
    static int on = -1;
    if (on < 0) { const char* ov = getenv("SSDB_RESIDENT_FILTER"); on = ov ? (atoi(ov) ? 1 : 0) : 1; }
    if (!on || !split || k != 3 || a.n_tiles != 1 || Hd < 16) return false;
    const int boxw = 8 + 2, boxh = 16 + 2;
    const int slot = (boxw * boxh * 128 + 1023) / 1024 * 1024;
    const int res = 9 * a.cblocks * a.block_n * 128;
    if (res % 1024 != 0 || res + 2 * slot > RING_BYTES) return false;
    const long long tiles = (long long)((Wd + 7) / 8) * ((Hd + 15) / 16) * B;
    const double eff = (double)B * Hd * Wd / (tiles * 128.0);
    if (eff < 0.9 || eff < 0.95 * cur_eff) return false;
    int dymin = dy[0], dxmin = dx[0];
    for (int t = 1; t < 9; ++t) { if (dy[t] < dymin) dymin = dy[t]; if (dx[t] < dxmin) dxmin = dx[t]; }
    a.TW = 8; a.TH = 16; a.TN = 1;
    a.tiles_x = (Wd + 7) / 8; a.tiles_y = (Hd + 15) / 16; a.tiles_n = B;
    a.mtu = 1; a.resb = 1; a.res_bytes = res;
    a.ngroups = 1; a.g_dy[0] = (signed char)dymin; a.g_dx[0] = (signed char)dxmin; a.g_nt[0] = 9;
    for (int t = 0; t < 9; ++t) a.r_win[t] = (unsigned char)((dy[t] - dymin) * boxw + (dx[t] - dxmin));
    a.a_rows = boxw * boxh; a.a_slot = slot; a.a_sbo = boxw * 128; a.b_tiles = 0;
    a.stage_bytes = slot;
    a.stages = (RING_BYTES - res) / slot; if (a.stages > MAX_STAGES) a.stages = MAX_STAGES;
    return true;
}

// row-window geometry: TW = 8 and TH*TN = 16 (128 rows); returns efficiency, 0 if impossible
double pick_window_tile(int B, int H, int W, int* th, int* tn) {
    double best = 0.0;
    for (int t = 1; t <= 16; t <<= 1) {
        int h = 16 / t;                          // TH = h, TN = t
        if (h > H && h > 1) continue;
        if (t > B) continue;
        long long tiles = (long long)((W + 7) / 8) * ((H + h - 1) / h) * ((B + t - 1) / t);
        double eff = (double)B * H * W / (tiles * 128.0);
        if (eff > best) { best = eff; *th = h; *tn = t; }
    }
    return best;
}

bool want_row_window(const ConvGeom& g, int block_n, int B, int Hd, int Wd, double plain_eff, int* th, int* tn) {
    if (const char* ov = getenv("SSDB_KW3")) { if (atoi(ov) == 0) return false; }
    if (g.k != 3 || g.dil != 1 || g.stride != 1 || block_n > 128) return false;
    double e = pick_window_tile(B, Hd, Wd, th, tn);
    return e >= 0.9 * plain_eff && e >= 0.6;
}

int block_n_for(int channels) {
    int n = (channels + 15) / 16 * 16;
    return n > MAX_N ? MAX_N : n;
}

bool tc_common_ok(const ConvGeom& g) {
    return (g.stride == 1 || g.stride == 2) && g.pad_t == g.pad_l && g.Cin % 32 == 0 && g.Cout % 16 == 0 && g.k >= 1 && g.k <= 3;
}

}  // namespace

bool conv_tc_supported_fprop(const ConvGeom& g) {
    if (!tc_common_ok(g)) return false;
    if (g.Cout > MAX_N && g.Cout % MAX_N != 0) return false;
    TileGeom t = pick_tile(g.B, g.Ho, g.Wo);
    return t.eff >= 0.45;
}

bool conv_tc_supported_dgrad(const ConvGeom& g) {
    if (!tc_common_ok(g) || g.Cout % 32 != 0) return false;
    if (g.Cin > MAX_N && g.Cin % MAX_N != 0) return false;
    const int s = g.stride;
    if (s == 2) {
        // every parity class of the destination needs at least one tap (true for k = 3, false for 1x1 stride 2)
        for (int py = 0; py < s; ++py) {
            bool any = false;
            for (int kh = 0; kh < g.k; ++kh) if ((((py + g.pad_t - kh * g.dil) % s) + s) % s == 0) any = true;
            if (!any) return false;
        }
    }
    TileGeom t = pick_tile(g.B, (g.H + s - 1) / s, (g.W + s - 1) / s);
    return t.eff >= 0.45;
}

int pack_filter_t(const float* w_hwio, int taps, int Cin, int Cout, int cout_pad, int fmt, float* w_t, cudaStream_t st) {
    SSDB_REQUIRE(fmt == ACT_F32 || Cin % 32 == 0, "split filter copies need Cin % 32 == 0");
    dim3 grid((Cin + 31) / 32, (cout_pad + 31) / 32, taps), block(32, 8);
    pack_filter_t_kernel<<<grid, block, 0, st>>>(w_hwio, taps, Cin, Cout, cout_pad, fmt, w_t);
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

// row-window mode with 64 < N <= 128: one M tile per unit keeps three 68 KB stages but re-reads the whole filter for
// every 128 pixels (the L2 -> SM traffic of conv2_x and the 4-box heads is then 70% filter bytes, and the kernel sits at
// the L2 -> SM rate); two M tiles share the filter tile at the price of a two-stage ring of 88 KB stages.  Measured on
// B200 (split mode, batch 64): conv2_1 fprop 0.605 -> 0.442 ms, conv2_2 fprop 1.140 -> 0.824, dgrad 1.169 -> 0.896,
// head0 fprop 0.358 -> 0.237.  SSDB_RW_MTU2=0 restores one tile per unit.
static int rw_wide_mtu(int mtu_from_cost_model) {
    const char* ov = getenv("SSDB_RW_MTU2");
    if (ov && !atoi(ov)) return 1;
    (void)mtu_from_cost_model;
    return 2;
}

static int split_terms_env() {
    const char* ov = getenv("SSDB_SPLIT_TERMS");       // bring-up switch, read per call
    const int t = ov ? atoi(ov) : 3;
    return (t < 1 || t > 3) ? 3 : t;
}

// Pair mode of the fprop / dgrad kernel (TcArgs::c2): split operands, N tile >= 128, enough units for two waves of the 74 SM pairs.
// Measured on B200 (vgg300, batch 64, same box, SSDB_TC_PAIR=0 / 1): see DESIGN.md section 4.
static void maybe_pair(TcArgs& a, bool split, bool scatter) {
    const char* ov = getenv("SSDB_TC_PAIR");                    // read per call (tests compare both paths): 0 off, 1 fprop + dgrad, 2 fprop only
    const int on = ov ? atoi(ov) : 1;
    if (!on || (on == 2 && a.mode != 0) || !split || scatter || a.resb || a.block_n < 128 || a.block_n % 16 != 0 || num_sms() % 2 != 0) return;
    const long long m_tiles = (long long)a.tiles_x * a.tiles_y * a.tiles_n;
    const long long units = ((m_tiles + a.mtu - 1) / a.mtu) * a.n_tiles;
    if (units < 2LL * num_sms()) return;
    a.c2 = 1;
    a.stage_bytes = (a.mtu * a.a_slot + a.b_tiles * a.block_n * 64 + 1023) / 1024 * 1024;
    a.stages = RING_BYTES / a.stage_bytes; if (a.stages > MAX_STAGES) a.stages = MAX_STAGES;
}

// can this fprop write the 2x2 / stride-2 max pool of its output instead of the output (see TcArgs::pool)?
bool conv_tc_fprop_can_pool(const ConvGeom& g, int fmt) {
    if (fmt != ACT_S32 || !conv_tc_supported_fprop(g) || (g.Ho & 1) || (g.Wo & 1) || g.Cout % 32 != 0) return false;
    const int block_n = block_n_for(g.Cout);
    TileGeom t = pick_tile(g.B, g.Ho, g.Wo);
    int wth = 0, wtn = 0;
    if (!want_row_window(g, block_n, g.B, g.Ho, g.Wo, t.eff, &wth, &wtn)) return false;
    return wth >= 2 && (wth & 1) == 0;
}

int conv_tc_fprop(const ConvGeom& g, const float* x, const float* w_t, int cout_pad, int fmt, const ConvEpilogue& ep, float* y, cudaStream_t st) {
    SSDB_REQUIRE(conv_tc_supported_fprop(g), "shape not supported by the tcgen05 fprop kernel");
    SSDB_REQUIRE(!ep.pool_dst || conv_tc_fprop_can_pool(g, fmt), "this layer cannot fuse its max pool");
    TileGeom t = pick_tile(g.B, g.Ho, g.Wo);
    TcArgs a{};
    a.TW = t.TW; a.TH = t.TH; a.TN = t.TN;
    a.tiles_x = (g.Wo + t.TW - 1) / t.TW; a.tiles_y = (g.Ho + t.TH - 1) / t.TH; a.tiles_n = (g.B + t.TN - 1) / t.TN;
    a.block_n = block_n_for(g.Cout); a.n_tiles = (g.Cout + a.block_n - 1) / a.block_n;
    SSDB_REQUIRE(cout_pad >= a.n_tiles * a.block_n, "transposed filter is not padded enough");
    a.Hd = g.Ho; a.Wd = g.Wo; a.Bn = g.B; a.Cd = g.Cout; a.cd_valid = g.Cout;
    SSDB_REQUIRE(g.k <= 3, "tcgen05 path supports 1x1 and 3x3 filters");
    a.cblocks = g.Cin / BLOCK_K; a.rows_per_tap = cout_pad;
    int tdy[9], tdx[9];
    for (int t = 0; t < g.k * g.k; ++t) { tdy[t] = (t / g.k) * g.dil - g.pad_t; tdx[t] = (t % g.k) * g.dil - g.pad_l; }
    a.sstride = g.stride; a.dscale = 1; a.dpy = a.dpx = 0; a.mode = 0;
    a.mtu = tc_mtu(a.block_n, (long long)a.tiles_x * a.tiles_y * a.tiles_n, a.n_tiles, g.k * g.k * a.cblocks, false, fmt == ACT_S32);
    int wth = 0, wtn = 0;
    const bool rw = want_row_window(g, a.block_n, g.B, g.Ho, g.Wo, t.eff, &wth, &wtn);
    if (rw) {
        t.TW = 8; t.TH = wth; t.TN = wtn; a.TW = 8; a.TH = wth; a.TN = wtn;
        a.tiles_x = (g.Wo + 7) / 8; a.tiles_y = (g.Ho + wth - 1) / wth; a.tiles_n = (g.B + wtn - 1) / wtn;
        if (a.block_n > 64) a.mtu = rw_wide_mtu(a.mtu);        // 2 x 20 KB + 3 x 16 KB would leave only two stages
    }
    fill_groups(a, g.k, tdy, tdx, rw);
    const bool resb = rw && !ep.scatter && try_resident_filter(a, g.k, tdy, tdx, g.B, g.Ho, g.Wo, fmt == ACT_S32,
                                                               (double)g.B * g.Ho * g.Wo / ((double)a.tiles_x * a.tiles_y * a.tiles_n * 128.0));
    if (resb) { t.TW = 8; t.TH = 16; t.TN = 1; }
    maybe_pair(a, fmt == ACT_S32, ep.scatter != 0);
    SSDB_REQUIRE(g.pad_t == g.pad_l, "tcgen05 path assumes equal top/left padding");
    a.dst = y; a.bias = ep.bias; a.mask = nullptr; a.relu = ep.relu; a.beta = 0; a.round_out = ep.round_tf32;
    a.split = fmt == ACT_S32 ? 1 : 0; a.split_terms = split_terms_env();
    a.pool = ep.pool_dst ? 1 : 0; a.pool_dst = ep.pool_dst; a.pool_code = ep.pool_code; a.Hp = g.Ho / 2; a.Wp = g.Wo / 2;
    a.scatter = ep.scatter; a.V = ep.V; a.n_valid = ep.n_valid; a.anchor_base = ep.anchor_base; a.A = ep.A;
    CUtensorMap ms, mw;
    int rc = encode_act_map(&ms, x, g.B, g.H, g.W, g.Cin, rw ? a.a_sbo / 128 : t.TW, resb ? t.TH + 2 : t.TH, t.TN, CU_TENSOR_MAP_SWIZZLE_128B, g.stride); if (rc) return rc;
    rc = encode_w_map(&mw, w_t, (long long)g.k * g.k * cout_pad, g.Cin, a.c2 ? a.block_n / 2 : a.block_n); if (rc) return rc;
    return launch_tc(ms, mw, a, st);
}

int conv_tc_dgrad(const ConvGeom& g, const float* dz, const float* w_hwio, int fmt, const float* mask_x, int beta, int round_out, float* dx, cudaStream_t st) {
    SSDB_REQUIRE(conv_tc_supported_dgrad(g), "shape not supported by the tcgen05 dgrad kernel");
    // A conv with stride s scatters: destination pixel (y, x) only sees taps with (y + pad - kh*dil) % s == 0.  Launch one
    // unit-stride contraction per parity class (py, px) of the destination, each with its own tap subset.
    const int s = g.stride;
    for (int py = 0; py < s; ++py) {
        for (int px = 0; px < s; ++px) {
            const int Hc = (g.H - py + s - 1) / s, Wc = (g.W - px + s - 1) / s;       // pixels of this class
            if (Hc <= 0 || Wc <= 0) continue;
            TileGeom t = pick_tile(g.B, Hc, Wc);
            TcArgs a{};
            a.TW = t.TW; a.TH = t.TH; a.TN = t.TN;
            a.tiles_x = (Wc + t.TW - 1) / t.TW; a.tiles_y = (Hc + t.TH - 1) / t.TH; a.tiles_n = (g.B + t.TN - 1) / t.TN;
            a.block_n = block_n_for(g.Cin); a.n_tiles = (g.Cin + a.block_n - 1) / a.block_n;
            a.Hd = g.H; a.Wd = g.W; a.Bn = g.B; a.Cd = g.Cin; a.cd_valid = g.Cin;
            a.cblocks = g.Cout / BLOCK_K; a.rows_per_tap = g.Cin;
            a.sstride = 1; a.dscale = s; a.dpy = py; a.dpx = px; a.mode = 1;
            a.mtu = tc_mtu(a.block_n, (long long)a.tiles_x * a.tiles_y * a.tiles_n, a.n_tiles, g.k * g.k * a.cblocks / (s * s), true, fmt == ACT_S32);
            SSDB_REQUIRE(g.k <= 3, "tcgen05 path supports 1x1 and 3x3 filters");
            int wth = 0, wtn = 0;
            bool resb = false;
            const bool rw = s == 1 && want_row_window(g, a.block_n, g.B, Hc, Wc, t.eff, &wth, &wtn);
            if (rw) {
                t.TW = 8; t.TH = wth; t.TN = wtn; a.TW = 8; a.TH = wth; a.TN = wtn;
                a.tiles_x = (Wc + 7) / 8; a.tiles_y = (Hc + wth - 1) / wth; a.tiles_n = (g.B + wtn - 1) / wtn;
                if (a.block_n > 64) a.mtu = rw_wide_mtu(a.mtu);
                int tdy[9], tdx[9];
                for (int tt = 0; tt < g.k * g.k; ++tt) { tdy[tt] = g.pad_t - (tt / g.k) * g.dil; tdx[tt] = g.pad_l - (tt % g.k) * g.dil; }
                fill_groups(a, g.k, tdy, tdx, true);
                resb = try_resident_filter(a, g.k, tdy, tdx, g.B, Hc, Wc, fmt == ACT_S32,
                                           (double)g.B * Hc * Wc / ((double)a.tiles_x * a.tiles_y * a.tiles_n * 128.0));
                if (resb) { t.TW = 8; t.TH = 16; t.TN = 1; }
            } else {
                int ntap = 0;
                a.ngroups = 0;
                for (int kh = 0; kh < g.k; ++kh) {
                    int ny = py + g.pad_t - kh * g.dil;
                    if (((ny % s) + s) % s) continue;
                    for (int kw = 0; kw < g.k; ++kw) {
                        int nx = px + g.pad_l - kw * g.dil;
                        if (((nx % s) + s) % s) continue;
                        // floor division: ny, nx are multiples of s here
                        a.g_dy[ntap] = (signed char)(ny / s); a.g_dx[ntap] = (signed char)(nx / s); a.g_nt[ntap] = 1;
                        a.g_win[ntap][0] = 0; a.g_w[ntap][0] = (unsigned char)(kh * g.k + kw); ++ntap;
                    }
                }
                SSDB_REQUIRE(ntap > 0, "tcgen05 dgrad: a destination parity class without any filter tap");
                a.ngroups = ntap;
                a.a_rows = a.TW * a.TH * a.TN; a.a_slot = A_BYTES; a.a_sbo = 1024; a.b_tiles = 1; a.use_bo = 0;
                a.stage_bytes = a.mtu * a.a_slot + a.block_n * 128; a.stage_bytes = (a.stage_bytes + 1023) / 1024 * 1024;
                a.stages = RING_BYTES / a.stage_bytes; if (a.stages > MAX_STAGES) a.stages = MAX_STAGES;
            }
            a.dst = dx; a.bias = nullptr; a.mask = mask_x; a.relu = 0; a.beta = beta; a.round_out = round_out;
            a.split = fmt == ACT_S32 ? 1 : 0; a.split_terms = split_terms_env();
            maybe_pair(a, fmt == ACT_S32, false);
            CUtensorMap ms, mw;
            int rc = encode_act_map(&ms, dz, g.B, g.Ho, g.Wo, g.Cout, rw ? a.a_sbo / 128 : t.TW, resb ? t.TH + 2 : t.TH, t.TN); if (rc) return rc;
            rc = encode_w_map(&mw, w_hwio, (long long)g.k * g.k * g.Cin, g.Cout, a.c2 ? a.block_n / 2 : a.block_n); if (rc) return rc;
            rc = launch_tc(ms, mw, a, st); if (rc) return rc;
        }
    }
    return SSDB_OK;
}

namespace {

struct PixGeom { int PW, PH, PN, P; double eff; };

// pixel box for the wgrad contraction: P = PW*PH*PN a multiple of 8, at most p_max, least waste
PixGeom pick_pix(int B, int H, int W, int p_max, int p_mult = 8) {
    PixGeom best{0, 0, 0, 0, 0.0};
    const double total = (double)B * H * W;
    for (int pw = 1; pw <= W && pw <= p_max; ++pw)
        for (int ph = 1; ph <= H && pw * ph <= p_max; ++ph)
            for (int pn = 1; pn <= B && pw * ph * pn <= p_max; ++pn) {
                int P = pw * ph * pn;
                if (P % p_mult) continue;
                long long tiles = (long long)((W + pw - 1) / pw) * ((H + ph - 1) / ph) * ((B + pn - 1) / pn);
                double eff = total / ((double)tiles * P);
                // a pipeline stage is one box: small boxes starve the tensor core (one MMA per 8 pixels and a
                // barrier round trip per stage), so the box size weighs more than a few percent of padding
                // measured on B200: boxes that span more than 4 images run ~2.4x slower (each image is megabytes
                // away: the TMA unit walks far-apart pages), so those only win when nothing else fits
                auto weight = [](int P_, int pn_) {
                    double w = P_ >= 32 ? 1.0 + 0.05 * (P_ - 32) / 32.0 : 0.25 + 0.5 * P_ / 32.0;
                    return pn_ > 4 ? 0.45 * w : w;
                };
                double score = eff * weight(P, pn);
                double bscore = best.P ? best.eff * weight(best.P, best.PN) : -1.0;
                if (score > bscore) best = PixGeom{pw, ph, pn, P, eff};
            }
    return best;
}

// 5-D view of an NHWC tensor: (c within a 32-channel block, x, y, image, channel block)
int encode_act_map5(CUtensorMap* m, const float* ptr, int B, int H, int W, int C, int PW, int PH, int PN, int nblk, int estride) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return SSDB_ECUDA; }
    cuuint64_t dims[5] = {32, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)(C / 32)};
    cuuint64_t strides[4] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4, 128};
    cuuint32_t box[5] = {32, (cuuint32_t)(PW * estride), (cuuint32_t)(PH * estride), (cuuint32_t)PN, (cuuint32_t)nblk};
    cuuint32_t es[5] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(5-D activation %dx%dx%dx%d box %d,%d,%d,%d) failed: %d", B, H, W, C, PW, PH, PN, nblk, (int)r); return SSDB_ECUDA; }
    return SSDB_OK;
}

// 5-D view of an ACT_S32 tensor for the MN-major (wgrad) operands: (32 bf16 of one part, x, y, image, q = 2 * channel block + part)
int encode_act_map5s(CUtensorMap* m, const float* ptr, int B, int H, int W, int C, int bw, int bh, int bn, int nq, int estride, int inner = 32) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return SSDB_ECUDA; }
    cuuint64_t dims[5] = {32, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)(C / 32 * 2)};
    cuuint64_t strides[4] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4, 64};
    cuuint32_t box[5] = {(cuuint32_t)inner, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn, (cuuint32_t)nq};     // inner = 16: half a block, SWIZZLE_32B
    cuuint32_t es[5] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<float*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, inner == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(split 5-D activation %dx%dx%dx%d box %d,%d,%d,%d) failed: %d", B, H, W, C, bw, bh, bn, nq, (int)r); return SSDB_ECUDA; }
    return SSDB_OK;
}

// 8 KB of bf16 1.0 followed by 8 KB of zeros: the (hi, lo) sub-blocks of the split kernels' bias slot
const float* ones_buffer_split() {
    static PerDevice<float*> d_pd;
    float*& d = d_pd.get();
    if (!d) {
        std::vector<uint32_t> h(4096, 0u);
        for (int i = 0; i < 2048; ++i) h[i] = 0x3f803f80u;
        if (cudaMalloc(reinterpret_cast<void**>(&d), h.size() * sizeof(uint32_t)) != cudaSuccess) return nullptr;
        cudaMemcpy(d, h.data(), h.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
    }
    return d;
}

const float* ones_buffer() {
    static PerDevice<float*> d_pd;
    float*& d = d_pd.get();
    if (!d) {
        std::vector<float> h(64 * 32, 1.0f);
        if (cudaMalloc(reinterpret_cast<void**>(&d), h.size() * sizeof(float)) != cudaSuccess) return nullptr;
        cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
    }
    return d;
}

struct WgRwPlan { WgRwArgs a; bool ok; };

// window plan: 3x3, dilation 1, stride 1, SAME, few output channels
WgRwPlan plan_wgrad_rw(const ConvGeom& g, int fmt) {
    WgRwPlan pl{}; pl.ok = false;
    if (const char* ov = getenv("SSDB_WG_RW")) { if (atoi(ov) == 0) return pl; }
    if (g.k != 3 || g.dil != 1 || g.stride != 1 || g.pad_t != 1 || g.pad_l != 1 || g.Cin % 32 != 0 || g.Cout % 32 != 0) return pl;
    // Measured on B200: MN-major tf32 MMAs issue at about half the K-major rate (a 128x256x8 MMA takes ~260 cycles), and the
    // plain wgrad kernel already runs the N = 256 layers at that limit; the window kernel pays 25% dummy-window MMAs for its
    // lower L2 traffic, which only wins where the plain kernel is traffic-bound: N <= 128 (SSDB_WG_RW_MAXN to experiment).
    { int maxn = 128; if (const char* ov = getenv("SSDB_WG_RW_MAXN")) maxn = atoi(ov); if (g.Cout > maxn) return pl; }
    if (g.Cout > 128 && g.Cout % 128 != 0) return pl;
    if (g.Ho != g.H || g.Wo != g.W) return pl;                // SAME 3x3 only; the VALID tails use the plain kernel
    WgRwArgs& a = pl.a;
    a.block_n = g.Cout > 128 ? 128 : g.Cout; a.n_tiles = g.Cout / a.block_n;
    a.Cin = g.Cin; a.Cout = g.Cout; a.pad = g.pad_t;
    a.cblocks = g.Cin / 32;
    const int noff = (a.block_n + 31) & ~31;
    a.cpu = a.cblocks < 4 ? a.cblocks : 4;
    while (a.cpu > 1 && ((3 * a.cpu + 1) * noff > TMEM_COLS || a.cblocks % a.cpu)) --a.cpu;   // 3 x cpu taps-rows + the bias accumulator
    if ((3 * a.cpu + 1) * noff > TMEM_COLS) return pl;
    a.cgroups = a.cblocks / a.cpu;
    // output lines per stage: as many as fit (<= 8) with at least 3 stages in the ring
    // (split kernel: one K = 16 MMA consumes two lines, so PH is even)
    const int ph_step = fmt == ACT_S32 ? 2 : 1;
    int PH = g.H < 8 ? (g.H + ph_step - 1) / ph_step * ph_step : 8;
    for (;; PH -= ph_step) {
        int sb = (a.cpu * 11 * (PH + 2) * 128 + 1023) / 1024 * 1024 + (a.block_n / 32) * 8 * PH * 128;
        sb = (sb + 1023) / 1024 * 1024;
        if (RING_BYTES / sb >= 3 || PH <= ph_step) { a.stage_bytes = sb; break; }
    }
    if (PH < 2 || RING_BYTES / a.stage_bytes < 2) return pl;
    a.terms = split_terms_env();
    a.PH = PH;
    a.stages = RING_BYTES / a.stage_bytes; if (a.stages > MAX_STAGES) a.stages = MAX_STAGES;
    a.ptx = (g.W + 7) / 8; a.pty = (g.H + PH - 1) / PH; a.ptn = g.B;
    double eff = (double)g.H * g.W / ((double)a.ptx * a.pty * 8 * PH);
    if (eff < 0.5) return pl;
    a.psize = 9LL * g.Cin * g.Cout + g.Cout;
    a.want_bias = 1;
    long long pix_tiles = (long long)a.ptx * a.pty * a.ptn;
    long long ut = (long long)a.cgroups * a.n_tiles;
    long long want = (2LL * num_sms() + ut - 1) / ut;
    long long max_splits = (pix_tiles + 7) / 8;
    if (want > max_splits) want = max_splits;
    if (want < 1) want = 1;
    if (want > 512) want = 512;
    a.tiles_per_split = (int)((pix_tiles + want - 1) / want);
    a.splits = (int)((pix_tiles + a.tiles_per_split - 1) / a.tiles_per_split);
    a.partial = nullptr;
    pl.ok = true;
    return pl;
}

// 5-D maps of the row-window kernel: x box (32, 11, PH, PN, cpu) and dz box (32, 8, PH, PN, N/32)
int encode_rw_map(CUtensorMap* m, const float* ptr, int B, int H, int W, int C, int bw, int PH, int PN, int nblk) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return SSDB_ECUDA; }
    cuuint64_t dims[5] = {32, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)(C / 32)};
    cuuint64_t strides[4] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4, 128};
    cuuint32_t box[5] = {32, (cuuint32_t)bw, (cuuint32_t)PH, (cuuint32_t)PN, (cuuint32_t)nblk};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(row-window %dx%dx%dx%d box %d,%d,%d,%d) failed: %d", B, H, W, C, bw, PH, PN, nblk, (int)r); return SSDB_ECUDA; }
    return SSDB_OK;
}

// Split the pixel range of a weight gradient so that the units fill whole waves of the persistent grid: between 2 and 8 units
// per SM, the count whose last wave is fullest (every split adds one partial filter to write and reduce, so fewer wins ties).
// `ut` = units per split, `min_tiles` = fewest pixel tiles a unit should still own.
// upper bound of what pick_splits can return for ANY batch size (the engine sizes its partial-sum workspace at max_batch, and a
// smaller batch may well pick more splits)
long long max_splits_bound(long long ut, int max_splits_cap) {
    long long hi = (8LL * num_sms() + ut - 1) / ut;
    if (hi > max_splits_cap) hi = max_splits_cap;
    return hi < 1 ? 1 : hi;
}

void pick_splits(long long pix_tiles, long long ut, int min_tiles, int max_splits_cap, int* tiles_per_split, int* splits, long long slots = 0) {
    const long long sms = slots > 0 ? slots : num_sms();     // CTAs (or CTA pairs) of the persistent grid
    long long max_splits = (pix_tiles + min_tiles - 1) / min_tiles;
    if (max_splits > max_splits_cap) max_splits = max_splits_cap;
    if (max_splits < 1) max_splits = 1;
    long long lo = (2 * sms + ut - 1) / ut, hi = (8 * sms + ut - 1) / ut;
    if (lo < 1) lo = 1;
    if (hi > max_splits) hi = max_splits;
    if (lo > hi) lo = hi;
    long long pick = lo; double pick_eff = 0.0;
    for (long long sp = lo; sp <= hi; ++sp) {
        const long long tps = (pix_tiles + sp - 1) / sp, real = (pix_tiles + tps - 1) / tps;
        const long long units = real * ut, waves = (units + sms - 1) / sms;
        const double eff = (double)units / (double)(waves * sms);
        if (eff > pick_eff + 0.02) { pick_eff = eff; pick = sp; }
    }
    *tiles_per_split = (int)((pix_tiles + pick - 1) / pick);
    *splits = (int)((pix_tiles + *tiles_per_split - 1) / *tiles_per_split);
}

struct WgR2Plan { WgR2Args a; bool ok; };

// rw2 plan: 3x3, stride 1, SAME (pad = dilation), split operands, Cout a multiple of 128, a pixel tiling that wastes little
WgR2Plan plan_wgrad_r2(const ConvGeom& g, int fmt) {
    WgR2Plan pl{}; pl.ok = false;
    if (fmt != ACT_S32) return pl;
    if (const char* ov = getenv("SSDB_WG_R2")) { if (atoi(ov) == 0) return pl; }
    if (g.k != 3 || g.stride != 1 || g.pad_t != g.dil || g.pad_l != g.dil || g.Cin % 32 != 0 || g.Cout % 128 != 0) return pl;
    if (g.Ho != g.H || g.Wo != g.W || g.dil < 1 || g.dil > 2) return pl;
    WgR2Args& a = pl.a;
    a.dil = g.dil; a.Cin = g.Cin; a.Cout = g.Cout; a.cblocks = g.Cin / 32; a.mblocks = g.Cout / 128;
    a.ptx = (g.W + 7) / 8; a.ptn = g.B;
    // even PH <= 8: the tallest box within 8 % of the best pixel efficiency (a stage of a 2-line box is only 16 pixels of MMA work)
    double eff_of[5] = {0, 0, 0, 0, 0}, best = 0.0;
    for (int ph = 2; ph <= 8; ph += 2) {
        const int pty = (g.H + ph - 1) / ph;
        eff_of[ph / 2] = (double)g.H * g.W / ((double)a.ptx * 8 * pty * ph);
        if (eff_of[ph / 2] > best) best = eff_of[ph / 2];
    }
    int PH = 0;
    for (int ph = 8; ph >= 2; ph -= 2) if (eff_of[ph / 2] >= 0.92 * best) { PH = ph; break; }
    double min_eff = 0.8;
    if (const char* ov = getenv("SSDB_WG_R2_EFF")) min_eff = atof(ov);
    if (PH == 0 || eff_of[PH / 2] < min_eff) return pl;
    a.PH = PH; a.pty = (g.H + PH - 1) / PH;
    const int bw = 8 + 2 * g.dil, bh = PH + 2 * g.dil;
    int sb = (2 * bw * bh * 64 + 1023) / 1024 * 1024 + 8 * 8 * PH * 64;
    sb = (sb + 1023) / 1024 * 1024;
    a.stage_bytes = sb;
    a.stages = RING_BYTES / sb; if (a.stages > MAX_STAGES) a.stages = MAX_STAGES;
    if (a.stages < 2) return pl;
    a.psize = 9LL * g.Cin * g.Cout + g.Cout;
    a.want_bias = 1; a.terms = split_terms_env();
    const long long pix_tiles = (long long)a.ptx * a.pty * a.ptn;
    const long long ut = (long long)a.cblocks * a.mblocks;
    pick_splits(pix_tiles, ut, 8, 512, &a.tiles_per_split, &a.splits);
    // SM-pair variant (cta_group::2): 256 output channels per unit, each CTA loads 16 of the 32 x channels
    int c2 = 1;
    if (const char* ov = getenv("SSDB_WG_C2")) c2 = atoi(ov) ? 1 : 0;
    if (c2 && g.Cout % 256 == 0 && num_sms() % 2 == 0) {
        int sb2 = (2 * bw * bh * 32 + 1023) / 1024 * 1024 + 8 * 8 * PH * 64;
        sb2 = (sb2 + 1023) / 1024 * 1024;
        a.c2 = 1; a.stage_bytes = sb2;
        a.stages = RING_BYTES / sb2; if (a.stages > MAX_STAGES) a.stages = MAX_STAGES;
        pick_splits(pix_tiles, ut / 2, 8, 512, &a.tiles_per_split, &a.splits, num_sms() / 2);
    }
    a.partial = nullptr;
    pl.ok = true;
    return pl;
}

struct WgPlan { WgArgs a; bool ok; };

WgPlan plan_wgrad(const ConvGeom& g, int fmt) {
    WgPlan pl{}; pl.ok = false;
    if ((g.stride != 1 && g.stride != 2) || g.Cin % 32 != 0 || g.Cout % 32 != 0 || g.pad_t != g.pad_l || g.k < 1 || g.k > 7) return pl;
    if (g.Cout > MAX_N && g.Cout % MAX_N != 0) return pl;
    WgArgs& a = pl.a;
    a.block_n = g.Cout > MAX_N ? MAX_N : g.Cout;
    a.n_tiles = g.Cout / a.block_n;
    a.mtu = a.block_n >= 128 ? 2 : 1;
    if (const char* ov = getenv("SSDB_WG_MTU")) { int v = atoi(ov); if (v == 1 || v == 2) a.mtu = v; }
    int blocks = 4 * a.mtu + a.block_n / 32;
    int p_max = stage_bytes_for(a.mtu) / (blocks * 128);
    const int p_mult = fmt == ACT_S32 ? 16 : 8;             // pixels per MMA (K of one instruction)
    p_max = p_max / p_mult * p_mult; if (p_max > 64) p_max = 64;
    if (g.stride == 2 && p_max > 32) p_max = 32;            // strided boxes: keep every box dimension <= 256 / stride
    PixGeom pg = pick_pix(g.B, g.Ho, g.Wo, p_max, p_mult);
    if (const char* ov = getenv("SSDB_WG_BOX")) {          // bring-up override: "PW,PH,PN"
        int a1 = 0, a2 = 0, a3 = 0;
        if (sscanf(ov, "%d,%d,%d", &a1, &a2, &a3) == 3 && a1 * a2 * a3 % p_mult == 0 && a1 * a2 * a3 <= p_max) {
            pg.PW = a1; pg.PH = a2; pg.PN = a3; pg.P = a1 * a2 * a3; pg.eff = 1.0;
        }
    }
    if (pg.P == 0 || pg.eff < 0.4) return pl;
    if (pg.PW * g.stride > 256 || pg.PH * g.stride > 256) return pl;
    a.PW = pg.PW; a.PH = pg.PH; a.PN = pg.PN; a.P = pg.P;
    a.ptx = (g.Wo + pg.PW - 1) / pg.PW; a.pty = (g.Ho + pg.PH - 1) / pg.PH; a.ptn = (g.B + pg.PN - 1) / pg.PN;
    a.cblocks = g.Cin / 32; a.taps = g.k * g.k; a.kdim = g.k;
    a.load_blocks = a.cblocks < 4 ? a.cblocks : 4;
    if (a.cblocks % a.load_blocks) return pl;
    a.bias_slot = a.taps * a.cblocks;                        // the all-ones slot comes after the real ones
    a.slots = a.bias_slot + 1; a.m_tiles = (a.slots + 4 * a.mtu - 1) / (4 * a.mtu);
    a.Cin = g.Cin; a.Cout = g.Cout;
    a.off0 = -g.pad_t; a.offstep = g.dil; a.sstride = g.stride;
    a.psize = (long long)a.taps * g.Cin * g.Cout + g.Cout;
    long long pix_tiles = (long long)a.ptx * a.pty * a.ptn;
    long long base_units = (long long)a.m_tiles * a.n_tiles;
    if (fmt == ACT_S32) pick_splits(pix_tiles, base_units, 8, 256, &a.tiles_per_split, &a.splits);
    else {
        long long want = (2LL * num_sms() + base_units - 1) / base_units;     // ~2 units per SM
        long long max_splits = (pix_tiles + 7) / 8;                            // at least 8 stages per unit
        if (want > max_splits) want = max_splits;
        if (want < 1) want = 1;
        if (want > 256) want = 256;
        a.tiles_per_split = (int)((pix_tiles + want - 1) / want);
        a.splits = (int)((pix_tiles + a.tiles_per_split - 1) / a.tiles_per_split);
    }
    a.partial = nullptr; a.ones = nullptr;
    pl.ok = true;
    return pl;
}

}  // namespace

bool conv_tc_supported_wgrad(const ConvGeom& g, int fmt) { return plan_wgrad_r2(g, fmt).ok || plan_wgrad_rw(g, fmt).ok || plan_wgrad(g, fmt).ok; }

size_t conv_tc_wgrad_ws(const ConvGeom& g, int fmt) {
    size_t need = 0;
    // the wave-fitted kernels (r2, split plain) are sized for the most splits any batch <= g.B can pick
    WgR2Plan r2 = plan_wgrad_r2(g, fmt);
    if (r2.ok) need = (size_t)max_splits_bound((long long)r2.a.cblocks * r2.a.mblocks, 512) * (size_t)r2.a.psize;
    WgRwPlan rw = plan_wgrad_rw(g, fmt);
    if (rw.ok) { size_t n1 = (size_t)rw.a.splits * (size_t)rw.a.psize; if (n1 > need) need = n1; }
    WgPlan pl = plan_wgrad(g, fmt);
    if (pl.ok) {
        size_t n2 = (size_t)pl.a.splits * (size_t)pl.a.psize;
        if (fmt == ACT_S32) {
            // with and without the bias slot (m_tiles differs by at most one)
            const long long ut = (long long)pl.a.m_tiles * pl.a.n_tiles, ut2 = (long long)(pl.a.m_tiles > 1 ? pl.a.m_tiles - 1 : 1) * pl.a.n_tiles;
            const long long b = max_splits_bound(ut, 256), b2 = max_splits_bound(ut2, 256);
            n2 = (size_t)(b > b2 ? b : b2) * (size_t)pl.a.psize;
        }
        if (n2 > need) need = n2;
    }
    return need;
}

int conv_tc_wgrad(const ConvGeom& g, const float* x, const float* dz, int fmt, float* dw, float* db, float* partial, cudaStream_t st) {
    const bool split = fmt == ACT_S32;
    WgR2Plan r2 = plan_wgrad_r2(g, fmt);
    if (r2.ok) {
        WgR2Args a = r2.a;
        a.partial = partial;
        a.want_bias = db ? 1 : 0;
        CUtensorMap mx, mz;
        int rc = encode_act_map5s(&mx, x, g.B, g.H, g.W, g.Cin, 8 + 2 * a.dil, a.PH + 2 * a.dil, 1, 2, 1, a.c2 ? 16 : 32); if (rc) return rc;
        rc = encode_act_map5s(&mz, dz, g.B, g.Ho, g.Wo, g.Cout, 8, a.PH, 1, 8, 1); if (rc) return rc;
        static PerDevice<bool> attr_r2_pd;
        bool& attr_r2 = attr_r2_pd.get();
        if (!attr_r2) {
            SSDB_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_r2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            SSDB_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_r2c2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            attr_r2 = true;
        }
        if (a.c2) {
            long long units = (long long)a.cblocks * (a.mblocks / 2) * a.splits;
            const long long pairs = num_sms() / 2;
            int grid = 2 * (int)(units < pairs ? units : pairs);
            conv_tc_wgrad_r2c2_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mx, mz, a);
        } else {
            long long units = (long long)a.cblocks * a.mblocks * a.splits;
            int grid = (int)(units < num_sms() ? units : num_sms());
            conv_tc_wgrad_r2_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mx, mz, a);
        }
        SSDB_LAUNCH_CHECK();
        long long nw = 9LL * g.Cin * g.Cout;
        reduce_partials_kernel<<<(unsigned)((a.psize / 4 + 255) / 256), 256, 0, st>>>(partial, a.psize, nw, a.splits, dw, db);
        SSDB_LAUNCH_CHECK();
        return SSDB_OK;
    }
    WgRwPlan rw = plan_wgrad_rw(g, fmt);
    if (rw.ok) {
        WgRwArgs a = rw.a;
        a.partial = partial;
        a.want_bias = db ? 1 : 0;
        CUtensorMap mx, mz;
        int rc;
        if (split) {
            rc = encode_act_map5s(&mx, x, g.B, g.H, g.W, g.Cin, 11, a.PH + 2, 1, 2 * a.cpu, 1); if (rc) return rc;
            rc = encode_act_map5s(&mz, dz, g.B, g.Ho, g.Wo, g.Cout, 8, a.PH, 1, 2 * (a.block_n / 32), 1); if (rc) return rc;
        } else {
            rc = encode_rw_map(&mx, x, g.B, g.H, g.W, g.Cin, 11, a.PH + 2, 1, a.cpu); if (rc) return rc;
            rc = encode_rw_map(&mz, dz, g.B, g.Ho, g.Wo, g.Cout, 8, a.PH, 1, a.block_n / 32); if (rc) return rc;
        }
        static PerDevice<bool> attr_rw_pd;
        bool& attr_rw = attr_rw_pd.get();
        if (!attr_rw) {
            SSDB_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_rw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            SSDB_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_rw_s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            attr_rw = true;
        }
        long long units = (long long)a.cgroups * a.n_tiles * a.splits;
        int grid = (int)(units < num_sms() ? units : num_sms());
        if (split) conv_tc_wgrad_rw_s_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mx, mz, a);
        else conv_tc_wgrad_rw_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mx, mz, a);
        SSDB_LAUNCH_CHECK();
        long long nw = 9LL * g.Cin * g.Cout;
        reduce_partials_kernel<<<(unsigned)((a.psize / 4 + 255) / 256), 256, 0, st>>>(partial, a.psize, nw, a.splits, dw, db);
        SSDB_LAUNCH_CHECK();
        return SSDB_OK;
    }
    WgPlan pl = plan_wgrad(g, fmt);
    SSDB_REQUIRE(pl.ok, "shape not supported by the tcgen05 wgrad kernel");
    WgArgs a = pl.a;
    a.partial = partial;
    a.ones = split ? ones_buffer_split() : ones_buffer();
    { const char* dbg = getenv("SSDB_WG_DEBUG"); a.debug = dbg ? atoi(dbg) : 0; }
    if (split) a.debug = split_terms_env() < 3 ? split_terms_env() : 0;
    SSDB_REQUIRE(a.ones != nullptr, "could not allocate the ones buffer");
    if (!db) { a.bias_slot = -1; a.slots = a.taps * a.cblocks; a.m_tiles = (a.slots + 4 * a.mtu - 1) / (4 * a.mtu); }
    CUtensorMap mx, mz;
    int rc;
    if (split) {
        rc = encode_act_map5s(&mx, x, g.B, g.H, g.W, g.Cin, a.PW * g.stride, a.PH * g.stride, a.PN, 2 * a.load_blocks, g.stride); if (rc) return rc;
        rc = encode_act_map5s(&mz, dz, g.B, g.Ho, g.Wo, g.Cout, a.PW, a.PH, a.PN, 2 * (a.block_n / 32), 1); if (rc) return rc;
    } else {
        rc = encode_act_map5(&mx, x, g.B, g.H, g.W, g.Cin, a.PW, a.PH, a.PN, a.load_blocks, g.stride); if (rc) return rc;
        rc = encode_act_map5(&mz, dz, g.B, g.Ho, g.Wo, g.Cout, a.PW, a.PH, a.PN, a.block_n / 32, 1); if (rc) return rc;
    }
    static PerDevice<bool> attr_pd;
    bool& attr = attr_pd.get();
    if (!attr) {
        SSDB_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        SSDB_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr = true;
    }
    long long units = (long long)a.m_tiles * a.n_tiles * a.splits;
    int grid = (int)(units < num_sms() ? units : num_sms());
    if (split) conv_tc_wgrad_s_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mx, mz, a);
    else conv_tc_wgrad_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mx, mz, a);
    SSDB_LAUNCH_CHECK();
    long long nw = (long long)a.taps * g.Cin * g.Cout;
    reduce_partials_kernel<<<(unsigned)((a.psize / 4 + 255) / 256), 256, 0, st>>>(partial, a.psize, nw, a.splits, dw, db);
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

}  // namespace ssdb
