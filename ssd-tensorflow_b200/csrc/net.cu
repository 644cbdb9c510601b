// The SSD-VGG engine: layer plan, flat parameter / gradient / momentum buffers,
// forward, multibox loss, backward and Momentum update, behind the C ABI of
// include/ssd_b200.h.  Replaces SSDVGG.build_from_vgg / build_optimizer and the
// tf.Session that runs them (reference ssdvgg.py:87-649, train.py:262-266,
// infer.py:225-227).
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace ssdb {

thread_local char g_err[1024] = "";
long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

struct MapSpec { int size; double scale; std::vector<double> ratios; };
struct Preset { std::string name; int image; std::vector<MapSpec> maps; double extra_scale; int num_anchors; };

// ssdutils.SSD_PRESETS (ssdutils.py:36-62)
const Preset* find_preset(const std::string& n) {
    static const Preset p300{"vgg300", 300,
        {{38, 0.1, {2, 0.5}}, {19, 0.2, {2, 3, 0.5, 1. / 3.}}, {10, 0.375, {2, 3, 0.5, 1. / 3.}},
         {5, 0.55, {2, 3, 0.5, 1. / 3.}}, {3, 0.725, {2, 0.5}}, {1, 0.9, {2, 0.5}}}, 1.075, 8732};
    static const Preset p512{"vgg512", 512,
        {{64, 0.07, {2, 0.5}}, {32, 0.15, {2, 3, 0.5, 1. / 3.}}, {16, 0.3, {2, 3, 0.5, 1. / 3.}},
         {8, 0.45, {2, 3, 0.5, 1. / 3.}}, {4, 0.6, {2, 3, 0.5, 1. / 3.}}, {2, 0.75, {2, 0.5}}, {1, 0.9, {2, 0.5}}}, 1.05, 24564};
    if (n == "vgg300") return &p300;
    if (n == "vgg512") return &p512;
    return nullptr;
}

struct Buf { int H, W, C; size_t off; bool relu_out; };

struct Master {            // one tensor of the flat buffer
    std::string name; int rank; int shape[4]; size_t off; size_t count; bool decay;
};
struct RefTensor {         // a tensor under the reference's variable name
    std::string name; int rank; int shape[4];
    int master;            // index into masters
    int col0, cols;        // column window inside the master's last dimension (classifier views)
};

enum OpType { OP_CONV = 0, OP_POOL = 1, OP_L2NORM = 2 };
struct Op {
    OpType type; std::string name;
    int in, out;                       // buffer ids; in == -1: the image batch; out == -1: head output tensor
    int k = 1, stride = 1, dil = 1, pad = 0;
    int cin = 0, cout = 0;             // cout = stored channel count (heads: padded)
    bool relu = false, head = false;
    int anchor_base = 0, nbox = 0;
    int w = -1, b = -1;                // master indices
    size_t wt_off = 0; int cout_pad = 0; bool has_wt = false;
    int pool_after = -1;               // conv: index of the 2x2/s2 pool that is the ONLY reader of its output (fusable), else -1
};

int same_pad_before(int n, int keff, int stride) {
    int out = (n + stride - 1) / stride;
    int total = (out - 1) * stride + keff - n;
    if (total < 0) total = 0;
    return total / 2;
}

}  // namespace
}  // namespace ssdb

using namespace ssdb;

struct ssdb_net {
    const Preset* preset = nullptr;
    int C = 20, V = 25, A = 0, S = 300, max_batch = 0;
    std::vector<Buf> bufs;
    std::vector<Op> ops;
    std::vector<Master> masters;
    std::vector<RefTensor> refs;
    std::map<std::string, int> ref_index;
    size_t n_flat = 0, act_floats_per_image = 0, wt_floats = 0;
    // device memory
    float *params = nullptr, *grads = nullptr, *moms = nullptr, *wt = nullptr;
    float *wr = nullptr;               // tf32-rounded copy of the parameters (dgrad B operand)
    // conv1_1 as a 1x1 tensor-core conv over an explicit 3x3x3 patch matrix (Cin = 3 cannot feed the MMA directly)
    float *patches = nullptr, *c1_w32 = nullptr, *c1_wt = nullptr, *c1_dw32 = nullptr;
    unsigned char* pool5_arg = nullptr;   // winning window cell of mod_pool5 (3x3 stride 1), one byte per output element
    std::vector<unsigned char*> pool_code; // per op: code bytes of the 2x2/s2 pools (null for every other op)
    bool round = false;                // tf32 mode: activations / gradients are stored tf32-rounded
    int fmt = ACT_S32;                 // storage format of activations / gradients / packed filters (common.cuh): split bf16
                                       // pairs (default), plain float32 in the tf32 and SIMT modes
    float *acts = nullptr, *gacts = nullptr;
    float *out = nullptr, *out_grad = nullptr, *result = nullptr, *dz_head = nullptr;
    std::vector<size_t> dz_head_off;   // per head (in backward order): offset of its dz buffer inside dz_head
    float* l2n_partial = nullptr;      // workspace of l2norm_bwd (the wgrad kernels' `partial` may be in use on the side stream)
    float *images_stage = nullptr, *labels_stage = nullptr;
    double* gt_stage = nullptr; int* gt_count_stage = nullptr;    // raw ground truth of the host entry points ([max_batch, 128, 5] + counts)
    int* match_stage = nullptr;                                   // [max_batch, A] owner GT per anchor (fused match), on request
    int *det_rows = nullptr, *det_counts = nullptr;               // ssdb_forward_detect_host: [max_batch, A, 8] / [max_batch, 2]
    float *partial = nullptr; size_t partial_floats = 0;
    float *small_ws = nullptr;         // [0..3] losses, [4..5] conf/loc, [6] l2 sum, then per-image + partials
    unsigned int* counter = nullptr;
    void* loss_ws = nullptr;           // workspace of the multibox loss kernels (multibox_loss_ws_bytes)
    void* det_ws = nullptr;            // workspace of the decode + NMS launched on n->result (ssdb_forward_detect_host; decode_nms_scratch_bytes at max_batch)
    unsigned char* decay_mask = nullptr;
    double* anchors = nullptr;
    float* host_small = nullptr;       // pinned
    bool wt_dirty = true, have_forward = false;
    bool inference = false;            // SSDB_FLAG_INFERENCE: forward / detection only, no training state
    bool fuse_pool = true;             // conv1_2 / conv2_2 write their max pool instead of their output (SSDB_FUSE_POOL=0: off)
    std::vector<char> fused_now;       // per op: this forward pass fused it away (its activation was not materialised)
    // CUDA graphs of the frozen forward + detection (ssdb_forward_detect_host on an inference handle), one per
    // (batch, threshold, cap, IoU); dropped when a parameter changes
    struct DetGraph { int B; float thr; int cap; double iou; bool warmed; cudaGraphExec_t exec; };
    std::vector<DetGraph> det_graphs;
    int conv_mode = SSDB_CONV_AUTO;
    int swap_rb = 1; float mean[3] = {103.939f, 116.779f, 123.68f};
    cudaStream_t own_stream = nullptr, copy_stream = nullptr, side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_side = nullptr, ev_l2 = nullptr;
    bool l2_pending = false;           // the L2 term of this step is being summed on the side stream (ev_l2)
    cudaEvent_t ev_labels = nullptr, ev_result = nullptr, ev_images = nullptr;
    cudaEvent_t ev_chunk[8] = {};      // image chunks of the host entry points
    // gradient buckets for an all-reduce that overlaps the rest of the backward (ssdb_grad_buckets): contiguous ranges of
    // the flat gradient buffer, final in the order the backward produces them (heads and deep layers first)
    static constexpr int MAX_BUCKETS = 4;
    int n_buckets = 0;
    long long bucket_begin[MAX_BUCKETS] = {}, bucket_end[MAX_BUCKETS] = {};
    int bucket_last_op[MAX_BUCKETS] = {};          // the backward has finished a bucket once this op's gradients are written
    cudaEvent_t ev_bucket[MAX_BUCKETS] = {};
    int last_B = 0;
    // per-op device timing (ssdb_profile_step)
    bool prof = false;
    struct ProfEntry { std::string label; cudaEvent_t a, b; long long launches; double flops; };
    std::vector<ProfEntry> prof_entries;

    float* act(int id, int /*B*/) { return acts + bufs[id].off * (size_t)max_batch; }
    float* gact(int id, int /*B*/) { return gacts + bufs[id].off * (size_t)max_batch; }
};

static void drop_det_graphs(ssdb_net* n) {
    for (auto& g : n->det_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    n->det_graphs.clear();
}

namespace ssdb {
namespace {

__global__ void finalize_losses_kernel(const float* conf_loc, const float* l2sum, float wd, float* out4) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float l2 = wd * l2sum[0];
        out4[0] = conf_loc[0] + conf_loc[1] + l2;   // total
        out4[1] = conf_loc[1];                       // localization
        out4[2] = conf_loc[0];                       // confidence
        out4[3] = l2;
    }
}

int add_buf(ssdb_net* n, int H, int W, int C, bool relu_out) {
    Buf b{H, W, C, n->act_floats_per_image, relu_out};
    n->act_floats_per_image += (size_t)H * W * C;
    n->act_floats_per_image = (n->act_floats_per_image + 63) / 64 * 64;
    n->bufs.push_back(b);
    return (int)n->bufs.size() - 1;
}

int add_master(ssdb_net* n, const std::string& name, int rank, const int* shape, bool decay) {
    Master m; m.name = name; m.rank = rank; m.count = 1;
    for (int i = 0; i < 4; ++i) { m.shape[i] = i < rank ? shape[i] : 1; m.count *= m.shape[i]; }
    m.off = n->n_flat; m.decay = decay;
    n->n_flat += (m.count + OPT_BLOCK - 1) / OPT_BLOCK * OPT_BLOCK;
    n->masters.push_back(m);
    return (int)n->masters.size() - 1;
}

void add_ref(ssdb_net* n, const std::string& name, int rank, const int* shape, int master, int col0, int cols) {
    RefTensor r; r.name = name; r.rank = rank;
    for (int i = 0; i < 4; ++i) r.shape[i] = i < rank ? shape[i] : 1;
    r.master = master; r.col0 = col0; r.cols = cols;
    n->ref_index[name] = (int)n->refs.size();
    n->refs.push_back(r);
}

// conv_map (ssdvgg.py:42-52) / VGG conv: returns the output buffer id
int add_conv(ssdb_net* n, const std::string& name, int in, int k, int cout, int stride, int dil, bool same, int out_override = 0) {
    Op op; op.type = OP_CONV; op.name = name; op.in = in; op.k = k; op.stride = stride; op.dil = dil; op.relu = true;
    int H = in < 0 ? n->S : n->bufs[in].H, W = in < 0 ? n->S : n->bufs[in].W;
    op.cin = in < 0 ? 3 : n->bufs[in].C; op.cout = cout;
    int keff = (k - 1) * dil + 1, Ho, Wo;
    if (same) { op.pad = same_pad_before(H, keff, stride); Ho = (H + stride - 1) / stride; Wo = (W + stride - 1) / stride; }
    else { op.pad = 0; Ho = (H - keff) / stride + 1; Wo = (W - keff) / stride + 1; }
    if (out_override) { Ho = Wo = out_override; }
    op.out = add_buf(n, Ho, Wo, cout, true);
    int fs[4] = {k, k, op.cin, cout};
    op.w = add_master(n, name + "/filter", 4, fs, true);
    int bs[1] = {cout};
    op.b = add_master(n, name + "/biases", 1, bs, false);
    add_ref(n, name + "/filter", 4, fs, op.w, 0, cout);
    add_ref(n, name + "/biases", 1, bs, op.b, 0, cout);
    n->ops.push_back(op);
    return op.out;
}

int add_pool(ssdb_net* n, const std::string& name, int in, int k, int stride) {
    Op op; op.type = OP_POOL; op.name = name; op.in = in; op.k = k; op.stride = stride;
    const Buf& b = n->bufs[in];
    op.pad = same_pad_before(b.H, k, stride);
    op.cin = op.cout = b.C;
    op.out = add_buf(n, (b.H + stride - 1) / stride, (b.W + stride - 1) / stride, b.C, false);
    n->ops.push_back(op);
    return op.out;
}

void build_plan(ssdb_net* n) {
    const Preset& P = *n->preset;
    int x = -1;
    x = add_conv(n, "conv1_1", x, 3, 64, 1, 1, true);  x = add_conv(n, "conv1_2", x, 3, 64, 1, 1, true);
    x = add_pool(n, "pool1", x, 2, 2);
    x = add_conv(n, "conv2_1", x, 3, 128, 1, 1, true); x = add_conv(n, "conv2_2", x, 3, 128, 1, 1, true);
    x = add_pool(n, "pool2", x, 2, 2);
    x = add_conv(n, "conv3_1", x, 3, 256, 1, 1, true); x = add_conv(n, "conv3_2", x, 3, 256, 1, 1, true);
    x = add_conv(n, "conv3_3", x, 3, 256, 1, 1, true);
    x = add_pool(n, "pool3", x, 2, 2);
    x = add_conv(n, "conv4_1", x, 3, 512, 1, 1, true); x = add_conv(n, "conv4_2", x, 3, 512, 1, 1, true);
    int c43 = add_conv(n, "conv4_3", x, 3, 512, 1, 1, true);
    x = add_pool(n, "pool4", c43, 2, 2);
    x = add_conv(n, "conv5_1", x, 3, 512, 1, 1, true); x = add_conv(n, "conv5_2", x, 3, 512, 1, 1, true);
    x = add_conv(n, "conv5_3", x, 3, 512, 1, 1, true);
    x = add_pool(n, "mod_pool5", x, 3, 1);                                   // ssdvgg.py:234
    x = add_conv(n, "mod_conv6", x, 3, 1024, 1, 6, true);                    // a-trous, rate 6 (ssdvgg.py:260)
    int c7 = add_conv(n, "mod_conv7", x, 1, 1024, 1, 1, true);
    const bool seven = P.maps.size() >= 7;
    x = add_conv(n, "conv8_1", c7, 1, 256, 1, 1, true);   int c82 = add_conv(n, "conv8_2", x, 3, 512, 2, 1, true);
    x = add_conv(n, "conv9_1", c82, 1, 128, 1, 1, true);  int c92 = add_conv(n, "conv9_2", x, 3, 256, 2, 1, true);
    x = add_conv(n, "conv10_1", c92, 1, 128, 1, 1, true);
    int c102 = add_conv(n, "conv10_2", x, 3, 256, seven ? 2 : 1, 1, seven);
    x = add_conv(n, "conv11_1", c102, 1, 128, 1, 1, true);
    int c112 = add_conv(n, "conv11_2", x, 3, 256, 1, 1, false);
    std::vector<int> fmaps;
    // l2_normalization of conv4_3 (ssdvgg.py:80-84,335-337)
    {
        Op op; op.type = OP_L2NORM; op.name = "l2_norm_conv4_3"; op.in = c43;
        const Buf& b = n->bufs[c43];
        op.cin = op.cout = b.C;
        op.out = add_buf(n, b.H, b.W, b.C, false);
        int ss[1] = {b.C};
        op.w = add_master(n, "l2_norm_conv4_3/scale", 1, ss, false);
        add_ref(n, "l2_norm_conv4_3/scale", 1, ss, op.w, 0, b.C);
        n->ops.push_back(op);
        fmaps.push_back(op.out);
    }
    fmaps.push_back(c7); fmaps.push_back(c82); fmaps.push_back(c92); fmaps.push_back(c102); fmaps.push_back(c112);
    if (seven) {
        x = add_conv(n, "conv12_1", c112, 1, 128, 1, 1, true);
        // zero-pad bottom/right by one, then 3x3 VALID (ssdvgg.py:326-331) == pad-after-only conv with a 1x1 output
        int c122 = add_conv(n, "conv12_2", x, 3, 256, 1, 1, false, 1);
        fmaps.push_back(c122);
    }
    // classifiers (ssdvgg.py:55-65,353-366): the box types of one map share a merged filter
    int base = 0;
    for (size_t i = 0; i < P.maps.size(); ++i) {
        int nbox = 2 + (int)P.maps[i].ratios.size();
        const Buf& fb = n->bufs[fmaps[i]];
        Op op; op.type = OP_CONV; op.name = "classifiers/map" + std::to_string(i); op.in = fmaps[i]; op.out = -1;
        op.k = 3; op.stride = 1; op.dil = 1; op.pad = 1; op.cin = fb.C; op.relu = false; op.head = true;
        op.nbox = nbox; op.anchor_base = base;
        op.cout = (nbox * n->V + 31) / 32 * 32;
        int fs[4] = {3, 3, fb.C, op.cout};
        op.w = add_master(n, op.name + "/filter", 4, fs, true);
        int bs[1] = {op.cout};
        op.b = add_master(n, op.name + "/biases", 1, bs, false);
        for (int j = 0; j < nbox; ++j) {
            std::string rn = "classifiers/classifier" + std::to_string(i) + "_" + std::to_string(j);
            int rfs[4] = {3, 3, fb.C, n->V};
            add_ref(n, rn + "/filter", 4, rfs, op.w, j * n->V, n->V);
            int rbs[1] = {n->V};
            add_ref(n, rn + "/biases", 1, rbs, op.b, j * n->V, n->V);
        }
        n->ops.push_back(op);
        base += nbox * fb.H * fb.W;
    }
    n->A = base;
    // conv -> 2x2/s2 pool pairs where the pool is the only reader of the conv output (conv1_2, conv2_2, conv3_3; conv4_3 also
    // feeds the L2 normalisation): candidates for the fused epilogue
    for (size_t i = 0; i < n->ops.size(); ++i) {
        Op& c = n->ops[i];
        if (c.type != OP_CONV || c.head || c.out < 0) continue;
        int readers = 0, pool = -1;
        for (size_t j = 0; j < n->ops.size(); ++j)
            if (n->ops[j].in == c.out) { ++readers; if (n->ops[j].type == OP_POOL && n->ops[j].k == 2 && n->ops[j].stride == 2 && n->ops[j].pad == 0) pool = (int)j; }
        if (readers == 1 && pool == (int)i + 1) c.pool_after = pool;
    }
    // transposed filter copies for the tcgen05 fprop kernel
    for (Op& op : n->ops) {
        if (op.type != OP_CONV || op.cin % 32 != 0) continue;
        int bn = op.cout > 256 ? 256 : (op.cout + 15) / 16 * 16;
        op.cout_pad = (op.cout + bn - 1) / bn * bn;
        op.wt_off = n->wt_floats; op.has_wt = true;
        n->wt_floats += (size_t)op.k * op.k * op.cout_pad * op.cin;
    }
}

ConvGeom geom_of(const ssdb_net* n, const Op& op, int B) {
    ConvGeom g;
    g.B = B;
    g.H = op.in < 0 ? n->S : n->bufs[op.in].H; g.W = op.in < 0 ? n->S : n->bufs[op.in].W; g.Cin = op.cin;
    if (op.out >= 0) { g.Ho = n->bufs[op.out].H; g.Wo = n->bufs[op.out].W; } else { g.Ho = g.H; g.Wo = g.W; }
    g.Cout = op.cout; g.k = op.k; g.stride = op.stride; g.dil = op.dil; g.pad_t = op.pad; g.pad_l = op.pad;
    return g;
}

bool use_tc(const ssdb_net* n, bool supported) {
    if (n->conv_mode == SSDB_CONV_SIMT) return false;
    return supported;
}

int repack_filters(ssdb_net* n, cudaStream_t st) {
    for (const Op& op : n->ops) {
        if (!op.has_wt) continue;
        int rc = pack_filter_t(n->params + n->masters[op.w].off, op.k * op.k, op.cin, op.cout, op.cout_pad, n->fmt, n->wt + op.wt_off, st);
        if (rc) return rc;
    }
    // dgrad B operand: the HWIO filters themselves, tf32-rounded or split along Cout (every tensor starts on a 1024-element boundary)
    if (n->round) { int rc = round_tf32_copy(n->params, n->wr, (long long)n->n_flat, st); if (rc) return rc; }
    else if (n->fmt == ACT_S32 && n->wr) { int rc = split_copy(n->params, n->wr, (long long)n->n_flat, st); if (rc) return rc; }
    if (n->patches) {
        const Op& c1 = n->ops[0];
        int rc = conv1_pad_filter(n->params + n->masters[c1.w].off, c1.cout, n->c1_w32, st); if (rc) return rc;
        rc = pack_filter_t(n->c1_w32, 1, 32, c1.cout, c1.cout, n->fmt, n->c1_wt, st); if (rc) return rc;
    }
    n->wt_dirty = false;
    return SSDB_OK;
}

struct ProfScope {
    ssdb_net* n; cudaStream_t st; int idx = -1; long long l0 = 0;
    ProfScope(ssdb_net* n_, cudaStream_t st_, const std::string& label, double flops = 0.0) : n(n_), st(st_) {
        if (!n->prof) return;
        ssdb_net::ProfEntry e; e.label = label; e.launches = 0; e.flops = flops;
        cudaEventCreate(&e.a); cudaEventCreate(&e.b);
        cudaEventRecord(e.a, st);
        n->prof_entries.push_back(e); idx = (int)n->prof_entries.size() - 1; l0 = g_launches;
    }
    ~ProfScope() {
        if (idx < 0) return;
        cudaEventRecord(n->prof_entries[idx].b, st);
        n->prof_entries[idx].launches = g_launches - l0;
    }
};

double conv_flops(const ConvGeom& g) { return 2.0 * g.B * g.Ho * g.Wo * (double)g.k * g.k * g.Cin * g.Cout; }

// conv1_1 (patch matrix + 1x1 tensor-core conv) of images [b0, b0 + Bc) only: every output pixel depends on its own image,
// so the host entry points run it chunk by chunk while the rest of the batch is still on its way over PCIe
int run_first_conv_chunk(ssdb_net* n, const float* images, int b0, int Bc, cudaStream_t st) {
    const Op& op = n->ops[0];
    SSDB_REQUIRE(op.type == OP_CONV && op.in < 0 && n->patches, "internal: chunked conv1_1 needs the tensor-core first layer");
    ConvGeom g = geom_of(n, op, Bc);
    ConvEpilogue ep;
    ep.bias = n->params + n->masters[op.b].off; ep.relu = op.relu ? 1 : 0; ep.round_tf32 = n->round ? 1 : 0;
    const size_t px = (size_t)b0 * n->S * n->S;
    float* patches = n->patches + px * 32;
    int rc = conv1_im2col(images + px * 3, Bc, n->S, n->swap_rb, n->mean, n->fmt, patches, st); if (rc) return rc;
    ConvGeom g1 = g; g1.Cin = 32; g1.k = 1; g1.pad_t = g1.pad_l = 0; g1.dil = 1;
    return conv_tc_fprop(g1, patches, n->c1_wt, op.cout, n->fmt, ep, n->act(op.out, n->max_batch) + px * op.cout, st);
}

// Two streams per step.  The classifier convolutions (forward) and EVERY weight-gradient kernel (backward) go to a side
// stream, ordered by events: a head starts when its feature map is written and runs beside the rest of the trunk; the
// wgrad of a layer starts when its dz is final and runs beside the dgrad chain.  The kernels are persistent and take a whole
// SM each, so two big layers simply queue; the gain is in the tails -- the last, partly filled wave of one kernel is topped
// up by the other stream's CTAs -- and in the small layers (extras, heads of the small maps: 12-85 us each on a handful of
// SMs), which now overlap each other.  Off while profiling per op (ssdb_profile_step) and with SSDB_DUAL_STREAM=0.
bool dual_stream(const ssdb_net* n) {
    static int on = -1;
    if (on < 0) { const char* ov = getenv("SSDB_DUAL_STREAM"); on = ov ? (atoi(ov) ? 1 : 0) : 1; }
    return on && !n->prof && n->side_stream != nullptr;
}

int run_one_forward_op(ssdb_net* n, const Op& op, const float* images, int B, cudaStream_t st, const Op* fused_pool = nullptr,
                       cudaEvent_t filters_ready = nullptr) {
    int rc = SSDB_OK;
    ProfScope ps(n, st, std::string("fwd:") + op.name, op.type == OP_CONV ? conv_flops(geom_of(n, op, B)) : 0.0);
    if (op.type == OP_CONV) {
        ConvGeom g = geom_of(n, op, B);
        ConvEpilogue ep;
        ep.bias = n->params + n->masters[op.b].off; ep.relu = op.relu ? 1 : 0;
        ep.round_tf32 = (n->round && !op.head) ? 1 : 0;
        const float* x = op.in < 0 ? images : n->act(op.in, B);
        float* y = op.out >= 0 ? n->act(op.out, B) : n->out;
        if (op.head) { ep.scatter = 1; ep.V = n->V; ep.n_valid = op.nbox * n->V; ep.anchor_base = op.anchor_base; ep.A = n->A; }
        if (op.in < 0) { ep.preprocess = 1; ep.swap_rb = n->swap_rb; ep.mean[0] = n->mean[0]; ep.mean[1] = n->mean[1]; ep.mean[2] = n->mean[2]; }
        if (op.in < 0 && n->patches) {
            rc = conv1_im2col(images, B, n->S, n->swap_rb, n->mean, n->fmt, n->patches, st); if (rc) return rc;
            if (filters_ready) SSDB_CUDA(cudaStreamWaitEvent(st, filters_ready, 0));       // the patch matrix needs no filters: it ran beside their re-split
            ConvGeom g1 = g; g1.Cin = 32; g1.k = 1; g1.pad_t = g1.pad_l = 0; g1.dil = 1;
            ConvEpilogue e1 = ep; e1.preprocess = 0;
            rc = conv_tc_fprop(g1, n->patches, n->c1_wt, op.cout, n->fmt, e1, y, st);
        } else if (op.has_wt && use_tc(n, conv_tc_supported_fprop(g))) {
            if (fused_pool) { ep.pool_dst = n->act(fused_pool->out, B); ep.pool_code = n->pool_code[fused_pool - n->ops.data()]; }
            rc = conv_tc_fprop(g, x, n->wt + op.wt_off, op.cout_pad, n->fmt, ep, y, st);
        } else
            rc = conv_simt_fprop(g, x, n->params + n->masters[op.w].off, n->fmt, ep, y, st);
    } else if (op.type == OP_POOL) {
        const Buf& bi = n->bufs[op.in]; const Buf& bo = n->bufs[op.out];
        if (op.stride == 1 && n->pool5_arg)
            rc = maxpool_fwd_arg(n->act(op.in, B), n->fmt, B, bi.H, bi.W, bi.C, op.k, op.stride, op.pad, op.pad, bo.H, bo.W, n->act(op.out, B), n->pool5_arg, st);
        else if (n->pool_code[&op - n->ops.data()])
            rc = maxpool2x2_fwd_code(n->act(op.in, B), n->fmt, B, bi.H, bi.W, bi.C, bo.H, bo.W, n->act(op.out, B), n->pool_code[&op - n->ops.data()], st);
        else
            rc = maxpool_fwd(n->act(op.in, B), n->fmt, B, bi.H, bi.W, bi.C, op.k, op.stride, op.pad, op.pad, bo.H, bo.W, n->act(op.out, B), st);
    } else {
        const Buf& bi = n->bufs[op.in];
        rc = l2norm_fwd(n->act(op.in, B), n->params + n->masters[op.w].off, n->fmt, (long long)B * bi.H * bi.W, bi.C, n->round ? 1 : 0, n->act(op.out, B), st);
    }
    return rc;
}

// want_l2: also start the L2 term of the loss (sum of squares of 26 M parameters, 0.1 ms; it depends on nothing the step
// computes) on the side stream; loss_and_finalize waits for it
int run_forward(ssdb_net* n, const float* images, int B, cudaStream_t st, bool skip_first = false, bool want_l2 = false) {
    SSDB_REQUIRE(B >= 1 && B <= n->max_batch, "batch size out of range");
    const bool two = dual_stream(n);
    cudaStream_t s2 = n->side_stream;
    cudaEvent_t filters_ready = nullptr;
    if (n->wt_dirty) {
        // the per-step re-split of the filters (58 small launches, 0.25 ms) runs on the side stream beside the patch matrix of
        // conv1_1, which needs no filters
        const bool beside = two && !skip_first && n->patches != nullptr;
        if (beside) { SSDB_CUDA(cudaEventRecord(n->ev_fork, st)); SSDB_CUDA(cudaStreamWaitEvent(s2, n->ev_fork, 0)); }
        { ProfScope ps(n, beside ? s2 : st, "repack"); int rc = repack_filters(n, beside ? s2 : st); if (rc) return rc; }
        if (beside) { SSDB_CUDA(cudaEventRecord(n->ev_side, s2)); filters_ready = n->ev_side; }
    }
    if (want_l2 && two) {
        SSDB_CUDA(cudaEventRecord(n->ev_fork, st));                  // the parameters are final in stream order here
        SSDB_CUDA(cudaStreamWaitEvent(s2, n->ev_fork, 0));
        int rc = l2_sum(n->params, (long long)n->n_flat, n->decay_mask, n->small_ws + 8, n->small_ws + 6, s2); if (rc) return rc;
        SSDB_CUDA(cudaEventRecord(n->ev_l2, s2));
        n->l2_pending = true;
    }
    bool forked = false;
    n->fused_now.assign(n->ops.size(), 0);
    for (const Op& op : n->ops) {
        const size_t oi = &op - n->ops.data();
        if (skip_first && oi == 0) continue;
        if (op.head && two) continue;                      // launched right after the op that writes its feature map (below)
        if (n->fused_now[oi]) continue;                    // a pool that the previous conv already computed
        // conv + the 2x2 pool that alone reads it: the conv's epilogue writes the pooled map and the pool's code bytes
        const Op* fp = nullptr;
        if (op.type == OP_CONV && op.pool_after >= 0 && n->fuse_pool && n->pool_code[op.pool_after] && op.has_wt && n->fmt == ACT_S32 &&
            use_tc(n, true) && conv_tc_fprop_can_pool(geom_of(n, op, B), n->fmt)) {
            fp = &n->ops[op.pool_after];
            n->fused_now[op.pool_after] = 1; n->fused_now[oi] = 2;
        }
        int rc = run_one_forward_op(n, op, images, B, st, fp, oi == 0 ? filters_ready : nullptr); if (rc) return rc;
        if (oi == 0 && filters_ready && !(op.in < 0 && n->patches)) SSDB_CUDA(cudaStreamWaitEvent(st, filters_ready, 0));
        if (!two) continue;
        for (const Op& h : n->ops) {
            if (!h.head || h.in != op.out) continue;
            SSDB_CUDA(cudaEventRecord(n->ev_fork, st));
            SSDB_CUDA(cudaStreamWaitEvent(s2, n->ev_fork, 0));
            rc = run_one_forward_op(n, h, images, B, s2); if (rc) return rc;
            forked = true;
        }
    }
    if (forked) {
        SSDB_CUDA(cudaEventRecord(n->ev_join, s2));
        SSDB_CUDA(cudaStreamWaitEvent(st, n->ev_join, 0));
    }
    n->have_forward = true; n->last_B = B;
    return SSDB_OK;
}

int run_backward(ssdb_net* n, int B, cudaStream_t st) {
    SSDB_REQUIRE(n->have_forward && n->last_B == B, "backward without a matching forward");
    const bool two = dual_stream(n);
    cudaStream_t sw = two ? n->side_stream : st;            // the stream of the weight-gradient kernels
    std::vector<char> written(n->bufs.size(), 0);
    int head_k = 0;
    for (int oi = (int)n->ops.size() - 1; oi >= 0; --oi) {
        const Op& op = n->ops[oi];
        int rc = SSDB_OK;
        if (op.type == OP_CONV) {
            ConvGeom g = geom_of(n, op, B);
            const float* dz;
            if (op.head) {
                // every head has its own dz buffer: its wgrad (side stream) and dgrad (main stream) read it while the next
                // head's gather already runs
                const Buf& fb = n->bufs[op.in];
                float* dzh = n->dz_head + n->dz_head_off[head_k++];
                rc = head_grad_gather(n->out_grad, B, n->A, n->V, op.anchor_base, fb.H * fb.W, op.nbox, op.cout, n->fmt, n->round ? 1 : 0, dzh, st);
                if (rc) return rc;
                dz = dzh;
            } else {
                SSDB_REQUIRE(written[op.out], "internal: gradient of a conv output was never produced");
                dz = n->gact(op.out, B);
            }
            const float* x = op.in < 0 ? nullptr : n->act(op.in, B);
            float* dw = n->grads + n->masters[op.w].off;
            float* db = n->grads + n->masters[op.b].off;
            long long pixels = (long long)B * g.Ho * g.Wo;
            if (two) {                                       // dz of this layer is final here: the wgrad may start
                SSDB_CUDA(cudaEventRecord(n->ev_fork, st));
                SSDB_CUDA(cudaStreamWaitEvent(sw, n->ev_fork, 0));
            }
            if (op.in < 0 && n->patches) {
                ProfScope ps(n, sw, std::string("bwd_w:") + op.name, conv_flops(g));
                ConvGeom g1 = g; g1.Cin = 32; g1.k = 1; g1.pad_t = g1.pad_l = 0; g1.dil = 1;
                rc = conv_tc_wgrad(g1, n->patches, dz, n->fmt, n->c1_dw32, db, n->partial, sw); if (rc) return rc;
                SSDB_CUDA(cudaMemcpyAsync(dw, n->c1_dw32, (size_t)27 * op.cout * sizeof(float), cudaMemcpyDeviceToDevice, sw));
            } else {
                const bool tcw = op.in >= 0 && use_tc(n, conv_tc_supported_wgrad(g, n->fmt));
                if (!tcw) { ProfScope ps(n, sw, std::string("bwd_b:") + op.name); rc = bias_grad(dz, n->fmt, pixels, op.cout, db, n->partial, sw); }
                if (rc) return rc;
                ProfScope* psw = new ProfScope(n, sw, std::string("bwd_w:") + op.name, conv_flops(g));
                ConvEpilogue ep;
                if (op.in < 0) { ep.preprocess = 1; ep.swap_rb = n->swap_rb; ep.mean[0] = n->mean[0]; ep.mean[1] = n->mean[1]; ep.mean[2] = n->mean[2]; }
                if (tcw)
                    rc = conv_tc_wgrad(g, x, dz, n->fmt, dw, db, n->partial, sw);
                else
                    rc = conv_simt_wgrad(g, op.in < 0 ? n->images_stage : x, dz, n->fmt, ep, dw, n->partial, sw);
                delete psw;
                if (rc) return rc;
            }
            if (op.in >= 0) {
                ProfScope psd(n, st, std::string("bwd_d:") + op.name, conv_flops(g));
                const float* mask = n->bufs[op.in].relu_out ? x : nullptr;
                int beta = written[op.in] ? 1 : 0;
                if (use_tc(n, conv_tc_supported_dgrad(g)))
                    rc = conv_tc_dgrad(g, dz, n->wr + n->masters[op.w].off, n->fmt, mask, beta, n->round ? 1 : 0, n->gact(op.in, B), st);
                else
                    rc = conv_simt_dgrad(g, dz, n->params + n->masters[op.w].off, n->fmt, mask, beta, n->round ? 1 : 0, n->gact(op.in, B), st);
                written[op.in] = 1;
            }
        } else if (op.type == OP_POOL) {
            ProfScope ps(n, st, std::string("bwd:") + op.name);
            SSDB_REQUIRE(written[op.out], "internal: gradient of a pool output was never produced");
            const Buf& bi = n->bufs[op.in]; const Buf& bo = n->bufs[op.out];
            if (op.stride == 1 && n->pool5_arg)
                rc = maxpool_bwd_arg(n->act(op.in, B), n->gact(op.out, B), n->pool5_arg, n->fmt, B, bi.H, bi.W, bi.C, op.k, op.stride, op.pad, op.pad, bo.H, bo.W,
                                     written[op.in] ? 1 : 0, bi.relu_out ? 1 : 0, n->round ? 1 : 0, n->gact(op.in, B), st);
            else if (n->pool_code[&op - n->ops.data()] && !written[op.in])
                rc = maxpool2x2_bwd_code(n->gact(op.out, B), n->pool_code[&op - n->ops.data()], n->fmt, B, bi.H, bi.W, bi.C, bo.H, bo.W, bi.relu_out ? 1 : 0,
                                         n->round ? 1 : 0, n->gact(op.in, B), st);
            else
            rc = maxpool_bwd(n->act(op.in, B), n->gact(op.out, B), n->fmt, B, bi.H, bi.W, bi.C, op.k, op.stride, op.pad, op.pad, bo.H, bo.W,
                             written[op.in] ? 1 : 0, bi.relu_out ? 1 : 0, n->round ? 1 : 0, n->gact(op.in, B), st);
            written[op.in] = 1;
        } else {
            // the scale gradient of the L2 normalisation is written on the main stream (its own small workspace); it belongs to
            // the first gradient bucket, whose event is recorded on the wgrad stream AFTER a wait on a later main-stream event
            ProfScope ps(n, st, std::string("bwd:") + op.name);
            SSDB_REQUIRE(written[op.out], "internal: gradient of the L2-norm output was never produced");
            const Buf& bi = n->bufs[op.in];
            rc = l2norm_bwd(n->act(op.in, B), n->params + n->masters[op.w].off, n->gact(op.out, B), n->fmt, (long long)B * bi.H * bi.W, bi.C,
                            written[op.in] ? 1 : 0, n->round ? 1 : 0, n->gact(op.in, B), n->grads + n->masters[op.w].off, n->l2n_partial, st);
            written[op.in] = 1;
        }
        if (rc) return rc;
        for (int b = 0; b < n->n_buckets; ++b)
            if (n->bucket_last_op[b] == oi) SSDB_CUDA(cudaEventRecord(n->ev_bucket[b], sw));
    }
    if (two) {                                               // the update (or the caller's all-reduce) needs every weight gradient
        SSDB_CUDA(cudaEventRecord(n->ev_join, sw));
        SSDB_CUDA(cudaStreamWaitEvent(st, n->ev_join, 0));
    }
    return SSDB_OK;
}

// images for the conv1_1 weight gradient: the backward needs the raw batch again
int remember_images(ssdb_net* n, const float* images, int B, cudaStream_t st) {
    if (images == n->images_stage) return SSDB_OK;
    SSDB_CUDA(cudaMemcpyAsync(n->images_stage, images, (size_t)B * n->S * n->S * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return SSDB_OK;
}

void anchors_host(const Preset& P, std::vector<double>& out) {
    // get_anchors_for_preset (ssdutils.py:76-117)
    for (size_t k = 0; k < P.maps.size(); ++k) {
        const MapSpec& m = P.maps[k];
        std::vector<std::pair<double, double>> sizes;
        std::vector<double> rs; rs.push_back(1.0); for (double r : m.ratios) rs.push_back(r);
        for (double r : rs) { double q = std::sqrt(r); sizes.push_back({m.scale * q, m.scale / q}); }
        double nxt = k + 1 < P.maps.size() ? P.maps[k + 1].scale : P.extra_scale;
        double sp = std::sqrt(m.scale * nxt);
        sizes.push_back({sp, sp});
        for (auto& s : sizes)
            for (int j = 0; j < m.size; ++j)
                for (int i = 0; i < m.size; ++i) {
                    out.push_back((i + 0.5) / (double)m.size); out.push_back((j + 0.5) / (double)m.size);
                    out.push_back(s.first); out.push_back(s.second);
                }
    }
}

}  // namespace
}  // namespace ssdb

// ============================================================================ C ABI
extern "C" {

int ssdb_version(void) { return 100; }
const char* ssdb_last_error(void) { return ssdb::g_err; }
long long ssdb_launch_count(void) { return ssdb::g_launches; }

int ssdb_device_ok(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("no CUDA device: this library has no CPU fallback"); return SSDB_ENOGPU; }
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10) {
        set_error("device compute capability %d.x is not sm_100: this library is built for B200 only", major);
        return SSDB_ENOGPU;
    }
    return SSDB_OK;
}

int ssdb_create(const char* preset, int num_classes, int max_batch, unsigned flags, ssdb_net** out) {
    SSDB_REQUIRE(preset && out && max_batch >= 1 && num_classes >= 1 && num_classes + 5 <= 32, "bad arguments");
    SSDB_REQUIRE((flags & ~(unsigned)SSDB_FLAG_INFERENCE) == 0, "unknown flags");
    int rc = ssdb_device_ok(); if (rc) return rc;
    const Preset* P = find_preset(preset);
    if (!P) { set_error("No such preset: %s", preset); return SSDB_ENOTFOUND; }
    ssdb_net* n = new ssdb_net();
    n->preset = P; n->C = num_classes; n->V = num_classes + 5; n->S = P->image; n->max_batch = max_batch;
    n->inference = (flags & SSDB_FLAG_INFERENCE) != 0;
    { const char* ov = getenv("SSDB_FUSE_POOL"); if (ov && !atoi(ov)) n->fuse_pool = false; }
    const bool train = !n->inference;
    // SSDB_CONV: (unset) / "split" = tensor cores with split bf16 operands (fp32-grade products, the product mode);
    // "tf32" = tensor cores with tf32 operands (10-bit significands: outside the 1e-3 parity bar, kept for comparison);
    // "simt" = fp32 CUDA-core kernels only
    const char* mode = getenv("SSDB_CONV");
    if (mode && !strcmp(mode, "simt")) { n->conv_mode = SSDB_CONV_SIMT; n->round = false; n->fmt = ACT_F32; }
    else if (mode && !strcmp(mode, "tf32")) { n->round = true; n->fmt = ACT_F32; }
    else if (mode && strcmp(mode, "split") && strcmp(mode, "auto") && mode[0]) { set_error("SSDB_CONV=%s: expected split, tf32 or simt", mode); delete n; return SSDB_EINVAL; }
    build_plan(n);
    if (n->A != P->num_anchors) { set_error("internal: anchor count %d != %d", n->A, P->num_anchors); delete n; return SSDB_EINVAL; }
    // workspace sizes
    size_t partial = (size_t)1184 * 1024 + 296 * 1024;
    size_t dzh = 0;
    for (const Op& op : n->ops) {
        if (op.type != OP_CONV) continue;
        ConvGeom g = geom_of(n, op, max_batch);
        size_t w = conv_simt_wgrad_ws(g); if (w > partial) partial = w;
        w = conv_tc_wgrad_ws(g, n->fmt); if (w > partial) partial = w;
    }
    for (int oi = (int)n->ops.size() - 1; oi >= 0; --oi) {       // one dz buffer per head, in the order the backward visits them
        const Op& op = n->ops[oi];
        if (op.type != OP_CONV || !op.head) continue;
        ConvGeom g = geom_of(n, op, max_batch);
        n->dz_head_off.push_back(dzh);
        dzh += ((size_t)max_batch * g.H * g.W * op.cout + 255) / 256 * 256;
    }
    n->partial_floats = partial;
    size_t bav = (size_t)max_batch * n->A * n->V;
    size_t img = (size_t)max_batch * n->S * n->S * 3;
#define ALLOC(ptr, count, type) SSDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&(ptr)), (size_t)(count) * sizeof(type)))
    // an inference handle (SSDB_FLAG_INFERENCE: the frozen model of export_model.py / detect.py) keeps only what the forward
    // and the detection kernels touch: no gradients, momentum, gradient activations, label / loss buffers or pool codes
    ALLOC(n->params, n->n_flat, float);
    ALLOC(n->wt, n->wt_floats ? n->wt_floats : 1, float);
    ALLOC(n->acts, n->act_floats_per_image * max_batch, float);
    ALLOC(n->out, bav, float); ALLOC(n->result, bav, float);
    ALLOC(n->images_stage, img, float);
    ALLOC(n->small_ws, 4096 + 2 * (size_t)max_batch, float);
    ALLOC(n->counter, 1, unsigned int); ALLOC(n->decay_mask, n->n_flat / OPT_BLOCK, unsigned char);
    ALLOC(n->anchors, (size_t)n->A * 4, double);
    if (train) {
        ALLOC(n->grads, n->n_flat, float); ALLOC(n->moms, n->n_flat, float); ALLOC(n->wr, n->n_flat, float);
        ALLOC(n->gacts, n->act_floats_per_image * max_batch, float);
        ALLOC(n->out_grad, bav, float); ALLOC(n->labels_stage, bav, float);
        ALLOC(n->dz_head, dzh ? dzh : 1, float);
        ALLOC(n->partial, partial, float); ALLOC(n->l2n_partial, (size_t)296 * 1024, float);
        ALLOC(n->gt_stage, (size_t)max_batch * 128 * 5, double); ALLOC(n->gt_count_stage, max_batch, int);
        ALLOC(n->match_stage, (size_t)max_batch * n->A, int);
        ALLOC(n->loss_ws, multibox_loss_ws_bytes(max_batch, n->A), unsigned char);
        SSDB_CUDA(cudaMemset(n->loss_ws, 0, multibox_loss_ws_bytes(max_batch, n->A)));
        for (const Op& op : n->ops)
            if (op.type == OP_POOL && op.stride == 1) { const Buf& bo = n->bufs[op.out]; ALLOC(n->pool5_arg, (size_t)max_batch * bo.H * bo.W * bo.C, unsigned char); }
    }
    // (inference handles keep the code bytes too: they are what lets conv1_2 / conv2_2 write their pool from the epilogue)
    n->pool_code.assign(n->ops.size(), nullptr);
    if (!getenv("SSDB_POOL_CODE") || atoi(getenv("SSDB_POOL_CODE")))
        for (size_t i = 0; i < n->ops.size(); ++i) {
            const Op& op = n->ops[i];
            if (op.type != OP_POOL || op.k != 2 || op.stride != 2 || op.pad != 0) continue;
            const Buf& bo = n->bufs[op.out];
            ALLOC(n->pool_code[i], (size_t)max_batch * bo.H * bo.W * bo.C, unsigned char);
        }
    if (n->conv_mode != SSDB_CONV_SIMT) {
        const Op& c1 = n->ops[0];
        ConvGeom g1 = geom_of(n, c1, max_batch); g1.Cin = 32; g1.k = 1; g1.pad_t = g1.pad_l = 0;
        if (conv_tc_supported_fprop(g1) && (!train || conv_tc_supported_wgrad(g1, n->fmt))) {
            ALLOC(n->patches, (size_t)max_batch * n->S * n->S * 32, float);
            ALLOC(n->c1_w32, 32 * c1.cout, float); ALLOC(n->c1_wt, 32 * c1.cout, float); ALLOC(n->c1_dw32, 32 * c1.cout, float);
            size_t w = train ? conv_tc_wgrad_ws(g1, n->fmt) : 0;
            if (w > n->partial_floats) { cudaFree(n->partial); n->partial_floats = w; ALLOC(n->partial, w, float); }
        }
    }
#undef ALLOC
    SSDB_CUDA(cudaMemset(n->params, 0, n->n_flat * sizeof(float)));
    SSDB_CUDA(cudaMemset(n->wt, 0, (n->wt_floats ? n->wt_floats : 1) * sizeof(float)));
    SSDB_CUDA(cudaMemset(n->counter, 0, sizeof(unsigned int)));
    if (train) {
        SSDB_CUDA(cudaMemset(n->grads, 0, n->n_flat * sizeof(float)));
        SSDB_CUDA(cudaMemset(n->moms, 0, n->n_flat * sizeof(float)));
        SSDB_CUDA(cudaMemset(n->gacts, 0, n->act_floats_per_image * max_batch * sizeof(float)));
    }
    std::vector<unsigned char> mask(n->n_flat / OPT_BLOCK, 0);
    for (const Master& m : n->masters)
        if (m.decay) for (size_t b = m.off / OPT_BLOCK; b < (m.off + m.count + OPT_BLOCK - 1) / OPT_BLOCK; ++b) mask[b] = 1;
    SSDB_CUDA(cudaMemcpy(n->decay_mask, mask.data(), mask.size(), cudaMemcpyHostToDevice));
    std::vector<double> anc; anchors_host(*P, anc);
    SSDB_CUDA(cudaMemcpy(n->anchors, anc.data(), anc.size() * sizeof(double), cudaMemcpyHostToDevice));
    SSDB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&n->host_small), 64 * sizeof(float)));
    SSDB_CUDA(cudaStreamCreateWithFlags(&n->own_stream, cudaStreamNonBlocking));
    SSDB_CUDA(cudaStreamCreateWithFlags(&n->copy_stream, cudaStreamNonBlocking));
    SSDB_CUDA(cudaStreamCreateWithFlags(&n->side_stream, cudaStreamNonBlocking));
    SSDB_CUDA(cudaEventCreateWithFlags(&n->ev_fork, cudaEventDisableTiming));
    SSDB_CUDA(cudaEventCreateWithFlags(&n->ev_join, cudaEventDisableTiming));
    SSDB_CUDA(cudaEventCreateWithFlags(&n->ev_side, cudaEventDisableTiming));
    SSDB_CUDA(cudaEventCreateWithFlags(&n->ev_l2, cudaEventDisableTiming));
    SSDB_CUDA(cudaEventCreateWithFlags(&n->ev_labels, cudaEventDisableTiming));
    SSDB_CUDA(cudaEventCreateWithFlags(&n->ev_result, cudaEventDisableTiming));
    SSDB_CUDA(cudaEventCreateWithFlags(&n->ev_images, cudaEventDisableTiming));
    for (int c = 0; c < 8; ++c) SSDB_CUDA(cudaEventCreateWithFlags(&n->ev_chunk[c], cudaEventDisableTiming));
    {   // buckets: [mod_conv6 .. end) (conv6/7, extras, scale, heads: final first), [conv4_1 .. mod_conv6), [start .. conv4_1)
        const char* cuts[] = {"mod_conv6", "conv4_1"};
        long long end = (long long)n->n_flat;
        for (const char* cname : cuts)
            for (size_t oi = 0; oi < n->ops.size(); ++oi)
                if (n->ops[oi].name == cname && n->ops[oi].w >= 0) {
                    const long long begin = (long long)n->masters[n->ops[oi].w].off;
                    n->bucket_begin[n->n_buckets] = begin; n->bucket_end[n->n_buckets] = end; n->bucket_last_op[n->n_buckets] = (int)oi;
                    end = begin; ++n->n_buckets;
                }
        n->bucket_begin[n->n_buckets] = 0; n->bucket_end[n->n_buckets] = end; n->bucket_last_op[n->n_buckets] = 0; ++n->n_buckets;
        for (int b = 0; b < n->n_buckets; ++b) SSDB_CUDA(cudaEventCreateWithFlags(&n->ev_bucket[b], cudaEventDisableTiming));
    }
    *out = n;
    return SSDB_OK;
}

int ssdb_destroy(ssdb_net* n) {
    if (!n) return SSDB_OK;
    cudaDeviceSynchronize();
    drop_det_graphs(n);
    void* ptrs[] = {n->pool5_arg, n->patches, n->c1_w32, n->c1_wt, n->c1_dw32, n->wr, n->params, n->grads, n->moms, n->wt, n->acts, n->gacts, n->out, n->out_grad, n->result, n->labels_stage,
                    n->dz_head, n->images_stage, n->partial, n->small_ws, n->counter, n->decay_mask, n->anchors, n->loss_ws, n->det_ws,
                    n->gt_stage, n->gt_count_stage, n->match_stage, n->det_rows, n->det_counts, n->l2n_partial};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (unsigned char* p : n->pool_code) if (p) cudaFree(p);
    if (n->host_small) cudaFreeHost(n->host_small);
    if (n->own_stream) cudaStreamDestroy(n->own_stream);
    if (n->copy_stream) cudaStreamDestroy(n->copy_stream);
    if (n->side_stream) cudaStreamDestroy(n->side_stream);
    if (n->ev_fork) cudaEventDestroy(n->ev_fork);
    if (n->ev_join) cudaEventDestroy(n->ev_join);
    if (n->ev_side) cudaEventDestroy(n->ev_side);
    if (n->ev_l2) cudaEventDestroy(n->ev_l2);
    if (n->ev_labels) cudaEventDestroy(n->ev_labels);
    if (n->ev_result) cudaEventDestroy(n->ev_result);
    if (n->ev_images) cudaEventDestroy(n->ev_images);
    for (int c = 0; c < 8; ++c) if (n->ev_chunk[c]) cudaEventDestroy(n->ev_chunk[c]);
    for (int b = 0; b < ssdb_net::MAX_BUCKETS; ++b) if (n->ev_bucket[b]) cudaEventDestroy(n->ev_bucket[b]);
    delete n;
    return SSDB_OK;
}

int ssdb_num_anchors(const ssdb_net* n) { return n ? n->A : SSDB_EINVAL; }
int ssdb_image_size(const ssdb_net* n) { return n ? n->S : SSDB_EINVAL; }
int ssdb_num_tensors(const ssdb_net* n) { return n ? (int)n->refs.size() : SSDB_EINVAL; }

int ssdb_tensor_info(const ssdb_net* n, int index, char* name_out, int name_cap, int* rank_out, int shape_out[4]) {
    SSDB_REQUIRE(n && index >= 0 && index < (int)n->refs.size(), "bad tensor index");
    const RefTensor& r = n->refs[index];
    if (name_out && name_cap > 0) { strncpy(name_out, r.name.c_str(), name_cap - 1); name_out[name_cap - 1] = 0; }
    if (rank_out) *rank_out = r.rank;
    if (shape_out) for (int i = 0; i < 4; ++i) shape_out[i] = r.shape[i];
    return SSDB_OK;
}

static int tensor_io(ssdb_net* n, const char* name, int which, float* host, long long count, bool write) {
    SSDB_REQUIRE(n && name && host && which >= 0 && which <= 2, "bad arguments");
    auto it = n->ref_index.find(name);
    if (it == n->ref_index.end()) { set_error("no such tensor: %s", name); return SSDB_ENOTFOUND; }
    const RefTensor& r = n->refs[it->second];
    const Master& m = n->masters[r.master];
    long long want = 1; for (int i = 0; i < r.rank; ++i) want *= r.shape[i];
    SSDB_REQUIRE(count == want, "element count does not match the tensor shape");
    SSDB_REQUIRE(which == 0 || !n->inference, "an inference handle has no gradients or momentum");
    float* base = (which == 0 ? n->params : which == 1 ? n->grads : n->moms) + m.off;
    SSDB_CUDA(cudaDeviceSynchronize());
    if (write && which == 0) drop_det_graphs(n);
    int last = m.shape[m.rank - 1];
    if (r.cols == last && r.col0 == 0) {
        if (write) SSDB_CUDA(cudaMemcpy(base, host, count * sizeof(float), cudaMemcpyHostToDevice));
        else SSDB_CUDA(cudaMemcpy(host, base, count * sizeof(float), cudaMemcpyDeviceToHost));
    } else {
        long long rows = (long long)m.count / last;
        if (write) SSDB_CUDA(cudaMemcpy2D(base + r.col0, (size_t)last * sizeof(float), host, (size_t)r.cols * sizeof(float),
                                          (size_t)r.cols * sizeof(float), (size_t)rows, cudaMemcpyHostToDevice));
        else SSDB_CUDA(cudaMemcpy2D(host, (size_t)r.cols * sizeof(float), base + r.col0, (size_t)last * sizeof(float),
                                    (size_t)r.cols * sizeof(float), (size_t)rows, cudaMemcpyDeviceToHost));
    }
    if (write && which == 0) n->wt_dirty = true;
    return SSDB_OK;
}

int ssdb_get_tensor(ssdb_net* n, const char* name, int which, float* host_out, long long count) { return tensor_io(n, name, which, host_out, count, false); }
int ssdb_set_tensor(ssdb_net* n, const char* name, int which, const float* host_in, long long count) { return tensor_io(n, name, which, const_cast<float*>(host_in), count, true); }

int ssdb_flat_buffer(ssdb_net* n, int which, void** dev_ptr_out, long long* count_out) {
    SSDB_REQUIRE(n && which >= 0 && which <= 2 && dev_ptr_out && count_out, "bad arguments");
    SSDB_REQUIRE(which == 0 || !n->inference, "an inference handle has no gradients or momentum");
    *dev_ptr_out = which == 0 ? n->params : which == 1 ? n->grads : n->moms;
    *count_out = (long long)n->n_flat;
    return SSDB_OK;
}

int ssdb_grad_buckets(const ssdb_net* n, int cap, long long* begin_out, long long* end_out) {
    SSDB_REQUIRE(n && begin_out && end_out && cap >= 1, "bad arguments");
    const int k = n->n_buckets < cap ? n->n_buckets : cap;
    for (int b = 0; b < k; ++b) { begin_out[b] = n->bucket_begin[b]; end_out[b] = n->bucket_end[b]; }
    return n->n_buckets;
}

int ssdb_wait_grad_bucket(ssdb_net* n, int bucket, void* stream) {
    SSDB_REQUIRE(n && bucket >= 0 && bucket < n->n_buckets, "bad bucket index");
    SSDB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, n->ev_bucket[bucket], 0));
    return SSDB_OK;
}

int ssdb_set_preprocess(ssdb_net* n, int swap_rb, const float mean[3]) {
    SSDB_REQUIRE(n && mean, "bad arguments");
    n->swap_rb = swap_rb ? 1 : 0; n->mean[0] = mean[0]; n->mean[1] = mean[1]; n->mean[2] = mean[2];
    return SSDB_OK;
}

int ssdb_forward(ssdb_net* n, const float* images_dev, int B, float* result_dev, void* stream) {
    SSDB_REQUIRE(n && images_dev, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = run_forward(n, images_dev, B, st); if (rc) return rc;
    rc = softmax_result(n->out, (long long)B * n->A, n->C, result_dev ? result_dev : n->result, st);
    return rc;
}

// Host entry points: the image batch goes up in four or eight chunks on the copy stream and conv1_1 runs chunk by chunk
// behind it on the compute stream, so only the first chunk's transfer (not the whole 69 MB at batch 64) is exposed (measured: 29.16 -> 28.43 ms per step end to end with four chunks).
// *first_done tells the caller to skip conv1_1 in run_forward.
static int upload_images_chunked(ssdb_net* n, const float* images_host, int B, cudaStream_t st, cudaStream_t cs, bool* first_done) {
    const size_t img = (size_t)n->S * n->S * 3;
    const char* ov = getenv("SSDB_CHUNKED_UPLOAD");
    const bool chunked = n->patches && B >= 8 && !(ov && atoi(ov) == 0);
    *first_done = false;
    if (!chunked) {
        SSDB_CUDA(cudaMemcpyAsync(n->images_stage, images_host, (size_t)B * img * sizeof(float), cudaMemcpyHostToDevice, cs));
        SSDB_CUDA(cudaEventRecord(n->ev_images, cs));
        SSDB_CUDA(cudaStreamWaitEvent(st, n->ev_images, 0));
        return SSDB_OK;
    }
    if (n->wt_dirty) { int rc = repack_filters(n, st); if (rc) return rc; }
    const int chunks = B >= 32 ? 8 : 4;
    const int per = (B + chunks - 1) / chunks;
    for (int c = 0; c < chunks; ++c) {
        const int b0 = c * per, bc = B - b0 < per ? B - b0 : per;
        if (bc <= 0) break;
        SSDB_CUDA(cudaMemcpyAsync(n->images_stage + (size_t)b0 * img, images_host + (size_t)b0 * img, (size_t)bc * img * sizeof(float),
                                  cudaMemcpyHostToDevice, cs));
        SSDB_CUDA(cudaEventRecord(n->ev_chunk[c], cs));
        SSDB_CUDA(cudaStreamWaitEvent(st, n->ev_chunk[c], 0));
        int rc = run_first_conv_chunk(n, n->images_stage, b0, bc, st); if (rc) return rc;
    }
    *first_done = true;
    return SSDB_OK;
}

int ssdb_forward_host(ssdb_net* n, const float* images_host, int B, float* result_host) {
    SSDB_REQUIRE(n && images_host && result_host && B >= 1 && B <= n->max_batch, "bad arguments");
    cudaStream_t st = n->own_stream;
    bool first_done = false;
    int rc = upload_images_chunked(n, images_host, B, st, n->copy_stream, &first_done); if (rc) return rc;
    rc = run_forward(n, n->images_stage, B, st, first_done); if (rc) return rc;
    rc = softmax_result(n->out, (long long)B * n->A, n->C, n->result, st); if (rc) return rc;
    SSDB_CUDA(cudaMemcpyAsync(result_host, n->result, (size_t)B * n->A * n->V * sizeof(float), cudaMemcpyDeviceToHost, st));
    SSDB_CUDA(cudaStreamSynchronize(st));
    return SSDB_OK;
}

int ssdb_read_output_host(ssdb_net* n, int B, float* output_host) {
    SSDB_REQUIRE(n && output_host && B >= 1 && B <= n->max_batch, "bad arguments");
    SSDB_REQUIRE(n->last_B >= B, "no forward pass has produced that many rows yet");
    SSDB_CUDA(cudaDeviceSynchronize());
    SSDB_CUDA(cudaMemcpy(output_host, n->out, (size_t)B * n->A * n->V * sizeof(float), cudaMemcpyDeviceToHost));
    return SSDB_OK;
}

// Test / diagnosis hook: any intermediate tensor of the last step as plain float32 on the host.
//   "<op name>"        activation written by that op (conv1_1 ... conv11_2, pool1..4, mod_pool5, l2_norm_conv4_3), NHWC
//   "grad:<op name>"   gradient with respect to that activation (after a backward)
//   "output" / "output_grad"   the raw head output [B, A, C+5] / its gradient
int ssdb_debug_read(ssdb_net* n, const char* name, int B, float* host_out, long long count) {
    SSDB_REQUIRE(n && name && host_out && B >= 1 && B <= n->max_batch, "bad arguments");
    SSDB_CUDA(cudaDeviceSynchronize());
    std::string nm(name);
    SSDB_REQUIRE(!(n->inference && nm == "output_grad"), "an inference handle has no gradients");
    if (nm == "output" || nm == "output_grad") {
        SSDB_REQUIRE(count == (long long)B * n->A * n->V, "element count does not match [B, A, C+5]");
        SSDB_CUDA(cudaMemcpy(host_out, nm == "output" ? n->out : n->out_grad, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost));
        return SSDB_OK;
    }
    const bool grad = nm.rfind("grad:", 0) == 0;
    SSDB_REQUIRE(!n->inference || (!grad && nm != "output_grad"), "an inference handle has no gradients");
    if (grad) nm = nm.substr(5);
    for (const Op& op : n->ops) {
        if (op.name != nm || op.out < 0) continue;
        const Buf& b = n->bufs[op.out];
        const long long want = (long long)B * b.H * b.W * b.C;
        SSDB_REQUIRE(count == want, "element count does not match the activation shape");
        const size_t oi = &op - n->ops.data();
        if (!grad && oi < n->fused_now.size() && n->fused_now[oi] == 2) {
            set_error("the activation of %s was not materialised: its max pool is fused into the convolution (SSDB_FUSE_POOL=0 disables it)", name);
            return SSDB_ESTATE;
        }
        const float* src = grad ? n->gact(op.out, B) : n->act(op.out, B);
        if (n->fmt == ACT_S32) {
            float* tmp = nullptr;
            SSDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&tmp), (size_t)want * sizeof(float)));
            int rc = unsplit_copy(src, tmp, want, nullptr);
            if (!rc && cudaMemcpy(host_out, tmp, (size_t)want * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("copy failed"); rc = SSDB_ECUDA; }
            cudaFree(tmp);
            return rc;
        }
        SSDB_CUDA(cudaMemcpy(host_out, src, (size_t)want * sizeof(float), cudaMemcpyDeviceToHost));
        return SSDB_OK;
    }
    set_error("no such activation: %s", name);
    return SSDB_ENOTFOUND;
}

int ssdb_debug_shape(const ssdb_net* n, const char* name, int shape_out[3]) {
    SSDB_REQUIRE(n && name && shape_out, "bad arguments");
    for (const Op& op : n->ops)
        if (op.name == name && op.out >= 0) {
            const Buf& b = n->bufs[op.out];
            shape_out[0] = b.H; shape_out[1] = b.W; shape_out[2] = b.C;
            return SSDB_OK;
        }
    set_error("no such activation: %s", name);
    return SSDB_ENOTFOUND;
}

static int loss_and_finalize(ssdb_net* n, const float* labels_dev, const double* gt_dev, const int* gt_count_dev, int G, int B,
                             float weight_decay, float grad_scale, bool want_grad, float* losses_out_dev, float* result_dev, cudaStream_t st,
                             int* match_out_dev = nullptr) {
    float* conf_loc = n->small_ws + 4;
    float* l2s = n->small_ws + 6;
    ProfScope ps(n, st, "loss");
    int rc = multibox_loss_launch(n->out, labels_dev, gt_dev, gt_count_dev, G, n->anchors, B, n->A, n->C, grad_scale, conf_loc,
                                  want_grad ? n->out_grad : nullptr, result_dev ? result_dev : n->result, match_out_dev, n->loss_ws, st);
    if (rc) return rc;
    if (n->l2_pending) { SSDB_CUDA(cudaStreamWaitEvent(st, n->ev_l2, 0)); n->l2_pending = false; }     // summed beside the forward
    else { rc = l2_sum(n->params, (long long)n->n_flat, n->decay_mask, n->small_ws + 8, l2s, st); if (rc) return rc; }
    finalize_losses_kernel<<<1, 32, 0, st>>>(conf_loc, l2s, weight_decay, losses_out_dev ? losses_out_dev : n->small_ws);
    SSDB_LAUNCH_CHECK();
    return SSDB_OK;
}

int ssdb_apply_update(ssdb_net* n, float lr, float momentum, float weight_decay, float grad_post_scale, void* stream) {
    SSDB_REQUIRE(n && !n->inference, "bad arguments (or an inference handle)");
    drop_det_graphs(n);
    ProfScope ps(n, (cudaStream_t)stream, "update");
    int rc = sgd_momentum(n->params, n->grads, n->moms, (long long)n->n_flat, n->decay_mask, lr, momentum, weight_decay, grad_post_scale,
                          (cudaStream_t)stream);
    n->wt_dirty = true; n->have_forward = false;
    return rc;
}

int ssdb_train_step(ssdb_net* n, const float* images_dev, const float* labels_dev, const double* gt_dev, const int* gt_count_dev, int G,
                    int B, float lr, float momentum, float weight_decay, float grad_scale, int apply_update, float* losses_out_dev,
                    float* result_dev, void* stream) {
    SSDB_REQUIRE(n && images_dev && (labels_dev || (gt_dev && gt_count_dev)), "bad arguments");
    SSDB_REQUIRE(!n->inference, "training entry point called on an inference handle");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = remember_images(n, images_dev, B, st); if (rc) return rc;
    rc = run_forward(n, images_dev, B, st, false, true); if (rc) return rc;
    rc = loss_and_finalize(n, labels_dev, gt_dev, gt_count_dev, G, B, weight_decay, grad_scale, true, losses_out_dev, result_dev, st); if (rc) return rc;
    rc = run_backward(n, B, st); if (rc) return rc;
    if (apply_update) rc = ssdb_apply_update(n, lr, momentum, weight_decay, 1.0f, stream);
    return rc;
}

// ground truth of the host entry points: every label id must be a class index (the reference would raise an IndexError on
// transforms.py:107; unchecked it would index out of bounds in the kernels) and every count within [0, G]
static int check_gt_host(const double* gt, const int* gt_count, int B, int G, int C) {
    SSDB_REQUIRE(G >= 1 && G <= 128, "G (ground-truth slots per image) must be in [1,128]");
    for (int b = 0; b < B; ++b) {
        SSDB_REQUIRE(gt_count[b] >= 0 && gt_count[b] <= G, "ground-truth count outside [0, G]");
        for (int k = 0; k < gt_count[b]; ++k) {
            const double id = gt[((size_t)b * G + k) * 5];
            if (!(id >= 0.0 && id < (double)C && id == (double)(int)id)) {
                set_error("ground-truth box %d of image %d has label id %g: expected an integer in [0, %d)", k, b, id, C);
                return SSDB_EINVAL;
            }
        }
    }
    return SSDB_OK;
}

static int train_step_host_impl(ssdb_net* n, const float* images_host, const float* labels_host, const double* gt_host, const int* gt_count_host,
                                int G, int B, float lr, float momentum, float weight_decay, int apply_update, float* losses_out_host,
                                float* result_host, int* match_out_host) {
    SSDB_REQUIRE(n && images_host && (labels_host || (gt_host && gt_count_host)) && B >= 1 && B <= n->max_batch, "bad arguments");
    SSDB_REQUIRE(!n->inference, "training entry point called on an inference handle");
    if (!labels_host) { int rc = check_gt_host(gt_host, gt_count_host, B, G, n->C); if (rc) return rc; }
    // copies ride a second stream: the labels arrive while the forward runs (they are first needed by the loss) and the
    // result leaves while the backward runs; only the image upload is on the critical path
    cudaStream_t st = n->own_stream, cs = n->copy_stream;
    const size_t bav = (size_t)B * n->A * n->V * sizeof(float);
    // both uploads share one DMA direction: images first (critical path, chunked so that conv1_1 starts after the first
    // quarter), labels (873 KB per image) or raw ground truth (40 bytes per box) behind them
    bool first_done = false;
    int rc = upload_images_chunked(n, images_host, B, st, cs, &first_done); if (rc) return rc;
    if (labels_host) SSDB_CUDA(cudaMemcpyAsync(n->labels_stage, labels_host, bav, cudaMemcpyHostToDevice, cs));
    else {
        SSDB_CUDA(cudaMemcpyAsync(n->gt_stage, gt_host, (size_t)B * G * 5 * sizeof(double), cudaMemcpyHostToDevice, cs));
        SSDB_CUDA(cudaMemcpyAsync(n->gt_count_stage, gt_count_host, (size_t)B * sizeof(int), cudaMemcpyHostToDevice, cs));
    }
    SSDB_CUDA(cudaEventRecord(n->ev_labels, cs));
    rc = run_forward(n, n->images_stage, B, st, first_done, true); if (rc) return rc;
    SSDB_CUDA(cudaStreamWaitEvent(st, n->ev_labels, 0));
    const bool want_grad = apply_update >= 0;
    if (labels_host) rc = loss_and_finalize(n, n->labels_stage, nullptr, nullptr, 0, B, weight_decay, 1.0f, want_grad, n->small_ws, n->result, st);
    else rc = loss_and_finalize(n, nullptr, n->gt_stage, n->gt_count_stage, G, B, weight_decay, 1.0f, want_grad, n->small_ws, n->result, st,
                                match_out_host ? n->match_stage : nullptr);
    if (rc) return rc;
    if (result_host || match_out_host) {
        SSDB_CUDA(cudaEventRecord(n->ev_result, st));
        SSDB_CUDA(cudaStreamWaitEvent(cs, n->ev_result, 0));
        if (result_host) SSDB_CUDA(cudaMemcpyAsync(result_host, n->result, bav, cudaMemcpyDeviceToHost, cs));
        if (match_out_host) SSDB_CUDA(cudaMemcpyAsync(match_out_host, n->match_stage, (size_t)B * n->A * sizeof(int), cudaMemcpyDeviceToHost, cs));
    }
    if (apply_update >= 0) { rc = run_backward(n, B, st); if (rc) return rc; }
    if (apply_update == 1) { rc = ssdb_apply_update(n, lr, momentum, weight_decay, 1.0f, st); if (rc) return rc; }
    SSDB_CUDA(cudaMemcpyAsync(n->host_small, n->small_ws, 4 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (apply_update == 2) return SSDB_OK;         // ssdb_train_step_host_begin: the caller finishes with ssdb_train_step_host_end
    SSDB_CUDA(cudaStreamSynchronize(st));
    SSDB_CUDA(cudaStreamSynchronize(cs));
    if (losses_out_host) memcpy(losses_out_host, n->host_small, 4 * sizeof(float));
    return SSDB_OK;
}

int ssdb_train_step_host_begin(ssdb_net* n, const float* images_host, const float* labels_host, const double* gt_host, const int* gt_count_host,
                               int G, int B, float weight_decay, float* result_host, int* match_out_host) {
    return train_step_host_impl(n, images_host, labels_host, gt_host, gt_count_host, G, B, 0.f, 0.f, weight_decay, 2, nullptr, result_host,
                                labels_host ? nullptr : match_out_host);
}

int ssdb_train_step_host_end(ssdb_net* n, float* losses_out_host) {
    SSDB_REQUIRE(n, "bad arguments");
    SSDB_CUDA(cudaStreamSynchronize(n->own_stream));
    SSDB_CUDA(cudaStreamSynchronize(n->copy_stream));
    if (losses_out_host) memcpy(losses_out_host, n->host_small, 4 * sizeof(float));
    return SSDB_OK;
}

int ssdb_train_step_host(ssdb_net* n, const float* images_host, const float* labels_host, int B, float lr, float momentum,
                         float weight_decay, float* losses_out_host, float* result_host) {
    SSDB_REQUIRE(labels_host, "labels are required (ssdb_train_step_host_gt takes raw ground truth)");
    return train_step_host_impl(n, images_host, labels_host, nullptr, nullptr, 0, B, lr, momentum, weight_decay, 1, losses_out_host, result_host, nullptr);
}

int ssdb_train_step_host_noupdate(ssdb_net* n, const float* images_host, const float* labels_host, int B, float weight_decay,
                                  float* losses_out_host, float* result_host) {
    SSDB_REQUIRE(labels_host, "labels are required (ssdb_train_step_host_gt takes raw ground truth)");
    return train_step_host_impl(n, images_host, labels_host, nullptr, nullptr, 0, B, 0.f, 0.f, weight_decay, 0, losses_out_host, result_host, nullptr);
}

int ssdb_train_step_host_gt(ssdb_net* n, const float* images_host, const double* gt_host, const int* gt_count_host, int G, int B,
                            float lr, float momentum, float weight_decay, int apply_update, float* losses_out_host, float* result_host,
                            int* match_out_host) {
    SSDB_REQUIRE(gt_host && gt_count_host, "ground truth is required");
    return train_step_host_impl(n, images_host, nullptr, gt_host, gt_count_host, G, B, lr, momentum, weight_decay,
                                apply_update > 0 ? 1 : (apply_update < 0 ? -1 : 0), losses_out_host, result_host, match_out_host);
}

// forward, then decode + top-k + class-wise NMS straight on the device-resident result: only the detections go back
int ssdb_forward_detect_host(ssdb_net* n, const float* images_host, int B, float conf_thr, int cap, double iou_thr, int* dets_out_host,
                             int* counts_out_host, float* result_host) {
    SSDB_REQUIRE(n && images_host && dets_out_host && counts_out_host && B >= 1 && B <= n->max_batch, "bad arguments");
    cudaStream_t st = n->own_stream;
    const int cap_eff = (cap > 0 && cap < n->A) ? cap : n->A;
    if (!n->det_rows) {
        SSDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&n->det_rows), (size_t)n->max_batch * n->A * 8 * sizeof(int)));
        SSDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&n->det_counts), (size_t)n->max_batch * 2 * sizeof(int)));
    }
    if (!n->det_ws) SSDB_CUDA(cudaMalloc(&n->det_ws, decode_nms_scratch_bytes(n->max_batch, n->A, 0)));
    int rc = SSDB_OK;
    // Frozen model: the whole device side (im2col, ~60 forward launches, softmax, decode + NMS) is one CUDA graph per
    // (batch, threshold, cap, IoU), captured on the second call with that key (the first runs eagerly: it sets the kernels'
    // shared-memory attributes and allocates the lazily created tables, which must not happen inside a capture).
    static int graphs_on = -1;
    if (graphs_on < 0) { const char* ov = getenv("SSDB_GRAPH"); graphs_on = ov ? (atoi(ov) ? 1 : 0) : 1; }
    if (n->inference && graphs_on && !result_host) {
        if (n->wt_dirty) { rc = repack_filters(n, st); if (rc) return rc; drop_det_graphs(n); }
        ssdb_net::DetGraph* dg = nullptr;
        for (auto& g : n->det_graphs) if (g.B == B && g.thr == conf_thr && g.cap == cap && g.iou == iou_thr) dg = &g;
        if (!dg) { n->det_graphs.push_back(ssdb_net::DetGraph{B, conf_thr, cap, iou_thr, false, nullptr}); dg = &n->det_graphs.back(); }
        SSDB_CUDA(cudaMemcpyAsync(n->images_stage, images_host, (size_t)B * n->S * n->S * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
        auto device_side = [&]() -> int {
            int r = run_forward(n, n->images_stage, B, st); if (r) return r;
            r = softmax_result(n->out, (long long)B * n->A, n->C, n->result, st); if (r) return r;
            SSDB_CUDA(cudaMemsetAsync(n->det_rows, 0, (size_t)B * cap_eff * 8 * sizeof(int), st));
            return decode_nms_launch(n->result, B, n->A, n->C, n->anchors, conf_thr, cap, iou_thr, n->det_rows, n->det_counts, n->det_ws,
                                     decode_nms_scratch_bytes(n->max_batch, n->A, 0), st);
        };
        if (!dg->warmed) { rc = device_side(); if (rc) return rc; dg->warmed = true; }
        else {
            if (!dg->exec) {
                cudaGraph_t graph = nullptr;
                SSDB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                rc = device_side();
                cudaError_t ce = cudaStreamEndCapture(st, &graph);
                if (rc || ce != cudaSuccess || !graph) {
                    if (graph) cudaGraphDestroy(graph);
                    if (!rc) { set_error("stream capture of the frozen forward failed: %s", cudaGetErrorString(ce)); rc = SSDB_ECUDA; }
                    return rc;
                }
                ce = cudaGraphInstantiate(&dg->exec, graph, 0);
                cudaGraphDestroy(graph);
                if (ce != cudaSuccess) { dg->exec = nullptr; set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ce)); return SSDB_ECUDA; }
            }
            SSDB_CUDA(cudaGraphLaunch(dg->exec, st));
            ++g_launches;
        }
        SSDB_CUDA(cudaMemcpyAsync(counts_out_host, n->det_counts, (size_t)B * 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        SSDB_CUDA(cudaMemcpyAsync(dets_out_host, n->det_rows, (size_t)B * cap_eff * 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
        SSDB_CUDA(cudaStreamSynchronize(st));
        return SSDB_OK;
    }
    bool first_done = false;
    rc = upload_images_chunked(n, images_host, B, st, n->copy_stream, &first_done); if (rc) return rc;
    rc = run_forward(n, n->images_stage, B, st, first_done); if (rc) return rc;
    rc = softmax_result(n->out, (long long)B * n->A, n->C, n->result, st); if (rc) return rc;
    if (result_host) {       // optional: the caller also wants net.result (it leaves on the copy stream, behind the kernels below)
        SSDB_CUDA(cudaEventRecord(n->ev_result, st));
        SSDB_CUDA(cudaStreamWaitEvent(n->copy_stream, n->ev_result, 0));
        SSDB_CUDA(cudaMemcpyAsync(result_host, n->result, (size_t)B * n->A * n->V * sizeof(float), cudaMemcpyDeviceToHost, n->copy_stream));
    }
    SSDB_CUDA(cudaMemsetAsync(n->det_rows, 0, (size_t)B * cap_eff * 8 * sizeof(int), st));
    rc = decode_nms_launch(n->result, B, n->A, n->C, n->anchors, conf_thr, cap, iou_thr, n->det_rows, n->det_counts, n->det_ws,
                           decode_nms_scratch_bytes(n->max_batch, n->A, 0), st);
    if (rc) return rc;
    SSDB_CUDA(cudaMemcpyAsync(counts_out_host, n->det_counts, (size_t)B * 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    SSDB_CUDA(cudaMemcpy2DAsync(dets_out_host, (size_t)cap_eff * 8 * sizeof(int), n->det_rows, (size_t)cap_eff * 8 * sizeof(int),
                                (size_t)cap_eff * 8 * sizeof(int), (size_t)B, cudaMemcpyDeviceToHost, st));
    SSDB_CUDA(cudaStreamSynchronize(st));
    SSDB_CUDA(cudaStreamSynchronize(n->copy_stream));
    return SSDB_OK;
}

int ssdb_eval_step(ssdb_net* n, const float* images_dev, const float* labels_dev, int B, float weight_decay, float* losses_out_dev,
                   float* result_dev, void* stream) {
    SSDB_REQUIRE(n && images_dev && labels_dev, "bad arguments");
    SSDB_REQUIRE(!n->inference, "loss entry point called on an inference handle");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = run_forward(n, images_dev, B, st); if (rc) return rc;
    return loss_and_finalize(n, labels_dev, nullptr, nullptr, 0, B, weight_decay, 1.0f, false, losses_out_dev, result_dev, st);
}

// ---------------------------------------------------------------- stateless ops
int ssdb_match_anchors(const double* gt_dev, const int* gt_count_dev, int B, int G, const double* anchors_prop_dev, int A, int C,
                       int* match_out_dev, float* labels_out_dev, void* stream) {
    SSDB_REQUIRE(gt_dev && gt_count_dev && anchors_prop_dev, "bad arguments");
    return match_anchors_launch(gt_dev, gt_count_dev, B, G, anchors_prop_dev, A, C, match_out_dev, labels_out_dev, (cudaStream_t)stream);
}

// Device scratch of the stateless *_host entry points: one grow-only arena per device (a process may drive several
// GPUs), carved per call; calls are serialised by the arena mutex (the reference is single-threaded per session anyway).
// Nothing is allocated or freed per call once the arena has grown to the largest request, and an error return leaks nothing.
namespace {
struct HostArena { unsigned char* p = nullptr; size_t cap = 0; };
std::mutex g_arena_mu;
std::map<int, HostArena> g_arenas;
int arena_get(size_t bytes, unsigned char** out) {
    int dev = 0;
    SSDB_CUDA(cudaGetDevice(&dev));
    HostArena& a = g_arenas[dev];
    if (bytes > a.cap) {
        if (a.p) { SSDB_CUDA(cudaDeviceSynchronize()); cudaFree(a.p); a.p = nullptr; a.cap = 0; }
        SSDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&a.p), bytes));
        a.cap = bytes;
    }
    *out = a.p;
    return SSDB_OK;
}
size_t up256(size_t v) { return (v + 255) / 256 * 256; }
}  // namespace

int ssdb_match_anchors_host(const double* gt, const int* gt_count, int B, int G, const double* anchors, int A, int C, int* match_out,
                            float* labels_out) {
    SSDB_REQUIRE(gt && gt_count && anchors && B >= 1 && G >= 1, "bad arguments");
    int rc = ssdb_device_ok(); if (rc) return rc;
    rc = check_gt_host(gt, gt_count, B, G, C); if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_arena_mu);
    const size_t s_gt = up256((size_t)B * G * 5 * 8), s_anc = up256((size_t)A * 4 * 8), s_cnt = up256((size_t)B * 4);
    const size_t s_match = match_out ? up256((size_t)B * A * 4) : 0, s_lab = labels_out ? up256((size_t)B * A * (C + 5) * 4) : 0;
    unsigned char* base = nullptr;
    rc = arena_get(s_gt + s_anc + s_cnt + s_match + s_lab, &base); if (rc) return rc;
    double* d_gt = reinterpret_cast<double*>(base); double* d_anc = reinterpret_cast<double*>(base + s_gt);
    int* d_cnt = reinterpret_cast<int*>(base + s_gt + s_anc);
    int* d_match = match_out ? reinterpret_cast<int*>(base + s_gt + s_anc + s_cnt) : nullptr;
    float* d_lab = labels_out ? reinterpret_cast<float*>(base + s_gt + s_anc + s_cnt + s_match) : nullptr;
    SSDB_CUDA(cudaMemcpyAsync(d_gt, gt, (size_t)B * G * 5 * 8, cudaMemcpyHostToDevice, nullptr));
    SSDB_CUDA(cudaMemcpyAsync(d_anc, anchors, (size_t)A * 4 * 8, cudaMemcpyHostToDevice, nullptr));
    SSDB_CUDA(cudaMemcpyAsync(d_cnt, gt_count, (size_t)B * 4, cudaMemcpyHostToDevice, nullptr));
    rc = match_anchors_launch(d_gt, d_cnt, B, G, d_anc, A, C, d_match, d_lab, nullptr); if (rc) return rc;
    if (match_out) SSDB_CUDA(cudaMemcpyAsync(match_out, d_match, (size_t)B * A * 4, cudaMemcpyDeviceToHost, nullptr));
    if (labels_out) SSDB_CUDA(cudaMemcpyAsync(labels_out, d_lab, (size_t)B * A * (C + 5) * 4, cudaMemcpyDeviceToHost, nullptr));
    SSDB_CUDA(cudaStreamSynchronize(nullptr));
    return SSDB_OK;
}

// grow-only scratch per stream: calls on one stream are ordered, calls on different streams never share a buffer
static int decode_scratch(cudaStream_t st, size_t need, void** out) {
    struct Scratch { void* p = nullptr; size_t cap = 0; };
    static std::map<cudaStream_t, Scratch> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    Scratch& s = cache[st];
    if (need > s.cap) {
        if (s.p) { SSDB_CUDA(cudaStreamSynchronize(st)); cudaFree(s.p); s.p = nullptr; s.cap = 0; }
        SSDB_CUDA(cudaMalloc(&s.p, need)); s.cap = need;
    }
    *out = s.p;
    return SSDB_OK;
}

int ssdb_decode_nms(const float* pred_dev, int B, int A, int C, const double* anchors_prop_dev, float conf_thr, int cap, double iou_thr,
                    int* dets_out_dev, int* counts_out_dev, void* stream) {
    SSDB_REQUIRE(pred_dev && anchors_prop_dev && dets_out_dev && counts_out_dev, "bad arguments");
    size_t sb = decode_nms_scratch_bytes(B, A, cap);
    void* scratch = nullptr;
    int rc = decode_scratch((cudaStream_t)stream, sb, &scratch); if (rc) return rc;
    return decode_nms_launch(pred_dev, B, A, C, anchors_prop_dev, conf_thr, cap, iou_thr, dets_out_dev, counts_out_dev, scratch, sb, (cudaStream_t)stream);
}

int ssdb_decode_nms_host(const float* pred, int B, int A, int C, const double* anchors, float conf_thr, int cap, double iou_thr,
                         int* dets_out, int* counts_out) {
    SSDB_REQUIRE(pred && anchors && dets_out && counts_out && B >= 1, "bad arguments");
    int rc = ssdb_device_ok(); if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_arena_mu);
    const int cap_eff = (cap > 0 && cap < A) ? cap : A;
    const size_t pb = (size_t)B * A * (C + 5) * 4, db = (size_t)B * cap_eff * 8 * 4, sb = decode_nms_scratch_bytes(B, A, cap);
    const size_t s_pred = up256(pb), s_anc = up256((size_t)A * 4 * 8), s_dets = up256(db), s_cnt = up256((size_t)B * 2 * 4);
    unsigned char* base = nullptr;
    rc = arena_get(s_pred + s_anc + s_dets + s_cnt + up256(sb), &base); if (rc) return rc;
    float* d_pred = reinterpret_cast<float*>(base); double* d_anc = reinterpret_cast<double*>(base + s_pred);
    int* d_dets = reinterpret_cast<int*>(base + s_pred + s_anc); int* d_cnt = reinterpret_cast<int*>(base + s_pred + s_anc + s_dets);
    void* scratch = base + s_pred + s_anc + s_dets + s_cnt;
    // page-locked host buffers (ssdb_pinned_alloc, torch pin_memory) make these DMA transfers at the full PCIe rate
    SSDB_CUDA(cudaMemcpyAsync(d_pred, pred, pb, cudaMemcpyHostToDevice, nullptr));
    SSDB_CUDA(cudaMemcpyAsync(d_anc, anchors, (size_t)A * 4 * 8, cudaMemcpyHostToDevice, nullptr));
    SSDB_CUDA(cudaMemsetAsync(d_dets, 0, db, nullptr));                  // rows beyond an image's count read as zeros
    rc = decode_nms_launch(d_pred, B, A, C, d_anc, conf_thr, cap, iou_thr, d_dets, d_cnt, scratch, sb, nullptr); if (rc) return rc;
    SSDB_CUDA(cudaMemcpyAsync(dets_out, d_dets, db, cudaMemcpyDeviceToHost, nullptr));
    SSDB_CUDA(cudaMemcpyAsync(counts_out, d_cnt, (size_t)B * 2 * 4, cudaMemcpyDeviceToHost, nullptr));
    SSDB_CUDA(cudaStreamSynchronize(nullptr));
    return SSDB_OK;
}

int ssdb_nms_host(const int* boxes, const int* labelid, const float* conf, int n, int nclass, double iou_thr, int* keep_out, int* count_out) {
    int rc = ssdb_device_ok(); if (rc) return rc;
    return nms_only_host(boxes, labelid, conf, n, nclass, iou_thr, keep_out, count_out);
}

// grow-only workspace of the stateless loss entry points (one caller thread per process, like the reference)
static int loss_ws(int B, int A, void** ws_out) {
    static PerDevice<void*> ws_pd; static PerDevice<size_t> cap_pd;
    void*& ws = ws_pd.get(); size_t& cap = cap_pd.get();
    const size_t need = multibox_loss_ws_bytes(B, A);
    if (need > cap) {
        if (ws) { SSDB_CUDA(cudaDeviceSynchronize()); cudaFree(ws); ws = nullptr; cap = 0; }
        SSDB_CUDA(cudaMalloc(&ws, need)); cap = need;
        SSDB_CUDA(cudaMemset(ws, 0, need));
    }
    *ws_out = ws;
    return SSDB_OK;
}

int ssdb_multibox_loss(const float* output_dev, const float* labels_dev, int B, int A, int C, float grad_scale, float* losses_out_dev,
                       float* grad_out_dev, float* result_out_dev, void* stream) {
    SSDB_REQUIRE(output_dev && labels_dev && losses_out_dev, "bad arguments");
    void* ws; int rc = loss_ws(B, A, &ws); if (rc) return rc;
    return multibox_loss_launch(output_dev, labels_dev, nullptr, nullptr, 0, nullptr, B, A, C, grad_scale, losses_out_dev, grad_out_dev,
                                result_out_dev, nullptr, ws, (cudaStream_t)stream);
}

int ssdb_multibox_loss_gt(const float* output_dev, const double* gt_dev, const int* gt_count_dev, int B, int G, const double* anchors_prop_dev,
                          int A, int C, float grad_scale, float* losses_out_dev, float* grad_out_dev, float* result_out_dev,
                          int* match_out_dev, void* stream) {
    SSDB_REQUIRE(output_dev && gt_dev && gt_count_dev && anchors_prop_dev && losses_out_dev, "bad arguments");
    void* ws; int rc = loss_ws(B, A, &ws); if (rc) return rc;
    return multibox_loss_launch(output_dev, nullptr, gt_dev, gt_count_dev, G, anchors_prop_dev, B, A, C, grad_scale, losses_out_dev,
                                grad_out_dev, result_out_dev, match_out_dev, ws, (cudaStream_t)stream);
}

static ConvGeom make_geom(int B, int H, int W, int Cin, int Cout, int k, int stride, int dil, int pad_t, int pad_l, int Ho, int Wo) {
    ConvGeom g; g.B = B; g.H = H; g.W = W; g.Cin = Cin; g.Ho = Ho; g.Wo = Wo; g.Cout = Cout; g.k = k; g.stride = stride; g.dil = dil;
    g.pad_t = pad_t; g.pad_l = pad_l;
    return g;
}

// temporary device buffers of the layer test hooks, released (stream-ordered) when the hook returns
struct TmpBufs {
    cudaStream_t st; std::vector<void*> ptrs;
    explicit TmpBufs(cudaStream_t s) : st(s) {}
    float* get(size_t floats) {
        void* p = nullptr;
        if (cudaMallocAsync(&p, (floats ? floats : 1) * sizeof(float), st) != cudaSuccess) return nullptr;
        ptrs.push_back(p);
        return static_cast<float*>(p);
    }
    ~TmpBufs() { for (void* p : ptrs) cudaFreeAsync(p, st); }
};

// which kernel family and operand format a test-hook `impl` selects
static bool hook_mode(int impl, bool tc_supported, bool* tc, int* fmt) {
    *tc = impl == SSDB_CONV_TC || impl == SSDB_CONV_TC_SPLIT || (impl == SSDB_CONV_AUTO && tc_supported);
    *fmt = (*tc && impl != SSDB_CONV_TC) ? ACT_S32 : ACT_F32;
    return !*tc || tc_supported;
}

// the engine's operand preparation: tf32 rounding (ACT_F32) or the bf16 split (ACT_S32; n must be a multiple of 32)
static int hook_prepare(const float* src, float* dst, long long n, int fmt, cudaStream_t st) {
    return fmt == ACT_S32 ? split_copy(src, dst, n, st) : round_tf32_copy(src, dst, n, st);
}

int ssdb_op_conv_fprop(int impl, const float* x, const float* w_hwio, const float* bias, int B, int H, int W, int Cin, int Cout, int k,
                       int stride, int dil, int pad_t, int pad_l, int Ho, int Wo, int relu, float* y, void* stream) {
    SSDB_REQUIRE(x && w_hwio && y, "bad arguments");
    ConvGeom g = make_geom(B, H, W, Cin, Cout, k, stride, dil, pad_t, pad_l, Ho, Wo);
    ConvEpilogue ep; ep.bias = bias; ep.relu = relu;
    cudaStream_t st = (cudaStream_t)stream;
    bool tc; int fmt;
    SSDB_REQUIRE(hook_mode(impl, conv_tc_supported_fprop(g), &tc, &fmt), "shape not supported by the tcgen05 kernel");
    if (!tc) return conv_simt_fprop(g, x, w_hwio, ACT_F32, ep, y, st);
    SSDB_REQUIRE(fmt == ACT_F32 || Cout % 32 == 0, "split mode needs Cout % 32 == 0");
    int bn = Cout > 256 ? 256 : (Cout + 15) / 16 * 16;
    int cout_pad = (Cout + bn - 1) / bn * bn;
    long long nx = (long long)B * H * W * Cin, ny = (long long)B * Ho * Wo * Cout;
    TmpBufs tmp(st);
    float* wt = tmp.get((size_t)k * k * cout_pad * Cin); float* xr = tmp.get((size_t)nx);
    float* ys = fmt == ACT_S32 ? tmp.get((size_t)ny) : y;
    SSDB_REQUIRE(wt && xr && ys, "out of device memory");
    int rc = pack_filter_t(w_hwio, k * k, Cin, Cout, cout_pad, fmt, wt, st);
    if (!rc) rc = hook_prepare(x, xr, nx, fmt, st);
    if (!rc) rc = conv_tc_fprop(g, xr, wt, cout_pad, fmt, ep, ys, st);
    if (!rc && fmt == ACT_S32) rc = unsplit_copy(ys, y, ny, st);
    return rc;
}

int ssdb_op_conv_dgrad(int impl, const float* dz, const float* w_hwio, const float* mask_x, int B, int H, int W, int Cin, int Cout, int k,
                       int stride, int dil, int pad_t, int pad_l, int Ho, int Wo, int beta, float* dx, void* stream) {
    SSDB_REQUIRE(dz && w_hwio && dx, "bad arguments");
    ConvGeom g = make_geom(B, H, W, Cin, Cout, k, stride, dil, pad_t, pad_l, Ho, Wo);
    cudaStream_t st = (cudaStream_t)stream;
    bool tc; int fmt;
    SSDB_REQUIRE(hook_mode(impl, conv_tc_supported_dgrad(g), &tc, &fmt), "shape not supported by the tcgen05 kernel");
    if (!tc) return conv_simt_dgrad(g, dz, w_hwio, ACT_F32, mask_x, beta, 0, dx, st);
    long long nz = (long long)B * Ho * Wo * Cout, nw = (long long)k * k * Cin * Cout, nx = (long long)B * H * W * Cin;
    TmpBufs tmp(st);
    float* zr = tmp.get((size_t)nz); float* wr = tmp.get((size_t)nw);
    SSDB_REQUIRE(zr && wr, "out of device memory");
    int rc = hook_prepare(dz, zr, nz, fmt, st);
    if (!rc) rc = hook_prepare(w_hwio, wr, nw, fmt, st);
    if (rc) return rc;
    if (fmt == ACT_F32) return conv_tc_dgrad(g, zr, wr, fmt, mask_x, beta, 0, dx, st);
    float* ms = mask_x ? tmp.get((size_t)nx) : nullptr; float* ds = tmp.get((size_t)nx);
    SSDB_REQUIRE(ds && (ms || !mask_x), "out of device memory");
    if (mask_x) { rc = split_copy(mask_x, ms, nx, st); if (rc) return rc; }
    if (beta) { rc = split_copy(dx, ds, nx, st); if (rc) return rc; }
    rc = conv_tc_dgrad(g, zr, wr, fmt, ms, beta, 0, ds, st);
    if (!rc) rc = unsplit_copy(ds, dx, nx, st);
    return rc;
}

int ssdb_op_conv_wgrad(int impl, const float* x, const float* dz, int B, int H, int W, int Cin, int Cout, int k, int stride, int dil,
                       int pad_t, int pad_l, int Ho, int Wo, float* dw, float* db, void* stream) {
    SSDB_REQUIRE(x && dz && dw, "bad arguments");
    ConvGeom g = make_geom(B, H, W, Cin, Cout, k, stride, dil, pad_t, pad_l, Ho, Wo);
    cudaStream_t st = (cudaStream_t)stream;
    bool tc = impl == SSDB_CONV_TC || impl == SSDB_CONV_TC_SPLIT;
    int fmt = impl == SSDB_CONV_TC ? ACT_F32 : ACT_S32;
    if (impl == SSDB_CONV_AUTO) { tc = conv_tc_supported_wgrad(g, ACT_S32); fmt = tc ? ACT_S32 : ACT_F32; }
    if (!tc) fmt = ACT_F32;
    size_t ws = tc ? conv_tc_wgrad_ws(g, fmt) : conv_simt_wgrad_ws(g);
    if (ws < (size_t)1184 * Cout) ws = (size_t)1184 * Cout;
    TmpBufs tmp(st);
    float* partial = tmp.get(ws);
    SSDB_REQUIRE(partial, "out of device memory");
    ConvEpilogue ep;
    int rc = SSDB_OK;
    if (tc) {
        SSDB_REQUIRE(conv_tc_supported_wgrad(g, fmt), "shape not supported by the tcgen05 wgrad kernel");
        long long nx = (long long)B * H * W * Cin, nz = (long long)B * Ho * Wo * Cout;
        float* xr = tmp.get((size_t)nx); float* zr = tmp.get((size_t)nz);
        SSDB_REQUIRE(xr && zr, "out of device memory");
        rc = hook_prepare(x, xr, nx, fmt, st);
        if (!rc) rc = hook_prepare(dz, zr, nz, fmt, st);
        if (!rc) rc = conv_tc_wgrad(g, xr, zr, fmt, dw, db, partial, st);      // db from the all-ones slot (of the prepared dz)
    } else {
        rc = conv_simt_wgrad(g, x, dz, ACT_F32, ep, dw, partial, st);
        if (!rc && db) rc = bias_grad(dz, ACT_F32, (long long)B * Ho * Wo, Cout, db, partial, st);
    }
    return rc;
}

// deterministic pseudo-random fill: value in [-1, 1), a fraction `zero_frac` of the elements exactly 0 (ReLU-like sparsity)
__global__ void bench_fill_kernel(float* p, long long n, unsigned seed, float zero_frac, float scale) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned h = (unsigned)i * 2654435761u ^ seed;
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    float u = (float)(h >> 8) * (1.0f / 16777216.0f);
    unsigned h2 = h * 0x9e3779b9u; h2 ^= h2 >> 15;
    float z = (float)(h2 >> 8) * (1.0f / 16777216.0f);
    p[i] = z < zero_frac ? 0.f : (2.f * u - 1.f) * scale;
}

int ssdb_op_conv_bench(int kind, int impl, int B, int H, int W, int Cin, int Cout, int k, int stride, int dil, int pad_t, int pad_l,
                       int Ho, int Wo, int with_mask, int beta, int iters, float* ms_out) {
    SSDB_REQUIRE(kind >= 0 && kind <= 2 && iters >= 1 && ms_out, "bad arguments");
    int rc = ssdb_device_ok(); if (rc) return rc;
    ConvGeom g = make_geom(B, H, W, Cin, Cout, k, stride, dil, pad_t, pad_l, Ho, Wo);
    const bool tc = impl != SSDB_CONV_SIMT;
    const int fmt = impl == SSDB_CONV_TC ? ACT_F32 : (impl == SSDB_CONV_SIMT ? ACT_F32 : ACT_S32);
    cudaStream_t st = nullptr;
    SSDB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const long long nx = (long long)B * H * W * Cin, ny = (long long)B * Ho * Wo * Cout, nw = (long long)k * k * Cin * Cout;
    const int bn = Cout > 256 ? 256 : (Cout + 15) / 16 * 16;
    const int cout_pad = (Cout + bn - 1) / bn * bn;
    size_t ws = tc ? conv_tc_wgrad_ws(g, fmt) : conv_simt_wgrad_ws(g);
    if (ws < (size_t)1184 * Cout) ws = (size_t)1184 * Cout;
    float *x = nullptr, *y = nullptr, *w = nullptr, *tmp = nullptr, *wt = nullptr, *wr = nullptr, *dw = nullptr, *db = nullptr, *partial = nullptr, *bias = nullptr, *dx = nullptr;
    const long long nmax = nx > ny ? nx : ny;
#define BALLOC(ptr, count) SSDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&(ptr)), (size_t)(count) * sizeof(float)))
    BALLOC(x, nx); BALLOC(y, ny); BALLOC(w, nw); BALLOC(tmp, nmax > nw ? nmax : nw); BALLOC(wt, (size_t)k * k * cout_pad * Cin); BALLOC(wr, nw);
    BALLOC(dw, nw); BALLOC(db, Cout); BALLOC(partial, ws); BALLOC(bias, Cout); BALLOC(dx, nx);
#undef BALLOC
    auto fill = [&](float* p, long long n, unsigned seed, float zf, float sc) {
        bench_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n, seed, zf, sc);
    };
    auto prep = [&](float* dst, long long n, unsigned seed, float zf, float sc) -> int {   // random data in the engine's operand format
        fill(tmp, n, seed, zf, sc);
        if (!tc) { SSDB_CUDA(cudaMemcpyAsync(dst, tmp, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st)); return SSDB_OK; }
        return hook_prepare(tmp, dst, n, fmt, st);
    };
    rc = prep(x, nx, 1u, 0.5f, 1.f);                                  // activations: half zeros (post-ReLU)
    if (!rc) rc = prep(y, ny, 2u, 0.f, 1.f);                          // dz: dense
    if (!rc) rc = prep(dx, nx, 5u, 0.f, 1.f);
    fill(w, nw, 3u, 0.f, 0.05f); fill(bias, Cout, 4u, 0.f, 0.1f);
    if (!rc && tc) rc = pack_filter_t(w, k * k, Cin, Cout, cout_pad, fmt, wt, st);
    if (!rc && tc) rc = hook_prepare(w, wr, nw, fmt, st);
    ConvEpilogue ep; ep.bias = bias; ep.relu = 1; ep.round_tf32 = (tc && fmt == ACT_F32) ? 1 : 0;
    auto once = [&]() -> int {
        if (kind == 0) return tc ? conv_tc_fprop(g, x, wt, cout_pad, fmt, ep, y, st) : conv_simt_fprop(g, x, w, ACT_F32, ep, y, st);
        if (kind == 1) return tc ? conv_tc_dgrad(g, y, wr, fmt, with_mask ? x : nullptr, beta, fmt == ACT_F32 ? 1 : 0, dx, st)
                                 : conv_simt_dgrad(g, y, w, ACT_F32, with_mask ? x : nullptr, beta, 0, dx, st);
        if (tc) return conv_tc_wgrad(g, x, y, fmt, dw, db, partial, st);
        ConvEpilogue e2;
        return conv_simt_wgrad(g, x, y, ACT_F32, e2, dw, partial, st);
    };
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2 && !rc; ++i) rc = once();
    if (!rc) {
        cudaEventRecord(e0, st);
        for (int i = 0; i < iters && !rc; ++i) rc = once();
        cudaEventRecord(e1, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error("conv bench failed: %s", cudaGetErrorString(e)); rc = SSDB_ECUDA; }
        else { float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); *ms_out = ms / iters; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaStreamSynchronize(st);
    for (float* p : {x, y, w, tmp, wt, wr, dw, db, partial, bias, dx}) cudaFree(p);
    cudaStreamDestroy(st);
    return rc;
}

int ssdb_pinned_alloc(long long bytes, void** host_ptr_out) {
    SSDB_REQUIRE(bytes > 0 && host_ptr_out, "bad arguments");
    int rc = ssdb_device_ok(); if (rc) return rc;
    SSDB_CUDA(cudaHostAlloc(host_ptr_out, (size_t)bytes, cudaHostAllocDefault));
    return SSDB_OK;
}

int ssdb_pinned_free(void* host_ptr) {
    if (host_ptr) SSDB_CUDA(cudaFreeHost(host_ptr));
    return SSDB_OK;
}

unsigned int ssdb_crc32c(unsigned int crc, const void* data_host, size_t bytes) {
    static uint32_t table[8][256];
    static std::once_flag once;
    std::call_once(once, [] {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1) ? 0x82f63b78u : 0u);
            table[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int t = 1; t < 8; ++t) table[t][i] = (table[t - 1][i] >> 8) ^ table[0][table[t - 1][i] & 0xff];
    });
    const unsigned char* p = static_cast<const unsigned char*>(data_host);
    uint32_t c = crc ^ 0xffffffffu;
    while (bytes >= 8) {                                     // slicing-by-8
        uint32_t lo, hi;
        memcpy(&lo, p, 4); memcpy(&hi, p + 4, 4);
        lo ^= c;
        c = table[7][lo & 0xff] ^ table[6][(lo >> 8) & 0xff] ^ table[5][(lo >> 16) & 0xff] ^ table[4][lo >> 24] ^
            table[3][hi & 0xff] ^ table[2][(hi >> 8) & 0xff] ^ table[1][(hi >> 16) & 0xff] ^ table[0][hi >> 24];
        p += 8; bytes -= 8;
    }
    while (bytes--) c = table[0][(c ^ *p++) & 0xff] ^ (c >> 8);
    return c ^ 0xffffffffu;
}

int ssdb_profile_step(ssdb_net* n, const float* images_dev, const float* labels_dev, int B, char (*names_out)[32], float* ms_out,
                      int* launches_out, int cap) {
    SSDB_REQUIRE(n && images_dev && labels_dev && names_out && ms_out && launches_out && cap > 0, "bad arguments");
    SSDB_REQUIRE(!n->inference, "training entry point called on an inference handle");
    cudaStream_t st = n->own_stream;
    SSDB_CUDA(cudaStreamSynchronize(st));
    n->prof = true; n->prof_entries.clear();
    int rc = ssdb_train_step(n, images_dev, labels_dev, nullptr, nullptr, 0, B, 0.00075f, 0.9f, 0.0005f, 1.0f, 1, nullptr, nullptr, st);
    n->prof = false;
    cudaError_t e = cudaStreamSynchronize(st);
    int count = 0;
    for (auto& pe : n->prof_entries) {
        float ms = 0.f;
        if (!rc && e == cudaSuccess) cudaEventElapsedTime(&ms, pe.a, pe.b);
        if (count < cap) {
            strncpy(names_out[count], pe.label.c_str(), 31); names_out[count][31] = 0;
            ms_out[count] = ms; launches_out[count] = (int)pe.launches; ++count;
        }
        cudaEventDestroy(pe.a); cudaEventDestroy(pe.b);
    }
    n->prof_entries.clear();
    if (rc) return rc;
    if (e != cudaSuccess) { set_error("profile step failed: %s", cudaGetErrorString(e)); return SSDB_ECUDA; }
    return count;
}

}  // extern "C"
