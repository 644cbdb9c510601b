"""CPU oracle for the dense half of the SSD hot path -- TEST INFRASTRUCTURE ONLY.

A torch-CPU restatement (float64 master, float32 on request) of the graph that
the reference's ``ssdvgg.py`` builds in TensorFlow 1.x: the VGG-16 trunk
(ssdvgg.py:190-207), the a-trous conv6/conv7 (ssdvgg.py:231-292), the extra
layers (ssdvgg.py:295-332), the L2 normalisation of conv4_3
(ssdvgg.py:80-84,335-337), the per-(map, box-type) 3x3 classifiers and the
output layout (ssdvgg.py:55-65,340-372), the multibox loss with 3:1 hard
negative mining (ssdvgg.py:68-71,375-580) and the Momentum update
(ssdvgg.py:585-588).  TF semantics (SAME/VALID padding, l2_normalize, l2_loss,
top_k, softmax-CE v2) follow SURVEY.md Appendix A.

Parity status: UNPINNED at the TensorFlow boundary.  TensorFlow 1.x is an
un-vendored, un-pinned third-party dependency of the reference (API use
brackets it to about 1.6 - 1.15), it is not installable here, the pretrained
VGG saved-model (ssdvgg.py:174) is absent, and the reference ships no test,
fixture or golden vector for this graph.  The restatement is therefore
anchored on the reference's call sites only; the one shipped known-answer --
the output row count 8732 / 24564 (ssdutils.py:48,61) -- is asserted.  The
input pre-processing that lives inside the third-party VGG graph is defined
explicitly here (``preprocess``) and mirrored by the CUDA engine.

Nothing in the product imports this file; tests, smoke() and bench.py's
CPU-baseline legs use it as the checker / the host-core baseline.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from box_oracle import PRESETS

VGG_MEAN_RGB = (123.68, 116.779, 103.939)


def conv_specs(preset_name, num_classes=20):
    """Ordered conv layer table: dict(name, k, cin, cout, stride, dilation, padding, relu, src).

    Trunk = VGG-16 configuration D; conv6/7, extras and classifiers per
    ssdvgg.py:231-372.  Classifiers are one layer per (map, box type) exactly
    as in the reference (scope 'classifiers/classifier{i}_{j}')."""
    maps = PRESETS[preset_name]['maps']
    L = []
    def add(name, k, cin, cout, stride=1, dil=1, padding='SAME', relu=True):
        L.append(dict(name=name, k=k, cin=cin, cout=cout, stride=stride, dilation=dil,
                      padding=padding, relu=relu))
    cfg = [('conv1', 2, 3, 64), ('conv2', 2, 64, 128), ('conv3', 3, 128, 256),
           ('conv4', 3, 256, 512), ('conv5', 3, 512, 512)]
    for blk, n, cin, cout in cfg:
        for i in range(n):
            add('%s_%d' % (blk, i + 1), 3, cin if i == 0 else cout, cout)
    add('mod_conv6', 3, 512, 1024, dil=6)
    add('mod_conv7', 1, 1024, 1024)
    seven = len(maps) >= 7
    add('conv8_1', 1, 1024, 256); add('conv8_2', 3, 256, 512, stride=2)
    add('conv9_1', 1, 512, 128); add('conv9_2', 3, 128, 256, stride=2)
    add('conv10_1', 1, 256, 128)
    add('conv10_2', 3, 128, 256, stride=2 if seven else 1, padding='SAME' if seven else 'VALID')
    add('conv11_1', 1, 256, 128); add('conv11_2', 3, 128, 256, padding='VALID')
    if seven:
        add('conv12_1', 1, 256, 128); add('conv12_2', 3, 128, 256, padding='VALID')
    src_c = [512, 1024, 512, 256, 256, 256, 256]
    for i, (fk, s, ratios) in enumerate(maps):
        for j in range(2 + len(ratios)):
            add('classifiers/classifier%d_%d' % (i, j), 3, src_c[i], num_classes + 5, relu=False)
    return L


def init_params(preset_name, num_classes=20, seed=7, dtype=torch.float64):
    """He-normal trunk + conv6/7 (stand-in for the pretrained VGG weights),
    Xavier-uniform extras / classifiers (ssdvgg.py:46,59), zero biases, scale 20."""
    g = np.random.default_rng(seed)
    P = {}
    for s in conv_specs(preset_name, num_classes):
        k, cin, cout = s['k'], s['cin'], s['cout']
        trunk = s['name'].startswith(('conv1_', 'conv2_', 'conv3_', 'conv4_', 'conv5_', 'mod_conv'))
        if trunk:
            w = g.normal(0, math.sqrt(2.0 / (k * k * cin)), (k, k, cin, cout))
            if s['name'] == 'conv1_1':
                w = w / (255.0 / math.sqrt(12.0))     # raw 0..255 pixels in, O(1) activations out
        else:
            lim = math.sqrt(6.0 / (k * k * cin + k * k * cout))
            w = g.uniform(-lim, lim, (k, k, cin, cout))
        P[s['name'] + '/filter'] = torch.tensor(w.astype(np.float32), dtype=dtype)
        P[s['name'] + '/biases'] = torch.zeros(cout, dtype=dtype)
    P['l2_norm_conv4_3/scale'] = torch.full((512,), 20.0, dtype=dtype)
    return P


def _same_pad(n, k_eff, stride):
    out = -(-n // stride)
    total = max((out - 1) * stride + k_eff - n, 0)
    return total // 2, total - total // 2


def conv_tf(x, w_hwio, b, stride=1, dilation=1, padding='SAME', relu=True, relu_mask=None):
    """tf.nn.conv2d / atrous_conv2d + bias_add (+ relu) on NCHW x with TF padding.
    relu_mask (0/1 tensor): use these ReLU decisions instead of the sign of this pass's own pre-activations."""
    k = w_hwio.shape[0]
    if padding == 'SAME':
        ke = (k - 1) * dilation + 1
        pt, pb = _same_pad(x.shape[2], ke, stride)
        pl, pr = _same_pad(x.shape[3], ke, stride)
        x = F.pad(x, (pl, pr, pt, pb))
    y = F.conv2d(x, w_hwio.permute(3, 2, 0, 1), b, stride=stride, dilation=dilation)
    if relu and relu_mask is not None:
        return y * relu_mask
    return F.relu(y) if relu else y


def max_pool_tf(x, k, stride, route_like=None):
    """tf.nn.max_pool padding='SAME' (pad value -inf).
    route_like: a tensor of x's shape whose window arg-maxes are used instead of x's own (x is gathered at them)."""
    pt, pb = _same_pad(x.shape[2], k, stride)
    pl, pr = _same_pad(x.shape[3], k, stride)
    x = F.pad(x, (pl, pr, pt, pb), value=float('-inf'))
    if route_like is None:
        return F.max_pool2d(x, k, stride)
    r = F.pad(route_like.to(x.dtype), (pl, pr, pt, pb), value=float('-inf'))
    _, idx = F.max_pool2d(r, k, stride, return_indices=True)
    return x.flatten(2).gather(2, idx.flatten(2)).view_as(idx)


def preprocess(x_nhwc):
    """Input stage of the third-party VGG graph, defined explicitly:
    split channels as R,G,B, subtract the ImageNet means, re-stack as B,G,R."""
    r, g, b = x_nhwc[..., 0], x_nhwc[..., 1], x_nhwc[..., 2]
    return torch.stack([b - VGG_MEAN_RGB[2], g - VGG_MEAN_RGB[1], r - VGG_MEAN_RGB[0]], dim=1)


def round_tf32(t):
    """Round to the tf32 grid (10 explicit mantissa bits), nearest with ties away from zero, like the engine's
    `cvt.rna.tf32.f32` at every producer (csrc/common.cuh tf32_rn); returned in the input dtype."""
    bits = t.to(torch.float32).contiguous().view(torch.int32)
    bits = (bits + 0x1000) & ~0x1fff
    return bits.view(torch.float32).to(t.dtype)


def forward(P, x_nhwc, preset_name, num_classes=20, taps=None, producer_round=None, decisions=None):
    """Returns output [B, A, C+5] (logits | offsets), pre-softmax (ssdvgg.py:365-366).
    `taps`, if a dict, receives the intermediate feature maps (NCHW) by name.
    `producer_round` (e.g. round_tf32) is applied where the engine rounds: to the pre-processed image, to every filter,
    to the output of every convolution that feeds another one (after bias + ReLU) and to the L2-norm output.  With exact
    products and wide accumulation this is the arithmetic model of the engine's tensor-core path: its distance from the
    plain float64 graph is the error floor of tf32 operands for this network, independent of any kernel, and the engine
    itself should match the model far more tightly than it matches float64.
    `decisions` (dict name -> NCHW activation tensor of ANOTHER forward pass of the same network, e.g. the engine's): every
    non-linear decision is taken from it -- the ReLU of conv <name> passes where decisions[<name>] > 0, the max-pool <name>
    routes to the window arg-max of its input there (decisions[<producer of its input>]).  The graph then is the same
    piecewise-linear function as the other pass, so gradients can be compared without decision flips."""
    maps = PRESETS[preset_name]['maps']
    spec = {s['name']: s for s in conv_specs(preset_name, num_classes)}
    rnd = producer_round if producer_round is not None else (lambda t: t)
    dec = decisions or {}
    def conv(name, x):
        s = spec[name]
        mask = (dec[name] > 0).to(x.dtype) if (name in dec and s['relu']) else None
        y = conv_tf(x, rnd(P[name + '/filter']), P[name + '/biases'], s['stride'], s['dilation'],
                    s['padding'], s['relu'], relu_mask=mask)
        if not name.startswith('classifiers/'):
            y = rnd(y)
        if taps is not None:
            taps[name] = y
        return y
    x = rnd(preprocess(x_nhwc.to(P['conv1_1/filter'].dtype)))
    x = conv('conv1_2', conv('conv1_1', x)); x = max_pool_tf(x, 2, 2, dec.get('conv1_2'))
    x = conv('conv2_2', conv('conv2_1', x)); x = max_pool_tf(x, 2, 2, dec.get('conv2_2'))
    x = conv('conv3_3', conv('conv3_2', conv('conv3_1', x))); x = max_pool_tf(x, 2, 2, dec.get('conv3_3'))
    c43 = conv('conv4_3', conv('conv4_2', conv('conv4_1', x))); x = max_pool_tf(c43, 2, 2, dec.get('conv4_3'))
    x = conv('conv5_3', conv('conv5_2', conv('conv5_1', x)))
    x = max_pool_tf(x, 3, 1, dec.get('conv5_3'))                   # mod_pool5
    c7 = conv('mod_conv7', conv('mod_conv6', x))
    c82 = conv('conv8_2', conv('conv8_1', c7))
    c92 = conv('conv9_2', conv('conv9_1', c82))
    c102 = conv('conv10_2', conv('conv10_1', c92))
    c112 = conv('conv11_2', conv('conv11_1', c102))
    # l2_normalization (ssdvgg.py:80-84): scale * x * rsqrt(max(sum x^2, 1e-12))
    ss = (c43 * c43).sum(dim=1, keepdim=True).clamp_min(1e-12)
    n43 = rnd(c43 * torch.rsqrt(ss) * P['l2_norm_conv4_3/scale'].view(1, -1, 1, 1))
    if taps is not None:
        taps['l2_norm_conv4_3'] = n43
    fmaps = [n43, c7, c82, c92, c102, c112]
    if len(maps) >= 7:
        y = conv('conv12_1', c112)
        y = F.pad(y, (0, 1, 0, 1))                                  # ssdvgg.py:327-329
        fmaps.append(conv('conv12_2', y))
    outs = []
    for i, (fk, s, ratios) in enumerate(maps):
        for j in range(2 + len(ratios)):
            y = conv('classifiers/classifier%d_%d' % (i, j), fmaps[i])    # [B, 25, H, W]
            outs.append(y.permute(0, 2, 3, 1).reshape(y.shape[0], fk * fk, -1))
    out = torch.cat(outs, dim=1)
    assert out.shape[1] == PRESETS[preset_name]['num_anchors']
    return out


def result_from_output(out, num_classes=20):
    """result = concat(softmax(logits), locator) (ssdvgg.py:368-372)."""
    nc = num_classes + 1
    return torch.cat([torch.softmax(out[..., :nc], dim=-1), out[..., nc:]], dim=-1)


def multibox_loss(out, labels, num_classes=20, selected=None):
    """(confidence_loss, localization_loss) restating ssdvgg.py:380-560.
    selected ([B, A] bool): use this set of hard negatives instead of this pass's own top_k (see forward(decisions=...))."""
    nc = num_classes + 1
    logits, loc = out[..., :nc], out[..., nc:]
    gt_cl, gt_loc = labels[..., :nc], labels[..., nc:]
    B, A = logits.shape[0], logits.shape[1]
    neg_num = (gt_cl[..., -1] != 0).sum(dim=1)
    pos_num = A - neg_num
    pos_mask = gt_cl[..., -1] == 0
    ce = torch.logsumexp(logits, dim=-1) - (gt_cl * logits).sum(dim=-1)
    zeros = torch.zeros_like(ce)
    pos_sum = torch.where(pos_mask, ce, zeros).sum(dim=-1)
    negatives = torch.where(~pos_mask, ce, zeros)
    # tf.nn.top_k(k=A): sorted descending, ties -> lower index first == stable descending sort
    top = torch.sort(negatives, dim=1, descending=True, stable=True)[0]
    kmax = torch.minimum(neg_num, 3 * pos_num).unsqueeze(1)
    keep = torch.arange(A).unsqueeze(0) < kmax
    neg_sum = torch.where(keep, top, torch.zeros_like(top)).sum(dim=-1)
    if selected is not None:
        neg_sum = torch.where(selected & ~pos_mask, ce, zeros).sum(dim=-1)
    safe = torch.where(pos_num == 0, torch.full_like(ce[:, 0], 10e-15), pos_num.to(ce.dtype))
    conf = torch.where(pos_num == 0, torch.zeros_like(pos_sum), (pos_sum + neg_sum) / safe)
    d = loc - gt_loc
    ad = d.abs()
    sl1 = torch.where(ad < 1, 0.5 * d * d, ad - 0.5).sum(dim=-1)
    loc_sum = torch.where(pos_mask, sl1, zeros).sum(dim=-1)
    locl = torch.where(pos_num == 0, torch.zeros_like(loc_sum), loc_sum / safe)
    return conf.mean(), locl.mean()


def l2_term(P):
    """sum over conv filters of sum(w^2)/2 -- tf.nn.l2_loss on filters only (App. A)."""
    return sum((v * v).sum() / 2 for k, v in P.items() if k.endswith('/filter'))


def losses(P, x_nhwc, labels, preset_name, num_classes=20, weight_decay=0.0005, decisions=None, selected=None):
    out = forward(P, x_nhwc, preset_name, num_classes, decisions=decisions)
    conf, loc = multibox_loss(out, labels.to(out.dtype), num_classes, selected=selected)
    l2 = weight_decay * l2_term(P)
    return dict(total=conf + loc + l2, confidence=conf, localization=loc, l2=l2), out


def train_step(P, V, x_nhwc, labels, preset_name, num_classes=20, lr=0.00075,
               momentum=0.9, weight_decay=0.0005, decisions=None, selected=None):
    """One MomentumOptimizer.minimize step (accum = mu*accum + g; var -= lr*accum).
    Mutates P and V in place; returns (losses, pre-update output, grads)."""
    for v in P.values():
        v.requires_grad_(True)
        v.grad = None
    L, out = losses(P, x_nhwc, labels, preset_name, num_classes, weight_decay, decisions, selected)
    L['total'].backward()
    grads = {k: v.grad.detach().clone() for k, v in P.items()}
    with torch.no_grad():
        for k, v in P.items():
            V[k].mul_(momentum).add_(grads[k])
            v.sub_(lr * V[k])
    for v in P.values():
        v.requires_grad_(False)
    return {k: float(v.detach()) for k, v in L.items()}, out.detach(), grads
