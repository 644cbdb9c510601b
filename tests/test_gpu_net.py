"""GPU parity of the whole engine against the torch restatement of the reference graph:
raw logits, offsets, softmax result, losses, every parameter gradient and one Momentum step.

north_star's bar: class logits / box offsets within 1e-3 (relative, fp32).  The engine's DEFAULT mode
('split': tensor cores with split bf16 operands) must meet it with a wide margin on both presets; the
fp32 CUDA-core mode ('simt') is the exact cross-check; 'tf32' is the comparison mode that does NOT meet the
bar (10-bit significands: ~1.5e-3) and is only checked against its own known floor."""
import json
import os

import numpy as np
import pytest
import torch

import box_oracle as bo
import net_oracle as no
import ssdb
import synth

pytestmark = pytest.mark.gpu

NORTH_STAR_TOL = 1e-3
# measured headroom of the product mode (B200: 1.3e-4 / 1.4e-4); a regression to tf32-grade arithmetic would trip this long before 1e-3
SPLIT_EXPECT = 3e-4


def _load(net, P):
    names = dict(net.tensors())
    assert set(names) == set(P), set(names) ^ set(P)
    for k, shape in names.items():
        assert tuple(P[k].shape) == shape, (k, shape, tuple(P[k].shape))
        net.set_tensor(k, P[k].detach().numpy().astype(np.float32))


def _relmax(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _make(preset, B, mode):
    if mode == 'default':
        os.environ.pop('SSDB_CONV', None)           # the product path: whatever the library picks with no override
    else:
        os.environ['SSDB_CONV'] = mode
    try:
        return ssdb.Net(preset, 20, max_batch=B)
    finally:
        os.environ.pop('SSDB_CONV', None)


@pytest.mark.parametrize('mode', ['default', 'simt', 'tf32'])
@pytest.mark.parametrize('preset,B', [('vgg300', 2), ('vgg512', 1)])
def test_forward_and_train_step(preset, B, mode):
    net = _make(preset, B, mode)
    side = bo.PRESETS[preset]['image']
    P = no.init_params(preset, dtype=torch.float64)
    _load(net, P)
    anc = bo.anchors(preset); aabs = bo.anchors_abs(anc)
    x = synth.images(0, B, side)
    labels = np.stack([bo.make_labels(synth.gt_boxes(i), anc, aabs, 20)[0] for i in range(B)])
    # ---------------- forward: raw logits, offsets, softmax
    res = net.forward_host(x).copy()
    raw = net.read_output(B)
    out = no.forward(P, torch.tensor(x), preset).numpy()
    ref = no.result_from_output(torch.tensor(out)).numpy()
    report = {'preset': preset, 'mode': mode, 'B': B}
    report['logits_rel'] = _relmax(raw[..., :21], out[..., :21])          # relative to max |logit|
    report['locator_rel'] = _relmax(raw[..., 21:], out[..., 21:])
    report['locator_rel_rms'] = float(np.sqrt(((raw[..., 21:] - out[..., 21:]) ** 2).mean() / (out[..., 21:] ** 2).mean()))
    report['softmax_abs'] = float(np.abs(res[..., :21] - ref[..., :21]).max())
    report['argmax_agree'] = float((res[..., :21].argmax(-1) == ref[..., :21].argmax(-1)).mean())
    assert np.array_equal(res[..., 21:], raw[..., 21:]), 'result offsets are not the raw head offsets'
    print('PARITY-FWD', json.dumps(report))
    if mode == 'tf32':
        # comparison mode: sits on the error floor of tf32 operands (profiles/r1_tf32_floor_*.json: 1.3e-3 / 1.6e-3 max-norm)
        assert report['logits_rel'] < 4e-3 and report['locator_rel'] < 4e-3, report
        assert report['argmax_agree'] > 0.99
    else:
        assert report['logits_rel'] < NORTH_STAR_TOL, report
        assert report['locator_rel'] < NORTH_STAR_TOL, report
        assert report['logits_rel'] < SPLIT_EXPECT and report['locator_rel'] < SPLIT_EXPECT, report
        # |logit| ~ 1e3 on this synthetic input: 2e-5 relative on a logit is 2e-2 absolute before the softmax
        assert report['argmax_agree'] > 0.9995, report
        assert report['softmax_abs'] < (3e-3 if mode == 'simt' else 5e-2), report
    # ---------------- one training step
    tol = 4e-3 if mode == 'tf32' else NORTH_STAR_TOL
    V = {k: torch.zeros_like(v) for k, v in P.items()}
    L, out0, grads = no.train_step(P, V, torch.tensor(x), torch.tensor(labels), preset, lr=0.00075, momentum=0.9, weight_decay=0.0005)
    res2, losses = net.train_step_host(x, labels, 0.00075, 0.9, 0.0005)
    for key, i in (('total', 0), ('localization', 1), ('confidence', 2), ('l2', 3)):
        assert abs(losses[i] - L[key]) <= tol * abs(L[key]) + 1e-6, (key, losses[i], L[key])
    # Gradients go through ReLU masks, max-pool arg-maxes and the hard-negative selection, so forward noise can flip
    # decisions; tests/test_gpu_grad.py removes the flips (margins) and checks 2e-3.  Here: the plain synthetic input.
    gtol = {'simt': 1e-2, 'default': 2e-2, 'tf32': 0.2}[mode]
    bad = []
    worst = (None, 0.0, 0.0)
    for k, shape in net.tensors():
        g = net.get_tensor(k, shape, ssdb.GRAD)
        want = grads[k].numpy()
        if k.endswith('/filter'):
            want = want - 0.0005 * (P[k].numpy() + 0.00075 * V[k].numpy())   # oracle grads include the L2 term of the pre-update weights
        e = _relmax(g, want)
        l2 = float(np.sqrt(((g - want) ** 2).sum() / max((want ** 2).sum(), 1e-60)))
        if e > worst[1]:
            worst = (k, e, l2)
        if e > gtol:
            bad.append((k, e))
    report['losses'] = [float(v) for v in losses]
    report['losses_ref'] = [L['total'], L['localization'], L['confidence'], L['l2']]
    report['worst_grad'] = [worst[0], worst[1], worst[2]]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(root, 'gpurun_out', 'net_parity_%s_%s.json' % (preset, mode)), 'w') as f:
        json.dump(report, f)
    print('PARITY', json.dumps(report))
    assert not bad, bad[:10]
    for k, shape in net.tensors():
        w = net.get_tensor(k, shape, ssdb.PARAM)
        ref_w = P[k].numpy()
        step = 0.00075 * np.abs(V[k].numpy()).max()            # size of the update that was applied
        assert np.abs(w - ref_w).max() <= gtol * step + 1e-7 * np.abs(ref_w).max(), k
    net.close()
