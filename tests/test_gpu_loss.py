"""GPU parity of the fused multibox loss (dense-label and fused-match variants) against the
float64 torch restatement of ssdvgg.py:380-580 and its autograd gradient."""
import numpy as np
import pytest
import torch

import box_oracle as bo
import net_oracle as no
import ssdb
import synth
from gpu_util import dev, ptr, rel_err

pytestmark = pytest.mark.gpu


def _case(preset, B, seed, maxg=8, scale=2.0):
    anc = bo.anchors(preset)
    aabs = bo.anchors_abs(anc)
    gts = [synth.gt_boxes(500 + seed * 100 + i, max_boxes=maxg) for i in range(B)]
    labels = np.stack([bo.make_labels(g, anc, aabs, 20)[0] for g in gts])
    rng = np.random.default_rng(seed)
    out = (rng.standard_normal((B, anc.shape[0], 25)) * scale).astype(np.float32)
    return anc, gts, labels, out


def _oracle(out, labels):
    o = torch.tensor(out, dtype=torch.float64, requires_grad=True)
    conf, loc = no.multibox_loss(o, torch.tensor(labels, dtype=torch.float64))
    (conf + loc).backward()
    res = no.result_from_output(o.detach()).numpy()
    return float(conf), float(loc), o.grad.numpy(), res


@pytest.mark.parametrize('preset,B', [('vgg300', 4), ('vgg512', 2)])
def test_dense_loss_matches_oracle(preset, B):
    anc, gts, labels, out = _case(preset, B, seed=1)
    conf, loc, grad, res = _oracle(out, labels)
    A = anc.shape[0]
    od, ld = dev(out), dev(labels)
    losses = torch.zeros(2, device='cuda'); g = torch.empty_like(od); r = torch.empty_like(od)
    ssdb.check(ssdb.lib().ssdb_multibox_loss(ptr(od), ptr(ld), B, A, 20, 1.0, ptr(losses), ptr(g), ptr(r), None))
    torch.cuda.synchronize()
    l = losses.cpu().numpy()
    assert abs(l[0] - conf) / conf < 2e-5 and abs(l[1] - loc) / loc < 2e-5, (l, conf, loc)
    assert rel_err(g.cpu().numpy(), grad) < 2e-5
    assert np.abs(r.cpu().numpy() - res).max() < 2e-6
    # the set of anchors that receive a confidence gradient must be identical (bit-exact selection)
    assert np.array_equal(np.abs(g.cpu().numpy()[..., :21]).sum(-1) > 0, np.abs(grad[..., :21]).sum(-1) > 0)


def test_loss_image_without_positives():
    anc, gts, labels, out = _case('vgg300', 3, seed=2)
    labels[1] = 0; labels[1, :, 20] = 1          # no positive anchors in image 1 -> both losses 0 for it
    conf, loc, grad, _ = _oracle(out, labels)
    od, ld = dev(out), dev(labels)
    losses = torch.zeros(2, device='cuda'); g = torch.empty_like(od)
    ssdb.check(ssdb.lib().ssdb_multibox_loss(ptr(od), ptr(ld), 3, anc.shape[0], 20, 1.0, ptr(losses), ptr(g), None, None))
    torch.cuda.synchronize()
    l = losses.cpu().numpy()
    assert abs(l[0] - conf) / conf < 2e-5 and abs(l[1] - loc) / loc < 2e-5
    assert np.all(g.cpu().numpy()[1] == 0)
    assert rel_err(g.cpu().numpy(), grad) < 2e-5


def test_loss_with_tied_negatives():
    # all-zero head output: every negative has the same CE = log(21); top_k keeps the lowest indices
    anc, gts, labels, out = _case('vgg300', 2, seed=3)
    out[:] = 0
    conf, loc, grad, _ = _oracle(out, labels)
    od, ld = dev(out), dev(labels)
    losses = torch.zeros(2, device='cuda'); g = torch.empty_like(od)
    ssdb.check(ssdb.lib().ssdb_multibox_loss(ptr(od), ptr(ld), 2, anc.shape[0], 20, 1.0, ptr(losses), ptr(g), None, None))
    torch.cuda.synchronize()
    l = losses.cpu().numpy()
    assert abs(l[0] - conf) / conf < 2e-5
    assert np.array_equal(np.abs(g.cpu().numpy()[..., :21]).sum(-1) > 0, np.abs(grad[..., :21]).sum(-1) > 0)


@pytest.mark.parametrize('preset,B', [('vgg300', 4), ('vgg512', 2)])
def test_fused_match_loss_equals_dense(preset, B):
    anc, gts, labels, out = _case(preset, B, seed=4, maxg=12)
    conf, loc, grad, _ = _oracle(out, labels)
    gt, cnt = synth.pack_gt(gts, 12)
    A = anc.shape[0]
    od, gd, cd, ad = dev(out), dev(gt), dev(cnt), dev(anc)
    losses = torch.zeros(2, device='cuda'); g = torch.empty_like(od)
    match = torch.empty((B, A), dtype=torch.int32, device='cuda')
    ssdb.check(ssdb.lib().ssdb_multibox_loss_gt(ptr(od), ptr(gd), ptr(cd), B, 12, ptr(ad), A, 20, 1.0, ptr(losses), ptr(g), None,
                                                ptr(match), None))
    torch.cuda.synchronize()
    l = losses.cpu().numpy()
    assert abs(l[0] - conf) / conf < 2e-5 and abs(l[1] - loc) / loc < 2e-5
    assert rel_err(g.cpu().numpy(), grad) < 2e-5
    aabs = bo.anchors_abs(anc)
    for i in range(B):
        assert np.array_equal(match[i].cpu().numpy(), bo.match_anchors(gts[i], anc, aabs))
