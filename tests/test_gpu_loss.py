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


@pytest.fixture(params=['v2', 'v1'], autouse=True)
def loss_impl(request, monkeypatch):
    """Every test runs against the tiled streaming kernels (default) and the one-CTA-per-image kernel."""
    monkeypatch.setenv('SSDB_LOSS', request.param)
    return request.param


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


def _run_dense(out, labels, C=20, offset_floats=0):
    """Dense-label loss through the C ABI; offset_floats > 0 shifts every device pointer off 16-byte alignment."""
    B, A, V = out.shape
    def shifted(a):
        buf = torch.zeros(a.size + offset_floats + 4, dtype=torch.float32, device='cuda')
        view = buf[offset_floats:offset_floats + a.size].view(a.shape)
        view.copy_(torch.from_numpy(a))
        return buf, view
    ob, od = shifted(out); lb, ld = shifted(labels)
    gb, g = shifted(np.zeros_like(out)); rb, r = shifted(np.zeros_like(out))
    losses = torch.zeros(2, device='cuda')
    ssdb.check(ssdb.lib().ssdb_multibox_loss(ptr(od), ptr(ld), B, A, C, 1.0, ptr(losses), ptr(g), ptr(r), None))
    torch.cuda.synchronize()
    return losses.cpu().numpy(), g.cpu().numpy(), r.cpu().numpy()


def test_streaming_kernels_equal_single_cta_kernel(monkeypatch, loss_impl):
    """v2 must pick exactly the same negatives and produce bit-identical gradient / result rows as v1."""
    if loss_impl == 'v1':
        pytest.skip('comparison runs once')
    anc, gts, labels, out = _case('vgg300', 5, seed=7)
    monkeypatch.setenv('SSDB_LOSS', 'v2')
    l2, g2, r2 = _run_dense(out, labels)
    monkeypatch.setenv('SSDB_LOSS', 'v1')
    l1, g1, r1 = _run_dense(out, labels)
    assert np.array_equal(g1, g2) and np.array_equal(r1, r2)
    assert np.allclose(l1, l2, rtol=2e-6)


@pytest.mark.parametrize('C,A,offset', [(20, 8732, 1), (5, 1003, 0), (7, 777, 3), (20, 300, 0)])
def test_generic_row_width_and_unaligned_buffers(C, A, offset):
    """Row widths other than 25, anchor counts that are not a multiple of the tile, and buffers that are not
    16-byte aligned (the TMA bulk path must step aside for plain loads)."""
    rng = np.random.default_rng(C * 1000 + A)
    B, V = 3, C + 5
    out = (rng.standard_normal((B, A, V)) * 2).astype(np.float32)
    labels = np.zeros((B, A, V), np.float32)
    cls = rng.integers(0, C, size=(B, A))
    posm = rng.random((B, A)) < 0.03
    labels[..., C] = 1
    bi, ai = np.nonzero(posm)
    labels[bi, ai, C] = 0
    labels[bi, ai, cls[bi, ai]] = 1
    labels[bi, ai, C + 1:] = rng.standard_normal((bi.size, 4)).astype(np.float32)
    o = torch.tensor(out, dtype=torch.float64, requires_grad=True)
    conf, loc = no.multibox_loss(o, torch.tensor(labels, dtype=torch.float64), num_classes=C)
    (conf + loc).backward()
    res = no.result_from_output(o.detach(), num_classes=C).numpy()
    l, g, r = _run_dense(out, labels, C=C, offset_floats=offset)
    assert abs(l[0] - float(conf)) / float(conf) < 2e-5 and abs(l[1] - float(loc)) / float(loc) < 2e-5
    assert rel_err(g, o.grad.numpy()) < 2e-5
    assert np.abs(r - res).max() < 2e-6
    assert np.array_equal(np.abs(g[..., :C + 1]).sum(-1) > 0, np.abs(o.grad.numpy()[..., :C + 1]).sum(-1) > 0)


def test_engine_sized_batch_64():
    """BASELINE.json configs[1] size: 64 images x 8732 anchors; checks the losses and the selected set."""
    anc, gts, labels, out = _case('vgg300', 64, seed=9)
    conf, loc, grad, _ = _oracle(out, labels)
    l, g, r = _run_dense(out, labels)
    assert abs(l[0] - conf) / conf < 2e-5 and abs(l[1] - loc) / loc < 2e-5
    assert rel_err(g, grad) < 2e-5
    assert np.array_equal(np.abs(g[..., :21]).sum(-1) > 0, np.abs(grad[..., :21]).sum(-1) > 0)
