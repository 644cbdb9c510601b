"""GPU parity of the box path: anchor matching / label creation and fused decode + NMS,
bit-exact against the committed reference fixtures and the NumPy oracle."""
import os

import numpy as np
import pytest

import box_oracle as bo
import ssdb
import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(params=['v2', 'v1'], autouse=True)
def nms_impl(request, monkeypatch):
    """Decode + NMS runs as scan kernel + per-image kernel (default) and as the single per-image kernel."""
    monkeypatch.setenv('SSDB_NMS', request.param)
    return request.param


def _labels_equal(got, want):
    """classes / background flag / linear offsets exact; log offsets within 1 float32 ulp
    (device log() vs libm log() may differ in the last float64 bit before rounding)."""
    assert np.array_equal(got[..., :21], want[..., :21])
    assert np.array_equal(got[..., 21:23], want[..., 21:23])
    a, b = got[..., 23:], want[..., 23:]
    ulp = np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))
    assert ulp.max() <= 1
    return int((ulp > 0).sum())


@pytest.mark.parametrize('preset,B,maxg', [('vgg300', 8, 16), ('vgg512', 3, 16), ('vgg300', 4, 3)])
def test_match_anchors_vs_oracle(preset, B, maxg):
    anc = bo.anchors(preset)
    aabs = bo.anchors_abs(anc)
    gts = [synth.gt_boxes(300 + i, max_boxes=maxg) for i in range(B)]
    gt, cnt = synth.pack_gt(gts, maxg)
    match, labels = ssdb.match_anchors_host(gt, cnt, anc, 20)
    for i in range(B):
        vec, m = bo.make_labels(gts[i], anc, aabs, 20)
        assert np.array_equal(match[i], m)
        _labels_equal(labels[i], vec)


def test_match_anchors_vs_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, 'match.npz'))
    keys = sorted(k[:-3] for k in g.files if k.endswith('_gt'))
    for key in keys:
        preset = key.split('_')[0]
        anc = bo.anchors(preset)
        gt = g[key + '_gt']
        arr, cnt = synth.pack_gt([gt], max(len(gt), 1))
        match, labels = ssdb.match_anchors_host(arr, cnt, anc, 20)
        pos = np.nonzero(match[0] >= 0)[0]
        assert np.array_equal(pos, g[key + '_pos']), key
        _labels_equal(labels[0][pos], g[key + '_rows'])
        neg = np.setdiff1d(np.arange(anc.shape[0]), pos)
        assert np.all(labels[0][neg, 20] == 1) and np.all(labels[0][neg, :20] == 0) and np.all(labels[0][neg, 21:] == 0)


def test_match_empty_and_unmatched():
    anc = bo.anchors('vgg300')
    gt = np.zeros((2, 4, 5)); cnt = np.array([0, 1], np.int32)
    gt[1, 0] = (3, 0.5, 0.5, 0.004, 0.004)          # too small to reach IoU 0.5 with any anchor
    match, labels = ssdb.match_anchors_host(gt, cnt, anc, 20)
    assert np.all(match == -1) and np.all(labels[..., 20] == 1)


def _rows_from_dets(dets, counts, i):
    n = counts[i, 0]
    d = dets[i, :n].astype(np.int64)
    rows = np.zeros((n, 7), np.int64)
    rows[:, 0] = d[:, 0] & 0xffffffff
    rows[:, 1:7] = d[:, 1:7]
    return rows


@pytest.mark.parametrize('preset,dist,thr,cap', [
    ('vgg300', 'U', 0.01, 200), ('vgg300', 'C', 0.01, 200), ('vgg300', 'C', 0.5, 200), ('vgg300', 'C', 0.3, None),
    ('vgg300', 'U', 0.97, None), ('vgg300', 'C', 0.01, 50), ('vgg512', 'C', 0.01, 200), ('vgg512', 'U', 0.01, 200),
    ('vgg300', 'C', 0.01, 256), ('vgg300', 'C', 0.01, 300), ('vgg300', 'C', 0.01, 7), ('vgg300', 'U', 0.01, 33)])
def test_decode_nms_vs_oracle(preset, dist, thr, cap):
    anc = bo.anchors(preset)
    B = 4 if preset == 'vgg300' else 2
    preds = np.stack([synth.pred_uniform(20 + i, anc.shape[0]) if dist == 'U' else synth.pred_clustered(20 + i, anc)
                      for i in range(B)])
    dets, counts = ssdb.decode_nms_host(preds, anc, thr, cap, 0.45)
    for i in range(B):
        rows, cand = bo.detect(preds[i], anc, thr, cap)
        assert counts[i, 1] == cand['idx'].shape[0]
        got = _rows_from_dets(dets, counts, i)
        assert got.shape == rows.shape and np.array_equal(got, rows), (preset, dist, thr, cap, i)


def test_decode_nms_vs_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, 'detect.npz'))
    cache = {}
    for key in [str(k) for k in g['cases']]:
        parts = key.split('_')
        preset, dist, i, thr = parts[0], parts[1], int(parts[2]), float(parts[3])
        cap = None if parts[4] == 'None' else int(parts[4])
        if preset not in cache:
            cache[preset] = bo.anchors(preset)
        anc = cache[preset]
        pred = synth.pred_uniform(i, anc.shape[0]) if dist == 'U' else synth.pred_clustered(i, anc)
        dets, counts = ssdb.decode_nms_host(pred[None], anc, thr, cap, 0.45)
        got = _rows_from_dets(dets, counts, 0)
        ref = g[key + '_rows']
        assert got.shape[0] == ref.shape[0], key
        assert np.array_equal(got[:, 0].astype(np.uint32).view(np.float32), g[key + '_conf32']), key
        assert np.array_equal(got[:, 1], ref[:, 1].astype(np.int64)), key
        w = (got[:, 3] - got[:, 2]).astype(np.float64); h = (got[:, 5] - got[:, 4]).astype(np.float64)
        assert np.array_equal((got[:, 2] + w / 2) / 1000, ref[:, 2]) and np.array_equal((got[:, 4] + h / 2) / 1000, ref[:, 3]), key
        assert np.array_equal(w / 1000, ref[:, 4]) and np.array_equal(h / 1000, ref[:, 5]), key


def test_decode_nms_nothing_above_threshold():
    anc = bo.anchors('vgg300')
    pred = synth.pred_uniform(1, anc.shape[0])[None]
    dets, counts = ssdb.decode_nms_host(pred, anc, 2.0, 200, 0.45)
    assert counts[0, 0] == 0 and counts[0, 1] == 0


def test_decode_nms_batch_128_property():
    """BASELINE.json configs[4] size (128 images): a sample of images is compared with the oracle row by row; every
    image satisfies the size-independent properties (kept <= candidates <= cap, confidence-descending inside a class)."""
    anc = bo.anchors('vgg300')
    preds = np.stack([synth.pred_clustered(3000 + i, anc) for i in range(128)])
    dets, counts = ssdb.decode_nms_host(preds, anc, 0.01, 200, 0.45)
    for i in range(0, 128, 9):
        rows, cand = bo.detect(preds[i], anc, 0.01, 200)
        got = _rows_from_dets(dets, counts, i)
        assert got.shape == rows.shape and np.array_equal(got, rows), i
    for i in range(128):
        n = counts[i, 0]
        assert 0 < n <= counts[i, 1] <= 200
        conf = dets[i, :n, 0].astype(np.uint32).view(np.float32)
        cls = dets[i, :n, 1]
        for c in np.unique(cls):
            assert np.all(np.diff(conf[cls == c]) <= 0), 'confidence order inside a class'
