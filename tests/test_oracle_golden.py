"""CPU: the NumPy oracle (oracle/box_oracle.py) against the committed fixtures that
oracle/make_golden.py produced from the REAL reference code, and -- when the
reference tree is present (authoring container) -- against that code directly."""
import hashlib
import os

import numpy as np
import pytest

import box_oracle as bo
import ref_loader
import synth

GRID = 1000


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


@pytest.mark.parametrize('preset', ['vgg300', 'vgg512'])
def test_anchor_count_known_answer(preset):
    # the only known-answers the reference ships: ssdutils.py:48,61
    assert bo.anchors(preset).shape[0] == {'vgg300': 8732, 'vgg512': 24564}[preset]


@pytest.mark.parametrize('preset', ['vgg300', 'vgg512'])
def test_anchors_match_reference_fixture(golden_dir, preset):
    g = _load(golden_dir, 'anchors.npz')
    a = bo.anchors(preset)
    assert np.array_equal(a, g[preset + '_prop'])
    assert np.array_equal(bo.anchors_abs(a), g[preset + '_abs'].astype(np.int64))


def _match_cases(g):
    keys = sorted(k[:-3] for k in g.files if k.endswith('_gt'))
    return keys


def test_label_creation_matches_reference_fixture(golden_dir):
    g = _load(golden_dir, 'match.npz')
    cache = {}
    n = 0
    for key in _match_cases(g):
        preset = key.split('_')[0]
        if preset not in cache:
            a = bo.anchors(preset)
            cache[preset] = (a, bo.anchors_abs(a))
        a, aabs = cache[preset]
        vec, match = bo.make_labels(g[key + '_gt'], a, aabs, 20)
        pos = np.nonzero(match >= 0)[0]
        assert np.array_equal(pos, g[key + '_pos']), key
        assert np.array_equal(vec[pos], g[key + '_rows']), key
        neg = np.setdiff1d(np.arange(len(vec)), pos)
        assert np.all(vec[neg, 20] == 1) and np.all(vec[neg, :20] == 0) and np.all(vec[neg, 21:] == 0)
        n += 1
    assert n >= 10


def _pred_for(key, cache):
    preset, dist, i = key.split('_')[0], key.split('_')[1], int(key.split('_')[2])
    if preset not in cache:
        cache[preset] = bo.anchors(preset)
    a = cache[preset]
    pred = synth.pred_uniform(i, a.shape[0]) if dist == 'U' else synth.pred_clustered(i, a)
    return pred, a


def test_decode_nms_matches_reference_fixture(golden_dir):
    g = _load(golden_dir, 'detect.npz')
    cache = {}
    for key in [str(k) for k in g['cases']]:
        parts = key.split('_')
        thr = float(parts[3])
        cap = None if parts[4] == 'None' else int(parts[4])
        pred, a = _pred_for(key, cache)
        digest = hashlib.sha256(np.ascontiguousarray(pred).tobytes()).digest()
        assert digest == g[key + '_insha'].tobytes(), 'synthetic generator drifted: ' + key
        rows, _ = bo.detect(pred, a, thr, cap)
        ref = g[key + '_rows']
        assert rows.shape[0] == ref.shape[0], key
        assert np.array_equal(rows[:, 0].astype(np.uint32).view(np.float32), g[key + '_conf32']), key
        assert np.array_equal(rows[:, 1], ref[:, 1].astype(np.int64)), key
        w = (rows[:, 3] - rows[:, 2]).astype(np.float64)
        h = (rows[:, 5] - rows[:, 4]).astype(np.float64)
        assert np.array_equal((rows[:, 2] + w / 2) / GRID, ref[:, 2]), key
        assert np.array_equal((rows[:, 4] + h / 2) / GRID, ref[:, 3]), key
        assert np.array_equal(w / GRID, ref[:, 4]) and np.array_equal(h / GRID, ref[:, 5]), key


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_oracle_vs_live_reference_random():
    ru, rs, rt = ref_loader.load()
    p = rs.get_preset_by_name('vgg300')
    ra = rs.get_anchors_for_preset(p)
    a = bo.anchors('vgg300')
    aabs = bo.anchors_abs(a)
    lc = rt.LabelCreatorTransform(preset=p, num_classes=20)
    for i in range(200, 212):
        gt = synth.gt_boxes(i, max_boxes=12)
        boxes = [ru.Box('x', int(r[0]), ru.Point(r[1], r[2]), ru.Size(r[3], r[4])) for r in gt]
        _, vec, _ = lc(None, None, ru.Sample('f', boxes, ru.Size(300, 300)))
        assert np.array_equal(vec, bo.make_labels(gt, a, aabs, 20)[0])
    for i in range(50, 53):
        pred = synth.pred_clustered(i, a)
        dets = rs.suppress_overlaps(rs.decode_boxes(pred.copy(), ra, 0.01, {}, 200))
        rows, _ = bo.detect(pred, a, 0.01, 200)
        assert len(dets) == len(rows)
        for (c, b), r in zip(dets, rows):
            cen, siz = ru.abs2prop(r[2], r[3], r[4], r[5], ru.Size(1000, 1000))
            assert np.float32(c).view(np.uint32) == r[0] and b.labelid == r[1]
            assert b.center == cen and b.size == siz
