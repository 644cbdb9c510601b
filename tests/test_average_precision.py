"""AP evaluation: the oracle restatement against the fixture produced by the REAL reference
average_precision.py (tests/golden/ap.npz, oracle/make_golden.py) and against the live reference when its tree is present;
the package's APCalculator (Box-tuple input and GPU-row input) against both."""
import os

import numpy as np
import pytest

import ap_oracle
import average_precision as ap
import box_oracle as bo
import ref_loader
import synth
from utils import Box, Point, Size, abs2prop, GRID

N_IMG = 24
LID2NAME = {i: 'c%d' % i for i in range(20)}


@pytest.fixture(scope='module')
def fixture(golden_dir):
    return np.load(os.path.join(golden_dir, 'ap.npz'))


def _inputs(g):
    gts = [[(int(r[0]), r[1], r[2], r[3], r[4]) for r in g['img%d_gt' % i]] for i in range(N_IMG)]
    dets = [[(np.float32(c), int(r[1]), r[2], r[3], r[4], r[5]) for r, c in zip(g['img%d_det' % i], g['img%d_conf32' % i])]
            for i in range(N_IMG)]
    return gts, dets


@pytest.mark.parametrize('tag,minoverlap', [('m50', 0.5), ('m70', 0.7)])
def test_oracle_reproduces_reference_fixture(fixture, tag, minoverlap):
    gts, dets = _inputs(fixture)
    aps = ap_oracle.compute_aps(gts, dets, minoverlap)
    assert sorted(aps) == [int(k) for k in fixture['aps_%s_ids' % tag]]
    assert np.array_equal(np.array([aps[int(k)] for k in fixture['aps_%s_ids' % tag]]), fixture['aps_' + tag])
    assert ap_oracle.aps2map(aps) == fixture['map_' + tag][0]


@pytest.mark.parametrize('tag,minoverlap', [('m50', 0.5), ('m70', 0.7)])
def test_package_calculator_reproduces_reference_fixture(fixture, tag, minoverlap):
    gts, dets = _inputs(fixture)
    calc = ap.APCalculator(minoverlap)
    for gt, det in zip(gts, dets):
        gt_boxes = [Box(LID2NAME[l], l, Point(cx, cy), Size(w, h)) for (l, cx, cy, w, h) in gt]
        boxes = [(c, Box(LID2NAME[l], l, Point(cx, cy), Size(w, h))) for (c, l, cx, cy, w, h) in det]
        calc.add_detections(gt_boxes, boxes)
    aps = calc.compute_aps()
    ids = [int(k) for k in fixture['aps_%s_ids' % tag]]
    assert sorted(aps) == sorted(LID2NAME[k] for k in ids)
    assert np.array_equal(np.array([aps[LID2NAME[k]] for k in ids]), fixture['aps_' + tag])
    assert ap.APs2mAP(aps) == fixture['map_' + tag][0]
    calc.clear()
    assert calc.compute_aps() == {} and ap.APs2mAP({}) == 0


def _preds(i, anc):
    pred, objs = synth.pred_clustered(5000 + i, anc, return_objects=True)
    if i % 3 == 2:
        pred = synth.pred_clustered(9000 + i, anc)
    return pred, objs


def test_batch_input_from_integer_rows_equals_fixture(fixture):
    """add_detections_batch consumes the integer rows of decode + NMS (here from the NumPy box oracle, which the GPU kernels
    match bit for bit): same APs as the reference computed from its own Box tuples."""
    anc = bo.anchors('vgg300')
    calc = ap.APCalculator(0.5)
    dets = np.zeros((N_IMG, 200, 8), np.int32); counts = np.zeros((N_IMG, 2), np.int32)
    gts = []
    for i in range(N_IMG):
        pred, objs = _preds(i, anc)
        rows, cand = bo.detect(pred, anc, 0.01, 200)
        dets[i, :len(rows), 0] = rows[:, 0].astype(np.uint32).view(np.int32)
        dets[i, :len(rows), 1:7] = rows[:, 1:7]
        counts[i] = (len(rows), cand['idx'].shape[0])
        gts.append(objs)
    calc.add_detections_batch(gts, dets, counts, LID2NAME)
    aps = calc.compute_aps()
    ids = [int(k) for k in fixture['aps_m50_ids']]
    assert np.array_equal(np.array([aps[LID2NAME[k]] for k in ids]), fixture['aps_m50'])


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present (GPU box): fixtures cover it')
def test_oracle_and_package_vs_live_reference_random():
    rap = ref_loader.load_ap()
    ru, _, _ = ref_loader.load()
    rng = np.random.default_rng(11)
    for trial in range(4):
        rcalc, calc = rap.APCalculator(0.5), ap.APCalculator(0.5)
        gts, dets = [], []
        used = set()
        for s in range(12):
            g = [(int(rng.integers(0, 5)), *rng.uniform(0.2, 0.8, 2), *rng.uniform(0.05, 0.4, 2)) for _ in range(int(rng.integers(0, 5)))]
            d = []
            for (l, cx, cy, w, h) in g:                                     # jittered copies of the GT + random boxes
                for _ in range(int(rng.integers(0, 4))):
                    j = rng.normal(0, 0.03, 4)
                    d.append((l, cx + j[0], cy + j[1], abs(w + j[2]) + 0.01, abs(h + j[3]) + 0.01))
            for _ in range(int(rng.integers(0, 6))):
                d.append((int(rng.integers(0, 6)), *rng.uniform(0.2, 0.8, 2), *rng.uniform(0.05, 0.4, 2)))
            rows = []
            for (l, cx, cy, w, h) in d:
                c = np.float32(rng.random())
                while (l, c) in used:
                    c = np.float32(rng.random())
                used.add((l, c))
                rows.append((c, l, cx, cy, w, h))
            gts.append(g); dets.append(rows)
            rcalc.add_detections([ru.Box('n%d' % l, l, ru.Point(cx, cy), ru.Size(w, h)) for (l, cx, cy, w, h) in g],
                                 [(c, ru.Box('n%d' % l, l, ru.Point(cx, cy), ru.Size(w, h))) for (c, l, cx, cy, w, h) in rows])
            calc.add_detections([Box('n%d' % l, l, Point(cx, cy), Size(w, h)) for (l, cx, cy, w, h) in g],
                                [(c, Box('n%d' % l, l, Point(cx, cy), Size(w, h))) for (c, l, cx, cy, w, h) in rows])
        want = rcalc.compute_aps()
        got_o = ap_oracle.compute_aps(gts, dets, 0.5)
        got_p = calc.compute_aps()
        assert {int(k[1:]): v for k, v in want.items()} == got_o
        assert want == got_p
        assert rap.APs2mAP(want) == ap.APs2mAP(got_p) == ap_oracle.aps2map(got_o)


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present (GPU box)')
def test_pascal_summary_files_equal_the_reference(tmp_path):
    import importlib.util
    import sys
    import cv2
    import pascal_summary as ps
    ru, rs, _ = ref_loader.load()
    saved = sys.modules.get('utils')
    sys.modules['utils'] = ru
    try:
        spec = importlib.util.spec_from_file_location('_ref_pascal_summary', os.path.join(ref_loader.REF_DIR, 'pascal_summary.py'))
        rps = importlib.util.module_from_spec(spec); spec.loader.exec_module(rps)
    finally:
        if saved is None:
            sys.modules.pop('utils', None)
        else:
            sys.modules['utils'] = saved
    rng = np.random.default_rng(5)
    ref, got = rps.PascalSummary(), ps.PascalSummary()
    for k, (w, h) in enumerate([(500, 375), (333, 500), (64, 48)]):
        fn = str(tmp_path / ('img_%d.x.jpg' % k))
        cv2.imwrite(fn, np.zeros((h, w, 3), np.uint8))
        rows = [(np.float32(rng.random()), 'c%d' % int(rng.integers(0, 3)), *rng.uniform(-0.1, 1.1, 2), *rng.uniform(0.01, 0.9, 2)) for _ in range(20)]
        ref.add_detections(fn, [(c, ru.Box(l, 0, ru.Point(cx, cy), ru.Size(bw, bh))) for (c, l, cx, cy, bw, bh) in rows])
        if k == 1:
            got.add_detections(fn, [(c, Box(l, 0, Point(cx, cy), Size(bw, bh))) for (c, l, cx, cy, bw, bh) in rows])       # via cv2
        else:
            got.add_detections(fn, [(c, Box(l, 0, Point(cx, cy), Size(bw, bh))) for (c, l, cx, cy, bw, bh) in rows], Size(w, h))
    (tmp_path / 'ref').mkdir(); (tmp_path / 'got').mkdir()
    ref.write_summary(str(tmp_path / 'ref')); got.write_summary(str(tmp_path / 'got'))
    names = sorted(os.listdir(tmp_path / 'ref'))
    assert names == sorted(os.listdir(tmp_path / 'got')) and len(names) == 3
    for n in names:
        assert (tmp_path / 'ref' / n).read_text() == (tmp_path / 'got' / n).read_text()
