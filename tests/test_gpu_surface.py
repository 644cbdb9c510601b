"""GPU: the reference-named Python surface (ssdutils / ssdvgg) end to end -- the calls train.py / infer.py make --
against the committed reference fixtures and the oracle."""
import os

import numpy as np
import pytest

import box_oracle as bo
import ssdutils
import synth
import utils
from ssdvgg import SSDVGG, GlobalStep, Session, piecewise_constant

pytestmark = pytest.mark.gpu


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, 'detect.npz'))
    for key in [str(k) for k in g['cases']]:
        parts = key.split('_')
        yield g, key, parts[0], parts[1], int(parts[2]), float(parts[3]), (None if parts[4] == 'None' else int(parts[4]))


def test_decode_boxes_and_suppress_overlaps_like_the_reference_scripts(golden_dir):
    """train.py:275-278 / infer.py:233-235: boxes = decode_boxes(...); boxes = suppress_overlaps(boxes)."""
    anchors = {}
    n = 0
    for g, key, preset, dist, i, thr, cap in _cases(golden_dir):
        if preset not in anchors:
            anchors[preset] = ssdutils.get_anchors_for_preset(ssdutils.get_preset_by_name(preset))
        anc = anchors[preset]
        arr = ssdutils.anchors_as_array(anc)
        pred = synth.pred_uniform(i, len(anc)) if dist == 'U' else synth.pred_clustered(i, arr)
        boxes = ssdutils.decode_boxes(pred, anc, thr, {3: 'car'}, cap)
        assert all(boxes[j][0] >= boxes[j + 1][0] for j in range(len(boxes) - 1))
        kept = ssdutils.suppress_overlaps(boxes)
        ref = g[key + '_rows']
        assert len(kept) == ref.shape[0], key
        for (conf, box), r, c32 in zip(kept, ref, g[key + '_conf32']):
            assert np.float32(conf) == c32 and int(box.labelid) == int(r[1]), key
            assert (box.center.x, box.center.y, box.size.w, box.size.h) == (r[2], r[3], r[4], r[5]), key
            assert box.label == ('car' if int(r[1]) == 3 else None)
        # the fused batched call gives the same list
        fused = ssdutils.detect_batch(pred, anc, thr, {3: 'car'}, cap)[0]
        assert [(float(c), b) for c, b in fused] == [(float(c), b) for c, b in kept], key
        n += 1
    assert n >= 8


def test_create_labels_like_label_creator_transform(golden_dir):
    g = np.load(os.path.join(golden_dir, 'match.npz'))
    key = sorted(k[:-3] for k in g.files if k.endswith('_gt'))[0]
    preset = key.split('_')[0]
    anc = ssdutils.get_anchors_for_preset(ssdutils.get_preset_by_name(preset))
    gt = g[key + '_gt']
    boxes = [utils.Box('x', int(r[0]), utils.Point(r[1], r[2]), utils.Size(r[3], r[4])) for r in gt]
    labels, match = ssdutils.create_labels([boxes, []], anc, 20)
    pos = np.nonzero(match[0] >= 0)[0]
    assert np.array_equal(pos, g[key + '_pos'])
    assert np.allclose(labels[0][pos], g[key + '_rows'], rtol=2e-7, atol=0)
    assert np.all(match[1] == -1) and np.all(labels[1][:, 20] == 1)


def test_session_run_train_eval_infer():
    preset = ssdutils.get_preset_by_name('vgg300')
    anc = ssdutils.get_anchors_for_preset(preset)
    with Session() as sess:
        net = SSDVGG(sess, preset)
        with pytest.raises(RuntimeError):
            sess.run(net.result, feed_dict={net.image_input: np.zeros((1, 300, 300, 3), np.float32)})
        net.build_from_vgg(None, 20)
        step = GlobalStep(0)
        net.build_optimizer(learning_rate=piecewise_constant(step, [2, 4], [1e-5, 1e-6, 1e-7]), weight_decay=0.0005,
                            momentum=0.9, global_step=step)
        x = synth.images(0, 2, 300)
        gts = [synth.gt_boxes(i) for i in range(2)]
        boxes = [[utils.Box(None, int(r[0]), utils.Point(r[1], r[2]), utils.Size(r[3], r[4])) for r in gt] for gt in gts]
        y, _ = ssdutils.create_labels(boxes, anc, 20)
        # inference fetch (infer.py:225-227)
        res = np.array(sess.run(net.result, feed_dict={net.image_input: x, net.keep_prob: 1}))
        assert res.shape == (2, 8732, 25) and np.allclose(res[..., :21].sum(-1), 1, atol=1e-4)
        # validation fetch (train.py:291-294): no update
        r2, l2 = sess.run([net.result, net.losses], feed_dict={net.image_input: x, net.labels: y})
        assert np.allclose(np.array(r2), res, atol=1e-6) and step.value == 0
        # training fetch (train.py:262-266)
        first = None
        for _ in range(3):
            r3, l3, _ = sess.run([net.result, net.losses, net.optimizer], feed_dict={net.image_input: x, net.labels: y})
            first = first or dict(l3)
        assert step.value == 3
        assert abs(first['total'] - l2['total']) < 1e-3 * abs(l2['total'])          # first step's losses are pre-update
        assert l3['total'] < first['total']                                         # and the same batch gets easier
        assert abs(l3['total'] - (l3['confidence'] + l3['localization'] + l3['l2'])) < 1e-3 * l3['total']
        with pytest.raises(ValueError):
            sess.run([net.loss], feed_dict={net.image_input: x})
        # save / restore round trip under the reference's variable names
        import tempfile
        d = tempfile.mkdtemp()
        net.save(os.path.join(d, 'final'))
        res_after = np.array(sess.run(net.result, feed_dict={net.image_input: x}))
    with Session() as sess2:
        net2 = SSDVGG(sess2, preset)
        net2.build_from_metagraph(None, os.path.join(d, 'final.npz'))
        res_rest = np.array(sess2.run(net2.result, feed_dict={net2.image_input: x}))
        assert np.array_equal(res_rest, res_after)


def test_ap_from_gpu_detections_equals_reference_fixture(golden_dir):
    """decode + NMS on the GPU -> APCalculator.add_detections_batch == the APs the REAL reference computed from its own
    decode_boxes / suppress_overlaps / APCalculator chain (tests/golden/ap.npz)."""
    import average_precision as ap
    g = np.load(os.path.join(golden_dir, 'ap.npz'))
    anc = bo.anchors('vgg300')
    lid2name = {i: 'c%d' % i for i in range(20)}
    preds, gts = [], []
    for i in range(24):
        pred, objs = synth.pred_clustered(5000 + i, anc, return_objects=True)
        if i % 3 == 2:
            pred = synth.pred_clustered(9000 + i, anc)
        preds.append(pred); gts.append(objs)
    dets, counts = ssdutils.detect_batch_rows(np.stack(preds), ssdutils.get_anchors_for_preset(ssdutils.get_preset_by_name('vgg300')), 0.01, 200)
    for tag, mo in (('m50', 0.5), ('m70', 0.7)):
        calc = ap.APCalculator(mo)
        calc.add_detections_batch(gts, dets, counts, lid2name)
        aps = calc.compute_aps()
        ids = [int(k) for k in g['aps_%s_ids' % tag]]
        assert np.array_equal(np.array([aps[lid2name[k]] for k in ids]), g['aps_' + tag])
        assert ap.APs2mAP(aps) == g['map_' + tag][0]
