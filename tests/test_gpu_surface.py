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


def _model(sess, lr=1e-5):
    preset = ssdutils.get_preset_by_name('vgg300')
    net = SSDVGG(sess, preset)
    net.build_from_vgg(None, 20)
    step = GlobalStep(0)
    net.build_optimizer(learning_rate=lr, weight_decay=0.0005, momentum=0.9, global_step=step)
    return net, step, ssdutils.get_anchors_for_preset(preset)


def test_gt_feed_is_the_dense_label_step_with_the_match_inside():
    """net.gt_boxes (raw ground truth, fused match + loss on the product path: ssdb_train_step_host_gt) against net.labels
    (dense labels built by LabelCreatorTransform's GPU port) and against the oracle's matching: same match indices,
    same losses, same result, same parameters after the step (transforms.py:72-114 -> train.py:262-266)."""
    x = synth.images(0, 3, 300)
    gts = [synth.gt_boxes(i) for i in range(2)] + [np.zeros((0, 5))]          # the third image has no box at all
    boxes = [[utils.Box(None, int(r[0]), utils.Point(r[1], r[2]), utils.Size(r[3], r[4])) for r in gt] for gt in gts]
    anc_arr = bo.anchors('vgg300'); aabs = bo.anchors_abs(anc_arr)
    p0 = _model(Session())[0].get_params()
    with Session() as s1:
        n1, _, anc = _model(s1)
        y, _ = ssdutils.create_labels(boxes, anc, 20)
        r1, l1, _ = s1.run([n1.result, n1.losses, n1.optimizer], feed_dict={n1.image_input: x, n1.labels: y})
        p1 = n1.get_params()
    with Session() as s2:
        n2, step, _ = _model(s2)
        # validation fetch with raw GT: forward + loss only, nothing moves
        rv, lv, mv = s2.run([n2.result, n2.losses, n2.match], feed_dict={n2.image_input: x, n2.gt_boxes: boxes})
        assert step.value == 0 and all(np.array_equal(v, p0[k]) for k, v in n2.get_params().items())
        r2, l2, _, m2 = s2.run([n2.result, n2.losses, n2.optimizer, n2.match], feed_dict={n2.image_input: x, n2.gt_boxes: boxes})
        p2 = n2.get_params()
        assert step.value == 1
    for b in range(3):
        want = bo.make_labels(gts[b], anc_arr, aabs, 20)[1] if len(gts[b]) else np.full(8732, -1)
        assert np.array_equal(m2[b], want), 'match indices of image %d differ from the oracle' % b
        assert np.array_equal(mv[b], want)
    assert np.array_equal(np.array(r1), np.array(r2)) and np.array_equal(np.array(rv), np.array(r2))
    for k in ('total', 'localization', 'confidence', 'l2'):
        assert abs(l1[k] - l2[k]) <= 2e-6 * abs(l1[k]), (k, l1[k], l2[k])
        assert abs(lv[k] - l2[k]) <= 2e-6 * abs(l1[k])
    # the two loss paths differ only in how the log() of the encoded offsets is rounded (<= 1 float32 ulp, tests/test_gpu_box.py)
    for k in p1:
        upd = np.abs(p1[k] - p0[k]).max()                      # size of the update both runs applied
        d = np.abs(p1[k] - p2[k]).max()
        assert d <= 1e-3 * upd + 2.4e-7 * max(np.abs(p1[k]).max(), 1e-30), (k, d, upd)
    # packed-array form of the same feed, and a bad label id
    gt, cnt = synth.pack_gt(gts, 8)
    with Session() as s3:
        n3, _, _ = _model(s3)
        _, l3 = s3.run([n3.result, n3.losses], feed_dict={n3.image_input: x, n3.gt_boxes: gt, n3.gt_counts: cnt})
        assert abs(l3['total'] - l2['total']) <= 2e-6 * abs(l2['total'])
        gt[0, 0, 0] = 20
        import ssdb
        with pytest.raises(ssdb.SSDBError):
            s3.run([n3.losses], feed_dict={n3.image_input: x, n3.gt_boxes: gt, n3.gt_counts: cnt})


def test_detect_on_the_device_resident_result_equals_the_two_step_flow():
    """SSDVGG.detect (ssdb_forward_detect_host: forward, then decode + NMS on the resident result; infer.py:225-235 as one
    flow) == sess.run(net.result) followed by detect_batch on the host copy, for a cap of 200 and for no cap."""
    x = synth.images(40, 3, 300)
    with Session() as sess:
        net, _, anc = _model(sess)
        res = sess.run(net.result, feed_dict={net.image_input: x, net.keep_prob: 1})
        for cap, thr in ((200, 0.01), (None, 0.02), (7, 0.01)):
            want = ssdutils.detect_batch(res, anc, thr, {}, cap)
            got = net.detect(x, thr, {}, cap)
            assert len(want) == len(got) == 3
            assert sum(len(w) for w in want) > 0
            for w, g in zip(want, got):
                assert [(float(c), b) for c, b in w] == [(float(c), b) for c, b in g]
            dets, counts = net.detect(x, thr, {}, cap, rows=True)
            d2, c2 = ssdutils.detect_batch_rows(res, anc, thr, cap)
            assert np.array_equal(counts, c2)
            for i in range(3):
                assert np.array_equal(dets[i, :counts[i, 0]], d2[i, :c2[i, 0]])


def test_fetched_arrays_are_the_callers_own():
    """tf.Session.run returns fresh arrays: a result must survive later runs, a larger batch (engine re-creation) and close()."""
    x1, x2 = synth.images(0, 2, 300), synth.images(7, 2, 300)
    sess = Session()
    net, _, _ = _model(sess)
    r1 = sess.run(net.result, feed_dict={net.image_input: x1})
    keep = np.array(r1)
    c1 = sess.run(net.classifier, feed_dict={net.image_input: x1})
    r2 = sess.run(net.result, feed_dict={net.image_input: x2})
    assert np.array_equal(r1, keep) and not np.array_equal(r2, keep)
    assert np.array_equal(c1, keep[..., :21])
    r3 = sess.run(net.result, feed_dict={net.image_input: synth.images(0, 3, 300)})      # larger batch: new engine
    assert np.allclose(r3[:2], keep, rtol=0, atol=1e-4)       # another batch size tiles the GEMMs differently: same values up to summation order
    lg = sess.run(net.logits, feed_dict={net.image_input: x1})                              # ssdvgg.py:366
    assert lg.shape == (2, 8732, 21)
    e = np.exp(lg - lg.max(-1, keepdims=True)); sm = e / e.sum(-1, keepdims=True)
    assert np.abs(sm - keep[..., :21]).max() < 1e-5
    sess.close()
    assert np.array_equal(r1, keep) and np.allclose(r3[:2], keep, rtol=0, atol=1e-4) and np.isfinite(r2).all()


def test_continue_training_restores_momentum_step_and_epoch(tmp_path):
    """train.py:101-134,191-193: the Saver restores the variables, the Momentum slots and global_step; a resumed run must
    take exactly the step the uninterrupted run takes."""
    x = synth.images(0, 2, 300)
    gt, cnt = synth.pack_gt([synth.gt_boxes(i) for i in range(2)], 8)
    for tf_ckpt in (False, True):
        with Session() as sess:
            net, step, _ = _model(sess, lr=piecewise_constant(GlobalStep(0), [1], [1e-5, 1e-6]))
            net._opt['lr'] = piecewise_constant(step, [1], [1e-5, 1e-6])
            feed = {net.image_input: x, net.gt_boxes: gt, net.gt_counts: cnt}
            sess.run([net.losses, net.optimizer], feed_dict=feed)
            sess.run([net.losses, net.optimizer], feed_dict=feed)
            net.epoch = 3
            path = str(tmp_path / ('e3_%d' % tf_ckpt))
            net.save(path, tf_checkpoint=tf_ckpt)
            sess.run([net.losses, net.optimizer], feed_dict=feed)
            want = net.get_params(); want_m = net.get_params(2)
        with Session() as s2:
            n2 = SSDVGG(s2, ssdutils.get_preset_by_name('vgg300'))
            n2.build_from_metagraph(None, path if tf_ckpt else path + '.npz')
            step2 = GlobalStep(0)
            n2.build_optimizer_from_metagraph(piecewise_constant(step2, [1], [1e-5, 1e-6]), 0.0005, 0.9, step2)
            assert step2.value == 2 and (tf_ckpt or n2.epoch == 3)
            s2.run([n2.losses, n2.optimizer], feed_dict={n2.image_input: x, n2.gt_boxes: gt, n2.gt_counts: cnt})
            got = n2.get_params(); got_m = n2.get_params(2)
        for k in want:
            assert np.array_equal(want[k], got[k]), k
            assert np.array_equal(want_m[k], got_m[k]), k


def test_frozen_model_export_detect_and_cuda_graph(tmp_path):
    """export_model.py -> detect.py (reference export_model.py:62-72, detect.py:90-112): the frozen, inference-only engine gives
    the same detections as the training engine; the CUDA-graph replay (third call) the same as the eager first call; training
    entry points refuse an inference handle; the CLI writes the reference's per-image text format."""
    import subprocess, sys
    import ssdb
    x = synth.images(11, 4, 300)
    with Session() as sess:
        net, _, anc = _model(sess)
        want = net.detect(x, 0.01, {}, 200, rows=True)
        d = str(tmp_path / 'final')
        net.save(d)
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'ssd-tensorflow_b200')
    frozen = str(tmp_path / 'model.frozen.npz')
    r = subprocess.run([sys.executable, os.path.join(pkg, 'export_model.py'), '--checkpoint-file', d + '.npz', '--output-file', frozen,
                        '--output-tensors', 'result/result:0'], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    with Session() as s2:
        n2 = SSDVGG(s2, ssdutils.get_preset_by_name('vgg300'))
        n2.build_from_frozen(frozen)
        with pytest.raises(RuntimeError):
            n2.build_optimizer()
        launches = []
        for it in range(4):                       # eager, capture + launch, replay, replay
            l0 = ssdb.launch_count()
            got = n2.detect(x, 0.01, {}, 200, rows=True)
            launches.append(ssdb.launch_count() - l0)
            assert np.array_equal(got[1], want[1])
            for i in range(4):
                assert np.array_equal(got[0][i, :got[1][i, 0]], want[0][i, :want[1][i, 0]]), (it, i)
        assert launches[0] > 50 and launches[2] == 1 and launches[3] == 1, launches       # one graph launch replaces the kernel list
        eng = n2._engine
        assert eng.inference
        with pytest.raises(ssdb.SSDBError):
            eng.train_step_host(x, np.zeros((4, 8732, 25), np.float32), 1e-5, 0.9, 0.0005)
        with pytest.raises(ssdb.SSDBError):
            eng.get_tensor('conv1_1/filter', (3, 3, 3, 64), ssdb.GRAD)
        # another batch size and another threshold get their own graphs; plain forward still works on the handle
        got1 = n2.detect(x[:1], 0.01, {}, 200, rows=True)
        assert np.array_equal(got1[0][0, :got1[1][0, 0]], want[0][0, :want[1][0, 0]])
        res = s2.run(n2.result, feed_dict={n2.image_input: x})
        assert res.shape == (4, 8732, 25)
    # the CLI on real image files
    import cv2
    files = []
    for i in range(2):
        f = str(tmp_path / ('img%d.png' % i))
        cv2.imwrite(f, np.clip(synth.images(20 + i, 1, 320)[0], 0, 255).astype(np.uint8))
        files.append(f)
    out = str(tmp_path / 'out')
    r = subprocess.run([sys.executable, os.path.join(pkg, 'detect.py'), '--model', frozen, '--output-dir', out, '--threshold', '0.01'] + files,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for f in files:
        base = os.path.basename(f)
        assert os.path.exists(os.path.join(out, base))
        lines = open(os.path.join(out, base + '.txt')).read().splitlines()
        assert 0 < len(lines) <= 200 and all(len(l.split()) == 6 for l in lines)


def test_cuda_graph_forward_with_cluster_launches_at_batch_32():
    """At batch 32 the N >= 128 layers of the forward run as clusters of two CTAs (cta_group::2 pair kernels, cudaLaunchKernelEx with
    a cluster dimension): the CUDA graph of an inference handle must capture and replay them, with the detections of the eager
    training engine."""
    import ssdb
    preset = 'vgg300'
    x = synth.images(40, 32, 300)
    nets = [ssdb.Net(preset, 20, max_batch=32), ssdb.Net(preset, 20, max_batch=32, inference=True)]
    P = SSDVGG(Session(), ssdutils.get_preset_by_name(preset))._initial_params(20, seed=3)
    for net in nets:
        for k, shape in net.tensors():
            net.set_tensor(k, P[k])
    want = nets[0].forward_detect_host(x, 0.01, 200, 0.45)
    launches = []
    for it in range(3):                               # eager, capture + launch, replay
        l0 = ssdb.launch_count()
        got = nets[1].forward_detect_host(x, 0.01, 200, 0.45)
        launches.append(ssdb.launch_count() - l0)
        assert np.array_equal(got[1], want[1]), it
        for i in range(32):
            assert np.array_equal(got[0][i, :got[1][i, 0]], want[0][i, :want[1][i, 0]]), (it, i)
    assert launches[0] > 50 and launches[2] == 1, launches
    for net in nets:
        net.close()
