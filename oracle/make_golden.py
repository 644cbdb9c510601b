"""Generate tests/golden/*.npz from the REAL reference code -- TEST INFRASTRUCTURE ONLY.

Run in the authoring container (needs /root/reference):
    python oracle/make_golden.py
The reference's NumPy code (ssdutils.py, utils.py, transforms.py) is imported
unmodified under a stub `tensorflow` module (oracle/ref_loader.py) and run on
seeded synthetic inputs (ssd-tensorflow_b200/synth.py).  Outputs:
  anchors.npz   G1: default boxes of both presets (+ 1000-grid integer bounds)
  match.npz     G2: LabelCreatorTransform label tensors, stored sparsely
  detect.npz    G3: decode_boxes + suppress_overlaps results
  ap.npz        G7: APCalculator.compute_aps of the reference on decode + NMS output of 24 clustered images
Inputs whose result would depend on NumPy's unspecified argsort tie order
(duplicate confidences inside the candidate set) are skipped, as SURVEY 8c asks.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, 'ssd-tensorflow_b200'))
import ref_loader  # noqa: E402
import synth       # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
DETECT_CASES = [  # (preset, dist, image index, thr, cap)
    ('vgg300', 'U', 0, 0.01, 200), ('vgg300', 'U', 1, 0.01, 200), ('vgg300', 'U', 2, 0.5, 200),
    ('vgg300', 'U', 3, 0.9, None), ('vgg300', 'C', 0, 0.01, 200), ('vgg300', 'C', 1, 0.01, 200),
    ('vgg300', 'C', 2, 0.5, 200), ('vgg300', 'C', 3, 0.3, None), ('vgg300', 'C', 4, 0.01, 50),
    ('vgg512', 'U', 0, 0.01, 200), ('vgg512', 'C', 0, 0.01, 200), ('vgg512', 'C', 1, 0.4, None),
]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ru, rs, rt = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    grid = ru.Size(1000, 1000)
    anchors = {}
    g1 = {}
    for name in ('vgg300', 'vgg512'):
        p = rs.get_preset_by_name(name)
        ra = rs.get_anchors_for_preset(p)
        anchors[name] = (p, ra)
        g1[name + '_prop'] = np.array([[a.center.x, a.center.y, a.size.w, a.size.h] for a in ra])
        g1[name + '_abs'] = rs.anchors2array(ra, grid).astype(np.int32)
    np.savez_compressed(os.path.join(OUT, 'anchors.npz'), **g1)

    g2 = {}
    for name, idxs, maxb in (('vgg300', range(8), 16), ('vgg300', range(100, 104), 3), ('vgg512', range(3), 16)):
        p, ra = anchors[name]
        lc = rt.LabelCreatorTransform(preset=p, num_classes=20)
        for i in idxs:
            gt = synth.gt_boxes(i, max_boxes=maxb)
            boxes = [ru.Box('x', int(g[0]), ru.Point(g[1], g[2]), ru.Size(g[3], g[4])) for g in gt]
            _, vec, _ = lc(None, None, ru.Sample('f', boxes, ru.Size(300, 300)))
            pos = np.nonzero(vec[:, 20] == 0)[0]
            key = '%s_%d_%d' % (name, i, maxb)
            g2[key + '_gt'] = gt
            g2[key + '_pos'] = pos.astype(np.int32)
            g2[key + '_rows'] = vec[pos]
            assert np.all(vec[np.setdiff1d(np.arange(len(vec)), pos), 20] == 1)
    np.savez_compressed(os.path.join(OUT, 'match.npz'), **g2)

    g3 = {}
    kept_cases = []
    for (name, dist, i, thr, cap) in DETECT_CASES:
        p, ra = anchors[name]
        prop = g1[name + '_prop']
        pred = synth.pred_uniform(i, len(ra)) if dist == 'U' else synth.pred_clustered(i, prop)
        nc = pred.shape[1] - 4
        cls = np.argmax(pred[:, :nc - 1], axis=1)
        conf = pred[np.arange(len(pred)), cls]
        srt = np.sort(conf)[::-1]
        ncand = int((srt >= np.float32(thr)).sum())
        if cap is not None:
            ncand = min(ncand, cap)
        head = srt[:ncand + 1]
        if np.any(head[1:] == head[:-1]):
            print('skip (tie in candidate set):', name, dist, i, thr, cap)
            continue
        work = pred.copy()
        dets = rs.suppress_overlaps(rs.decode_boxes(work, ra, thr, {}, cap))
        rows = np.zeros((len(dets), 6), np.float64)
        for r, (c, b) in enumerate(dets):
            rows[r] = (float(c), b.labelid, b.center.x, b.center.y, b.size.w, b.size.h)
        key = '%s_%s_%d_%g_%s' % (name, dist, i, thr, cap)
        g3[key + '_rows'] = rows
        g3[key + '_conf32'] = np.array([c for c, _ in dets], np.float32)
        g3[key + '_insha'] = np.frombuffer(bytes.fromhex(sha(pred)), np.uint8)
        kept_cases.append(key)
        print(key, 'candidates', ncand, 'kept', len(dets))
    g3['cases'] = np.array(kept_cases)
    np.savez_compressed(os.path.join(OUT, 'detect.npz'), **g3)
    make_ap(ru, rs, anchors, g1)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


AP_IMAGES = 24


def make_ap(ru, rs, anchors, g1):
    """G7: the reference's APCalculator fed with the reference's own decode_boxes + suppress_overlaps output; ground truth =
    the objects the clustered prediction generator planted.  Stored: per image the GT rows, the detection rows
    (conf, labelid, cx, cy, w, h) and the resulting per-class APs."""
    rap = ref_loader.load_ap()
    p, ra = anchors['vgg300']
    prop = g1['vgg300_prop']
    lid2name = {i: 'c%d' % i for i in range(20)}
    out = {}
    for minoverlap, tag in ((0.5, 'm50'), (0.7, 'm70')):
        calc = rap.APCalculator(minoverlap)
        all_conf = []
        for i in range(AP_IMAGES):
            pred, objs = synth.pred_clustered(5000 + i, prop, return_objects=True)
            if i % 3 == 2:        # every third image: detections unrelated to the ground truth (false positives at all confidences)
                pred = synth.pred_clustered(9000 + i, prop)
            gt_boxes = [ru.Box(lid2name[int(o[0])], int(o[0]), ru.Point(o[1], o[2]), ru.Size(o[3], o[4])) for o in objs]
            dets = rs.suppress_overlaps(rs.decode_boxes(pred.copy(), ra, 0.01, lid2name, 200))
            calc.add_detections(gt_boxes, dets)
            rows = np.array([(float(c), b.labelid, b.center.x, b.center.y, b.size.w, b.size.h) for c, b in dets], np.float64).reshape(-1, 6)
            out['img%d_gt' % i] = objs
            out['img%d_det' % i] = rows
            out['img%d_conf32' % i] = np.array([c for c, _ in dets], np.float32)
            all_conf.extend((b.labelid, np.float32(c)) for c, b in dets)
        assert len(set(all_conf)) == len(all_conf), 'duplicate (class, confidence): argsort tie order would matter'
        aps = calc.compute_aps()
        ids = sorted(int(k[1:]) for k in aps)
        out['aps_%s_ids' % tag] = np.array(ids, np.int32)
        out['aps_%s' % tag] = np.array([aps['c%d' % k] for k in ids], np.float64)
        out['map_%s' % tag] = np.array([rap.APs2mAP(aps)], np.float64)
        print('AP fixture', tag, 'classes', len(ids), 'mAP', rap.APs2mAP(aps))
    np.savez_compressed(os.path.join(OUT, 'ap.npz'), **out)


if __name__ == '__main__':
    main()
