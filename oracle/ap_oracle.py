"""CPU restatement of the reference's VOC07 11-point average precision -- TEST INFRASTRUCTURE ONLY.

Follows average_precision.py:30-42 (APs2mAP) and :45-192 (APCalculator) of the reference: detections are
re-quantised through utils.prop2abs on the 1000x1000 grid (:75), grouped per class; only classes that
have ground truth are scored (:110); detections are visited in descending confidence (:118-121), each
is matched to the ground-truth box of its own image with the highest inclusive-pixel IoU
(ssdutils.jaccard_overlap, first maximum) and counts as a true positive when that IoU is not below
`minoverlap` and the box is still unmatched (:133-160); precision / recall from cumulative sums, AP =
mean over r in {0, 0.1, .. 1} of max precision at recall >= r (:165-176).

Pinned against the real reference code (oracle/ref_loader.load_ap, which supplies the removed `np.int`
alias the reference needs under NumPy 2) in tests/test_oracle_golden.py and through tests/golden/ap.npz.
Plain Python loops on purpose: this is the checker, not the product (ssd-tensorflow_b200/average_precision.py).
"""
import numpy as np


def prop2abs_1000(cx, cy, w, h):
    """utils.prop2abs on Size(1000, 1000) (utils.py:100-108): int() truncation of float64 expressions."""
    hw = w * 1000 / 2
    hh = h * 1000 / 2
    px = cx * 1000
    py = cy * 1000
    return int(px - hw), int(px + hw), int(py - hh), int(py + hh)


def iou_incl(a, b):
    """ssdutils.jaccard_overlap (ssdutils.py:138-152) for two (xmin, xmax, ymin, ymax) boxes; float64 quotient."""
    area_a = (a[1] - a[0] + 1) * (a[3] - a[2] + 1)
    area_b = (b[1] - b[0] + 1) * (b[3] - b[2] + 1)
    w = max(0, min(a[1], b[1]) - max(a[0], b[0]) + 1)
    h = max(0, min(a[3], b[3]) - max(a[2], b[2]) + 1)
    inter = w * h
    return float(inter) / float(area_a + area_b - inter)


def compute_aps(gt_per_sample, det_per_sample, minoverlap=0.5):
    """gt_per_sample[s]  = [(label, cx, cy, w, h), ...]          proportional boxes (float)
    det_per_sample[s] = [(conf, label, cx, cy, w, h), ...]    detections of sample s
    -> {label: ap} over the labels that occur in the ground truth (average_precision.py:84-181)."""
    counts, gt_map = {}, {}
    for s, boxes in enumerate(gt_per_sample):
        for (label, cx, cy, w, h) in boxes:
            counts[label] = counts.get(label, 0) + 1
            gt_map.setdefault(label, {}).setdefault(s, []).append(prop2abs_1000(cx, cy, w, h))
    dets = {}
    for s, boxes in enumerate(det_per_sample):
        for (conf, label, cx, cy, w, h) in boxes:
            dets.setdefault(label, []).append((np.float32(conf), s, prop2abs_1000(cx, cy, w, h)))
    aps = {}
    for label in gt_map:
        rows = dets.get(label, [])
        order = sorted(range(len(rows)), key=lambda i: -float(rows[i][0]))     # distinct confidences assumed (argsort ties)
        matched = {s: [False] * len(v) for s, v in gt_map[label].items()}
        tp, fp = [], []
        for i in order:
            _, s, box = rows[i]
            if s not in gt_map[label]:
                tp.append(0); fp.append(1); continue
            ious = [iou_incl(box, g) for g in gt_map[label][s]]
            j = int(np.argmax(ious))
            if ious[j] < minoverlap or matched[s][j]:
                tp.append(0); fp.append(1); continue
            matched[s][j] = True
            tp.append(1); fp.append(0)
        tps, fps = np.cumsum(np.array(tp, np.float64)), np.cumsum(np.array(fp, np.float64))
        recall = tps / counts[label]
        prec = tps / (tps + fps) if len(tp) else tps
        ap = 0.0
        for r in np.arange(0, 1.1, 0.1):
            sel = prec[recall >= r]
            if len(sel) > 0:
                ap += np.amax(sel)
        aps[label] = ap / 11.0
    return aps


def aps2map(aps):
    """APs2mAP (average_precision.py:30-42)."""
    return sum(aps.values()) / len(aps) if len(aps) else 0
