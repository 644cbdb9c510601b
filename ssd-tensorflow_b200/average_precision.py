"""VOC07 11-point average precision: the step right after NMS in the reference's loops
(train.py:278,303 `ap_calc.add_detections`, infer.py:260-264) -- SURVEY.md 8f row 3.

Mirrors the public names of the reference's ``average_precision.py`` (``APs2mAP`` :30-42, ``APCalculator`` :45-192:
``add_detections(gt_boxes, boxes)``, ``compute_aps()``, ``clear()``) and adds ``add_detections_batch`` which takes the
integer rows the GPU decode + NMS kernels return (``ssdb.decode_nms_host`` / ``ssdutils.detect_batch``) without building
``Box`` tuples.  Host-side NumPy: one image contributes at most a few hundred boxes, so there is no device work here; the
boxes arrive already on the 1000x1000 grid, and every IoU decision is exact integer arithmetic (IoU >= t with I, U < 2^21
is decided on the float64 quotient of two exactly represented integers, as the reference does).

Differences from the reference that are deliberate: ``np.bool`` / ``np.int`` (removed from NumPy; the reference file
raises on NumPy >= 1.24) are not used, and equal confidences are ordered by arrival (stable sort) instead of NumPy's
unspecified quicksort order.
"""
from collections import defaultdict

import numpy as np

from utils import GRID, Point, Size, prop2abs


def APs2mAP(aps):
    """Mean of the per-class APs (average_precision.py:30-42)."""
    n = len(aps)
    return sum(aps.values()) / n if n else 0


def _requantise(box_int):
    """[n,4] (xmin, xmax, ymin, ymax) ints of normalised boxes -> what utils.prop2abs(abs2prop(.)) yields: the reference
    stores detections as proportional Box tuples and converts them back on the 1000 grid (average_precision.py:75),
    which is not always a round trip (float64 truncation)."""
    b = np.asarray(box_int, np.float64).reshape(-1, 4)
    w, h = b[:, 1] - b[:, 0], b[:, 3] - b[:, 2]
    cx, cy = (b[:, 0] + w / 2) / GRID.w, (b[:, 2] + h / 2) / GRID.h
    sw, sh = w / GRID.w, h / GRID.h
    hw, hh = sw * GRID.w / 2, sh * GRID.h / 2
    px, py = cx * GRID.w, cy * GRID.h
    return np.stack([np.trunc(px - hw), np.trunc(px + hw), np.trunc(py - hh), np.trunc(py + hh)], axis=1).astype(np.int64)


class APCalculator:
    """Average precision of object detections as used by the PASCAL VOC 2007 challenge."""

    def __init__(self, minoverlap=0.5):
        self.minoverlap = minoverlap
        self.clear()

    def clear(self):
        """Forget every detection and ground-truth box added so far (average_precision.py:184-192)."""
        self._det = defaultdict(lambda: ([], [], []))       # label -> (boxes [4] int, confidences, sample ids)
        self._gt = []                                       # per sample: [(label, (xmin, xmax, ymin, ymax))]

    # ---- reference-shaped input: Box tuples ----
    def add_detections(self, gt_boxes, boxes):
        """gt_boxes: the sample's ground-truth ``Box`` list; boxes: [(confidence, Box)] as decode_boxes /
        suppress_overlaps return them (average_precision.py:65-81)."""
        sid = len(self._gt)
        self._gt.append([(b.label, prop2abs(b.center, b.size, GRID)) for b in gt_boxes])
        for conf, box in boxes:
            d = self._det[box.label]
            d[0].append(prop2abs(box.center, box.size, GRID)); d[1].append(conf); d[2].append(sid)

    # ---- GPU-shaped input: integer rows of the decode + NMS kernels ----
    def add_detections_batch(self, gt_boxes_batch, dets, counts, lid2name=None):
        """gt_boxes_batch: per image an array [G,5] (labelid, cx, cy, w, h) or a ``Box`` list; dets [B,cap,8] int32 and
        counts [B,2] from ``ssdb.decode_nms_host`` (row = confidence bits, labelid, xmin, xmax, ymin, ymax, anchor, rank).
        Labels are ``lid2name[labelid]`` when a map is given, else the integer ids."""
        name = (lambda i: lid2name[int(i)]) if lid2name else int
        for b, gts in enumerate(gt_boxes_batch):
            sid = len(self._gt)
            rows = []
            for g in gts:
                if hasattr(g, 'center'):
                    rows.append((g.label, prop2abs(g.center, g.size, GRID)))
                else:
                    rows.append((name(g[0]), prop2abs(Point(float(g[1]), float(g[2])), Size(float(g[3]), float(g[4])), GRID)))
            self._gt.append(rows)
            n = int(counts[b, 0])
            if n == 0:
                continue
            d = np.asarray(dets[b, :n])
            conf = d[:, 0].astype(np.int32).view(np.float32)
            boxes = _requantise(d[:, 2:6])
            for k in range(n):
                e = self._det[name(d[k, 1])]
                e[0].append(tuple(boxes[k])); e[1].append(conf[k]); e[2].append(sid)

    def compute_aps(self):
        """{label: AP} over the labels present in the ground truth (average_precision.py:84-181)."""
        counts = defaultdict(int)
        gt_map = defaultdict(dict)
        for sid, boxes in enumerate(self._gt):
            per = defaultdict(list)
            for label, b in boxes:
                counts[label] += 1
                per[label].append(b)
            for label, v in per.items():
                gt_map[label][sid] = (np.array(v, np.int64).reshape(-1, 4), np.zeros(len(v), bool))
        aps = {}
        for label, per_sample in gt_map.items():
            boxes, confs, sids = self._det[label] if label in self._det else ([], [], [])
            n = len(confs)
            params = np.array(boxes, np.int64).reshape(-1, 4)
            confs = np.array(confs, np.float32)
            order = np.argsort(-confs, kind='stable')
            tp = np.zeros(n)
            for rank, i in enumerate(order):
                entry = per_sample.get(sids[i])
                if entry is None:
                    continue
                gt, matched = entry
                box = params[i]
                w = np.maximum(0, np.minimum(box[1], gt[:, 1]) - np.maximum(box[0], gt[:, 0]) + 1)
                h = np.maximum(0, np.minimum(box[3], gt[:, 3]) - np.maximum(box[2], gt[:, 2]) + 1)
                inter = w * h
                union = (box[1] - box[0] + 1) * (box[3] - box[2] + 1) + (gt[:, 1] - gt[:, 0] + 1) * (gt[:, 3] - gt[:, 2] + 1) - inter
                iou = inter / union
                j = int(np.argmax(iou))
                if iou[j] < self.minoverlap or matched[j]:
                    continue
                matched[j] = True
                tp[rank] = 1
            tps = np.cumsum(tp)
            fps = np.cumsum(1 - tp)
            recall = tps / counts[label]
            prec = tps / (tps + fps) if n else tps
            ap = 0.0
            for r in np.arange(0, 1.1, 0.1):
                sel = prec[recall >= r]
                if len(sel) > 0:
                    ap += np.amax(sel)
            aps[label] = ap / 11.
        return aps
