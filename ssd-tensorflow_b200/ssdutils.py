"""Box geometry, anchors, matching, decoding and NMS of the SSD hot path.

Same public names, argument order and defaults as the reference's
``ssdutils.py`` -- ``SSD_PRESETS``, ``get_preset_by_name`` (and the
``get_preset`` spelling BASELINE.json uses), ``get_anchors_for_preset``,
``anchors2array``, ``box2array``, ``jaccard_overlap``, ``compute_overlap``,
``compute_location``, ``decode_location``, ``decode_boxes``,
``non_maximum_suppression``, ``suppress_overlaps`` -- but everything that
loops over anchors runs on the B200 through libssd_b200 (``ssdb``): anchor
matching, decode + top-k and class-wise NMS.  There is no NumPy fallback; the
calls raise ``ssdb.SSDBError`` when the library or the GPU is missing.

The static configuration (presets, the anchor list) is host data computed
once, exactly as in the reference (ssdutils.py:32-117).
"""
from collections import namedtuple
from math import exp, log, sqrt

import numpy as np

import ssdb
from utils import GRID, Box, Overlap, Point, Score, Size, abs2prop, prop2abs

SSDMap = namedtuple('SSDMap', ['size', 'scale', 'aspect_ratios'])
SSDPreset = namedtuple('SSDPreset', ['name', 'image_size', 'maps', 'extra_scale', 'num_anchors'])
Anchor = namedtuple('Anchor', ['center', 'size', 'x', 'y', 'scale', 'map'])


def _preset(name, side, maps, extra_scale, num_anchors):
    return SSDPreset(name=name, image_size=Size(side, side),
                     maps=[SSDMap(Size(m, m), s, list(r)) for m, s, r in maps],
                     extra_scale=extra_scale, num_anchors=num_anchors)


_R2 = (2, 0.5)
_R4 = (2, 3, 0.5, 1. / 3.)
SSD_PRESETS = {
    'vgg300': _preset('vgg300', 300, [(38, 0.1, _R2), (19, 0.2, _R4), (10, 0.375, _R4), (5, 0.55, _R4),
                                     (3, 0.725, _R2), (1, 0.9, _R2)], 1.075, 8732),
    'vgg512': _preset('vgg512', 512, [(64, 0.07, _R2), (32, 0.15, _R4), (16, 0.3, _R4), (8, 0.45, _R4),
                                     (4, 0.6, _R4), (2, 0.75, _R2), (1, 0.9, _R2)], 1.05, 24564),
}


def get_preset_by_name(pname):
    """ssdutils.py:70-73 -- RuntimeError for an unknown preset, like the reference."""
    if pname not in SSD_PRESETS:
        raise RuntimeError('No such preset: ' + pname)
    return SSD_PRESETS[pname]


get_preset = get_preset_by_name   # the name BASELINE.json's north_star uses


def get_anchors_for_preset(preset):
    """Default boxes in the reference's order: map -> box type [ratio 1, aspect
    ratios..., s'-square] -> row -> column (ssdutils.py:76-117)."""
    anchors = []
    nmaps = len(preset.maps)
    for k, m in enumerate(preset.maps):
        s = m.scale
        shapes = [(s * sqrt(r), s / sqrt(r)) for r in [1] + list(m.aspect_ratios)]
        nxt = preset.maps[k + 1].scale if k + 1 < nmaps else preset.extra_scale
        shapes.append((sqrt(s * nxt),) * 2)
        fk = m.size[0]
        for (w, h) in shapes:
            for j in range(fk):
                cy = (j + 0.5) / float(fk)
                for i in range(fk):
                    anchors.append(Anchor(Point((i + 0.5) / float(fk), cy), Size(w, h), i, j, s, k))
    return anchors


def anchors_as_array(anchors):
    """[A,4] float64 (cx, cy, w, h): the form the GPU entry points take."""
    if isinstance(anchors, np.ndarray):
        return np.ascontiguousarray(anchors, np.float64)
    return np.array([[a.center.x, a.center.y, a.size.w, a.size.h] for a in anchors], np.float64)


def anchors2array(anchors, img_size):
    """Absolute (xmin, xmax, ymin, ymax) per anchor as float64 (ssdutils.py:120-130)."""
    arr = np.zeros((len(anchors), 4))
    for i, a in enumerate(anchors):
        arr[i] = prop2abs(a.center, a.size, img_size)
    return arr


def box2array(box, img_size):
    return np.array(prop2abs(box.center, box.size, img_size))


# ------------------------------------------------------------------ matching (GPU)
def create_labels(gt_boxes_per_image, anchors, num_classes):
    """Batched LabelCreatorTransform (transforms.py:72-114) on the GPU.

    gt_boxes_per_image: list (one entry per image) of lists of ``Box``.
    Returns (labels [B, A, num_classes+5] float32, match [B, A] int32)."""
    anc = anchors_as_array(anchors)
    B = len(gt_boxes_per_image)
    G = max(1, max((len(b) for b in gt_boxes_per_image), default=1))
    gt = np.zeros((B, G, 5), np.float64)
    cnt = np.zeros(B, np.int32)
    for i, boxes in enumerate(gt_boxes_per_image):
        cnt[i] = len(boxes)
        for j, b in enumerate(boxes):
            gt[i, j] = (b.labelid, b.center.x, b.center.y, b.size.w, b.size.h)
    match, labels = ssdb.match_anchors_host(gt, cnt, anc, num_classes)
    return labels, match


def jaccard_overlap(box_arr, anchors_arr):
    """Inclusive-pixel IoU of one absolute box against absolute anchors
    (ssdutils.py:138-152).  A per-GT helper of the reference's matcher; the
    product matcher is ``create_labels`` / ``ssdb_match_anchors`` -- this
    vectorised form only exists so that callers of ``compute_overlap`` keep working."""
    a = np.asarray(anchors_arr, np.float64)
    b = np.asarray(box_arr, np.float64)
    iw = np.maximum(0, np.minimum(b[1], a[:, 1]) - np.maximum(b[0], a[:, 0]) + 1)
    ih = np.maximum(0, np.minimum(b[3], a[:, 3]) - np.maximum(b[2], a[:, 2]) + 1)
    inter = iw * ih
    union = (b[1] - b[0] + 1) * (b[3] - b[2] + 1) + (a[:, 1] - a[:, 0] + 1) * (a[:, 3] - a[:, 2] + 1) - inter
    return inter / union


def compute_overlap(box_arr, anchors_arr, threshold):
    """Overlap(best, good) of one box against all anchors (ssdutils.py:155-170)."""
    iou = jaccard_overlap(box_arr, anchors_arr)
    good = [Score(i, iou[i]) for i in np.nonzero(iou > threshold)[0]]
    b = int(np.argmax(iou))
    return Overlap(Score(b, iou[b]) if iou[b] > threshold else None, good)


def compute_location(box, anchor):
    """Offset encoding with variances 0.1 / 0.2 (ssdutils.py:173-179)."""
    return np.array([(box.center.x - anchor.center.x) / anchor.size.w * 10,
                     (box.center.y - anchor.center.y) / anchor.size.h * 10,
                     log(box.size.w / anchor.size.w) * 5,
                     log(box.size.h / anchor.size.h) * 5])


def decode_location(box, anchor):
    """Inverse of compute_location for ONE box (ssdutils.py:182-189); like the
    reference it clamps the offsets to <= 100 in place."""
    box[box > 100] = 100
    x = box[0] / 10 * anchor.size.w + anchor.center.x
    y = box[1] / 10 * anchor.size.h + anchor.center.y
    return Point(x, y), Size(exp(box[2] / 5) * anchor.size.w, exp(box[3] / 5) * anchor.size.h)


# ------------------------------------------------------------------ decode + NMS (GPU)
def _boxes_from_rows(rows, lid2name):
    out = []
    for r in rows:
        conf = np.array(r[0], np.int32).view(np.float32)[()]
        cid = int(r[1])
        center, size = abs2prop(int(r[2]), int(r[3]), int(r[4]), int(r[5]), GRID)
        out.append((conf, Box(lid2name.get(cid), np.int64(cid), center, size)))
    return out


def detect_batch(pred, anchors, confidence_threshold=0.01, lid2name={}, detections_cap=200, overlap_threshold=0.45):
    """decode_boxes + suppress_overlaps for a whole batch in ONE kernel launch.

    pred: [B, A, C+5] float32.  Returns a list (per image) of the reference's
    ``[(confidence, Box), ...]`` lists, in the reference's output order."""
    pred = np.asarray(pred, np.float32)
    if pred.ndim == 2:
        pred = pred[None]
    dets, counts = ssdb.decode_nms_host(pred, anchors_as_array(anchors), confidence_threshold, detections_cap,
                                        overlap_threshold)
    return [_boxes_from_rows(dets[i, :counts[i, 0]], lid2name) for i in range(pred.shape[0])]


def detect_batch_rows(pred, anchors, confidence_threshold=0.01, detections_cap=200, overlap_threshold=0.45):
    """decode_boxes + suppress_overlaps for a batch, as the kernels' integer rows: (dets [B,cap,8] int32, counts [B,2])
    with row = (confidence bits, labelid, xmin, xmax, ymin, ymax on the 1000 grid, anchor, confidence rank) --
    what ``average_precision.APCalculator.add_detections_batch`` consumes without building Box tuples."""
    pred = np.asarray(pred, np.float32)
    if pred.ndim == 2:
        pred = pred[None]
    return ssdb.decode_nms_host(pred, anchors_as_array(anchors), confidence_threshold, detections_cap, overlap_threshold)


def decode_boxes(pred, anchors, confidence_threshold=0.01, lid2name={}, detections_cap=200):
    """Decode boxes from one image's predictions (ssdutils.py:192-229), on the GPU.

    Returns ``[(confidence, Box), ...]`` in descending confidence.  The NMS
    threshold is set above 1 so that nothing is suppressed; the rows come back
    class-grouped and are re-ordered by their confidence rank.  Ties in confidence
    go to the lower anchor index (NumPy's argsort leaves them unspecified)."""
    pred = np.asarray(pred, np.float32)
    dets, counts = ssdb.decode_nms_host(pred[None], anchors_as_array(anchors), confidence_threshold, detections_cap, 2.0)
    rows = dets[0, :counts[0, 0]]
    rows = rows[np.argsort(rows[:, 7], kind='stable')]
    return _boxes_from_rows(rows, lid2name)


def non_maximum_suppression(boxes, overlap_threshold):
    """Greedy NMS of one class's ``(confidence, Box)`` list (ssdutils.py:232-307).
    Runs through the same GPU kernel as ``suppress_overlaps``."""
    return _nms_gpu(boxes, overlap_threshold, single_class=True)


def suppress_overlaps(boxes):
    """Class-wise NMS at IoU 0.45 (ssdutils.py:310-318)."""
    return _nms_gpu(boxes, 0.45, single_class=False)


def _nms_gpu(boxes, thr, single_class):
    """Host side: the reference's own re-quantisation of each box through prop2abs on the
    1000x1000 grid (float64 scalars, a handful of boxes); device side: sort, greedy sweep
    and output ordering (ssdb_nms_host)."""
    if not boxes:
        return []
    abs_boxes = np.array([prop2abs(b[1].center, b[1].size, GRID) for b in boxes], np.int32)
    ids = np.zeros(len(boxes), np.int32) if single_class else np.array([int(b[1].labelid) for b in boxes], np.int32)
    conf = np.array([b[0] for b in boxes], np.float32)
    keep = ssdb.nms_host(abs_boxes, ids, conf, thr)
    return [boxes[int(i)] for i in keep]
