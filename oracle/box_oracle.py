"""CPU oracle for the box half of the SSD hot path -- TEST INFRASTRUCTURE ONLY.

A NumPy restatement of the reference's anchor generation, anchor matching
(label creation), box decoding and class-wise greedy NMS.  Nothing under
``oracle/`` is product code: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import it, and only as the checker.

Parity status: PINNED.  Every function here is checked against the reference's
own code, imported from /root/reference under a stub ``tensorflow`` module
(``oracle/ref_loader.py``), by ``tests/test_oracle_vs_reference.py`` in this
container, and against the committed fixtures ``tests/golden/*.npz`` that
``oracle/make_golden.py`` generated from that real reference code.

Each function cites the reference lines it restates.  The arithmetic follows
what the reference does under this container's NumPy 2.x scalar promotion
(NEP 50) -- e.g. decoded centres are float32 while decoded sizes are float64
 -- because that is the behaviour the committed
goldens were produced with.
"""
import math

import numpy as np

GRID = 1000  # virtual image side used for match / NMS (transforms.py:67, ssdutils.py:241)

# (map size, scale, aspect ratios) per preset -- ssdutils.py:36-62
PRESETS = {
    'vgg300': dict(image=300, extra_scale=1.075, num_anchors=8732, maps=[
        (38, 0.1, [2, 0.5]), (19, 0.2, [2, 3, 0.5, 1. / 3.]),
        (10, 0.375, [2, 3, 0.5, 1. / 3.]), (5, 0.55, [2, 3, 0.5, 1. / 3.]),
        (3, 0.725, [2, 0.5]), (1, 0.9, [2, 0.5])]),
    'vgg512': dict(image=512, extra_scale=1.05, num_anchors=24564, maps=[
        (64, 0.07, [2, 0.5]), (32, 0.15, [2, 3, 0.5, 1. / 3.]),
        (16, 0.3, [2, 3, 0.5, 1. / 3.]), (8, 0.45, [2, 3, 0.5, 1. / 3.]),
        (4, 0.6, [2, 3, 0.5, 1. / 3.]), (2, 0.75, [2, 0.5]), (1, 0.9, [2, 0.5])]),
}


def anchors(preset_name):
    """Default boxes as an [A,4] float64 array (cx, cy, w, h), proportional.

    Restates get_anchors_for_preset (ssdutils.py:76-117): per map, box types
    are [ratio 1, *aspect ratios, s'-square]; index order is map -> box type ->
    row (y) -> column (x).
    """
    p = PRESETS[preset_name]
    maps = p['maps']
    out = []
    for k, (fk, s, ratios) in enumerate(maps):
        nxt = maps[k + 1][1] if k + 1 < len(maps) else p['extra_scale']
        sizes = []
        for r in [1] + list(ratios):
            q = math.sqrt(r)
            sizes.append((s * q, s / q))
        sp = math.sqrt(s * nxt)
        sizes.append((sp, sp))
        for (w, h) in sizes:
            for j in range(fk):
                cy = (j + 0.5) / float(fk)
                for i in range(fk):
                    cx = (i + 0.5) / float(fk)
                    out.append((cx, cy, w, h))
    arr = np.array(out, dtype=np.float64)
    assert arr.shape[0] == p['num_anchors']
    return arr


def prop2abs_int(cx, cy, w, h, side=GRID):
    """float64 proportional box(es) -> truncated integer bounds
    (xmin, xmax, ymin, ymax); restates prop2abs (utils.py:100-108)."""
    cx = np.asarray(cx, np.float64); cy = np.asarray(cy, np.float64)
    hw = np.asarray(w, np.float64) * side / 2
    hh = np.asarray(h, np.float64) * side / 2
    px = cx * side
    py = cy * side
    t = lambda v: np.trunc(v).astype(np.int64)
    return t(px - hw), t(px + hw), t(py - hh), t(py + hh)


def anchors_abs(anc):
    """[A,4] int64 (xmin,xmax,ymin,ymax) on the 1000-grid; anchors2array (ssdutils.py:120-130)."""
    return np.stack(prop2abs_int(anc[:, 0], anc[:, 1], anc[:, 2], anc[:, 3]), axis=1)


def iou_1000(box, anc_abs):
    """Inclusive-pixel IoU of one integer box against all anchors, float64;
    jaccard_overlap (ssdutils.py:138-152)."""
    a = anc_abs.astype(np.float64)
    b = [float(v) for v in box]
    area_a = (a[:, 1] - a[:, 0] + 1) * (a[:, 3] - a[:, 2] + 1)
    area_b = (b[1] - b[0] + 1) * (b[3] - b[2] + 1)
    iw = np.maximum(0, np.minimum(b[1], a[:, 1]) - np.maximum(b[0], a[:, 0]) + 1)
    ih = np.maximum(0, np.minimum(b[3], a[:, 3]) - np.maximum(b[2], a[:, 2]) + 1)
    inter = iw * ih
    return inter / (area_b + area_a - inter)


def match_anchors(gt, anc, anc_abs, threshold=0.5):
    """Anchor matching.  gt: [G,5] float64 rows (labelid, cx, cy, w, h).

    Returns match[A] int32 (-1 = background, else the GT row that owns the
    anchor).  Restates compute_overlap (ssdutils.py:155-170) plus the two-pass
    conflict resolution of LabelCreatorTransform / process_overlap
    (transforms.py:47-54,72-114): pass 1 every anchor with IoU > thr takes the
    GT with the strictly highest IoU (earlier GT wins ties); pass 2, with a
    fresh score table, every GT's arg-max anchor (first maximum, only if its
    IoU > thr) is handed to that GT, again strictly-higher-wins.
    """
    A = anc.shape[0]
    match = np.full(A, -1, np.int32)
    G = len(gt)
    ious = np.zeros((G, A))
    for g in range(G):
        ious[g] = iou_1000([int(v) for v in prop2abs_int(*gt[g, 1:5])], anc_abs)
    score = np.full(A, -1.0)
    for g in range(G):
        take = (ious[g] > threshold) & (ious[g] > score)
        match[take] = g
        score[take] = ious[g][take]
    score2 = {}
    for g in range(G):
        a = int(np.argmax(ious[g])) if A else 0
        s = ious[g, a]
        if not s > threshold:
            continue
        if a in score2 and score2[a] >= s:
            continue
        score2[a] = s
        match[a] = g
    return match


def encode_offsets(gt_row, anc_row):
    """compute_location (ssdutils.py:173-179): float64, variances 0.1 / 0.2."""
    _, bx, by, bw, bh = [float(v) for v in gt_row]
    ax, ay, aw, ah = [float(v) for v in anc_row]
    return ((bx - ax) / aw * 10, (by - ay) / ah * 10,
            math.log(bw / aw) * 5, math.log(bh / ah) * 5)


def make_labels(gt, anc, anc_abs, num_classes):
    """Dense label tensor [A, num_classes+5] float32 (transforms.py:72-114).

    Columns: one-hot object class (0..C-1), background flag (C), 4 offsets.
    Also returns the match vector."""
    A = anc.shape[0]
    match = match_anchors(gt, anc, anc_abs)
    vec = np.zeros((A, num_classes + 5), np.float32)
    vec[:, num_classes] = 1
    for a in np.nonzero(match >= 0)[0]:
        g = match[a]
        vec[a, :num_classes + 1] = 0
        vec[a, int(gt[g, 0])] = 1
        vec[a, num_classes + 1:] = encode_offsets(gt[g], anc[a])
    return vec, match


def _order_desc(conf):
    """Descending confidence, ties -> lower index first.  The reference uses
    np.argsort(...)[::-1] whose tie order is unspecified; this
    is the deterministic rule the CUDA kernel implements, and the golden
    generator rejects inputs where ties could matter."""
    idx = np.arange(conf.shape[0])
    return np.lexsort((idx, -conf.astype(np.float64)))


def decode_candidates(pred, anc, conf_thr=0.01, cap=200):
    """Arg-max class, top-`cap` by confidence, threshold, offset decode and
    1000-grid normalisation.  pred: [A, C+5] float32 (softmax scores | 4 offsets).

    Returns dict(idx, conf (float32), cls, box (int64 [n,4] xmin,xmax,ymin,ymax
    after normalize_box), nms (int64 [n,4] bounds as NMS re-derives them)).
    Restates decode_boxes / decode_location (ssdutils.py:182-229),
    normalize_box / prop2abs / abs2prop (utils.py:85-135) and the prop2abs call
    at ssdutils.py:243-249, with this container's NumPy-2 promotion: centre in
    float32, size in float64, `centre - half_size` rounded to float32.
    """
    pred = np.asarray(pred, np.float32)
    nc = pred.shape[1] - 4
    cls_all = np.argmax(pred[:, :nc - 1], axis=1)
    conf_all = pred[np.arange(pred.shape[0]), cls_all]
    order = _order_desc(conf_all)
    if cap is not None:
        order = order[:cap]
    below = np.nonzero(conf_all[order] < np.float32(conf_thr))[0]
    if below.size:
        order = order[:below[0]]
    n = order.shape[0]
    off = np.minimum(pred[order, nc:], np.float32(100)) if n else np.zeros((0, 4), np.float32)
    off = np.where(np.isnan(pred[order, nc:]), pred[order, nc:], off) if n else off
    a = anc[order]
    f32 = np.float32
    x = (off[:, 0] / f32(10)) * a[:, 0 + 2].astype(f32) + a[:, 0].astype(f32)
    y = (off[:, 1] / f32(10)) * a[:, 1 + 2].astype(f32) + a[:, 1].astype(f32)
    e2 = (off[:, 2] / f32(5)).astype(np.float64)
    e3 = (off[:, 3] / f32(5)).astype(np.float64)
    w = np.array([math.exp(v) for v in e2], np.float64) * a[:, 2]
    h = np.array([math.exp(v) for v in e3], np.float64) * a[:, 3]
    # normalize_box -> prop2abs on the 1000 grid (mixed f32 / f64, see docstring)
    px = x * f32(GRID)
    py = y * f32(GRID)
    hw = (w * GRID / 2).astype(f32)
    hh = (h * GRID / 2).astype(f32)
    t = lambda v: np.trunc(v.astype(np.float64)).astype(np.int64)
    xmin, xmax, ymin, ymax = t(px - hw), t(px + hw), t(py - hh), t(py + hh)
    xmin = np.maximum(xmin, 0); xmax = np.minimum(xmax, GRID - 1)
    ymin = np.maximum(ymin, 0); ymax = np.minimum(ymax, GRID - 1)
    xmin = np.minimum(xmin, xmax); ymin = np.minimum(ymin, ymax)
    box = np.stack([xmin, xmax, ymin, ymax], axis=1) if n else np.zeros((0, 4), np.int64)
    # abs2prop (float64) then the NMS stage's prop2abs (float64) -- does NOT always round-trip
    bw = (xmax - xmin).astype(np.float64); bh = (ymax - ymin).astype(np.float64)
    pcx = (xmin.astype(np.float64) + bw / 2) / GRID
    pcy = (ymin.astype(np.float64) + bh / 2) / GRID
    nms = np.stack(prop2abs_int(pcx, pcy, bw / GRID, bh / GRID), axis=1) if n else np.zeros((0, 4), np.int64)
    return dict(idx=order.astype(np.int64), conf=conf_all[order], cls=cls_all[order].astype(np.int64),
                box=box, nms=nms)


def nms_classwise(cand, iou_thr=0.45):
    """Class-wise greedy NMS.  Returns positions into `cand` of the kept boxes,
    in the reference's output order: classes by first appearance in confidence
    order, boxes by descending confidence within a class.  Restates
    non_maximum_suppression / suppress_overlaps (ssdutils.py:232-318)."""
    n = cand['idx'].shape[0]
    out = []
    seen = []
    for c in cand['cls']:
        if int(c) not in seen:
            seen.append(int(c))
    b = cand['nms'].astype(np.float64)
    area = (b[:, 1] - b[:, 0] + 1) * (b[:, 3] - b[:, 2] + 1) if n else np.zeros(0)
    for c in seen:
        members = [i for i in range(n) if int(cand['cls'][i]) == c]   # already conf-descending
        alive = list(members)
        while alive:
            i = alive.pop(0)
            out.append(i)
            rest = []
            for j in alive:
                iw = max(0.0, min(b[i, 1], b[j, 1]) - max(b[i, 0], b[j, 0]) + 1)
                ih = max(0.0, min(b[i, 3], b[j, 3]) - max(b[i, 2], b[j, 2]) + 1)
                inter = iw * ih
                if not (inter / (area[i] + area[j] - inter) > iou_thr):
                    rest.append(j)
            alive = rest
    return np.array(out, np.int64)


def detect(pred, anc, conf_thr=0.01, cap=200, iou_thr=0.45):
    """decode + suppress: rows [conf_bits(u32 view of f32), cls, xmin, xmax, ymin, ymax, anchor]."""
    cand = decode_candidates(pred, anc, conf_thr, cap)
    keep = nms_classwise(cand, iou_thr)
    rows = np.zeros((keep.shape[0], 7), np.int64)
    if keep.size:
        rows[:, 0] = cand['conf'][keep].view(np.uint32).astype(np.int64)
        rows[:, 1] = cand['cls'][keep]
        rows[:, 2:6] = cand['box'][keep]
        rows[:, 6] = cand['idx'][keep]
    return rows, cand
