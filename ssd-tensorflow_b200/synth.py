"""Seeded synthetic inputs for the benchmarks and parity tests.

Pure NumPy, no device work.  Every image uses its own generator
``numpy.random.default_rng(1234 + global_image_index)`` so shards of a
data-parallel batch are reproducible on any rank.
"""
import math

import numpy as np

BASE_SEED = 1234


def rng_for(image_index):
    return np.random.default_rng(BASE_SEED + int(image_index))


def images(first_index, count, side):
    """[count, side, side, 3] float32, U[0,255) -- what cv2 would hand the net."""
    out = np.empty((count, side, side, 3), np.float32)
    for i in range(count):
        out[i] = rng_for(first_index + i).random((side, side, 3), np.float32) * np.float32(255)
    return out


def gt_boxes(image_index, max_boxes=8, num_classes=20):
    """[G,5] float64 rows (labelid, cx, cy, w, h): G~U{1..max}, w,h~U[0.1,0.6],
    centre uniform such that the box stays inside the image."""
    r = np.random.default_rng(BASE_SEED + 7919 * 1000 + int(image_index))
    g = int(r.integers(1, max_boxes + 1))
    out = np.zeros((g, 5), np.float64)
    for k in range(g):
        w, h = r.uniform(0.1, 0.6, 2)
        cx = r.uniform(w / 2, 1 - w / 2)
        cy = r.uniform(h / 2, 1 - h / 2)
        out[k] = (int(r.integers(0, num_classes)), cx, cy, w, h)
    return out


def pack_gt(gts, max_boxes):
    """list of [G,5] -> ([B,max,5] float64 zero padded, [B] int32 counts)."""
    b = len(gts)
    arr = np.zeros((b, max_boxes, 5), np.float64)
    cnt = np.zeros(b, np.int32)
    for i, g in enumerate(gts):
        arr[i, :len(g)] = g
        cnt[i] = len(g)
    return arr, cnt


def _softmax(z):
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


def pred_uniform(image_index, num_anchors, num_classes=20):
    """Distribution U: logits ~ N(0, 3^2) -> softmax; offsets ~ N(0,1).  [A, C+5] float32."""
    r = np.random.default_rng(BASE_SEED + 104729 * 1000 + int(image_index))
    z = r.normal(0, 3, (num_anchors, num_classes + 1))
    loc = r.normal(0, 1, (num_anchors, 4)).astype(np.float32)
    return np.concatenate([_softmax(z), loc], axis=1)


def pred_clustered(image_index, anchors, num_classes=20, objects=8, return_objects=False):
    """Distribution C: `objects` boxes per image; anchors with IoU > 0.4 to an
    object get logit +6 on its class and offsets = encode(object) + N(0, 0.3^2);
    all others logit +6 on background.  anchors: [A,4] float64 (cx,cy,w,h).
    return_objects: also the [objects,5] (labelid, cx, cy, w, h) rows -- the ground truth of the image for AP evaluation."""
    r = np.random.default_rng(BASE_SEED + 15485863 * 100 + int(image_index))
    a = anchors
    n = a.shape[0]
    z = r.normal(0, 1, (n, num_classes + 1))
    loc = r.normal(0, 0.3, (n, 4))
    z[:, num_classes] += 6
    ax0, ax1 = a[:, 0] - a[:, 2] / 2, a[:, 0] + a[:, 2] / 2
    ay0, ay1 = a[:, 1] - a[:, 3] / 2, a[:, 1] + a[:, 3] / 2
    objs = np.zeros((objects, 5), np.float64)
    for k in range(objects):
        w, h = r.uniform(0.1, 0.6, 2)
        cx = r.uniform(w / 2, 1 - w / 2)
        cy = r.uniform(h / 2, 1 - h / 2)
        c = int(r.integers(0, num_classes))
        objs[k] = (c, cx, cy, w, h)
        iw = np.maximum(0, np.minimum(ax1, cx + w / 2) - np.maximum(ax0, cx - w / 2))
        ih = np.maximum(0, np.minimum(ay1, cy + h / 2) - np.maximum(ay0, cy - h / 2))
        inter = iw * ih
        iou = inter / (a[:, 2] * a[:, 3] + w * h - inter)
        hit = iou > 0.4
        z[hit, num_classes] -= 6
        z[hit, c] += 6
        loc[hit, 0] += (cx - a[hit, 0]) / a[hit, 2] * 10
        loc[hit, 1] += (cy - a[hit, 1]) / a[hit, 3] * 10
        loc[hit, 2] += np.log(w / a[hit, 2]) * 5
        loc[hit, 3] += np.log(h / a[hit, 3]) * 5
    pred = np.concatenate([_softmax(z), loc.astype(np.float32)], axis=1)
    return (pred, objs) if return_objects else pred
