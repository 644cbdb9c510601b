"""CPU proofs-by-test of the arithmetic shortcuts the CUDA kernels take on the box path.

The reference decides everything with rounded float64 divisions of integer-valued operands
(ssdutils.py:138-170 jaccard_overlap / compute_overlap, :232-307 non_maximum_suppression).  The kernels
(csrc/loss.cu iou_int / iu_greater / match_one_int, csrc/detect.cu lazy greedy NMS) avoid float64:

  * IoU > 0.5            <=>  2*I > U                       (integers)
  * IoU_1 > IoU_2        <=>  I_1*U_2 > I_2*U_1             (int64 cross-multiplication)
  * IoU > thr            decided by an fp32 filter, float64 division only within 1e-6 of the threshold

Here the device logic is restated in NumPy with the same dtypes and checked against the reference's float64
expressions on random, exhaustive-small and adversarial inputs.
"""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

MAX_U = 2 * 1001 * 1001          # union of two boxes on the 1000x1000 grid (coordinates may poke out a little)


def _pairs(rng, n):
    u = rng.integers(1, MAX_U, n, dtype=np.int64)
    i = (rng.random(n) * (u + 1)).astype(np.int64)
    return np.minimum(i, u), u


def test_half_threshold_is_integer_compare():
    rng = np.random.default_rng(0)
    i, u = _pairs(rng, 2_000_000)
    # adversarial: exactly one half, and one unit either side
    ue = rng.integers(1, MAX_U // 2, 200_000, dtype=np.int64) * 2
    i = np.concatenate([i, ue // 2, ue // 2 + 1, np.maximum(ue // 2 - 1, 0)])
    u = np.concatenate([u, ue, ue, ue])
    ref = (i.astype(np.float64) / u.astype(np.float64)) > 0.5
    assert np.array_equal(ref, 2 * i > u)


def test_cross_multiplication_orders_like_float64_division():
    rng = np.random.default_rng(1)
    i1, u1 = _pairs(rng, 2_000_000)
    i2, u2 = _pairs(rng, 2_000_000)
    # adversarial: equal ratios with different denominators, and neighbours of equal ratios
    k = rng.integers(1, 1000, 300_000, dtype=np.int64)
    a = rng.integers(0, 1000, 300_000, dtype=np.int64)
    b = rng.integers(1, 1000, 300_000, dtype=np.int64)
    a = np.minimum(a, b)
    i1 = np.concatenate([i1, a * k, a * k + 1]); u1 = np.concatenate([u1, b * k, b * k])
    i2 = np.concatenate([i2, a, a]); u2 = np.concatenate([u2, b, b])
    q1 = i1.astype(np.float64) / u1.astype(np.float64)
    q2 = i2.astype(np.float64) / u2.astype(np.float64)
    assert np.array_equal(q1 > q2, i1 * u2 > i2 * u1)
    assert np.array_equal(q1 == q2, i1 * u2 == i2 * u1)          # ties: first maximum / earlier GT keeps


def _filter_decision(inter, uni, thr):
    """detect.cu, lazy greedy NMS: the hit test of one (kept box, later class mate) pair."""
    thr_f = np.float32(thr)
    fi = inter.astype(np.float32)
    lim = thr_f * uni.astype(np.float32)
    exact = (inter.astype(np.float64) / uni.astype(np.float64)) > thr
    sure_hit = (fi > lim * np.float32(1.000001)) & (lim >= 0)
    sure_miss = fi < lim * np.float32(0.999999)
    out = np.where(inter == 0, 0.0 > thr, np.where(sure_hit, True, np.where(sure_miss, False, exact)))
    undecided = ~(inter == 0) & ~sure_hit & ~sure_miss
    return out, exact, undecided


@pytest.mark.parametrize('thr', [0.45, 0.5, 0.3, 0.7, 0.05, 0.999, 0.0])
def test_fp32_filter_equals_float64_division(thr):
    rng = np.random.default_rng(int(thr * 1000) + 7)
    i, u = _pairs(rng, 1_000_000)
    # adversarial: the integers closest to thr * uni
    ua = rng.integers(1, MAX_U, 500_000, dtype=np.int64)
    near = np.floor(thr * ua).astype(np.int64)
    i = np.concatenate([i, near, near + 1, np.maximum(near - 1, 0), np.zeros(1000, np.int64)])
    u = np.concatenate([u, ua, ua, ua, rng.integers(1, MAX_U, 1000, dtype=np.int64)])
    i = np.minimum(i, u)
    got, exact, undecided = _filter_decision(i, u, thr)
    assert np.array_equal(got, exact)
    # the float64 division must stay a rare path for the kernel to be cheap
    assert undecided[:1_000_000].mean() < 1e-3


@settings(max_examples=300, deadline=None)
@given(st.integers(1, 1001), st.integers(1, 1001), st.integers(1, 1001), st.integers(1, 1001),
       st.integers(-40, 1040), st.integers(-40, 1040), st.integers(-40, 1040), st.integers(-40, 1040))
def test_inclusive_pixel_iou_int_matches_reference_formula(w1, h1, w2, h2, x1, y1, x2, y2):
    """iou_int (loss.cu) vs jaccard_overlap (ssdutils.py:138-152): +1 inclusive widths, float64 I / U."""
    a = np.array([x1, x1 + w1 - 1, y1, y1 + h1 - 1], np.int64)          # xmin, xmax, ymin, ymax
    b = np.array([x2, x2 + w2 - 1, y2, y2 + h2 - 1], np.int64)
    # reference formula on float64 arrays
    af, bf = a.astype(np.float64), b.astype(np.float64)
    area_a = (af[1] - af[0] + 1) * (af[3] - af[2] + 1)
    area_b = (bf[1] - bf[0] + 1) * (bf[3] - bf[2] + 1)
    iw = max(min(af[1], bf[1]) - max(af[0], bf[0]) + 1, 0.0)
    ih = max(min(af[3], bf[3]) - max(af[2], bf[2]) + 1, 0.0)
    ref = iw * ih / (area_a + area_b - iw * ih)
    # device formula in int32
    i32 = np.int32
    ia = i32(a[1] - a[0] + 1) * i32(a[3] - a[2] + 1); ib = i32(b[1] - b[0] + 1) * i32(b[3] - b[2] + 1)
    jw = max(int(min(a[1], b[1]) - max(a[0], b[0]) + 1), 0); jh = max(int(min(a[3], b[3]) - max(a[2], b[2]) + 1), 0)
    inter = jw * jh; uni = int(ia) + int(ib) - inter
    assert uni > 0 and uni < 2 ** 31
    assert ref == inter / uni
    assert (ref > 0.5) == (2 * inter > uni)
    assert (ref > 0.45) == bool(_filter_decision(np.array([inter]), np.array([uni]), 0.45)[0][0])


def test_half_over_1000_table_equals_the_reference_divisions():
    """detect.cu decode: (x0 + (x1-x0)/2) / 1000 and (x1-x0) / 1000 (utils.abs2prop, utils.py:85-97) come from a table of
    (k / 2) / 1000, k = x0 + x1 resp. 2 * (x1 - x0); and x * 0.5 replaces x / 2."""
    table = (np.arange(2000, dtype=np.float64) / 2.0) / 1000.0
    x0, x1 = np.meshgrid(np.arange(1000), np.arange(1000), indexing='ij')
    ok = x0 <= x1
    x0, x1 = x0[ok].astype(np.float64), x1[ok].astype(np.float64)
    bw = x1 - x0
    assert np.array_equal((x0 + bw / 2.0) / 1000.0, table[(x0 + x1).astype(np.int64)])
    assert np.array_equal(bw / 1000.0, table[(2 * bw).astype(np.int64)])
    rng = np.random.default_rng(3)
    v = rng.standard_normal(1_000_000) * 10.0 ** rng.integers(-5, 6, 1_000_000)
    assert np.array_equal(v / 2.0, v * 0.5)


def _radix_select_model(keys, valid, k, small=256):
    """Python model of csrc/select.cuh radix_select_kth: two 8-bit passes, then counting among the survivors when they are
    few, else two more passes.  Returns (prefix, remaining)."""
    keys = keys.astype(np.uint64)
    prefix, mask, remaining = 0, 0, k
    for p in range(4):
        if p == 2:
            surv = keys[valid & ((keys & mask) == prefix)]
            if surv.size <= small:
                for me in surv:
                    gt, eq = int((surv > me).sum()), int((surv == me).sum())
                    if gt < remaining <= gt + eq:
                        return int(me), remaining - gt
                raise AssertionError('no survivor satisfies the rank condition')
        shift = 24 - 8 * p
        part = valid & ((keys & mask) == prefix)
        hist = np.bincount(((keys[part] >> shift) & 255).astype(np.int64), minlength=256)
        cum = 0
        for b in range(255, -1, -1):
            if cum + hist[b] >= remaining:
                break
            cum += hist[b]
        prefix |= b << shift; mask |= 255 << shift; remaining -= cum
    return prefix, remaining


@pytest.mark.parametrize('seed', range(6))
def test_radix_select_model_finds_kth_largest_with_ties(seed):
    rng = np.random.default_rng(seed)
    n = 8732
    if seed % 3 == 0:        # heavy ties (few distinct values), like an all-zero head output
        keys = rng.integers(0x3f000000, 0x3f000010, n, dtype=np.int64)
    elif seed % 3 == 1:      # positive floats in a narrow range, like cross entropies
        keys = (rng.random(n).astype(np.float32) * 4 + 1).view(np.uint32).astype(np.int64) | 0x80000000
    else:                    # survivors of the 16-bit prefix exceed the counting path: falls back to four passes
        keys = 0x3f800000 + rng.integers(0, 1 << 14, n, dtype=np.int64)
    valid = rng.random(n) < 0.9
    for k in (1, 7, 200, int(valid.sum())):
        prefix, remaining = _radix_select_model(keys, valid, k)
        srt = np.sort(keys[valid])[::-1]
        assert prefix == srt[k - 1]
        assert remaining == k - int((keys[valid] > prefix).sum()) and remaining >= 1


def _lazy_nms_model(cls, nms_boxes, thr, P=256):
    """Python model of the fast path of csrc/detect.cu decode_nms_kernel after the sort: per-class member masks (cmask),
    lazy greedy pass over compacted member lists with the fp32-filtered IoU test (rmask), then the popcount ranking
    (kcnt / kbase).  Returns the candidate positions in output order."""
    n = len(cls)
    W = 8
    cmask = np.zeros((64, W), np.uint32); rmask = np.zeros((64, W), np.uint32)
    first_pos = np.full(64, 0x7fffffff, np.int64)
    for i, c in enumerate(cls):
        cmask[c, i >> 5] |= np.uint32(1 << (i & 31)); first_pos[c] = min(first_pos[c], i)
    def hit(i, j):
        a, b = nms_boxes[i], nms_boxes[j]
        iw = max(int(min(a[1], b[1]) - max(a[0], b[0]) + 1), 0); ih = max(int(min(a[3], b[3]) - max(a[2], b[2]) + 1), 0)
        inter = iw * ih
        uni = int((a[1] - a[0] + 1) * (a[3] - a[2] + 1) + (b[1] - b[0] + 1) * (b[3] - b[2] + 1)) - inter
        return bool(_filter_decision(np.array([inter]), np.array([uni]), thr)[0][0])
    for c in range(64):
        ml = [w * 32 + k for w in range(W) for k in range(32) if (int(cmask[c, w]) >> k) & 1]
        dead = [False] * len(ml)
        for pos in range(len(ml)):
            if dead[pos]:
                continue                                  # only alive (= kept) members test their later class mates
            for q in range(pos + 1, len(ml)):
                if not dead[q] and hit(ml[pos], ml[q]):
                    dead[q] = True
        for q, j in enumerate(ml):
            if dead[q]:
                rmask[c, j >> 5] |= np.uint32(1 << (j & 31))
    alive = cmask & ~rmask
    kcnt = np.array([sum(bin(int(v)).count('1') for v in alive[c]) for c in range(64)])
    kbase = np.array([sum(kcnt[c2] for c2 in range(64) if first_pos[c2] < first_pos[c]) for c in range(64)])
    out = [None] * int(kcnt.sum())
    for i in range(n):
        c, w = cls[i], i >> 5
        if not (int(alive[c, w]) >> (i & 31)) & 1:
            continue
        rnk = kbase[c] + bin(int(alive[c, w]) & ((1 << (i & 31)) - 1)).count('1') + sum(bin(int(alive[c, w2])).count('1') for w2 in range(w))
        out[rnk] = i
    return out


@pytest.mark.parametrize('seed', range(5))
def test_lazy_per_class_nms_model_equals_reference_order(seed):
    """The lazy per-class pass + popcount ranking produce the reference's kept set AND output order
    (ssdutils.py:232-318: classes by first appearance, confidence-descending inside a class)."""
    import box_oracle as bo
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(1, 201))
    centers = rng.uniform(100, 900, (6, 2))
    which = rng.integers(0, 6, n)
    cxy = centers[which] + rng.normal(0, 25, (n, 2))
    wh = rng.uniform(40, 300, (n, 2))
    box = np.stack([cxy[:, 0] - wh[:, 0] / 2, cxy[:, 0] + wh[:, 0] / 2, cxy[:, 1] - wh[:, 1] / 2, cxy[:, 1] + wh[:, 1] / 2], axis=1)
    box = np.clip(np.trunc(box), 0, 999).astype(np.int64)
    cls = rng.integers(0, 1 + seed * 4, n)                  # seed 0: a single class holds every candidate
    cand = dict(idx=np.arange(n), cls=cls, nms=box)
    want = bo.nms_classwise(cand, 0.45)
    got = _lazy_nms_model(cls, box, 0.45)
    assert list(want) == got
