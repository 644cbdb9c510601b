"""Geometry helpers and record types of the SSD hot path (host side).

Mirrors the public names of the reference's ``utils.py`` that the hot path
uses (reference ``utils.py:64-135``): the record types ``Label Size Point
Sample Box Score Overlap`` and the proportional <-> absolute box conversions.
Only the pure-Python scalar helpers live here; everything that touches more
than a handful of boxes runs on the GPU through ``ssdb`` (the C-ABI binding).

TensorFlow-session plumbing, drawing and TensorBoard summary classes of the
reference file are out of scope.
"""
import argparse
import math
from collections import namedtuple

Label = namedtuple('Label', ['name', 'color'])
Size = namedtuple('Size', ['w', 'h'])
Point = namedtuple('Point', ['x', 'y'])
Sample = namedtuple('Sample', ['filename', 'boxes', 'imgsize'])
Box = namedtuple('Box', ['label', 'labelid', 'center', 'size'])
Score = namedtuple('Score', ['idx', 'score'])
Overlap = namedtuple('Overlap', ['best', 'good'])

#: every box comparison of the path (match, NMS, AP) happens on this virtual grid
#: (reference transforms.py:67, ssdutils.py:241, utils.py:122)
GRID = Size(1000, 1000)


def str2bool(v):
    """CLI helper with the reference's accepted spellings (utils.py:73-82)."""
    s = v.lower()
    if s in ('yes', 'true', 't', 'y', '1'):
        return True
    if s in ('no', 'false', 'f', 'n', '0'):
        return False
    raise argparse.ArgumentTypeError('Boolean value expected.')


def abs2prop(xmin, xmax, ymin, ymax, imgsize):
    """Absolute min/max bounds -> proportional centre/size (utils.py:85-97)."""
    w = float(xmax - xmin)
    h = float(ymax - ymin)
    cx = (float(xmin) + w / 2) / imgsize.w
    cy = (float(ymin) + h / 2) / imgsize.h
    return Point(cx, cy), Size(w / imgsize.w, h / imgsize.h)


def prop2abs(center, size, imgsize):
    """Proportional centre/size -> absolute bounds, truncated toward zero by
    ``int()`` exactly like the reference (utils.py:100-108)."""
    hw = size.w * imgsize.w / 2
    hh = size.h * imgsize.h / 2
    cx = center.x * imgsize.w
    cy = center.y * imgsize.h
    return int(cx - hw), int(cx + hw), int(cy - hh), int(cy + hh)


def box_is_valid(box):
    """False when any coordinate is NaN/Inf (utils.py:111-115)."""
    return all(math.isfinite(v) for v in
               (box.center.x, box.center.y, box.size.w, box.size.h))


def normalize_box(box):
    """Clamp a box to the 1000x1000 grid and re-quantise it (utils.py:118-135)."""
    if not box_is_valid(box):
        return box
    xmin, xmax, ymin, ymax = prop2abs(box.center, box.size, GRID)
    xmin = max(xmin, 0)
    xmax = min(xmax, GRID.w - 1)
    ymin = max(ymin, 0)
    ymax = min(ymax, GRID.h - 1)
    xmin = min(xmin, xmax)
    ymin = min(ymin, ymax)
    center, size = abs2prop(xmin, xmax, ymin, ymax, GRID)
    return Box(box.label, box.labelid, center, size)
