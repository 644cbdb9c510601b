"""CPU: the C-ABI library loads and exports every symbol include/ssd_b200.h declares (no compute
without a GPU), the ctypes table covers the header, the product fails loudly without a GPU, and
the host-side geometry mirrors the reference surface."""
import os
import re

import numpy as np
import pytest

import box_oracle as bo
import ssdb
import ssdutils
import utils

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'ssd_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ssdb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    l = ssdb.lib()
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(l, s), 'libssd_b200.so does not export ' + s


def test_ctypes_table_matches_header():
    assert sorted(ssdb.SIGNATURES) == _header_symbols()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    assert ssdb.lib().ssdb_device_ok() != 0
    with pytest.raises(ssdb.SSDBError):
        ssdb.Net('vgg300', 20, 1)
    anc = ssdutils.get_anchors_for_preset(ssdutils.get_preset_by_name('vgg300'))
    with pytest.raises(ssdb.SSDBError):
        ssdutils.decode_boxes(np.zeros((8732, 25), np.float32), anc)
    with pytest.raises(ssdb.SSDBError):
        ssdutils.suppress_overlaps([(np.float32(0.9), utils.Box('a', 1, utils.Point(.5, .5), utils.Size(.1, .1)))])


def test_presets_and_anchor_list():
    with pytest.raises(RuntimeError):
        ssdutils.get_preset_by_name('vgg999')
    assert ssdutils.get_preset is ssdutils.get_preset_by_name
    for name in ('vgg300', 'vgg512'):
        p = ssdutils.get_preset_by_name(name)
        a = ssdutils.get_anchors_for_preset(p)
        assert len(a) == p.num_anchors
        arr = ssdutils.anchors_as_array(a)
        assert np.array_equal(arr, bo.anchors(name))                 # host config == oracle == reference fixture
        assert a[0].map == 0 and a[-1].map == len(p.maps) - 1 and a[1].x == 1 and a[1].y == 0
        assert np.array_equal(ssdutils.anchors2array(a[:50], utils.Size(1000, 1000)).astype(np.int64), bo.anchors_abs(arr[:50]))


def test_geometry_helpers():
    c, s = utils.abs2prop(100, 300, 50, 150, utils.Size(1000, 1000))
    assert (c.x, c.y, s.w, s.h) == (0.2, 0.1, 0.2, 0.1)
    assert utils.prop2abs(utils.Point(0.5, 0.5), utils.Size(0.25, 0.5), utils.Size(300, 300)) == (112, 187, 75, 225)
    assert utils.prop2abs(utils.Point(0.01, 0.01), utils.Size(0.1, 0.1), utils.Size(1000, 1000))[0] == -40
    b = utils.normalize_box(utils.Box('x', 3, utils.Point(0.99, 0.5), utils.Size(0.2, 2.0)))
    assert utils.prop2abs(b.center, b.size, utils.Size(1000, 1000))[1] <= 999
    nanbox = utils.Box('x', 3, utils.Point(float('nan'), 0.5), utils.Size(0.2, 0.2))
    assert utils.normalize_box(nanbox) is nanbox
    assert utils.str2bool('Yes') and not utils.str2bool('0')


def test_compute_overlap_and_location_helpers():
    a = ssdutils.get_anchors_for_preset(ssdutils.get_preset_by_name('vgg300'))
    arr = ssdutils.anchors2array(a, utils.Size(1000, 1000))
    box = utils.Box('x', 1, utils.Point(0.5, 0.5), utils.Size(0.3, 0.3))
    ov = ssdutils.compute_overlap(ssdutils.box2array(box, utils.Size(1000, 1000)), arr, 0.5)
    iou = bo.iou_1000([350, 650, 350, 650], bo.anchors_abs(bo.anchors('vgg300')))
    assert ov.best.idx == int(np.argmax(iou)) and [s.idx for s in ov.good] == list(np.nonzero(iou > 0.5)[0])
    loc = ssdutils.compute_location(box, a[ov.best.idx])
    assert np.allclose(loc, bo.encode_offsets((1, .5, .5, .3, .3), bo.anchors('vgg300')[ov.best.idx]), rtol=0, atol=0)
    p, s = ssdutils.decode_location(np.array(loc, np.float32), a[ov.best.idx])
    assert abs(p.x - 0.5) < 1e-6 and abs(s.w - 0.3) < 1e-6
