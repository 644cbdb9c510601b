"""ctypes binding of libssd_b200.so (include/ssd_b200.h).

There is NO CPU fallback: importing this module only loads the library; every
compute entry point fails loudly (``SSDBError``) when no B200 is present or the
library is missing.  Device pointers are plain integers (e.g. ``tensor.data_ptr()``).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libssd_b200.so')

CONV_AUTO, CONV_SIMT, CONV_TC, CONV_TC_SPLIT = 0, 1, 2, 3
PARAM, GRAD, MOMENTUM = 0, 1, 2
FLAG_INFERENCE = 1


class SSDBError(RuntimeError):
    pass


_lib = None


def lib():
    """The loaded library.  Raises if it was not built (python ssd-tensorflow_b200/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SSDBError('libssd_b200.so is not built: run `python ssd-tensorflow_b200/build.py` '
                            '(or __graft_entry__.build()); there is no CPU fallback')
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_d = C.c_double
_ll = C.c_longlong

# name -> (restype, argtypes); mirrors include/ssd_b200.h one to one
SIGNATURES = {
    'ssdb_crc32c': (C.c_uint32, [C.c_uint32, _p, C.c_size_t]),
    'ssdb_version': (_i, []),
    'ssdb_last_error': (C.c_char_p, []),
    'ssdb_device_ok': (_i, []),
    'ssdb_match_anchors': (_i, [_p, _p, _i, _i, _p, _i, _i, _p, _p, _p]),
    'ssdb_match_anchors_host': (_i, [_p, _p, _i, _i, _p, _i, _i, _p, _p]),
    'ssdb_decode_nms': (_i, [_p, _i, _i, _i, _p, _f, _i, _d, _p, _p, _p]),
    'ssdb_decode_nms_host': (_i, [_p, _i, _i, _i, _p, _f, _i, _d, _p, _p]),
    'ssdb_nms_host': (_i, [_p, _p, _p, _i, _i, _d, _p, _p]),
    'ssdb_multibox_loss': (_i, [_p, _p, _i, _i, _i, _f, _p, _p, _p, _p]),
    'ssdb_multibox_loss_gt': (_i, [_p, _p, _p, _i, _i, _p, _i, _i, _f, _p, _p, _p, _p, _p]),
    'ssdb_op_conv_fprop': (_i, [_i, _p, _p, _p] + [_i] * 13 + [_p, _p]),
    'ssdb_op_conv_dgrad': (_i, [_i, _p, _p, _p] + [_i] * 13 + [_p, _p]),
    'ssdb_op_conv_wgrad': (_i, [_i, _p, _p] + [_i] * 12 + [_p, _p, _p]),
    'ssdb_op_conv_bench': (_i, [_i] * 17 + [C.POINTER(_f)]),
    'ssdb_create': (_i, [C.c_char_p, _i, _i, C.c_uint, C.POINTER(_p)]),
    'ssdb_destroy': (_i, [_p]),
    'ssdb_num_anchors': (_i, [_p]),
    'ssdb_image_size': (_i, [_p]),
    'ssdb_num_tensors': (_i, [_p]),
    'ssdb_tensor_info': (_i, [_p, _i, C.c_char_p, _i, C.POINTER(_i), C.POINTER(_i)]),
    'ssdb_get_tensor': (_i, [_p, C.c_char_p, _i, _p, _ll]),
    'ssdb_set_tensor': (_i, [_p, C.c_char_p, _i, _p, _ll]),
    'ssdb_flat_buffer': (_i, [_p, _i, C.POINTER(_p), C.POINTER(_ll)]),
    'ssdb_set_preprocess': (_i, [_p, _i, C.POINTER(_f)]),
    'ssdb_grad_buckets': (_i, [_p, _i, C.POINTER(_ll), C.POINTER(_ll)]),
    'ssdb_wait_grad_bucket': (_i, [_p, _i, _p]),
    'ssdb_forward': (_i, [_p, _p, _i, _p, _p]),
    'ssdb_forward_host': (_i, [_p, _p, _i, _p]),
    'ssdb_read_output_host': (_i, [_p, _i, _p]),
    'ssdb_train_step': (_i, [_p, _p, _p, _p, _p, _i, _i, _f, _f, _f, _f, _i, _p, _p, _p]),
    'ssdb_train_step_host': (_i, [_p, _p, _p, _i, _f, _f, _f, _p, _p]),
    'ssdb_train_step_host_noupdate': (_i, [_p, _p, _p, _i, _f, _p, _p]),
    'ssdb_train_step_host_gt': (_i, [_p, _p, _p, _p, _i, _i, _f, _f, _f, _i, _p, _p, _p]),
    'ssdb_forward_detect_host': (_i, [_p, _p, _i, _f, _i, _d, _p, _p, _p]),
    'ssdb_train_step_host_begin': (_i, [_p, _p, _p, _p, _p, _i, _i, _f, _p, _p]),
    'ssdb_train_step_host_end': (_i, [_p, _p]),
    'ssdb_eval_step': (_i, [_p, _p, _p, _i, _f, _p, _p, _p]),
    'ssdb_apply_update': (_i, [_p, _f, _f, _f, _f, _p]),
    'ssdb_pinned_alloc': (_i, [_ll, C.POINTER(_p)]),
    'ssdb_pinned_free': (_i, [_p]),
    'ssdb_debug_read': (_i, [_p, C.c_char_p, _i, _p, _ll]),
    'ssdb_debug_shape': (_i, [_p, C.c_char_p, C.POINTER(_i)]),
    'ssdb_launch_count': (_ll, []),
    'ssdb_profile_step': (_i, [_p, _p, _p, _i, _p, _p, _p, _i]),
}


def _declare(l):
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(l, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args


def check(rc):
    if rc != 0:
        raise SSDBError('libssd_b200 error %d: %s' % (rc, lib().ssdb_last_error().decode('utf-8', 'replace')))


def require_device():
    check(lib().ssdb_device_ok())


def _np(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a, a.ctypes.data_as(_p)


# ---------------------------------------------------------------- host-buffer box path
def match_anchors_host(gt, gt_count, anchors_prop, num_classes, want_labels=True, want_match=True):
    """gt [B,G,5] float64, gt_count [B] -> (match [B,A] int32, labels [B,A,C+5] float32)."""
    gt, pgt = _np(gt, np.float64)
    cnt, pcnt = _np(gt_count, np.int32)
    anc, panc = _np(anchors_prop, np.float64)
    B, G = gt.shape[0], gt.shape[1]
    A = anc.shape[0]
    match = np.empty((B, A), np.int32) if want_match else None
    labels = np.empty((B, A, num_classes + 5), np.float32) if want_labels else None
    check(lib().ssdb_match_anchors_host(pgt, pcnt, B, G, panc, A, num_classes,
                                        match.ctypes.data_as(_p) if want_match else None,
                                        labels.ctypes.data_as(_p) if want_labels else None))
    return match, labels


def decode_nms_host(pred, anchors_prop, conf_thr=0.01, cap=200, iou_thr=0.45):
    """pred [B,A,C+5] float32 -> (dets [B,cap_eff,8] int32, counts [B,2] int32)."""
    pred, pp = _np(pred, np.float32)
    anc, panc = _np(anchors_prop, np.float64)
    B, A, V = pred.shape
    cap_i = int(cap) if cap is not None else 0
    cap_eff = cap_i if 0 < cap_i < A else A
    dets = np.zeros((B, cap_eff, 8), np.int32)
    counts = np.zeros((B, 2), np.int32)
    check(lib().ssdb_decode_nms_host(pp, B, A, V - 5, panc, float(np.float32(conf_thr)), cap_i, float(iou_thr),
                                     dets.ctypes.data_as(_p), counts.ctypes.data_as(_p)))
    return dets, counts


def nms_host(boxes_abs, labelid, conf, iou_thr):
    """boxes_abs [n,4] int32, labelid [n] int32, conf [n] float32 -> kept indices in reference order."""
    b, pb = _np(boxes_abs, np.int32)
    l, pl = _np(labelid, np.int32)
    c, pc = _np(conf, np.float32)
    n = b.shape[0]
    keep = np.zeros(n, np.int32)
    cnt = np.zeros(1, np.int32)
    check(lib().ssdb_nms_host(pb, pl, pc, n, int(l.max()) + 1, float(iou_thr), keep.ctypes.data_as(_p), cnt.ctypes.data_as(_p)))
    return keep[:cnt[0]]


class PinnedArray:
    """float32 NumPy view of page-locked host memory owned by the library (explicit free)."""
    def __init__(self, shape):
        self.ptr = _p()
        n = int(np.prod(shape))
        check(lib().ssdb_pinned_alloc(n * 4, C.byref(self.ptr)))
        buf = (C.c_float * n).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=np.float32).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().ssdb_pinned_free(self.ptr)
            self.ptr = _p()


class _PinnedBlock:
    """Page-locked host block that lives as long as any NumPy array made from it: arrays created with ``np.asarray(block)``
    keep the block as their base, and the memory is released only when the last of them (and the pool entry) is gone."""
    def __init__(self, count):
        self.count = int(count)
        self.ptr = _p()
        check(lib().ssdb_pinned_alloc(self.count * 4, C.byref(self.ptr)))
        self.__array_interface__ = {'shape': (self.count,), 'typestr': '<f4', 'data': (self.ptr.value, False), 'version': 3}

    def __del__(self):
        try:
            if self.ptr:
                lib().ssdb_pinned_free(self.ptr)
                self.ptr = _p()
        except Exception:
            pass


class PinnedPool:
    """Result buffers of the host entry points.  ``tf.Session.run`` hands out fresh arrays, so a result must neither be
    overwritten by the next run nor die with the engine: a block is reused only when no array refers to it any more
    (the usual training loop drops the previous result every step, so the steady state is one or two blocks and no
    allocation or copy per step)."""
    MAX_BLOCKS = 8

    def __init__(self, count):
        self.count = int(count)
        self.blocks = []

    def take(self):
        import sys
        for b in self.blocks:
            if sys.getrefcount(b) <= 3:          # the list, the loop variable, the call argument: no array holds it
                return np.asarray(b)
        if len(self.blocks) < self.MAX_BLOCKS:
            b = _PinnedBlock(self.count)
            self.blocks.append(b)
            return np.asarray(b)
        return np.empty(self.count, np.float32)   # every block is still referenced by the caller: plain (pageable) memory

    def clear(self):
        self.blocks = []                           # blocks that are still referenced stay alive through their arrays


# ---------------------------------------------------------------- engine handle
class Net:
    """Owner of one ssdb_net handle (one per GPU, not thread-safe)."""

    def __init__(self, preset, num_classes=20, max_batch=8, inference=False):
        """inference=True: a frozen-model handle (SSDB_FLAG_INFERENCE): forward / detection only, half the memory, the
        device side of forward_detect_host replayed as one CUDA graph."""
        require_device()
        self._h = _p()
        self.inference = bool(inference)
        check(lib().ssdb_create(preset.encode(), int(num_classes), int(max_batch), FLAG_INFERENCE if inference else 0, C.byref(self._h)))
        self.preset = preset
        self.num_classes = int(num_classes)
        self.max_batch = int(max_batch)
        self.num_anchors = lib().ssdb_num_anchors(self._h)
        self.image_size = lib().ssdb_image_size(self._h)
        self.row = self.num_classes + 5
        self._pinned_result = None

    def result_buffer(self, B):
        """A page-locked [B, A, C+5] array for one result.  The array owns a reference to its memory block: it stays
        valid after later calls and after close(); the block returns to the pool when the caller drops the array."""
        if self._pinned_result is None:
            self._pinned_result = PinnedPool(self.max_batch * self.num_anchors * self.row)
        return self._pinned_result.take()[:B * self.num_anchors * self.row].reshape(B, self.num_anchors, self.row)

    def close(self):
        if self._pinned_result is not None:
            self._pinned_result.clear()
            self._pinned_result = None
        if self._h:
            lib().ssdb_destroy(self._h)
            self._h = _p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tensors(self):
        out = []
        name = C.create_string_buffer(128)
        rank = _i()
        shape = (_i * 4)()
        for k in range(lib().ssdb_num_tensors(self._h)):
            check(lib().ssdb_tensor_info(self._h, k, name, 128, C.byref(rank), shape))
            out.append((name.value.decode(), tuple(shape[:rank.value])))
        return out

    def set_tensor(self, name, array, which=PARAM):
        a, pa = _np(array, np.float32)
        check(lib().ssdb_set_tensor(self._h, name.encode(), which, pa, a.size))

    def get_tensor(self, name, shape, which=PARAM):
        a = np.empty(shape, np.float32)
        check(lib().ssdb_get_tensor(self._h, name.encode(), which, a.ctypes.data_as(_p), a.size))
        return a

    def flat_buffer(self, which):
        ptr = _p()
        cnt = _ll()
        check(lib().ssdb_flat_buffer(self._h, which, C.byref(ptr), C.byref(cnt)))
        return ptr.value, cnt.value

    def grad_buckets(self):
        """[(begin, end), ...] float offsets into the flat gradient buffer, in the order the backward completes them."""
        b = (_ll * 8)(); e = (_ll * 8)()
        n = lib().ssdb_grad_buckets(self._h, 8, b, e)
        if n < 0:
            check(n)
        return [(int(b[k]), int(e[k])) for k in range(n)]

    def wait_grad_bucket(self, bucket, stream):
        """Make `stream` (a raw cudaStream_t) wait until bucket `bucket` of the last train_step is complete."""
        check(lib().ssdb_wait_grad_bucket(self._h, int(bucket), stream))

    def set_preprocess(self, swap_rb, mean):
        m = (_f * 3)(*[float(v) for v in mean])
        check(lib().ssdb_set_preprocess(self._h, int(bool(swap_rb)), m))

    # host-buffer calls (what the SSDVGG facade uses: copies are inside)
    def forward_host(self, images):
        x, px = _np(images, np.float32)
        B = x.shape[0]
        res = self.result_buffer(B)
        check(lib().ssdb_forward_host(self._h, px, B, res.ctypes.data_as(_p)))
        return res

    def read_output(self, B):
        """Raw head output [B, A, C+5] (logits | offsets, pre-softmax) of the last forward / train / eval call."""
        out = np.empty((B, self.num_anchors, self.row), np.float32)
        check(lib().ssdb_read_output_host(self._h, B, out.ctypes.data_as(_p)))
        return out

    def debug_read(self, name, B):
        """An intermediate tensor of the last step as float32: '<op>' -> [B,H,W,C] activation, 'grad:<op>' its gradient,
        'output' / 'output_grad' -> [B,A,C+5]."""
        if name in ('output', 'output_grad'):
            shape = (B, self.num_anchors, self.row)
        else:
            hwc = (_i * 3)()
            check(lib().ssdb_debug_shape(self._h, name.split(':')[-1].encode(), hwc))
            shape = (B, hwc[0], hwc[1], hwc[2])
        out = np.empty(shape, np.float32)
        check(lib().ssdb_debug_read(self._h, name.encode(), B, out.ctypes.data_as(_p), out.size))
        return out

    def train_step_host(self, images, labels, lr, momentum, weight_decay, want_result=True, result_out=None):
        x, px = _np(images, np.float32)
        y, py = _np(labels, np.float32)
        B = x.shape[0]
        res = result_out if result_out is not None else (self.result_buffer(B) if want_result else None)
        losses = np.empty(4, np.float32)
        check(lib().ssdb_train_step_host(self._h, px, py, B, lr, momentum, weight_decay, losses.ctypes.data_as(_p),
                                         res.ctypes.data_as(_p) if res is not None else None))
        return res, losses

    def train_step_host_noupdate(self, images, labels, weight_decay, result_out=None):
        """forward + loss + backward from host buffers (overlapped copies), gradients left in the flat buffer."""
        x, px = _np(images, np.float32)
        y, py = _np(labels, np.float32)
        B = x.shape[0]
        res = result_out if result_out is not None else self.result_buffer(B)
        losses = np.empty(4, np.float32)
        check(lib().ssdb_train_step_host_noupdate(self._h, px, py, B, weight_decay, losses.ctypes.data_as(_p), res.ctypes.data_as(_p)))
        return res, losses

    def train_step_host_gt(self, images, gt, gt_count, lr=0.0, momentum=0.0, weight_decay=0.0005, apply_update=True,
                           want_result=True, result_out=None, want_match=False):
        """The training step fed with raw ground truth [B,G,5] float64 + counts [B] (fused anchor matching):
        returns (result or None, losses[4], match [B,A] or None).  apply_update: True / 1 full step, False / 0 backward
        only (gradients stay in the flat buffer), -1 forward + loss only."""
        x, px = _np(images, np.float32)
        g, pg = _np(gt, np.float64)
        c, pc = _np(gt_count, np.int32)
        B = x.shape[0]
        if g.ndim != 3 or g.shape[0] != B or g.shape[2] != 5 or c.shape != (B,):
            raise ValueError('gt must be [B, G, 5] and gt_count [B]')
        res = result_out if result_out is not None else (self.result_buffer(B) if want_result else None)
        match = np.empty((B, self.num_anchors), np.int32) if want_match else None
        losses = np.empty(4, np.float32)
        mode = apply_update if isinstance(apply_update, int) and not isinstance(apply_update, bool) else (1 if apply_update else 0)
        check(lib().ssdb_train_step_host_gt(self._h, px, pg, pc, g.shape[1], B, lr, momentum, weight_decay, mode,
                                            losses.ctypes.data_as(_p), res.ctypes.data_as(_p) if res is not None else None,
                                            match.ctypes.data_as(_p) if want_match else None))
        return res, losses, match

    def train_step_host_begin(self, images, labels=None, gt=None, gt_count=None, weight_decay=0.0005, want_result=True):
        """Enqueue upload + forward + loss + backward from host buffers and return without waiting (data-parallel callers
        overlap the gradient all-reduce); returns the result array that train_step_host_end() completes."""
        x, px = _np(images, np.float32)
        B = x.shape[0]
        self._pending = [x]                       # the host buffers must outlive the asynchronous copies
        if labels is not None:
            y, py = _np(labels, np.float32)
            pg = pc = None; G = 0
            self._pending.append(y)
        else:
            g, pg = _np(gt, np.float64)
            c, pc = _np(gt_count, np.int32)
            py = None; G = g.shape[1]
            self._pending += [g, c]
        res = self.result_buffer(B) if want_result else None
        check(lib().ssdb_train_step_host_begin(self._h, px, py, pg, pc, G, B, weight_decay,
                                               res.ctypes.data_as(_p) if res is not None else None, None))
        return res

    def train_step_host_end(self):
        losses = np.empty(4, np.float32)
        check(lib().ssdb_train_step_host_end(self._h, losses.ctypes.data_as(_p)))
        self._pending = None
        return losses

    def forward_detect_host(self, images, conf_thr=0.01, cap=200, iou_thr=0.45, want_result=False):
        """forward + decode + class-wise NMS with the result tensor kept on the device:
        (dets [B,cap_eff,8] int32, counts [B,2] int32[, result])."""
        x, px = _np(images, np.float32)
        B = x.shape[0]
        cap_i = int(cap) if cap is not None else 0
        cap_eff = cap_i if 0 < cap_i < self.num_anchors else self.num_anchors
        dets = np.zeros((B, cap_eff, 8), np.int32)
        counts = np.zeros((B, 2), np.int32)
        res = self.result_buffer(B) if want_result else None
        check(lib().ssdb_forward_detect_host(self._h, px, B, float(np.float32(conf_thr)), cap_i, float(iou_thr),
                                             dets.ctypes.data_as(_p), counts.ctypes.data_as(_p),
                                             res.ctypes.data_as(_p) if want_result else None))
        return (dets, counts, res) if want_result else (dets, counts)

    # device-pointer calls
    def forward(self, images_ptr, B, result_ptr=None, stream=None):
        check(lib().ssdb_forward(self._h, images_ptr, B, result_ptr, stream))

    def train_step(self, images_ptr, B, labels_ptr=None, gt_ptr=None, gt_count_ptr=None, G=0, lr=0.00075, momentum=0.9,
                   weight_decay=0.0005, grad_scale=1.0, apply_update=True, losses_ptr=None, result_ptr=None, stream=None):
        check(lib().ssdb_train_step(self._h, images_ptr, labels_ptr, gt_ptr, gt_count_ptr, G, B, lr, momentum, weight_decay,
                                    grad_scale, 1 if apply_update else 0, losses_ptr, result_ptr, stream))

    def eval_step(self, images_ptr, labels_ptr, B, weight_decay=0.0005, losses_ptr=None, result_ptr=None, stream=None):
        check(lib().ssdb_eval_step(self._h, images_ptr, labels_ptr, B, weight_decay, losses_ptr, result_ptr, stream))

    def profile_step(self, images_ptr, labels_ptr, B, cap=512):
        """One training step with CUDA events around every op: list of (label, ms, launches)."""
        names = ((C.c_char * 32) * cap)()
        ms = (_f * cap)()
        launches = (_i * cap)()
        n = lib().ssdb_profile_step(self._h, images_ptr, labels_ptr, B, names, ms, launches, cap)
        if n < 0:
            check(n)
        return [(names[k].value.decode(), float(ms[k]), int(launches[k])) for k in range(n)]

    def apply_update(self, lr, momentum, weight_decay, grad_post_scale=1.0, stream=None):
        check(lib().ssdb_apply_update(self._h, lr, momentum, weight_decay, grad_post_scale, stream))


def launch_count():
    return int(lib().ssdb_launch_count())
