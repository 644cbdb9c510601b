"""SSDVGG -- the reference's model class (ssdvgg.py:87-649) on the B200 engine.

Keeps the Python surface the reference's train.py / infer.py consume:

    net = SSDVGG(session, preset)
    net.build_from_vgg(vgg_dir, num_classes)            # ssdvgg.py:96
    net.build_optimizer(learning_rate=..., weight_decay=..., momentum=..., global_step=...)   # :375
    result, losses, _ = session.run([net.result, net.losses, net.optimizer],
                                    feed_dict={net.image_input: x, net.labels: y})     # train.py:262-266
    result = session.run(net.result, feed_dict={net.image_input: x, net.keep_prob: 1})   # infer.py:225-227

but there is no TensorFlow graph behind it: ``Session.run`` resolves the fetches to
ONE call into libssd_b200 (forward / forward+loss / forward+loss+backward+update,
hand-written sm_100a kernels), with the host<->device copies of the feeds and
fetches inside that call.  ``Session`` is the small stand-in for ``tf.Session``
that the re-authored entry scripts construct (the reference constructs
``tf.Session()`` itself, train.py:166 / infer.py:211).
"""
import math
import os

import numpy as np

import ssdb
import tf_bundle
import vgg_import
from ssdutils import get_preset_by_name


class Fetch:
    """A named handle usable as a fetch or as a feed_dict key (stands in for a tf.Tensor)."""
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return '<ssd_b200 tensor %s>' % self.name


class Session:
    """Minimal tf.Session stand-in: ``run(fetches, feed_dict)`` over one SSDVGG."""
    def __init__(self):
        self.model = None

    def run(self, fetches, feed_dict=None):
        if self.model is None:
            raise RuntimeError('no model has been built in this session')
        return self.model._run(fetches, feed_dict or {})

    def close(self):
        if self.model is not None:
            self.model.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class GlobalStep:
    """tf.Variable(0, trainable=False) stand-in (train.py:169)."""
    def __init__(self, value=0):
        self.value = int(value)


def piecewise_constant(global_step, boundaries, values):
    """tf.train.piecewise_constant (train.py:43-47,187): x <= b0 -> v0; b0 < x <= b1 -> v1; ..."""
    def lr():
        x = global_step.value
        for b, v in zip(boundaries, values):
            if x <= b:
                return v
        return values[-1]
    return lr


INPUT_STD = 255.0 / math.sqrt(12.0)     # standard deviation of a U[0,255) pixel


def _trunk_scope(name):
    return name.startswith(('conv1_', 'conv2_', 'conv3_', 'conv4_', 'conv5_', 'mod_conv'))


def _flatten(f, out):
    """fetch structure (handles nested in lists / tuples / dicts, like tf.Session.run accepts) -> flat list of handles"""
    if isinstance(f, (list, tuple)):
        for v in f:
            _flatten(v, out)
    elif isinstance(f, dict):
        for v in f.values():
            _flatten(v, out)
    else:
        out.append(f)


def _assemble(f, values):
    """the fetch structure with every handle replaced by its value"""
    if isinstance(f, (list, tuple)):
        return [_assemble(v, values) for v in f]
    if isinstance(f, dict):
        return {k: _assemble(v, values) for k, v in f.items()}
    return values.get(f)


def _pack_gt(gt, counts, B):
    """net.gt_boxes feed -> ([B, G, 5] float64, [B] int32).  Accepts the packed array (+ counts; all rows valid if omitted)
    or per-image sequences of utils.Box / (labelid, cx, cy, w, h) rows."""
    if isinstance(gt, np.ndarray) and gt.ndim == 3:
        arr = np.ascontiguousarray(gt, np.float64)
        cnt = np.full(B, arr.shape[1], np.int32) if counts is None else np.ascontiguousarray(counts, np.int32)
        if arr.shape[0] != B or arr.shape[2] != 5 or cnt.shape != (B,):
            raise ValueError('gt_boxes must be [B, G, 5] with gt_counts [B]')
        return arr, cnt
    if len(gt) != B:
        raise ValueError('gt_boxes must hold one box list per image')
    G = max(1, max(len(g) for g in gt))
    arr = np.zeros((B, G, 5), np.float64)
    cnt = np.zeros(B, np.int32)
    for b, boxes in enumerate(gt):
        cnt[b] = len(boxes)
        for k, bx in enumerate(boxes):
            if hasattr(bx, 'center'):
                arr[b, k] = (bx.labelid, bx.center.x, bx.center.y, bx.size.w, bx.size.h)
            else:
                arr[b, k] = bx
    return arr, cnt


class SSDVGG:
    def __init__(self, session, preset):
        self.preset = preset if not isinstance(preset, str) else get_preset_by_name(preset)
        self.session = session
        session.model = self
        self._built = False
        self._engine = None
        self._host_params = None     # name -> ndarray, until an engine exists (and to survive re-creation)
        self._opt = None
        self._frozen = False         # build_from_frozen: inference-only engine (no training state, CUDA-graph forward)
        self._host_momentum = None   # name -> ndarray restored from a checkpoint, loaded into the engine when it is created
        self._restored_step = 0
        self.epoch = 0               # epochs completed (kept in checkpoints so that --continue-training resumes the loop)
        self._build_names()
        # fetch / feed handles (same attribute names as the reference)
        self.image_input = Fetch('image_input:0')
        self.keep_prob = Fetch('keep_prob:0')
        self.labels = Fetch('labels:0')
        # raw ground truth instead of the dense label tensor (fused anchor matching on the GPU): feed either
        # net.gt_boxes = [B, G, 5] float64 rows (labelid, cx, cy, w, h) with net.gt_counts = [B] int32, or net.gt_boxes = the
        # reference's own per-image lists of utils.Box (what train_generator yields as `gt`, train.py:257)
        self.gt_boxes = Fetch('gt_boxes')
        self.gt_counts = Fetch('gt_counts')
        self.match = Fetch('match')          # fetchable with net.gt_boxes: [B, A] owner GT per anchor, -1 = background
        self.result = Fetch('result/result:0')
        self.logits = Fetch('output/logits')
        self.classifier = Fetch('result/classifier')
        self.locator = Fetch('result/locator')
        self.loss = Fetch('total_loss/loss:0')
        self.confidence_loss = Fetch('confidence_loss/confidence_loss:0')
        self.localization_loss = Fetch('localization_loss/localization_loss:0')
        self.l2_loss = Fetch('total_loss/l2_loss:0')
        self.optimizer = Fetch('optimizer/optimizer')
        self.losses = {'total': self.loss, 'localization': self.localization_loss,
                       'confidence': self.confidence_loss, 'l2': self.l2_loss}

    # ------------------------------------------------------------------ building
    def build_from_vgg(self, vgg_dir, num_classes, a_trous=True, progress_hook='tqdm'):
        """ssdvgg.py:96-118.  The reference downloads a pretrained VGG-16 saved-model
        (ssdvgg.py:153-187, network access) and decimates fc6/fc7 into conv6/conv7.
        Here: if ``<vgg_dir>/vgg/variables/variables.index`` exists (the saved-model the reference
        downloads), its variables are read straight from the TensorFlow tensor bundle and fc6 / fc7
        are decimated like ssdvgg.py:245-280 (vgg_import.py, no TensorFlow needed); else if
        ``<vgg_dir>/vgg16_ssd_init.npz`` exists its tensors (reference variable names) are loaded;
        otherwise the trunk gets a He-normal stand-in (no network in this environment).  The new
        layers always get the reference's Xavier-uniform / zero-bias / scale-20 initialisers
        (ssdvgg.py:46-47,59-60,336)."""
        if not a_trous:
            raise NotImplementedError('only the a-trous variant (the reference default, used by both CLIs) is built')
        self.num_classes = num_classes + 1
        self.num_vars = num_classes + 5
        self._host_params = self._initial_params(num_classes, seed=7)
        path = os.path.join(vgg_dir or '', 'vgg16_ssd_init.npz')
        if vgg_dir and vgg_import.find_bundle(vgg_dir):
            for k, v in vgg_import.load_vgg_dir(vgg_dir).items():
                if k in self._host_params and self._host_params[k].shape == v.shape:
                    self._host_params[k] = v
                else:
                    raise ValueError('VGG variable %s has shape %s, the engine expects %s' %
                                     (k, v.shape, self._host_params[k].shape if k in self._host_params else None))
        elif vgg_dir and os.path.exists(path):
            with np.load(path) as z:
                for k in z.files:
                    if k in self._host_params:
                        self._host_params[k] = z[k].astype(np.float32)
        self._built = True

    def build_from_metagraph(self, metagraph_file, checkpoint_file):
        """ssdvgg.py:120-130: restore a trained model.  `checkpoint_file` is an .npz written
        by ``save`` (tensors under the reference's variable names); the metagraph argument is
        accepted for signature compatibility and ignored (there is no TF graph to import)."""
        if os.path.exists(checkpoint_file + '.index'):            # a TensorFlow checkpoint-V2 bundle (ours or the reference's)
            t = tf_bundle.read_bundle(checkpoint_file)
            epoch = 0
        else:
            with np.load(checkpoint_file if checkpoint_file.endswith('.npz') else checkpoint_file + '.npz') as z:
                t = {k: z[k] for k in z.files if not k.startswith('__')}
                epoch = int(z['__epoch']) if '__epoch' in z.files else 0
        ok = ('/filter', '/biases', '/scale')
        self._host_params = {k: np.asarray(v, np.float32) for k, v in t.items() if k.endswith(ok)}
        # optimizer state, restored like the reference's Saver does (train.py:101-134,191-193): Momentum slots + global_step
        self._host_momentum = {k[:-len('/Momentum')]: np.asarray(v, np.float32) for k, v in t.items()
                               if k.endswith('/Momentum') and k[:-len('/Momentum')].endswith(ok)}
        self._restored_step = int(np.asarray(t['global_step']).reshape(-1)[0]) if 'global_step' in t else 0
        self.epoch = epoch
        row = int(self._host_params['classifiers/classifier0_0/biases'].shape[0])
        self.num_vars = row
        self.num_classes = row - 4
        self._built = True

    def export_frozen(self, path):
        """export_model.py:62-72: freeze the model for deployment.  The reference folds the variables of a checkpoint into a
        GraphDef (.pb); here the frozen model is the weights alone (reference variable names, no optimizer state) plus the
        preset and class count -- everything ``build_from_frozen`` / detect.py need."""
        arrays = self.get_params()
        arrays['__num_vars'] = np.array(self.num_vars)
        arrays['__preset'] = np.array(self.preset.name)
        arrays['__frozen'] = np.array(1)
        np.savez(path, **arrays)

    def build_from_frozen(self, path):
        """detect.py:62-92: load a frozen model.  The engine behind it is an inference handle: no gradient / momentum /
        label buffers, and ``detect`` replays the forward + decode + NMS of a batch as one CUDA graph."""
        with np.load(path) as z:
            if '__frozen' not in z.files:
                raise ValueError('%s is not a frozen model (use export_model.py)' % path)
            if str(z['__preset']) != self.preset.name:
                raise ValueError('frozen model is for preset %s, this SSDVGG is %s' % (z['__preset'], self.preset.name))
            self._host_params = {k: z[k].astype(np.float32) for k in z.files if not k.startswith('__')}
            row = int(z['__num_vars'])
        self.num_vars = row
        self.num_classes = row - 4
        self._frozen = True
        self._built = True

    def build_optimizer(self, learning_rate=0.001, weight_decay=0.0005, momentum=0.9, global_step=None):
        if self._frozen:
            raise RuntimeError('a frozen model cannot be trained')
        """ssdvgg.py:375-599.  `learning_rate` is a float or a zero-argument callable
        (see piecewise_constant)."""
        self._opt = dict(lr=learning_rate, wd=float(weight_decay), mu=float(momentum), step=global_step)
        if global_step is not None and self._restored_step and global_step.value == 0:
            global_step.value = self._restored_step        # a restored model continues its learning-rate schedule

    def build_optimizer_from_metagraph(self, learning_rate=0.00075, weight_decay=0.0005, momentum=0.9, global_step=None):
        """ssdvgg.py:133-150: re-attach the optimizer of a restored model.  The reference finds its hyper-parameters in the
        imported metagraph; there is no graph here, so they are arguments (defaults of train.py:69-76).  The Momentum
        accumulators and global_step come from the checkpoint (build_from_metagraph)."""
        self.build_optimizer(learning_rate, weight_decay, momentum, global_step)

    def build_summaries(self, restore):
        """ssdvgg.py:625-649: TensorBoard histograms are observability, out of scope; a handle is
        returned so callers that fetch it keep working (it evaluates to None)."""
        return Fetch('net_summaries/net_summaries:0')

    def save(self, path, tf_checkpoint=False):
        """Write every trainable tensor under the reference's variable names: an .npz (default), or with
        ``tf_checkpoint=True`` a TensorFlow checkpoint-V2 bundle ``<path>.index`` / ``<path>.data-00000-of-00001`` like the
        reference's ``saver.save(sess, 'e<N>.ckpt')`` (train.py:336-343), readable by TensorFlow tools."""
        arrays = self.get_params()
        # the optimizer state the reference's Saver keeps (train.py:208): Momentum slots '<var>/Momentum' and global_step
        if not self._frozen:
            mom = self.get_params(ssdb.MOMENTUM) if self._engine is not None else (self._host_momentum or {})
            for k, v in mom.items():
                arrays[k + '/Momentum'] = v
        step = self._opt['step'].value if self._opt is not None and self._opt.get('step') is not None else self._restored_step
        arrays['global_step'] = np.array(step, np.int64)
        if tf_checkpoint:
            tf_bundle.write_bundle(path, arrays)
            return
        arrays['__num_vars'] = np.array(self.num_vars)
        arrays['__epoch'] = np.array(self.epoch, np.int64)
        np.savez(path if path.endswith('.npz') else path + '.npz', **arrays)

    def close(self):
        if self._engine is not None:
            self._host_params = self.get_params()
            self._host_momentum = None if self._frozen else self.get_params(ssdb.MOMENTUM)
            self._engine.close()
            self._engine = None

    # ------------------------------------------------------------------ parameters
    def _conv_table(self, num_classes):
        maps = self.preset.maps
        seven = len(maps) >= 7
        t = []
        cin = 3
        for blk, n, cout in (('conv1', 2, 64), ('conv2', 2, 128), ('conv3', 3, 256), ('conv4', 3, 512), ('conv5', 3, 512)):
            for i in range(n):
                t.append(('%s_%d' % (blk, i + 1), 3, cin, cout)); cin = cout
        t += [('mod_conv6', 3, 512, 1024), ('mod_conv7', 1, 1024, 1024), ('conv8_1', 1, 1024, 256), ('conv8_2', 3, 256, 512),
              ('conv9_1', 1, 512, 128), ('conv9_2', 3, 128, 256), ('conv10_1', 1, 256, 128), ('conv10_2', 3, 128, 256),
              ('conv11_1', 1, 256, 128), ('conv11_2', 3, 128, 256)]
        if seven:
            t += [('conv12_1', 1, 256, 128), ('conv12_2', 3, 128, 256)]
        src = [512, 1024, 512, 256, 256, 256, 256]
        for i, m in enumerate(maps):
            for j in range(2 + len(m.aspect_ratios)):
                t.append(('classifiers/classifier%d_%d' % (i, j), 3, src[i], num_classes + 5))
        return t

    def _initial_params(self, num_classes, seed):
        rng = np.random.default_rng(seed)
        P = {}
        for name, k, cin, cout in self._conv_table(num_classes):
            if _trunk_scope(name):
                w = rng.normal(0, math.sqrt(2.0 / (k * k * cin)), (k, k, cin, cout))
                if name == 'conv1_1':
                    w /= INPUT_STD           # raw 0..255 pixels: keep the activations O(1) like the pretrained net
            else:
                lim = math.sqrt(6.0 / (k * k * cin + k * k * cout))      # xavier_initializer(), uniform
                w = rng.uniform(-lim, lim, (k, k, cin, cout))
            P[name + '/filter'] = w.astype(np.float32)
            P[name + '/biases'] = np.zeros(cout, np.float32)
        P['l2_norm_conv4_3/scale'] = np.full(512, 20.0, np.float32)
        return P

    def set_params(self, params):
        """name -> array under the reference's variable names (e.g. 'conv4_3/filter' [3,3,512,512])."""
        if self._engine is not None:
            for k, v in params.items():
                self._engine.set_tensor(k, v)
        else:
            self._host_params.update({k: np.asarray(v, np.float32) for k, v in params.items()})

    def get_params(self, which=ssdb.PARAM):
        if self._engine is None:
            return dict(self._host_params)
        return {k: self._engine.get_tensor(k, shape, which) for k, shape in self._engine.tensors()}

    def _ensure_engine(self, batch):
        if not self._built:
            raise RuntimeError('call build_from_vgg or build_from_metagraph first')
        if self._engine is not None and batch <= self._engine.max_batch:
            return self._engine
        state = None
        if self._engine is not None:
            state = [self.get_params(ssdb.PARAM), None if self._frozen else self.get_params(ssdb.MOMENTUM)]
            self._engine.close()
        eng = ssdb.Net(self.preset.name, self.num_vars - 5, max_batch=batch, inference=self._frozen)
        src = state[0] if state else self._host_params
        for k, shape in eng.tensors():
            if tuple(src[k].shape) != tuple(shape):
                raise ValueError('tensor %s has shape %s, expected %s' % (k, src[k].shape, shape))
            eng.set_tensor(k, src[k])
            if self._frozen:
                continue
            if state:
                eng.set_tensor(k, state[1][k], ssdb.MOMENTUM)
            elif self._host_momentum and k in self._host_momentum:
                eng.set_tensor(k, self._host_momentum[k], ssdb.MOMENTUM)
        self._engine = eng
        return eng

    # ------------------------------------------------------------------ execution
    def _run(self, fetches, feed):
        # (no recursive closures here: a function that refers to itself forms a reference cycle with everything it captured,
        # and the result arrays -- page-locked pool blocks -- would then wait for the cyclic collector instead of being
        # released when the caller drops them)
        flat = []
        _flatten(fetches, flat)
        if self.image_input not in feed:
            raise ValueError('feed_dict must provide net.image_input')
        x = np.ascontiguousarray(feed[self.image_input], np.float32)
        if x.ndim != 4 or x.shape[1] != self.preset.image_size.h or x.shape[2] != self.preset.image_size.w or x.shape[3] != 3:
            raise ValueError('image_input must be [B, %d, %d, 3]' % (self.preset.image_size.h, self.preset.image_size.w))
        loss_handles = set(self.losses.values())
        want_train = any(f is self.optimizer for f in flat)
        want_loss = any(f in loss_handles for f in flat)
        want_match = any(f is self.match for f in flat)
        eng = self._ensure_engine(x.shape[0])
        B = x.shape[0]
        values = {}
        if want_train or want_loss:
            if self.labels not in feed and self.gt_boxes not in feed:
                raise ValueError('feed_dict must provide net.labels (or net.gt_boxes) for loss / optimizer fetches')
            if self._opt is None:
                raise RuntimeError('call build_optimizer first')
            lr = self._opt['lr']() if callable(self._opt['lr']) else float(self._opt['lr'])
            if self.gt_boxes in feed:
                gt, cnt = _pack_gt(feed[self.gt_boxes], feed.get(self.gt_counts), B)
                res, ls, match = eng.train_step_host_gt(x, gt, cnt, lr, self._opt['mu'], self._opt['wd'],
                                                        apply_update=1 if want_train else -1, want_match=want_match)
                values[self.match] = match
            else:
                if want_match:
                    raise ValueError('net.match is produced by the fused matcher: feed net.gt_boxes')
                y = np.ascontiguousarray(feed[self.labels], np.float32)
                if want_train:
                    res, ls = eng.train_step_host(x, y, lr, self._opt['mu'], self._opt['wd'])
                else:
                    res, ls = self._eval_host(eng, x, y)
            if want_train and self._opt['step'] is not None:
                self._opt['step'].value += 1
            values[self.loss], values[self.localization_loss] = ls[0], ls[1]
            values[self.confidence_loss], values[self.l2_loss] = ls[2], ls[3]
        else:
            res = eng.forward_host(x)
        nc = self.num_vars - 4
        # like tf.Session.run, every fetch is an array of its own: `res` owns its (pooled, page-locked) memory block,
        # so it survives later runs and close()
        values[self.result] = res
        values[self.classifier] = res[..., :nc]
        values[self.locator] = res[..., nc:]
        values[self.optimizer] = None
        if any(f is self.logits for f in flat):
            values[self.logits] = eng.read_output(B)[..., :nc]      # ssdvgg.py:366: the pre-softmax class scores

        return _assemble(fetches, values)

    def detect(self, images, confidence_threshold=0.01, lid2name={}, detections_cap=200, overlap_threshold=0.45, rows=False):
        """infer.py:225-235 in one call: forward, then decode_boxes + suppress_overlaps on the device-resident result
        (ssdb_forward_detect_host); only the detections cross PCIe.  Returns per image the reference's
        ``[(confidence, Box), ...]`` list (or, with rows=True, the kernels' integer rows + counts)."""
        from ssdutils import _boxes_from_rows
        x = np.ascontiguousarray(images, np.float32)
        eng = self._ensure_engine(x.shape[0])
        dets, counts = eng.forward_detect_host(x, confidence_threshold, detections_cap, overlap_threshold)
        if rows:
            return dets, counts
        return [_boxes_from_rows(dets[i, :counts[i, 0]], lid2name) for i in range(x.shape[0])]

    def _eval_host(self, eng, x, y):
        import torch                      # device-memory plumbing only
        xd = torch.from_numpy(x).cuda(); yd = torch.from_numpy(y).cuda()
        res = torch.empty((x.shape[0], eng.num_anchors, eng.row), dtype=torch.float32, device='cuda')
        ls = torch.empty(4, dtype=torch.float32, device='cuda')
        st = torch.cuda.current_stream().cuda_stream
        eng.eval_step(xd.data_ptr(), yd.data_ptr(), x.shape[0], self._opt['wd'], ls.data_ptr(), res.data_ptr(), st)
        return res.cpu().numpy(), ls.cpu().numpy()

    # ------------------------------------------------------------------ names
    def _build_names(self):
        """ssdvgg.py:602-622."""
        self.original_scopes = ['conv1_1', 'conv1_2', 'conv2_1', 'conv2_2', 'conv3_1', 'conv3_2', 'conv3_3', 'conv4_1',
                                'conv4_2', 'conv4_3', 'conv5_1', 'conv5_2', 'conv5_3', 'mod_conv6', 'mod_conv7']
        self.new_scopes = ['conv8_1', 'conv8_2', 'conv9_1', 'conv9_2', 'conv10_1', 'conv10_2', 'conv11_1', 'conv11_2']
        if len(self.preset.maps) == 7:
            self.new_scopes += ['conv12_1', 'conv12_2']
        for i, m in enumerate(self.preset.maps):
            for j in range(2 + len(m.aspect_ratios)):
                self.new_scopes.append('classifiers/classifier{}_{}'.format(i, j))
