"""Pretrained VGG-16 -> SSD initial parameters without TensorFlow.

The reference loads the Udacity VGG-16 saved-model (``<vgg_dir>/vgg``, tag ``vgg16``; ssdvgg.py:190-207), keeps conv1_1 ..
conv5_3 as they are and turns the fully connected fc6 / fc7 into the a-trous conv6 / 1x1 conv7 by decimation
(ssdvgg.py:245-253, 273-280):

    mod_conv6/filter[h, w, :, i] = fc6/weights[3h, 3w, :, 4i]      h, w < 3, i < 1024      [7,7,512,4096] -> [3,3,512,1024]
    mod_conv6/biases[i]          = fc6/biases[4i]
    mod_conv7/filter[0, 0, j, i] = fc7/weights[0, 0, 4j, 4i]       i, j < 1024             [1,1,4096,4096] -> [1,1,1024,1024]
    mod_conv7/biases[i]          = fc7/biases[4i]

``load_vgg_dir`` reads the variables of that saved-model straight from its tensor bundle (tf_bundle.py) and returns the
tensors under the names the engine uses (= the reference's variable scopes).
"""
import os

import numpy as np

import tf_bundle

VGG_CONVS = ['conv1_1', 'conv1_2', 'conv2_1', 'conv2_2', 'conv3_1', 'conv3_2', 'conv3_3',
             'conv4_1', 'conv4_2', 'conv4_3', 'conv5_1', 'conv5_2', 'conv5_3']


def decimate_fc6(w, b):
    """fc6 [7,7,Cin,4N], [4N] -> conv6 [3,3,Cin,N], [N] (ssdvgg.py:245-253)."""
    w = np.asarray(w); b = np.asarray(b)
    return np.ascontiguousarray(w[0:7:3, 0:7:3, :, 0::4]), np.ascontiguousarray(b[0::4])


def decimate_fc7(w, b):
    """fc7 [1,1,4N,4N], [4N] -> conv7 [1,1,N,N], [N] (ssdvgg.py:273-280)."""
    w = np.asarray(w); b = np.asarray(b)
    return np.ascontiguousarray(w[:, :, 0::4, 0::4]), np.ascontiguousarray(b[0::4])


def vgg_variables_to_params(variables):
    """{'conv1_1/filter': .., 'conv1_1/biases': .., 'fc6/weights': .., 'fc6/biases': .., 'fc7/weights': .., ..} (the
    variable names of the saved-model, ssdvgg.py:193-200,637) -> engine tensors (float32)."""
    out = {}
    for l in VGG_CONVS:
        for part in ('filter', 'biases'):
            key = '%s/%s' % (l, part)
            if key not in variables:
                raise KeyError('VGG variable %s is missing from the bundle (has: %s ...)' % (key, sorted(variables)[:6]))
            out[key] = np.asarray(variables[key], np.float32)
    w6, b6 = decimate_fc6(variables['fc6/weights'], variables['fc6/biases'])
    w7, b7 = decimate_fc7(variables['fc7/weights'], variables['fc7/biases'])
    out['mod_conv6/filter'] = w6.astype(np.float32); out['mod_conv6/biases'] = b6.astype(np.float32)
    out['mod_conv7/filter'] = w7.astype(np.float32); out['mod_conv7/biases'] = b7.astype(np.float32)
    return out


def find_bundle(vgg_dir):
    """Prefix of the saved-model's variable bundle under ``vgg_dir`` (the directory train.py's --vgg-dir names), or None."""
    for sub in ('vgg/variables/variables', 'variables/variables'):
        prefix = os.path.join(vgg_dir, sub)
        if os.path.exists(prefix + '.index'):
            return prefix
    return None


def load_vgg_dir(vgg_dir):
    """Engine tensors from ``<vgg_dir>/vgg/variables/variables.{index,data-00000-of-00001}``."""
    prefix = find_bundle(vgg_dir)
    if prefix is None:
        raise FileNotFoundError('no VGG saved-model variables under ' + vgg_dir)
    wanted = {'%s/%s' % (l, p) for l in VGG_CONVS for p in ('filter', 'biases')}
    wanted |= {'fc6/weights', 'fc6/biases', 'fc7/weights', 'fc7/biases'}
    return vgg_variables_to_params(tf_bundle.read_bundle(prefix, names=wanted))
