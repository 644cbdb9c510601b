"""Error floor of tf32 operands for the SSD-VGG forward, computed on the CPU: the float64 oracle rounding to tf32 exactly where the
engine does (pre-processed image, filters, conv outputs that feed convs, L2-norm output; exact products, float64 accumulation) against the plain float64 oracle, on
the inputs of tests/test_gpu_net.py.  The engine's measured error (profiles/net_parity_*_auto.json) should sit on this
floor; what exceeds it would be a kernel defect.   python tools/tf32_floor.py [vgg300|vgg512] [B]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'ssd-tensorflow_b200'))
import box_oracle as bo   # noqa: E402
import net_oracle as no   # noqa: E402
import synth              # noqa: E402

preset = sys.argv[1] if len(sys.argv) > 1 else 'vgg300'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
side = bo.PRESETS[preset]['image']
P = no.init_params(preset, dtype=torch.float64)
x = torch.tensor(synth.images(0, B, side))
with torch.no_grad():
    ref = no.result_from_output(no.forward(P, x, preset)).numpy()
    got = no.result_from_output(no.forward(P, x, preset, producer_round=no.round_tf32)).numpy()
rep = {'preset': preset, 'B': B, 'model': 'float64 oracle with tf32 rounding at every producer (the engine\'s arithmetic model) vs plain float64 oracle',
       'softmax_abs': float(np.abs(got[..., :21] - ref[..., :21]).max()),
       'locator_rel': float(np.abs(got[..., 21:] - ref[..., 21:]).max() / np.abs(ref[..., 21:]).max()),
       'locator_rel_rms': float(np.sqrt(((got[..., 21:] - ref[..., 21:]) ** 2).mean() / (ref[..., 21:] ** 2).mean())),
       'argmax_agree': float((got[..., :21].argmax(-1) == ref[..., :21].argmax(-1)).mean())}
eng = os.path.join(ROOT, 'profiles', 'net_parity_%s_auto.json' % preset)
if os.path.exists(eng):
    e = json.load(open(eng))
    rep['engine_measured'] = {k: e[k] for k in ('softmax_abs', 'locator_rel', 'locator_rel_rms', 'argmax_agree') if k in e}
print(json.dumps(rep))
