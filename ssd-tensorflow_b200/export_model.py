#!/usr/bin/env python
"""Freeze a trained model for deployment (reference export_model.py:25-72).

The reference imports the metagraph, restores the checkpoint and folds the variables into a GraphDef
(``convert_variables_to_constants``) that detect.py then runs.  There is no TensorFlow graph here: the frozen model is
the weights alone under the reference's variable names (no Momentum slots, no global_step) plus preset and class count,
loaded by ``SSDVGG.build_from_frozen`` into an inference-only engine.  Same flags as the reference."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ssdutils   # noqa: E402
from ssdvgg import SSDVGG, Session   # noqa: E402


def main():
    ap = argparse.ArgumentParser(description='Export a frozen model')
    ap.add_argument('--metagraph-file', default='final.ckpt.meta', help='accepted for compatibility (there is no metagraph)')
    ap.add_argument('--checkpoint-file', default='final.npz', help='.npz written by train.py, or a TensorFlow checkpoint prefix')
    ap.add_argument('--output-file', default='model.frozen.npz')
    ap.add_argument('--output-tensors', nargs='+', default=['result/result'],
                    help='accepted for compatibility: the frozen model always serves image_input -> result/result')
    ap.add_argument('--preset', default='vgg300')
    args = ap.parse_args()
    print('[i] Checkpoint file: ', args.checkpoint_file)
    print('[i] Output file:     ', args.output_file)
    print('[i] Output tensors:  ', args.output_tensors)
    if not (os.path.exists(args.checkpoint_file) or os.path.exists(args.checkpoint_file + '.npz') or os.path.exists(args.checkpoint_file + '.index')):
        print('[!] Cannot find file:', args.checkpoint_file)
        return 1
    if any(t.split(':')[0] not in ('result/result', 'result') for t in args.output_tensors):
        print('[!] Only result/result can be an output of the frozen model')
        return 1
    net = SSDVGG(Session(), ssdutils.get_preset_by_name(args.preset))
    net.build_from_metagraph(args.metagraph_file, args.checkpoint_file)
    net.export_frozen(args.output_file)
    print('[i] Wrote', args.output_file)
    return 0


if __name__ == '__main__':
    sys.exit(main())
