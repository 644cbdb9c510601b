#!/usr/bin/env python
"""Detection with a frozen model (reference detect.py:40-128): read image files, resize to the preset's input size, run the
network, decode + suppress, keep the first 200 boxes, write ``<name>.txt`` (label labelid cx cy w h) and the annotated image.

The reference imports a frozen GraphDef into a tf.Session; here ``SSDVGG.build_from_frozen`` creates an inference-only
engine and ``SSDVGG.detect`` does forward + decode_boxes + suppress_overlaps in one call with the result tensor kept on the
device (from the second batch of a given size on, as a single CUDA graph launch)."""
import argparse
import os
import pickle
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ssdutils   # noqa: E402
from ssdvgg import SSDVGG, Session   # noqa: E402

VOC_LABELS = ['aeroplane', 'bicycle', 'bird', 'boat', 'bottle', 'bus', 'car', 'cat', 'chair', 'cow', 'diningtable', 'dog',
              'horse', 'motorbike', 'person', 'pottedplant', 'sheep', 'sofa', 'train', 'tvmonitor']    # source_pascal_voc.py:37-57


def draw_box(img, box, color):
    """utils.py:138-148: the box and its label on the image."""
    import cv2
    h, w = img.shape[:2]
    xmin, xmax, ymin, ymax = ssdutils.prop2abs(box.center, box.size, ssdutils.Size(w, h))
    cv2.rectangle(img, (xmin, ymin), (xmax, ymax), color, 2)
    cv2.rectangle(img, (xmin - 1, ymin), (xmax + 1, ymin - 20), color, cv2.FILLED)
    cv2.putText(img, str(box.label), (xmin + 5, ymin - 5), cv2.FONT_HERSHEY_SIMPLEX, 0.5, (255, 255, 255), 1)


def main():
    ap = argparse.ArgumentParser(description='SSD inference')
    ap.add_argument('files', nargs='*')
    ap.add_argument('--model', default='model.frozen.npz', help='frozen model written by export_model.py')
    ap.add_argument('--training-data', default='', help="the reference's training-data pickle (preset, colors, lid2name); optional")
    ap.add_argument('--output-dir', default='test-out')
    ap.add_argument('--batch-size', type=int, default=32)
    ap.add_argument('--preset', default='vgg300')
    ap.add_argument('--threshold', type=float, default=0.5)
    args = ap.parse_args()
    print('[i] Model:         ', args.model)
    print('[i] Output dir:    ', args.output_dir)
    print('[i] Batch size:    ', args.batch_size)
    import cv2
    lid2name = {i: n for i, n in enumerate(VOC_LABELS)}
    colors = {n: (int(37 * i % 255), int(97 * i % 255), int(181 * i % 255)) for i, n in lid2name.items()}
    preset = ssdutils.get_preset_by_name(args.preset)
    if args.training_data:
        with open(args.training_data, 'rb') as f:
            data = pickle.load(f)
        lid2name, colors = data['lid2name'], data['colors']
        preset = ssdutils.get_preset_by_name(data['preset'].name)
    os.makedirs(args.output_dir, exist_ok=True)
    side = (preset.image_size.w, preset.image_size.h)
    with Session() as sess:
        net = SSDVGG(sess, preset)
        net.build_from_frozen(args.model)
        for i in range(0, len(args.files), args.batch_size):
            names = args.files[i:i + args.batch_size]
            originals = [cv2.imread(f) for f in names]
            if any(o is None for o in originals):
                print('[!] Cannot read', [f for f, o in zip(names, originals) if o is None])
                return 1
            batch = np.array([cv2.resize(o, side) for o in originals], np.float32)
            dets = net.detect(batch, args.threshold, lid2name, None)         # detect.py:111-112: no cap, then [:200]
            for name, img, boxes in zip(names, originals, dets):
                base = os.path.basename(name)
                with open(os.path.join(args.output_dir, base + '.txt'), 'w') as f:
                    for conf, box in boxes[:200]:
                        draw_box(img, box, colors.get(box.label, (0, 255, 0)))
                        f.write('{} {} {} {} {} {}\n'.format(box.label, box.labelid, box.center.x, box.center.y, box.size.w, box.size.h))
                cv2.imwrite(os.path.join(args.output_dir, base), img)
            print('[i] %d images: %s detections' % (len(names), [len(b[:200]) for b in dets]))
    return 0


if __name__ == '__main__':
    sys.exit(main())
