"""Pascal VOC submission files (``comp4_det_test_<class>.txt``) from the detections that come out of NMS -- the export half
of SURVEY.md 8f row 3 (reference pascal_summary.py:27-65, called from infer.py:262-264).

Same class / method names as the reference.  ``add_detections`` accepts the image size directly (the reference re-reads
every image with cv2 only to learn its size); without it the file is opened like the reference does."""
import os
from collections import defaultdict, namedtuple

from utils import Size, prop2abs

Detection = namedtuple('Detection', ['fileid', 'confidence', 'left', 'top', 'right', 'bottom'])


def _clamp(v, hi):
    return 0 if v < 0 else (hi - 1 if v >= hi else v)


class PascalSummary:
    def __init__(self):
        self.boxes = defaultdict(list)

    def add_detections(self, filename, boxes, img_size=None):
        """boxes: [(confidence, Box)] of one image; coordinates are scaled to the image, clamped to it and written
        1-based (pascal_summary.py:36-53)."""
        fileid = ''.join(os.path.basename(filename).split('.')[:-1])
        if img_size is None:
            import cv2
            img = cv2.imread(filename)
            img_size = Size(img.shape[1], img.shape[0])
        for conf, box in boxes:
            xmin, xmax, ymin, ymax = prop2abs(box.center, box.size, img_size)
            xmin, xmax = _clamp(xmin, img_size.w), _clamp(xmax, img_size.w)
            ymin, ymax = _clamp(ymin, img_size.h), _clamp(ymax, img_size.h)
            self.boxes[box.label].append(Detection(fileid, conf, float(xmin + 1), float(ymin + 1), float(xmax + 1), float(ymax + 1)))

    def write_summary(self, target_dir):
        """One file per class, one line per detection: ``<image id> <confidence> <left> <top> <right> <bottom>``."""
        for label, dets in self.boxes.items():
            with open(os.path.join(target_dir, 'comp4_det_test_' + label + '.txt'), 'w') as f:
                for d in dets:
                    f.write('{} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f}\n'.format(d.fileid, d.confidence, d.left, d.top, d.right, d.bottom))
