"""Per-layer roofline table from the CUDA-event profile of one training step (profiles/r1_per_layer_events_b64.json,
written by tools/quick_bench.py): FLOPs and compulsory HBM bytes of every conv launch, the time each bound allows, the
measured time, and which bound the layer is closest to.  CPU only:  python tools/layer_roofline.py [events.json] [out.txt]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ssd-tensorflow_b200'))
from ssdutils import get_preset_by_name   # noqa: E402


def layers(preset_name, num_classes=20):
    """(name, k, stride, Cin, Cout, Hin, Hout) of every conv of the preset, heads merged per feature map like the engine."""
    p = get_preset_by_name(preset_name)
    S = p.image_size.w
    out = []
    h, cin = S, 3
    for blk, n, cout in (('conv1', 2, 64), ('conv2', 2, 128), ('conv3', 3, 256), ('conv4', 3, 512), ('conv5', 3, 512)):
        for i in range(n):
            out.append(('%s_%d' % (blk, i + 1), 3, 1, cin, cout, h, h)); cin = cout
        if blk != 'conv5':
            h = (h + 1) // 2
    out.append(('mod_conv6', 3, 1, 512, 1024, h, h)); out.append(('mod_conv7', 1, 1, 1024, 1024, h, h))
    sizes = [m.size.w for m in p.maps]
    extras = [('conv8', 1024, 256, 512), ('conv9', 512, 128, 256), ('conv10', 256, 128, 256), ('conv11', 256, 128, 256), ('conv12', 256, 128, 256)]
    hin = h
    for (name, c0, c1, c2), hout in zip(extras, sizes[2:]):
        out.append((name + '_1', 1, 1, c0, c1, hin, hin))
        out.append((name + '_2', 3, 2 if hout * 2 >= hin else 1, c1, c2, hin, hout))
        hin = hout
    src = [512, 1024, 512, 256, 256, 256, 256]
    for i, m in enumerate(p.maps):
        out.append(('classifiers/map%d' % i, 3, 1, src[i], (2 + len(m.aspect_ratios)) * (num_classes + 5), m.size.w, m.size.w))
    return out


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r1_per_layer_events_b64.json')
    dst = sys.argv[2] if len(sys.argv) > 2 else None
    d = json.load(open(src))
    B = d['B']
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    tf32 = peaks.get('bf16_tflops_sustained', 1344.5) / 2 * 1e12
    hbm = peaks.get('hbm_gbs', 6524.6) * 1e9
    geo = {n: (k, s, ci, co, hi, ho) for n, k, s, ci, co, hi, ho in layers(d['preset'])}
    rows = ['# per-layer roofline, %s batch %d: tensor bound = FLOPs / %.0f TFLOP/s (sustained tf32), HBM bound = compulsory bytes / %.0f GB/s'
            % (d['preset'], B, tf32 / 1e12, hbm / 1e9),
            '# compulsory bytes: fprop x + y; dgrad dz + dx (+ x as ReLU mask); wgrad x + dz; fp32 NHWC, filters ignored',
            '%-26s %8s %8s %8s %8s %7s %6s  %s' % ('op', 'GFLOP', 'MB', 't_mma', 't_hbm', 't_meas', 'frac', 'nearest bound')]
    tot_meas = tot_bound = 0.0
    for label, ms, _ in d['profile']:
        phase, _, name = label.partition(':')
        if phase not in ('fwd', 'bwd_d', 'bwd_w') or name not in geo:
            continue
        k, s, ci, co, hi, ho = geo[name]
        flops = 2.0 * B * ho * ho * k * k * ci * co
        x, y = 4.0 * B * hi * hi * ci, 4.0 * B * ho * ho * co
        byts = x + y if phase != 'bwd_d' else 2 * x + y
        t_mma, t_hbm = flops / tf32 * 1e3, byts / hbm * 1e3
        bound = max(t_mma, t_hbm)
        tot_meas += ms; tot_bound += bound
        rows.append('%-26s %8.1f %8.0f %8.3f %8.3f %7.3f %6.2f  %s' % (label, flops / 1e9, byts / 1e6, t_mma, t_hbm, ms, bound / ms,
                                                                      'tensor' if t_mma >= t_hbm else 'hbm'))
    rows.append('%-26s %8s %8s %8s %8s %7.3f %6.2f  (sum of bounds %.3f ms)' % ('all conv launches', '', '', '', '', tot_meas, tot_bound / tot_meas, tot_bound))
    text = '\n'.join(rows) + '\n'
    if dst:
        open(dst, 'w').write(text)
    print(text)


if __name__ == '__main__':
    main()
