"""Per-layer device time of the conv kernels on the real layer shapes (vgg300 batch 64 / vgg512 batch 32), through
ssdb_op_conv_bench: the bare kernels on engine-format synthetic operands.  Prints ms, TFLOP/s (algorithmic, fp32-equivalent).
    python tools/layer_bench.py [preset] [B] [impl: split|tf32] [filter-substring ...]"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ssd-tensorflow_b200'))
import ssdb

def same_pad(n, k, stride=1, dil=1):
    keff = (k - 1) * dil + 1
    out = -(-n // stride)
    total = max((out - 1) * stride + keff - n, 0)
    return total // 2, out

def layers(preset):
    S = 300 if preset == 'vgg300' else 512
    L = []
    def add(name, H, cin, cout, k=3, stride=1, dil=1, same=True, first=False):
        if same: pad, Ho = same_pad(H, k, stride, dil)
        else: pad, Ho = 0, (H - ((k - 1) * dil + 1)) // stride + 1
        L.append(dict(name=name, H=H, cin=cin, cout=cout, k=k, stride=stride, dil=dil, pad=pad, Ho=Ho, first=first))
        return Ho
    h = S
    add('conv1_1', h, 32, 64, k=1, first=True); add('conv1_2', h, 64, 64); h = -(-h // 2)
    add('conv2_1', h, 64, 128); add('conv2_2', h, 128, 128); h = -(-h // 2)
    add('conv3_1', h, 128, 256); add('conv3_2', h, 256, 256); add('conv3_3', h, 256, 256); h = -(-h // 2)
    add('conv4_1', h, 256, 512); add('conv4_2', h, 512, 512); add('conv4_3', h, 512, 512); h4 = h; h = -(-h // 2)
    add('conv5_1', h, 512, 512); add('conv5_2', h, 512, 512); add('conv5_3', h, 512, 512)
    add('mod_conv6', h, 512, 1024, dil=6); add('mod_conv7', h, 1024, 1024, k=1); h7 = h
    add('conv8_1', h, 1024, 256, k=1); h8 = add('conv8_2', h, 256, 512, stride=2)
    add('conv9_1', h8, 512, 128, k=1); h9 = add('conv9_2', h8, 128, 256, stride=2)
    add('conv10_1', h9, 256, 128, k=1)
    seven = preset == 'vgg512'
    h10 = add('conv10_2', h9, 128, 256, stride=2 if seven else 1, same=seven)
    add('conv11_1', h10, 256, 128, k=1); h11 = add('conv11_2', h10, 128, 256, same=False)
    maps = [(h4, 512, 4), (h7, 1024, 6), (h8, 512, 6), (h9, 256, 6), (h10, 256, 6 if seven else 4), (h11, 256, 4)]
    if seven:
        add('conv12_1', h11, 256, 128, k=1)
        maps.append((1, 256, 4))
    for i, (hh, c, nb) in enumerate(maps):
        add('head%d' % i, hh, c, (nb * 25 + 31) // 32 * 32)
    return L

if __name__ == '__main__':
    preset = sys.argv[1] if len(sys.argv) > 1 else 'vgg300'
    B = int(sys.argv[2]) if len(sys.argv) > 2 else (64 if preset == 'vgg300' else 32)
    impl = {'split': ssdb.CONV_TC_SPLIT, 'tf32': ssdb.CONV_TC, 'simt': ssdb.CONV_SIMT}[sys.argv[3] if len(sys.argv) > 3 else 'split']
    filt = [a for a in sys.argv[4:] if not a.startswith('--')]
    iters = 1 if '--once' in sys.argv else 5            # --once: 2 warm-up launches + 1 (short runs under ncu)
    ssdb.require_device()
    out = []
    tot = [0.0, 0.0, 0.0]
    for l in layers(preset):
        if filt and not any(f in l['name'] for f in filt):
            continue
        gf = 2.0 * B * l['Ho'] * l['Ho'] * l['k'] ** 2 * l['cin'] * l['cout'] / 1e9
        row = {'name': l['name'], 'gflop': gf}
        line = '%-10s H%-3d %4d->%-4d k%d s%d d%d  %7.1f GF ' % (l['name'], l['H'], l['cin'], l['cout'], l['k'], l['stride'], l['dil'], gf)
        for kind, kn in ((0, 'fprop'), (1, 'dgrad'), (2, 'wgrad')):
            if kind == 1 and l['first']:
                continue
            ms = ctypes.c_float(0)
            mask = 0 if l['name'] in ('conv1_1',) else 1
            rc = ssdb.lib().ssdb_op_conv_bench(kind, impl, B, l['H'], l['H'], l['cin'], l['cout'], l['k'], l['stride'], l['dil'], l['pad'], l['pad'],
                                               l['Ho'], l['Ho'], mask, 0, iters, ctypes.byref(ms))
            if rc:
                line += ' %s ERR(%s)' % (kn, ssdb.lib().ssdb_last_error().decode()[:60]); continue
            row[kn] = ms.value; tot[kind] += ms.value
            line += ' %s %6.3f ms %5.0f TF' % (kn, ms.value, gf / ms.value)
        out.append(row)
        print(line, flush=True)
    print('total fprop %.2f dgrad %.2f wgrad %.2f ms' % tuple(tot))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'layer_bench_%s_%d.json' % (preset, B)), 'w'), indent=1)
