"""Bring-up of the split-operand (bf16 hi/lo) tensor-core kernels: every conv kernel on a few shapes against a float64
CPU reference, with 1, 2 and 3 product terms (SSDB_SPLIT_TERMS), printing the max-norm relative error of each.
Expected: terms=1 ~ 4e-3 (bf16 x bf16), terms=2 ~ 2e-3 (one operand exact), terms=3 ~ 1e-5."""
import os, sys, traceback
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'ssd-tensorflow_b200'), os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import ssdb
from gpu_util import conv_case, rel_err, run_dgrad, run_fprop, run_wgrad, torch_conv_ref

CASES = [  # (B, H, Cin, Cout, k, stride, dil, padding)
    (2, 20, 64, 64, 3, 1, 1, 'SAME'),      # rw wgrad
    (4, 10, 256, 128, 1, 1, 1, 'SAME'),    # 1x1, plain wgrad
    (3, 19, 128, 512, 3, 1, 1, 'SAME'),    # two N tiles, plain wgrad
    (4, 19, 256, 512, 3, 2, 1, 'SAME'),    # stride 2
    (2, 19, 64, 96, 3, 1, 6, 'SAME'),      # dilation
    (2, 38, 64, 128, 3, 1, 1, 'SAME'),     # rw wgrad N = 128
]
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
for case in CASES:
    B, H, Cin, Cout, k, stride, dil, padding = case
    x, w, b, pad, Ho = conv_case(B, H, Cin, Cout, k, stride, dil, padding, seed=H + Cin)
    xt, wt, bt, z = torch_conv_ref(x, w, b, stride, dil, pad, Ho, relu=False)
    rng = np.random.default_rng(1)
    dz = rng.standard_normal((B, Ho, Ho, Cout), dtype=np.float32)
    z.backward(torch.tensor(dz).permute(0, 3, 1, 2).to(z.dtype))
    yref = z.detach().permute(0, 2, 3, 1).numpy()
    dxref = xt.grad.permute(0, 2, 3, 1).numpy(); dwref = wt.grad.numpy(); dbref = bt.grad.numpy()
    for terms in (1, 2, 3):
        os.environ['SSDB_SPLIT_TERMS'] = str(terms)
        line = 'case %s terms %d:' % (case, terms)
        for name, fn in (('fprop', lambda: rel_err(run_fprop(ssdb.CONV_TC_SPLIT, x, w, b, k, stride, dil, pad, Ho, relu=False), yref)),
                         ('dgrad', lambda: rel_err(run_dgrad(ssdb.CONV_TC_SPLIT, dz, w, None, x.shape, k, stride, dil, pad), dxref)),
                         ('wgrad', lambda: tuple(rel_err(a, r) for a, r in zip(run_wgrad(ssdb.CONV_TC_SPLIT, x, dz, k, stride, dil, pad), (dwref, dbref))))):
            if which != 'all' and which != name:
                continue
            try:
                e = fn()
                line += ' %s %s' % (name, ('%.2e' % e) if not isinstance(e, tuple) else '(%.2e, bias %.2e)' % e)
            except Exception as ex:     # keep going: one broken kernel must not hide the others
                line += ' %s ERROR %s' % (name, str(ex)[:120])
                if 'CUDA' in str(ex) or 'cuda' in str(ex):
                    print(line, flush=True); traceback.print_exc(); sys.exit(1)
        print(line, flush=True)
os.environ.pop('SSDB_SPLIT_TERMS', None)

# Where does the remaining ~1e-5 per kernel come from?  Operands that ARE bf16 numbers (low parts exactly zero) make every
# product exact in fp32: what is left is the tensor core's accumulation of the K products.
def bf16_round(a):
    t = torch.tensor(a).to(torch.bfloat16).to(torch.float32).numpy()
    return t
for case in [(2, 38, 512, 512, 3, 1, 1, 'SAME'), (2, 38, 64, 512, 3, 1, 1, 'SAME')]:
    B, H, Cin, Cout, k, stride, dil, padding = case
    x, w, b, pad, Ho = conv_case(B, H, Cin, Cout, k, stride, dil, padding, seed=3)
    xb, wb = bf16_round(x), bf16_round(w)
    xt, wt, bt, z = torch_conv_ref(xb, wb, b * 0, stride, dil, pad, Ho, relu=False)
    yref = z.detach().permute(0, 2, 3, 1).numpy()
    e_split = rel_err(run_fprop(ssdb.CONV_TC_SPLIT, xb, wb, b * 0, k, stride, dil, pad, Ho, relu=False), yref)
    e_simt = rel_err(run_fprop(ssdb.CONV_SIMT, xb, wb, b * 0, k, stride, dil, pad, Ho, relu=False), yref)
    print('exact-product case %s (K = %d): tensor-core accumulation error %.2e, fp32 CUDA-core FMA chain %.2e' % (case, k * k * Cin, e_split, e_simt), flush=True)
