"""Helpers shared by the -m gpu parity tests (torch is only the device-memory plumbing)."""
import ctypes

import numpy as np
import torch

import ssdb


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def same_pad(n, k, stride=1, dil=1):
    keff = (k - 1) * dil + 1
    out = -(-n // stride)
    total = max((out - 1) * stride + keff - n, 0)
    return total // 2, out


def conv_case(B, H, Cin, Cout, k, stride=1, dil=1, padding='SAME', seed=0):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, H, H, Cin), dtype=np.float32)
    w = (rng.standard_normal((k, k, Cin, Cout), dtype=np.float32) / np.sqrt(k * k * Cin)).astype(np.float32)
    b = rng.standard_normal(Cout, dtype=np.float32) * 0.1
    if padding == 'SAME':
        pad, Ho = same_pad(H, k, stride, dil)
    else:
        pad, Ho = 0, (H - ((k - 1) * dil + 1)) // stride + 1
    return x, w, b, pad, Ho


def run_fprop(impl, x, w, b, k, stride, dil, pad, Ho, relu=True):
    B, H, _, Cin = x.shape
    Cout = w.shape[3]
    xd, wd, bd = dev(x), dev(w), dev(b)
    y = torch.empty((B, Ho, Ho, Cout), dtype=torch.float32, device='cuda')
    ssdb.check(ssdb.lib().ssdb_op_conv_fprop(impl, ptr(xd), ptr(wd), ptr(bd), B, H, H, Cin, Cout, k, stride, dil, pad, pad,
                                             Ho, Ho, 1 if relu else 0, ptr(y), None))
    torch.cuda.synchronize()
    return y.cpu().numpy()


def run_dgrad(impl, dz, w, mask, shape_x, k, stride, dil, pad, beta=0, dx0=None):
    B, H, _, Cin = shape_x
    Ho, Cout = dz.shape[1], dz.shape[3]
    dzd, wd = dev(dz), dev(w)
    md = dev(mask) if mask is not None else None
    dx = dev(dx0) if dx0 is not None else torch.zeros(shape_x, dtype=torch.float32, device='cuda')
    ssdb.check(ssdb.lib().ssdb_op_conv_dgrad(impl, ptr(dzd), ptr(wd), ptr(md), B, H, H, Cin, Cout, k, stride, dil, pad, pad,
                                             Ho, Ho, beta, ptr(dx), None))
    torch.cuda.synchronize()
    return dx.cpu().numpy()


def run_wgrad(impl, x, dz, k, stride, dil, pad):
    B, H, _, Cin = x.shape
    Ho, Cout = dz.shape[1], dz.shape[3]
    xd, dzd = dev(x), dev(dz)
    dw = torch.empty((k, k, Cin, Cout), dtype=torch.float32, device='cuda')
    db = torch.empty((Cout,), dtype=torch.float32, device='cuda')
    ssdb.check(ssdb.lib().ssdb_op_conv_wgrad(impl, ptr(xd), ptr(dzd), B, H, H, Cin, Cout, k, stride, dil, pad, pad, Ho, Ho,
                                             ptr(dw), ptr(db), None))
    torch.cuda.synchronize()
    return dw.cpu().numpy(), db.cpu().numpy()


def torch_conv_ref(x, w, b, stride, dil, pad, Ho, relu=True, dtype=torch.float64):
    """CPU reference of one TF-style conv with explicit pad-before and fixed output size."""
    import torch.nn.functional as F
    xt = torch.tensor(x).permute(0, 3, 1, 2).to(dtype).requires_grad_(True)
    wt = torch.tensor(w).to(dtype).requires_grad_(True)
    bt = torch.tensor(b).to(dtype).requires_grad_(True)
    k = w.shape[0]
    keff = (k - 1) * dil + 1
    H = x.shape[1]
    after = max((Ho - 1) * stride + keff - H - pad, 0)
    xp = F.pad(xt, (pad, after, pad, after))
    y = F.conv2d(xp, wt.permute(3, 2, 0, 1), bt, stride=stride, dilation=dil)[:, :, :Ho, :Ho]
    if relu:
        y = F.relu(y)
    return xt, wt, bt, y


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
