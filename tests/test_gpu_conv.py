"""GPU parity of the convolution kernels (SIMT and tcgen05) against a float64 torch-CPU
restatement of tf.nn.conv2d / atrous_conv2d + bias + relu and its gradients."""
import os

import numpy as np
import pytest
import torch

import ssdb
from gpu_util import conv_case, rel_err, run_dgrad, run_fprop, run_wgrad, torch_conv_ref

pytestmark = pytest.mark.gpu

FP32_TOL = 2e-5     # fp32 accumulate, different summation order
TF32_TOL = 2e-3     # tf32 operands (10-bit mantissa), fp32 accumulate; relative to max |ref|
SPLIT_TOL = 5e-5    # split bf16 operands (hi + lo = 16 significand bits, three MMAs per product), fp32 accumulate
TC_IMPLS = [(ssdb.CONV_TC_SPLIT, SPLIT_TOL), (ssdb.CONV_TC, TF32_TOL)]     # the engine's default first
TC_IDS = ['split', 'tf32']

# (B, H, Cin, Cout, k, stride, dil, padding)
SIMT_CASES = [
    (2, 20, 3, 64, 3, 1, 1, 'SAME'),       # conv1_1-like (Cin = 3)
    (2, 19, 64, 64, 3, 1, 1, 'SAME'),
    (2, 19, 32, 64, 3, 2, 1, 'SAME'),      # stride 2, pad 1/1
    (2, 10, 32, 64, 3, 2, 1, 'SAME'),      # stride 2, pad 0/1
    (2, 9, 32, 32, 3, 1, 6, 'SAME'),       # dilation 6
    (3, 5, 64, 32, 3, 1, 1, 'VALID'),
    (2, 7, 64, 128, 1, 1, 1, 'SAME'),
]
TC_CASES = [
    (2, 20, 64, 64, 3, 1, 1, 'SAME'),
    (2, 38, 64, 128, 3, 1, 1, 'SAME'),
    (3, 19, 128, 512, 3, 1, 1, 'SAME'),    # two N tiles of 256
    (2, 19, 64, 96, 3, 1, 6, 'SAME'),      # dilation 6, N = 96
    (4, 10, 256, 128, 1, 1, 1, 'SAME'),    # 1x1
    (8, 5, 64, 160, 3, 1, 1, 'SAME'),      # head-like N = 160, tile spans images
    (2, 33, 32, 64, 3, 1, 1, 'VALID'),
]


def _check_all(impl, case, tol):
    B, H, Cin, Cout, k, stride, dil, padding = case
    x, w, b, pad, Ho = conv_case(B, H, Cin, Cout, k, stride, dil, padding, seed=H + Cin)
    xt, wt, bt, y = torch_conv_ref(x, w, b, stride, dil, pad, Ho, relu=True)
    got = run_fprop(impl, x, w, b, k, stride, dil, pad, Ho, relu=True)
    ref = y.detach().permute(0, 2, 3, 1).numpy()
    assert rel_err(got, ref) < tol, ('fprop', case, rel_err(got, ref))
    # backward through the pre-activation: dz arbitrary, no relu
    xt, wt, bt, z = torch_conv_ref(x, w, b, stride, dil, pad, Ho, relu=False)
    rng = np.random.default_rng(1)
    dz = rng.standard_normal((B, Ho, Ho, Cout), dtype=np.float32)
    z.backward(torch.tensor(dz).permute(0, 3, 1, 2).to(z.dtype))
    dw, db = run_wgrad(impl, x, dz, k, stride, dil, pad)
    assert rel_err(dw, wt.grad.numpy()) < tol, ('wgrad', case, rel_err(dw, wt.grad.numpy()))
    assert rel_err(db, bt.grad.numpy()) < max(tol, FP32_TOL * 10), ('bgrad', case, rel_err(db, bt.grad.numpy()))
    if Cin % 4 == 0:
        dx_ref = xt.grad.permute(0, 2, 3, 1).numpy()
        mask = x
        got = run_dgrad(impl, dz, w, mask, x.shape, k, stride, dil, pad, beta=0)
        want = dx_ref * (x > 0)
        assert rel_err(got, want) < tol, ('dgrad', case, rel_err(got, want))
        old = rng.standard_normal(x.shape, dtype=np.float32)
        got = run_dgrad(impl, dz, w, None, x.shape, k, stride, dil, pad, beta=1, dx0=old)
        assert rel_err(got, dx_ref + old) < tol, ('dgrad beta', case)


@pytest.mark.parametrize('case', SIMT_CASES)
def test_simt_conv_matches_oracle(case):
    _check_all(ssdb.CONV_SIMT, case, FP32_TOL)


@pytest.mark.parametrize('impl,tol', TC_IMPLS, ids=TC_IDS)
@pytest.mark.parametrize('case', TC_CASES)
def test_tcgen05_conv_matches_oracle(case, impl, tol):
    _check_all(impl, case, tol)


# real layer shapes of vgg300 (reduced batch) straight against the float64 oracle, split mode:
# conv1_2, conv4_3, mod_conv6 (dilation 6), conv7 (1x1), head of map 1 (6 boxes, N = 160), conv8_2 (stride 2), conv11_2 (VALID)
ORACLE_NET_CASES = [
    (1, 300, 64, 64, 3, 1, 1, 'SAME'), (2, 38, 512, 512, 3, 1, 1, 'SAME'), (2, 19, 512, 1024, 3, 1, 6, 'SAME'),
    (4, 19, 1024, 1024, 1, 1, 1, 'SAME'), (4, 19, 1024, 160, 3, 1, 1, 'SAME'), (4, 19, 256, 512, 3, 2, 1, 'SAME'),
    (64, 3, 128, 256, 3, 1, 1, 'VALID'),
]


@pytest.mark.parametrize('case', ORACLE_NET_CASES)
def test_split_conv_matches_oracle_on_network_shapes(case):
    _check_all(ssdb.CONV_TC_SPLIT, case, SPLIT_TOL)


# every distinct stride-1 conv shape of vgg300 / vgg512 (H, Cin, Cout, k, dil, padding) at a small batch:
# the tensor-core kernels against the CUDA-core kernels on the same device data (no CPU oracle in the loop)
NET_SHAPES = [
    (300, 64, 64, 3, 1, 'SAME'), (150, 64, 128, 3, 1, 'SAME'), (150, 128, 128, 3, 1, 'SAME'), (75, 128, 256, 3, 1, 'SAME'),
    (75, 256, 256, 3, 1, 'SAME'), (38, 256, 512, 3, 1, 'SAME'), (38, 512, 512, 3, 1, 'SAME'), (19, 512, 512, 3, 1, 'SAME'),
    (19, 512, 1024, 3, 6, 'SAME'), (19, 1024, 1024, 1, 1, 'SAME'), (19, 1024, 256, 1, 1, 'SAME'), (10, 512, 128, 1, 1, 'SAME'),
    (5, 256, 128, 1, 1, 'SAME'), (5, 128, 256, 3, 1, 'VALID'),
    (38, 512, 128, 3, 1, 'SAME'), (19, 1024, 160, 3, 1, 'SAME'), (10, 512, 160, 3, 1, 'SAME'), (5, 256, 160, 3, 1, 'SAME'),
    (64, 512, 512, 3, 1, 'SAME'), (32, 512, 1024, 3, 6, 'SAME'), (16, 512, 160, 3, 1, 'SAME'), (8, 256, 160, 3, 1, 'SAME'),
]


@pytest.mark.parametrize('impl,tol', TC_IMPLS, ids=TC_IDS)
@pytest.mark.parametrize('shape', NET_SHAPES)
def test_tcgen05_matches_simt_on_network_shapes(shape, impl, tol):
    H, Cin, Cout, k, dil, padding = shape
    B = 2 if H >= 64 else (4 if H >= 19 else 16)
    x, w, b, pad, Ho = conv_case(B, H, Cin, Cout, k, 1, dil, padding, seed=H * 7 + Cin)
    rng = np.random.default_rng(3)
    dz = rng.standard_normal((B, Ho, Ho, Cout), dtype=np.float32)
    y_s = run_fprop(ssdb.CONV_SIMT, x, w, b, k, 1, dil, pad, Ho)
    y_t = run_fprop(impl, x, w, b, k, 1, dil, pad, Ho)
    assert rel_err(y_t, y_s) < tol, ('fprop', shape, rel_err(y_t, y_s))
    d_s = run_dgrad(ssdb.CONV_SIMT, dz, w, x, x.shape, k, 1, dil, pad)
    d_t = run_dgrad(impl, dz, w, x, x.shape, k, 1, dil, pad)
    assert rel_err(d_t, d_s) < tol, ('dgrad', shape, rel_err(d_t, d_s))
    w_s, _ = run_wgrad(ssdb.CONV_SIMT, x, dz, k, 1, dil, pad)
    w_t, _ = run_wgrad(impl, x, dz, k, 1, dil, pad)
    assert rel_err(w_t, w_s) < tol, ('wgrad', shape, rel_err(w_t, w_s))


@pytest.mark.parametrize('impl,tol', TC_IMPLS, ids=TC_IDS)
@pytest.mark.parametrize('case', [(4, 19, 256, 512, 3, 2, 1, 'SAME'), (4, 10, 128, 256, 3, 2, 1, 'SAME'), (2, 32, 64, 64, 3, 2, 1, 'SAME')])
def test_tcgen05_stride2_matches_simt(case, impl, tol):
    B, H, Cin, Cout, k, stride, dil, padding = case
    x, w, b, pad, Ho = conv_case(B, H, Cin, Cout, k, stride, dil, padding, seed=11)
    rng = np.random.default_rng(5)
    dz = rng.standard_normal((B, Ho, Ho, Cout), dtype=np.float32)
    w_s, b_s = run_wgrad(ssdb.CONV_SIMT, x, dz, k, stride, dil, pad)
    w_t, b_t = run_wgrad(impl, x, dz, k, stride, dil, pad)
    assert rel_err(w_t, w_s) < tol, ('wgrad s2', case, rel_err(w_t, w_s))
    assert rel_err(b_t, b_s) < tol, ('bias s2', case, rel_err(b_t, b_s))
    y_s = run_fprop(ssdb.CONV_SIMT, x, w, b, k, stride, dil, pad, Ho)
    y_t = run_fprop(impl, x, w, b, k, stride, dil, pad, Ho)
    assert rel_err(y_t, y_s) < tol, ('fprop s2', case, rel_err(y_t, y_s))
    d_s = run_dgrad(ssdb.CONV_SIMT, dz, w, x, x.shape, k, stride, dil, pad)
    d_t = run_dgrad(impl, dz, w, x, x.shape, k, stride, dil, pad)
    assert rel_err(d_t, d_s) < tol, ('dgrad s2', case, rel_err(d_t, d_s))
    old = rng.standard_normal(x.shape, dtype=np.float32)
    d_s = run_dgrad(ssdb.CONV_SIMT, dz, w, None, x.shape, k, stride, dil, pad, beta=1, dx0=old)
    d_t = run_dgrad(impl, dz, w, None, x.shape, k, stride, dil, pad, beta=1, dx0=old)
    assert rel_err(d_t, d_s) < tol, ('dgrad s2 beta', case, rel_err(d_t, d_s))


@pytest.mark.parametrize('impl,tol', TC_IMPLS, ids=TC_IDS)
@pytest.mark.parametrize('case', [(2, 64, 64, 64), (3, 40, 128, 128), (2, 38, 512, 128), (4, 19, 64, 96), (2, 150, 64, 128), (2, 38, 256, 512), (4, 19, 512, 1024), (2, 75, 128, 256)])
def test_tcgen05_row_window_wgrad_matches_simt(case, impl, tol):
    """3x3 SAME layers with <= 128 output channels take the row-window wgrad kernel (one x box per filter row)."""
    B, H, Cin, Cout = case
    x, w, b, pad, Ho = conv_case(B, H, Cin, Cout, 3, 1, 1, 'SAME', seed=H + Cout)
    rng = np.random.default_rng(9)
    dz = rng.standard_normal((B, Ho, Ho, Cout), dtype=np.float32)
    w_s, b_s = run_wgrad(ssdb.CONV_SIMT, x, dz, 3, 1, 1, pad)
    os.environ['SSDB_WG_RW_MAXN'] = '1024'          # exercise the window kernel on the wide layers too (N tiles of 128)
    try:
        w_t, b_t = run_wgrad(impl, x, dz, 3, 1, 1, pad)
    finally:
        os.environ.pop('SSDB_WG_RW_MAXN', None)
    assert rel_err(w_t, w_s) < tol, ('wgrad rw', case, rel_err(w_t, w_s))
    assert rel_err(b_t, b_s) < tol, ('bias rw', case, rel_err(b_t, b_s))


def _with_env(name, value, fn):
    old = os.environ.get(name)
    os.environ[name] = value
    try:
        return fn()
    finally:
        if old is None:
            os.environ.pop(name, None)
        else:
            os.environ[name] = old


@pytest.mark.parametrize('case', [(64, 38, 256, 512), (16, 75, 128, 256), (8, 150, 128, 128), (64, 38, 512, 128)])
def test_sm_pair_kernels_equal_single_cta_kernels(case):
    """The cta_group::2 variants (clusters of two CTAs: conv_tc_kernel<.., PAIR> for fprop / dgrad, conv_tc_wgrad_r2c2_kernel for
    the 3x3 weight gradients with Cout % 256 == 0) contract the same products in the same order as the single-CTA kernels:
    every output must be bit-identical.  Shapes with at least two waves of units, so that the pair paths are the ones taken."""
    B, H, Cin, Cout = case
    x, w, b, pad, Ho = conv_case(B, H, Cin, Cout, 3, 1, 1, 'SAME', seed=H + Cout)
    rng = np.random.default_rng(21)
    dz = rng.standard_normal((B, Ho, Ho, Cout), dtype=np.float32)
    impl = ssdb.CONV_TC_SPLIT
    got = {}
    for pair in ('1', '0'):
        got[pair] = _with_env('SSDB_TC_PAIR', pair, lambda: _with_env('SSDB_WG_C2', pair, lambda: (
            run_fprop(impl, x, w, b, 3, 1, 1, pad, Ho),
            run_dgrad(impl, dz, w, x, x.shape, 3, 1, 1, pad),
            run_wgrad(impl, x, dz, 3, 1, 1, pad))))
    y1, d1, (w1, b1) = got['1']
    y0, d0, (w0, b0) = got['0']
    assert np.array_equal(y1, y0), ('fprop', rel_err(y1, y0))
    assert np.array_equal(d1, d0), ('dgrad', rel_err(d1, d0))
    assert np.array_equal(w1, w0) and np.array_equal(b1, b0), ('wgrad', rel_err(w1, w0))
    # and they are right: against the CUDA-core kernels on the same data
    y_s = run_fprop(ssdb.CONV_SIMT, x, w, b, 3, 1, 1, pad, Ho)
    assert rel_err(y1, y_s) < SPLIT_TOL
    w_s, _ = run_wgrad(ssdb.CONV_SIMT, x, dz, 3, 1, 1, pad)
    assert rel_err(w1, w_s) < SPLIT_TOL
