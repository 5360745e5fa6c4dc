"""Per-kernel parity of the CUDA path (through the C-ABI) against the oracle, on seeded inputs.

Tolerances: integer / index work is bit-exact; fp32-parity mode (bf16x3 MMA, fp32 activations) is held to 1e-4
relative (north_star); throughput mode (bf16) is compared with an fp64 evaluation of the SAME bf16-rounded operands.
"""
import math

import numpy
import pytest
import torch
import torch.nn.functional as F

import oracle
from oracle import ref_ops as R
from util import synthetic_metas, busy_corner_map, nchw, nhwc, relerr

pytestmark = pytest.mark.gpu

TOL_FP32 = 1e-4   # north_star: fp32 activations / gradients within 1e-4 relative
TOL_BF16 = 2e-5   # vs fp64 arithmetic on the same bf16-rounded operands (fp32 accumulation error only)


def _ops():
    from denet_b200 import ops
    return ops


# ------------------------------------------------------------------------------------------------ convolution
CONV_CASES = [
    # n, h, w, cin, cout, k, stride, pad
    (2, 16, 16, 64, 64, 3, 1, 1),
    (3, 14, 14, 96, 100, 3, 1, 1),
    (2, 24, 24, 200, 85, 1, 1, 0),
    (2, 32, 32, 64, 128, 3, 2, 1),
    (2, 30, 30, 72, 128, 3, 2, 1),
    (2, 32, 32, 64, 128, 1, 2, 0),
    (1, 20, 20, 128, 40, 7, 1, 0),
    (2, 9, 9, 256, 512, 3, 1, 2),
    (1, 64, 64, 64, 64, 3, 1, 1),      # wgrad row-shared kernel: 64-pixel patch in one image row
    (2, 32, 32, 128, 136, 3, 1, 1),    # 32 x 2 patch, two Cout tiles
    (4, 8, 8, 64, 96, 3, 1, 1),        # 8 x 8 patch: one MMA spans two image rows
    (2, 16, 16, 72, 64, 5, 1, 2),      # 5 taps per filter row
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("split", [False, True])
def test_conv_fprop_dgrad_wgrad(cuda, case, split):
    ops = _ops()
    n, h, w, cin, cout, k, s, pad = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    bias = torch.randn(cout, generator=g)
    oh = math.ceil((h + 2 * pad - k + 1) / s)
    ow = math.ceil((w + 2 * pad - k + 1) / s)
    dy = torch.randn(n, cout, oh, ow, generator=g)
    if split:
        xr, wr, dyr, tol = x, wt, dy, TOL_FP32
        xd = ops.act_operand(nhwc(x, torch.float32, cuda))
        dyd = ops.act_operand(nhwc(dy, torch.float32, cuda))
    else:
        xr, wr, dyr, tol = x.bfloat16().float(), wt.bfloat16().float(), dy.bfloat16().float(), TOL_BF16
        xd = ops.ActOperand(nhwc(x, torch.bfloat16, cuda))
        dyd = ops.ActOperand(nhwc(dy, torch.bfloat16, cuda))
    wdev = wt.to(cuda)
    wop = ops.conv_weight_prep(wdev, 0, split)
    y = ops.conv2d_fprop(xd, wop, (pad, pad), (oh, ow), torch.float32, stride=(s, s), bias=bias.to(cuda))
    xg = xr.double().requires_grad_(True)
    wg = wr.double().requires_grad_(True)
    yref = R.conv2d(xg, wg, (s, s), pad, bias.double())
    assert tuple(yref.shape) == (n, cout, oh, ow)
    assert relerr(nchw(y), yref.detach()) < tol
    dxref, dwref = torch.autograd.grad(yref, (xg, wg), dyr.double())

    dw = ops.conv2d_wgrad(dyd, xd, k, k, (pad, pad), (s, s))
    assert relerr(dw, dwref) < tol
    # deferred split-K reduction (one multi-tensor launch, as ModelCNN.backward runs it): same fixed order -> same bits
    class Owner:
        pass
    pending, dw2 = [], torch.full((cout, cin, k, k), 3.0, device=cuda)
    ops.conv2d_wgrad(dyd, xd, k, k, (pad, pad), (s, s), dw=dw2, defer=(pending, Owner()))
    assert len(pending) == 1
    ops.wgrad_reduce_pending(pending)
    assert pending == [] and torch.equal(dw2, dw)

    wop_d = ops.conv_weight_prep(wdev, 1, split)
    if s == 1:
        dx = ops.conv2d_fprop(dyd, wop_d, (k - 1 - pad, k - 1 - pad), (h, w), torch.float32)
    elif k == 1:
        dxc = ops.conv2d_fprop(dyd, wop_d, (0, 0), (oh, ow), torch.float32)
        dx = ops.dilate(dxc, (s, s), (h, w))
    else:
        hd, wd = (oh - 1) * s + 1, (ow - 1) * s + 1
        dil = ops.ActOperand(ops.dilate(dyd.hi, (s, s), (hd, wd)),
                             None if dyd.lo is None else ops.dilate(dyd.lo, (s, s), (hd, wd)))
        dx = ops.conv2d_fprop(dil, wop_d, (k - 1 - pad, k - 1 - pad), (h, w), torch.float32)
    assert relerr(nchw(dx), dxref) < tol
    if s > 1 and k > 1:
        # the same data gradient as one stride-1 correlation per parity class on the UNDILATED dy (+ a residual)
        classes = ops.dgrad_parity_classes((h, w), (k, k), (s, s), (pad, pad))
        assert classes is not None and len(classes) == s * s
        records, by_class = [], {}
        for c in classes:
            a, b, r0, s0, rc, sc = c[:6]
            hi = torch.empty((cin, rc * sc, (cout + 63) // 64 * 64), dtype=torch.bfloat16, device=cuda)
            op = ops.ConvOperand(hi, torch.empty_like(hi) if split else None, cin, cout, rc, sc)
            by_class[(a, b)] = op
            records.append((wdev, op, 3, ops.parity_class_code(c, (s, s))))
        ops.conv_weight_prep_records(records)
        add_to = torch.randn(n, cin, h, w, generator=g)
        addd = nhwc(add_to if split else add_to.bfloat16().float(), torch.float32, cuda)
        dx2 = torch.full_like(addd, 9.0)
        for (a, b, r0, s0, rc, sc, ph, pw, hc, wc) in classes:
            ops.conv2d_fprop(dyd, by_class[(a, b)], (ph, pw), (hc, wc), torch.float32, residual=addd, out=dx2,
                             scatter=(h, w, s, s, a, b))
        want = dxref + (add_to if split else add_to.bfloat16().float()).double()
        assert relerr(nchw(dx2), want) < tol


ROWFOLD_CASES = [
    # n, cin, h, w, cout, k, stride, pad
    (2, 3, 32, 32, 16, 7, 2, 3),      # the ResNet stem shape (C.B[64,7,2]) in small
    (3, 3, 16, 16, 128, 3, 1, 1),     # cfg1's first layer C[128,3]: stride 1 -> 8-channel padding
    (2, 1, 20, 28, 40, 5, 2, 2),      # one input channel, non-square image
    (2, 4, 64, 64, 64, 7, 2, 3),
]


@pytest.mark.parametrize("case", ROWFOLD_CASES)
@pytest.mark.parametrize("split", [False, True])
def test_conv_rowfold_stem(cuda, case, split):
    """row-folded stem convolution (overlapping TMA windows over the zero-padded image) vs the oracle conv"""
    ops = _ops()
    n, cin, h, w, cout, k, s, pad = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.rand(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    bias = torch.randn(cout, generator=g)
    oh = math.ceil((h + 2 * pad - k + 1) / s)
    ow = math.ceil((w + 2 * pad - k + 1) / s)
    dy = torch.randn(n, cout, oh, ow, generator=g)
    geom = ops.rowfold_geometry((h, w), cin, (k, k), (s, s), (pad, pad), (oh, ow))
    assert geom is not None
    img = ops.PaddedImage(n, cin, h, w, geom[0], (pad, pad), geom[1], geom[2], split, cuda).fill(x.to(cuda))
    if split:
        xr, wr, dyr, tol = x, wt, dy, TOL_FP32
        dyd = ops.act_operand(nhwc(dy, torch.float32, cuda))
    else:
        xr, wr, dyr, tol = x.bfloat16().float(), wt.bfloat16().float(), dy.bfloat16().float(), TOL_BF16
        dyd = ops.ActOperand(nhwc(dy, torch.bfloat16, cuda))
    wdev = wt.to(cuda)
    wop = ops.conv_weight_prep_rowfold(wdev, geom[0], split)
    ssum, ssq = torch.zeros(cout, device=cuda), torch.zeros(cout, device=cuda)
    y = ops.conv2d_rowfold_fprop(img, wop, (s, s), (oh, ow), torch.float32, bias=bias.to(cuda), stats=(ssum, ssq))
    xg, wg = xr.double().requires_grad_(True), wr.double().requires_grad_(True)
    yref = R.conv2d(xg, wg, (s, s), pad, bias.double())
    assert relerr(nchw(y), yref.detach()) < tol
    assert relerr(ssum, yref.detach().sum(dim=(0, 2, 3))) < 1e-4
    assert relerr(ssq, (yref.detach() ** 2).sum(dim=(0, 2, 3))) < 1e-4
    _, dwref = torch.autograd.grad(yref, (xg, wg), dyr.double())
    dw = torch.full((cout, cin, k, k), 7.0, device=cuda)
    ops.conv2d_rowfold_wgrad(dyd, img, k, k, (s, s), dw)
    assert relerr(dw, dwref) < tol
    class Owner:
        pass
    pending, dw2 = [], torch.full((cout, cin, k, k), 3.0, device=cuda)
    ops.conv2d_rowfold_wgrad(dyd, img, k, k, (s, s), dw2, defer=(pending, Owner()))
    ops.wgrad_reduce_pending(pending)
    assert torch.equal(dw2, dw)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", [(2, 16, 16, 64, 64, 3, True), (2, 24, 24, 96, 136, 3, False), (3, 12, 20, 200, 72, 1, True),
                                  (2, 16, 16, 320, 512, 3, False)])
def test_dgrad_with_fused_bn_backward_statistics(cuda, dtype, case):
    """conv dgrad whose epilogue masks the gradient and accumulates the batch-norm backward sums
    (denet_conv2d_dgrad_bnbwd) + denet_bn_backward_sums  ==  plain dgrad + denet_bn_backward, for both ways of
    getting the ReLU mask (forward output when a residual was added, recomputed from x otherwise)"""
    ops = _ops()
    n, h, w, cin, cout, k, with_res = case        # the conv maps cin -> cout; the batch-norm layer has cin channels
    g = torch.Generator().manual_seed(sum(case[:6]))
    pad = k // 2
    split = dtype == torch.float32
    xbn = (torch.randn(n, cin, h, w, generator=g) * 1.5 + 0.3).to(dtype).float()
    res = torch.randn(n, cin, h, w, generator=g).to(dtype).float()
    dy = torch.randn(n, cout, h, w, generator=g).to(dtype).float()
    add_to = torch.randn(n, cin, h, w, generator=g).to(dtype).float()
    wt = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cout * k * k)).to(cuda)
    gamma, beta = (torch.rand(cin, generator=g) + 0.5).to(cuda), torch.randn(cin, generator=g).to(cuda)
    xd, resd, addd = nhwc(xbn, dtype, cuda), nhwc(res, dtype, cuda), nhwc(add_to, dtype, cuda)
    dyd = ops.act_operand(nhwc(dy, dtype, cuda)) if split else ops.ActOperand(nhwc(dy, dtype, cuda))
    mean, invstd = torch.empty(cin, device=cuda), torch.empty(cin, device=cuda)
    ops.bn_stats(xd, 1e-5, mean, invstd)
    yout = ops.bn_apply(xd, mean, invstd, gamma, beta, residual=resd if with_res else None, relu=True)
    wop_d = ops.conv_weight_prep(wt, 1, split)
    pd = (k - 1 - pad, k - 1 - pad)
    # reference: plain dgrad (+ add_to), then the two-pass batch-norm backward
    dz = ops.conv2d_fprop(dyd, wop_d, pd, (h, w), dtype, residual=addd)
    dg_ref, db_ref = torch.zeros(cin, device=cuda), torch.zeros(cin, device=cuda)
    dx_ref, dres_ref = ops.bn_backward(dz, yout if with_res else None, xd, mean, invstd, gamma, True, dg_ref, db_ref,
                                       want_dres=True, beta=beta)
    # fused
    s0, s1 = torch.zeros(cin, device=cuda), torch.zeros(cin, device=cuda)
    fuse = ops.BnBwdFuse(xd, yout if with_res else None, mean, invstd, gamma, beta, True, s0, s1)
    dzm = ops.conv2d_fprop(dyd, wop_d, pd, (h, w), dtype, residual=addd, bn_bwd=fuse)
    dg, db = torch.zeros(cin, device=cuda), torch.zeros(cin, device=cuda)
    dx = ops.bn_backward_sums(dzm, xd, mean, invstd, gamma, s0, s1, dg, db)
    tol = 2e-5 if split else 1e-2
    assert relerr(dzm[..., :cin], dres_ref[..., :cin]) < (1e-6 if split else 1e-6), "masked gradient"
    assert relerr(dg, dg_ref) < tol and relerr(db, db_ref) < tol
    assert relerr(dx[..., :cin], dx_ref[..., :cin]) < tol


def test_conv_epilogue_residual_relu_stats(cuda):
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    n, h, w, cin, cout = 2, 12, 12, 64, 96
    x = torch.randn(n, cin, h, w, generator=g).bfloat16().float()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / 24).bfloat16().float()
    res = torch.randn(n, cout, h, w, generator=g).bfloat16().float()
    xd = ops.ActOperand(nhwc(x, torch.bfloat16, cuda))
    wop = ops.conv_weight_prep(wt.to(cuda), 0, False)
    ssum = torch.zeros(cout, device=cuda)
    ssq = torch.zeros(cout, device=cuda)
    y = ops.conv2d_fprop(xd, wop, (1, 1), (h, w), torch.bfloat16, residual=nhwc(res, torch.bfloat16, cuda), relu=True,
                         stats=(ssum, ssq))
    yc = R.conv2d(x.double(), wt.double(), (1, 1), "half")
    yref = torch.relu(yc + res.double())
    assert relerr(nchw(y), yref) < 4e-3   # bf16 output rounding
    assert relerr(ssum, yc.sum(dim=(0, 2, 3))) < 1e-4
    assert relerr(ssq, (yc * yc).sum(dim=(0, 2, 3))) < 1e-4


def test_im2col_path_matches_direct_conv(cuda):
    """the 3-channel stem: im2col + 1x1 GEMM == reference convolution (7x7 stride 2, half border)"""
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    n, h, w, cout = 2, 34, 30, 32
    x = torch.rand(n, 3, h, w, generator=g)
    wt = torch.randn(cout, 3, 7, 7, generator=g) / 12
    xd = nhwc(x, torch.float32, cuda)
    oh, ow = math.ceil((h + 6 - 7 + 1) / 2), math.ceil((w + 6 - 7 + 1) / 2)
    col = ops.im2col(xd, 7, 7, (2, 2), (3, 3), (oh, ow))
    w2 = ops.weight_to_im2col(wt.to(cuda))
    wop = ops.conv_weight_prep(w2, 0, True)
    y = ops.conv2d_fprop(ops.act_operand(col), wop, (0, 0), (oh, ow), torch.float32)
    xg = x.double().requires_grad_(True)
    wg = wt.double().requires_grad_(True)
    yref = R.conv2d(xg, wg, (2, 2), "half")
    assert relerr(nchw(y), yref.detach()) < TOL_FP32
    dy = torch.randn(n, cout, oh, ow, generator=g)
    dxref, dwref = torch.autograd.grad(yref, (xg, wg), dy.double())
    dyd = ops.act_operand(nhwc(dy, torch.float32, cuda))
    dw2 = ops.conv2d_wgrad(dyd, ops.act_operand(col), 1, 1, (0, 0))
    dw = torch.empty(cout, 3, 7, 7, device=cuda)
    ops.weight_grad_from_im2col(dw2, dw)
    assert relerr(dw, dwref) < TOL_FP32
    wop_d = ops.conv_weight_prep(w2, 1, True)
    dcol = ops.conv2d_fprop(dyd, wop_d, (0, 0), (oh, ow), torch.float32)
    dx = ops.col2im(dcol, (n, h, w, 3), 7, 7, (2, 2), (3, 3))
    assert relerr(nchw(dx), dxref) < TOL_FP32


# ------------------------------------------------------------------------------------------------ batch norm
def test_bn_known_answer_from_reference_test(cuda):
    """reference denet/layer/batch_norm.py:131-154: U(0,1) seed 1002, (64,128,32,32): mean running stdinv 1.24641"""
    ops = _ops()
    numpy.random.seed(1002)
    x = numpy.random.uniform(0.0, 1.0, (64, 128, 32, 32)).astype(numpy.float32)
    xd = nhwc(x, torch.float32, cuda)
    c = 128
    mean = torch.empty(c, device=cuda)
    invstd = torch.empty(c, device=cuda)
    rm = torch.zeros(c, device=cuda)
    rs = torch.ones(c, device=cuda)
    ops.bn_stats(xd, 1e-5, mean, invstd, rm, rs, 0.9)
    y = ops.bn_apply(xd, mean, invstd, torch.ones(c, device=cuda), torch.zeros(c, device=cuda))
    yh = nchw(y)
    assert abs(yh.mean().item()) < 1e-4 and abs(yh.std().item() - 1.0) < 1e-4
    assert abs(rs.mean().item() - 1.24641) < 1e-4
    assert abs(rm.mean().item() - 0.1 * x.mean()) < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(4, 64, 9, 9), (3, 100, 5, 7), (2, 512, 4, 4)])
def test_bn_forward_backward(cuda, dtype, shape):
    ops = _ops()
    g = torch.Generator().manual_seed(shape[1])
    n, c, h, w = shape
    x = (torch.randn(shape, generator=g) * 2 + 0.5).to(dtype).float()
    res = torch.randn(shape, generator=g).to(dtype).float()
    dy = torch.randn(shape, generator=g).to(dtype).float()
    gamma = torch.rand(c, generator=g) + 0.5
    beta = torch.randn(c, generator=g)
    xd, resd, dyd = nhwc(x, dtype, cuda), nhwc(res, dtype, cuda), nhwc(dy, dtype, cuda)
    mean = torch.empty(c, device=cuda)
    invstd = torch.empty(c, device=cuda)
    ops.bn_stats(xd, 1e-5, mean, invstd)
    y = ops.bn_apply(xd, mean, invstd, gamma.to(cuda), beta.to(cuda), residual=resd, relu=True)
    xg = x.double().requires_grad_(True)
    gg = gamma.double().requires_grad_(True)
    bg = beta.double().requires_grad_(True)
    rg = res.double().requires_grad_(True)
    yb, m_ref, is_ref = R.batchnorm_train(xg, gg, bg, 1e-5)
    yref = torch.relu(yb + rg)
    tol = TOL_FP32 if dtype == torch.float32 else 1e-2
    assert relerr(mean, m_ref.detach()) < 1e-5 and relerr(invstd, is_ref.detach()) < 1e-5
    assert relerr(nchw(y), yref.detach()) < tol
    dgamma = torch.zeros(c, device=cuda)
    dbeta = torch.zeros(c, device=cuda)
    dx, dres = ops.bn_backward(dyd, y, xd, mean, invstd, gamma.to(cuda), True, dgamma, dbeta, want_dres=True)
    # the mask comes from the stored (possibly bf16-rounded) output: use the device's own output sign pattern
    mask = (nchw(y) > 0).double()
    yref2 = (yb + rg) * mask
    dxr, dgr, dbr, drr = torch.autograd.grad(yref2, (xg, gg, bg, rg), dy.double())
    assert relerr(nchw(dx), dxr) < tol and relerr(nchw(dres), drr) < tol
    assert relerr(dgamma, dgr) < tol and relerr(dbeta, dbr) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(4, 64, 9, 9), (3, 100, 5, 7), (2, 512, 4, 4), (2, 2304, 3, 3)])
def test_bn_apply_from_epilogue_sums(cuda, dtype, shape):
    """bn_apply_sums (statistics finalised inside the apply launch) == bn_finalize_sums + bn_apply, bit for bit:
    output, published mean / invstd and the running-statistics update"""
    ops = _ops()
    g = torch.Generator().manual_seed(shape[1] + 1)
    n, c, h, w = shape
    x = (torch.randn(shape, generator=g) * 2 + 0.5).to(dtype).float()
    res = torch.randn(shape, generator=g).to(dtype).float()
    xd, resd = nhwc(x, dtype, cuda), nhwc(res, dtype, cuda)
    gamma, beta = (torch.rand(c, generator=g) + 0.5).to(cuda), torch.randn(c, generator=g).to(cuda)
    xf = xd[..., :c].float().reshape(-1, c)
    ssum, ssq = xf.sum(0).contiguous(), (xf * xf).sum(0).contiguous()
    M = xf.shape[0]
    mean_a, is_a, rm_a, rs_a = (torch.empty(c, device=cuda), torch.empty(c, device=cuda),
                                torch.full((c,), 0.25, device=cuda), torch.full((c,), 1.5, device=cuda))
    ops.bn_finalize_sums(ssum, ssq, M, 1e-5, mean_a, is_a, rm_a, rs_a, 0.9)
    ya = ops.bn_apply(xd, mean_a, is_a, gamma, beta, residual=resd, relu=True)
    mean_b, is_b, rm_b, rs_b = (torch.empty(c, device=cuda), torch.empty(c, device=cuda),
                                torch.full((c,), 0.25, device=cuda), torch.full((c,), 1.5, device=cuda))
    yb = ops.bn_apply_sums(xd, ssum, ssq, 1e-5, gamma, beta, mean_b, is_b, rm_b, rs_b, 0.9, residual=resd, relu=True)
    assert torch.equal(mean_a, mean_b) and torch.equal(is_a, is_b)
    assert torch.equal(rm_a, rm_b) and torch.equal(rs_a, rs_b)
    assert torch.equal(ya[..., :c], yb[..., :c])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bn_backward_mask_recomputed_from_x(cuda, dtype):
    """BN+ReLU without a residual: the backward recomputes the relu mask from x (yout = None) and must give exactly
    what the mask read back from the stored forward output gives"""
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    shape = (4, 72, 12, 10)
    c = shape[1]
    x = (torch.randn(shape, generator=g) * 1.5 - 0.2).to(dtype).float()
    dy = torch.randn(shape, generator=g).to(dtype).float()
    gamma = (torch.rand(c, generator=g) + 0.5).to(cuda)
    gamma[::5] *= -1                                        # negative scales flip the sign relation
    beta = torch.randn(c, generator=g).to(cuda)
    xd, dyd = nhwc(x, dtype, cuda), nhwc(dy, dtype, cuda)
    mean, invstd = torch.empty(c, device=cuda), torch.empty(c, device=cuda)
    ops.bn_stats(xd, 1e-5, mean, invstd)
    y = ops.bn_apply(xd, mean, invstd, gamma, beta, relu=True)
    outs = []
    for yout in (y, None):
        dgamma, dbeta = torch.zeros(c, device=cuda), torch.zeros(c, device=cuda)
        dx, _ = ops.bn_backward(dyd, yout, xd, mean, invstd, gamma, True, dgamma, dbeta, beta=beta)
        outs.append((dx.clone(), dgamma, dbeta))
    assert (nchw(y) > 0).float().mean().item() > 0.2
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


def test_bn_test_mode_quirk(cuda):
    ops = _ops()
    rs = torch.rand(32) + 0.5
    out = ops.bn_inference_invstd(rs.to(cuda), 1e-5)
    ref = 1.0 / torch.sqrt((1.0 / rs.double()) ** 2 + 1e-5)
    assert relerr(out, ref) < 1e-6


# ------------------------------------------------------------------------------------------------ pooling / pool-inv
@pytest.mark.parametrize("mode,size,stride,pad", [("max", (3, 3), (2, 2), (1, 1)), ("max", (2, 2), (2, 2), (0, 0)),
                                                  ("average_inc_pad", (7, 7), (7, 7), (0, 0)),
                                                  ("average_inc_pad", (3, 3), (2, 2), (1, 1))])
def test_pool(cuda, mode, size, stride, pad):
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 40, 14, 14, generator=g)
    oh, ow = R.pool_out_hw((14, 14), size, stride, pad)
    xd = nhwc(x, torch.float32, cuda)
    y, arg = ops.pool_fwd(xd, 0 if mode == "max" else 1, size, stride, pad, (oh, ow))
    xg = x.clone().requires_grad_(True)
    yref = R.pool2d(xg, size, stride, pad, mode)
    assert torch.equal(nchw(y), yref.detach()) or relerr(nchw(y), yref.detach()) < 1e-6
    dy = torch.randn(yref.shape, generator=g)
    dxr, = torch.autograd.grad(yref, xg, dy)
    dx = ops.pool_bwd(nhwc(dy, torch.float32, cuda), 0 if mode == "max" else 1, size, stride, pad, tuple(xd.shape), arg)
    assert relerr(nchw(dx), dxr) < 1e-6


@pytest.mark.parametrize("hw", [(14, 14), (13, 17), (32, 30)])
def test_maxpool_3x3s2_bf16_quad_backward(cuda, hw):
    """the bf16 3x3/stride-2/pad-1 max-pool backward (one thread per 2x2 input quad) gives exactly the fp32 gather
    kernel's result rounded to bf16 (same windows, same summation order), for even and odd extents"""
    ops = _ops()
    g = torch.Generator().manual_seed(sum(hw))
    h, w = hw
    x = torch.randn(3, 64, h, w, generator=g).bfloat16().float()
    oh, ow = R.pool_out_hw((h, w), (3, 3), (2, 2), (1, 1))
    dy = torch.randn(3, 64, oh, ow, generator=g).bfloat16().float()
    x32, dy32 = nhwc(x, torch.float32, cuda), nhwc(dy, torch.float32, cuda)
    y32, arg32 = ops.pool_fwd(x32, 0, (3, 3), (2, 2), (1, 1), (oh, ow))
    dx32 = ops.pool_bwd(dy32, 0, (3, 3), (2, 2), (1, 1), tuple(x32.shape), arg32)
    xb, dyb = nhwc(x, torch.bfloat16, cuda), nhwc(dy, torch.bfloat16, cuda)
    yb, argb = ops.pool_fwd(xb, 0, (3, 3), (2, 2), (1, 1), (oh, ow))
    assert torch.equal(argb, arg32) and torch.equal(yb.float(), y32)
    dxb = ops.pool_bwd(dyb, 0, (3, 3), (2, 2), (1, 1), tuple(xb.shape), argb)
    assert dxb.dtype == torch.bfloat16
    assert torch.equal(dxb[..., :64], dx32[..., :64].bfloat16())


def test_pool_inv_matches_reference_kernels(cuda):
    """recipe of the reference's own A/B block (pool_inv.py:43-88): (4,64,4,4), 2x2, seed 1"""
    ops = _ops()
    numpy.random.seed(1)
    x = numpy.random.uniform(0, 1, (4, 64, 4, 4)).astype(numpy.float32)
    y = ops.pool_inv_fwd(nhwc(x, torch.float32, cuda), (2, 2))
    assert numpy.array_equal(nchw(y).numpy(), oracle.pool_inv_fwd(x, 2, 2))
    assert numpy.array_equal(oracle.pool_inv_fwd(x, 2, 2), numpy.repeat(numpy.repeat(x, 2, axis=2), 2, axis=3))
    dy = numpy.random.uniform(-1, 1, (4, 64, 8, 8)).astype(numpy.float32)
    dx = ops.pool_inv_bwd(nhwc(dy, torch.float32, cuda), (2, 2))
    assert numpy.array_equal(nchw(dx).numpy(), oracle.pool_inv_bwd(dy, 2, 2))   # same fp32 summation order
    ref = oracle.reference_cuda()
    if ref is not None:   # the reference's own kernel text, compiled for this GPU
        xd = torch.from_numpy(x).cuda()
        r = torch.empty(4, 64, 8, 8, device="cuda")
        assert ref.refcuda_pool_inv_fwd_2x2(ctypes_ptr(xd), ctypes_ptr(r), 4, 64, 4, 4) == 0
        assert torch.equal(r.cpu(), nchw(y))
        dyd = torch.from_numpy(dy).cuda()
        r2 = torch.empty(4, 64, 4, 4, device="cuda")
        assert ref.refcuda_pool_inv_bwd_2x2(ctypes_ptr(dyd), ctypes_ptr(r2), 4, 64, 4, 4) == 0
        assert torch.equal(r2.cpu(), nchw(dx))


def ctypes_ptr(t):
    import ctypes
    return ctypes.c_void_p(t.data_ptr())


# ------------------------------------------------------------------------------------------------ sparse sample
def _sparse_inputs(B, Fc, H, W, sn, seed=1):
    """the reference's A/B recipe (denet_sparse.py:228-247): boxes x0,y0~U(0,1), x1~U(x0,1), y1~U(y0,1)"""
    import random
    numpy.random.seed(seed)
    random.seed(seed)
    fmap = numpy.random.uniform(-1, 1, (B, Fc, H, W)).astype(numpy.float32)
    bbox = numpy.zeros((B, sn, sn, 4), dtype=numpy.float32)
    for b in range(B):
        for j in range(sn):
            for i in range(sn):
                x0, y0 = random.uniform(0, 1), random.uniform(0, 1)
                bbox[b, j, i] = (x0, y0, random.uniform(x0, 1), random.uniform(y0, 1))
    return fmap, bbox


@pytest.mark.parametrize("gs,H", [(7, 32), (7, 64), (3, 16), (10, 40)])
def test_sparse_sample_indices_bit_exact(cuda, gs, H):
    ops = _ops()
    _, bbox = _sparse_inputs(4, 8, H, H, 12, seed=gs)
    # boxes that land on .5 exactly (grid 7 => multiples of 1/6 of integer box extents): the lroundf / FMA cases
    bbox[0, 0, :, 0] = numpy.arange(12) / H
    bbox[0, 0, :, 2] = (numpy.arange(12) + 3 + numpy.arange(12) % 5) / H
    ys, xs = ops.sparse_sample_index(torch.from_numpy(bbox).to(cuda), gs, H, H)
    rys, rxs = oracle.sparse_sample_index(bbox, gs, H, H)
    assert numpy.array_equal(ys.cpu().numpy().reshape(rys.shape), rys)
    assert numpy.array_equal(xs.cpu().numpy().reshape(rxs.shape), rxs)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_sparse_sample_fwd_bwd(cuda, dtype):
    ops = _ops()
    B, Fc, H, W, sn, gs = 4, 32, 32, 32, 8, 7
    fmap, bbox = _sparse_inputs(B, Fc, H, W, sn)
    if dtype == torch.bfloat16:
        fmap = torch.from_numpy(fmap).bfloat16().float().numpy()
    bd = torch.from_numpy(bbox).to(cuda)
    out = ops.sparse_sample_fwd(nhwc(fmap, dtype, cuda), bd, gs)
    ref = oracle.sparse_sample_fwd(fmap, bbox, gs)                      # (B, gs*gs*F+2, sn, sn)
    got = nchw(out).numpy()
    if dtype == torch.float32:
        assert numpy.array_equal(got, ref)                              # a pure copy: exact
    else:
        assert numpy.array_equal(got[:, :-2], ref[:, :-2])
        assert numpy.allclose(got[:, -2:], ref[:, -2:], rtol=1e-2)
    dy = numpy.random.uniform(-1, 1, ref.shape).astype(numpy.float32)
    if dtype == torch.bfloat16:
        dy = torch.from_numpy(dy).bfloat16().float().numpy()
    dfmap = ops.sparse_sample_bwd(nhwc(dy, dtype, cuda), bd, gs, (B, H, W, Fc))
    dref = oracle.sparse_sample_bwd(dy, bbox, gs, (B, Fc, H, W))
    assert relerr(dfmap.permute(0, 3, 1, 2), dref) < 1e-6               # fp32 atomics vs double accumulation


def test_sparse_sample_matches_reference_cuda_kernels(cuda):
    """reference A/B block recipe B=32,F=64,H=W=32,sn=24,gs=7 against the reference's own kernel text (oracle/_ref)"""
    ops = _ops()
    ref = oracle.reference_cuda()
    if ref is None:
        pytest.skip("oracle/_ref/libref_cuda_kernels.so not built")
    B, Fc, H, W, sn, gs = 32, 64, 32, 32, 24, 7
    fmap, bbox = _sparse_inputs(B, Fc, H, W, sn)
    fd, bd = torch.from_numpy(fmap).cuda(), torch.from_numpy(bbox).cuda()
    r = torch.empty(B, gs * gs * Fc + 2, sn, sn, device="cuda")
    assert ref.refcuda_sparse_sample_fwd_7(ctypes_ptr(fd), ctypes_ptr(bd), ctypes_ptr(r), B, Fc, H, W, sn) == 0
    out = ops.sparse_sample_fwd(nhwc(fmap, torch.float32, cuda), bd, gs)
    assert torch.equal(nchw(out), r.cpu())
    assert numpy.array_equal(r.cpu().numpy(), oracle.sparse_sample_fwd(fmap, bbox, gs))   # pins the C restatement
    dy = torch.rand(r.shape, device="cuda") - 0.5
    r2 = torch.empty(B, Fc, H, W, device="cuda")
    assert ref.refcuda_sparse_sample_bwd_7(ctypes_ptr(dy), ctypes_ptr(bd), ctypes_ptr(r2), B, Fc, H, W, sn) == 0
    dfmap = ops.sparse_sample_bwd(nhwc(dy.cpu(), torch.float32, cuda), bd, gs, (B, H, W, Fc))
    assert relerr(dfmap.permute(0, 3, 1, 2), r2) < 1e-5                 # both are unordered fp32 atomic sums


# ------------------------------------------------------------------------------------------------ build_samples
def _check_samples(cp, sample_num, max_corners=1024, local_max=0, thr=0.01, cluster_threshold=1.0):
    ops = _ops()
    B = cp.shape[0]
    K = sample_num * sample_num
    pr, bbox, ibox, count, ncand = [t.cpu().numpy() for t in
                                    ops.build_samples(torch.from_numpy(cp).cuda(), thr, sample_num, max_corners,
                                                      local_max, cluster_threshold)]
    ref, ref_ncand = oracle.build_samples(cp, thr, sample_num, max_corners, local_max, cluster_threshold)
    for b in range(B):
        assert ncand[b] == ref_ncand[b], "number of unique candidate boxes differs"
        assert count[b] == len(ref[b])
        n = count[b]
        got = {tuple(int(v) for v in ibox[b, i]): i for i in range(n)}
        want = {(int(s["ix0"]), int(s["iy0"]), int(s["ix1"]), int(s["iy1"])): s for s in ref[b]}
        if ref_ncand[b] > K and cluster_threshold >= 1.0:
            # std::partial_sort is unstable: boxes tied with the K-th score are interchangeable
            cut = ref[b][-1]["pr"]
            got_strict = {k for k, i in got.items() if pr[b, i] > cut}
            want_strict = {k for k, s in want.items() if s["pr"] > cut}
            assert got_strict == want_strict
            assert all(pr[b, i] >= cut for i in got.values())
        else:
            assert set(got) == set(want)
        assert numpy.all(numpy.diff(pr[b, :n]) <= 0), "scores must be sorted descending"
        for k, i in got.items():
            if k in want:
                s = want[k]
                assert pr[b, i] == s["pr"], "score not bit-exact"
                assert tuple(bbox[b, i]) == (s["x0"], s["y0"], s["x1"], s["y1"]), "normalised box not bit-exact"
    return count, ncand


@pytest.mark.parametrize("k,H,sn", [(4, 32, 8), (12, 64, 24), (64, 64, 24), (40, 128, 48)])
def test_build_samples_vs_oracle(cuda, k, H, sn):
    cp = busy_corner_map(3, H, H, k, seed=k)
    count, ncand = _check_samples(cp, sn)
    assert ncand.max() > 0


@pytest.mark.parametrize("k,H,sn", [(6, 32, 8), (20, 64, 24), (48, 64, 24)])
def test_build_samples_centre_corners(cuda, k, H, sn):
    """DNC.C: five maps per image - centres pair with every corner type (denet_sparse.cc:377-468) and the centre
    probability joins every score (:296-303); vs the C restatement and the reference's compiled extension"""
    cp = busy_corner_map(3, H, H, k, seed=100 + k, corner_num=5)
    count, ncand = _check_samples(cp, sn)
    assert ncand.max() > 0
    ref_cc = oracle.reference_cc()
    if ref_cc is not None:
        ops = _ops()
        pr, bbox, ibox, cnt, nc = [t.cpu().numpy() for t in ops.build_samples(torch.from_numpy(cp).cuda(), 0.01, sn)]
        ref = ref_cc.build_samples(3, cp, 0.01, sn, 1024, 0, 1.0)
        for b in range(3):
            assert cnt[b] == len(ref[b])
            want = {tuple(numpy.float32(v) for v in bb): numpy.float32(p) for p, bb in ref[b]}
            cut = min(want.values()) if want else 0
            for i in range(cnt[b]):
                if pr[b, i] > cut or nc[b] <= sn * sn:
                    assert want[tuple(bbox[b, i])] == pr[b, i]


@pytest.mark.parametrize("k,H,sn,cthr,cn", [(12, 32, 4, 0.7, 4), (20, 48, 6, 0.5, 4), (30, 64, 8, 0.7, 4), (16, 32, 8, 0.3, 4),
                                            (40, 64, 24, 0.7, 4), (14, 32, 5, 0.6, 5), (48, 128, 12, 0.7, 4)])
def test_build_samples_clustering(cuda, k, H, sn, cthr, cn):
    """nmsThreshold < 1: apply_cluster (denet_sparse.cc:165-242) as connected components on the device, against the C
    restatement (which the CPU tests pin to the reference's compiled extension) and that extension itself"""
    cp = busy_corner_map(3, H, H, k, seed=200 + k, corner_num=cn)
    count, ncand = _check_samples(cp, sn, cluster_threshold=cthr)
    assert (ncand > sn * sn).any(), "the case must exercise clustering"
    ref_cc = oracle.reference_cc()
    if ref_cc is not None:
        ops = _ops()
        pr, bbox, ibox, cnt, nc = [t.cpu().numpy() for t in
                                   ops.build_samples(torch.from_numpy(cp).cuda(), 0.01, sn, 1024, 0, cthr)]
        ref = ref_cc.build_samples(3, cp, 0.01, sn, 1024, 0, cthr)
        for b in range(3):
            assert cnt[b] == len(ref[b])
            assert {(numpy.float32(p), tuple(numpy.float32(v) for v in bb)) for p, bb in ref[b]} == \
                {(pr[b, i], tuple(bbox[b, i])) for i in range(cnt[b])}


def test_build_samples_edge_cases(cuda):
    # untrained net: bias 5 everywhere (denet_corner.py:42-47) -> no corner passes the threshold -> zero samples
    from util import log_softmax_corner
    cp = log_softmax_corner(numpy.full((2, 4, 32, 32), 5.0, numpy.float32))
    count, ncand = _check_samples(cp, 8)
    assert count.sum() == 0 and ncand.sum() == 0
    # more than max_corners candidates per type (radix-select path) and > sort-buffer candidates (multi-pass select)
    rng = numpy.random.RandomState(7)
    z = rng.randn(2, 4, 32, 32).astype(numpy.float32) * 3
    from util import log_softmax_corner
    cp = log_softmax_corner(z)
    _check_samples(cp, 8, max_corners=64, thr=0.3)
    _check_samples(cp, 24, max_corners=200, thr=0.2)
    # local-max filter (reference loop with exclusive upper bounds)
    _check_samples(busy_corner_map(2, 32, 32, 30, seed=5), 8, local_max=2)
    # heavily quantised logits: many exactly tied scores
    _check_samples(busy_corner_map(2, 32, 32, 40, seed=9, quantize=2), 8)


def test_build_samples_matches_compiled_reference(cuda):
    """the reference's own C++ extension (oracle/_ref, compiled unmodified) on the same maps"""
    ref_cc = oracle.reference_cc()
    if ref_cc is None:
        pytest.skip("oracle/_ref/denet_sparse*.so not built")
    ops = _ops()
    cp = busy_corner_map(4, 64, 64, 16, seed=11)
    sn = 24
    pr, bbox, ibox, count, ncand = [t.cpu().numpy() for t in ops.build_samples(torch.from_numpy(cp).cuda(), 0.01, sn)]
    ref = ref_cc.build_samples(4, cp, 0.01, sn, 1024, 0, 1.0)
    for b in range(4):
        assert count[b] == len(ref[b])
        want = {}
        for p, bb in ref[b]:
            want[tuple(numpy.float32(v) for v in bb)] = numpy.float32(p)
        cut = min(want.values())
        for i in range(count[b]):
            key = tuple(bbox[b, i])
            if pr[b, i] > cut or ncand[b] <= sn * sn:
                assert key in want and want[key] == pr[b, i]


# ------------------------------------------------------------------------------------------------ targets
@pytest.mark.parametrize("B,H,sn,classes,use_bbox", [(32, 64, 24, 80, True), (3, 16, 5, 20, False), (2, 32, 8, 4, True)])
def test_device_targets_bit_exact(cuda, B, H, sn, classes, use_bbox):
    """denet_corner_target / denet_detect_target == the oracle's restatement of the host builders, bit for bit"""
    import random
    ops = _ops()
    metas = synthetic_metas(B, classes, seed=B + sn)
    metas[0]["bbox"], metas[0]["class"] = [], []                      # an image without objects
    rnd = random.Random(5)
    K = sn * sn
    samples = []
    for b in range(B):
        lst = []
        for i in range(K):
            gts = metas[b]["bbox"]
            if gts and i % 3 == 0:                                    # jittered ground truth: positives, shared RoIs
                g = gts[rnd.randrange(len(gts))]
                j = [rnd.uniform(-0.03, 0.03) for _ in range(4)]
                lst.append((0.5, (g[0] + j[0], g[1] + j[1], g[2] + j[2], g[3] + j[3])))
            elif gts and i % 7 == 1:
                lst.append((1.0, tuple(gts[rnd.randrange(len(gts))])))    # exact ground truth (IoU 1)
            else:
                x0, y0 = rnd.uniform(0, 1), rnd.uniform(0, 1)
                lst.append((0.0, (x0, y0, rnd.uniform(x0, 1), rnd.uniform(y0, 1))))
        samples.append(lst)
    if len(metas[1]["bbox"]) >= 2:                                    # two objects of different class on one RoI
        metas[1]["bbox"][1] = metas[1]["bbox"][0]
        metas[1]["class"][1] = (metas[1]["class"][0] + 1) % classes
    G = ops.MAX_GT
    box = numpy.zeros((B, G, 4)); cls = numpy.zeros((B, G), numpy.int32); cnt = numpy.zeros((B,), numpy.int32)
    for b, m in enumerate(metas):
        cnt[b] = len(m["bbox"])
        if cnt[b]:
            box[b, :cnt[b]] = m["bbox"]
            cls[b, :cnt[b]] = m["class"]
    gt = (torch.from_numpy(box).cuda(), torch.from_numpy(cls).cuda(), torch.from_numpy(cnt).cuda())
    for cn in (4, 5):
        tgt = torch.empty((B, 2, cn, H, H), device="cuda")
        ops.corner_target(gt, cn, H, H, tgt)
        ref = R.corner_target(metas, (B, 2, cn, H, H), cn == 5)[1]
        assert numpy.array_equal(tgt.cpu().numpy().reshape(-1), ref)
    s64 = torch.from_numpy(numpy.array([[bb for _, bb in lst] for lst in samples], dtype=numpy.float64)).cuda()
    det = torch.empty((B, classes + 1, sn, sn), device="cuda")
    valid = torch.empty((B, sn, sn), device="cuda") if use_bbox else None
    reg = torch.empty((B, 8, sn, sn), device="cuda") if use_bbox else None
    ops.detect_target(gt, s64, sn, classes, 0.5, 0.5, use_bbox, det, valid, reg)
    ref = R.detect_target(metas, samples, B, sn, classes, 0.5, use_bbox)[1]
    got = torch.cat([t.reshape(-1) for t in (det, valid, reg) if t is not None]).cpu().numpy()
    n0 = det.numel()
    assert numpy.array_equal(got[:n0], ref[:n0]), "class targets"
    assert (ref[:n0].reshape(B, classes + 1, -1)[:, :classes] > 0).sum() > 0, "the case must contain positives"
    assert numpy.array_equal(got, ref)
    # v2 targets: joint fitness (classNum*5+1 channels, denet_detect.py:179-182) and independent fitness (:187-191)
    for thr in (0.5, 0.7):
        detj = torch.empty((B, classes * 5 + 1, sn, sn), device="cuda")
        ops.detect_target(gt, s64, sn, classes, thr, thr, use_bbox, detj, valid, reg, fit_mode=1)
        ref = R.detect_target(metas, samples, B, sn, classes, thr, use_bbox, use_jointfit=True)[1]
        got = torch.cat([t.reshape(-1) for t in (detj, valid, reg) if t is not None]).cpu().numpy()
        assert numpy.array_equal(got, ref), "joint fitness targets"
        bins = ref[:detj.numel()].reshape(B, classes * 5 + 1, -1)[:, :classes * 5].reshape(B, classes, 5, -1)
        assert ((bins > 0).sum(axis=(0, 1, 3)) > 0).sum() >= 3, "several fitness bins must occur"
        fit = torch.empty((B, 6, sn, sn), device="cuda")
        ops.detect_target(gt, s64, sn, classes, thr, thr, use_bbox, det, valid, reg, fit_mode=2, fit=fit)
        ref = R.detect_target(metas, samples, B, sn, classes, thr, use_bbox, use_indfit=True)[1]
        got = torch.cat([t.reshape(-1) for t in (det, valid, reg, fit) if t is not None]).cpu().numpy()
        assert numpy.array_equal(got, ref), "independent fitness targets"


# ------------------------------------------------------------------------------------------------ costs
def test_corner_logprob_and_cost(cuda):
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    B, H, W, cn, Fc = 3, 16, 16, 4, 12
    z = torch.randn(B, cn + Fc, H, W, generator=g) * 3
    zd = nhwc(z, torch.float32, cuda)
    cp = ops.corner_logprob(zd, cn)
    zg = z.double().requires_grad_(True)
    lh = torch.stack([zg[:, :cn], -zg[:, :cn]], dim=1)
    ref = R.log_softmax(lh, 1)
    assert relerr(cp, ref.detach()) < 1e-6
    t = torch.rand(B, 2, cn, H, W, generator=g) / (H * W * cn)
    cost_ref = -(t.double() * ref).sum(dim=(1, 2, 3, 4)).mean() / math.log(2) * 100.0
    dref, = torch.autograd.grad(cost_ref * 0.5, zg)
    dz = ops.alloc_like(zd)
    dz.zero_()
    cost = torch.zeros(1, device=cuda)
    ops.corner_cost(zd, cn, t.to(cuda).contiguous(), 100.0, 0.5, dz, cost)
    assert abs(cost.item() - cost_ref.item()) < 1e-4 * abs(cost_ref.item())
    assert relerr(nchw(dz)[:, :cn], dref[:, :cn]) < 1e-5


@pytest.mark.parametrize("box_mode,nfit,joint", [(0, 0, False), (1, 0, False), (2, 0, False), (1, 6, False),
                                                 (2, 0, True), (0, 6, False)])
def test_detect_cost(cuda, box_mode, nfit, joint):
    """denet_detect_cost_v2 == autograd through the oracle's restatement of get_errors / cost (denet_detect.py:238-313):
    Fast R-CNN and bounded-IoU box losses, the independent-fitness head, joint-fitness channel count"""
    ops = _ops()
    g = torch.Generator().manual_seed(4 + box_mode + nfit)
    B, sn = 3, 6
    s0 = 20 * 5 + 1 if joint else 21
    s1 = 4 if box_mode else 0
    o = torch.randn(B, s0 + s1 + nfit, sn, sn, generator=g)
    if box_mode:
        o[:, s0:s0 + 4] *= 0.3
    od = nhwc(o, torch.float32, cuda)
    t_det = torch.rand(B, s0, sn, sn, generator=g)
    t_det = t_det / t_det.sum(dim=1, keepdim=True) / (sn * sn)
    valid = (torch.rand(B, sn, sn, generator=g) > 0.5).float() / (sn * sn)
    sb = torch.rand(B, sn, sn, 4, generator=g) * 0.5
    sb[..., 2:] += sb[..., :2] + 0.05
    reg = torch.rand(B, 8, sn, sn, generator=g) * 0.5 + 0.1
    if box_mode == 2:      # targets near the samples so that both branches of every switch / minimum occur
        reg[:, 0] = 0.5 * (sb[..., 0] + sb[..., 2]) + (torch.rand(B, sn, sn, generator=g) - 0.5) * 0.2
        reg[:, 1] = 0.5 * (sb[..., 1] + sb[..., 3]) + (torch.rand(B, sn, sn, generator=g) - 0.5) * 0.2
        reg[:, 2] = (sb[..., 2] - sb[..., 0]) * (0.5 + torch.rand(B, sn, sn, generator=g))
        reg[:, 3] = (sb[..., 3] - sb[..., 1]) * (0.5 + torch.rand(B, sn, sn, generator=g))
    t_fit = torch.rand(B, max(nfit, 1), sn, sn, generator=g)
    t_fit = t_fit / t_fit.sum(dim=1, keepdim=True) / (sn * sn)
    parts = [t_det.reshape(-1)]
    if box_mode:
        parts += [valid.reshape(-1), reg.reshape(-1)]
    if nfit:
        parts += [t_fit.reshape(-1)]
    yt = torch.cat(parts).double()
    og = o.double().requires_grad_(True)
    det_pr = R.log_softmax(og[:, :s0], 1)
    fit_pr = R.log_softmax(og[:, s0 + s1:], 1) if nfit else None
    cf, bf, ff = 1.5, 2.0 if box_mode else 0.0, 0.8 if nfit else 0.0
    e_det, e_box, e_fit = R.detect_errors(det_pr, og[:, s0:s0 + 4] if box_mode else None, fit_pr, sb.double(), yt, bf,
                                          box_mode == 2)
    det = cf * e_det.sum() / B
    box = bf * e_box.sum() / B if box_mode else torch.zeros((), dtype=torch.float64)
    fitc = ff * e_fit.sum() / B if nfit else torch.zeros((), dtype=torch.float64)
    dref, = torch.autograd.grad((det + box + fitc) * 0.7, og)
    dout = ops.alloc_like(od)
    cost3 = torch.zeros(3, device=cuda)
    ops.detect_cost(od, sn, s0, box_mode, t_det.to(cuda).contiguous(), valid.to(cuda).contiguous() if box_mode else None,
                    reg.to(cuda).contiguous() if box_mode else None, cf, bf, 0.7, dout, cost3, nfit=nfit,
                    target_fit=t_fit.to(cuda).contiguous() if nfit else None, fit_factor=ff,
                    sample_bbox=sb.to(cuda).contiguous() if box_mode == 2 else None)
    assert abs(cost3[0].item() - det.item()) < 1e-4 * abs(det.item())
    if box_mode:
        assert box.item() > 0 and abs(cost3[1].item() - box.item()) < 1e-4 * abs(box.item())
    if nfit:
        assert abs(cost3[2].item() - fitc.item()) < 1e-4 * abs(fitc.item())
    assert relerr(nchw(dout), dref) < 1e-5
    if box_mode:
        assert relerr(nchw(dout)[:, s0:s0 + 4], dref[:, s0:s0 + 4]) < 1e-4


def test_softmax_nll(cuda):
    ops = _ops()
    g = torch.Generator().manual_seed(6)
    B, C = 37, 1000
    o = torch.randn(B, C, 1, 1, generator=g) * 4
    label = torch.randint(0, C, (B,), generator=g)
    od = nhwc(o, torch.float32, cuda)
    og = o.double().reshape(B, C).requires_grad_(True)
    logp = R.log_softmax(og, 1)
    cost_ref = -logp[torch.arange(B), label].mean()
    dref, = torch.autograd.grad(cost_ref * 2.0, og)
    dout = ops.alloc_like(od)
    lp = torch.empty(B, C, device=cuda)
    cost = torch.zeros(1, device=cuda)
    ops.softmax_nll(od, C, label.int().to(cuda), 2.0, dout, lp, cost)
    assert abs(cost.item() - cost_ref.item()) < 1e-5 * abs(cost_ref.item())
    assert relerr(lp, logp.detach()) < 1e-6
    assert relerr(nchw(dout).reshape(B, C), dref) < 1e-5


# ------------------------------------------------------------------------------------------------ solver
@pytest.mark.parametrize("solver", ["sgd", "nesterov", "adam"])
def test_solver_update(cuda, solver):
    from denet_b200.model import model_cnn
    from denet_b200 import layer as layer_mod
    numpy.random.seed(3)
    model = model_cnn.ModelCNN()
    model.batch_size, model.class_num = 2, 5
    model.build("C.B[8,3] BN A P.A R".split(), (3, 8, 8), "relu", "half", ["he-backward"])
    model.to_device(precision="fp32")
    model.build_train_func(solver, [])
    ref = [(p.detach().cpu().double().clone(), w) for p, w in model.train_params]
    ms = [torch.zeros_like(p) for p, _ in ref]
    vs = [torch.zeros_like(p) for p, _ in ref]
    g = torch.Generator().manual_seed(1)
    for it in range(3):
        grads = [torch.randn(p.shape, generator=g) for p, _ in ref]
        for (p, _), gr in zip(model.train_params, grads):
            p.grad.copy_(gr.to(cuda))
        model.solver_step(0.05, [0.9, 0.99], 1e-3, it, grad_scale=0.5)
        for i, ((p, is_w), gr) in enumerate(zip(ref, grads)):
            out = R.solver_update(p, 0.5 * gr.double(), ms[i], solver, it, 0.05, [0.9, 0.99], 1e-3, is_w, vs[i])
            ref[i] = (out[0], is_w)
            ms[i] = out[1]
            if solver == "adam":
                vs[i] = out[2]
        for (p, _), (pr, _) in zip(model.train_params, ref):
            assert relerr(p, pr) < 1e-5


def test_layout_and_misc(cuda):
    ops = _ops()
    x = torch.randn(3, 10, 7, 5)
    xd = ops.nchw_to_nhwc(x.to(cuda), torch.float32)
    assert torch.equal(nchw(xd), x)
    assert torch.equal(ops.nhwc_to_nchw(xd).cpu(), x)
    a, b = nhwc(x, torch.float32, cuda), nhwc(x * 2 - 1, torch.float32, cuda)
    assert torch.equal(nchw(ops.add(a, b, relu=True)), torch.relu(x + (x * 2 - 1)))
    y = ops.relu_fwd(a)
    assert numpy.array_equal(nchw(y).numpy(), oracle.relu_inplace(x.numpy()))
    dy = torch.randn_like(x)
    assert torch.equal(nchw(ops.relu_bwd(nhwc(dy, torch.float32, cuda), y)), dy * (x > 0))
    s = torch.zeros(10, device=cuda)
    ops.colsum(a, s)
    assert relerr(s, x.double().sum(dim=(0, 2, 3))) < 1e-6
    d = ops.dilate(a, (2, 2), (13, 9))
    ref = torch.zeros(3, 10, 13, 9)
    ref[:, :, ::2, ::2] = x
    assert torch.equal(nchw(d), ref)
    assert torch.equal(nchw(ops.convert(ops.convert(a, torch.bfloat16), torch.float32)), x.bfloat16().float())
