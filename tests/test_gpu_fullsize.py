"""Parity at the shapes the benchmark actually runs.

Every distinct convolution (filter, input, stride) of BASELINE cfg3 (DeNet-34 skip, batch 32) and cfg5 (DeNet-101 wide,
batch 8) is run through the ConvLayer object itself - forward, data gradient, filter gradient - so that the template
instances, persistent-CTA rounds (TMEM double-buffer / mbarrier phase wrap), split-K plans, parity-class dgrads and the
row-folded stem are exactly the ones the bench launches, in both precision modes.  The checker is an fp64 evaluation
of oracle.ref_ops.conv2d on the SAME operands, run on the GPU for speed (test-only; the product path never sees it).
Full-size batch-norm, pool-inv, sparse-sample and build_samples cases follow.  The measured worst errors are written to
gpurun_out/r2_fullsize_parity.json (committed under profiles/).

Error measures: `frob` = ||got - ref||_F / ||ref||_F (the north_star's "within 1e-4 relative" is read on this, like
round 1) and `linf` = max|got - ref| / max|ref| (element-wise worst case against the tensor's scale; reported and
bounded too, since a Frobenius ratio can hide a few wrong elements).
"""
import json
import math
import os

import numpy
import pytest
import torch

import oracle
from oracle import ref_ops as R
from util import busy_corner_map

pytestmark = pytest.mark.gpu

# frob; bf16 mode is compared on the same bf16-rounded operands.  The TMEM accumulator truncates on every accumulation
# step (measured relative bias ~ steps x 6e-8: 3.7e-5 after the 2048 K=16 steps of a 32 768-pixel wgrad split), hence
# 5e-5 and not a few ulp for the long-K cfg5 filter gradients; the fp32-parity mode accumulates its two correction
# terms FIRST so that only the main term's steps count, and meets 1e-4 on every shape.
TOL = {"fp32": 1e-4, "bf16": 5e-5}
TOL_LINF = {"fp32": 2e-4, "bf16": 1e-4}
REPORT = {}


def _err(got, ref):
    got, ref = got.double(), ref.double()
    d = got - ref
    return (d.norm() / (ref.norm() + 1e-300)).item(), (d.abs().max() / (ref.abs().max() + 1e-300)).item()


def _report(section, key, **vals):
    REPORT.setdefault(section, {})[key] = vals
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "r2_fullsize_parity.json"), "w") as f:
        json.dump(REPORT, f, indent=1, sort_keys=True)


def _distinct_convs(workload, batch):
    """[(filter_shape, input_shape, stride, border, use_bias, is_first, count)] of a BASELINE workload"""
    from denet_b200.model import model_cnn, recipes
    desc, data_shape, _, classes, convert, _ = recipes.WORKLOADS[workload]
    numpy.random.seed(1)
    m = model_cnn.ModelCNN()
    m.batch_size, m.class_num = batch, classes
    m.build(desc.split(), data_shape, "relu", "half", [0.0])       # zero weights: shapes only, no RNG cost
    if convert:
        m.convert_bn_relu()
    seen = {}
    for l in model_cnn._walk(m.layers):
        if l.type_name == "conv" and l.enabled:
            key = (l.filter_shape, l.input_shape, l.stride, l.border_mode if not isinstance(l.border_mode, list)
                   else tuple(l.border_mode), l.use_bias, bool(l.is_first))
            seen[key] = seen.get(key, 0) + 1
    return [k + (n,) for k, n in seen.items()]


def _conv_cases():
    cases = []
    for workload, batch in (("denet34-skip", 32), ("denet101-wide", 8)):
        for c in _distinct_convs(workload, batch):
            cases.append((workload,) + c)
    return cases


CONV_CASES = _conv_cases()


def _case_id(c):
    w, fs, ishape, st = c[0], c[1], c[2], c[3]
    return "%s-%dx%dx%dx%d-in%dx%d-s%d" % (w.split("-")[0], fs[0], fs[1], fs[2], fs[3], ishape[2], ishape[3], st[0])


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("case", CONV_CASES, ids=_case_id)
def test_conv_at_bench_shape(cuda, case, precision):
    from denet_b200 import layer as layer_mod, ops
    from denet_b200.layer import InitialLayer
    from denet_b200.layer.convolution import ConvLayer
    workload, fs, ishape, stride, border, use_bias, is_first, count = case
    layer_mod.set_precision(precision)
    layer_mod.set_train(True)
    layer_mod.set_wgrad_pending(None)
    try:
        n, ci, h, w = ishape
        co, _, R_, S_ = fs
        g = torch.Generator(device="cuda").manual_seed((co * 131 + ci * 7 + h) % 100003)
        init = InitialLayer(None, ishape)
        init.is_model_input = is_first
        numpy.random.seed(0)
        layer = ConvLayer([init], fs, stride, use_bias, border, 0.0)
        wt = torch.randn(fs, device="cuda", generator=g) / math.sqrt(ci * R_ * S_)
        if is_first:
            x = torch.rand((n, ci, h, w), device="cuda", generator=g)          # images: U(0,1)
        else:
            x = torch.randn((n, ci, h, w), device="cuda", generator=g)
        _, _, oh, ow = layer.output_shape
        dy = torch.randn((n, co, oh, ow), device="cuda", generator=g)
        bias = torch.randn((co,), device="cuda", generator=g) if use_bias else None
        if precision == "bf16":                                                  # same rounded operands on both sides
            x, wt, dy = x.bfloat16().float(), wt.bfloat16().float(), dy.bfloat16().float()
        layer.to("cuda")
        with torch.no_grad():
            layer.omega.copy_(wt)
            if use_bias:
                layer.beta.copy_(bias)
        layer.omega.grad = torch.zeros_like(layer.omega)
        if use_bias:
            layer.beta.grad = torch.zeros_like(layer.beta)
        layer_mod.bump_param_version()
        adt = layer_mod.act_dtype()

        def to_nhwc(t):
            out = ops.alloc_nhwc(t.shape[0], t.shape[2], t.shape[3], t.shape[1], adt, "cuda", zero=True)
            out.copy_(t.permute(0, 2, 3, 1))
            return out
        if layer.rowfold is not None:
            cp, hp, wp = layer.rowfold
            xin = ops.PaddedImage(n, ci, h, w, cp, layer.pad, hp, wp, precision == "fp32", torch.device("cuda")).fill(
                x.contiguous())
        else:
            xin = to_nhwc(x)
        with torch.no_grad():
            y = layer.forward(xin)
            dx = layer.backward(to_nhwc(dy))
        torch.cuda.synchronize()
        # fp64 checker on the GPU (oracle.ref_ops.conv2d: true convolution = correlation with the flipped filter)
        xg = x.double().requires_grad_(not is_first)
        wg = wt.double().requires_grad_(True)
        yref = R.conv2d(xg, wg, stride, border, None if bias is None else bias.double())
        assert tuple(yref.shape) == tuple(layer.output_shape)
        grads = torch.autograd.grad(yref, (wg,) if is_first else (wg, xg), dy.double())
        yref = yref.detach()
        res = {}
        yf = y.float().permute(0, 3, 1, 2)
        if precision == "bf16":
            # the bf16 output is the rounding of an fp32 accumulator: compare against the rounded reference, one ulp
            # of slack comes from accumulation-order noise right at a rounding boundary
            res["y"] = _err(yf, yref.float().bfloat16().float())
            assert res["y"][0] < 3e-3 and res["y"][1] < 8e-3, res
        else:
            res["y"] = _err(yf, yref)
            assert res["y"][0] < TOL[precision] and res["y"][1] < TOL_LINF[precision], res
        res["dw"] = _err(layer.omega.grad, grads[0])
        assert res["dw"][0] < TOL[precision] and res["dw"][1] < TOL_LINF[precision], res
        if use_bias:
            res["dbias"] = _err(layer.beta.grad, dy.double().sum(dim=(0, 2, 3)))
            assert res["dbias"][0] < 1e-4, res
        if not is_first:
            dxf = dx.float().permute(0, 3, 1, 2)
            if precision == "bf16":
                res["dx"] = _err(dxf, grads[1].float().bfloat16().float())
                assert res["dx"][0] < 3e-3 and res["dx"][1] < 8e-3, res
            else:
                res["dx"] = _err(dxf, grads[1])
                assert res["dx"][0] < TOL[precision] and res["dx"][1] < TOL_LINF[precision], res
        _report("conv_" + precision, _case_id(case), layers=count,
                **{k: {"frob": v[0], "linf": v[1]} for k, v in res.items()})
    finally:
        layer_mod.set_precision("bf16")
        layer_mod.set_train(False)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(32, 64, 128, 128), (32, 256, 32, 32), (32, 1536, 24, 24), (8, 2048, 48, 48)])
def test_bn_full_size(cuda, dtype, shape):
    """batch-norm forward / backward (+ residual + ReLU) at cfg3 / cfg5 tensor sizes against the fp64 restatement"""
    from denet_b200 import ops
    n, c, h, w = shape
    g = torch.Generator(device="cuda").manual_seed(c)
    x = (torch.randn(shape, device="cuda", generator=g) * 2 + 0.5).to(dtype).float()
    res = torch.randn(shape, device="cuda", generator=g).to(dtype).float()
    dy = torch.randn(shape, device="cuda", generator=g).to(dtype).float()
    gamma = torch.rand(c, device="cuda", generator=g) + 0.5
    beta = torch.randn(c, device="cuda", generator=g)

    def to_nhwc(t):
        out = ops.alloc_nhwc(n, h, w, c, dtype, "cuda", zero=True)
        out.copy_(t.permute(0, 2, 3, 1))
        return out
    xd, resd, dyd = to_nhwc(x), to_nhwc(res), to_nhwc(dy)
    mean, invstd = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    ops.bn_stats(xd, 1e-5, mean, invstd)
    y = ops.bn_apply(xd, mean, invstd, gamma, beta, residual=resd, relu=True)
    xg, gg, bg, rg = [t.double().requires_grad_(True) for t in (x, gamma, beta, res)]
    yb, m_ref, is_ref = R.batchnorm_train(xg, gg, bg, 1e-5)
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    e_mean, e_is = _err(mean, m_ref.detach()), _err(invstd, is_ref.detach())
    assert e_mean[0] < 1e-5 and e_is[0] < 1e-5
    yn = y.float().permute(0, 3, 1, 2)
    e_y = _err(yn, torch.relu(yb + rg).detach())
    assert e_y[0] < tol
    dgamma, dbeta = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    dx, dres = ops.bn_backward(dyd, y, xd, mean, invstd, gamma, True, dgamma, dbeta, want_dres=True)
    mask = (yn > 0).double()
    dxr, dgr, dbr, drr = torch.autograd.grad((yb + rg) * mask, (xg, gg, bg, rg), dy.double())
    e_dx, e_dr = _err(dx.float().permute(0, 3, 1, 2), dxr), _err(dres.float().permute(0, 3, 1, 2), drr)
    e_dg, e_db = _err(dgamma, dgr), _err(dbeta, dbr)
    assert e_dx[0] < tol and e_dr[0] < tol and e_dg[0] < tol and e_db[0] < tol
    _report("bn_" + ("fp32" if dtype == torch.float32 else "bf16"), "x".join(map(str, shape)),
            mean=e_mean[0], invstd=e_is[0], y=e_y[0], dx=e_dx[0], dres=e_dr[0], dgamma=e_dg[0], dbeta=e_db[0])


def test_pool_inv_full_size(cuda):
    """cfg3's two pool-inv layers, bit-exact (copy / fixed-order 4-term sum) vs the reference's own kernel text"""
    import ctypes
    from denet_b200 import ops
    ref = oracle.reference_cuda()
    for shape in [(32, 512, 16, 16), (32, 256, 32, 32)]:
        n, c, h, w = shape
        g = torch.Generator(device="cuda").manual_seed(h)
        x = torch.rand(shape, device="cuda", generator=g)
        xd = ops.alloc_nhwc(n, h, w, c, torch.float32, "cuda")
        xd.copy_(x.permute(0, 2, 3, 1))
        y = ops.pool_inv_fwd(xd, (2, 2)).permute(0, 3, 1, 2)
        want = x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
        assert torch.equal(y, want)
        dy = torch.rand((n, c, 2 * h, 2 * w), device="cuda", generator=g) - 0.5
        dyd = ops.alloc_nhwc(n, 2 * h, 2 * w, c, torch.float32, "cuda")
        dyd.copy_(dy.permute(0, 2, 3, 1))
        dx = ops.pool_inv_bwd(dyd, (2, 2)).permute(0, 3, 1, 2).contiguous()
        if ref is not None:
            r = torch.empty(shape, device="cuda")
            assert ref.refcuda_pool_inv_bwd_2x2(ctypes.c_void_p(dy.data_ptr()), ctypes.c_void_p(r.data_ptr()), n, c, h,
                                                w) == 0
            assert torch.equal(dx, r)
        else:
            d = dy.view(n, c, h, 2, w, 2)
            assert torch.equal(dx, ((d[:, :, :, 0, :, 0] + d[:, :, :, 0, :, 1]) + d[:, :, :, 1, :, 0]) +
                               d[:, :, :, 1, :, 1])
    _report("pool_inv", "cfg3", bit_exact=True, vs_reference_kernel=ref is not None)


def test_sparse_sample_full_size(cuda):
    """cfg3's gather: B=32, F=96, H=W=64, sn=24, gs=7 (173 MB bf16 / 347 MB fp32 output) vs the reference kernel text"""
    import ctypes
    import random
    from denet_b200 import ops
    ref = oracle.reference_cuda()
    if ref is None:
        pytest.skip("oracle/_ref/libref_cuda_kernels.so not built")
    B, Fc, H, W, sn, gs = 32, 96, 64, 64, 24, 7
    g = torch.Generator(device="cuda").manual_seed(3)
    fmap = torch.rand((B, Fc, H, W), device="cuda", generator=g) * 2 - 1
    rs = numpy.random.RandomState(5)
    x0, y0 = rs.uniform(0, 1, (B, sn, sn)), rs.uniform(0, 1, (B, sn, sn))
    x1, y1 = x0 + (1 - x0) * rs.uniform(0, 1, x0.shape), y0 + (1 - y0) * rs.uniform(0, 1, x0.shape)
    bbox = torch.from_numpy(numpy.stack([x0, y0, x1, y1], axis=-1).astype(numpy.float32)).cuda()
    r = torch.empty(B, gs * gs * Fc + 2, sn, sn, device="cuda")
    vp = ctypes.c_void_p
    assert ref.refcuda_sparse_sample_fwd_7(vp(fmap.data_ptr()), vp(bbox.data_ptr()), vp(r.data_ptr()), B, Fc, H, W, sn) == 0
    fd = ops.alloc_nhwc(B, H, W, Fc, torch.float32, "cuda")
    fd.copy_(fmap.permute(0, 2, 3, 1))
    out = ops.sparse_sample_fwd(fd, bbox, gs)
    assert torch.equal(out.permute(0, 3, 1, 2), r)                      # payload copy + fp32 index math: bit-exact
    dy = torch.rand(r.shape, device="cuda", generator=g) - 0.5
    r2 = torch.empty(B, Fc, H, W, device="cuda")
    assert ref.refcuda_sparse_sample_bwd_7(vp(dy.data_ptr()), vp(bbox.data_ptr()), vp(r2.data_ptr()), B, Fc, H, W, sn) == 0
    dyd = ops.alloc_nhwc(B, sn, sn, gs * gs * Fc + 2, torch.float32, "cuda")
    dyd.copy_(dy.permute(0, 2, 3, 1))
    dfmap = ops.sparse_sample_bwd(dyd, bbox, gs, (B, H, W, Fc))
    e = _err(dfmap.permute(0, 3, 1, 2), r2)
    assert e[0] < 1e-5                                                  # both are unordered fp32 atomic sums
    _report("sparse_sample", "B32_F96_H64_sn24_gs7", fwd_bit_exact=True, bwd_frob=e[0], bwd_linf=e[1])


def test_build_samples_full_size(cuda):
    """cfg3's sampler geometry (B=32, 64x64 maps, sn=24, up to 1024 corners per type) vs the C restatement and, where
    no K-th-score tie is involved, the reference's compiled extension"""
    from test_gpu_kernels import _check_samples
    cp = busy_corner_map(32, 64, 64, 48, seed=21)
    count, ncand = _check_samples(cp, 24)
    assert ncand.max() > 576                       # the top-K cut is exercised
    _report("build_samples", "B32_H64_sn24", images=32, max_candidates=int(ncand.max()), bit_exact=True)
