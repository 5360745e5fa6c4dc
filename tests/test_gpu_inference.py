"""Inference tail: detect-layer outputs, per-class NMS / soft-NMS and DeNetDetectLayer.get_detections
on the GPU against the oracle (python restatement of denet_detect.cc, pinned to the reference's own extension compiled
unmodified - used directly too when oracle/_ref travelled to this box)."""
import numpy
import pytest
import torch

import oracle
from oracle import ref_ops as R
from util import nms_inputs, synthetic_metas

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("soft", [0, 1])
@pytest.mark.parametrize("shape", [(8, 80, 24), (2, 20, 48), (3, 5, 6)])
def test_detections_nms_bit_exact(cuda, shape, soft):
    from denet_b200 import common
    B, classes, sn = shape
    ext = common.import_c("denet_detect.cc")
    det_pr, bbox, num = nms_inputs(B, classes, sn, seed=B + sn + soft)
    fitness = (det_pr + numpy.float32(0.25) * numpy.random.RandomState(1).randn(*det_pr.shape).astype(numpy.float32))
    cc = oracle.reference_detect_cc()
    for pr_thr, nms_thr in [(0.05, 0.5), (0.3, 0.3), (0.02, 1.0)]:
        got = ext.build_detections_nms(pr_thr, nms_thr, soft, det_pr, fitness, bbox, num)
        if cc is not None:
            ref = cc.build_detections_nms(pr_thr, nms_thr, soft, det_pr, fitness, bbox, num)
            assert [len(d) for d in got] == [len(d) for d in ref]
            for b in range(B):
                for g, r in zip(got[b], ref[b]):
                    assert g[1] == r[1] and numpy.float32(g[0]) == numpy.float32(r[0]), (g, r)
                    assert tuple(numpy.float32(v) for v in g[2]) == tuple(numpy.float32(v) for v in r[2])
        if B * classes * sn * sn <= 200000:          # the python restatement is slow: small / medium cases only
            want = R.build_detections_nms(pr_thr, nms_thr, soft, det_pr, fitness, bbox, num)
            flat = bbox.reshape(B, -1, 4)
            for b in range(B):
                assert len(got[b]) == len(want[b]) and len(got[b]) > 0
                for g, (logs, cls, k) in zip(got[b], want[b]):
                    assert g[1] == cls and numpy.float32(g[0]) == R.libm_expf(logs)
                    assert tuple(numpy.float32(v) for v in g[2]) == tuple(flat[b, k])
        else:
            assert cc is not None or sum(len(d) for d in got) > 0


def test_detect_outputs_vs_restatement(cuda):
    """log_softmax over the class channels (theano_util.py:27-29) and the box decode (denet_detect.py:87-100)"""
    from denet_b200 import ops
    B, sn, classes = 4, 8, 20
    s0 = classes + 1
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(B, sn, sn, s0 + 4, generator=g) * 2
    logits[..., s0:] *= 0.2
    sb = torch.rand(B, sn, sn, 4, generator=g) * 0.5
    sb[..., 2:] += sb[..., :2] + 0.05
    lg = ops.alloc_nhwc(B, sn, sn, s0 + 4, torch.float32, cuda)
    lg.copy_(logits)
    det_pr, fitness, bbox = ops.detect_outputs(lg, sn, s0, True, sb.to(cuda))
    assert fitness is det_pr
    z = logits.double().permute(0, 3, 1, 2)
    want = R.log_softmax(z[:, :s0], 1)
    assert (det_pr.cpu().double() - want).abs().max() < 1e-5
    reg = z[:, s0:]
    s = sb.double()
    cx, cy, w, h = 0.5 * (s[..., 0] + s[..., 2]), 0.5 * (s[..., 1] + s[..., 3]), s[..., 2] - s[..., 0], s[..., 3] - s[..., 1]
    pcx, pcy = reg[:, 0] * w + cx, reg[:, 1] * h + cy
    pw, ph = torch.exp(reg[:, 2]) * w, torch.exp(reg[:, 3]) * h
    wantb = torch.stack([pcx - pw * 0.5, pcy - ph * 0.5, pcx + pw * 0.5, pcy + ph * 0.5], dim=-1)
    assert (bbox.cpu().double() - wantb).abs().max() < 1e-5
    _, _, plain = ops.detect_outputs(lg, sn, s0, False, sb.to(cuda))
    assert torch.equal(plain.cpu(), sb)


@pytest.mark.parametrize("joint,indfit,use_bbox", [(True, False, False), (True, False, True), (False, True, True),
                                                   (False, True, False)])
def test_detect_outputs_fitness_heads(cuda, joint, indfit, use_bbox):
    """the v2 heads of get_detections (denet_detect.py:332-348 joint fitness, :392-397 independent fitness) against the
    oracle's numpy / torch-CPU float32 restatement (the reference evaluates these in Theano + numpy float32: 1e-5)"""
    from denet_b200 import ops
    B, sn, classes, thr0 = 3, 7, 20, 0.6
    s0 = classes * 5 + 1 if joint else classes + 1
    nout = s0 + (4 if use_bbox else 0) + (6 if indfit else 0)
    g = torch.Generator().manual_seed(9)
    logits = torch.randn(B, sn, sn, nout, generator=g) * 2
    if use_bbox:
        logits[..., s0:s0 + 4] *= 0.2
    sb = torch.rand(B, sn, sn, 4, generator=g) * 0.5
    sb[..., 2:] += sb[..., :2] + 0.05
    lg = ops.alloc_nhwc(B, sn, sn, nout, torch.float32, cuda)
    lg.copy_(logits)
    det_pr, fitness, bbox = ops.detect_outputs(lg, sn, s0, use_bbox, sb.to(cuda), class_num=classes,
                                               fit_mode=(1 if joint else 0) | (2 if indfit else 0), thr0=thr0)
    want_pr, want_fit, want_box = R.detect_outputs(logits.permute(0, 3, 1, 2).numpy(), sb.numpy(), classes, thr0, use_bbox,
                                                   joint, indfit)
    assert det_pr.shape == (B, classes + 1, sn, sn) and fitness.shape == det_pr.shape
    assert numpy.abs(det_pr.cpu().numpy() - want_pr).max() < 1e-5
    nf = want_fit.shape[1]
    assert numpy.abs(fitness.cpu().numpy()[:, :nf] - want_fit).max() < 1e-5
    assert numpy.abs(bbox.cpu().numpy() - want_box).max() < 1e-5
    assert numpy.abs(fitness.cpu().numpy()[:, :classes] - det_pr.cpu().numpy()[:, :classes]).max() > 1e-3


def test_get_detections_end_to_end(cuda):
    """reference entry point DeNetDetectLayer.get_detections(model, data_x, data_m, params): one test-mode pass of a small
    DSS detector whose corner detector fires; the lists are re-derived from the layer's own outputs by the oracle NMS"""
    from test_gpu_model import DENET_SMALL, build
    from denet_b200.layer import get_param, set_param
    model = build(DENET_SMALL.replace("DND[0.5,1,1]", "DND[0.5,1,1]"), (3, 128, 128), 4, 20, "fp32", convert=True)
    dnc = [l for l in model.layers if l.type_name == "denet-corner"][0]
    conv = dnc.layers[1]
    rng = numpy.random.RandomState(4)
    w = get_param(conv.omega).copy()
    w[:4] = rng.randn(*w[:4].shape) * 0.3
    set_param(conv.omega, w)
    b = get_param(conv.beta).copy()
    b[:4] = 2.0
    set_param(conv.beta, b)
    dnd = model.layers[-1]
    wd = get_param(dnd.layers[0].omega).copy()
    wd[:] = rng.randn(*wd.shape) * 0.05                         # a classifier that is not all-null
    set_param(dnd.layers[0].omega, wd)
    model.to_device(precision="fp32")
    numpy.random.seed(3)
    x = numpy.random.uniform(0, 1, (4, 3, 128, 128)).astype(numpy.float32)
    metas = synthetic_metas(4, 20, seed=3, max_boxes=4)
    for soft in (0, 1):
        params = {"prThreshold": 0.03, "nmsThreshold": 0.5, "useSoftNMS": soft}
        res = dnd.get_detections(model, x, metas, params)
        assert len(res) == 4 and all(r["meta"] is metas[i] for i, r in enumerate(res))
        det_pr, fitness, bbox, counts = model.detect_forward(x, dnd)
        counts = counts.cpu().numpy()
        assert counts.max() > 0, "the corner detector of this test is expected to fire"
        want = R.build_detections_nms(0.03, 0.5, soft, det_pr.cpu().numpy(), fitness.cpu().numpy(), bbox.cpu().numpy(),
                                      counts)
        flat = bbox.cpu().numpy().reshape(4, -1, 4)
        total = 0
        for i in range(4):
            dets = res[i]["detections"]
            assert len(dets) == len(want[i])
            total += len(dets)
            for d, (logs, cls, k) in zip(dets, want[i]):
                assert isinstance(d[0], float) and d[1] == cls and numpy.float32(d[0]) == R.libm_expf(logs)
                assert tuple(numpy.float32(v) for v in d[2]) == tuple(flat[i, k])
        assert total > 0
    # sample boxes past the per-image count are zero boxes, the first `count` are the sampler's ranked RoIs
    sp = dnd.sparse_layer
    sb = sp.sample_bbox.cpu().numpy().reshape(4, -1, 4)
    for i in range(4):
        assert (sb[i, counts[i]:] == 0).all()


# ------------------------------------------------------------------------------------------------ golden fixtures
import glob
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NMS_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "nms_*.npz")))
SAMPLE_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "build_samples_*.npz")))


@pytest.mark.parametrize("path", NMS_FIXTURES, ids=[os.path.basename(p)[:-4] for p in NMS_FIXTURES])
def test_detections_nms_vs_reference_golden(cuda, path):
    """the CUDA NMS against detection lists the reference's own compiled extension produced (tests/golden/make_golden.py)"""
    from denet_b200 import common
    g = numpy.load(path)
    got = common.import_c("denet_detect.cc").build_detections_nms(
        float(g["pr_threshold"]), float(g["nms_threshold"]), int(g["use_soft_nms"]), g["det_pr"], g["fitness"], g["bbox"],
        [int(v) for v in g["bbox_num"]])
    for b, dets in enumerate(got):
        n = int(g["count"][b])
        assert len(dets) == n
        for i, (pr, cls, bb) in enumerate(dets):
            assert cls == g["cls"][b, i] and numpy.float32(pr) == g["score"][b, i]
            assert tuple(numpy.float32(v) for v in bb) == tuple(g["box"][b, i])


@pytest.mark.parametrize("path", SAMPLE_FIXTURES, ids=[os.path.basename(p)[:-4] for p in SAMPLE_FIXTURES])
def test_build_samples_vs_reference_golden(cuda, path):
    """the CUDA sampler (4 and 5 corner maps) against the outputs of the reference's compiled build_samples"""
    from denet_b200 import ops
    g = numpy.load(path)
    sn = int(g["sample_num"])
    K = sn * sn
    pr, bbox, ibox, count, ncand = [t.cpu().numpy() for t in ops.build_samples(
        torch.from_numpy(g["corner_pr"]).cuda(), float(g["threshold"]), sn, int(g["max_corners"]), int(g["local_max"]))]
    for b in range(g["corner_pr"].shape[0]):
        n = int(g["count"][b])
        assert count[b] == n
        assert numpy.array_equal(pr[b, :n], g["pr"][b, :n])          # the score sequence is bit-exact
        got = {tuple(bbox[b, i]): pr[b, i] for i in range(n)}
        want = {tuple(g["bbox"][b, i]): g["pr"][b, i] for i in range(n)}
        if ncand[b] <= K:
            assert got == want
        else:
            cut = g["pr"][b, n - 1]                                   # ties at the K-th score may permute boxes
            assert {k: v for k, v in got.items() if v > cut} == {k: v for k, v in want.items() if v > cut}
