"""The oracle is pinned before it is trusted: its restatements are checked against the golden vectors produced by the
REAL reference (tests/golden/make_golden.py: the reference's compiled denet_sparse.cc), against the reference's one
numeric known-answer (batch_norm.py:153), and against independent formulations of the same op.  CPU only."""
import glob
import math
import os
import subprocess
import sys
import tempfile

import numpy
import pytest
import torch

import oracle
from oracle import ref_ops as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "build_samples_*.npz")))


def test_golden_fixtures_present():
    assert len(FIXTURES) >= 5


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_build_samples_restatement_vs_reference_golden(path):
    """oracle/ref_kernels.c ref_build_samples == the reference's build_samples on the same corner maps"""
    g = numpy.load(path)
    sn = int(g["sample_num"])
    K = sn * sn
    res, ncand = oracle.build_samples(g["corner_pr"], float(g["threshold"]), sn, int(g["max_corners"]),
                                      int(g["local_max"]))
    for b, mine in enumerate(res):
        n = int(g["count"][b])
        assert len(mine) == n
        want_pr, want_bbox = g["pr"][b, :n], g["bbox"][b, :n]
        # the score sequence is identical bit for bit (ties permute boxes, never scores)
        assert numpy.array_equal(mine["pr"], want_pr)
        assert numpy.all(want_pr[:-1] >= want_pr[1:])
        got = {(s["x0"], s["y0"], s["x1"], s["y1"]): s["pr"] for s in mine}
        want = {tuple(want_bbox[i]): want_pr[i] for i in range(n)}
        if ncand[b] <= K:
            assert got == want                       # nothing was cut: the box sets are identical
        else:
            cut = want_pr[-1]                        # boxes tied with the K-th score may legitimately differ
            assert {k: v for k, v in got.items() if v > cut} == {k: v for k, v in want.items() if v > cut}
    # build_bbox_array (denet_sparse.cc:670-699)
    samples = [[(float(g["pr"][b, i]), tuple(g["bbox"][b, i])) for i in range(int(g["count"][b]))]
               for b in range(len(res))]
    assert numpy.array_equal(R.bbox_array(samples, len(res), sn), g["bbox_array"])


NMS_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "nms_*.npz")))


@pytest.mark.parametrize("path", NMS_FIXTURES, ids=[os.path.basename(p)[:-4] for p in NMS_FIXTURES])
def test_nms_restatement_vs_reference_golden(path):
    """oracle.ref_ops.build_detections_nms == the detection lists the reference's compiled denet_detect.cc produced"""
    g = numpy.load(path)
    got = R.build_detections_nms(float(g["pr_threshold"]), float(g["nms_threshold"]), int(g["use_soft_nms"]),
                                 g["det_pr"], g["fitness"], g["bbox"], g["bbox_num"])
    flat = g["bbox"].reshape(g["bbox"].shape[0], -1, 4)
    for b, dets in enumerate(got):
        n = int(g["count"][b])
        assert len(dets) == n
        for i, (logs, cls, k) in enumerate(dets):
            assert cls == g["cls"][b, i] and R.libm_expf(logs) == g["score"][b, i]
            assert numpy.array_equal(flat[b, k], g["box"][b, i])


def test_build_samples_restatement_vs_live_reference():
    """same check against the compiled reference itself when oracle/_ref is present (it is in the build container)"""
    ref = oracle.reference_cc()
    if ref is None:
        pytest.skip("oracle/_ref/denet_sparse*.so not built")
    from util import busy_corner_map
    cp = busy_corner_map(3, 48, 48, 10, seed=21)
    want = ref.build_samples(3, cp, 0.01, 12, 1024, 0, 1.0)
    res, ncand = oracle.build_samples(cp, 0.01, 12)
    for b in range(3):
        assert [numpy.float32(p) for p, _ in want[b]] == list(res[b]["pr"])
        if ncand[b] <= 144:
            assert {tuple(numpy.float32(v) for v in bb) for _, bb in want[b]} == \
                {(s["x0"], s["y0"], s["x1"], s["y1"]) for s in res[b]}


@pytest.mark.parametrize("k,H,sn,cthr,cn", [(12, 32, 4, 0.7, 4), (30, 64, 8, 0.7, 4), (16, 32, 8, 0.3, 4), (14, 32, 5, 0.6, 5)])
def test_clustering_and_centre_restatement_vs_live_reference(k, H, sn, cthr, cn):
    """ref_kernels.c apply_cluster / centre-corner search == the reference's compiled extension (sequence for sequence)"""
    ref = oracle.reference_cc()
    if ref is None:
        pytest.skip("oracle/_ref/denet_sparse*.so not built")
    from util import busy_corner_map
    cp = busy_corner_map(3, H, H, k, seed=200 + k, corner_num=cn)
    want = ref.build_samples(3, cp, 0.01, sn, 1024, 0, cthr)
    res, ncand = oracle.build_samples(cp, 0.01, sn, 1024, 0, cthr)
    assert (ncand > sn * sn).any()
    for b in range(3):
        mine = [(s["pr"], (s["x0"], s["y0"], s["x1"], s["y1"])) for s in res[b]]
        theirs = [(numpy.float32(p), tuple(numpy.float32(v) for v in bb)) for p, bb in want[b]]
        assert mine == theirs


def test_batchnorm_known_answer():
    """reference denet/layer/batch_norm.py:131-153: U(0,1) seed 1002, (64,128,32,32): after one training step the
    mean of the running inverse std is 1.24641 (momentum 0.9 from 1.0), |y.mean| < 1e-4, |y.std - 1| < 1e-4"""
    g = numpy.load(os.path.join(GOLDEN, "bn_known_answer.npz"))
    numpy.random.seed(int(g["seed"]))
    x = torch.from_numpy(numpy.random.uniform(0.0, 1.0, tuple(g["shape"])).astype(numpy.float32))
    c = x.shape[1]
    y, mean, invstd = R.batchnorm_train(x, torch.ones(c), torch.zeros(c), 1e-5)
    assert abs(float(y.mean())) < 1e-4 and abs(float(y.std()) - 1.0) < 1e-4
    run_mean = R.bn_running_update(torch.zeros(c), mean, 0.9)
    run_stdinv = R.bn_running_update(torch.ones(c), invstd, 0.9)
    assert abs(float(run_stdinv.mean()) - float(g["expected_mean_running_stdinv"])) < 5e-6
    assert torch.allclose(run_mean, 0.1 * x.mean(dim=(0, 2, 3)), atol=1e-6)


@pytest.mark.parametrize("case", [(2, 3, 9, 11, 5, 3, 1, "half"), (1, 4, 12, 12, 6, 5, 2, "half"),
                                  (2, 2, 7, 7, 3, 3, 1, "valid"), (1, 3, 8, 8, 4, 3, 1, "full"),
                                  (1, 2, 9, 9, 3, 4, 1, "same")])
def test_conv_restatement_vs_naive_true_convolution(case):
    """ref_ops.conv2d (torch correlation with flipped filters) == a direct loop evaluation of the TRUE convolution
    y[n,o,i,j] = sum x[n,c,i*s+p-a, j*s+q-b] * w[o,c,a,b] that Theano's conv2d defines (convolution.py:83)"""
    n, c, h, w, o, k, s, border = case
    g = torch.Generator().manual_seed(k * 100 + h)
    x = torch.randn(n, c, h, w, generator=g, dtype=torch.float64)
    wt = torch.randn(o, c, k, k, generator=g, dtype=torch.float64)
    y = R.conv2d(x, wt, (s, s), border)
    if border == "same":
        full = R.conv2d_naive(x, wt, (1, 1), (k - 1, k - 1))
        o0 = (k - 1) // 2
        ref = full[:, :, o0:o0 + h, o0:o0 + w]
    else:
        ref = R.conv2d_naive(x, wt, (s, s), R.conv_pad((k, k), border))
    assert tuple(y.shape[2:]) == R.conv_out_hw((h, w), (k, k), (s, s), border)
    assert torch.allclose(y, torch.as_tensor(ref), atol=1e-10)


def test_sparse_sample_c_restatement_vs_torch_restatement():
    """two independent restatements of k_sparse_sample (denet_sparse_op.py:42-85) agree, incl. the .5 rounding cases"""
    rng = numpy.random.RandomState(3)
    B, Fc, H, W, sn, gs = 2, 5, 16, 16, 4, 7
    fmap = rng.randn(B, Fc, H, W).astype(numpy.float32)
    bbox = rng.rand(B, sn, sn, 4).astype(numpy.float32)
    bbox[..., 2:] = bbox[..., :2] + (1 - bbox[..., :2]) * rng.rand(B, sn, sn, 2).astype(numpy.float32)
    bbox[0, 0, :, 0] = numpy.arange(sn) / W
    bbox[0, 0, :, 2] = (numpy.arange(sn) + 4) / W
    a = oracle.sparse_sample_fwd(fmap, bbox, gs)
    b = R.sparse_sample(torch.from_numpy(fmap), bbox, gs).numpy()
    assert numpy.array_equal(a, b)
    dy = rng.randn(*a.shape).astype(numpy.float32)
    d1 = oracle.sparse_sample_bwd(dy, bbox, gs, fmap.shape)
    f = torch.from_numpy(fmap).double().requires_grad_(True)
    (R.sparse_sample(f, bbox, gs) * torch.from_numpy(dy).double()).sum().backward()
    assert numpy.allclose(d1, f.grad.numpy(), rtol=1e-5, atol=1e-6)


def test_pool_inv_restatement():
    rng = numpy.random.RandomState(1)                       # recipe of the reference's A/B block, pool_inv.py:50-57
    x = rng.uniform(-1, 1, (4, 64, 4, 4)).astype(numpy.float32)
    y = oracle.pool_inv_fwd(x, 2, 2)
    assert numpy.array_equal(y, R.pool_inv(torch.from_numpy(x), (2, 2)).numpy())
    dy = rng.uniform(-1, 1, y.shape).astype(numpy.float32)
    dx = oracle.pool_inv_bwd(dy, 2, 2)
    assert numpy.allclose(dx, dy.reshape(4, 64, 4, 2, 4, 2).sum(axis=(3, 5)), rtol=1e-5, atol=1e-6)


def test_expf_glibc_restatement_host():
    """denet_b200/csrc/expf_glibc.cuh (host path) == libm expf bit for bit on 4 M sampled arguments in [0, 90)"""
    src = r'''
#include "expf_glibc.cuh"
#include <stdio.h>
int main() { long bad = 0; for (uint32_t u = 0; u < 0x42b40000u; u += 257) { float x; memcpy(&x, &u, 4);
  float a = expf(x), b = dn::expf_glibc(x); if (memcmp(&a, &b, 4)) bad++; }
  float big = 100.f; if (!(dn::expf_glibc(big) == expf(big))) bad++; printf("%ld\n", bad); return 0; }
'''
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "denet_b200", "csrc")
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.cc"), "w").write(src)
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-I", inc, os.path.join(d, "t.cc"), "-o",
                               os.path.join(d, "t"), "-lm"])
        assert subprocess.check_output([os.path.join(d, "t")]).decode().strip() == "0"


def test_reference_train_step_runs_and_learns():
    """oracle.ref_train.RefTrainer (the bench's CPU arm): a small DeNet trains for a few steps, cost decreases"""
    import random
    from denet_b200.model import model_cnn
    from oracle.ref_train import RefTrainer
    from util import synthetic_metas
    numpy.random.seed(3)
    model = model_cnn.ModelCNN()
    model.batch_size, model.class_num = 2, 5
    desc = ("C.B[8,7,2] BN A P[3,2,1] nRSN.O[1,8,3] SKIPSRC[0] nRSN.O[1,16,3,2] PI[2] C[8,3] SKIP[0] BNA DNC[8,100] "
            "DNS[3,4,0.01,0.1] C[16,1] BNA DND[0.5,1,1]")
    model.build(desc.split(), (3, 64, 64), "relu", "half", ["he-backward"])
    model.convert_bn_relu()
    js = model.export_json()["layers"]
    random.seed(1)
    tr = RefTrainer(js, (2, 3, 64, 64), 5, solver="nesterov", dtype=torch.float64)
    x = numpy.random.uniform(0, 1, (2, 3, 64, 64)).astype(numpy.float32)
    metas = synthetic_metas(2, 5, 2, max_boxes=3)
    costs = [tr.train_step(x, metas, it, 0.02, [0.9, 0.9], 1e-4)[0] for it in range(6)]
    assert all(math.isfinite(c) for c in costs) and costs[-1] < costs[0]


@pytest.mark.parametrize("soft", [0, 1])
def test_nms_restatement_matches_compiled_reference(soft):
    """oracle.ref_ops.build_detections_nms (python restatement of denet_detect.cc) is pinned to the reference's own
    extension compiled unmodified (oracle/_ref): identical lists, scores and boxes bit for bit"""
    import oracle
    cc = oracle.reference_detect_cc()
    if cc is None:
        pytest.skip("oracle/_ref/denet_detect*.so not built")
    from util import nms_inputs
    det_pr, bbox, num = nms_inputs(2, 5, 6, seed=3 + soft)
    fitness = det_pr.copy()
    for pr_thr, nms_thr in [(0.05, 0.5), (0.2, 0.3), (0.01, 1.0)]:
        ref = cc.build_detections_nms(pr_thr, nms_thr, soft, det_pr, fitness, bbox, num)
        mine = R.build_detections_nms(pr_thr, nms_thr, soft, det_pr, fitness, bbox, num)
        assert len(ref) == len(mine) == 2
        for b in range(2):
            assert len(ref[b]) == len(mine[b]) and len(ref[b]) > 0
            for (pr, cls, bb), (logs, mcls, k) in zip(ref[b], mine[b]):
                assert cls == mcls
                assert numpy.float32(pr) == R.libm_expf(logs)
                assert tuple(numpy.float32(v) for v in bb) == tuple(bbox[b].reshape(-1, 4)[k])
