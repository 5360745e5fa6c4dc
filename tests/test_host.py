"""Host-side logic of the Layer API mirror on the CPU: model-desc grammar and shapes, the reference's recipes, JSON
round trip, the host target builders and the python-`random` post-processing against the oracle's restatements."""
import math
import random

import numpy
import pytest
import torch

from oracle import ref_ops as R
from util import synthetic_metas


def build(desc, shape, batch, classes, convert=False, seed=1):
    from denet_b200.model import model_cnn
    numpy.random.seed(seed)
    m = model_cnn.ModelCNN()
    m.batch_size, m.class_num = batch, classes
    m.build(desc.split(), shape, "relu", "half", ["he-backward"])
    if convert:
        m.convert_bn_relu()
    return m


def test_recipes_shapes_and_flops():
    """the BASELINE.json workloads (cfg1, cfg2, cfg3/4, cfg5) parse through parse_desc; conv FLOPs match SURVEY.md §8d"""
    from denet_b200.model import model_cnn, recipes
    want = {"cifar-cnn": ((32, 10), 0.920e9), "resnet34": ((256, 1000), 21.74e9), "denet34-skip": (None, 163.05e9),
            "denet101-wide": (None, 841.85e9)}
    for name, (desc, shape, batch, classes, convert, _) in recipes.WORKLOADS.items():
        b = 2 if name != "cifar-cnn" else batch
        m = build(desc, shape, b, classes, convert)
        convs = [l for l in model_cnn._walk(m.layers) if l.type_name == "conv" and l.enabled]
        fwd = sum(l.fprop_flops() for l in convs) / b
        train = 3 * fwd - sum(l.fprop_flops() for l in convs if l.is_first) / b
        assert abs(train - want[name][1]) / want[name][1] < 2e-3, (name, train)
        if name == "denet34-skip":
            assert m.layers[-1].type_name == "denet-detect"
            dns = [l for l in m.layers if l.type_name == "denet-sparse"][0]
            assert dns.output_shape == (b, 7 * 7 * 96 + 2, 24, 24)
            dnc = [l for l in m.layers if l.type_name == "denet-corner"][0]
            assert dnc.corner_shape == (b, 2, 4, 64, 64)
            assert m.get_parameter_num() > 32e6
        elif name == "denet101-wide":
            dns = [l for l in m.layers if l.type_name == "denet-sparse"][0]
            assert dns.output_shape == (b, 7 * 7 * 128 + 2, 48, 48)
            dnc = [l for l in m.layers if l.type_name == "denet-corner"][0]
            assert dnc.corner_shape == (b, 2, 4, 128, 128)
        else:
            assert tuple(m.get_output_shape()) == (b, classes) or tuple(m.get_output_shape())[:2] == (b, classes)


def test_desc_grammar_tags_and_errors():
    from denet_b200.model import model_cnn
    m = build("C.B[8,3] BN A P[2] C.X[4,3,1,2,1] BNA PI[2] SPLIT R", (3, 16, 16), 2, 5)
    types = [l.type_name for l in m.layers[1:]]
    assert types[:4] == ["conv", "batchnorm", "activation", "pool"]
    assert m.layers[1].use_bias and m.layers[1].filter_shape == (8, 3, 3, 3)
    assert m.layers[5].filter_shape == (4, 8, 3, 1) and m.layers[5].stride == (2, 1)
    with pytest.raises(Exception):
        build("C[8,3] NOPE[1]", (3, 8, 8), 1, 2)


def test_json_roundtrip_on_host():
    from denet_b200.model import model_cnn
    m = build("C.B[8,7,2] BN A P[3,2,1] nRSN.O[1,8,3] nRSN.O[1,16,3,2] P.A[2] R.TB", (3, 32, 32), 2, 4)
    js = m.export_json()
    m2 = model_cnn.load_from_json(js, batch_size=2)
    js2 = m2.export_json()
    assert [l["type"] for l in js["layers"]] == [l["type"] for l in js2["layers"]]
    w1 = numpy.asarray(js["layers"][0]["weight"])
    assert numpy.array_equal(w1, numpy.asarray(js2["layers"][0]["weight"])) and w1.shape == (8, 3, 7, 7)


def _denet(batch=3, classes=6, dnd="DND[0.5,1,1]"):
    desc = ("C.B[8,7,2] BN A P[3,2,1] nRSN.O[1,8,3] SKIPSRC[0] nRSN.O[1,16,3,2] PI[2] C[8,3] SKIP[0] BNA DNC[8,100] "
            "DNS[3,4,0.01,0.25] C[16,1] BNA " + dnd)
    return build(desc, (3, 64, 64), batch, classes, True)


def test_host_target_builders_match_oracle():
    m = _denet()
    metas = synthetic_metas(3, 6, seed=4, max_boxes=5)
    metas[1]["bbox"], metas[1]["class"] = [], []
    dnc = [l for l in m.layers if l.type_name == "denet-corner"][0]
    dns = [l for l in m.layers if l.type_name == "denet-sparse"][0]
    dnd = [l for l in m.layers if l.type_name == "denet-detect"][0]
    assert numpy.array_equal(dnc.get_target_host(metas)[1], R.corner_target(metas, dnc.corner_shape, False)[1])
    rnd = random.Random(2)
    k = dns.sample_count
    pr = numpy.zeros((3, k))
    bbox = numpy.zeros((3, k, 4))
    for b in range(3):
        for i in range(k):
            if metas[b]["bbox"] and i % 2 == 0:
                g = metas[b]["bbox"][i % len(metas[b]["bbox"])]
                bbox[b, i] = [g[0] + rnd.uniform(-.02, .02), g[1], g[2], g[3] + rnd.uniform(-.02, .02)]
            else:
                x0, y0 = rnd.uniform(0, 1), rnd.uniform(0, 1)
                bbox[b, i] = [x0, y0, rnd.uniform(x0, 1), rnd.uniform(y0, 1)]
    dns.sample_pr_host, dns.sample_bbox_host = pr, bbox
    samples = [[(0.0, tuple(bbox[b, i])) for i in range(k)] for b in range(3)]
    want = R.detect_target(metas, samples, 3, dns.sample_num, 6, 0.5, True)[1]
    got = dnd.get_target_host(metas)[1]
    assert numpy.array_equal(got, want)
    assert (want[:3 * 7 * k].reshape(3, 7, k)[:, :6] > 0).any()


@pytest.mark.parametrize("dnd,joint,indfit", [("DND.J[0.5,1,1]", True, False), ("DND.JB[0.6,1,1]", True, False),
                                              ("DND[0.5,1,1,0.5]", False, True), ("DND.B[0.5,1,0,2]", False, True)])
def test_host_detect_target_v2_matches_oracle(dnd, joint, indfit):
    """joint-fitness / independent-fitness targets (denet_detect.py:58-66,177-191) of the host builder, the head's
    channel count, and the export keys the reference writes (:131-139)"""
    m = _denet(dnd=dnd)
    l = [x for x in m.layers if x.type_name == "denet-detect"][0]
    dns = [x for x in m.layers if x.type_name == "denet-sparse"][0]
    use_bbox = l.bbox_factor > 0
    s0 = 6 * 5 + 1 if joint else 7
    assert l.layers[0].filter_shape[0] == s0 + (4 if use_bbox else 0) + (6 if indfit else 0)
    assert l.use_jointfit == joint and l.use_indfit == indfit and l.use_bounded_iou == ("B" in dnd.split("[")[0])
    js = l.export_json()
    assert js["useJointFitness"] == joint and js["useBoundedIoU"] == l.use_bounded_iou
    assert js["fitnessFactor"] == l.indfit_factor
    from denet_b200.model import model_cnn
    l2 = [x for x in model_cnn.load_from_json(m.export_json(), batch_size=3).layers if x.type_name == "denet-detect"][0]
    assert (l2.use_jointfit, l2.use_indfit, l2.use_bounded_iou) == (l.use_jointfit, l.use_indfit, l.use_bounded_iou)
    assert l2.layers[0].filter_shape == l.layers[0].filter_shape and l2.det_shape == l.det_shape
    metas = synthetic_metas(3, 6, seed=7, max_boxes=5)
    rnd = random.Random(3)
    k = dns.sample_count
    bbox = numpy.zeros((3, k, 4))
    for b in range(3):
        for i in range(k):
            if metas[b]["bbox"] and i % 3 != 2:
                g = metas[b]["bbox"][i % len(metas[b]["bbox"])]
                j = 0.06 * (i % 5) / 4.0                    # IoU spread over the fitness bins
                bbox[b, i] = [g[0] + rnd.uniform(-j, j), g[1] + rnd.uniform(-j, j), g[2], g[3] + rnd.uniform(-j, j)]
            else:
                x0, y0 = rnd.uniform(0, 1), rnd.uniform(0, 1)
                bbox[b, i] = [x0, y0, rnd.uniform(x0, 1), rnd.uniform(y0, 1)]
    dns.sample_pr_host, dns.sample_bbox_host = numpy.zeros((3, k)), bbox
    samples = [[(0.0, tuple(bbox[b, i])) for i in range(k)] for b in range(3)]
    thr = l.overlap_threshold
    want = R.detect_target(metas, samples, 3, dns.sample_num, 6, thr, use_bbox, use_jointfit=joint, use_indfit=indfit)[1]
    got = l.get_target_host(metas)[1]
    assert got.shape == want.shape and numpy.array_equal(got, want)
    det = want[:3 * s0 * k].reshape(3, s0, k)
    assert (det[:, :s0 - 1] > 0).any()
    if joint:
        assert len(set(numpy.nonzero(det[:, :s0 - 1])[1] % 5)) >= 3, "several fitness bins must occur"


@pytest.mark.parametrize("counts", [[0, 3, 16], [16, 16, 16], [2, 14, 5], [13, 16, 1]])
def test_sparse_postprocess_consumes_python_random_like_the_reference(counts, monkeypatch):
    """DeNetSparseLayer.finish_target == the oracle's loop restatement of denet_sparse.py:184-201, including the exact
    consumption of python's `random` stream (both the vectorised whole-batch path and the per-image path)"""
    m = _denet()
    dns = [l for l in m.layers if l.type_name == "denet-sparse"][0]
    k = dns.sample_count                                   # 16, random_sample 0.25 -> keep at most 12
    metas = synthetic_metas(3, 6, seed=9, max_boxes=3)
    rs = numpy.random.RandomState(1)
    pr32 = numpy.sort(rs.rand(3, k).astype(numpy.float32), axis=1)[:, ::-1].copy()
    bbox32 = rs.rand(3, k, 4).astype(numpy.float32)
    captured = {}
    monkeypatch.setattr(dns, "set_samples_arrays", lambda pr, bbox: captured.update(pr=pr.copy(), bbox=bbox.copy()))
    random.seed(77)
    dns.finish_target(metas, pr32, bbox32, numpy.array(counts))
    after_mine = random.random()
    lists = [[(float(pr32[b, i]), tuple(float(v) for v in bbox32[b, i])) for i in range(counts[b])] for b in range(3)]
    random.seed(77)
    want = R.sparse_postprocess(lists, metas, k, dns.random_sample, True, random)
    after_ref = random.random()
    assert after_mine == after_ref, "python random stream consumed differently"
    for b in range(3):
        assert len(want[b]) == k
        for i in range(k):
            assert captured["pr"][b, i] == want[b][i][0]
            assert tuple(captured["bbox"][b, i]) == tuple(want[b][i][1])


def test_py_random_doubles_is_pythons_stream():
    from denet_b200.layer.denet_sparse import py_random_doubles
    for k in (1, 31, 32, 1000):
        random.seed(5)
        a = py_random_doubles(k)
        nxt = random.random()
        random.seed(5)
        b = [random.random() for _ in range(k)]
        assert list(a) == b and nxt == random.random()


def test_no_cpu_fallback():
    from denet_b200 import lib, ops
    m = _denet()
    if not torch.cuda.is_available():
        with pytest.raises(lib.DenetError):
            m.to_device()
    with pytest.raises(lib.DenetError):
        ops.act_operand(torch.zeros(1, 2, 2, 8))


def test_native_pyrandom_matches_the_interpreter():
    """csrc/pyrandom.cu restates CPython's Mersenne Twister, random.random, _randbelow and both variants of
    random.sample: same results AND the same stream position afterwards, for populations around powers of two"""
    from denet_b200.layer.denet_sparse import py_random_sample, mt_export, mt_import
    from denet_b200 import lib
    import ctypes
    for seed, (n, k) in enumerate([(576, 519), (576, 576), (1, 1), (2, 1), (512, 100), (513, 3), (64, 5), (2304, 2074),
                                   (5000, 6), (100000, 50), (1023, 1000), (21, 21), (22, 5), (30, 0)]):
        random.seed(seed)
        want = random.sample(range(n), k)
        after = random.random()
        random.seed(seed)
        got = py_random_sample(n, k)
        assert list(got) == want, (n, k)
        assert random.random() == after, (n, k)
    random.seed(9)
    want = [random.random() for _ in range(1500)]      # crosses a 624-word regeneration
    random.seed(9)
    mt, pos, version, gauss = mt_export()
    out = numpy.empty(1500)
    lib.call("denet_pyrandom_random", mt.ctypes.data, ctypes.addressof(pos), 1500, out.ctypes.data)
    mt_import(mt, pos, version, gauss)
    assert list(out) == want and random.random() == [random.seed(9), [random.random() for _ in range(1501)]][1][-1]


def test_strided_dgrad_parity_classes_restate_the_data_gradient():
    """ops.dgrad_parity_classes: the data gradient of a strided (true) convolution equals one stride-1 correlation per
    parity class of the input pixel on the UNDILATED dy, with the class's taps r0 + s*t and leading pad - checked on the
    CPU against autograd of the oracle convolution (the CUDA path uses exactly these class parameters)"""
    import torch
    import torch.nn.functional as F
    from denet_b200 import ops
    from oracle import ref_ops as R
    g = torch.Generator().manual_seed(7)
    for (h, w, k, s, pad) in [(16, 16, 3, 2, 1), (15, 18, 3, 2, 1), (20, 20, 5, 2, 2), (21, 17, 7, 2, 3), (18, 18, 3, 3, 1),
                              (16, 16, 4, 2, 1)]:
        cin, cout, n = 3, 4, 2
        x = torch.randn(n, cin, h, w, generator=g, dtype=torch.float64, requires_grad=True)
        wt = torch.randn(cout, cin, k, k, generator=g, dtype=torch.float64)
        y = R.conv2d(x, wt, (s, s), pad, None)
        dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
        (dx_ref,) = torch.autograd.grad(y, x, dy)
        classes = ops.dgrad_parity_classes((h, w), (k, k), (s, s), (pad, pad))
        assert classes is not None and len(classes) == s * s
        dx = torch.zeros_like(dx_ref)
        for (a, b, r0, s0, rc, sc, ph, pw, hc, wc) in classes:
            sub = wt[:, :, r0::s, s0::s]                       # (cout, cin, rc, sc): taps of this class
            assert tuple(sub.shape[2:]) == (rc, sc)
            kern = sub.permute(1, 0, 2, 3)                      # correlation dy (cout channels) -> dx (cin channels)
            need_h, need_w = hc + rc - 1, wc + sc - 1           # out[i] = sum_t in[i + t - ph] K[t], i < hc
            dyp = F.pad(dy, (pw, max(0, need_w - pw - dy.shape[3]), ph, max(0, need_h - ph - dy.shape[2])))
            out = F.conv2d(dyp[:, :, :need_h, :need_w], kern)
            assert tuple(out.shape[2:]) == (hc, wc)
            dx[:, :, a::s, b::s] = out
        assert float((dx - dx_ref).abs().max()) < 1e-10, (h, w, k, s, pad)


@pytest.mark.parametrize("counts", [[0, 0, 0], [10, 300, 64], [64, 64, 64, 64], [700, 0, 33]])
def test_random_ahead_equals_live_generator(counts):
    """finish_target's generator work done AHEAD of time (denet_pyrandom_ahead + denet_sparse_postprocess_ahead) yields
    the same boxes and leaves python's `random` in the same state as the live path - for any number of words consumed,
    also across several 624-word regenerations and from any starting position inside a block"""
    import ctypes
    from denet_b200 import lib
    from denet_b200.layer.denet_sparse import RandomAhead, mt_export, mt_import
    B, K, n_keep = len(counts), 64, 58
    rs = numpy.random.RandomState(5)
    pr32 = rs.rand(B, K).astype(numpy.float32)
    bbox32 = rs.rand(B, K, 4).astype(numpy.float32)
    count = numpy.array(counts, dtype=numpy.int64)
    for seed, burn in [(1, 0), (2, 311), (3, 623)]:
        random.seed(seed)
        for _ in range(burn):
            random.getrandbits(32)
        start = random.getstate()
        # live
        pr_a, bb_a = numpy.zeros((B, K)), numpy.zeros((B, K, 4))
        mt, pos, version, gauss = mt_export()
        lib.call("denet_sparse_postprocess", mt.ctypes.data, ctypes.addressof(pos), pr32.ctypes.data, bbox32.ctypes.data,
                 count.ctypes.data, B, K, n_keep, pr_a.ctypes.data, bb_a.ctypes.data)
        mt_import(mt, pos, version, gauss)
        end_live = random.getstate()
        # ahead
        random.setstate(start)
        ahead = RandomAhead(B * 12 * K + 1248)
        assert ahead.still_valid()
        pr_b, bb_b = numpy.zeros((B, K)), numpy.zeros((B, K, 4))
        used = ctypes.c_longlong(0)
        rc = lib.call("denet_sparse_postprocess_ahead", ahead.words.ctypes.data, ahead.nwords, ctypes.addressof(used),
                      pr32.ctypes.data, bbox32.ctypes.data, count.ctypes.data, B, K, n_keep, pr_b.ctypes.data,
                      bb_b.ctypes.data, allow=(1,))
        assert rc == 0
        ahead.commit(int(used.value))
        assert numpy.array_equal(pr_a, pr_b) and numpy.array_equal(bb_a, bb_b)
        assert random.getstate() == end_live
        # a buffer that is too small reports it instead of returning garbage silently
        random.setstate(start)
        small = RandomAhead(0)
        rc = lib.call("denet_sparse_postprocess_ahead", small.words.ctypes.data, min(small.nwords, 40),
                      ctypes.addressof(used), pr32.ctypes.data, bbox32.ctypes.data, count.ctypes.data, B, K, n_keep,
                      pr_b.ctypes.data, bb_b.ctypes.data, allow=(1,))
        assert rc == (1 if sum(K - min(c, n_keep) for c in counts) > 0 else 0)
        random.getrandbits(32)
        assert not small.still_valid()
