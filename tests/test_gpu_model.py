"""Whole-model parity: a training step through the Layer API on the CUDA path vs the oracle's RefModel (torch-CPU
fp64 restatement of the Theano graph + autograd), on the same weights (exchanged through the reference's JSON layer
schema), inputs, RoIs and targets.  fp32-parity mode is held to north_star's 1e-4 relative per tensor for the shallow
configs; for the deeper stacks the measured error is asserted against the stated looser bound."""
import random

import numpy
import pytest
import torch

from oracle import ref_ops as R
from oracle.ref_model import RefModel
from util import relerr, synthetic_metas

pytestmark = pytest.mark.gpu

CFG1 = "C[128,3] BN A P[2] C[256,3] BN A P[2] C[512,3] BN A P.A R"                      # BASELINE.json configs[0]
RESNET_SMALL = "C.B[16,7,2] BN A P[3,2,1] nRSN.O[2,16,3] nRSN.O[2,32,3,2] nRSN.O[1,64,3,2,16] P.A[2] R.TB"
DENET_SMALL = ("C.B[32,7,2] BN A P[3,2,1] nRSN.O[1,32,3] nRSN.O[1,64,3,2] SKIPSRC[0] nRSN.O[1,128,3,2] SKIPSRC[1] "
               "nRSN.O[1,256,3,2] PI[2] C[64,3] SKIP[1] BNA PI[2] C[32,3] SKIP[0] BNA DNC[32,100] "
               "DNS[7,8,0.01,0.1] C[128,1] BNA C.B[64,1] BNA DND[0.5,1,1]")


def build(desc, data_shape, batch, classes, precision, convert=False, seed=1):
    from denet_b200.model import model_cnn
    numpy.random.seed(seed)
    model = model_cnn.ModelCNN()
    model.batch_size, model.class_num = batch, classes
    model.build(desc.split(), data_shape, "relu", "half", ["he-backward"])
    if convert:
        model.convert_bn_relu()
    return model


def named_params(model):
    """{RefModel path.name: (param, is_weight)} by the same depth-first walk over the exported layer lists"""
    out = {}

    def walk(layer, path):
        t = layer.type_name
        if t == "conv":
            out[path + ".weight"] = layer.omega
            if layer.use_bias:
                out[path + ".bias"] = layer.beta
        elif t in ("batchnorm", "batchnorm-relu") and layer.enabled:
            out[path + ".gamma"] = layer.omega
            out[path + ".bias"] = layer.beta
            out[path + ".mean"] = layer.mean
            out[path + ".std"] = layer.stdinv
        for i, sub in enumerate(layer.layers):
            walk(sub, path + "/" + str(i))
    for i, l in enumerate(model.layers[1:]):
        walk(l, str(i))
    return out


def relu_masks(model):
    """{RefModel node path: (B,C,H,W) bool} sign of every ReLU input as THIS run saw it - the oracle uses them only for
    pre-activations within 1e-4 of zero (see RefModel.relu_masks), where fp32 and fp64 legitimately disagree"""
    out = {}

    def walk(layer, path):
        if layer.type_name in ("activation", "batchnorm-relu", "resnet") and torch.is_tensor(layer.output) \
                and layer.output.dim() == 4:
            out[path] = (layer.output.float().permute(0, 3, 1, 2) > 0).cpu().numpy()
        for i, sub in enumerate(layer.layers):
            walk(sub, path + "/" + str(i))
    for i, l in enumerate(model.layers[1:]):
        walk(l, str(i))
    return out


def pool_argmax(model):
    """{RefModel node path: (B,C,oh,ow) window tap of each maximum} as THIS run picked them (near-ties only matter)"""
    out = {}
    for i, l in enumerate(model.layers[1:]):
        if l.type_name == "pool" and getattr(l, "last_argmax", None) is not None:
            b, c, oh, ow = l.output_shape
            out[str(i)] = l.last_argmax.reshape(b, oh, ow, c).permute(0, 3, 1, 2).cpu().numpy()
    return out


def oracle_targets(model, metas):
    """[(layer type, (yt_index, yt_value))] in cost-layer order, built by the ORACLE's restatement of the reference's
    host target builders from the metas and the RoIs the step used; the targets the CUDA path built for itself
    (on the device, csrc/targets.cu) must equal them bit for bit"""
    out = []
    dns = [l for l in model.layers if l.type_name == "denet-sparse"]
    for l in model.layers:
        if l.type_name == "denet-corner":
            t = R.corner_target(metas, l.corner_shape, l.use_center)
        elif l.type_name == "denet-detect":
            t = R.detect_target(metas, dns[0].sample_bbox_list, l.batch_size, l.sample_num, l.class_num,
                                l.overlap_threshold, l.use_bbox_reg, use_jointfit=l.use_jointfit,
                                use_indfit=l.use_indfit)
        elif l.type_name == "regression":
            classes = l.output_shape[1]
            t = (numpy.array([b * classes + int(m["image_class"]) for b, m in enumerate(metas)], dtype=numpy.int64),
                 numpy.array([], dtype=numpy.float32))
        else:
            continue
        if hasattr(l, "last_target"):
            mine = l.last_target()
            assert numpy.array_equal(mine[1], t[1]), "%s: target built by the CUDA path differs from the oracle" % \
                l.type_name
        out.append((l.type_name, t))
    return out


def run_step(model, x, metas, solver="nesterov", lr=0.05, mom=(0.9, 0.9), decay=1e-4, it=1, seed=5):
    """GPU train step; returns what is needed to replay it on the oracle"""
    js = model.export_json()["layers"]
    before = {k: p.detach().cpu().double().clone() for k, p in named_params(model).items()}
    random.seed(seed)
    numpy.random.seed(seed)
    cost, costs = model.train_step(x, metas, 0, it, lr, list(mom), decay)
    captured = oracle_targets(model, metas)
    return js, before, captured, cost, costs


MODEL_REPORT = {}


def report(name, **vals):
    """measured whole-model errors -> gpurun_out/r2_model_parity.json (committed under profiles/)"""
    import json
    import os
    MODEL_REPORT[name] = vals
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "r2_model_parity.json"), "w") as f:
        json.dump(MODEL_REPORT, f, indent=1, sort_keys=True)


def fp32_noise_floor(model, js, x, targets, grads64, sample_bbox=None, cost_factors=None):
    """What plain fp32 arithmetic (the precision the reference's cuDNN path computes in) costs on this very step: the
    oracle's RefModel evaluated in float32, same ReLU masks / pool taps, against its float64 evaluation.
    Returns {param: rel err}.  A deviation of the CUDA path of the same size as this floor IS parity with an fp32
    reference; the north_star's 1e-4 is asserted wherever the floor allows it."""
    ref32 = RefModel(js, x.shape, model.class_num, dtype=torch.float32)
    ref32.relu_masks = relu_masks(model)
    ref32.pool_argmax = pool_argmax(model)
    kw = {} if cost_factors is None else {"cost_factors": cost_factors}
    _, _, g32, _ = ref32.train_gradients(x, targets, sample_bbox=sample_bbox, **kw)
    return {n: relerr(g32[n], g) for n, g in grads64.items() if g.norm().item() >= 1e-12}


def compare(model, js, before, captured, x, cost, costs, solver, lr, mom, decay, it, tol, sample_bbox=None, tag=None):
    ref = RefModel(js, x.shape, model.class_num, dtype=torch.float64)
    ref.relu_masks = relu_masks(model)
    ref.pool_argmax = pool_argmax(model)
    targets = [t for _, t in captured]
    total, ref_costs, grads, out = ref.train_gradients(x, targets, sample_bbox=sample_bbox)
    floor = fp32_noise_floor(model, js, x, targets, grads, sample_bbox)
    assert abs(cost - total) <= tol * abs(total), (cost, total)
    for c, rc in zip(costs, ref_costs):
        assert abs(c - rc) <= tol * max(abs(rc), 1e-6), (costs, ref_costs)
    mine = named_params(model)
    worst = 0.0
    ref_named = {n: (p, w) for n, p, w in ref.named_params()}
    gmax = max(g.norm().item() for g in grads.values())
    for name, g in grads.items():
        p = mine[name]
        scale = g.norm().item()
        if scale < 1e-12:
            # analytically zero (e.g. a conv bias feeding a batch norm): fp32 leaves cancellation noise only
            assert p.grad.norm().item() < 1e-6 * gmax, name
            continue
        e = relerr(p.grad, g)
        worst = max(worst, e)
        # north_star bound, widened only where a plain fp32 evaluation of the same graph is itself further away
        assert e < max(tol, 3.0 * floor[name]), "gradient of %s: rel err %.3e (fp32 floor %.3e)" % (name, e, floor[name])
        # solver step (momentum buffers start at zero) replayed on the oracle's gradient, and - to pin the update
        # rule itself independently of the gradient error - on the gradient the CUDA path produced
        p_new = R.solver_update(before[name], g, torch.zeros_like(g), solver, it, lr, list(mom), decay,
                                ref_named[name][1])[0]
        assert relerr(p, p_new) < tol, name
        g_mine = p.grad.detach().cpu().double()
        p_own = R.solver_update(before[name], g_mine, torch.zeros_like(g), solver, it, lr, list(mom), decay,
                                ref_named[name][1])[0]
        assert relerr(p, p_own) < 1e-6, name
    for path, (m_new, s_new) in ref.bn_updates.items():
        assert relerr(mine[path + ".mean"], m_new) < tol
        assert relerr(mine[path + ".std"], s_new) < tol
    if tag:
        report(tag, worst_gradient_rel_err=worst, worst_fp32_floor=max(floor.values()),
               over_1e4=sorted(n for n, g in grads.items() if g.norm().item() >= 1e-12 and
                               relerr(mine[n].grad, g) >= 1e-4), tol=tol, cost=cost, oracle_cost=float(total))
    return worst


def test_cfg1_cifar_cnn_train_step_fp32(cuda):
    """BASELINE.json configs[0]: 3-layer CIFAR10 CNN, batch 32, synthetic 3x32x32, 10 classes"""
    model = build(CFG1, (3, 32, 32), 32, 10, "fp32")
    model.to_device(precision="fp32")
    model.build_train_func("sgd", [])
    numpy.random.seed(1)
    x = numpy.random.uniform(0, 1, (32, 3, 32, 32)).astype(numpy.float32)
    metas = [{"image_class": int(c), "bbox": [], "class": []} for c in numpy.random.randint(0, 10, 32)]
    js, before, cap, cost, costs = run_step(model, x, metas, "sgd", 0.1, (0.9, 0.9), 1e-4, 0)
    worst = compare(model, js, before, cap, x, cost, costs, "sgd", 0.1, (0.9, 0.9), 1e-4, 0, tol=1e-4, tag="cfg1_cifar_cnn")
    print("cfg1 worst gradient rel err %.2e" % worst)


@pytest.mark.parametrize("convert", [False, True])
def test_resnet_classifier_train_step_fp32(cuda, convert):
    model = build(RESNET_SMALL, (3, 64, 64), 8, 10, "fp32", convert)
    model.to_device(precision="fp32")
    model.build_train_func("nesterov", [])
    numpy.random.seed(2)
    x = numpy.random.uniform(0, 1, (8, 3, 64, 64)).astype(numpy.float32)
    metas = [{"image_class": int(c), "bbox": [], "class": []} for c in numpy.random.randint(0, 10, 8)]
    js, before, cap, cost, costs = run_step(model, x, metas, "nesterov", 0.05, (0.9, 0.9), 1e-4, 1)
    worst = compare(model, js, before, cap, x, cost, costs, "nesterov", 0.05, (0.9, 0.9), 1e-4, 1, tol=1e-4,
                    tag="resnet_small_convert%d" % int(convert))
    print("resnet worst gradient rel err %.2e" % worst)


def _denet_step(cuda, precision, tol, centre=False, dnd=None):
    from denet_b200.layer import set_param, get_param
    desc = DENET_SMALL.replace("DNC[32,100]", "DNC.C[32,100]") if centre else DENET_SMALL
    if dnd:
        desc = desc.replace("DND[0.5,1,1]", dnd)
    model = build(desc, (3, 128, 128), 4, 20, precision, convert=True)
    if dnd:
        # the detect head starts at zero (denet_detect.py:71): give the v2 losses something to differentiate
        head = [l for l in model.layers if l.type_name == "denet-detect"][0].layers[0]
        rs = numpy.random.RandomState(11)
        set_param(head.omega, (rs.randn(*get_param(head.omega).shape) * 0.05).astype(numpy.float32))
        set_param(head.beta, (rs.randn(*get_param(head.beta).shape) * 0.1).astype(numpy.float32))
    # a corner detector that fires: random corner rows, bias near the decision boundary
    dnc = [l for l in model.layers if l.type_name == "denet-corner"][0]
    assert dnc.corner_num == (5 if centre else 4)
    cn = dnc.corner_num
    conv = dnc.layers[1]
    rng = numpy.random.RandomState(4)
    w = get_param(conv.omega).copy()
    w[:cn] = rng.randn(*w[:cn].shape) * 0.3
    set_param(conv.omega, w)
    b = get_param(conv.beta).copy()
    b[:cn] = 2.0
    set_param(conv.beta, b)
    model.to_device(precision=precision)
    model.build_train_func("nesterov", [1.0, 0.5])
    numpy.random.seed(3)
    x = numpy.random.uniform(0, 1, (4, 3, 128, 128)).astype(numpy.float32)
    metas = synthetic_metas(4, 20, seed=3, max_boxes=4)
    js, before, cap, cost, costs = run_step(model, x, metas, "nesterov", 0.02, (0.9, 0.9), 1e-4, 1)
    dns = [l for l in model.layers if l.type_name == "denet-sparse"][0]
    bbox = dns.sample_bbox_host.astype(numpy.float32).reshape(4, 8, 8, 4)
    return model, js, before, cap, x, cost, costs, bbox


@pytest.mark.parametrize("centre,dnd", [(False, None), (True, None), (False, "DND.J[0.5,1,1]"),
                                        (False, "DND.B[0.5,1,1,0.5]")],
                         ids=["DNC", "DNC.C", "DND.J", "DND.B+fitness"])
def test_denet_train_step_fp32(cuda, centre, dnd):
    """conv stack + skip + pool-inv + DNC/DNS/DND head, forward/backward/update vs the oracle on the same RoIs
    (DNC.C: the v2 corner layer with a fifth, box-centre map feeding sampler, target and cost; DND.J: joint-fitness
    classes; DND.B + fitness factor: bounded-IoU box loss and the independent-fitness head)"""
    model, js, before, cap, x, cost, costs, bbox = _denet_step(cuda, "fp32", 1e-4, centre, dnd)
    ref = RefModel(js, x.shape, 20, dtype=torch.float64)
    ref.relu_masks = relu_masks(model)
    ref.pool_argmax = pool_argmax(model)
    targets = [t for _, t in cap]
    total, ref_costs, grads, out = ref.train_gradients(x, targets, sample_bbox=bbox, cost_factors=[1.0, 0.5])
    floor = fp32_noise_floor(model, js, x, targets, grads, sample_bbox=bbox, cost_factors=[1.0, 0.5])
    assert abs(cost - total) < 1e-4 * abs(total), (cost, total, costs, ref_costs)
    mine = named_params(model)
    worst, over = 0.0, []
    for name, g in grads.items():
        if g.norm().item() < 1e-12:
            continue
        e = relerr(mine[name].grad, g)
        worst = max(worst, e)
        if e >= 1e-4:
            over.append(name)
        assert e < max(1e-4, 3.0 * floor[name]), "gradient of %s: rel err %.3e (fp32 floor %.3e)" % (name, e, floor[name])
    report("denet_small" + ("_centre" if centre else "") + ("_" + dnd.split("[")[0] if dnd else ""), worst_gradient_rel_err=worst, worst_fp32_floor=max(floor.values()), over_1e4=sorted(over),
           tol=1e-4, cost=cost, oracle_cost=float(total))
    print("denet worst gradient rel err %.2e (fp32 floor %.2e)" % (worst, max(floor.values())))


def test_denet_train_step_bf16_tracks_oracle(cuda):
    """throughput mode (bf16 activations / MMA): costs within 2 %, gradients well aligned with the fp64 oracle"""
    model, js, before, cap, x, cost, costs, bbox = _denet_step(cuda, "bf16", None)
    ref = RefModel(js, x.shape, 20, dtype=torch.float64)
    ref.relu_masks = relu_masks(model)
    ref.pool_argmax = pool_argmax(model)
    ref.relu_mask_tol = 2e-2     # bf16 activations: 8 mantissa bits
    targets = [t for _, t in cap]
    total, ref_costs, grads, out = ref.train_gradients(x, targets, sample_bbox=bbox, cost_factors=[1.0, 0.5])
    assert abs(cost - total) < 2e-2 * abs(total), (cost, total)
    mine = named_params(model)
    cos = []
    for name, g in grads.items():
        if g.norm().item() < 1e-12 or g.numel() < 64:
            continue
        a = mine[name].grad.detach().cpu().double().flatten()
        cos.append(torch.dot(a, g.flatten()).item() / (a.norm().item() * g.norm().item() + 1e-30))
    assert min(cos) > 0.95 and sum(cos) / len(cos) > 0.99, (min(cos), sum(cos) / len(cos))


def test_json_roundtrip_and_predict(cuda):
    from denet_b200.model import model_cnn
    model = build(CFG1, (3, 32, 32), 4, 10, "fp32")
    model.to_device(precision="fp32")
    x = numpy.random.RandomState(0).uniform(0, 1, (4, 3, 32, 32)).astype(numpy.float32)
    p0 = model.predict_output_step(x)
    assert p0.shape == (4, 10) and numpy.allclose(p0.sum(axis=1), 1.0, atol=1e-5)
    js = model.export_json()
    m2 = model_cnn.load_from_json(js, batch_size=4)
    m2.to_device(precision="fp32")
    assert numpy.array_equal(m2.predict_output_step(x), p0)
    ref = RefModel(js["layers"], x.shape, 10, dtype=torch.float64)
    out = ref.forward(x, train=False)
    assert relerr(p0, out["output"].detach()) < 1e-4


def _steps(model, x, metas, n, graphs):
    import random as _r
    _r.seed(11)
    model.enable_cuda_graphs(graphs)
    out = []
    for it in range(n):
        out.append(model.train_step(x, metas, 0, it, 0.002, [0.9, 0.9], 1e-4))   # small lr: the tiny models are chaotic
    return out


@pytest.mark.parametrize("which", ["denet", "classifier"])
def test_cuda_graph_step_equals_eager_step(cuda, which):
    """the captured-graph training step (ModelCNN.enable_cuda_graphs) follows the eager step.  With the batch-norm
    statistics taken by the deterministic two-stage reduction the classifier must match BIT FOR BIT (costs and
    parameters after 3 steps); the conv-epilogue statistics (fp32 atomics, the throughput default) and the sparse
    scatter of the DeNet head are order-nondeterministic in both modes, so there the graphed run only has to stay within
    the run-to-run spread of these tiny, chaotic models"""
    from denet_b200 import layer as layer_mod

    def make():
        if which == "denet":
            m = build(DENET_SMALL, (3, 128, 128), 4, 20, "bf16", convert=True)
        else:
            m = build(RESNET_SMALL, (3, 64, 64), 8, 10, "bf16", False)
        m.to_device(precision="bf16")
        m.build_train_func("nesterov", [])
        return m
    numpy.random.seed(3)
    if which == "denet":
        x = numpy.random.uniform(0, 1, (4, 3, 128, 128)).astype(numpy.float32)
        metas = synthetic_metas(4, 20, seed=3, max_boxes=4)
    else:
        x = numpy.random.uniform(0, 1, (8, 3, 64, 64)).astype(numpy.float32)
        metas = [{"image_class": int(c), "bbox": [], "class": []} for c in numpy.random.randint(0, 10, 8)]
    try:
        if which == "classifier":
            layer_mod.set_fuse_bn_stats(False)
            a = make()
            ca = _steps(a, x, metas, 3, graphs=False)
            b = make()
            cb = _steps(b, x, metas, 3, graphs=True)
            assert b._graphs is not None, "the graphs were not captured"
            assert ca == cb, (ca, cb)
            pa, pb = named_params(a), named_params(b)
            for name in pa:
                assert torch.equal(pa[name], pb[name]), name
        layer_mod.set_fuse_bn_stats(True)
        a = make()
        ca = _steps(a, x, metas, 3, graphs=False)
        a2 = make()
        ca2 = _steps(a2, x, metas, 3, graphs=False)
        b = make()
        cb = _steps(b, x, metas, 3, graphs=True)     # step 0 eager warm-up, step 1 captures + replays, step 2 replays
        assert b._graphs is not None, "the graphs were not captured"
        for (t0, _), (t1, _), (t2, _) in zip(ca, ca2, cb):
            assert abs(t0 - t2) <= 3 * abs(t0 - t1) + 0.1 * abs(t0) + 1e-3, (ca, ca2, cb)
        pa, pa2, pb = named_params(a), named_params(a2), named_params(b)
        for name in pa:
            if pa[name].norm().item() > 1e-2:
                assert relerr(pb[name], pa[name]) < 3 * relerr(pa2[name], pa[name]) + 5e-2, name
    finally:
        layer_mod.set_fuse_bn_stats(True)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_one_launch_operand_preparation_matches_per_layer_kernels(cuda, precision):
    """ModelCNN.prepare_operands (all conv layers in one launch, denet_conv_weight_prep_multi) writes exactly the
    operands of the per-layer kernels: fprop / dgrad layouts, the row-folded stem, bf16 hi (+ lo in parity mode)"""
    from denet_b200 import ops
    model = build(DENET_SMALL, (3, 128, 128), 2, 20, precision, convert=True)
    model.to_device(precision=precision)
    model.prepare_operands()
    split = precision == "fp32"
    seen = 0
    for l in model._prep_layers:
        if l.rowfold is not None:
            ref = [(l._wop_f, ops.conv_weight_prep_rowfold(l.omega, l.rowfold[0], split))]
        else:
            ref = [(l._wop_f, ops.conv_weight_prep(l.omega, 0, split))]
            if not l.is_first and l.dgrad_classes() is not None:
                # strided conv: one dgrad operand per parity class, B[ci][(tr,ts)][co] = W[co][ci][r0+sh*tr][s0+sw*ts]
                cout, cin = l.filter_shape[:2]
                for (a, b, r0, s0, rc, sc) in [c[:6] for c in l.dgrad_classes()]:
                    sub = l.omega.detach()[:, :, r0::l.stride[0], s0::l.stride[1]]
                    assert tuple(sub.shape[2:]) == (rc, sc)
                    want = torch.zeros((cin, rc * sc, (cout + 63) // 64 * 64), device=sub.device)
                    want[:, :, :cout] = sub.permute(1, 2, 3, 0).reshape(cin, rc * sc, cout)
                    hi = want.bfloat16()
                    lo = (want - hi.float()).bfloat16() if split else None
                    ref.append((l._wop_dc[(a, b)], ops.ConvOperand(hi, lo, cin, cout, rc, sc)))
            elif not l.is_first:
                ref.append((l._wop_d, ops.conv_weight_prep(l.omega, 1, split)))
        for got, want in ref:
            assert torch.equal(got.hi, want.hi), l.filter_shape
            assert (got.lo is None) == (want.lo is None)
            if split:
                assert torch.equal(got.lo, want.lo), l.filter_shape
            seen += 1
    assert seen > 20


def test_denet101_wide_recipe_trains(cuda):
    """BASELINE.json configs[4]: the DeNet-101 'wide' recipe (bottleneck ResNet-101, three skip levels, stride-4 corner
    map, 48 x 48 RoIs per image) builds from its model-desc string and takes training steps at a reduced image size"""
    from denet_b200.model import recipes
    desc, _, _, classes, convert, solver = recipes.WORKLOADS["denet101-wide"]
    model = build(desc, (3, 256, 256), 2, classes, "bf16", convert=convert)
    model.to_device(precision="bf16")
    model.build_train_func(solver, [])
    assert [l.output_shape for l in model.layers if l.type_name == "denet-sparse"] == [(2, 7 * 7 * 128 + 2, 48, 48)]
    numpy.random.seed(5)
    x = numpy.random.uniform(0, 1, (2, 3, 256, 256)).astype(numpy.float32)
    metas = synthetic_metas(2, classes, seed=5, max_boxes=4)
    costs = [model.train_step(x, metas, 0, it, 0.01, [0.9, 0.9], 1e-4)[0] for it in range(3)]
    assert all(numpy.isfinite(c) for c in costs), costs
    assert costs[-1] < 1.02 * costs[0], costs
