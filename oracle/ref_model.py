"""ORACLE - TEST INFRASTRUCTURE ONLY.

Whole-model CPU reference: interprets a model in the reference's JSON layer schema (what every layer's export_json
produces, e.g. convolution.py:125-136, batch_norm.py:109-122, resnet.py:155-169) with torch-CPU ops from
oracle.ref_ops, NCHW, and differentiates the summed cost with autograd - the role theano.function + tensor.grad play
in ModelCNN.build_train_func (model_cnn.py:205-337).  Used to check activations, costs, gradients and one solver
step of the CUDA path, and as the CPU baseline arm of bench.py.
"""
import math

import numpy as np
import torch

from . import ref_ops as R


def _t(a, dtype):
    return torch.tensor(np.asarray(a), dtype=dtype)


class _Node:
    def __init__(self, kind, js, path):
        self.kind, self.js, self.path = kind, js, path
        self.params = {}      # name -> tensor (requires_grad for trainable)
        self.children = []


class RefModel:
    """json_layers: list of layer dicts (reference schema). input_shape: (B, C, H, W)."""

    def __init__(self, json_layers, input_shape, class_num, dtype=torch.float32):
        self.dtype = dtype
        self.input_shape = tuple(input_shape)
        self.class_num = class_num
        self.nodes = [self._build(js, str(i)) for i, js in enumerate(json_layers)]
        # ReLU is discontinuous in its gradient: an fp32 run and this fp64 run disagree on the sign of the few
        # pre-activations that are ~1e-6 from zero, and one flipped mask element moves a gradient by far more than
        # the arithmetic error being tested.  relu_masks {node path: bool (B,C,H,W)} lets a test hand over the masks
        # of the run under test; they are used ONLY where |pre-activation| < relu_mask_tol * max|pre-activation|.
        self.relu_masks = None
        self.relu_mask_tol = 1e-4
        self.relu_overrides = 0
        self.pool_argmax = None      # same idea for max-pool gradient routing: {node path: tap index (B,C,oh,ow)}

    # ------------------------------------------------------------------ construction
    def _build(self, js, path):
        kind = js["type"]
        node = _Node(kind, js, path)
        d = self.dtype
        if kind == "conv":
            node.params["weight"] = _t(js["weight"], d).requires_grad_(True)
            if js.get("useBias", False):
                node.params["bias"] = _t(js["bias"], d).requires_grad_(True)
        elif kind in ("batchnorm", "batchnorm-relu"):
            if js.get("enabled", True):
                node.params["gamma"] = _t(js["gamma"], d).requires_grad_(True)
                node.params["bias"] = _t(js["bias"], d).requires_grad_(True)
                node.params["mean"] = _t(js["mean"], d)
                node.params["std"] = _t(js["std"], d)
        for i, sub in enumerate(js.get("layers", [])):
            node.children.append(self._build(sub, path + "/" + str(i)))
        return node

    def named_params(self, trainable_only=True):
        """[(path.name, tensor, is_weight)] in depth-first JSON order; weights = conv filters (decayed),
        everything else counts as a bias (batch_norm.py:106-107, model_cnn.py:308-324)"""
        out = []

        def walk(n):
            for name, p in n.params.items():
                if trainable_only and not p.requires_grad:
                    continue
                out.append((n.path + "." + name, p, n.kind == "conv" and name == "weight"))
            for c in n.children:
                walk(c)
        for n in self.nodes:
            walk(n)
        return out

    # ------------------------------------------------------------------ layer semantics
    def _relu(self, y, path):
        if not self.relu_masks or path not in self.relu_masks:
            return R.relu(y)
        with torch.no_grad():
            given = torch.as_tensor(np.asarray(self.relu_masks[path])).reshape(y.shape).bool()
            own = y > 0
            ambiguous = y.abs() < self.relu_mask_tol * y.abs().max()
            mask = torch.where(ambiguous, given, own)
            self.relu_overrides += int((mask != own).sum())
        return y * mask.to(y.dtype)

    def _conv(self, node, x):
        js = node.js
        return R.conv2d(x, node.params["weight"], tuple(js.get("stride", (1, 1))), _border(js.get("border", "half")),
                        node.params.get("bias"))

    def _bn(self, node, x, train, apply_relu):
        js = node.js
        if not js.get("enabled", True):
            return x
        p = node.params
        if train:
            y, mean, invstd = R.batchnorm_train(x, p["gamma"], p["bias"], js.get("eps", 1e-5))
            mom = js.get("momentum", 0.9)
            self.bn_updates[node.path] = (R.bn_running_update(p["mean"], mean.detach(), mom),
                                          R.bn_running_update(p["std"], invstd.detach(), mom))
        else:
            y = R.batchnorm_test(x, p["gamma"], p["bias"], p["mean"], p["std"], js.get("eps", 1e-5))
        return self._relu(y, node.path) if apply_relu else y

    def _resnet(self, node, x, train):
        """resnet.py:52-113, 'original' (post-activation) versions incl. the bnrelu-converted form"""
        js = node.js
        version = js.get("version", "original")
        assert "pre-activation" not in version, "oracle: pre-activation resnet blocks not restated"
        assert js.get("activation", "relu") == "relu"
        subs = [c for c in node.children if c.kind not in ("initial", "identity")]
        it = iter(subs)
        y = x
        nconv = 3 if js.get("bottleneck", 0) > 0 else 2
        for ci in range(nconv):
            c = next(it)
            assert c.kind == "conv", c.kind
            y = self._conv(c, y)
            b = next(it)
            last = ci == nconv - 1
            if last:
                assert b.kind == "batchnorm"
                y = self._bn(b, y, train, False)
            elif b.kind == "batchnorm-relu":
                y = self._bn(b, y, train, True)
            else:
                assert b.kind == "batchnorm"
                y = self._bn(b, y, train, False)
                a = next(it)
                assert a.kind == "activation"
                y = self._relu(y, a.path)
        rest = list(it)
        if rest:  # projection shortcut: 1x1 conv (+BN for 'original') on the block input
            s = self._conv(rest[0], x)
            if len(rest) > 1:
                s = self._bn(rest[1], s, train, False)
        else:
            s = x
        return self._relu(s + y, node.path)

    # ------------------------------------------------------------------ forward
    def forward(self, x, sample_bbox=None, train=True, stop_at_corner=False):
        """x: numpy/torch (B,C,H,W).  Returns dict with every head output that exists in the model."""
        h = x if torch.is_tensor(x) else _t(x, self.dtype)
        self.bn_updates = {}
        out = {}
        skips = {}
        self.acts = []
        for node in self.nodes:
            k, js = node.kind, node.js
            if k == "conv":
                h = self._conv(node, h)
            elif k == "batchnorm":
                h = self._bn(node, h, train, False)
            elif k == "batchnorm-relu":
                h = self._bn(node, h, train, True)
            elif k == "activation":
                assert js.get("activation", "relu") == "relu"
                h = self._relu(h, node.path)
            elif k == "pool":
                stride = js["stride"] or js["size"]
                if js.get("mode", "max") == "max" and self.pool_argmax and node.path in self.pool_argmax:
                    h, moved = R.max_pool2d_routed(h, js["size"], stride, js.get("pad", (0, 0)),
                                                   self.pool_argmax[node.path], self.relu_mask_tol)
                    self.relu_overrides += moved
                else:
                    h = R.pool2d(h, js["size"], stride, js.get("pad", (0, 0)), js.get("mode", "max"))
            elif k == "pool-inv":
                h = R.pool_inv(h, js["size"])
            elif k == "resnet":
                h = self._resnet(node, h, train)
            elif k == "skip-src":
                skips[js["index"]] = h
            elif k == "skip":
                s = skips[js["index"]]
                assert js.get("combineMode", "proj-add") == "proj-add"
                convs = [c for c in node.children if c.kind == "conv"]
                h = h + (self._conv(convs[0], s) if convs else s)   # skip.py:78-86
            elif k in ("split", "identity", "initial"):
                pass
            elif k == "denet-corner":
                conv = [c for c in node.children if c.kind == "conv"][0]
                cn = 5 if js.get("useCenter", False) else 4
                o = self._conv(conv, h)
                lh = o[:, :cn]
                lh = torch.stack([lh, -lh], dim=1)                   # denet_corner.py:50-52
                out["corner_pr"] = R.log_softmax(lh, 1)              # (B,2,cn,H,W)
                out["sample"] = o[:, cn:]                            # :58
                out["corner_factor"] = js.get("costFactor", 1.0)
                if stop_at_corner:
                    return out
            elif k == "denet-sparse":
                assert sample_bbox is not None, "sample_bbox needed past the denet-sparse layer"
                out["sample_bbox"] = np.asarray(sample_bbox, np.float32)
                h = R.sparse_sample(out["sample"], out["sample_bbox"], js["gridSize"])
                out["sparse"] = h
            elif k == "denet-detect":
                conv = [c for c in node.children if c.kind == "conv"][0]
                o = self._conv(conv, h)
                joint = js.get("useJointFitness", False)              # denet_detect.py:58-66
                nf = 5 if joint else 6
                s0 = js["classNum"] * nf + 1 if joint else js["classNum"] + 1
                out["det_pr"] = R.log_softmax(o[:, :s0], 1)          # :76-78
                s1 = 0
                if js.get("bboxFactor", 0.0) > 0.0:
                    out["bbox_reg"] = o[:, s0:s0 + 4]
                    s1 = 4
                if js.get("fitnessFactor", 0.0) > 0.0:               # :100-104
                    out["indfit_pr"] = R.log_softmax(o[:, s0 + s1:s0 + s1 + nf], 1)
                out["detect_js"] = js
            elif k == "regression":
                if js.get("valid"):
                    cols = [h[:, :, v[1], v[2]] for v in js["valid"]]
                    xr = torch.stack(cols, dim=2)
                else:
                    xr = h
                out["log_pr"] = R.log_softmax(xr, 1)                 # regression.py:41
                h = torch.exp(out["log_pr"])
                if h.dim() > 2:
                    h = h.mean(dim=tuple(range(2, h.dim())))
            else:
                raise Exception("oracle: layer type not restated: " + k)
            self.acts.append(h)
        out["output"] = h
        return out

    # ------------------------------------------------------------------ costs
    def costs(self, out, targets):
        """targets: list of (yt_index, yt_value) in cost-layer order (corner, detect, regression as present).
        Returns list of scalar costs in the same order (before cost_factors)."""
        costs = []
        ti = 0
        for node in self.nodes:
            k, js = node.kind, node.js
            if k == "denet-corner":
                yt = _t(targets[ti][1], self.dtype).reshape(out["corner_pr"].shape)
                ti += 1
                c = -(yt * out["corner_pr"]).sum(dim=(1, 2, 3, 4)).mean() / math.log(2)   # denet_corner.py:130
                costs.append(js.get("costFactor", 1.0) * c)
            elif k == "denet-detect":
                v = _t(targets[ti][1], self.dtype)
                ti += 1
                B = out["det_pr"].shape[0]
                bf = js.get("bboxFactor", 0.0)
                ff = js.get("fitnessFactor", 0.0)
                det_err, bbox_err, fit_err = R.detect_errors(
                    out["det_pr"], out.get("bbox_reg") if bf > 0.0 else None, out.get("indfit_pr") if ff > 0.0 else None,
                    _t(out["sample_bbox"], self.dtype), v, bf, js.get("useBoundedIoU", False))
                cost = js.get("costFactor", 1.0) * det_err.sum() / B                        # denet_detect.py:308
                if bbox_err is not None:
                    cost = cost + bf * bbox_err.sum() / B                                    # :310 (factor twice)
                if fit_err is not None:
                    cost = cost + ff * fit_err.sum() / B                                     # :312
                costs.append(cost)
            elif k == "regression":
                idx = torch.as_tensor(np.asarray(targets[ti][0]), dtype=torch.long)
                ti += 1
                costs.append(-(out["log_pr"].flatten()[idx]).mean())                        # regression.py:98
        return costs

    def train_gradients(self, x, targets, sample_bbox=None, cost_factors=None):
        """forward (train mode) + summed cost + gradients for every trainable parameter."""
        for _, p, _ in self.named_params():
            p.grad = None
        out = self.forward(x, sample_bbox=sample_bbox, train=True)
        costs = self.costs(out, targets)
        cf = cost_factors or [1.0] * len(costs)
        total = sum(f * c for f, c in zip(cf, costs))
        total.backward()
        grads = {name: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
                 for name, p, _ in self.named_params()}
        return float(total.detach()), [float(c.detach()) for c in costs], grads, out


def _border(b):
    if isinstance(b, (list, tuple)):
        return tuple(b)
    return b
