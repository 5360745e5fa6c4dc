"""ORACLE - TEST INFRASTRUCTURE ONLY.

One training iteration the way the reference runs it on the CPU (ModelCNN.train_step, model_cnn.py:407-445), built
from the restatements in this package:

  1. DeNetSparseLayer.get_samples: a FIRST forward pass up to the corner layer in train mode (the reference compiles a
     separate corner_func for it, denet_sparse.py:117-129), then build_samples - the reference's own C++ extension
     compiled into oracle/_ref when available, else the C restatement (denet_sparse.py:141);
  2. python-side post-processing with the `random` module (denet_sparse.py:184-201), build_bbox_array;
  3. host target builders (denet_corner.py:81-123, denet_detect.py:147-235);
  4. the SECOND forward pass + backward (theano.function 'train_step', model_cnn.py:399) + solver update
     (model_cnn.py:282-305, 320-324) + batch-norm running statistics.

Used by bench.py as the CPU baseline (`cpu_baseline`, `--impl reference`) and by tests as a multi-step checker.
"""
import random

import numpy as np
import torch

from . import build_samples as port_build_samples
from . import reference_cc
from . import ref_ops as R
from .ref_model import RefModel


class RefTrainer:
    def __init__(self, json_layers, input_shape, class_num, solver="nesterov", dtype=torch.float32,
                 use_reference_cc=True, rng=random):
        self.model = RefModel(json_layers, input_shape, class_num, dtype=dtype)
        self.solver = solver
        self.class_num = class_num
        self.rng = rng
        self.input_shape = tuple(input_shape)
        self.kinds = {n.kind: n for n in self.model.nodes}
        self.ref_cc = reference_cc() if use_reference_cc else None
        self.momenta = {name: torch.zeros_like(p) for name, p, _ in self.model.named_params()}
        self.momenta2 = {name: torch.zeros_like(p) for name, p, _ in self.model.named_params()}
        self.sample_impl = "reference denet_sparse.cc (oracle/_ref)" if self.ref_cc is not None else "C restatement"

    def _samples(self, x):
        js = self.kinds["denet-sparse"].js
        with torch.no_grad():
            out = self.model.forward(x, train=True, stop_at_corner=True)      # corner_func: extra forward pass
        cp = np.ascontiguousarray(out["corner_pr"].detach().float().numpy())
        B = cp.shape[0]
        sn = js["sampleNum"]
        if self.ref_cc is not None:
            samples = self.ref_cc.build_samples(B, cp, float(js["cornerThreshold"]), sn, 1024, int(js["localMax"]),
                                                float(js.get("nmsThreshold", 1.0)))
            samples = [list(s) for s in samples]
        else:
            res, _ = port_build_samples(cp, js["cornerThreshold"], sn, 1024, js["localMax"])
            samples = [[(float(s["pr"]), (float(s["x0"]), float(s["y0"]), float(s["x1"]), float(s["y1"])))
                        for s in img] for img in res]
        return samples, cp.shape

    def train_step(self, x, metas, iteration, lr, momentum, decay):
        """-> (cost, [costs]); x numpy (B,C,H,W) fp32"""
        m = self.model
        B = x.shape[0]
        targets, bbox, cost_factors = [], None, None
        if "denet-sparse" in self.kinds:
            sjs = self.kinds["denet-sparse"].js
            cjs = self.kinds["denet-corner"].js
            djs = self.kinds["denet-detect"].js
            samples, cshape = self._samples(x)
            sn = sjs["sampleNum"]
            samples = R.sparse_postprocess(samples, metas, sn * sn, sjs["randomSample"], sjs.get("sampleGT", True),
                                           self.rng)
            bbox = R.bbox_array(samples, B, sn)
            targets.append(R.corner_target(metas, cshape, cjs.get("useCenter", False)))
            targets.append(R.detect_target(metas, samples, B, sn, djs["classNum"], djs["overlapThreshold"],
                                           djs.get("bboxFactor", 0.0) > 0.0))
        elif "regression" in self.kinds:
            out_hw = 1
            idx = [int(meta["image_class"]) + b * self.class_num * out_hw for b, meta in enumerate(metas)]
            targets.append((np.asarray(idx, dtype=np.int64), np.array([], dtype=np.float32)))
        total, costs, grads, _ = m.train_gradients(x, targets, sample_bbox=bbox, cost_factors=cost_factors)
        with torch.no_grad():
            for name, p, is_weight in m.named_params():
                res = R.solver_update(p, grads[name], self.momenta[name], self.solver, iteration, lr, list(momentum),
                                      decay, is_weight, self.momenta2[name])
                p.copy_(res[0])
                self.momenta[name] = res[1]
                if len(res) > 2:
                    self.momenta2[name] = res[2]
            nodes = {}

            def walk(n):
                nodes[n.path] = n
                for c in n.children:
                    walk(c)
            for n in m.nodes:
                walk(n)
            for path, (mean, stdinv) in m.bn_updates.items():
                nodes[path].params["mean"].copy_(mean)
                nodes[path].params["std"].copy_(stdinv)
        return total, costs
