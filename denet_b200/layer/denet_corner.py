"""DeNet corner layer 'DNC[sample_feat,cost_factor,dropout]' (reference denet/layer/denet_corner.py:17-134).

A 1x1 convolution with bias maps the backbone features to corner_num (4, or 5 with tag C) corner logits followed by
sample_feat sampling features.  corner_pr = log_softmax([+z, -z]) as (B,2,corner_num,H,W) (denet_corner.py:50-53) feeds
the sparse layer's sampler and the corner cost (-sum(t*logp) per image, batch mean, / ln 2, * cost_factor; :126-134).
Like the reference the layer is a pass-through in the layer list: the sparse layer picks up `sample` from here.
"""
import math

import numpy
import torch

from .. import ops
from . import (AbstractLayer, InitialLayer, d2h, device_targets, get_ground_truth, get_train, h2d, set_param,
               get_param)
from .convolution import ConvLayer


class DeNetCornerLayer(AbstractLayer):
    type_name = "denet-corner"
    has_cost = True

    def __init__(self, layers, sample_feat=512, cost_factor=1, dropout=0.0, use_center=False, json_param={}):
        super().__init__(layer_index=len(layers))
        self.input = layers[-1].output
        self.input_shape = self.output_shape = tuple(layers[-1].output_shape)
        self.batch_size, self.features, self.height, self.width = self.input_shape

        self.sample_feat = json_param.get("sampleFeat", sample_feat)
        self.cost_factor = json_param.get("costFactor", cost_factor)
        self.use_center = json_param.get("useCenter", use_center)
        self.dropout = json_param.get("dropout", dropout)
        self.corner_num = 5 if self.use_center else 4

        self.layers.append(InitialLayer(None, self.input_shape))
        conv = ConvLayer(self.layers, (self.corner_num + self.sample_feat, self.features, 1, 1), (1, 1), True, False)
        conv.out_fp32 = True
        self.layers.append(conv)
        # corner logits start at "almost surely not a corner": zero weights, bias 5 (denet_corner.py:42-47)
        omega = get_param(conv.omega).copy()
        omega[:self.corner_num] = 0.0
        set_param(conv.omega, omega)
        beta = get_param(conv.beta).copy()
        beta[:self.corner_num] = 5.0
        set_param(conv.beta, beta)

        self.corner_shape = (self.batch_size, 2, self.corner_num, self.height, self.width)
        self.sample_shape = (self.batch_size, self.sample_feat, self.height, self.width)
        self.corner_pr = None      # (B,2,corner_num,H,W) fp32 device tensor of the last forward
        self.sample = None         # (B,H,W,sample_feat) view of the conv output
        self.grad_factor = 1.0
        self.cost_value = None
        self._target = None
        self._target_dev = None     # persistent (B,2,cn,H,W) buffer filled by denet_corner_target
        self._z = self._dz = None
        self._sample_grad = None

    @staticmethod
    def parse_desc(layers, name, tags, params):
        if name != "DNC":
            return False
        layers.append(DeNetCornerLayer(layers, params.get(0, 512), params.get(1, 1.0), params.get(2, 0.0),
                                       "C" in tags))
        return True

    def export_json(self):
        json = super().export_json()
        json.update({"sampleFeat": self.sample_feat, "useCenter": self.use_center, "costFactor": self.cost_factor,
                     "dropout": self.dropout})
        return json

    def get_target(self, model, samples, metas):
        """one-hot corner maps of the ground-truth boxes (denet_corner.py:81-123).  With device targets on, the same
        map is produced by denet_corner_target inside forward() and this returns None."""
        if device_targets() and self.dropout == 0.0:
            self._target = None
            return None
        return self.get_target_host(metas)

    def get_target_host(self, metas):
        corner_pr = numpy.zeros(self.corner_shape, dtype=numpy.float32)
        W, H = self.width, self.height
        for b, meta in enumerate(metas):
            for bbox in meta["bbox"]:
                x0 = int(round(bbox[0] * W))
                y0 = int(round(bbox[1] * H))
                x1 = max(x0, int(round(bbox[2] * W)) - 1)
                y1 = max(y0, int(round(bbox[3] * H)) - 1)
                x0v, y0v = 0 <= x0 < W, 0 <= y0 < H
                x1v, y1v = 0 <= x1 < W, 0 <= y1 < H
                if x0v and y0v:
                    corner_pr[b, 1, 0, y0, x0] = 1.0
                if x1v and y0v:
                    corner_pr[b, 1, 1, y0, x1] = 1.0
                if x0v and y1v:
                    corner_pr[b, 1, 2, y1, x0] = 1.0
                if x1v and y1v:
                    corner_pr[b, 1, 3, y1, x1] = 1.0
                if self.use_center:
                    cx = int(round((bbox[0] + bbox[2]) * 0.5 * W))
                    cy = int(round((bbox[1] + bbox[3]) * 0.5 * H))
                    if 0 <= cx < W and 0 <= cy < H:
                        corner_pr[b, 1, 4, cy, cx] = 1.0
        corner_pr[:, 0] = 1.0 - corner_pr[:, 1]
        corner_pr /= W * H * self.corner_num
        if self.dropout > 0.0:
            mask = numpy.random.binomial(1, 1.0 - self.dropout, (self.corner_shape[0],) + self.corner_shape[2:])
            corner_pr *= mask.astype(numpy.float32)[:, None] / (1.0 - self.dropout)
        return numpy.array([], dtype=numpy.int64), corner_pr.flatten()

    def set_target(self, yt_index, yt_value):
        self._target = h2d(numpy.ascontiguousarray(yt_value, dtype=numpy.float32))

    def forward(self, x):
        self.input = self.output = x
        conv = self.layers[1]
        z = conv.forward(x)                                   # (B,H,W,cn+F) fp32
        self._z = z
        self.corner_pr = ops.corner_logprob(z, self.corner_num)
        self.sample = z[..., self.corner_num:]
        if self.cost_value is None:
            self.cost_value = torch.zeros((1,), dtype=torch.float32, device=x.device)
        if get_train():
            if device_targets() and self.dropout == 0.0:
                if self._target_dev is None:
                    self._target_dev = torch.empty(self.corner_shape, dtype=torch.float32, device=x.device)
                self._target = ops.corner_target(get_ground_truth(), self.corner_num, self.height, self.width,
                                                 self._target_dev)
            assert self._target is not None, "denet-corner: get_target/set_target must precede a training forward"
            self._dz = ops.alloc_like(z)
            ops.corner_cost(z, self.corner_num, self._target, float(self.cost_factor), self.grad_factor, self._dz,
                            self.cost_value)
        return x

    def cost(self, yt_index=None, yt_value=None):
        return self.cost_value

    def last_target(self):
        """(yt_index, yt_value) of the last training forward in the reference's format, whichever side built it"""
        return numpy.array([], dtype=numpy.int64), d2h(self._target.reshape(-1))

    def set_sample_grad(self, dfmap):
        """gradient wrt `sample` (B,H,W,F) fp32 from the sparse layer's scatter"""
        self._sample_grad = dfmap

    def backward(self, dy):
        dz = self._dz
        if self._sample_grad is not None:
            ops.convert(self._sample_grad, dz.dtype, out=dz[..., self.corner_num:])
        else:
            dz[..., self.corner_num:].zero_()
        self._sample_grad = None
        dx = self.layers[1].backward(dz)
        self._z = self._dz = None
        self.sample = None
        return dx if dy is None else ops.add(dx, dy)
