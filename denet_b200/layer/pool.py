"""Pooling layer 'P' / 'P.A' (reference denet/layer/pool.py:10-69): cuDNN max / average_inc_pad semantics,
output extent floor((in + 2*pad - size)/stride) + 1 (ignore_border) as in pool.py:28-34."""
import math

from .. import ops
from . import AbstractLayer


class PoolLayer(AbstractLayer):
    type_name = "pool"

    def __init__(self, layers, size=(2, 2), stride=None, pad=(0, 0), mode="max", ignore_border=True, json_param={}):
        super().__init__(layer_index=len(layers))
        self.input = layers[-1].output
        self.input_shape = tuple(layers[-1].output_shape)
        size = json_param.get("size", size)
        # README.md:98 - a bare 'P' / 'P.A' pools over the whole input
        self.size = tuple(self.input_shape[2 + i] if size[i] is None else int(size[i]) for i in range(2))
        self.pad = tuple(json_param.get("pad", pad))
        self.ignore_border = json_param.get("ignoreBorder", ignore_border)
        self.mode = json_param.get("mode", mode)
        stride = json_param.get("stride", stride)
        if stride is None or stride[0] is None:
            stride = self.size
        self.stride = tuple(int(v) for v in stride)
        if self.mode not in ("max", "average_inc_pad"):
            raise Exception("unsupported pool mode: " + str(self.mode))
        if self.ignore_border:
            h = int(math.floor((self.input_shape[2] + 2 * self.pad[0] - self.size[0]) / self.stride[0])) + 1
            w = int(math.floor((self.input_shape[3] + 2 * self.pad[1] - self.size[1]) / self.stride[1])) + 1
        else:
            h = int(math.ceil((self.input_shape[2] + 2 * self.pad[0]) / self.stride[0]))
            w = int(math.ceil((self.input_shape[3] + 2 * self.pad[1]) / self.stride[1]))
        self.output_shape = (self.input_shape[0], self.input_shape[1], h, w)
        self._argmax = None
        self.last_argmax = None

    @staticmethod
    def parse_desc(layers, name, tags, params):
        if name != "P":
            return False
        size = (params.get(0), params.get(0))
        stride = (params.get(1, size[0]), params.get(1, size[0]))
        pad = (params.get(2, 0), params.get(2, 0))
        mode = "average_inc_pad" if "A" in tags else "max"
        ignore_border = bool("B" not in tags)
        layers.append(PoolLayer(layers, size, stride, pad, ignore_border=ignore_border, mode=mode))
        return True

    def export_json(self):
        json = super().export_json()
        json.update({"mode": self.mode, "size": self.size, "stride": self.stride, "pad": self.pad,
                     "ignoreBorder": self.ignore_border})
        return json

    def forward(self, x):
        self.input = x
        mode = 0 if self.mode == "max" else 1
        y, self._argmax = ops.pool_fwd(x, mode, self.size, self.stride, self.pad, self.output_shape[2:])
        self._in_shape = tuple(x.shape)
        self.last_argmax = self._argmax    # (N,Ho,Wo,C) uint8 window tap of the maximum; kept for inspection / tests
        self.output = y
        return y

    def backward(self, dy):
        mode = 0 if self.mode == "max" else 1
        dx = ops.pool_bwd(dy, mode, self.size, self.stride, self.pad, self._in_shape, self._argmax)
        self._argmax = None
        return dx
