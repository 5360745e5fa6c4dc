"""Classification head 'R' (reference denet/layer/regression.py:10-98): a full-extent 'valid' convolution to classNum
channels followed by log-softmax; cost = -mean(log_pr[target]) (regression.py:97-98).  Output = class probabilities.
The kernel computes cost, log-probabilities and the gradient wrt the logits in one pass (csrc/loss.cu)."""
import numpy
import torch

from .. import ops
from . import AbstractLayer, get_train, h2d
from .convolution import ConvLayer


class RegressionLayer(AbstractLayer):
    type_name = "regression"
    has_cost = True

    def __init__(self, layers, use_center=True, valid=[], json_param={}):
        super().__init__(layer_index=len(layers))
        self.input = layers[-1].output
        self.input_shape = tuple(layers[-1].output_shape)
        if use_center:
            valid = [(0, self.input_shape[-2] // 2, self.input_shape[-1] // 2)]
        self.valid = json_param.get("valid", valid)
        if len(self.valid) > 0:
            self.log_pr_shape = (self.input_shape[0], self.input_shape[1], len(self.valid))
        else:
            self.log_pr_shape = self.input_shape
        if tuple(self.input_shape[2:]) != (1, 1):
            raise Exception("regression layer on the B200 hot path expects a 1x1 spatial input, got %s"
                            % str(self.input_shape))
        self.output_shape = (self.log_pr_shape[0], self.log_pr_shape[1])
        self.grad_factor = 1.0
        self._label = None
        self.cost_value = None
        self.log_pr = None
        if isinstance(layers[-1], ConvLayer):
            layers[-1].out_fp32 = True   # logits stay fp32

    @staticmethod
    def parse_desc(layers, name, tags, params):
        if name != "R":
            return False
        use_bias = bool("B" in tags)
        use_center = bool("C" in tags)
        prev = layers[-1].output_shape
        filter_shape = (params["classNum"], prev[1], params.get(0, prev[2]), params.get(0, prev[3]))
        layers.append(ConvLayer(layers, filter_shape, (1, 1), use_bias, "valid", params["wb"]))
        layers.append(RegressionLayer(layers, use_center))
        return True

    def export_json(self):
        json = super().export_json()
        json.update({"valid": self.valid})
        return json

    def get_target(self, model, samples, metas):
        """flat indices of the target class of every sample (regression.py:75-94)"""
        yt_index = [numpy.ravel_multi_index((b, metas[b]["image_class"]), self.output_shape)
                    for b in range(len(metas))]
        return numpy.array(yt_index, dtype=numpy.int64), numpy.array([], dtype=numpy.float32)

    def set_target(self, yt_index, yt_value):
        classes = self.output_shape[1]
        label = (numpy.asarray(yt_index, dtype=numpy.int64) % classes).astype(numpy.int32)
        self._label = h2d(label, slot="regression/label" + self._slot_ns)

    def forward(self, x):
        self.input = x
        b, classes = self.output_shape
        self.log_pr = torch.empty((b, classes), dtype=torch.float32, device=x.device)
        if self.cost_value is None:
            self.cost_value = torch.zeros((1,), dtype=torch.float32, device=x.device)
        if get_train():
            assert self._label is not None, "regression layer: get_target/set_target must precede a training forward"
            self._dlogits = ops.alloc_like(x)
            ops.softmax_nll(x, classes, self._label, self.grad_factor, self._dlogits, self.log_pr, self.cost_value)
            self.output = None
        else:
            label = torch.zeros((b,), dtype=torch.int32, device=x.device)
            ops.softmax_nll(x, classes, label, 0.0, None, self.log_pr, self.cost_value)
            self.output = torch.exp(self.log_pr)   # prediction output only (not on the training path)
        return self.output

    def cost(self, yt_index=None, yt_value=None):
        return self.cost_value

    def backward(self, dy):
        d, self._dlogits = self._dlogits, None
        return d
