"""Batch normalisation layers: BatchNormLayer ('BN', type "batchnorm") and the fused BatchNormReluLayer ('BNA', type
"batchnorm-relu"), with the reference's contract (denet/layer/batch_norm.py:12-128, batch_norm_relu.py:86-168):

  * train mode: per-channel batch mean / biased variance over (N,H,W), y = gamma*(x-mean)*invstd + beta with
    invstd = 1/sqrt(var+eps) (cuDNN spatial BN), followed by ReLU for 'BNA';
  * the running statistics are an EMA of the batch MEAN and of the batch INVERSE STD (batch_norm.py:75-76) and are
    exported under the JSON keys "mean" / "std";
  * test mode: var = (1/stdinv)^2 handed to cuDNN inference, which adds eps again (batch_norm.py:50-52);
  * gradient of 'BNA': dy masked by y > 0, then the BN gradient (batch_norm_relu.py:50-54).

Statistics come either from the producing convolution's epilogue (throughput mode, one pass less over the
activation) or from a deterministic two-stage reduction (parity mode).
"""
import numpy
import torch

from .. import ops
from . import AbstractLayer, fuse_bn_backward, fuse_bn_stats, get_param, get_train, new_param, set_param


class BatchNormLayer(AbstractLayer):
    type_name = "batchnorm"
    apply_relu = False

    def __init__(self, layers, momentum=0.9, eps=1e-5, renorm_max_r=1.0, renorm_max_d=0.0, renorm_max_it=10,
                 json_param={}):
        super().__init__(layer_index=len(layers))
        self.input = layers[-1].output
        self.input_shape = tuple(layers[-1].output_shape)
        self.enabled = json_param.get("enabled", True)
        self.momentum = json_param.get("momentum", momentum)
        self.renorm_max_r = json_param.get("renormMaxR", renorm_max_r)
        self.renorm_max_d = json_param.get("renormMaxD", renorm_max_d)
        self.renorm_max_it = json_param.get("renormMaxIt", renorm_max_it)
        self.eps = json_param.get("eps", eps)
        self.output_shape = self.input_shape
        self._init_params()

    def _init_params(self):
        c = self.input_shape[1]
        if self.enabled:
            self.omega = new_param(numpy.ones((c,)))
            self.beta = new_param(numpy.zeros((c,)))
            self.mean = new_param(numpy.zeros((c,)))
            self.stdinv = new_param(numpy.ones((c,)))
        self._fused = None          # (sum, sqsum) views into the model's per-step statistics buffer
        self._fused_ready = False
        self._saved = None
        self._bwd_sums = None       # (sum dz', sum dz' * xhat) views: backward statistics from a dgrad epilogue
        self._bwd_presummed = False

    @staticmethod
    def parse_desc(layers, name, tags, params):
        if name != "BN":
            return False
        layers.append(BatchNormLayer(layers, params.get(0, 0.9), params.get(1, 1e-5), params.get(2, 1),
                                     params.get(3, 0), params.get(4, 0)))
        return True

    def params(self):
        return [self.omega, self.beta, self.mean, self.stdinv] if self.enabled else []

    def updates(self, cost=None):
        return [(self.mean, "ema(batch mean)"), (self.stdinv, "ema(batch inverse std)")] if self.enabled else []

    def weights(self):
        return []

    def biases(self):
        return [self.omega, self.beta] if self.enabled else []

    def export_json(self):
        json = super().export_json()
        json.update({"momentum": self.momentum, "eps": self.eps})
        if self.enabled:
            json.update({"mean": get_param(self.mean), "std": get_param(self.stdinv),
                         "gamma": get_param(self.omega), "bias": get_param(self.beta)})
        if self.type_name == "batchnorm":
            json.update({"renormMaxR": self.renorm_max_r, "renormMaxD": self.renorm_max_d,
                         "renormMaxIt": self.renorm_max_it, "enabled": self.enabled})
        return json

    def import_json(self, json_param):
        if self.enabled:
            set_param(self.omega, json_param["gamma"])
            set_param(self.beta, json_param["bias"])
            set_param(self.mean, json_param["mean"])
            set_param(self.stdinv, json_param["std"])

    # ---------------------------------------------------------------------------------------------- execution
    def wants_fused_stats(self):
        ok = self.enabled and get_train() and fuse_bn_stats() and self._fused is not None
        if ok:
            self._fused_ready = True
        return ok

    def fused_stat_buffers(self):
        return self._fused

    def forward(self, x, residual=None, relu=None):
        """y = [relu](bn(x) [+ residual]); residual / relu let a ResNet block fold its add + ReLU into this pass"""
        self.input = x
        if not self.enabled:
            self.output = x
            return x
        relu = self.apply_relu if relu is None else relu
        c = self.input_shape[1]
        if get_train():
            mean = torch.empty((c,), dtype=torch.float32, device=x.device)
            invstd = torch.empty_like(mean)
            if self._fused_ready:
                # the conv epilogue left per-channel sums: statistics are finalised inside the apply launch
                y = ops.bn_apply_sums(x, self._fused[0], self._fused[1], self.eps, self.omega, self.beta, mean, invstd,
                                      self.mean, self.stdinv, self.momentum, residual=residual, relu=relu)
                self._fused_ready = False
            else:
                ops.bn_stats(x, self.eps, mean, invstd, self.mean, self.stdinv, self.momentum)
                y = ops.bn_apply(x, mean, invstd, self.omega, self.beta, residual=residual, relu=relu)
            # backward needs y only for the relu mask, and only when a residual was added (otherwise the mask is
            # recomputed from x: one tensor read less in both backward passes)
            self._saved = (x, y if (relu and residual is not None) else None, mean, invstd, relu)
        else:
            invstd = ops.bn_inference_invstd(self.stdinv, self.eps)
            y = ops.bn_apply(x, self.mean, invstd, self.omega, self.beta, residual=residual, relu=relu)
        self.output = y
        return y

    MAX_FUSED_BWD_CHANNELS = 512      # the dgrad epilogue keeps the per-channel constants in shared memory

    def bwd_fuse_args(self):
        """ops.BnBwdFuse for the dgrad that produces this layer's output gradient: its epilogue then applies the ReLU
        mask and accumulates the two per-channel sums, and backward() below skips its own reduction pass.  None when
        not applicable (parity mode, statistics buffers not linked, too many channels, no saved forward state)."""
        if not (self.enabled and self._saved is not None and self._bwd_sums is not None and fuse_bn_backward()):
            return None
        x, y, mean, invstd, relu = self._saved
        if x.shape[-1] > self.MAX_FUSED_BWD_CHANNELS:
            return None
        self._bwd_presummed = True
        return ops.BnBwdFuse(x, y, mean, invstd, self.omega, self.beta, relu, self._bwd_sums[0], self._bwd_sums[1])

    def backward(self, dy, want_dres=False):
        """returns dx, or (dx, dres) when want_dres (dres = masked gradient flowing into the fused residual input)"""
        if not self.enabled:
            return (dy, dy) if want_dres else dy
        x, y, mean, invstd, relu = self._saved
        self._saved = None
        if self._bwd_presummed:
            # dy arrived masked and the sums are complete (ops.conv2d_fprop(..., bn_bwd=...)): one pass left
            self._bwd_presummed = False
            dx = ops.bn_backward_sums(dy, x, mean, invstd, self.omega, self._bwd_sums[0], self._bwd_sums[1],
                                      self.omega.grad, self.beta.grad)
            return (dx, dy) if want_dres else dx
        dx, dres = ops.bn_backward(dy, y, x, mean, invstd, self.omega, relu, self.omega.grad, self.beta.grad,
                                   want_dres=want_dres, beta=self.beta)
        return (dx, dres) if want_dres else dx


class BatchNormReluLayer(BatchNormLayer):
    type_name = "batchnorm-relu"
    apply_relu = True

    def __init__(self, layers, momentum=0.9, eps=1e-5, json_param={}):
        AbstractLayer.__init__(self, layer_index=len(layers))
        self.input = layers[-1].output
        self.input_shape = tuple(layers[-1].output_shape)
        self.enabled = json_param.get("enabled", True)
        self.momentum = json_param.get("momentum", momentum)
        self.eps = json_param.get("eps", eps)
        self.output_shape = self.input_shape
        self._init_params()

    @staticmethod
    def parse_desc(layers, name, tags, params):
        if name != "BNA":
            return False
        layers.append(BatchNormReluLayer(layers, params.get(0, 0.9), params.get(1, 1e-5)))
        return True
