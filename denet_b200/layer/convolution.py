"""Convolution layer: same constructor / parse_desc / JSON contract as the reference ConvLayer
(denet/layer/convolution.py:10-136); fprop / dgrad / wgrad run as tcgen05 implicit GEMMs (csrc/conv_tc.cu).

Theano's conv2d flips the filter (true convolution, convolution.py:83).  The fp32 master weight `omega` keeps the
reference layout (Cout, Cin, R, S) of that true convolution, so checkpoints are interchangeable; the flip is folded
into the bf16 GEMM operands that denet_conv_weight_prep derives from it after every solver step.
"""
import math

import numpy
import torch

from .. import lib, ops
from . import (AbstractLayer, act_dtype, get_param, get_precision, new_param, param_version, set_param,
               wgrad_pending, wgrad_side)


def conv_output_hw(in_hw, size, stride, border_mode):
    """output extent exactly as convolution.py:55-74 computes it, plus the zero padding that implements the border"""
    out, pad = [], []
    for i in range(2):
        n, k, s = in_hw[i], size[i], stride[i]
        if border_mode == "valid":
            p, o = 0, math.ceil((n - k + 1) / s)
        elif border_mode == "full":
            p, o = k - 1, math.ceil((n + k - 1) / s)
        elif border_mode == "half":
            p = k // 2
            o = math.ceil((n + 2 * p - k + 1) / s)
        elif border_mode == "same":
            assert tuple(stride) == (1, 1)
            # full convolution cropped at (k-1)//2  ==  correlation with leading pad (k-1) - (k-1)//2
            p, o = (k - 1) - (k - 1) // 2, n
        elif isinstance(border_mode, (int, bool)):
            p = int(border_mode)
            o = math.ceil((n + 2 * p - k + 1) / s)
        elif isinstance(border_mode, (tuple, list)):
            p = int(border_mode[i])
            o = math.ceil((n + 2 * p - k + 1) / s)
        else:
            raise Exception("Unknown border mode: " + str(border_mode))
        out.append(int(o))
        pad.append(p)
    return tuple(out), tuple(pad)


class ConvLayer(AbstractLayer):
    type_name = "conv"
    IM2COL_MAX_CIN = 16   # below this the 64-channel K chunks of the TMA path would be mostly zero padding

    def __init__(self, layers, filter_shape=None, filter_stride=(1, 1), use_bias=False, border_mode="half",
                 wb="he-backward", json_param={}):
        super().__init__(layer_index=len(layers))
        self.input = layers[-1].output
        self.input_shape = tuple(layers[-1].output_shape)
        self.is_first = getattr(layers[-1], "is_model_input", False)   # no data gradient needed for the images

        self.border_mode = json_param.get("border", border_mode)
        if isinstance(self.border_mode, list):
            self.border_mode = tuple(self.border_mode)
        self.filter_shape = tuple(int(v) for v in json_param.get("shape", filter_shape))
        self.stride = tuple(int(v) for v in json_param.get("stride", filter_stride))
        self.use_bias = bool(json_param.get("useBias", use_bias))
        self.size = (self.filter_shape[2], self.filter_shape[3])
        self.enabled = json_param.get("enabled", True)

        # weight initialisation, same numpy.random call sequence as convolution.py:28-46
        fs = self.filter_shape
        if type(wb) is float or type(wb) is int:
            self.w_bound = float(wb)
        elif "he-forward" in wb:
            self.w_bound = math.sqrt(2.0 / (fs[2] * fs[3] * fs[1]))
        elif "he-backward" in wb:
            self.w_bound = math.sqrt(2.0 / (fs[2] * fs[3] * fs[0]))
        elif "xavier-forward" in wb:
            self.w_bound = math.sqrt(1.0 / (fs[2] * fs[3] * fs[1]))
        elif "xavier-backward" in wb:
            self.w_bound = math.sqrt(1.0 / (fs[2] * fs[3] * fs[0]))
        else:
            raise Exception("Unknown weight initialisation: " + str(wb))
        if self.w_bound > 0:
            if type(wb) is str and "uniform" in wb:
                w = numpy.random.uniform(-self.w_bound, self.w_bound, size=fs)
            else:
                w = numpy.random.normal(0.0, self.w_bound, size=fs)
        else:
            w = numpy.zeros(shape=fs)
        self.omega = new_param(w)
        if self.use_bias:
            self.beta = new_param(numpy.zeros((fs[0],)))

        (oh, ow), self.pad = conv_output_hw(self.input_shape[2:], self.size, self.stride, self.border_mode)
        self.output_shape = (self.input_shape[0], fs[0], oh, ow)
        self.output = None
        self.out_fp32 = False          # logits layers (DNC / DND / R) keep an fp32 output in bf16 mode
        self.stat_consumer = None      # BatchNorm layer fed by this conv's epilogue statistics (set by link pass)
        self._wver = -1
        self._wop_f = self._wop_d = self._w2 = None
        self._wop_dc = None            # strided multi-tap conv: dgrad operands per parity class {(a, b): ConvOperand}

    @staticmethod
    def parse_desc(layers, name, tags, params):
        if name != "C":
            return False
        use_bias = bool("B" in tags)
        cin = layers[-1].output_shape[1]
        if bool("X" in tags):
            filter_shape = (params.get(0), cin, params.get(1), params.get(2))
            filter_stride = (params.get(3, 1), params.get(4, 1))
        else:
            filter_shape = (params.get(0), cin, params.get(1, 1), params.get(1, 1))
            filter_stride = (params.get(2, 1), params.get(2, 1))
        layers.append(ConvLayer(layers, filter_shape, filter_stride, use_bias, params["borderMode"], params["wb"]))
        return True

    def fprop_flops(self):
        """algorithmic FLOPs of one forward pass: 2 * N * Cout * Ho * Wo * Cin * R * S"""
        n, co, oh, ow = self.output_shape
        return 2.0 * n * co * oh * ow * self.filter_shape[1] * self.size[0] * self.size[1]

    def weights(self):
        return super().weights() + ([self.omega] if self.enabled else [])

    def biases(self):
        return super().biases() + ([self.beta] if self.use_bias and self.enabled else [])

    def import_json(self, json_param):
        super().import_json(json_param)
        if self.use_bias:
            set_param(self.beta, json_param["bias"])
        set_param(self.omega, json_param["weight"])

    def export_json(self):
        json = super().export_json()
        json.update({"shape": self.filter_shape, "stride": self.stride, "border": self.border_mode,
                     "enabled": self.enabled, "useBias": self.use_bias,
                     "bias": get_param(self.beta) if self.use_bias else None, "weight": get_param(self.omega)})
        return json

    # ---------------------------------------------------------------------------------------------- execution
    @property
    def rowfold(self):
        """(Cp, Hp, Wp) when this is the image stem and runs as a row-folded conv over a zero-padded NHWC-Cp image
        (no im2col matrix, see csrc/conv_tc.cu), else None"""
        if not (self.is_first and self.enabled and self.size != (1, 1)) or not isinstance(self.pad[0], int):
            return None
        return ops.rowfold_geometry(self.input_shape[2:], self.filter_shape[1], self.size, self.stride, self.pad,
                                    self.output_shape[2:])

    @property
    def use_im2col(self):
        return self.filter_shape[1] <= self.IM2COL_MAX_CIN and self.size != (1, 1) and self.rowfold is None

    def dgrad_classes(self):
        """parity classes of the strided data gradient (ops.dgrad_parity_classes), or None when this layer's dgrad does
        not use them (stride 1, 1x1 filters, the image stem, im2col variant, a class without taps)"""
        if getattr(self, "_dclasses", 0) == 0:
            cls = None
            if self.enabled and not self.is_first and self.stride != (1, 1) and self.size != (1, 1) and \
                    not self.use_im2col and isinstance(self.pad[0], int) and max(self.stride) <= 15 and \
                    max(self.size) <= 15:
                cls = ops.dgrad_parity_classes(self.input_shape[2:], self.size, self.stride, self.pad)
                if cls is not None:
                    # each class is its own launch of 8 x 16-pixel tiles: worth it only while a class fills the GPU
                    # (measured: 64 tiles per class lose to one launch on the zero-dilated gradient)
                    hc, wc = cls[0][8], cls[0][9]
                    if self.input_shape[0] * ((hc + 15) // 16) * ((wc + 7) // 8) < 148:
                        cls = None
            object.__setattr__(self, "_dclasses", cls)
        return self._dclasses

    def _class_operands(self, split):
        """allocate (once) the per-class dgrad operands; returns the prep records [(w, operand, 3, code)]"""
        cout, cin, R, S = self.filter_shape
        if self._wop_dc is None or (next(iter(self._wop_dc.values())).lo is not None) != split:
            self._wop_dc = {}
            for c in self.dgrad_classes():
                a, b, r0, s0, rc, sc = c[:6]
                hi = torch.empty((cin, rc * sc, (cout + 63) // 64 * 64), dtype=torch.bfloat16, device=self.omega.device)
                self._wop_dc[(a, b)] = ops.ConvOperand(hi, torch.empty_like(hi) if split else None, cin, cout, rc, sc)
        return [(self.omega, self._wop_dc[(c[0], c[1])], 3, ops.parity_class_code(c, self.stride))
                for c in self.dgrad_classes()]

    def _operands(self):
        """bf16 GEMM operands derived from omega; refreshed when the parameters changed"""
        if self._wver != param_version():
            split = get_precision() == "fp32"
            if self._wop_f is not None and (self._wop_f.lo is not None) != split:
                self._wop_f = self._wop_d = None
            if self.dgrad_classes() is not None:
                ops.conv_weight_prep_records(self._class_operands(split))
            w = self.omega
            if self.rowfold is not None:
                self._wop_f = ops.conv_weight_prep_rowfold(w, self.rowfold[0], split, self._wop_f)
                self._wver = param_version()
                return self._wop_f, None
            if self.use_im2col:
                self._w2 = ops.weight_to_im2col(w, self._w2)
                w = self._w2
            self._wop_f = ops.conv_weight_prep(w, 0, split, self._wop_f)
            if not self.is_first:
                self._wop_d = ops.conv_weight_prep(w, 1, split, self._wop_d)
            self._wver = param_version()
        return self._wop_f, self._wop_d

    def prep_entries(self):
        """(w, operand, mode, Cp) records for ModelCNN's one-launch operand preparation, or None when this layer
        prepares its own operands (im2col variant); allocates the operand buffers on first use"""
        if not self.enabled or self.use_im2col:
            return None
        split = get_precision() == "fp32"
        if self._wop_f is not None and (self._wop_f.lo is not None) != split:
            self._wop_f = self._wop_d = None
        cout, cin, R, S = self.filter_shape
        dev = self.omega.device

        def alloc(shape, rows, kin):
            hi = torch.empty(shape, dtype=torch.bfloat16, device=dev)
            return ops.ConvOperand(hi, torch.empty_like(hi) if split else None, rows, kin, R, S)
        if self.rowfold is not None:
            if self._wop_f is None:
                self._wop_f = alloc((cout, R, 64), cout, cin)
            return [(self.omega, self._wop_f, 2, self.rowfold[0])]
        if self._wop_f is None:
            self._wop_f = alloc((cout, R * S, (cin + 63) // 64 * 64), cout, cin)
        out = [(self.omega, self._wop_f, 0, 0)]
        if not self.is_first:
            if self.dgrad_classes() is not None:
                out += self._class_operands(split)       # strided: one small operand per parity class
            else:
                if self._wop_d is None:
                    self._wop_d = alloc((cin, R * S, (cout + 63) // 64 * 64), cin, cout)
                out.append((self.omega, self._wop_d, 1, 0))
        return out

    def mark_operands_current(self):
        self._wver = param_version()

    def _as_operand(self, t):
        """activation / gradient tensor -> MMA operand of the current precision mode"""
        if get_precision() == "bf16" and t.dtype == torch.float32:
            t = ops.convert(t, torch.bfloat16)
        return ops.act_operand(t)

    def forward(self, x, residual=None, relu=False):
        self.input = x
        if not self.enabled:
            self.output = x
            return x
        wop_f, _ = self._operands()
        lib.set_tag(("fprop", self))
        n, ci, h, w = self.input_shape
        _, co, oh, ow = self.output_shape
        out_dtype = torch.float32 if self.out_fp32 else act_dtype()
        stats = None
        if self.stat_consumer is not None and self.stat_consumer.wants_fused_stats():
            stats = self.stat_consumer.fused_stat_buffers()
        bias = self.beta if self.use_bias else None
        if self.rowfold is not None:
            assert isinstance(x, ops.PaddedImage) and residual is None, "the stem conv reads the padded model input"
            self._xop = x
            y = ops.conv2d_rowfold_fprop(x, wop_f, self.stride, (oh, ow), out_dtype, bias=bias, relu=relu, stats=stats)
        elif self.use_im2col:
            col = ops.im2col(x, self.size[0], self.size[1], self.stride, self.pad, (oh, ow))
            self._xop = self._as_operand(col)
            y = ops.conv2d_fprop(self._xop, wop_f, (0, 0), (oh, ow), out_dtype, bias=bias, residual=residual,
                                 relu=relu, stats=stats)
        else:
            self._xop = self._as_operand(x)
            y = ops.conv2d_fprop(self._xop, wop_f, self.pad, (oh, ow), out_dtype, stride=self.stride, bias=bias,
                                 residual=residual, relu=relu, stats=stats)
        self.output = y
        return y

    accepts_bn_next = True

    def backward(self, dy, add_to=None, bn_next=None):
        """dy: gradient wrt the conv output (before any fused residual / relu).  Writes omega.grad (and beta.grad),
        returns the gradient wrt the input (+ add_to, fused into the dgrad epilogue where possible).
        bn_next: the batch-norm layer that produced this conv's input and will consume the returned gradient next; where
        the dgrad runs as one stride-1 correlation its epilogue takes over the first pass of that layer's backward."""
        if not self.enabled:
            return dy if add_to is None else ops.add(dy, add_to)
        _, wop_d = self._operands()
        lib.set_tag(("bwd", self))
        n, ci, h, w = self.input_shape
        R, S = self.size
        dyop = self._as_operand(dy)
        if self.use_bias and self.beta.grad is not None:
            ops.colsum(dy, self.beta.grad)
        gdt = act_dtype()
        dx = None
        pending = wgrad_pending()
        defer = (pending, self) if pending is not None else None
        if self.rowfold is not None:
            ops.conv2d_rowfold_wgrad(dyop, self._xop, R, S, self.stride, self.omega.grad, defer=defer)
        elif self.use_im2col:
            dw2 = ops.conv2d_wgrad(dyop, self._xop, 1, 1, (0, 0))
            ops.weight_grad_from_im2col(dw2, self.omega.grad)
            if not self.is_first:
                dcol = ops.conv2d_fprop(dyop, wop_d, (0, 0), dy.shape[1:3], gdt)
                dx = ops.col2im(dcol, (n, h, w, ci), R, S, self.stride, self.pad)
                if add_to is not None:
                    dx = ops.add(dx, add_to, out=dx)
        else:
            side = wgrad_side() if defer is not None else None
            if side is not None:
                # the filter gradient only feeds the (deferred) reduction at the end of the pass: run it on the side
                # stream, next to the data-gradient / batch-norm chain of the layers below
                stream, keep = side
                ready = torch.cuda.Event()
                ready.record(torch.cuda.current_stream())
                stream.wait_event(ready)
                with ops.on_stream(stream):
                    ops.conv2d_wgrad(dyop, self._xop, R, S, self.pad, self.stride, dw=self.omega.grad, defer=defer)
                keep.append((dyop, self._xop, dy))       # both operands must outlive the side stream's kernel
            else:
                ops.conv2d_wgrad(dyop, self._xop, R, S, self.pad, self.stride, dw=self.omega.grad, defer=defer)
            if not self.is_first:
                if self.stride == (1, 1):
                    dx = ops.conv2d_fprop(dyop, wop_d, (R - 1 - self.pad[0], S - 1 - self.pad[1]), (h, w), gdt,
                                          residual=add_to, bn_bwd=bn_next.bwd_fuse_args() if bn_next is not None else None)
                elif (R, S) == (1, 1) and self.pad == (0, 0):
                    # a 1x1 convolution commutes with the zero insertion: GEMM at the small resolution, then scatter
                    dxc = ops.conv2d_fprop(dyop, wop_d, (0, 0), dy.shape[1:3], gdt)
                    dx = ops.dilate(dxc, self.stride, (h, w), add=add_to)
                elif self.dgrad_classes() is not None and self._wop_dc is not None:
                    # strided conv: one stride-1 correlation per parity class of the input pixel, on the undilated dy,
                    # each scattering its pixels (and adding add_to there) into dx
                    dx = ops.alloc_nhwc(n, h, w, ci, gdt, dy.device)
                    for (a, b, r0, s0, rc, sc, ph, pw, hc, wc) in self.dgrad_classes():
                        ops.conv2d_fprop(dyop, self._wop_dc[(a, b)], (ph, pw), (hc, wc), gdt, residual=add_to, out=dx,
                                         scatter=(h, w, self.stride[0], self.stride[1], a, b))
                else:
                    oh, ow = dy.shape[1:3]
                    hd, wd = (oh - 1) * self.stride[0] + 1, (ow - 1) * self.stride[1] + 1
                    dyd = ops.ActOperand(ops.dilate(dyop.hi, self.stride, (hd, wd)),
                                         None if dyop.lo is None else ops.dilate(dyop.lo, self.stride, (hd, wd)))
                    dx = ops.conv2d_fprop(dyd, wop_d, (R - 1 - self.pad[0], S - 1 - self.pad[1]), (h, w), gdt,
                                          residual=add_to, bn_bwd=bn_next.bwd_fuse_args() if bn_next is not None else None)
        self._xop = None
        return dx
