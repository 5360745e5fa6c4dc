"""ResNet block 'RSN[f,k,s,bottleneck]' / 'nRSN[n,f,k,s,bottleneck]' (reference denet/layer/resnet.py:13-169).

Sub-layer list, JSON schema and wiring follow the reference: conv -> BN(+ReLU) -> conv -> BN (three convs for the
bottleneck form), 1x1 projection shortcut (+BN in the 'original' design) when the shape changes, out = relu(x + y)
('original', tag O) or x + y ('pre-activation').  Execution differences: the residual add + ReLU is folded into the
last batch-norm's apply pass, and in backward the shortcut gradient is folded into the first conv's dgrad epilogue.
"""
from .. import ops
from . import AbstractLayer, InitialLayer
from .activation import ActivationLayer
from .batch_norm import BatchNormLayer, BatchNormReluLayer
from .convolution import ConvLayer


class ResnetLayer(AbstractLayer):
    type_name = "resnet"

    def __init__(self, layers, filter_shape=None, stride=(1, 1), bottleneck=0, activation="relu", version="original",
                 json_param={}):
        super().__init__(layer_index=len(layers))
        self.input = layers[-1].output
        self.input_shape = tuple(layers[-1].output_shape)
        self.filter_shape = tuple(json_param.get("shape", filter_shape))
        self.stride = tuple(json_param.get("stride", stride))
        self.bottleneck = json_param.get("bottleneck", bottleneck)
        self.version = json_param.get("version", version)
        self.activation = json_param.get("activation", activation)
        self.bn_json_param = json_param.get("bnParam", {"enabled": json_param.get("enableBatchNorm", True)})
        if self.activation != "relu":
            raise Exception("resnet blocks on the B200 hot path use relu, got " + str(self.activation))

        fs = self.filter_shape
        if self.bottleneck > 0:
            self.size = (fs[2], fs[3])
            shape0 = (self.bottleneck, fs[1], 1, 1)
            shape1 = (self.bottleneck, self.bottleneck, fs[2], fs[3])
            shape2 = (fs[0], self.bottleneck, 1, 1)
        else:
            self.size = (fs[2] * 2 - 1, fs[3] * 2 - 1)
            shape0 = fs
            shape1 = (fs[0], fs[0], fs[2], fs[3])
            shape2 = None

        preact = "pre-activation" in self.version
        bnrelu = "bnrelu" in self.version
        L = self.layers
        L.append(InitialLayer(None, self.input_shape))

        def add_bn_act():
            if bnrelu:
                L.append(BatchNormReluLayer(L, json_param=self.bn_json_param))
            else:
                L.append(BatchNormLayer(L, json_param=self.bn_json_param))
                L.append(ActivationLayer(L, self.activation))

        if preact:
            add_bn_act()
        self._tap = len(L)   # the pre-activation shortcut taps layers[0:2] (resnet.py:91-92)
        L.append(ConvLayer(L, filter_shape=shape0, filter_stride=self.stride, border_mode="half", use_bias=False))
        add_bn_act()
        L.append(ConvLayer(L, filter_shape=shape1, border_mode="half", use_bias=False))
        if self.bottleneck > 0:
            add_bn_act()
            L.append(ConvLayer(L, filter_shape=shape2, border_mode="half", use_bias=False))
        if not preact:
            L.append(BatchNormLayer(L, json_param=self.bn_json_param))
        self._main_end = len(L)
        y_shape = tuple(L[-1].output_shape)

        self._has_proj = self.input_shape != y_shape
        if self._has_proj:
            input_layers = list(L[0:2]) if preact else [InitialLayer(None, self.input_shape)]
            L.append(ConvLayer(input_layers, filter_shape=(y_shape[1], self.input_shape[1], 1, 1),
                               filter_stride=self.stride, use_bias=False, border_mode="half"))
            if "original" in self.version:
                L.append(BatchNormLayer(L, json_param=self.bn_json_param))
        self.output_shape = y_shape
        self._preact = preact

    @staticmethod
    def parse_desc(layers, name, tags, params):
        if name == "RSN":
            version = "original" if "O" in tags else "pre-activation"
            filter_shape = (params.get(0), layers[-1].output_shape[1], params.get(1), params.get(1))
            filter_stride = (params.get(2, 1), params.get(2, 1))
            layers.append(ResnetLayer(layers, filter_shape, filter_stride, params.get(3, 0), params["activation"],
                                      version))
            return True
        if name == "nRSN":
            version = "original" if "O" in tags else "pre-activation"
            bottleneck = params.get(4, 0)
            for i in range(params.get(0)):
                filter_shape = (params.get(1), layers[-1].output_shape[1], params.get(2), params.get(2))
                filter_stride = (params.get(3, 1), params.get(3, 1)) if i == 0 else (1, 1)
                layers.append(ResnetLayer(layers, filter_shape, filter_stride, bottleneck, params["activation"],
                                          version))
            return True
        return False

    def updates(self, cost=None):
        return sum([layer.updates(cost) for layer in self.layers], [])

    def weights(self):
        return sum([layer.weights() for layer in self.layers], [])

    def biases(self):
        return sum([layer.biases() for layer in self.layers], [])

    def import_json(self, json_param):
        n = 0
        for json_layer in json_param["layers"]:
            if json_layer["type"] == "identity":   # introduced by old versions, ignored (resnet.py:147-149)
                continue
            assert json_layer["type"] == self.layers[n].type_name, (json_layer["type"], self.layers[n].type_name)
            self.layers[n].import_json(json_layer)
            n += 1

    def export_json(self):
        json = super().export_json()
        json.update({"shape": self.filter_shape, "stride": self.stride, "bottleneck": self.bottleneck,
                     "bnParam": self.bn_json_param, "activation": self.activation, "version": self.version})
        return json

    # ---------------------------------------------------------------------------------------------- execution
    def forward(self, x):
        self.input = x
        L = self.layers
        if self._preact:
            # x -> BN/ReLU -> main path; shortcut from the raw input or (projection) from layers[1]'s output
            h = x
            tap = x
            for i in range(1, self._main_end):
                h = L[i].forward(h)
                if i == 1:
                    tap = h
            s = L[self._main_end].forward(tap) if self._has_proj else x
            self.output = ops.add(s, h)
            return self.output
        h = x
        for i in range(1, self._main_end - 1):
            h = L[i].forward(h)
        if self._has_proj:
            s = x
            for i in range(self._main_end, len(L)):
                s = L[i].forward(s)
        else:
            s = x
        # relu(shortcut + bn(y)) in the last batch-norm's apply pass (resnet.py:113)
        last_bn = L[self._main_end - 1]
        if last_bn.enabled:
            self.output = last_bn.forward(h, residual=s, relu=True)
        else:
            self.output = ops.add(s, h, relu=True)
        return self.output

    @property
    def accepts_bn_next(self):
        return not self._preact

    def fusable_last_bn(self):
        """the batch-norm layer whose backward consumes the gradient wrt this block's output (original version only)"""
        if self._preact:
            return None
        bn = self.layers[self._main_end - 1]
        return bn if bn.enabled else None

    def backward(self, dy, bn_next=None):
        """bn_next: batch-norm layer (of the preceding block / layer) that consumes the returned gradient, see
        ConvLayer.backward"""
        L = self.layers
        if self._preact:
            ds = dy
            dh = dy
            d_tap = None
            if self._has_proj:
                d_tap = L[self._main_end].backward(ds)
                ds = None
            for i in range(self._main_end - 1, 1, -1):
                dh = L[i].backward(dh)
            if d_tap is not None:
                dh = ops.add(dh, d_tap)
            dh = L[1].backward(dh)
            return dh if ds is None else ops.add(dh, ds)
        last_bn = L[self._main_end - 1]
        if last_bn.enabled:
            dh, ds = last_bn.backward(dy, want_dres=True)
        else:
            dh = ds = ops.relu_bwd(dy, self.output)
        if self._has_proj:
            for i in range(len(L) - 1, self._main_end - 1, -1):
                ds = L[i].backward(ds)
        for i in range(self._main_end - 2, 1, -1):
            prev = L[i - 1]
            if isinstance(L[i], ConvLayer) and isinstance(prev, BatchNormLayer) and prev.enabled and i - 1 > 1:
                dh = L[i].backward(dh, bn_next=prev)      # dgrad epilogue starts the backward of the BN before it
            else:
                dh = L[i].backward(dh)
        # first conv's dgrad epilogue adds the shortcut gradient (and starts the backward of the preceding block's BN)
        return L[1].backward(dh, add_to=ds, bn_next=bn_next)
