"""Skip connections 'SKIPSRC[idx]' / 'SKIP[idx]' (reference denet/layer/skip.py:9-116): SKIP adds the tensor marked
by the SKIPSRC with the same index, through a 1x1 projection convolution when the channel counts differ
(skip.py:78-86).  The add is folded into the projection conv's epilogue."""
from .. import ops
from . import AbstractLayer, InitialLayer
from .convolution import ConvLayer


class SkipSrcLayer(AbstractLayer):
    type_name = "skip-src"

    def __init__(self, layers, skip_index=0, split=False, json_param={}):
        super().__init__(layer_index=len(layers))
        self.skip_index = json_param.get("index", skip_index)
        self.split = json_param.get("split", split)   # 'X' tag: VRAM split point in the reference, identity here
        self.has_split = False
        self.input = layers[-1].output
        self.input_shape = self.output_shape = tuple(layers[-1].output_shape)
        self.skip = None
        self._skip_grad = None

    def export_json(self):
        j = super().export_json()
        j.update({"index": self.skip_index, "split": self.split})
        return j

    @staticmethod
    def parse_desc(layers, name, tags, params):
        if name != "SKIPSRC":
            return False
        layers.append(SkipSrcLayer(layers, params.get(0, 0), "X" in tags))
        return True

    def forward(self, x):
        self.input = self.output = self.skip = x
        self._skip_grad = None
        return x

    def add_skip_grad(self, g):
        self._skip_grad = g if self._skip_grad is None else ops.add(self._skip_grad, g)

    def backward(self, dy):
        g, self._skip_grad = self._skip_grad, None
        self.skip = None
        if g is None:
            return dy
        return g if dy is None else ops.add(dy, g)


class SkipLayer(AbstractLayer):
    type_name = "skip"

    def __init__(self, layers, skip_index=0, combine_mode="proj-add", json_param={}):
        super().__init__(layer_index=len(layers))
        self.combine_mode = json_param.get("combineMode", combine_mode)
        self.skip_index = json_param.get("index", skip_index)
        found = None
        for layer in layers:
            if layer.type_name == "skip-src" and layer.skip_index == self.skip_index:
                found = layer
                break
        assert found is not None
        object.__setattr__(self, "skip_layer", found)   # a reference only: not a sub-module of this layer
        self.x_shape = tuple(layers[-1].output_shape)
        self.y_shape = tuple(self.skip_layer.output_shape)
        if self.combine_mode == "proj-add":
            self.output_shape = self.x_shape
            if self.y_shape[1] != self.x_shape[1]:
                self.layers.append(InitialLayer(None, self.y_shape))
                self.layers.append(ConvLayer(self.layers, filter_shape=(self.x_shape[1], self.y_shape[1], 1, 1)))
        elif self.combine_mode == "concat":
            raise Exception("skip combine mode 'concat' is not on the B200 hot path")
        else:
            raise Exception("Unknown combine mode: %s" % self.combine_mode)

    def export_json(self):
        j = super().export_json()
        j.update({"index": self.skip_index, "combineMode": self.combine_mode})
        return j

    @staticmethod
    def parse_desc(layers, name, tags, params):
        if name != "SKIP":
            return False
        layers.append(SkipLayer(layers, params.get(0, 0)))
        return True

    def forward(self, x):
        self.input = x
        y = self.skip_layer.skip
        if len(self.layers) > 0:
            self.output = self.layers[1].forward(y, residual=x)   # x + proj(y) in one epilogue
        else:
            self.output = ops.add(x, y)
        return self.output

    def backward(self, dy):
        if len(self.layers) > 0:
            self.skip_layer.add_skip_grad(self.layers[1].backward(dy))
        else:
            self.skip_layer.add_skip_grad(dy)
        return dy
