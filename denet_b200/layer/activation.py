"""Activation layer 'A' (reference denet/layer/activation.py:9-56).  The hot path only uses ReLU
(tensor.nnet.relu == 0.5*(x+|x|), activation.py:32-34) and 'none'."""
from .. import ops
from . import AbstractLayer


class ActivationLayer(AbstractLayer):
    type_name = "activation"
    SUPPORTED = ("relu", "relu-safe", "none")

    def __init__(self, layers, activation="relu", json_param={}):
        super().__init__(layer_index=len(layers))
        self.input = layers[-1].output
        self.input_shape = tuple(layers[-1].output_shape)
        self.activation = json_param.get("activation", activation)
        if self.activation not in self.SUPPORTED:
            raise Exception("Unknown / unsupported activation type on the B200 hot path:", self.activation)
        self.output_shape = self.input_shape

    @staticmethod
    def parse_desc(layers, name, tags, params):
        if name != "A":
            return False
        layers.append(ActivationLayer(layers, params["activation"]))
        return True

    def export_json(self):
        json = super().export_json()
        json.update({"activation": self.activation})
        return json

    def forward(self, x):
        self.input = x
        self.output = x if self.activation == "none" else ops.relu_fwd(x)
        return self.output

    def backward(self, dy):
        if self.activation == "none":
            return dy
        return ops.relu_bwd(dy, self.output)
