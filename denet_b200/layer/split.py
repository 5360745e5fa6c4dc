"""'SPLIT' layer (reference denet/layer/split.py:7-46).  In the reference it cuts the Theano graph into separately
compiled forward/backward functions to fit 12 GB GPUs; with 180 GB of HBM the semantics are the identity."""
from . import AbstractLayer


class SplitLayer(AbstractLayer):
    type_name = "split"

    def __init__(self, layers, json_param={}):
        super().__init__(layer_index=len(layers))
        self.enabled = json_param.get("enabled", True)
        self.has_split = False
        self.input = layers[-1].output
        self.output_shape = self.input_shape = tuple(layers[-1].output_shape)

    def export_json(self):
        json = super().export_json()
        json.update({"enabled": self.enabled})
        return json

    @staticmethod
    def parse_desc(layers, name, tags, params):
        if name != "SPLIT":
            return False
        layers.append(SplitLayer(layers))
        return True
