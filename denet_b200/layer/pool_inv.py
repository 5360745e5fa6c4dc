"""Pool-inv layer 'PI[n]': nearest-neighbour n x n upsampling (reference denet/layer/pool_inv.py:10-41; kernels
k_pool_inv_WxH / k_pool_inv_grad_WxH, pool_inv_op.py:38-63, 144-169)."""
from .. import ops
from . import AbstractLayer


class PoolInvLayer(AbstractLayer):
    type_name = "pool-inv"

    def __init__(self, layers, size=(2, 2), json_param={}):
        super().__init__(layer_index=len(layers))
        self.input = layers[-1].output
        self.input_shape = tuple(layers[-1].output_shape)
        self.size = tuple(json_param.get("size", size))   # (size_w, size_h)
        self.output_shape = (self.input_shape[0], self.input_shape[1], self.size[1] * self.input_shape[2],
                             self.size[0] * self.input_shape[3])

    @staticmethod
    def parse_desc(layers, name, tags, params):
        if name != "PI":
            return False
        layers.append(PoolInvLayer(layers, (params.get(0), params.get(0))))
        return True

    def export_json(self):
        json = super().export_json()
        json.update({"size": self.size})
        return json

    def forward(self, x):
        self.input = x
        self.output = ops.pool_inv_fwd(x, self.size)
        return self.output

    def backward(self, dy):
        return ops.pool_inv_bwd(dy, self.size)
