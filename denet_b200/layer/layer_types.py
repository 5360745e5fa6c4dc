"""Layer registry (reference denet/layer/layer_types.py:17-25).  ModelCNN.build_layer offers every model-desc token
to each class's parse_desc in this order.  Dropout, Border, CropMirror and Deconv are not on the hot path
 and are not registered."""
from . import IdentityLayer, InitialLayer  # noqa: F401
from .activation import ActivationLayer
from .batch_norm import BatchNormLayer, BatchNormReluLayer
from .convolution import ConvLayer
from .denet_corner import DeNetCornerLayer
from .denet_detect import DeNetDetectLayer
from .denet_sparse import DeNetSparseLayer
from .pool import PoolLayer
from .pool_inv import PoolInvLayer
from .regression import RegressionLayer
from .resnet import ResnetLayer
from .skip import SkipLayer, SkipSrcLayer
from .split import SplitLayer

layer_types = [IdentityLayer, ConvLayer, PoolLayer, PoolInvLayer, RegressionLayer, ActivationLayer, BatchNormLayer,
               BatchNormReluLayer, ResnetLayer, SplitLayer, SkipLayer, SkipSrcLayer]

# DeNet detection layers
layer_types += [DeNetCornerLayer, DeNetSparseLayer, DeNetDetectLayer]
