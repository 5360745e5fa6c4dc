"""denet_b200: B200-native implementation of the DeNet training hot path behind the reference's Layer API.

Package map (mirrors the reference's `denet` package for the hot path only):
  denet_b200.layer   - Layer plugin API: parse_desc() / type_name / export_json / import_json / get_target / cost
  denet_b200.model   - ModelCNN (build, build_train_func, train_step, train_epoch) and the model-train driver
  denet_b200.multi   - data-parallel replicas with NCCL gradient all-reduce (replaces denet/multi)
  denet_b200.common  - find_layers, IoU helpers, gz-JSON checkpoint format
  denet_b200.ops/lib - tensor-level wrappers / ctypes binding of the C-ABI kernels (csrc/, include/denet_b200.h)
"""
__version__ = "0.1.0"
