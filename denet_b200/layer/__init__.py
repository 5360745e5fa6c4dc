"""Layer plugin API of the reference (denet/layer/__init__.py:64-143), re-hosted on PyTorch nn.Modules.

What stays identical for callers (ModelCNN / bin/model-train): the `layer_types` registry, per-class static
`parse_desc(layers, name, tags, params) -> bool`, `type_name`, `layer_index`, `layers` (sub-layers), `has_split`,
`input_shape` / `output_shape` as (B, C, H, W) tuples, `weights() / biases() / params() / updates(cost)`,
`cost(yt_index, yt_value)`, `get_target(model, data_x, metas)`, `export_json() / import_json()`.

What changes underneath: the reference extends a symbolic Theano graph in every constructor and lets Theano
differentiate it; here every layer is an nn.Module with an explicit `forward(x)` and `backward(dy)` that enqueue
hand-written sm_100a kernels (denet_b200.ops -> C-ABI) on the current CUDA stream.  Activations travel as NHWC device
tensors (bf16 in throughput mode, fp32 in parity mode); `output` holds the tensor of the last forward pass instead of
a symbolic variable.  There is no autograd graph and no CPU fallback.
"""
import itertools

import numpy
import torch
from torch import nn

# ---------------------------------------------------------------------------------------------------------------
# global switches (the reference keeps train flag / epoch / iteration as module-level Theano symbols,
# denet/layer/__init__.py:5-28)
_state = {"train": False, "epoch": 0, "iteration": 0, "precision": "bf16", "device": "cuda", "param_version": 0,
          "fuse_bn_stats": True, "device_targets": True, "gt": None, "wgrad_pending": None,
          "fuse_bn_bwd": False, "wgrad_side": None}


def get_train():
    return _state["train"]


def set_train(v):
    _state["train"] = bool(v)


def wgrad_pending():
    """list collecting the filter gradients whose split-K reduction is deferred to one multi-tensor launch
    (ModelCNN.backward owns it), or None: every conv layer reduces its own gradient right away"""
    return _state["wgrad_pending"]


def set_wgrad_pending(v):
    _state["wgrad_pending"] = v


def wgrad_side():
    """(side stream, keep-alive list) while ModelCNN.backward runs the filter gradients on a second stream, else None"""
    return _state["wgrad_side"]


def set_wgrad_side(v):
    _state["wgrad_side"] = v


def fuse_bn_backward():
    """let dgrad epilogues take over the reduction pass of the batch-norm backward (throughput mode only)"""
    return _state["fuse_bn_bwd"] and fuse_bn_stats()


def set_fuse_bn_backward(v):
    _state["fuse_bn_bwd"] = bool(v)


def device_targets():
    """build the corner / detection targets on the device from the ground-truth boxes (csrc/targets.cu) instead of in
    numpy on the host; active when the model has uploaded the boxes of the current batch (set_ground_truth)"""
    return _state["device_targets"] and _state["gt"] is not None


def set_device_targets(v):
    _state["device_targets"] = bool(v)


def get_ground_truth():
    return _state["gt"]


def set_ground_truth(gt):
    """(gt_bbox (B,G,4) f64, gt_class (B,G) i32, gt_count (B) i32) device tensors of the current batch, or None"""
    _state["gt"] = gt


def get_epoch():
    return _state["epoch"]


def set_epoch(v):
    _state["epoch"] = int(v)


def get_iteration():
    return _state["iteration"]


def set_iteration(v):
    _state["iteration"] = int(v)


def get_precision():
    """'bf16' (throughput: bf16 activations, bf16 x bf16 -> fp32 MMA) or 'fp32' (parity: fp32 activations,
    error-compensated bf16x3 MMA)"""
    return _state["precision"]


def set_precision(p):
    assert p in ("bf16", "fp32"), p
    _state["precision"] = p


def act_dtype():
    return torch.bfloat16 if _state["precision"] == "bf16" else torch.float32


def get_device():
    return _state["device"]


def set_device(d):
    _state["device"] = d


def param_version():
    return _state["param_version"]


def bump_param_version():
    """parameters changed (solver step / import_json): cached GEMM operands of the conv layers are stale"""
    _state["param_version"] += 1


def fuse_bn_stats():
    return _state["fuse_bn_stats"] and _state["precision"] == "bf16"


def set_fuse_bn_stats(v):
    _state["fuse_bn_stats"] = bool(v)


def set_rng_seed(v):
    torch.manual_seed(v)


def import_json(json_layers, x, x_shape, layer_range=None):
    """rebuild a layer list from exported JSON (reference denet/layer/__init__.py:31-60)"""
    if layer_range is None:
        layer_start, layer_end = 0, len(json_layers)
    elif type(layer_range) is tuple:
        layer_start, layer_end = layer_range[0], min(len(json_layers), layer_range[1])
    elif type(layer_range) is int:
        layer_start, layer_end = 0, min(len(json_layers), layer_range)
    else:
        raise Exception("Unknown layer range format:", layer_range)

    from .layer_types import layer_types
    layers = [InitialLayer(x, x_shape)]
    for layer_json in json_layers[layer_start:layer_end]:
        layer = None
        for layer_type in layer_types:
            if layer_json["type"] == layer_type.type_name:
                layer = layer_type(layers, json_param=layer_json)
                break
        assert layer is not None, "ERROR Unknown layer type: " + layer_json["type"]
        layer.import_json(layer_json)
        layers.append(layer)
    return layers


# host <-> device traffic of the training step, counted so that bench.py can report it (e2e.h2d/d2h_bytes_per_step)
transfer_bytes = {"h2d": 0, "d2h": 0}


_slots = {}     # slot name -> (pinned host tensor, device tensor, copy-done event)
_slot_uid = itertools.count(1)
_frozen_ns = set()


def new_slot_namespace():
    """suffix that makes the slot names of one model / layer INSTANCE unique: two models (or two sparse layers) with
    equal shapes must not alias each other's staging buffers"""
    return "#%d" % next(_slot_uid)


def freeze_slots(ns, on=True):
    """after a CUDA-graph capture the device tensors of the namespace's slots are baked into the graphs: a shape change
    must not silently re-allocate them (the graphs would keep reading the old tensor)"""
    (_frozen_ns.add if on else _frozen_ns.discard)(ns)


class SlotFrozenError(RuntimeError):
    pass


def _slot_entry(slot, t, dev):
    ent = _slots.get(slot)
    if ent is None or ent[1].shape != t.shape or ent[1].dtype != t.dtype:
        if ent is not None and "#" in slot and slot[slot.rindex("#"):] in _frozen_ns:
            raise SlotFrozenError("slot %s is an input of captured CUDA graphs: shape %s -> %s is not allowed (disable "
                                  "the graphs or keep the batch shape)" % (slot, tuple(ent[1].shape), tuple(t.shape)))
        ent = (torch.empty(t.shape, dtype=t.dtype).pin_memory(), torch.empty(t.shape, dtype=t.dtype, device=dev),
               torch.cuda.Event())
        _slots[slot] = ent
    return ent


def h2d(array, device=None, slot=None):
    """host numpy array / CPU tensor -> device tensor, asynchronous on the current stream.

    With `slot`, the copy goes through a PERSISTENT pinned staging buffer into a PERSISTENT device tensor (same
    address every step: no cudaHostAlloc per call, and the device tensor can be an input of a captured CUDA graph).
    Without it a fresh pinned buffer is used.  CUDA tensors pass through (or are copied into the slot)."""
    t = array if torch.is_tensor(array) else torch.from_numpy(numpy.ascontiguousarray(array))
    dev = device if device is not None else (_state["device"] or "cuda")
    if slot is None:
        if t.is_cuda:
            return t
        transfer_bytes["h2d"] += t.numel() * t.element_size()
        if not t.is_pinned():
            t = t.pin_memory()
        return t.to(dev, non_blocking=True)
    pinned, dst, done = _slot_entry(slot, t, dev)
    if t.is_cuda:
        if t.data_ptr() != dst.data_ptr():
            dst.copy_(t, non_blocking=True)
        return dst
    transfer_bytes["h2d"] += t.numel() * t.element_size()
    if t.is_pinned():
        dst.copy_(t, non_blocking=True)          # caller-owned pinned memory (e.g. the image batch): no staging copy
        return dst
    done.synchronize()                            # the previous copy out of the staging buffer has finished
    pinned.copy_(t)
    dst.copy_(pinned, non_blocking=True)
    done.record()
    return dst


def h2d_on_stream(array, device, slot, stream, after=None):
    """like h2d(..., slot=...) for HOST data, but enqueued on `stream` (a copy stream) after the event `after` (the last
    consumer of the slot's device tensor): the transfer overlaps whatever the compute stream is still doing.  Returns
    (device tensor, event that fires when the data has landed); the consumer's stream must wait on that event."""
    t = array if torch.is_tensor(array) else torch.from_numpy(numpy.ascontiguousarray(array))
    assert not t.is_cuda
    pinned, dst, done = _slot_entry(slot, t, device)
    transfer_bytes["h2d"] += t.numel() * t.element_size()
    ready = torch.cuda.Event()
    with torch.cuda.stream(stream):
        if after is not None:
            stream.wait_event(after)
        if t.is_pinned():
            dst.copy_(t, non_blocking=True)      # caller-owned pinned memory: no staging copy
        else:
            done.synchronize()                    # the previous copy out of the staging buffer has finished
            pinned.copy_(t)
            dst.copy_(pinned, non_blocking=True)
            done.record(stream)
        ready.record(stream)
    return dst, ready


def slot_tensor(slot):
    ent = _slots.get(slot)
    return None if ent is None else ent[1]


def d2h(tensor, slot=None):
    """device tensor -> numpy (synchronises the current stream); with `slot` through a persistent pinned buffer"""
    transfer_bytes["d2h"] += tensor.numel() * tensor.element_size()
    if slot is None:
        return tensor.cpu().numpy()
    key = ("d2h", slot)
    ent = _slots.get(key)
    if ent is None or ent[0].shape != tensor.shape or ent[0].dtype != tensor.dtype:
        ent = (torch.empty(tensor.shape, dtype=tensor.dtype).pin_memory(), None, None)
        _slots[key] = ent
    ent[0].copy_(tensor, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return ent[0].numpy()


def new_param(array):
    """fp32 master parameter in the reference's layout; gradients are managed by the layers, not autograd"""
    t = torch.as_tensor(numpy.ascontiguousarray(numpy.asarray(array, dtype=numpy.float32)))
    return nn.Parameter(t.clone(), requires_grad=False)


def set_param(p, array):
    with torch.no_grad():
        p.copy_(torch.as_tensor(numpy.asarray(array, dtype=numpy.float32)).reshape(p.shape))
    bump_param_version()


def get_param(p):
    return p.detach().cpu().numpy()


class AbstractLayer(nn.Module):
    """reference AbstractLayer (denet/layer/__init__.py:64-143) + explicit forward / backward"""
    type_name = "abstract"
    has_cost = False

    def __init__(self, layer_index, has_split=False):
        super().__init__()
        self.output = self.input = None
        self.output_shape = self.input_shape = None
        self.has_split = has_split
        self.layers = nn.ModuleList()
        self.layer_index = layer_index
        object.__setattr__(self, "_slot_ns", new_slot_namespace())      # per-instance staging-buffer names

    def __str__(self):
        groups = {"int": [], "str": [], "float": [], "bool": [], "tuple": []}
        for k, v in self.__dict__.items():
            if k in ("has_split", "layer_index", "output_shape", "training") or k.startswith("_"):
                continue
            if type(v) is int:
                groups["int"].append(k + ": %i" % v)
            elif type(v) is str:
                groups["str"].append(k + ": " + v)
            elif type(v) is float:
                groups["float"].append(k + ": %.3f" % v)
            elif type(v) is bool:
                groups["bool"].append(k + ": %s" % v)
            elif type(v) is tuple:
                groups["tuple"].append(k + ": " + str(v))
        text = ""
        for name in ("tuple", "str", "int", "float", "bool"):
            if groups[name]:
                text += " " + " ".join(sorted(groups[name]))
        return "%i:" % self.layer_index + self.type_name + " - " + text

    __repr__ = __str__

    # ---- parameter bookkeeping (same meaning as the reference)
    def weights(self):
        return sum([x.weights() for x in self.layers], [])

    def biases(self):
        return sum([x.biases() for x in self.layers], [])

    def params(self):
        return self.weights() + self.biases()

    def updates(self, cost=None):
        """[(state tensor, description)] updated as a side effect of a training forward pass (BN running stats)"""
        return sum([x.updates(cost) for x in self.layers], [])

    def split_forward(self):
        return []

    def split_backward(self, cost, known_grads):
        return []

    def split_known_grads(self):
        return {}

    # ---- costs / targets
    def cost(self, yt_index, yt_value):
        return None

    def get_target(self, model, samples, metas):
        return None

    def set_target(self, yt_index, yt_value):
        pass

    # ---- execution
    def forward(self, x):
        self.output = x
        return x

    def backward(self, dy):
        return dy

    # ---- load / save
    def export_json(self):
        return {"type": type(self).type_name, "layers": [layer.export_json() for layer in self.layers]}

    def import_json(self, json_param):
        if "layers" in json_param:
            for i, json_layer in enumerate(json_param["layers"]):
                self.layers[i].import_json(json_layer)


class InitialLayer(AbstractLayer):
    type_name = "initial"

    def __init__(self, x, x_shape, json_param={}):
        super().__init__(layer_index=0)
        self.output = self.input = x
        self.output_shape = self.input_shape = tuple(x_shape)


class IdentityLayer(AbstractLayer):
    type_name = "identity"

    def __init__(self, layers, json_param={}):
        super().__init__(layer_index=len(layers))
        self.output = self.input = layers[-1].output
        self.output_shape = self.input_shape = layers[-1].output_shape

    @staticmethod
    def parse_desc(layers, name, tags, params):
        return False
