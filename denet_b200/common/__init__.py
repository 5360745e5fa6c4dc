"""Helpers shared by layers and model (reference denet/common/__init__.py, hot-path subset)."""
import math
import time


class Timer:
    """wall-clock marks (reference common/__init__.py:16-46)"""

    def __init__(self):
        self.reset()

    def reset(self):
        self.times = [time.time()]

    def mark(self):
        self.times.append(time.time())

    def current(self):
        return time.time() - self.times[0]

    def current_ms(self):
        return 1000.0 * self.current()

    def delta_ms(self, index):
        return 1000.0 * (self.times[index + 1] - self.times[index])


def convert_num(s):
    """'12' -> 12, '0.5' -> 0.5, anything else stays a string (model-desc parameters)"""
    try:
        return int(s)
    except ValueError:
        try:
            return float(s)
        except ValueError:
            return s


def find_layers(layers, layer_names, warn_missing=False):
    """first layer of each requested type_name (reference common/__init__.py:65-86)"""
    single = isinstance(layer_names, str)
    names = [layer_names] if single else list(layer_names)
    found = [None] * len(names)
    for layer in layers:
        for i, name in enumerate(names):
            if found[i] is None and layer.type_name == name:
                found[i] = layer
    if warn_missing:
        missed = [names[i] for i, f in enumerate(found) if f is None]
        if missed:
            raise Exception("Could not find layers of name: ", missed)
    return found[0] if single else found


def overlap(bbox0, bbox1=(0, 0, 1, 1)):
    dx = max(0, min(bbox0[2], bbox1[2]) - max(bbox0[0], bbox1[0]))
    dy = max(0, min(bbox0[3], bbox1[3]) - max(bbox0[1], bbox1[1]))
    return dx * dy


def overlap_iou(bbox0, bbox1=(0, 0, 1, 1)):
    """area of intersection / area of union (reference common/__init__.py:103-109)"""
    a0 = (bbox0[2] - bbox0[0]) * (bbox0[3] - bbox0[1])
    a1 = (bbox1[2] - bbox1[0]) * (bbox1[3] - bbox1[1])
    ai = overlap(bbox0, bbox1)
    return ai / (a0 + a1 - ai)


def ndarray_unpack(v, shapes):
    index = 0
    r = []
    for shape in shapes:
        size = int(math.prod(shape))
        r.append(v[index:index + size].reshape(shape))
        index += size
    return r


def get_params_dict(params):
    """'a=1,b' -> {'a': 1, 'b': True} (reference common/__init__.py:200-208)"""
    out = {}
    for item in params.split(","):
        pv = item.split("=")
        out[pv[0]] = True if len(pv) == 1 else convert_num(pv[1])
    return out


def import_c(fname):
    """reference common.import_c (common/__init__.py:171-195) JIT-compiles a CPython extension from a .cc file next to
    the layer; here the same call returns the GPU-backed module with the extension's functions and signatures
    (denet_sparse.cc -> denet_b200.layer.denet_sparse_c)"""
    import importlib
    import os
    base = os.path.splitext(os.path.basename(fname))[0]
    table = {"denet_sparse": "denet_b200.layer.denet_sparse_c", "denet_detect": "denet_b200.layer.denet_detect_c"}
    if base not in table:
        raise ImportError("import_c: no B200 replacement for %s" % fname)
    return importlib.import_module(table[base])
