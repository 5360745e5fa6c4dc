"""dev helper: eager vs CUDA-graph training steps of the small classifier under kernel-variant switches"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy, torch
import test_gpu_model as T
from util import relerr
from denet_b200 import lib, layer as layer_mod
L = lib.load()
def make(defer=True):
    m = T.build(T.RESNET_SMALL, (3, 64, 64), 8, 10, "bf16", False)
    m.to_device(precision="bf16"); m.build_train_func("nesterov", [])
    m.defer_wgrad_reduce = defer
    return m
numpy.random.seed(3)
x = numpy.random.uniform(0, 1, (8, 3, 64, 64)).astype(numpy.float32)
metas = [{"image_class": int(c), "bbox": [], "class": []} for c in numpy.random.randint(0, 10, 8)]
for label, fmode, defer, fuse in [("default", 3, True, True), ("fprop old", 0, True, True), ("no defer", 3, False, True),
                                   ("no bn fuse", 3, True, False), ("all old", 0, False, False)]:
    L.denet_conv2d_fprop_set_mode(fmode)
    layer_mod.set_fuse_bn_stats(fuse)
    out = {}
    for name, g in [("eager1", False), ("eager2", False), ("graph", True), ("graph2", True)]:
        m = make(defer)
        out[name] = (T._steps(m, x, metas, 4, g), T.named_params(m))
    print(label)
    for name in out:
        print("   %-7s" % name, [round(c[0], 5) for c in out[name][0]])
    for a, b in [("eager1", "eager2"), ("eager1", "graph"), ("graph", "graph2")]:
        worst = max((relerr(out[b][1][k], out[a][1][k]), k) for k in out[a][1] if out[a][1][k].norm().item() > 1e-2)
        print("   %s vs %s: worst param rel diff %.3e (%s)" % (a, b, worst[0], worst[1]))
