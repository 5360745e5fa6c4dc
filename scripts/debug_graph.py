import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy, torch
import test_gpu_model as T
from util import synthetic_metas, relerr
def make():
    m = T.build(T.DENET_SMALL, (3, 128, 128), 4, 20, "bf16", convert=True)
    m.to_device(precision="bf16"); m.build_train_func("nesterov", [])
    return m
numpy.random.seed(3)
x = numpy.random.uniform(0, 1, (4, 3, 128, 128)).astype(numpy.float32)
metas = synthetic_metas(4, 20, seed=3, max_boxes=4)
runs = {}
for name, g in [("eager1", False), ("eager2", False), ("graph", True)]:
    m = make()
    runs[name] = (T._steps(m, x, metas, 6, g), T.named_params(m))
    print(name, [round(c[0], 5) for c in runs[name][0]])
for a, b in [("eager1", "eager2"), ("eager1", "graph")]:
    worst = max(relerr(runs[b][1][k], runs[a][1][k]) for k in runs[a][1])
    print(a, b, "worst param rel diff %.3e" % worst)
