"""dev helper: per-parameter gradient errors of one fp32-parity training step vs the oracle"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy, torch
import test_gpu_model as T
from oracle.ref_model import RefModel
from util import relerr

which = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
if which == "cfg1":
    model = T.build(T.CFG1, (3, 32, 32), 32, 10, "fp32")
    model.to_device(precision="fp32"); model.build_train_func("sgd", [])
    numpy.random.seed(1)
    x = numpy.random.uniform(0, 1, (32, 3, 32, 32)).astype(numpy.float32)
    metas = [{"image_class": int(c), "bbox": [], "class": []} for c in numpy.random.randint(0, 10, 32)]
    js, before, cap, cost, costs = T.run_step(model, x, metas, "sgd", 0.1, (0.9, 0.9), 1e-4, 0)
else:
    model = T.build(T.RESNET_SMALL, (3, 64, 64), 8, 10, "fp32", False)
    model.to_device(precision="fp32"); model.build_train_func("nesterov", [])
    numpy.random.seed(2)
    x = numpy.random.uniform(0, 1, (8, 3, 64, 64)).astype(numpy.float32)
    metas = [{"image_class": int(c), "bbox": [], "class": []} for c in numpy.random.randint(0, 10, 8)]
    js, before, cap, cost, costs = T.run_step(model, x, metas, "nesterov", 0.05, (0.9, 0.9), 1e-4, 1)
ref = RefModel(js, x.shape, model.class_num, dtype=torch.float64)
ref.relu_masks = T.relu_masks(model)
ref.pool_argmax = T.pool_argmax(model)
total, ref_costs, grads, out = ref.train_gradients(x, [t for _, t in cap])
print("cost", cost, total, "relu overrides", ref.relu_overrides)
mine = T.named_params(model)
for name, g in grads.items():
    print("%-20s %-22s rel %.3e  |g| %.3e" % (name, tuple(g.shape), relerr(mine[name].grad, g), g.norm().item()))
