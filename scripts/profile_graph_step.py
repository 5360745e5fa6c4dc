"""dev helper: host timeline of the graphed training step"""
import sys, os, time, random
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy, torch
import bench
from denet_b200 import layer as layer_mod, ops

model, data_shape, batch, classes, solver = bench.build_model("denet34-skip", 0)
model.to_device(torch.device("cuda", 0), precision="bf16")
model.build_train_func(solver, [])
random.seed(1)
x, metas = bench.synthetic_batch(batch, data_shape, classes, 1)
xd = torch.from_numpy(x).cuda()
hp = bench.SOLVER_HP
model.enable_cuda_graphs(True)
for it in range(6):
    model._train_step_device(xd, metas, 0, it, hp["lr"], hp["momentum"], hp["decay"])
torch.cuda.synchronize()
si = model._sparse_index()
sp = model.layers[si]
ga, gb = model._graphs
rows = []
for it in range(6, 26):
    t = [time.perf_counter()]
    layer_mod.set_train(True)
    ops.pin_stream(True)
    layer_mod.h2d(xd, model.device, slot="model/image")
    gt = model.upload_metas(metas); layer_mod.set_ground_truth(gt)
    hp_dev = model._write_hp(it, hp["lr"], hp["momentum"], hp["decay"])
    t.append(time.perf_counter())
    ga.replay()
    t.append(time.perf_counter())
    packed = sp.collect_samples()
    t.append(time.perf_counter())
    sp.finish_target(metas, *packed)
    t.append(time.perf_counter())
    gb.replay()
    t.append(time.perf_counter())
    torch.cuda.synchronize()
    t.append(time.perf_counter())
    ops.pin_stream(False)
    rows.append([1e3 * (b - a) for a, b in zip(t[:-1], t[1:])])
print("phase ms: inputs | ga.replay() | wait A (d2h) | finish_target | gb.replay() | wait B")
for r in rows:
    print("  ".join("%6.2f" % v for v in r), "  total %.2f" % sum(r))
