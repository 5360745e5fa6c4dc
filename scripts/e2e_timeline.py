"""dev helper: host-side timeline of graphed training steps fed from pinned host memory"""
import sys, os, random, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy, torch
import bench
from denet_b200 import lib, layer as layer_mod

model, data_shape, batch, classes, solver = bench.build_model("denet34-skip", 0)
model.to_device(torch.device("cuda", 0), precision="bf16")
model.build_train_func(solver, [])
random.seed(1)
x, metas = bench.synthetic_batch(batch, data_shape, classes, 1)
xp = torch.from_numpy(x).pin_memory()
xd = torch.from_numpy(x).cuda()
hp = bench.SOLVER_HP
model.enable_cuda_graphs(True)
it = 0
def host_step():
    global it
    c = model.train_step(xp, metas, 0, it, hp["lr"], hp["momentum"], hp["decay"])[0]; it += 1; return c
def host_step_nocost():
    global it
    model._train_step_device(xp, metas, 0, it, hp["lr"], hp["momentum"], hp["decay"]); it += 1
def dev_step():
    global it
    model._train_step_device(xd, metas, 0, it, hp["lr"], hp["momentum"], hp["decay"]); it += 1
for _ in range(8):
    host_step()
for name, fn in [("device batch", dev_step), ("host batch, costs read", host_step), ("host batch, no cost read", host_step_nocost),
                 ("device batch", dev_step), ("host batch, costs read", host_step)]:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    ts = []
    for _ in range(10):
        t1 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t1) * 1e3)
    b.record(); torch.cuda.synchronize()
    print("%-26s %.2f ms/step (events)  %.2f ms/step (wall)  host time per call: %s" % (
        name, a.elapsed_time(b) / 10, (time.perf_counter() - t0) * 100, " ".join("%.1f" % t for t in ts)))
# where does the host spend its time inside one graphed step?
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    host_step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
