"""dev helper: where does the host time of one training step go? (cProfile of steady-state steps + phase timers)"""
import cProfile, pstats, sys, os, time, random
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy, torch
import bench
from denet_b200 import lib

model, data_shape, batch, classes, solver = bench.build_model("denet34-skip", 0)
model.to_device(torch.device("cuda", 0), precision="bf16")
model.build_train_func(solver, [])
random.seed(1)
x, metas = bench.synthetic_batch(batch, data_shape, classes, 1)
xd = torch.from_numpy(x).cuda()
hp = bench.SOLVER_HP
for it in range(3):
    model._train_step_device(xd, metas, 0, it, hp["lr"], hp["momentum"], hp["decay"])
torch.cuda.synchronize()

# phase timers: host time to ISSUE each phase (no sync) and wall time with sync
for l in model.layers[1:]:
    if l.type_name in ("denet-sparse", "denet-corner", "denet-detect"):
        orig = l.get_target
        def wrapped(m, dx, dm, _o=orig, _l=l):
            t0 = time.perf_counter(); r = _o(m, dx, dm); dt = time.perf_counter() - t0
            print("  get_target %-14s %.2f ms" % (_l.type_name, dt * 1e3))
            return r
        l.get_target = wrapped
for it in range(3, 5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    model._train_step_device(xd, metas, 0, it, hp["lr"], hp["momentum"], hp["decay"])
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("step: host issue %.2f ms, +drain %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
pr = cProfile.Profile()
pr.enable()
for it in range(5, 8):
    model._train_step_device(xd, metas, 0, it, hp["lr"], hp["momentum"], hp["decay"])
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(45)
