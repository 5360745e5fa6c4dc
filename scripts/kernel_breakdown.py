"""dev helper: warm, in-pipeline device time of every C-ABI entry point over eager training steps (CUDA events around
each call on its stream; unlike an ncu launch list the caches are in their steady state)"""
import sys, os, random, collections
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy, torch
import bench
from denet_b200 import lib

steps = 5
model, data_shape, batch, classes, solver = bench.build_model("denet34-skip", 0)
model.to_device(torch.device("cuda", 0), precision="bf16")
model.build_train_func(solver, [])
random.seed(1)
x, metas = bench.synthetic_batch(batch, data_shape, classes, 1)
xd = torch.from_numpy(x).cuda()
hp = bench.SOLVER_HP
for it in range(4):
    model._train_step_device(xd, metas, 0, it, hp["lr"], hp["momentum"], hp["decay"])
torch.cuda.synchronize()
lib.load()
names = [n for n in lib.SIGNATURES if n.startswith("denet_")]
lib.start_timing(names)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for it in range(4, 4 + steps):
    model._train_step_device(xd, metas, 0, it, hp["lr"], hp["momentum"], hp["decay"])
b.record()
t = lib.stop_timing()
tot = a.elapsed_time(b) / steps
rows = sorted(((sum(ms for ms, _ in evs) / steps, len(evs) // steps, n) for n, evs in t.items() if evs), reverse=True)
acc = sum(r[0] for r in rows)
print("eager step %.2f ms; sum of bracketed calls %.2f ms" % (tot, acc))
for ms, cnt, n in rows:
    print("%8.3f ms %5.1f%%  x%-4d %s" % (ms, 100 * ms / acc, cnt, n))
# batch-norm calls by tensor size
for key in ("denet_bn_backward", "denet_bn_apply"):
    by = collections.defaultdict(lambda: [0.0, 0])
    for ms, tag in t.get(key, []):
        by[str(tag[1].input_shape) if tag and hasattr(tag[1], "input_shape") else "?"][0] += ms / steps
    for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:8]:
        print("   %s %s: %.3f ms/step" % (key, k, v[0]))
