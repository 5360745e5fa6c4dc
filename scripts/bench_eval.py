"""eval rate of the DSS detector (the only speed figure the reference publishes: README.md:120-128, DeNet-34 skip 82 Hz
on a Titan X at batch 8 through model-predict): images/s of DeNetDetectLayer.get_detections on synthetic 512x512 images,
host batch in, detection lists out (upload, forward, device sampler, head, NMS, list assembly all inside the timed loop)"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy, torch
import bench
from denet_b200.layer import get_param, set_param

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
workload = sys.argv[2] if len(sys.argv) > 2 else "denet34-skip"
model, data_shape, batch, classes, solver = bench.build_model(workload, batch)
# an untrained corner head never fires (bias 5): give it random corner weights so that the sampler, gather and NMS all
# have work (busy regime, a few hundred RoIs per image)
dnc = [l for l in model.layers if l.type_name == "denet-corner"][0]
rng = numpy.random.RandomState(4)
w = get_param(dnc.layers[1].omega).copy(); w[:4] = rng.randn(*w[:4].shape) * 0.05; set_param(dnc.layers[1].omega, w)
b = get_param(dnc.layers[1].beta).copy(); b[:4] = 3.0; set_param(dnc.layers[1].beta, b)
dnd = model.layers[-1]
wd = get_param(dnd.layers[0].omega).copy(); wd[:] = rng.randn(*wd.shape) * 0.02; set_param(dnd.layers[0].omega, wd)
model.to_device(torch.device("cuda", 0), precision="bf16")
x, metas = bench.synthetic_batch(batch, data_shape, classes, 1)
xp = torch.from_numpy(x).pin_memory()
params = {"prThreshold": 0.05, "nmsThreshold": 0.5}
for _ in range(3):
    res = dnd.get_detections(model, xp, metas, params)
torch.cuda.synchronize()
steps = 20
t0 = time.perf_counter()
for _ in range(steps):
    res = dnd.get_detections(model, xp, metas, params)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
_, _, _, counts = model.detect_forward(xp, dnd)
print(json.dumps({"metric": "denet_eval_images_per_sec", "workload": workload, "batch": batch, "value": batch * steps / dt,
                  "ms_per_batch": 1000 * dt / steps, "rois_per_image": [int(c) for c in counts.cpu()],
                  "detections_per_image": [len(r["detections"]) for r in res], "precision": "bf16",
                  "reference_published": "82 Hz (DeNet-34 skip, Titan X, batch 8, README.md:122)"}))
