"""dev helper (context, not product): cuDNN (through torch, bf16, channels_last) timed once per distinct conv shape of a
bench workload - the "practical upper reference" of SURVEY.md §2a - summed to a per-step conv time that compares with
bench.py's `roofline.conv_ms_per_step`.  Nothing under denet_b200/ calls cuDNN.

usage: python scripts/cudnn_context.py [workload] [batch] > profiles/rN_cudnn_context.txt"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch

import bench
from denet_b200.model.model_cnn import _walk

workload = sys.argv[1] if len(sys.argv) > 1 else "denet34-skip"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
model, data_shape, batch, classes, solver = bench.build_model(workload, batch)
cuda = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = True


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


shapes = {}
for l in _walk(model.layers):
    if l.type_name != "conv" or not l.enabled:
        continue
    pad = l.pad if isinstance(l.pad[0], int) else (0, 0)
    key = (tuple(l.input_shape), tuple(l.filter_shape), tuple(l.stride), tuple(pad), bool(l.is_first))
    shapes[key] = shapes.get(key, 0) + 1

tot = {"fprop": [0.0, 0.0], "dgrad": [0.0, 0.0], "wgrad": [0.0, 0.0]}
print("# cuDNN %s via torch %s, bf16 NHWC, %s batch %d; per-shape times are isolated back-to-back launches" % (
    torch.backends.cudnn.version(), torch.__version__, workload, batch))
print("# %-24s %-20s s  n | fprop ms TF/s | dgrad ms TF/s | wgrad ms TF/s" % ("input", "filter"))
for (ishape, fshape, stride, pad, first), n in sorted(shapes.items(), key=lambda kv: -kv[1]):
    nb, cin, h, w = ishape
    cout, _, r, s = fshape
    x = torch.randn(nb, cin, h, w, device=cuda).bfloat16().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, r, s, device=cuda) * 0.05).bfloat16().contiguous(memory_format=torch.channels_last)
    y = torch.nn.functional.conv2d(x, wt, None, stride, pad)
    dy = torch.randn_like(y)
    flops = 2.0 * y.numel() * cin * r * s
    args = (dy, x, wt, None, list(stride), list(pad), [1, 1], False, [0, 0], 1)
    t_f = timeit(lambda: torch.nn.functional.conv2d(x, wt, None, stride, pad))
    t_d = 0.0 if first else timeit(lambda: torch.ops.aten.convolution_backward(*args, [True, False, False]))
    t_w = timeit(lambda: torch.ops.aten.convolution_backward(*args, [False, True, False]))
    for k, t in (("fprop", t_f), ("dgrad", t_d), ("wgrad", t_w)):
        if t > 0:
            tot[k][0] += t * n
            tot[k][1] += flops * n
    tf = lambda t: flops / t / 1e9 if t > 0 else 0.0
    print("%-26s %-20s %d x%-2d | %7.3f %5.0f | %7.3f %5.0f | %7.3f %5.0f" % (
        ishape, fshape, stride[0], n, t_f, tf(t_f), t_d, tf(t_d), t_w, tf(t_w)), flush=True)
    del x, wt, y, dy
ms = sum(v[0] for v in tot.values())
fl = sum(v[1] for v in tot.values())
for k, (t, f) in tot.items():
    print("# cuDNN %s: %.3f ms/step, %.0f TFLOP/s" % (k, t, f / t / 1e9))
print("# cuDNN conv stack: %.3f ms/step, %.0f TFLOP/s (sum of isolated kernels, no batch-norm statistics in the "
      "epilogue, no residual / ReLU fusion, bf16 weight-gradient output); compare bench.py roofline.conv_ms_per_step" % (
          ms, fl / ms / 1e9))
