"""dev helper: short-K tiles (the parity-class launches of the strided data gradients are 1-4 taps x 2 chunks): time and
profiling knobs (denet_conv2d_fprop_set_mode: 1 cheap epilogue, 4 no MMAs, 8 no TMA loads) as CUDA-graph replays"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch

from denet_b200 import lib, ops

L = lib.load()
cuda = torch.device("cuda:0")


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1000.0


for (n, h, w, cin, cout, k) in [(32, 64, 64, 128, 64, 2), (32, 64, 64, 128, 64, 1), (32, 32, 32, 256, 128, 2),
                                (32, 64, 64, 128, 64, 3)]:
    x = ops.ActOperand(torch.randn(n, h, w, cin, device=cuda).bfloat16())
    wt = torch.randn(cout, cin, k, k, device=cuda) * 0.05
    wop = ops.conv_weight_prep(wt, 0, False)
    out = ops.alloc_nhwc(n, h, w, cout, torch.bfloat16, cuda)
    res = torch.randn_like(out)
    flops = 2.0 * n * h * w * cin * cout * k * k
    pad = ((k - 1) // 2, (k - 1) // 2)
    row = []
    for kn in (0, 1, 4, 8, 13):
        L.denet_conv2d_fprop_set_mode(15 | (kn << 4))
        row.append("k%d:%.1f" % (kn, timeit(lambda: ops.conv2d_fprop(x, wop, pad, (h, w), torch.bfloat16, out=out))))
    L.denet_conv2d_fprop_set_mode(15)
    t_res = timeit(lambda: ops.conv2d_fprop(x, wop, pad, (h, w), torch.bfloat16, residual=res, out=out))
    t_reg = None
    L.denet_conv2d_fprop_set_mode(7)
    t_reg = timeit(lambda: ops.conv2d_fprop(x, wop, pad, (h, w), torch.bfloat16, out=out))
    L.denet_conv2d_fprop_set_mode(15)
    print("%s: us %s | +residual %.1f | register epilogue %.1f | %.0f TFLOP/s" % (
        (n, h, w, cin, cout, k), " ".join(row), t_res, t_reg, flops / float(row[0].split(":")[1]) / 1e6), flush=True)
