"""dev helper: filter-gradient kernel A/B as CUDA-graph replays: denet_conv2d_wgrad_set_mode 1 (default: CTA pairs for
the 256-wide tiles) vs 129 (bit 7: one CTA per tile), partial sums only (no reduction)"""
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


class Own:
    pass


for (n, h, w, cin, cout, k) in [(32, 32, 32, 256, 256, 3), (32, 16, 16, 512, 512, 3), (32, 32, 32, 512, 256, 3),
                                (32, 24, 24, 4706, 1536, 1), (32, 24, 24, 1536, 1024, 1)]:
    xt = ops.alloc_nhwc(n, h, w, cin, torch.bfloat16, cuda)
    xt.copy_(torch.randn(n, h, w, cin, device=cuda))
    dyt = ops.alloc_nhwc(n, h, w, cout, torch.bfloat16, cuda)
    dyt.copy_(torch.randn(n, h, w, cout, device=cuda))
    x, dy = ops.ActOperand(xt), ops.ActOperand(dyt)
    dw = torch.empty(cout, cin, k, k, device=cuda)
    flops = 2.0 * n * h * w * cin * cout * k * k
    pad = ((k - 1) // 2, (k - 1) // 2)
    res = {}
    outs = {}
    for mode in (129, 1, 129 | 2, 1 | 2, 129 | 4, 1 | 4):
        L.denet_conv2d_wgrad_set_mode(mode)
        if mode in (1, 129):
            ops.conv2d_wgrad(dy, x, k, k, pad, (1, 1), dw=dw)
            outs[mode] = dw.clone()
        pend = []
        res[mode] = timeit(lambda: (ops.conv2d_wgrad(dy, x, k, k, pad, (1, 1), dw=dw, defer=(pend, Own)), pend.clear()))
    L.denet_conv2d_wgrad_set_mode(1)
    err = ((outs[1] - outs[129]).norm() / outs[129].norm()).item()
    print("%s: one CTA %.1f us (%.0f TF) noMMA %.1f noTMA %.1f | pairs %.1f us (%.0f TF) noMMA %.1f noTMA %.1f | rel diff %.1e" % (
        (n, h, w, cin, cout, k), res[129], flops / res[129] / 1e6, res[131], res[133], res[1], flops / res[1] / 1e6,
        res[3], res[5], err), flush=True)
