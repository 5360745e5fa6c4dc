"""dev helper: wgrad kernel time by mode (bit0 row-shared, bit1 no MMA, bit2 no TMA)"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from denet_b200 import ops, lib
L = lib.load()
cuda = torch.device("cuda:0")
for (n, h, w, cin, cout, k) in [(32, 128, 128, 64, 64, 3), (32, 32, 32, 256, 256, 3), (32, 64, 64, 128, 128, 3)]:
    x = ops.ActOperand(torch.randn(n, h, w, cin, device=cuda).bfloat16())
    dy = ops.ActOperand(torch.randn(n, h, w, cout, device=cuda).bfloat16())
    dw = torch.empty(cout, cin, k, k, device=cuda)
    flops = 2.0 * n * h * w * cin * cout * k * k
    for mode in [0, 1, 8, 9, 2, 3, 4, 5]:
        L.denet_conv2d_wgrad_set_mode(mode)
        for _ in range(3):
            ops.conv2d_wgrad(dy, x, k, k, (1, 1), (1, 1), dw=dw)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            ops.conv2d_wgrad(dy, x, k, k, (1, 1), (1, 1), dw=dw)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print("%s mode rows=%d noMMA=%d noTMA=%d noEpiStore=%d : %.3f ms  %.0f TFLOP/s" % ((n, h, w, cin, cout, k), mode & 1, (mode >> 1) & 1, (mode >> 2) & 1, (mode >> 3) & 1, ms, flops / ms / 1e9))
L.denet_conv2d_wgrad_set_mode(1)
