"""dev helper: wgrad kernel time by mode (bit0 row-shared, bit1 no MMA, bit2 no TMA, bit3 no store, bit4 legacy splits,
bit5 per-tap MMAs in the row-shared kernel); timed per call with CUDA events inside one stream"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from denet_b200 import ops, lib
L = lib.load()
cuda = torch.device("cuda:0")
for (n, h, w, cin, cout, k) in [(32, 128, 128, 64, 64, 3), (32, 64, 64, 128, 128, 3), (32, 32, 32, 256, 256, 3), (32, 16, 16, 512, 512, 3)]:
    x = ops.ActOperand(torch.randn(n, h, w, cin, device=cuda).bfloat16())
    dy = ops.ActOperand(torch.randn(n, h, w, cout, device=cuda).bfloat16())
    dw = torch.empty(cout, cin, k, k, device=cuda)
    flops = 2.0 * n * h * w * cin * cout * k * k
    ref = None
    pend = []
    class Own: pass
    for mode in [1, 65, 3, 5, 67, 69]:
        L.denet_conv2d_wgrad_set_mode(mode)
        for _ in range(3):
            ops.conv2d_wgrad(dy, x, k, k, (1, 1), (1, 1), dw=dw)
        if mode in (1, 65, 0):
            if ref is None:
                ref = dw.clone()
            else:
                err = ((dw - ref).norm() / ref.norm()).item()
                assert err < 1e-4, (mode, err)
        evs = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.conv2d_wgrad(dy, x, k, k, (1, 1), (1, 1), dw=dw, defer=(pend, Own)); b.record(); pend.clear()
            evs.append((a, b))
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
        print("%s mode %2d rows=%d noMMA=%d noTMA=%d noStore=%d pertap=%d : %.3f ms  %.0f TFLOP/s (incl. reduce)" % (
            (n, h, w, cin, cout, k), mode, mode & 1, (mode >> 1) & 1, (mode >> 2) & 1, (mode >> 3) & 1, (mode >> 5) & 1,
            ms, flops / ms / 1e9), flush=True)
L.denet_conv2d_wgrad_set_mode(1)
