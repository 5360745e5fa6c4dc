import sys, os, math
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from denet_b200 import ops, lib
from util import nhwc, relerr
cuda = torch.device("cuda:0")
L = lib.load()
for case in [(1, 64, 64, 64, 64, 3, 1, 1), (2, 16, 16, 64, 64, 3, 1, 1), (4, 8, 8, 64, 96, 3, 1, 1)]:
    n, h, w, cin, cout, k, s, pad = case
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, cin, h, w, generator=g).bfloat16().float()
    dy = torch.randn(n, cout, h, w, generator=g).bfloat16().float()
    xd = ops.ActOperand(nhwc(x, torch.bfloat16, cuda)); dyd = ops.ActOperand(nhwc(dy, torch.bfloat16, cuda))
    L.denet_conv2d_wgrad_set_mode(0)
    ref = ops.conv2d_wgrad(dyd, xd, k, k, (pad, pad), (s, s)).clone()
    for mode in [(1,)]:
        L.denet_conv2d_wgrad_set_mode(*mode)
        dw = ops.conv2d_wgrad(dyd, xd, k, k, (pad, pad), (s, s))
        torch.cuda.synchronize()
        print(case, "mode", mode, "total rel", "%.3e" % relerr(dw, ref),
              "per tap:", [["%.1e" % relerr(dw[:, :, r, c], ref[:, :, r, c]) for c in range(k)] for r in range(k)])
L.denet_conv2d_wgrad_set_mode(1)
