import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from denet_b200 import ops, lib
L = lib.load()
mode = int(sys.argv[1])
cuda = torch.device("cuda:0")
n, h, w, cin, cout, k = 32, 128, 128, 64, 64, 3
x = ops.ActOperand(torch.randn(n, h, w, cin, device=cuda).bfloat16())
dy = ops.ActOperand(torch.randn(n, h, w, cout, device=cuda).bfloat16())
dw = torch.empty(cout, cin, k, k, device=cuda)
L.denet_conv2d_wgrad_set_mode(mode)
for _ in range(3):
    ops.conv2d_wgrad(dy, x, k, k, (1, 1), (1, 1), dw=dw)
torch.cuda.synchronize()
