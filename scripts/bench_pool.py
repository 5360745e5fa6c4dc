import sys, os
sys.path.insert(0, "/root/repo")
import torch
from denet_b200 import ops
cuda = torch.device("cuda:0")
x = torch.randn(32, 256, 256, 64, device=cuda).bfloat16()
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1000
y, arg = ops.pool_fwd(x, 0, (3, 3), (2, 2), (1, 1), (128, 128))
print("maxpool fwd 3x3s2 bf16: %.1f us" % timeit(lambda: ops.pool_fwd(x, 0, (3, 3), (2, 2), (1, 1), (128, 128))))
print("maxpool fwd 3x3s2 bf16 (generic kernel via pad 0 geometry 2x2s2): %.1f us" % timeit(lambda: ops.pool_fwd(x, 0, (2, 2), (2, 2), (0, 0), (128, 128))))
dy = torch.randn_like(y)
print("maxpool bwd: %.1f us" % timeit(lambda: ops.pool_bwd(dy, 0, (3, 3), (2, 2), (1, 1), tuple(x.shape), arg)))
