"""dev helper: batch-norm forward apply (statistics finalised in the launch) and backward on the tensor shapes of the
DeNet-34 step, timed as back-to-back launches (small tensors stay in L2 like inside the step).
usage: python scripts/bench_bn.py [mode ...]   (modes = denet_bn_set_mode values to compare, default: current)"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch

from denet_b200 import lib, ops

L = lib.load()
cuda = torch.device("cuda:0")
shapes = [(32, 256, 256, 64), (32, 128, 128, 64), (32, 64, 64, 128), (32, 32, 32, 256), (32, 16, 16, 512)]
modes = [int(a) for a in sys.argv[1:]] or [None]


def timeit(fn, reps=20):
    """CUDA-graph replay of `reps` launches (the host's ctypes call costs more than the small kernels run)"""
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


for (n, h, w, c) in shapes:
    x = torch.randn(n, h, w, c, device=cuda).bfloat16()
    dy = torch.randn(n, h, w, c, device=cuda).bfloat16()
    res = torch.randn(n, h, w, c, device=cuda).bfloat16()
    y, dx = torch.empty_like(x), torch.empty_like(x)
    M = n * h * w
    xf = x.float().reshape(M, c)
    sums, sq = xf.sum(0).contiguous(), (xf * xf).sum(0).contiguous()
    gamma, beta = torch.rand(c, device=cuda) + 0.5, torch.randn(c, device=cuda) * 0.1
    mean, invstd = torch.zeros(c, device=cuda), torch.zeros(c, device=cuda)
    dg, db = torch.zeros(c, device=cuda), torch.zeros(c, device=cuda)
    mb = x.numel() * 2 / 1e6
    for mode in modes:
        if mode is not None:
            L.denet_bn_set_mode(mode)
        t_a = timeit(lambda: ops.bn_apply_sums(x, sums, sq, 1e-5, gamma, beta, mean, invstd, relu=True, out=y))
        t_r = timeit(lambda: ops.bn_apply_sums(x, sums, sq, 1e-5, gamma, beta, mean, invstd, residual=res, relu=True,
                                               out=y))
        t_b = timeit(lambda: ops.bn_backward(dy, None, x, mean, invstd, gamma, True, dg, db, dx=dx, beta=beta))
        t_c = timeit(lambda: ops.bn_backward(dy, y, x, mean, invstd, gamma, True, dg, db, want_dres=True, dx=dx))
        print("%-20s %6.1f MB mode %s | apply %6.1f us (%4.0f GB/s)  apply+res %6.1f us | bwd %6.1f us (%4.0f GB/s)  "
              "bwd+y+dres %6.1f us" % ((n, h, w, c), mb, mode, t_a, 2 * mb / t_a * 1e3, t_r, t_b, 3 * mb / t_b * 1e3,
                                       t_c), flush=True)
