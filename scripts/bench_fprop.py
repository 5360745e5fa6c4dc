"""dev helper: fprop kernel time by kernel variant and profiling knob.
mode = variant | knobs << 4;  variant: 0 one box per tap, 1 tap-group (halo) streaming filters, 3 + resident filters
knobs: 1 epilogue = TMEM read only, 2 no statistics, 4 no MMAs, 8 no TMA loads"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from denet_b200 import ops, lib
L = lib.load()
cuda = torch.device("cuda:0")
cases = [(32, 128, 128, 64, 64, 3), (32, 64, 64, 128, 128, 3), (32, 32, 32, 256, 256, 3), (32, 16, 16, 512, 512, 3)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
knob_sets = [1, 4, 8, 5, 9, 13]


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for (n, h, w, cin, cout, k) in cases:
    x = ops.ActOperand(torch.randn(n, h, w, cin, device=cuda).bfloat16())
    wt = torch.randn(cout, cin, k, k, device=cuda) * 0.05
    wop = ops.conv_weight_prep(wt, 0, False)
    out = ops.alloc_nhwc(n, h, w, cout, torch.bfloat16, cuda)
    s0, s1 = torch.zeros(cout, device=cuda), torch.zeros(cout, device=cuda)
    flops = 2.0 * n * h * w * cin * cout * k * k
    for stats in (True,):
        for variant in (7, 15):
            row = []
            for kn in knob_sets:
                L.denet_conv2d_fprop_set_mode(variant | (kn << 4))
                ms = timeit(lambda: ops.conv2d_fprop(x, wop, (k // 2, k // 2), (h, w), torch.bfloat16,
                                                     stats=(s0, s1) if stats else None, out=out))
                row.append("k%d:%.3f" % (kn, ms))
            L.denet_conv2d_fprop_set_mode(variant)
            ms = timeit(lambda: ops.conv2d_fprop(x, wop, (k // 2, k // 2), (h, w), torch.bfloat16,
                                                 stats=(s0, s1) if stats else None, out=out))
            print("%s stats=%d variant=%d: %.3f ms %.0f TFLOP/s | %s" % ((n, h, w, cin, cout, k), stats, variant, ms,
                                                                       flops / ms / 1e9, " ".join(row)), flush=True)
L.denet_conv2d_fprop_set_mode(15)
sys.exit(0)
# stem
n, h, w, cin, cout, k, s, pad = 32, 512, 512, 3, 64, 7, 2, 3
oh = ow = 256
geom = ops.rowfold_geometry((h, w), cin, (k, k), (s, s), (pad, pad), (oh, ow))
img = ops.PaddedImage(n, cin, h, w, geom[0], (pad, pad), geom[1], geom[2], False, cuda).fill(torch.rand(n, cin, h, w, device=cuda))
wt = torch.randn(cout, cin, k, k, device=cuda) * 0.05
wop = ops.conv_weight_prep_rowfold(wt, geom[0], False)
s0, s1 = torch.zeros(cout, device=cuda), torch.zeros(cout, device=cuda)
bias = torch.zeros(cout, device=cuda)
flops = 2.0 * n * oh * ow * cin * cout * k * k
for variant in (0, 1, 3):
    row = []
    for kn in knob_sets:
        L.denet_conv2d_fprop_set_mode(variant | (kn << 4))
        ms = timeit(lambda: ops.conv2d_rowfold_fprop(img, wop, (s, s), (oh, ow), torch.bfloat16, bias=bias, stats=(s0, s1)))
        row.append("k%d:%.3f" % (kn, ms))
    L.denet_conv2d_fprop_set_mode(variant)
    ms = timeit(lambda: ops.conv2d_rowfold_fprop(img, wop, (s, s), (oh, ow), torch.bfloat16, bias=bias, stats=(s0, s1)))
    print("stem variant=%d: %.3f ms %.0f TFLOP/s | %s" % (variant, ms, flops / ms / 1e9, " ".join(row)), flush=True)
L.denet_conv2d_fprop_set_mode(3)
