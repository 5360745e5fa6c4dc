"""dev helper: per-tile clock stamps of CTA 0's roles in conv_fprop_halo_kernel (see denet_conv2d_fprop_set_timeline)"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from denet_b200 import ops, lib
L = lib.load()
cuda = torch.device("cuda:0")
buf = torch.zeros(3 * 64 * 4, dtype=torch.int64, device=cuda)
cases = [(32, 128, 128, 64, 64, 3), (32, 64, 64, 128, 128, 3)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
for (n, h, w, cin, cout, k) in cases:
    x = ops.ActOperand(torch.randn(n, h, w, cin, device=cuda).bfloat16())
    wt = torch.randn(cout, cin, k, k, device=cuda) * 0.05
    wop = ops.conv_weight_prep(wt, 0, False)
    out = ops.alloc_nhwc(n, h, w, cout, torch.bfloat16, cuda)
    for mode in (3, 3 | (13 << 4), 3 | (4 << 4), 1):
        L.denet_conv2d_fprop_set_mode(mode)
        L.denet_conv2d_fprop_set_timeline(None)
        for _ in range(2):
            ops.conv2d_fprop(x, wop, ((k - 1) // 2, (k - 1) // 2), (h, w), torch.bfloat16, out=out)
        buf.zero_()
        L.denet_conv2d_fprop_set_timeline(buf.data_ptr())
        ops.conv2d_fprop(x, wop, ((k - 1) // 2, (k - 1) // 2), (h, w), torch.bfloat16, out=out)
        torch.cuda.synchronize()
        L.denet_conv2d_fprop_set_timeline(None)
        t = buf.cpu().view(3, 64, 4)
        t0 = int(t[t > 0].min())
        print("case", (n, h, w, cin, cout, k), "mode", mode & 15, "knobs", mode >> 4)
        for i in range(14):
            pr, mm, ep = t[0, i], t[1, i], t[2, i]
            f = lambda v: "%6d" % (int(v) - t0) if int(v) > 0 else "     -"
            print("  tile %2d | prod start %s issued %s | mma start %s tempty-ok %s a-full-ok %s committed %s | epi start %s tfull-ok %s done %s"
                  % (i, f(pr[0]), f(pr[1]), f(mm[0]), f(mm[1]), f(mm[2]), f(mm[3]), f(ep[0]), f(ep[1]), f(ep[2])))
L.denet_conv2d_fprop_set_mode(3)
