"""Developer check for the tcgen05 conv kernels on a GPU box (not part of the test-suite).

Compares fprop / dgrad / wgrad against an fp64 torch reference on the device. Run under gpurun:
    timeout 300 python scripts/dev_conv_check.py
"""
import sys
import os
import time
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from denet_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def ref_conv(x_nhwc, w, pad):
    # reference semantics: true convolution == correlation with the flipped filter
    x = x_nhwc.double().permute(0, 3, 1, 2)
    y = F.conv2d(x, w.double().flip(2, 3), padding=pad)
    return y.permute(0, 2, 3, 1).contiguous()


def relerr(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def check(name, err, tol):
    ok = err < tol
    print("%-58s rel_err %.3e  (tol %.0e) %s" % (name, err, tol, "OK" if ok else "FAIL"), flush=True)
    return ok


def run_case(n, h, w_, cin, cout, k, pad, split):
    all_ok = True
    x = ops.alloc_nhwc(n, h, w_, cin, torch.float32, zero=True)
    x.copy_(torch.randn(n, h, w_, cin, device=dev))
    wt = torch.randn(cout, cin, k, k, device=dev) * (1.0 / (cin * k * k) ** 0.5)
    ho, wo = h + 2 * pad - k + 1, w_ + 2 * pad - k + 1
    tag = "n%d %dx%d cin%d cout%d k%d p%d %s" % (n, h, w_, cin, cout, k, pad, "bf16x3" if split else "bf16")
    if split:
        xop = ops.act_operand(x)
        xr, wr = x, wt
        tol = 2e-5
    else:
        xb = ops.alloc_nhwc(n, h, w_, cin, torch.bfloat16, zero=True)
        xb.copy_(x)
        xop = ops.ActOperand(xb)
        xr, wr = xb.float(), wt.to(torch.bfloat16).float()
        tol = 2e-5
    wop = ops.conv_weight_prep(wt, 0, split)
    bias = torch.randn(cout, device=dev)
    y = ops.conv2d_fprop(xop, wop, pad, pad, (ho, wo), torch.float32, bias=bias)
    torch.cuda.synchronize()
    yref = ref_conv(xr, wr, pad) + bias.double()
    all_ok &= check("fprop " + tag, relerr(y, yref), tol)

    # dgrad: correlate dy with the mode-1 operand, pad' = k-1-pad
    dy = ops.alloc_nhwc(n, ho, wo, cout, torch.float32, zero=True)
    dy.copy_(torch.randn(n, ho, wo, cout, device=dev))
    if split:
        dyop = ops.act_operand(dy)
        dyr = dy
    else:
        dyb = ops.alloc_nhwc(n, ho, wo, cout, torch.bfloat16, zero=True)
        dyb.copy_(dy)
        dyop = ops.ActOperand(dyb)
        dyr = dyb.float()
    wop_d = ops.conv_weight_prep(wt, 1, split)
    dx = ops.conv2d_fprop(dyop, wop_d, k - 1 - pad, k - 1 - pad, (h, w_), torch.float32)
    torch.cuda.synchronize()
    xg = xr.double().permute(0, 3, 1, 2).requires_grad_(True)
    yg = F.conv2d(xg, wr.double().flip(2, 3), padding=pad)
    dx_ref, = torch.autograd.grad(yg, xg, dyr.double().permute(0, 3, 1, 2))
    all_ok &= check("dgrad " + tag, relerr(dx, dx_ref.permute(0, 2, 3, 1)), tol)

    # wgrad
    dw = ops.conv2d_wgrad(dyop, xop, k, k, pad, pad)
    torch.cuda.synchronize()
    wg = wr.double().requires_grad_(True)
    yg = F.conv2d(xr.double().permute(0, 3, 1, 2), wg.flip(2, 3), padding=pad)
    dw_ref, = torch.autograd.grad(yg, wg, dyr.double().permute(0, 3, 1, 2))
    all_ok &= check("wgrad " + tag, relerr(dw, dw_ref), tol)
    return all_ok


def bench(n, h, w_, cin, cout, k, pad, iters=20):
    x = torch.randn(n, h, w_, cin, device=dev).to(torch.bfloat16)
    wt = torch.randn(cout, cin, k, k, device=dev) * 0.05
    xop = ops.ActOperand(x)
    wop = ops.conv_weight_prep(wt, 0, False)
    ho, wo = h + 2 * pad - k + 1, w_ + 2 * pad - k + 1
    y = torch.empty(n, ho, wo, cout, device=dev, dtype=torch.bfloat16)
    dyop = ops.ActOperand(torch.randn(n, ho, wo, cout, device=dev).to(torch.bfloat16))
    flops = 2.0 * n * ho * wo * cout * cin * k * k
    for what in ("fprop", "wgrad"):
        def run():
            if what == "fprop":
                ops.conv2d_fprop(xop, wop, pad, pad, (ho, wo), torch.bfloat16, out=y)
            else:
                ops.conv2d_wgrad(dyop, xop, k, k, pad, pad)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print("bench %s n%d %dx%d cin%d cout%d k%d: %.3f ms  %.1f TFLOP/s" %
              (what, n, h, w_, cin, cout, k, ms, flops / ms / 1e9), flush=True)


if __name__ == "__main__":
    t0 = time.time()
    ok = True
    cases = [
        (1, 1, 512, 128, 128, 1, 0),
        (2, 16, 16, 64, 64, 3, 1),
        (4, 32, 32, 128, 256, 3, 1),
        (3, 14, 14, 96, 100, 3, 1),
        (8, 8, 8, 256, 512, 3, 1),
        (2, 24, 24, 200, 85, 1, 0),
        (32, 7, 7, 64, 40, 7, 0),
    ]
    for split in (False, True):
        for c in cases:
            try:
                ok &= run_case(*c, split)
            except Exception as e:  # keep going: one broken case should not hide the others
                ok = False
                print("EXCEPTION in case", c, split, repr(e), flush=True)
    print("ALL OK" if ok else "SOME FAILED", "(%.1fs)" % (time.time() - t0), flush=True)
    if "--bench" in sys.argv:
        bench(32, 32, 32, 256, 256, 3, 1)
        bench(32, 64, 64, 128, 128, 3, 1)
        bench(32, 128, 128, 64, 64, 3, 1)
        bench(32, 16, 16, 512, 512, 3, 1)
        bench(1, 1, 18432, 4736, 1536, 1, 0)
        bench(1, 1, 18432, 1536, 1024, 1, 0)
