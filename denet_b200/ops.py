"""Tensor-level wrappers over the C-ABI (include/denet_b200.h).

torch is used for device memory and streams only; every op here is one or more calls into libdenet_b200.so
on the current CUDA stream.  Activations are NHWC tensors (N, H, W, C) whose last-dim pitch may be padded.
"""
import torch

from . import lib
from .lib import DENET_BF16, DENET_F32, call


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _dtype_code(t):
    if t.dtype == torch.float32:
        return DENET_F32
    if t.dtype == torch.bfloat16:
        return DENET_BF16
    raise TypeError("unsupported dtype %s" % t.dtype)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise lib.DenetError("denet_b200 ops need CUDA tensors (there is no CPU fallback)")


def _pitch(t):
    """pixel pitch (elements) of an NHWC tensor whose channel dim is contiguous."""
    assert t.stride(-1) == 1, "channel dimension must be contiguous"
    ld = t.stride(-2)
    n, h, w, _ = t.shape
    assert t.stride(1) == ld * w and (n == 1 or t.stride(0) == ld * w * h), "tensor is not pixel-contiguous NHWC"
    return ld


# ---------------------------------------------------------------------------------------------- conv operands
class ConvOperand:
    """GEMM B operand prepared from reference-layout filters (hi [+ lo] bf16)."""

    def __init__(self, hi, lo, rows, kin, R, S):
        self.hi, self.lo, self.rows, self.kin, self.R, self.S = hi, lo, rows, kin, R, S


def conv_weight_prep(w, mode, split):
    """w: (Cout, Cin, R, S) fp32 reference filters. mode 0 = fprop operand, 1 = dgrad operand."""
    _require_cuda(w)
    assert w.dtype == torch.float32 and w.is_contiguous()
    cout, cin, R, S = w.shape
    rows, kin = (cout, cin) if mode == 0 else (cin, cout)
    kp = (kin + 63) // 64 * 64
    hi = torch.empty((rows, R * S, kp), dtype=torch.bfloat16, device=w.device)
    lo = torch.empty_like(hi) if split else None
    call("denet_conv_weight_prep", w.data_ptr(), cout, cin, R, S, mode, hi.data_ptr(), _ptr(lo), _stream())
    return ConvOperand(hi, lo, rows, kin, R, S)


class ActOperand:
    """bf16 view(s) of an activation: hi only (throughput mode) or hi+lo (fp32 parity mode)."""

    def __init__(self, hi, lo=None):
        self.hi, self.lo = hi, lo

    @property
    def shape(self):
        return self.hi.shape


def alloc_nhwc(n, h, w, c, dtype, device="cuda", zero=False):
    """NHWC tensor whose pixel pitch is padded to a multiple of 8 elements (TMA needs 16-byte strides)."""
    ld = (c + 7) // 8 * 8
    buf = (torch.zeros if zero or ld != c else torch.empty)((n, h, w, ld), dtype=dtype, device=device)
    return buf[..., :c] if ld != c else buf


def _padded_base(x):
    """the (N, H, W, ld) buffer behind a channel-sliced NHWC view made by alloc_nhwc"""
    n, h, w, _ = x.shape
    ld = _pitch(x)
    return torch.as_strided(x, (n, h, w, ld), (h * w * ld, w * ld, ld, 1))


def act_operand(x):
    """NHWC activation -> ActOperand. bf16 tensors are used as-is; fp32 tensors are split into hi/lo."""
    _require_cuda(x)
    if x.dtype == torch.bfloat16:
        return ActOperand(x)
    assert x.dtype == torch.float32
    base = _padded_base(x)
    hi = torch.empty(base.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty_like(hi)
    call("denet_split_bf16", base.data_ptr(), hi.data_ptr(), lo.data_ptr(), base.numel(), _stream())
    c = x.shape[-1]
    return ActOperand(hi[..., :c], lo[..., :c])


def conv2d_fprop(xop, wop, pad_h, pad_w, out_hw, out_dtype, bias=None, residual=None, relu=False, stats=None,
                 out=None):
    """Stride-1 correlation with a prepared operand (see denet_conv2d_fprop). Returns NHWC (N, Ho, Wo, rows)."""
    x = xop.hi
    n, hi_, wi_, cin = x.shape
    assert cin == wop.kin, (cin, wop.kin)
    assert (xop.lo is None) == (wop.lo is None), "operand split modes differ"
    ho, wo = out_hw
    cout = wop.rows
    if out is None:
        out = torch.empty((n, ho, wo, cout), dtype=out_dtype, device=x.device)
    ldx = _pitch(x)
    ldy = _pitch(out)
    if residual is not None:
        assert residual.shape == out.shape and residual.dtype == out.dtype and _pitch(residual) == ldy
    # 1x1 convolutions are plain GEMMs over all pixels: flatten so that M tiles are 128 consecutive pixels
    if wop.R == 1 and wop.S == 1 and pad_h == 0 and pad_w == 0 and (ho, wo) == (hi_, wi_):
        n_, h_, w_ = 1, 1, n * hi_ * wi_
        ho_, wo_ = 1, w_
    else:
        n_, h_, w_ = n, hi_, wi_
        ho_, wo_ = ho, wo
    s0, s1 = (stats if stats is not None else (None, None))
    call("denet_conv2d_fprop", x.data_ptr(), _ptr(xop.lo), n_, h_, w_, cin, ldx,
         wop.hi.data_ptr(), _ptr(wop.lo), cout, wop.R, wop.S, pad_h, pad_w,
         out.data_ptr(), _dtype_code(out), ldy, ho_, wo_, _ptr(bias), _ptr(residual), int(relu),
         _ptr(s0), _ptr(s1), _stream())
    return out


_wgrad_ws = {}


def _workspace(nbytes, device):
    ws = _wgrad_ws.get(device)
    if ws is None or ws.numel() * 4 < nbytes:
        ws = torch.empty(((nbytes + 3) // 4,), dtype=torch.float32, device=device)
        _wgrad_ws[device] = ws
    return ws


def conv2d_wgrad(dyop, xop, R, S, pad_h, pad_w, dw=None, accumulate=False):
    """Filter gradient in the reference layout (Cout, Cin, R, S). dy: (N,Ho,Wo,Cout), x: (N,Hi,Wi,Cin)."""
    dy, x = dyop.hi, xop.hi
    n, ho, wo, cout = dy.shape
    n2, hi_, wi_, cin = x.shape
    assert n == n2
    assert (dyop.lo is None) == (xop.lo is None)
    if dw is None:
        dw = torch.empty((cout, cin, R, S), dtype=torch.float32, device=x.device)
        accumulate = False
    if R == 1 and S == 1 and pad_h == 0 and pad_w == 0 and (ho, wo) == (hi_, wi_):
        n_, ho_, wo_, hi2, wi2 = 1, 1, n * ho * wo, 1, n * ho * wo
    else:
        n_, ho_, wo_, hi2, wi2 = n, ho, wo, hi_, wi_
    nbytes = lib.load().denet_conv2d_wgrad_workspace(n_, ho_, wo_, cout, cin, R, S)
    ws = _workspace(nbytes, x.device)
    call("denet_conv2d_wgrad", dy.data_ptr(), _ptr(dyop.lo), n_, ho_, wo_, cout, _pitch(dy),
         x.data_ptr(), _ptr(xop.lo), hi2, wi2, cin, _pitch(x), R, S, pad_h, pad_w,
         dw.data_ptr(), int(accumulate), ws.data_ptr(), ws.numel() * 4, _stream())
    return dw
