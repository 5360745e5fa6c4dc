"""Tensor-level wrappers over the C-ABI (include/denet_b200.h).

torch is used for device memory and streams only; every op here is one or more calls into libdenet_b200.so
on the current CUDA stream.  Activations are NHWC tensors (N, H, W, C) whose last-dim pitch may be padded to a
multiple of 8 elements (TMA needs 16-byte strides).  Nothing here falls back to torch arithmetic.
"""
import numpy
import torch

from . import lib
from .lib import DENET_BF16, DENET_F32, call


_stream_handle = None


def _stream():
    """raw handle of the stream the kernels are enqueued on.  torch.cuda.current_stream() costs ~15 us of host time
    per call, so ModelCNN pins it once per step with pin_stream(); unpinned callers pay the lookup."""
    if _stream_handle is not None:
        return _stream_handle
    return torch.cuda.current_stream().cuda_stream


def pin_stream(on=True):
    """cache (or release) the current stream handle for subsequent ops"""
    global _stream_handle
    _stream_handle = torch.cuda.current_stream().cuda_stream if on else None


class on_stream:
    """context manager: the ops inside are enqueued on `stream` (a torch.cuda.Stream) instead of the pinned one"""

    def __init__(self, stream):
        self.handle = stream.cuda_stream

    def __enter__(self):
        global _stream_handle
        self.prev = _stream_handle
        _stream_handle = self.handle

    def __exit__(self, *exc):
        global _stream_handle
        _stream_handle = self.prev


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
    assert (w == 1 or t.stride(2) == ld) and (h == 1 or t.stride(1) == ld * w) and \
        (n == 1 or t.stride(0) == ld * w * h), "tensor is not pixel-contiguous NHWC"
    return ld


def _rows(t):
    n, h, w, _ = t.shape
    return n * h * w


def alloc_nhwc(n, h, w, c, dtype, device="cuda", zero=False):
    """NHWC tensor whose pixel pitch is padded to a multiple of 8 elements."""
    ld = (c + 7) // 8 * 8
    if zero:
        buf = torch.zeros((n, h, w, ld), dtype=dtype, device=device)
    else:
        buf = torch.empty((n, h, w, ld), dtype=dtype, device=device)
        if ld != c:
            buf[..., c:].zero_()       # only the pad columns (the 4706-channel RoI tensor has 6 of them, not 174 MB)
    return buf[..., :c] if ld != c else buf


def alloc_like(x, dtype=None):
    n, h, w, c = x.shape
    return alloc_nhwc(n, h, w, c, dtype or x.dtype, x.device)


def _padded_base(x):
    """the (N, H, W, ld) buffer behind a channel-sliced NHWC view made by alloc_nhwc"""
    n, h, w, _ = x.shape
    ld = _pitch(x)
    return torch.as_strided(x, (n, h, w, ld), (h * w * ld, w * ld, ld, 1), x.storage_offset())


# ---------------------------------------------------------------------------------------------- workspaces
_ws = {}


def workspace(nbytes, device, tag="ws"):
    key = (tag, str(device))
    ws = _ws.get(key)
    if ws is None or ws.numel() * 4 < nbytes:
        ws = torch.empty((max(1024, (nbytes + 3) // 4),), dtype=torch.float32, device=device)
        _ws[key] = ws
    return ws


# ---------------------------------------------------------------------------------------------- conv operands
class ConvOperand:
    """GEMM B operand prepared from reference-layout filters (hi [+ lo] bf16)."""

    def __init__(self, hi, lo, rows, kin, R, S):
        self.hi, self.lo, self.rows, self.kin, self.R, self.S = hi, lo, rows, kin, R, S


def conv_weight_prep(w, mode, split, out=None):
    """w: (Cout, Cin, R, S) fp32 reference filters. mode 0 = fprop operand, 1 = dgrad operand."""
    _require_cuda(w)
    assert w.dtype == torch.float32 and w.is_contiguous()
    cout, cin, R, S = w.shape
    rows, kin = (cout, cin) if mode == 0 else (cin, cout)
    kp = (kin + 63) // 64 * 64
    if out is None:
        hi = torch.empty((rows, R * S, kp), dtype=torch.bfloat16, device=w.device)
        lo = torch.empty_like(hi) if split else None
        out = ConvOperand(hi, lo, rows, kin, R, S)
    call("denet_conv_weight_prep", w.data_ptr(), cout, cin, R, S, mode, out.hi.data_ptr(), _ptr(out.lo), _stream())
    return out


def dgrad_parity_classes(in_hw, size, stride, pad):
    """Data gradient of a strided convolution as one stride-1 correlation per parity class of the input pixel.
    With dx[h] = sum_r dyd[h + r - (R-1-pad)] W[r] on the zero-dilated dyd, only the taps r = r0 + s*t with
    r0 = (R-1-pad - a) mod s reach the pixels h = s*i + a, reading dy[i + o] with o = (a + r - (R-1-pad)) / s.
    Returns [(a, b, r0, s0, Rc, Sc, pad_h, pad_w, Hc, Wc)] or None when a class has no tap (its pixels are zero)."""
    out = []
    per_dim = []
    for d in range(2):
        n, k, st, p = in_hw[d], size[d], stride[d], pad[d]
        dim = []
        for a in range(st):
            r0 = (k - 1 - p - a) % st
            cnt = len(range(r0, k, st))
            if cnt == 0:
                return None
            o_min = (a + r0 - (k - 1 - p)) // st
            assert (a + r0 - (k - 1 - p)) % st == 0
            dim.append((a, r0, cnt, -o_min, (n - a + st - 1) // st))
        per_dim.append(dim)
    for (a, r0, rc, ph, hc) in per_dim[0]:
        for (b, s0, sc, pw, wc) in per_dim[1]:
            out.append((a, b, r0, s0, rc, sc, ph, pw, hc, wc))
    return out


def parity_class_code(cls, stride):
    a, b, r0, s0, rc, sc = cls[:6]
    return r0 | (s0 << 4) | (stride[0] << 8) | (stride[1] << 12) | (rc << 16) | (sc << 20)


def conv_weight_prep_records(records):
    """denet_conv_weight_prep_multi for a list of (w, operand, mode, cp) records (table built on the fly: the lazy,
    per-layer path; ModelCNN.prepare_operands keeps one cached table for the whole model)"""
    L = lib.load()
    chunk = L.denet_weight_prep_chunk()
    assert L.denet_weight_prep_entry_bytes() == 56
    table = numpy.zeros((len(records), 7), dtype=numpy.int64)
    ints = table.view(numpy.int32).reshape(len(records), 14)
    block_entry, block_offset = [], []
    for i, (w, op, mode, cp) in enumerate(records):
        cout, cin, R, S = w.shape
        items = op.hi.numel() // op.hi.shape[1]
        table[i, 0] = w.data_ptr()
        table[i, 1] = op.hi.data_ptr()
        table[i, 2] = op.lo.data_ptr() if op.lo is not None else 0
        table[i, 3] = items
        ints[i, 8:14] = [cout, cin, R, S, mode, cp]
        for o in range(0, items, chunk):
            block_entry.append(i)
            block_offset.append(o)
    dev = records[0][0].device
    t = (torch.from_numpy(table).to(dev), torch.tensor(block_entry, dtype=torch.int32, device=dev),
         torch.tensor(block_offset, dtype=torch.int64, device=dev))
    call("denet_conv_weight_prep_multi", t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), len(block_entry), _stream())
    return t


class ActOperand:
    """bf16 view(s) of an activation: hi only (throughput mode) or hi+lo (fp32 parity mode)."""

    def __init__(self, hi, lo=None):
        self.hi, self.lo = hi, lo

    @property
    def shape(self):
        return self.hi.shape


def act_operand(x):
    """NHWC activation -> ActOperand. bf16 tensors are used as-is; fp32 tensors are split into hi/lo."""
    _require_cuda(x)
    if isinstance(x, ActOperand):
        return x
    if x.dtype == torch.bfloat16:
        return ActOperand(x)
    assert x.dtype == torch.float32
    base = _padded_base(x)
    hi = torch.empty(base.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty_like(hi)
    call("denet_split_bf16", base.data_ptr(), hi.data_ptr(), lo.data_ptr(), base.numel(), _stream())
    c = x.shape[-1]
    return ActOperand(hi[..., :c], lo[..., :c])


class BnBwdFuse:
    """arguments of the batch-norm backward statistics fused into a dgrad epilogue (denet_conv2d_dgrad_bnbwd): the
    batch-norm layer's saved input x, forward output yout (or None: mask recomputed from x), batch mean / invstd, gamma,
    beta, relu flag and the two zeroed per-channel accumulators"""

    def __init__(self, x, yout, mean, invstd, gamma, beta, relu, sum_dz, sum_dz_xhat):
        self.x, self.yout, self.mean, self.invstd, self.gamma, self.beta = x, yout, mean, invstd, gamma, beta
        self.relu, self.sum_dz, self.sum_dz_xhat = relu, sum_dz, sum_dz_xhat


def conv2d_fprop(xop, wop, pad, out_hw, out_dtype, stride=(1, 1), bias=None, residual=None, relu=False, stats=None,
                 out=None, bn_bwd=None, scatter=None):
    """Correlation with a prepared operand (see denet_conv2d_fprop). Returns NHWC (N, Ho, Wo, rows).
    bn_bwd (BnBwdFuse): this call is the dgrad feeding a batch-norm layer's backward; its epilogue masks the gradient
    and accumulates that layer's two backward sums (denet_conv2d_dgrad_bnbwd)."""
    x = xop.hi
    n, hi_, wi_, cin = x.shape
    assert cin == wop.kin, (cin, wop.kin)
    assert (xop.lo is None) == (wop.lo is None), "operand split modes differ"
    ho, wo = out_hw
    cout = wop.rows
    if scatter is not None:
        # output pixel (h, w) of this launch is pixel (h*osh + ooh, w*osw + oow) of `out` (N, Hf, Wf, cout)
        hf, wf, osh, osw, ooh, oow = scatter
        assert out is not None and tuple(out.shape) == (n, hf, wf, cout) and tuple(stride) == (1, 1)
        assert bias is None and not relu and stats is None and bn_bwd is None
        assert residual is None or (residual.shape == out.shape and residual.dtype == out.dtype and
                                    _pitch(residual) == _pitch(out))
        call("denet_conv2d_fprop_scatter", x.data_ptr(), _ptr(xop.lo), n, hi_, wi_, cin, _pitch(x),
             wop.hi.data_ptr(), _ptr(wop.lo), cout, wop.R, wop.S, pad[0], pad[1], out.data_ptr(), _dtype_code(out),
             _pitch(out), ho, wo, hf, wf, osh, osw, ooh, oow, _ptr(residual), _stream())
        return out
    if out is None:
        out = alloc_nhwc(n, ho, wo, cout, out_dtype, x.device)
    ldx = _pitch(x)
    ldy = _pitch(out)
    if residual is not None:
        assert residual.shape == out.shape and residual.dtype == out.dtype and _pitch(residual) == ldy
    # 1x1 stride-1 convolutions are plain GEMMs over all pixels: flatten so that M tiles are 128 consecutive pixels
    if wop.R == 1 and wop.S == 1 and tuple(pad) == (0, 0) and tuple(stride) == (1, 1) and (ho, wo) == (hi_, wi_):
        n_, h_, w_ = 1, 1, n * hi_ * wi_
        ho_, wo_ = 1, w_
    else:
        n_, h_, w_ = n, hi_, wi_
        ho_, wo_ = ho, wo
    if bn_bwd is not None:
        assert tuple(stride) == (1, 1) and bias is None and not relu and stats is None
        b = bn_bwd
        assert b.x.shape == out.shape and b.x.dtype == out.dtype and _pitch(b.x) == ldy
        assert b.yout is None or (b.yout.shape == out.shape and b.yout.dtype == out.dtype and _pitch(b.yout) == ldy)
        call("denet_conv2d_dgrad_bnbwd", x.data_ptr(), _ptr(xop.lo), n_, h_, w_, cin, ldx,
             wop.hi.data_ptr(), _ptr(wop.lo), cout, wop.R, wop.S, pad[0], pad[1], out.data_ptr(), _dtype_code(out), ldy,
             ho_, wo_, _ptr(residual), b.x.data_ptr(), _ptr(b.yout), b.mean.data_ptr(), b.invstd.data_ptr(),
             b.gamma.data_ptr(), _ptr(b.beta), int(b.relu), b.sum_dz.data_ptr(), b.sum_dz_xhat.data_ptr(), _stream())
        return out
    s0, s1 = (stats if stats is not None else (None, None))
    call("denet_conv2d_fprop", x.data_ptr(), _ptr(xop.lo), n_, h_, w_, cin, ldx,
         wop.hi.data_ptr(), _ptr(wop.lo), cout, wop.R, wop.S, pad[0], pad[1], stride[0], stride[1],
         out.data_ptr(), _dtype_code(out), ldy, ho_, wo_, _ptr(bias), _ptr(residual), int(relu),
         _ptr(s0), _ptr(s1), _stream())
    return out


class WgradPending:
    """split-K partial sums of one filter gradient waiting for the multi-tensor reduction (denet_wgrad_reduce_multi)"""

    def __init__(self, ws, dw, splits, cout, cin, R, S, ldws, mode=0, cp=0, accumulate=False):
        self.ws, self.dw, self.splits, self.cout, self.cin, self.R, self.S = ws, dw, splits, cout, cin, R, S
        self.ldws, self.mode, self.cp, self.accumulate = ldws, mode, cp, accumulate

    def key(self):
        return (self.ws.data_ptr(), self.dw.data_ptr(), self.splits, self.cout, self.cin, self.R, self.S, self.ldws,
                self.mode, self.cp, int(self.accumulate))


def _own_workspace(owner, nbytes, device):
    """per-layer persistent workspace (deferred reductions must not share one scratch buffer)"""
    ws = getattr(owner, "_wgrad_ws", None)
    if ws is None or ws.numel() * 4 < nbytes or ws.device != device:
        ws = torch.empty((max(1024, (nbytes + 3) // 4),), dtype=torch.float32, device=device)
        owner._wgrad_ws = ws
    return ws


def conv2d_wgrad(dyop, xop, R, S, pad, stride=(1, 1), dw=None, accumulate=False, defer=None):
    """Filter gradient in the reference layout (Cout, Cin, R, S). dy: (N,Ho,Wo,Cout), x: (N,Hi,Wi,Cin).
    defer = (pending list, owner): leave the split-K partials in the owner's private workspace and append a
    WgradPending record instead of reducing now (ModelCNN reduces all layers in one launch)."""
    dy, x = dyop.hi, xop.hi
    n, ho, wo, cout = dy.shape
    n2, hi_, wi_, cin = x.shape
    assert n == n2
    assert (dyop.lo is None) == (xop.lo is None)
    if dw is None:
        dw = torch.empty((cout, cin, R, S), dtype=torch.float32, device=x.device)
        accumulate = False
    if R == 1 and S == 1 and tuple(pad) == (0, 0) and tuple(stride) == (1, 1) and (ho, wo) == (hi_, wi_):
        n_, ho_, wo_, hi2, wi2 = 1, 1, n * ho * wo, 1, n * ho * wo
    else:
        n_, ho_, wo_, hi2, wi2 = n, ho, wo, hi_, wi_
    L = lib.load()
    nbytes = L.denet_conv2d_wgrad_workspace(n_, ho_, wo_, cout, cin, R, S)
    if defer is not None:
        pending, owner = defer
        ws = _own_workspace(owner, nbytes, x.device)
        splits = L.denet_conv2d_wgrad_splits(n_, ho_, wo_, cout, cin, R, S, stride[0], stride[1])
        pending.append(WgradPending(ws, dw, splits, cout, cin, R, S, (cin + 3) // 4 * 4, 0, 0, accumulate))
        dw_ptr = None
    else:
        ws = workspace(nbytes, x.device, "wgrad")
        dw_ptr = dw.data_ptr()
    if defer is not None and accumulate and len(defer[0]) > 1:
        wgrad_reduce_pending(defer[0][:-1])      # a launch must not hold two writers of one gradient tensor
        del defer[0][:-1]
    call("denet_conv2d_wgrad", dy.data_ptr(), _ptr(dyop.lo), n_, ho_, wo_, cout, _pitch(dy),
         x.data_ptr(), _ptr(xop.lo), hi2, wi2, cin, _pitch(x), R, S, pad[0], pad[1], stride[0], stride[1],
         dw_ptr, int(accumulate), ws.data_ptr(), ws.numel() * 4, _stream())
    if defer is not None and accumulate:
        wgrad_reduce_pending(defer[0])
    return dw


_reduce_tables = {}


def wgrad_reduce_pending(pending):
    """one launch reducing every WgradPending record of the list (fixed summation order, reference filter layout);
    the device table of a given list of records is built once and reused (the step structure is static)"""
    if not pending:
        return
    key = tuple(e.key() for e in pending)
    tab = _reduce_tables.get(key)
    if tab is None:
        L = lib.load()
        assert L.denet_wgrad_reduce_entry_bytes() == 64
        chunk = L.denet_wgrad_reduce_chunk()
        items = L.denet_wgrad_reduce_items()
        group = L.denet_wgrad_reduce_group()
        dev = pending[0].ws.device
        table = numpy.zeros((len(pending), 8), dtype=numpy.int64)
        ints = table.view(numpy.int32).reshape(len(pending), 16)
        block_entry, block_offset = [], []
        for i, e in enumerate(pending):
            total = e.cout * e.cin * e.R * e.S
            table[i, 0] = e.ws.data_ptr()
            table[i, 1] = e.dw.data_ptr()
            table[i, 2] = total
            ints[i, 6:16] = [e.splits, e.cout, e.cin, e.R, e.S, e.ldws, e.mode, e.cp, int(e.accumulate), 0]
            if e.mode == 0 and e.R * e.S > 1:      # tiled path: units of (co, channel group)
                units, per_block = e.cout * ((e.cin + group - 1) // group), items
            else:
                units, per_block = total, chunk
            for o in range(0, units, per_block):
                block_entry.append(i)
                block_offset.append(o)
        tab = (torch.from_numpy(table).to(dev), torch.tensor(block_entry, dtype=torch.int32, device=dev),
               torch.tensor(block_offset, dtype=torch.int64, device=dev), len(block_entry))
        if len(_reduce_tables) > 256:
            _reduce_tables.clear()
        _reduce_tables[key] = tab
    entries, be, bo, nblocks = tab
    call("denet_wgrad_reduce_multi", entries.data_ptr(), be.data_ptr(), bo.data_ptr(), nblocks, _stream())
    del pending[:]


# ---------------------------------------------------------------------------------------------- row-folded stem conv
class PaddedImage:
    """zero-padded NHWC-Cp bf16 image batch (hi [+ lo]) read by the row-folded convolution through overlapping TMA
    windows; `shape` is the logical NCHW-style (N, H, W, C) of the unpadded image"""

    def __init__(self, n, c, h, w, cp, pad, hp, wp, split, device):
        self.n, self.c, self.h, self.w, self.cp, self.pad, self.hp, self.wp = n, c, h, w, cp, pad, hp, wp
        self.hi = torch.zeros((n, hp, wp, cp), dtype=torch.bfloat16, device=device)
        self.lo = torch.zeros_like(self.hi) if split else None
        self.shape = (n, h, w, c)
        self.device = device

    def fill(self, x_nchw):
        """x_nchw: (N,C,H,W) fp32 contiguous device tensor; only the interior is written, the border stays zero"""
        assert x_nchw.dtype == torch.float32 and x_nchw.is_contiguous() and tuple(x_nchw.shape) == \
            (self.n, self.c, self.h, self.w)
        call("denet_nchw_to_padded_nhwc", x_nchw.data_ptr(), self.n, self.c, self.h, self.w, self.cp, self.pad[0],
             self.pad[1], self.hp, self.wp, self.hi.data_ptr(), _ptr(self.lo), _stream())
        return self


def rowfold_geometry(in_hw, cin, size, stride, pad, out_hw):
    """(Cp, Hp, Wp) of the padded buffer a row-folded conv needs, or None when the shape does not qualify"""
    R, S = size
    for cp in (4, 8):
        if cin <= cp and (stride[1] * cp) % 8 == 0 and (S * cp + 7) // 8 * 8 <= 64:
            kf = (S * cp + 7) // 8 * 8
            hp = max(in_hw[0] + 2 * pad[0], (out_hw[0] - 1) * stride[0] + R)
            wp = max(in_hw[1] + 2 * pad[1], -(-((out_hw[1] - 1) * stride[1] * cp + kf) // cp))
            while (wp * cp) % 8:
                wp += 1
            return cp, hp, wp
    return None


def conv_weight_prep_rowfold(w, cp, split, out=None):
    cout, cin, R, S = w.shape
    if out is None:
        hi = torch.empty((cout, R, 64), dtype=torch.bfloat16, device=w.device)
        out = ConvOperand(hi, torch.empty_like(hi) if split else None, cout, cin, R, S)
    call("denet_conv_weight_prep_rowfold", w.data_ptr(), cout, cin, R, S, cp, out.hi.data_ptr(), _ptr(out.lo),
         _stream())
    return out


def conv2d_rowfold_fprop(img, wop, stride, out_hw, out_dtype, bias=None, relu=False, stats=None):
    assert (img.lo is None) == (wop.lo is None), "operand split modes differ"
    ho, wo = out_hw
    out = alloc_nhwc(img.n, ho, wo, wop.rows, out_dtype, img.device)
    s0, s1 = (stats if stats is not None else (None, None))
    call("denet_conv2d_rowfold_fprop", img.hi.data_ptr(), _ptr(img.lo), img.n, img.hp, img.wp, img.cp, img.c,
         wop.hi.data_ptr(), _ptr(wop.lo), wop.rows, wop.R, wop.S, stride[0], stride[1], out.data_ptr(),
         _dtype_code(out), _pitch(out), ho, wo, _ptr(bias), int(relu), _ptr(s0), _ptr(s1), _stream())
    return out


def conv2d_rowfold_wgrad(dyop, img, R, S, stride, dw, accumulate=False, defer=None):
    dy = dyop.hi
    n, ho, wo, cout = dy.shape
    assert (dyop.lo is None) == (img.lo is None)
    L = lib.load()
    nbytes = L.denet_conv2d_rowfold_wgrad_workspace(n, ho, wo, cout, R, stride[0])
    if defer is not None:
        pending, owner = defer
        ws = _own_workspace(owner, nbytes, dy.device)
        splits = L.denet_conv2d_rowfold_wgrad_splits(n, ho, wo, cout, R, stride[0])
        pending.append(WgradPending(ws, dw, splits, cout, img.c, R, S, 64, 1, img.cp, accumulate))
        dw_ptr = None
    else:
        ws = workspace(nbytes, dy.device, "wgrad")
        dw_ptr = dw.data_ptr()
    call("denet_conv2d_rowfold_wgrad", dy.data_ptr(), _ptr(dyop.lo), n, ho, wo, cout, _pitch(dy), img.hi.data_ptr(),
         _ptr(img.lo), img.hp, img.wp, img.cp, img.c, R, S, stride[0], stride[1], dw_ptr, int(accumulate),
         ws.data_ptr(), ws.numel() * 4, _stream())
    return dw


def dilate(x, stride, out_hw, add=None):
    """zero-insertion upsampling (strided dgrad helper): y[:, h*sh, w*sw] = x[:, h, w] (+ add, a tensor like y)"""
    n, h, w, c = x.shape
    y = alloc_nhwc(n, out_hw[0], out_hw[1], c, x.dtype, x.device)
    assert add is None or (add.shape == y.shape and add.dtype == y.dtype and _pitch(add) == _pitch(y))
    call("denet_dilate_add", x.data_ptr(), _dtype_code(x), n, h, w, c, _pitch(x), stride[0], stride[1], _ptr(add),
         y.data_ptr(), out_hw[0], out_hw[1], _pitch(y), _stream())
    return y


def im2col(x, R, S, stride, pad, out_hw):
    """NHWC -> column matrix (N, Ho, Wo, R*S*C) (pitch padded to 8, tail zeroed)"""
    n, h, w, c = x.shape
    ho, wo = out_hw
    k = R * S * c
    col = alloc_nhwc(n, ho, wo, k, x.dtype, x.device)
    call("denet_im2col", x.data_ptr(), _dtype_code(x), n, h, w, c, _pitch(x), R, S, stride[0], stride[1], pad[0],
         pad[1], ho, wo, col.data_ptr(), _pitch(col), _stream())
    return col


def col2im(dcol, x_shape, R, S, stride, pad):
    n, h, w, c = x_shape
    _, ho, wo, _ = dcol.shape
    dx = alloc_nhwc(n, h, w, c, dcol.dtype, dcol.device)
    call("denet_col2im", dcol.data_ptr(), _dtype_code(dcol), _pitch(dcol), n, h, w, c, _pitch(dx), R, S, stride[0],
         stride[1], pad[0], pad[1], ho, wo, dx.data_ptr(), _stream())
    return dx


def weight_to_im2col(w, out=None):
    cout, cin, R, S = w.shape
    if out is None:
        out = torch.empty((cout, R * S * cin, 1, 1), dtype=torch.float32, device=w.device)
    call("denet_weight_to_im2col", w.data_ptr(), cout, cin, R, S, out.data_ptr(), _stream())
    return out


def weight_grad_from_im2col(dw2, dw, accumulate=False):
    cout, cin, R, S = dw.shape
    call("denet_weight_grad_from_im2col", dw2.data_ptr(), cout, cin, R, S, dw.data_ptr(), int(accumulate), _stream())
    return dw


# ---------------------------------------------------------------------------------------------- batch norm
def _bn_ws(M, C, device):
    nbytes = lib.load().denet_bn_workspace_bytes(M, C)
    return workspace(nbytes, device, "bn")


def bn_stats(x, eps, mean, invstd, run_mean=None, run_stdinv=None, momentum=0.9):
    M, C = _rows(x), x.shape[-1]
    ws = _bn_ws(M, C, x.device)
    call("denet_bn_stats", x.data_ptr(), _dtype_code(x), M, C, _pitch(x), eps, mean.data_ptr(), invstd.data_ptr(),
         _ptr(run_mean), _ptr(run_stdinv), momentum, ws.data_ptr(), ws.numel() * 4, _stream())


def bn_finalize_sums(sums, sqsums, M, eps, mean, invstd, run_mean=None, run_stdinv=None, momentum=0.9):
    call("denet_bn_finalize_sums", sums.data_ptr(), sqsums.data_ptr(), M, sums.numel(), eps, mean.data_ptr(),
         invstd.data_ptr(), _ptr(run_mean), _ptr(run_stdinv), momentum, _stream())


def bn_apply(x, mean, invstd, gamma, beta, residual=None, relu=False, out=None):
    if out is None:
        out = alloc_like(x)
    assert _pitch(out) == _pitch(x) and (residual is None or _pitch(residual) == _pitch(x))
    call("denet_bn_apply", x.data_ptr(), _dtype_code(x), _rows(x), x.shape[-1], _pitch(x), mean.data_ptr(),
         invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(residual), int(relu), out.data_ptr(), _stream())
    return out


def bn_apply_sums(x, sums, sqsums, eps, gamma, beta, mean, invstd, run_mean=None, run_stdinv=None, momentum=0.9,
                  residual=None, relu=False, out=None):
    """bn_apply fed with the conv epilogue's per-channel sums: finalises the statistics inside the apply launch"""
    if out is None:
        out = alloc_like(x)
    assert _pitch(out) == _pitch(x) and (residual is None or _pitch(residual) == _pitch(x))
    call("denet_bn_apply_sums", x.data_ptr(), _dtype_code(x), _rows(x), x.shape[-1], _pitch(x), sums.data_ptr(),
         sqsums.data_ptr(), eps, gamma.data_ptr(), beta.data_ptr(), _ptr(residual), int(relu), out.data_ptr(),
         mean.data_ptr(), invstd.data_ptr(), _ptr(run_mean), _ptr(run_stdinv), momentum, _stream())
    return out


def bn_inference_invstd(run_stdinv, eps):
    out = torch.empty_like(run_stdinv)
    call("denet_bn_inference_invstd", run_stdinv.data_ptr(), eps, out.data_ptr(), run_stdinv.numel(), _stream())
    return out


def bn_backward(dy, yout, x, mean, invstd, gamma, relu, dgamma, dbeta, accumulate=False, want_dres=False, dx=None,
                beta=None):
    """yout=None with relu: the mask is recomputed from x (needs beta; only valid when the forward had no residual)"""
    M, C = _rows(x), x.shape[-1]
    ws = _bn_ws(M, C, x.device)
    if dx is None:
        dx = alloc_like(x)
    dres = alloc_like(x) if want_dres else None
    ld = _pitch(x)
    assert _pitch(dy) == ld and _pitch(dx) == ld and (yout is None or _pitch(yout) == ld)
    call("denet_bn_backward", dy.data_ptr(), _ptr(yout), x.data_ptr(), _dtype_code(x), M, C, ld, mean.data_ptr(),
         invstd.data_ptr(), gamma.data_ptr(), _ptr(beta), int(relu), dx.data_ptr(), _ptr(dres), _ptr(dgamma),
         _ptr(dbeta),
         int(accumulate), ws.data_ptr(), ws.numel() * 4, _stream())
    return dx, dres


def bn_backward_sums(dy, x, mean, invstd, gamma, sum_dy, sum_dy_xhat, dgamma, dbeta, accumulate=False, dx=None):
    """second half of the batch-norm backward from sums the dgrad epilogue accumulated; dy is already masked"""
    M, C = _rows(x), x.shape[-1]
    if dx is None:
        dx = alloc_like(x)
    ld = _pitch(x)
    assert _pitch(dy) == ld and _pitch(dx) == ld
    call("denet_bn_backward_sums", dy.data_ptr(), x.data_ptr(), _dtype_code(x), M, C, ld, mean.data_ptr(),
         invstd.data_ptr(), gamma.data_ptr(), sum_dy.data_ptr(), sum_dy_xhat.data_ptr(), dx.data_ptr(), _ptr(dgamma),
         _ptr(dbeta), int(accumulate), _stream())
    return dx


# ---------------------------------------------------------------------------------------------- elementwise
def relu_fwd(x, out=None):
    if out is None:
        out = alloc_like(x)
    call("denet_relu_fwd", x.data_ptr(), _dtype_code(x), _rows(x), x.shape[-1], _pitch(x), out.data_ptr(), _stream())
    return out


def relu_bwd(dy, y, out=None):
    if out is None:
        out = alloc_like(dy)
    assert _pitch(dy) == _pitch(y) == _pitch(out)
    call("denet_relu_bwd", dy.data_ptr(), y.data_ptr(), _dtype_code(dy), _rows(dy), dy.shape[-1], _pitch(dy),
         out.data_ptr(), _stream())
    return out


def add(a, b, relu=False, out=None):
    if out is None:
        out = alloc_like(a)
    assert a.shape == b.shape and _pitch(a) == _pitch(b) == _pitch(out) and a.dtype == b.dtype
    call("denet_add", a.data_ptr(), b.data_ptr(), _dtype_code(a), _rows(a), a.shape[-1], _pitch(a), int(relu),
         out.data_ptr(), _stream())
    return out


def colsum(x, out, accumulate=False):
    M, C = _rows(x), x.shape[-1]
    ws = _bn_ws(M, C, x.device)
    call("denet_colsum", x.data_ptr(), _dtype_code(x), M, C, _pitch(x), out.data_ptr(), int(accumulate),
         ws.data_ptr(), ws.numel() * 4, _stream())
    return out


def convert(x, dtype, out=None):
    if out is None:
        out = alloc_like(x, dtype)
    call("denet_convert", x.data_ptr(), _dtype_code(x), _rows(x), x.shape[-1], _pitch(x), out.data_ptr(),
         _dtype_code(out), _pitch(out), _stream())
    return out


def nchw_to_nhwc(x, dtype):
    """x: (N,C,H,W) fp32 contiguous device tensor -> NHWC activation"""
    assert x.dtype == torch.float32 and x.is_contiguous()
    n, c, h, w = x.shape
    y = alloc_nhwc(n, h, w, c, dtype, x.device, zero=True)
    call("denet_nchw_to_nhwc", x.data_ptr(), n, c, h, w, y.data_ptr(), _dtype_code(y), _pitch(y), _stream())
    return y


def nhwc_to_nchw(x):
    n, h, w, c = x.shape
    y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    call("denet_nhwc_to_nchw", x.data_ptr(), _dtype_code(x), _pitch(x), n, c, h, w, y.data_ptr(), _stream())
    return y


# ---------------------------------------------------------------------------------------------- pooling
def pool_fwd(x, mode, size, stride, pad, out_hw):
    n, h, w, c = x.shape
    ho, wo = out_hw
    y = alloc_nhwc(n, ho, wo, c, x.dtype, x.device)
    argmax = torch.empty((n * ho * wo * c,), dtype=torch.uint8, device=x.device) if mode == 0 else None
    call("denet_pool_fwd", x.data_ptr(), _dtype_code(x), n, h, w, c, _pitch(x), mode, size[0], size[1], stride[0],
         stride[1], pad[0], pad[1], y.data_ptr(), ho, wo, _pitch(y), _ptr(argmax), _stream())
    return y, argmax


def pool_bwd(dy, mode, size, stride, pad, x_shape, argmax):
    n, h, w, c = x_shape
    _, ho, wo, _ = dy.shape
    dx = alloc_nhwc(n, h, w, c, dy.dtype, dy.device)
    call("denet_pool_bwd", dy.data_ptr(), _dtype_code(dy), n, h, w, c, _pitch(dx), mode, size[0], size[1], stride[0],
         stride[1], pad[0], pad[1], ho, wo, _pitch(dy), _ptr(argmax), dx.data_ptr(), _stream())
    return dx


def pool_inv_fwd(x, size):
    """size = (size_w, size_h) as in the reference PoolInvLayer"""
    n, h, w, c = x.shape
    y = alloc_nhwc(n, h * size[1], w * size[0], c, x.dtype, x.device)
    call("denet_pool_inv_fwd", x.data_ptr(), _dtype_code(x), n, h, w, c, _pitch(x), size[0], size[1], y.data_ptr(),
         _pitch(y), _stream())
    return y


def pool_inv_bwd(dy, size):
    n, rh, rw, c = dy.shape
    h, w = rh // size[1], rw // size[0]
    dx = alloc_nhwc(n, h, w, c, dy.dtype, dy.device)
    call("denet_pool_inv_bwd", dy.data_ptr(), _dtype_code(dy), n, h, w, c, _pitch(dx), size[0], size[1],
         dx.data_ptr(), _pitch(dy), _stream())
    return dx


# ---------------------------------------------------------------------------------------------- DSS head
def sparse_sample_fwd(fmap, bbox, gs, out=None, out_dtype=None):
    """fmap (B,H,W,F), bbox (B,sn,sn,4) fp32 device -> (B, sn, sn, gs*gs*F+2)"""
    b, h, w, f = fmap.shape
    sn = bbox.shape[1]
    assert bbox.dtype == torch.float32 and bbox.is_contiguous()
    oc = gs * gs * f + 2
    if out is None:
        out = alloc_nhwc(b, sn, sn, oc, out_dtype or fmap.dtype, fmap.device)
    call("denet_sparse_sample_fwd", fmap.data_ptr(), _dtype_code(fmap), b, h, w, f, _pitch(fmap), bbox.data_ptr(),
         sn * sn, gs, out.data_ptr(), _dtype_code(out), _pitch(out), _stream())
    return out


def sparse_sample_bwd(dy, bbox, gs, fmap_shape):
    """dy (B,sn,sn,gs*gs*F+2) -> dfmap (B,H,W,F) fp32 (dense)"""
    b, h, w, f = fmap_shape
    sn = bbox.shape[1]
    dfmap = torch.empty((b, h, w, f), dtype=torch.float32, device=dy.device)
    call("denet_sparse_sample_bwd", dy.data_ptr(), _dtype_code(dy), _pitch(dy), bbox.data_ptr(), b, h, w, f, sn * sn,
         gs, dfmap.data_ptr(), _stream())
    return dfmap


def sparse_sample_index(bbox, gs, H, W):
    nroi = bbox.numel() // 4
    ys = torch.empty((nroi, gs), dtype=torch.int32, device=bbox.device)
    xs = torch.empty_like(ys)
    call("denet_sparse_sample_index", bbox.data_ptr(), nroi, gs, H, W, ys.data_ptr(), xs.data_ptr(), _stream())
    return ys, xs


def build_samples(corner_pr, corner_threshold, sample_num, max_corners=1024, local_max=0, cluster_threshold=1.0,
                  out=None):
    """corner_pr (B,2,4|5,H,W) fp32 device (5 = with the centre map of DNC.C).
    Returns (pr (B,K), bbox (B,K,4), ibox (B,K,4) int32, count (B), ncand (B))"""
    assert corner_pr.dtype == torch.float32 and corner_pr.is_contiguous() and corner_pr.shape[1] == 2 and \
        corner_pr.shape[2] in (4, 5)
    b, _, cn, h, w = corner_pr.shape
    k = sample_num * sample_num
    dev = corner_pr.device
    if out is not None:
        pr, bbox, count = out            # caller-owned (e.g. views of one packed buffer): pr (b,k) f32, bbox (b,k,4) f32, count (b) i32
        assert pr.is_contiguous() and bbox.is_contiguous() and count.is_contiguous() and count.dtype == torch.int32
    else:
        pr = torch.empty((b, k), dtype=torch.float32, device=dev)
        bbox = torch.empty((b, k, 4), dtype=torch.float32, device=dev)
        count = torch.empty((b,), dtype=torch.int32, device=dev)
    ibox = torch.empty((b, k, 4), dtype=torch.int32, device=dev)
    ncand = torch.empty((b,), dtype=torch.int32, device=dev)
    if cluster_threshold < 1.0:
        nbytes = lib.load().denet_build_samples_cluster_workspace(b, h, w, max_corners, sample_num)
    else:
        nbytes = lib.load().denet_build_samples_workspace(b, h, w, max_corners)
    ws = workspace(nbytes, dev, "samples")
    call("denet_build_samples_cluster", corner_pr.data_ptr(), b, cn, h, w, corner_threshold, sample_num, max_corners,
         local_max, float(cluster_threshold), pr.data_ptr(), bbox.data_ptr(), ibox.data_ptr(), count.data_ptr(),
         ncand.data_ptr(), ws.data_ptr(), ws.numel() * 4, _stream())
    return pr, bbox, ibox, count, ncand


# ---------------------------------------------------------------------------------------------- targets
MAX_GT = 64   # ground-truth boxes per image the device target builders handle (csrc/targets.cu kMaxGt)


def corner_target(gt, cn, H, W, out):
    """gt = (gt_bbox (B,G,4) f64, gt_class (B,G) i32, gt_count (B) i32) device tensors -> out (B,2,cn,H,W) fp32"""
    gt_bbox, _, gt_count = gt
    b, g = gt_bbox.shape[:2]
    call("denet_corner_target", gt_bbox.data_ptr(), gt_count.data_ptr(), b, g, cn, H, W, out.data_ptr(), _stream())
    return out


def detect_target(gt, sample_bbox64, sn, class_num, thr0, thr1, use_bbox, det, valid, reg, fit_mode=0, fit=None):
    """fit_mode bit0 joint fitness (det has classNum*5+1 channels), bit1 independent fitness (fit (B,6,sn,sn))"""
    gt_bbox, gt_class, gt_count = gt
    b, g = gt_bbox.shape[:2]
    assert sample_bbox64.dtype == torch.float64 and sample_bbox64.is_contiguous()
    call("denet_detect_target_v2", gt_bbox.data_ptr(), gt_class.data_ptr(), gt_count.data_ptr(),
         sample_bbox64.data_ptr(), b, g, sn, class_num, float(thr0), float(thr1), int(use_bbox), int(fit_mode),
         det.data_ptr(), _ptr(valid), _ptr(reg), _ptr(fit), _stream())


# ---------------------------------------------------------------------------------------------- costs
def _loss_ws(device):
    return workspace(lib.load().denet_loss_workspace_bytes(), device, "loss")


def corner_logprob(z, cn):
    """z (B,H,W,>=cn) -> (B,2,cn,H,W) fp32"""
    b, h, w, _ = z.shape
    out = torch.empty((b, 2, cn, h, w), dtype=torch.float32, device=z.device)
    call("denet_corner_logprob", z.data_ptr(), _dtype_code(z), _pitch(z), b, cn, h, w, out.data_ptr(), _stream())
    return out


def corner_cost(z, cn, target, cost_factor, grad_factor, dz, cost):
    b, h, w, _ = z.shape
    assert _pitch(dz) == _pitch(z) and dz.dtype == z.dtype
    ws = _loss_ws(z.device)
    call("denet_corner_cost", z.data_ptr(), _dtype_code(z), _pitch(z), b, cn, h, w, target.data_ptr(), cost_factor,
         grad_factor, dz.data_ptr(), cost.data_ptr(), ws.data_ptr(), _stream())


def detect_cost(o, sn, s0, box_mode, target_det, target_valid, target_reg, cost_factor, bbox_factor, grad_factor, dout,
                cost3, nfit=0, target_fit=None, fit_factor=0.0, sample_bbox=None):
    """box_mode 0 none / 1 Fast R-CNN / 2 bounded IoU (needs sample_bbox (B,sn,sn,4) fp32); nfit independent-fitness
    logits after the box outputs; cost3 (3) fp32 = {detection, box, fitness} cost"""
    b = o.shape[0]
    assert _pitch(dout) == _pitch(o) and dout.dtype == o.dtype and cost3.numel() >= 3
    assert sample_bbox is None or (sample_bbox.dtype == torch.float32 and sample_bbox.is_contiguous())
    ws = _loss_ws(o.device)
    call("denet_detect_cost_v2", o.data_ptr(), _dtype_code(o), _pitch(o), b, sn, s0, int(box_mode), int(nfit),
         _ptr(sample_bbox), target_det.data_ptr(), _ptr(target_valid), _ptr(target_reg), _ptr(target_fit), cost_factor,
         bbox_factor, fit_factor, grad_factor, dout.data_ptr(), dout.shape[-1], cost3.data_ptr(), ws.data_ptr(),
         _stream())


def softmax_nll(o, classes, label, grad_factor, dout, logp, cost):
    """o: (B, 1, 1, classes) logits; label int32 (B)"""
    b = o.shape[0]
    ws = _loss_ws(o.device)
    call("denet_softmax_nll", o.data_ptr(), _dtype_code(o), _pitch(o), b, classes, label.data_ptr(), grad_factor,
         _ptr(dout), _ptr(logp), cost.data_ptr(), ws.data_ptr(), _stream())


# ---------------------------------------------------------------------------------------------- inference tail
def detect_outputs(logits, sn, s0, use_bbox, sample_bbox, class_num=None, fit_mode=0, thr0=0.5):
    """detect-layer logits (B,sn,sn,>=s0[+4][+6]) fp32 -> (det_pr (B,classNum+1,sn,sn) log-probabilities,
    fitness (same shape; det_pr itself without a fitness head), bbox (B,sn,sn,4)).
    fit_mode bit0: joint fitness (s0 = classNum*5+1), bit1: independent fitness"""
    assert logits.dtype == torch.float32
    b = logits.shape[0]
    class_num = s0 - 1 if class_num is None else class_num
    det_pr = torch.empty((b, class_num + 1, sn, sn), dtype=torch.float32, device=logits.device)
    fitness = torch.empty_like(det_pr) if fit_mode else None
    bbox = torch.empty((b, sn, sn, 4), dtype=torch.float32, device=logits.device)
    call("denet_detect_outputs_v2", logits.data_ptr(), _pitch(logits), b, sn, s0, int(use_bbox), class_num,
         int(fit_mode), float(thr0), sample_bbox.data_ptr(), det_pr.data_ptr(), _ptr(fitness), bbox.data_ptr(), _stream())
    return det_pr, (fitness if fit_mode else det_pr), bbox


def detections_nms(det_pr, fitness, bbox, bbox_num, pr_threshold, nms_threshold, use_soft_nms):
    """det_pr / fitness (B, classes+1, sn, sn) fp32 device, bbox (B,sn,sn,4) fp32 device, bbox_num (B) int32 device.
    Returns (score (B,classes,K), index (B,classes,K) int32, count (B,classes) int32) device tensors."""
    assert det_pr.dtype == torch.float32 and det_pr.is_contiguous() and fitness.is_contiguous()
    b, c1, sn, _ = det_pr.shape
    k, classes = sn * sn, c1 - 1
    dev = det_pr.device
    score = torch.empty((b, classes, k), dtype=torch.float32, device=dev)
    index = torch.empty((b, classes, k), dtype=torch.int32, device=dev)
    count = torch.empty((b, classes), dtype=torch.int32, device=dev)
    call("denet_detections_nms", det_pr.data_ptr(), fitness.data_ptr(), c1 * k, k, 1, bbox.data_ptr(), bbox_num.data_ptr(),
         b, classes, k, float(pr_threshold), float(nms_threshold), int(use_soft_nms), score.data_ptr(), index.data_ptr(),
         count.data_ptr(), _stream())
    return score, index, count
