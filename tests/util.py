"""shared helpers for the parity tests"""
import random

import numpy
import torch


def log_softmax_corner(z):
    """(B,C,H,W) logits -> (B,2,C,H,W) fp32 log_softmax([z,-z]) like theano_util.log_softmax (float32 arithmetic)"""
    z = z.astype(numpy.float32)
    lh = numpy.stack([z, -z], axis=1)
    m = lh.max(axis=1, keepdims=True)
    d = lh - m
    return (d - numpy.log(numpy.exp(d).sum(axis=1, keepdims=True, dtype=numpy.float32))).astype(numpy.float32)


def busy_corner_map(B, H, W, k, seed, quantize=None):
    """SURVEY.md §8d 'busy' corner map: logits 5+N(0,1) (not a corner) with k strong corners per type"""
    rng = numpy.random.RandomState(seed)
    z = 5.0 + rng.randn(B, 4, H, W).astype(numpy.float32)
    for b in range(B):
        for c in range(4):
            for _ in range(k):
                z[b, c, rng.randint(H), rng.randint(W)] = -1.0 - 4.0 * rng.rand()
    if quantize:
        z = numpy.round(z * quantize) / quantize
    return log_softmax_corner(z)


def synthetic_metas(B, classes, seed, max_boxes=8):
    """SURVEY.md §8d metas recipe: 1-8 GT boxes/img, x0,y0~U(0,0.7), w,h~U(0.1,0.3) clipped to 1"""
    rnd = random.Random(seed)
    metas = []
    for _ in range(B):
        boxes, cls = [], []
        for _ in range(rnd.randint(1, max_boxes)):
            x0, y0 = rnd.uniform(0, 0.7), rnd.uniform(0, 0.7)
            boxes.append((x0, y0, min(1.0, x0 + rnd.uniform(0.1, 0.3)), min(1.0, y0 + rnd.uniform(0.1, 0.3))))
            cls.append(rnd.randint(0, classes - 1))
        metas.append({"bbox": boxes, "class": cls, "image_class": cls[0]})
    return metas


def relerr(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def nhwc(x_nchw, dtype, device):
    """numpy/torch NCHW -> padded-pitch NHWC device tensor"""
    from denet_b200 import ops
    x = torch.as_tensor(x_nchw).float()
    n, c, h, w = x.shape
    out = ops.alloc_nhwc(n, h, w, c, dtype, device, zero=True)
    out.copy_(x.permute(0, 2, 3, 1).to(device))
    return out


def nchw(x_nhwc):
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous().cpu()
