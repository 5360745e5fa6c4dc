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


def busy_corner_map(B, H, W, k, seed, quantize=None, corner_num=4):
    """SURVEY.md §8d 'busy' corner map: logits 5+N(0,1) (not a corner) with k strong corners per type
    (corner_num = 5 adds the centre map of DNC.C)"""
    rng = numpy.random.RandomState(seed)
    z = 5.0 + rng.randn(B, corner_num, H, W).astype(numpy.float32)
    for b in range(B):
        for c in range(corner_num):
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


def nms_inputs(B, classes, sn, seed, crowd=True):
    """random detect-layer outputs: log-softmax class scores, boxes clustered around a few centres so that NMS bites"""
    rs = numpy.random.RandomState(seed)
    z = rs.randn(B, classes + 1, sn, sn).astype(numpy.float32) * 2
    z[:, classes] += 1.0
    m = z.max(axis=1, keepdims=True)
    det_pr = ((z - m) - numpy.log(numpy.exp(z - m).sum(axis=1, keepdims=True))).astype(numpy.float32)
    cx, cy = rs.uniform(0.2, 0.8, (B, 6)), rs.uniform(0.2, 0.8, (B, 6))
    which = rs.randint(0, 6, (B, sn, sn))
    bx = numpy.take_along_axis(cx, which.reshape(B, -1), 1).reshape(B, sn, sn) + rs.randn(B, sn, sn) * (0.03 if crowd else 0.3)
    by = numpy.take_along_axis(cy, which.reshape(B, -1), 1).reshape(B, sn, sn) + rs.randn(B, sn, sn) * (0.03 if crowd else 0.3)
    w, h = rs.uniform(0.05, 0.3, (B, sn, sn)), rs.uniform(0.05, 0.3, (B, sn, sn))
    bbox = numpy.stack([bx - w / 2, by - h / 2, bx + w / 2, by + h / 2], axis=-1).astype(numpy.float32)
    num = [int(v) for v in rs.randint(sn * sn // 2, sn * sn + 1, B)]
    return det_pr, bbox, num
