"""ORACLE - TEST INFRASTRUCTURE ONLY.

torch-CPU / numpy restatement of the Theano side of the reference's hot path, in the reference's own NCHW layout.
Gradients come from torch autograd (the reference uses theano tensor.grad, model_cnn.py:318).  Every function
cites the reference lines it follows (paths relative to the reference repository).

Pinning: the arithmetic below is performed in the reference by Theano (git fadc8be4, README.md:17-24) + cuDNN 5.1,
neither of which is vendored or installable here, so conv / pooling / log-softmax / solver are *parity unpinned*
by reference-run outputs.  They are pinned only by (i) the reference's single numeric known-answer test
(batch_norm.py:131-154, mean running inverse-std 1.24641 - reproduced in tests/test_oracle.py) and (ii) cross-checks
between two independent restatements (torch conv2d vs a naive numpy loop).  See oracle/README.md.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

import oracle

# ----------------------------------------------------------------------------------------------------------------
# shapes


def conv_out_hw(in_hw, k_hw, stride, border_mode):
    """output (h, w) exactly as convolution.py:55-74 computes it"""
    out = []
    for i in range(2):
        n, k, s = in_hw[i], k_hw[i], stride[i]
        if border_mode == "valid":
            o = math.ceil((n - k + 1) / s)
        elif border_mode == "full":
            o = math.ceil((n + k - 1) / s)
        elif border_mode == "half":
            o = math.ceil((n + 2 * (k // 2) - k + 1) / s)
        elif border_mode == "same":
            assert tuple(stride) == (1, 1)
            o = n
        elif isinstance(border_mode, int):
            o = math.ceil((n + 2 * border_mode - k + 1) / s)
        elif isinstance(border_mode, (tuple, list)):
            o = math.ceil((n + 2 * border_mode[i] - k + 1) / s)
        else:
            raise Exception("Unknown border mode: " + str(border_mode))
        out.append(o)
    return tuple(out)


def conv_pad(k_hw, border_mode):
    if border_mode == "valid":
        return (0, 0)
    if border_mode == "full":
        return (k_hw[0] - 1, k_hw[1] - 1)
    if border_mode == "half":
        return (k_hw[0] // 2, k_hw[1] // 2)
    if isinstance(border_mode, (int, bool)):   # denet_corner.py:39 passes border_mode=False (== pad 0)
        return (int(border_mode), int(border_mode))
    if isinstance(border_mode, (tuple, list)):
        return tuple(int(v) for v in border_mode)
    raise Exception("Unknown border mode: " + str(border_mode))


# ----------------------------------------------------------------------------------------------------------------
# layers (forward functions on torch CPU tensors, NCHW)


def conv2d(x, w, stride=(1, 1), border_mode="half", bias=None):
    """tensor.nnet.conv2d(input, filters, subsample, border_mode) with Theano's default filter_flip=True
    (convolution.py:76-92): a TRUE convolution = correlation with the spatially flipped filter."""
    wf = torch.flip(w, dims=(2, 3))
    if border_mode == "same":  # convolution.py:76-80: full convolution, then crop
        k = w.shape[2:]
        y = F.conv2d(x, wf, padding=(k[0] - 1, k[1] - 1))
        y0, x0 = (k[0] - 1) // 2, (k[1] - 1) // 2
        y = y[:, :, y0:y0 + x.shape[2], x0:x0 + x.shape[3]]
    else:
        y = F.conv2d(x, wf, stride=tuple(stride), padding=conv_pad(w.shape[2:], border_mode))
    if bias is not None:
        y = y + bias[None, :, None, None]
    return y


def conv2d_naive(x, w, stride=(1, 1), pad=(0, 0)):
    """independent numpy restatement (direct loops, float64) used only to cross-check conv2d on tiny shapes"""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    n, c, h, ww = x.shape
    co, ci, kh, kw = w.shape
    xp = np.zeros((n, c, h + 2 * pad[0], ww + 2 * pad[1]))
    xp[:, :, pad[0]:pad[0] + h, pad[1]:pad[1] + ww] = x
    oh = (h + 2 * pad[0] - kh) // stride[0] + 1
    ow = (ww + 2 * pad[1] - kw) // stride[1] + 1
    y = np.zeros((n, co, oh, ow))
    for i in range(oh):
        for j in range(ow):
            patch = xp[:, :, i * stride[0]:i * stride[0] + kh, j * stride[1]:j * stride[1] + kw]
            # true convolution: filter index runs backwards over the window
            y[:, :, i, j] = np.einsum("ncrs,ocrs->no", patch, w[:, :, ::-1, ::-1])
    return y


def batchnorm_train(x, gamma, beta, eps):
    """cuDNN spatial BN forward-training as the reference calls it (batch_norm.py:47-52):
    returns y, batch mean, batch inverse std (1/sqrt(biased var + eps))."""
    mean = x.mean(dim=(0, 2, 3))
    var = ((x - mean[None, :, None, None]) ** 2).mean(dim=(0, 2, 3))
    invstd = 1.0 / torch.sqrt(var + eps)
    y = (x - mean[None, :, None, None]) * (gamma * invstd)[None, :, None, None] + beta[None, :, None, None]
    return y, mean, invstd


def batchnorm_test(x, gamma, beta, run_mean, run_stdinv, eps):
    """batch_norm.py:50-52: var = (1/stdinv)^2 handed to cuDNN inference, which adds eps again"""
    var = (1.0 / run_stdinv) ** 2
    return (x - run_mean[None, :, None, None]) * (gamma / torch.sqrt(var + eps))[None, :, None, None] + \
        beta[None, :, None, None]


def bn_running_update(run, batch_value, momentum):
    """batch_norm.py:75-76: EMA of the batch mean and of the batch INVERSE STD (not the variance)"""
    return momentum * run + (1.0 - momentum) * batch_value


def relu(x):
    """tensor.nnet.relu == 0.5*(x+|x|) (activation.py:32-34) and k_relu (batch_norm_relu.py:34-39)"""
    return 0.5 * (x + x.abs())


def pool_out_hw(in_hw, size, stride, pad, ignore_border=True):
    """pool.py:28-34"""
    if ignore_border:
        return tuple(int(math.floor((in_hw[i] + 2 * pad[i] - size[i]) / stride[i])) + 1 for i in range(2))
    return tuple(int(math.ceil((in_hw[i] + 2 * pad[i]) / stride[i])) for i in range(2))


def pool2d(x, size, stride, pad, mode):
    """dnn_pool (pool.py:36-38): 'max' pads with -inf, 'average_inc_pad' counts the zero padding"""
    if mode == "max":
        return F.max_pool2d(x, kernel_size=tuple(size), stride=tuple(stride), padding=tuple(pad))
    if mode == "average_inc_pad":
        return F.avg_pool2d(x, kernel_size=tuple(size), stride=tuple(stride), padding=tuple(pad),
                            count_include_pad=True)
    raise Exception("unsupported pool mode " + mode)


def max_pool2d_routed(x, size, stride, pad, given, tol):
    """max pooling whose gradient routing follows `given` (window tap index r*kw+s, (B,C,oh,ow)) wherever the two
    largest values of a window are closer than tol * max|x| - there an fp32 run and an fp64 run may pick different
    taps, which re-routes a gradient element although both forward values agree.  Returns (y, number of re-routes)."""
    kh, kw = size
    B, C, H, W = x.shape
    oh, ow = pool_out_hw((H, W), size, stride, pad)
    xp = F.pad(x, (pad[1], pad[1], pad[0], pad[0]), value=float("-inf"))
    win = xp.unfold(2, kh, stride[0]).unfold(3, kw, stride[1])[:, :, :oh, :ow].reshape(B, C, oh, ow, kh * kw)
    with torch.no_grad():
        top = win.topk(2, dim=-1)
        own = top.indices[..., 0]
        given = torch.as_tensor(given).reshape(own.shape).long().clamp(0, kh * kw - 1)
        eps = tol * x.abs().max()
        near = (top.values[..., 0] - top.values[..., 1]) < eps
        ok = (top.values[..., 0] - win.gather(-1, given.unsqueeze(-1)).squeeze(-1)) < eps
        idx = torch.where(near & ok, given, own)
        moved = int((idx != own).sum())
    return win.gather(-1, idx.unsqueeze(-1)).squeeze(-1), moved


def pool_inv(x, size):
    """pool_inv.py:26 (CPU path): repeat along h by size[1], along w by size[0]  == k_pool_inv"""
    return x.repeat_interleave(size[1], dim=2).repeat_interleave(size[0], dim=3)


def log_softmax(x, axis):
    """theano_util.py:27-29 / regression.py:66-68"""
    xdev = x - x.max(dim=axis, keepdim=True)[0]
    return xdev - torch.log(torch.sum(torch.exp(xdev), dim=axis, keepdim=True))


def smooth_l1(x):
    """theano_util.py:32-34"""
    xa = x.abs()
    return torch.where(xa < 1, 0.5 * x ** 2, xa - 0.5)


def sparse_sample(fmap, bbox, gs):
    """DeNetSparseOp forward (denet_sparse_op.py:42-85) as a differentiable gather: the integer grid positions come
    from the C restatement of the kernel's float sequence, the gather itself from torch indexing (so autograd
    yields what k_sparse_sample_grad accumulates).  fmap (B,F,H,W), bbox numpy (B,sn,sn,4) -> (B, gs*gs*F+2, sn, sn)"""
    B, Fc, H, W = fmap.shape
    sn = bbox.shape[1]
    ys, xs = oracle.sparse_sample_index(bbox, gs, H, W)  # (B,sn,sn,gs)
    ys = torch.from_numpy(ys.astype(np.int64))
    xs = torch.from_numpy(xs.astype(np.int64))
    bidx = torch.arange(B)[:, None, None, None, None]
    yy = ys[:, :, :, :, None].expand(B, sn, sn, gs, gs)
    xx = xs[:, :, :, None, :].expand(B, sn, sn, gs, gs)
    g = fmap.permute(0, 2, 3, 1)[bidx, yy, xx]            # (B,sn,sn,gs,gs,F)
    g = g.reshape(B, sn, sn, gs * gs * Fc).permute(0, 3, 1, 2)
    bb = torch.from_numpy(np.asarray(bbox, np.float32)).to(fmap.dtype)
    bh = (bb[..., 3] - bb[..., 1])[:, None]
    bw = (bb[..., 2] - bb[..., 0])[:, None]
    return torch.cat([g, bh, bw], dim=1)


# ----------------------------------------------------------------------------------------------------------------
# host-side target builders (straight loop restatements)


def corner_target(metas, corner_shape, use_center=False):
    """DeNetCornerLayer.get_target, denet_corner.py:81-123 (dropout = 0)"""
    B, _, corner_num, height, width = corner_shape
    corner_pr = np.zeros(corner_shape, dtype=np.float32)
    for b, meta in enumerate(metas):
        for bbox in meta["bbox"]:
            x0 = int(round(bbox[0] * width))
            y0 = int(round(bbox[1] * height))
            x1 = max(x0, int(round(bbox[2] * width)) - 1)
            y1 = max(y0, int(round(bbox[3] * height)) - 1)
            x0v, y0v = 0 <= x0 < width, 0 <= y0 < height
            x1v, y1v = 0 <= x1 < width, 0 <= y1 < height
            if x0v and y0v:
                corner_pr[b, 1, 0, y0, x0] = 1.0
            if x1v and y0v:
                corner_pr[b, 1, 1, y0, x1] = 1.0
            if x0v and y1v:
                corner_pr[b, 1, 2, y1, x0] = 1.0
            if x1v and y1v:
                corner_pr[b, 1, 3, y1, x1] = 1.0
            if use_center:
                cx = int(round((bbox[0] + bbox[2]) * 0.5 * width))
                cy = int(round((bbox[1] + bbox[3]) * 0.5 * height))
                if 0 <= cx < width and 0 <= cy < height:
                    corner_pr[b, 1, 4, cy, cx] = 1.0
    corner_pr[:, 0] = 1.0 - corner_pr[:, 1]
    corner_pr /= width * height * corner_num
    return np.array([], dtype=np.int64), corner_pr.flatten()


def overlap_iou_matrix(obj_bboxs, sample_bboxs):
    """theano_util.py:38-59, float32 like the compiled Theano function (allow_input_downcast)"""
    x = np.array(obj_bboxs, dtype=np.float32)
    y = np.array(sample_bboxs, dtype=np.float32)
    x_area = (x[:, 2] - x[:, 0]) * (x[:, 3] - x[:, 1])
    y_area = (y[:, 2] - y[:, 0]) * (y[:, 3] - y[:, 1])
    dx = np.maximum(np.minimum(x[:, None, 2], y[None, :, 2]) - np.maximum(x[:, None, 0], y[None, :, 0]), 0)
    dy = np.maximum(np.minimum(x[:, None, 3], y[None, :, 3]) - np.maximum(x[:, None, 1], y[None, :, 1]), 0)
    inter = dx * dy
    union = x_area[:, None] + y_area[None, :] - inter
    return inter / union


def detect_target(metas, sample_bbox_list, batch_size, sample_num, class_num, overlap_threshold, use_bbox_reg,
                  use_jointfit=False, use_indfit=False):
    """DeNetDetectLayer.get_target, denet_detect.py:147-235, incl. the joint-fitness (:179-182) and independent-fitness
    (:187-191) targets.  overlap_threshold: scalar or pair (the reference indexes [0]/[1] although parse_desc passes a
    scalar).  sample_f (:177) mixes a numpy float32 scalar with python floats: under the numpy the reference ran on
    (1.x, value-based scalar promotion) that arithmetic is float64, which is what float(...) restates here."""
    thr = overlap_threshold if isinstance(overlap_threshold, (tuple, list)) else (overlap_threshold,
                                                                                  overlap_threshold)
    fitness_num = 5 if use_jointfit else 6                                              # :58-66
    null_class = class_num * fitness_num if use_jointfit else class_num
    det_shape = (batch_size, null_class + 1, sample_num, sample_num)
    det_pr = np.zeros(det_shape, dtype=np.float32)
    det_pr[:, null_class] = 1.0
    if use_bbox_reg:
        bbox_valid = np.zeros((batch_size, sample_num, sample_num), dtype=np.float32)
        bbox_reg = np.zeros((batch_size, 8, sample_num, sample_num), dtype=np.float32)
        bbox_reg[:, [2, 3, 6, 7]] = 1.0
    if use_indfit:
        indfit_pr = np.zeros((batch_size, fitness_num, sample_num, sample_num), dtype=np.float32)
        indfit_pr[:, 0] = 1.0
    for b, meta in enumerate(metas):
        samples = [bbox for _, bbox in sample_bbox_list[b]]
        if len(meta["bbox"]) > 0 and len(samples) > 0:
            overlap = overlap_iou_matrix(meta["bbox"], samples)
            bbox_indexs, sample_indexs = np.where(overlap > thr[0])
            for obj, index in zip(bbox_indexs.tolist(), sample_indexs.tolist()):
                si, sj = index % sample_num, index // sample_num
                sample_cls = meta["class"][obj]
                sample_f = (float(overlap[obj, index]) - thr[0]) / (1.0 - thr[0])        # :177
                if use_jointfit:
                    f = max(0, min(int(fitness_num * sample_f), fitness_num - 1))
                    det_pr[b, sample_cls * fitness_num + f, sj, si] = 1.0
                else:
                    det_pr[b, sample_cls, sj, si] = 1.0
                det_pr[b, null_class, sj, si] = 0.0
                if use_indfit:
                    f = 1 + int(math.floor((fitness_num - 1) * sample_f))
                    f = max(1, min(f, fitness_num - 1))
                    indfit_pr[b, 0, sj, si] = 0.0
                    indfit_pr[b, f, sj, si] = 1.0
            if use_bbox_reg:
                overlap_max = overlap.argmax(axis=0)
                for index in range(len(samples)):
                    obj = overlap_max[index]
                    if overlap[obj, index] <= thr[1]:
                        continue
                    sample, target = samples[index], meta["bbox"][obj]
                    si, sj = index % sample_num, index // sample_num
                    bbox_valid[b, sj, si] = 1.0
                    bbox_reg[b, 0, sj, si] = 0.5 * (target[0] + target[2])
                    bbox_reg[b, 1, sj, si] = 0.5 * (target[1] + target[3])
                    bbox_reg[b, 2, sj, si] = target[2] - target[0]
                    bbox_reg[b, 3, sj, si] = target[3] - target[1]
                    bbox_reg[b, 4, sj, si] = 0.5 * (sample[0] + sample[2])
                    bbox_reg[b, 5, sj, si] = 0.5 * (sample[1] + sample[3])
                    bbox_reg[b, 6, sj, si] = sample[2] - sample[0]
                    bbox_reg[b, 7, sj, si] = sample[3] - sample[1]
    det_pr /= det_pr.sum(axis=1)[:, None]
    if use_indfit:
        indfit_pr /= indfit_pr.sum(axis=1)[:, None]
    nfactor = sample_num * sample_num
    det_pr /= nfactor
    yt_value = det_pr.flatten()
    if use_bbox_reg:
        bbox_valid /= nfactor
        yt_value = np.concatenate((yt_value, bbox_valid.flatten(), bbox_reg.flatten()))
    if use_indfit:
        indfit_pr /= nfactor
        yt_value = np.concatenate((yt_value, indfit_pr.flatten()))
    return np.array([], dtype=np.int64), yt_value


def detect_errors(det_pr, bbox_reg, indfit_pr, sample_bbox, yt_value, bbox_factor, use_bounded_iou):
    """DeNetDetectLayer.get_errors, denet_detect.py:238-301, on torch tensors (differentiable).
    det_pr (B,s0,sn,sn) log-softmax, bbox_reg (B,4,sn,sn) or None, indfit_pr (B,nf,sn,sn) log-softmax or None,
    sample_bbox (B,sn,sn,4); returns (det_errors, bbox_errors | None, indfit_errors | None), each (B,sn,sn)."""
    B, s0, sn, _ = det_pr.shape
    n0, n1 = B * s0 * sn * sn, B * sn * sn
    v = yt_value
    det_t = v[:n0].reshape(det_pr.shape)
    det_errors = -(det_t * det_pr).sum(dim=1) / math.log(s0)                              # :257
    off = n0
    bbox_errors = None
    if bbox_reg is not None:
        valid = v[off:off + n1].reshape(B, sn, sn)
        reg = v[off + n1:off + 9 * n1].reshape(B, 8, sn, sn)
        off += 9 * n1
        tgt, smp = reg[:, 0:4], reg[:, 4:8]
        if use_bounded_iou:                                                                # :266-286
            sb = sample_bbox
            sample_cx, sample_cy = 0.5 * (sb[..., 0] + sb[..., 2]), 0.5 * (sb[..., 1] + sb[..., 3])     # :84-87
            sample_w, sample_h = sb[..., 2] - sb[..., 0], sb[..., 3] - sb[..., 1]
            pcx, pcy = bbox_reg[:, 0] * sample_w + sample_cx, bbox_reg[:, 1] * sample_h + sample_cy     # :89-92
            pw, ph = torch.exp(bbox_reg[:, 2]) * sample_w, torch.exp(bbox_reg[:, 3]) * sample_h
            x0, y0, x1, y1 = pcx - pw * 0.5, pcy - ph * 0.5, pcx + pw * 0.5, pcy + ph * 0.5             # :93-96
            predict_x, predict_y, predict_w, predict_h = 0.5 * (x0 + x1), 0.5 * (y0 + y1), x1 - x0, y1 - y0
            dx, dy = tgt[:, 0] - predict_x, tgt[:, 1] - predict_y
            eps = 0.001
            cost_x = torch.where(dx >= 0.0, 2 * dx / (tgt[:, 2] + dx + eps), -2 * dx / (tgt[:, 2] - dx + eps))
            cost_y = torch.where(dy >= 0.0, 2 * dy / (tgt[:, 3] + dy + eps), -2 * dy / (tgt[:, 3] - dy + eps))
            cost_w = 1.0 - torch.minimum(tgt[:, 2] / (predict_w + eps), predict_w / (tgt[:, 2] + eps))
            cost_h = 1.0 - torch.minimum(tgt[:, 3] / (predict_h + eps), predict_h / (tgt[:, 3] + eps))
            cost = torch.stack([cost_x, cost_y, cost_w, cost_h], dim=1)
            bbox_errors = bbox_factor * valid * smooth_l1(cost).sum(dim=1)
        else:
            tx = (tgt[:, 0] - smp[:, 0]) / smp[:, 2]                                      # :289-292
            ty = (tgt[:, 1] - smp[:, 1]) / smp[:, 3]
            tw = torch.log(tgt[:, 2] / smp[:, 2])
            th = torch.log(tgt[:, 3] / smp[:, 3])
            dt = torch.stack([tx, ty, tw, th], dim=1) - bbox_reg
            bbox_errors = bbox_factor * valid * smooth_l1(dt).sum(dim=1)                   # :295
    indfit_errors = None
    if indfit_pr is not None:
        nf = indfit_pr.shape[1]
        fit_t = v[off:off + nf * n1].reshape(indfit_pr.shape)
        indfit_errors = -(fit_t * indfit_pr).sum(dim=1) / math.log(nf)                    # :299
    return det_errors, bbox_errors, indfit_errors


def detect_outputs(logits, sample_bbox, class_num, overlap_threshold, use_bbox_reg, use_jointfit, use_indfit):
    """What get_detections hands to build_detections_nms (denet_detect.py:330-399), numpy float32:
    logits (B, s0+s1+s2, sn, sn), sample_bbox (B,sn,sn,4) -> det_pr (B,classNum+1,sn,sn), fitness (B,classNum[+1],sn,sn),
    bboxs (B,sn,sn,4)."""
    thr0 = overlap_threshold[0] if isinstance(overlap_threshold, (tuple, list)) else overlap_threshold
    o = torch.from_numpy(np.asarray(logits, np.float32))
    fitness_num = 5 if use_jointfit else 6
    s0 = class_num * fitness_num + 1 if use_jointfit else class_num + 1
    s1 = 4 if use_bbox_reg else 0
    lp = log_softmax(o[:, :s0], 1)
    B, _, sn, _ = o.shape
    if use_jointfit:                                                                       # :332-348
        det_fit = lp[:, :class_num * fitness_num].reshape(B, class_num, fitness_num, sn, sn)
        m = det_fit.max(dim=2)[0]
        det_pr = m + torch.log(torch.sum(torch.exp(det_fit - m[:, :, None]), dim=2))
        det_pr = torch.cat([det_pr, lp[:, class_num * fitness_num][:, None]], dim=1)
        val = torch.tensor([thr0 + i * (1.0 - thr0) / fitness_num for i in range(fitness_num)], dtype=torch.float32)
        fitness = torch.log(torch.sum(torch.exp(det_fit) * val[None, None, :, None, None], dim=2))
    else:
        det_pr = lp
        fitness = lp.clone()
    sb = torch.from_numpy(np.asarray(sample_bbox, np.float32))
    if use_bbox_reg:                                                                       # :80-97
        r = o[:, s0:s0 + 4]
        scx, scy = 0.5 * (sb[..., 0] + sb[..., 2]), 0.5 * (sb[..., 1] + sb[..., 3])
        sw, sh = sb[..., 2] - sb[..., 0], sb[..., 3] - sb[..., 1]
        pcx, pcy = r[:, 0] * sw + scx, r[:, 1] * sh + scy
        pw, ph = torch.exp(r[:, 2]) * sw, torch.exp(r[:, 3]) * sh
        bboxs = torch.stack([pcx - pw * 0.5, pcy - ph * 0.5, pcx + pw * 0.5, pcy + ph * 0.5], dim=-1)
    else:
        bboxs = sb
    det_pr, fitness, bboxs = det_pr.numpy(), fitness.numpy(), bboxs.numpy()
    if use_indfit:                                                                         # :392-397
        indfit_pr = torch.exp(log_softmax(o[:, s0 + s1:s0 + s1 + fitness_num], 1)).numpy()
        fitness_val = np.array([0.0] + [thr0 + i * (1.0 - thr0) / (fitness_num - 1) for i in range(fitness_num - 1)])
        fitness_exp = np.sum(indfit_pr * fitness_val[None, :, None, None], axis=1).astype(np.float32)
        fitness = fitness + np.log(fitness_exp)[:, None, :, :]
    return det_pr, fitness, bboxs


def sparse_postprocess(sample_bboxs, metas, sample_count, random_sample, sample_gt, rng):
    """DeNetSparseLayer.get_target post-processing, denet_sparse.py:184-201. rng = python `random` module/instance.
    Mutates and returns the per-image lists of (pr, (x0,y0,x1,y1))."""
    for b, meta in enumerate(metas):
        n = sample_count - math.floor(random_sample * sample_count)
        if len(sample_bboxs[b]) > n:
            sample_bboxs[b] = rng.sample(sample_bboxs[b], n)
        while len(sample_bboxs[b]) < sample_count:
            x0 = rng.uniform(0.0, 1.0)
            y0 = rng.uniform(0.0, 1.0)
            x1 = rng.uniform(x0, 1.0)
            y1 = rng.uniform(y0, 1.0)
            sample_bboxs[b].append((0.0, (x0, y0, x1, y1)))
        if sample_gt:
            for index, bbox in enumerate(meta["bbox"]):
                sample_bboxs[b][-(index + 1)] = (1.0, tuple(bbox))
    return sample_bboxs


def bbox_array(sample_bboxs, batch_size, sample_num):
    """build_bbox_array, denet_sparse.cc:670-699: sample i -> (i // sn, i % sn)"""
    out = np.zeros((batch_size, sample_num, sample_num, 4), dtype=np.float32)
    for b in range(batch_size):
        for i, (_, bbox) in enumerate(sample_bboxs[b]):
            out[b, i // sample_num, i % sample_num] = np.asarray(bbox, dtype=np.float32)
    return out


# ----------------------------------------------------------------------------------------------------------------
# solver (model_cnn.py:282-305, 320-324)


def solver_update(p, g, m, solver, iteration, lr, momentum, decay, is_weight, v=None):
    """one parameter tensor; returns (p_new, m_new[, v_new]).  numpy or torch tensors."""
    if is_weight:
        g = g + decay * p
    rho = momentum[0] if iteration > 0 else 0.0
    if solver in ("torch", "nesterov"):
        m_new = rho * m + g
        p_new = p - lr * (g + momentum[0] * m_new)
        return p_new, m_new
    if solver == "adam":
        eps = 1e-8
        m_new = momentum[0] * m + (1.0 - momentum[0]) * g
        v_new = momentum[1] * v + (1.0 - momentum[1]) * (g * g)
        m_hat = m_new / (1.0 - momentum[0] ** (iteration + 1))
        v_hat = v_new / (1.0 - momentum[1] ** (iteration + 1))
        return p - lr * m_hat / (v_hat ** 0.5 + eps), m_new, v_new
    m_new = rho * m + (1.0 - rho) * g
    return p - lr * m_new, m_new


# ----------------------------------------------------------------------------------------------------------------
# inference tail (test infrastructure like everything here): restatement of denet/layer/denet_detect.cc


def _libm_logf(x):
    import ctypes
    import ctypes.util
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.logf.restype, libm.logf.argtypes = ctypes.c_float, [ctypes.c_float]
    return np.float32(libm.logf(ctypes.c_float(x)))


def libm_expf(x):
    """std::exp(float) as the reference's extension evaluates it"""
    import ctypes
    import ctypes.util
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.expf.restype, libm.expf.argtypes = ctypes.c_float, [ctypes.c_float]
    return np.float32(libm.expf(ctypes.c_float(float(x))))


def _iou_f32(a, b):
    """denet_detect.cc:12-31 in float32, operation for operation"""
    f = np.float32
    dx = max(f(0.0), f(min(a[2], b[2]) - max(a[0], b[0])))
    dy = max(f(0.0), f(min(a[3], b[3]) - max(a[1], b[1])))
    ai = f(dx * dy)
    aa = f(f(a[2] - a[0]) * f(a[3] - a[1]))
    ab = f(f(b[2] - b[0]) * f(b[3] - b[1]))
    au = f(f(aa + ab) - ai)
    with np.errstate(divide="ignore", invalid="ignore"):
        return f(ai / au)


def build_detections_nms(pr_threshold, nms_threshold, use_soft_nms, det_pr, fitness, bbox, bbox_num):
    """denet_detect.cc:101-173 (+ perform_nms :74-99, perform_soft_nms :35-72).  det_pr / fitness (B,C+1,sn,sn) fp32,
    bbox (B,sn,sn,4) fp32.  Returns list[B] of [(fitness score in LOG basis as float32, cls, sample index k)] in the
    reference's output order; the caller applies exp (the reference's std::exp(float) is libm expf)."""
    f = np.float32
    det_pr, fitness, bbox = np.asarray(det_pr, f), np.asarray(fitness, f), np.asarray(bbox, f)
    log_thr = _libm_logf(pr_threshold)           # std::log(float) of the reference = libm logf
    B, C1, sn, _ = det_pr.shape
    out = []
    for b in range(B):
        dets = []
        nb = int(bbox_num[b])
        for cls in range(C1 - 1):
            inst = []
            for k in range(min(nb, sn * sn)):
                j, i = divmod(k, sn)
                if det_pr[b, cls, j, i] >= log_thr:
                    inst.append([f(fitness[b, cls, j, i]), bbox[b, j, i], k])
            if not (0.0 < nms_threshold < 1.0) or not inst:
                keep = inst
            elif use_soft_nms:
                keep, rest = [], list(inst)
                thr, discard = f(nms_threshold), f(-6.9)
                while rest:
                    m = 0
                    for t in range(1, len(rest)):
                        if rest[t][0] > rest[m][0]:
                            m = t
                    M = rest.pop(m)
                    keep.append(M)
                    for r in rest:
                        iou = _iou_f32(M[1], r[1])
                        r[0] = f(r[0] - f(f(iou * iou) / thr))
                    rest = [r for r in rest if not (r[0] < discard)]
            else:
                thr = f(nms_threshold)
                keep = []
                for a in inst:
                    if not any((a[0] < o[0]) and (_iou_f32(a[1], o[1]) > thr) for o in inst):
                        keep.append(a)
            dets += [(s, cls, k) for s, _, k in keep]
        out.append(dets)
    return out
