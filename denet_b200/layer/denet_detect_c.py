"""Drop-in for the reference's CPython extension `denet_detect` (denet/layer/denet_detect.cc:101-200), backed by the
CUDA NMS kernel (csrc/detect_nms.cu): one CTA per (image, class) instead of one CPU thread for everything.

    detlists = c_code.build_detections_nms(pr_threshold, nms_threshold, use_soft_nms, det_pr, fitness, bboxs,
                                           sample_bbox_num)                       # denet_detect.py:405

Same argument order, same return value: list[B] of lists of (pr, cls, (x0, y0, x1, y1)) python objects, class-major,
bit-identical scores and boxes.  Inputs may be the host ndarrays the reference passes or CUDA tensors.
"""
import numpy
import torch

from .. import lib, ops


def _dev(a, dtype):
    if torch.is_tensor(a):
        return a.to(device="cuda", dtype=dtype).contiguous()
    return torch.from_numpy(numpy.ascontiguousarray(a, dtype={torch.float32: numpy.float32,
                                                               torch.int32: numpy.int32}[dtype])).cuda()


def build_detections_nms(pr_threshold, nms_threshold, use_soft_nms, det_pr, fitness, bboxs, sample_bbox_num):
    if not torch.cuda.is_available():
        raise lib.DenetError("denet_detect.build_detections_nms: needs a CUDA device (no CPU fallback)")
    det_pr, fitness, bboxs = _dev(det_pr, torch.float32), _dev(fitness, torch.float32), _dev(bboxs, torch.float32)
    num = _dev(numpy.asarray(sample_bbox_num, dtype=numpy.int32) if not torch.is_tensor(sample_bbox_num)
               else sample_bbox_num, torch.int32)
    score, index, count = ops.detections_nms(det_pr, fitness, bboxs, num, pr_threshold, nms_threshold, use_soft_nms)
    score, index, count, box = score.cpu().numpy(), index.cpu().numpy(), count.cpu().numpy(), bboxs.cpu().numpy()
    b, classes = count.shape
    box = box.reshape(b, -1, 4)
    out = []
    for i in range(b):
        dets = []
        for cls in range(classes):
            for n in range(int(count[i, cls])):
                bb = box[i, index[i, cls, n]]
                dets.append((float(score[i, cls, n]), cls, (float(bb[0]), float(bb[1]), float(bb[2]), float(bb[3]))))
        out.append(dets)
    return out
