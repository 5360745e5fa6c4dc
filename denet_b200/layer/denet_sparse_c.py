"""Drop-in for the reference's CPython extension `denet_sparse` (denet/layer/denet_sparse.cc:559-706), backed by the
CUDA sampler (csrc/build_samples.cu) instead of one std::thread per image.

Same three functions, same argument order and the same return objects, so the reference call site
(denet/layer/denet_sparse.py:129-141)

    sample_bboxs = c_code.build_samples(self.thread_num, corner_pr, self.corner_threshold, self.sample_num,
                                        self.corner_max, self.local_max, self.nms_threshold)

keeps working when `common.import_c("denet_sparse.cc")` hands out this module.  corner_pr may be the host ndarray the
reference passes (uploaded here) or a CUDA tensor that never left the device.
"""
import numpy
import torch

from .. import lib, ops


def init_logging(fname):
    """denet_sparse.cc:21-30 opens a log file for LOG_PRINT; the CUDA path has nothing to log there"""
    return None


def build_samples(thread_num, corner_pr, corner_threshold, sample_num, max_corners, local_max, cluster_threshold):
    """-> list[B] of list[<= sample_num^2] of (pr, (x0, y0, x1, y1)), python floats, sorted by pr descending.
    thread_num is accepted for signature compatibility (the GPU kernel runs one CTA per image x corner type)."""
    if not torch.cuda.is_available():
        raise lib.DenetError("denet_sparse.build_samples: needs a CUDA device (no CPU fallback)")
    if torch.is_tensor(corner_pr):
        cp = corner_pr.to(device="cuda", dtype=torch.float32).contiguous()
    else:
        cp = torch.from_numpy(numpy.ascontiguousarray(corner_pr, dtype=numpy.float32)).cuda()
    if cp.dim() != 5 or cp.shape[1] != 2:
        raise ValueError("build_samples: corner_pr must be (B,2,corner_num,H,W), got %s" % (tuple(cp.shape),))
    pr, bbox, _, count, _ = ops.build_samples(cp, float(corner_threshold), int(sample_num), int(max_corners),
                                              int(local_max), float(cluster_threshold))
    pr, bbox, count = pr.cpu().numpy(), bbox.cpu().numpy(), count.cpu().numpy()
    return [[(float(pr[b, i]), (float(bbox[b, i, 0]), float(bbox[b, i, 1]), float(bbox[b, i, 2]), float(bbox[b, i, 3])))
             for i in range(int(count[b]))] for b in range(cp.shape[0])]


def build_bbox_array(samples, bbox):
    """denet_sparse.cc:670-699: sample i of image b -> bbox[b, i // sn, i % sn, :] (in place, float32)"""
    sn = bbox.shape[1]
    for b, image in enumerate(samples):
        for i, s in enumerate(image):
            bbox[b, i // sn, i % sn, :] = s[1]
    return None
