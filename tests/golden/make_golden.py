"""Generates tests/golden/*.npz from the REAL reference, run once in the build container (needs /root/reference and
oracle/_ref built by oracle/build_ref.py).  The GPU box has no /root/reference; the fixtures travel instead.

  build_samples_*.npz : corner maps (fp32 log-probabilities) + the output of the reference's own C++ extension
                        denet_sparse.build_samples (denet/layer/denet_sparse.cc:559-668, compiled unmodified) and of
                        build_bbox_array (:670-699).
  bn_known_answer.npz : the input recipe of the one numeric known-answer in the reference (batch_norm.py:131-153:
                        U(0,1) seed 1002, shape (64,128,32,32) -> mean running inverse-std 1.24641).

Usage:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from oracle import build_ref  # noqa: E402
from util import busy_corner_map, log_softmax_corner  # noqa: E402


def samples_to_arrays(samples, K):
    B = len(samples)
    pr = numpy.zeros((B, K), numpy.float32)
    bbox = numpy.zeros((B, K, 4), numpy.float32)
    count = numpy.zeros((B,), numpy.int32)
    for b, s in enumerate(samples):
        count[b] = len(s)
        for i, (p, bb) in enumerate(s):
            pr[b, i] = p
            bbox[b, i] = bb
    return pr, bbox, count


def main():
    build_ref.build_all()
    ref = oracle.reference_cc()
    assert ref is not None, "the reference extension did not build"
    cases = {
        "build_samples_busy4_h32_sn8": (busy_corner_map(3, 32, 32, 4, seed=4), 0.01, 8, 1024, 0),
        "build_samples_busy12_h32_sn8": (busy_corner_map(2, 32, 32, 12, seed=12), 0.01, 8, 1024, 0),
        "build_samples_busy40_h64_sn24": (busy_corner_map(2, 64, 64, 40, seed=40), 0.01, 24, 1024, 0),
        "build_samples_localmax_h32_sn8": (busy_corner_map(2, 32, 32, 30, seed=5), 0.01, 8, 1024, 2),
        "build_samples_select_h32_sn8": (log_softmax_corner(numpy.random.RandomState(7).randn(2, 4, 32, 32).astype(
            numpy.float32) * 3), 0.3, 8, 64, 0),
        "build_samples_empty_h16_sn4": (log_softmax_corner(numpy.full((1, 4, 16, 16), 5.0, numpy.float32)), 0.01, 4,
                                        1024, 0),
        # DNC.C: five maps per image, the fifth are box centres (denet_sparse.cc:296-303,377-468)
        "build_samples_centre10_h32_sn8": (busy_corner_map(3, 32, 32, 10, seed=51, corner_num=5), 0.01, 8, 1024, 0),
        "build_samples_centre24_h48_sn12": (busy_corner_map(2, 48, 48, 24, seed=52, corner_num=5), 0.01, 12, 1024, 0),
    }
    for name, (cp, thr, sn, maxc, lm) in cases.items():
        if os.path.exists(os.path.join(HERE, name + ".npz")) and "--all" not in sys.argv:
            continue                                  # fixtures are immutable once committed
        samples = ref.build_samples(cp.shape[0], cp, thr, sn, maxc, lm, 1.0)
        pr, bbox, count = samples_to_arrays(samples, sn * sn)
        arr = numpy.zeros((cp.shape[0], sn, sn, 4), numpy.float32)
        ref.build_bbox_array([list(s) for s in samples], arr)
        numpy.savez_compressed(os.path.join(HERE, name + ".npz"), corner_pr=cp, threshold=numpy.float32(thr),
                               sample_num=sn, max_corners=maxc, local_max=lm, pr=pr, bbox=bbox, count=count,
                               bbox_array=arr)
        print(name, "samples per image", count.tolist())
    if not os.path.exists(os.path.join(HERE, "bn_known_answer.npz")) or "--all" in sys.argv:
        numpy.savez_compressed(os.path.join(HERE, "bn_known_answer.npz"), seed=1002,
                               shape=numpy.array([64, 128, 32, 32]),
                               expected_mean_running_stdinv=numpy.float32(1.24641))
    # detection lists of the reference's own NMS extension (denet/layer/denet_detect.cc:101-173, compiled unmodified)
    det = oracle.reference_detect_cc()
    assert det is not None, "the reference NMS extension did not build"
    from util import nms_inputs
    for name, (B, classes, sn, seed, pr_thr, nms_thr, soft) in {
            "nms_hard_b3_c7_sn8": (3, 7, 8, 11, 0.05, 0.5, 0),
            "nms_soft_b3_c7_sn8": (3, 7, 8, 12, 0.05, 0.5, 1),
            "nms_hard_b2_c20_sn12": (2, 20, 12, 13, 0.1, 0.3, 0),
            "nms_off_b2_c5_sn6": (2, 5, 6, 14, 0.02, 1.0, 0)}.items():
        if os.path.exists(os.path.join(HERE, name + ".npz")) and "--all" not in sys.argv:
            continue
        det_pr, bbox, num = nms_inputs(B, classes, sn, seed)
        fitness = (det_pr + numpy.float32(0.25) * numpy.random.RandomState(seed).randn(*det_pr.shape)
                   .astype(numpy.float32)).astype(numpy.float32)
        lists = det.build_detections_nms(pr_thr, nms_thr, soft, det_pr, fitness, bbox, num)
        n = max(len(d) for d in lists)
        score = numpy.zeros((B, n), numpy.float32)
        cls = numpy.full((B, n), -1, numpy.int32)
        box = numpy.zeros((B, n, 4), numpy.float32)
        for b, d in enumerate(lists):
            for i, (p, c, bb) in enumerate(d):
                score[b, i], cls[b, i], box[b, i] = p, c, bb
        numpy.savez_compressed(os.path.join(HERE, name + ".npz"), det_pr=det_pr, fitness=fitness, bbox=bbox,
                               bbox_num=numpy.array(num, numpy.int32), pr_threshold=numpy.float32(pr_thr),
                               nms_threshold=numpy.float32(nms_thr), use_soft_nms=soft, score=score, cls=cls, box=box,
                               count=numpy.array([len(d) for d in lists], numpy.int32))
        print(name, "detections per image", [len(d) for d in lists])


if __name__ == "__main__":
    main()
