"""DeNet sparse layer 'DNS[grid,sample_num,corner_thr,random_sample,local_max,nms_thr]'
(reference denet/layer/denet_sparse.py:26-219).

get_target() turns the corner maps of THIS forward pass into the RoI set: the reference re-runs the whole backbone in a
second compiled function, copies corner_pr to the host and calls the C++ extension build_samples (:117-145); here
the maps stay in HBM and denet_build_samples ranks the boxes on the device (csrc/build_samples.cu).  Only the ranked
RoIs (a few hundred KB) visit the host, where the reference's python `random` post-processing runs unchanged in
meaning AND in random-stream consumption (:184-201): sub-sample to make room for random boxes, pad with random boxes,
overwrite the tail with the ground truth.  forward() is the sparse RoI feature gather (DeNetSparseOp,
denet_sparse_op.py:42-85), backward() its scatter-add (:171-212).
"""
import ctypes
import math
import random

import numpy
import torch

from .. import common, ops
from ..lib import call
from . import AbstractLayer, act_dtype, d2h, get_train, h2d


def py_random_doubles(k):
    """the next k values of python's random.random(), advanced exactly as k calls would (vectorised through numpy's
    MT19937, which shares CPython's generator and its 53-bit double construction)"""
    if k <= 0:
        return numpy.empty((0,), dtype=numpy.float64)
    if k < 32:
        return numpy.array([random.random() for _ in range(k)], dtype=numpy.float64)
    version, internal, gauss = random.getstate()
    rs = numpy.random.RandomState()
    rs.set_state(("MT19937", numpy.array(internal[:-1], dtype=numpy.uint32), int(internal[-1])))
    out = rs.random_sample(k)
    st = rs.get_state()
    random.setstate((version, tuple(st[1].tolist()) + (int(st[2]),), gauss))
    return out


def mt_export():
    """python `random` state -> (624-word numpy array, c_int position, version, gauss_next) for the native helpers"""
    version, internal, gauss = random.getstate()
    arr = numpy.array(internal, dtype=numpy.uint32)
    return numpy.ascontiguousarray(arr[:624]), ctypes.c_int(int(arr[624])), version, gauss


def mt_import(mt, pos, version, gauss):
    random.setstate((version, tuple(mt.tolist()) + (int(pos.value),), gauss))


class RandomAhead:
    """python `random`'s Mersenne-Twister stepped ahead of time (denet_pyrandom_ahead): tempered words + the state after
    every regeneration, generated while the GPU is busy, consumed by finish_target between the two graphs"""

    def __init__(self, nwords_needed):
        version, internal, gauss = random.getstate()
        self.version, self.gauss = version, gauss
        self.internal_head = internal[:8] + (internal[624],)
        mt = numpy.array(internal[:624], dtype=numpy.uint32)
        self.pos0 = int(internal[624])
        self.mt0 = mt
        nblocks = max(1, -(-max(0, nwords_needed - (624 - self.pos0)) // 624) + 1)
        self.states = numpy.empty((nblocks, 624), dtype=numpy.uint32)
        self.words = numpy.empty(((624 - self.pos0) + 624 * nblocks,), dtype=numpy.uint32)
        n = ctypes.c_longlong(0)
        call("denet_pyrandom_ahead", mt.ctypes.data, self.pos0, nblocks, self.states.ctypes.data,
             self.words.ctypes.data, ctypes.addressof(n))
        self.nwords = int(n.value)

    def still_valid(self):
        """nobody consumed python's `random` since the words were generated"""
        internal = random.getstate()[1]
        return internal[:8] + (internal[624],) == self.internal_head

    def commit(self, used):
        """advance the interpreter's generator by `used` words"""
        first = 624 - self.pos0
        if used <= first:
            mt, pos = self.mt0, self.pos0 + used
        else:
            u = used - first
            j = (u - 1) // 624
            mt, pos = self.states[j], u - 624 * j
        random.setstate((self.version, tuple(mt.tolist()) + (int(pos),), self.gauss))


def py_random_sample(n, k):
    """random.sample(range(n), k), natively, consuming the interpreter's stream identically"""
    mt, pos, version, gauss = mt_export()
    out = numpy.empty((k,), dtype=numpy.int32)
    call("denet_pyrandom_sample", mt.ctypes.data, ctypes.addressof(pos), int(n), int(k), out.ctypes.data)
    mt_import(mt, pos, version, gauss)
    return out


class DeNetSparseLayer(AbstractLayer):
    type_name = "denet-sparse"

    def __init__(self, layers, grid_size=3, sample_num=16, corner_threshold=0.01, random_sample=0.0, local_max=0,
                 nms_threshold=0.7, sample_gt=True, version="v2", json_param={}):
        super().__init__(layer_index=len(layers))
        self.input = layers[-1].output
        self.input_shape = tuple(layers[-1].output_shape)
        self.batch_size = self.input_shape[0]

        self.grid_size = json_param.get("gridSize", grid_size)
        self.sample_num = json_param.get("sampleNum", sample_num)
        self.sample_gt = json_param.get("sampleGT", sample_gt)
        self.corner_threshold = json_param.get("cornerThreshold", corner_threshold)
        self.nms_threshold = json_param.get("nmsThreshold", nms_threshold)
        self.random_sample = json_param.get("randomSample", random_sample)
        self.local_max = json_param.get("localMax", local_max)
        self.version = json_param.get("version", version)

        self.corner_max = 1024
        self.thread_num = self.batch_size
        self.sample_count = self.sample_num * self.sample_num

        corner_layer = common.find_layers(layers, "denet-corner", True)     # 4 corner maps, 5 with centres (DNC.C)
        object.__setattr__(self, "corner_layer", corner_layer)

        self.sample_bbox = None           # (B,sn,sn,4) fp32 device tensor consumed by the gather kernel
        self.sample_bbox64 = None         # (B,sn*sn,4) float64 device copy for denet_detect_target
        self.sample_bbox_host = None      # (B,sn*sn,4) float64: what the reference keeps as python floats
        self.sample_pr_host = None        # (B,sn*sn)   float64
        self._sample_bbox_list = None
        self._packed_dev = None
        self.output_feat = self.grid_size * self.grid_size * corner_layer.sample_shape[1] + 2
        self.output_shape = (self.batch_size, self.output_feat, self.sample_num, self.sample_num)

    @staticmethod
    def parse_desc(layers, name, tags, params):
        if name != "DNS":
            return False
        layers.append(DeNetSparseLayer(layers, params.get(0, 3), params.get(1, 4), params.get(2, 0.01),
                                       params.get(3, 0.1), params.get(4, 0), params.get(5, 1.0), "G" not in tags))
        return True

    def export_json(self):
        json = super().export_json()
        json.update({"gridSize": self.grid_size, "sampleNum": self.sample_num, "sampleGT": self.sample_gt,
                     "localMax": self.local_max, "cornerThreshold": self.corner_threshold,
                     "randomSample": self.random_sample, "nmsThreshold": self.nms_threshold,
                     "version": self.version})
        return json

    # ---------------------------------------------------------------------------------------------- sampling
    def enqueue_samples(self, corner_pr=None):
        """enqueue the device sampler on the corner maps of this forward pass; results land in persistent device
        buffers packed as [pr (B,K) | bbox (B,K,4) | count (B)] fp32 (capturable in a CUDA graph: no host access)"""
        if corner_pr is None:
            corner_pr = self.corner_layer.corner_pr
        assert corner_pr is not None, "denet-sparse: the corner layer has not run a forward pass yet"
        b, k = self.batch_size, self.sample_count
        if self._packed_dev is None:
            self._packed_dev = torch.empty((5 * b * k + b,), dtype=torch.float32, device=corner_pr.device)
        p = self._packed_dev        # the sampler writes straight into the packed buffer (counts as int32 bit patterns)
        ops.build_samples(corner_pr, self.corner_threshold, self.sample_num, self.corner_max, self.local_max,
                          self.nms_threshold, out=(p[:b * k].view(b, k), p[b * k:5 * b * k].view(b, k, 4),
                                                   p[5 * b * k:].view(torch.int32)))
        return self._packed_dev

    def collect_samples(self):
        """device -> host copy of the packed sampler output (one small synchronising copy):
        pr (B,K) f32, bbox (B,K,4) f32, count (B)"""
        packed = d2h(self._packed_dev, slot="denet-sparse/samples" + self._slot_ns)
        b, k = self.batch_size, self.sample_count
        return (packed[:b * k].reshape(b, k), packed[b * k:5 * b * k].reshape(b, k, 4),
                packed[5 * b * k:].view(numpy.int32).astype(numpy.int64))

    def get_samples_arrays(self, corner_pr=None):
        self.enqueue_samples(corner_pr)
        return self.collect_samples()

    def get_samples(self, data_x=None, train=False, store_shared=False):
        """reference return format (denet_sparse.py:117-145): per image a list of (pr, (x0,y0,x1,y1))"""
        pr, bbox, count = self.get_samples_arrays()
        return [[(float(pr[b, i]), tuple(float(v) for v in bbox[b, i])) for i in range(int(count[b]))]
                for b in range(self.batch_size)]

    @property
    def sample_bbox_list(self):
        if self._sample_bbox_list is None and self.sample_bbox_host is not None:
            self._sample_bbox_list = [[(float(self.sample_pr_host[b, i]), tuple(float(v) for v in
                                                                           self.sample_bbox_host[b, i]))
                                       for i in range(self.sample_count)] for b in range(self.batch_size)]
        return self._sample_bbox_list

    def set_samples_arrays(self, pr, bbox):
        """pr (B,K) / bbox (B,K,4) float64 host arrays; row i of image b lands at (i // sn, i % sn)
        (build_bbox_array, denet_sparse.cc:670-699)"""
        self.sample_pr_host, self.sample_bbox_host = pr, bbox
        self._sample_bbox_list = None
        arr = numpy.ascontiguousarray(bbox.astype(numpy.float32).reshape(self.batch_size, self.sample_num,
                                                                         self.sample_num, 4))
        self.sample_bbox = h2d(arr, slot="denet-sparse/bbox32" + self._slot_ns)
        # the doubles feed the device-side detection targets (python floats in the reference, denet_detect.py:200-212)
        self.sample_bbox64 = h2d(numpy.ascontiguousarray(bbox, dtype=numpy.float64), slot="denet-sparse/bbox64" + self._slot_ns)
        return arr

    def set_samples(self, sample_bboxs):
        """reference entry point: list (per image) of lists of (pr, bbox)"""
        k = self.sample_count
        pr = numpy.zeros((self.batch_size, k), dtype=numpy.float64)
        bbox = numpy.zeros((self.batch_size, k, 4), dtype=numpy.float64)
        for b, samples in enumerate(sample_bboxs):
            for i, (p, bb) in enumerate(samples[:k]):
                pr[b, i] = p
                bbox[b, i] = bb
        return self.set_samples_arrays(pr, bbox)

    def sample_for_inference(self):
        """test-time sampling entirely on the device: the ranked RoIs of the corner maps of this pass become the sample
        boxes (slots past the per-image count are zero boxes, like the reference's zero-initialised bbox array,
        denet_sparse.py:151-157); returns the per-image counts (B) int32 on the device"""
        pr, bbox, _, count, _ = ops.build_samples(self.corner_layer.corner_pr, self.corner_threshold, self.sample_num,
                                                  self.corner_max, self.local_max, self.nms_threshold)
        b, k = self.batch_size, self.sample_count
        valid = (torch.arange(k, device=bbox.device)[None, :] < count[:, None]).unsqueeze(-1)
        self.sample_bbox = torch.where(valid, bbox, torch.zeros_like(bbox)).reshape(b, self.sample_num, self.sample_num,
                                                                                   4).contiguous()
        self.sample_bbox64 = None
        self.sample_bbox_host = self.sample_pr_host = self._sample_bbox_list = None
        return count

    def get_target(self, model, data_x, metas):
        """denet_sparse.py:164-206, vectorised; consumes python's `random` stream exactly like the reference loops"""
        return self.finish_target(metas, *self.get_samples_arrays())

    def random_ahead(self):
        """generate the random words one step can consume at most, ahead of time (call while the GPU is busy)"""
        k = self.sample_count
        return RandomAhead(self.batch_size * (8 * k + 4 * k) + 1248)

    def finish_target(self, metas, pr32, bbox32, count, ahead=None):
        """host half of get_target: the reference's python-`random` post-processing of the ranked RoIs, then the
        upload of the final (B,sn,sn,4) box tensor.  ahead: a RandomAhead generated for this step (same stream, the
        generator work already done)"""
        k = self.sample_count
        n_keep = k - math.floor(self.random_sample * k)
        pr = numpy.zeros((self.batch_size, k), dtype=numpy.float64)
        bbox = numpy.zeros((self.batch_size, k, 4), dtype=numpy.float64)
        nb = len(metas)
        count = numpy.ascontiguousarray(count[:nb], dtype=numpy.int64)
        pr32 = numpy.ascontiguousarray(pr32[:nb], dtype=numpy.float32)
        bbox32 = numpy.ascontiguousarray(bbox32[:nb], dtype=numpy.float32)
        # sub-sample / pad with python-`random` semantics for the whole batch in one native call
        # (csrc/pyrandom.cu: the interpreter's Mersenne-Twister state goes in and comes back advanced exactly as the
        # reference's loops over random.sample / random.uniform would have advanced it)
        done = False
        if ahead is not None and ahead.still_valid():
            used = ctypes.c_longlong(0)
            rc = call("denet_sparse_postprocess_ahead", ahead.words.ctypes.data, ahead.nwords, ctypes.addressof(used),
                      pr32.ctypes.data, bbox32.ctypes.data, count.ctypes.data, nb, k, n_keep, pr.ctypes.data,
                      bbox.ctypes.data, allow=(1,))
            if rc == 0:
                ahead.commit(int(used.value))
                done = True
        if not done:
            mt, pos, version, gauss = mt_export()
            call("denet_sparse_postprocess", mt.ctypes.data, ctypes.addressof(pos), pr32.ctypes.data, bbox32.ctypes.data,
                 count.ctypes.data, nb, k, n_keep, pr.ctypes.data, bbox.ctypes.data)
            mt_import(mt, pos, version, gauss)
        if self.sample_gt:
            for b, meta in enumerate(metas):
                if len(meta["bbox"]) > k:
                    # the reference's list assignment raises here too (denet_sparse.py:199-201); numpy would wrap the
                    # negative index and let ground-truth entries overwrite each other
                    raise IndexError("denet-sparse: image %d has %d ground-truth boxes but only %d samples"
                                     % (b, len(meta["bbox"]), k))
                for index, gt in enumerate(meta["bbox"]):       # the LAST len(GT) slots become the ground truth
                    pr[b, k - (index + 1)] = 1.0
                    bbox[b, k - (index + 1)] = gt
        self.set_samples_arrays(pr, bbox)
        return None

    # ---------------------------------------------------------------------------------------------- execution
    def forward(self, x):
        self.input = x
        assert self.sample_bbox is not None, "denet-sparse: get_target()/set_samples() must precede forward()"
        fmap = self.corner_layer.sample
        self._fmap_shape = tuple(fmap.shape)
        self.output = ops.sparse_sample_fwd(fmap, self.sample_bbox, self.grid_size, out_dtype=act_dtype())
        return self.output

    def backward(self, dy):
        dfmap = ops.sparse_sample_bwd(dy, self.sample_bbox, self.grid_size, self._fmap_shape)
        self.corner_layer.set_sample_grad(dfmap)
        return None   # the reference op has no gradient wrt the boxes, and the layer input is not used (:38)
