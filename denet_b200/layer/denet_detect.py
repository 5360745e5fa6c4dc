"""DeNet detect layer 'DND[overlap_thr,cost_factor,bbox_factor,fitness_factor]'
(reference denet/layer/denet_detect.py:24-313; training path only).

A zero-initialised 1x1 convolution with bias classifies every RoI into classNum+1 classes (+4 box regression
outputs when bbox_factor > 0).  get_target assigns classes / box targets by IoU against the ground truth
(:147-235); the cost is -sum(t*logp)/ln(s0) per RoI summed / batch * cost_factor, plus the Fast R-CNN smooth-L1 box
loss (:238-313, bbox_factor applied in :295 and again in :310).  get_detections (:316-424) is the inference tail:
one test-mode forward, device sampler, head, per-class NMS on the GPU (csrc/detect_nms.cu).

v2 variants (the "J" / "B" tags and the fourth parameter of DND): joint fitness (:58-61, 179-182, 332-348) - every
class splits into 5 IoU-fitness bins, classNum*5+1 outputs; independent fitness (:100-104, 187-191, 297-299,
392-397) - a second 6-way softmax head; bounded-IoU box loss (:266-286) on the decoded box.
"""
import numpy
import torch

from .. import common, ops
from . import AbstractLayer, InitialLayer, d2h, device_targets, get_ground_truth, get_train, h2d
from .convolution import ConvLayer


def overlap_iou_matrix(obj_bboxs, sample_bboxs):
    """IoU matrix in float32 with the operation order of the reference's compiled Theano function
    (common/theano_util.py:38-59)"""
    x = numpy.asarray(obj_bboxs, dtype=numpy.float32)
    y = numpy.asarray(sample_bboxs, dtype=numpy.float32)
    x_area = (x[:, 2] - x[:, 0]) * (x[:, 3] - x[:, 1])
    y_area = (y[:, 2] - y[:, 0]) * (y[:, 3] - y[:, 1])
    dx = numpy.maximum(numpy.minimum(x[:, None, 2], y[None, :, 2]) - numpy.maximum(x[:, None, 0], y[None, :, 0]), 0)
    dy = numpy.maximum(numpy.minimum(x[:, None, 3], y[None, :, 3]) - numpy.maximum(x[:, None, 1], y[None, :, 1]), 0)
    inter = dx * dy
    union = x_area[:, None] + y_area[None, :] - inter
    with numpy.errstate(divide="ignore", invalid="ignore"):
        return inter / union


class DeNetDetectLayer(AbstractLayer):
    type_name = "denet-detect"
    has_cost = True

    def __init__(self, layers, class_num=10, overlap_threshold=0.5, cost_factor=1.0, bbox_factor=0.0, indfit_factor=0.0,
                 use_jointfit=False, use_bounded_iou=False, json_param={}):
        super().__init__(layer_index=len(layers))
        self.input = layers[-1].output
        self.input_shape = self.output_shape = tuple(layers[-1].output_shape)

        self.cost_factor = json_param.get("costFactor", cost_factor)
        self.bbox_factor = json_param.get("bboxFactor", bbox_factor)
        self.class_num = json_param.get("classNum", class_num)
        self.overlap_threshold = json_param.get("overlapThreshold", overlap_threshold)
        self.use_jointfit = json_param.get("useJointFitness", use_jointfit)
        self.use_bounded_iou = json_param.get("useBoundedIoU", use_bounded_iou)
        self.indfit_factor = json_param.get("fitnessFactor", indfit_factor)
        self.use_indfit = self.indfit_factor > 0.0
        assert not (self.use_indfit and self.use_jointfit), "Cannot enable both fitness methods at once!"

        sparse_layer = common.find_layers(layers, "denet-sparse", False)
        assert sparse_layer is not None, "Error: Requires denet-sparse layer to be specified before denet-detect layer!"
        object.__setattr__(self, "sparse_layer", sparse_layer)

        self.use_bbox_reg = self.bbox_factor > 0.0
        self.batch_size = sparse_layer.batch_size
        self.sample_num = sparse_layer.sample_num
        if self.use_jointfit:                   # :58-66
            self.fitness_num = 5
            self.null_class = self.class_num * self.fitness_num
        else:
            self.fitness_num = 6
            self.null_class = self.class_num
        s0 = self.null_class + 1
        s1 = 4 if self.use_bbox_reg else 0
        s2 = self.fitness_num if self.use_indfit else 0
        conv = ConvLayer([InitialLayer(None, self.input_shape)], (s0 + s1 + s2, self.input_shape[1], 1, 1), (1, 1), True,
                         "valid", 0.0)
        conv.out_fp32 = True
        self.layers.append(conv)
        self.det_shape = (self.batch_size, s0, self.sample_num, self.sample_num)
        if self.use_bbox_reg:
            self.bbox_shape = (self.batch_size, s1, self.sample_num, self.sample_num)
        if self.use_indfit:
            self.indfit_shape = (self.batch_size, s2, self.sample_num, self.sample_num)
        self.fit_mode = (1 if self.use_jointfit else 0) | (2 if self.use_indfit else 0)
        self.box_mode = 0 if not self.use_bbox_reg else (2 if self.use_bounded_iou else 1)
        self.grad_factor = 1.0
        self.cost_value = None     # device tensor [detection, box, fitness cost] of the last training forward
        self._targets = None
        self._targets_dev = None    # persistent target buffers filled by denet_detect_target
        self._dout = None
        self.logits = None

    @staticmethod
    def parse_desc(layers, name, tags, params):
        if name != "DND":
            return False
        layers.append(DeNetDetectLayer(layers, params.get("classNum"), params.get(0, 0.5), params.get(1, 1.0),
                                       params.get(2, 0.0), params.get(3, 0.0), "J" in tags, "B" in tags))
        return True

    def import_json(self, json_param):
        super().import_json(json_param)
        if "conv" in json_param:   # backward compatibility (denet_detect.py:127-129)
            self.layers[0].import_json(json_param["conv"])

    def export_json(self):
        json = super().export_json()
        json.update({"costFactor": self.cost_factor, "bboxFactor": self.bbox_factor,
                     "fitnessFactor": self.indfit_factor, "useJointFitness": self.use_jointfit,
                     "useBoundedIoU": self.use_bounded_iou, "classNum": self.class_num,
                     "overlapThreshold": self.overlap_threshold})
        return json

    def _thresholds(self):
        # the reference indexes overlap_threshold[0] / [1] although parse_desc passes a scalar: accept both
        t = self.overlap_threshold
        return (t[0], t[1]) if isinstance(t, (tuple, list)) else (t, t)

    def get_target(self, model, samples, metas):
        """denet_detect.py:147-235.  With device targets on, denet_detect_target builds the same arrays inside forward()
        from the ground-truth boxes and the RoIs already in HBM, and this returns None."""
        if device_targets():
            self._targets = None
            return None
        return self.get_target_host(metas)

    def get_target_host(self, metas):
        """denet_detect.py:147-235 on the host arrays of the sparse layer (no per-RoI python tuples)"""
        thr0, thr1 = self._thresholds()
        sn = self.sample_num
        det_pr = numpy.zeros(self.det_shape, dtype=numpy.float32)
        det_pr[:, self.null_class] = 1.0
        if self.use_bbox_reg:
            bbox_valid = numpy.zeros((self.batch_size, sn, sn), dtype=numpy.float32)
            bbox_reg = numpy.zeros((self.batch_size, 8, sn, sn), dtype=numpy.float32)
            bbox_reg[:, [2, 3, 6, 7]] = 1.0
        if self.use_indfit:
            indfit_pr = numpy.zeros(self.indfit_shape, dtype=numpy.float32)
            indfit_pr[:, 0] = 1.0
        all_samples = self.sparse_layer.sample_bbox_host            # (B,K,4) float64
        for b, meta in enumerate(metas):
            samples = all_samples[b]
            if len(meta["bbox"]) == 0 or len(samples) == 0:
                continue
            overlap = overlap_iou_matrix(meta["bbox"], samples)
            bbox_indexs, sample_indexs = numpy.where(overlap > thr0)
            if len(bbox_indexs) > 0:
                cls = numpy.asarray(meta["class"], dtype=numpy.int64)[bbox_indexs]
                if self.fit_mode:
                    # :177 in float64: the reference mixes a numpy float32 scalar with python floats (numpy 1.x promotes
                    # that to float64)
                    sample_f = (overlap[bbox_indexs, sample_indexs].astype(numpy.float64) - thr0) / (1.0 - thr0)
                if self.use_jointfit:               # :179-182 (int() truncates toward zero)
                    f = numpy.clip(numpy.trunc(self.fitness_num * sample_f).astype(numpy.int64), 0, self.fitness_num - 1)
                    cls = cls * self.fitness_num + f
                det_pr[b, cls, sample_indexs // sn, sample_indexs % sn] = 1.0
                det_pr[b, self.null_class, sample_indexs // sn, sample_indexs % sn] = 0.0
                if self.use_indfit:                 # :187-191
                    f = 1 + numpy.floor((self.fitness_num - 1) * sample_f).astype(numpy.int64)
                    f = numpy.clip(f, 1, self.fitness_num - 1)
                    indfit_pr[b, 0, sample_indexs // sn, sample_indexs % sn] = 0.0
                    indfit_pr[b, f, sample_indexs // sn, sample_indexs % sn] = 1.0
            if self.use_bbox_reg:
                overlap_max = overlap.argmax(axis=0)
                index = numpy.arange(len(samples))
                sel = index[overlap[overlap_max, index] > thr1]
                if len(sel) > 0:
                    target = numpy.asarray(meta["bbox"], dtype=numpy.float64)[overlap_max[sel]]
                    sample = samples[sel]
                    sj, si = sel // sn, sel % sn
                    bbox_valid[b, sj, si] = 1.0
                    bbox_reg[b, 0, sj, si] = 0.5 * (target[:, 0] + target[:, 2])
                    bbox_reg[b, 1, sj, si] = 0.5 * (target[:, 1] + target[:, 3])
                    bbox_reg[b, 2, sj, si] = target[:, 2] - target[:, 0]
                    bbox_reg[b, 3, sj, si] = target[:, 3] - target[:, 1]
                    bbox_reg[b, 4, sj, si] = 0.5 * (sample[:, 0] + sample[:, 2])
                    bbox_reg[b, 5, sj, si] = 0.5 * (sample[:, 1] + sample[:, 3])
                    bbox_reg[b, 6, sj, si] = sample[:, 2] - sample[:, 0]
                    bbox_reg[b, 7, sj, si] = sample[:, 3] - sample[:, 1]
        det_pr /= det_pr.sum(axis=1)[:, None]
        nfactor = sn * sn
        det_pr /= nfactor
        yt_value = det_pr.flatten()
        if self.use_bbox_reg:
            bbox_valid /= nfactor
            yt_value = numpy.concatenate((yt_value, bbox_valid.flatten(), bbox_reg.flatten()))
        if self.use_indfit:
            indfit_pr /= indfit_pr.sum(axis=1)[:, None]
            indfit_pr /= nfactor
            yt_value = numpy.concatenate((yt_value, indfit_pr.flatten()))
        return numpy.array([], dtype=numpy.int64), yt_value

    def set_target(self, yt_index, yt_value):
        v = h2d(numpy.ascontiguousarray(yt_value, dtype=numpy.float32))
        n0 = int(numpy.prod(self.det_shape))
        n1 = self.batch_size * self.sample_num * self.sample_num
        off = n0
        t_valid = t_reg = t_fit = None
        if self.use_bbox_reg:
            t_valid, t_reg = v[n0:n0 + n1], v[n0 + n1:n0 + 9 * n1]
            off += 9 * n1
        if self.use_indfit:
            t_fit = v[off:off + self.fitness_num * n1]
        self._targets = (v[:n0], t_valid, t_reg, t_fit)

    def forward(self, x):
        self.input = self.output = x
        o = self.layers[0].forward(x)            # (B,sn,sn,s0+s1) fp32
        self.logits = o
        if self.cost_value is None:
            self.cost_value = torch.zeros((3,), dtype=torch.float32, device=x.device)
        if get_train():
            if device_targets():
                if self._targets_dev is None:
                    dev, sn = x.device, self.sample_num
                    self._targets_dev = (
                        torch.empty(self.det_shape, dtype=torch.float32, device=dev),
                        torch.empty((self.batch_size, sn, sn), dtype=torch.float32, device=dev) if self.use_bbox_reg
                        else None,
                        torch.empty((self.batch_size, 8, sn, sn), dtype=torch.float32, device=dev) if self.use_bbox_reg
                        else None,
                        torch.empty(self.indfit_shape, dtype=torch.float32, device=dev) if self.use_indfit else None)
                thr0, thr1 = self._thresholds()
                t_det, t_valid, t_reg, t_fit = self._targets_dev
                ops.detect_target(get_ground_truth(), self.sparse_layer.sample_bbox64, self.sample_num, self.class_num,
                                  thr0, thr1, self.use_bbox_reg, t_det, t_valid, t_reg, fit_mode=self.fit_mode, fit=t_fit)
                self._targets = self._targets_dev
            assert self._targets is not None, "denet-detect: get_target/set_target must precede a training forward"
            self._dout = ops.alloc_like(o)
            t_det, t_valid, t_reg, t_fit = self._targets
            ops.detect_cost(o, self.sample_num, self.det_shape[1], self.box_mode, t_det, t_valid, t_reg,
                            float(self.cost_factor), float(self.bbox_factor), self.grad_factor, self._dout,
                            self.cost_value, nfit=self.fitness_num if self.use_indfit else 0, target_fit=t_fit,
                            fit_factor=float(self.indfit_factor),
                            sample_bbox=self.sparse_layer.sample_bbox if self.box_mode == 2 else None)
        return x

    def cost(self, yt_index=None, yt_value=None):
        return None if self.cost_value is None else self.cost_value.sum()

    def last_target(self):
        """(yt_index, yt_value) of the last training forward in the reference's flattened format (:229-235)"""
        parts = [t.reshape(-1) for t in self._targets if t is not None]
        return numpy.array([], dtype=numpy.int64), d2h(torch.cat(parts))

    def backward(self, dy):
        dx = self.layers[0].backward(self._dout)
        self._dout = None
        return dx if dy is None else ops.add(dx, dy)

    # ---------------------------------------------------------------------------------------------- inference
    def get_detections(self, model, data_x, data_m, params):
        """reference get_detections (denet_detect.py:316-424): most likely (class, box) instances of every image.

        params: prThreshold (0.01), nmsThreshold (0.5), cornerThreshold (the sparse layer's), cornerMax (1024),
        useSoftNMS (0).  Returns [{"detections": [(pr, cls, (x0, y0, x1, y1)), ...], "meta": data_m[i]}].

        The reference compiles two Theano functions (corner maps, then the head on cached features); here ONE
        test-mode forward pass pauses at the sparse layer: the device sampler ranks the RoIs from the corner maps of
        that pass, the head classifies them, denet_detect_outputs produces log-probabilities + decoded boxes and
        denet_detections_nms runs the per-class NMS - only the final detection lists leave the GPU."""
        from . import denet_detect_c
        pr_threshold = params.get("prThreshold", 0.01)
        nms_threshold = params.get("nmsThreshold", 0.5)
        sp = self.sparse_layer
        corner_threshold = params.get("cornerThreshold", sp.corner_threshold)
        corner_max = params.get("cornerMax", 1024)
        use_soft_nms = params.get("useSoftNMS", 0) == 1
        det_pr, fitness, bboxs, counts = model.detect_forward(data_x, self, corner_threshold, corner_max)
        detlists = denet_detect_c.build_detections_nms(pr_threshold, nms_threshold, int(use_soft_nms), det_pr, fitness,
                                                       bboxs, counts)
        data_m = data_m if data_m is not None else [None] * len(detlists)
        return [{"detections": detlist, "meta": data_m[i]} for i, detlist in enumerate(detlists)]

    def detect_outputs(self):
        """(det_pr (B,classNum+1,sn,sn), fitness, bbox (B,sn,sn,4)) device tensors of the last test-mode forward
        (the outputs of the reference's detect_func, denet_detect.py:330-362, with the host arithmetic of :378-397 folded
        in; fitness = det_pr without a fitness head)"""
        thr0, _ = self._thresholds()
        return ops.detect_outputs(self.logits, self.sample_num, self.det_shape[1], self.use_bbox_reg,
                                  self.sparse_layer.sample_bbox, class_num=self.class_num, fit_mode=self.fit_mode,
                                  thr0=thr0)
