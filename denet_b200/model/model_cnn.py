"""ModelCNN: the reference's model container and training step (denet/model/model_cnn.py:86-571) on the B200 kernels.

Same driver-facing surface: initialize(args, ...), ModelCNN.build / build_layer (model-desc grammar TYPE.TAGS[ARGS],
model_cnn.py:122-157), export_json / import_json (+ load_from_file / save_to_file, gz-JSON checkpoints),
build_train_func(solver_mode, cost_factors), train_step(data_x, data_m, epoch, it, lr, momentum, decay) ->
(cost, [costs]), train_epoch, predict_output_step / predict_output / predict_label.

What replaces theano.function + tensor.grad: ONE eager forward pass over the layer list (the reference runs the
backbone twice per iteration, once in DeNetSparseLayer.corner_func and once in train_step; both in train mode with
identical batch statistics, so one pass that pauses at the sparse layer for sampling is result-equivalent), an
explicit reverse pass calling each layer's backward(), and one multi-tensor solver kernel
(model_cnn.py:282-305: sgd / torch|nesterov / adam; L2 decay on weights only, :320-324).  With torch.distributed
initialised the gradients are all-reduced over NCCL in buckets that overlap the remaining backward pass
(denet_b200/multi), replacing denet/multi's host-side parameter averaging.
"""
import ctypes
import getpass
import math
import os
import random
import time

import numpy
import torch

from .. import common, layer as layer_mod, lib, ops
from ..common import json_util
from ..layer import InitialLayer
from ..layer.layer_types import layer_types

SOLVER_CODES = {"sgd": 0, "torch": 1, "nesterov": 1, "adam": 2}


def load_from_json(json_obj, batch_size=32, layer_range=None):
    model = ModelCNN()
    model.batch_size = batch_size
    model.import_json(json_obj, layer_range)
    return model


def load_from_file(fname, batch_size=32, layer_range=None):
    model = load_from_json(json_util.json_from_gz(fname), batch_size, layer_range)
    model.fname = fname
    return model


def save_to_file(model, fname, compresslevel=9):
    json_util.json_to_gz(fname, model.export_json(), compresslevel)


def initialize(args, data_shape, class_labels, class_num):
    """reference model_cnn.initialize (:46-83)"""
    if args.model is None:
        model = ModelCNN()
        model.batch_size = args.batch_size
        model.class_labels = class_labels
        model.class_num = class_num
        try:
            n = int(args.border_mode)
            border_mode = (n, n)
        except ValueError:
            border_mode = args.border_mode
        model.build(args.model_desc, data_shape, args.activation, border_mode, list(args.weight_init))
    else:
        model = load_from_file(args.model, args.batch_size)
        model.class_labels = class_labels
        model.class_num = class_num
        assert tuple(data_shape) == tuple(model.data_shape), "Mismatching data shapes in .mdl and data: " + \
            str(data_shape) + "!=" + str(model.data_shape)
    model.skip_layer_updates = getattr(args, "skip_layer_updates", [])
    return model


def _walk(layers):
    """depth-first over a layer list and its sub-layers"""
    for l in layers:
        yield l
        yield from _walk(l.layers)


def link_fusions(layers):
    """conv -> batch-norm pairs: the conv epilogue accumulates the statistics its batch-norm consumes"""
    from ..layer.batch_norm import BatchNormLayer
    from ..layer.convolution import ConvLayer
    bns = []
    for seq in [layers] + [l.layers for l in _walk(layers)]:
        seq = list(seq)
        for a, b in zip(seq[:-1], seq[1:]):
            if isinstance(a, ConvLayer) and isinstance(b, BatchNormLayer) and b.enabled and a.enabled \
                    and not a.out_fp32:
                object.__setattr__(a, "stat_consumer", b)
                bns.append(b)
    return bns


def _bn_consumer(prev):
    """the batch-norm layer whose backward directly consumes the gradient wrt `prev`'s output, or None"""
    from ..layer.batch_norm import BatchNormLayer
    from ..layer.resnet import ResnetLayer
    if isinstance(prev, BatchNormLayer):
        return prev if prev.enabled else None
    if isinstance(prev, ResnetLayer):
        return prev.fusable_last_bn()
    return None


class ModelCNN:

    def __init__(self):
        self.batch_size = 0
        self.iteration = 0
        self.class_labels = None
        self.data_shape = None
        self.class_num = 0
        self.rng_seed = random.randint(1, 9999)
        layer_mod.set_rng_seed(self.rng_seed)

        self.gradient_clip = 0.0
        self.skip_layer_updates = []
        self.bias_decay = False
        self.layers = []
        self.distort_mode = []
        self.func = {}
        self.input = None
        self.device = None
        self.ddp = None             # denet_b200.multi.GradientAllReduce when running data parallel
        self.defer_wgrad_reduce = True   # one multi-tensor split-K reduction launch instead of one per conv layer
        # dgrad epilogues take over the reduction pass of the batch-norm backward (layer.set_fuse_bn_backward; the
        # environment variable DENET_FUSE_BN_BWD=0/1 overrides the default for A/B measurements)
        if os.environ.get("DENET_FUSE_BN_BWD"):
            layer_mod.set_fuse_bn_backward(os.environ["DENET_FUSE_BN_BWD"] != "0")
        self.fuse_bn_backward = True
        # filter gradients on a second stream, concurrent with the data-gradient / batch-norm chain (A/B: DENET_OVERLAP_WGRAD)
        self.overlap_wgrad = os.environ.get("DENET_OVERLAP_WGRAD", "0") != "0"
        self._wgrad_stream = None
        self._ready = False
        self.last_costs_device = None
        self._image = None          # padded input buffer of a row-folded stem conv
        self._static_inputs = False # copy every batch into one persistent device buffer (needed by CUDA graphs)
        self._graphs = None         # captured CUDA graphs of the training step (enable_cuda_graphs)
        self._prep_table = None     # device table of the one-launch conv operand preparation
        self._prep_version = -1
        self._prep_precision = None
        self._prep_layers = []
        self._use_graphs = False
        self._copy_stream = None       # graphed steps: image upload / cost download off the compute stream
        self._img_consumed = self._cost_ready = self._cost_copied = None
        self._img_consumed_valid = False
        self._costs_pinned = None
        self._host_costs = None
        self._slot_ns = layer_mod.new_slot_namespace()     # staging buffers (image, ground truth, scalars) of THIS model

    # ---------------------------------------------------------------------------------------------- shapes
    def get_input_shape(self):
        assert self.data_shape is not None, "Data shape hasn't been set!"
        return tuple([self.batch_size] + list(self.data_shape))

    def get_output_shape(self):
        return self.layers[-1].output_shape

    def get_parameter_num(self):
        n = 0
        for layer in self.layers:
            for param in layer.params():
                n += param.numel()
        return n

    # ---------------------------------------------------------------------------------------------- building
    def _initial_layer(self):
        init = InitialLayer(None, self.get_input_shape())
        init.is_model_input = True
        return init

    def build_layer(self, layer_desc, layers, activation, border_mode, wb):
        """TYPE.TAGS[a,b,...] -> parse_desc of every registered layer type (model_cnn.py:122-145)"""
        p_start = layer_desc.find("[")
        p_end = layer_desc.find("]")
        layer_params = {"classNum": self.class_num, "activation": activation, "borderMode": border_mode, "wb": wb}
        if p_start > 0 and p_end > p_start:
            layer_type = layer_desc[:p_start]
            for i, p in enumerate(layer_desc[(p_start + 1):p_end].split(",")):
                layer_params[i] = common.convert_num(p)
        else:
            layer_type = layer_desc
        t_index = layer_type.find(".")
        if t_index > 0:
            layer_tags = layer_type[(t_index + 1):]
            layer_type = layer_type[:t_index]
        else:
            layer_tags = ""
        for layer in layer_types:
            if layer.parse_desc(layers, layer_type, layer_tags, layer_params):
                return
        raise Exception("Invalid layer - type: ", layer_type, "tags:", layer_tags, "params:", layer_params)

    def build(self, model_desc, data_shape, activation="relu", border_mode="valid", weight_init="he-forward"):
        if isinstance(model_desc, str):
            model_desc = model_desc.split()
        if isinstance(weight_init, str):
            weight_init = [weight_init]
        self.model_desc = " ".join(model_desc)
        self.data_shape = tuple(data_shape)
        self.layers = [self._initial_layer()]
        for i, layer_desc in enumerate(model_desc):
            wb = weight_init[min(len(weight_init) - 1, i)]
            self.build_layer(layer_desc, self.layers, activation, border_mode, wb)
        self._ready = False

    def export_json(self):
        self._sync_params_to_host()
        json_layers = [self.layers[index].export_json() for index in range(1, len(self.layers))]
        from time import gmtime, strftime
        json_obj = {"classifierType": "CNN", "classLabels": self.class_labels, "classNum": self.class_num,
                    "dataShape": self.data_shape, "date": strftime("%Y-%m-%d %H:%M:%S", gmtime()),
                    "user": getpass.getuser()}
        json_obj.update({"version": 3, "layers": json_layers})
        return json_obj

    def import_json(self, json_obj, layer_range=None):
        self.func = {}
        if json_obj.get("version", 0) == 0:
            raise Exception("Old format model file detected, no compatibility!")
        self.class_labels = json_obj["classLabels"]
        if "imageSize" in json_obj and "imageMode" in json_obj:
            width, height = json_obj["imageSize"][0], json_obj["imageSize"][1]
            self.data_shape = ({"RGB": 3, "L": 1}[json_obj.get("imageMode", "RGB")], width, height)
        elif "dataShape" in json_obj:
            self.data_shape = tuple(json_obj["dataShape"])
        else:
            assert False, "Bad mdl file, Cannot determine input data shape!"
        assert json_obj.get("imageBorder", 0) == 0
        self.class_num = json_obj.get("classNum", len(self.class_labels) if self.class_labels else 0)
        layers = layer_mod.import_json(json_obj["layers"], None, self.get_input_shape(), layer_range)
        layers[0].is_model_input = True
        # the first real layer was built before the flag existed: refresh it
        for l in layers[1:2]:
            if hasattr(l, "is_first"):
                l.is_first = True
        self.layers = layers
        self._ready = False

    def convert_bn_relu(self):
        """merge batchnorm + relu pairs into batchnorm-relu layers, at the top level and inside 'original' ResNet
        blocks - what `model-modify --convert-bn-relu` does to build the DeNet stacks (reference model/modify.py:70-110)"""
        js = self.export_json()

        def as_bnrelu(j):
            j = dict(j)
            j["type"] = "batchnorm-relu"
            return j

        def is_bn_relu_pair(a, b):
            return a["type"] == "batchnorm" and b is not None and b["type"] == "activation" and \
                b.get("activation") == "relu"

        src, out, i = js["layers"], [], 0
        while i < len(src):
            cur = src[i]
            nxt = src[i + 1] if i + 1 < len(src) else None
            if is_bn_relu_pair(cur, nxt):
                out.append(as_bnrelu(cur))
                i += 2
                continue
            if cur["type"] == "resnet" and "bnrelu" not in cur["version"] and "pre-activation" not in cur["version"]:
                cur = dict(cur)
                sub = [l for l in cur["layers"] if l["type"] != "identity"]
                sub[2] = as_bnrelu(sub[2])
                del sub[3]
                if cur["bottleneck"] > 0:
                    sub[4] = as_bnrelu(sub[4])
                    del sub[5]
                cur["layers"] = sub
                cur["version"] = cur["version"] + ",bnrelu"
            out.append(cur)
            i += 1
        js["layers"] = out
        batch_size = self.batch_size
        self.import_json(js)
        self.batch_size = batch_size
        return self

    def _sync_params_to_host(self):
        if self.device is not None:
            torch.cuda.synchronize()

    # ---------------------------------------------------------------------------------------------- device setup
    def to_device(self, device=None, precision=None):
        """move parameters to the GPU, allocate one flat fp32 gradient buffer and the solver tables"""
        if not torch.cuda.is_available():
            raise lib.DenetError("denet_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        lib.load()
        if precision is not None:
            layer_mod.set_precision(precision)
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        layer_mod.set_device(self.device)
        for l in _walk(self.layers):
            l.to(self.device)
        # trainable parameters in the reference's order: per layer weights() then biases() (model_cnn.py:308-316)
        self.train_params = []
        owners = []
        for index, l in enumerate(self.layers):
            if index in self.skip_layer_updates:
                continue
            mine = [(p, True) for p in l.weights()] + [(p, False) for p in l.biases()]
            self.train_params += mine
            owners += [index] * len(mine)
        total = sum(p.numel() for p, _ in self.train_params)
        # every tensor starts on a 16-byte boundary inside the flat buffer
        offsets, off = [], 0
        self.layer_grad_ranges = {}     # top-level layer index -> [start, end) of its gradients in flat_grad
        for (p, _), owner in zip(self.train_params, owners):
            offsets.append(off)
            start = self.layer_grad_ranges.get(owner, (off, off))[0]
            off += (p.numel() + 3) // 4 * 4
            self.layer_grad_ranges[owner] = (start, off)
        self.flat_grad = torch.zeros((max(off, 4),), dtype=torch.float32, device=self.device)
        self.grad_offsets = offsets
        for (p, _), o in zip(self.train_params, offsets):
            p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)
        self.num_trainable = total
        # per-step batch-norm statistics buffer (conv epilogue -> bn finalize), zeroed once per step
        bns = link_fusions(self.layers)
        from ..layer.batch_norm import BatchNormLayer
        all_bns = [l for l in _walk(self.layers) if isinstance(l, BatchNormLayer) and l.enabled]
        csum = sum(2 * b.input_shape[1] for b in bns) + sum(2 * b.input_shape[1] for b in all_bns)
        self.bn_stat_buffer = torch.zeros((max(csum, 2),), dtype=torch.float32, device=self.device)
        o = 0
        for b in bns:
            c = b.input_shape[1]
            b._fused = (self.bn_stat_buffer[o:o + c], self.bn_stat_buffer[o + c:o + 2 * c])
            o += 2 * c
        for b in all_bns:      # backward statistics accumulated by the dgrad epilogue that produces the layer's dy
            c = b.input_shape[1]
            b._bwd_sums = (self.bn_stat_buffer[o:o + c], self.bn_stat_buffer[o + c:o + 2 * c])
            o += 2 * c
        layer_mod.bump_param_version()
        self._ready = True
        return self

    def enable_data_parallel(self, bucket_bytes=None, average_bn_stats=True, group=None):
        """gradient all-reduce across the ranks of torch.distributed (NCCL on GPUs).  bucket_bytes > 0: buckets are
        reduced on a side stream as the backward pass retires their layers; 0: ONE all-reduce of the whole flat
        gradient after the backward pass (131 MB over NVLink/NVSwitch is ~0.35 ms on 8 GPUs - less than what the
        overlapped variant loses when the collective's CTAs share the SMs with the persistent conv kernels).
        Default from DENET_DDP_BUCKET_MB, 0 = one all-reduce: measured on 8 B200 (profiles/r2_scaling.md) 22130 vs
        21932 images/s device-resident and 22016 vs 18785 end to end against 32 MB overlapped buckets."""
        from ..multi import GradientAllReduce
        if bucket_bytes is None:
            bucket_bytes = int(float(os.environ.get("DENET_DDP_BUCKET_MB", "0")) * (1 << 20))
        if bucket_bytes <= 0:
            bucket_bytes = 1 << 62
        if not self._ready:
            self.to_device()
        ranges = [(index, r[0], r[1]) for index, r in sorted(self.layer_grad_ranges.items())]
        extra = []
        if average_bn_stats:
            for l in _walk(self.layers):
                extra += [t for t, _ in l.updates()] if not len(l.layers) else []
            if extra and all(t.is_cuda and t.dtype == torch.float32 for t in extra):
                # move the running statistics of all batch-norm layers into ONE flat buffer (the parameters become
                # views of it): their per-step average is then a single in-place all-reduce
                flat = torch.empty((sum(t.numel() for t in extra),), dtype=torch.float32, device=extra[0].device)
                o = 0
                for t in extra:
                    n = t.numel()
                    flat[o:o + n].copy_(t.detach().reshape(-1))
                    t.data = flat[o:o + n].view(t.shape)
                    o += n
                self._bn_running_flat = flat
                extra = [flat]
        self.ddp = GradientAllReduce(self.flat_grad, ranges, bucket_bytes, extra, group)
        # every replica must start from ONE model (the reference's workers all load the shared model,
        # multi/worker.py:85-122): fresh models draw their weights from the unseeded global numpy.random, so without
        # this the ranks would apply an averaged gradient to different weights and never converge
        self.sync_state_from_rank0(group)
        return self.ddp

    def state_tensors(self):
        """every tensor that defines the training state of this replica: parameters + batch-norm running statistics
        (layer.params()), and the solver momenta once build_train_func() has created them"""
        out = []
        for l in _walk(self.layers):
            if not len(l.layers):
                out += list(l.params())
        seen, uniq = set(), []
        for t in out + list(getattr(self, "momenta", None) or []) + list(getattr(self, "momenta2", None) or []):
            if id(t) not in seen:
                seen.add(id(t))
                uniq.append(t)
        return uniq

    def sync_state_from_rank0(self, group=None):
        """broadcast parameters, running statistics and momenta from rank 0 (no-op without torch.distributed)"""
        from ..multi import ddp as ddp_mod
        n = ddp_mod.broadcast_state(self.state_tensors(), src=0, group=group)
        if n:
            layer_mod.bump_param_version()
        return n

    def _build_solver_tables(self):
        chunk = lib.load().denet_solver_chunk()
        entry_bytes = lib.load().denet_solver_entry_bytes()
        assert entry_bytes == 48, entry_bytes
        n = len(self.train_params)
        self.momenta = [torch.zeros_like(p) for p, _ in self.train_params]
        self.momenta2 = [torch.zeros_like(p) for p, _ in self.train_params] if self.solver_mode == "adam" else None
        table = numpy.zeros((n, 6), dtype=numpy.int64)   # p, g, m, v, n, (is_weight | pad<<32)
        block_tensor, block_offset = [], []
        for i, (p, is_weight) in enumerate(self.train_params):
            table[i, 0] = p.data_ptr()
            table[i, 1] = p.grad.data_ptr()
            table[i, 2] = self.momenta[i].data_ptr()
            table[i, 3] = self.momenta2[i].data_ptr() if self.momenta2 is not None else 0
            table[i, 4] = p.numel()
            table[i, 5] = 1 if is_weight else 0
            for o in range(0, p.numel(), chunk):
                block_tensor.append(i)
                block_offset.append(o)
        self._solver_entries = torch.from_numpy(table).to(self.device)
        self._solver_block_tensor = torch.tensor(block_tensor, dtype=torch.int32, device=self.device)
        self._solver_block_offset = torch.tensor(block_offset, dtype=torch.int64, device=self.device)
        self._solver_nblocks = len(block_tensor)
        if self.ddp is not None:
            self.sync_state_from_rank0(self.ddp.group)

    def build_train_func(self, solver_mode="sgd", cost_factors=[], use_acc_mode=False, skip_build=False):
        """collect the cost layers and prepare the solver (model_cnn.py:205-405)"""
        if solver_mode not in SOLVER_CODES:
            solver_mode = "sgd"   # the reference falls through to sgd for unknown names (:301-305)
        self.solver_mode = solver_mode
        self.cost_layers = [l for l in self.layers if l.has_cost]
        self.cost_layer_names = [l.type_name for l in self.cost_layers]
        self.cost_factors = [1.0] * len(self.cost_layers) if len(cost_factors) == 0 else [float(c) for c in
                                                                                            cost_factors]
        assert len(self.cost_factors) == len(self.cost_layers), \
            "Different number of cost factors (%i) and cost layers (%i)" % (len(self.cost_factors),
                                                                            len(self.cost_layers))
        for l, f in zip(self.cost_layers, self.cost_factors):
            l.grad_factor = f
        self.use_split_mode = False
        self.use_acc_mode = use_acc_mode
        self._cost_factor_t = None
        if skip_build:
            return
        if not self._ready:
            self.to_device()
        self._build_solver_tables()
        self.func["train_step"] = self._train_step_device

    # ---------------------------------------------------------------------------------------------- execution
    def upload(self, data_x):
        """host NCHW fp32 batch (numpy or pinned tensor) -> NHWC device activation"""
        if isinstance(data_x, numpy.ndarray):
            data_x = numpy.ascontiguousarray(data_x, dtype=numpy.float32)
        t = layer_mod.h2d(data_x, self.device, slot="model/image" + self._slot_ns if self._static_inputs else None).contiguous()
        first = self.layers[1] if len(self.layers) > 1 else None
        geom = getattr(first, "rowfold", None)
        if geom is not None:
            # image stem: keep the batch zero-padded in NHWC-Cp, the layout the row-folded conv's TMA windows read
            n, c, h, w = t.shape
            split = layer_mod.get_precision() == "fp32"
            img = self._image
            if img is None or (img.n, img.c, img.h, img.w, (img.lo is not None)) != (n, c, h, w, split):
                img = self._image = ops.PaddedImage(n, c, h, w, geom[0], first.pad, geom[1], geom[2], split, t.device)
            return img.fill(t)
        return ops.nchw_to_nhwc(t, layer_mod.act_dtype())

    def upload_metas(self, data_m):
        """ground-truth boxes / classes of the batch as fixed-shape device arrays for the device-side target builders
        (a few KB); None when there are no metas, no DSS head, or an image has more boxes than the kernels handle"""
        if not data_m or not layer_mod._state["device_targets"] or \
                not any(l.type_name == "denet-corner" for l in self.layers):
            return None
        G = ops.MAX_GT
        b = len(data_m)
        if max(len(m.get("bbox", [])) for m in data_m) > G:
            return None
        box = numpy.zeros((b, G, 4), dtype=numpy.float64)
        cls = numpy.zeros((b, G), dtype=numpy.int32)
        cnt = numpy.zeros((b,), dtype=numpy.int32)
        for i, m in enumerate(data_m):
            n = len(m["bbox"])
            cnt[i] = n
            if n:
                box[i, :n] = numpy.asarray(m["bbox"], dtype=numpy.float64)
                cls[i, :n] = numpy.asarray(m["class"], dtype=numpy.int32)
        return (layer_mod.h2d(box, self.device, slot="model/gt_bbox" + self._slot_ns), layer_mod.h2d(cls, self.device, slot="model/gt_class" + self._slot_ns),
                layer_mod.h2d(cnt, self.device, slot="model/gt_count" + self._slot_ns))

    def prepare_operands(self):
        """bf16 GEMM operands of every conv layer from the fp32 master weights in ONE kernel launch (after each
        solver step); layers not covered (im2col variant) refresh their own operands lazily"""
        if self._prep_version == layer_mod.param_version() or self.device is None:
            return
        precision = layer_mod.get_precision()
        if self._prep_table is None or self._prep_precision != precision:
            convs = [l for l in _walk(self.layers) if l.type_name == "conv"]
            records = []
            self._prep_layers = []
            for l in convs:
                ent = l.prep_entries()
                if ent is not None:
                    records += ent
                    self._prep_layers.append(l)
            ebytes = lib.load().denet_weight_prep_entry_bytes()
            chunk = lib.load().denet_weight_prep_chunk()
            assert ebytes == 56, ebytes
            table = numpy.zeros((max(len(records), 1), 7), dtype=numpy.int64)
            ints = table.view(numpy.int32).reshape(len(table), 14)
            block_entry, block_offset = [], []
            for i, (w, op, mode, cp) in enumerate(records):
                cout, cin, R, S = w.shape
                table[i, 0] = w.data_ptr()
                table[i, 1] = op.hi.data_ptr()
                table[i, 2] = op.lo.data_ptr() if op.lo is not None else 0
                items = op.hi.numel() // op.hi.shape[1]      # (operand row, K column) pairs: operands are [rows][taps][K]
                table[i, 3] = items
                ints[i, 8:14] = [cout, cin, R, S, mode, cp]
                for o in range(0, items, chunk):
                    block_entry.append(i)
                    block_offset.append(o)
            self._prep_table = (torch.from_numpy(table).to(self.device),
                                torch.tensor(block_entry, dtype=torch.int32, device=self.device),
                                torch.tensor(block_offset, dtype=torch.int64, device=self.device), len(block_entry))
            self._prep_precision = precision
        entries, be, bo, nblocks = self._prep_table
        lib.call("denet_conv_weight_prep_multi", entries.data_ptr(), be.data_ptr(), bo.data_ptr(), nblocks,
                 ops._stream())
        for l in self._prep_layers:
            l.mark_operands_current()
        self._prep_version = layer_mod.param_version()

    def forward(self, data_x, data_m=None, train=False):
        """one pass over the layer list; in train mode every layer's get_target runs right before its forward so
        that the sparse layer can sample from the corner maps of this very pass"""
        layer_mod.set_train(train)
        layer_mod.set_ground_truth(self.upload_metas(data_m) if train else None)
        self.prepare_operands()
        x = self.upload(data_x)
        self.layers[0].output = x
        return self.forward_layers(x, 1, len(self.layers), data_x, data_m, train)

    def forward_layers(self, x, start, end, data_x=None, data_m=None, train=False, with_targets=True):
        for l in self.layers[start:end]:
            if train and with_targets:
                target = l.get_target(self, data_x, data_m)
                if target is not None:
                    l.set_target(*target)
            x = l.forward(x)
        return x

    def backward(self):
        """reverse pass over the layer list.  The split-K reductions of the filter gradients are deferred and run as
        ONE multi-tensor launch: at the end of the pass, or - with data parallelism - whenever a top-level layer
        retires, right before its gradients may enter an all-reduce bucket."""
        dy = None
        hook = self.ddp.layer_done if self.ddp is not None else None
        pending = [] if self.defer_wgrad_reduce else None
        layer_mod.set_wgrad_pending(pending)
        side = keep = None
        if self.overlap_wgrad and pending is not None and not lib.timing_active():
            if self._wgrad_stream is None:
                self._wgrad_stream = torch.cuda.Stream()
            side, keep = self._wgrad_stream, []
            layer_mod.set_wgrad_side((side, keep))

        def join_side():
            # the reductions (and anything that reads the gradients) wait for the filter gradients of the side stream
            if side is not None:
                done = torch.cuda.Event()
                done.record(side)
                torch.cuda.current_stream().wait_event(done)
        try:
            for index in range(len(self.layers) - 1, 0, -1):
                layer = self.layers[index]
                bn_next = _bn_consumer(self.layers[index - 1]) if self.fuse_bn_backward else None
                if bn_next is not None and getattr(layer, "accepts_bn_next", False):
                    dy = layer.backward(dy, bn_next=bn_next)
                else:
                    dy = layer.backward(dy)
                if hook is not None:
                    if pending and self.ddp.will_launch(index):
                        join_side()
                        ops.wgrad_reduce_pending(pending)
                    hook(index)
            join_side()
            if pending:
                ops.wgrad_reduce_pending(pending)
        finally:
            layer_mod.set_wgrad_pending(None)
            layer_mod.set_wgrad_side(None)
        return dy

    def solver_step(self, learning_rate, momentum, decay, iteration, grad_scale=1.0, hp_dev=None):
        """one multi-tensor update launch; with hp_dev the per-step scalars are read from device memory (CUDA graphs)"""
        if hp_dev is not None:
            lib.call("denet_solver_update_dev", self._solver_entries.data_ptr(), self._solver_block_tensor.data_ptr(),
                     self._solver_block_offset.data_ptr(), self._solver_nblocks, SOLVER_CODES[self.solver_mode],
                     hp_dev.data_ptr(), int(self.bias_decay), ops._stream())
        else:
            mom = list(momentum) + [0.0, 0.0]
            lib.call("denet_solver_update", self._solver_entries.data_ptr(), self._solver_block_tensor.data_ptr(),
                     self._solver_block_offset.data_ptr(), self._solver_nblocks, SOLVER_CODES[self.solver_mode],
                     float(learning_rate), float(mom[0]), float(mom[1]), float(decay), int(iteration),
                     int(self.bias_decay), float(grad_scale), ops._stream())
        layer_mod.bump_param_version()

    def _pack_costs(self):
        """[total, cost_0, ...] device tensor in ONE launch of our own (denet_pack_costs); every cost layer keeps its
        terms in a persistent device tensor `cost_value` whose address is tabled once"""
        if self._cost_factor_t is None:
            vals = [l.cost_value for l in self.cost_layers]
            assert all(v is not None and v.is_cuda and v.dtype == torch.float32 for v in vals)
            dev = vals[0].device
            self._cost_table = (torch.tensor([v.data_ptr() for v in vals], dtype=torch.int64, device=dev),
                                torch.tensor([v.numel() for v in vals], dtype=torch.int32, device=dev), vals)
            self._cost_factor_t = torch.tensor(self.cost_factors, dtype=torch.float32, device=dev)
            self._cost_out = torch.zeros((1 + len(vals),), dtype=torch.float32, device=dev)
        ptrs, lens, vals = self._cost_table
        lib.call("denet_pack_costs", ptrs.data_ptr(), lens.data_ptr(), self._cost_factor_t.data_ptr(), len(vals),
                 self._cost_out.data_ptr(), ops._stream())
        return self._cost_out

    def _train_step_device(self, data_x, data_m, epoch, it, learning_rate, momentum, decay):
        """forward + backward + update on the device; returns the device tensor [total, cost_0, cost_1, ...]"""
        layer_mod.set_epoch(epoch)
        layer_mod.set_iteration(it)
        if tuple(data_x.shape) != self.get_input_shape():
            # layer shapes (and captured graphs) are static; the reference's Dataset.export pads the last batch to a
            # full one (dataset/__init__.py:349-366), so a short batch is a caller error
            raise ValueError("train_step: batch shape %s does not match the model input %s" %
                             (tuple(data_x.shape), self.get_input_shape()))
        if self.gradient_clip > 0:
            raise NotImplementedError("gradient_clip > 0 (reference model_cnn.py grad_clip) is not implemented on the "
                                      "B200 path; refusing to train without it")
        self._host_costs = None        # set again by a graphed step that fetched its costs early (see train_step)
        ops.pin_stream(True)
        try:
            with torch.no_grad():
                if self._use_graphs:
                    return self._train_step_graphed(data_x, data_m, it, learning_rate, momentum, decay)
                return self._train_step_eager(data_x, data_m, it, learning_rate, momentum, decay)
        finally:
            ops.pin_stream(False)
            layer_mod.set_train(False)

    def _train_step_eager(self, data_x, data_m, it, learning_rate, momentum, decay):
        self.bn_stat_buffer.zero_()
        self.forward(data_x, data_m, train=True)
        if self.ddp is not None:
            self.ddp.begin_step()
        self.backward()
        grad_scale = 1.0
        if self.ddp is not None:
            grad_scale = self.ddp.finish_step()
        self.solver_step(learning_rate, momentum, decay, it, grad_scale)
        self.last_costs_device = self._pack_costs()
        return self.last_costs_device

    # ---------------------------------------------------------------------------------------------- CUDA graphs
    def enable_cuda_graphs(self, on=True):
        """Replay the training step from captured CUDA graphs instead of ~520 eager launches.  The step is cut where
        the HOST has to look at device results: after the corner layer + device sampler (the reference's
        python-`random` post-processing of the RoIs runs on the host between the two graphs).  Models without a
        sparse layer are one graph.  Inputs (image batch, ground truth, RoI boxes, solver scalars) live in persistent
        device buffers refreshed by small copies before each replay.  The first graphed step runs eagerly (creates
        every persistent buffer), the second captures."""
        self._use_graphs = bool(on)
        self._static_inputs = bool(on)
        self._graphs = None
        self._graph_warm = False

    def _sparse_index(self):
        for i, l in enumerate(self.layers):
            if l.type_name == "denet-sparse":
                return i
        return None

    def _write_hp(self, it, learning_rate, momentum, decay):
        mom = list(momentum) + [0.0, 0.0]
        world = self.ddp.world if self.ddp is not None else 1
        hp = numpy.array([learning_rate, mom[0], mom[1], decay, float(it), 1.0 / world], dtype=numpy.float32)
        return layer_mod.h2d(hp, self.device, slot="model/hp" + self._slot_ns)

    def _segment_a(self, si):
        """image -> ... -> corner layer (+ device targets, corner cost) -> device sampler"""
        self.bn_stat_buffer.zero_()
        self._prep_version = -1                  # the operand preparation is part of the captured graph
        self.prepare_operands()
        if self._image_outside_graph():
            x = self._image              # filled by the eager conversion kernel that precedes this graph's replay
        else:
            x = self.upload(layer_mod.slot_tensor("model/image" + self._slot_ns))
        self.layers[0].output = x
        end = si if si is not None else len(self.layers)
        x = self.forward_layers(x, 1, end, train=True, with_targets=False)
        if si is not None:
            self.layers[si].enqueue_samples()
        return x

    def _image_outside_graph(self):
        """image stems keep the batch in a persistent zero-padded NHWC buffer (ops.PaddedImage): its fill kernel runs
        eagerly right before graph A, so the fp32 staging tensor is free again after ONE kernel instead of after the
        whole forward trunk - the next step's host->device upload gets (almost) the whole step to complete"""
        first = self.layers[1] if len(self.layers) > 1 else None
        return getattr(first, "rowfold", None) is not None and self._image is not None

    def _segment_b1(self, x, si):
        """sparse gather -> head -> costs (the end of the forward pass)"""
        if si is not None:
            x = self.forward_layers(x, si, len(self.layers), train=True, with_targets=False)
        self._g_costs = self._pack_costs()          # persistent output tensor of denet_pack_costs

    def _segment_b2(self, hp_dev):
        """backward (+ gradient all-reduce) -> solver"""
        if self.ddp is not None:
            self.ddp.begin_step()
        self.backward()
        if self.ddp is not None:
            self.ddp.finish_step()
        self.solver_step(None, None, None, None, hp_dev=hp_dev)

    def _train_step_graphed(self, data_x, data_m, it, learning_rate, momentum, decay):
        """Replays the captured step.  Two things are taken off the critical path of a caller that feeds HOST batches
        and reads the costs back (train_step): the image upload runs on a copy stream and only waits for the previous
        step's stem to have consumed the staging tensor, and the costs - final once the forward pass is done - are
        copied to pinned host memory on that stream while the backward pass and the update still run; train_step
        returns as soon as they have landed, so the next call's upload overlaps this step's backward pass."""
        if not self._graph_warm:
            # eager step through the persistent input buffers: allocates them, the workspaces and the cost tensors
            self._graph_warm = True
            self._host_costs = None
            return self._train_step_eager(data_x, data_m, it, learning_rate, momentum, decay)
        si = self._sparse_index()
        if si is not None and data_m and max(len(m.get("bbox", [])) for m in data_m) > ops.MAX_GT:
            # the captured corner / detection target kernels read the persistent ground-truth slots, which hold at
            # most MAX_GT boxes per image: this batch takes the eager step with the host target builders instead
            # (replaying the graphs would silently train on the previous batch's targets)
            self._host_costs = None
            return self._train_step_eager(data_x, data_m, it, learning_rate, momentum, decay)
        layer_mod.set_train(True)
        cur = torch.cuda.current_stream()
        # refresh the static inputs of this step
        img_ready = None
        host_image = not (torch.is_tensor(data_x) and data_x.is_cuda)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
            self._img_consumed = torch.cuda.Event()
            self._cost_ready = torch.cuda.Event()
            self._cost_copied = torch.cuda.Event()
        if host_image:
            if isinstance(data_x, numpy.ndarray):
                data_x = numpy.ascontiguousarray(data_x, dtype=numpy.float32)
            _, img_ready = layer_mod.h2d_on_stream(data_x, self.device, "model/image" + self._slot_ns, self._copy_stream,
                                                   self._img_consumed if self._img_consumed_valid else None)
        else:
            layer_mod.h2d(data_x, self.device, slot="model/image" + self._slot_ns)
        gt = self.upload_metas(data_m)
        layer_mod.set_ground_truth(gt)
        for l in self.layers[1:]:
            if l.type_name == "regression":          # labels are the only host-built target left
                l.set_target(*l.get_target(self, data_x, data_m))
        hp_dev = self._write_hp(it, learning_rate, momentum, decay)
        if self._graphs is None:
            if img_ready is not None:
                img_ready.synchronize()
            self._capture(si, hp_dev)
        ga, gb1, gb2 = self._graphs
        if img_ready is not None:
            cur.wait_event(img_ready)
        if self._image_outside_graph():
            self.upload(layer_mod.slot_tensor("model/image" + self._slot_ns))     # one kernel into the persistent buffer
            self._img_consumed.record(cur)           # the staging tensor of the image may be overwritten from here on
            ga.replay()
        else:
            ga.replay()
            self._img_consumed.record(cur)
        self._img_consumed_valid = True
        if si is not None:
            sp = self.layers[si]
            ahead = sp.random_ahead()                # generator work while the GPU runs graph A
            sp.finish_target(data_m, *sp.collect_samples(), ahead=ahead)
            gb1.replay()
        self._host_costs = None
        if host_image:
            # costs are final here: fetch them on the copy stream while the backward pass runs
            if self._costs_pinned is None or self._costs_pinned.shape != self._g_costs.shape:
                self._costs_pinned = torch.empty(self._g_costs.shape, dtype=torch.float32).pin_memory()
            self._cost_ready.record(cur)
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(self._cost_ready)
                self._costs_pinned.copy_(self._g_costs, non_blocking=True)
                self._cost_copied.record(self._copy_stream)
            self._host_costs = (self._costs_pinned, self._cost_copied)
        gb2.replay()
        layer_mod.bump_param_version()
        self.last_costs_device = self._g_costs
        return self.last_costs_device

    def _capture(self, si, hp_dev):
        assert any(l.type_name == "denet-corner" for l in self.layers) == (layer_mod.get_ground_truth() is not None), \
            "CUDA graphs need the device-side target builders (<= %d boxes per image)" % ops.MAX_GT
        for l in _walk(self.layers):
            if hasattr(l, "_wver"):
                l._wver = -1                         # weight operand preparation must be part of the graph
        self._g_costs = torch.zeros((1 + len(self.cost_layers),), dtype=torch.float32, device=self.device)
        torch.cuda.synchronize()
        pool = torch.cuda.graph_pool_handle()
        ga = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ga, pool=pool):
            ops.pin_stream(True)
            x = self._segment_a(si)
            if si is None:
                self._segment_b1(x, si)
        gb1 = None
        if si is not None:
            # the RoI box buffers are persistent slots created by the eager warm-up step: record against them
            gb1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gb1, pool=pool):
                ops.pin_stream(True)
                self._segment_b1(x, si)
        gb2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gb2, pool=pool):
            ops.pin_stream(True)
            self._segment_b2(hp_dev)
        ops.pin_stream(True)
        self._graphs = (ga, gb1, gb2)
        # the staging tensors are baked into the graphs now: a later shape change must fail loudly, not re-allocate
        layer_mod.freeze_slots(self._slot_ns)
        for l in _walk(self.layers):
            layer_mod.freeze_slots(l._slot_ns)

    def train_step(self, data_x, data_m, epoch, it, learning_rate, momentum, decay):
        """reference contract (model_cnn.py:407-445): returns (cost, [layer costs]) as python floats"""
        assert "train_step" in self.func, "Call build_train_func() before calling train_step()"
        dev = self._train_step_device(data_x, data_m, epoch, it, learning_rate, momentum, decay)
        if self._host_costs is not None:
            # graphed step: the costs were copied to pinned memory right after the forward pass; the backward pass and
            # the update may still be running (every later call is ordered behind them on the compute stream)
            pinned, landed = self._host_costs
            self._host_costs = None
            landed.synchronize()
            layer_mod.transfer_bytes["d2h"] += pinned.numel() * 4
            costs = pinned.numpy().copy()
        else:
            costs = layer_mod.d2h(dev)
        return float(costs[0]), [float(c) for c in costs[1:]]

    def train_epoch(self, dataset, epoch, learning_rate, momentum=[0, 1, 0], decay=0.0, solver_mode="sgd"):
        dataset_x, dataset_m, dataset_size = dataset.export(self.batch_size)
        index_num = math.ceil(dataset_size / self.batch_size)
        total_cost = 0
        for index in range(index_num):
            data_x = dataset_x[index * self.batch_size:(index + 1) * self.batch_size]
            data_m = dataset_m[index * self.batch_size:(index + 1) * self.batch_size]
            cost, _ = self.train_step(data_x, data_m, epoch, self.iteration, learning_rate, momentum, decay)
            if math.isnan(cost):   # watch out for GPUs randomly producing NaN (model_cnn.py:463)
                raise Exception("ERROR: Cost is NaN")
            total_cost += cost
            self.iteration += 1
        return total_cost

    # ---------------------------------------------------------------------------------------------- prediction
    def detect_forward(self, data_x, detect_layer, corner_threshold=None, corner_max=None):
        """test-mode forward of a DSS detector for DeNetDetectLayer.get_detections: backbone -> corner maps -> device
        RoI sampler (no random / ground-truth boxes at test time, denet_sparse.py:117-145 with train=False) -> sparse
        gather -> head.  Returns (det_pr, fitness, bbox, sample counts) as device tensors."""
        if not self._ready:
            self.to_device()
        si = self._sparse_index()
        assert si is not None, "detect_forward: the model has no denet-sparse layer"
        sp = self.layers[si]
        ops.pin_stream(True)
        try:
            with torch.no_grad():
                layer_mod.set_train(False)
                layer_mod.set_ground_truth(None)
                self.prepare_operands()
                x = self.upload(data_x)
                self.layers[0].output = x
                x = self.forward_layers(x, 1, si, train=False)
                saved = (sp.corner_threshold, sp.corner_max)
                try:
                    if corner_threshold is not None:
                        sp.corner_threshold = corner_threshold
                    if corner_max is not None:
                        sp.corner_max = corner_max
                    counts = sp.sample_for_inference()
                finally:
                    sp.corner_threshold, sp.corner_max = saved
                self.forward_layers(x, si, len(self.layers), train=False)
                det_pr, fitness, bbox = detect_layer.detect_outputs()
        finally:
            ops.pin_stream(False)
        return det_pr, fitness, bbox, counts

    def predict_output_step(self, data_x):
        if not self._ready:
            self.to_device()
        with torch.no_grad():
            out = self.forward(data_x, None, train=False)
        if out.dim() == 4:
            out = ops.nhwc_to_nchw(out)
        return out.float().cpu().numpy()

    def predict_output(self, dataset):
        dataset_x, dataset_y, dataset_size = dataset.export(self.batch_size)
        n = math.ceil(dataset_size / self.batch_size)
        pr = [self.predict_output_step(dataset_x[i * self.batch_size:(i + 1) * self.batch_size]) for i in range(n)]
        pr = numpy.concatenate(pr, axis=0)
        return pr[:dataset_size]

    def predict_label(self, dataset):
        pr = self.predict_output(dataset)
        assert pr.ndim == 2
        return [int(numpy.argmax(pr[i])) for i in range(pr.shape[0])]
