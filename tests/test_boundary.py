"""Drop-in boundary at driver level.

* CPU, in the build container only (needs /root/reference): the reference's OWN driver denet/model/train.py (what
  bin/model-train executes) runs UNMODIFIED under denet_b200.compat's `denet` alias - argument parsing, the reference's
  dataset loader and logging, then `model_cnn.initialize(args, ...)` building the model through our parse_desc - up to the
  point where build_train_func needs the GPU (there is no CPU fallback, so it must raise DenetError there).
* GPU: the same driver sequence train.py:111-133 (`initialize` -> `build_train_func` -> `train_epoch`) with an argparse
  namespace produced by a parser that restates bin/model-train's flags (train.py:50-83), on a fake dataset object with
  the reference Dataset interface (export / shuffle / subset_num / load_from_subset).
"""
import argparse
import math
import os
import random
import sys

import numpy
import pytest

REF = "/root/reference"


def model_train_parser():
    """the flags of bin/model-train that reach the model (reference denet/model/train.py:50-83), same names/defaults"""
    p = argparse.ArgumentParser()
    p.add_argument("--model", default=None)
    p.add_argument("--cost-factors", default=[], nargs="+")
    p.add_argument("--thread-num", type=int, default=1)
    p.add_argument("--border-mode", default="valid")
    p.add_argument("--output-prefix", default="./model")
    p.add_argument("--activation", default="relu")
    p.add_argument("--solver", type=str, default="nesterov")
    p.add_argument("--weight-init", nargs="+", default=["he-backward"])
    p.add_argument("--learn-rate", type=float, default=0.1)
    p.add_argument("--learn-momentum", type=float, default=[0.0, 0.0], nargs="+")
    p.add_argument("--learn-anneal", type=float, default=1)
    p.add_argument("--learn-decay", type=float, default=0.0)
    p.add_argument("--epochs", type=int, default=30)
    p.add_argument("--batch-size", type=int, default=32)
    p.add_argument("--seed", type=int, default=23455)
    p.add_argument("--skip-layer-updates", type=int, nargs="+", default=[])
    p.add_argument("--model-desc", default=["C[100,7]", "P[2]", "C[150,4]", "P[2]", "C[250,4]", "P[2]", "C[300,1]", "R"],
                   nargs="+", type=str)
    return p


class FakeDataset:
    """the part of the reference's DatasetAbstract that ModelCNN.train_epoch / predict_output touch
    (dataset/__init__.py:349-366: export pads the last batch to a full one with random samples)"""

    def __init__(self, n, shape, classes, seed):
        rs = numpy.random.RandomState(seed)
        self.labels = rs.randint(0, classes, size=n)
        # class-dependent mean so that a few SGD steps can actually reduce the cost
        self.x = (rs.uniform(0, 1, (n,) + tuple(shape)) * 0.5 + self.labels[:, None, None, None] / (2.0 * classes)) \
            .astype(numpy.float32)
        self.class_labels = {str(i): i for i in range(classes)}
        self.subset_num = 1

    def __len__(self):
        return len(self.x)

    def get_data_shape(self):
        return self.x.shape[1:]

    def get_class_num(self):
        return len(self.class_labels)

    def shuffle(self):
        pass

    def load_from_subset(self, subset):
        pass

    def export(self, batch_size=1):
        size = batch_size * math.ceil(len(self) / batch_size)
        idx = [i if i < len(self) else random.randint(0, len(self) - 1) for i in range(size)]
        metas = [{"image_class": int(self.labels[i]), "partial": False} for i in idx]
        return self.x[idx], metas, len(self)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree exists in the build container only")
def test_reference_train_driver_runs_unmodified_up_to_the_device(tmp_path, capsys):
    from PIL import Image
    from denet_b200 import compat, lib
    rs = numpy.random.RandomState(0)
    for cls in ("cat", "dog"):
        os.makedirs(tmp_path / "train" / cls)
        for i in range(4):
            Image.fromarray(rs.randint(0, 255, (16, 16, 3), dtype=numpy.uint8)).save(tmp_path / "train" / cls / ("%d.ppm" % i))
    argv = ["--train", str(tmp_path / "train"), "--extension", "ppm", "--model-desc", "C[8,3]", "BN", "A", "P[2]",
            "C.B[16,3]", "BNA", "P.A", "R", "--border-mode", "half", "--batch-size", "4", "--epochs", "1",
            "--solver", "nesterov", "--output-prefix", str(tmp_path / "model")]
    import torch
    try:
        if torch.cuda.is_available():
            compat.run_reference_script(REF, "denet/model/train.py", argv)       # trains for real on a GPU box
            assert os.path.exists(str(tmp_path / "model") + "_epoch000_final.mdl.gz")
        else:
            with pytest.raises(lib.DenetError, match="CUDA device"):
                compat.run_reference_script(REF, "denet/model/train.py", argv)
        # the reference driver got as far as our ModelCNN: its own import line resolved to denet_b200
        import denet.model.model_cnn as aliased
        from denet_b200.model import model_cnn
        assert aliased is model_cnn
        import denet.layer.layer_types as lt
        assert [t.type_name for t in lt.layer_types].count("denet-sparse") == 1
        import denet.dataset as ref_dataset                                       # the reference's, unmodified
        assert ref_dataset.__file__.startswith(REF)
    finally:
        compat.uninstall_alias()


def test_alias_without_a_reference_tree():
    """on a box without the reference sources the alias still serves the hot-path modules"""
    from denet_b200 import compat
    compat.install_alias(None)
    try:
        import denet.layer.denet_sparse as a
        import denet.model.model_cnn as b
        import denet.common as c
        from denet_b200.layer import denet_sparse
        from denet_b200.model import model_cnn
        assert a is denet_sparse and b is model_cnn
        assert c.find_layers is not None and c.import_c("x/denet_sparse.cc").build_samples is not None
    finally:
        compat.uninstall_alias()


@pytest.mark.gpu
def test_model_train_driver_sequence(cuda, tmp_path):
    """train.py:84-85,111-160 with our model_cnn: seeds, initialize, build_train_func, epochs of train_epoch, save"""
    from denet_b200.model import model_cnn
    args = model_train_parser().parse_args(
        ["--model-desc", "C[16,3]", "BN", "A", "P[2]", "C.B[32,3]", "BNA", "P.A", "R", "--border-mode", "half",
         "--batch-size", "8", "--epochs", "3", "--learn-rate", "0.05", "--learn-momentum", "0.9", "0.9",
         "--learn-decay", "1e-4", "--solver", "nesterov", "--output-prefix", str(tmp_path / "m")])
    random.seed(args.seed)
    numpy.random.seed(args.seed)
    data = FakeDataset(20, (3, 16, 16), 4, seed=3)          # 20 samples, batch 8: the last batch is padded by export()
    model = model_cnn.initialize(args, data.get_data_shape(), data.class_labels, data.get_class_num())
    model.build_train_func(args.solver, args.cost_factors)
    costs, lr = [], args.learn_rate
    for epoch in range(args.epochs):
        data.shuffle()
        for subset in range(data.subset_num):
            data.load_from_subset(subset)
            costs.append(model.train_epoch(data, epoch, lr, args.learn_momentum, args.learn_decay))
        lr *= args.learn_anneal
        model_cnn.save_to_file(model, args.output_prefix + "_epoch%03i.mdl.gz" % epoch)
    assert model.iteration == 3 * 3 and all(math.isfinite(c) for c in costs)
    assert costs[-1] < costs[0], costs
    labels = model.predict_label(data)
    assert len(labels) == len(data)
    # the checkpoint re-loads through the reference's entry point and reproduces the predictions
    again = model_cnn.load_from_file(args.output_prefix + "_epoch002.mdl.gz", args.batch_size)
    assert again.predict_label(data) == labels
    # a short batch is a caller error (Dataset.export pads), reported instead of silently mis-shaped
    with pytest.raises(ValueError):
        model.train_step(data.x[:3], [{"image_class": 0}] * 3, 0, 0, 0.1, [0.9, 0.9], 0.0)


@pytest.mark.gpu
def test_denet_sparse_extension_signature(cuda):
    """common.import_c('denet_sparse.cc').build_samples(...) with the reference's exact signature and list return
    (denet_sparse.cc:559-571) against the compiled reference, plus build_bbox_array"""
    import oracle
    from util import busy_corner_map
    from denet_b200 import common
    ext = common.import_c("/any/where/denet_sparse.cc")
    cp = busy_corner_map(3, 32, 32, 12, seed=5)
    got = ext.build_samples(3, cp, 0.01, 8, 1024, 0, 1.0)
    cc = oracle.reference_cc()           # the reference's own extension, compiled unmodified (oracle/_ref)
    ref = cc.build_samples(3, cp, 0.01, 8, 1024, 0, 1.0) if cc is not None else None
    assert isinstance(got, list) and len(got) == 3
    for b in range(3):
        assert all(isinstance(s[0], float) and len(s[1]) == 4 for s in got[b])
        prs = [s[0] for s in got[b]]
        assert prs == sorted(prs, reverse=True) and len(got[b]) <= 64
    if ref is not None:
        for b in range(3):
            if len(ref[b]) < 64:          # no cut at the K-th score: the sets must agree exactly
                assert {s[1] for s in got[b]} == {tuple(s[1]) for s in ref[b]}
    bbox = numpy.zeros((3, 8, 8, 4), dtype=numpy.float32)
    ext.build_bbox_array(got, bbox)
    for b in range(3):
        for i, s in enumerate(got[b]):
            assert tuple(bbox[b, i // 8, i % 8]) == tuple(numpy.float32(v) for v in s[1])
