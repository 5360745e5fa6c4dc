"""ORACLE - TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's algorithms for the DeNet training hot path, plus loaders for the real
reference binaries built into oracle/_ref/ (see build_ref.py).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; the product (denet_b200/) never does.
"""
import ctypes
import glob
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_clib = None


class RefSample(ctypes.Structure):
    _fields_ = [("pr", ctypes.c_float), ("x0", ctypes.c_float), ("y0", ctypes.c_float), ("x1", ctypes.c_float),
                ("y1", ctypes.c_float), ("ix0", ctypes.c_int32), ("iy0", ctypes.c_int32), ("ix1", ctypes.c_int32),
                ("iy1", ctypes.c_int32), ("order", ctypes.c_int32)]


SAMPLE_DTYPE = np.dtype([("pr", "f4"), ("x0", "f4"), ("y0", "f4"), ("x1", "f4"), ("y1", "f4"), ("ix0", "i4"),
                         ("iy0", "i4"), ("ix1", "i4"), ("iy1", "i4"), ("order", "i4")])


def clib():
    """the plain-C restatement (oracle/ref_kernels.c); built on demand"""
    global _clib
    if _clib is None:
        path = os.path.join(HERE, "libdenet_oracle.so")
        if not os.path.exists(path):
            from . import build_ref
            build_ref.build_restatement()
        _clib = ctypes.CDLL(path)
    return _clib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def sparse_sample_index(bbox, gs, H, W):
    """integer grid coordinates of k_sparse_sample: returns ys, xs of shape (B, sn, sn, gs) int32"""
    bbox = _f32(bbox)
    B, sn = bbox.shape[0], bbox.shape[1]
    ys = np.empty((B, sn, sn, gs), np.int32)
    xs = np.empty((B, sn, sn, gs), np.int32)
    clib().ref_sparse_sample_index(_p(bbox), B, sn, gs, H, W, _p(ys), _p(xs))
    return ys, xs


def sparse_sample_fwd(fmap, bbox, gs):
    """fmap (B,F,H,W), bbox (B,sn,sn,4) -> (B, gs*gs*F+2, sn, sn); reference denet_sparse_op.py:42-85"""
    fmap, bbox = _f32(fmap), _f32(bbox)
    B, F, H, W = fmap.shape
    sn = bbox.shape[1]
    out = np.empty((B, gs * gs * F + 2, sn, sn), np.float32)
    clib().ref_sparse_sample_fwd(_p(fmap), _p(bbox), _p(out), B, F, H, W, sn, gs)
    return out


def sparse_sample_bwd(dy, bbox, gs, fmap_shape):
    """dy (B, gs*gs*F+2, sn, sn) -> dfmap (B,F,H,W); reference denet_sparse_op.py:171-212"""
    dy, bbox = _f32(dy), _f32(bbox)
    B, F, H, W = fmap_shape
    sn = bbox.shape[1]
    out = np.empty((B, F, H, W), np.float32)
    clib().ref_sparse_sample_bwd(_p(dy), _p(bbox), _p(out), B, F, H, W, sn, gs)
    return out


def pool_inv_fwd(x, sw, sh):
    x = _f32(x)
    B, F, H, W = x.shape
    out = np.empty((B, F, H * sh, W * sw), np.float32)
    clib().ref_pool_inv_fwd(_p(x), _p(out), B, F, H, W, sw, sh)
    return out


def pool_inv_bwd(dy, sw, sh):
    dy = _f32(dy)
    B, F, RH, RW = dy.shape
    H, W = RH // sh, RW // sw
    out = np.empty((B, F, H, W), np.float32)
    clib().ref_pool_inv_bwd(_p(dy), _p(out), B, F, H, W, sw, sh)
    return out


def relu_inplace(x):
    x = _f32(x).copy()
    clib().ref_relu_inplace(_p(x), ctypes.c_size_t(x.size))
    return x


def build_samples(corner_pr, corner_threshold, sample_num, max_corners=1024, local_max=0, cluster_threshold=1.0):
    """C restatement of denet_sparse.cc build_samples (4 corner types, or 5 with the centre map of DNC.C;
    cluster_threshold < 1 runs the reference's apply_cluster on over-full sample sets).

    corner_pr (B,2,4,H,W) fp32 log-probabilities.  Returns a list (per image) of structured arrays with fields
    pr,x0,y0,x1,y1 (normalised floats) and ix0,iy0,ix1,iy1 (integer corner positions), sorted by pr descending,
    and the per-image number of unique candidate boxes before the top-k."""
    cp = _f32(corner_pr)
    B, two, C, H, W = cp.shape
    assert two == 2
    sc = sample_num * sample_num
    out = np.zeros((B, sc), SAMPLE_DTYPE)
    counts = np.zeros((B,), np.int32)
    ncand = np.zeros((B,), np.int32)
    lib = clib()
    lib.ref_build_samples.restype = ctypes.c_int
    rc = lib.ref_build_samples(_p(cp), B, C, H, W, ctypes.c_float(corner_threshold), sample_num, max_corners,
                               local_max, ctypes.c_float(cluster_threshold), _p(out), _p(counts), _p(ncand))
    if rc != 0:
        raise ValueError("oracle.build_samples: unsupported arguments (corner_num must be 4 or 5)")
    return [out[b, :counts[b]].copy() for b in range(B)], ncand


# ------------------------------------------------------------------------------------------- real reference
_ref_cc = None
_ref_cuda = None


def reference_cc():
    """the reference's own denet_sparse C++ extension compiled unmodified (oracle/_ref), or None if not built"""
    global _ref_cc
    if _ref_cc is None:
        hits = sorted(glob.glob(os.path.join(REF_DIR, "denet_sparse*.so")))
        if not hits:
            return None
        spec = importlib.util.spec_from_file_location("denet_sparse", hits[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.init_logging(os.devnull)  # the reference writes a log through a global FILE*; must be opened first
        _ref_cc = mod
    return _ref_cc


_ref_detect_cc = None


def reference_detect_cc():
    """the reference's own denet_detect C++ extension (NMS) compiled unmodified (oracle/_ref), or None if not built"""
    global _ref_detect_cc
    if _ref_detect_cc is None:
        hits = sorted(glob.glob(os.path.join(REF_DIR, "denet_detect*.so")))
        if not hits:
            return None
        spec = importlib.util.spec_from_file_location("denet_detect", hits[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref_detect_cc = mod
    return _ref_detect_cc


def reference_cuda():
    """the reference's inline CUDA kernels compiled for the GPU box (oracle/_ref), or None if not built"""
    global _ref_cuda
    if _ref_cuda is None:
        path = os.path.join(REF_DIR, "libref_cuda_kernels.so")
        if not os.path.exists(path):
            return None
        lib = ctypes.CDLL(path)
        # the launchers take size_t dims: without argtypes ctypes passes 32-bit ints and stack-passed ones are garbage
        vp, sz = ctypes.c_void_p, ctypes.c_size_t
        for gs in (3, 7, 10):
            for d in ("fwd", "bwd"):
                fn = getattr(lib, "refcuda_sparse_sample_%s_%d" % (d, gs))
                fn.argtypes, fn.restype = [vp, vp, vp, sz, sz, sz, sz, sz], ctypes.c_int
        for d in ("fwd", "bwd"):
            fn = getattr(lib, "refcuda_pool_inv_%s_2x2" % d)
            fn.argtypes, fn.restype = [vp, vp, sz, sz, sz, sz], ctypes.c_int
        lib.refcuda_relu.argtypes, lib.refcuda_relu.restype = [vp, sz], ctypes.c_int
        _ref_cuda = lib
    return _ref_cuda
