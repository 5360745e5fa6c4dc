"""ORACLE - TEST INFRASTRUCTURE ONLY.  Build recipe for the checkers.

  1. oracle/libdenet_oracle.so      <- oracle/ref_kernels.c  (our plain-C restatement)                 [always]
  2. oracle/_ref/denet_sparse*.so   <- /root/reference/denet/layer/denet_sparse.cc, compiled UNMODIFIED from where
                                       it lies (CPython extension: build_samples / build_bbox_array)    [if present]
  3. oracle/_ref/libref_cuda_kernels.so <- the reference's inline CUDA kernel text (k_sparse_sample<gs>,
                                       k_sparse_sample_grad<gs>, k_pool_inv*, k_relu) extracted at build time from
                                       the Python string literals in denet_sparse_op.py / pool_inv_op.py /
                                       batch_norm_relu.py, wrapped in extern "C" launchers that use the reference's
                                       own launch geometry, compiled with nvcc for sm_100a.                [if present]

No reference source is copied into the repository: generated files and binaries go to oracle/_ref/ only, which is
git-ignored (but not gpurun-ignored, so the binaries travel to the GPU box).
"""
import os
import re
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("DENET_REFERENCE", "/root/reference")
REF_OUT = os.path.join(HERE, "_ref")

STUB_HEADER = """/* stand-in for Theano's theano_mod_helper.h: the reference .cc includes it but uses none of its macros */
#ifndef THEANO_MOD_HELPER
#define THEANO_MOD_HELPER
#define MOD_PUBLIC __attribute__((visibility("default")))
#define THEANO_EXTERN extern "C"
#define THEANO_RTYPE void
#endif
"""


def run(cmd):
    print("[oracle/build]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def newer(target, *sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_restatement():
    src = os.path.join(HERE, "ref_kernels.c")
    out = os.path.join(HERE, "libdenet_oracle.so")
    if not newer(out, src):
        # -ffp-contract=off: the float sequence is written out explicitly (fmaf where the CUDA kernel has an FMA)
        run(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-ffp-contract=off", "-fno-math-errno", src, "-o", out,
             "-lm"])
    return out


def build_reference_cc(name="denet_sparse"):
    """one of the reference's own CPython extensions (denet_sparse: RoI sampling, denet_detect: NMS), compiled from
    /root/reference unmodified"""
    import numpy
    src = os.path.join(REF_ROOT, "denet", "layer", name + ".cc")
    if not os.path.exists(src):
        return None
    os.makedirs(os.path.join(REF_OUT, "stub"), exist_ok=True)
    stub = os.path.join(REF_OUT, "stub", "theano_mod_helper.h")
    with open(stub, "w") as f:
        f.write(STUB_HEADER)
    ext = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    out = os.path.join(REF_OUT, name + ext)
    if not newer(out, src):
        # flags = the reference's common.import_c (denet/common/__init__.py:182); the two -D map numpy-1 aliases
        # removed in numpy 2 (used at denet_sparse.cc:573,677)
        run(["g++", "-std=c++11", "-fPIC", "-O3", "-fno-math-errno", "-Wno-unused-label", "-Wno-unused-variable",
             "-Wno-write-strings", "-shared", "-pthread",
             "-DNPY_IN_ARRAY=NPY_ARRAY_IN_ARRAY", "-DNPY_INOUT_ARRAY=NPY_ARRAY_INOUT_ARRAY",
             "-I" + os.path.join(REF_OUT, "stub"), "-I" + sysconfig.get_paths()["include"],
             "-I" + numpy.get_include(), src, "-o", out])
    return out


def _support_code(py_path, class_name):
    """return the raw (unformatted) string literal returned by <class_name>.c_support_code in a reference module"""
    text = open(py_path).read()
    cls = text.index("class " + class_name)
    fn = text.index("def c_support_code", cls)
    m = re.compile(r'"""(.*?)"""', re.S).search(text, fn)
    return m.group(1)


LAUNCHERS = r"""
#include <cmath>
// launch geometry = the reference c_code blocks: 1024 threads, ceil(total/1024) blocks, contiguous NCHW strides
static inline dim3 ref_grid(size_t total) { return dim3((unsigned)((total + 1023) / 1024), 1, 1); }
static inline int ref_finish() {
    cudaError_t e = cudaGetLastError();   // launch failures (e.g. too many registers for 1024 threads) must surface
    if (e != cudaSuccess) return (int)e;
    return (int)cudaDeviceSynchronize();
}

#define SPARSE_LAUNCHER(GS)                                                                                       \
extern "C" int refcuda_sparse_sample_fwd_##GS(float* fmap, float* bbox, float* r, size_t bs, size_t fn, size_t h,  \
                                              size_t w, size_t sn) {                                               \
    size_t oc = fn * GS * GS + 2;                                                                                  \
    k_sparse_sample##GS<<<ref_grid(bs * sn * sn), 1024>>>(fmap, fn * h * w, h * w, w, 1, bbox, sn * sn * 4, sn * 4, \
                                                          4, 1, r, oc * sn * sn, sn * sn, sn, 1, fn, h, w, sn, bs); \
    return ref_finish();                                                                           \
}                                                                                                                  \
extern "C" int refcuda_sparse_sample_bwd_##GS(float* dy, float* bbox, float* r, size_t bs, size_t fn, size_t h,    \
                                              size_t w, size_t sn) {                                               \
    size_t oc = fn * GS * GS + 2;                                                                                  \
    cudaMemset(r, 0, 4 * bs * fn * h * w);                                                                         \
    k_sparse_sample_grad##GS<<<ref_grid(bs * sn * sn), 1024>>>(dy, oc * sn * sn, sn * sn, sn, 1, bbox, sn * sn * 4, \
                                                               sn * 4, 4, 1, r, fn * h * w, h * w, w, 1, fn, h, w, \
                                                               sn, bs);                                            \
    return ref_finish();                                                                           \
}

#define POOLINV_LAUNCHER(SW, SH)                                                                                   \
extern "C" int refcuda_pool_inv_fwd_##SW##x##SH(float* x, float* r, size_t bs, size_t fn, size_t h, size_t w) {     \
    k_pool_inv_##SW##x##SH<<<ref_grid(bs * h * w), 1024>>>(x, fn * h * w, h * w, w, 1, r, fn * h * SH * w * SW,     \
                                                           h * SH * w * SW, w * SW, 1, bs, fn, h, w);              \
    return ref_finish();                                                                           \
}                                                                                                                  \
extern "C" int refcuda_pool_inv_bwd_##SW##x##SH(float* dy, float* r, size_t bs, size_t fn, size_t h, size_t w) {    \
    k_pool_inv_grad_##SW##x##SH<<<ref_grid(bs * h * w), 1024>>>(dy, fn * h * SH * w * SW, h * SH * w * SW, w * SW,  \
                                                                1, r, fn * h * w, h * w, w, 1, bs, fn, h, w);      \
    return ref_finish();                                                                           \
}

extern "C" int refcuda_relu(float* x, size_t n) {
    k_relu<<<ref_grid(n), 1024>>>(x, n);
    return ref_finish();
}
"""

GRID_SIZES = (3, 7, 10)
POOL_SIZES = ((2, 2),)


def build_reference_cuda():
    """the reference's inline CUDA kernels, text extracted at build time, compiled for the GPU box"""
    layer = os.path.join(REF_ROOT, "denet", "layer")
    sparse_py = os.path.join(layer, "denet_sparse_op.py")
    pool_py = os.path.join(layer, "pool_inv_op.py")
    bn_py = os.path.join(layer, "batch_norm_relu.py")
    if not all(os.path.exists(p) for p in (sparse_py, pool_py, bn_py)):
        return None
    os.makedirs(REF_OUT, exist_ok=True)
    out = os.path.join(REF_OUT, "libref_cuda_kernels.so")
    gen = os.path.join(REF_OUT, "ref_cuda_kernels.cu")
    if newer(out, sparse_py, pool_py, bn_py, os.path.abspath(__file__)):
        return out
    parts = ["// GENERATED by oracle/build_ref.py from the reference's Python string literals - do not commit\n"
             "#include <cuda_runtime.h>\n#include <cstdio>\n"]
    fwd, bwd = _support_code(sparse_py, "DeNetSparseOp"), _support_code(sparse_py, "DeNetSparseGradOp")
    for gs in GRID_SIZES:
        parts.append(fwd % (gs, gs))   # same substitution as denet_sparse_op.py:85
        parts.append(bwd % (gs, gs))   # denet_sparse_op.py:212
    pfwd, pbwd = _support_code(pool_py, "PoolInvOp"), _support_code(pool_py, "PoolInvGradOp")
    for size_w, size_h in POOL_SIZES:
        parts.append(pfwd % {"size_w": size_w, "size_h": size_h})   # pool_inv_op.py:63
        parts.append(pbwd % {"size_w": size_w, "size_h": size_h})   # pool_inv_op.py:169
    relu = _support_code(bn_py, "BatchNormReluOp")
    parts.append(relu)
    parts.append(LAUNCHERS)
    for gs in GRID_SIZES:
        parts.append("SPARSE_LAUNCHER(%d)\n" % gs)
    for size_w, size_h in POOL_SIZES:
        parts.append("POOLINV_LAUNCHER(%d, %d)\n" % (size_w, size_h))
    with open(gen, "w") as f:
        f.write("\n".join(parts))
    # default nvcc flags (fmad on), like Theano's nvcc invocation (SURVEY.md §2a "Build flags")
    # -maxrregcount 64: the reference launches 1024 threads per block, which needs <= 64 registers per thread
    run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-maxrregcount", "64", "-shared", "-Xcompiler",
         "-fPIC", "-w", gen, "-o", out])
    return out


def build_all(verbose=True):
    built = {"restatement": build_restatement()}
    if os.path.isdir(REF_ROOT):
        built["reference_cc"] = build_reference_cc("denet_sparse")
        built["reference_detect_cc"] = build_reference_cc("denet_detect")
        try:
            built["reference_cuda"] = build_reference_cuda()
        except Exception as e:  # the CUDA harness is a bonus pin; never block the build on it
            print("[oracle/build] reference CUDA kernels not built:", repr(e), file=sys.stderr)
            built["reference_cuda"] = None
    elif verbose:
        print("[oracle/build] %s not present: using prebuilt oracle/_ref if any" % REF_ROOT)
    return built


if __name__ == "__main__":
    print(build_all())
