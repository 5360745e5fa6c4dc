"""ctypes binding of the C-ABI library (include/denet_b200.h).

The product path has NO CPU fallback: if libdenet_b200.so is missing or a call fails, this raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdenet_b200.so")

DENET_F32 = 0
DENET_BF16 = 1

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_size_t = ctypes.c_size_t
c_float = ctypes.c_float

# name -> (restype, argtypes); mirrors include/denet_b200.h one to one.
SIGNATURES = {
    "denet_last_error": (ctypes.c_char_p, []),
    "denet_abi_version": (c_int, []),
    "denet_conv_weight_prep": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "denet_split_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_void_p]),
    "denet_conv2d_fprop": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_ll,
                                   c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_void_p, c_int,
                                   c_void_p, c_void_p, c_void_p]),
    "denet_conv2d_wgrad_workspace": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "denet_conv2d_wgrad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_ll,
                                   c_void_p, c_void_p, c_int, c_int, c_int, c_ll, c_int, c_int, c_int, c_int,
                                   c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
}


class DenetError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (once) and attach the prototypes. Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DenetError(
            "denet_b200: %s not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().denet_last_error()
        raise DenetError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def call(name, *args):
    """Call an int-returning entry point and raise DenetError on a non-zero status."""
    fn = getattr(load(), name)
    check(fn(*args), name)
