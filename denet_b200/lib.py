"""ctypes binding of the C-ABI library (include/denet_b200.h).

The prototypes are read from the header itself, so the binding mirrors it one to one.  The product path has NO CPU
fallback: if libdenet_b200.so is missing or a call fails, this raises.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdenet_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "denet_b200.h")

DENET_F32 = 0
DENET_BF16 = 1

_SCALARS = {"int": ctypes.c_int, "double": ctypes.c_double, "long long": ctypes.c_longlong, "float": ctypes.c_float, "size_t": ctypes.c_size_t, "uint32_t": ctypes.c_uint32,
            "cudaStream_t": ctypes.c_void_p}


def _ctype(decl):
    decl = decl.strip()
    if "*" in decl:
        return ctypes.c_char_p if decl.replace(" ", "") == "constchar*" else ctypes.c_void_p
    decl = re.sub(r"\bconst\b", "", decl).strip()
    return _SCALARS[decl]


def parse_header(path=HEADER_PATH):
    """{name: (restype, [argtypes])} for every function declared in the C header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w \*]*?)\b(denet_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        argtypes = []
        if args.strip() not in ("", "void"):
            for a in args.split(","):
                a = a.strip()
                # drop the parameter name (last identifier), keep the type
                t = re.sub(r"\s*\b\w+$", "", a) if not a.endswith("*") else a
                argtypes.append(_ctype(t))
        out[name] = (_ctype(ret), argtypes)
    return out


class DenetError(RuntimeError):
    pass


_lib = None
SIGNATURES = None


def load():
    """Load the shared library (once) and attach the prototypes. Raises if it is not built."""
    global _lib, SIGNATURES
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DenetError(
            "denet_b200: %s not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    SIGNATURES = parse_header()
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.denet_abi_version() != 4:
        raise DenetError("denet_b200: ABI version mismatch (library %d, binding 4) - rebuild" % lib.denet_abi_version())
    _lib = lib
    # profiling / A-B switches (kernel variants only; every variant is a CUDA kernel of this library)
    if os.environ.get("DENET_FPROP_MODE"):
        lib.denet_conv2d_fprop_set_mode(int(os.environ["DENET_FPROP_MODE"]))
    if os.environ.get("DENET_PDL"):
        lib.denet_set_pdl(int(os.environ["DENET_PDL"]))
    if os.environ.get("DENET_BN_MODE"):
        lib.denet_bn_set_mode(int(os.environ["DENET_BN_MODE"]))
    if os.environ.get("DENET_WGRAD_MODE"):
        lib.denet_conv2d_wgrad_set_mode(int(os.environ["DENET_WGRAD_MODE"]))
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().denet_last_error()
        raise DenetError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


_launches = 0


def launch_count():
    """number of C-ABI calls issued so far by this process (bench.py reports the per-step delta)"""
    return _launches


# ---------------------------------------------------------------------------------------------- per-entry timing
# bench.py brackets selected entry points with CUDA events on the stream they are enqueued on (the roofline figure
# needs the device duration of the conv kernels measured live inside the timed region).
_timed = None     # {entry point name: [(start_event, end_event, tag)]} while timing is on
_timed_in_graph = False


def start_timing(names, in_graph=False):
    """in_graph: bracket only the calls made while the stream is being CAPTURED, with external events (event-record
    nodes of the CUDA graph): every replay of that graph re-records them, read_timing() after a replay returns the
    device durations inside the graph - eager launches of kernels shorter than the host's launch cost (~20 us through
    ctypes + tensor-map encoding) measure the host instead."""
    global _timed, _timed_in_graph
    _timed = {n: [] for n in names}
    _timed_in_graph = bool(in_graph)


def read_timing():
    """-> {name: [(milliseconds, tag)]} of the event pairs recorded so far (last replay for in-graph events);
    synchronises the device, keeps timing on"""
    import torch
    torch.cuda.synchronize()
    return {n: [(a.elapsed_time(b), tag) for a, b, tag in evs] for n, evs in (_timed or {}).items()}


def stop_timing():
    """-> {name: [(milliseconds, tag)]}; synchronises the device"""
    global _timed
    out = read_timing()
    _timed = None
    return out


def timing_active():
    return _timed is not None


_tag = None


def set_tag(tag):
    """label attached to the timed calls that follow (bench.py: which layer / pass issued the kernel)"""
    global _tag
    _tag = tag


def call(name, *args, allow=()):
    """Call an int-returning entry point and raise DenetError on a non-zero status (statuses listed in `allow` are
    returned to the caller instead)."""
    global _launches
    _launches += 1
    fn = getattr(load(), name)
    if allow:
        rc = fn(*args)
        if rc != 0 and rc not in allow:
            check(rc, name)
        return rc
    if _timed is not None and name in _timed:
        import torch
        capturing = torch.cuda.is_current_stream_capturing()
        if _timed_in_graph and not capturing:
            check(fn(*args), name)
            return 0
        a = torch.cuda.Event(enable_timing=True, external=capturing)
        b = torch.cuda.Event(enable_timing=True, external=capturing)
        a.record()
        rc = fn(*args)
        b.record()
        _timed[name].append((a, b, _tag))
        check(rc, name)
        return 0
    check(fn(*args), name)
    return 0
