"""The C-ABI library loads and exports exactly what include/denet_b200.h declares (no compute calls: CPU-only box)."""
import ctypes
import os
import subprocess

from denet_b200 import lib


def test_library_present_and_loads():
    assert os.path.exists(lib.LIB_PATH), "build it first: python -c 'import __graft_entry__ as g; g.build()'"
    l = lib.load()
    import re
    header = open(lib.HEADER_PATH).read()
    assert l.denet_abi_version() == int(re.search(r"#define DENET_ABI_VERSION (\d+)", header).group(1))
    assert l.denet_last_error() is not None


def test_every_declared_symbol_is_exported():
    sigs = lib.parse_header()
    assert len(sigs) >= 40
    l = ctypes.CDLL(lib.LIB_PATH)
    for name in sigs:
        assert hasattr(l, name), "declared in include/denet_b200.h but not exported: " + name


def test_every_exported_symbol_is_declared():
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib.LIB_PATH]).decode()
    exported = {line.split()[-1] for line in out.splitlines() if " T denet_" in line}
    declared = set(lib.parse_header())
    assert exported == declared, (exported - declared, declared - exported)


def test_host_side_queries_need_no_gpu():
    l = lib.load()
    assert l.denet_solver_entry_bytes() == 48
    assert l.denet_solver_chunk() == 4096
    assert l.denet_loss_workspace_bytes() > 0
    assert l.denet_bn_workspace_bytes(1000, 64) > 0
    assert l.denet_build_samples_workspace(2, 64, 64, 1024) == 2 * 4 * 1024 * 4 + 2 * 4 * 4


def test_argument_errors_are_reported_not_thrown():
    l = lib.load()
    rc = l.denet_conv_weight_prep(None, 1, 1, 1, 1, 0, None, None, None)
    assert rc == -1 and b"null pointer" in l.denet_last_error()


def test_product_has_no_oracle_import():
    root = os.path.dirname(lib.LIB_PATH)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(dirpath, f)
