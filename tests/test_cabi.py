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
    assert l.denet_build_samples_workspace(2, 64, 64, 1024) == 2 * 5 * 1024 * 4 + 2 * 5 * 4   # sized for 5 corner maps (DNC.C)


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


def test_wgrad_split_plan_fills_whole_waves():
    """the split-K factor of the filter gradient (plan query of the C-ABI, no GPU needed: 148 SMs assumed) puts the
    DeNet-34 layers on whole waves of the persistent grid - the former rule ceil(2*SMs / tiles) left 297..306 tiles on
    148 SMs, i.e. three rounds for two waves of work"""
    from denet_b200 import lib
    L = lib.load()
    sms = 148
    # (N, Ho, Wo, Cout, Cin, k): the stride-1 3x3 layers of the ResNet-34 stages at 512 x 512 input, batch 32
    for (n, ho, wo, cout, cin, k) in [(32, 128, 128, 64, 64, 3), (32, 64, 64, 128, 128, 3), (32, 32, 32, 256, 256, 3),
                                      (32, 16, 16, 512, 512, 3)]:
        splits = L.denet_conv2d_wgrad_splits(n, ho, wo, cout, cin, k, k, 1, 1)
        bn = 64 if cin <= 64 else (128 if cin <= 128 else 256)
        rows_kernel = cin <= 64                                  # one tile per filter ROW, else one per tap
        base = -(-cout // 128) * -(-cin // bn) * (k if rows_kernel else k * k)
        tiles = base * splits
        rounds = -(-tiles // sms)
        assert tiles / (rounds * sms) >= 0.95, (cout, cin, splits, tiles)
        ws = L.denet_conv2d_wgrad_workspace(n, ho, wo, cout, cin, k, k)
        assert ws >= splits * cout * cin * k * k * 4
    # the row-folded stem: every filter row in one tile -> splits = persistent CTAs
    assert L.denet_conv2d_rowfold_wgrad_splits(32, 256, 256, 64, 7, 2) == sms
