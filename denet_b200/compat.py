"""`import denet...` compatibility: lets the reference's OWN drivers (bin/model-train -> denet/model/train.py) run on
the B200 path without editing them.

    python -m denet_b200.compat /path/to/lachlants-denet denet/model/train.py --train ... --model-desc ...

install_alias() registers a meta-path finder that resolves

    denet.layer[.*], denet.model.model_cnn, denet.multi, denet.common[.json_util]   ->  denet_b200.<same path>
    everything else under `denet.` (dataset loaders, logging, the driver scripts)     ->  the reference tree, unmodified

The hot path (layers, ModelCNN, the native extension) is ours; the control plane around it (datasets, logging,
argument parsing) stays the reference's own code, which is exactly the drop-in boundary SURVEY.md §8(b) names.  The
reference drivers `import theano` at module level without using it (model/train.py:10); with stub_theano=True an empty
module satisfies that import when Theano is not installed (nothing on the B200 path touches it).
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

# reference module -> denet_b200 module.  Prefix matches: 'denet.layer.convolution' -> 'denet_b200.layer.convolution'
_OURS = ("denet.layer", "denet.model.model_cnn", "denet.multi")
# modules of denet.common that exist here; anything else in denet.common (logging, image_util) is the reference's
_OURS_EXACT = ("denet.common.json_util",)


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        return importlib.import_module(self.target)

    def exec_module(self, module):
        pass


class _CommonLoader(importlib.abc.Loader):
    """denet.common = the reference's package (logging, import_c ...) with the hot-path helpers overridden by ours and
    import_c() handing out the GPU-backed extension modules instead of JIT-compiling the reference's .cc files"""

    def __init__(self, ref_init):
        self.ref_init = ref_init

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        from . import common as ours
        if self.ref_init is not None:
            with open(self.ref_init) as f:
                src = f.read()
            try:
                exec(compile(src, self.ref_init, "exec"), module.__dict__)
            except ImportError:
                pass            # optional imports of the reference's helpers (e.g. theano-only utilities)
            module.import_c = ours.import_c       # the native extensions are the GPU-backed ones, not a JIT build
            return
        for k, v in vars(ours).items():
            if not k.startswith("_"):
                setattr(module, k, v)


class DenetAliasFinder(importlib.abc.MetaPathFinder):
    def __init__(self, reference_root=None):
        self.ref = os.path.join(reference_root, "denet") if reference_root else None

    def _ref_path(self, fullname):
        if self.ref is None:
            return None, False
        rel = fullname.split(".")[1:]
        base = os.path.join(self.ref, *rel)
        if os.path.isdir(base) and os.path.exists(os.path.join(base, "__init__.py")):
            return os.path.join(base, "__init__.py"), True
        if os.path.exists(base + ".py"):
            return base + ".py", False
        return None, False

    def find_spec(self, fullname, path=None, target=None):
        if fullname != "denet" and not fullname.startswith("denet."):
            return None
        if fullname == "denet":
            spec = importlib.machinery.ModuleSpec("denet", _PackageLoader(), is_package=True)
            spec.submodule_search_locations = []
            return spec
        if fullname == "denet.common":
            ref_init, _ = self._ref_path(fullname)
            spec = importlib.machinery.ModuleSpec(fullname, _CommonLoader(ref_init), is_package=True)
            spec.submodule_search_locations = [os.path.dirname(ref_init)] if ref_init else []
            return spec
        if fullname == "denet.model":
            # a namespace of its own: model_cnn is ours, the driver scripts (train.py ...) are the reference's
            spec = importlib.machinery.ModuleSpec(fullname, _PackageLoader(), is_package=True)
            spec.submodule_search_locations = [os.path.join(self.ref, "model")] if self.ref else []
            return spec
        for prefix in _OURS:
            if fullname == prefix or fullname.startswith(prefix + "."):
                return importlib.machinery.ModuleSpec(fullname, _AliasLoader("denet_b200" + fullname[5:]),
                                                      is_package=(fullname in ("denet.layer", "denet.multi")))
        if fullname in _OURS_EXACT:
            return importlib.machinery.ModuleSpec(fullname, _AliasLoader("denet_b200" + fullname[5:]))
        fpath, is_pkg = self._ref_path(fullname)
        if fpath is None:
            return None
        return importlib.util.spec_from_file_location(
            fullname, fpath, submodule_search_locations=[os.path.dirname(fpath)] if is_pkg else None)


class _PackageLoader(importlib.abc.Loader):
    def create_module(self, spec):
        return None

    def exec_module(self, module):
        pass


_installed = None


def _legacy_environment_shims():
    """the reference is 2017-era code: names it uses that today's libraries dropped (its control plane, not ours)"""
    try:
        from PIL import Image
        if not hasattr(Image, "ANTIALIAS"):
            Image.ANTIALIAS = Image.LANCZOS          # dataset/augment.py:20; removed in Pillow 10
    except ImportError:
        pass


def install_alias(reference_root=None, stub_theano=False):
    """make `import denet...` resolve as described in the module docstring; idempotent"""
    global _installed
    if _installed is not None:
        sys.meta_path.remove(_installed)
        for name in [m for m in sys.modules if m == "denet" or m.startswith("denet.")]:
            del sys.modules[name]
    if reference_root:
        _legacy_environment_shims()
    _installed = DenetAliasFinder(reference_root)
    sys.meta_path.insert(0, _installed)
    if stub_theano and "theano" not in sys.modules:
        try:
            importlib.import_module("theano")
        except ImportError:
            sys.modules["theano"] = types.ModuleType("theano")     # imported, never used, by the reference drivers
    return _installed


def uninstall_alias():
    global _installed
    if _installed is not None:
        sys.meta_path.remove(_installed)
        _installed = None
    for name in [m for m in sys.modules if m == "denet" or m.startswith("denet.")]:
        del sys.modules[name]
    t = sys.modules.get("theano")
    if t is not None and getattr(t, "__file__", None) is None:
        del sys.modules["theano"]


def run_reference_script(reference_root, script, argv):
    """run one of the reference's driver scripts (path relative to the reference root, e.g. denet/model/train.py) as
    __main__ with `denet` aliased"""
    import runpy
    install_alias(reference_root, stub_theano=True)
    old = sys.argv
    sys.argv = [os.path.join(reference_root, script)] + list(argv)
    try:
        runpy.run_path(sys.argv[0], run_name="__main__")
    finally:
        sys.argv = old


if __name__ == "__main__":
    if len(sys.argv) < 3:
        raise SystemExit(__doc__)
    run_reference_script(sys.argv[1], sys.argv[2], sys.argv[3:])
