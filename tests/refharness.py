"""Import harness for the UNMODIFIED reference under /root/reference (build container only).

The reference is a Python program, so it can be imported here (CPU) to pin the oracle and to
generate the golden fixtures under tests/golden/. It does not exist on the GPU box: nothing that
runs there may import this module (tests that use it skip when /root/reference is absent).

Missing third-party modules the reference imports at module scope are stubbed (none of them is
called on the render/loss path): termcolor, vigra, mcubes, trimesh, chamfer_3D, matplotlib, seaborn.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(os.path.dirname(_HERE), "oracle", "_ref", "py")   # sourceless bytecode, see oracle/build_ref.py
REF_ROOT = os.environ.get("SC_REFERENCE_ROOT", "/root/reference")
STAGED = False
if not os.path.isfile(os.path.join(REF_ROOT, "model", "renderer.py")) and os.path.isfile(os.path.join(STAGED_ROOT, "model", "renderer.refbc")):
    REF_ROOT, STAGED = STAGED_ROOT, True       # the GPU box: the bytecode build_ref.stage_python() compiled from the reference


def reference_available():
    return STAGED or os.path.isfile(os.path.join(REF_ROOT, "model", "renderer.py"))


def source_available():
    """True only where the reference SOURCES are (the build container): tests that read reference text need this."""
    return os.path.isfile(os.path.join(REF_ROOT, "model", "renderer.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    """Put the reference on sys.path (after stubbing absent deps) and return its key modules."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _stub("termcolor", colored=lambda s, *a, **k: str(s))
    for name in ("vigra", "mcubes", "trimesh", "seaborn"):
        _stub(name)
    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
    if "chamfer_3D" not in sys.modules:
        _stub("chamfer_3D")
    if STAGED:
        sys.path.insert(0, os.path.dirname(_HERE))
        from oracle import build_ref
        build_ref.install_staged_importer()
    elif REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    mods = types.SimpleNamespace()
    mods.implicit = importlib.import_module("model.implicit")
    mods.renderer = importlib.import_module("model.renderer")
    mods.camera = importlib.import_module("utils.camera")
    mods.loss = importlib.import_module("model.loss")
    mods.util = importlib.import_module("utils.util")
    return mods


def load_reference_opt(H=None, W=None, **overrides):
    """EasyDict options from the reference's own YAML (options/pix3d/config.yaml), CPU device."""
    mods = import_reference()
    y = os.path.join(REF_ROOT, "options", "pix3d", "config.yaml")
    if os.path.isfile(y):
        import yaml
        with open(y) as f:
            opt = mods.util.EasyDict(yaml.safe_load(f))
    else:
        import json
        with open(y[:-5] + ".json") as f:
            opt = mods.util.EasyDict(json.load(f))
    opt.device = "cpu"
    opt.H, opt.W = opt.image_size
    if H is not None:
        opt.H = H
    if W is not None:
        opt.W = W
    for k, v in overrides.items():
        node = opt
        keys = k.split(".")
        for kk in keys[:-1]:
            node = node[kk]
        node[keys[-1]] = v
    return opt


def import_reference_graph():
    """model.graph needs torchvision weights=None (reference hard-codes pretrained=True → network)."""
    import torchvision
    mods = import_reference()
    for name in ("resnet18", "resnet34"):
        orig = getattr(torchvision.models, name)
        if getattr(orig, "_sc_patched", False):
            continue

        def patched(*a, _orig=orig, **k):
            k.pop("pretrained", None)
            k["weights"] = None
            return _orig(**k)
        patched._sc_patched = True
        setattr(torchvision.models, name, patched)
    import importlib
    mods.graph = importlib.import_module("model.graph")
    return mods
