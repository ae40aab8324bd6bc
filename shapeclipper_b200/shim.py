"""Makes the reference's own entry points (train.py / evaluate.py) run on this package unchanged:

    import shapeclipper_b200.shim; shapeclipper_b200.shim.install()

registers `model.renderer`, `model.implicit`, `chamfer_3D` and `clip` in sys.modules before the reference imports them
(model/graph.py:10-12, utils/eval_3D.py:6, CLIP_anno.py:7). `clip` is only registered when no real openai/CLIP package is
importable (the reference's annotator then gets this package's image tower: `clip.load(name, device)` -> (model, preprocess))."""
import sys


def install():
    from . import chamfer_3D, implicit, renderer
    sys.modules["model.renderer"] = renderer
    sys.modules["model.implicit"] = implicit
    sys.modules["chamfer_3D"] = chamfer_3D
    if "clip" not in sys.modules:
        import importlib.util
        if importlib.util.find_spec("clip") is None:
            from . import clip
            sys.modules["clip"] = clip
    return renderer, implicit, chamfer_3D
