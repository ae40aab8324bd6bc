"""Makes the reference's own entry points (train.py / evaluate.py) run on this package unchanged:

    import shapeclipper_b200.shim; shapeclipper_b200.shim.install()

registers `model.renderer`, `model.implicit` and `chamfer_3D` in sys.modules before the reference imports them
(model/graph.py:10-12, utils/eval_3D.py:6)."""
import sys


def install():
    from . import chamfer_3D, implicit, renderer
    sys.modules["model.renderer"] = renderer
    sys.modules["model.implicit"] = implicit
    sys.modules["chamfer_3D"] = chamfer_3D
    return renderer, implicit, chamfer_3D
