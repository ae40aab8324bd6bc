"""ORACLE support (build container only): compiles the UNMODIFIED reference chamfer3D extension from the
sources where they lie under /root/reference into oracle/_ref/chamfer_3D_ref*.so (git-ignored; it travels to
the GPU box with the gpurun snapshot, where /root/reference does not exist).

It is a torch C++/CUDA extension (ATen tensors + pybind11), so the recipe is two compiler invocations with
torch's include/library paths — the reference's own setup.py is not run, no source is copied.
On the GPU box the module is imported by tests/ and bench.py as the GPU-side checker and as the kernel to beat.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SC_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
NAME = "chamfer_3D_ref"


def build(force=False):
    src_cu = os.path.join(REF, "external", "chamfer3D", "chamfer3D.cu")
    src_cpp = os.path.join(REF, "external", "chamfer3D", "chamfer_cuda.cpp")
    so = os.path.join(OUT, NAME + ".so")
    if not os.path.isfile(src_cu):
        return so if os.path.isfile(so) else None
    if os.path.isfile(so) and not force:
        return so
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT, exist_ok=True)
    inc = []
    for p in ce.include_paths("cuda") + [sysconfig.get_paths()["include"]]:
        inc += ["-isystem", p]
    defs = ["-DTORCH_EXTENSION_NAME=" + NAME, "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    o_cu, o_cpp = os.path.join(OUT, "chamfer3D.o"), os.path.join(OUT, "chamfer_cuda.o")
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                           "-Xcompiler", "-fPIC", "-w"] + defs + inc + ["-c", src_cu, "-o", o_cu])
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-w"] + defs + inc + ["-c", src_cpp, "-o", o_cpp])
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    subprocess.check_call(["g++", "-shared", o_cu, o_cpp, "-o", so, "-L" + libdir, "-L/usr/local/cuda/lib64",
                           "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
                           "-Wl,-rpath," + libdir])
    os.remove(o_cu)
    os.remove(o_cpp)
    return so


def load():
    """Import the compiled reference module (needs torch imported first and a CUDA device to be useful)."""
    import importlib.util
    import torch  # noqa: F401
    so = os.path.join(OUT, NAME + ".so")
    if not os.path.isfile(so):
        return None
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------------------------------
# The reference's Python modules as SOURCELESS bytecode: oracle/_ref/py/{model,utils,data}/*.refbc + CLIP_anno.refbc, compiled
# by py_compile straight from /root/reference (nothing is copied as source; oracle/_ref is git-ignored and, like the .so
# above, travels to the GPU box with the gpurun snapshot — which drops *.pyc, hence the suffix and the small importer below). tests/test_dropin_gpu.py imports them there to run the
# reference's OWN Graph / eval_3D / NN_annotator unshimmed on the CPU and shimmed on the GPU. The option tree is staged as
# data (YAML -> JSON).
PY_OUT = os.path.join(OUT, "py")
PY_TREES = ("model", "utils", "data")
PY_FILES = ("CLIP_anno.py",)
SUFFIX = ".refbc"


def stage_python(force=False):
    import json
    import py_compile
    if not os.path.isdir(os.path.join(REF, "model")):
        return PY_OUT if os.path.isdir(PY_OUT) else None
    n = 0
    todo = [(t, f) for t in PY_TREES for f in sorted(os.listdir(os.path.join(REF, t))) if f.endswith(".py")]
    todo += [("", f) for f in PY_FILES]
    for sub, f in todo:
        src = os.path.join(REF, sub, f)
        dst = os.path.join(PY_OUT, sub, f[:-3] + SUFFIX)
        if os.path.isfile(dst) and not force and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile="<reference>/" + os.path.join(sub, f), doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        n += 1
    import yaml
    for rel in (("options", "pix3d", "config.yaml"), ("options", "clip", "pix3d.yaml")):
        src = os.path.join(REF, *rel)
        dst = os.path.join(PY_OUT, *rel)[:-5] + ".json"
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(src) as fi, open(dst, "w") as fo:
            json.dump(yaml.safe_load(fi), fo)
    return PY_OUT


def staged_available():
    return os.path.isfile(os.path.join(PY_OUT, "model", "renderer" + SUFFIX))


def install_staged_importer():
    """Makes `import model.graph`, `import utils.eval_3D`, `import CLIP_anno` ... resolve to the staged bytecode."""
    import importlib.abc
    import importlib.machinery
    import marshal

    class Loader(importlib.abc.Loader):
        def __init__(self, path):
            self.path = path

        def create_module(self, spec):
            return None

        def exec_module(self, module):
            with open(self.path, "rb") as f:
                code = marshal.loads(f.read()[16:])            # 16-byte pyc header, then the code object
            exec(code, module.__dict__)

    class Finder(importlib.abc.MetaPathFinder):
        def find_spec(self, fullname, path=None, target=None):
            base = os.path.join(PY_OUT, *fullname.split("."))
            if os.path.isfile(base + SUFFIX):
                return importlib.machinery.ModuleSpec(fullname, Loader(base + SUFFIX), origin=base + SUFFIX)
            if os.path.isdir(base) and fullname.split(".")[0] in PY_TREES:
                init = os.path.join(base, "__init__" + SUFFIX)
                spec = importlib.machinery.ModuleSpec(fullname, Loader(init) if os.path.isfile(init) else None,
                                                      origin=base, is_package=True)
                spec.submodule_search_locations = [base]
                return spec
            return None

    if not any(type(f).__name__ == "Finder" and getattr(f, "_sc_staged", False) for f in sys.meta_path):
        f = Finder()
        f._sc_staged = True
        sys.meta_path.insert(0, f)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(stage_python(force="--force" in sys.argv))
