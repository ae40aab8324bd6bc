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


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
