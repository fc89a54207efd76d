"""Compiles the reference's ONLY native code — the 1-D (soft-)NMS CPU extension, detection/eval_detection/csrc/nms_cpu.cpp —
from the source WHERE IT LIES under /root/reference into oracle/_ref/nms_1d_cpu.so. TEST INFRASTRUCTURE ONLY.

    python oracle/build_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

The reference builds it with `python setup.py install` (detection/eval_detection/setup.py:7-19: a torch CppExtension named
`nms_1d_cpu`, one source file, -fopenmp). This recipe is the same compilation done by hand with g++ (the reference's own
setup.py is not run; nothing is copied into the repository): torch + pybind11 headers of the interpreter in use, the same module
name (`import nms_1d_cpu`, as detection/eval_detection/nms.py:5 does). The output directory is git-ignored but travels to the GPU
box with the repository snapshot, where the same image (same torch) loads it. Used by tests/ (to pin oracle/nms_oracle.py and as
the checker at sizes the pure-Python oracle is too slow for) and by tools/nms_bench.py as the CPU baseline (kind "reference").
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
SRC = "/root/reference/detection/eval_detection/csrc/nms_cpu.cpp"
OUT = os.path.join(OUT_DIR, "nms_1d_cpu.so")


def build(force: bool = False) -> str:
    """Returns the path of the built module ('' when the reference tree is absent and nothing was built before)."""
    if not os.path.exists(SRC):
        return OUT if os.path.exists(OUT) else ""
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT_DIR, exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}"]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fopenmp", "-DTORCH_EXTENSION_NAME=nms_1d_cpu",
           "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={abi}", *inc, SRC, "-o", OUT,
           f"-L{libdir}", "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        raise RuntimeError("g++ failed on the reference NMS source:\n" + r.stdout[-4000:])
    return OUT


def load():
    """The compiled reference module (nms, softnms), or None when it has not been built."""
    if not os.path.exists(OUT):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("nms_1d_cpu", OUT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv) or "reference tree absent: nothing built")
