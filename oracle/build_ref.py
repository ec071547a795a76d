"""Compile the reference's own native code into oracle/_ref/ (checker only).

The reference's single native component, drift/util/_fast_tools.pyx (Cython + OpenMP),
is cythonized and compiled from where it lies under /root/reference -- no reference
source is copied into the repository; only the built extension lands in oracle/_ref/
(git-ignored, but it travels to the GPU box).  The reference's build system is not run.

The spherical-harmonic transform of the reference lives in healpy/libsharp (external,
not installable offline) and therefore cannot be built here: the CPU baseline uses the
numpy port in oracle/sht.py for that part ("Oracle").
"""

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = "/root/reference/drift/util/_fast_tools.pyx"


def build(verbose=False):
    if not os.path.exists(SRC):
        return None  # GPU box: use the prebuilt extension if it travelled
    import numpy as np

    os.makedirs(OUT, exist_ok=True)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    target = os.path.join(OUT, "_fast_tools" + ext)
    if os.path.exists(target) and os.path.getmtime(target) >= os.path.getmtime(SRC):
        return target
    cfile = os.path.join(OUT, "_fast_tools.c")
    subprocess.check_call([sys.executable, "-m", "cython", "-3", SRC, "-o", cfile],
                          stdout=None if verbose else subprocess.DEVNULL,
                          stderr=None if verbose else subprocess.DEVNULL)
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-w",
                           "-I", sysconfig.get_paths()["include"], "-I", np.get_include(), cfile, "-o", target])
    os.remove(cfile)
    return target


def load():
    """Import the compiled reference extension (or return None if it is not there).
    The .pyx imports ``cora.util.coord.thetaphi_plane_cart`` at module level; ``cora`` is
    not installable, so the oracle's restatement is injected under that name."""
    import glob
    import importlib.util
    import types

    hits = glob.glob(os.path.join(OUT, "_fast_tools*.so"))
    if not hits:
        return None
    from . import beam as obeam

    for name in ("cora", "cora.util", "cora.util.coord"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["cora.util.coord"].thetaphi_plane_cart = obeam.thetaphi_plane_cart
    spec = importlib.util.spec_from_file_location("_fast_tools", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose=True))
