#!/usr/bin/env python
"""Build the reference's UNMODIFIED Python package (Cython shim + cmfrec/__init__.py) against libcmfrec_b200, in a
scratch directory outside this repository.  Proves the drop-in boundary: `from cmfrec import CMF` then runs its ALS
fits, MostPopular fits and topN on the GPU library, everything else on the reference's own C code.

    python integration/build_dropin.py [--ref /root/reference] [--out /tmp/cmfrec_dropin]

How it links (nothing is copied into this repo):
  * the reference's src/*.c are compiled into a static helper library with
      -Dfit_collective_explicit_als=ref_fit_collective_explicit_als   (same for ..._implicit_als, fit_most_popular, topN)
    so that every entry point this repo does not replace still resolves to reference code, and the replaced ones
    remain reachable under ref_* names;
  * the Cython shim (cmfrec/cfuns_{double,float}_plusblas.pyx, which includes wrapper_untyped.pxi) is compiled WITHOUT
    those defines and linked against cmfrec_b200/lib/libcmfrec_b200_{f64,f32}.so, so its calls to the four entry points
    bind to this repo's symbols (reference call sites: cmfrec/wrapper_untyped.pxi:1193, 1383, 2549, 2876).
"""
import argparse
import os
import shutil
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPLACED = ["fit_collective_explicit_als", "fit_collective_implicit_als", "fit_most_popular", "topN", "predict_multiple",
            "predict_X_old_collective_explicit"]

SETUP = '''
import numpy as np
from setuptools import setup, Extension
from Cython.Distutils import build_ext
LIBDIR = {libdir!r}
RENAMES = [(name, "ref_" + name) for name in {replaced!r}]
COMMON = [("_FOR_PYTHON", None), ("NDEBUG", None), ("NO_FINDBLAS", None), ("AVOID_BLAS_SYR", None)]
CFILES = ["collective.c", "common.c", "offsets.c", "helpers.c", "lbfgs.c", "cblas_wrappers.c"]
CFLAGS = ["-O3", "-fopenmp", "-std=c99", "-fno-math-errno", "-fno-trapping-math", "-ffp-contract=fast", "-fPIC", "-w"]
libs, exts = [], []
for tag, macro, prec in (("double", "USE_DOUBLE", "f64"), ("float", "USE_FLOAT", "f32")):
    # one copy of the C sources per precision (src_double/, src_float/): build_clib names object files after their source
    # path, so two libraries built from the SAME paths would silently share the objects of whichever was compiled first
    sdir = "src_" + tag
    libs.append(("cmfrec_ref_" + tag, dict(sources=[sdir + "/" + f for f in CFILES], include_dirs=[sdir],
                                           macros=COMMON + [(macro, None)] + RENAMES, cflags=CFLAGS)))
    exts.append(Extension("cmfrec.wrapper_" + tag, sources=["cmfrec/cfuns_%s_plusblas.pyx" % tag],
                          include_dirs=[np.get_include(), "src"], define_macros=COMMON + [(macro, None)],
                          libraries=["cmfrec_ref_" + tag, "cmfrec_b200_" + prec, "gomp"], library_dirs=[LIBDIR],
                          runtime_library_dirs=[LIBDIR], extra_compile_args=["-O2", "-fopenmp", "-w"],
                          extra_link_args=["-fopenmp"]))
setup(name="cmfrec", packages=["cmfrec"], libraries=libs, ext_modules=exts, cmdclass=dict(build_ext=build_ext))
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default="/tmp/cmfrec_dropin")
    ap.add_argument("--install", default=None, help="copy the importable package (cmfrec/__init__.py + the two extension "
                    "modules, nothing else) into this directory, e.g. integration/_dropin (git-ignored, shipped to the GPU box)")
    a = ap.parse_args()
    if os.path.exists(a.out):
        shutil.rmtree(a.out)
    shutil.copytree(a.ref, a.out, ignore=shutil.ignore_patterns(".git", "docs", "benchmark", "R", "man", "vignettes"))
    subprocess.run(["chmod", "-R", "u+w", a.out], check=True)
    for tag in ("double", "float"):
        shutil.copytree(os.path.join(a.out, "src"), os.path.join(a.out, "src_" + tag))
    open(os.path.join(a.out, "setup_dropin.py"), "w").write(
        textwrap.dedent(SETUP).format(libdir=os.path.join(ROOT, "cmfrec_b200", "lib"), replaced=REPLACED))
    env = dict(os.environ, CC="/usr/bin/gcc", LDSHARED="/usr/bin/gcc -shared")
    subprocess.run([sys.executable, "setup_dropin.py", "build_clib", "build_ext", "--inplace"], cwd=a.out, env=env, check=True)
    print("built; use it with  PYTHONPATH=%s  (import cmfrec)" % a.out)
    if a.install:
        dst = os.path.join(a.install, "cmfrec")
        os.makedirs(dst, exist_ok=True)
        for name in os.listdir(os.path.join(a.out, "cmfrec")):
            if name == "__init__.py" or (name.startswith("wrapper_") and name.endswith(".so")):
                shutil.copy2(os.path.join(a.out, "cmfrec", name), os.path.join(dst, name))
        print("installed into %s" % a.install)


if __name__ == "__main__":
    main()
