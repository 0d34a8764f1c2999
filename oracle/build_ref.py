"""Build the reference's own Cython sampler kernels into ``oracle/_ref/``.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pyblip_b200/`` may import this.

The reference (amspector100/pyblip v0.4.2) is a Python package whose hot
loops are five Cython extension modules (``/root/reference/setup.py:33-55``).
This recipe compiles them *from the sources where they lie* under
``/root/reference`` (read-only), writes every intermediate file to a scratch
directory under ``/tmp`` and copies only the resulting ``.so`` binaries into
``oracle/_ref/pyblip/...`` (git-ignored, but shipped to the GPU box by
``gpurun``).  No reference source file is copied into the repository; the
``__init__.py`` files written below are empty package markers of our own (the
reference's ``pyblip/__init__.py`` imports cvxpy, which is not installed).

Usage:  python oracle/build_ref.py [--force]
"""
import os
import shutil
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("PYBLIP_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

# (module name, .pyx path relative to the reference root) -- setup.py:33-55
EXTENSIONS = [
    ("pyblip.cython_utils._truncnorm", "pyblip/cython_utils/_truncnorm.pyx"),
    ("pyblip.cython_utils._update_hparams", "pyblip/cython_utils/_update_hparams.pyx"),
    ("pyblip.linear._linear", "pyblip/linear/_linear.pyx"),
    ("pyblip.linear._linear_multi", "pyblip/linear/_linear_multi.pyx"),
    ("pyblip.nprior._nprior", "pyblip/nprior/_nprior.pyx"),
]
PACKAGES = ["pyblip", "pyblip/cython_utils", "pyblip/linear", "pyblip/nprior",
            "pyblip/probit"]


def _so_path(modname):
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return os.path.join(OUT, *modname.split(".")) + suffix


def is_built():
    return all(os.path.exists(_so_path(m)) for m, _ in EXTENSIONS)


def build(force=False, verbose=False):
    """Compile the reference extensions. Returns True if oracle/_ref is usable."""
    if is_built() and not force:
        return True
    if not os.path.isdir(os.path.join(REF_SRC, "pyblip")):
        return False  # GPU box: only the prebuilt files exist
    import numpy as np
    from Cython.Build import cythonize
    from setuptools import Extension
    from setuptools.dist import Distribution

    scratch = tempfile.mkdtemp(prefix="pyblip_ref_build_")
    cwd = os.getcwd()
    try:
        # cythonize wants paths relative to cwd for sane module layout
        os.chdir(REF_SRC)
        exts = [Extension(name, sources=[src]) for name, src in EXTENSIONS]
        exts = cythonize(
            exts,
            compiler_directives={"language_level": 3, "embedsignature": True},
            build_dir=os.path.join(scratch, "cy"),
            quiet=not verbose,
        )
        dist = Distribution({"ext_modules": exts,
                             "include_dirs": [np.get_include()]})
        cmd = dist.get_command_obj("build_ext")
        cmd.build_temp = os.path.join(scratch, "tmp")
        cmd.build_lib = os.path.join(scratch, "lib")
        cmd.include_dirs = [np.get_include()]
        cmd.ensure_finalized()
        cmd.run()
        for pkg in PACKAGES:
            d = os.path.join(OUT, pkg)
            os.makedirs(d, exist_ok=True)
            init = os.path.join(d, "__init__.py")
            if not os.path.exists(init):
                with open(init, "w") as f:
                    f.write("# empty package marker written by oracle/build_ref.py\n")
        for name, _ in EXTENSIONS:
            rel = os.path.relpath(_so_path(name), OUT)
            shutil.copy2(os.path.join(scratch, "lib", rel), _so_path(name))
    finally:
        os.chdir(cwd)
        shutil.rmtree(scratch, ignore_errors=True)
    return is_built()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("oracle/_ref built" if ok else "oracle/_ref NOT built (no reference sources)")
    sys.exit(0 if ok else 1)
