#!/usr/bin/env python
"""Compile the reference's OWN Cython sources of the counting path, where they lie under /root/reference,
into oracle/_ref/pyref/ (git-ignored; TEST INFRASTRUCTURE ONLY, see oracle/ref_stubs/README.md).

  python oracle/build_pyref.py            # idempotent; prints "pyref: built" / "pyref: up to date" / the reason it cannot

What is built, from unmodified sources:
  plastid/genomics/c_common.pyx, roitools.pyx, map_factories.pyx  (compile-time env PYSAM10=True, directives
  embedsignature + language_level 3 as in the reference's setup.py:236-239, 300-318)
against the stand-in `pysam.libcalignmentfile` of oracle/ref_stubs (pysam itself is absent and does no arithmetic on
this path beyond the CIGAR walk, which is pinned separately to the reference's vendored htslib).
No reference source is copied into the repo: Cython reads the .pyx/.pxd files in place and writes C files and
extension modules only under oracle/_ref/pyref/.
"""
import glob
import os
import re
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PLASTID_REF", "/root/reference")
OUT = os.path.join(HERE, "_ref", "pyref")
STUBS = os.path.join(HERE, "ref_stubs")
MODULES = ["plastid.genomics.c_common", "plastid.genomics.roitools", "plastid.genomics.map_factories"]


def ext_path(modname, root):
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return os.path.join(root, *modname.split(".")) + suffix


def up_to_date():
    targets = [ext_path(m, OUT) for m in MODULES] + [ext_path("pysam.libcalignmentfile", OUT)]
    if not all(os.path.exists(t) for t in targets):
        return False
    newest_src = max(os.path.getmtime(p) for p in glob.glob(os.path.join(REF, "plastid", "genomics", "*.p*"))
                     + glob.glob(os.path.join(STUBS, "pysam", "*")) + [os.path.abspath(__file__)])
    return min(os.path.getmtime(t) for t in targets) >= newest_src


def build():
    if not os.path.isdir(os.path.join(REF, "plastid", "genomics")):
        print("pyref: reference tree absent (%s): nothing to build" % REF)
        return False
    if up_to_date():
        print("pyref: up to date")
        return True
    import numpy
    from setuptools import Extension
    from setuptools.dist import Distribution
    from Cython.Build import cythonize
    os.makedirs(OUT, exist_ok=True)
    build_c = os.path.join(OUT, "_c")
    # numpy >= 2 dropped `int_t` / `long_t` from its Cython declarations; the reference's .pyx files name them
    # (map_factories.pyx:151-154).  A build-time copy of numpy's own __init__.pxd with the two typedefs restored
    # (both were `npy_long`) is put in front of the include path — numpy's file, not the reference's, is what is
    # adjusted.
    shim = os.path.join(OUT, "_pxd", "numpy")
    os.makedirs(shim, exist_ok=True)
    with open(os.path.join(os.path.dirname(numpy.__file__), "__init__.pxd")) as fh:
        pxd = fh.read()
    if "ctypedef npy_long int_t" not in pxd and not re.search(r"ctypedef\s+npy_long\s+int_t", pxd):
        pxd += "\nctypedef npy_long int_t\nctypedef npy_long long_t\nctypedef npy_ulong uint_t\nctypedef npy_ulong ulong_t\n"
    with open(os.path.join(shim, "__init__.pxd"), "w") as fh:
        fh.write(pxd)
    exts = [Extension("pysam.libcalignmentfile", [os.path.join(STUBS, "pysam", "libcalignmentfile.pyx")])]
    for m in MODULES:
        src = os.path.join(REF, *m.split(".")) + ".pyx"
        exts.append(Extension(m, [src], include_dirs=[numpy.get_include()],
                              define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")]))
    # the sources still use the Python 2 builtin `long` (roitools.pyx:554): let unknown names resolve at run time;
    # the loader (oracle/pyref.py) provides `builtins.long = int`
    import Cython.Compiler.Options as cy_options
    cy_options.error_on_unknown_names = False
    cwd = os.getcwd()
    os.chdir(REF)          # module names are derived from paths relative to the package root
    try:
        exts = cythonize(exts, build_dir=build_c, include_path=[os.path.dirname(shim), STUBS, REF],
                         compile_time_env={"PYSAM10": True},
                         compiler_directives={"embedsignature": True, "language_level": 3},
                         quiet=True, force=True)
    finally:
        os.chdir(cwd)
    dist = Distribution({"ext_modules": exts})
    cmd = dist.get_command_obj("build_ext")
    cmd.build_lib = OUT
    cmd.build_temp = os.path.join(OUT, "_o")
    cmd.ensure_finalized()
    cmd.run()
    # the stand-in's python half travels next to its extension module
    import shutil
    for name in ("__init__.py", "libctabix.py"):
        shutil.copy(os.path.join(STUBS, "pysam", name), os.path.join(OUT, "pysam", name))
    print("pyref: built %s" % ", ".join(MODULES))
    return True


if __name__ == "__main__":
    try:
        ok = build()
    except Exception as exc:           # a checker that cannot be built must not break build()
        print("pyref: build failed: %r" % (exc,))
        ok = False
    sys.exit(0 if ok else 1)
