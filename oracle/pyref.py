"""Loader of the REFERENCE ITSELF for generating golden vectors (TEST INFRASTRUCTURE ONLY; this container only —
/root/reference does not exist on the GPU box, so nothing at test time may call this; the fixtures it produces
are committed under tests/golden/ together with the scripts that made them).

``load()`` makes ``import plastid.genomics.map_factories`` & co. resolve to

* the reference's own Cython modules compiled by ``oracle/build_pyref.py`` (``oracle/_ref/pyref``: c_common,
  roitools, map_factories — unmodified sources), and
* the reference's own pure-Python modules read in place from ``/root/reference/plastid`` (genome_array.py,
  genome_hash.py, bin/*.py, util/...),

with stand-ins (``oracle/ref_stubs``) for the third-party imports this image lacks (pysam, Bio, matplotlib,
termcolor, twobitreader) and for the kent-backed BigWig/BigBed readers, which are not on the counting path.
Environment shims the old sources need on Python 3.12 / numpy 2.3: ``builtins.long``, ``numpy.int/float/long``
(aliases numpy removed in 1.24), ``inspect.getargspec``.  No reference file is modified or copied.
"""
import builtins
import importlib
import inspect
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PLASTID_REF", "/root/reference")
OUT = os.path.join(HERE, "_ref", "pyref")
STUBS = os.path.join(HERE, "ref_stubs")

_loaded = False


def available():
    return os.path.isdir(os.path.join(REF, "plastid")) and os.path.isdir(os.path.join(OUT, "plastid", "genomics"))


def load():
    """Idempotent.  Returns the ``plastid`` package shell; raises RuntimeError when the reference is absent."""
    global _loaded
    if _loaded:
        return sys.modules["plastid"]
    if not available():
        raise RuntimeError("reference tree or oracle/_ref/pyref missing: run `python oracle/build_pyref.py` in the build container")
    import numpy
    builtins.long = int
    for name, val in (("int", int), ("float", float), ("long", int)):
        if name not in numpy.__dict__:
            setattr(numpy, name, val)
    if not hasattr(inspect, "getargspec"):
        inspect.getargspec = inspect.getfullargspec
    for p in (STUBS, OUT):          # OUT ends up first: its pysam/ holds the compiled stand-in
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore", message="pkg_resources is deprecated")
    # package shells: the reference's plastid/__init__.py imports every reader and matplotlib; the counting path
    # needs none of that, so the top package is an empty shell whose submodules come from the two trees
    pkg = types.ModuleType("plastid")
    pkg.__path__ = [os.path.join(OUT, "plastid"), os.path.join(REF, "plastid")]
    pkg.__version__ = "0.6.1"
    sys.modules["plastid"] = pkg
    gen = types.ModuleType("plastid.genomics")
    gen.__path__ = [os.path.join(OUT, "plastid", "genomics"), os.path.join(REF, "plastid", "genomics")]
    sys.modules["plastid.genomics"] = gen
    pkg.genomics = gen
    # plotting: colors.py is read in place (roitools imports three helpers from it); plots.py draws with matplotlib
    # and is replaced by a module whose every attribute is a callable that draws nothing
    plo = types.ModuleType("plastid.plotting")
    plo.__path__ = [os.path.join(REF, "plastid", "plotting")]
    sys.modules["plastid.plotting"] = plo
    plots = types.ModuleType("plastid.plotting.plots")

    def _no_plot(name):
        if name.startswith("__"):
            raise AttributeError(name)

        def absorb(*a, **k):                      # figures are not part of any comparison
            from unittest import mock
            axes = mock.MagicMock()                # unpacks as (ax1, ax2) where a caller expects several axes
            axes.__iter__.side_effect = lambda: iter((mock.MagicMock(), mock.MagicMock()))
            return mock.MagicMock(), axes
        return absorb
    plots.__getattr__ = _no_plot
    sys.modules["plastid.plotting.plots"] = plots
    # kent-backed readers (BigWig / BigBed): not on the path; names only
    for mod, names in (("plastid.readers.bigwig", ["BigWigReader"]), ("plastid.readers.bigbed", ["BigBedReader"]),
                       ("plastid.readers.bbifile", ["BBIFile"])):
        m = types.ModuleType(mod)
        for n in names:
            setattr(m, n, type(n, (object,), {"__init__": lambda self, *a, **k: (_ for _ in ()).throw(
                IOError("kent-backed reader stubbed out in oracle/pyref.py"))}))
        sys.modules[mod] = m
    _loaded = True
    return pkg


def modules():
    """The reference modules of the counting path, imported: dict name -> module."""
    load()
    names = ["plastid.genomics.roitools", "plastid.genomics.map_factories", "plastid.genomics.genome_array",
             "plastid.genomics.genome_hash", "plastid.bin.counts_in_region", "plastid.bin.cs", "plastid.bin.metagene",
             "plastid.bin.psite", "plastid.bin.phase_by_size"]
    return {n.split(".")[-1]: importlib.import_module(n) for n in names}
