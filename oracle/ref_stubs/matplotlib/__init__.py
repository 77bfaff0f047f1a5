"""Stand-in for matplotlib (see ../README.md): nothing here draws.  Every submodule is a mock that absorbs the
figure-drawing calls the reference's programs make after they have written their tables."""
import sys
import types
from unittest import mock

__version__ = "0.0"


class _Cycle(object):
    def by_key(self):
        return {"color": ["#000000", "#222222", "#444444"]}


class _RcParams(dict):
    def __missing__(self, key):
        if key == "axes.prop_cycle":
            return _Cycle()
        return mock.MagicMock()


rcParams = _RcParams()


def use(*args, **kwargs):
    pass


class _Absorb(mock.MagicMock):
    """MagicMock whose `subplots()` unpacks into (figure, axes)."""

    def subplots(self, *a, **k):
        return mock.MagicMock(), mock.MagicMock()


def _submodule(name):
    m = _Absorb()
    m.__name__ = "matplotlib." + name
    m.__file__ = __file__
    m.__path__ = []
    sys.modules["matplotlib." + name] = m
    return m


for _n in ("pyplot", "cm", "gridspec", "patches", "ticker", "artist", "figure", "axes", "collections", "lines", "path",
           "transforms", "backend_bases", "backends", "font_manager", "text"):
    globals()[_n] = _submodule(_n)

style = types.ModuleType("matplotlib.style")
style.available = []
style.use = lambda *a, **k: None
sys.modules["matplotlib.style"] = style
