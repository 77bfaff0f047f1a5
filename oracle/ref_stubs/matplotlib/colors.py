"""Stand-in for matplotlib.colors: hex strings only (what SegmentChain.as_bed hands to colorConverter)."""
import numpy


class _Converter(object):
    def to_rgba(self, c, alpha=None):
        if isinstance(c, str) and c.startswith("#") and len(c) == 7:
            return tuple(int(c[i:i + 2], 16) / 255.0 for i in (1, 3, 5)) + (1.0,)
        if isinstance(c, (tuple, list)) and len(c) in (3, 4):
            return tuple(float(x) for x in c)[:3] + (1.0,)
        raise ValueError("matplotlib stand-in: cannot convert color %r" % (c,))

    def to_rgb(self, c):
        return self.to_rgba(c)[:3]

    def to_rgba_array(self, c, alpha=None):
        return numpy.array([self.to_rgba(c)])


colorConverter = _Converter()


class Normalize(object):
    def __init__(self, *a, **k):
        pass


class LinearSegmentedColormap(object):
    def __init__(self, *a, **k):
        pass
