"""pyplot stand-in: figures, axes and artists are inert objects that accept any method call."""
import sys


class _Inert:
    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        return _Inert()

    def __iter__(self):
        return iter((_Inert(), _Inert()))

    def __getitem__(self, i):
        return _Inert()


_TAB10 = [(0.12, 0.47, 0.71, 1.0), (1.0, 0.5, 0.05, 1.0), (0.17, 0.63, 0.17, 1.0), (0.84, 0.15, 0.16, 1.0), (0.58, 0.4, 0.74, 1.0),
          (0.55, 0.34, 0.29, 1.0), (0.89, 0.47, 0.76, 1.0), (0.5, 0.5, 0.5, 1.0), (0.74, 0.74, 0.13, 1.0), (0.09, 0.75, 0.81, 1.0)]


def get_cmap(name=None, lut=None):
    return lambda i: _TAB10[int(i) % 10]


def figure(*a, **k):
    return _Inert()


def gcf():
    return _Inert()


def gca():
    return _Inert()


def subplots(*a, **k):
    return _Inert(), _Inert()


def __getattr__(name):                      # plt.plot, plt.axis, plt.scatter, plt.savefig, ...: accepted, ignored
    if name.startswith("__"):
        raise AttributeError(name)
    return _Inert()
