"""No-op stand-in for ``matplotlib`` (absent here): the reference scripts only plot loss curves and
confusion matrices to files (ref:ssp_vit2spn_tiny.py:11, ref:octmnist_ft_vit2spn.py:160-167)."""


class _Null:
    def __getattr__(self, name):
        return self

    def __call__(self, *a, **k):
        return self

    def __iter__(self):
        return iter((self, self))

    def __getitem__(self, i):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __float__(self):
        return 0.0


__version__ = "0.0-v2s-stub"
cm = _Null()
rcParams = {}


def use(*a, **k):
    return None
