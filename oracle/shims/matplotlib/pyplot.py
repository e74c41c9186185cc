"""No-op pyplot: plotting is visualisation only (out of scope)."""


class _Noop:
    def __getattr__(self, name):
        return _noop

    def __call__(self, *a, **k):
        return self

    def __iter__(self):
        return iter(())


def _noop(*a, **k):
    return _Noop()


def __getattr__(name):
    return _noop
