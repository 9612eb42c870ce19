"""Stand-in: every attribute is a no-op callable (the oracle never plots)."""


class Axes:
    pass


def __getattr__(name):
    def _noop(*args, **kwargs):
        raise RuntimeError("matplotlib stub: plotting is not available in the oracle build")
    return _noop
