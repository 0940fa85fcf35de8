def __getattr__(name):
    def _inert(*a, **k):
        raise RuntimeError(f"matplotlib.pyplot.{name} called: plotting is outside the oracle's scope")
    return _inert
