"""Serial stand-in for pathos.multiprocessing.ProcessingPool."""


class ProcessingPool:
    def __init__(self, *args, **kwargs):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def map(self, fn, *iterables):
        return [fn(*a) for a in zip(*iterables)]

    def close(self):
        pass

    def join(self):
        pass

    def clear(self):
        pass
