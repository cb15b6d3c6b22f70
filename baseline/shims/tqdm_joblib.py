import contextlib
import joblib


def ParallelPbar(desc=None):
    return lambda **kw: joblib.Parallel(**kw)


@contextlib.contextmanager
def tqdm_joblib(*a, **k):
    yield
