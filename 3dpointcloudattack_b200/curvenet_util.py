"""Drop-in for model/curvenet_util.py:10-26 (CurveNet kNN on xyz)."""
from .dgcnn import knn as _knn


def knn(x, k):
    """curvenet_util.py:10-17: returns k + 1 columns."""
    return _knn(x, k + 1)


def normal_knn(x, k):
    """curvenet_util.py:20-26."""
    return _knn(x, k)
