"""Drop-in for model/curvenet_util.py:10-26 (CurveNet kNN on xyz), the gather stage of
LPFA.group_feature (:206-236) and farthest_point_sample (:69-90, start index 0)."""
from . import functional as F
from .dgcnn import knn as _knn


def knn(x, k):
    """curvenet_util.py:10-17: returns k + 1 columns."""
    return _knn(x, k + 1)


def normal_knn(x, k):
    """curvenet_util.py:20-26."""
    return _knn(x, k)


def lpfa_point_feature(xyz, idx):
    """curvenet_util.py:219-227: xyz[B,3,N], idx[B,N,k] -> cat(points, neighbours, neighbours - points)
    as [B,9,N,k]."""
    return F.edge_feature(xyz, idx, (F.EDGE_CENTER, F.EDGE_NEIGHBOR, F.EDGE_DIFF))


def lpfa_feature_diff(x, idx):
    """curvenet_util.py:229-234: x[B,C,N] -> (neighbour feature - x) as [B,C,N,k]."""
    return F.edge_feature(x, idx, (F.EDGE_DIFF,))


def group_feature(self, x, xyz, idx):
    """Replacement body for LPFA.group_feature (curvenet_util.py:206-236); `self` is the
    reference's LPFA module (k, initial, xyz2feature are read from it)."""
    import torch.nn.functional as nnF
    if idx is None:
        idx = knn(xyz, k=self.k)[:, :, :self.k]
    point_feature = lpfa_point_feature(xyz, idx)
    if self.initial:
        return point_feature
    feature = lpfa_feature_diff(x, idx)
    point_feature = self.xyz2feature(point_feature)
    return nnF.leaky_relu(feature + point_feature, 0.2)


def farthest_point_sample(xyz, npoint):
    """curvenet_util.py:69-90 (start index 0) -> centroids [B,npoint] int64."""
    return F.farthest_point_sample(xyz, npoint, None).long()
