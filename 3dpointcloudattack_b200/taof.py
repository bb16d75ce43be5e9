"""Drop-in for the graph helpers of attack/AOF/TAOF_attack.py:13-52 (same in attack/AOF/Eval_AOF.py).

`knn` is model/dgcnn.py's formulation (see dgcnn.knn).  `get_Laplace_from_pc` assembles the dense
graph Laplacian of the symmetrised 30-NN graph with Gaussian weights in one kernel pass over the index
tensor instead of materialising [B,N,N,3] differences, a scatter mask and its transpose; the dense
eigendecomposition stays with torch (`torch.linalg.eigh` replaces the removed `torch.symeig`; it is a
LAPACK/cuSOLVER call, not part of the point-set distance path).
"""
import torch

from . import functional as F
from .dgcnn import knn  # noqa: F401


def laplacian_from_pc(ori_pc, k=30):
    """ori_pc [B,3,N] -> L [B,N,N] = D - A (TAOF_attack.py:36-50)."""
    pc = ori_pc.detach()
    pts = pc.transpose(2, 1)
    _, idx = F.knn(pts, pts, k, form=F.FORM_COL_ROW, norm=F.NORM_MULSUM)
    return F.graph_laplacian(pts, idx)


def get_Laplace_from_pc(ori_pc):
    """TAOF_attack.py:31-52 -> (eigenvalues [B,N] ascending, eigenvectors [B,N,N])."""
    with torch.no_grad():
        L = laplacian_from_pc(ori_pc, 30)
        e, v = torch.linalg.eigh(L)
    return e.to(ori_pc), v.to(ori_pc)
