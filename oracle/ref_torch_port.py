"""Op-for-op torch restatement of the reference's CPU implementation of the benchmarked path
-- TEST / BENCH INFRASTRUCTURE ONLY (bench.py cpu_baseline and --impl reference legs).

/root/reference does not travel to the GPU box, so the reference's own torch path
(attack/CW/CW_utils/distance.py:15-70) is restated here with the same ATen op sequence: three
torch.bmm (two of them only to read their diagonals), diagonal gather, broadcast add,
torch.min along both axes, mean / max.  Same memory behaviour ([B,N,M] matrices materialised and
saved for autograd), same arithmetic (checked bit-exact against the reference in
tests/test_oracle_golden.py::test_ref_torch_port_matches_golden).
"""
import torch


def batch_pairwise_dist(x, y):
    """distance.py:15-32."""
    _, num_points_x, _ = x.size()
    _, num_points_y, _ = y.size()
    xx = torch.bmm(x, x.transpose(2, 1))
    yy = torch.bmm(y, y.transpose(2, 1))
    zz = torch.bmm(x, y.transpose(2, 1))
    diag_ind_x = torch.arange(0, num_points_x, device=x.device)
    diag_ind_y = torch.arange(0, num_points_y, device=x.device)
    rx = xx[:, diag_ind_x, diag_ind_x].unsqueeze(1).expand_as(zz.transpose(2, 1))
    ry = yy[:, diag_ind_y, diag_ind_y].unsqueeze(1).expand_as(zz)
    return rx.transpose(2, 1) + ry - 2 * zz


def chamfer_distance(preds, gts):
    """distance.py:40-50."""
    P = batch_pairwise_dist(gts, preds)
    mins, _ = torch.min(P, 1)
    loss1 = torch.mean(mins, dim=1)
    mins, _ = torch.min(P, 2)
    loss2 = torch.mean(mins, dim=1)
    return loss1, loss2


def hausdorff_distance(preds, gts):
    """distance.py:58-70."""
    P = batch_pairwise_dist(gts, preds)
    mins, _ = torch.min(P, 1)
    loss1 = torch.max(mins, dim=1)[0]
    mins, _ = torch.min(P, 2)
    loss2 = torch.max(mins, dim=1)[0]
    return loss1, loss2


def chamfer_hausdorff_fwd_bwd(adv, ori):
    """One benchmark step of the reference path: Chamfer + Hausdorff (both directions) forward
    and backward w.r.t. adv -- the reference builds the pair matrix twice."""
    adv = adv.detach().requires_grad_(True)
    c1, c2 = chamfer_distance(adv, ori)
    h1, h2 = hausdorff_distance(adv, ori)
    loss = (c1 + c2 + h1 + h2).sum()
    loss.backward()
    return torch.stack([c1, c2, h1, h2]).detach(), adv.grad
