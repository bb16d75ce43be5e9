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


# ------------------------------------------------------------------ CW iteration (CPU baseline)
class _ChamferHausdorffAvg:
    """dist_func used by the CW benchmark: ChamferDist('avg') + HausdorffDist('avg') with per-sample
    weights (attack/CW/CW_utils/dist_utils.py:49-109), CPU restatement."""

    def __call__(self, adv_pc, ori_pc, weights):
        c1, c2 = chamfer_distance(adv_pc, ori_pc)
        h1, h2 = hausdorff_distance(adv_pc, ori_pc)
        loss = ((c1 + c2) / 2. + (h1 + h2) / 2.) * weights.float()
        return loss.mean()


def cw_iterations_cpu(model, data, target, num_iter, attack_lr=1e-2, init_weight=10., kappa=30., budget=0.18):
    """`num_iter` iterations of the reference's CW loop body (attack/CW/CW_attack.py:111-178) on CPU,
    B = data.shape[0] (the reference runs B = 1), including its host-side best tracking."""
    import numpy as np
    B, K = data.shape[:2]
    ori = data.float().transpose(1, 2).contiguous().detach()
    adv = (ori.clone() + torch.randn((B, 3, K)) * 1e-7).requires_grad_(True)
    opt = torch.optim.Adam([adv], lr=attack_lr, weight_decay=0.)
    weight = np.ones((B,)) * init_weight
    bestdist = np.array([1e10] * B); bestscore = np.array([-1] * B)
    dist_func = _ChamferHausdorffAvg()
    label_val = target.numpy()
    for _ in range(num_iter):
        logits = model(adv)[0]
        pred = torch.argmax(logits, dim=1)
        dist_val = torch.sqrt(torch.sum((adv - ori) ** 2, dim=[1, 2])).detach().numpy()
        pred_val = pred.detach().numpy()
        for e, (dist, p, label) in enumerate(zip(dist_val, pred_val, label_val)):
            if dist < bestdist[e] and p != label:
                bestdist[e] = dist; bestscore[e] = p
        one_hot = torch.zeros_like(logits).scatter_(1, target.view(-1, 1), 1.)
        real = torch.sum(one_hot * logits, dim=1)
        other = torch.max((1. - one_hot) * logits - one_hot * 10000., dim=1)[0]
        adv_loss = torch.clamp(real - other + kappa, min=0.).mean()
        dist_loss = dist_func(adv.transpose(1, 2).contiguous(), ori.transpose(1, 2).contiguous(),
                              torch.from_numpy(weight))
        loss = adv_loss + dist_loss
        opt.zero_grad()
        loss.backward()
        opt.step()
        with torch.no_grad():
            diff = adv - ori
            norm = torch.sum(diff ** 2, dim=1) ** 0.5
            scale = torch.clamp(budget / (norm + 1e-9), max=1.)
            adv.data = ori + diff * scale[:, None, :]
    return float(loss.detach())
