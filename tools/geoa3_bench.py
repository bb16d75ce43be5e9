"""GeoA3-shaped geometry loss (attack/GeoA3/GeoA3_attack.py:103-183: CD + 0.1 HD + curvature, k=16) forward+backward:
this package's loss_utils (one cached NN-1 sweep serves the 4-6 repeats of the adv->ori query) vs. the reference's
torch formulation (knn_utils.py matmul + topk) on the same GPU.  Development tool."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
pcd = importlib.import_module("3dpointcloudattack_b200")
synth = importlib.import_module("3dpointcloudattack_b200.synth")
from knn_bench import timeit  # noqa: E402
LU = pcd.loss_utils


def ref_knn_points(p1, p2, K):                       # attack/GeoA3/knn_utils.py:10-55
    inner = -2 * torch.matmul(p1, p2.transpose(2, 1))
    p1_2 = torch.sum(p1 ** 2, dim=2, keepdim=True).transpose(2, 1)
    p2_2 = torch.sum(p2 ** 2, dim=2, keepdim=True).transpose(2, 1)
    dist = p1_2 + inner + p2_2.transpose(2, 1)
    v, i = (-dist).topk(K, dim=-1)
    return -v, i


def ref_gather(x, idx):                              # knn_gather: x[B,M,U], idx[B,L,K] -> [B,L,K,U]
    B, M, U = x.shape
    _, L, K = idx.shape
    return x[:, :, None].expand(B, M, K, U).gather(1, idx[:, :, :, None].expand(B, L, K, U))


def ref_loss(adv, ori, normal, ori_kappa, k=16):     # loss_utils.py:36-105 composition
    a, o = adv.permute(0, 2, 1), ori.permute(0, 2, 1)
    d1, _ = ref_knn_points(a, o, 1); d2, _ = ref_knn_points(o, a, 1)
    cd = d1.squeeze(-1).mean(-1) + d2.squeeze(-1).mean(-1)
    hd = ref_knn_points(a, o, 1)[0].squeeze(-1).max(-1)[0]
    _, i1 = ref_knn_points(a, o, 1)
    nrm = ref_gather(normal.permute(0, 2, 1), i1).permute(0, 3, 1, 2).squeeze(3)
    _, ik = ref_knn_points(a, a, k + 1)
    nn = ref_gather(a, ik).permute(0, 3, 1, 2)[:, :, :, 1:]
    vec = nn - adv.unsqueeze(3)
    vec = vec / torch.sqrt(torch.sum(vec ** 2, dim=1, keepdim=True) + 1e-12)
    kappa = torch.abs((vec * nrm.unsqueeze(3)).sum(1)).mean(2)
    _, i2 = ref_knn_points(a, o, 1)
    ok = ref_gather(ori_kappa.unsqueeze(2), i2).view(adv.shape[0], -1)
    curv = ((kappa - ok) ** 2).mean(-1)
    return (cd + 0.1 * hd + curv).sum()


def our_loss(adv, ori, normal, ori_kappa, k=16):
    cd = LU.chamfer_loss(adv, ori)
    hd = LU.hausdorff_loss(adv, ori)
    adv_kappa, _ = LU._get_kappa_adv(adv, ori, normal, k)
    curv = LU.curvature_loss(adv, ori, adv_kappa, ori_kappa)
    return (cd + 0.1 * hd + curv).sum()


def main():
    for (B, N) in [(32, 2048), (128, 2048)]:
        ori = synth.face_clouds(B, N, seed=3).cuda().transpose(1, 2).contiguous()
        adv = (ori + 0.01 * torch.randn_like(ori)).requires_grad_(True)
        normal = torch.nn.functional.normalize(torch.randn_like(ori), dim=1)
        with torch.no_grad():
            ori_kappa = LU._get_kappa_ori(ori, normal, 16)

        def run(fn):
            adv.grad = None
            l = fn(adv, ori, normal, ori_kappa)
            l.backward()
            return l

        l1 = run(our_loss); g1 = adv.grad.clone()
        t1 = timeit(lambda: run(our_loss), reps=5)
        try:
            l2 = run(ref_loss); g2 = adv.grad.clone()
            t2 = timeit(lambda: run(ref_loss), reps=3)
            extra = f"torch formulation {t2:8.2f} ms  speed-up {t2 / t1:5.1f}x  loss rel diff {abs(float(l1 - l2)) / abs(float(l2)):.1e}  grad rel {float((g1 - g2).abs().max() / g2.abs().max()):.1e}"
        except RuntimeError as e:
            extra = "torch formulation failed: " + str(e)[:60]
        print(f"GeoA3 geometry loss fwd+bwd B={B} N={N} k=16: ours {t1:7.3f} ms   {extra}", flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
