"""BASELINE configs[2]: kNN attack (attack/KNN/KNN_attack.py loop) vs. a PointNet++ SSG victim, B=64, N=1024,
ChamferkNNDist(knn_k=16): iterations/s of the device-resident loop with this package's kernels (Chamfer sweep, k-NN
outlier loss, FPS, ball query) vs. the same loop on the reference's torch formulations, same weights, same GPU.
Development tool."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
pcd = importlib.import_module("3dpointcloudattack_b200")
synth = importlib.import_module("3dpointcloudattack_b200.synth")
import victims  # noqa: E402
CL = pcd.cw_loop


class TorchChamferkNN(torch.nn.Module):
    """attack/CW/CW_utils/dist_utils.py:189-223 on the reference's formulations (bmm matrices, topk)."""

    def __init__(self, k=16, alpha=1.05, w1=5., w2=3.):
        super().__init__()
        self.k, self.alpha, self.w1, self.w2 = k, alpha, w1, w2

    def forward(self, adv, ori, weights=None, batch_avg=False):
        x, y = ori, adv                                                     # P[b,i,j] = |gts_i|^2 + |preds_j|^2 - 2 g.p
        zz = torch.bmm(x, y.transpose(2, 1))
        rx = torch.sum(x * x, -1)[:, :, None]; ry = torch.sum(y * y, -1)[:, None, :]
        P = rx + ry - 2 * zz
        chamfer = torch.min(P, 1)[0].mean(1)                                # adv2ori
        pc = adv.transpose(2, 1)
        inner = -2. * torch.matmul(pc.transpose(2, 1), pc)
        xx = torch.sum(pc ** 2, dim=1, keepdim=True)
        dist = xx + inner + xx.transpose(2, 1)
        neg_value, _ = (-dist).topk(k=self.k + 1, dim=-1)
        value = torch.mean(-(neg_value[..., 1:]), dim=-1)
        with torch.no_grad():
            thr = value.mean(-1) + self.alpha * value.std(-1)
            mask = (value > thr[:, None]).float()
        knn = torch.mean(value * mask, dim=1)
        return chamfer * self.w1 + knn * self.w2


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    B, N, iters = 64, 1024, 30
    torch.manual_seed(0)
    ours_v = victims.PointNet2SSGVictim(pcd.pointnet2_utils.sample_and_group).cuda().eval()
    ref_v = victims.PointNet2SSGVictim(victims.torch_sample_and_group).cuda().eval()
    ref_v.load_state_dict(ours_v.state_dict())
    data = synth.face_clouds(B, N, seed=77).cuda()
    with torch.no_grad():
        torch.manual_seed(5); target = ours_v(data.transpose(1, 2))[0].argmax(1)
    res = {}
    for name, victim, dist in (("ours", ours_v, pcd.dist_utils.ChamferkNNDist(knn_k=16)), ("torch", ref_v, TorchChamferkNN(16))):
        atk = CL.KNNAttack(victim, CL.UntargetedLogitsAdvLoss(kappa=15.), dist, CL.ProjectInnerClipLinf(0.1), attack_lr=1e-3, num_iter=iters)
        torch.manual_seed(9); atk.attack(data, target, seed=1)            # warm-up
        torch.manual_seed(9); adv, _ = atk.attack(data, target, seed=1)
        res[name] = (iters / (atk.loop_ms * 1e-3), adv)
    d = (res["ours"][1] - res["torch"][1]).abs()
    print(f"kNN attack vs PointNet++ SSG, B={B} N={N}, ChamferkNNDist(k=16): ours {res['ours'][0]:7.1f} it/s   torch formulations "
          f"{res['torch'][0]:7.1f} it/s   speed-up {res['ours'][0] / res['torch'][0]:.1f}x   adversarial clouds agree on "
          f"{float((d < 1e-4).float().mean()):.3f} of the coordinates (FPS start draws seeded identically)")


if __name__ == "__main__":
    main()
