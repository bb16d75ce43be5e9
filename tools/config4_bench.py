"""BASELINE configs[3] shape: one GeoA3-style iteration (attack/GeoA3/GeoA3_attack.py:103-183: victim forward, margin
loss, Chamfer + 0.1 Hausdorff + curvature k=16, backward, Adam step) against a DGCNN victim (k=20 edge-conv graph),
N=2048: this package's kernels vs. the reference's torch formulations on the same GPU.  Development tool."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
pcd = importlib.import_module("3dpointcloudattack_b200")
synth = importlib.import_module("3dpointcloudattack_b200.synth")
import victims  # noqa: E402
from geoa3_bench import our_loss, ref_loss  # noqa: E402
from knn_bench import timeit  # noqa: E402
LU = pcd.loss_utils


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    N = 2048
    for B in (16, 64):
        torch.manual_seed(0)
        ours_v = victims.DGCNNVictim(lambda x, k: pcd.dgcnn.get_graph_feature(x, k=k), k=20).cuda().eval()
        ref_v = victims.DGCNNVictim(victims.torch_graph_feature, k=20).cuda().eval()
        ref_v.load_state_dict(ours_v.state_dict())
        for m in (ours_v, ref_v):
            for p in m.parameters():
                p.requires_grad_(False)
        ori = synth.face_clouds(B, N, seed=3).cuda().transpose(1, 2).contiguous()
        normal = torch.nn.functional.normalize(torch.randn_like(ori), dim=1)
        with torch.no_grad():
            ori_kappa = LU._get_kappa_ori(ori, normal, 16)
            target = ours_v(ori)[0].argmax(1)
        res = {}
        for name, victim, geo in (("ours", ours_v, our_loss), ("torch", ref_v, ref_loss)):
            adv = (ori + 0.01 * torch.randn_like(ori)).requires_grad_(True)
            opt = torch.optim.Adam([adv], lr=1e-2)

            def it():
                logp = victim(adv)[0]
                cls = torch.nn.functional.nll_loss(logp, target, reduction="sum") * -1.0
                loss = cls + 10.0 * geo(adv, ori, normal, ori_kappa)
                opt.zero_grad()
                loss.backward()
                opt.step()

            res[name] = timeit(it, reps=5 if name == "ours" else 3)
            del adv, opt
            torch.cuda.empty_cache()
        print(f"GeoA3-style iteration vs DGCNN, B={B} N={N}: ours {res['ours']:8.2f} ms ({1e3 / res['ours']:6.2f} it/s)   torch formulations "
              f"{res['torch']:8.2f} ms ({1e3 / res['torch']:6.2f} it/s)   speed-up {res['torch'] / res['ours']:.2f}x", flush=True)
        del ours_v, ref_v
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
