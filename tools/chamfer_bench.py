"""BASELINE configs[1] on one GPU: Chamfer + Hausdorff forward+backward, B=32, N=M=4096, this package (eager and CUDA
graph) vs. the reference's torch formulation (attack/CW/CW_utils/distance.py:15-70: three bmm's, broadcast adds, min /
mean / max, autograd) on the same device.  Development tool; the formulation is restated here, nothing is imported
from oracle/ or the reference."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
pcd = importlib.import_module("3dpointcloudattack_b200")
synth = importlib.import_module("3dpointcloudattack_b200.synth")
from knn_bench import timeit  # noqa: E402


def ref_pairwise(x, y):                                   # distance.py:15-32
    xx = torch.bmm(x, x.transpose(2, 1)); yy = torch.bmm(y, y.transpose(2, 1)); zz = torch.bmm(x, y.transpose(2, 1))
    dx = torch.arange(0, x.shape[1], device=x.device); dy = torch.arange(0, y.shape[1], device=x.device)
    rx = xx[:, dx, dx].unsqueeze(1).expand_as(zz.transpose(2, 1)); ry = yy[:, dy, dy].unsqueeze(1).expand_as(zz)
    return rx.transpose(2, 1) + ry - 2 * zz


def ref_losses(preds, gts):                               # ChamferDistance.forward + HausdorffDistance.forward
    out = []
    for red in (torch.mean, lambda t, dim: torch.max(t, dim=dim)[0]):
        P = ref_pairwise(gts, preds)
        out += [red(torch.min(P, 1)[0], dim=1), red(torch.min(P, 2)[0], dim=1)]
    return out


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    for (B, N) in [(32, 4096), (8, 4096)]:
        ori = synth.face_clouds(B, N, seed=1234).cuda()
        adv = synth.perturb(ori.cpu(), 0.01, seed=99).cuda().requires_grad_(True)

        def ours():
            adv.grad = None
            c1, c2 = pcd.distance.chamfer(adv, ori); h1, h2 = pcd.distance.hausdorff(adv, ori)
            (c1 + c2 + h1 + h2).sum().backward()
            return c1, c2, h1, h2

        def ref():
            adv.grad = None
            l = ref_losses(adv, ori)
            (l[0] + l[1] + l[2] + l[3]).sum().backward()
            return l

        o = [t.detach().clone() for t in ours()]; go = adv.grad.clone()
        r = [t.detach().clone() for t in ref()]; gr = adv.grad.clone()
        t_ours = timeit(ours, reps=10)
        g = pcd.graph.GraphedLoss(lambda a, b: ((lambda c, h: ((c[0] + c[1] + h[0] + h[1]).sum(), ()))(pcd.distance.chamfer(a, b), pcd.distance.hausdorff(a, b))), adv, ori)
        t_graph = timeit(g.replay, reps=10)
        t_ref = timeit(ref, reps=3)
        pairs = B * N * N
        print(f"chamfer+hausdorff fwd+bwd B={B} N=M={N}: ours eager {t_ours * 1e3:7.1f} us, graph {t_graph * 1e3:7.1f} us ({pairs / t_graph / 1e9:.2f} Tpair/s)   "
              f"torch formulation on the same GPU {t_ref:8.2f} ms   speed-up {t_ref / t_graph:6.0f}x (graph) / {t_ref / t_ours:5.0f}x (eager)   "
              f"losses equal {all(torch.equal(a, b) for a, b in zip(o[2:], r[2:]))} (hausdorff), chamfer rel {max(float(((a - b).abs() / b.abs()).max()) for a, b in zip(o[:2], r[:2])):.1e}, "
              f"grad rel {float((go - gr).abs().max() / gr.abs().max()):.1e}", flush=True)
        del ori, adv
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
