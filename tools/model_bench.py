"""Model-level effect of the drop-in (development tool): DGCNN-shaped victim forward + backward w.r.t. the input
cloud with this package's k-NN + edge-feature kernels vs. the reference's torch formulation, same weights, same GPU.
(BASELINE configs[3] shape: k=20 edge-conv graph, N=2048.)"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
pcd = importlib.import_module("3dpointcloudattack_b200")
synth = importlib.import_module("3dpointcloudattack_b200.synth")
import victims  # noqa: E402
from knn_bench import timeit  # noqa: E402


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    N, k = 2048, 20
    for B in (16, 64):
        torch.manual_seed(0)
        ours = victims.DGCNNVictim(lambda x, kk: pcd.dgcnn.get_graph_feature(x, k=kk), k=k).cuda().eval()
        ref = victims.DGCNNVictim(victims.torch_graph_feature, k=k).cuda().eval()
        ref.load_state_dict(ours.state_dict())
        for m in (ours, ref):
            for p in m.parameters():
                p.requires_grad_(False)
        x = synth.face_clouds(B, N, seed=5).cuda().transpose(1, 2).contiguous().requires_grad_(True)

        def run(m):
            x.grad = None
            out = m(x)[0]
            out[:, 0].sum().backward()
            return out

        o1 = run(ours); g1 = x.grad.clone()
        o2 = run(ref); g2 = x.grad.clone()
        same_graph = float((pcd.dgcnn.knn(x.detach(), k) == (lambda t: t)(victims.torch_graph_feature.__globals__["torch"].topk(
            -(torch.sum(x.detach() ** 2, 1, keepdim=True).transpose(2, 1) - 2 * torch.matmul(x.detach().transpose(2, 1), x.detach())
              + torch.sum(x.detach() ** 2, 1, keepdim=True)), k)[1])).float().mean())
        t_ours = timeit(lambda: run(ours), reps=5)
        t_ref = timeit(lambda: run(ref), reps=3)
        print(f"DGCNN fwd+bwd B={B} N={N} k={k}: ours {t_ours:8.2f} ms   torch formulation {t_ref:8.2f} ms   speed-up {t_ref / t_ours:.2f}x   "
              f"max|dlogp| {float((o1 - o2).abs().max()):.2e}  grad rel {float((g1 - g2).abs().max() / g2.abs().max()):.2e}  "
              f"layer-1 graph index agreement {same_graph:.5f}", flush=True)
        del ours, ref, x, o1, o2, g1, g2
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
