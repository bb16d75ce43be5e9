"""Edge-feature / FPS timing (development tool): ours vs. the reference's torch formulation on the same GPU,
with achieved HBM GB/s against the algorithmic bytes (output written once / gradient read once)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcd = importlib.import_module("3dpointcloudattack_b200")
F = pcd.functional
from tools.knn_bench import timeit  # noqa: E402


def ref_graph_feature(x, idx):            # model/dgcnn.py:203-227 formulation (device = x.device)
    B, C, N = x.shape
    k = idx.shape[2]
    idx = (idx + torch.arange(0, B, device=x.device).view(-1, 1, 1) * N).view(-1)
    xt = x.transpose(2, 1).contiguous()
    feature = xt.view(B * N, -1)[idx, :].view(B, N, k, C)
    xr = xt.view(B, N, 1, C).repeat(1, 1, k, 1)
    return torch.cat((feature - xr, xr), dim=3).permute(0, 3, 1, 2).contiguous()


def ref_fps(xyz, npoint):                 # model/pointnet2_utils.py:59-81 formulation
    B, N, _ = xyz.shape
    dev = xyz.device
    centroids = torch.zeros(B, npoint, dtype=torch.long, device=dev)
    distance = torch.ones(B, N, device=dev) * 1e10
    farthest = torch.zeros(B, dtype=torch.long, device=dev)
    bi = torch.arange(B, dtype=torch.long, device=dev)
    for i in range(npoint):
        centroids[:, i] = farthest
        centroid = xyz[bi, farthest, :].view(B, 1, 3)
        dist = torch.sum((xyz - centroid) ** 2, -1)
        mask = dist < distance
        distance[mask] = dist[mask]
        farthest = torch.max(distance, -1)[1]
    return centroids


def main():
    for (B, C, N, k) in [(128, 3, 2048, 20), (128, 64, 2048, 20), (128, 128, 2048, 20), (16, 64, 2048, 20)]:
        x = torch.randn(B, C, N, device="cuda", requires_grad=True)
        idx = torch.randint(0, N, (B, N, k), device="cuda", dtype=torch.int32)
        ops = (F.EDGE_DIFF, F.EDGE_CENTER)
        out = F.edge_feature(x, idx, ops)
        g = torch.randn_like(out)
        out_bytes = out.numel() * 4
        fwd = timeit(lambda: F.edge_feature(x, idx, ops))
        bwd = timeit(lambda: torch.autograd.grad(out, x, g, retain_graph=True))
        del out
        idx64 = idx.long()
        try:
            rf = timeit(lambda: ref_graph_feature(x, idx64), reps=3)
            ro = ref_graph_feature(x, idx64)
            rb = timeit(lambda: torch.autograd.grad(ro, x, g, retain_graph=True), reps=3)
            del ro
        except RuntimeError as e:
            rf = rb = float("nan"); print("reference failed:", str(e)[:80])
        print(f"edge B={B} C={C} N={N} k={k}: out {out_bytes / 1e9:.2f} GB  fwd {fwd * 1e3:8.1f} us ({out_bytes / fwd / 1e6:6.0f} GB/s)"
              f"  bwd {bwd * 1e3:8.1f} us ({out_bytes / bwd / 1e6:6.0f} GB/s)   torch-GPU reference fwd {rf * 1e3:8.1f} bwd {rb * 1e3:8.1f} us"
              f"  speed-up {rf / fwd:.1f}x / {rb / bwd:.1f}x", flush=True)
        del g, x, idx, idx64
        torch.cuda.empty_cache()
    for (B, N, S) in [(64, 1024, 512), (64, 512, 128), (128, 2048, 512), (32, 4096, 1024), (8, 16384, 1024)]:
        xyz = torch.rand(B, N, 3, device="cuda")
        ours = timeit(lambda: F.farthest_point_sample(xyz, S))
        ref = timeit(lambda: ref_fps(xyz, S), reps=2)
        # torch's GPU sum over the 3 coordinates may round in another order than its CPU sum (which the
        # kernel follows bit for bit, tests/test_gpu_parity.py): report agreement instead of asserting
        agree = (F.farthest_point_sample(xyz, S).long() == ref_fps(xyz, S)).float().mean().item()
        print(f"fps B={B} N={N} npoint={S}: ours {ours * 1e3:8.1f} us ({ours * 1e6 / S:6.0f} ns/iter)   torch-GPU reference {ref * 1e3:9.1f} us"
              f"  speed-up {ref / ours:.1f}x  index agreement with torch-GPU {agree:.4f}", flush=True)


if __name__ == "__main__":
    main()
