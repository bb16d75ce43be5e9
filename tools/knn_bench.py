"""k-NN / ball-query timing (development tool): ours vs. the reference's torch formulation on the same GPU."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcd = importlib.import_module("3dpointcloudattack_b200")
F = pcd.functional


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2]


def ref_knn(x, k):          # model/dgcnn.py:194-200 formulation
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    return (-xx - inner - xx.transpose(2, 1)).topk(k=k, dim=-1)[1]


def breakdown():
    """per-kernel device time of our k-NN path (torch profiler)"""
    from torch.profiler import profile, ProfilerActivity
    for (B, N, C, K) in [(128, 2048, 3, 20), (64, 1024, 3, 17), (128, 2048, 64, 20)]:
        x = torch.randn(B, C, N, device="cuda")
        pts = x.transpose(1, 2)
        for _ in range(3):
            F.knn(pts, pts, K)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                F.knn(pts, pts, K)
            torch.cuda.synchronize()
        print(f"--- B={B} N={N} C={C} K={K}")
        for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:6]:
            print(f"   {e.device_time_total / e.count:9.1f} us x{e.count:3d}  {e.key[:90]}")


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    if "--breakdown" in sys.argv:
        return breakdown()
    for (B, N, C, K) in [(64, 1024, 3, 17), (128, 2048, 3, 20), (128, 2048, 64, 20), (128, 2048, 128, 20), (32, 4096, 3, 17)]:
        x = torch.randn(B, C, N, device="cuda")
        pts = x.transpose(1, 2)
        ours = timeit(lambda: F.knn(pts, pts, K))
        try:
            ref = timeit(lambda: ref_knn(x, K), reps=5)
        except RuntimeError as e:
            ref = float("nan")
        pairs = B * N * N
        print(f"knn B={B} N={N} C={C} K={K}: ours {ours*1e3:9.1f} us ({pairs/ours/1e9:6.2f} Tpair/s, {(2*C+2)*pairs/ours/1e9:6.1f} TFLOP/s)"
              f"   torch-GPU reference {ref*1e3:9.1f} us   speed-up {ref/ours:5.1f}x")
    xyz = torch.rand(64, 1024, 3, device="cuda"); new = xyz[:, :512].contiguous()
    ours = timeit(lambda: F.ball_query(0.2, 32, xyz, new))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    def ref_ball():
        Bq, Nq, _ = xyz.shape; S = new.shape[1]
        gi = torch.arange(Nq, device="cuda").view(1, 1, Nq).repeat([Bq, S, 1])
        d = -2 * torch.matmul(new, xyz.permute(0, 2, 1)); d += torch.sum(new ** 2, -1).view(Bq, S, 1); d += torch.sum(xyz ** 2, -1).view(Bq, 1, Nq)
        gi[d > 0.2 ** 2] = Nq
        gi = gi.sort(dim=-1)[0][:, :, :32]
        return gi
    ref = timeit(ref_ball, reps=5)
    print(f"ball query B=64 S=512 N=1024 ns=32: ours {ours*1e3:8.1f} us  torch-GPU reference {ref*1e3:8.1f} us  speed-up {ref/ours:5.1f}x")


if __name__ == "__main__":
    main()
