"""Edge-feature backward: gather form (inverted graph) vs shared-memory atomics (development tool)."""
import importlib, os, sys
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcd = importlib.import_module("3dpointcloudattack_b200")
F = pcd.functional
from tools.knn_bench import timeit  # noqa: E402

for (B, C, N, k) in [(128, 64, 2048, 20), (128, 128, 2048, 20), (128, 64, 1024, 20), (16, 64, 2048, 20), (128, 3, 2048, 20), (32, 64, 4096, 16)]:
    x = torch.randn(B, C, N, device="cuda", requires_grad=True)
    idx = pcd.dgcnn.knn(torch.randn(B, 3, N, device="cuda"), k).int()        # a real k-NN graph (locality as in the victims)
    ops = (F.EDGE_DIFF, F.EDGE_CENTER)
    out = F.edge_feature(x, idx, ops)
    g = torch.randn_like(out)
    nbytes = out.numel() * 4
    res = {}
    for gather in (True, False):
        F.deterministic_edge_backward(gather)
        t = timeit(lambda: torch.autograd.grad(out, x, g, retain_graph=True))
        res[gather] = (t, torch.autograd.grad(out, x, g, retain_graph=True)[0])
    F.deterministic_edge_backward(False)
    err = ((res[True][1] - res[False][1]).abs().max() / res[False][1].abs().max()).item()
    print(f"B={B} C={C} N={N} k={k}: g {nbytes / 1e9:.2f} GB  gather {res[True][0] * 1e3:8.1f} us ({nbytes / res[True][0] / 1e6:6.0f} GB/s)"
          f"  atomics {res[False][0] * 1e3:8.1f} us ({nbytes / res[False][0] / 1e6:6.0f} GB/s)  max rel diff {err:.1e}", flush=True)
    if (B, C, N) == (128, 64, 2048):
        F.deterministic_edge_backward(True)
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            torch.autograd.grad(out, x, g, retain_graph=True)
            torch.cuda.synchronize()
        for e in prof.key_averages():
            if "edge" in e.key:
                print(f"    {e.key[:60]:60s} {e.device_time_total:9.1f} us")
        F.deterministic_edge_backward(False)
    del out, g, x, idx
    torch.cuda.empty_cache()
