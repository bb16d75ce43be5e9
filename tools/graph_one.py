"""one edge-feature forward+backward and one FPS call (profiling target): python tools/graph_one.py B C N k"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcd = importlib.import_module("3dpointcloudattack_b200")
F = pcd.functional
B, C, N, k = [int(a) for a in sys.argv[1:5]]
x = torch.randn(B, C, N, device="cuda", requires_grad=True)
idx = torch.randint(0, N, (B, N, k), device="cuda", dtype=torch.int32)
for _ in range(2):
    out = F.edge_feature(x, idx, (F.EDGE_DIFF, F.EDGE_CENTER))
    (g,) = torch.autograd.grad(out, x, torch.ones_like(out))
F.deterministic_edge_backward(True)                      # the gather form once (edge_csr_build + edge_feature_bwd_gather)
out = F.edge_feature(x, idx, (F.EDGE_DIFF, F.EDGE_CENTER))
(g2,) = torch.autograd.grad(out, x, torch.ones_like(out))
F.deterministic_edge_backward(False)
cf = torch.randn(B, 3, N, device="cuda")
adv = cf + 0.03 * torch.randn_like(cf)
F.clip_points_(adv.clone(), cf, 0.03)
F.clip_points_(adv.clone(), cf, 0.03, F.CLIP_PROJECT_LINF, normal=cf)
F.clip_points_(adv.clone(), cf, 0.5, F.CLIP_L2)
F.lp_clip(adv - cf, 0.02)
F.offset_proj(adv - cf, cf, idx[:, :, 0])
xyz = torch.rand(64, 1024, 3, device="cuda")
for _ in range(2):
    s = F.farthest_point_sample(xyz, 512)
torch.cuda.synchronize()
print("ok", float(g.sum()), int(s.sum()))
